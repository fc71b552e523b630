// tile_util.cuh -- helpers shared by the float4 ("fast") tile kernels.
//
// Fast tile kernels keep every shared-memory stage buffer in rows whose x extent is a multiple of 4 and
// whose origin is 16-byte aligned, so that each thread produces 4 horizontally adjacent outputs from
// aligned float4 loads. They need W % 4 == 0 and 16-byte aligned image slabs; other shapes use the
// generic scalar kernels. Border semantics (fill_border, akaze/src/types/image.rs:239-260) are applied
// after each pass by fix_border below, only in blocks whose regions touch a clamp band.
#pragma once

#include <cuda_runtime.h>

namespace akz {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 sub4(const float4& a, const float4& b) {
    return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
}

// Packed f32x2 multiply (FMUL2, new on sm_100): two IEEE products in one issue slot, each rounded exactly like a scalar
// FMUL. Only the PRODUCTS of the 3-tap filters are packed; their sums stay scalar FADDs, because ptxas 12.9 contracts
// mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false (profiles/r1z_pipes_microbench.txt), which would break
// the bit-exact match with the reference's unfused arithmetic.
__device__ __forceinline__ void mul2(float a0, float a1, float k, float& p0, float& p1) {
    asm("{.reg .b64 ra, rk, rp;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rk, {%4, %4};\n\t"
        "mul.rn.f32x2 rp, ra, rk;\n\t"
        "mov.b64 {%0, %1}, rp;}"
        : "=f"(p0), "=f"(p1)
        : "f"(a0), "f"(a1), "f"(k));
}
__device__ __forceinline__ void mul2v(float a0, float a1, float b0, float b1, float& p0, float& p1) {
    asm("{.reg .b64 ra, rb, rp;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rp, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rp;}"
        : "=f"(p0), "=f"(p1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// Packed f32x2 add / subtract (FADD2): only for operands that are NOT products computed in the same straight-line code
// (loaded, shuffled, loop-carried or themselves sums) -- ptxas would contract a packed product + packed add into FFMA2.
__device__ __forceinline__ void add2(float a0, float a1, float b0, float b1, float& r0, float& r1) {
    asm("{.reg .b64 ra, rb, rr;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rr, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rr;}"
        : "=f"(r0), "=f"(r1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void sub2(float a0, float a1, float b0, float b1, float& r0, float& r1) {
    asm("{.reg .b64 ra, rb, rr;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "sub.rn.f32x2 rr, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rr;}"
        : "=f"(r0), "=f"(r1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// AKZ_FAST_MATH (opt-in build, libakaze_b200_fast.so): the filters use packed fused multiply-adds -- one rounding per tap
// instead of two, 3 instead of 7 instructions per pixel pair and 3-tap filter. Results then differ from the reference in
// the last bits (tools/fast_math_report.py measures by how much); the default build never defines it.
__device__ __forceinline__ void fma2(float a0, float a1, float k, float c0, float c1, float& r0, float& r1) {
    asm("{.reg .b64 ra, rk, rc, rr;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rk, {%4, %4};\n\t"
        "mov.b64 rc, {%5, %6};\n\t"
        "fma.rn.f32x2 rr, ra, rk, rc;\n\t"
        "mov.b64 {%0, %1}, rr;}"
        : "=f"(r0), "=f"(r1)
        : "f"(a0), "f"(a1), "f"(k), "f"(c0), "f"(c1));
}
// out[j] = (k0 * a[j] + k1 * b[j]) + k2 * c[j], j = 0..3, in tap order (products packed in pairs, sums scalar)
__device__ __forceinline__ void tap3x4(float k0, float k1, float k2, const float (&a)[4], const float (&b)[4], const float (&c)[4],
                                       float (&out)[4]) {
#ifdef AKZ_FAST_MATH
#pragma unroll
    for (int j = 0; j < 4; j += 2) {
        float p0, p1;
        mul2(a[j], a[j + 1], k0, p0, p1);
        fma2(b[j], b[j + 1], k1, p0, p1, p0, p1);
        fma2(c[j], c[j + 1], k2, p0, p1, out[j], out[j + 1]);
    }
    return;
#endif
#pragma unroll
    for (int j = 0; j < 4; j += 2) {
        float pa0, pa1, pb0, pb1, pc0, pc1;
        mul2(a[j], a[j + 1], k0, pa0, pa1);
        mul2(b[j], b[j + 1], k1, pb0, pb1);
        mul2(c[j], c[j + 1], k2, pc0, pc1);
        out[j] = (pa0 + pb0) + pc0;
        out[j + 1] = (pa1 + pb1) + pc1;
    }
}

// out[j] = (((k0 * a[j] + k1 * b[j]) + k2 * c[j]) + k3 * d[j]) + k4 * e[j], the 5-tap form of tap3x4
__device__ __forceinline__ void tap5x4(float k0, float k1, float k2, float k3, float k4, const float (&a)[4], const float (&b)[4],
                                       const float (&c)[4], const float (&d)[4], const float (&e)[4], float (&out)[4]) {
#ifdef AKZ_FAST_MATH
#pragma unroll
    for (int j = 0; j < 4; j += 2) {
        float p0, p1;
        mul2(a[j], a[j + 1], k0, p0, p1);
        fma2(b[j], b[j + 1], k1, p0, p1, p0, p1);
        fma2(c[j], c[j + 1], k2, p0, p1, p0, p1);
        fma2(d[j], d[j + 1], k3, p0, p1, p0, p1);
        fma2(e[j], e[j + 1], k4, p0, p1, out[j], out[j + 1]);
    }
    return;
#endif
#pragma unroll
    for (int j = 0; j < 4; j += 2) {
        float pa0, pa1, pb0, pb1, pc0, pc1, pd0, pd1, pe0, pe1;
        mul2(a[j], a[j + 1], k0, pa0, pa1);
        mul2(b[j], b[j + 1], k1, pb0, pb1);
        mul2(c[j], c[j + 1], k2, pc0, pc1);
        mul2(d[j], d[j + 1], k3, pd0, pd1);
        mul2(e[j], e[j + 1], k4, pe0, pe1);
        out[j] = (((pa0 + pb0) + pc0) + pd0) + pe0;
        out[j + 1] = (((pa1 + pb1) + pc1) + pd1) + pe1;
    }
}

// v[0..11] = row[cx-4 .. cx+7] (cx a multiple of 4)
__device__ __forceinline__ void load12(const float* row, int cx, float (&v)[12]) {
    const float4 a = ld4(row + cx - 4), b = ld4(row + cx), c = ld4(row + cx + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w;
}

// 3-tap vertical pass on float4 rows: (k0*a + k1*b) + k2*c per component, in tap order
__device__ __forceinline__ float4 vtap3(const float4& a, const float4& b, const float4& c, float k0, float k1, float k2) {
    return make_float4((k0 * a.x + k1 * b.x) + k2 * c.x, (k0 * a.y + k1 * b.y) + k2 * c.y, (k0 * a.z + k1 * b.z) + k2 * c.z,
                       (k0 * a.w + k1 * b.w) + k2 * c.w);
}

// fill_border of one filter pass applied to a shared-memory stage buffer (image.rs:239-260 in closed
// form): positions of the region whose coordinate is inside the hw-wide clamp band of the IMAGE take
// the value at the clamped coordinate -- rows first, then columns, like the reference. The region is
// [rx0, rx0+rw) x [ry0, ry0+rh) in image coordinates, stored with pitch P from buf. Called by every
// thread of a block whose region touches a clamp band (block-uniform); contains two barriers.
__device__ __forceinline__ void fix_border(float* buf, int P, int rx0, int ry0, int rw, int rh, int W, int H, int hw, int tid,
                                           int nthreads) {
    const int xa = max(rx0, 0), xb = min(rx0 + rw, W);
    const int ya = max(ry0, 0), yb = min(ry0 + rh, H);
    {   // rows above hw copy row hw, rows below H-1-hw copy row H-1-hw
        const int ntop = max(0, min(hw, yb) - ya);
        const int bot0 = max(H - hw, ya);
        const int nbot = max(0, yb - bot0);
        const int wv = xb - xa;
        for (int i = tid; i < (ntop + nbot) * wv; i += nthreads) {
            const int r = i / wv, x = xa + (i - r * wv);
            const int y = r < ntop ? ya + r : bot0 + (r - ntop);
            const int sy = r < ntop ? hw : H - 1 - hw;
            buf[(y - ry0) * P + (x - rx0)] = buf[(sy - ry0) * P + (x - rx0)];
        }
    }
    __syncthreads();
    {   // columns left of hw copy column hw, right of W-1-hw copy column W-1-hw
        const int nl = max(0, min(hw, xb) - xa);
        const int r0 = max(W - hw, xa);
        const int nr = max(0, xb - r0);
        const int hv = yb - ya;
        for (int i = tid; i < (nl + nr) * hv; i += nthreads) {
            const int c = i / hv, y = ya + (i - c * hv);
            const int x = c < nl ? xa + c : r0 + (c - nl);
            const int sx = c < nl ? hw : W - 1 - hw;
            buf[(y - ry0) * P + (x - rx0)] = buf[(y - ry0) * P + (sx - rx0)];
        }
    }
    __syncthreads();
}

}  // namespace akz
