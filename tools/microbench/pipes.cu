// Pipe-rate micro-benchmark for the stencil kernels' instruction mix (B200, sm_100a).
// Each kernel runs N_IT iterations of 8 independent dependency chains per thread; reports warp-instructions
// per clock per SM at 4, 8, 16, 32 warps per SM. Build: nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -O3
#include <cstdio>
#include <algorithm>
#include <cuda_runtime.h>
#define N_IT 512
__device__ __forceinline__ void add2(float& x, float& y, float a, float b) {
    asm volatile("{.reg .b64 ra, rb; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%3}; add.rn.f32x2 ra, ra, rb; mov.b64 {%0,%1}, ra;}" : "+f"(x), "+f"(y) : "f"(a), "f"(b));
}
__device__ __forceinline__ void mul2(float& x, float& y, float a, float b) {
    asm volatile("{.reg .b64 ra, rb; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%3}; mul.rn.f32x2 ra, ra, rb; mov.b64 {%0,%1}, ra;}" : "+f"(x), "+f"(y) : "f"(a), "f"(b));
}
template <int MODE>
__global__ void k(float* out, float p, float q, long long* cyc) {
    __shared__ float4 sm[8][128];
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = threadIdx.x * 0.001f + i;
    const int lane = threadIdx.x;
    sm[0][lane & 127] = make_float4(1, 2, 3, 4);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N_IT; it++) {
        if (MODE == 0) {  // 16 FADD
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = __fadd_rn(v[i], p);
        } else if (MODE == 1) {  // 16 FMUL
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = __fmul_rn(v[i], p);
        } else if (MODE == 2) {  // 8 FADD2
#pragma unroll
            for (int i = 0; i < 16; i += 2) add2(v[i], v[i + 1], p, q);
        } else if (MODE == 3) {  // 8 FMUL2
#pragma unroll
            for (int i = 0; i < 16; i += 2) mul2(v[i], v[i + 1], p, q);
        } else if (MODE == 4) {  // 8 FMUL + 8 FADD interleaved (separate chains)
#pragma unroll
            for (int i = 0; i < 16; i += 2) { v[i] = __fmul_rn(v[i], p); v[i + 1] = __fadd_rn(v[i + 1], q); }
        } else if (MODE == 5) {  // 16 FFMA
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = __fmaf_rn(v[i], p, q);
        } else if (MODE == 6) {  // 16 SHFL
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = __shfl_up_sync(0xffffffffu, v[i], 1);
        } else if (MODE == 7) {  // 4 LDS.128 + 4 STS.128
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float4 t = sm[i][threadIdx.x & 127];
                v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
                sm[4 + i][threadIdx.x & 127] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
        } else if (MODE == 8) {  // stencil pattern scalar: (p*a + q*b) + p*c on 4 columns: 12 FMUL + 8 FADD
#pragma unroll
            for (int i = 0; i < 4; i++) v[i] = __fadd_rn(__fadd_rn(__fmul_rn(p, v[4 + i]), __fmul_rn(q, v[8 + i])), __fmul_rn(p, v[12 + i]));
#pragma unroll
            for (int i = 0; i < 4; i++) v[12 + i] = v[8 + i], v[8 + i] = v[4 + i], v[4 + i] = v[i];
        } else if (MODE == 9) {  // same, packed multiplies + scalar adds: 6 FMUL2 + 8 FADD
#pragma unroll
            for (int i = 0; i < 4; i += 2) {
                float a0 = v[4 + i], a1 = v[5 + i], b0 = v[8 + i], b1 = v[9 + i], c0 = v[12 + i], c1 = v[13 + i];
                mul2(a0, a1, p, p); mul2(b0, b1, q, q); mul2(c0, c1, p, p);
                v[i] = __fadd_rn(__fadd_rn(a0, b0), c0);
                v[i + 1] = __fadd_rn(__fadd_rn(a1, b1), c1);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) v[12 + i] = v[8 + i], v[8 + i] = v[4 + i], v[4 + i] = v[i];
        } else if (MODE == 10) {  // 8 FADD2 + 8 FADD mixed
#pragma unroll
            for (int i = 0; i < 8; i += 2) add2(v[i], v[i + 1], p, q);
#pragma unroll
            for (int i = 8; i < 16; i++) v[i] = __fadd_rn(v[i], p);
        } else if (MODE == 11) {  // 8 FMUL2 + 8 SHFL mixed
#pragma unroll
            for (int i = 0; i < 8; i += 2) mul2(v[i], v[i + 1], p, q);
#pragma unroll
            for (int i = 8; i < 16; i++) v[i] = __shfl_up_sync(0xffffffffu, v[i], 1);
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) { cyc[(blockIdx.x * 32 + (threadIdx.x >> 5)) * 2] = t0; cyc[(blockIdx.x * 32 + (threadIdx.x >> 5)) * 2 + 1] = t1; }
}
template <int MODE>
void run(const char* name, int inst_per_it) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4 * sizeof(float));
    cudaMalloc(&cyc, 148 * 64 * sizeof(long long));
    printf("%-44s", name);
    for (int wps : {4, 8, 16, 32}) {
        // one block per SM with wps warps
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<MODE><<<148, wps * 32>>>(out, 1.0001f, 0.9999f, cyc);
        cudaEventRecord(e0);
        k<MODE><<<148, wps * 32>>>(out, 1.0001f, 0.9999f, cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        static long long h[148 * 64]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double mc = 0;
        for (int b = 0; b < 148; b++) { long long lo = h[b * 64], hi = h[b * 64 + 1];
            for (int w = 0; w < wps; w++) { lo = std::min(lo, h[(b * 32 + w) * 2]); hi = std::max(hi, h[(b * 32 + w) * 2 + 1]); }
            mc += (double)(hi - lo); }
        mc /= 148;
        printf("  w%-2d %6.2f inst/clk/SM", wps, (double)inst_per_it * N_IT * wps / mc);
    }
    printf("\n");
    cudaError_t e = cudaGetLastError(); if (e) printf("err %s\n", cudaGetErrorString(e));
}
int main() {
    run<0>("16 FADD", 16); run<1>("16 FMUL", 16); run<2>("8 FADD2", 8); run<3>("8 FMUL2", 8);
    run<4>("8 FMUL + 8 FADD", 16); run<5>("16 FFMA", 16); run<6>("16 SHFL", 16); run<7>("4 LDS.128 + 4 STS.128 (+16 FADD)", 8);
    run<8>("stencil scalar 12 FMUL + 8 FADD (+12 MOV?)", 20); run<9>("stencil packed 6 FMUL2 + 8 FADD", 14);
    run<10>("4 FADD2 + 8 FADD", 12); run<11>("4 FMUL2 + 8 SHFL", 12);
    return 0;
}
