// detector.cu -- scale-normalised Hessian determinant and ordered candidate emission.
//
// Replaces (reference paths relative to the akaze-rust repository):
//   compute_multiscale_derivatives / detector_response   akaze/src/ops/detector_response.rs:8-55
//   the threshold + 4-neighbour test and the is_out test of find_scale_space_extrema
//                                                        akaze/src/ops/scale_space_extrema.rs:32-42, 80-88
//
// Per level, one kernel evaluates the whole derivative chain in shared memory:
//   Lx  = V_off (H_main(Lsmooth))   "x_order"  (derivatives.rs:41-47; the reference's Lx is d/dy, Q1)
//   Ly  = V_main(H_off (Lsmooth))   "y_order"  (derivatives.rs:59-65)
//   Lxx = V_off (H_main(Lx)),  Lyy = V_main(H_off(Ly)),  Lxy = V_main(H_off(Lx))
//   Ldet = ((Lxx*Lyy) - (Lxy*Lxy)) * (s^4 as f32)
// with the scaled Scharr taps [n,0..,wn,0..,n] / [-1,0..0,1] applied in tap order; zero taps add an
// exact +0 and are skipped. Every pass clamps like fill_border (see scale_space.cu). Candidates are
// written as one bit per pixel; a second set of kernels turns the bitmask into a list in raster order,
// which the order-dependent cache pass of the reference needs.
#include <algorithm>
#include <cstdlib>

#include <type_traits>

#include "common.cuh"
#include "tile_util.cuh"

namespace akz {
namespace {

constexpr int DT = 32;  // tile edge
constexpr int NTX = 32, NTY = 8;

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct DetParams {
    int W, H, s;
    float n, wn;  // scharr_main_axis_kernel(s)
    float quat;   // (s*s*s*s) as f32
    float thr;
    int xmin, xmax, ymin, ymax;
    int wpr;
    int fast;  // interior 64x32 tiles are computed by k_detector_fast
};

__global__ void __launch_bounds__(NTX* NTY)
k_detector(const float* __restrict__ lsmooth, size_t img_px, float* __restrict__ oLx, float* __restrict__ oLy,
           float* __restrict__ oLdet, float* __restrict__ oLxx, float* __restrict__ oLyy, float* __restrict__ oLxy,
           unsigned int* __restrict__ mask, size_t mask_img_words, DetParams p) {
    extern __shared__ float smem[];
    const int s = p.s, W = p.W, H = p.H;
    const int HL = 2 * s + 1;
    const int PW = DT + 2 * HL;
    const int N = PW * PW;
    float* b0 = smem;
    float* b1 = b0 + N;
    float* b2 = b1 + N;
    float* b3 = b2 + N;
    float* b4 = b3 + N;
    const int tx0 = min((int)blockIdx.x * DT, W - DT);
    const int ty0 = min((int)blockIdx.y * DT, H - DT);
    const int img = blockIdx.z;
    auto idx = [&](int x, int y) { return (y - ty0 + HL) * PW + (x - tx0 + HL); };
    const float* L = lsmooth + (size_t)img * img_px;

    {  // stage 0: Lsmooth over tile +- (2s+1)
        const int xa = max(tx0 - HL, 0), xb = min(tx0 + DT + HL, W);
        const int ya = max(ty0 - HL, 0), yb = min(ty0 + DT + HL, H);
        for (int y = ya + threadIdx.y; y < yb; y += NTY)
            for (int x = xa + threadIdx.x; x < xb; x += NTX) b0[idx(x, y)] = L[(size_t)y * W + x];
    }
    __syncthreads();
    {  // stage 1: A = H_main(L) -> b1, Bo = H_off(L) -> b2 : x in tile+-(s+1), y in tile+-(2s+1)
        const int xa = max(tx0 - s - 1, 0), xb = min(tx0 + DT + s + 1, W);
        const int ya = max(ty0 - HL, 0), yb = min(ty0 + DT + HL, H);
        for (int y = ya + threadIdx.y; y < yb; y += NTY)
            for (int x = xa + threadIdx.x; x < xb; x += NTX) {
                const int cx = clampi(x, s, W - 1 - s), cy = clampi(y, s, H - 1 - s);
                const float l = b0[idx(cx - s, cy)], c = b0[idx(cx, cy)], r = b0[idx(cx + s, cy)];
                float acc = 0.0f + p.n * l;
                acc = acc + p.wn * c;
                acc = acc + p.n * r;
                b1[idx(x, y)] = acc;
                b2[idx(x, y)] = r - l;
            }
    }
    __syncthreads();
    {  // stage 2: Lx = V_off(A) -> b3, Ly = V_main(Bo) -> b4 : tile +- (s+1)
        const int xa = max(tx0 - s - 1, 0), xb = min(tx0 + DT + s + 1, W);
        const int ya = max(ty0 - s - 1, 0), yb = min(ty0 + DT + s + 1, H);
        float* ox = oLx + (size_t)img * img_px;
        float* oy = oLy + (size_t)img * img_px;
        for (int y = ya + threadIdx.y; y < yb; y += NTY)
            for (int x = xa + threadIdx.x; x < xb; x += NTX) {
                const int cx = clampi(x, s, W - 1 - s), cy = clampi(y, s, H - 1 - s);
                const float lx = b1[idx(cx, cy + s)] - b1[idx(cx, cy - s)];
                float acc = 0.0f + p.n * b2[idx(cx, cy - s)];
                acc = acc + p.wn * b2[idx(cx, cy)];
                acc = acc + p.n * b2[idx(cx, cy + s)];
                b3[idx(x, y)] = lx;
                b4[idx(x, y)] = acc;
                if (x >= tx0 && x < tx0 + DT && y >= ty0 && y < ty0 + DT) {
                    ox[(size_t)y * W + x] = lx;
                    oy[(size_t)y * W + x] = acc;
                }
            }
    }
    __syncthreads();
    {  // stage 3: C = H_main(Lx) -> b0, D = H_off(Ly) -> b1, E = H_off(Lx) -> b2 : x tile+-1, y tile+-(s+1)
        const int xa = max(tx0 - 1, 0), xb = min(tx0 + DT + 1, W);
        const int ya = max(ty0 - s - 1, 0), yb = min(ty0 + DT + s + 1, H);
        for (int y = ya + threadIdx.y; y < yb; y += NTY)
            for (int x = xa + threadIdx.x; x < xb; x += NTX) {
                const int cx = clampi(x, s, W - 1 - s), cy = clampi(y, s, H - 1 - s);
                const float l = b3[idx(cx - s, cy)], c = b3[idx(cx, cy)], r = b3[idx(cx + s, cy)];
                float acc = 0.0f + p.n * l;
                acc = acc + p.wn * c;
                acc = acc + p.n * r;
                b0[idx(x, y)] = acc;
                b2[idx(x, y)] = r - l;
                b1[idx(x, y)] = b4[idx(cx + s, cy)] - b4[idx(cx - s, cy)];
            }
    }
    __syncthreads();
    {  // stage 4: Ldet over tile +- 1 -> b3 (Lx in b3 is dead after stage 3)
        const int xa = max(tx0 - 1, 0), xb = min(tx0 + DT + 1, W);
        const int ya = max(ty0 - 1, 0), yb = min(ty0 + DT + 1, H);
        float* od = oLdet + (size_t)img * img_px;
        for (int y = ya + threadIdx.y; y < yb; y += NTY)
            for (int x = xa + threadIdx.x; x < xb; x += NTX) {
                const int cx = clampi(x, s, W - 1 - s), cy = clampi(y, s, H - 1 - s);
                const float lxx = b0[idx(cx, cy + s)] - b0[idx(cx, cy - s)];
                float lyy = 0.0f + p.n * b1[idx(cx, cy - s)];
                lyy = lyy + p.wn * b1[idx(cx, cy)];
                lyy = lyy + p.n * b1[idx(cx, cy + s)];
                float lxy = 0.0f + p.n * b2[idx(cx, cy - s)];
                lxy = lxy + p.wn * b2[idx(cx, cy)];
                lxy = lxy + p.n * b2[idx(cx, cy + s)];
                const float det = ((lxx * lyy) - (lxy * lxy)) * p.quat;  // detector_response.rs:52
                b3[idx(x, y)] = det;
                if (x >= tx0 && x < tx0 + DT && y >= ty0 && y < ty0 + DT) {
                    const size_t o = (size_t)y * W + x;
                    od[o] = det;
                    if (oLxx != nullptr) {
                        oLxx[(size_t)img * img_px + o] = lxx;
                        oLyy[(size_t)img * img_px + o] = lyy;
                        oLxy[(size_t)img * img_px + o] = lxy;
                    }
                }
            }
    }
    __syncthreads();
    {  // stage 5: threshold + strict 4-neighbour maximum + is_out (scale_space_extrema.rs:36-41, 80-87)
        unsigned int* m = mask + (size_t)img * mask_img_words;
        for (int y = ty0 + threadIdx.y; y < ty0 + DT; y += NTY) {
            const int x = tx0 + threadIdx.x;  // DT == NTX: one warp covers one tile row
            bool cand = false;
            if (x >= p.xmin && x <= p.xmax && y >= p.ymin && y <= p.ymax) {
                const float v = b3[idx(x, y)];
                cand = v > p.thr && v > b3[idx(x + 1, y)] && v > b3[idx(x - 1, y)] && v > b3[idx(x, y - 1)] &&
                       v > b3[idx(x, y + 1)];
            }
            const unsigned int bal = __ballot_sync(0xffffffffu, cand);
            if (bal != 0 && threadIdx.x == 0) {
                const int w0 = tx0 >> 5, sh = tx0 & 31;
                atomicOr(&m[(size_t)y * p.wpr + w0], bal << sh);
                if (sh != 0 && (bal >> (32 - sh)) != 0) atomicOr(&m[(size_t)y * p.wpr + w0 + 1], bal >> (32 - sh));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Fast path: interior 64x32 tiles, compile-time Scharr scale S, every shared-memory access a float4
// (4 pixels per thread), no clamp arithmetic. Same operation order as k_detector -> same bits.
// ------------------------------------------------------------------------------------------------
template <int S>
struct DetGeo {
    // x halos are multiples of 4 so that every access is an aligned float4: Ldet is produced on
    // [-4, 68) (needs [-1, 65)), so Lx/Ly are needed on [-4-S, 68+S) and Lsmooth S further out
    static constexpr int HX1 = ((4 + S) + 3) & ~3;      // x halo of A/Bo/Lx/Ly
    static constexpr int HLX = ((HX1 + S) + 3) & ~3;    // x halo of the Lsmooth tile
    static constexpr int HLY = 2 * S + 1;
    static constexpr int HY2 = S + 1;                   // y halo of Lx/Ly/C/D/E
    static constexpr int P0 = 64 + 2 * HLX, R0 = 32 + 2 * HLY;
    static constexpr int P1 = 64 + 2 * HX1, R2 = 32 + 2 * HY2;
    static constexpr int N0 = P0 * R0, N1 = P1 * R0, N2 = P1 * R2;
    static constexpr int FLOATS = N0 + 2 * N1 + 2 * N2;
};

// 12 consecutive values v[0..11] = in[x-4 .. x+7]; main-axis taps of outputs x..x+3 (tap order -S, 0, +S)
template <int S>
__device__ __forceinline__ float4 hmain(const float (&v)[12], float n, float wn) {
    float4 r;
    r.x = (n * v[4 - S] + wn * v[4]) + n * v[4 + S];
    r.y = (n * v[5 - S] + wn * v[5]) + n * v[5 + S];
    r.z = (n * v[6 - S] + wn * v[6]) + n * v[6 + S];
    r.w = (n * v[7 - S] + wn * v[7]) + n * v[7 + S];
    return r;
}
template <int S>
__device__ __forceinline__ float4 hoff(const float (&v)[12]) {
    return make_float4(v[4 + S] - v[4 - S], v[5 + S] - v[5 - S], v[6 + S] - v[6 - S], v[7 + S] - v[7 - S]);
}
__device__ __forceinline__ float4 vmain(const float4& a, const float4& b, const float4& c, float n, float wn) {
    return make_float4((n * a.x + wn * b.x) + n * c.x, (n * a.y + wn * b.y) + n * c.y, (n * a.z + wn * b.z) + n * c.z,
                       (n * a.w + wn * b.w) + n * c.w);
}
template <int S>
__global__ void __launch_bounds__(256)
k_detector_fast(const float* __restrict__ lsmooth, size_t img_px, float* __restrict__ oLx, float* __restrict__ oLy,
                float* __restrict__ oLdet, unsigned int* __restrict__ mask, size_t mask_img_words, DetParams p) {
    using G = DetGeo<S>;
    extern __shared__ __align__(16) float smem[];
    const int W = p.W, H = p.H;
    float* b0 = smem;
    float* b1 = b0 + G::N0;
    float* b2 = b1 + G::N1;
    float* b3 = b2 + G::N1;
    float* b4 = b3 + G::N2;
    // full-size tiles, the last one of a row/column shifted inwards (x0 stays a multiple of 4: W % 4 == 0)
    const int x0 = min((int)blockIdx.x * 64, W - 64), y0 = min((int)blockIdx.y * 32, H - 32);
    const int img = blockIdx.z;
    const int tid = threadIdx.x;
    const float n = p.n, wn = p.wn;
    const size_t ibase = (size_t)img * img_px;
    // does any stage region leave the image or touch a clamp band? (block-uniform)
    const bool border = x0 - G::HLX < S || x0 + 64 + G::HLX > W - S || y0 - G::HLY < S || y0 + 32 + G::HLY > H - S;

    {  // stage 0: Lsmooth rows y0-HLY.., cols x0-HLX.. (P0 wide) -> b0 (out-of-image groups are never read by valid outputs)
        const float* L = lsmooth + ibase;
        constexpr int GW = G::P0 / 4, TOT = GW * G::R0, IT = (TOT + 255) / 256;
#pragma unroll
        for (int it = 0; it < IT; it++) {
            const int g = tid + 256 * it;
            if (g < TOT) {
                const int gy = g / GW, gx = g - gy * GW;
                const int y = y0 - G::HLY + gy, x = x0 - G::HLX + 4 * gx;
                float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (!border || (y >= 0 && y < H && x >= 0 && x < W)) v = ld4(L + (size_t)y * W + x);
                st4(b0 + gy * G::P0 + 4 * gx, v);
            }
        }
    }
    __syncthreads();
    {  // stage 1: A = H_main(L) -> b1, Bo = H_off(L) -> b2 : cols x0-HX1.. (P1 wide), all R0 rows
        constexpr int GW = G::P1 / 4, TOT = GW * G::R0, IT = (TOT + 255) / 256;
#pragma unroll
        for (int it = 0; it < IT; it++) {
            const int g = tid + 256 * it;
            if (g < TOT) {
                const int gy = g / GW, gx = g - gy * GW;
                float v[12];
                load12(b0 + gy * G::P0, G::HLX - G::HX1 + 4 * gx, v);
                st4(b1 + gy * G::P1 + 4 * gx, hmain<S>(v, n, wn));
                st4(b2 + gy * G::P1 + 4 * gx, hoff<S>(v));
            }
        }
    }
    __syncthreads();
    if (border) {
        fix_border(b1, G::P1, x0 - G::HX1, y0 - G::HLY, G::P1, G::R0, W, H, S, tid, 256);
        fix_border(b2, G::P1, x0 - G::HX1, y0 - G::HLY, G::P1, G::R0, W, H, S, tid, 256);
    }
    {  // stage 2: Lx = V_off(A) -> b3, Ly = V_main(Bo) -> b4 : rows y0-HY2.. (R2 rows)
        constexpr int GW = G::P1 / 4, TOT = GW * G::R2, IT = (TOT + 255) / 256;
#pragma unroll
        for (int it = 0; it < IT; it++) {
            const int g = tid + 256 * it;
            if (g < TOT) {
                const int gy = g / GW, gx = g - gy * GW;
                const int ry = gy + (G::HLY - G::HY2);  // row in b1/b2
                const float* a = b1 + ry * G::P1 + 4 * gx;
                const float* b = b2 + ry * G::P1 + 4 * gx;
                st4(b3 + gy * G::P1 + 4 * gx, sub4(ld4(a + S * G::P1), ld4(a - S * G::P1)));
                st4(b4 + gy * G::P1 + 4 * gx, vmain(ld4(b - S * G::P1), ld4(b), ld4(b + S * G::P1), n, wn));
            }
        }
    }
    __syncthreads();
    if (border) {
        fix_border(b3, G::P1, x0 - G::HX1, y0 - G::HY2, G::P1, G::R2, W, H, S, tid, 256);
        fix_border(b4, G::P1, x0 - G::HX1, y0 - G::HY2, G::P1, G::R2, W, H, S, tid, 256);
    }
    {  // stage 3: C = H_main(Lx) -> b0, D = H_off(Ly) -> b1, E = H_off(Lx) -> b2 : cols x0-4.. (72 wide), R2 rows;
       // the centre float4 of each group is the final Lx / Ly of that position -> global store for tile pixels
        constexpr int GW = 72 / 4, TOT = GW * G::R2, IT = (TOT + 255) / 256;
        float* ox = oLx + ibase;
        float* oy = oLy + ibase;
#pragma unroll
        for (int it = 0; it < IT; it++) {
            const int g = tid + 256 * it;
            if (g < TOT) {
                const int gy = g / GW, gx = g - gy * GW;
                const int lxp = 4 * gx - 4, lyp = gy - G::HY2;  // tile-local position of the group
                const bool in_tile = lxp >= 0 && lxp < 64 && lyp >= 0 && lyp < 32;
                const size_t o = (size_t)(y0 + lyp) * W + x0 + lxp;
                float v[12];
                load12(b3 + gy * G::P1, G::HX1 - 4 + 4 * gx, v);
                st4(b0 + gy * G::P0 + 4 * gx, hmain<S>(v, n, wn));
                st4(b2 + gy * G::P1 + 4 * gx, hoff<S>(v));
                if (in_tile) st4(ox + o, make_float4(v[4], v[5], v[6], v[7]));
                load12(b4 + gy * G::P1, G::HX1 - 4 + 4 * gx, v);
                st4(b1 + gy * G::P1 + 4 * gx, hoff<S>(v));
                if (in_tile) st4(oy + o, make_float4(v[4], v[5], v[6], v[7]));
            }
        }
    }
    __syncthreads();
    if (border) {
        fix_border(b0, G::P0, x0 - 4, y0 - G::HY2, 72, G::R2, W, H, S, tid, 256);
        fix_border(b1, G::P1, x0 - 4, y0 - G::HY2, 72, G::R2, W, H, S, tid, 256);
        fix_border(b2, G::P1, x0 - 4, y0 - G::HY2, 72, G::R2, W, H, S, tid, 256);
    }
    {  // stage 4: Ldet over cols x0-4.. (72 wide), rows y0-1.. (34 rows) -> b3
        constexpr int GW = 72 / 4, TOT = GW * 34, IT = (TOT + 255) / 256;
#pragma unroll
        for (int it = 0; it < IT; it++) {
            const int g = tid + 256 * it;
            if (g < TOT) {
                const int gy = g / GW, gx = g - gy * GW;
                const int ry = gy + (G::HY2 - 1);  // row in b0/b1/b2 (stage-3 rows)
                const float* c = b0 + ry * G::P0 + 4 * gx;
                const float* d = b1 + ry * G::P1 + 4 * gx;
                const float* e = b2 + ry * G::P1 + 4 * gx;
                const float4 lxx = sub4(ld4(c + S * G::P0), ld4(c - S * G::P0));
                const float4 lyy = vmain(ld4(d - S * G::P1), ld4(d), ld4(d + S * G::P1), n, wn);
                const float4 lxy = vmain(ld4(e - S * G::P1), ld4(e), ld4(e + S * G::P1), n, wn);
                float4 det;  // detector_response.rs:52
                det.x = ((lxx.x * lyy.x) - (lxy.x * lxy.x)) * p.quat;
                det.y = ((lxx.y * lyy.y) - (lxy.y * lxy.y)) * p.quat;
                det.z = ((lxx.z * lyy.z) - (lxy.z * lxy.z)) * p.quat;
                det.w = ((lxx.w * lyy.w) - (lxy.w * lxy.w)) * p.quat;
                st4(b3 + gy * G::P1 + 4 * gx, det);
            }
        }
    }
    __syncthreads();
    // Lxx/Lyy/Lxy are clamp-replicated like every pass output, and so is their pointwise product
    if (border) fix_border(b3, G::P1, x0 - 4, y0 - 1, 72, 34, W, H, S, tid, 256);
    {  // stage 5: Ldet store + threshold + strict 4-neighbour maximum + is_out; b3 holds Ldet at (col lx+4, row ly+1)
        unsigned int* m = mask + (size_t)img * mask_img_words;
        float* od = oLdet + ibase;
        const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int ly = wid + 8 * k;
            const int y = y0 + ly;
#pragma unroll
            for (int hx = 0; hx < 2; hx++) {
                const int lx = lane + 32 * hx, x = x0 + lx;
                const float* q = b3 + (ly + 1) * G::P1 + lx + 4;
                const float v = q[0];
                od[(size_t)y * W + x] = v;
                bool cand = false;
                if (x >= p.xmin && x <= p.xmax && y >= p.ymin && y <= p.ymax)
                    cand = v > p.thr && v > q[1] && v > q[-1] && v > q[-G::P1] && v > q[G::P1];
                const unsigned int bal = __ballot_sync(0xffffffffu, cand);
                if (bal != 0 && lane == 0) {
                    const int xs = x0 + 32 * hx, w0 = xs >> 5, sh = xs & 31;
                    atomicOr(&m[(size_t)y * p.wpr + w0], bal << sh);
                    if (sh != 0 && (bal >> (32 - sh)) != 0) atomicOr(&m[(size_t)y * p.wpr + w0 + 1], bal >> (32 - sh));
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Streaming path (default): one warp owns a strip of 128 columns (4 adjacent columns per lane, float4
// I/O) and marches down the rows, like k_fed. The whole derivative chain is a software pipeline in the
// row index: when Lsmooth row c arrives,
//   A, Bo (row c)            <- H taps of Lsmooth row c (neighbour lanes by shuffle)
//   Lx, Ly (row c-S)         <- V taps of A / Bo rows c-2S, c-S, c
//   C, E, D (row c-S)        <- H taps of that Lx / Ly row
//   Lxx, Lxy, Lyy, Ldet (row c-2S) <- V taps of C / E / D rows c-3S, c-2S, c-S
//   candidates (row c-2S-1)  <- Ldet rows c-2S-2 .. c-2S kept in registers
// The rows a V tap needs again later live in per-lane ring buffers (each lane reads back only what it
// wrote itself: no barriers in the row loop) -- in TENSOR MEMORY in the default kernel (k_detector_tmem,
// RingTmem below), in shared memory in k_detector_stream (RingSmem); nothing is recomputed vertically
// inside a segment, and the ring traffic is 13 128-bit accesses per 4 pixels instead of ~45 shared-memory
// accesses in the tile kernel, which ncu shows to be shared-memory bound (profiles/r1f). fill_border: every pass of
// the reference clamps with the same half-width S, so rows < S / > H-1-S of every intermediate equal
// rows S / H-1-S: ring reads clamp their row index, rows beyond H-1-S reuse the registers of row
// H-1-S, and border rows are written when the row they replicate is produced. Columns: the H-pass
// outputs of columns < S / > W-1-S are replaced by those of columns S / W-1-S (one shuffle per array,
// only in strips that touch the image border); V passes inherit it.
// ------------------------------------------------------------------------------------------------
constexpr int DS_W = 128;
template <int S>
struct StreamGeo {
    static constexpr int HX = ((2 * S + 1) + 3) & ~3;  // x halo: Ldet is needed one column beyond the strip's outputs
    static constexpr int UX = DS_W - 2 * HX;           // output columns per strip
    static constexpr int D = 2 * S;                    // ring depth: exactly the rows a V pass looks back (row c reuses the slot of row c - 2S)
};

// ext[0..3] = left lane's v, ext[4..7] = v, ext[8..11] = right lane's v (only what a +-S tap touches is fetched)
template <int S>
__device__ __forceinline__ void h_neighbours(const float (&v)[4], float (&ext)[12]) {
#pragma unroll
    for (int j = 0; j < 12; j++) ext[j] = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; j++) ext[4 + j] = v[j];
#pragma unroll
    for (int j = 4 - S; j < 4; j++) ext[j] = __shfl_up_sync(0xffffffffu, v[j], 1);
#pragma unroll
    for (int j = 0; j < S; j++) ext[8 + j] = __shfl_down_sync(0xffffffffu, v[j], 1);
}

// out[j] = er[j] - el[j]: for even S the operands pair up as they sit in their float4 registers (packed FSUB2, neither
// operand is a product); for odd S column j and j+1 straddle two float4s and a packed form would need moves
template <int S>
__device__ __forceinline__ void diff4(const float (&er)[4], const float (&el)[4], float (&out)[4]) {
    if (S % 2 == 0) {
        sub2(er[0], er[1], el[0], el[1], out[0], out[1]);
        sub2(er[2], er[3], el[2], el[3], out[2], out[3]);
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) out[j] = er[j] - el[j];
    }
}

// Row c of the pipeline. STEADY: every stage is active, no ring read clamps, no border rows to replicate,
// -> no row tests besides the three warp-uniform flags of DetSteadyPtrs (stores / candidate test on or off), and the ring
// handles of rows c and c-S (which rows c-2S and c-3S share) come in as sl[0..1], advanced by the caller.
// BORDER: the strip touches the first/last S columns of the image. The steady loop is deliberately NOT
// unrolled: an 8x unrolled body (35 KB of SASS per variant) ran 55 % slower on instruction-fetch stalls
// (profiles/r1i: stall_no_inst 29 % of samples).
template <int S>
struct DetStreamCtx {
    int W, H, lane, x0, Ya, Yb, ylo, yhi;
    bool xin, xout, has_l, has_r;
    int lane_l, lane_r;
    unsigned int colmask;  // bit j: column x0+j may emit a candidate
    float n, wn, quat, thr;
    int ymin, ymax, wpr;
    const float* L;
    float *ox, *oy, *od;
    unsigned int* m;
};

template <int S>
struct DetStreamRegs {
    float a[4], bo[4], lx[4], ly[4], cc[4], ee[4], dd[4], det_m[4], det_0[4], det_p[4];
};

template <int S, bool BORDER>
__device__ __forceinline__ void det_fix_cols(const DetStreamCtx<S>& k, float (&v)[4]) {
    if (!BORDER) return;
    constexpr unsigned int FULL = 0xffffffffu;
    if (k.has_l) {
        const float t = __shfl_sync(FULL, v[S & 3], k.lane_l);
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (k.x0 + j < S) v[j] = t;
    }
    if (k.has_r) {
        const float t = __shfl_sync(FULL, v[(3 - S) & 3], k.lane_r & 31);
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (k.x0 + j > k.W - 1 - S) v[j] = t;
    }
}

// row o of an output plane plus the border rows that replicate it (fill_border in y)
template <int S>
__device__ __forceinline__ void det_store_rows(const DetStreamCtx<S>& k, float* plane, int o, const float (&v)[4]) {
    if (!k.xout) return;
    const float4 q = make_float4(v[0], v[1], v[2], v[3]);
    if (o >= k.Ya && o < k.Yb) st4(plane + (size_t)o * k.W + k.x0, q);
    if (o == k.ylo)
        for (int r = max(0, k.Ya); r < min(k.ylo, k.Yb); r++) st4(plane + (size_t)r * k.W + k.x0, q);
    if (o == k.yhi)
        for (int r = max(k.yhi + 1, k.Ya); r < min(k.H, k.Yb); r++) st4(plane + (size_t)r * k.W + k.x0, q);
}

// ---- where the five per-lane rings (A, Bo, C, E, D; 2S rows of one float4 per lane each) live -------------------------
// RingSmem: shared memory, one warp per block (the original layout). RingTmem: TENSOR MEMORY used as a per-lane delay
// line -- each warp of a four-warp CTA owns its quarter of the 128 TMEM lanes, a ring row is four 32-bit columns, and
// tcgen05.ld/st.32x32b.x4 moves exactly the float4-per-thread shape the rings need. The detector is bound by the
// shared-memory/LSU pipe (82 % of its peak, profiles/r1z_ncu_fullload.txt); the ring accesses are 52 of its ~92 pipe
// cycles per row, and tools/microbench/tmem_ring.cu measured the same 8-load/5-store mix at 382 B/clk/SM in TMEM
// against 97 in shared memory. For S = 4 the rings need 160 columns; ring A (one read, one write per row) stays in
// shared memory so that 128 columns suffice and four CTAs (16 warps) fit an SM's 512 columns for every S.
template <int S>
struct RingSmem {
    float4 (*ring)[StreamGeo<S>::D][32];
    int lane;
    // a ring row is named by a handle: here simply the slot index
    __device__ __forceinline__ int handle(int slot) const { return slot; }
    __device__ __forceinline__ int next(int h) const { return (h + 1 == StreamGeo<S>::D) ? 0 : h + 1; }
    template <int A>
    __device__ __forceinline__ float4 load(int h) const { return ring[A][h][lane]; }
    template <int A>
    __device__ __forceinline__ void store(int h, const float4& v) const { ring[A][h][lane] = v; }
    __device__ __forceinline__ int uniform(int h) const { return h; }
    __device__ __forceinline__ void row_begin() const {}
    __device__ __forceinline__ void fence3(float4&, float4&, float4&) const {}
    __device__ __forceinline__ void fence5(float4&, float4&, float4&, float4&, float4&) const {}
};

// NSM = how many of the five rings (in the order A, Bo, C, E, D) live in shared memory instead; the others take
// (5 - NSM) * 2S * 4 tensor-memory columns of the warp's 32 lanes.
template <int S, int NSM>
struct RingTmem {
    static constexpr int D = StreamGeo<S>::D;
    static constexpr int COLS = (5 - NSM) * D * 4;  // tensor-memory columns one warp needs
    unsigned int base;                // TMEM address: this warp's first lane, first of its columns
    float4 (*ring_sm)[D][32];         // [NSM] rings of this warp in shared memory
    int lane;
    // a ring row is named by a handle = the TMEM address of its four columns in the first tensor-memory array; the other
    // arrays sit at compile-time column offsets that go into the instruction's immediate field (tmem[UR + imm]), so a steady
    // row moves two handles to uniform registers instead of one address per access
    __device__ __forceinline__ int handle(int slot) const { return (int)(base + 4u * (unsigned int)slot); }
    __device__ __forceinline__ int next(int h) const { return (h + 4 == (int)(base + 4u * D)) ? (int)base : h + 4; }
    template <int A>
    __device__ __forceinline__ float4 load(int h) const {
        if (A < NSM) return ring_sm[A < NSM ? A : 0][(h - (int)base) >> 2][lane];
        unsigned int r0, r1, r2, r3;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4 + %5];"
                     : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                     : "r"(h), "n"((A < NSM ? 0 : A - NSM) * D * 4)
                     : "memory");
        return make_float4(__uint_as_float(r0), __uint_as_float(r1), __uint_as_float(r2), __uint_as_float(r3));
    }
    template <int A>
    __device__ __forceinline__ void store(int h, const float4& v) const {
        if (A < NSM) {
            ring_sm[A < NSM ? A : 0][(h - (int)base) >> 2][lane] = v;
            return;
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0 + %1], {%2, %3, %4, %5};" ::"r"(h), "n"((A < NSM ? 0 : A - NSM) * D * 4),
                     "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w))
                     : "memory");
    }
    // LDTM/STTM take their address from a UNIFORM register; a handle kept in an ordinary register costs one R2UR per access
    // (23 per row). A warp-wide max of the (already warp-uniform) handle is a single CREDUX that lands in a uniform
    // register, which all accesses of the row then share.
    __device__ __forceinline__ int uniform(int h) const { return (int)__reduce_max_sync(0xffffffffu, (unsigned int)h); }
    // a row's stores are read again S rows later at the earliest: one wait per row covers them
    __device__ __forceinline__ void row_begin() const { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
    // the loaded registers are in/out operands of the wait, so that no use of them can be scheduled above it
    __device__ __forceinline__ void fence3(float4& a, float4& b, float4& c) const {
        asm volatile("tcgen05.wait::ld.sync.aligned;"
                     : "+f"(a.x), "+f"(a.y), "+f"(a.z), "+f"(a.w), "+f"(b.x), "+f"(b.y), "+f"(b.z), "+f"(b.w), "+f"(c.x), "+f"(c.y), "+f"(c.z), "+f"(c.w)
                     :
                     : "memory");
    }
    __device__ __forceinline__ void fence5(float4& a, float4& b, float4& c, float4& d, float4& e) const {
        asm volatile("tcgen05.wait::ld.sync.aligned;"
                     : "+f"(a.x), "+f"(a.y), "+f"(a.z), "+f"(a.w), "+f"(b.x), "+f"(b.y), "+f"(b.z), "+f"(b.w), "+f"(c.x), "+f"(c.y), "+f"(c.z), "+f"(c.w),
                       "+f"(d.x), "+f"(d.y), "+f"(d.z), "+f"(d.w), "+f"(e.x), "+f"(e.y), "+f"(e.z), "+f"(e.w)
                     :
                     : "memory");
    }
};

// where the steady rows store: Lx/Ly row c-S, Ldet row c-2S, mask word of row c-2S-1 (advanced one row per step by the
// caller, so the steady body does no 64-bit index arithmetic)
struct DetSteadyPtrs {
    float *px, *py, *pd;
    unsigned int* pm;
    bool st1, st2, cd;  // this step's Lx/Ly row, Ldet row, candidate row belong to the segment (warp-uniform)
};

template <int S, bool STEADY, bool BORDER, class RG>
__device__ __forceinline__ void det_stream_step(const DetStreamCtx<S>& k, DetStreamRegs<S>& R, const RG& rg,
                                                int c, const float4& Lc, const int (&sl)[2], const DetSteadyPtrs& sp) {
    using G = StreamGeo<S>;
    constexpr unsigned int FULL = 0xffffffffu;
    constexpr int D = G::D;
    const float n = k.n, wn = k.wn;
    const int lane = k.lane;
    // ring slots of rows c, c-S, c-2S, c-3S
    const int s0 = rg.uniform(STEADY ? sl[0] : rg.handle(c % D));
    const int s1 = rg.uniform(STEADY ? sl[1] : rg.handle((c - S) % D));
    const int o1 = c - S, o2 = c - 2 * S, o3 = o2 - 1;
    // ---- A = H_main(Lsmooth), Bo = H_off(Lsmooth), row c (rows beyond yhi keep the registers of row yhi)
    if (STEADY || c <= k.yhi) {
        const float v[4] = {Lc.x, Lc.y, Lc.z, Lc.w};
        float e[12];
        h_neighbours<S>(v, e);
        {
            const float el[4] = {e[4 - S], e[5 - S], e[6 - S], e[7 - S]}, er[4] = {e[4 + S], e[5 + S], e[6 + S], e[7 + S]};
            tap3x4(n, wn, n, el, v, er, R.a);
            diff4<S>(er, el, R.bo);
        }
        det_fix_cols<S, BORDER>(k, R.a);
        det_fix_cols<S, BORDER>(k, R.bo);
    }
    // ---- Lx = V_off(A), Ly = V_main(Bo), row o1 = c - S
    const bool row1 = STEADY || (o1 >= k.ylo && o1 <= k.yhi);
    if (row1) {
        const int rm = STEADY ? s0 : rg.uniform(rg.handle(max(o1 - S, k.ylo) % D));  // row c - 2S shares the slot of row c
        float4 a_m = rg.template load<0>(rm), b_m = rg.template load<1>(rm), b_0 = rg.template load<1>(s1);
        rg.fence3(a_m, b_m, b_0);
        sub2(R.a[0], R.a[1], a_m.x, a_m.y, R.lx[0], R.lx[1]);  // A is a sum, a_m comes out of the ring: nothing to contract
        sub2(R.a[2], R.a[3], a_m.z, a_m.w, R.lx[2], R.lx[3]);
        {
            const float bm[4] = {b_m.x, b_m.y, b_m.z, b_m.w}, b0[4] = {b_0.x, b_0.y, b_0.z, b_0.w};
            tap3x4(n, wn, n, bm, b0, R.bo, R.ly);
        }
        if (STEADY) {
            if (k.xout && sp.st1) {
                st4(sp.px, make_float4(R.lx[0], R.lx[1], R.lx[2], R.lx[3]));
                st4(sp.py, make_float4(R.ly[0], R.ly[1], R.ly[2], R.ly[3]));
            }
        } else {
            det_store_rows<S>(k, k.ox, o1, R.lx);
            det_store_rows<S>(k, k.oy, o1, R.ly);
        }
    }
    if (STEADY || c <= k.yhi) {  // after the reads above: row c may reuse the slot of row c - 2S
        rg.template store<0>(s0, make_float4(R.a[0], R.a[1], R.a[2], R.a[3]));
        rg.template store<1>(s0, make_float4(R.bo[0], R.bo[1], R.bo[2], R.bo[3]));
    }
    // ---- C = H_main(Lx), E = H_off(Lx), D = H_off(Ly), row o1
    if (row1) {
        float e[12];
        h_neighbours<S>(R.lx, e);
        {
            const float el[4] = {e[4 - S], e[5 - S], e[6 - S], e[7 - S]}, er[4] = {e[4 + S], e[5 + S], e[6 + S], e[7 + S]};
            tap3x4(n, wn, n, el, R.lx, er, R.cc);
            diff4<S>(er, el, R.ee);
        }
        h_neighbours<S>(R.ly, e);
        {
            const float el[4] = {e[4 - S], e[5 - S], e[6 - S], e[7 - S]}, er[4] = {e[4 + S], e[5 + S], e[6 + S], e[7 + S]};
            diff4<S>(er, el, R.dd);
        }
        det_fix_cols<S, BORDER>(k, R.cc);
        det_fix_cols<S, BORDER>(k, R.ee);
        det_fix_cols<S, BORDER>(k, R.dd);
    }
    // ---- Lxx = V_off(C), Lxy = V_main(E), Lyy = V_main(D), Ldet, row o2 = c - 2S
    if (STEADY || (o2 >= k.ylo && o2 <= k.yhi)) {
        const int rm = STEADY ? s1 : rg.uniform(rg.handle(max(o2 - S, k.ylo) % D));  // row c - 3S shares the slot of row c - S
        const int r0 = STEADY ? s0 : rg.uniform(rg.handle((o2 + 2 * D) % D));
        float4 c_m = rg.template load<2>(rm), e_m = rg.template load<3>(rm), e_0 = rg.template load<3>(r0);
        float4 d_m = rg.template load<4>(rm), d_0 = rg.template load<4>(r0);
        rg.fence5(c_m, e_m, e_0, d_m, d_0);
        const float cm[4] = {c_m.x, c_m.y, c_m.z, c_m.w}, em[4] = {e_m.x, e_m.y, e_m.z, e_m.w}, e0[4] = {e_0.x, e_0.y, e_0.z, e_0.w};
        const float dm[4] = {d_m.x, d_m.y, d_m.z, d_m.w}, d0[4] = {d_0.x, d_0.y, d_0.z, d_0.w};
        float lyy4[4], lxy4[4];
        tap3x4(n, wn, n, dm, d0, R.dd, lyy4);
        tap3x4(n, wn, n, em, e0, R.ee, lxy4);
        // Ldet = ((Lxx * Lyy) - (Lxy * Lxy)) * s^4 (detector_response.rs:52): Lxx = C - C_m and the three products packed in
        // pairs, the one subtraction of products scalar
        float lxx4[4], p1[4], p2[4], df[4];
        sub2(R.cc[0], R.cc[1], cm[0], cm[1], lxx4[0], lxx4[1]);
        sub2(R.cc[2], R.cc[3], cm[2], cm[3], lxx4[2], lxx4[3]);
        mul2v(lxx4[0], lxx4[1], lyy4[0], lyy4[1], p1[0], p1[1]);
        mul2v(lxx4[2], lxx4[3], lyy4[2], lyy4[3], p1[2], p1[3]);
#ifndef AKZ_FAST_MATH
        mul2v(lxy4[0], lxy4[1], lxy4[0], lxy4[1], p2[0], p2[1]);
        mul2v(lxy4[2], lxy4[3], lxy4[2], lxy4[3], p2[2], p2[3]);
#endif
#pragma unroll
        for (int j = 0; j < 4; j++) {
#ifdef AKZ_FAST_MATH
            df[j] = fmaf(-lxy4[j], lxy4[j], p1[j]);
            (void)p2;
#else
            df[j] = p1[j] - p2[j];
#endif
            R.det_m[j] = R.det_0[j];
            R.det_0[j] = R.det_p[j];
        }
        mul2(df[0], df[1], k.quat, R.det_p[0], R.det_p[1]);
        mul2(df[2], df[3], k.quat, R.det_p[2], R.det_p[3]);
        if (STEADY) {
            if (k.xout && sp.st2) st4(sp.pd, make_float4(R.det_p[0], R.det_p[1], R.det_p[2], R.det_p[3]));
        } else {
            det_store_rows<S>(k, k.od, o2, R.det_p);
        }
    } else if (o2 > k.yhi) {  // Ldet(o2) = Ldet(yhi)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            R.det_m[j] = R.det_0[j];
            R.det_0[j] = R.det_p[j];
        }
    }
    if (row1) {
        rg.template store<2>(s1, make_float4(R.cc[0], R.cc[1], R.cc[2], R.cc[3]));
        rg.template store<3>(s1, make_float4(R.ee[0], R.ee[1], R.ee[2], R.ee[3]));
        rg.template store<4>(s1, make_float4(R.dd[0], R.dd[1], R.dd[2], R.dd[3]));
    }
    // ---- candidates of row o3 = o2 - 1: threshold + strict 4-neighbour maximum + is_out
    if (STEADY ? sp.cd : (o3 >= k.Ya && o3 < k.Yb && o3 >= k.ymin && o3 <= k.ymax)) {
        const float left = __shfl_up_sync(FULL, R.det_0[3], 1), right = __shfl_down_sync(FULL, R.det_0[0], 1);
        unsigned int nib = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float v = R.det_0[j];
            const float l = j > 0 ? R.det_0[j > 0 ? j - 1 : 0] : left;
            const float r = j < 3 ? R.det_0[j < 3 ? j + 1 : 3] : right;
            // v > thr && v > r && v > l && v > up && v > down  ==  v > max(thr, r, l, up, down): Ldet of a u8 image is
            // finite, so no NaN can make the two forms differ (two 3-input FMNMX and one compare instead of five)
            const float hi = fmaxf(fmaxf(fmaxf(k.thr, r), l), fmaxf(R.det_m[j], R.det_p[j]));
            nib |= (v > hi ? 1u : 0u) << j;
        }
        nib &= k.colmask;
        if (nib) {
            if (STEADY) atomicOr(sp.pm, nib << (k.x0 & 31));
            else atomicOr(&k.m[(size_t)o3 * k.wpr + (k.x0 >> 5)], nib << (k.x0 & 31));
        }
    }
}

// Lsmooth rows reach the pipeline through a 4-deep cp.async queue in shared memory (each lane copies and later
// reads only its own 16 bytes: no barrier). Row c+3 is requested while row c is consumed. A register queue does
// not work here: rotating it with moves touches the load's destination in the same step (27 % of all stall
// samples in profiles/r1j), and renaming it by unrolling the loop 4x makes the body miss the instruction cache
// (profiles/r1i, r1k: 40-55 % slower).
__device__ __forceinline__ void cp_async16(float4* smem_dst, const float* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned int)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait3() { asm volatile("cp.async.wait_group 3;" ::: "memory"); }

template <int S, bool BORDER, class RG>
__device__ __forceinline__ void det_stream_run(const DetStreamCtx<S>& k, const RG& rg, float4 (*lq)[32], int c_begin, int c_end) {
    using G = StreamGeo<S>;
    DetStreamRegs<S> R;
#pragma unroll
    for (int j = 0; j < 4; j++)
        R.a[j] = R.bo[j] = R.lx[j] = R.ly[j] = R.cc[j] = R.ee[j] = R.dd[j] = R.det_m[j] = R.det_0[j] = R.det_p[j] = 0.0f;
    const int lane = k.lane;
#pragma unroll
    for (int i = 0; i < 4; i++) lq[i][lane] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);  // lanes outside the image never copy
    auto request_row = [&](int c) {  // one commit group per row, empty when there is nothing to copy
        if (k.xin && c <= k.yhi) cp_async16(&lq[c & 3][lane], k.L + (size_t)c * k.W + k.x0);
        cp_async_commit();
    };
    request_row(c_begin);
    request_row(c_begin + 1);
    request_row(c_begin + 2);
    // rows [c_lo, c_hi] are steady (see det_stream_step); the others run the fully guarded step
    // (a segment's warm-up rows and the rows outside the candidate band are steady rows with their stores and the
    // candidate test switched off by three warp-uniform flags: the guarded step costs ~4x a steady one, and with
    // the flags only the 3S rows at the top and the 3S+4 rows at the bottom of the IMAGE still take it)
    const int c_lo = max(4 * S, c_begin);
    const int c_hi = min(k.yhi - 3, c_end);
    const int lo3 = max(k.Ya, k.ymin), hi3 = min(k.Yb - 1, k.ymax);
    const unsigned int seg_rows = (unsigned int)(k.Yb - k.Ya), cand_rows = hi3 >= lo3 ? (unsigned int)(hi3 - lo3 + 1) : 0u;
    int c = c_begin;
    const int no_slots[2] = {0, 0};
    const DetSteadyPtrs no_ptrs = {nullptr, nullptr, nullptr, nullptr, false, false, false};
    auto generic_until = [&](int stop) {  // rows c .. stop-1
        for (; c < stop; c++) {
            request_row(c + 3);
            cp_async_wait3();
            const float4 Lc = lq[c & 3][lane];
            rg.row_begin();
            det_stream_step<S, false, BORDER>(k, R, rg, c, Lc, no_slots, no_ptrs);
        }
    };
    if (c_lo <= c_hi) {
        generic_until(max(c_lo, c_begin));
        int sl[2] = {rg.handle(c % G::D), rg.handle((c - S) % G::D)};  // c >= 4S here
        const float* pl = k.L + (size_t)(c + 3) * k.W + k.x0;  // row c + 3 <= yhi
        DetSteadyPtrs sp;
        sp.px = k.ox + (ptrdiff_t)(c - S) * k.W + k.x0;
        sp.py = k.oy + (ptrdiff_t)(c - S) * k.W + k.x0;
        sp.pd = k.od + (ptrdiff_t)(c - 2 * S) * k.W + k.x0;
        sp.pm = k.m + (ptrdiff_t)(c - 2 * S - 1) * k.wpr + (k.x0 >> 5);
        auto advance = [&]() {
#pragma unroll
            for (int i = 0; i < 2; i++) sl[i] = rg.next(sl[i]);
        };
#pragma unroll 1
        for (; c <= c_hi; c++) {
            if (k.xin) cp_async16(&lq[(c + 3) & 3][lane], pl);
            cp_async_commit();
            pl += k.W;
            cp_async_wait3();
            const float4 Lc = lq[c & 3][lane];
            sp.st1 = (unsigned int)(c - S - k.Ya) < seg_rows;
            sp.st2 = (unsigned int)(c - 2 * S - k.Ya) < seg_rows;
            sp.cd = (unsigned int)(c - 2 * S - 1 - lo3) < cand_rows;
            rg.row_begin();
            det_stream_step<S, true, BORDER>(k, R, rg, c, Lc, sl, sp);
            advance();
            sp.px += k.W;
            sp.py += k.W;
            sp.pd += k.W;
            sp.pm += k.wpr;
        }
    }
    generic_until(c_end + 1);
}

// strip `si`, segment `sj` of image `img` -> the per-warp context
template <int S>
__device__ __forceinline__ void det_stream_ctx(DetStreamCtx<S>& k, const DetParams& p, int si, int sj, int img, int n_seg, int RL,
                                               const float* lsmooth, size_t img_px, float* oLx, float* oLy, float* oLdet,
                                               unsigned int* mask, size_t mask_img_words) {
    using G = StreamGeo<S>;
    k.lane = threadIdx.x & 31;
    k.W = p.W;
    k.H = p.H;
    k.n = p.n;
    k.wn = p.wn;
    k.quat = p.quat;
    k.thr = p.thr;
    k.ymin = p.ymin;
    k.ymax = p.ymax;
    k.wpr = p.wpr;
    const int xb = si * G::UX - G::HX;  // first column of the strip (multiple of 4, may be negative)
    k.x0 = xb + 4 * k.lane;
    k.Ya = sj * RL;
    k.Yb = (sj == n_seg - 1) ? p.H : k.Ya + RL;
    k.xin = k.x0 >= 0 && k.x0 < p.W;
    k.xout = k.xin && k.x0 >= si * G::UX && k.x0 < (si + 1) * G::UX;
    k.ylo = S;
    k.yhi = p.H - 1 - S;  // rows every pass computes; the others replicate them
    k.has_l = xb < S;
    k.has_r = xb + DS_W - 1 > p.W - 1 - S;
    k.lane_l = (S - xb) >> 2;            // lane holding column S (component S & 3)
    k.lane_r = (p.W - 1 - S - xb) >> 2;  // lane holding column W-1-S (component (3 - S) & 3: W % 4 == 0)
    k.colmask = 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (k.xout && k.x0 + j >= p.xmin && k.x0 + j <= p.xmax) k.colmask |= 1u << j;
    const size_t ibase = (size_t)img * img_px;
    k.L = lsmooth + ibase;
    k.ox = oLx + ibase;
    k.oy = oLy + ibase;
    k.od = oLdet + ibase;
    k.m = mask + (size_t)img * mask_img_words;
}

template <int S>
__global__ void __launch_bounds__(32)
k_detector_stream(const float* __restrict__ lsmooth, size_t img_px, float* __restrict__ oLx, float* __restrict__ oLy,
                  float* __restrict__ oLdet, unsigned int* __restrict__ mask, size_t mask_img_words, DetParams p, int strips_x,
                  int n_seg, int RL) {
    using G = StreamGeo<S>;
    __shared__ float4 ring[5][G::D][32];  // A, Bo, C, E, D
    __shared__ float4 lq[4][32];          // cp.async queue of Lsmooth rows
    DetStreamCtx<S> k;
    det_stream_ctx<S>(k, p, blockIdx.x % strips_x, blockIdx.x / strips_x, blockIdx.z, n_seg, RL, lsmooth, img_px, oLx, oLy, oLdet, mask,
                      mask_img_words);
    const RingSmem<S> rg{ring, k.lane};
    const int c_begin = max(k.ylo, k.Ya - 1 - 2 * S), c_end = k.Yb + 2 * S;
    if (k.has_l || k.has_r)
        det_stream_run<S, true>(k, rg, lq, c_begin, c_end);
    else
        det_stream_run<S, false>(k, rg, lq, c_begin, c_end);
}

// The same stream with the rings in tensor memory (RingTmem). A warp reaches only the 32 tensor-memory lanes of its quarter
// (warp id mod 4), so a CTA of WPC warps stacks WPC / 4 warps per quarter side by side in the columns.
//   WPC = 4:  four warps, 128 columns per CTA, four CTAs per SM = 16 warps (ring A of S = 4 in shared memory)
//   WPC = 20: one CTA per SM with all 512 columns, 20 warps at 96 registers (the kernel needs no more: 0 spill bytes); the
//             rings that do not fit 512 / 5 = 102 columns per warp go to shared memory (none for S = 2, A for S = 3, A and Bo
//             for S = 4)
template <int S, int WPC>
struct DetTmemCfg {
    static constexpr int GROUPS = WPC / 4;  // warps per lane quarter
    static constexpr int NSM = (WPC == 4) ? (S == 4 ? 1 : 0) : (S <= 2 ? 0 : (S == 3 ? 1 : 2));
    using RT = RingTmem<S, NSM>;
    static constexpr int NEED = GROUPS * RT::COLS;
    static constexpr int ALLOC = NEED <= 32 ? 32 : (NEED <= 64 ? 64 : (NEED <= 128 ? 128 : (NEED <= 256 ? 256 : 512)));
    static_assert(NEED <= 512, "the rings of one CTA must fit the SM's 512 tensor-memory columns");
};

template <int S, int WPC>
__global__ void __launch_bounds__(WPC * 32, WPC == 4 ? 4 : 1)
k_detector_tmem(const float* __restrict__ lsmooth, size_t img_px, float* __restrict__ oLx, float* __restrict__ oLy,
                float* __restrict__ oLdet, unsigned int* __restrict__ mask, size_t mask_img_words, DetParams p, int strips_x,
                int n_seg, int RL) {
    using G = StreamGeo<S>;
    using CFG = DetTmemCfg<S, WPC>;
    using RT = typename CFG::RT;
    constexpr int NSM = CFG::NSM;
    extern __shared__ float4 det_smem[];  // [WPC][4][32] cp.async queues of Lsmooth rows, then [WPC][NSM][D][32] rings
    float4(*lq)[4][32] = reinterpret_cast<float4(*)[4][32]>(det_smem);
    float4(*ring_sm)[NSM > 0 ? NSM : 1][G::D][32] = reinterpret_cast<float4(*)[NSM > 0 ? NSM : 1][G::D][32]>(det_smem + WPC * 4 * 32);
    __shared__ unsigned int tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned int)__cvta_generic_to_shared(&tmem_slot)), "r"((unsigned int)CFG::ALLOC) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int strip = blockIdx.x * WPC + warp;
    if (strip < strips_x * n_seg) {  // surplus warps of the last CTA only take part in the barriers
        DetStreamCtx<S> k;
        det_stream_ctx<S>(k, p, strip % strips_x, strip / strips_x, blockIdx.z, n_seg, RL, lsmooth, img_px, oLx, oLy, oLdet, mask,
                          mask_img_words);
        const unsigned int tbase = tmem_slot + ((unsigned int)((warp & 3) * 32) << 16) + (unsigned int)((warp >> 2) * RT::COLS);
        const RT rg{tbase, ring_sm[NSM > 0 ? warp : 0], k.lane};
        const int c_begin = max(k.ylo, k.Ya - 1 - 2 * S), c_end = k.Yb + 2 * S;
        if (k.has_l || k.has_r)
            det_stream_run<S, true>(k, rg, lq[warp], c_begin, c_end);
        else
            det_stream_run<S, false>(k, rg, lq[warp], c_begin, c_end);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"((unsigned int)CFG::ALLOC) : "memory");
}

template <int S, int WPC>
static size_t det_tmem_smem_bytes() {
    return sizeof(float4) * 32 * ((size_t)WPC * 4 + (size_t)WPC * DetTmemCfg<S, WPC>::NSM * StreamGeo<S>::D);
}

// ---- bitmask -> ordered list ------------------------------------------------------------------
// rows are enumerated level-major; one warp per row
__device__ __forceinline__ int find_level(const PlanDev* plan, int grow, int* row_in_level) {
    int l = 0, base = 0;
    while (l < plan->n_levels - 1 && grow >= base + plan->lv[l].h) {
        base += plan->lv[l].h;
        l++;
    }
    *row_in_level = grow - base;
    return l;
}

// (the three kernels work on the global rows [row0, row1): the levels [l0, l1) of one launch_compact call)
__global__ void k_rowcount(const unsigned int* __restrict__ mask, const PlanDev* __restrict__ plan,
                           unsigned int* __restrict__ rowcount, int total_rows, int row0, int row1) {
    const int lane = threadIdx.x & 31;
    const int grow = row0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int img = blockIdx.y;
    if (grow >= row1) return;
    int y;
    const int l = find_level(plan, grow, &y);
    const LevelDev& lv = plan->lv[l];
    unsigned int cnt = 0;
    if (y >= lv.ymin && y <= lv.ymax) {
        const unsigned int* m = mask + (size_t)img * plan->mask_words + lv.mask_off + (size_t)y * lv.wpr;
        for (int w = lane; w < lv.wpr; w += 32) cnt += __popc(m[w]);
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) rowcount[(size_t)img * total_rows + grow] = cnt;
}

// one block per image: exclusive scan of the row counts, per-level offsets, total
__global__ void __launch_bounds__(1024)
k_rowscan(unsigned int* __restrict__ rowcount, const PlanDev* __restrict__ plan, unsigned int* __restrict__ level_off,
          unsigned int* __restrict__ n_total, unsigned int* __restrict__ err_flags, int total_rows, unsigned int cand_cap, int row0,
          int row1, int l0, int l1) {
    __shared__ unsigned int warp_sums[32];
    __shared__ unsigned int carry;
    const int img = blockIdx.x;
    unsigned int* rc = rowcount + (size_t)img * total_rows;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = l0 > 0 ? level_off[(size_t)img * (kMaxLevels + 1) + l0] : 0u;  // where the levels before l0 ended
    __syncthreads();
    for (int base = row0; base < row1; base += 1024) {
        const int i = base + tid;
        const unsigned int v = (i < row1) ? rc[i] : 0;
        unsigned int incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            unsigned int ws = warp_sums[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int t = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += t;
            }
            warp_sums[lane] = ws;  // inclusive
        }
        __syncthreads();
        const unsigned int woff = wid ? warp_sums[wid - 1] : 0;
        const unsigned int excl = carry + woff + incl - v;
        if (i < row1) rc[i] = excl;
        __syncthreads();
        if (tid == 1023) carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) {
        // per-level offsets from the scanned rows; the entry after the last level of this call is the running total
        int row = row0;
        for (int l = l0; l < l1; l++) {
            level_off[(size_t)img * (kMaxLevels + 1) + l] = (row < row1) ? rc[row] : carry;
            row += plan->lv[l].h;
        }
        level_off[(size_t)img * (kMaxLevels + 1) + l1] = carry;
        if (l1 == plan->n_levels) {
            for (int l = l1; l <= kMaxLevels; l++) level_off[(size_t)img * (kMaxLevels + 1) + l] = carry;
            n_total[img] = carry;
        }
        if (carry > cand_cap) atomicOr(&err_flags[img], (unsigned int)kErrCandOverflow);
    }
}

__global__ void k_scatter(const unsigned int* __restrict__ mask, const PlanDev* __restrict__ plan,
                          const unsigned int* __restrict__ rowoff, unsigned int* __restrict__ cand, int total_rows,
                          unsigned int cand_cap, int row0, int row1) {
    const int lane = threadIdx.x & 31;
    const int grow = row0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int img = blockIdx.y;
    if (grow >= row1) return;
    int y;
    const int l = find_level(plan, grow, &y);
    const LevelDev& lv = plan->lv[l];
    if (y < lv.ymin || y > lv.ymax) return;
    const unsigned int* m = mask + (size_t)img * plan->mask_words + lv.mask_off + (size_t)y * lv.wpr;
    unsigned int pos = rowoff[(size_t)img * total_rows + grow];
    unsigned int* out = cand + (size_t)img * cand_cap;
    for (int wb = 0; wb < lv.wpr; wb += 32) {
        const int w = wb + lane;
        unsigned int bits = (w < lv.wpr) ? m[w] : 0u;
        const unsigned int c = __popc(bits);
        unsigned int incl = c;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        unsigned int at = pos + incl - c;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            if (at < cand_cap) out[at] = (unsigned int)(y * lv.w + w * 32 + b);
            at++;
        }
        pos += __shfl_sync(0xffffffffu, incl, 31);
    }
}

}  // namespace

// opt in to > 48 KB dynamic shared memory (call once per device, after cudaSetDevice)
template <int S>
static cudaError_t det_tmem_attributes() {
    cudaError_t e = cudaFuncSetAttribute(k_detector_tmem<S, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)det_tmem_smem_bytes<S, 20>());
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_detector_tmem<S, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)det_tmem_smem_bytes<S, 4>());
}

cudaError_t init_detector_attributes() {
    {
        cudaError_t e0 = det_tmem_attributes<2>();
        if (e0 != cudaSuccess) return e0;
        e0 = det_tmem_attributes<3>();
        if (e0 != cudaSuccess) return e0;
        e0 = det_tmem_attributes<4>();
        if (e0 != cudaSuccess) return e0;
    }
    const int pw = DT + 2 * (2 * kMaxDetScale + 1);
    cudaError_t e = cudaFuncSetAttribute(k_detector, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * pw * pw * (int)sizeof(float));
    if (e != cudaSuccess) return e;
    // all pipeline kernels ask for the largest shared-memory carveout: an SM cannot change its L1/shared split while any
    // block is resident, and the cache pass (k_dedup_smem, one long-lived warp per image on every SM) used to pin the
    // small split chosen for it, which cut the streaming kernels' resident warps for as long as it ran
    if (getenv("AKZ_NO_CARVEOUT") == nullptr) {
        e = cudaFuncSetAttribute(k_detector_stream<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_detector_stream<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_detector_stream<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_rowcount, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_rowscan, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_scatter, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
    }
    e = cudaFuncSetAttribute(k_detector_fast<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, DetGeo<2>::FLOATS * (int)sizeof(float));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_detector_fast<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, DetGeo<3>::FLOATS * (int)sizeof(float));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_detector_fast<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, DetGeo<4>::FLOATS * (int)sizeof(float));
}

int launch_detector(const Launch& L, const Plan& P, const Buffers& B, int level) {
    const LevelDev& lv = P.dev.lv[level];
    DetParams p;
    p.W = lv.w;
    p.H = lv.h;
    p.s = lv.s_det;
    p.n = P.sch_n[lv.s_det];
    p.wn = P.sch_wn[lv.s_det];
    const uint32_t s = (uint32_t)lv.s_det;
    p.quat = (float)(s * s * s * s);
    p.thr = P.dev.det_threshold;
    p.xmin = lv.xmin;
    p.xmax = lv.xmax;
    p.ymin = lv.ymin;
    p.ymax = lv.ymax;
    p.wpr = lv.wpr;
    const size_t img_px = (size_t)lv.w * lv.h;
    // float4 kernel: rows and image slabs 16-byte aligned, compile-time Scharr scale; otherwise (odd widths,
    // unusual scales, or keep-evolutions mode that also wants Lxx/Lyy/Lxy) the generic kernel
    const bool fast = lv.w % 4 == 0 && img_px % 4 == 0 && lv.s_det >= 2 && lv.s_det <= 4 && !B.keep;
    p.fast = fast ? 1 : 0;
    const size_t off = (size_t)lv.off * L.batch;
    // level 0: Lsmooth is Lt (lib.rs:58)
    const float* ls = (level == 0) ? B.Lt : lsmooth_slab(P, B, L.batch, level);
    unsigned int* mk = B.mask + lv.mask_off;
    // streaming kernel: additionally needs the candidate rows (and their two neighbours) to be rows the passes compute
    static const bool force_tile = getenv("AKZ_DETECTOR_TILE") != nullptr;  // A/B switch for profiling
    const bool stream = fast && !force_tile && lv.ymin >= lv.s_det + 1 && lv.ymax <= lv.h - 2 - lv.s_det && lv.h >= 4 * lv.s_det + 4;
    if (stream) {
        // segment height: every segment re-runs 4S+2 warm-up rows, so take the tallest one that still gives the GPU a
        // few waves of one-warp blocks (AKZ_DET_RL overrides, for profiling)
        static const int rl_env = getenv("AKZ_DET_RL") ? atoi(getenv("AKZ_DET_RL")) : 0;
        int RL = 256;
        {
            const int sx128 = (lv.w + 103) / 104;
            auto warps = [&](int rl) { return (long long)sx128 * std::max(1, (lv.h - 1 - lv.s_det) / rl) * L.batch; };
            while (RL > 32 && warps(RL) < 148LL * 16 * 3) RL >>= 1;
            // a single image or a small batch: down to 8-row segments (2-3 x the rows per segment with their warm-up, but the
            // kernel is as long as ONE warp's walk) while the grid is still short of one wave
            static const bool no_short = getenv("AKZ_NO_SHORT_SEGMENTS") != nullptr;  // A/B switch
            while (!no_short && RL > min_segment_rows() && warps(RL) < 148LL * 16) RL >>= 1;
        }
        if (rl_env > 0) RL = rl_env;
        // equal-height segments (the last one used to take the remainder: 312 rows against 256 at 1080p, and a CTA of the
        // tensor-memory kernel waits for its slowest warp)
        int n_seg = std::max(1, (lv.h + RL / 2) / RL);
        RL = (lv.h + n_seg - 1) / n_seg;
        while (n_seg > 1 && lv.h - (n_seg - 1) * RL < 2 * lv.s_det + 4) n_seg--;  // the last segment takes what is left

        float* px = B.Lx + off;
        float* py = B.Ly + off;
        float* pd = B.Ldet + off;
        static const bool rings_in_smem = getenv("AKZ_DET_SMEM") != nullptr;  // A/B switch: the shared-memory ring kernel
        auto go = [&](auto s_tag) {
            constexpr int S = decltype(s_tag)::value;
            const int sx = (lv.w + StreamGeo<S>::UX - 1) / StreamGeo<S>::UX;
            if (rings_in_smem)
                k_detector_stream<S><<<dim3(sx * n_seg, 1, L.batch), 32, 0, L.stream>>>(ls, img_px, px, py, pd, mk, (size_t)P.dev.mask_words, p, sx, n_seg, RL);
            else
            {
                // Default: four-warp CTAs, four per SM (16 warps). AKZ_DET_WPC=20: one 20-warp CTA per SM with the warps' ring
                // columns packed side by side (96 registers, no spills) -- 25 % more resident warps, and measured SLOWER:
                // detector 0.0424 -> 0.0482 ms per image (profiles/r2_ab.txt). One block per SM lives as long as its slowest
                // of 20 warps, and for S >= 3 one or two of the five rings fall back to shared memory, the pipe the tensor-
                // memory rings were introduced to relieve.
                static const int wpc = getenv("AKZ_DET_WPC") ? atoi(getenv("AKZ_DET_WPC")) : 4;
                if (wpc == 4)
                    k_detector_tmem<S, 4><<<dim3((sx * n_seg + 3) / 4, 1, L.batch), 4 * 32, det_tmem_smem_bytes<S, 4>(), L.stream>>>(ls, img_px, px, py, pd, mk, (size_t)P.dev.mask_words, p, sx, n_seg, RL);
                else
                    k_detector_tmem<S, 20><<<dim3((sx * n_seg + 19) / 20, 1, L.batch), 20 * 32, det_tmem_smem_bytes<S, 20>(), L.stream>>>(ls, img_px, px, py, pd, mk, (size_t)P.dev.mask_words, p, sx, n_seg, RL);
            }
        };
        if (lv.s_det == 2) go(std::integral_constant<int, 2>{});
        else if (lv.s_det == 3) go(std::integral_constant<int, 3>{});
        else go(std::integral_constant<int, 4>{});
        return 1;
    }
    if (fast) {
        dim3 gf((lv.w + 63) / 64, (lv.h + 31) / 32, L.batch);
        if (lv.s_det == 2)
            k_detector_fast<2><<<gf, 256, DetGeo<2>::FLOATS * sizeof(float), L.stream>>>(ls, img_px, B.Lx + off, B.Ly + off, B.Ldet + off, mk, (size_t)P.dev.mask_words, p);
        else if (lv.s_det == 3)
            k_detector_fast<3><<<gf, 256, DetGeo<3>::FLOATS * sizeof(float), L.stream>>>(ls, img_px, B.Lx + off, B.Ly + off, B.Ldet + off, mk, (size_t)P.dev.mask_words, p);
        else
            k_detector_fast<4><<<gf, 256, DetGeo<4>::FLOATS * sizeof(float), L.stream>>>(ls, img_px, B.Lx + off, B.Ly + off, B.Ldet + off, mk, (size_t)P.dev.mask_words, p);
        return 1;
    }
    const int HL = 2 * lv.s_det + 1, PW = DT + 2 * HL;
    const size_t smem = (size_t)5 * PW * PW * sizeof(float);
    float* xx = B.keep ? B.Lxx + off : nullptr;
    float* yy = B.keep ? B.Lyy + off : nullptr;
    float* xy = B.keep ? B.Lxy + off : nullptr;
    dim3 grid((lv.w + DT - 1) / DT, (lv.h + DT - 1) / DT, L.batch);
    k_detector<<<grid, dim3(NTX, NTY), smem, L.stream>>>(ls, img_px, B.Lx + off, B.Ly + off, B.Ldet + off, xx, yy, xy, mk,
                                                         (size_t)P.dev.mask_words, p);
    return 1;
}

int launch_compact(const Launch& L, const Plan& P, const Buffers& B, int l0, int l1) {
    // candidate lists of the levels [l0, l1), appended behind those of the levels before l0 (compacted by an earlier call)
    if (l1 < 0) l1 = P.dev.n_levels;
    int total_rows = 0, row0 = 0, row1 = 0;
    for (int l = 0; l < P.dev.n_levels; l++) {
        if (l < l0) row0 += P.dev.lv[l].h;
        if (l < l1) row1 += P.dev.lv[l].h;
        total_rows += P.dev.lv[l].h;
    }
    unsigned int* rowcount = B.rowcount;  // [B][total_rows]
    const int wpb = 8;
    dim3 grid((row1 - row0 + wpb - 1) / wpb, L.batch);
    k_rowcount<<<grid, wpb * 32, 0, L.stream>>>(B.mask, B.plan_dev, rowcount, total_rows, row0, row1);
    k_rowscan<<<L.batch, 1024, 0, L.stream>>>(rowcount, B.plan_dev, B.cand_level_count, B.n_cand_total, B.err_flags,
                                              total_rows, L.cand_cap, row0, row1, l0, l1);
    k_scatter<<<grid, wpb * 32, 0, L.stream>>>(B.mask, B.plan_dev, rowcount, B.cand, total_rows, L.cand_cap, row0, row1);
    return 3;
}

}  // namespace akz
