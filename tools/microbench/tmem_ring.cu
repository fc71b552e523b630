// Can tensor memory serve as a per-lane delay line? Throughput of tcgen05.ld / tcgen05.st .32x32b.x4 (one float4 per
// thread, the access shape of the detector's shared-memory rings) against LDS.128 / STS.128, per SM, with 4-warp CTAs
// (each warp owns its quarter of the 128 TMEM lanes) and 1..4 CTAs per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ring tmem_ring.cu
#include <cstdio>
#include <algorithm>
#include <cuda_runtime.h>
#define N_IT 2048
__device__ __forceinline__ float4 lds128(const float4* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned int)__cvta_generic_to_shared(p)));
    return v;
}
__device__ __forceinline__ void sts128(float4* p, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"((unsigned int)__cvta_generic_to_shared(p)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(128) k(float* out, long long* cyc, int ncols) {
    __shared__ unsigned int slot;
    __shared__ float4 ring[8][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned int)__cvta_generic_to_shared(&slot)), "r"((unsigned int)ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned int base = slot + ((unsigned int)(warp * 32) << 16);  // this warp's lanes
    float4 v = make_float4(threadIdx.x, 1.f, 2.f, 3.f), acc = make_float4(0, 0, 0, 0);
    // initialise 8 ring rows
    for (int r = 0; r < 8; r++) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + r * 4), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)) : "memory");
        ring[r][threadIdx.x] = v;
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    long long t0 = clock64();
    int s = 0;
#pragma unroll 1
    for (int it = 0; it < N_IT; it++) {
        if (MODE == 0) {  // 8 TMEM loads (x4) + 1 wait
            unsigned int r[32];
#pragma unroll
            for (int q = 0; q < 8; q++)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[4 * q]), "=r"(r[4 * q + 1]), "=r"(r[4 * q + 2]), "=r"(r[4 * q + 3]) : "r"(base + ((s + q) & 7) * 4));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 8; q++) acc.x += __uint_as_float(r[4 * q]) + __uint_as_float(r[4 * q + 3]);
        } else if (MODE == 1) {  // 8 TMEM stores (x4) + 1 wait
#pragma unroll
            for (int q = 0; q < 8; q++)
                asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + ((s + q) & 7) * 4), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(acc.x)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            acc.x += 1.0f;
        } else if (MODE == 2) {  // 8 LDS.128
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const float4 t = lds128(&ring[(s + q) & 7][threadIdx.x]);
                acc.x += t.x + t.w;
            }
        } else if (MODE == 3) {  // 8 STS.128
#pragma unroll
            for (int q = 0; q < 8; q++) sts128(&ring[(s + q) & 7][threadIdx.x], make_float4(v.x, acc.x, v.z, v.w));
            acc.x += 1.0f;
        } else if (MODE == 4) {  // the detector's mix in TMEM: 8 loads, wait, 5 stores (store wait at the top of the next iteration)
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            unsigned int r[32];
#pragma unroll
            for (int q = 0; q < 8; q++)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[4 * q]), "=r"(r[4 * q + 1]), "=r"(r[4 * q + 2]), "=r"(r[4 * q + 3]) : "r"(base + ((s + q) & 7) * 4));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 8; q++) acc.x += __uint_as_float(r[4 * q]) + __uint_as_float(r[4 * q + 3]);
#pragma unroll
            for (int q = 0; q < 5; q++)
                asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + ((s + q) & 7) * 4), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(acc.x)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)) : "memory");
        } else if (MODE == 5) {  // the same mix in shared memory
            float4 t[8];
#pragma unroll
            for (int q = 0; q < 8; q++) t[q] = lds128(&ring[(s + q) & 7][threadIdx.x]);
#pragma unroll
            for (int q = 0; q < 8; q++) acc.x += t[q].x + t[q].w;
#pragma unroll
            for (int q = 0; q < 5; q++) sts128(&ring[(s + q) & 7][threadIdx.x], make_float4(v.x, acc.x, v.z, v.w));
        }
        s = (s + 1) & 7;
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x;
    if (lane == 0) {
        cyc[(blockIdx.x * 4 + warp) * 2] = t0;
        cyc[(blockIdx.x * 4 + warp) * 2 + 1] = t1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"((unsigned int)ncols) : "memory");
}
template <int MODE>
void run(const char* name, int acc_per_it) {
    float* out; long long* cyc;
    const int max_blocks = 148 * 4;
    cudaMalloc(&out, max_blocks * 128 * sizeof(float));
    cudaMalloc(&cyc, max_blocks * 8 * sizeof(long long));
    printf("%-52s", name);
    for (int bps : {1, 2, 4}) {  // CTAs per SM (grid = 148 * bps; each CTA allocates 128 TMEM columns)
        k<MODE><<<148 * bps, 128>>>(out, cyc, 128);
        k<MODE><<<148 * bps, 128>>>(out, cyc, 128);
        cudaError_t e = cudaDeviceSynchronize();
        if (e) { printf("err %s\n", cudaGetErrorString(e)); return; }
        static long long h[148 * 4 * 8];
        cudaMemcpy(h, cyc, 148 * bps * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
        // per CTA: cycles from first start to last end; bytes moved per CTA = 4 warps * 512 B * accesses
        double mc = 0;
        for (int b = 0; b < 148 * bps; b++) {
            long long lo = h[b * 8], hi = h[b * 8 + 1];
            for (int w = 0; w < 4; w++) { lo = std::min(lo, h[(b * 4 + w) * 2]); hi = std::max(hi, h[(b * 4 + w) * 2 + 1]); }
            mc += (double)(hi - lo);
        }
        mc /= 148 * bps;
        const double bytes_per_sm = (double)bps * 4 * 512.0 * acc_per_it * N_IT;  // assumes the bps CTAs of an SM run concurrently
        printf("  %d CTA/SM: %7.1f B/clk/SM", bps, bytes_per_sm / mc);
    }
    printf("\n");
}
int main() {
    run<0>("8 tcgen05.ld.32x32b.x4 + wait", 8);
    run<1>("8 tcgen05.st.32x32b.x4 + wait", 8);
    run<2>("8 LDS.128", 8);
    run<3>("8 STS.128", 8);
    run<4>("detector mix in TMEM (8 ld, wait, 5 st)", 13);
    run<5>("detector mix in shared memory (8 LDS.128, 5 STS.128)", 13);
    return 0;
}
