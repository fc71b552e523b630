// matcher_tc.cu -- brute-force Hamming top-2 on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same contract as matcher.cu (descriptor_match's inner loops, akaze/src/ops/feature_matching.rs:37-50,
// 113-123): per query the two smallest Hamming distances over the database and the LOWEST database index
// attaining the minimum. The integer-popc kernel of matcher.cu is bounded by 16 POPC/clk/SM; this path
// turns the distance matrix into an exact int8 GEMM instead:
//
//   every descriptor bit becomes one int8: query bit b -> +1 / -1, database bit b -> -1 / +1, so
//   acc[i][j] = sum_k A[i][k] * B[j][k] = (#differing bits) - (#equal bits) = 2 * ham(i, j) - 512
//   over the 512 bit positions of a padded 64-byte row (padding bits are equal, they cancel exactly).
//   Accumulation is s32 in TMEM: no rounding anywhere, distances are bit-exact.
//
// Operands are expanded once per call into "shared-memory images": 128 descriptors x 512 int8 as four
// K-chunks of [128 rows x 128 B] in the canonical K-major SWIZZLE_128B layout, so that one plain bulk copy
// (cp.async.bulk, no tensor map) lands a ready-to-use MMA operand.
//
// Kernel: one CTA owns TC_MT query tiles of 128 rows (resident in shared memory for the whole kernel) and
// streams every database tile of its part through a 4-stage mbarrier ring. Warp-specialised:
//   warp 0    bulk-copy producer
//   warp 1    TMEM allocation + single-thread tcgen05.mma issue (kind::i8, M=128, N=128, K=32 x 16)
//   warps 2-5 epilogue: tcgen05.ld of the s32 accumulators (row = TMEM lane = thread), key = acc*64 + column,
//             running two smallest keys per row with VIMNMX/VIMNMX3, merged into (best, second, index) per tile
// Accumulators are double-buffered in TMEM (2 x TC_MT x 128 columns = all 512), so the epilogue of tile t
// overlaps the MMAs of tile t+1.
#include "common.cuh"

namespace akz {
namespace {

constexpr int TC_ROWS = 128;                   // descriptors per operand tile (= UMMA M = UMMA N)
constexpr int TC_CHUNK_BYTES = TC_ROWS * 128;  // one K-chunk: 128 rows x 128 int8
constexpr int TC_TILE_BYTES = 4 * TC_CHUNK_BYTES;  // 64 KB image per 128 descriptors
constexpr int TC_MT = 2;                       // query tiles per CTA
constexpr int TC_STAGES = 4;                   // database K-chunk ring
constexpr int TC_THREADS = 192;
constexpr int TC_SMEM = TC_MT * TC_TILE_BYTES + TC_STAGES * TC_CHUNK_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr unsigned int kSentinel = 10000u;     // distance_threshold hard-wired by lib.rs:264
constexpr int kHamBias = 256;                  // ham = (acc >> 1) + 256

// ---- operand expansion -----------------------------------------------------------------------------
// one thread per (descriptor, 16-bit group): 2 packed bytes -> 16 int8 at the swizzled position
template <bool IS_DB>
__global__ void k_match_expand(const uint8_t* __restrict__ rows, unsigned long long n, unsigned long long n_padded,
                               uint8_t* __restrict__ image) {
    const unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long d = gid >> 5;
    if (d >= n_padded) return;
    const int grp = (int)(gid & 31);  // 16-bit group 0..31 of the 512-bit row
    const int chunk = grp >> 3, g = grp & 7;
    const unsigned long long tile = d >> 7;
    const int r = (int)(d & 127);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (d < n) {
        const unsigned int bits = *reinterpret_cast<const unsigned short*>(rows + d * kDescStride + 2 * grp);
        // bit -> byte mask (0x01 per set bit), then +1/-1: queries 1 -> 0x01, 0 -> 0xFF; database 1 -> 0xFF, 0 -> 0x01
        const unsigned int base = IS_DB ? 0x01010101u : 0xFFFFFFFFu;
        unsigned int w[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const unsigned int m = (((bits >> (4 * i)) & 0xFu) * 0x00204081u) & 0x01010101u;
            w[i] = base ^ (m * 0xFEu);
        }
        v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    const size_t off = (size_t)tile * TC_TILE_BYTES + (size_t)chunk * TC_CHUNK_BYTES + (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128 +
                       (size_t)((g ^ (r & 7)) * 16);
    *reinterpret_cast<uint4*>(image + off) = v;
}

// ---- PTX helpers --------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded spin: a protocol bug must not hang the GPU (traps after ~seconds instead)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (unsigned int spin = 0; spin < (1u << 24); spin++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, SWIZZLE_128B operand descriptor: rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);  // start address
    d |= (uint64_t)1 << 16;                   // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset
    d |= (uint64_t)1 << 46;                   // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
    return d;
}
// kind::i8 instruction descriptor: D = s32, A = B = signed int8, both K-major, M = 128, N = 128
constexpr uint32_t kIdescI8 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_ROWS >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);

#define TC_LD32(R, TADDR)                                                                                                       \
    asm volatile(                                                                                                               \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                               \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "  \
        "%25, %26, %27, %28, %29, %30, %31}, [%32];"                                                                            \
        : "=r"(R[0]), "=r"(R[1]), "=r"(R[2]), "=r"(R[3]), "=r"(R[4]), "=r"(R[5]), "=r"(R[6]), "=r"(R[7]), "=r"(R[8]),           \
          "=r"(R[9]), "=r"(R[10]), "=r"(R[11]), "=r"(R[12]), "=r"(R[13]), "=r"(R[14]), "=r"(R[15]), "=r"(R[16]), "=r"(R[17]),   \
          "=r"(R[18]), "=r"(R[19]), "=r"(R[20]), "=r"(R[21]), "=r"(R[22]), "=r"(R[23]), "=r"(R[24]), "=r"(R[25]), "=r"(R[26]),  \
          "=r"(R[27]), "=r"(R[28]), "=r"(R[29]), "=r"(R[30]), "=r"(R[31])                                                       \
        : "r"(TADDR)                                                                                                            \
        : "memory")

// the registers of the loads in flight are in/out operands, so that no use of them can be scheduled above the wait
#define TC_WAIT_LD(R)                                                                                                          \
    asm volatile("tcgen05.wait::ld.sync.aligned;"                                                                              \
                 : "+r"(R[0]), "+r"(R[1]), "+r"(R[2]), "+r"(R[3]), "+r"(R[4]), "+r"(R[5]), "+r"(R[6]), "+r"(R[7]), "+r"(R[8]), \
                   "+r"(R[9]), "+r"(R[10]), "+r"(R[11]), "+r"(R[12]), "+r"(R[13]), "+r"(R[14]), "+r"(R[15]), "+r"(R[16]),      \
                   "+r"(R[17]), "+r"(R[18]), "+r"(R[19]), "+r"(R[20]), "+r"(R[21]), "+r"(R[22]), "+r"(R[23]), "+r"(R[24]),     \
                   "+r"(R[25]), "+r"(R[26]), "+r"(R[27]), "+r"(R[28]), "+r"(R[29]), "+r"(R[30]), "+r"(R[31])                   \
                 :                                                                                                             \
                 : "memory")

// two smallest keys of {k1, k2} U the 32 keys acc[c]*64 + col0 + c (acc is even: key order = (ham, column) order)
template <bool MASKED>
__device__ __forceinline__ void top2_32(const uint32_t (&r)[32], int col0, int valid, int& k1, int& k2) {
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
        int a = (int)r[c] * 64 + (col0 + c);
        int b = (int)r[c + 1] * 64 + (col0 + c + 1);
        if (MASKED) {
            if (col0 + c >= valid) a = 0x7fffffff;
            if (col0 + c + 1 >= valid) b = 0x7fffffff;
        }
        const int mab = min(a, b), Mab = max(a, b);
        const int t1 = max(k1, mab);
        k1 = min(k1, mab);
        k2 = min(min(k2, Mab), t1);
    }
}

// q_img: TC_MT consecutive query tiles per blockIdx.x; db_img: database tiles; out[part][nq]
__global__ void __launch_bounds__(TC_THREADS, 1)
k_match_tc(const uint8_t* __restrict__ q_img, unsigned long long nq, const uint8_t* __restrict__ db_img, unsigned long long ndb,
           unsigned int tiles_per_part, unsigned int db_index_base, akz_top2* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = base;                                   // TC_MT tiles, resident
    const uint32_t sB = sA + TC_MT * TC_TILE_BYTES;             // TC_STAGES K-chunks
    const uint32_t bars = sB + TC_STAGES * TC_CHUNK_BYTES;
    const uint32_t bar_a = bars;                                // A landed
    const uint32_t bar_full = bars + 8;                         // [TC_STAGES] chunk landed
    const uint32_t bar_empty = bar_full + 8 * TC_STAGES;        // [TC_STAGES] chunk consumed by the MMAs
    const uint32_t bar_tfull = bar_empty + 8 * TC_STAGES;       // [2] accumulators complete
    const uint32_t bar_tempty = bar_tfull + 16;                 // [2] accumulators drained by the epilogue
    const uint32_t tmem_slot = bar_tempty + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const unsigned int n_db_tiles = (unsigned int)((ndb + TC_ROWS - 1) / TC_ROWS);
    const unsigned int t_beg = blockIdx.y * tiles_per_part;
    const unsigned int t_end = min(n_db_tiles, t_beg + tiles_per_part);
    const unsigned int n_tiles = t_end > t_beg ? t_end - t_beg : 0;

    if (threadIdx.x == 0) {
        mbar_init(bar_a, 1);
        for (int s = 0; s < TC_STAGES; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(bar_tfull + 8 * s, 1);
            mbar_init(bar_tempty + 8 * s, 4);  // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        // ===== producer: query tiles once, then the database K-chunks =====
        if (lane == 0) {
            mbar_expect_tx(bar_a, TC_MT * TC_TILE_BYTES);
            const uint8_t* qa = q_img + (size_t)blockIdx.x * TC_MT * TC_TILE_BYTES;
            for (int i = 0; i < TC_MT * 4; i++) bulk_g2s(sA + i * TC_CHUNK_BYTES, qa + (size_t)i * TC_CHUNK_BYTES, TC_CHUNK_BYTES, bar_a);
            unsigned int it = 0;
            for (unsigned int t = 0; t < n_tiles; t++) {
                const uint8_t* src = db_img + (size_t)(t_beg + t) * TC_TILE_BYTES;
                for (int c = 0; c < 4; c++, it++) {
                    const int s = it % TC_STAGES;
                    mbar_wait(bar_empty + 8 * s, ((it / TC_STAGES) & 1) ^ 1);
                    mbar_expect_tx(bar_full + 8 * s, TC_CHUNK_BYTES);
                    bulk_g2s(sB + s * TC_CHUNK_BYTES, src + (size_t)c * TC_CHUNK_BYTES, TC_CHUNK_BYTES, bar_full + 8 * s);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            mbar_wait(bar_a, 0);
            unsigned int it = 0;
            for (unsigned int t = 0; t < n_tiles; t++) {
                const int as = t & 1;
                mbar_wait(bar_tempty + 8 * as, ((t >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int c = 0; c < 4; c++, it++) {
                    const int s = it % TC_STAGES;
                    mbar_wait(bar_full + 8 * s, (it / TC_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t bsm = sB + s * TC_CHUNK_BYTES;
#pragma unroll
                    for (int m = 0; m < TC_MT; m++) {
                        const uint32_t asm_ = sA + m * TC_TILE_BYTES + c * TC_CHUNK_BYTES;
                        const uint32_t d_tmem = tmem_base + (uint32_t)((as * TC_MT + m) * TC_ROWS);
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            tc_mma_i8(d_tmem, tc_smem_desc(asm_ + 32 * k), tc_smem_desc(bsm + 32 * k), kIdescI8, (c | k) != 0);
                    }
                    tc_commit(bar_empty + 8 * s);  // chunk free once these MMAs have read it
                }
                tc_commit(bar_tfull + 8 * as);     // accumulators of tile t complete
            }
        }
    } else {
        // ===== epilogue: thread = query row (TMEM lane), TC_MT rows per thread =====
        const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;
        int best[TC_MT], second[TC_MT];
        unsigned int bidx[TC_MT];
#pragma unroll
        for (int m = 0; m < TC_MT; m++) {
            best[m] = (int)kSentinel - kHamBias;
            second[m] = (int)kSentinel - kHamBias;
            bidx[m] = 0;
        }
        for (unsigned int t = 0; t < n_tiles; t++) {
            const int as = t & 1;
            const unsigned int col_base = (t_beg + t) * TC_ROWS;
            const int valid = (int)min((unsigned long long)TC_ROWS, ndb - col_base);
            mbar_wait(bar_tfull + 8 * as, (t >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int m = 0; m < TC_MT; m++) {
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((as * TC_MT + m) * TC_ROWS);
                int k1 = 0x7fffffff, k2 = 0x7fffffff;
                uint32_t r0[32], r1[32];
                TC_LD32(r0, taddr);
                TC_WAIT_LD(r0);
                TC_LD32(r1, taddr + 32);
                if (valid == TC_ROWS) top2_32<false>(r0, 0, valid, k1, k2); else top2_32<true>(r0, 0, valid, k1, k2);
                TC_WAIT_LD(r1);
                TC_LD32(r0, taddr + 64);
                if (valid == TC_ROWS) top2_32<false>(r1, 32, valid, k1, k2); else top2_32<true>(r1, 32, valid, k1, k2);
                TC_WAIT_LD(r0);
                TC_LD32(r1, taddr + 96);
                if (valid == TC_ROWS) top2_32<false>(r0, 64, valid, k1, k2); else top2_32<true>(r0, 64, valid, k1, k2);
                TC_WAIT_LD(r1);
                if (valid == TC_ROWS) top2_32<false>(r1, 96, valid, k1, k2); else top2_32<true>(r1, 96, valid, k1, k2);
                // merge the tile's two smallest into the running result (feature_matching.rs:43-49; tiles ascend, so a
                // strict compare keeps the lowest index among equal distances)
                const int d1 = k1 >> 7, d2 = k2 >> 7;
                if (d1 < best[m]) {
                    second[m] = min(best[m], d2);
                    best[m] = d1;
                    bidx[m] = col_base + (unsigned int)(k1 & 127);
                } else {
                    second[m] = min(second[m], d1);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * as);
        }
#pragma unroll
        for (int m = 0; m < TC_MT; m++) {
            const unsigned long long qi = ((unsigned long long)blockIdx.x * TC_MT + m) * TC_ROWS + row;
            if (qi < nq) {
                akz_top2 o;
                o.best_idx = bidx[m] + db_index_base;
                o.best = (uint16_t)min(best[m] + kHamBias, (int)kSentinel);
                o.second = (uint16_t)min(second[m] + kHamBias, (int)kSentinel);
                out[(unsigned long long)blockIdx.y * nq + qi] = o;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace

cudaError_t init_matcher_tc_attributes() {
    return cudaFuncSetAttribute(k_match_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
}

// bytes of the expanded image of n descriptors when used as queries / as database
size_t match_tc_query_image_bytes(uint64_t nq) {
    const uint64_t tiles = (nq + TC_ROWS - 1) / TC_ROWS;
    return (size_t)((tiles + TC_MT - 1) / TC_MT * TC_MT) * TC_TILE_BYTES;
}
size_t match_tc_db_image_bytes(uint64_t ndb) { return (size_t)((ndb + TC_ROWS - 1) / TC_ROWS) * TC_TILE_BYTES; }

int match_tc_parts(uint64_t nq, uint64_t ndb) {
    const uint64_t ctas = (nq + (uint64_t)TC_ROWS * TC_MT - 1) / ((uint64_t)TC_ROWS * TC_MT);
    const uint64_t db_tiles = (ndb + TC_ROWS - 1) / TC_ROWS;
    uint64_t parts = 1;
    if (ctas < 148) parts = (148 + ctas - 1) / ctas;
    if (parts > db_tiles) parts = db_tiles;
    if (parts < 1) parts = 1;
    if (parts > 65535) parts = 65535;
    return (int)parts;
}

// d_q, d_db: 64-byte rows; q_img/db_img: scratch of match_tc_*_image_bytes; d_out: akz_top2[n_parts][nq]
int launch_match_tc(cudaStream_t s, const uint8_t* d_q, uint64_t nq, const uint8_t* d_db, uint64_t ndb, uint32_t db_index_base,
                    uint8_t* q_img, uint8_t* db_img, akz_top2* d_out, int n_parts) {
    if (nq == 0 || ndb == 0) return 0;
    const uint64_t nq_pad = match_tc_query_image_bytes(nq) / TC_TILE_BYTES * TC_ROWS;
    const uint64_t ndb_pad = match_tc_db_image_bytes(ndb) / TC_TILE_BYTES * TC_ROWS;
    k_match_expand<false><<<(unsigned int)((nq_pad * 32 + 255) / 256), 256, 0, s>>>(d_q, nq, nq_pad, q_img);
    k_match_expand<true><<<(unsigned int)((ndb_pad * 32 + 255) / 256), 256, 0, s>>>(d_db, ndb, ndb_pad, db_img);
    const unsigned int db_tiles = (unsigned int)(ndb_pad / TC_ROWS);
    const unsigned int tiles_per_part = (db_tiles + n_parts - 1) / n_parts;
    dim3 grid((unsigned int)(nq_pad / (TC_ROWS * TC_MT)), (unsigned int)n_parts);
    k_match_tc<<<grid, TC_THREADS, TC_SMEM, s>>>(q_img, nq, db_img, ndb, tiles_per_part, db_index_base, d_out);
    return 3;
}

}  // namespace akz
