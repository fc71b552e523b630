// matcher.cu -- brute-force Hamming top-2 on the integer popc pipe.
//
// Replaces descriptor_match's inner loops and hamming_distance
// (akaze/src/ops/feature_matching.rs:37-50, 113-123). Result per query: the two smallest values of the
// multiset {d_j} U {10000, 10000} and the LOWEST j attaining the minimum (the early bail-out of
// hamming_distance never changes that). Descriptors are padded to 64-byte rows
// (486 bits used); 16 x (XOR + POPC32) per pair.
//
// Layout: each thread keeps MQ query descriptors in registers (16 words each); database descriptors
// are staged through shared memory in tiles and read with broadcast 128-bit loads, so one LDS.128
// feeds 4*MQ POPCs. The database can be split into parts along blockIdx.y when there are too few
// queries to fill the GPU; parts (and multi-GPU shards) are merged by k_merge with the sequential
// scan's tie rule.
#include "common.cuh"

namespace akz {
namespace {

constexpr int MQ = 4;      // queries per thread
constexpr int MT = 128;    // threads per block
constexpr int MDB = 128;   // database descriptors per shared-memory tile (8 KB)
constexpr unsigned int kSentinel = 10000u;  // distance_threshold hard-wired by lib.rs:264

__device__ __forceinline__ unsigned int ham16(const uint4 (&q)[4], const uint4& d0, const uint4& d1, const uint4& d2,
                                              const uint4& d3) {
    unsigned int s = __popc(q[0].x ^ d0.x) + __popc(q[0].y ^ d0.y) + __popc(q[0].z ^ d0.z) + __popc(q[0].w ^ d0.w);
    s += __popc(q[1].x ^ d1.x) + __popc(q[1].y ^ d1.y) + __popc(q[1].z ^ d1.z) + __popc(q[1].w ^ d1.w);
    s += __popc(q[2].x ^ d2.x) + __popc(q[2].y ^ d2.y) + __popc(q[2].z ^ d2.z) + __popc(q[2].w ^ d2.w);
    s += __popc(q[3].x ^ d3.x) + __popc(q[3].y ^ d3.y) + __popc(q[3].z ^ d3.z) + __popc(q[3].w ^ d3.w);
    return s;
}

__global__ void __launch_bounds__(MT)
k_match_top2(const uint4* __restrict__ q, unsigned long long nq, const uint4* __restrict__ db, unsigned long long ndb,
             unsigned long long part_len, unsigned int db_index_base, akz_top2* __restrict__ out) {
    __shared__ uint4 s_db[MDB * 4];
    const unsigned long long q0 = ((unsigned long long)blockIdx.x * MT + threadIdx.x) * MQ;
    const unsigned long long j_beg = (unsigned long long)blockIdx.y * part_len;
    const unsigned long long j_end = min(ndb, j_beg + part_len);
    uint4 qr[MQ][4];
    unsigned int best[MQ], second[MQ], bidx[MQ];
#pragma unroll
    for (int m = 0; m < MQ; m++) {
        const unsigned long long qi = min(q0 + m, nq - 1);  // tail threads recompute the last query
#pragma unroll
        for (int w = 0; w < 4; w++) qr[m][w] = q[qi * 4 + w];
        best[m] = kSentinel;
        second[m] = kSentinel;
        bidx[m] = 0;
    }
    for (unsigned long long t0 = j_beg; t0 < j_end; t0 += MDB) {
        const int cnt = (int)min((unsigned long long)MDB, j_end - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 4; i += MT) s_db[i] = db[t0 * 4 + i];
        __syncthreads();
        for (int j = 0; j < cnt; j++) {
            const uint4 d0 = s_db[j * 4 + 0], d1 = s_db[j * 4 + 1], d2 = s_db[j * 4 + 2], d3 = s_db[j * 4 + 3];
            const unsigned int jj = (unsigned int)(t0 + j);
#pragma unroll
            for (int m = 0; m < MQ; m++) {
                const unsigned int d = ham16(qr[m], d0, d1, d2, d3);
                // feature_matching.rs:43-49
                if (d < best[m]) {
                    second[m] = best[m];
                    best[m] = d;
                    bidx[m] = jj;
                } else if (d < second[m]) {
                    second[m] = d;
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < MQ; m++) {
        if (q0 + m < nq) {
            akz_top2 r;
            r.best_idx = bidx[m] + db_index_base;
            r.best = (uint16_t)best[m];
            r.second = (uint16_t)second[m];
            out[(unsigned long long)blockIdx.y * nq + q0 + m] = r;
        }
    }
}

// merge parts ordered by ascending database index range; lowest index wins ties
__global__ void k_merge_top2(const akz_top2* __restrict__ parts, unsigned int n_parts, unsigned long long nq,
                             akz_top2* __restrict__ out) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    unsigned int best = kSentinel, second = kSentinel, bidx = 0;
    for (unsigned int p = 0; p < n_parts; p++) {
        const akz_top2 r = parts[(unsigned long long)p * nq + i];
        // feed the part's two smallest distances through the same update rule, best first
        unsigned int d = r.best;
        if (d < best) {
            second = best;
            best = d;
            bidx = r.best_idx;
        } else if (d < second) {
            second = d;
        }
        d = r.second;
        if (d < second) second = d;  // r.second >= r.best >= best: can never become the new best
    }
    if (n_parts > 0 && best == kSentinel) bidx = parts[i].best_idx;
    akz_top2 o;
    o.best_idx = bidx;
    o.best = (uint16_t)best;
    o.second = (uint16_t)second;
    out[i] = o;
}

}  // namespace

// d_out must hold n_parts * nq records when n_parts > 1 (see match_parts)
int match_parts(uint64_t nq, uint64_t ndb) {
    if (nq == 0 || ndb == 0) return 1;
    const uint64_t qblocks = (nq + (uint64_t)MT * MQ - 1) / ((uint64_t)MT * MQ);
    uint64_t parts = 1;
    const uint64_t want = 148 * 4;
    if (qblocks < want) parts = (want + qblocks - 1) / qblocks;
    const uint64_t max_parts = (ndb + 4 * MDB - 1) / (4 * MDB);
    if (parts > max_parts) parts = max_parts;
    if (parts < 1) parts = 1;
    if (parts > 65535) parts = 65535;
    return (int)parts;
}

int launch_match_top2(cudaStream_t s, const uint8_t* d_q, uint64_t nq, const uint8_t* d_db, uint64_t ndb,
                      uint32_t db_index_base, akz_top2* d_out, int n_parts) {
    if (nq == 0) return 0;
    const uint64_t qblocks = (nq + (uint64_t)MT * MQ - 1) / ((uint64_t)MT * MQ);
    const uint64_t part_len = ndb == 0 ? 1 : (ndb + n_parts - 1) / n_parts;
    dim3 grid((unsigned int)qblocks, (unsigned int)n_parts);
    k_match_top2<<<grid, MT, 0, s>>>((const uint4*)d_q, nq, (const uint4*)d_db, ndb, part_len, db_index_base, d_out);
    return 1;
}

// rows of `stride` bytes with `desc_len` valid ones -> rows of 64 bytes, zero padded (one thread per 4 output bytes)
__global__ void k_repack_rows(const uint8_t* __restrict__ src, size_t stride, uint32_t desc_len, uint64_t n, uint32_t* __restrict__ dst) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 16) return;
    const uint64_t row = t >> 4;
    const uint32_t col = (uint32_t)(t & 15) * 4;
    const uint8_t* r = src + row * stride;
    uint32_t v = 0;
#pragma unroll
    for (int b = 0; b < 4; b++)
        if (col + b < desc_len) v |= (uint32_t)r[col + b] << (8 * b);
    dst[t] = v;
}

int launch_repack_rows(cudaStream_t s, const uint8_t* d_src, size_t stride, uint32_t desc_len, uint64_t n, uint8_t* d_dst) {
    if (n == 0) return 0;
    const uint64_t threads = n * 16;
    k_repack_rows<<<(unsigned int)((threads + 255) / 256), 256, 0, s>>>(d_src, stride, desc_len, n, reinterpret_cast<uint32_t*>(d_dst));
    return 1;
}

int launch_merge_top2(cudaStream_t s, const akz_top2* d_parts, uint32_t n_parts, uint64_t nq, akz_top2* d_out) {
    if (nq == 0) return 0;
    k_merge_top2<<<(unsigned int)((nq + 255) / 256), 256, 0, s>>>(d_parts, n_parts, nq, d_out);
    return 1;
}

}  // namespace akz
