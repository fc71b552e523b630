// keypoints.cu -- cross-scale suppression, "sub-pixel" refinement, dominant orientation, MLDB descriptor.
//
// Replaces (reference paths relative to the akaze-rust repository):
//   find_scale_space_extrema, cache pass + upper-scale filter  akaze/src/ops/scale_space_extrema.rs:43-132
//   do_subpixel_refinement                                      akaze/src/ops/scale_space_extrema.rs:141-189
//   compute_main_orientation                                    akaze/src/ops/scale_space_extrema.rs:274-329
//   extract_descriptors .. mldb_binary_comparisons              akaze/src/ops/descriptors.rs:14-175
//
// The cache pass of the reference is a greedy, insertion-order dependent scan: for each
// candidate, in level-major raster order, the FIRST cache slot holding a keypoint of the same or the
// previous class_id within `size` pixels decides (replace-in-slot if the candidate is stronger, else
// drop). That semantics is kept exactly: one warp walks one image's ordered candidate list; the linear
// cache scan is replaced by a spatial hash (two grids, one per class parity, since only classes L and
// L-1 can match) whose cells are searched in parallel by the lanes for the LOWEST matching slot.
#include <cstdlib>

#include "common.cuh"

namespace akz {
namespace {

__device__ __forceinline__ bool image_fits_smem_pass(const PlanDev* plan, const unsigned int* lo);
__device__ __forceinline__ bool image_fits_level_pass(const PlanDev* plan, const unsigned int* lo, int n);

// ------------------------------------------------------------------------------------------------
// K5a: greedy cache pass, one warp per image, working set in global memory (fallback for images whose
// busiest level does not fit the shared-memory pools of k_dedup_smem below)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
k_dedup(const PlanDev* __restrict__ plan, const unsigned int* __restrict__ cand, const unsigned int* __restrict__ level_off,
        const float* __restrict__ ldet_plane, int batch, unsigned int cand_cap, unsigned int kp_cap,
        volatile float* c_x, volatile float* c_y, volatile float* c_resp, volatile int* c_cls, volatile int* c_next,
        volatile int* grid, unsigned int* __restrict__ n_cache, unsigned int* __restrict__ err_flags, int fast_pass,
        unsigned int* __restrict__ upper_done) {
    const int img = blockIdx.x;
    const int lane = threadIdx.x;
    const unsigned int FULL = 0xffffffffu;
    const int gw = plan->grid_w, gh = plan->grid_h, gshift = plan->grid_shift;
    const int gcells = gw * gh;
    c_x += (size_t)img * kp_cap;
    c_y += (size_t)img * kp_cap;
    c_resp += (size_t)img * kp_cap;
    c_cls += (size_t)img * kp_cap;
    c_next += (size_t)img * kp_cap;
    grid += (size_t)img * 2 * gcells;
    const unsigned int* cl = cand + (size_t)img * cand_cap;
    const unsigned int* lo = level_off + (size_t)img * (kMaxLevels + 1);
    if (err_flags[img] & kErrCandOverflow) {
        if (lane == 0) n_cache[img] = 0;
        return;
    }
    // handled by the fast pass that runs beside this kernel (1 = k_dedup_levels, 0 = k_dedup_smem)
    if (fast_pass == 1 ? image_fits_level_pass(plan, lo, plan->n_levels) : image_fits_smem_pass(plan, lo)) return;
    if (lane == 0) upper_done[img] = 0;  // k_filter_refine runs the upper-scale scan for this image
    unsigned int n = 0;  // cache length (uniform across lanes)
    bool overflow = false;

    for (int L = 0; L < plan->n_levels && !overflow; L++) {
        const LevelDev& lv = plan->lv[L];
        volatile int* g_cur = grid + (size_t)(L & 1) * gcells;   // entries of class L
        volatile int* g_prev = grid + (size_t)((L + 1) & 1) * gcells;  // entries of class L-1
        for (int i = lane; i < gcells; i += 32) g_cur[i] = -1;  // drops class L-2
        __syncwarp();
        const float ratio = lv.ratio, size = lv.kp_size, size_sq = lv.size_sq, hr = lv.half_ratio_m1;
        const float* ldet = ldet_plane + (size_t)lv.off * batch + (size_t)img * lv.w * lv.h;
        const unsigned int beg = lo[L], end = lo[L + 1];
        for (unsigned int base = beg; base < end && !overflow; base += 32) {
            // prefetch up to 32 candidates
            unsigned int my_flat = 0;
            float my_resp = 0.0f;
            if (base + lane < end) {
                my_flat = cl[base + lane];
                my_resp = fabsf(ldet[my_flat]);  // scale_space_extrema.rs:44
            }
            const int cnt = min(32u, end - base);
            for (int k = 0; k < cnt; k++) {
                const unsigned int flat = __shfl_sync(FULL, my_flat, k);
                const float resp = __shfl_sync(FULL, my_resp, k);
                const int px = (int)(flat % (unsigned int)lv.w), py = (int)(flat / (unsigned int)lv.w);
                // :62-65 compares keypoint.point * ratio (level coords scaled) with the cached full-res point
                const float qx = (float)px * ratio, qy = (float)py * ratio;
                const int cx0 = max(0, ((int)floorf(qx - size) - 1) >> gshift);
                const int cx1 = min(gw - 1, ((int)floorf(qx + size) + 1) >> gshift);
                const int cy0 = max(0, ((int)floorf(qy - size) - 1) >> gshift);
                const int cy1 = min(gh - 1, ((int)floorf(qy + size) + 1) >> gshift);
                const int nx = cx1 - cx0 + 1, ny = cy1 - cy0 + 1;
                const int ncell = nx * ny;
                const int ntot = (L > 0) ? 2 * ncell : ncell;
                unsigned int best = 0xffffffffu;
                for (int c = lane; c < ntot; c += 32) {
                    const int which = c >= ncell;
                    const int cc = which ? c - ncell : c;
                    const int cell = (cy0 + cc / nx) * gw + (cx0 + cc % nx);
                    int e = which ? g_prev[cell] : g_cur[cell];
                    while (e >= 0) {
                        const int ecls = c_cls[e];
                        if (ecls == L || ecls + 1 == L) {
                            const float dx = qx - c_x[e], dy = qy - c_y[e];
                            const float dist = dx * dx + dy * dy;
                            if (dist <= size_sq && (unsigned int)e < best) best = (unsigned int)e;
                        }
                        e = c_next[e];
                    }
                }
                best = __reduce_min_sync(FULL, best);
                bool do_write = false;
                unsigned int slot = 0;
                if (best == 0xffffffffu) {
                    slot = n;
                    if (slot >= kp_cap) {
                        overflow = true;
                        break;
                    }
                    n++;
                    do_write = true;
                } else if (resp > c_resp[best]) {  // :67
                    slot = best;
                    do_write = true;
                    if (lane == 0) {
                        // unlink `best` from the list it is on (class and position of the old occupant)
                        const int ocls = c_cls[best];
                        volatile int* g_old = grid + (size_t)(ocls & 1) * gcells;
                        const int ocx = min(gw - 1, max(0, (int)c_x[best] >> gshift));
                        const int ocy = min(gh - 1, max(0, (int)c_y[best] >> gshift));
                        const int ocell = ocy * gw + ocx;
                        int e = g_old[ocell];
                        if (e == (int)best) {
                            g_old[ocell] = c_next[best];
                        } else {
                            while (e >= 0) {
                                const int nxt = c_next[e];
                                if (nxt == (int)best) {
                                    c_next[e] = c_next[best];
                                    break;
                                }
                                e = nxt;
                            }
                        }
                    }
                }
                if (do_write && lane == 0) {
                    // :89-92 full-resolution point
                    const float fx = (float)px * ratio + hr, fy = (float)py * ratio + hr;
                    c_x[slot] = fx;
                    c_y[slot] = fy;
                    c_resp[slot] = resp;
                    c_cls[slot] = L;
                    const int ncx = min(gw - 1, max(0, (int)fx >> gshift));
                    const int ncy = min(gh - 1, max(0, (int)fy >> gshift));
                    const int ncl = ncy * gw + ncx;
                    c_next[slot] = g_cur[ncl];
                    g_cur[ncl] = (int)slot;
                }
                __syncwarp();
            }
        }
    }
    if (lane == 0) {
        n_cache[img] = n;
        if (overflow) atomicOr(&err_flags[img], (unsigned int)kErrKpOverflow);
    }
}

// ------------------------------------------------------------------------------------------------
// K5a': the same greedy pass with the whole working set in shared memory. Only keypoints of classes
// L-1 and L can match a level-L candidate, so two pools (by class parity) of at most kPool entries
// each hold everything the pass touches; a pool is flushed to the global cache arrays (indexed by
// slot) when its class falls out of reach. Hash-grid heads are u16 pool indices. An image whose
// busiest level has more than kPool candidates is left to the global-memory kernel above.
// ------------------------------------------------------------------------------------------------
constexpr int kGroups = 8;   // candidates decided per step of the cache pass (lanes per candidate = 32 / kGroups)
constexpr int kMaxGridCells = 8448;
constexpr unsigned short kNil = 0xffffu;

__device__ __forceinline__ bool image_fits_smem_pass(const PlanDev* plan, const unsigned int* lo) {
    if (plan->grid_w * plan->grid_h > kMaxGridCells) return false;
    for (int l = 0; l < plan->n_levels; l++)
        if (lo[l + 1] - lo[l] > (unsigned int)plan->pool_cap) return false;
    return true;
}

// Only the hash-grid heads live in shared memory (34 KB). The two pools are per-image scratch in global
// memory: an entry is touched once or twice per candidate and stays in L1/L2, while a 181 KB shared-memory
// working set kept every other kernel of the pipeline (which runs concurrently on the second stream) off the
// SMs that hosted a cache-pass block. Lanes of the one warp exchange pool entries through global memory
// ordered by __syncwarp().
struct DedupSmem {
    unsigned short heads[2][kMaxGridCells];
};
// per-image slab of dedup_pool_bytes(): two pools (class parity) of plan.pool_cap entries each; pool_cap follows the image
// size (PlanDev::pool_cap, <= 65534 because pool indices are u16) so that a 3840x2160 frame's busiest level fits too
struct DedupPool {
    float* px[2];
    float* py[2];
    float* presp[2];
    unsigned int* pslot[2];  // global cache slot; 0xffffffff = dead (replaced)
    unsigned short* pnext[2];
    __device__ __forceinline__ DedupPool(unsigned char* slab, int cap) {
        float* f = reinterpret_cast<float*>(slab);
        px[0] = f;
        px[1] = f + cap;
        py[0] = f + 2 * cap;
        py[1] = f + 3 * cap;
        presp[0] = f + 4 * cap;
        presp[1] = f + 5 * cap;
        pslot[0] = reinterpret_cast<unsigned int*>(f + 6 * cap);
        pslot[1] = pslot[0] + cap;
        pnext[0] = reinterpret_cast<unsigned short*>(f + 8 * cap);
        pnext[1] = pnext[0] + cap;
    }
};
constexpr size_t kPoolBytesPerEntry = 8 * 4 + 2 * 2;

__global__ void __launch_bounds__(32)
k_dedup_smem(const PlanDev* __restrict__ plan, const unsigned int* __restrict__ cand, const unsigned int* __restrict__ level_off,
             const float* __restrict__ ldet_plane, int batch, unsigned int cand_cap, unsigned int kp_cap, float* __restrict__ c_x,
             float* __restrict__ c_y, float* __restrict__ c_resp, int* __restrict__ c_cls, unsigned int* __restrict__ n_cache,
             unsigned int* __restrict__ err_flags, unsigned char* pools, unsigned int* __restrict__ upper_done) {
    __shared__ DedupSmem S;
    const int img = blockIdx.x;
    const DedupPool G(pools + (size_t)img * ((size_t)plan->pool_cap * kPoolBytesPerEntry), plan->pool_cap);
    const int lane = threadIdx.x;
    const unsigned int FULL = 0xffffffffu;
    const unsigned int* cl = cand + (size_t)img * cand_cap;
    const unsigned int* lo = level_off + (size_t)img * (kMaxLevels + 1);
    if (err_flags[img] & kErrCandOverflow) {
        if (lane == 0) n_cache[img] = 0;
        return;
    }
    if (!image_fits_smem_pass(plan, lo)) return;  // handled by k_dedup
    if (lane == 0) upper_done[img] = 0;
    const int gw = plan->grid_w, gh = plan->grid_h, gshift = plan->grid_shift;
    const int gcells = gw * gh;
    c_x += (size_t)img * kp_cap;
    c_y += (size_t)img * kp_cap;
    c_resp += (size_t)img * kp_cap;
    c_cls += (size_t)img * kp_cap;
    unsigned int n = 0;
    int cnt[2] = {0, 0};
    bool overflow = false;

    auto flush = [&](int which, int cls) {
        for (int i = lane; i < cnt[which]; i += 32) {
            const unsigned int s = G.pslot[which][i];
            if (s != 0xffffffffu) {
                c_x[s] = G.px[which][i];
                c_y[s] = G.py[which][i];
                c_resp[s] = G.presp[which][i];
                c_cls[s] = cls;
            }
        }
    };

    const int nl = plan->n_levels;
    for (int L = 0; L < nl && !overflow; L++) {
        const LevelDev& lv = plan->lv[L];
        const int cur = L & 1, prv = cur ^ 1;
        if (L >= 2) flush(cur, L - 2);  // class L-2 can no longer be matched or replaced
        cnt[cur] = 0;
        for (int i = lane; i < gcells; i += 32) S.heads[cur][i] = kNil;
        __syncwarp();
        const float ratio = lv.ratio, size = lv.kp_size, size_sq = lv.size_sq, hr = lv.half_ratio_m1;
        const float* ldet = ldet_plane + (size_t)lv.off * batch + (size_t)img * lv.w * lv.h;
        const unsigned int beg = lo[L], end = lo[L + 1];
        // software-pipelined prefetch of 32 candidates (flat index + |Ldet|)
        unsigned int nx_flat = 0;
        float nx_resp = 0.0f;
        if (beg + lane < end) {
            nx_flat = cl[beg + lane];
            nx_resp = fabsf(ldet[nx_flat]);
        }
        // kGroups candidates are decided per step, one per group of 32/kGroups lanes, against the state
        // BEFORE the step; a candidate whose decision could be changed by the write of an earlier candidate
        // of the same step (see the conflict rule below) ends the step: only the candidates before it
        // commit, it is re-decided at the head of the next step. The sequential semantics is kept exactly.
        constexpr int GL = 32 / kGroups;  // lanes per group
        const int grp = lane / GL, sub = lane % GL;
        const bool leader = sub == 0;
        for (unsigned int base = beg; base < end && !overflow; base += 32) {
            const unsigned int my_flat = nx_flat;
            const float my_resp = nx_resp;
            if (base + 32 + lane < end) {
                nx_flat = cl[base + 32 + lane];
                nx_resp = fabsf(ldet[nx_flat]);
            }
            const int n_here = min(32u, end - base);
            int k0 = 0;
            while (k0 < n_here) {
                const int k = k0 + grp;
                const bool active = k < n_here;
                const unsigned int flat = __shfl_sync(FULL, my_flat, k & 31);
                const float resp = __shfl_sync(FULL, my_resp, k & 31);
                const int px = (int)(flat % (unsigned int)lv.w), py = (int)(flat / (unsigned int)lv.w);
                const float qx = (float)px * ratio, qy = (float)py * ratio;
                unsigned int best = 0xffffffffu;  // lowest matching slot
                unsigned int best_ref = 0;        // (pool << 16) | pool index of that slot
                if (active) {
                    const int cx0 = max(0, ((int)floorf(qx - size) - 1) >> gshift);
                    const int cx1 = min(gw - 1, ((int)floorf(qx + size) + 1) >> gshift);
                    const int cy0 = max(0, ((int)floorf(qy - size) - 1) >> gshift);
                    const int cy1 = min(gh - 1, ((int)floorf(qy + size) + 1) >> gshift);
                    const int nx = cx1 - cx0 + 1;
                    const int ncell = nx * (cy1 - cy0 + 1);
                    const int ntot = (L > 0) ? 2 * ncell : ncell;
                    for (int c = sub; c < ntot; c += GL) {
                        const int which = (c >= ncell) ? prv : cur;
                        const int cc = (c >= ncell) ? c - ncell : c;
                        const int cell = (cy0 + cc / nx) * gw + (cx0 + cc % nx);
                        unsigned short e = S.heads[which][cell];
                        while (e != kNil) {
                            const float dx = qx - G.px[which][e], dy = qy - G.py[which][e];
                            const float dist = dx * dx + dy * dy;
                            const unsigned int s = G.pslot[which][e];  // 0xffffffff = replaced (dead) entry: never < best
                            if (dist <= size_sq && s < best) {
                                best = s;
                                best_ref = ((unsigned int)which << 16) | e;
                            }
                            e = G.pnext[which][e];
                        }
                    }
                }
#pragma unroll
                for (int o = 1; o < GL; o <<= 1) {
                    const unsigned int ob = __shfl_xor_sync(FULL, best, o);
                    const unsigned int orf = __shfl_xor_sync(FULL, best_ref, o);
                    if (ob < best) {
                        best = ob;
                        best_ref = orf;
                    }
                }
                // decision against the pre-step state: 0 = drop, 1 = append, 2 = replace slot `best`
                const int bw = (int)(best_ref >> 16);
                const unsigned short be = (unsigned short)(best_ref & 0xffffu);
                int act = 0;
                if (active) {
                    if (best == 0xffffffffu) act = 1;
                    else if (resp > G.presp[bw][be]) act = 2;  // scale_space_extrema.rs:67
                }
                const float fx = (float)px * ratio + hr, fy = (float)py * ratio + hr;  // :89-92
                // conflict rule: an earlier candidate a of this step that writes changes the decision of b only if
                // its new entry lies within `size` of b (b would see it), or it replaces the very slot b matched
                bool conflict = false;
#pragma unroll
                for (int a = 0; a < kGroups - 1; a++) {
                    const int a_act = __shfl_sync(FULL, act, a * GL);
                    const float a_fx = __shfl_sync(FULL, fx, a * GL), a_fy = __shfl_sync(FULL, fy, a * GL);
                    const unsigned int a_best = __shfl_sync(FULL, best, a * GL);
                    if (a < grp && a_act != 0 && active) {
                        const float dx = qx - a_fx, dy = qy - a_fy;
                        const float dist = dx * dx + dy * dy;
                        if (dist <= size_sq || (a_act == 2 && a_best == best)) conflict = true;
                    }
                }
                const unsigned int cmask = __ballot_sync(FULL, conflict && leader);
                const int n_act = min(kGroups, n_here - k0);
                const int n_commit = cmask ? min(n_act, (__ffs(cmask) - 1) / GL) : n_act;  // >= 1: group 0 never conflicts
                const bool commits = leader && grp < n_commit;
                const unsigned int amask = __ballot_sync(FULL, commits && act == 1);
                const unsigned int wmask = __ballot_sync(FULL, commits && act != 0);
                const unsigned int lt = (1u << lane) - 1u;
                if (n + __popc(amask) > kp_cap) {
                    overflow = true;
                    break;
                }
                if (commits && act != 0) {
                    const unsigned int slot = (act == 1) ? n + __popc(amask & lt) : best;
                    const int at = cnt[cur] + __popc(wmask & lt);  // <= candidates of this level <= pool_cap
                    if (act == 2) G.pslot[bw][be] = 0xffffffffu;  // the old occupant is dead; it stays on its cell list
                    G.px[cur][at] = fx;
                    G.py[cur][at] = fy;
                    G.presp[cur][at] = resp;
                    G.pslot[cur][at] = slot;
                    const int ncx = min(gw - 1, max(0, (int)fx >> gshift));
                    const int ncy = min(gh - 1, max(0, (int)fy >> gshift));
                    const int ncl = ncy * gw + ncx;
                    // writers that hash to the same cell link in lane order, the others all at once
                    const unsigned int same = __match_any_sync(wmask, ncl);
                    const int rank = __popc(same & lt), nsame = __popc(same);
                    for (int r = 0; r < nsame; r++) {
                        if (r == rank) {
                            G.pnext[cur][at] = S.heads[cur][ncl];
                            S.heads[cur][ncl] = (unsigned short)at;
                        }
                        __syncwarp(same);
                    }
                }
                n += __popc(amask);
                cnt[cur] += __popc(wmask);
                k0 += n_commit;
                __syncwarp();
            }
        }
    }
    // flush what is still resident: classes nl-2 (if any) and nl-1
    if (nl >= 2) flush((nl - 2) & 1, nl - 2);
    flush((nl - 1) & 1, nl - 1);
    if (lane == 0) {
        n_cache[img] = n;
        if (overflow) atomicOr(&err_flags[img], (unsigned int)kErrKpOverflow);
    }
}

// ------------------------------------------------------------------------------------------------
// K5a'': the same greedy pass, one WARP PER LEVEL, pipelined down the image (default).
//
// The sequential pass visits the levels one after the other, but a level-L candidate at full-resolution row Y only ever
// reads or changes cache entries of classes L and L-1 that lie within size_L of it. So level L may run concurrently with
// level L-1 as long as it stays far enough BEHIND it: when the level L-1 warp works on the candidate at row P, everything
// it will still use or change lies at rows >= P - size_{L-1}, and the level-L warp at row Y uses rows <= Y + size_L. With
//     P >= Y + size_L + size_{L-1} + 4 * ratio_{L-1} + 6
// the two never meet, hence level L sees exactly the state the sequential pass would show it: every level-(L-1) decision
// that can influence it has been taken, none of its own writes can influence a pending level-(L-1) decision. Each warp
// publishes the row of its first undecided candidate (s_progress), its follower spins on it. Within a warp the pass is
// k_dedup_smem's (KG speculative candidates per step, exact conflict rule).
//
// Slots: the sequential pass numbers cache slots in append order = (level, running index inside the level). That pair is
// the KEY of an entry (a replacement keeps the key of the slot it takes over), "lowest slot first" is "lowest key first",
// and the final slot numbers are the keys compacted by a prefix sum over the levels' append counts once all warps are
// through.
//
// Pools: entries live in per-level pools (class L = written by warp L: appended or replaced into). A warp walks its
// candidates in raster order, so its pool is SORTED BY ROW as it grows; a table in shared memory holds, for every row of the
// level, the index of the first entry at or below it. "All entries within size of a point" is then one contiguous index
// range of the pool (the rows the circle touches, all columns): 16-byte records, independent loads, no lists to chase.
// The first version hashed the entries into cell lists; chasing them cost two dependent L2 round trips per hop (stores do
// not allocate in L1, so freshly written entries always miss) and left a step of eight candidates at 4.4 k cycles:
// 0.94 ms for a 1080p frame, 4.2 ms for a 3840x2160 one. Row ranges: see profiles/README.md.
// ------------------------------------------------------------------------------------------------
constexpr unsigned int kDeadKey = 0xffffffffu;
constexpr int kKeyShift = 20;               // key = (level of the append << 20) | index of the append inside its level
constexpr unsigned int kMaxLevelCands = (1u << kKeyShift) - 1u;  // the index field of a key
typedef unsigned int row_t;  // row-table entry (u16 would halve the tables and send levels beyond 65 534 candidates to k_dedup)

// (the levels [0, n): a first part of the pass only knows the offsets of its own levels)
__device__ __forceinline__ bool image_fits_level_pass(const PlanDev* plan, const unsigned int* lo, int n) {
    for (int l = 0; l < n; l++)
        if (lo[l + 1] - lo[l] > kMaxLevelCands) return false;
    return true;
}

// one pool entry: the point as :89-92 stores it (x, y), |response|, key -- 16 bytes, one load.
// Entry e of level l sits at index lo[l] + e of the image's slab (cand_cap entries).
constexpr size_t kLevelPoolBytesPerCand = sizeof(uint4);

template <int KG, int MAXT>  // candidates decided per step (32 / KG lanes each); threads per block (32 per level)
// (86 registers at 512 threads: one block per SM. Capping them at 64 for two blocks per SM was measured: a 256-image sub-batch's
// pass 1.6 -> 1.2 ms, hidden either way, but one image 0.57 -> 0.64 ms and one 3840x2160 image 2.6 -> 3.4 ms.)
__global__ void __launch_bounds__(MAXT)
k_dedup_levels(const PlanDev* __restrict__ plan, const unsigned int* __restrict__ cand, const unsigned int* __restrict__ level_off,
               const float* __restrict__ ldet_plane, int batch, unsigned int cand_cap, unsigned int kp_cap, float* __restrict__ c_x,
               float* __restrict__ c_y, float* __restrict__ c_resp, int* __restrict__ c_cls, unsigned int* __restrict__ n_cache,
               unsigned int* __restrict__ err_flags, unsigned char* pools, unsigned int* __restrict__ keep_flag,
               unsigned int* __restrict__ upper_done, size_t slab_bytes, int l0, int l1) {
    // This launch walks the levels [l0, l1). One launch for all of them is the plain pass. A sub-batch with nothing to hide
    // behind (a single image, the last of a call) runs it in two: the first octave as soon as its detectors are through, next
    // to the stencil kernels of the other octaves, the rest at the end. The first part leaves its row tables and counts in
    // the image's slab, the part that ends at the last level reloads them and assigns the final slots for every level.
    extern __shared__ uint4 s_dyn[];  // per level KG step records (float4), then the row tables (u32, level l at ltab_off[l])
    __shared__ volatile int s_progress[kMaxLevels];
    __shared__ unsigned int s_appends[kMaxLevels], s_entries[kMaxLevels], s_base[kMaxLevels + 1];
    __shared__ int s_ok;
    const unsigned int FULL = 0xffffffffu;
    const int img = blockIdx.x;
    const int nl = plan->n_levels;
    const bool final_part = l1 == nl;  // (then the block has a warp for every level, else one for each of [l0, l1))
    const int lane = threadIdx.x & 31, L = (threadIdx.x >> 5) + (final_part ? 0 : l0);  // one warp per level
    float4* s_step = reinterpret_cast<float4*>(s_dyn) + (threadIdx.x >> 5) * KG;
    row_t* s_rows = reinterpret_cast<row_t*>(s_dyn + nl * KG);
    const unsigned int* cl = cand + (size_t)img * cand_cap;
    const unsigned int* lo = level_off + (size_t)img * (kMaxLevels + 1);
    unsigned char* slab = pools + (size_t)img * slab_bytes;  // pool entries, then the row tables, then (appends, entries) per level
    float4* pool = reinterpret_cast<float4*>(slab);
    row_t* g_rows = reinterpret_cast<row_t*>(slab + (size_t)cand_cap * kLevelPoolBytesPerCand);
    unsigned int* g_meta = reinterpret_cast<unsigned int*>(g_rows + plan->ltab_off[nl]);
    if (threadIdx.x == 0) s_ok = !(err_flags[img] & kErrCandOverflow) && image_fits_level_pass(plan, lo, l1);
    if (threadIdx.x < kMaxLevels) {
        s_progress[threadIdx.x] = threadIdx.x < l0 ? 0x7fffffff : -1;  // the levels of an earlier part are finished
        s_appends[threadIdx.x] = threadIdx.x < l0 ? g_meta[2 * threadIdx.x] : 0u;
    }
    for (int i = threadIdx.x; i < plan->ltab_off[l0]; i += blockDim.x) s_rows[i] = g_rows[i];
    __syncthreads();
    if (!s_ok) {  // candidate overflow: nothing to do; a level with more candidates than a key's index field holds: k_dedup takes the image
        if (threadIdx.x == 0 && final_part && (err_flags[img] & kErrCandOverflow)) {
            n_cache[img] = 0;
            upper_done[img] = 0;
        }
        return;
    }
    c_x += (size_t)img * kp_cap;
    c_y += (size_t)img * kp_cap;
    c_resp += (size_t)img * kp_cap;
    c_cls += (size_t)img * kp_cap;

    const LevelDev& lv = plan->lv[L];
    const int Lp = L > 0 ? L - 1 : 0;
    const LevelDev& pv = plan->lv[Lp];
    row_t* t_cur = s_rows + plan->ltab_off[L];        // t_cur[r] = entries of this pool above row r
    const row_t* t_prv = s_rows + plan->ltab_off[Lp];
    const unsigned int beg = lo[L], end = lo[L + 1];
    const unsigned int pbeg = L > 0 ? lo[L - 1] : 0u;  // pool base of level L-1
    const float ratio = lv.ratio, size = lv.kp_size, size_sq = lv.size_sq, hr = lv.half_ratio_m1;
    const float inv_ratio = 1.0f / ratio, p_ratio = pv.ratio, p_hr = pv.half_ratio_m1, p_inv = 1.0f / pv.ratio;
    const int p_rows = pv.h;
    // an entry whose row lies more than size above or below a point cannot be within size of it; rows (y * ratio + hr) and
    // points are exact in f32, the 0.01 px cover the rounding of the squared distance and of size * size
    const float reach = size + 0.01f;
    const int margin = L > 0 ? (int)ceilf(size + pv.kp_size) + 4 * (int)p_ratio + 6 : 0;
    const float* ldet = ldet_plane + (size_t)lv.off * batch + (size_t)img * lv.w * lv.h;
    const unsigned int lw = (unsigned int)lv.w;
    const bool walks = L >= l0;  // (a final part: the warps of the levels before l0 only take part in the slot assignment)
    unsigned int n_app = walks ? 0u : g_meta[2 * L];      // appends of this level (uniform across the warp)
    unsigned int cnt = walks ? 0u : g_meta[2 * L + 1];    // pool entries of this level (appends + replacements)
    int filled = 0;          // t_cur[0 .. filled] are final
    int seen = -1;           // the last progress of level L-1 this warp has read
    if (lane == 0 && walks) t_cur[0] = 0;

    // the 32 candidates of a batch, one per lane: position in the level, response
    int nx_px = 0, nx_py = 0;
    float nx_resp = 0.0f;
    if (walks && beg + lane < end) {
        const unsigned int f = cl[beg + lane];
        nx_py = (int)(f / lw);
        nx_px = (int)(f - (unsigned int)nx_py * lw);
        nx_resp = fabsf(ldet[f]);  // scale_space_extrema.rs:44
    }
    constexpr int GL = 32 / KG;  // lanes per candidate
    const int grp = lane / GL, sub = lane % GL;
    const bool leader = sub == 0;
    const unsigned int lt = (1u << lane) - 1u;
#ifdef AKZ_DEDUP_STATS
    long long t_wait = 0, t_search = 0, t_conf = 0, t_commit = 0, t_all0 = clock64();
    unsigned int n_steps = 0, n_scanned = 0;
#define DSTAT(x) x
#else
#define DSTAT(x)
#endif
    for (unsigned int base = walks ? beg : end; base < end; base += 32) {
        const int my_px = nx_px, my_py = nx_py;
        const float my_resp = nx_resp;
        if (base + 32 + lane < end) {
            const unsigned int f = cl[base + 32 + lane];
            nx_py = (int)(f / lw);
            nx_px = (int)(f - (unsigned int)nx_py * lw);
            nx_resp = fabsf(ldet[f]);
        }
        const int n_here = min(32u, end - base);
        int k0 = 0;
        while (k0 < n_here) {
            DSTAT(long long t0 = clock64(); n_steps++;)
            // Rows up to the first undecided candidate's are complete: close their table entries (published below, once the
            // search has given the last steps' pool writes time to land); wait until level L-1 is `margin` rows past this
            // step's last candidate.
            const int r0 = __shfl_sync(FULL, my_py, k0);
#pragma unroll 1
            for (int r = filled + 1 + lane; r <= r0; r += 32) t_cur[r] = cnt;
            filled = r0;
            if (L > 0) {
                const int rl = __shfl_sync(FULL, my_py, min(k0 + KG - 1, n_here - 1));
                const int need = (int)((float)rl * ratio) + margin;
                if (seen < need) {  // (a value seen earlier was followed by a fence then: everything up to it is visible)
                    while ((seen = s_progress[L - 1]) < need) __nanosleep(20);
                    __threadfence_block();
                }
            }
            __syncwarp();
            DSTAT(long long t1 = clock64(); t_wait += t1 - t0;)
            const int k = k0 + grp;
            const bool active = k < n_here;
            const int px = __shfl_sync(FULL, my_px, k & 31), py = __shfl_sync(FULL, my_py, k & 31);
            const float resp = __shfl_sync(FULL, my_resp, k & 31);
            const float qx = (float)px * ratio, qy = (float)py * ratio;  // :62-65 compares the level point * ratio
            unsigned int best = kDeadKey;  // lowest matching key (= lowest slot)
            unsigned int best_at = 0;      // its index in the slab
            float best_resp = 0.0f;
            if (active) {
                // this level's entries from the first row the circle can touch (not past the rows closed above) to the newest,
                // then the rows of level L-1 it can touch; one index space, eight independent loads in flight per lane
                const int r_lo = max(0, min(r0, (int)ceilf((qy - reach - hr) * inv_ratio)));
                const unsigned int i0 = beg + t_cur[r_lo];
                const unsigned int n_cur = beg + cnt - i0;
                unsigned int j0 = 0, n_tot = n_cur;
                if (L > 0) {
                    const int p_lo = max(0, (int)ceilf((qy - reach - p_hr) * p_inv));
                    const int p_hi = min(p_rows - 1, (int)floorf((qy + reach - p_hr) * p_inv));
                    if (p_lo <= p_hi) {
                        j0 = pbeg + t_prv[p_lo];
                        n_tot += pbeg + t_prv[p_hi + 1] - j0;
                    }
                }
                const unsigned int j_off = j0 - n_cur;  // index t >= n_cur of the joint space is entry j0 + (t - n_cur)
                constexpr int UN = 8;
#pragma unroll 1
                for (unsigned int t = sub; t < n_tot; t += UN * GL) {
                    float4 rec[UN];
                    unsigned int at[UN];
#pragma unroll
                    for (int u = 0; u < UN; u++) {
                        const unsigned int idx = min(t + u * GL, n_tot - 1u);  // past the end: the last entry once more
                        at[u] = idx + (idx < n_cur ? i0 : j_off);
                        rec[u] = pool[at[u]];
                    }
#pragma unroll
                    for (int u = 0; u < UN; u++) {
                        const float dx = qx - rec[u].x, dy = qy - rec[u].y;
                        const float dist = dx * dx + dy * dy;
                        const unsigned int key = __float_as_uint(rec[u].w);
                        DSTAT(n_scanned++;)
                        if (dist <= size_sq && key < best) {  // kDeadKey = replaced entry: never < best
                            best = key;
                            best_at = at[u];
                            best_resp = rec[u].z;
                        }
                    }
                }
            }
#pragma unroll
            for (int o = 1; o < GL; o <<= 1) {
                const unsigned int ob = __shfl_xor_sync(FULL, best, o);
                const unsigned int oa = __shfl_xor_sync(FULL, best_at, o);
                const float orr = __shfl_xor_sync(FULL, best_resp, o);
                if (ob < best) {
                    best = ob;
                    best_at = oa;
                    best_resp = orr;
                }
            }
            int act = 0;  // 0 = drop, 1 = append, 2 = replace the entry holding key `best`
            if (active) {
                if (best == kDeadKey) act = 1;
                else if (resp > best_resp) act = 2;  // :67
            }
            const float fx = (float)px * ratio + hr, fy = (float)py * ratio + hr;  // :89-92
            DSTAT(__syncwarp(); long long t2 = clock64(); t_search += t2 - t1;)
            // conflict rule of k_dedup_smem: an earlier candidate a of this step that writes changes b's decision only if
            // its new entry lies within `size` of b, or it replaces the very slot b matched
            if (leader) s_step[grp] = make_float4(fx, fy, __int_as_float(act), __uint_as_float(best));
            __threadfence_block();  // the table rows closed above and every earlier step's pool writes, then the row
            if (lane == 0) s_progress[L] = (int)((float)r0 * ratio);
            __syncwarp();
            bool conflict = false;
            if (active) {
                for (int a = sub; a < grp; a += GL) {
                    const float4 w = s_step[a];
                    const int a_act = __float_as_int(w.z);
                    const float dx = qx - w.x, dy = qy - w.y;
                    const float dist = dx * dx + dy * dy;
                    if (a_act != 0 && (dist <= size_sq || (a_act == 2 && __float_as_uint(w.w) == best))) conflict = true;
                }
            }
            const unsigned int cmask = __ballot_sync(FULL, conflict);
            DSTAT(long long t3 = clock64(); t_conf += t3 - t2;)
            const int n_act = min(KG, n_here - k0);
            const int n_commit = cmask ? min(n_act, (__ffs(cmask) - 1) / GL) : n_act;  // >= 1: group 0 never conflicts
            const bool commits = leader && grp < n_commit;
            const unsigned int amask = __ballot_sync(FULL, commits && act == 1);
            const unsigned int wmask = __ballot_sync(FULL, commits && act != 0);
            const unsigned int e = cnt + __popc(wmask & lt);  // the pool index an entry of this candidate gets
            const int py_before = __shfl_sync(FULL, my_py, (k + 31) & 31);  // row of the candidate before this one
            if (commits) {
                if (grp > 0) {
#pragma unroll 1
                    for (int r = py_before + 1; r <= py; r++) t_cur[r] = e;  // a new row starts here
                }
                if (act != 0) {
                    const unsigned int key = (act == 1) ? (((unsigned int)L << kKeyShift) | (n_app + __popc(amask & lt))) : best;
                    if (act == 2) reinterpret_cast<unsigned int*>(pool + best_at)[3] = kDeadKey;  // the old occupant is dead
                    pool[beg + e] = make_float4(fx, fy, resp, __uint_as_float(key));
                }
            }
            filled = __shfl_sync(FULL, my_py, (k0 + n_commit - 1) & 31);
            n_app += __popc(amask);
            cnt += __popc(wmask);
            k0 += n_commit;
            __syncwarp();
            DSTAT(t_commit += clock64() - t3;)
        }
    }
    if (walks) {
#pragma unroll 1
        for (int r = filled + 1 + lane; r <= lv.h; r += 32) t_cur[r] = cnt;
    }
#ifdef AKZ_DEDUP_STATS
    {
        const unsigned int scanned = __reduce_add_sync(FULL, n_scanned);
        if (img == 0 && lane == 0)
            printf("L%2d cand %6u steps %5u appends %5u entries %5u scanned %8u | kcycles total %7lld wait %7lld search %7lld conflict %6lld commit %6lld\n", L,
                   end - beg, n_steps, n_app, cnt, scanned, (clock64() - t_all0) / 1000, t_wait / 1000, t_search / 1000, t_conf / 1000, t_commit / 1000);
    }
#endif
    __syncwarp();
    __threadfence_block();
    if (lane == 0) {
        s_appends[L] = n_app;
        s_entries[L] = cnt;
        s_progress[L] = 0x7fffffff;  // level finished
    }
    if (!final_part) {  // leave the table and the counts for the part that follows
        for (int r = lane; r <= lv.h; r += 32) g_rows[plan->ltab_off[L] + r] = t_cur[r];
        if (lane == 0) {
            g_meta[2 * L] = n_app;
            g_meta[2 * L + 1] = cnt;
        }
        return;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int acc = 0;
        for (int l = 0; l < nl; l++) {
            s_base[l] = acc;
            acc += s_appends[l];
        }
        s_base[nl] = acc;
        if (acc > kp_cap) {
            atomicOr(&err_flags[img], (unsigned int)kErrKpOverflow);
            n_cache[img] = 0;
        } else {
            n_cache[img] = acc;
        }
        upper_done[img] = 1;  // the upper-scale filter below replaces k_filter_refine's scan for this image
    }
    __syncthreads();
    if (s_base[nl] > kp_cap) return;
    // Every live entry goes to its final slot: base of the level that appended it + its index there. The upper-scale filter
    // (scale_space_extrema.rs:111-129) rides along: entry i is dropped if a class i+1 entry in a slot >= i lies within size_i
    // of it -- the class L+1 entries are exactly the live entries of pool L+1, and its row table is still in shared memory,
    // so the scan over the whole class (quadratic in the keypoint count; 13 x the time for 4.8 x the keypoints at 3840x2160)
    // becomes a look at the rows the circle touches.
    // All threads of the block share the work level by level (the busiest level has ~10 x the entries of the quietest).
    keep_flag += (size_t)img * kp_cap;
    constexpr unsigned int kIdxMask = (1u << kKeyShift) - 1u;
    for (int l = 0; l < nl; l++) {
        const LevelDev& wl = plan->lv[l];
        const bool has_up = l + 1 < nl;
        const int lu = has_up ? l + 1 : l;
        const LevelDev& uv = plan->lv[lu];
        const row_t* t_up = s_rows + plan->ltab_off[lu];
        const unsigned int lbeg = lo[l], ubeg = lo[lu], n_l = s_entries[l];
        const float l_size_sq = wl.size_sq, l_reach = wl.kp_size + 0.01f;
        const float u_hr = uv.half_ratio_m1, u_inv = 1.0f / uv.ratio;
        for (unsigned int e = threadIdx.x; e < n_l; e += blockDim.x) {
            const float4 rec = pool[lbeg + e];
            const unsigned int key = __float_as_uint(rec.w);
            if (key == kDeadKey) continue;
            const unsigned int slot = s_base[key >> kKeyShift] + (key & kIdxMask);
            const float xi = rec.x, yi = rec.y;
            bool repeated = false;
            if (has_up) {
                const int u_lo = max(0, (int)ceilf((yi - l_reach - u_hr) * u_inv));
                const int u_hi = min(uv.h - 1, (int)floorf((yi + l_reach - u_hr) * u_inv));
                const unsigned int j0 = u_lo <= u_hi ? ubeg + t_up[u_lo] : 0u;
                const unsigned int j1 = u_lo <= u_hi ? ubeg + t_up[u_hi + 1] : 0u;
                for (unsigned int j = j0; j < j1 && !repeated; j += 4) {
                    float4 up[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) up[u] = pool[min(j + u, j1 - 1u)];  // (past the end: the last record once more)
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const unsigned int ukey = __float_as_uint(up[u].w);
                        const float dx = xi - up[u].x, dy = yi - up[u].y;
                        const float dist = dx * dx + dy * dy;
                        // :115 scans the slots j >= i
                        if (ukey != kDeadKey && dist <= l_size_sq && s_base[ukey >> kKeyShift] + (ukey & kIdxMask) >= slot) repeated = true;
                    }
                }
            }
            c_x[slot] = xi;
            c_y[slot] = yi;
            c_resp[slot] = rec.z;
            c_cls[slot] = l;
            keep_flag[slot] = repeated ? 0u : 1u;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K5b: upper-scale filter (:111-129) + pseudo sub-pixel offset (:141-178). One warp per cache slot.
// keep_flag[i] = 1 iff the slot survives; refined point written back into c_x/c_y.
// ------------------------------------------------------------------------------------------------
// slot range [lo, hi] occupied by each class_id (slots are handed out in level order, so the keypoints of
// class c+1 sit in a narrow slot range; only replaced older slots fall outside the bulk)
__global__ void __launch_bounds__(256)
k_class_ranges(const int* __restrict__ c_cls, const unsigned int* __restrict__ n_cache, unsigned int kp_cap,
               unsigned int* __restrict__ cls_range) {
    __shared__ unsigned int lo[kMaxLevels], hi[kMaxLevels];
    const int img = blockIdx.x;
    const unsigned int n = n_cache[img];
    if (threadIdx.x < kMaxLevels) {
        lo[threadIdx.x] = 0xffffffffu;
        hi[threadIdx.x] = 0u;
    }
    __syncthreads();
    for (unsigned int i = threadIdx.x; i < n; i += blockDim.x) {
        const int c = c_cls[(size_t)img * kp_cap + i];
        atomicMin(&lo[c], i);
        atomicMax(&hi[c], i);
    }
    __syncthreads();
    if (threadIdx.x < kMaxLevels) {
        cls_range[((size_t)img * kMaxLevels + threadIdx.x) * 2 + 0] = lo[threadIdx.x];
        cls_range[((size_t)img * kMaxLevels + threadIdx.x) * 2 + 1] = hi[threadIdx.x];
    }
}

__global__ void k_filter_refine(const PlanDev* __restrict__ plan, const float* __restrict__ ldet_plane, int batch,
                                unsigned int kp_cap, const float* __restrict__ c_x, const float* __restrict__ c_y,
                                const int* __restrict__ c_cls, const unsigned int* __restrict__ n_cache,
                                const unsigned int* __restrict__ cls_range, float* __restrict__ r_x, float* __restrict__ r_y,
                                unsigned int* __restrict__ keep_flag, const unsigned int* __restrict__ upper_done) {
    const int img = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const unsigned int n = n_cache[img];
    const size_t o = (size_t)img * kp_cap;
    const bool scanned = upper_done[img] != 0;  // k_dedup_levels already ran the upper-scale filter: keep_flag[i] holds its verdict
    const int warps = (blockDim.x >> 5) * gridDim.x;
    for (unsigned int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) {
        const int cls = c_cls[o + i];
        const LevelDev& lv = plan->lv[cls];
        const float xi = c_x[o + i], yi = c_y[o + i];
        const float size_sq = lv.size_sq;
        bool rep = scanned && keep_flag[o + i] == 0u;
        // :115 scans slots j >= i for class_id + 1; those all lie inside that class's slot range
        unsigned int jb = i, je = 0;
        if (!scanned && cls + 1 < plan->n_levels) {
            const unsigned int rl = cls_range[((size_t)img * kMaxLevels + cls + 1) * 2 + 0];
            const unsigned int rh = cls_range[((size_t)img * kMaxLevels + cls + 1) * 2 + 1];
            if (rl != 0xffffffffu) {
                jb = max(i, rl);
                je = min(n, rh + 1);
            }
        }
        for (unsigned int j0 = jb; j0 < je && !rep; j0 += 32) {
            const unsigned int j = j0 + lane;
            bool hit = false;
            if (j < je && c_cls[o + j] == cls + 1) {
                const float dx = xi - c_x[o + j], dy = yi - c_y[o + j];
                const float dist = dx * dx + dy * dy;
                hit = dist <= size_sq;
            }
            rep = __any_sync(0xffffffffu, hit);
        }
        if (lane == 0) {
            bool keep = !rep;
            const float ratio = lv.ratio;
            // :147-149 (the division is exact: ratio is a power of two)
            const int x = (int)roundf(xi / ratio), y = (int)roundf(yi / ratio);
            const float* D = ldet_plane + (size_t)lv.off * batch + (size_t)img * lv.w * lv.h;
            const float x_p = D[(size_t)y * lv.w + x + 1], x_m = D[(size_t)y * lv.w + x - 1];
            const float y_p = D[(size_t)(y + 1) * lv.w + x], y_m = D[(size_t)(y - 1) * lv.w + x];
            const float d_x = 0.5f * (x_p - x_m), d_y = 0.5f * (y_p - y_m);
            const float b0 = -d_x, b1 = -d_y;  // lu.solve's result is dropped (:168-169)
            if (!(fabsf(b0) <= 1.0f && fabsf(b1) <= 1.0f)) keep = false;
            float nx = (float)x + b0, ny = (float)y + b1;
            nx = nx * ratio + lv.half_ratio_m1;
            ny = ny * ratio + lv.half_ratio_m1;
            r_x[o + i] = nx;
            r_y[o + i] = ny;
            keep_flag[o + i] = keep ? 1u : 0u;
        }
    }
}

// one block per image: exclusive scan of keep_flag -> output slot, survivors counted
__global__ void __launch_bounds__(1024)
k_keep_scan(unsigned int* __restrict__ keep_flag, const unsigned int* __restrict__ n_cache, unsigned int* __restrict__ n_kp,
            unsigned int kp_cap) {
    __shared__ unsigned int warp_sums[32];
    __shared__ unsigned int carry;
    const int img = blockIdx.x;
    const unsigned int n = n_cache[img];
    unsigned int* kf = keep_flag + (size_t)img * kp_cap;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (unsigned int base = 0; base < n; base += 1024) {
        const unsigned int i = base + tid;
        const unsigned int v = (i < n) ? kf[i] : 0;
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
            warp_sums[lane] = ws;
        }
        __syncthreads();
        const unsigned int excl = carry + (wid ? warp_sums[wid - 1] : 0) + incl - v;
        // encode: bit 31 = keep, low bits = output position
        if (i < n) kf[i] = (v ? 0x80000000u : 0u) | excl;
        __syncthreads();
        if (tid == 1023) carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) n_kp[img] = carry;
}

// GAUSS25 of the reference (scale_space_extrema.rs:207-271; numeric table = data)
__constant__ float c_gauss25[7][7] = {
    {0.02546481f, 0.02350698f, 0.01849125f, 0.01239505f, 0.00708017f, 0.00344629f, 0.00142946f},
    {0.02350698f, 0.02169968f, 0.01706957f, 0.01144208f, 0.00653582f, 0.00318132f, 0.00131956f},
    {0.01849125f, 0.01706957f, 0.01342740f, 0.00900066f, 0.00514126f, 0.00250252f, 0.00103800f},
    {0.01239505f, 0.01144208f, 0.00900066f, 0.00603332f, 0.00344629f, 0.00167749f, 0.00069579f},
    {0.00708017f, 0.00653582f, 0.00514126f, 0.00344629f, 0.00196855f, 0.00095820f, 0.00039744f},
    {0.00344629f, 0.00318132f, 0.00250252f, 0.00167749f, 0.00095820f, 0.00046640f, 0.00019346f},
    {0.00142946f, 0.00131956f, 0.00103800f, 0.00069579f, 0.00039744f, 0.00019346f, 0.00008024f},
};

// ------------------------------------------------------------------------------------------------
// K5c: dominant orientation, one thread per surviving keypoint (:274-329).
// angs[k] = atan2(res_y,res_y) is pi/4 for res_y>0 and <= 0 otherwise, so a sample is inside a window
// iff res_y>0 and the window contains pi/4; which of the 42 windows do is decided once on the host
// with the reference's own f32 arithmetic (plan.orient_window_mask). sum_x/sum_y are never reset
// between windows (:301-302), so the samples are re-accumulated, in sample order, once per such window.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_orientation(const PlanDev* __restrict__ plan, const float* __restrict__ lx_plane, const float* __restrict__ ly_plane,
              int batch, unsigned int kp_cap, const float* __restrict__ r_x, const float* __restrict__ r_y,
              const float* __restrict__ c_resp, const int* __restrict__ c_cls, const unsigned int* __restrict__ keep_flag,
              const unsigned int* __restrict__ n_cache, akz_keypoint* __restrict__ kps, unsigned int* __restrict__ err_flags) {
    const int img = blockIdx.y;
    const unsigned int n = n_cache[img];
    const size_t o = (size_t)img * kp_cap;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned int kf = keep_flag[o + i];
        if (!(kf & 0x80000000u)) continue;
        const unsigned int pos = kf & 0x7fffffffu;
        const int cls = c_cls[o + i];
        const LevelDev& lv = plan->lv[cls];
        const float ratio = lv.ratio, s = lv.s_smp;
        const float ptx = r_x[o + i], pty = r_y[o + i];
        const float xf = ptx / ratio, yf = pty / ratio;
        const float* Lx = lx_plane + (size_t)lv.off * batch + (size_t)img * lv.w * lv.h;
        const float* Ly = ly_plane + (size_t)lv.off * batch + (size_t)img * lv.w * lv.h;
        // The 109 samples sit on an 11 x 11 lattice (|a|, |b| <= 5: a*a + b*b < 36 excludes +-6), so only 11 column and 11
        // row coordinates are rounded and range-checked (a column/row is used by at least the sample on the other axis'
        // centre line, so "any of the 11 out of range" is exactly "any sample out of range").
        int ixs[11], rows[11];
        bool oob = false;
#pragma unroll
        for (int t = 0; t < 11; t++) {
            int ix = (int)roundf(xf + (float)(t - 5) * s);
            int iy = (int)roundf(yf + (float)(t - 5) * s);
            if (ix < 0 || iy < 0 || ix >= lv.w || iy >= lv.h) {
                oob = true;
                ix = min(max(ix, 0), lv.w - 1);
                iy = min(max(iy, 0), lv.h - 1);
            }
            ixs[t] = ix;
            rows[t] = iy * lv.w;
        }
        // A sample takes part iff res_y > 0; the others are replaced by +0.0, which leaves a running f32 sum unchanged
        // (the sums start at +0.0 and can never become -0.0), so the accumulation below needs no per-sample test.
        float rx[109], ry[109];
        {
            int idx = 0;
#pragma unroll
            for (int a = -5; a <= 5; a++)
#pragma unroll
                for (int b = -5; b <= 5; b++)
                    if (a * a + b * b < 36) {
                        const int at = rows[b + 5] + ixs[a + 5];
                        const float gw = c_gauss25[a < 0 ? -a : a][b < 0 ? -b : b];
                        const float vx = gw * Lx[at], vy = gw * Ly[at];
                        const bool in = vy > 0.0f;
                        rx[idx] = in ? vx : 0.0f;
                        ry[idx] = in ? vy : 0.0f;
                        idx++;
                    }
        }
        float sum_x = 0.0f, sum_y = 0.0f, maxv = 0.0f, best_x = 0.0f, best_y = 0.0f;
        bool found = false;
        // windows that do not contain pi/4 leave the sums, hence val, unchanged: only the others can raise the maximum
        const int n_active = __popcll(plan->orient_window_mask & ((plan->n_orient_windows >= 64) ? ~0ull : ((1ull << plan->n_orient_windows) - 1ull)));
        for (int w = 0; w < n_active; w++) {
#pragma unroll
            for (int k = 0; k < 109; k++) {
                sum_x = sum_x + rx[k];
                sum_y = sum_y + ry[k];
            }
            const float val = sum_x * sum_x + sum_y * sum_y;
            if (val > maxv) {
                maxv = val;
                best_x = sum_x;
                best_y = sum_y;
                found = true;
            }
        }
        // f32::atan2 -> libm atan2f; evaluated in f64 and rounded once (correctly rounded except for vanishingly rare
        // double-rounding cases), once, for the window that won
        const float angle = found ? (float)atan2((double)best_y, (double)best_x) : 0.0f;
        akz_keypoint kp;
        kp.x = ptx;
        kp.y = pty;
        kp.response = c_resp[o + i];
        kp.size = lv.kp_size;
        kp.octave = (uint32_t)lv.octave;
        kp.class_id = (uint32_t)cls;
        kp.angle = angle;
        kps[o + pos] = kp;
        if (oob) atomicOr(&err_flags[img], (unsigned int)kErrBounds);
    }
}

// ------------------------------------------------------------------------------------------------
// K6: MLDB descriptor, one warp per keypoint (descriptors.rs:37-175).
//
// The three grids of the reference (2x2 cells of p x p samples, 3x3 of ceil(2p/3)^2, 4x4 of ceil(p/2)^2)
// all sample the SAME lattice: sample (k, l) sits at the rotated offset ((l+0.5), (k+0.5)) * scale whatever
// the grid, with k, l in [-p, -p + M), M = max over grids of cells*step (21 for p = 10). So
//   1. the warp evaluates the M x M lattice once (rotated position, rounded gather of Lt, Lx, Ly, rotated
//      derivative) into shared memory -- 441 gathers instead of the reference's 1241. Adjacent lanes take
//      adjacent lattice points along the lattice axis that is closer to the image x axis (k if |cos| >=
//      |sin|, else l), so that one warp-wide gather touches few cache lines;
//   2. one lane per (cell, channel) adds that cell's samples SEQUENTIALLY from shared memory (k outer, l
//      inner, f32: exactly the reference's summation order) and divides by the sample count;
//   3. the comparisons are bit-packed with warp ballots through a (value i, value j) table per bit
//      (LSB-first: grid level, then channel, then pairs i<j, descriptors.rs:161-173).
// ------------------------------------------------------------------------------------------------
constexpr int kDescWarps = 8;
constexpr int kDescM = 21;          // lattice edge for pattern 10 (3 cells x 7)
constexpr int kDescCS = 463;        // channel stride in shared memory (>= 21*21; chosen for few bank conflicts in phase 2)

// sums of one grid level: NC x NC cells of STEP x STEP lattice points, lane t owns (cell, channel) = (t / nch, t % nch)
template <int STEP, int NC>
__device__ __forceinline__ void desc_cell_sums(const float* smp, float* val, int M, int nch, int lane) {
    for (int t = lane; t < NC * NC * nch; t += 32) {
        const int c = t / nch, ch = t - c * nch;
        const int ci = c / NC, cj = c - ci * NC;
        const float* v = smp + ch * kDescCS + (ci * STEP) * M + cj * STEP;
        float acc = 0.0f;
#pragma unroll 2
        for (int kk = 0; kk < STEP; kk++) {
#pragma unroll
            for (int ll = 0; ll < STEP; ll++) acc = acc + v[kk * M + ll];
        }
        val[c * 3 + ch] = acc / (float)(STEP * STEP);
    }
}
__device__ __forceinline__ void desc_cell_sums_rt(const float* smp, float* val, int M, int nch, int lane, int step, int nc) {
    const float ns = (float)(step * step);
    for (int t = lane; t < nc * nc * nch; t += 32) {
        const int c = t / nch, ch = t - c * nch;
        const int ci = c / nc, cj = c - ci * nc;
        const float* v = smp + ch * kDescCS + (ci * step) * M + cj * step;
        float acc = 0.0f;
        for (int kk = 0; kk < step; kk++)
            for (int ll = 0; ll < step; ll++) acc = acc + v[kk * M + ll];
        val[c * 3 + ch] = acc / ns;
    }
}

// DEF: Config::default() descriptor geometry (pattern 10, 3 channels) as compile-time constants -- the divisions by the
// lattice edge and the channel count fold into multiplies (they were 17 % of the kernel's instructions)
template <bool DEF>
__global__ void __launch_bounds__(32 * kDescWarps)
k_descriptor(const PlanDev* __restrict__ plan, const float* __restrict__ lt_plane, const float* __restrict__ lx_plane,
             const float* __restrict__ ly_plane, int batch, unsigned int kp_cap, const akz_keypoint* __restrict__ kps,
             const unsigned int* __restrict__ n_kp, uint8_t* __restrict__ desc, unsigned int* __restrict__ err_flags) {
    __shared__ float s_smp[kDescWarps][3 * kDescCS];
    __shared__ float s_val[kDescWarps][29 * 3 + 1];
    __shared__ unsigned short s_cmp[512];  // per descriptor bit: (index of value i) | (index of value j) << 8
    const int img = blockIdx.y;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned int n = min(n_kp[img], kp_cap);
    const int warps = (blockDim.x >> 5) * gridDim.x;
    const int nch = DEF ? 3 : plan->channels;
    const int pattern = DEF ? 10 : plan->pattern_size;
    // sample_size per grid level: ceil(pattern * {1, 2/3, 1/2}) in f32 (descriptors.rs:50,61)
    const float pf = (float)pattern;
    const int st0 = (int)ceilf(pf * 1.0f), st1 = (int)ceilf(pf * (2.0f / 3.0f)), st2 = (int)ceilf(pf * (1.0f / 2.0f));
    const int M = DEF ? kDescM : max(2 * st0, max(3 * st1, 4 * st2));  // <= kDescM (validated on the host)
    const int nbits = 162 * nch;
    {   // bit -> (i, j) table, identical for every keypoint
        for (int d = threadIdx.x; d < 512; d += blockDim.x) {
            unsigned short e = 0;
            if (d < nbits) {
                int rem = d, cell_base = 0, g = 0;
                for (; g < 3; g++) {
                    const int cnt = (g + 2) * (g + 2), pairs = cnt * (cnt - 1) / 2;
                    if (rem < nch * pairs) break;
                    rem -= nch * pairs;
                    cell_base += cnt;
                }
                const int cnt = (g + 2) * (g + 2), pairs = cnt * (cnt - 1) / 2;
                const int pos = rem / pairs;
                int pidx = rem - pos * pairs, i = 0;
                while (pidx >= cnt - 1 - i) {  // unrank pidx -> (i, j), i < j, row-major over i
                    pidx -= cnt - 1 - i;
                    i++;
                }
                const int j = i + 1 + pidx;
                e = (unsigned short)(((cell_base + i) * 3 + pos) | (((cell_base + j) * 3 + pos) << 8));
            }
            s_cmp[d] = e;
        }
    }
    __syncthreads();
    float* smp = s_smp[wib];
    float* val = s_val[wib];
    for (unsigned int kidx = blockIdx.x * (blockDim.x >> 5) + wib; kidx < n; kidx += warps) {
        const akz_keypoint kp = kps[(size_t)img * kp_cap + kidx];
        const LevelDev& lv = plan->lv[kp.class_id];
        const float ratio = lv.ratio;
        const float scale = lv.s_smp;
        const int W = lv.w, H = lv.h;
        const float xf = kp.x / ratio, yf = kp.y / ratio;
        const float co = (float)cos((double)kp.angle), si = (float)sin((double)kp.angle);
        const size_t base = (size_t)lv.off * batch + (size_t)img * W * H;
        const float* Lt = lt_plane + base;
        const float* Lx = lx_plane + base;
        const float* Ly = ly_plane + base;
        bool oob = false;
        // ---- phase 1: the M x M lattice; fast lane index runs along the lattice axis closest to image x
        const bool k_fast = fabsf(co) >= fabsf(si);
        // Lanes cover the lattice in 8 x 4 blocks (8 along the fast axis): a warp-wide gather then touches the cache lines
        // of a compact patch -- about (8 |sin| + 4 |cos|) * scale image rows -- instead of those of a 32-point line that
        // climbs up to 15 * scale rows at 45 degrees (simulated over random angles: 156 instead of 193 L1 wavefronts per
        // plane and keypoint; the kernel is bound by exactly those). Two blocks per lane and iteration: the six gathers
        // are issued before the first of them is used.
        const int bx = lane & 7, by = lane >> 3;
        const int nbx = (M + 7) >> 3, nby = (M + 3) >> 2;  // 3 x 6 blocks for M = 21
        for (int blk = 0; blk < nbx * nby; blk += 2) {
            float v_t[2], v_x[2], v_y[2];
            int so[2];
            bool act[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int bl = blk + u;
                const int bj = bl / nbx, bi = bl - bj * nbx;  // block row (slow axis), block column (fast axis)
                const int a = bj * 4 + by, b = bi * 8 + bx;
                act[u] = bl < nbx * nby && a < M && b < M;
                so[u] = 0;
                v_t[u] = v_x[u] = v_y[u] = 0.0f;
                if (act[u]) {
                    const int kk = k_fast ? b : a, ll = k_fast ? a : b;
                    const float lf = (float)(ll - pattern) + 0.5f, kf = (float)(kk - pattern) + 0.5f;
                    const float sample_y = yf + (lf * co * scale + kf * si * scale);
                    const float sample_x = xf + (-lf * si * scale + kf * co * scale);
                    int y1 = (int)roundf(sample_y), x1 = (int)roundf(sample_x);
                    if (x1 < 0 || y1 < 0 || x1 >= W || y1 >= H) {
                        oob = true;
                        x1 = min(max(x1, 0), W - 1);
                        y1 = min(max(y1, 0), H - 1);
                    }
                    const int at = y1 * W + x1;  // one level image is far below 2^31 pixels
                    so[u] = kk * M + ll;
                    v_t[u] = Lt[at];
                    if (nch > 1) {
                        v_x[u] = Lx[at];
                        v_y[u] = Ly[at];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; u++)
                if (act[u]) {
                    smp[so[u]] = v_t[u];
                    if (nch > 1) {
                        const float rx = v_x[u], ry = v_y[u];
                        if (nch == 2) {
                            smp[kDescCS + so[u]] = sqrtf(rx * rx + ry * ry);
                        } else {
                            const float rry = rx * co + ry * si;
                            const float rrx = -rx * si + ry * co;
                            smp[kDescCS + so[u]] = rrx;
                            smp[2 * kDescCS + so[u]] = rry;
                        }
                    }
                }
        }
        __syncwarp();
        // ---- phase 2: sequential sums per (cell, channel); values of grid g start at cell_base(g) = 0, 4, 13
        if (DEF || pattern == 10) {
            desc_cell_sums<10, 2>(smp, val, M, nch, lane);
            desc_cell_sums<7, 3>(smp, val + 4 * 3, M, nch, lane);
            desc_cell_sums<5, 4>(smp, val + 13 * 3, M, nch, lane);
        } else {
            desc_cell_sums_rt(smp, val, M, nch, lane, st0, 2);
            desc_cell_sums_rt(smp, val + 4 * 3, M, nch, lane, st1, 3);
            desc_cell_sums_rt(smp, val + 13 * 3, M, nch, lane, st2, 4);
        }
        __syncwarp();
        // ---- phase 3: bit d of the descriptor = values[i] > values[j]; word w is the ballot of bits 32w..32w+31
        unsigned int my_word = 0;
#pragma unroll
        for (int w = 0; w < 16; w++) {
            const int d = w * 32 + lane;
            const unsigned int e = s_cmp[d];
            const bool bit = d < nbits && val[e & 255u] > val[e >> 8];
            const unsigned int word = __ballot_sync(0xffffffffu, bit);
            if (lane == w) my_word = word;
        }
        if (lane < 16) {
            unsigned int* out = (unsigned int*)(desc + ((size_t)img * kp_cap + kidx) * kDescStride);
            out[lane] = my_word;
        }
        if (oob) atomicOr(&err_flags[img], (unsigned int)kErrBounds);
        __syncwarp();
    }
}

}  // namespace

template <int KG, int MAXT>
static cudaError_t level_pass_attributes() {
    // see init_detector_attributes: the cache pass must not pin a small shared-memory carveout on the SMs it lives on
    cudaError_t e = cudaFuncSetAttribute(k_dedup_levels<KG, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess || getenv("AKZ_NO_CARVEOUT") != nullptr) return e;
    return cudaFuncSetAttribute(k_dedup_levels<KG, MAXT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

cudaError_t init_keypoint_attributes() {
    cudaError_t e = level_pass_attributes<8, 512>();
    if (e != cudaSuccess) return e;
    e = level_pass_attributes<8, 1024>();
    if (e != cudaSuccess) return e;
    e = level_pass_attributes<16, 512>();
    if (e != cudaSuccess) return e;
    e = level_pass_attributes<16, 1024>();
    if (e != cudaSuccess) return e;
    if (getenv("AKZ_NO_CARVEOUT") != nullptr) return cudaSuccess;
    e = cudaFuncSetAttribute(k_dedup_smem, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_dedup, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

size_t dedup_pool_bytes(const Plan& P) { return (size_t)P.dev.pool_cap * kPoolBytesPerEntry; }
size_t dedup_level_pool_bytes(const Plan& P, uint32_t cand_cap) {  // per image: pool entries, row tables, (appends, entries) per level
    const size_t b = (size_t)cand_cap * kLevelPoolBytesPerCand + (size_t)P.dev.ltab_off[P.dev.n_levels] * sizeof(unsigned int) +
                     2 * kMaxLevels * sizeof(unsigned int);
    return (b + 255) & ~(size_t)255;
}
template <int KG>
static size_t level_pass_smem(const Plan& P) {
    return (size_t)P.dev.n_levels * KG * sizeof(uint4) + (size_t)P.dev.ltab_off[P.dev.n_levels] * sizeof(unsigned int);
}

template <int KG>
static void launch_level_pass(const Launch& L, const Plan& P, const Buffers& B, int l0, int l1) {
    // 512-thread blocks (the default 16 levels) may use 128 registers per thread, 1024-thread ones 64
    const size_t smem = level_pass_smem<KG>(P);
    const size_t slab = dedup_level_pool_bytes(P, L.cand_cap);
    const int warps = l1 == P.dev.n_levels ? l1 : l1 - l0;  // the part that assigns the final slots has a warp per level
    if (P.dev.n_levels <= 16)
        k_dedup_levels<KG, 512><<<L.batch, 32 * warps, smem, L.stream>>>(B.plan_dev, B.cand, B.cand_level_count, B.Ldet, L.batch, L.cand_cap,
                                                                         L.kp_cap, B.c_x, B.c_y, B.c_resp, B.c_cls, B.n_cache, B.err_flags,
                                                                         B.level_pool, B.keep_flag, B.upper_done, slab, l0, l1);
    else
        k_dedup_levels<KG, 1024><<<L.batch, 32 * warps, smem, L.stream>>>(B.plan_dev, B.cand, B.cand_level_count, B.Ldet, L.batch, L.cand_cap,
                                                                          L.kp_cap, B.c_x, B.c_y, B.c_resp, B.c_cls, B.n_cache, B.err_flags,
                                                                          B.level_pool, B.keep_flag, B.upper_done, slab, l0, l1);
}

static bool level_pass_enabled(const Plan& P) {
    static const bool single_warp = getenv("AKZ_DEDUP_SINGLE") != nullptr;  // A/B switch: the one-warp-per-image pass
    return !single_warp && level_pass_smem<16>(P) <= 200 * 1024;
}

// the first octave's levels go first when the pass is split (they hold ~60 % of the candidates and are known after a
// quarter of the sub-batch's stencil time at the default 4 x 4 levels)
int dedup_split_level(const Plan& P) {
    // Opt-in (AKZ_SPLIT_PASS=1). Measured: the second octave's levels, not the first's, are the critical path of the pass (they
    // scan more rows of the level before them per candidate), so the part that has to wait for the last detector is nearly as
    // long as the whole pass and no longer overlaps the first octave's levels: a single 3840x2160 image takes 7.0 ms against
    // 5.6 ms unsplit (the same with the first part alone on its SM), 32 of them per step run at 1413 against 1394 images/s, a
    // single 1080p image 1.73 against 1.70 ms. The pass stays in one piece by default.
    static const bool split = getenv("AKZ_SPLIT_PASS") != nullptr;
    if (!split || !level_pass_enabled(P)) return 0;
    int l = 1;
    while (l < P.dev.n_levels && !P.dev.lv[l].new_octave) l++;
    return l < P.dev.n_levels ? l : 0;
}

int launch_dedup(const Launch& L, const Plan& P, const Buffers& B, int l0, int l1) {
    // every image is handled by exactly one of the kernels launched here (image_fits_*_pass)
    static const int groups = getenv("AKZ_DEDUP_GROUPS") ? atoi(getenv("AKZ_DEDUP_GROUPS")) : 8;  // A/B switch: candidates per step
    if (l1 < 0) l1 = P.dev.n_levels;
    const bool levels = level_pass_enabled(P);
    if (levels && groups == 16) {
        launch_level_pass<16>(L, P, B, l0, l1);
    } else if (levels) {
        launch_level_pass<8>(L, P, B, l0, l1);
    } else {
        k_dedup_smem<<<L.batch, 32, 0, L.stream>>>(B.plan_dev, B.cand, B.cand_level_count, B.Ldet, L.batch, L.cand_cap, L.kp_cap, B.c_x,
                                                   B.c_y, B.c_resp, B.c_cls, B.n_cache, B.err_flags, B.dedup_pool, B.upper_done);
    }
    if (l1 != P.dev.n_levels) return 1;  // (only the level pass is ever launched in parts)
    k_dedup<<<L.batch, 32, 0, L.stream>>>(B.plan_dev, B.cand, B.cand_level_count, B.Ldet, L.batch, L.cand_cap, L.kp_cap,
                                           B.c_x, B.c_y, B.c_resp, B.c_cls, B.c_next, B.grid, B.n_cache, B.err_flags, levels ? 1 : 0, B.upper_done);
    return 2;
}

int launch_finalize(const Launch& L, const Plan& P, const Buffers& B) {
    float* r_x = B.r_x;
    float* r_y = B.r_y;
    dim3 g1(64, L.batch);
    k_class_ranges<<<L.batch, 256, 0, L.stream>>>(B.c_cls, B.n_cache, L.kp_cap, B.cls_range);
    k_filter_refine<<<g1, 256, 0, L.stream>>>(B.plan_dev, B.Ldet, L.batch, L.kp_cap, B.c_x, B.c_y, B.c_cls, B.n_cache, B.cls_range,
                                              r_x, r_y, B.keep_flag, B.upper_done);
    k_keep_scan<<<L.batch, 1024, 0, L.stream>>>(B.keep_flag, B.n_cache, B.n_kp, L.kp_cap);
    // a single image or a small batch spreads over more, hence shorter, blocks (16 blocks per image fill the GPU from ~20 images)
    dim3 g3(std::min(128, std::max(16, 148 * 2 / std::max(1, L.batch))), L.batch);
    k_orientation<<<g3, 128, 0, L.stream>>>(B.plan_dev, B.Lx, B.Ly, L.batch, L.kp_cap, r_x, r_y, B.c_resp, B.c_cls, B.keep_flag,
                                            B.n_cache, B.kps, B.err_flags);
    return 4;
}

int launch_descriptors(const Launch& L, const Plan& P, const Buffers& B) {
    dim3 g(std::min(256, std::max(32, 148 * 4 / std::max(1, L.batch))), L.batch);  // small batches: more blocks per image
    if (P.dev.channels == 3 && P.dev.pattern_size == 10)
        k_descriptor<true><<<g, 256, 0, L.stream>>>(B.plan_dev, B.Lt, B.Lx, B.Ly, L.batch, L.kp_cap, B.kps, B.n_kp, B.desc, B.err_flags);
    else
        k_descriptor<false><<<g, 256, 0, L.stream>>>(B.plan_dev, B.Lt, B.Lx, B.Ly, L.batch, L.kp_cap, B.kps, B.n_kp, B.desc, B.err_flags);
    return 1;
}

}  // namespace akz
