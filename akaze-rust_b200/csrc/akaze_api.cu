// akaze_api.cu -- C ABI (include/akaze_b200.h), level-table construction and launch orchestration.
//
// Host-side restatement of the reference's tiny scalar set-up code (all f64/f32 exactly as written):
//   Config::default                 akaze/src/types/evolution.rs:41-54
//   EvolutionStep::new              akaze/src/types/evolution.rs:101-126
//   allocate_evolutions             akaze/src/types/evolution.rs:135-161
//   fed_tau_by_process_time ...     akaze/src/ops/fed_tau.rs:27-106
//   gaussian / gaussian_kernel      akaze/src/types/image.rs:341-365
//   scharr_main_axis_kernel         akaze/src/ops/derivatives.rs:91-101
// and the orchestration of create_nonlinear_scale_space / find_image_keypoints / extract_features
// (akaze/src/lib.rs:49-194) as a fixed sequence of kernel launches on one stream.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <memory>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "nccl_dyn.h"

namespace akz {

static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
    return fail(AKZ_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                              \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

// ---- FED schedule (fed_tau.rs) -------------------------------------------------------------------
static bool is_prime(uint64_t n) {
    if (n < 2) return false;
    for (uint64_t d = 2; d * d <= n; d++)
        if (n % d == 0) return false;
    return true;
}

// fed_tau.rs:61-106 with reordering = true
static bool fed_tau_internal(size_t n, double scale, double tau_max, std::vector<double>* tau) {
    tau->assign(n, 0.0);
    if (n == 0) return true;
    std::vector<double> tauh(n);
    const double c = 1.0 / (4.0 * (double)n + 2.0);
    const double d = scale * tau_max / 2.0;
    const double pi = 3.14159265358979323846264338327950288;
    for (size_t k = 0; k < n; k++) {
        const double h = cos(pi * (2.0 * (double)k + 1.0) * c);
        tauh[k] = d / (h * h);
    }
    const size_t kappa = n / 2;
    if (kappa == 0) return false;  // n == 1: the reference underflows and never terminates (Q12)
    size_t prime = n + 1;
    while (!is_prime(prime)) prime += 1;
    size_t k = 0;
    for (size_t t = 0; t < n; t++) {
        // usize arithmetic of the reference: a product that is a multiple of `prime` wraps to
        // usize::MAX and is skipped by `index >= n`
        size_t m = ((k + 1) * kappa) % prime;
        while (m == 0 || m - 1 >= n) {
            k += 1;
            m = ((k + 1) * kappa) % prime;
        }
        (*tau)[t] = tauh[m - 1];
        k += 1;
    }
    return true;
}

// fed_tau.rs:27-49
static bool fed_tau_by_process_time(double T, int M, double tau_max, std::vector<double>* tau) {
    const double t = T / (double)M;
    const double nf = ceil(sqrt(3.0 * t / tau_max + 0.25) - 0.5 - 1.0e-8) + 0.5;
    const size_t n = nf > 0.0 ? (size_t)nf : 0;
    const double scale = 3.0 * t / (tau_max * (double)(n * (n + 1)));
    return fed_tau_internal(n, scale, tau_max, tau);
}

// image.rs:341-365, f32 throughout
static void gaussian_kernel(float r, int size, float* k) {
    const int hw = size / 2;
    const float pi = 3.14159265358979323846f;
    float sum = 0.0f;
    for (int i = -hw; i <= hw; i++) {
        const float x = (float)i;
        const float a = 1.0f / (sqrtf(2.0f * pi) * r);
        const float v = a * expf(-(x * x) / (2.0f * (r * r)));
        k[i + hw] = v;
        sum += v;
    }
    for (int i = 0; i < size; i++) k[i] /= sum;
}

std::string build_plan(uint32_t w, uint32_t h, const akz_config& cfg, Plan* P) {
    if (cfg.num_sublevels == 0 || cfg.max_octave_evolution == 0) return "num_sublevels and max_octave_evolution must be >= 1";
    if ((uint64_t)cfg.num_sublevels * cfg.max_octave_evolution > (uint64_t)kMaxLevels) return "too many evolution levels (max 32)";
    if (cfg.descriptor_channels < 1 || cfg.descriptor_channels > 3) return "descriptor_channels must be 1, 2 or 3";
    if (cfg.contrast_factor_num_bins < 1 || cfg.contrast_factor_num_bins > (uint64_t)kMaxBins) return "contrast_factor_num_bins must be in 1..1024";
    if (!(cfg.base_scale_offset > 0.0) || cfg.base_scale_offset > 4.0) return "base_scale_offset must be in (0, 4]";
    if (w < 64 || h < 32) return "image must be at least 64x32";
    if (cfg.descriptor_pattern_size < 2 || cfg.descriptor_pattern_size > 64) return "descriptor_pattern_size must be in 2..64";
    {
        const float pf = (float)cfg.descriptor_pattern_size;
        const int p = (int)cfg.descriptor_pattern_size;
        const int s0 = (int)ceilf(pf * 1.0f), s1 = (int)ceilf(pf * (2.0f / 3.0f)), s2 = (int)ceilf(pf * (1.0f / 2.0f));
        const int n0 = (2 * p + s0 - 1) / s0, n1 = (2 * p + s1 - 1) / s1, n2 = (2 * p + s2 - 1) / s2;
        if (n0 != 2 || n1 != 3 || n2 != 4) return "descriptor_pattern_size does not give the 2x2/3x3/4x4 MLDB grids";
        if (2 * s0 > 21 || 3 * s1 > 21 || 4 * s2 > 21) return "descriptor_pattern_size too large (max 10)";  // kDescM lattice of k_descriptor
    }
    P->w = w;
    P->h = h;
    P->cfg = cfg;
    PlanDev& D = P->dev;
    memset(&D, 0, sizeof(D));
    P->host.clear();

    // evolution.rs:135-150
    for (uint32_t i = 0; i < cfg.max_octave_evolution; i++) {
        const double rfactor = 1.0 / pow(2.0, (double)i);
        const uint32_t level_height = (uint32_t)((double)h * rfactor);
        const uint32_t level_width = (uint32_t)((double)w * rfactor);
        if ((level_width >= 80 && level_height >= 40) || i == 0) {
            for (uint32_t j = 0; j < cfg.num_sublevels; j++) {
                LevelHost lh;
                // evolution.rs:101-113
                const double esigma = cfg.base_scale_offset * pow(2.0, (double)j / (double)cfg.num_sublevels + (double)i);
                lh.info.esigma = esigma;
                lh.info.etime = 0.5 * (esigma * esigma);
                lh.info.octave = i;
                lh.info.sublevel = j;
                lh.info.sigma_size = (uint32_t)round(esigma);
                lh.info.n_steps = 0;
                // image.rs:103-104 applied once per octave: floor halving
                lh.info.width = w >> i;
                lh.info.height = h >> i;
                P->host.push_back(lh);
            }
        } else {
            break;
        }
    }
    const int nl = (int)P->host.size();
    // evolution.rs:151-153
    for (int i = 1; i < nl; i++) {
        const double ttime = P->host[i].info.etime - P->host[i - 1].info.etime;
        if (!fed_tau_by_process_time(ttime, 1, 0.25, &P->host[i].tau)) return "FED schedule with a single step (the reference does not terminate on it)";
        if (P->host[i].tau.size() > (size_t)kMaxFedSteps) return "too many FED steps per level";
        P->host[i].info.n_steps = (uint32_t)P->host[i].tau.size();
        P->host[i].half_tau.clear();
        for (double t : P->host[i].tau) P->host[i].half_tau.push_back(0.5f * (float)t);  // nonlinear_diffusion.rs:67
    }

    // filter taps
    {
        const float r0 = (float)cfg.base_scale_offset;  // lib.rs:56
        const int size0 = (int)ceilf(r0) * 2 + 1;       // image.rs:376
        if (size0 > kMaxGaussTaps) return "base_scale_offset too large";
        P->gauss0_n = size0;
        gaussian_kernel(r0, size0, P->gauss0);
        gaussian_kernel(1.0f, 3, P->gauss1);
        for (int s = 1; s <= kMaxDetScale; s++) {
            const double wgt = 10.0 / 3.0;  // derivatives.rs:94-99
            const double norm = 1.0 / (2.0 * (double)s * (wgt + 2.0));
            P->sch_n[s] = (float)norm;
            P->sch_wn[s] = (float)(wgt * norm);
        }
        P->sch_n[0] = P->sch_wn[0] = 0.0f;
    }

    D.n_levels = nl;
    D.w0 = (int)w;
    D.h0 = (int)h;
    D.channels = (int)cfg.descriptor_channels;
    D.pattern_size = (int)cfg.descriptor_pattern_size;
    D.desc_len = (int)((162 * cfg.descriptor_channels + 7) / 8);
    D.n_bins = (int)cfg.contrast_factor_num_bins;
    D.det_threshold = (float)cfg.detector_threshold;
    D.percentile = cfg.contrast_percentile;
    unsigned long long off = 0, moff = 0;
    const float smax = 10.0f * sqrtf(2.0f);  // scale_space_extrema.rs:14
    for (int l = 0; l < nl; l++) {
        LevelDev& lv = D.lv[l];
        const akz_level_info& in = P->host[l].info;
        lv.w = (int)in.width;
        lv.h = (int)in.height;
        if (lv.w < 64 || lv.h < 32) return "evolution level smaller than one tile (64x32)";
        lv.octave = (int)in.octave;
        // detector_response.rs:22-24
        const double ratio_d = pow(2.0, (double)in.octave);
        const double sd = round(in.esigma * cfg.derivative_factor / ratio_d);
        if (!(sd >= 1.0) || sd > (double)kMaxDetScale) return "detector Scharr scale out of range 1..6 (derivative_factor / scale settings)";
        lv.s_det = (int)sd;
        lv.wpr = (lv.w + 31) / 32;
        lv.new_octave = (l > 0 && P->host[l].info.octave > P->host[l - 1].info.octave) ? 1 : 0;
        lv.n_steps = (int)in.n_steps;
        lv.ratio = powf(2.0f, (float)in.octave);                        // scale_space_extrema.rs:51
        lv.kp_size = (float)(in.esigma * cfg.derivative_factor);        // :45
        lv.size_sq = lv.kp_size * lv.kp_size;                           // :66
        lv.s_smp = roundf(0.5f * lv.kp_size / lv.ratio);                // :280, descriptors.rs:52
        lv.half_ratio_m1 = 0.5f * (lv.ratio - 1.0f);                    // :90
        // is_out (:80-87) evaluated per coordinate with the reference's own f32 expressions
        const float sigma_size = roundf(lv.kp_size / lv.ratio);         // :52
        lv.xmin = lv.w;
        lv.xmax = -1;
        for (int x = 1; x < lv.w - 1; x++) {
            const float left_x = roundf((float)x - smax * sigma_size) - 1.0f;
            const float right_x = roundf((float)x + smax * sigma_size) + 1.0f;
            if (!(left_x < 0.0f || right_x >= (float)lv.w)) {
                lv.xmin = std::min(lv.xmin, x);
                lv.xmax = std::max(lv.xmax, x);
            }
        }
        lv.ymin = lv.h;
        lv.ymax = -1;
        for (int y = 1; y < lv.h - 1; y++) {
            const float up_y = roundf((float)y - smax * sigma_size) - 1.0f;
            const float down_y = roundf((float)y + smax * sigma_size) + 1.0f;
            if (!(up_y < 0.0f || down_y >= (float)lv.h)) {
                lv.ymin = std::min(lv.ymin, y);
                lv.ymax = std::max(lv.ymax, y);
            }
        }
        lv.off = off;
        lv.mask_off = moff;
        off += (unsigned long long)lv.w * lv.h;
        moff += (unsigned long long)lv.wpr * lv.h;
    }
    D.plane_px = off;
    D.mask_words = moff;

    // orientation windows (scale_space_extrema.rs:300-319): which windows contain atan2(v,v), v > 0
    {
        const float PI = 3.14159265358979323846f;
        const float ang = atan2f(1.0f, 1.0f);
        float ang1 = 0.0f;
        int wi = 0;
        unsigned long long m = 0;
        while (ang1 < 2.0f * PI) {
            const float ang2 = (ang1 + PI / 3.0f > 2.0f * PI) ? (ang1 - 5.0f * PI / 3.0f) : (ang1 + PI / 3.0f);
            ang1 += 0.15f;
            const bool in = (ang1 < ang2 && ang1 < ang && ang < ang2) ||
                            (ang2 < ang1 && ((ang > 0.0f && ang < ang2) || (ang > ang1 && ang < 2.0f * PI)));
            if (in && wi < 64) m |= 1ull << wi;
            wi++;
        }
        if (wi > 64) return "internal: too many orientation windows";
        D.orient_window_mask = m;
        D.n_orient_windows = wi;
    }
    // dedup hash grid: cells of 16 px, doubled until the grid fits the shared-memory pass (8448 cells)
    D.grid_shift = 4;
    while ((((int)w >> D.grid_shift) + 1) * (((int)h >> D.grid_shift) + 1) > 8448) D.grid_shift++;
    D.grid_w = ((int)w >> D.grid_shift) + 1;
    D.grid_h = ((int)h >> D.grid_shift) + 1;
    // cache-pass pools: room for the candidates of the busiest level (a photo-like 1080p frame has ~2 k per level, a
    // 3840x2160 one ~10 k); images beyond it take the global-memory pass
    D.pool_cap = (int)std::min<uint64_t>(64512, std::max<uint64_t>(4096, (((uint64_t)w * h / 384) + 1023) & ~1023ull));
    // level-pipelined cache pass: one row table per level in shared memory (a 1080p frame: 32 KB, 3840x2160: 65 KB)
    D.ltab_off[0] = 0;
    for (int l = 0; l < nl; l++) D.ltab_off[l + 1] = D.ltab_off[l] + ((D.lv[l].h + 2 + 7) & ~7);
    return "";
}

}  // namespace akz

using namespace akz;

// ---- context -----------------------------------------------------------------------------------
// One extraction call is cut into sub-batches that flow through a software pipeline (run_pipeline):
//   S(i) (stream `stream`):     stencil work -- level 0, contrast, per level prep / FED / detector, candidate compaction;
//   D(i) (stream `stream_kp`):  the latency-bound cache pass, one warp per image, started after S(i) and F(i-1);
//   F(i) (stream `stream`):     filter/refine, orientation, descriptors, enqueued after S(i+1).
// Two "lanes" of work buffers alternate between consecutive sub-batches, so D(i) overlaps S(i+1) and nothing else does.
// Results of all images of the call live in context-level arrays.
struct Lane {
    Buffers buf;
    std::vector<void*> allocs;
    int alloc_batch = 0;
    cudaEvent_t ev_stencil = nullptr;  // stage A of the sub-batch using this lane has finished
    cudaEvent_t ev_dedup = nullptr;    // the cache pass of the sub-batch using this lane has finished
    cudaEvent_t ev_compact = nullptr;  // the candidate lists of the first part of a two-part cache pass are written
    cudaEvent_t ev_done = nullptr;     // stage B of the sub-batch using this lane has finished
    cudaEvent_t ev_prep[kMaxLevels] = {nullptr};  // Lsmooth_l (and for level 1 the consumers of the stored gradients) ready
    cudaEvent_t ev_det[kMaxLevels] = {nullptr};   // detector(l) has finished
    bool busy = false;
};

// per-image statistics mirrored into pinned host memory after stage B of each sub-batch
struct HostStats {
    unsigned int* n_kp = nullptr;       // [n]
    unsigned int* n_cache = nullptr;
    unsigned int* n_cand = nullptr;
    unsigned int* err = nullptr;
    double* kcontrast = nullptr;        // [n][kMaxLevels]
    void* base = nullptr;
    int alloc_n = 0;
};

// pinned host slab that receives the keypoints + descriptors of one sub-batch; shared by the akz_features
// handles cut from it and recycled by the context's pool once they have all been freed
struct PinnedChunk {
    void* p = nullptr;
    size_t cap = 0;
    ~PinnedChunk() {
        if (p) cudaFreeHost(p);
    }
};

struct Results {
    akz_keypoint* kps = nullptr;        // [n][kp_cap]
    uint8_t* desc = nullptr;            // [n][kp_cap][64]
    unsigned int* n_kp = nullptr;       // [n]
    unsigned int* n_cache = nullptr;
    unsigned int* n_cand = nullptr;
    unsigned int* err = nullptr;
    double* kcontrast = nullptr;        // [n][kMaxLevels]
    uint8_t* in_u8 = nullptr;           // staging for host inputs, [n][h][w]
    float* in_f32 = nullptr;
    std::vector<void*> allocs;
    int alloc_n = 0;
};

struct akz_context {
    int device = 0;
    uint32_t max_w = 0, max_h = 0, max_batch = 0, flags = 0;
    std::recursive_mutex mu;           // every entry point that touches the context holds it: thread-safe per context
    uint64_t id = 0;                   // unique per akz_create (handles check it: an address can be reused)
    uint32_t cand_cap = 262144, kp_cap = 65536;
    bool caps_auto = true;             // capacities follow the image size until akz_context_set_limits is called
    uint32_t sub_batch = 128;          // images per pipeline sub-batch
    bool sub_batch_auto = true;        // chosen from the free device memory when the plan changes (see prepare)
    cudaStream_t stream = nullptr;     // stage A, copies, and the stream callers may time on
    cudaStream_t stream_kp = nullptr;  // stage B
    cudaStream_t stream_det = nullptr; // the detectors: detector(l) only needs Lsmooth_l, so it runs beside FED(l) / prep(l+1)
    cudaStream_t stream_copy = nullptr;  // host -> device staging of the inputs, one event per sub-batch
    cudaStream_t stream_d2h = nullptr;   // device -> host copies of the results, sub-batch by sub-batch
    std::vector<cudaEvent_t> ev_copy;
    std::vector<cudaEvent_t> ev_stats;   // statistics of sub-batch i are in pinned memory (implies its stage B is done)
    HostStats hs;
    std::vector<std::shared_ptr<PinnedChunk>> pinned_pool;
    uint64_t launches = 0;
    uint64_t generation = 0;
    bool have_plan = false;
    Plan plan;
    Lane lane[2];
    Results res;
    int cur_batch = 0;
    uint32_t n_sub_batches = 0;
    // a call that is one sub-batch long (a single image, a small batch) is launch-bound: ~90 launches and ~70 event
    // operations cost the host more than the kernels cost the GPU. Its launch sequence is captured once into a CUDA graph
    // and replayed while nothing it depends on changes (GraphKey).
    struct GraphKey {
        uint64_t alloc_epoch = 0;
        const void* in = nullptr;
        const void* stats = nullptr;
        size_t stride = 0;
        uint32_t n = 0, cand_cap = 0, kp_cap = 0;
        bool is_u8 = false;
        bool operator==(const GraphKey& o) const {
            return alloc_epoch == o.alloc_epoch && in == o.in && stats == o.stats && stride == o.stride && n == o.n && cand_cap == o.cand_cap &&
                   kp_cap == o.kp_cap && is_u8 == o.is_u8;
        }
    };
    struct CachedGraph {
        GraphKey key;
        cudaGraphExec_t exec = nullptr;
        int launches = 0;
        uint64_t last_use = 0;
    };
    std::vector<CachedGraph> graphs;  // a few of them, least recently used out first: callers that rotate input buffers
    uint64_t graph_replays = 0;
    std::vector<std::pair<uint32_t, uint32_t>> sched;  // (first image, count) of every sub-batch of the current call
    // per-stage timing
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
    std::vector<int> ev_stage;       // stage of each used event pair
    std::vector<int> ev_launches;
    size_t ev_used = 0;
    double stage_ms[AKZ_NUM_STAGES] = {0};
    uint64_t stage_launches[AKZ_NUM_STAGES] = {0};
    // matcher scratch
    void* m_q = nullptr;
    void* m_db = nullptr;
    void* m_parts = nullptr;
    void* m_out = nullptr;
    size_t m_q_cap = 0, m_db_cap = 0, m_parts_cap = 0, m_out_cap = 0;
    void* m_qimg = nullptr;   // int8 operand images of the tensor-core matcher (matcher_tc.cu)
    void* m_dbimg = nullptr;
    size_t m_qimg_cap = 0, m_dbimg_cap = 0;
    int match_path = AKZ_MATCH_AUTO;
    // multi-GPU matching (akz_context_comm_init*, akz_match_top2_sharded*)
    ncclComm_t comm = nullptr;
    int comm_rank = 0, comm_size = 1;
    void* m_gather = nullptr;
    size_t m_gather_cap = 0;
    void* m_stage = nullptr;  // host descriptor rows as uploaded, before they are repacked to 64-byte rows
    size_t m_stage_cap = 0;
};

// live contexts by id: a handle that outlives its context must fail cleanly instead of touching freed memory
static std::mutex g_registry_mu;
static std::unordered_map<uint64_t, akz_context*> g_registry;
static std::atomic<uint64_t> g_next_ctx_id{1};

static akz_context* live_context(uint64_t id) {
    std::lock_guard<std::mutex> g(g_registry_mu);
    auto it = g_registry.find(id);
    return it == g_registry.end() ? nullptr : it->second;
}

struct akz_features {
    akz_context* ctx = nullptr;
    uint64_t ctx_id = 0;
    uint64_t generation = 0;
    int img = 0, batch = 0;            // index inside its pipeline sub-batch, images of that sub-batch
    int lane = 0;                      // which set of work buffers the sub-batch used
    uint32_t sub_batch_index = 0, n_sub_batches = 1;
    std::shared_ptr<PinnedChunk> chunk;                      // owns the memory kps/desc point into
    const akz_keypoint* kps = nullptr;
    const uint8_t* desc = nullptr;
    size_t n = 0;
    std::shared_ptr<const std::vector<LevelHost>> levels;   // shared by all features of one call
    uint32_t desc_len = 0;
    double contrast = 0.0;
    uint64_t n_cand = 0, n_cache = 0;
};

static void free_lane(Lane& ln) {
    for (void* p : ln.allocs) cudaFree(p);
    ln.allocs.clear();
    ln.buf = Buffers();
    ln.alloc_batch = 0;
    ln.busy = false;
}
static void free_results(Results& r) {
    for (void* p : r.allocs) cudaFree(p);
    r = Results();
}
static void free_buffers(akz_context* c) {
    c->generation++;  // handles of earlier extractions must not read the freed planes
    free_lane(c->lane[0]);
    free_lane(c->lane[1]);
    free_results(c->res);
}

static std::atomic<uint64_t> g_alloc_epoch{1};  // bumped by every device allocation: a captured graph holds raw pointers

template <class T>
static cudaError_t dalloc(std::vector<void*>& owner, T** p, size_t count) {
    void* v = nullptr;
    g_alloc_epoch++;
    cudaError_t e = cudaMalloc(&v, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    owner.push_back(v);
    *p = (T*)v;
    return cudaSuccess;
}

static int ensure_lane(akz_context* c, Lane& ln, int batch) {
    if (batch <= ln.alloc_batch) return AKZ_OK;
    free_lane(ln);
    const Plan& P = c->plan;
    Buffers& B = ln.buf;
    std::vector<void*>& A = ln.allocs;
    const size_t nb = (size_t)batch;
    const size_t plane = (size_t)P.dev.plane_px * nb;
    const size_t n0 = (size_t)P.w * P.h * nb;
    B.keep = (c->flags & AKZ_KEEP_EVOLUTIONS) != 0;
    CK(dalloc(A, &B.Lt, plane));
    CK(dalloc(A, &B.Lx, plane));
    CK(dalloc(A, &B.Ly, plane));
    CK(dalloc(A, &B.Ldet, plane));
    CK(dalloc(A, &B.Lsmooth, B.keep ? plane : n0));
    if (!B.keep) CK(dalloc(A, &B.Lsmooth2, n0));
    CK(dalloc(A, &B.Lflow, B.keep ? plane : n0));
    CK(dalloc(A, &B.Ltmp, n0));
    if (B.keep) {
        CK(dalloc(A, &B.Lxx, plane));
        CK(dalloc(A, &B.Lyy, plane));
        CK(dalloc(A, &B.Lxy, plane));
        CK(dalloc(A, &B.Lstep, plane));
        CK(cudaMemset(B.Lstep, 0, plane * sizeof(float)));
    }
    CK(dalloc(A, &B.hmax_bits, nb));
    CK(dalloc(A, &B.hist, nb * kMaxBins));
    CK(dalloc(A, &B.contrast_thr, nb * (kMaxBins + 1)));
    CK(dalloc(A, &B.fine_hist, nb * contrast_fine_bins()));
    CK(dalloc(A, &B.contrast_resolved, nb));
    CK(dalloc(A, &B.contrast_npoints, nb));
    CK(dalloc(A, &B.mask, (size_t)P.dev.mask_words * nb));
    CK(dalloc(A, &B.cand, (size_t)c->cand_cap * nb));
    size_t total_rows = 0;
    for (int l = 0; l < P.dev.n_levels; l++) total_rows += P.dev.lv[l].h;
    CK(dalloc(A, &B.rowcount, total_rows * nb));
    CK(dalloc(A, &B.cand_level_count, nb * (kMaxLevels + 1)));
    const size_t kc = (size_t)c->kp_cap * nb;
    CK(dalloc(A, &B.c_x, kc));
    CK(dalloc(A, &B.c_y, kc));
    CK(dalloc(A, &B.c_resp, kc));
    CK(dalloc(A, &B.r_x, kc));
    CK(dalloc(A, &B.r_y, kc));
    CK(dalloc(A, &B.c_cls, kc));
    CK(dalloc(A, &B.c_next, kc));
    CK(dalloc(A, &B.grid, nb * 2 * (size_t)P.dev.grid_w * P.dev.grid_h));
    CK(dalloc(A, &B.dedup_pool, nb * dedup_pool_bytes(P)));
    CK(dalloc(A, &B.level_pool, nb * dedup_level_pool_bytes(P, c->cand_cap)));
    CK(dalloc(A, &B.keep_flag, kc));
    CK(dalloc(A, &B.cls_range, nb * kMaxLevels * 2));
    CK(dalloc(A, &B.upper_done, nb));
    CK(dalloc(A, &B.plan_dev, 1));
    CK(cudaMemcpy(B.plan_dev, &P.dev, sizeof(PlanDev), cudaMemcpyHostToDevice));
    if (!ln.ev_stencil) {
        CK(cudaEventCreateWithFlags(&ln.ev_stencil, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ln.ev_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ln.ev_dedup, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ln.ev_compact, cudaEventDisableTiming));
        for (int l = 0; l < kMaxLevels; l++) {
            CK(cudaEventCreateWithFlags(&ln.ev_prep[l], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ln.ev_det[l], cudaEventDisableTiming));
        }
    }
    ln.alloc_batch = batch;
    return AKZ_OK;
}

static int ensure_results(akz_context* c, int n) {
    Results& R = c->res;
    if (n <= R.alloc_n) return AKZ_OK;
    free_results(R);
    const size_t nb = (size_t)n, kc = (size_t)c->kp_cap * nb;
    const size_t n0 = (size_t)c->plan.w * c->plan.h * nb;
    CK(dalloc(R.allocs, &R.kps, kc));
    CK(dalloc(R.allocs, &R.desc, kc * kDescStride));
    CK(dalloc(R.allocs, &R.n_kp, nb));
    CK(dalloc(R.allocs, &R.n_cache, nb));
    CK(dalloc(R.allocs, &R.n_cand, nb));
    CK(dalloc(R.allocs, &R.err, nb));
    CK(dalloc(R.allocs, &R.kcontrast, nb * kMaxLevels));
    CK(dalloc(R.allocs, &R.in_u8, n0));
    CK(dalloc(R.allocs, &R.in_f32, (size_t)c->plan.w * c->plan.h));  // akz_extract_f32 is single-image
    R.alloc_n = n;
    HostStats& H = c->hs;
    if (n > H.alloc_n) {
        if (H.base) cudaFreeHost(H.base);
        H = HostStats();
        const size_t bytes = nb * (kMaxLevels * sizeof(double) + 4 * sizeof(unsigned int));
        CK(cudaHostAlloc(&H.base, bytes, cudaHostAllocDefault));
        H.kcontrast = (double*)H.base;
        H.n_kp = (unsigned int*)(H.kcontrast + nb * kMaxLevels);
        H.n_cache = H.n_kp + nb;
        H.n_cand = H.n_cache + nb;
        H.err = H.n_cand + nb;
        H.alloc_n = n;
    }
    return AKZ_OK;
}

// images per sub-batch for a call of n images
static uint32_t sub_batch_of(const akz_context* c, uint32_t n) {
    if (c->flags & AKZ_KEEP_EVOLUTIONS) return n;  // evolutions of every image of the call stay resident
    return std::min(n, c->sub_batch);
}

// The sub-batches of a call of n images. Device-resident inputs: equal sub-batches of `sub_batch` images. Host inputs:
// the first sub-batch's upload and the last one's result download cannot hide behind anything, so the schedule ramps
// up (m/4, m/2, m, m, ...) and ends on a short sub-batch; uploads run at about twice the extraction rate at 1080p, which
// is what a doubling ramp needs to keep the kernels fed.
static void make_schedule(akz_context* c, uint32_t n, bool host_io) {
    c->sched.clear();
    const uint32_t m = sub_batch_of(c, n);
    static const bool no_ramp = getenv("AKZ_NO_RAMP") != nullptr;  // A/B switch
    // a sub-batch must keep the stencil stream busy for as long as the cache pass of its predecessor runs (the pass is
    // latency-bound: about as long for 8 images as for 256), which takes ~16 images at any size
    constexpr uint32_t kMinSub = 16;
    // (device-resident calls that fit one sub-batch stay one sub-batch: splitting 32 x 3840x2160 into two halves to hide
    // the first cache pass was measured slower, 1 202 -> 1 032 images/s -- 16-image launches underfill the small octaves)
    if (!host_io || no_ramp || (c->flags & AKZ_KEEP_EVOLUTIONS) || n < 2 * kMinSub) {
        for (uint32_t i0 = 0; i0 < n; i0 += m) c->sched.push_back({i0, std::min(m, n - i0)});
        return;
    }
    if (n < 4 * kMinSub) {  // two halves: the second upload and the first download overlap the kernels
        const uint32_t h = std::min(m, (n + 1) / 2);
        for (uint32_t i0 = 0; i0 < n; i0 += h) c->sched.push_back({i0, std::min(h, n - i0)});
        return;
    }
    const uint32_t tail = std::max<uint32_t>(kMinSub, std::min<uint32_t>(m / 4, n / 8));
    uint32_t i0 = 0, step = tail;
    while (i0 < n - tail) {
        const uint32_t cnt = std::min(step, n - tail - i0);
        c->sched.push_back({i0, cnt});
        i0 += cnt;
        step = std::min<uint32_t>(m, step * 2);
    }
    c->sched.push_back({i0, n - i0});
}

static bool same_cfg(const akz_config& a, const akz_config& b) { return memcmp(&a, &b, sizeof(akz_config)) == 0; }


static int prepare(akz_context* c, uint32_t n, uint32_t w, uint32_t h, const akz_config* cfg, bool host_io) {
    if (!c || !cfg) return fail(AKZ_ERR_INVALID, "null argument");
    if (n == 0) return fail(AKZ_ERR_INVALID, "empty batch");
    if (n > c->max_batch) return fail(AKZ_ERR_CAPACITY, "batch larger than the context's max_batch");
    if (w > c->max_w || h > c->max_h) return fail(AKZ_ERR_CAPACITY, "image larger than the context's max size");
    CK(cudaSetDevice(c->device));
    if (!c->have_plan || c->plan.w != w || c->plan.h != h || !same_cfg(c->plan.cfg, *cfg)) {
        akz_config z;
        memset(&z, 0, sizeof(z));  // normalise struct padding before memcmp-style comparison
        z.num_sublevels = cfg->num_sublevels;
        z.max_octave_evolution = cfg->max_octave_evolution;
        z.base_scale_offset = cfg->base_scale_offset;
        z.initial_contrast = cfg->initial_contrast;
        z.contrast_percentile = cfg->contrast_percentile;
        z.contrast_factor_num_bins = cfg->contrast_factor_num_bins;
        z.derivative_factor = cfg->derivative_factor;
        z.detector_threshold = cfg->detector_threshold;
        z.descriptor_channels = cfg->descriptor_channels;
        z.descriptor_pattern_size = cfg->descriptor_pattern_size;
        Plan np;
        std::string err = build_plan(w, h, z, &np);
        if (!err.empty()) return fail(AKZ_ERR_INVALID, err);
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaStreamSynchronize(c->stream_kp));
        free_buffers(c);
        c->plan = np;
        c->have_plan = true;
        if (c->caps_auto) {
            // the reference's Vec<Keypoint> has no capacity; the defaults follow the image size (one keypoint per 16
            // pixels, four candidates per keypoint) so that a raw 3840x2160 frame works without akz_context_set_limits
            const uint64_t px = (uint64_t)w * h;
            c->kp_cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(65536, (px / 16 + 4095) & ~4095ull), 1u << 22);
            c->cand_cap = 4 * c->kp_cap;
        }
        if (c->sub_batch_auto) {
            // Large sub-batches fill the GPU in the small octaves, give the latency-bound cache pass one warp per
            // image to hide behind and shorten the pipeline tail: up to 256 images per lane, as long as the two
            // lanes take at most half of the memory that is free right now (B200: 180 GB).
            size_t free_b = 0, total_b = 0;
            CK(cudaMemGetInfo(&free_b, &total_b));
            const size_t n0 = (size_t)w * h;
            const size_t per_image = 4 * (4 * (size_t)np.dev.plane_px + 4 * n0) + 4 * (size_t)np.dev.mask_words + 4 * (size_t)c->cand_cap +
                                     40 * (size_t)c->kp_cap + dedup_pool_bytes(np) + dedup_level_pool_bytes(np, c->cand_cap) + (1 << 16);
            const size_t fit = (free_b / 2) / (2 * per_image);
            c->sub_batch = (uint32_t)std::max<size_t>(16, std::min<size_t>(256, fit));
        }
    }
    make_schedule(c, n, host_io);
    uint32_t m = 0;  // the largest sub-batch of the call: what a lane of work buffers must hold
    for (const auto& sb : c->sched) m = std::max(m, sb.second);
    int rc = ensure_lane(c, c->lane[0], (int)m);
    if (rc != AKZ_OK) return rc;
    if (c->sched.size() > 1) {
        rc = ensure_lane(c, c->lane[1], (int)m);
        if (rc != AKZ_OK) return rc;
    }
    return ensure_results(c, (int)n);
}

// brackets one stage's launches with events when timing is on
struct StageTimer {
    akz_context* c;
    cudaStream_t st;
    size_t slot = (size_t)-1;
    StageTimer(akz_context* ctx, int stage, cudaStream_t stream) : c(ctx), st(stream) {
        if (!c->timing) return;
        if (c->ev_used == c->ev_pool.size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            c->ev_pool.push_back({a, b});
            c->ev_stage.push_back(0);
            c->ev_launches.push_back(0);
        }
        slot = c->ev_used++;
        c->ev_stage[slot] = stage;
        cudaEventRecord(c->ev_pool[slot].first, st);
    }
    void done(int launches) {
        if (slot == (size_t)-1) return;
        c->ev_launches[slot] = launches;
        cudaEventRecord(c->ev_pool[slot].second, st);
    }
};

// folds finished event pairs into the per-stage totals (streams must be idle)
static void harvest_timing(akz_context* c) {
    for (size_t i = 0; i < c->ev_used; i++) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, c->ev_pool[i].first, c->ev_pool[i].second) == cudaSuccess) {
            c->stage_ms[c->ev_stage[i]] += (double)ms;
            c->stage_launches[c->ev_stage[i]] += (uint64_t)c->ev_launches[i];
        }
    }
    c->ev_used = 0;
}

// runs the whole pipeline on inputs already in device memory; results stay on the device (c->res).
// On return everything has been ISSUED and c->stream waits for all of it.
// `upload`, when given, enqueues the host -> device copy of one sub-batch on the copy stream and records ev_copy[sb]; it is
// called one sub-batch ahead of the kernels that consume it.
static int issue_pipeline(akz_context* c, uint32_t n, const void* d_in, bool is_u8, size_t in_stride,
                          const std::function<int(uint32_t)>& upload, bool capturing, int* launches) {
    const bool wait_copies = (bool)upload;
    const Plan& P = c->plan;
    const Results& R = c->res;
    const size_t in_img = in_stride * P.h * (is_u8 ? 1 : sizeof(float));  // bytes per input image
    int k = 0, j;
#define STAGE(st, strm, expr)          \
    {                                  \
        StageTimer t__(c, st, strm);   \
        j = (expr);                    \
        t__.done(j);                   \
        k += j;                        \
    }
    // Stage B of a sub-batch is split: only its latency-bound cache pass runs on the side stream (one warp per image,
    // hidden behind whatever the main stream does next); filter/refine, orientation and descriptors follow on the MAIN
    // stream after the next sub-batch's stencil work. Running those throughput kernels concurrently with the stencil
    // kernels was measured 4 % slower than not overlapping anything (5144 vs 5328 images/s, 1024 x 1080p): they only
    // take SMs, shared memory and L1 from each other. Two fully independent in-order lanes (one stream each) lose too:
    // 5866 vs 5992 images/s.
    struct Pending {
        Lane* ln;
        Buffers B;
        uint32_t i0, cnt, sb;
    };
    Pending prev{};
    bool have_prev = false;
    auto finish_sub_batch = [&](const Pending& pd) -> int {
        const Buffers& B = pd.B;
        CK(cudaStreamWaitEvent(c->stream, pd.ln->ev_dedup, 0));
        Launch LB{c->stream, (int)pd.cnt, c->cand_cap, c->kp_cap};
        STAGE(AKZ_STAGE_FINALIZE, c->stream, launch_finalize(LB, P, B));
        STAGE(AKZ_STAGE_DESCRIPTOR, c->stream, launch_descriptors(LB, P, B));
        CK(cudaEventRecord(pd.ln->ev_done, c->stream));
        // mirror this sub-batch's statistics into pinned memory (side stream: tiny copies must not stall the kernels);
        // the host reads them after ev_stats[sb]
        CK(cudaStreamWaitEvent(c->stream_kp, pd.ln->ev_done, 0));
        const HostStats& H = c->hs;
        const uint32_t i0 = pd.i0, cnt = pd.cnt;
        CK(cudaMemcpyAsync(H.n_kp + i0, B.n_kp, cnt * sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream_kp));
        CK(cudaMemcpyAsync(H.n_cache + i0, B.n_cache, cnt * sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream_kp));
        CK(cudaMemcpyAsync(H.n_cand + i0, B.n_cand_total, cnt * sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream_kp));
        CK(cudaMemcpyAsync(H.err + i0, B.err_flags, cnt * sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream_kp));
        CK(cudaMemcpyAsync(H.kcontrast + (size_t)i0 * kMaxLevels, B.kcontrast, (size_t)cnt * kMaxLevels * sizeof(double),
                           cudaMemcpyDeviceToHost, c->stream_kp));
        while (c->ev_stats.size() <= pd.sb) {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            c->ev_stats.push_back(e);
        }
        CK(cudaEventRecord(c->ev_stats[pd.sb], c->stream_kp));
        return AKZ_OK;
    };
    if (upload) {
        int rc = upload(0);
        if (rc != AKZ_OK) return rc;
    }
    for (uint32_t sb = 0; sb < (uint32_t)c->sched.size(); sb++) {
        Lane& ln = c->lane[sb & 1];
        const uint32_t i0 = c->sched[sb].first, cnt = c->sched[sb].second;
        if (upload && sb + 1 < (uint32_t)c->sched.size()) {
            int rc = upload(sb + 1);
            if (rc != AKZ_OK) return rc;
        }
        if (ln.busy) CK(cudaStreamWaitEvent(c->stream, ln.ev_done, 0));  // the lane's previous sub-batch must be through stage B
        if (wait_copies) CK(cudaStreamWaitEvent(c->stream, c->ev_copy[sb], 0));  // its inputs must have arrived
        Buffers B = ln.buf;
        B.kps = R.kps + (size_t)i0 * c->kp_cap;
        B.desc = R.desc + (size_t)i0 * c->kp_cap * kDescStride;
        B.n_kp = R.n_kp + i0;
        B.n_cache = R.n_cache + i0;
        B.n_cand_total = R.n_cand + i0;
        B.err_flags = R.err + i0;
        B.kcontrast = R.kcontrast + (size_t)i0 * kMaxLevels;
        const void* in = (const uint8_t*)d_in + (size_t)i0 * in_img;
        Launch LA{c->stream, (int)cnt, c->cand_cap, c->kp_cap};
        CK(cudaMemsetAsync(B.mask, 0, (size_t)P.dev.mask_words * cnt * sizeof(unsigned int), c->stream));
        CK(cudaMemsetAsync(B.err_flags, 0, cnt * sizeof(unsigned int), c->stream));
        // The detectors run on their own stream: detector(l) reads Lsmooth_l only, so it overlaps FED(l) and prep(l+1) and
        // fills the SMs the other kernel's last wave leaves idle (every streaming kernel takes a whole register file per
        // SM, so they only ever meet in those tails; the small octaves are one to two waves long). Lsmooth alternates
        // between two scratch planes, prep(l) waits for detector(l-2). Level 1's detector overwrites the Lx/Ly planes
        // that hold the contrast pass's gradients, so it waits for their consumers (FED(1)). Timing mode keeps one stream.
        static const bool det_inline = getenv("AKZ_DET_INLINE") != nullptr;  // A/B switch: detectors on the main stream
        const bool split = !c->timing && !det_inline && !B.keep;
        cudaStream_t sd = split ? c->stream_det : c->stream;
        Launch LD{sd, (int)cnt, c->cand_cap, c->kp_cap};
        STAGE(AKZ_STAGE_LEVEL0, c->stream, launch_level0(LA, P, B, in, is_u8, in_stride));
        if (split) {
            CK(cudaEventRecord(ln.ev_prep[0], c->stream));
            CK(cudaStreamWaitEvent(sd, ln.ev_prep[0], 0));
        }
        STAGE(AKZ_STAGE_CONTRAST, c->stream, launch_contrast(LA, P, B));
        STAGE(AKZ_STAGE_DETECTOR, sd, launch_detector(LD, P, B, 0));
        if (split) CK(cudaEventRecord(ln.ev_det[0], sd));
        // A sub-batch with nothing behind it (a single image, the last of a call) cannot hide its cache pass behind the next
        // one's stencil kernels. It runs the pass in two parts instead: the first octave's levels (most of the candidates) as
        // soon as their detectors are through, on the side stream, next to the stencil kernels of the other octaves.
        const bool last = (sb + 1 == (uint32_t)c->sched.size());
        const int pass_cut = (split && last) ? dedup_split_level(P) : 0;
        auto first_part = [&](int l) -> int {
            if (l != pass_cut - 1) return AKZ_OK;
            CK(cudaStreamWaitEvent(c->stream_kp, ln.ev_det[l], 0));
            Launch LK{c->stream_kp, (int)cnt, c->cand_cap, c->kp_cap};
            k += launch_compact(LK, P, B, 0, pass_cut);
            CK(cudaEventRecord(ln.ev_compact, c->stream_kp));
            k += launch_dedup(LK, P, B, 0, pass_cut);
            return AKZ_OK;
        };
        if (pass_cut) {
            int rc = first_part(0);
            if (rc != AKZ_OK) return rc;
        }
        for (int l = 1; l < P.dev.n_levels; l++) {
            if (split && l >= 2) CK(cudaStreamWaitEvent(c->stream, ln.ev_det[l - 2], 0));  // its Lsmooth scratch plane is free again
            STAGE(AKZ_STAGE_PREP, c->stream, launch_prep(LA, P, B, l));
            if (split && l >= 2) CK(cudaEventRecord(ln.ev_prep[l], c->stream));
            STAGE(AKZ_STAGE_FED, c->stream, launch_fed(LA, P, B, l));
            if (split) {
                if (l == 1) CK(cudaEventRecord(ln.ev_prep[l], c->stream));
                CK(cudaStreamWaitEvent(sd, ln.ev_prep[l], 0));
            }
            STAGE(AKZ_STAGE_DETECTOR, sd, launch_detector(LD, P, B, l));
            if (split) CK(cudaEventRecord(ln.ev_det[l], sd));
            if (pass_cut) {
                int rc = first_part(l);
                if (rc != AKZ_OK) return rc;
            }
        }
        if (split) CK(cudaStreamWaitEvent(c->stream, ln.ev_det[P.dev.n_levels - 1], 0));
        if (pass_cut) CK(cudaStreamWaitEvent(c->stream, ln.ev_compact, 0));  // the lists of the other levels go behind the first part's
        STAGE(AKZ_STAGE_COMPACT, c->stream, launch_compact(LA, P, B, pass_cut, P.dev.n_levels));
        CK(cudaEventRecord(ln.ev_stencil, c->stream));
        // Order on the main stream: S(0) S(1) F(0) S(2) F(1) ... ; the cache pass D(i) starts on the side stream once
        // both S(i) and F(i-1) are through, so that it runs next to the stencil kernels of S(i+1) (one-warp and
        // four-warp blocks that still fit beside it) and not next to the descriptor kernel, whose four 256-thread blocks
        // need the whole register file of an SM: with a cache-pass warp resident only three fit (measured: descriptors
        // +22 %, filter/orientation +20 %). The last sub-batch has nothing behind it, so its cache pass goes first.
        auto launch_cache_pass = [&]() -> int {
            CK(cudaStreamWaitEvent(c->stream_kp, ln.ev_stencil, 0));
            if (have_prev && !last) CK(cudaStreamWaitEvent(c->stream_kp, prev.ln->ev_done, 0));
            Launch LB{c->stream_kp, (int)cnt, c->cand_cap, c->kp_cap};
            STAGE(AKZ_STAGE_DEDUP, c->stream_kp, launch_dedup(LB, P, B, pass_cut, P.dev.n_levels));
            CK(cudaEventRecord(ln.ev_dedup, c->stream_kp));
            // timing mode (and the A/B switch) runs the cache pass alone, so that every stage's event pair brackets its kernels only
            static const bool serial_lanes = getenv("AKZ_SERIAL_LANES") != nullptr;
            static const bool timing_overlap = getenv("AKZ_TIMING_OVERLAP") != nullptr;  // time the stages as they really overlap
            if ((c->timing && !timing_overlap) || serial_lanes) CK(cudaStreamWaitEvent(c->stream, ln.ev_dedup, 0));
            return AKZ_OK;
        };
        if (last) {
            int rc = launch_cache_pass();
            if (rc != AKZ_OK) return rc;
        }
        if (have_prev) {
            int rc = finish_sub_batch(prev);
            if (rc != AKZ_OK) return rc;
        }
        if (!last) {
            int rc = launch_cache_pass();
            if (rc != AKZ_OK) return rc;
        }
        ln.busy = true;
        prev = Pending{&ln, B, i0, cnt, sb};
        have_prev = true;
    }
    if (have_prev) {
        int rc = finish_sub_batch(prev);
        if (rc != AKZ_OK) return rc;
    }
#undef STAGE
    for (int l = 0; l < 2; l++)
        if (c->lane[l].busy) {
            CK(cudaStreamWaitEvent(c->stream, c->lane[l].ev_done, 0));
            c->lane[l].busy = false;
        }
    // a capture ends on c->stream with every forked stream joined: the statistics copies are the last thing on the side stream
    if (capturing) CK(cudaStreamWaitEvent(c->stream, c->ev_stats[c->sched.size() - 1], 0));
    *launches = k;
    CK(cudaGetLastError());
    return AKZ_OK;
}

static void drop_graphs(akz_context* c) {
    for (auto& g : c->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    c->graphs.clear();
}

static int run_pipeline(akz_context* c, uint32_t n, const void* d_in, bool is_u8, size_t in_stride,
                        const std::function<int(uint32_t)>& upload = nullptr) {
    static const bool no_graph = getenv("AKZ_NO_GRAPH") != nullptr;  // A/B switch: always issue the launches one by one
    c->generation++;
    c->cur_batch = (int)n;
    c->n_sub_batches = (uint32_t)c->sched.size();
    int k = 0;
    const bool graph_ok = !no_graph && !c->timing && c->sched.size() == 1 && !(c->flags & AKZ_KEEP_EVOLUTIONS);
    if (!graph_ok) {
        int rc = issue_pipeline(c, n, d_in, is_u8, in_stride, upload, false, &k);
        c->launches += (uint64_t)k;
        return rc;
    }
    if (upload) {  // the inputs travel outside the graph, on the copy stream as always
        int rc = upload(0);
        if (rc != AKZ_OK) return rc;
        CK(cudaStreamWaitEvent(c->stream, c->ev_copy[0], 0));
    }
    akz_context::GraphKey key;
    key.alloc_epoch = g_alloc_epoch.load();
    key.in = d_in;
    key.stats = c->hs.n_kp;
    key.stride = in_stride;
    key.n = n;
    key.cand_cap = c->cand_cap;
    key.kp_cap = c->kp_cap;
    key.is_u8 = is_u8;
    constexpr size_t kMaxGraphs = 4;
    akz_context::CachedGraph* hit = nullptr;
    for (auto& g : c->graphs)
        if (g.key == key) hit = &g;
    if (!hit) {
        if (c->graphs.size() && !(c->graphs[0].key.alloc_epoch == key.alloc_epoch)) drop_graphs(c);  // buffers moved: all are stale
        if (c->graphs.size() >= kMaxGraphs) {
            size_t lru = 0;
            for (size_t i = 1; i < c->graphs.size(); i++)
                if (c->graphs[i].last_use < c->graphs[lru].last_use) lru = i;
            cudaGraphExecDestroy(c->graphs[lru].exec);
            c->graphs.erase(c->graphs.begin() + lru);
        }
        CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
        int rc = issue_pipeline(c, n, d_in, is_u8, in_stride, nullptr, true, &k);
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamEndCapture(c->stream, &g);
        if (rc != AKZ_OK) {
            if (g) cudaGraphDestroy(g);
            return rc;
        }
        CK(e);
        akz_context::CachedGraph cg;
        e = cudaGraphInstantiate(&cg.exec, g, 0);
        cudaGraphDestroy(g);
        CK(e);
        cg.key = key;
        cg.launches = k;
        c->graphs.push_back(cg);
        hit = &c->graphs.back();
    }
    hit->last_use = ++c->graph_replays;
    CK(cudaGraphLaunch(hit->exec, c->stream));
    CK(cudaEventRecord(c->ev_stats[0], c->stream));  // what the host waits for before it reads the statistics
    c->launches += (uint64_t)hit->launches;
    return AKZ_OK;
}

static int check_err_flags(const akz_context* c, uint32_t i0, uint32_t i1) {
    for (uint32_t i = i0; i < i1; i++) {
        const unsigned int e = c->hs.err[i];
        if (e & kErrCandOverflow) return fail(AKZ_ERR_CAPACITY, "candidate list overflow (raise max_candidates)");
        if (e & kErrKpOverflow) return fail(AKZ_ERR_CAPACITY, "keypoint cache overflow (raise max_keypoints)");
        if (e & kErrBounds) return fail(AKZ_ERR_BOUNDS, "a descriptor/orientation sample fell outside the image (the reference panics here)");
    }
    return AKZ_OK;
}

// waits for the whole call; the per-image statistics are then valid in c->hs
static int finish_device_call(akz_context* c, uint32_t n) {
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamSynchronize(c->stream_kp));
    if (c->timing) harvest_timing(c);
    return check_err_flags(c, 0, n);
}

// a pinned slab of at least `bytes` from the pool (slabs whose features have all been freed are reused)
static int get_chunk(akz_context* c, size_t bytes, std::shared_ptr<PinnedChunk>* out) {
    std::shared_ptr<PinnedChunk> best;
    for (auto& ch : c->pinned_pool)
        if (ch.use_count() == 1 && ch->cap >= bytes && (!best || ch->cap < best->cap)) best = ch;
    if (!best) {
        // drop idle slabs that were too small before growing the pool
        std::vector<std::shared_ptr<PinnedChunk>> keep;
        for (auto& ch : c->pinned_pool)
            if (ch.use_count() > 1 || ch->cap >= bytes) keep.push_back(ch);
        c->pinned_pool.swap(keep);
        best = std::make_shared<PinnedChunk>();
        const size_t cap = std::max<size_t>(bytes + bytes / 4, 1 << 16);
        cudaError_t e = cudaHostAlloc(&best->p, cap, cudaHostAllocDefault);
        if (e != cudaSuccess) return fail(AKZ_ERR_NOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
        best->cap = cap;
        c->pinned_pool.push_back(best);
    }
    *out = best;
    return AKZ_OK;
}

// Host side of a host-buffer call, sub-batch by sub-batch while later sub-batches are still running:
// wait for the sub-batch's statistics, then pull exactly n_kp keypoints + descriptors per image into one
// pinned slab on the D2H stream. Handles are cut from the slabs at the end.
static int collect_features(akz_context* c, uint32_t n, akz_features** outs) {
    const Results& R = c->res;
    const HostStats& H = c->hs;
    struct Slot {
        std::shared_ptr<PinnedChunk> chunk;
        size_t kp_off, desc_off;
    };
    std::vector<Slot> slots(n);
    int rc = AKZ_OK;
    for (uint32_t sb = 0; sb < (uint32_t)c->sched.size() && rc == AKZ_OK; sb++) {
        const uint32_t i0 = c->sched[sb].first, i1 = i0 + c->sched[sb].second;
        CK(cudaEventSynchronize(c->ev_stats[sb]));
        rc = check_err_flags(c, i0, i1);
        if (rc != AKZ_OK) break;
        size_t total = 0;
        for (uint32_t i = i0; i < i1; i++) total += H.n_kp[i];
        std::shared_ptr<PinnedChunk> chunk;
        rc = get_chunk(c, total * (sizeof(akz_keypoint) + kDescStride) + 64, &chunk);
        if (rc != AKZ_OK) break;
        // descriptors first (64-byte rows stay 64-byte aligned), keypoints behind them
        size_t doff = 0, koff = total * kDescStride;
        CK(cudaStreamWaitEvent(c->stream_d2h, c->ev_stats[sb], 0));
        for (uint32_t i = i0; i < i1; i++) {
            const size_t nk = H.n_kp[i];
            slots[i] = Slot{chunk, koff, doff};
            if (nk) {
                CK(cudaMemcpyAsync((char*)chunk->p + doff, R.desc + (size_t)i * c->kp_cap * kDescStride, nk * kDescStride,
                                   cudaMemcpyDeviceToHost, c->stream_d2h));
                CK(cudaMemcpyAsync((char*)chunk->p + koff, R.kps + (size_t)i * c->kp_cap, nk * sizeof(akz_keypoint),
                                   cudaMemcpyDeviceToHost, c->stream_d2h));
            }
            doff += nk * kDescStride;
            koff += nk * sizeof(akz_keypoint);
        }
    }
    // the call is over (also on error paths): nothing of it may still be in flight when we return
    cudaError_t e1 = cudaStreamSynchronize(c->stream);
    cudaError_t e2 = cudaStreamSynchronize(c->stream_kp);
    cudaError_t e3 = cudaStreamSynchronize(c->stream_d2h);
    if (c->timing) harvest_timing(c);
    if (rc != AKZ_OK) return rc;
    CK(e1);
    CK(e2);
    CK(e3);
    auto levels = std::make_shared<const std::vector<LevelHost>>(c->plan.host);
    std::vector<std::unique_ptr<akz_features>> fs;
    uint32_t sb_of = 0;
    for (uint32_t i = 0; i < n; i++) {
        while (i >= c->sched[sb_of].first + c->sched[sb_of].second) sb_of++;
        std::unique_ptr<akz_features> f(new akz_features());
        f->ctx = c;
        f->ctx_id = c->id;
        f->generation = c->generation;
        f->sub_batch_index = sb_of;
        f->n_sub_batches = (uint32_t)c->sched.size();
        f->lane = (int)(sb_of & 1);
        f->img = (int)(i - c->sched[sb_of].first);
        f->batch = (int)c->sched[sb_of].second;
        f->levels = levels;
        f->desc_len = (uint32_t)c->plan.dev.desc_len;
        f->contrast = H.kcontrast[(size_t)i * kMaxLevels];
        f->n_cand = H.n_cand[i];
        f->n_cache = H.n_cache[i];
        f->n = H.n_kp[i];
        f->chunk = slots[i].chunk;
        f->kps = (const akz_keypoint*)((const char*)slots[i].chunk->p + slots[i].kp_off);
        f->desc = (const uint8_t*)slots[i].chunk->p + slots[i].desc_off;
        fs.push_back(std::move(f));
    }
    for (uint32_t i = 0; i < n; i++) outs[i] = fs[i].release();
    return AKZ_OK;
}

#define LOCK(c) std::lock_guard<std::recursive_mutex> lk__((c)->mu)

extern "C" {

const char* akz_last_error(void) { return g_last_error.c_str(); }
const char* akz_version(void) { return "akaze_b200 0.1 (sm_100a)"; }

int akz_default_config(akz_config* cfg) {
    if (!cfg) return fail(AKZ_ERR_INVALID, "null config");
    memset(cfg, 0, sizeof(*cfg));
    cfg->num_sublevels = 4;
    cfg->max_octave_evolution = 4;
    cfg->base_scale_offset = 1.6;
    cfg->initial_contrast = 0.001;
    cfg->contrast_percentile = 0.7;
    cfg->contrast_factor_num_bins = 300;
    cfg->derivative_factor = 1.5;
    cfg->detector_threshold = 0.001;
    cfg->descriptor_channels = 3;
    cfg->descriptor_pattern_size = 10;
    return AKZ_OK;
}

int akz_create(int device, uint32_t max_width, uint32_t max_height, uint32_t max_batch, uint32_t flags, akz_context** out) {
    if (!out) return fail(AKZ_ERR_INVALID, "null out");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return fail(AKZ_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
    if (device < 0 || device >= count) return fail(AKZ_ERR_INVALID, "bad device index");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(AKZ_ERR_CUDA, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + ", the kernels are built for sm_100a only");
    if (max_batch == 0 || max_width < 64 || max_height < 32) return fail(AKZ_ERR_INVALID, "max_batch >= 1, max size >= 64x32 required");
    std::unique_ptr<akz_context> c(new akz_context());
    c->device = device;
    c->max_w = max_width;
    c->max_h = max_height;
    c->max_batch = max_batch;
    c->flags = flags;
    if (const char* sbe = getenv("AKZ_SUB_BATCH")) {
        const int v = atoi(sbe);
        if (v > 0) {
            c->sub_batch = (uint32_t)v;
            c->sub_batch_auto = false;
        }
    }
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->stream_kp, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->stream_det, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->stream_copy, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->stream_d2h, cudaStreamNonBlocking));
    CK(init_detector_attributes());
    CK(init_scale_space_attributes());
    CK(init_keypoint_attributes());
    CK(init_matcher_tc_attributes());
    c->id = g_next_ctx_id.fetch_add(1);
    {
        std::lock_guard<std::mutex> g(g_registry_mu);
        g_registry[c->id] = c.get();
    }
    *out = c.release();
    return AKZ_OK;
}

void akz_destroy(akz_context* c) {
    if (!c) return;
    {
        std::lock_guard<std::mutex> g(g_registry_mu);
        g_registry.erase(c->id);
    }
    { std::lock_guard<std::recursive_mutex> lk(c->mu); }  // let a call in flight on another thread finish
    cudaSetDevice(c->device);
    if (c->comm) nccl_dyn::api().CommDestroy(c->comm);
    cudaFree(c->m_gather);
    cudaFree(c->m_stage);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->stream_kp);
    free_buffers(c);
    for (int l = 0; l < 2; l++) {
        if (c->lane[l].ev_stencil) cudaEventDestroy(c->lane[l].ev_stencil);
        if (c->lane[l].ev_done) cudaEventDestroy(c->lane[l].ev_done);
        if (c->lane[l].ev_dedup) cudaEventDestroy(c->lane[l].ev_dedup);
        if (c->lane[l].ev_compact) cudaEventDestroy(c->lane[l].ev_compact);
        for (int i = 0; i < kMaxLevels; i++) {
            if (c->lane[l].ev_prep[i]) cudaEventDestroy(c->lane[l].ev_prep[i]);
            if (c->lane[l].ev_det[i]) cudaEventDestroy(c->lane[l].ev_det[i]);
        }
    }
    drop_graphs(c);
    cudaStreamDestroy(c->stream_kp);
    cudaStreamSynchronize(c->stream_det);
    cudaStreamDestroy(c->stream_det);
    for (cudaEvent_t e : c->ev_copy) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_stats) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream_copy);
    cudaStreamSynchronize(c->stream_d2h);
    cudaStreamDestroy(c->stream_d2h);
    if (c->hs.base) cudaFreeHost(c->hs.base);
    c->pinned_pool.clear();  // slabs still referenced by live akz_features are freed with the last of them
    cudaFree(c->m_q);
    cudaFree(c->m_db);
    cudaFree(c->m_parts);
    cudaFree(c->m_out);
    cudaFree(c->m_qimg);
    cudaFree(c->m_dbimg);
    for (auto& e : c->ev_pool) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    cudaStreamDestroy(c->stream);
    delete c;
}

void* akz_context_stream(akz_context* c) { return c ? (void*)c->stream : nullptr; }
uint64_t akz_context_launch_count(const akz_context* c) { return c ? c->launches : 0; }

int akz_context_enable_timing(akz_context* c, int enable) {
    if (!c) return fail(AKZ_ERR_INVALID, "null context");
    LOCK(c);
    c->timing = enable != 0;
    return AKZ_OK;
}

int akz_context_stage_times(akz_context* c, double* ms, uint64_t* launches, int reset) {
    if (!c) return fail(AKZ_ERR_INVALID, "null context");
    LOCK(c);
    for (int i = 0; i < AKZ_NUM_STAGES; i++) {
        if (ms) ms[i] = c->stage_ms[i];
        if (launches) launches[i] = c->stage_launches[i];
        if (reset) {
            c->stage_ms[i] = 0.0;
            c->stage_launches[i] = 0;
        }
    }
    return AKZ_OK;
}

int akz_context_set_limits(akz_context* c, uint32_t max_candidates, uint32_t max_keypoints) {
    if (!c || max_candidates == 0 || max_keypoints == 0) return fail(AKZ_ERR_INVALID, "bad limits");
    if (max_keypoints > 0x7fffffffu) return fail(AKZ_ERR_INVALID, "max_keypoints too large");
    if (max_candidates < max_keypoints) max_candidates = max_keypoints;
    LOCK(c);
    c->cand_cap = max_candidates;
    c->kp_cap = max_keypoints;
    c->caps_auto = false;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->stream_kp);
    free_buffers(c);
    return AKZ_OK;
}

int akz_context_set_match_path(akz_context* c, int path) {
    if (!c || path < AKZ_MATCH_AUTO || path > AKZ_MATCH_TENSOR) return fail(AKZ_ERR_INVALID, "bad match path");
    LOCK(c);
    c->match_path = path;
    return AKZ_OK;
}

int akz_context_set_sub_batch(akz_context* c, uint32_t images) {
    if (!c || images == 0) return fail(AKZ_ERR_INVALID, "bad sub-batch size");
    LOCK(c);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->stream_kp);
    c->sub_batch = images;
    c->sub_batch_auto = false;
    return AKZ_OK;
}

int akz_extract_batch_u8(akz_context* c, uint32_t n, const uint8_t* const* grays, uint32_t w, uint32_t h, size_t stride,
                         const akz_config* cfg, akz_features** outs) {
    if (!c || !grays || !outs) return fail(AKZ_ERR_INVALID, "null argument");
    if (stride < w) return fail(AKZ_ERR_INVALID, "stride < width");
    LOCK(c);
    int rc = prepare(c, n, w, h, cfg, true);
    if (rc != AKZ_OK) return rc;
    // uploads run on their own stream, one event per pipeline sub-batch, so that the copy of sub-batch i+1 overlaps the
    // kernels of sub-batch i (asynchronous when the caller's images are in pinned memory); images that follow each other
    // in host memory without row padding travel as ONE copy per sub-batch
    for (uint32_t i = 0; i < n; i++)
        if (!grays[i]) return fail(AKZ_ERR_INVALID, "null image");
    auto upload = [&](uint32_t sb) -> int {
        const uint32_t i0 = c->sched[sb].first, cnt = c->sched[sb].second;
        const size_t img = (size_t)w * h;
        bool contiguous = stride == w;
        for (uint32_t i = i0 + 1; contiguous && i < i0 + cnt; i++) contiguous = grays[i] == grays[i - 1] + img;
        if (contiguous) {
            CK(cudaMemcpyAsync(c->res.in_u8 + (size_t)i0 * img, grays[i0], (size_t)cnt * img, cudaMemcpyHostToDevice, c->stream_copy));
        } else {
            for (uint32_t i = i0; i < i0 + cnt; i++)
                CK(cudaMemcpy2DAsync(c->res.in_u8 + (size_t)i * img, w, grays[i], stride, w, h, cudaMemcpyHostToDevice, c->stream_copy));
        }
        while (c->ev_copy.size() <= sb) {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            c->ev_copy.push_back(e);
        }
        CK(cudaEventRecord(c->ev_copy[sb], c->stream_copy));
        return AKZ_OK;
    };
    rc = run_pipeline(c, n, c->res.in_u8, true, w, upload);
    if (rc != AKZ_OK) return rc;
    return collect_features(c, n, outs);
}

int akz_extract_u8(akz_context* c, const uint8_t* gray, uint32_t w, uint32_t h, size_t stride, const akz_config* cfg,
                   akz_features** out) {
    const uint8_t* p[1] = {gray};
    return akz_extract_batch_u8(c, 1, p, w, h, stride, cfg, out);
}

int akz_extract_f32(akz_context* c, const float* unit_gray, uint32_t w, uint32_t h, const akz_config* cfg, akz_features** out) {
    if (!c || !unit_gray || !out) return fail(AKZ_ERR_INVALID, "null argument");
    LOCK(c);
    int rc = prepare(c, 1, w, h, cfg, false);
    if (rc != AKZ_OK) return rc;
    CK(cudaMemcpyAsync(c->res.in_f32, unit_gray, (size_t)w * h * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    rc = run_pipeline(c, 1, c->res.in_f32, false, w);
    if (rc != AKZ_OK) return rc;
    return collect_features(c, 1, out);
}

int akz_extract_batch_u8_device(akz_context* c, uint32_t n, const void* d_grays, uint32_t w, uint32_t h, size_t stride,
                                const akz_config* cfg, uint32_t* counts) {
    if (!c || !d_grays || !counts) return fail(AKZ_ERR_INVALID, "null argument");
    if (stride < w) return fail(AKZ_ERR_INVALID, "stride < width");
    LOCK(c);
    int rc = prepare(c, n, w, h, cfg, false);
    if (rc != AKZ_OK) return rc;
    rc = run_pipeline(c, n, d_grays, true, stride);
    if (rc != AKZ_OK) return rc;
    rc = finish_device_call(c, n);
    if (rc != AKZ_OK) return rc;
    for (uint32_t i = 0; i < n; i++) counts[i] = c->hs.n_kp[i];
    return AKZ_OK;
}

int akz_context_device_results(akz_context* c, void** d_keypoints, void** d_descriptors, uint32_t* kp_capacity) {
    if (!c) return fail(AKZ_ERR_INVALID, "null context");
    LOCK(c);
    if (!c->res.kps) return fail(AKZ_ERR_INVALID, "no extraction has run on this context");
    if (d_keypoints) *d_keypoints = c->res.kps;
    if (d_descriptors) *d_descriptors = c->res.desc;
    if (kp_capacity) *kp_capacity = c->kp_cap;
    return AKZ_OK;
}

uint64_t akz_features_count(const akz_features* f) { return f ? f->n : 0; }
const akz_keypoint* akz_features_keypoints(const akz_features* f) { return f ? f->kps : nullptr; }
const uint8_t* akz_features_descriptors(const akz_features* f) { return f ? f->desc : nullptr; }
uint32_t akz_features_descriptor_len(const akz_features* f) { return f ? f->desc_len : 0; }
uint32_t akz_features_num_levels(const akz_features* f) { return f ? (uint32_t)f->levels->size() : 0; }
int akz_features_level_info(const akz_features* f, uint32_t level, akz_level_info* out) {
    if (!f || !out || level >= f->levels->size()) return fail(AKZ_ERR_INVALID, "bad level");
    *out = (*f->levels)[level].info;
    return AKZ_OK;
}
int akz_features_fed_tau(const akz_features* f, uint32_t level, double* out, uint32_t cap) {
    if (!f || level >= f->levels->size() || (!out && cap)) return fail(AKZ_ERR_INVALID, "bad level");
    const std::vector<double>& t = (*f->levels)[level].tau;
    for (uint32_t i = 0; i < cap && i < t.size(); i++) out[i] = t[i];
    return AKZ_OK;
}
double akz_features_contrast_factor(const akz_features* f) { return f ? f->contrast : 0.0; }
uint64_t akz_features_num_candidates(const akz_features* f) { return f ? f->n_cand : 0; }
uint64_t akz_features_num_cache(const akz_features* f) { return f ? f->n_cache : 0; }

int akz_features_evolution_download(const akz_features* f, uint32_t level, int kind, float* dst) {
    if (!f || !dst || level >= f->levels->size()) return fail(AKZ_ERR_INVALID, "bad argument");
    akz_context* c = live_context(f->ctx_id);
    if (!c || c != f->ctx) return fail(AKZ_ERR_INVALID, "the context of these features has been destroyed");
    LOCK(c);
    if (c->generation != f->generation) return fail(AKZ_ERR_INVALID, "evolutions were overwritten by a later extraction (or the context's buffers were resized)");
    const bool keep = (c->flags & AKZ_KEEP_EVOLUTIONS) != 0;
    // Without AKZ_KEEP_EVOLUTIONS only the four planes the keypoint stages sample (Lt, Lx, Ly, Ldet; Lsmooth_0 is Lt_0,
    // lib.rs:58) are persistent, and only for the images of the last sub-batch on each of the two work-buffer lanes.
    if (!keep) {
        const bool plane_kept = kind == AKZ_LT || kind == AKZ_LX || kind == AKZ_LY || kind == AKZ_LDET || (kind == AKZ_LSMOOTH && level == 0);
        if (!plane_kept) return fail(AKZ_ERR_INVALID, "only Lt, Lx, Ly and Ldet stay resident without AKZ_KEEP_EVOLUTIONS");
        if (f->sub_batch_index + 2 < f->n_sub_batches) return fail(AKZ_ERR_INVALID, "the evolutions of this image were overwritten by a later sub-batch of the same call");
    }
    CK(cudaSetDevice(c->device));
    const Buffers& B = c->lane[f->lane].buf;  // keep-evolutions mode runs the whole call as one sub-batch on lane 0
    const float* plane = nullptr;
    switch (kind) {
        case AKZ_LT: plane = B.Lt; break;
        case AKZ_LSMOOTH: plane = (level == 0) ? B.Lt : B.Lsmooth; break;  // lib.rs:58 (levels >= 1: keep mode only, checked above)
        case AKZ_LX: plane = B.Lx; break;
        case AKZ_LY: plane = B.Ly; break;
        case AKZ_LXX: plane = B.Lxx; break;
        case AKZ_LYY: plane = B.Lyy; break;
        case AKZ_LXY: plane = B.Lxy; break;
        case AKZ_LFLOW: plane = B.Lflow; break;
        case AKZ_LSTEP: plane = B.Lstep; break;
        case AKZ_LDET: plane = B.Ldet; break;
        default: return fail(AKZ_ERR_INVALID, "bad image kind");
    }
    if (level == 0 && (kind == AKZ_LFLOW || kind == AKZ_LSTEP)) return fail(AKZ_ERR_INVALID, "level 0 has no Lflow/Lstep (0x0 in the reference)");
    if (!plane) return fail(AKZ_ERR_INVALID, "image plane is not resident");
    const LevelDev& lv = c->plan.dev.lv[level];
    const size_t px = (size_t)lv.w * lv.h;
    const float* src = plane + (size_t)lv.off * f->batch + (size_t)f->img * px;
    CK(cudaMemcpyAsync(dst, src, px * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return AKZ_OK;
}

void akz_features_free(akz_features* f) { delete f; }

// ---- matching ------------------------------------------------------------------------------------
static int grow(void** p, size_t* cap, size_t need) {
    if (need <= *cap) return AKZ_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    CK(cudaMalloc(p, need));
    *cap = need;
    return AKZ_OK;
}

int akz_match_top2_device(akz_context* c, const void* d_q, uint64_t nq, const void* d_db, uint64_t ndb, uint32_t db_index_base,
                          void* d_out) {
    if (!c || (nq && (!d_q || !d_out)) || (ndb && !d_db)) return fail(AKZ_ERR_INVALID, "null argument");
    if (nq == 0) return AKZ_OK;
    LOCK(c);
    if (ndb > 0xffffffffull || nq > 0xffffffffull) return fail(AKZ_ERR_CAPACITY, "more than 2^32 descriptors");
    CK(cudaSetDevice(c->device));
    // tensor-core path (matcher_tc.cu) unless the problem is tiny
    // (measured through the host call, tools/match_latency.py: equal at 32 x 32, tensor 0.06 vs popc 0.10 ms at 128 x 128, 0.14 vs 0.40 ms at
    // the 7 395 x 5 629 descriptors of the reference's test images)
    const bool tensor = ndb > 0 && (c->match_path == AKZ_MATCH_TENSOR || (c->match_path == AKZ_MATCH_AUTO && nq * ndb >= (1ull << 12)));
    const int parts = tensor ? match_tc_parts(nq, ndb) : match_parts(nq, ndb);
    akz_top2* dst = (akz_top2*)d_out;
    if (parts > 1) {
        int rc = grow(&c->m_parts, &c->m_parts_cap, (size_t)parts * nq * sizeof(akz_top2));
        if (rc != AKZ_OK) return rc;
        dst = (akz_top2*)c->m_parts;
    }
    if (tensor) {
        int rc = grow(&c->m_qimg, &c->m_qimg_cap, match_tc_query_image_bytes(nq));
        if (rc != AKZ_OK) return rc;
        rc = grow(&c->m_dbimg, &c->m_dbimg_cap, match_tc_db_image_bytes(ndb));
        if (rc != AKZ_OK) return rc;
        c->launches += launch_match_tc(c->stream, (const uint8_t*)d_q, nq, (const uint8_t*)d_db, ndb, db_index_base, (uint8_t*)c->m_qimg,
                                       (uint8_t*)c->m_dbimg, dst, parts);
    } else {
        c->launches += launch_match_top2(c->stream, (const uint8_t*)d_q, nq, (const uint8_t*)d_db, ndb, db_index_base, dst, parts);
    }
    if (parts > 1) c->launches += launch_merge_top2(c->stream, (const akz_top2*)c->m_parts, (uint32_t)parts, nq, (akz_top2*)d_out);
    CK(cudaGetLastError());
    return AKZ_OK;
}

int akz_merge_top2_device(akz_context* c, const void* d_parts, uint32_t n_parts, uint64_t nq, void* d_out) {
    if (!c || (nq && (!d_parts || !d_out)) || n_parts == 0) return fail(AKZ_ERR_INVALID, "bad argument");
    LOCK(c);
    CK(cudaSetDevice(c->device));
    c->launches += launch_merge_top2(c->stream, (const akz_top2*)d_parts, n_parts, nq, (akz_top2*)d_out);
    CK(cudaGetLastError());
    return AKZ_OK;
}

// host rows (stride, desc_len) -> device rows of 64 bytes, zero padded
static int upload_padded(akz_context* c, const uint8_t* src, uint64_t n, uint32_t desc_len, size_t stride, void** dbuf, size_t* cap) {
    int rc = grow(dbuf, cap, std::max<size_t>((size_t)n * kDescStride, 64));
    if (rc != AKZ_OK) return rc;
    if (n == 0) return AKZ_OK;
    if (stride == (size_t)kDescStride && desc_len == (uint32_t)kDescStride) {
        CK(cudaMemcpyAsync(*dbuf, src, (size_t)n * kDescStride, cudaMemcpyHostToDevice, c->stream));
    } else {
        // One linear copy of the rows as they lie in host memory, repacked on the device. (A 2-D copy of 61-byte rows moves
        // them one DMA descriptor each: 3.7 ms for the 7 395 x 5 629 descriptors of the reference's two test images, against
        // 0.1 ms for the match itself.)
        const size_t bytes = (size_t)(n - 1) * stride + desc_len;
        rc = grow(&c->m_stage, &c->m_stage_cap, bytes);
        if (rc != AKZ_OK) return rc;
        CK(cudaMemcpyAsync(c->m_stage, src, bytes, cudaMemcpyHostToDevice, c->stream));
        c->launches += launch_repack_rows(c->stream, (const uint8_t*)c->m_stage, stride, desc_len, n, (uint8_t*)*dbuf);
    }
    return AKZ_OK;
}

int akz_match_top2(akz_context* c, const uint8_t* q, uint64_t nq, const uint8_t* db, uint64_t ndb, uint32_t desc_len, size_t stride,
                   akz_top2* out) {
    if (!c || (nq && (!q || !out)) || (ndb && !db)) return fail(AKZ_ERR_INVALID, "null argument");
    if (desc_len == 0 || desc_len > (uint32_t)kDescStride || stride < desc_len) return fail(AKZ_ERR_INVALID, "desc_len must be 1..64 and <= stride");
    if (nq == 0) return AKZ_OK;
    LOCK(c);
    CK(cudaSetDevice(c->device));
    int rc = upload_padded(c, q, nq, desc_len, stride, &c->m_q, &c->m_q_cap);
    if (rc != AKZ_OK) return rc;
    rc = upload_padded(c, db, ndb, desc_len, stride, &c->m_db, &c->m_db_cap);
    if (rc != AKZ_OK) return rc;
    rc = grow(&c->m_out, &c->m_out_cap, (size_t)nq * sizeof(akz_top2));
    if (rc != AKZ_OK) return rc;
    rc = akz_match_top2_device(c, c->m_q, nq, c->m_db, ndb, 0, c->m_out);
    if (rc != AKZ_OK) return rc;
    CK(cudaMemcpyAsync(out, c->m_out, (size_t)nq * sizeof(akz_top2), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return AKZ_OK;
}

int akz_descriptor_match(akz_context* c, const uint8_t* d0, uint64_t n0, const uint8_t* d1, uint64_t n1, uint32_t desc_len,
                         size_t stride, uint64_t distance_threshold, double lowes_ratio, akz_match* out, uint64_t* n_out) {
    if (!n_out || (n0 && !out)) return fail(AKZ_ERR_INVALID, "null argument");
    *n_out = 0;
    std::vector<akz_top2> t(n0);
    int rc = akz_match_top2(c, d0, n0, d1, n1, desc_len, stride, t.data());
    if (rc != AKZ_OK) return rc;
    // feature_matching.rs:38-50 seeds min and second-to-min with distance_threshold T: the pair it ends up with is the two
    // smallest values of {d_j} U {T, T}. The device seeds with 10000 (lib.rs:264), far above any Hamming distance of 64
    // bytes, so its (best, second) are the true two smallest distances whenever they exist; re-seeding with T is a min().
    const uint64_t T = distance_threshold;
    const double r2 = lowes_ratio * lowes_ratio;  // powi(2), feature_matching.rs:61
    uint64_t k = 0;
    for (uint64_t i = 0; i < n0; i++) {
        const uint64_t d_best = n1 >= 1 ? t[i].best : UINT64_MAX, d_second = n1 >= 2 ? t[i].second : UINT64_MAX;
        const uint64_t mn = std::min(d_best, T), sc = std::min(d_second, T);
        if ((double)mn < (double)sc * r2) {
            if (mn < T) {  // then mn is a real distance and best_idx the lowest index attaining it
                out[k].index_0 = i;
                out[k].index_1 = t[i].best_idx;
                out[k].distance = (double)mn;
                k++;
            }
        }
    }
    *n_out = k;
    return AKZ_OK;
}

// ---- RANSAC -----------------------------------------------------------------------------
namespace {
struct Xorshift128Plus {  // the `random` crate's default source (xorshift128+ seeded [42, 69]); ^0.12 is not under /root/reference
    uint64_t s0 = 42, s1 = 69;
    uint64_t read_u64() {
        uint64_t x = s0;
        const uint64_t y = s1;
        s0 = y;
        x ^= x << 23;
        x ^= x >> 17;
        x ^= y ^ (y >> 26);
        s1 = x;
        return x + y;
    }
};
}  // namespace

int akz_remove_outliers(akz_context* c, const akz_keypoint* kp0, uint64_t n0, const akz_keypoint* kp1, uint64_t n1, const akz_match* matches,
                        uint64_t n_matches, uint64_t num_trials, float epsilon_model, float epsilon_inlier, int sampling, akz_match* out,
                        uint64_t* n_out, float* model) {
    if (!c || !n_out || (n_matches && (!matches || !out || !kp0 || !kp1))) return fail(AKZ_ERR_INVALID, "null argument");
    if (sampling != AKZ_RANSAC_REFERENCE && sampling != AKZ_RANSAC_ADVANCING) return fail(AKZ_ERR_INVALID, "bad sampling policy");
    *n_out = 0;
    if (model) memset(model, 0, 9 * sizeof(float));
    if (n_matches < 8) {  // estimate_fundamental_matrix.rs:107-110: not enough points, the matches come back untouched
        for (uint64_t i = 0; i < n_matches; i++) out[i] = matches[i];
        *n_out = n_matches;
        return AKZ_OK;
    }
    if (n_matches > 0xffffffffull || num_trials > (1ull << 24)) return fail(AKZ_ERR_CAPACITY, "too many matches or trials");
    std::vector<float2> pl(n_matches), pr(n_matches);
    for (uint64_t i = 0; i < n_matches; i++) {
        if (matches[i].index_0 >= n0 || matches[i].index_1 >= n1) return fail(AKZ_ERR_INVALID, "match index outside the keypoint arrays");
        pl[i] = make_float2(kp0[matches[i].index_0].x, kp0[matches[i].index_0].y);
        pr[i] = make_float2(kp1[matches[i].index_1].x, kp1[matches[i].index_1].y);
    }
    // which eight matches every trial draws (:116-125), ascending inside a trial
    const uint32_t nt = (uint32_t)num_trials;
    std::vector<unsigned int> samples((size_t)nt * 8);
    Xorshift128Plus running;
    for (uint32_t t = 0; t < nt; t++) {
        Xorshift128Plus fresh;
        Xorshift128Plus& src = sampling == AKZ_RANSAC_REFERENCE ? fresh : running;
        unsigned int chosen[8];
        int k = 0;
        while (k < 8) {
            const unsigned int v = (unsigned int)(src.read_u64() % n_matches);
            bool dup = false;
            for (int j = 0; j < k; j++) dup = dup || chosen[j] == v;
            if (!dup) chosen[k++] = v;
        }
        std::sort(chosen, chosen + 8);
        memcpy(&samples[(size_t)t * 8], chosen, sizeof(chosen));
    }
    LOCK(c);
    CK(cudaSetDevice(c->device));
    const size_t b_pts = n_matches * sizeof(float2), b_smp = samples.size() * sizeof(unsigned int);
    const size_t b_models = (size_t)(nt + 1) * 9 * sizeof(float), b_u32 = (size_t)nt * sizeof(unsigned int);
    const size_t need = 2 * b_pts + b_smp + b_models + 2 * b_u32 + n_matches + 256;
    int rc = grow(&c->m_parts, &c->m_parts_cap, need);  // the matcher's scratch: no match job is in flight on a locked context
    if (rc != AKZ_OK) return rc;
    char* base = (char*)c->m_parts;
    float2* d_pl = (float2*)base;
    float2* d_pr = (float2*)(base + b_pts);
    float* d_models = (float*)(base + 2 * b_pts);                      // [nt] + the final model behind them
    unsigned int* d_samples = (unsigned int*)(base + 2 * b_pts + b_models);
    unsigned int* d_ok = (unsigned int*)((char*)d_samples + b_smp);
    unsigned int* d_counts = d_ok + nt;
    unsigned char* d_mask = (unsigned char*)(d_counts + nt);
    CK(cudaMemcpyAsync(d_pl, pl.data(), b_pts, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_pr, pr.data(), b_pts, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_samples, samples.data(), b_smp, cudaMemcpyHostToDevice, c->stream));
    c->launches += launch_ransac_models(c->stream, d_pl, d_pr, d_samples, nt, epsilon_model, d_models, d_ok);
    c->launches += launch_ransac_count(c->stream, d_pl, d_pr, (unsigned int)n_matches, d_models, d_ok, nt, epsilon_inlier, d_counts);
    std::vector<unsigned int> counts(nt), ok(nt);
    CK(cudaMemcpyAsync(counts.data(), d_counts, b_u32, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(ok.data(), d_ok, b_u32, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    // :150-153: the first trial with a strictly larger inlier count wins; without any model final_model stays all zeros
    unsigned int best = 0;
    int best_t = -1;
    for (uint32_t t = 0; t < nt; t++)
        if (ok[t] && counts[t] > best) {
            best = counts[t];
            best_t = (int)t;
        }
    float* d_final = d_models + (size_t)nt * 9;
    float final_model[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (best_t >= 0) CK(cudaMemcpyAsync(final_model, d_models + (size_t)best_t * 9, sizeof(final_model), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpyAsync(d_final, final_model, sizeof(final_model), cudaMemcpyHostToDevice, c->stream));
    c->launches += launch_ransac_mask(c->stream, d_pl, d_pr, (unsigned int)n_matches, d_final, epsilon_inlier, d_mask);
    std::vector<unsigned char> mask(n_matches);
    CK(cudaMemcpyAsync(mask.data(), d_mask, n_matches, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    uint64_t k = 0;
    for (uint64_t i = 0; i < n_matches; i++)
        if (mask[i]) out[k++] = matches[i];
    *n_out = k;
    if (model) memcpy(model, final_model, sizeof(final_model));
    return AKZ_OK;
}

// ---- multi-GPU matching -------------------------------------------------------
// The database is partitioned contiguously by index over the ranks (one context per GPU), the queries are replicated.
// Every rank computes its shard's top-2 records, the 8-byte records are all-gathered with NCCL over NVLink (8 MB per
// rank per 1 M queries) and merged with the sequential scan's tie rule (lowest database index wins), so every rank
// ends up with the result of the unsharded scan (feature_matching.rs:37-50), bit for bit.
#define NCK(call)                                                                                             \
    do {                                                                                                      \
        ncclResult_t r__ = (call);                                                                            \
        if (r__ != 0) return fail(AKZ_ERR_CUDA, std::string(#call) + ": " + nccl_dyn::api().GetErrorString(r__)); \
    } while (0)

static int need_nccl() {
    if (!nccl_dyn::api().ok()) return fail(AKZ_ERR_CUDA, nccl_dyn::api().error);
    return AKZ_OK;
}

int akz_comm_unique_id(uint8_t* id) {
    if (!id) return fail(AKZ_ERR_INVALID, "null id");
    int rc = need_nccl();
    if (rc != AKZ_OK) return rc;
    ncclUniqueId u;
    NCK(nccl_dyn::api().GetUniqueId(&u));
    static_assert(sizeof(u) == AKZ_COMM_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
    memcpy(id, &u, sizeof(u));
    return AKZ_OK;
}

static int drop_comm(akz_context* c) {
    if (c->comm) {
        CK(cudaSetDevice(c->device));
        CK(cudaStreamSynchronize(c->stream));
        nccl_dyn::api().CommDestroy(c->comm);
        c->comm = nullptr;
    }
    c->comm_rank = 0;
    c->comm_size = 1;
    return AKZ_OK;
}

int akz_context_comm_init(akz_context* c, const uint8_t* id, int rank, int n_ranks) {
    if (!c || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(AKZ_ERR_INVALID, "bad communicator arguments");
    int rc = need_nccl();
    if (rc != AKZ_OK) return rc;
    LOCK(c);
    rc = drop_comm(c);
    if (rc != AKZ_OK) return rc;
    CK(cudaSetDevice(c->device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    NCK(nccl_dyn::api().CommInitRank(&c->comm, n_ranks, u, rank));
    c->comm_rank = rank;
    c->comm_size = n_ranks;
    return AKZ_OK;
}

int akz_context_comm_init_all(akz_context* const* ctxs, int n) {
    if (!ctxs || n < 1) return fail(AKZ_ERR_INVALID, "bad communicator arguments");
    int rc = need_nccl();
    if (rc != AKZ_OK) return rc;
    std::vector<int> devs(n);
    for (int i = 0; i < n; i++) {
        if (!ctxs[i]) return fail(AKZ_ERR_INVALID, "null context");
        for (int j = 0; j < i; j++)
            if (ctxs[j]->device == ctxs[i]->device) return fail(AKZ_ERR_INVALID, "one context per device, please");
        rc = drop_comm(ctxs[i]);
        if (rc != AKZ_OK) return rc;
        devs[i] = ctxs[i]->device;
    }
    std::vector<ncclComm_t> comms(n, nullptr);
    NCK(nccl_dyn::api().CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; i++) {
        ctxs[i]->comm = comms[i];
        ctxs[i]->comm_rank = i;
        ctxs[i]->comm_size = n;
    }
    return AKZ_OK;
}

int akz_context_comm_destroy(akz_context* c) {
    if (!c) return fail(AKZ_ERR_INVALID, "null context");
    LOCK(c);
    return drop_comm(c);
}

// the three phases of one rank's share; the single-process driver below interleaves them over its contexts
static int sharded_match_phase(akz_context* c, const void* d_q, uint64_t nq, const void* d_db_shard, uint64_t ndb_shard, uint32_t db_index_base) {
    int rc = grow(&c->m_gather, &c->m_gather_cap, (size_t)(c->comm_size + 1) * nq * sizeof(akz_top2));
    if (rc != AKZ_OK) return rc;
    akz_top2* mine = (akz_top2*)c->m_gather + (size_t)c->comm_size * nq;  // this rank's records, behind the gathered ones
    // an empty shard (more ranks than descriptors) yields the seeds {10000, 10000}: the popc kernel runs zero tiles
    return akz_match_top2_device(c, d_q, nq, d_db_shard, ndb_shard, db_index_base, mine);
}
static int sharded_gather_phase(akz_context* c, uint64_t nq) {
    const akz_top2* mine = (const akz_top2*)c->m_gather + (size_t)c->comm_size * nq;
    NCK(nccl_dyn::api().AllGather(mine, c->m_gather, nq * sizeof(akz_top2), nccl_dyn::kUint8, c->comm, c->stream));
    return AKZ_OK;
}
static int sharded_merge_phase(akz_context* c, uint64_t nq, void* d_out) {
    return akz_merge_top2_device(c, c->m_gather, (uint32_t)c->comm_size, nq, d_out);  // ranks are ordered by database index range
}

int akz_match_top2_sharded_device(akz_context* c, const void* d_q, uint64_t nq, const void* d_db_shard, uint64_t ndb_shard,
                                  uint32_t db_index_base, void* d_out) {
    if (!c || (nq && (!d_q || !d_out)) || (ndb_shard && !d_db_shard)) return fail(AKZ_ERR_INVALID, "null argument");
    LOCK(c);
    if (!c->comm) return fail(AKZ_ERR_INVALID, "no communicator: call akz_context_comm_init first");
    if (nq == 0) return AKZ_OK;
    CK(cudaSetDevice(c->device));
    int rc = sharded_match_phase(c, d_q, nq, d_db_shard, ndb_shard, db_index_base);
    if (rc != AKZ_OK) return rc;
    rc = sharded_gather_phase(c, nq);
    if (rc != AKZ_OK) return rc;
    return sharded_merge_phase(c, nq, d_out);
}

int akz_match_top2_sharded(akz_context* const* ctxs, int n_gpu, const uint8_t* q, uint64_t nq, const uint8_t* db, uint64_t ndb,
                           uint32_t desc_len, size_t stride, akz_top2* out) {
    if (!ctxs || n_gpu < 1 || (nq && (!q || !out)) || (ndb && !db)) return fail(AKZ_ERR_INVALID, "null argument");
    if (desc_len == 0 || desc_len > (uint32_t)kDescStride || stride < desc_len) return fail(AKZ_ERR_INVALID, "desc_len must be 1..64 and <= stride");
    for (int i = 0; i < n_gpu; i++)
        if (!ctxs[i] || !ctxs[i]->comm || ctxs[i]->comm_size != n_gpu || ctxs[i]->comm_rank != i)
            return fail(AKZ_ERR_INVALID, "contexts must share a communicator of n_gpu ranks (akz_context_comm_init_all), in rank order");
    if (nq == 0) return AKZ_OK;
    const uint64_t per = (ndb + (uint64_t)n_gpu - 1) / (uint64_t)n_gpu;  // contiguous shards by index
    std::vector<std::unique_lock<std::recursive_mutex>> locks;
    for (int i = 0; i < n_gpu; i++) locks.emplace_back(ctxs[i]->mu);
    int rc;
    for (int i = 0; i < n_gpu; i++) {  // uploads + shard kernels: asynchronous on every device's own stream
        akz_context* c = ctxs[i];
        const uint64_t lo = std::min(ndb, (uint64_t)i * per), hi = std::min(ndb, lo + per);
        CK(cudaSetDevice(c->device));
        rc = upload_padded(c, q, nq, desc_len, stride, &c->m_q, &c->m_q_cap);
        if (rc != AKZ_OK) return rc;
        rc = upload_padded(c, db + lo * stride, hi - lo, desc_len, stride, &c->m_db, &c->m_db_cap);
        if (rc != AKZ_OK) return rc;
        rc = sharded_match_phase(c, c->m_q, nq, c->m_db, hi - lo, (uint32_t)lo);
        if (rc != AKZ_OK) return rc;
    }
    NCK(nccl_dyn::api().GroupStart());  // one thread drives all ranks: the collective must be issued as a group
    for (int i = 0; i < n_gpu; i++) {
        rc = sharded_gather_phase(ctxs[i], nq);
        if (rc != AKZ_OK) {
            nccl_dyn::api().GroupEnd();
            return rc;
        }
    }
    NCK(nccl_dyn::api().GroupEnd());
    akz_context* c0 = ctxs[0];
    CK(cudaSetDevice(c0->device));
    rc = grow(&c0->m_out, &c0->m_out_cap, (size_t)nq * sizeof(akz_top2));
    if (rc != AKZ_OK) return rc;
    rc = sharded_merge_phase(c0, nq, c0->m_out);
    if (rc != AKZ_OK) return rc;
    CK(cudaMemcpyAsync(out, c0->m_out, (size_t)nq * sizeof(akz_top2), cudaMemcpyDeviceToHost, c0->stream));
    for (int i = 0; i < n_gpu; i++) {
        CK(cudaSetDevice(ctxs[i]->device));
        CK(cudaStreamSynchronize(ctxs[i]->stream));
    }
    return AKZ_OK;
}

}  // extern "C"
