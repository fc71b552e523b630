// common.cuh -- shared declarations of the B200 A-KAZE engine (internal; the public ABI is
// include/akaze_b200.h). All device arithmetic that must match the reference bit for bit is written
// with explicit __fmul_rn/__fadd_rn-equivalent operation order and the library is compiled with
// --fmad=false so that nvcc never contracts a*b+c (rustc does not either).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/akaze_b200.h"

namespace akz {

constexpr int kMaxLevels = 32;
constexpr int kMaxFedSteps = 64;   // per level (default config needs 29)
constexpr int kMaxGaussTaps = 9;   // base_scale_offset <= 4
constexpr int kMaxDetScale = 6;    // Scharr scale in the detector (default config needs 4)
constexpr int kMaxBins = 1024;     // contrast histogram bins (default 300)
constexpr int kDescStride = AKZ_DESCRIPTOR_STRIDE;

// ---- per-level constants, host + device -------------------------------------------------------
struct LevelDev {
    int w, h;             // level size
    int octave;           // evolution.octave
    int s_det;            // Scharr scale of the detector (detector_response.rs:22-24)
    int wpr;              // candidate-mask words per row
    int xmin, xmax;       // inclusive x range passing the is_out test (scale_space_extrema.rs:80-87)
    int ymin, ymax;
    int new_octave;       // level starts a new octave (lib.rs:80)
    int n_steps;
    float ratio;          // 2^octave as f32
    float kp_size;        // (esigma * derivative_factor) as f32
    float size_sq;        // kp_size * kp_size (f32 product)
    float s_smp;          // round(0.5 * size / ratio): sample step of orientation and descriptor
    float half_ratio_m1;  // 0.5f * (ratio - 1.0f)
    unsigned long long off;       // float offset of this level inside a per-image plane
    unsigned long long mask_off;  // word offset inside a per-image candidate mask
};

struct PlanDev {
    int n_levels;
    int w0, h0;
    int channels;          // descriptor_channels
    int pattern_size;      // descriptor_pattern_size
    int desc_len;          // (162*channels+7)/8
    int n_bins;
    float det_threshold;   // detector_threshold as f32
    double percentile;
    unsigned long long plane_px;    // sum of level pixels (per image)
    unsigned long long mask_words;  // per image
    unsigned long long orient_window_mask;  // bit w set: sliding window w contains atan2(v,v), v>0
    int n_orient_windows;
    int grid_shift;        // dedup hash grid: cell = 1<<grid_shift full-resolution pixels
    int grid_w, grid_h;
    int pool_cap;          // entries per class-parity pool of the single-warp cache pass (per image)
    // level-pipelined cache pass: one row table of u32 pool indices per level in shared memory, level l (lv[l].h + 2
    // entries) at ltab_off[l]; ltab_off[n_levels] = entries in total
    int ltab_off[kMaxLevels + 1];
    LevelDev lv[kMaxLevels];
};

struct LevelHost {
    akz_level_info info;
    std::vector<double> tau;
    std::vector<float> half_tau;  // 0.5f * (tau as f32)  (nonlinear_diffusion.rs:67)
};

struct Plan {
    uint32_t w = 0, h = 0;
    akz_config cfg{};
    PlanDev dev{};
    std::vector<LevelHost> host;
    float gauss0[kMaxGaussTaps];  // taps of gaussian_blur(image, base_scale_offset)
    int gauss0_n = 0;
    float gauss1[3];              // taps of gaussian_blur(.., 1.0)
    float sch_n[kMaxDetScale + 1], sch_wn[kMaxDetScale + 1];  // Scharr main-axis taps per scale
};

// builds the level table; returns "" or an error message (host_plan.cpp part of akaze_api.cu)
std::string build_plan(uint32_t w, uint32_t h, const akz_config& cfg, Plan* out);

// ---- device buffers of one context -------------------------------------------------------------
struct Buffers {
    // persistent per batch: [level][image][N_l] inside each plane
    float* Lt = nullptr;
    float* Lx = nullptr;
    float* Ly = nullptr;
    float* Ldet = nullptr;
    // per-level scratch (octave-0 sized, per image) unless keep_evolutions
    float* Lsmooth = nullptr;
    float* Lsmooth2 = nullptr;  // second scratch (odd levels): detector(l) on its own stream may still read Lsmooth_l while
                                // prep(l+1) writes Lsmooth_{l+1}
    float* Lflow = nullptr;
    float* Ltmp = nullptr;
    // keep_evolutions extras (plane layout)
    float* Lxx = nullptr;
    float* Lyy = nullptr;
    float* Lxy = nullptr;
    float* Lstep = nullptr;
    bool keep = false;
    // input staging
    uint8_t* in_u8 = nullptr;
    float* in_f32 = nullptr;
    // contrast
    unsigned long long* hmax_bits = nullptr;  // [B] f64 bits (values >= 0 order like u64)
    double* contrast_thr = nullptr;           // [B][kMaxBins + 1] smallest squared gradient magnitude that falls into bin b
    unsigned int* hist = nullptr;             // [B][n_bins]
    double* kcontrast = nullptr;              // [B][kMaxLevels] contrast factor per level
    unsigned int* fine_hist = nullptr;        // [B][contrast_fine_bins()] hmax-independent fine histogram of g2 (scale_space.cu)
    int* contrast_resolved = nullptr;         // [B] bins walked, decided from the fine histogram, or -1
    unsigned long long* contrast_npoints = nullptr;  // [B]
    // candidates
    unsigned int* mask = nullptr;       // [B][mask_words]
    unsigned int* cand = nullptr;       // [B][cand_cap] packed flat index, level-major raster order
    unsigned int* rowcount = nullptr;   // [B][sum of level heights] scratch of the compaction
    unsigned int* cand_level_count = nullptr;  // [B][kMaxLevels+1] exclusive offsets per level
    // dedup cache
    float* c_x = nullptr;               // [B][kp_cap]
    float* c_y = nullptr;
    float* c_resp = nullptr;
    float* r_x = nullptr;               // [B][kp_cap] refined points (sub-pixel step)
    float* r_y = nullptr;
    int* c_cls = nullptr;
    int* c_next = nullptr;
    int* grid = nullptr;                // [B][2][grid_cells]
    unsigned char* dedup_pool = nullptr;  // [B][dedup_pool_bytes()] pools of the single-warp cache pass
    unsigned char* level_pool = nullptr;  // [B][dedup_level_pool_bytes()] per-level pools of the level-pipelined cache pass
    unsigned int* n_cache = nullptr;    // [B]
    unsigned int* n_cand_total = nullptr;  // [B]
    unsigned int* keep_flag = nullptr;  // [B][kp_cap]
    unsigned int* cls_range = nullptr;  // [B][kMaxLevels][2] slot range per class_id
    unsigned int* upper_done = nullptr; // [B] 1: the cache pass already ran the upper-scale filter (verdicts in keep_flag)
    unsigned int* n_kp = nullptr;       // [B]
    unsigned int* err_flags = nullptr;  // [B]
    akz_keypoint* kps = nullptr;        // [B][kp_cap]
    uint8_t* desc = nullptr;            // [B][kp_cap][64]
    PlanDev* plan_dev = nullptr;
    float* half_tau_dev = nullptr;      // [kMaxLevels][kMaxFedSteps]
};

enum ErrBits : unsigned int {
    kErrCandOverflow = 1u,
    kErrKpOverflow = 2u,
    kErrBounds = 4u,
};

// ---- launchers (each returns the number of kernels launched) ---------------------------------
struct Launch {
    cudaStream_t stream;
    int batch;          // images in this launch
    uint32_t cand_cap;  // per image
    uint32_t kp_cap;    // per image
};

// scale_space.cu
int launch_level0(const Launch& L, const Plan& P, const Buffers& B, const void* d_in, bool is_u8, size_t in_stride);
int launch_contrast(const Launch& L, const Plan& P, const Buffers& B);
size_t contrast_fine_bins();
int launch_prep(const Launch& L, const Plan& P, const Buffers& B, int level);
int launch_fed(const Launch& L, const Plan& P, const Buffers& B, int level);
// detector.cu
cudaError_t init_detector_attributes();
cudaError_t init_scale_space_attributes();
int launch_detector(const Launch& L, const Plan& P, const Buffers& B, int level);
// shortest segment the streaming kernels cut a small grid into (AKZ_MIN_SEGMENT overrides, for profiling)
inline int min_segment_rows() {
    static const int v = getenv("AKZ_MIN_SEGMENT") ? std::max(2, atoi(getenv("AKZ_MIN_SEGMENT")) & ~1) : 8;
    return v;
}
int launch_compact(const Launch& L, const Plan& P, const Buffers& B, int l0 = 0, int l1 = -1);  // levels [l0, l1), -1 = all
// keypoints.cu
cudaError_t init_keypoint_attributes();
size_t dedup_pool_bytes(const Plan& P);
size_t dedup_level_pool_bytes(const Plan& P, uint32_t cand_cap);
// levels [l0, l1) of the cache pass; the call that ends at the last level also assigns the final slots. split_level(): where a
// two-part pass is cut (0 = this plan does not split)
int launch_dedup(const Launch& L, const Plan& P, const Buffers& B, int l0 = 0, int l1 = -1);
int dedup_split_level(const Plan& P);
int launch_finalize(const Launch& L, const Plan& P, const Buffers& B);
int launch_descriptors(const Launch& L, const Plan& P, const Buffers& B);
// matcher.cu
int match_parts(uint64_t nq, uint64_t ndb);  // how many database parts the grid is split into
// d_out: akz_top2[n_parts][nq]
int launch_match_top2(cudaStream_t s, const uint8_t* d_q, uint64_t nq, const uint8_t* d_db, uint64_t ndb,
                      uint32_t db_index_base, akz_top2* d_out, int n_parts);
int launch_merge_top2(cudaStream_t s, const akz_top2* d_parts, uint32_t n_parts, uint64_t nq, akz_top2* d_out);
int launch_repack_rows(cudaStream_t s, const uint8_t* d_src, size_t stride, uint32_t desc_len, uint64_t n, uint8_t* d_dst);
// matcher_tc.cu (tcgen05 int8 path)
cudaError_t init_matcher_tc_attributes();
size_t match_tc_query_image_bytes(uint64_t nq);
size_t match_tc_db_image_bytes(uint64_t ndb);
int match_tc_parts(uint64_t nq, uint64_t ndb);
int launch_match_tc(cudaStream_t s, const uint8_t* d_q, uint64_t nq, const uint8_t* d_db, uint64_t ndb, uint32_t db_index_base,
                    uint8_t* q_img, uint8_t* db_img, akz_top2* d_out, int n_parts);

// where level `level`'s Lsmooth lives: per-level slabs when evolutions are kept, else one of two scratch planes by parity
static inline float* lsmooth_slab(const Plan& P, const Buffers& B, int batch, int level) {
    if (B.keep) return B.Lsmooth + (size_t)P.dev.lv[level].off * (size_t)batch;
    return (level & 1) ? B.Lsmooth2 : B.Lsmooth;
}

// ransac.cu
int launch_ransac_models(cudaStream_t s, const float2* pl, const float2* pr, const unsigned int* samples, unsigned int n_trials,
                         float epsilon, float* models, unsigned int* ok);
int launch_ransac_count(cudaStream_t s, const float2* pl, const float2* pr, unsigned int n_matches, const float* models,
                        const unsigned int* ok, unsigned int n_trials, float epsilon_inlier, unsigned int* counts);
int launch_ransac_mask(cudaStream_t s, const float2* pl, const float2* pr, unsigned int n_matches, const float* model, float epsilon_inlier,
                       unsigned char* mask);

// helpers to address [level][image] slabs
static inline size_t plane_off(const Plan& P, int batch, int level, int img) {
    return (size_t)P.dev.lv[level].off * (size_t)batch + (size_t)img * (size_t)P.dev.lv[level].w * P.dev.lv[level].h;
}

}  // namespace akz
