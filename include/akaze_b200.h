/*
 * akaze_b200.h -- C ABI of the B200-native A-KAZE engine (libakaze_b200.so, CUDA sm_100a).
 *
 * Drop-in boundary for the hot path of indianajohn/akaze-rust. The reference has no FFI seam of its
 * own; the boundary is its public Rust API, and these entry points are what a thin Rust `ffi.rs`
 * binds. Each entry point cites the reference interface it replaces
 * (paths relative to the reference repository).
 *
 * Conventions: every function returns an int status (AKZ_OK == 0); akz_last_error() gives a
 * thread-local message for the last failure on the calling thread. Inputs are caller-owned plain
 * buffers; outputs are either caller-provided buffers or library-owned handles released with the
 * matching akz_*_free. A context is bound to one CUDA device. Thread safety: every entry point that
 * takes a context locks it, so one context may be shared by several threads (calls are serialised
 * per context; use one context per thread or per GPU for concurrency); akz_features accessors only
 * read host memory owned by the handle and need no lock; akz_last_error() is per thread.
 * There is NO CPU fallback: every entry point that computes fails with AKZ_ERR_CUDA if
 * no usable sm_100 device is present.
 */
#ifndef AKAZE_B200_H
#define AKAZE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AKZ_DESCRIPTOR_STRIDE 64 /* bytes per descriptor row in every dense descriptor array (61 used) */

enum akz_status {
    AKZ_OK = 0,
    AKZ_ERR_INVALID = 1,  /* bad argument */
    AKZ_ERR_CUDA = 2,     /* CUDA runtime / driver error, or no device */
    AKZ_ERR_CAPACITY = 3, /* image, batch, candidate or keypoint count exceeds the context's limits */
    AKZ_ERR_BOUNDS = 4,   /* input on which the reference would panic (out-of-image sample) */
    AKZ_ERR_NOMEM = 5
};

/* types::evolution::Config, field for field (akaze/src/types/evolution.rs:8-38; defaults :41-54). */
typedef struct akz_config {
    uint32_t num_sublevels;
    uint32_t max_octave_evolution;
    double base_scale_offset;
    double initial_contrast; /* never read by the reference either */
    double contrast_percentile;
    uint64_t contrast_factor_num_bins;
    double derivative_factor;
    double detector_threshold;
    uint64_t descriptor_channels;
    uint64_t descriptor_pattern_size;
} akz_config;

/* types::keypoint::Keypoint (akaze/src/types/keypoint.rs:8-30); usize fields narrowed to u32. */
typedef struct akz_keypoint {
    float x, y; /* point.0, point.1 */
    float response;
    float size;
    uint32_t octave;
    uint32_t class_id;
    float angle;
} akz_keypoint;

/* types::feature_match::Match (akaze/src/types/feature_match.rs:9-16). */
typedef struct akz_match {
    uint64_t index_0;
    uint64_t index_1;
    double distance;
} akz_match;

/* Per-level metadata of types::evolution::EvolutionStep (akaze/src/types/evolution.rs:59-92). */
typedef struct akz_level_info {
    uint32_t octave, sublevel, sigma_size;
    uint32_t width, height;
    uint32_t n_steps; /* fed_tau_steps.len() */
    double esigma, etime;
} akz_level_info;

/* The ten images of an EvolutionStep (evolution.rs:70-89). */
enum akz_image_kind {
    AKZ_LT = 0, AKZ_LSMOOTH = 1, AKZ_LX = 2, AKZ_LY = 3, AKZ_LXX = 4, AKZ_LYY = 5, AKZ_LXY = 6,
    AKZ_LFLOW = 7, AKZ_LSTEP = 8, AKZ_LDET = 9
};

/* Per-query result of the brute-force matcher before the Lowe test (feature_matching.rs:37-50):
 * the two smallest Hamming distances of {d_j} U {10000,10000} and the lowest j attaining the minimum. */
typedef struct akz_top2 {
    uint32_t best_idx;
    uint16_t best;
    uint16_t second;
} akz_top2;

/* akz_create flags */
#define AKZ_KEEP_EVOLUTIONS 1u /* retain all ten images of every level so that
                                  akz_features_evolution_download works (debug / full drop-in of the
                                  Vec<EvolutionStep> return value; costs memory and bandwidth) */

typedef struct akz_context akz_context;
typedef struct akz_features akz_features;

/* ---- housekeeping --------------------------------------------------------------------------- */
const char *akz_last_error(void);
const char *akz_version(void);
/* Config::default() (evolution.rs:41-54) */
int akz_default_config(akz_config *cfg);
/* Creates an engine on `device` able to extract from images up to max_width x max_height, up to
 * max_batch images per call. Fails with AKZ_ERR_CUDA when there is no sm_100 GPU. */
int akz_create(int device, uint32_t max_width, uint32_t max_height, uint32_t max_batch, uint32_t flags,
               akz_context **out);
void akz_destroy(akz_context *ctx);
/* The CUDA stream (cudaStream_t) all work of this context is issued on; for event timing. */
void *akz_context_stream(akz_context *ctx);
/* Number of kernel launches issued by this context since creation (bench.py's gpu_launches). */
uint64_t akz_context_launch_count(const akz_context *ctx);
/* Per-stage device timing: when enabled, every extraction brackets each stage's launches with CUDA events
 * on the context's stream. akz_context_stage_times copies the accumulated milliseconds and launch counts
 * (AKZ_NUM_STAGES entries each, order of enum akz_stage) and optionally resets them. */
enum akz_stage {
    AKZ_STAGE_LEVEL0 = 0, AKZ_STAGE_CONTRAST = 1, AKZ_STAGE_PREP = 2, AKZ_STAGE_FED = 3, AKZ_STAGE_DETECTOR = 4,
    AKZ_STAGE_COMPACT = 5, AKZ_STAGE_DEDUP = 6, AKZ_STAGE_FINALIZE = 7, AKZ_STAGE_DESCRIPTOR = 8, AKZ_NUM_STAGES = 9
};
int akz_context_enable_timing(akz_context *ctx, int enable);
int akz_context_stage_times(akz_context *ctx, double *ms, uint64_t *launches, int reset);
/* Per-image capacities of the candidate list (4-neighbour maxima inside the descriptor margin) and of
 * the keypoint cache; exceeding either makes the extraction fail with AKZ_ERR_CAPACITY instead of
 * truncating (the reference's Vec<Keypoint> is unbounded, scale_space_extrema.rs:17). Defaults follow the
 * image size: max(65536, pixels / 16) keypoints and four times as many candidates (a raw 3840x2160 frame
 * gets 518400 / 2073600); this call pins them instead. Call before the first extraction. */
int akz_context_set_limits(akz_context *ctx, uint32_t max_candidates, uint32_t max_keypoints);

/* A call of n images is cut into sub-batches of `images` that flow through a two-stage software pipeline
 * (stencil stages of sub-batch i+1 overlap the keypoint stages and the result download of sub-batch i).
 * Default: up to 256, chosen so that both lanes fit in half of the free device memory; memory per in-flight image is about 0.2 GB at 1080p. Environment override: AKZ_SUB_BATCH. */
int akz_context_set_sub_batch(akz_context *ctx, uint32_t images);

/* Which kernel akz_match_top2* runs: AUTO picks the tcgen05 int8 path (matcher_tc.cu) for nq*ndb >= 2^12 pairs
 * and the integer-popc path (matcher.cu) below that; both are bit-exact. */
enum akz_match_path { AKZ_MATCH_AUTO = 0, AKZ_MATCH_POPC = 1, AKZ_MATCH_TENSOR = 2 };
int akz_context_set_match_path(akz_context *ctx, int path);

/* ---- extraction: replaces akaze::extract_features (akaze/src/lib.rs:167-194) from the
 *      GrayFloatImage on; decode + to_luma (lib.rs:171, image.rs:128) stay with the host ------- */
/* gray: 8-bit luma, row stride in bytes; the u8 -> unit float conversion of
 * create_unit_float_image (image.rs:127-140) runs on the device. */
int akz_extract_u8(akz_context *ctx, const uint8_t *gray, uint32_t width, uint32_t height, size_t stride,
                   const akz_config *cfg, akz_features **out);
/* unit_gray: exactly the GrayFloatImage buffer of image.rs:127-140 (row-major, width*height). */
int akz_extract_f32(akz_context *ctx, const float *unit_gray, uint32_t width, uint32_t height,
                    const akz_config *cfg, akz_features **out);
/* n images of identical size; outs[n] receives one handle per image. Uploads, kernels and result downloads
 * of consecutive sub-batches overlap; keypoints and descriptors land in pinned host slabs owned by the
 * handles (pass pinned images for fully asynchronous uploads). */
int akz_extract_batch_u8(akz_context *ctx, uint32_t n, const uint8_t *const *grays, uint32_t width,
                         uint32_t height, size_t stride, const akz_config *cfg, akz_features **outs);
/* Throughput path: the n images already sit in device memory (n * height * stride bytes, contiguous);
 * keypoints and descriptors stay on the device and only the per-image counts come back.
 * counts[n] receives the number of keypoints per image. Device result pointers (valid until the next
 * extraction on this context) via akz_context_device_results. */
int akz_extract_batch_u8_device(akz_context *ctx, uint32_t n, const void *d_grays, uint32_t width,
                                uint32_t height, size_t stride, const akz_config *cfg, uint32_t *counts);
/* d_keypoints: akz_keypoint[n][kp_capacity]; d_descriptors: uint8_t[n][kp_capacity][64]. */
int akz_context_device_results(akz_context *ctx, void **d_keypoints, void **d_descriptors,
                               uint32_t *kp_capacity);

uint64_t akz_features_count(const akz_features *f);
const akz_keypoint *akz_features_keypoints(const akz_features *f);
/* count x AKZ_DESCRIPTOR_STRIDE bytes; bytes [descriptor_len, 64) of every row are zero */
const uint8_t *akz_features_descriptors(const akz_features *f);
uint32_t akz_features_descriptor_len(const akz_features *f); /* (162*channels+7)/8, descriptors.rs:42-46 */
uint32_t akz_features_num_levels(const akz_features *f);
int akz_features_level_info(const akz_features *f, uint32_t level, akz_level_info *out);
/* fed_tau_steps of a level (evolution.rs:91,153); copies min(cap, n_steps) doubles */
int akz_features_fed_tau(const akz_features *f, uint32_t level, double *out, uint32_t cap);
double akz_features_contrast_factor(const akz_features *f); /* contrast_factor.rs:18-71 result */
/* candidate / cache statistics of find_scale_space_extrema (scale_space_extrema.rs:12-132) */
uint64_t akz_features_num_candidates(const akz_features *f);
uint64_t akz_features_num_cache(const akz_features *f);
/* Copies one image of one EvolutionStep (width*height floats) to dst; must be called before the next
 * extraction on the same context (AKZ_ERR_INVALID afterwards, or once the context is destroyed). All ten
 * images need AKZ_KEEP_EVOLUTIONS. Without it the four planes the keypoint stages sample -- Lt, Lx, Ly, Ldet
 * (and Lsmooth of level 0, which is Lt_0, lib.rs:58) -- are still served, for the images of the last two
 * pipeline sub-batches of the call (every image of a call of up to 2 x sub-batch images). */
int akz_features_evolution_download(const akz_features *f, uint32_t level, int kind, float *dst);
void akz_features_free(akz_features *f);

/* ---- matching: replaces ops::feature_matching::descriptor_match
 *      (akaze/src/ops/feature_matching.rs:23-94), the first half of akaze::match_features
 *      (lib.rs:252-266); RANSAC (lib.rs:267) stays with the host --------------------------------- */
/* Brute-force Hamming top-2. q, db: n x stride bytes, the first desc_len bytes of each row compared
 * (desc_len <= 64). out[nq]. Host buffers. */
int akz_match_top2(akz_context *ctx, const uint8_t *q, uint64_t nq, const uint8_t *db, uint64_t ndb,
                   uint32_t desc_len, size_t stride, akz_top2 *out);
/* Same on device-resident descriptors padded to 64-byte rows (bytes >= desc_len must be zero);
 * d_out: akz_top2[nq] in device memory; best_idx is offset by db_index_base (for sharded databases). */
int akz_match_top2_device(akz_context *ctx, const void *d_q, uint64_t nq, const void *d_db, uint64_t ndb,
                          uint32_t db_index_base, void *d_out);
/* Merges per-shard results: d_parts is akz_top2[n_parts][nq] (shards ordered by ascending database
 * index range); d_out: akz_top2[nq]. Same tie rule as the sequential scan (lowest index wins). */
int akz_merge_top2_device(akz_context *ctx, const void *d_parts, uint32_t n_parts, uint64_t nq, void *d_out);
/* descriptor_match proper: top-2 on the device, then the f64 Lowe-ratio and threshold tests
 * (feature_matching.rs:61-80) on the host, for any distance_threshold (akaze::match_features passes
 * 10000, lib.rs:264). out must hold n0 entries; *n_out receives the count. */
int akz_descriptor_match(akz_context *ctx, const uint8_t *d0, uint64_t n0, const uint8_t *d1, uint64_t n1,
                         uint32_t desc_len, size_t stride, uint64_t distance_threshold, double lowes_ratio,
                         akz_match *out, uint64_t *n_out);

/* ---- RANSAC: replaces ops::estimate_fundamental_matrix::remove_outliers
 *      (akaze/src/ops/estimate_fundamental_matrix.rs:99-165; with estimate_fundamental_matrix :17-69 and
 *      evaluate_model :79-83), the second half of akaze::match_features (lib.rs:267-274). All trials run
 *      in parallel: one thread per hypothesis (8 x 9 system, SVD, rank test), one block per hypothesis for
 *      the inlier count. Fewer than 8 matches are returned untouched (:107-110).
 *      The reference builds a FRESH default random source on every trial (:118), so all its trials draw the
 *      same eight matches: AKZ_RANSAC_REFERENCE reproduces that, AKZ_RANSAC_ADVANCING runs one source on
 *      across the trials (num_trials distinct hypotheses). The source is the `random` crate's default,
 *      xorshift128+ seeded [42, 69] (third party, unpinned); the eight matches are used in ascending index
 *      order (the reference's HashSet order is process-random). kp0 / kp1: the keypoint arrays the matches
 *      index; out must hold n_matches entries; model (may be NULL) receives the winning 3x3 matrix row by row. */
enum akz_ransac_sampling { AKZ_RANSAC_REFERENCE = 0, AKZ_RANSAC_ADVANCING = 1 };
int akz_remove_outliers(akz_context *ctx, const akz_keypoint *kp0, uint64_t n0, const akz_keypoint *kp1, uint64_t n1,
                        const akz_match *matches, uint64_t n_matches, uint64_t num_trials, float epsilon_model,
                        float epsilon_inlier, int sampling, akz_match *out, uint64_t *n_out, float *model);

/* ---- multi-GPU matching: the database is partitioned contiguously by index over the
 *      GPUs, queries are replicated, every GPU runs the top-2 scan on its shard, the 8-byte records are
 *      all-gathered with NCCL over NVLink and merged with the sequential scan's tie rule (lowest index),
 *      so the result equals descriptor_match's unsharded scan (feature_matching.rs:37-50) bit for bit.
 *      NCCL (libnccl.so.2) is bound at run time; environment AKZ_NCCL_LIB overrides the library path. */
#define AKZ_COMM_UNIQUE_ID_BYTES 128
/* ncclGetUniqueId: call on one rank, ship the 128 bytes to the others by any means. */
int akz_comm_unique_id(uint8_t *id);
/* One process (or thread) per GPU: joins this context to a communicator of n_ranks (ncclCommInitRank). */
int akz_context_comm_init(akz_context *ctx, const uint8_t *id, int rank, int n_ranks);
/* One process driving n GPUs: a communicator over n contexts on n distinct devices (ncclCommInitAll);
 * ctxs[i] becomes rank i. */
int akz_context_comm_init_all(akz_context *const *ctxs, int n);
int akz_context_comm_destroy(akz_context *ctx);
/* Per rank, device buffers: d_db_shard holds the rank's slice [db_index_base, db_index_base + ndb_shard)
 * of the database (rank order = ascending index order), d_q all nq queries; shard scan, ncclAllGather
 * and merge are issued on the context's stream; every rank's d_out receives akz_top2[nq]. */
int akz_match_top2_sharded_device(akz_context *ctx, const void *d_q, uint64_t nq, const void *d_db_shard,
                                  uint64_t ndb_shard, uint32_t db_index_base, void *d_out);
/* Single-process form with host buffers: shards db over the n_gpu contexts of one communicator
 * (akz_context_comm_init_all), out[nq] as akz_match_top2. */
int akz_match_top2_sharded(akz_context *const *ctxs, int n_gpu, const uint8_t *q, uint64_t nq, const uint8_t *db,
                           uint64_t ndb, uint32_t desc_len, size_t stride, akz_top2 *out);

#ifdef __cplusplus
}
#endif
#endif
