/*
 * akaze_oracle.h -- CPU restatement of the akaze-rust hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle and the CPU baseline ("port") for the B200 engine. It is NOT part of the
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it. The product library (libakaze_b200.so) never links or calls anything in oracle/.
 *
 * PARITY UNPINNED: the reference (indianajohn/akaze-rust) cannot be built here (no cargo/rustc, no
 * vendored crates) and its own tests hold no golden keypoints/descriptors/matches. The only
 * known-answer vectors in the reference are the Gaussian taps for (sigma=3, size=7)
 * (akaze/src/types/image.rs:486-502) and the Scharr(1) taps (akaze/src/ops/derivatives.rs:12,22); both
 * are checked in tests/test_oracle_pins.py. Everything else follows the reference source line by line
 * (citations on each function) and is cross-checked by an independent numpy restatement in tests/.
 */
#ifndef AKAZE_ORACLE_H
#define AKAZE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* field-for-field mirror of types::evolution::Config (akaze/src/types/evolution.rs:8-38) */
typedef struct akzo_config {
    uint32_t num_sublevels;
    uint32_t max_octave_evolution;
    double base_scale_offset;
    double initial_contrast;
    double contrast_percentile;
    uint64_t contrast_factor_num_bins;
    double derivative_factor;
    double detector_threshold;
    uint64_t descriptor_channels;
    uint64_t descriptor_pattern_size;
} akzo_config;

/* mirror of types::keypoint::Keypoint (akaze/src/types/keypoint.rs:8-30) */
typedef struct akzo_keypoint {
    float x, y;
    float response;
    float size;
    uint32_t octave;
    uint32_t class_id;
    float angle;
} akzo_keypoint;

/* mirror of types::feature_match::Match (akaze/src/types/feature_match.rs:9-16) */
typedef struct akzo_match {
    uint64_t index_0;
    uint64_t index_1;
    double distance;
} akzo_match;

typedef struct akzo_level_info {
    uint32_t octave, sublevel, sigma_size;
    uint32_t width, height;
    uint32_t n_steps;
    double esigma, etime;
} akzo_level_info;

enum akzo_image_kind {
    AKZO_LT = 0, AKZO_LSMOOTH = 1, AKZO_LX = 2, AKZO_LY = 3, AKZO_LXX = 4, AKZO_LYY = 5,
    AKZO_LXY = 6, AKZO_LFLOW = 7, AKZO_LSTEP = 8, AKZO_LDET = 9
};

typedef struct akzo_result akzo_result;

void akzo_default_config(akzo_config *cfg);

/* small pieces, exposed for unit tests */
void akzo_gaussian_kernel(float r, int kernel_size, float *out);
void akzo_scharr_main_axis_kernel(uint32_t scale, float *out);
void akzo_scharr_off_axis_kernel(uint32_t scale, float *out);
int akzo_fed_tau_by_process_time(double T, int M, double tau_max, int reordering, double *out, int cap);
void akzo_unit_float_from_u8(const uint8_t *gray, size_t n, float *out);
void akzo_horizontal_filter(const float *in, int w, int h, const float *kernel, int ksize, float *out);
void akzo_vertical_filter(const float *in, int w, int h, const float *kernel, int ksize, float *out);
void akzo_gaussian_blur(const float *in, int w, int h, float r, float *out);
void akzo_scharr(const float *in, int w, int h, int x_order, int y_order, uint32_t sigma_size, float *out);
void akzo_half_size(const float *in, int w, int h, float *out);
void akzo_pm_g2(const float *lx, const float *ly, size_t n, double k, float *out);
double akzo_compute_contrast_factor(const float *img, int w, int h, double percentile, double scale,
                                    uint64_t nbins);
void akzo_calculate_step(float *lt, const float *lflow, float *lstep, int w, int h, double step_size);

/* full pipeline; unit_gray is the GrayFloatImage buffer produced by create_unit_float_image.
 * threads: worker threads for the multiscale-derivative stage (the only threaded stage of the
 * reference, detector_response.rs:16-29); <=1 means serial. stop_after: 0 = everything,
 * 1 = stop after the scale space, 2 = stop after the detector response. Returns NULL on allocation
 * failure; akzo_result_status()!=0 marks inputs on which the reference would have panicked. */
akzo_result *akzo_extract(const float *unit_gray, uint32_t w, uint32_t h, const akzo_config *cfg,
                          int threads, int stop_after);
void akzo_result_free(akzo_result *r);
int akzo_result_status(const akzo_result *r);
uint32_t akzo_result_num_levels(const akzo_result *r);
void akzo_result_level_info(const akzo_result *r, uint32_t level, akzo_level_info *out);
const double *akzo_result_fed_tau(const akzo_result *r, uint32_t level);
const float *akzo_result_image(const akzo_result *r, uint32_t level, int kind);
double akzo_result_contrast_factor(const akzo_result *r);
uint64_t akzo_result_num_candidates(const akzo_result *r); /* 4-neighbour maxima seen (incl. rejected) */
uint64_t akzo_result_num_cache(const akzo_result *r);      /* keypoint_cache length before upper filter */
uint64_t akzo_result_num_keypoints(const akzo_result *r);
const akzo_keypoint *akzo_result_keypoints(const akzo_result *r);
uint64_t akzo_result_descriptor_len(const akzo_result *r);
const uint8_t *akzo_result_descriptors(const akzo_result *r); /* n x descriptor_len, dense */

/* matcher: raw top-2 (what the GPU kernel returns) and the full descriptor_match */
void akzo_match_top2(const uint8_t *q, uint64_t nq, const uint8_t *db, uint64_t ndb, uint64_t desc_len,
                     uint64_t stride, uint32_t *best_idx, uint32_t *best, uint32_t *second);
uint64_t akzo_descriptor_match(const uint8_t *d0, uint64_t n0, const uint8_t *d1, uint64_t n1,
                               uint64_t desc_len, uint64_t stride, uint64_t distance_threshold,
                               double lowes_ratio, akzo_match *out);

#ifdef __cplusplus
}
#endif
#endif
