/*
 * akaze_oracle.c -- plain-C restatement of the akaze-rust CPU hot path. TEST INFRASTRUCTURE ONLY
 * (see akaze_oracle.h). PARITY UNPINNED beyond the Gaussian/Scharr tap vectors: the reference cannot
 * be executed in this environment.
 *
 * Every function cites the reference file:line it follows (paths relative to the reference repo).
 * Loop structure, operand order and f32/f64 widths follow the Rust source literally, including its
 * quirks (flattened-buffer filter sweeps, swapped Scharr axes, no-op LU solve, atan2(y,y), running
 * orientation sums). Build with -ffp-contract=off and without fast-math: rustc never contracts
 * a*b+c into an FMA, so neither may the compiler here.
 */
#include "akaze_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * GrayFloatImage (types/image.rs:32-36, 78-101)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    float *buf;
    int w, h;
} img_t;

static img_t img_new(int w, int h) {
    img_t im;
    im.w = w;
    im.h = h;
    im.buf = (float *)calloc((size_t)w * (size_t)h + 1, sizeof(float));
    return im;
}
static void img_free(img_t *im) {
    free(im->buf);
    im->buf = NULL;
    im->w = im->h = 0;
}
static img_t img_clone(const img_t *src) {
    img_t im = img_new(src->w, src->h);
    memcpy(im.buf, src->buf, (size_t)src->w * src->h * sizeof(float));
    return im;
}
static inline float img_get(const img_t *im, size_t x, size_t y) { return im->buf[(size_t)im->w * y + x]; }
static inline void img_put(img_t *im, size_t x, size_t y, float v) { im->buf[(size_t)im->w * y + x] = v; }

/* Rust `as usize` on an f32: saturating, NaN -> 0 */
static size_t f32_as_usize(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 18446744073709551616.0f) return SIZE_MAX;
    return (size_t)v;
}

/* types/evolution.rs:41-54 */
void akzo_default_config(akzo_config *c) {
    c->num_sublevels = 4;
    c->max_octave_evolution = 4;
    c->base_scale_offset = 1.6;
    c->initial_contrast = 0.001;
    c->contrast_percentile = 0.7;
    c->contrast_factor_num_bins = 300;
    c->derivative_factor = 1.5;
    c->detector_threshold = 0.001;
    c->descriptor_channels = 3;
    c->descriptor_pattern_size = 10;
}

/* types/image.rs:127-140 -- f32::from(v) * 1f32 / 255f32 */
void akzo_unit_float_from_u8(const uint8_t *gray, size_t n, float *out) {
    for (size_t i = 0; i < n; i++) out[i] = ((float)gray[i] * 1.0f) / 255.0f;
}

/* types/image.rs:102-118 -- sum order (2x,2y),(2x,2y+1),(2x+1,2y),(2x+1,2y+1), then /4 */
static img_t half_size(const img_t *src) {
    int width = src->w / 2, height = src->h / 2;
    img_t out = img_new(width, height);
    for (int x = 0; x < width; x++) {
        for (int y = 0; y < height; y++) {
            float val = 0.0f;
            for (int xs = 2 * x; xs < 2 * x + 2; xs++)
                for (int ys = 2 * y; ys < 2 * y + 2; ys++) val += img_get(src, xs, ys);
            img_put(&out, x, y, val / 4.0f);
        }
    }
    return out;
}
void akzo_half_size(const float *in, int w, int h, float *out) {
    img_t s = {(float *)in, w, h};
    img_t o = half_size(&s);
    memcpy(out, o.buf, (size_t)o.w * o.h * sizeof(float));
    img_free(&o);
}

/* types/image.rs:239-260 */
static void fill_border(img_t *o, size_t hw) {
    for (size_t x = 0; x < (size_t)o->w; x++) {
        float plus = img_get(o, x, hw);
        float minus = img_get(o, x, (size_t)o->h - hw - 1);
        for (size_t y = 0; y < hw; y++) img_put(o, x, y, plus);
        for (size_t y = (size_t)o->h - hw; y < (size_t)o->h; y++) img_put(o, x, y, minus);
    }
    for (size_t y = 0; y < (size_t)o->h; y++) {
        float plus = img_get(o, hw, y);
        float minus = img_get(o, (size_t)o->w - hw - 1, y);
        for (size_t x = 0; x < hw; x++) img_put(o, x, y, plus);
        for (size_t x = (size_t)o->w - hw; x < (size_t)o->w; x++) img_put(o, x, y, minus);
    }
}

/* types/image.rs:270-295 -- one full sweep of the FLATTENED buffer per tap (rows wrap), zeroed output,
 * separate f32 multiply and add, then fill_border */
static img_t horizontal_filter(const img_t *image, const float *kernel, int ksize) {
    long hw = ksize / 2;
    long w = image->w, h = image->h;
    img_t output = img_new(image->w, image->h);
    for (long k = -hw; k <= hw; k++) {
        float kv = kernel[k + hw];
        float *out_ptr = output.buf + hw;
        const float *in_ptr = image->buf + hw + k;
        long n = (w * h - hw - 1) - hw;
        for (long t = 0; t < n; t++) out_ptr[t] += kv * in_ptr[t];
    }
    fill_border(&output, (size_t)hw);
    return output;
}

/* types/image.rs:305-332 */
static img_t vertical_filter(const img_t *image, const float *kernel, int ksize) {
    long hw = ksize / 2;
    long w = image->w, h = image->h;
    img_t output = img_new(image->w, image->h);
    for (long k = -hw; k <= hw; k++) {
        float kv = kernel[k + hw];
        float *out_ptr = output.buf + hw * w;
        const float *in_ptr = image->buf + hw * w + k * w;
        long n = (w * h - hw * w - 1) - hw * w;
        for (long t = 0; t < n; t++) out_ptr[t] += kv * in_ptr[t];
    }
    fill_border(&output, (size_t)hw);
    return output;
}
void akzo_horizontal_filter(const float *in, int w, int h, const float *kernel, int ksize, float *out) {
    img_t s = {(float *)in, w, h};
    img_t o = horizontal_filter(&s, kernel, ksize);
    memcpy(out, o.buf, (size_t)w * h * sizeof(float));
    img_free(&o);
}
void akzo_vertical_filter(const float *in, int w, int h, const float *kernel, int ksize, float *out) {
    img_t s = {(float *)in, w, h};
    img_t o = vertical_filter(&s, kernel, ksize);
    memcpy(out, o.buf, (size_t)w * h * sizeof(float));
    img_free(&o);
}

/* types/image.rs:341-343 -- all f32 */
static float gaussian(float x, float r) {
    float pi = 3.14159265358979323846f;
    float a = 1.0f / (sqrtf(2.0f * pi) * r);
    return a * expf(-(x * x) / (2.0f * (r * r)));
}
/* types/image.rs:352-365 */
void akzo_gaussian_kernel(float r, int kernel_size, float *kernel) {
    int hw = kernel_size / 2;
    float sum = 0.0f;
    for (int i = 0; i < kernel_size; i++) kernel[i] = 0.0f;
    for (int i = -hw; i <= hw; i++) {
        float val = gaussian((float)i, r);
        kernel[i + hw] = val;
        sum += val;
    }
    for (int i = 0; i < kernel_size; i++) kernel[i] /= sum;
}
/* types/image.rs:374-380 */
static img_t gaussian_blur(const img_t *image, float r) {
    int kernel_size = (int)f32_as_usize(ceilf(r)) * 2 + 1;
    float kernel[64];
    akzo_gaussian_kernel(r, kernel_size, kernel);
    img_t hz = horizontal_filter(image, kernel, kernel_size);
    img_t out = vertical_filter(&hz, kernel, kernel_size);
    img_free(&hz);
    return out;
}
void akzo_gaussian_blur(const float *in, int w, int h, float r, float *out) {
    img_t s = {(float *)in, w, h};
    img_t o = gaussian_blur(&s, r);
    memcpy(out, o.buf, (size_t)w * h * sizeof(float));
    img_free(&o);
}

/* ------------------------------------------------------------------------------------------------
 * Scharr (ops/derivatives.rs)
 * ---------------------------------------------------------------------------------------------- */
/* ops/derivatives.rs:74-82 */
void akzo_scharr_off_axis_kernel(uint32_t scale, float *kernel) {
    size_t size = 3 + 2 * (size_t)(scale - 1);
    for (size_t i = 0; i < size; i++) kernel[i] = 0.0f;
    kernel[0] = -1.0f;
    kernel[size / 2] = 0.0f;
    kernel[size - 1] = 1.0f;
}
/* ops/derivatives.rs:91-101 -- w and norm are f64, cast per tap */
void akzo_scharr_main_axis_kernel(uint32_t scale, float *kernel) {
    size_t size = 3 + 2 * (size_t)(scale - 1);
    double w = 10.0 / 3.0;
    double norm = 1.0 / (2.0 * (double)scale * (w + 2.0));
    for (size_t i = 0; i < size; i++) kernel[i] = 0.0f;
    kernel[0] = (float)norm;
    kernel[size / 2] = (float)(w * norm);
    kernel[size - 1] = (float)norm;
}
/* ops/derivatives.rs:41-47 -- "horizontal": main-axis kernel along x, difference kernel along y */
static img_t scharr_horizontal(const img_t *image, uint32_t s) {
    float km[64], ko[64];
    int size = 3 + 2 * (int)(s - 1);
    akzo_scharr_main_axis_kernel(s, km);
    akzo_scharr_off_axis_kernel(s, ko);
    img_t hz = horizontal_filter(image, km, size);
    img_t out = vertical_filter(&hz, ko, size);
    img_free(&hz);
    return out;
}
/* ops/derivatives.rs:59-65 */
static img_t scharr_vertical(const img_t *image, uint32_t s) {
    float km[64], ko[64];
    int size = 3 + 2 * (int)(s - 1);
    akzo_scharr_main_axis_kernel(s, km);
    akzo_scharr_off_axis_kernel(s, ko);
    img_t hz = horizontal_filter(image, ko, size);
    img_t out = vertical_filter(&hz, km, size);
    img_free(&hz);
    return out;
}
/* ops/derivatives.rs:112-130 (the x&&y branch is never taken on the hot path: sqrt_squared of two
 * horizontal derivatives, image.rs:212-231) */
static img_t scharr(const img_t *image, int x_order, int y_order, uint32_t s) {
    if (x_order && y_order) {
        img_t hz = scharr_horizontal(image, s);
        img_t vt = scharr_horizontal(image, s);
        size_t n = (size_t)image->w * image->h;
        for (size_t i = 0; i < n; i++) vt.buf[i] = sqrtf(vt.buf[i] * vt.buf[i] + hz.buf[i] * hz.buf[i]);
        img_free(&hz);
        return vt;
    } else if (x_order) {
        return scharr_horizontal(image, s);
    } else if (y_order) {
        return scharr_vertical(image, s);
    }
    return img_new(image->w, image->h);
}
void akzo_scharr(const float *in, int w, int h, int x_order, int y_order, uint32_t s, float *out) {
    img_t src = {(float *)in, w, h};
    img_t o = scharr(&src, x_order, y_order, s);
    memcpy(out, o.buf, (size_t)w * h * sizeof(float));
    img_free(&o);
}

/* ------------------------------------------------------------------------------------------------
 * Contrast factor (ops/contrast_factor.rs:18-71), all f64
 * ---------------------------------------------------------------------------------------------- */
static double compute_contrast_factor(const img_t *image, double percentile, double hist_scale,
                                      uint64_t num_bins) {
    double num_points = 0.0, hmax = 0.0;
    double *histogram = (double *)calloc(num_bins ? num_bins : 1, sizeof(double));
    img_t g = gaussian_blur(image, (float)hist_scale);
    img_t Lx = scharr(&g, 1, 0, 1);
    img_t Ly = scharr(&g, 0, 1, 1);
    for (int y = 1; y < g.h - 1; y++)
        for (int x = 1; x < g.w - 1; x++) {
            double lx = (double)img_get(&Lx, x, y), ly = (double)img_get(&Ly, x, y);
            double modg = sqrt(lx * lx + ly * ly);
            if (modg > hmax) hmax = modg;
        }
    for (int y = 1; y < g.h - 1; y++)
        for (int x = 1; x < g.w - 1; x++) {
            double lx = (double)img_get(&Lx, x, y), ly = (double)img_get(&Ly, x, y);
            double modg = sqrt(lx * lx + ly * ly);
            if (modg != 0.0) {
                double b = floor((double)num_bins * (modg / hmax));
                uint64_t bin = (b > 0.0) ? (uint64_t)b : 0;
                if (bin == num_bins) bin -= 1;
                histogram[bin] += 1.0;
                num_points += 1.0;
            }
        }
    uint64_t threshold = (uint64_t)(num_points * percentile);
    uint64_t k = 0, num_elements = 0;
    while (num_elements < threshold && k < num_bins) {
        num_elements += (uint64_t)histogram[k];
        k += 1;
    }
    img_free(&g);
    img_free(&Lx);
    img_free(&Ly);
    free(histogram);
    if (num_elements >= threshold) return hmax * (double)k / (double)num_bins;
    return 0.03;
}
double akzo_compute_contrast_factor(const float *img, int w, int h, double percentile, double scale,
                                    uint64_t nbins) {
    img_t s = {(float *)img, w, h};
    return compute_contrast_factor(&s, percentile, scale, nbins);
}

/* lib.rs:26-41 -- f64 island, cast at the end */
static img_t pm_g2(const img_t *Lx, const img_t *Ly, double k) {
    img_t dst = img_new(Lx->w, Lx->h);
    double inverse_k = 1.0 / (k * k);
    for (int y = 0; y < Lx->h; y++)
        for (int x = 0; x < Lx->w; x++) {
            double lx = (double)img_get(Lx, x, y), ly = (double)img_get(Ly, x, y);
            double d = 1.0 / (1.0 + inverse_k * (lx * lx + ly * ly));
            img_put(&dst, x, y, (float)d);
        }
    return dst;
}
void akzo_pm_g2(const float *lx, const float *ly, size_t n, double k, float *out) {
    double inverse_k = 1.0 / (k * k);
    for (size_t i = 0; i < n; i++) {
        double a = (double)lx[i], b = (double)ly[i];
        out[i] = (float)(1.0 / (1.0 + inverse_k * (a * a + b * b)));
    }
}

/* ------------------------------------------------------------------------------------------------
 * FED schedule (ops/fed_tau.rs)
 * ---------------------------------------------------------------------------------------------- */
static int is_prime_u64(uint64_t n) { /* primal::is_prime, restated */
    if (n < 2) return 0;
    for (uint64_t d = 2; d * d <= n; d++)
        if (n % d == 0) return 0;
    return 1;
}
/* ops/fed_tau.rs:61-106. The reference's `index` is a usize: ((k+1)*kappa) % prime - 1 wraps to
 * usize::MAX when the product is a multiple of prime and is then skipped by `index >= n`
 * (release build); signed arithmetic with index<0 treated as "skip" is the same thing. n==1 never
 * terminates in the reference (kappa==0); we return -1. */
static int fed_tau_internal(size_t n, double scale, double tau_max, int reordering, double *tau, int cap) {
    if (n == 0) return 0;
    if ((int)n > cap) return -2;
    double *tauh = (double *)calloc(n, sizeof(double));
    double c = 1.0 / (4.0 * (double)n + 2.0);
    double d = scale * tau_max / 2.0;
    const double PI = 3.14159265358979323846264338327950288;
    for (size_t k = 0; k < n; k++) {
        double hh = cos(PI * (2.0 * (double)k + 1.0) * c);
        if (reordering) tauh[k] = d / (hh * hh);
        else tau[k] = d / (hh * hh);
    }
    if (reordering) {
        size_t kappa = n / 2;
        size_t prime = n + 1;
        if (kappa == 0) {
            free(tauh);
            return -1;
        }
        while (!is_prime_u64(prime)) prime += 1;
        size_t k = 0;
        for (size_t t = 0; t < n; t++) {
            long long index = (long long)(((k + 1) * kappa) % prime) - 1;
            while (index < 0 || (size_t)index >= n) {
                k += 1;
                index = (long long)(((k + 1) * kappa) % prime) - 1;
            }
            tau[t] = tauh[index];
            k += 1;
        }
    }
    free(tauh);
    return (int)n;
}
/* ops/fed_tau.rs:43-49 */
static int fed_tau_by_cycle_time(double t, double tau_max, int reordering, double *out, int cap) {
    double nf = ceil(sqrt(3.0 * t / tau_max + 0.25) - 0.5 - 1.0e-8) + 0.5;
    size_t n = (nf > 0.0) ? (size_t)nf : 0;
    double scale = 3.0 * t / (tau_max * (double)(n * (n + 1)));
    return fed_tau_internal(n, scale, tau_max, reordering, out, cap);
}
/* ops/fed_tau.rs:27-30 */
int akzo_fed_tau_by_process_time(double T, int M, double tau_max, int reordering, double *out, int cap) {
    return fed_tau_by_cycle_time(T / (double)M, tau_max, reordering, out, cap);
}

/* ------------------------------------------------------------------------------------------------
 * EvolutionStep (types/evolution.rs:59-161)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    double etime, esigma;
    uint32_t octave, sublevel, sigma_size;
    img_t Lt, Lsmooth, Lx, Ly, Lxx, Lyy, Lxy, Lflow, Lstep, Ldet;
    double *fed_tau_steps;
    int n_steps;
} evo_t;

struct akzo_result {
    int status;
    akzo_config cfg;
    evo_t *evo;
    uint32_t n_levels;
    double contrast_factor;
    uint64_t n_candidates, n_cache;
    akzo_keypoint *kps;
    uint64_t n_kps;
    uint8_t *desc;
    uint64_t desc_len;
};

/* types/evolution.rs:101-126 */
static evo_t evo_new(uint32_t octave, uint32_t sublevel, const akzo_config *o) {
    evo_t e;
    memset(&e, 0, sizeof(e));
    e.esigma = o->base_scale_offset * pow(2.0, (double)sublevel / (double)o->num_sublevels + (double)octave);
    e.etime = 0.5 * (e.esigma * e.esigma);
    e.octave = octave;
    e.sublevel = sublevel;
    e.sigma_size = (uint32_t)round(e.esigma);
    return e;
}

/* types/evolution.rs:135-161 */
static int allocate_evolutions(akzo_result *r, uint32_t width, uint32_t height) {
    const akzo_config *o = &r->cfg;
    uint32_t cap = o->max_octave_evolution * o->num_sublevels;
    r->evo = (evo_t *)calloc(cap ? cap : 1, sizeof(evo_t));
    r->n_levels = 0;
    for (uint32_t i = 0; i < o->max_octave_evolution; i++) {
        double rfactor = 1.0 / pow(2.0, (double)i);
        uint32_t level_height = (uint32_t)((double)height * rfactor);
        uint32_t level_width = (uint32_t)((double)width * rfactor);
        if ((level_width >= 80 && level_height >= 40) || i == 0) {
            for (uint32_t j = 0; j < o->num_sublevels; j++) r->evo[r->n_levels++] = evo_new(i, j, o);
        } else {
            break;
        }
    }
    for (uint32_t i = 1; i < r->n_levels; i++) {
        double ttime = r->evo[i].etime - r->evo[i - 1].etime;
        double tmp[4096];
        int n = akzo_fed_tau_by_process_time(ttime, 1, 0.25, 1, tmp, 4096);
        if (n < 0) return -1;
        r->evo[i].fed_tau_steps = (double *)malloc(sizeof(double) * (n ? n : 1));
        memcpy(r->evo[i].fed_tau_steps, tmp, sizeof(double) * n);
        r->evo[i].n_steps = n;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Nonlinear diffusion step (ops/nonlinear_diffusion.rs:15-173)
 * ---------------------------------------------------------------------------------------------- */
/* :149-173 -- (c[x+px0,y+py0] + c[x+px1,y+py1]) * (Ld[x+px2,y+py2] - Ld[x+px3,y+py3]) */
static inline float nd_eval(const img_t *c, const img_t *Ld, int x, int y, const int px[4], const int py[4]) {
    float c0 = img_get(c, (size_t)(x + px[0]), (size_t)(y + py[0]));
    float c1 = img_get(c, (size_t)(x + px[1]), (size_t)(y + py[1]));
    float l0 = img_get(Ld, (size_t)(x + px[2]), (size_t)(y + py[2]));
    float l1 = img_get(Ld, (size_t)(x + px[3]), (size_t)(y + py[3]));
    return (c0 + c1) * (l0 - l1);
}

static void calculate_step(img_t *Ld, const img_t *c, img_t *Lstep, double step_size) {
    static const int Z[4] = {0, 0, 0, 0};
    static const int P[4] = {0, 1, 1, 0};
    static const int N[4] = {-1, 0, 0, -1};
    static const int PM[4] = {0, -1, -1, 0};
    int w = Lstep->w, h = Lstep->h;
    int xend = w - 1, yend = h - 1;
    float st = (float)step_size;
    /* middle (:30-81) */
    for (int y = 1; y < h - 1; y++) {
        const float *Lrow = Ld->buf + (size_t)w * y;
        const float *Lup = Ld->buf + (size_t)w * (y - 1);
        const float *Ldn = Ld->buf + (size_t)w * (y + 1);
        const float *crow = c->buf + (size_t)w * y;
        const float *cup = c->buf + (size_t)w * (y - 1);
        const float *cdn = c->buf + (size_t)w * (y + 1);
        float *srow = Lstep->buf + (size_t)w * y;
        for (int x = 1; x < w - 1; x++) {
            float x_pos = (crow[x] + crow[x + 1]) * (Lrow[x + 1] - Lrow[x]);
            float x_neg = (crow[x - 1] + crow[x]) * (Lrow[x] - Lrow[x - 1]);
            float y_pos = (crow[x] + cdn[x]) * (Ldn[x] - Lrow[x]);
            float y_neg = (cup[x] + crow[x]) * (Lrow[x] - Lup[x]);
            srow[x] = 0.5f * st * (x_pos - x_neg + y_pos - y_neg);
        }
    }
    /* first row (:83-102) */
    for (int x = 1; x < w - 1; x++) {
        float x_pos = nd_eval(c, Ld, x, 0, P, Z);
        float y_pos = nd_eval(c, Ld, x, 0, Z, P);
        float x_neg = nd_eval(c, Ld, x, 0, N, Z);
        img_put(Lstep, x, 0, 0.5f * st * (x_pos - x_neg + y_pos));
    }
    {
        float x_pos = nd_eval(c, Ld, 0, 0, P, Z);
        float y_pos = nd_eval(c, Ld, 0, 0, Z, P);
        img_put(Lstep, 0, 0, 0.5f * st * (x_pos + y_pos));
    }
    {
        float y_pos = nd_eval(c, Ld, xend, 0, Z, P);
        float x_neg = nd_eval(c, Ld, xend, 0, N, Z);
        img_put(Lstep, xend, 0, 0.5f * st * (-x_neg + y_pos));
    }
    /* last row (:104-119) */
    for (int x = 1; x < w - 1; x++) {
        float x_pos = nd_eval(c, Ld, x, yend, P, Z);
        float y_pos = nd_eval(c, Ld, x, yend, Z, PM);
        float x_neg = nd_eval(c, Ld, x, yend, N, Z);
        img_put(Lstep, x, yend, 0.5f * st * (x_pos - x_neg + y_pos));
    }
    {
        float x_pos = nd_eval(c, Ld, 0, yend, P, Z);
        float y_pos = nd_eval(c, Ld, 0, yend, Z, PM);
        img_put(Lstep, 0, yend, 0.5f * st * (x_pos + y_pos));
    }
    {
        float y_pos = nd_eval(c, Ld, xend, yend, Z, PM);
        float x_neg = nd_eval(c, Ld, xend, yend, N, Z);
        img_put(Lstep, xend, yend, 0.5f * st * (-x_neg + y_pos));
    }
    /* first and last columns (:121-138) */
    for (int y = 1; y < h - 1; y++) {
        {
            float x_pos = nd_eval(c, Ld, 0, y, P, Z);
            float y_pos = nd_eval(c, Ld, 0, y, Z, P);
            float y_neg = nd_eval(c, Ld, 0, y, Z, N);
            img_put(Lstep, 0, y, 0.5f * st * (x_pos + y_pos - y_neg));
        }
        {
            float y_pos = nd_eval(c, Ld, xend, y, Z, P);
            float x_neg = nd_eval(c, Ld, xend, y, N, Z);
            float y_neg = nd_eval(c, Ld, xend, y, Z, N);
            img_put(Lstep, xend, y, 0.5f * st * (-x_neg + y_pos - y_neg));
        }
    }
    /* :140-143 */
    size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; i++) Ld->buf[i] += Lstep->buf[i];
}
void akzo_calculate_step(float *lt, const float *lflow, float *lstep, int w, int h, double step_size) {
    img_t a = {lt, w, h}, b = {(float *)lflow, w, h}, s = {lstep, w, h};
    calculate_step(&a, &b, &s, step_size);
}

/* ------------------------------------------------------------------------------------------------
 * Scale space (lib.rs:49-120)
 * ---------------------------------------------------------------------------------------------- */
static void create_nonlinear_scale_space(akzo_result *r, const img_t *image) {
    const akzo_config *o = &r->cfg;
    evo_t *ev = r->evo;
    ev[0].Lt = gaussian_blur(image, (float)o->base_scale_offset);
    ev[0].Lsmooth = img_clone(&ev[0].Lt);
    double contrast_factor =
        compute_contrast_factor(&ev[0].Lsmooth, o->contrast_percentile, 1.0, o->contrast_factor_num_bins);
    r->contrast_factor = contrast_factor;
    for (uint32_t i = 1; i < r->n_levels; i++) {
        if (ev[i].octave > ev[i - 1].octave) {
            ev[i].Lt = half_size(&ev[i - 1].Lt);
            contrast_factor *= 0.75;
        } else {
            ev[i].Lt = img_clone(&ev[i - 1].Lt);
        }
        ev[i].Lsmooth = gaussian_blur(&ev[i].Lt, 1.0f);
        ev[i].Lx = scharr(&ev[i].Lsmooth, 1, 0, 1);
        ev[i].Ly = scharr(&ev[i].Lsmooth, 0, 1, 1);
        ev[i].Lflow = pm_g2(&ev[i].Lx, &ev[i].Ly, contrast_factor);
        ev[i].Lstep = img_new(ev[i].Lt.w, ev[i].Lt.h);
        for (int j = 0; j < ev[i].n_steps; j++)
            calculate_step(&ev[i].Lt, &ev[i].Lflow, &ev[i].Lstep, ev[i].fed_tau_steps[j]);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Detector response (ops/detector_response.rs)
 * ---------------------------------------------------------------------------------------------- */
static uint32_t detector_sigma_size(const evo_t *e, const akzo_config *o) {
    double ratio = pow(2.0, (double)e->octave);
    return (uint32_t)round(e->esigma * o->derivative_factor / ratio);
}
/* :8-14 */
static void multiscale_derivatives_for_evolution(evo_t *e, uint32_t s) {
    img_t t;
    t = scharr(&e->Lsmooth, 1, 0, s); img_free(&e->Lx); e->Lx = t;
    t = scharr(&e->Lsmooth, 0, 1, s); img_free(&e->Ly); e->Ly = t;
    t = scharr(&e->Lx, 1, 0, s); img_free(&e->Lxx); e->Lxx = t;
    t = scharr(&e->Ly, 0, 1, s); img_free(&e->Lyy); e->Lyy = t;
    t = scharr(&e->Lx, 0, 1, s); img_free(&e->Lxy); e->Lxy = t;
}
typedef struct {
    akzo_result *r;
    volatile int next;
    pthread_mutex_t mu;
} deriv_pool_t;
static void *deriv_worker(void *arg) {
    deriv_pool_t *p = (deriv_pool_t *)arg;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        int i = p->next++;
        pthread_mutex_unlock(&p->mu);
        if (i >= (int)p->r->n_levels) break;
        multiscale_derivatives_for_evolution(&p->r->evo[i], detector_sigma_size(&p->r->evo[i], &p->r->cfg));
    }
    return NULL;
}
/* :16-29 (scoped thread pool, one job per level) and :38-55 */
static void detector_response(akzo_result *r, int threads) {
    if (threads <= 1) {
        for (uint32_t i = 0; i < r->n_levels; i++)
            multiscale_derivatives_for_evolution(&r->evo[i], detector_sigma_size(&r->evo[i], &r->cfg));
    } else {
        deriv_pool_t pool;
        pool.r = r;
        pool.next = 0;
        pthread_mutex_init(&pool.mu, NULL);
        if (threads > 64) threads = 64;
        pthread_t th[64];
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, deriv_worker, &pool);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
        pthread_mutex_destroy(&pool.mu);
    }
    for (uint32_t i = 0; i < r->n_levels; i++) {
        evo_t *e = &r->evo[i];
        uint32_t s = detector_sigma_size(e, &r->cfg);
        uint32_t quat = s * s * s * s;
        float q = (float)quat;
        img_free(&e->Ldet);
        e->Ldet = img_new(e->Lxx.w, e->Lxx.h);
        size_t n = (size_t)e->Lxx.w * e->Lxx.h;
        for (size_t k = 0; k < n; k++)
            e->Ldet.buf[k] = ((e->Lxx.buf[k] * e->Lyy.buf[k]) - (e->Lxy.buf[k] * e->Lxy.buf[k])) * q;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Extrema, "sub-pixel", orientation (ops/scale_space_extrema.rs)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    akzo_keypoint *v;
    size_t n, cap;
} kpvec_t;
static void kp_push(kpvec_t *a, akzo_keypoint k) {
    if (a->n == a->cap) {
        a->cap = a->cap ? a->cap * 2 : 1024;
        a->v = (akzo_keypoint *)realloc(a->v, a->cap * sizeof(akzo_keypoint));
    }
    a->v[a->n++] = k;
}

/* :12-132 */
static kpvec_t find_scale_space_extrema(akzo_result *r) {
    const akzo_config *o = &r->cfg;
    kpvec_t cache = {0, 0, 0};
    float smax = 10.0f * sqrtf(2.0f);
    float thr = (float)o->detector_threshold;
    r->n_candidates = 0;
    for (uint32_t e_id = 0; e_id < r->n_levels; e_id++) {
        evo_t *ev = &r->evo[e_id];
        size_t w = (size_t)ev->Ldet.w, h = (size_t)ev->Ldet.h;
        const float *b = ev->Ldet.buf;
        size_t len = w * h;
        for (size_t i = w + 1; i + w + 1 < len; i++) {
            size_t x = i % w, y = i / w;
            float x_i = b[i], x_p = b[i + 1], x_m = b[i - 1], y_m = b[i - w], y_p = b[i + w];
            if (x != 0 && x != w && x_i > thr && x_i > x_p && x_i > x_m && x_i > y_m && x_i > y_p) {
                r->n_candidates++;
                akzo_keypoint kp;
                kp.response = fabsf(x_i);
                kp.size = (float)(ev->esigma * o->derivative_factor);
                kp.octave = ev->octave;
                kp.class_id = e_id;
                kp.x = (float)x;
                kp.y = (float)y;
                kp.angle = 0.0f;
                float ratio = powf(2.0f, (float)ev->octave);
                float sigma_size = roundf(kp.size / ratio);
                size_t id_repeated = 0;
                int is_repeated = 0, is_extremum = 1;
                for (size_t k = 0; k < cache.n; k++) {
                    const akzo_keypoint *pk = &cache.v[k];
                    if (kp.class_id == pk->class_id || (kp.class_id != 0 && kp.class_id - 1 == pk->class_id)) {
                        float dist = (kp.x * ratio - pk->x) * (kp.x * ratio - pk->x) +
                                     (kp.y * ratio - pk->y) * (kp.y * ratio - pk->y);
                        if (dist <= kp.size * kp.size) {
                            if (kp.response > pk->response) {
                                id_repeated = k;
                                is_repeated = 1;
                            } else {
                                is_extremum = 0;
                            }
                            break;
                        }
                    }
                }
                if (is_extremum) {
                    float left_x = roundf(kp.x - smax * sigma_size) - 1.0f;
                    float right_x = roundf(kp.x + smax * sigma_size) + 1.0f;
                    float up_y = roundf(kp.y - smax * sigma_size) - 1.0f;
                    float down_y = roundf(kp.y + smax * sigma_size) + 1.0f;
                    int is_out = left_x < 0.0f || right_x >= (float)w || up_y < 0.0f || down_y >= (float)h;
                    if (!is_out) {
                        kp.x = kp.x * ratio + 0.5f * (ratio - 1.0f);
                        kp.y = kp.y * ratio + 0.5f * (ratio - 1.0f);
                        if (!is_repeated) kp_push(&cache, kp);
                        else cache.v[id_repeated] = kp;
                    }
                }
            }
        }
    }
    r->n_cache = cache.n;
    /* :111-129 upper-scale filter; note the inner scan starts at slot i */
    kpvec_t out = {0, 0, 0};
    for (size_t i = 0; i < cache.n; i++) {
        int is_repeated = 0;
        akzo_keypoint ki = cache.v[i];
        for (size_t j = i; j < cache.n; j++) {
            const akzo_keypoint *kj = &cache.v[j];
            if (ki.class_id + 1 == kj->class_id) {
                float dist = (ki.x - kj->x) * (ki.x - kj->x) + (ki.y - kj->y) * (ki.y - kj->y);
                if (dist <= ki.size * ki.size) {
                    is_repeated = 1;
                    break;
                }
            }
        }
        if (!is_repeated) kp_push(&out, ki);
    }
    free(cache.v);
    return out;
}

/* :207-271 (numeric table from the reference; data, not code) */
static const float GAUSS25[7][7] = {
    {0.02546481f, 0.02350698f, 0.01849125f, 0.01239505f, 0.00708017f, 0.00344629f, 0.00142946f},
    {0.02350698f, 0.02169968f, 0.01706957f, 0.01144208f, 0.00653582f, 0.00318132f, 0.00131956f},
    {0.01849125f, 0.01706957f, 0.01342740f, 0.00900066f, 0.00514126f, 0.00250252f, 0.00103800f},
    {0.01239505f, 0.01144208f, 0.00900066f, 0.00603332f, 0.00344629f, 0.00167749f, 0.00069579f},
    {0.00708017f, 0.00653582f, 0.00514126f, 0.00344629f, 0.00196855f, 0.00095820f, 0.00039744f},
    {0.00344629f, 0.00318132f, 0.00250252f, 0.00167749f, 0.00095820f, 0.00046640f, 0.00019346f},
    {0.00142946f, 0.00131956f, 0.00103800f, 0.00069579f, 0.00039744f, 0.00019346f, 0.00008024f},
};

/* :274-329 -- atan2(res_y,res_y) and the never-reset running sums are reproduced as written */
static int compute_main_orientation(akzo_keypoint *kp, const evo_t *evo) {
    float res_x[109], res_y[109], angs[109];
    static const int id[13] = {6, 5, 4, 3, 2, 1, 0, 1, 2, 3, 4, 5, 6};
    const float PI = 3.14159265358979323846f;
    const evo_t *e = &evo[kp->class_id];
    float ratio = (float)(1u << e->octave);
    float s = roundf(0.5f * kp->size / ratio);
    float xf = kp->x / ratio, yf = kp->y / ratio;
    int idx = 0;
    size_t npx = (size_t)e->Lx.w * e->Lx.h;
    for (int i = -6; i <= 6; i++)
        for (int j = -6; j <= 6; j++)
            if (i * i + j * j < 36) {
                size_t iy = f32_as_usize(roundf(yf + (float)j * s));
                size_t ix = f32_as_usize(roundf(xf + (float)i * s));
                float gweight = GAUSS25[id[i + 6]][id[j + 6]];
                size_t at = (size_t)e->Lx.w * iy + ix;
                if (at >= npx) return -1; /* the reference would panic (index out of bounds) */
                res_x[idx] = gweight * e->Lx.buf[at];
                res_y[idx] = gweight * e->Ly.buf[at];
                angs[idx] = atan2f(res_y[idx], res_y[idx]);
                idx++;
            }
    float ang1 = 0.0f, sum_x = 0.0f, sum_y = 0.0f, max = 0.0f;
    while (ang1 < 2.0f * PI) {
        float ang2 = (ang1 + PI / 3.0f > 2.0f * PI) ? (ang1 - 5.0f * PI / 3.0f) : (ang1 + PI / 3.0f);
        ang1 += 0.15f;
        for (int k = 0; k < 109; k++) {
            float ang = angs[k];
            if ((ang1 < ang2 && ang1 < ang && ang < ang2) ||
                (ang2 < ang1 && ((ang > 0.0f && ang < ang2) || (ang > ang1 && ang < 2.0f * PI)))) {
                sum_x += res_x[k];
                sum_y += res_y[k];
            }
        }
        float val = sum_x * sum_x + sum_y * sum_y;
        if (val > max) {
            max = val;
            kp->angle = atan2f(sum_y, sum_x);
        }
    }
    return 0;
}

/* :141-189 -- lu.solve(&b) returns a value that is dropped; b keeps (-d_x,-d_y) */
static int do_subpixel_refinement(akzo_result *r, const kpvec_t *in, kpvec_t *result) {
    for (size_t n = 0; n < in->n; n++) {
        const akzo_keypoint *kp = &in->v[n];
        float ratio = powf(2.0f, (float)kp->octave);
        size_t x = f32_as_usize(roundf(kp->x / ratio));
        size_t y = f32_as_usize(roundf(kp->y / ratio));
        const img_t *D = &r->evo[kp->class_id].Ldet;
        float x_p = img_get(D, x + 1, y), x_m = img_get(D, x - 1, y);
        float y_p = img_get(D, x, y + 1), y_m = img_get(D, x, y - 1);
        float d_x = 0.5f * (x_p - x_m);
        float d_y = 0.5f * (y_p - y_m);
        float b0 = -d_x, b1 = -d_y;
        if (fabsf(b0) <= 1.0f && fabsf(b1) <= 1.0f) {
            akzo_keypoint c = *kp;
            c.x = (float)x + b0;
            c.y = (float)y + b1;
            c.x = c.x * ratio + 0.5f * (ratio - 1.0f);
            c.y = c.y * ratio + 0.5f * (ratio - 1.0f);
            kp_push(result, c);
        }
    }
    for (size_t n = 0; n < result->n; n++)
        if (compute_main_orientation(&result->v[n], r->evo) != 0) return -1;
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * MLDB descriptor (ops/descriptors.rs)
 * ---------------------------------------------------------------------------------------------- */
/* :87-151 */
static int mldb_fill_values(float *values, size_t sample_step, size_t level, float xf, float yf, float co,
                            float si, float scale, const akzo_config *o, const evo_t *evo) {
    int pattern_size = (int)o->descriptor_pattern_size;
    size_t nr_channels = (size_t)o->descriptor_channels;
    size_t valuepos = 0;
    const evo_t *e = &evo[level];
    size_t W = (size_t)e->Lt.w, H = (size_t)e->Lt.h;
    for (int i = -pattern_size; i < pattern_size; i += (int)sample_step)
        for (int j = -pattern_size; j < pattern_size; j += (int)sample_step) {
            float di = 0.0f, dx = 0.0f, dy = 0.0f;
            size_t nsamples = 0;
            for (int k = i; k < i + (int)sample_step; k++)
                for (int l = j; l < j + (int)sample_step; l++) {
                    float lf = (float)l + 0.5f;
                    float kf = (float)k + 0.5f;
                    float sample_y = yf + (lf * co * scale + kf * si * scale);
                    float sample_x = xf + (-lf * si * scale + kf * co * scale);
                    long long y1 = (long long)roundf(sample_y);
                    long long x1 = (long long)roundf(sample_x);
                    if (y1 < 0 || x1 < 0 || (size_t)x1 >= W || (size_t)y1 >= H) {
                        /* x1 >= W would silently alias the next row in the reference unless past the
                         * buffer end; we flag all of these so that parity cases stay in-bounds */
                        return -1;
                    }
                    float ri = img_get(&e->Lt, (size_t)x1, (size_t)y1);
                    di += ri;
                    if (nr_channels > 1) {
                        float rx = img_get(&e->Lx, (size_t)x1, (size_t)y1);
                        float ry = img_get(&e->Ly, (size_t)x1, (size_t)y1);
                        if (nr_channels == 2) {
                            dx += sqrtf(rx * rx + ry * ry);
                        } else {
                            float rry = rx * co + ry * si;
                            float rrx = -rx * si + ry * co;
                            dx += rrx;
                            dy += rry;
                        }
                    }
                    nsamples += 1;
                }
            di /= (float)nsamples;
            dx /= (float)nsamples;
            dy /= (float)nsamples;
            values[valuepos] = di;
            if (nr_channels > 1) values[valuepos + 1] = dx;
            if (nr_channels > 2) values[valuepos + 2] = dy;
            valuepos += nr_channels;
        }
    return 0;
}
/* :154-175 */
static void mldb_binary_comparisons(const float *values, uint8_t *descriptor, size_t count, size_t *dpos,
                                    size_t nr_channels) {
    for (size_t pos = 0; pos < nr_channels; pos++)
        for (size_t i = 0; i < count; i++) {
            float ival = values[nr_channels * i + pos];
            for (size_t j = i + 1; j < count; j++) {
                uint8_t res = (ival > values[nr_channels * j + pos]) ? 1 : 0;
                descriptor[*dpos >> 3] |= (uint8_t)(res << (*dpos & 7));
                *dpos += 1;
            }
        }
}
/* :37-83 */
static int get_mldb_descriptor(const akzo_keypoint *kp, const evo_t *evo, const akzo_config *o, uint8_t *out) {
    float values[16 * 3];
    const float size_mult[3] = {1.0f, 2.0f / 3.0f, 1.0f / 2.0f};
    float ratio = (float)(1u << kp->octave);
    float scale = roundf(0.5f * kp->size / ratio);
    float xf = kp->x / ratio, yf = kp->y / ratio;
    float co = cosf(kp->angle), si = sinf(kp->angle);
    size_t dpos = 0;
    float pattern_size = (float)o->descriptor_pattern_size;
    memset(values, 0, sizeof(values));
    for (size_t lvl = 0; lvl < 3; lvl++) {
        size_t val_count = (lvl + 2) * (lvl + 2);
        size_t sample_size = f32_as_usize(ceilf(pattern_size * size_mult[lvl]));
        if (mldb_fill_values(values, sample_size, kp->class_id, xf, yf, co, si, scale, o, evo) != 0) return -1;
        mldb_binary_comparisons(values, out, val_count, &dpos, (size_t)o->descriptor_channels);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * extract_features (lib.rs:167-194), after create_unit_float_image
 * ---------------------------------------------------------------------------------------------- */
akzo_result *akzo_extract(const float *unit_gray, uint32_t w, uint32_t h, const akzo_config *cfg, int threads,
                          int stop_after) {
    akzo_result *r = (akzo_result *)calloc(1, sizeof(akzo_result));
    if (!r) return NULL;
    r->cfg = *cfg;
    if (allocate_evolutions(r, w, h) != 0) {
        r->status = -1;
        return r;
    }
    img_t image = {(float *)unit_gray, (int)w, (int)h};
    create_nonlinear_scale_space(r, &image);
    if (stop_after == 1) return r;
    detector_response(r, threads);
    if (stop_after == 2) return r;
    kpvec_t cand = find_scale_space_extrema(r);
    kpvec_t kps = {0, 0, 0};
    if (do_subpixel_refinement(r, &cand, &kps) != 0) r->status = -2;
    free(cand.v);
    r->kps = kps.v;
    r->n_kps = kps.n;
    /* descriptors.rs:14-27,42-46 */
    size_t t = (6 + 36 + 120) * (size_t)cfg->descriptor_channels;
    r->desc_len = (t + 7) / 8;
    r->desc = (uint8_t *)calloc(r->n_kps * r->desc_len + 1, 1);
    if (r->status == 0)
        for (size_t i = 0; i < r->n_kps; i++)
            if (get_mldb_descriptor(&r->kps[i], r->evo, cfg, r->desc + i * r->desc_len) != 0) {
                r->status = -3;
                break;
            }
    return r;
}

void akzo_result_free(akzo_result *r) {
    if (!r) return;
    for (uint32_t i = 0; i < r->n_levels; i++) {
        evo_t *e = &r->evo[i];
        img_free(&e->Lt); img_free(&e->Lsmooth); img_free(&e->Lx); img_free(&e->Ly); img_free(&e->Lxx);
        img_free(&e->Lyy); img_free(&e->Lxy); img_free(&e->Lflow); img_free(&e->Lstep); img_free(&e->Ldet);
        free(e->fed_tau_steps);
    }
    free(r->evo);
    free(r->kps);
    free(r->desc);
    free(r);
}
int akzo_result_status(const akzo_result *r) { return r->status; }
uint32_t akzo_result_num_levels(const akzo_result *r) { return r->n_levels; }
void akzo_result_level_info(const akzo_result *r, uint32_t level, akzo_level_info *out) {
    const evo_t *e = &r->evo[level];
    out->octave = e->octave;
    out->sublevel = e->sublevel;
    out->sigma_size = e->sigma_size;
    out->width = (uint32_t)e->Lt.w;
    out->height = (uint32_t)e->Lt.h;
    out->n_steps = (uint32_t)e->n_steps;
    out->esigma = e->esigma;
    out->etime = e->etime;
}
const double *akzo_result_fed_tau(const akzo_result *r, uint32_t level) { return r->evo[level].fed_tau_steps; }
const float *akzo_result_image(const akzo_result *r, uint32_t level, int kind) {
    const evo_t *e = &r->evo[level];
    const img_t *im[10] = {&e->Lt, &e->Lsmooth, &e->Lx, &e->Ly, &e->Lxx, &e->Lyy, &e->Lxy, &e->Lflow, &e->Lstep, &e->Ldet};
    if (kind < 0 || kind > 9) return NULL;
    if (im[kind]->w == 0) return NULL;
    return im[kind]->buf;
}
double akzo_result_contrast_factor(const akzo_result *r) { return r->contrast_factor; }
uint64_t akzo_result_num_candidates(const akzo_result *r) { return r->n_candidates; }
uint64_t akzo_result_num_cache(const akzo_result *r) { return r->n_cache; }
uint64_t akzo_result_num_keypoints(const akzo_result *r) { return r->n_kps; }
const akzo_keypoint *akzo_result_keypoints(const akzo_result *r) { return r->kps; }
uint64_t akzo_result_descriptor_len(const akzo_result *r) { return r->desc_len; }
const uint8_t *akzo_result_descriptors(const akzo_result *r) { return r->desc; }

/* ------------------------------------------------------------------------------------------------
 * Matcher (ops/feature_matching.rs)
 * ---------------------------------------------------------------------------------------------- */
/* :113-123 -- byte-wise popcount with early bail-out */
static size_t hamming_distance(const uint8_t *d0, const uint8_t *d1, size_t len, size_t bailout) {
    size_t distance = 0;
    for (size_t i = 0; i < len; i++) {
        distance += (size_t)__builtin_popcount((unsigned)(d0[i] ^ d1[i]));
        if (distance > bailout) break;
    }
    return distance;
}
/* :37-50 for one query */
static void top2_one(const uint8_t *d0, const uint8_t *db, uint64_t ndb, uint64_t len, uint64_t stride,
                     size_t thr, size_t *min_d, size_t *min_j, size_t *second) {
    size_t min_distance = thr, mj = 0, second_to_min = thr;
    for (uint64_t j = 0; j < ndb; j++) {
        size_t d = hamming_distance(d0, db + j * stride, len, second_to_min);
        if (d < min_distance) {
            second_to_min = min_distance;
            min_distance = d;
            mj = j;
        } else if (d < second_to_min) {
            second_to_min = d;
        }
    }
    *min_d = min_distance;
    *min_j = mj;
    *second = second_to_min;
}
void akzo_match_top2(const uint8_t *q, uint64_t nq, const uint8_t *db, uint64_t ndb, uint64_t desc_len,
                     uint64_t stride, uint32_t *best_idx, uint32_t *best, uint32_t *second) {
    for (uint64_t i = 0; i < nq; i++) {
        size_t a, b, c;
        top2_one(q + i * stride, db, ndb, desc_len, stride, 10000, &a, &b, &c);
        best[i] = (uint32_t)a;
        best_idx[i] = (uint32_t)b;
        second[i] = (uint32_t)c;
    }
}
/* :23-94 */
uint64_t akzo_descriptor_match(const uint8_t *d0, uint64_t n0, const uint8_t *d1, uint64_t n1, uint64_t desc_len,
                               uint64_t stride, uint64_t distance_threshold, double lowes_ratio, akzo_match *out) {
    uint64_t n_out = 0;
    double r2 = lowes_ratio * lowes_ratio; /* powi(2) */
    for (uint64_t i = 0; i < n0; i++) {
        size_t min_distance, min_j, second;
        top2_one(d0 + i * stride, d1, n1, desc_len, stride, (size_t)distance_threshold, &min_distance, &min_j, &second);
        if ((double)min_distance < (double)second * r2) {
            if (min_distance < (size_t)distance_threshold) {
                out[n_out].index_0 = i;
                out[n_out].index_1 = min_j;
                out[n_out].distance = (double)min_distance;
                n_out++;
            }
        }
    }
    return n_out;
}
