// akaze_b200.hpp -- C++ host mirror of the akaze-rust crate's public surface on top of the C ABI
// (akaze_b200.h). The reference host is Rust; this image has no Rust toolchain, so the compiled-language
// host side is C++17, header-only, with the crate's names, argument meaning and error behaviour:
//
//   akaze::types::evolution::Config            akaze/src/types/evolution.rs:8-55        -> akaze::Config
//   akaze::types::keypoint::{Keypoint,Descriptor}  types/keypoint.rs:8-36               -> akaze::Keypoint, akaze::Descriptor
//   akaze::types::feature_match::Match         types/feature_match.rs:9-16              -> akaze::Match
//   akaze::types::image::GrayFloatImage        types/image.rs:32-36                     -> akaze::GrayFloatImage
//   akaze::types::evolution::EvolutionStep     types/evolution.rs:59-92                 -> akaze::EvolutionStep
//   akaze::extract_features                    akaze/src/lib.rs:167-194                 -> akaze::extract_features
//   akaze::ops::feature_matching::descriptor_match  ops/feature_matching.rs:23-94      -> akaze::descriptor_match
//   akaze::ops::estimate_fundamental_matrix::{estimate_fundamental_matrix, remove_outliers}
//                                              ops/estimate_fundamental_matrix.rs:17-165 -> same names (host, as in the crate)
//   akaze::match_features                      akaze/src/lib.rs:252-275                 -> akaze::match_features
//   akaze_util::{Features, (de)serialize_*_to/from_file}  akaze-util/src/lib.rs:11-67   -> akaze_util::* (bincode 1.1 layout)
//
// Differences, all at the boundary SURVEY.md section 8(b) draws: extract_features takes the decoded
// GrayFloatImage (or 8-bit luma) instead of a path, because JPEG decode + to_luma stay with the host's image
// library; where the reference panics (unwrap) these functions throw std::runtime_error carrying
// akz_last_error(). There is no CPU fallback: every function needs the CUDA library and an sm_100 device.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <set>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "akaze_b200.h"

namespace akaze {

// types/evolution.rs:8-55 (field for field, Default::default())
struct Config {
    uint32_t num_sublevels = 4;
    uint32_t max_octave_evolution = 4;
    double base_scale_offset = 1.6;
    double initial_contrast = 0.001;
    double contrast_percentile = 0.7;
    size_t contrast_factor_num_bins = 300;
    double derivative_factor = 1.5;
    double detector_threshold = 0.001;
    size_t descriptor_channels = 3;
    size_t descriptor_pattern_size = 10;
    akz_config to_c() const {
        akz_config c;
        std::memset(&c, 0, sizeof(c));
        c.num_sublevels = num_sublevels;
        c.max_octave_evolution = max_octave_evolution;
        c.base_scale_offset = base_scale_offset;
        c.initial_contrast = initial_contrast;
        c.contrast_percentile = contrast_percentile;
        c.contrast_factor_num_bins = contrast_factor_num_bins;
        c.derivative_factor = derivative_factor;
        c.detector_threshold = detector_threshold;
        c.descriptor_channels = descriptor_channels;
        c.descriptor_pattern_size = descriptor_pattern_size;
        return c;
    }
};

// types/keypoint.rs:8-30
struct Keypoint {
    std::pair<float, float> point{0.0f, 0.0f};
    float response = 0.0f;
    float size = 0.0f;
    size_t octave = 0;
    size_t class_id = 0;
    float angle = 0.0f;
};
// types/keypoint.rs:34-36
struct Descriptor {
    std::vector<uint8_t> vector;
};
// types/feature_match.rs:9-16
struct Match {
    size_t index_0 = 0;
    size_t index_1 = 0;
    double distance = 0.0;
};
// types/image.rs:32-36 (row-major unit floats)
struct GrayFloatImage {
    uint32_t width = 0, height = 0;
    std::vector<float> buffer;
    float get(uint32_t x, uint32_t y) const { return buffer[(size_t)y * width + x]; }
};
// create_unit_float_image (types/image.rs:127-140) for callers that hold 8-bit luma
inline GrayFloatImage create_unit_float_image(const uint8_t* luma, uint32_t width, uint32_t height, size_t stride) {
    GrayFloatImage g;
    g.width = width;
    g.height = height;
    g.buffer.resize((size_t)width * height);
    for (uint32_t y = 0; y < height; y++)
        for (uint32_t x = 0; x < width; x++) g.buffer[(size_t)y * width + x] = ((float)luma[(size_t)y * stride + x] * 1.0f) / 255.0f;
    return g;
}
// types/evolution.rs:59-92; the ten images are downloaded on demand (they exist on the device only when the
// engine was created with keep_evolutions)
struct EvolutionStep {
    double etime = 0.0, esigma = 0.0;
    uint32_t octave = 0, sublevel = 0, sigma_size = 0;
    uint32_t width = 0, height = 0;
    std::vector<double> fed_tau_steps;
};

[[noreturn]] inline void fail(const char* what) { throw std::runtime_error(std::string(what) + ": " + akz_last_error()); }

// RAII over akz_context. The library serialises calls per context (see the thread-safety note in akaze_b200.h), so one
// Engine may be shared between threads; use one per thread or per GPU for concurrency.
class Engine {
public:
    explicit Engine(int device = 0, uint32_t max_width = 4096, uint32_t max_height = 4096, uint32_t max_batch = 1, bool keep_evolutions = false) {
        if (akz_create(device, max_width, max_height, max_batch, keep_evolutions ? AKZ_KEEP_EVOLUTIONS : 0u, &ctx_) != AKZ_OK) fail("akz_create");
    }
    ~Engine() {
        if (ctx_) akz_destroy(ctx_);
    }
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;
    akz_context* get() const { return ctx_; }

private:
    akz_context* ctx_ = nullptr;
};

inline Engine& default_engine() {
    static thread_local Engine e;
    return e;
}

namespace detail {
inline std::tuple<std::vector<EvolutionStep>, std::vector<Keypoint>, std::vector<Descriptor>> unpack(akz_features* f) {
    std::vector<EvolutionStep> evo(akz_features_num_levels(f));
    for (uint32_t l = 0; l < evo.size(); l++) {
        akz_level_info li;
        if (akz_features_level_info(f, l, &li) != AKZ_OK) fail("akz_features_level_info");
        EvolutionStep& e = evo[l];
        e.etime = li.etime;
        e.esigma = li.esigma;
        e.octave = li.octave;
        e.sublevel = li.sublevel;
        e.sigma_size = li.sigma_size;
        e.width = li.width;
        e.height = li.height;
        e.fed_tau_steps.resize(li.n_steps);
        if (li.n_steps && akz_features_fed_tau(f, l, e.fed_tau_steps.data(), li.n_steps) != AKZ_OK) fail("akz_features_fed_tau");
    }
    const uint64_t n = akz_features_count(f);
    const akz_keypoint* k = akz_features_keypoints(f);
    const uint8_t* d = akz_features_descriptors(f);
    const uint32_t len = akz_features_descriptor_len(f);
    std::vector<Keypoint> kps(n);
    std::vector<Descriptor> desc(n);
    for (uint64_t i = 0; i < n; i++) {
        kps[i].point = {k[i].x, k[i].y};
        kps[i].response = k[i].response;
        kps[i].size = k[i].size;
        kps[i].octave = k[i].octave;
        kps[i].class_id = k[i].class_id;
        kps[i].angle = k[i].angle;
        desc[i].vector.assign(d + i * AKZ_DESCRIPTOR_STRIDE, d + i * AKZ_DESCRIPTOR_STRIDE + len);
    }
    akz_features_free(f);
    return {std::move(evo), std::move(kps), std::move(desc)};
}
}  // namespace detail

// akaze::extract_features (lib.rs:167-194) from the GrayFloatImage on
inline std::tuple<std::vector<EvolutionStep>, std::vector<Keypoint>, std::vector<Descriptor>> extract_features(
    const GrayFloatImage& image, Config options = Config(), Engine& engine = default_engine()) {
    const akz_config c = options.to_c();
    akz_features* f = nullptr;
    if (akz_extract_f32(engine.get(), image.buffer.data(), image.width, image.height, &c, &f) != AKZ_OK) fail("extract_features");
    return detail::unpack(f);
}
// same, from 8-bit luma (the u8 -> unit float conversion of image.rs:127-140 runs on the device)
inline std::tuple<std::vector<EvolutionStep>, std::vector<Keypoint>, std::vector<Descriptor>> extract_features(
    const uint8_t* luma, uint32_t width, uint32_t height, size_t stride, Config options = Config(), Engine& engine = default_engine()) {
    const akz_config c = options.to_c();
    akz_features* f = nullptr;
    if (akz_extract_u8(engine.get(), luma, width, height, stride, &c, &f) != AKZ_OK) fail("extract_features");
    return detail::unpack(f);
}

// ops::feature_matching::descriptor_match (feature_matching.rs:23-94): top-2 on the GPU, f64 Lowe test on the host
inline std::vector<Match> descriptor_match(const std::vector<Descriptor>& d0, const std::vector<Descriptor>& d1, size_t distance_threshold,
                                           double lowes_ratio, Engine& engine = default_engine()) {
    auto pack = [](const std::vector<Descriptor>& d, uint32_t* len) {
        std::vector<uint8_t> rows(std::max<size_t>(d.size(), 1) * AKZ_DESCRIPTOR_STRIDE, 0);
        for (size_t i = 0; i < d.size(); i++) {
            if (d[i].vector.size() > AKZ_DESCRIPTOR_STRIDE) throw std::runtime_error("descriptor longer than 64 bytes");
            *len = std::max<uint32_t>(*len, (uint32_t)d[i].vector.size());
            std::memcpy(rows.data() + i * AKZ_DESCRIPTOR_STRIDE, d[i].vector.data(), d[i].vector.size());
        }
        return rows;
    };
    uint32_t len = 1;
    const std::vector<uint8_t> r0 = pack(d0, &len), r1 = pack(d1, &len);
    std::vector<akz_match> out(std::max<size_t>(d0.size(), 1));
    uint64_t n = 0;
    if (akz_descriptor_match(engine.get(), r0.data(), d0.size(), r1.data(), d1.size(), len, AKZ_DESCRIPTOR_STRIDE, distance_threshold, lowes_ratio,
                             out.data(), &n) != AKZ_OK)
        fail("descriptor_match");
    std::vector<Match> m(n);
    for (uint64_t i = 0; i < n; i++) m[i] = Match{(size_t)out[i].index_0, (size_t)out[i].index_1, out[i].distance};
    return m;
}

// ---- RANSAC, host code as in the crate (ops/estimate_fundamental_matrix.rs) ---------------------------------
namespace detail {
// singular values and right singular vectors of an 8x9 f32 matrix through the symmetric eigenproblem of A^T A
// (cyclic Jacobi, f64 accumulation). nalgebra ^0.16's SVD is not under /root/reference: parity unpinned.
inline void svd_8x9(const float (&a)[8][9], double (&sv)[9], double (&v)[9][9]) {
    double m[9][9];
    for (int i = 0; i < 9; i++)
        for (int j = 0; j < 9; j++) {
            double s = 0.0;
            for (int r = 0; r < 8; r++) s += (double)a[r][i] * (double)a[r][j];
            m[i][j] = s;
            v[i][j] = i == j ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (int p = 0; p < 9; p++)
            for (int q = p + 1; q < 9; q++) off += m[p][q] * m[p][q];
        if (off < 1e-30) break;
        for (int p = 0; p < 9; p++)
            for (int q = p + 1; q < 9; q++) {
                if (std::fabs(m[p][q]) < 1e-300) continue;
                const double theta = (m[q][q] - m[p][p]) / (2.0 * m[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 9; k++) {
                    const double mkp = m[k][p], mkq = m[k][q];
                    m[k][p] = c * mkp - s * mkq;
                    m[k][q] = s * mkp + c * mkq;
                }
                for (int k = 0; k < 9; k++) {
                    const double mpk = m[p][k], mqk = m[q][k];
                    m[p][k] = c * mpk - s * mqk;
                    m[q][k] = s * mpk + c * mqk;
                }
                for (int k = 0; k < 9; k++) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 9; i++) sv[i] = std::sqrt(std::max(0.0, m[i][i]));
}
// the `random` crate's default source (xorshift128+ seeded [42, 69]); ^0.12 is not under /root/reference
struct Xorshift128Plus {
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
}  // namespace detail

struct Matrix3 {
    float m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
};

// estimate_fundamental_matrix.rs:17-69: eight matches -> F, or false when the system is rank deficient
inline bool estimate_fundamental_matrix(const std::vector<Keypoint>& k0, const std::vector<Keypoint>& k1, const std::vector<Match>& eight,
                                        float epsilon, Matrix3* out) {
    if (eight.size() != 8) return false;
    float a[8][9];
    for (int i = 0; i < 8; i++) {
        const float x0 = k0[eight[i].index_0].point.first, y0 = k0[eight[i].index_0].point.second;
        const float x1 = k1[eight[i].index_1].point.first, y1 = k1[eight[i].index_1].point.second;
        const float row[9] = {x0 * x1, x0 * y1, x0, y0 * x1, y0 * y1, y0, x1, y1, 1.0f};
        std::memcpy(a[i], row, sizeof(row));
    }
    double sv[9], v[9][9];
    detail::svd_8x9(a, sv, v);
    // the eight largest singular values must exceed epsilon (:44-49); the null vector belongs to the smallest of them
    int order[9] = {0, 1, 2, 3, 4, 5, 6, 7, 8};
    std::sort(order, order + 9, [&](int x, int y) { return sv[x] > sv[y]; });
    int rank = 0;
    for (int i = 0; i < 8; i++) rank += (float)sv[order[i]] > epsilon ? 1 : 0;
    if (rank != 8) return false;
    const int col = order[7];  // smallest of the 8 singular values of the 8x9 system
    float f[9];
    for (int i = 0; i < 9; i++) f[i] = (float)v[i][col];
    // Matrix3::new takes its arguments row by row (:55-65)
    const Matrix3 r = {{{f[0], f[3], f[6]}, {f[1], f[4], f[7]}, {f[2], f[5], f[8]}}};
    *out = r;
    return true;
}
// evaluate_model (:79-83): |p_r^T F p_l|
inline float evaluate_model(const Matrix3& F, const Keypoint& l, const Keypoint& r) {
    const float pl[3] = {l.point.first, l.point.second, 1.0f}, pr[3] = {r.point.first, r.point.second, 1.0f};
    float acc = 0.0f;
    for (int i = 0; i < 3; i++) {
        float row = 0.0f;
        for (int j = 0; j < 3; j++) row += F.m[i][j] * pl[j];
        acc += pr[i] * row;
    }
    return std::fabs(acc);
}
// remove_outliers (:99-165). Like the reference, a fresh default random source is created on every trial, so every
// trial draws the same eight matches; they are used in ascending index order (the reference's HashSet order is
// process-random).
inline std::vector<Match> remove_outliers(const std::vector<Keypoint>& k0, const std::vector<Keypoint>& k1, const std::vector<Match>& matches,
                                          size_t num_trials, float epsilon_model, float epsilon_inlier) {
    if (matches.size() < 8) return matches;
    size_t max_inliers = 0;
    Matrix3 final_model;
    std::set<size_t> last;
    bool have_last = false, last_ok = false;
    Matrix3 last_model;
    size_t last_count = 0;
    for (size_t t = 0; t < num_trials; t++) {
        detail::Xorshift128Plus src;
        std::set<size_t> chosen;
        while (chosen.size() < 8) chosen.insert((size_t)(src.read_u64() % matches.size()));
        if (!have_last || chosen != last) {
            std::vector<Match> eight;
            for (size_t j : chosen) eight.push_back(matches[j]);
            last_ok = estimate_fundamental_matrix(k0, k1, eight, epsilon_model, &last_model);
            last_count = 0;
            if (last_ok)
                for (const Match& m : matches) last_count += evaluate_model(last_model, k0[m.index_0], k1[m.index_1]) < epsilon_inlier ? 1 : 0;
            last = chosen;
            have_last = true;
        }
        if (last_ok && last_count > max_inliers) {
            max_inliers = last_count;
            final_model = last_model;
        }
    }
    std::vector<Match> inliers;
    for (const Match& m : matches)
        if (evaluate_model(final_model, k0[m.index_0], k1[m.index_1]) < epsilon_inlier) inliers.push_back(m);
    return inliers;
}

// remove_outliers on the GPU (akz_remove_outliers): every trial's hypothesis and inlier count in parallel. With
// AKZ_RANSAC_REFERENCE it returns exactly what remove_outliers above returns (same samples, same Jacobi SVD, same operation
// order); AKZ_RANSAC_ADVANCING lets the random source run on across the trials, i.e. num_trials distinct hypotheses.
inline std::vector<Match> remove_outliers_b200(const std::vector<Keypoint>& k0, const std::vector<Keypoint>& k1, const std::vector<Match>& matches,
                                               size_t num_trials, float epsilon_model, float epsilon_inlier, int sampling = AKZ_RANSAC_REFERENCE,
                                               Engine& engine = default_engine(), Matrix3* model = nullptr) {
    auto pack = [](const std::vector<Keypoint>& k) {
        std::vector<akz_keypoint> r(k.size());
        for (size_t i = 0; i < k.size(); i++)
            r[i] = akz_keypoint{k[i].point.first, k[i].point.second, k[i].response, k[i].size, (uint32_t)k[i].octave, (uint32_t)k[i].class_id, k[i].angle};
        return r;
    };
    const std::vector<akz_keypoint> a = pack(k0), b = pack(k1);
    std::vector<akz_match> in(matches.size()), out(std::max<size_t>(matches.size(), 1));
    for (size_t i = 0; i < matches.size(); i++) in[i] = akz_match{(uint64_t)matches[i].index_0, (uint64_t)matches[i].index_1, matches[i].distance};
    uint64_t n = 0;
    float f[9];
    if (akz_remove_outliers(engine.get(), a.data(), a.size(), b.data(), b.size(), in.data(), in.size(), num_trials, epsilon_model, epsilon_inlier,
                            sampling, out.data(), &n, f) != AKZ_OK)
        fail("remove_outliers");
    if (model)
        for (int i = 0; i < 9; i++) model->m[i / 3][i % 3] = f[i];
    std::vector<Match> r(n);
    for (uint64_t i = 0; i < n; i++) r[i] = Match{(size_t)out[i].index_0, (size_t)out[i].index_1, out[i].distance};
    return r;
}

// akaze::match_features (lib.rs:252-275): descriptor_match with the hard-wired distance threshold 10000, then RANSAC
// with epsilon_model = 0.05 (lib.rs:267-274)
inline std::vector<Match> match_features(const std::vector<Keypoint>& keypoints_0, const std::vector<Descriptor>& descriptors_0,
                                         const std::vector<Keypoint>& keypoints_1, const std::vector<Descriptor>& descriptors_1,
                                         double lowes_ratio, size_t ransac_trials, float ransac_epsilon_inliers, Engine& engine = default_engine()) {
    const std::vector<Match> m = descriptor_match(descriptors_0, descriptors_1, 10000, lowes_ratio, engine);
    return remove_outliers(keypoints_0, keypoints_1, m, ransac_trials, 0.05f, ransac_epsilon_inliers);
}

}  // namespace akaze

// ---- akaze-util/src/lib.rs:11-67: on-disk formats (bincode 1.1 defaults: little endian, u64 lengths, usize as u64) ----
namespace akaze_util {

struct Features {
    std::vector<akaze::Keypoint> keypoints;
    std::vector<akaze::Descriptor> descriptors;
};

namespace detail {
struct Writer {
    std::vector<uint8_t> b;
    template <class T>
    void put(T v) {
        const uint8_t* p = reinterpret_cast<const uint8_t*>(&v);
        b.insert(b.end(), p, p + sizeof(T));  // little-endian host (x86-64 / aarch64)
    }
};
struct Reader {
    const std::vector<uint8_t>& b;
    size_t at = 0;
    template <class T>
    T get() {
        if (at + sizeof(T) > b.size()) throw std::runtime_error("bincode: unexpected end of file");
        T v;
        std::memcpy(&v, b.data() + at, sizeof(T));
        at += sizeof(T);
        return v;
    }
};
inline void write_file(const std::string& path, const std::vector<uint8_t>& b) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot create " + path);
    const size_t w = std::fwrite(b.data(), 1, b.size(), f);
    std::fclose(f);
    if (w != b.size()) throw std::runtime_error("short write to " + path);
}
inline std::vector<uint8_t> read_file(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::vector<uint8_t> b;
    uint8_t buf[65536];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) b.insert(b.end(), buf, buf + n);
    std::fclose(f);
    return b;
}
}  // namespace detail

// Keypoint = (f32, f32), f32, f32, u64, u64, f32 = 36 bytes; Features = u64 n + n x 36 B + u64 n + n x (u64 len + bytes)
inline std::vector<uint8_t> serialize_features(const Features& f) {
    detail::Writer w;
    w.put<uint64_t>(f.keypoints.size());
    for (const akaze::Keypoint& k : f.keypoints) {
        w.put<float>(k.point.first);
        w.put<float>(k.point.second);
        w.put<float>(k.response);
        w.put<float>(k.size);
        w.put<uint64_t>(k.octave);
        w.put<uint64_t>(k.class_id);
        w.put<float>(k.angle);
    }
    w.put<uint64_t>(f.descriptors.size());
    for (const akaze::Descriptor& d : f.descriptors) {
        w.put<uint64_t>(d.vector.size());
        w.b.insert(w.b.end(), d.vector.begin(), d.vector.end());
    }
    return w.b;
}
inline Features deserialize_features(const std::vector<uint8_t>& bytes) {
    detail::Reader r{bytes};
    Features f;
    f.keypoints.resize(r.get<uint64_t>());
    for (akaze::Keypoint& k : f.keypoints) {
        k.point.first = r.get<float>();
        k.point.second = r.get<float>();
        k.response = r.get<float>();
        k.size = r.get<float>();
        k.octave = (size_t)r.get<uint64_t>();
        k.class_id = (size_t)r.get<uint64_t>();
        k.angle = r.get<float>();
    }
    f.descriptors.resize(r.get<uint64_t>());
    for (akaze::Descriptor& d : f.descriptors) {
        const uint64_t n = r.get<uint64_t>();
        if (r.at + n > bytes.size()) throw std::runtime_error("bincode: descriptor runs past the end of the file");
        d.vector.assign(bytes.begin() + r.at, bytes.begin() + r.at + n);
        r.at += n;
    }
    return f;
}
// Match = u64, u64, f64 = 24 bytes; Vec<Match> = u64 n + n x 24 B
inline std::vector<uint8_t> serialize_matches(const std::vector<akaze::Match>& m) {
    detail::Writer w;
    w.put<uint64_t>(m.size());
    for (const akaze::Match& x : m) {
        w.put<uint64_t>(x.index_0);
        w.put<uint64_t>(x.index_1);
        w.put<double>(x.distance);
    }
    return w.b;
}
inline std::vector<akaze::Match> deserialize_matches(const std::vector<uint8_t>& bytes) {
    detail::Reader r{bytes};
    std::vector<akaze::Match> m(r.get<uint64_t>());
    for (akaze::Match& x : m) {
        x.index_0 = (size_t)r.get<uint64_t>();
        x.index_1 = (size_t)r.get<uint64_t>();
        x.distance = r.get<double>();
    }
    return m;
}
// the file functions of akaze-util/src/lib.rs (binary flavour; the ".json" flavour is written by the Python host
// layer, akaze_rust_b200.formats)
inline void serialize_features_to_file(const Features& f, const std::string& path) { detail::write_file(path, serialize_features(f)); }
inline Features deserialize_features_from_file(const std::string& path) { return deserialize_features(detail::read_file(path)); }
inline void serialize_matches_to_file(const std::vector<akaze::Match>& m, const std::string& path) { detail::write_file(path, serialize_matches(m)); }
inline std::vector<akaze::Match> deserialize_matches_from_file(const std::string& path) { return deserialize_matches(detail::read_file(path)); }

}  // namespace akaze_util
