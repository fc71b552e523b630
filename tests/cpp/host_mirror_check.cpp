// Exercises the C++ host mirror (include/akaze_b200.hpp) for the Python test-suite:
//   host_mirror_check formats <out_dir>
//       no GPU: writes features.bin / matches.bin from fixed values (compared byte for byte with the Python writer),
//       reads them back, and runs remove_outliers on a synthetic two-view scene (prints "ransac <kept> <of>")
//   host_mirror_check gpu <w> <h> <img0.u8> <img1.u8> <out_dir>
//       extract_features on two raw 8-bit luma files, descriptor_match + match_features, results as bincode files
#include <cstdio>
#include <cstdlib>
#include <string>

#include "akaze_b200.hpp"

static std::vector<uint8_t> slurp(const char* p) { return akaze_util::detail::read_file(p); }

int main(int argc, char** argv) {
    try {
        if (argc >= 3 && std::string(argv[1]) == "formats") {
            const std::string out = argv[2];
            akaze_util::Features f;
            for (int i = 0; i < 5; i++) {
                akaze::Keypoint k;
                k.point = {10.5f + i, 20.25f * i};
                k.response = 0.001f * (i + 1);
                k.size = 2.4f * (1 << (i % 3));
                k.octave = i % 3;
                k.class_id = 4 * (i % 3) + 1;
                k.angle = 0.1f * i - 0.2f;
                f.keypoints.push_back(k);
                akaze::Descriptor d;
                d.vector.resize(61);
                for (int j = 0; j < 61; j++) d.vector[j] = (uint8_t)(i * 37 + j * 11);
                f.descriptors.push_back(d);
            }
            akaze_util::serialize_features_to_file(f, out + "/features.bin");
            const akaze_util::Features g = akaze_util::deserialize_features_from_file(out + "/features.bin");
            if (g.keypoints.size() != 5 || g.descriptors[4].vector != f.descriptors[4].vector || g.keypoints[3].angle != f.keypoints[3].angle) return 3;
            std::vector<akaze::Match> m = {{0, 3, 12.0}, {1, 1, 0.0}, {4, 2, 77.0}};
            akaze_util::serialize_matches_to_file(m, out + "/matches.bin");
            if (akaze_util::deserialize_matches_from_file(out + "/matches.bin")[2].distance != 77.0) return 4;
            // RANSAC: points related by a pure horizontal shift satisfy x'^T F x = 0 for F = [t]_x; 40 inliers + 10 gross outliers
            std::vector<akaze::Keypoint> k0, k1;
            std::vector<akaze::Match> cand;
            unsigned s = 12345;
            auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((s >> 8) & 0xffff) / 65535.0f; };
            for (int i = 0; i < 50; i++) {
                akaze::Keypoint a, b;
                a.point = {rnd() * 640.0f, rnd() * 480.0f};
                b.point = {a.point.first + 5.0f + 20.0f * rnd(), a.point.second};
                if (i >= 40) b.point.second += 60.0f + 100.0f * rnd();
                k0.push_back(a);
                k1.push_back(b);
                cand.push_back({(size_t)i, (size_t)i, 10.0});
            }
            const std::vector<akaze::Match> kept = akaze::remove_outliers(k0, k1, cand, 10, 1e-7f, 3.0f);
            std::printf("ransac %zu %zu\n", kept.size(), cand.size());
            return 0;
        }
        if (argc >= 7 && std::string(argv[1]) == "gpu") {
            const uint32_t w = (uint32_t)std::atoi(argv[2]), h = (uint32_t)std::atoi(argv[3]);
            const std::vector<uint8_t> i0 = slurp(argv[4]), i1 = slurp(argv[5]);
            const std::string out = argv[6];
            if (i0.size() != (size_t)w * h || i1.size() != (size_t)w * h) return 5;
            akaze::Engine eng(0, w, h, 1);
            auto [e0, kp0, d0] = akaze::extract_features(i0.data(), w, h, w, akaze::Config(), eng);
            // the second image goes through the GrayFloatImage entry point (create_unit_float_image on the host)
            auto [e1, kp1, d1] = akaze::extract_features(akaze::create_unit_float_image(i1.data(), w, h, w), akaze::Config(), eng);
            akaze_util::serialize_features_to_file({kp0, d0}, out + "/features0.bin");
            akaze_util::serialize_features_to_file({kp1, d1}, out + "/features1.bin");
            const std::vector<akaze::Match> dm = akaze::descriptor_match(d0, d1, 10000, 0.86, eng);
            akaze_util::serialize_matches_to_file(dm, out + "/descriptor_matches.bin");
            const std::vector<akaze::Match> mf = akaze::match_features(kp0, d0, kp1, d1, 0.86, 1000, 3.0f, eng);
            akaze_util::serialize_matches_to_file(mf, out + "/matches.bin");
            std::printf("levels %zu %zu keypoints %zu %zu descriptor_matches %zu matches %zu tau0 %zu\n", e0.size(), e1.size(), kp0.size(), kp1.size(),
                        dm.size(), mf.size(), e0.size() > 1 ? e0[1].fed_tau_steps.size() : 0);
            // RANSAC on the GPU: with the reference's sampling it must return exactly the host mirror's inliers; with an advancing
            // source it tries 1000 distinct hypotheses, the first of which is the reference's, so it keeps at least as many
            const std::vector<akaze::Match> host = akaze::remove_outliers(kp0, kp1, dm, 1000, 0.05f, 3.0f);
            akaze::Matrix3 f_ref, f_adv;
            const std::vector<akaze::Match> gpu = akaze::remove_outliers_b200(kp0, kp1, dm, 1000, 0.05f, 3.0f, AKZ_RANSAC_REFERENCE, eng, &f_ref);
            const std::vector<akaze::Match> adv = akaze::remove_outliers_b200(kp0, kp1, dm, 1000, 0.05f, 3.0f, AKZ_RANSAC_ADVANCING, eng, &f_adv);
            bool same = host.size() == gpu.size() && host.size() == mf.size();
            for (size_t i = 0; same && i < host.size(); i++) same = host[i].index_0 == gpu[i].index_0 && host[i].index_1 == gpu[i].index_1;
            // every match the advancing run kept is an inlier of the model it returned (evaluate_model on the host, same arithmetic)
            size_t consistent = 0;
            for (const akaze::Match& m : adv) consistent += akaze::evaluate_model(f_adv, kp0[m.index_0], kp1[m.index_1]) < 3.0f ? 1 : 0;
            akaze_util::serialize_matches_to_file(adv, out + "/matches_ransac_advancing.bin");
            std::printf("ransac host %zu gpu %zu same %d advancing %zu consistent %zu\n", host.size(), gpu.size(), same ? 1 : 0, adv.size(), consistent);
            return 0;
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
    std::fprintf(stderr, "usage: host_mirror_check formats <dir> | gpu <w> <h> <img0> <img1> <dir>\n");
    return 1;
}
