// ransac.cu -- RANSAC outlier removal with the fundamental-matrix model, all trials in parallel.
//
// Replaces the loops of ops::estimate_fundamental_matrix::remove_outliers (akaze/src/ops/estimate_fundamental_matrix.rs:
// 99-165), estimate_fundamental_matrix (:17-69) and evaluate_model (:79-83), the second half of akaze::match_features
// (akaze/src/lib.rs:267-274):
//   k_ransac_models   one thread per trial: the 8 x 9 system of its eight matches, its singular values and right singular
//                     vectors, the rank test and the model. nalgebra ^0.16's SVD is not under /root/reference (parity
//                     unpinned); like the C++ host mirror (include/akaze_b200.hpp: detail::svd_8x9) this is the cyclic
//                     Jacobi eigen-solution of A^T A in f64, operation for operation, so host mirror and device agree bit
//                     for bit. As in the reference the model is the right singular vector of the SMALLEST OF THE EIGHT
//                     singular values of the thin SVD (:47-66), not the null vector of the 8 x 9 system.
//   k_ransac_count    one block per trial: |p_r^T F p_l| < epsilon_inlier over all matches (:140-149)
//   k_ransac_mask     the final inlier set under the winning model (:153-164)
// Which eight matches a trial draws is decided on the host (akaze_api.cu: the `random` crate's default source,
// xorshift128+ seeded [42, 69]); the reference creates a FRESH default source per trial (:118), so all its trials draw the
// same eight matches -- AKZ_RANSAC_REFERENCE reproduces that, AKZ_RANSAC_ADVANCING lets one source run on across the trials.
#include "common.cuh"

namespace akz {
namespace {

__device__ void svd_8x9(const float (&a)[8][9], double (&sv)[9], double (&v)[9][9]) {
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
                if (fabs(m[p][q]) < 1e-300) continue;
                const double theta = (m[q][q] - m[p][p]) / (2.0 * m[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
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
    for (int i = 0; i < 9; i++) sv[i] = sqrt(fmax(0.0, m[i][i]));
}

// models[t][9] row-major 3x3 (Matrix3::new takes its arguments row by row, :55-65); ok[t] = the rank test passed
__global__ void k_ransac_models(const float2* __restrict__ pl, const float2* __restrict__ pr, const unsigned int* __restrict__ samples,
                                unsigned int n_trials, float epsilon, float* __restrict__ models, unsigned int* __restrict__ ok) {
    const unsigned int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_trials) return;
    float a[8][9];
    for (int i = 0; i < 8; i++) {
        const unsigned int mi = samples[(size_t)t * 8 + i];
        const float x0 = pl[mi].x, y0 = pl[mi].y, x1 = pr[mi].x, y1 = pr[mi].y;
        a[i][0] = x0 * x1; a[i][1] = x0 * y1; a[i][2] = x0;
        a[i][3] = y0 * x1; a[i][4] = y0 * y1; a[i][5] = y0;
        a[i][6] = x1;      a[i][7] = y1;      a[i][8] = 1.0f;
    }
    double sv[9], v[9][9];
    svd_8x9(a, sv, v);
    // order[] = indices by descending singular value (the host mirror's std::sort; singular values of a real system are distinct)
    int order[9];
    for (int i = 0; i < 9; i++) order[i] = i;
    for (int i = 1; i < 9; i++) {
        const int key = order[i];
        int j = i - 1;
        while (j >= 0 && sv[order[j]] < sv[key]) {
            order[j + 1] = order[j];
            j--;
        }
        order[j + 1] = key;
    }
    int rank = 0;
    for (int i = 0; i < 8; i++) rank += ((float)sv[order[i]] > epsilon) ? 1 : 0;  // svd.rank(epsilon) != 8 -> None (:43-44)
    const int col = order[7];
    float f[9];
    for (int i = 0; i < 9; i++) f[i] = (float)v[i][col];
    float* o = models + (size_t)t * 9;
    o[0] = f[0]; o[1] = f[3]; o[2] = f[6];
    o[3] = f[1]; o[4] = f[4]; o[5] = f[7];
    o[6] = f[2]; o[7] = f[5]; o[8] = f[8];
    ok[t] = rank == 8 ? 1u : 0u;
}

// evaluate_model (:79-83), in the host mirror's operation order
__device__ __forceinline__ float model_error(const float* F, float2 l, float2 r) {
    const float plv[3] = {l.x, l.y, 1.0f}, prv[3] = {r.x, r.y, 1.0f};
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float row = 0.0f;
#pragma unroll
        for (int j = 0; j < 3; j++) row += F[i * 3 + j] * plv[j];
        acc += prv[i] * row;
    }
    return fabsf(acc);
}

__global__ void __launch_bounds__(128)
k_ransac_count(const float2* __restrict__ pl, const float2* __restrict__ pr, unsigned int n_matches, const float* __restrict__ models,
               const unsigned int* __restrict__ ok, float epsilon_inlier, unsigned int* __restrict__ counts) {
    __shared__ float F[9];
    __shared__ unsigned int total;
    const unsigned int t = blockIdx.x;
    if (threadIdx.x < 9) F[threadIdx.x] = models[(size_t)t * 9 + threadIdx.x];
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    unsigned int c = 0;
    if (ok[t])
        for (unsigned int i = threadIdx.x; i < n_matches; i += blockDim.x) c += model_error(F, pl[i], pr[i]) < epsilon_inlier ? 1u : 0u;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&total, c);
    __syncthreads();
    if (threadIdx.x == 0) counts[t] = total;
}

__global__ void k_ransac_mask(const float2* __restrict__ pl, const float2* __restrict__ pr, unsigned int n_matches, const float* __restrict__ model,
                              float epsilon_inlier, unsigned char* __restrict__ mask) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_matches) return;
    float F[9];
    for (int k = 0; k < 9; k++) F[k] = model[k];
    mask[i] = model_error(F, pl[i], pr[i]) < epsilon_inlier ? 1 : 0;
}

}  // namespace

int launch_ransac_models(cudaStream_t s, const float2* pl, const float2* pr, const unsigned int* samples, unsigned int n_trials,
                         float epsilon, float* models, unsigned int* ok) {
    if (n_trials == 0) return 0;
    k_ransac_models<<<(n_trials + 63) / 64, 64, 0, s>>>(pl, pr, samples, n_trials, epsilon, models, ok);
    return 1;
}
int launch_ransac_count(cudaStream_t s, const float2* pl, const float2* pr, unsigned int n_matches, const float* models,
                        const unsigned int* ok, unsigned int n_trials, float epsilon_inlier, unsigned int* counts) {
    if (n_trials == 0) return 0;
    k_ransac_count<<<n_trials, 128, 0, s>>>(pl, pr, n_matches, models, ok, epsilon_inlier, counts);
    return 1;
}
int launch_ransac_mask(cudaStream_t s, const float2* pl, const float2* pr, unsigned int n_matches, const float* model, float epsilon_inlier,
                       unsigned char* mask) {
    if (n_matches == 0) return 0;
    k_ransac_mask<<<(n_matches + 255) / 256, 256, 0, s>>>(pl, pr, n_matches, model, epsilon_inlier, mask);
    return 1;
}

}  // namespace akz
