// K19: divergence / curl estimates of a vector field sampled on a 3-D point cloud.
//
// Replaces compute_vectorfield_features (reference examples/eeg_example/eeg_utils.py:46-80): a KD-tree query of the
// k nearest neighbours per point (k = 5 there) followed by Python loops
//     delta_position = positions[i] - point ;  delta_vector = normalized_vectors[i] - normalized_vectors[ref]
//     div  += dot(delta_vector, delta_position)   / norm(delta_position)**2
//     curl += cross(delta_vector, delta_position) / norm(delta_position)**2          ; both divided by k at the end.
// The reference indexes normalized_vectors[0] (row 0, not row j) for `ref` (eeg_utils.py:67); `ref_row0` = 1 reproduces
// that, 0 uses row j.  Neighbour lists come from the exact kNN kernel K2 (self excluded, ascending (distance, index) --
// the order KDTree.query returns), so the sums run in the reference's order.  One thread per point; gather-bound:
// 8 n (k+1) * 6 bytes read, 32 n bytes written.
#include "common.cuh"

namespace rvgp {

__device__ __forceinline__ void load_unit3(const double* __restrict__ v, int64_t r, double out[3]) {
    const double a = v[r * 3], b = v[r * 3 + 1], c = v[r * 3 + 2];
    // np.linalg.norm(axis=1): sqrt of the sum of squares in column order
    const double m = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c)));
    out[0] = a / m; out[1] = b / m; out[2] = c / m;
}

__global__ void __launch_bounds__(256)
vf_features_kernel(int n, int k, const double* __restrict__ pos, const double* __restrict__ vec, const int* __restrict__ knn,
                   int ref_row0, double* __restrict__ div, double* __restrict__ curl) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double ref[3];
    load_unit3(vec, ref_row0 ? 0 : j, ref);
    const double px = pos[(int64_t)j * 3], py = pos[(int64_t)j * 3 + 1], pz = pos[(int64_t)j * 3 + 2];
    double dsum = 0.0, c0 = 0.0, c1 = 0.0, c2 = 0.0;
    for (int t = 0; t < k; ++t) {
        const int i = __ldg(knn + (int64_t)j * k + t);
        const double dx = pos[(int64_t)i * 3] - px, dy = pos[(int64_t)i * 3 + 1] - py, dz = pos[(int64_t)i * 3 + 2] - pz;
        double nv[3];
        load_unit3(vec, i, nv);
        const double vx = nv[0] - ref[0], vy = nv[1] - ref[1], vz = nv[2] - ref[2];
        // np.linalg.norm(delta_position)**2: square of the rounded square root
        const double nr = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
        const double r2 = __dmul_rn(nr, nr);
        const double dot = __dadd_rn(__dadd_rn(__dmul_rn(vx, dx), __dmul_rn(vy, dy)), __dmul_rn(vz, dz));
        dsum = __dadd_rn(dsum, dot / r2);
        c0 = __dadd_rn(c0, __dsub_rn(__dmul_rn(vy, dz), __dmul_rn(vz, dy)) / r2);      // np.cross(delta_vector, delta_position)
        c1 = __dadd_rn(c1, __dsub_rn(__dmul_rn(vz, dx), __dmul_rn(vx, dz)) / r2);
        c2 = __dadd_rn(c2, __dsub_rn(__dmul_rn(vx, dy), __dmul_rn(vy, dx)) / r2);
    }
    div[j] = dsum / k;
    curl[(int64_t)j * 3] = c0 / k;
    curl[(int64_t)j * 3 + 1] = c1 / k;
    curl[(int64_t)j * 3 + 2] = c2 / k;
}

}  // namespace rvgp

using namespace rvgp;

extern "C" int rvgp_vectorfield_features_f64(rvgp_handle_t hh, int n, int k, const double* positions, const double* vectors,
                                             const int32_t* knn, int ref_row0, double* div, double* curl) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, n >= 1 && k >= 1 && positions && vectors && knn && div && curl, "vectorfield_features: bad arguments");
    vf_features_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(n, k, positions, vectors, knn, ref_row0, div, curl);
    RVGP_LAUNCH_OK(h, "vf_features_kernel");
    return RVGP_OK;
}
