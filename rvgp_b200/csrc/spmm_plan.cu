// Row-group merge plan: the UNION of the sorted column lists of R consecutive (Morton-ordered) block rows, each union entry
// a (column, R-bit row mask) pair.  The FP64-MMA SpMM (spmm_mma.cu) cuts these union lists into its k-steps
// (rvgp_bsr_mma_pack); adjacent rows share most neighbours, so at R = 4 the union holds about half as many entries as the
// four rows together (reuse 1.88x at C4).
#include <cub/cub.cuh>

#include "common.cuh"

namespace rvgp {

// ---- plan -------------------------------------------------------------------------------------------------------
template <int R>
__global__ void merge_count_kernel(const int* __restrict__ indptr, const int* __restrict__ indices, int n, int ngroups,
                                   int* __restrict__ ulen, int2* __restrict__ uent, const int* __restrict__ gptr) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    int p[R], pe[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = g * R + r;
        p[r] = (row < n) ? indptr[row] : 0;
        pe[r] = (row < n) ? indptr[row + 1] : 0;
    }
    int cnt = 0;
    const int base = gptr ? gptr[g] : 0;
    for (;;) {
        int mn = 0x7fffffff;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (p[r] < pe[r]) mn = min(mn, indices[p[r]] & 0x7fffffff);
        if (mn == 0x7fffffff) break;
        int mask = 0;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (p[r] < pe[r] && (indices[p[r]] & 0x7fffffff) == mn) { mask |= 1 << r; ++p[r]; }
        if (uent) uent[base + cnt] = make_int2(mn, mask);
        ++cnt;
    }
    if (ulen) ulen[g] = cnt;
}

}  // namespace rvgp

using namespace rvgp;

extern "C" int64_t rvgp_bsr_merge_plan_workspace_bytes(int nbrows, int R) {
    const int ngroups = (nbrows + R - 1) / R;
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, (int*)nullptr, (int*)nullptr, ngroups + 1);
    return (int64_t)(((size_t)(ngroups + 1) * 4 + 255) / 256 * 256 + (b + 255) / 256 * 256);
}

// Row-group merge plan.  Pass 1 (uent == NULL): fills gptr (ngroups+1, exclusive prefix of the union lengths); read
// gptr[ngroups] for the total number of union entries, allocate uent (int32 pairs (column, row mask)), then call again
// with uent to fill it.  `indices` may carry the ROT2 flip bit (ignored here).
extern "C" int rvgp_bsr_merge_plan(rvgp_handle_t hh, int nbrows, const int32_t* indptr, const int32_t* indices, int R,
                                   int32_t* gptr, int32_t* uent, void* workspace, int64_t workspace_bytes) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, R == 4 || R == 8, "merge_plan: R must be 4 or 8");
    const int ngroups = cdiv(nbrows, R);
    if (ngroups == 0) return RVGP_OK;
    if (uent == nullptr) {
        if (rvgp_bsr_merge_plan_workspace_bytes(nbrows, R) > workspace_bytes)
            return set_error(h, RVGP_ERR_CAPACITY, "merge_plan: workspace too small%s%s");
        int* ulen = (int*)workspace;
        void* cubtmp = (char*)workspace + ((size_t)(ngroups + 1) * 4 + 255) / 256 * 256;
        size_t cb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, cb, (int*)nullptr, (int*)nullptr, ngroups + 1);
        RVGP_CUDA_OK(h, cudaMemsetAsync(ulen + ngroups, 0, sizeof(int), h->stream));
        if (R == 4) merge_count_kernel<4><<<cdiv(ngroups, 128), 128, 0, h->stream>>>(indptr, indices, nbrows, ngroups, ulen, nullptr, nullptr);
        else merge_count_kernel<8><<<cdiv(ngroups, 128), 128, 0, h->stream>>>(indptr, indices, nbrows, ngroups, ulen, nullptr, nullptr);
        RVGP_LAUNCH_OK(h, "merge_count_kernel");
        RVGP_CUDA_OK(h, cub::DeviceScan::ExclusiveSum(cubtmp, cb, ulen, gptr, ngroups + 1, h->stream));
        h->launches++;
    } else {
        if (R == 4) merge_count_kernel<4><<<cdiv(ngroups, 128), 128, 0, h->stream>>>(indptr, indices, nbrows, ngroups, nullptr, reinterpret_cast<int2*>(uent), gptr);
        else merge_count_kernel<8><<<cdiv(ngroups, 128), 128, 0, h->stream>>>(indptr, indices, nbrows, ngroups, nullptr, reinterpret_cast<int2*>(uent), gptr);
        RVGP_LAUNCH_OK(h, "merge_count_kernel");
    }
    return RVGP_OK;
}
