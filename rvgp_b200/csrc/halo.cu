// Row-sharded Chebyshev filter with the halo exchange done by our own kernels over NVLink peer memory.
//
// The first sharded filter (rvgp_b200/distributed.py: pack kernel -> NCCL all_to_all -> SpMM -> boundary-row SpMM) issues
// five host calls per polynomial degree; at 8 GPUs the GPU work per degree is ~0.11 ms and the step was bound by the host
// (0.27 ms, profiles/r01_scaling_c4.txt).  Here the whole recurrence of a panel is ONE C call that only enqueues kernels:
//
//   per degree:   signal  - tell the neighbour ranks "my rows of X are final"   (release store into THEIR flag array)
//                 wait    - spin until every neighbour has said the same          (acquire loads of MY flag array)
//                 pull    - copy the halo rows of X straight out of the owners' HBM (peer-mapped pointers, no pack, no NCCL)
//                 spmm    - the fused step on [local | halo]
//
// The extended block vectors live in cudaMalloc memory shared through CUDA IPC (rvgp_ipc_*), so every rank holds
// peer-mapped pointers to the others' buffers.  The three rotating buffers make the write-after-read hazard impossible: a
// rank overwrites the buffer its neighbours pulled from at step e only at step e+2, which it cannot start before the
// neighbours have signalled step e+2, i.e. finished step e+1, i.e. finished the pull of step e (stream order).  A barrier
// (signal + wait) at the start of a call protects the copy-in of the next panel the same way.  Epochs are a 64-bit counter
// that only grows; all ranks run the same degrees, so they agree on it without communicating.  The spin has a wall-clock
// timeout (globaltimer) and raises through `err` instead of hanging the GPU when a peer dies.
#include "common.cuh"

namespace rvgp {

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void halo_signal_kernel(unsigned long long* const* __restrict__ peer_slots, int n_peers, unsigned long long epoch) {
    const int i = threadIdx.x;
    if (i >= n_peers) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(peer_slots[i]), "l"(epoch) : "memory");
}

__global__ void halo_wait_kernel(const unsigned long long* __restrict__ flags, const int* __restrict__ wait_idx, int n_peers,
                                 unsigned long long epoch, unsigned long long timeout_ns, int* __restrict__ err) {
    const int i = threadIdx.x;
    if (i >= n_peers) return;
    const unsigned long long* f = flags + wait_idx[i];
    const unsigned long long t0 = globaltimer_ns();
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
        if (v >= epoch) break;
        if (globaltimer_ns() - t0 > timeout_ns) { atomicOr(err, 1); break; }
        __nanosleep(200);
    }
}

// One warp per halo node: copy its row (row_doubles contiguous doubles, a multiple of 2, 16-byte aligned) from the owner.
// Volatile loads: the same peer addresses are re-read every third step and must not be served from this SM's L1.
__global__ void halo_pull_kernel(int n_halo, int row_doubles, const int64_t* __restrict__ src, double* __restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const int node = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (node >= n_halo) return;
    const double* s = reinterpret_cast<const double*>(src[node]);
    double* d = dst + (int64_t)node * row_doubles;
    for (int c = lane * 2; c < row_doubles; c += 64) {
        double a, b;
        asm volatile("ld.volatile.global.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(s + c));
        *reinterpret_cast<double2*>(d + c) = make_double2(a, b);
    }
}

int spmm_dispatch_public(Handle* h, int nbrows, int d, const int* indptr, const int* indices, const double* vals,
                         const double* X, int64_t ldx, const double* W, int64_t ldw, double* Y, int64_t ldy, int ncols,
                         double alpha, double beta, double gamma);
int spmm_mma_native_dispatch(Handle* h, int nbrows, const int* kptr, const int* kcols, const double* afrag, int rotc,
                             const double* X, int64_t nsx, const double* W, int64_t nsw, double* Y, int64_t nsy,
                             int ncols, double alpha, double beta, double gamma, int reverse);
int spmm_mma_native_dispatch_list(Handle* h, int nbrows, const int* kptr, const int* kcols, const double* afrag, int rotc,
                                  const double* X, int64_t nsx, const double* W, int64_t nsw, double* Y, int64_t nsy,
                                  int ncols, double alpha, double beta, double gamma, int reverse, const int* glist, int nlist);
int native_convert(Handle* h, bool to_native, int nbrows, int ncols, double* V, int64_t ldv, double* Xn, int64_t ns);

static int halo_sync(Handle* h, const rvgp_halo_ctx* c, unsigned long long epoch) {
    if (c->n_peers <= 0) return RVGP_OK;
    halo_signal_kernel<<<1, 32, 0, h->stream>>>(reinterpret_cast<unsigned long long* const*>(c->peer_slots), c->n_peers, epoch);
    RVGP_LAUNCH_OK(h, "halo_signal_kernel");
    halo_wait_kernel<<<1, 32, 0, h->stream>>>(reinterpret_cast<const unsigned long long*>(c->flags), c->wait_idx, c->n_peers, epoch,
                                              (unsigned long long)c->timeout_ms * 1000000ull, c->err);
    RVGP_LAUNCH_OK(h, "halo_wait_kernel");
    return RVGP_OK;
}

static int halo_pull(Handle* h, const rvgp_halo_ctx* c, int slot) {
    if (c->n_halo <= 0) return RVGP_OK;
    const int row_doubles = c->d * c->ncols;
    double* dst = c->E[slot] + (int64_t)c->n_loc * row_doubles;
    halo_pull_kernel<<<cdiv(c->n_halo, 8), 256, 0, h->stream>>>(c->n_halo, row_doubles, c->pull_src[slot], dst);
    RVGP_LAUNCH_OK(h, "halo_pull_kernel");
    return RVGP_OK;
}

// one fused step on the extended buffers: E[y][:local] = alpha A E[x] + beta E[x] + gamma E[w]
static int halo_step(Handle* h, const rvgp_halo_ctx* c, unsigned long long epoch, int x, int w, int y, double alpha, double beta,
                     double gamma) {
    int rc;
    const double* W = (w >= 0) ? c->E[w] : nullptr;
    if (c->kptr != nullptr && c->glist_interior != nullptr && c->glist_boundary != nullptr) {
        // Overlapped order (MMA plans with interior / boundary group lists).  The neighbours only ever pull my BOUNDARY rows (the
        // pattern is symmetric: a row they need is a row of mine that references one of theirs), so the interior groups of this
        // step may run -- and overwrite interior rows of E[y] -- before anybody has been waited for; the boundary groups keep the
        // three-buffer guarantee (they are written after the wait).  By the time the interior launch (most of the rows) is done
        // the neighbours have long signalled: their skew and the pull cost nothing.
        const int rotc = c->rotc;
        const int64_t ns = (rotc == 2) ? (int64_t)c->ncols : 2 * (int64_t)c->ncols;
        const int ncn = (rotc == 2) ? c->ncols / 2 : c->ncols;
        if (c->n_peers > 0) {
            halo_signal_kernel<<<1, 32, 0, h->stream>>>(reinterpret_cast<unsigned long long* const*>(c->peer_slots), c->n_peers, epoch);
            RVGP_LAUNCH_OK(h, "halo_signal_kernel");
        }
        rc = spmm_mma_native_dispatch_list(h, c->n_loc, c->kptr, c->kcols, c->afrag, rotc, c->E[x], ns, W, ns, c->E[y], ns, ncn, alpha,
                                           beta, gamma, 0, c->glist_interior, c->n_interior);
        if (rc) return rc;
        if (c->n_peers > 0) {
            halo_wait_kernel<<<1, 32, 0, h->stream>>>(reinterpret_cast<const unsigned long long*>(c->flags), c->wait_idx, c->n_peers,
                                                      epoch, (unsigned long long)c->timeout_ms * 1000000ull, c->err);
            RVGP_LAUNCH_OK(h, "halo_wait_kernel");
        }
        rc = halo_pull(h, c, x);
        if (rc) return rc;
        return spmm_mma_native_dispatch_list(h, c->n_loc, c->kptr, c->kcols, c->afrag, rotc, c->E[x], ns, W, ns, c->E[y], ns, ncn, alpha,
                                             beta, gamma, 0, c->glist_boundary, c->n_boundary);
    }
    // (a single signal + wait + pull kernel was tried in round 2: on 2 B200s it was SLOWER and erratic -- every block of the
    // pull grid spins on the flags while occupying an SM -- so the three small kernels stay; profiles/r02g_*)
    rc = halo_sync(h, c, epoch);
    if (rc) return rc;
    rc = halo_pull(h, c, x);
    if (rc) return rc;
    if (c->kptr != nullptr && c->rotc == 2) {
        // scalar pattern-mode Laplacian as L (x) I_2: the row-major extended buffers ARE native panels with ncols / 2 columns
        const int64_t ns = (int64_t)c->ncols;
        return spmm_mma_native_dispatch(h, c->n_loc, c->kptr, c->kcols, c->afrag, 2, c->E[x], ns, W, ns, c->E[y], ns, c->ncols / 2,
                                        alpha, beta, gamma, 0);
    }
    if (c->kptr != nullptr) {
        const int64_t ns = 2 * (int64_t)c->ncols;
        return spmm_mma_native_dispatch(h, c->n_loc, c->kptr, c->kcols, c->afrag, c->rotc, c->E[x], ns, W, ns, c->E[y], ns, c->ncols,
                                        alpha, beta, gamma, 0);
    }
    return spmm_dispatch_public(h, c->n_loc, c->d, c->indptr, c->indices, c->vals, c->E[x], c->ncols, W, c->ncols, c->E[y],
                                c->ncols, c->ncols, alpha, beta, gamma);
}

static int halo_copy_in(Handle* h, const rvgp_halo_ctx* c, const double* V, int64_t ldv, int slot) {
    if (c->n_loc == 0) return RVGP_OK;
    if (c->kptr != nullptr && c->rotc != 2) return native_convert(h, true, c->n_loc, c->ncols, const_cast<double*>(V), ldv, c->E[slot], 2 * (int64_t)c->ncols);
    RVGP_CUDA_OK(h, cudaMemcpy2DAsync(c->E[slot], (size_t)c->ncols * sizeof(double), V, ldv * sizeof(double),
                                      (size_t)c->ncols * sizeof(double), (size_t)c->n_loc * c->d, cudaMemcpyDeviceToDevice, h->stream));
    return RVGP_OK;
}

static int halo_copy_out(Handle* h, const rvgp_halo_ctx* c, double* V, int64_t ldv, int slot) {
    if (c->n_loc == 0) return RVGP_OK;
    if (c->kptr != nullptr && c->rotc != 2) return native_convert(h, false, c->n_loc, c->ncols, V, ldv, c->E[slot], 2 * (int64_t)c->ncols);
    RVGP_CUDA_OK(h, cudaMemcpy2DAsync(V, ldv * sizeof(double), c->E[slot], (size_t)c->ncols * sizeof(double),
                                      (size_t)c->ncols * sizeof(double), (size_t)c->n_loc * c->d, cudaMemcpyDeviceToDevice, h->stream));
    return RVGP_OK;
}

static int halo_check(Handle* h, const rvgp_halo_ctx* c) {
    RVGP_REQUIRE(h, c != nullptr && c->ncols >= 2 && c->ncols % 2 == 0 && c->d >= 1, "halo: bad context (ncols must be even)");
    RVGP_REQUIRE(h, c->E[0] && c->E[1] && c->E[2], "halo: the three extended buffers are required");
    RVGP_REQUIRE(h, c->n_peers <= 32, "halo: at most 32 neighbour ranks");
    RVGP_REQUIRE(h, c->kptr == nullptr || (c->rotc != 2 && c->d == 2 && c->ncols % 16 == 0) ||
                        (c->rotc == 2 && c->d == 1 && c->ncols % 32 == 0),
                 "halo: the MMA plan needs d == 2 and ncols % 16 == 0, or the scalar pattern plan (rotc == 2) with ncols % 32 == 0");
    return RVGP_OK;
}

}  // namespace rvgp

using namespace rvgp;

// ---- CUDA IPC plumbing for the shared extended buffers / flag arrays ------------------------------------------------
extern "C" int rvgp_ipc_alloc(rvgp_handle_t hh, int64_t bytes, void** dptr, uint8_t* handle64) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, bytes > 0 && dptr && handle64, "ipc_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    RVGP_CUDA_OK(h, cudaMalloc(dptr, (size_t)bytes));
    RVGP_CUDA_OK(h, cudaMemset(*dptr, 0, (size_t)bytes));
    cudaIpcMemHandle_t mh;
    RVGP_CUDA_OK(h, cudaIpcGetMemHandle(&mh, *dptr));
    memcpy(handle64, &mh, 64);
    return RVGP_OK;
}

extern "C" int rvgp_ipc_open(rvgp_handle_t hh, const uint8_t* handle64, void** dptr) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, dptr && handle64, "ipc_open: bad arguments");
    cudaIpcMemHandle_t mh;
    memcpy(&mh, handle64, 64);
    RVGP_CUDA_OK(h, cudaIpcOpenMemHandle(dptr, mh, cudaIpcMemLazyEnablePeerAccess));
    return RVGP_OK;
}

extern "C" int rvgp_ipc_close(rvgp_handle_t hh, void* dptr) {
    Handle* h = H(hh);
    if (dptr) RVGP_CUDA_OK(h, cudaIpcCloseMemHandle(dptr));
    return RVGP_OK;
}

extern "C" int rvgp_ipc_free(rvgp_handle_t hh, void* dptr) {
    Handle* h = H(hh);
    if (dptr) RVGP_CUDA_OK(h, cudaFree(dptr));
    return RVGP_OK;
}

// Cross-rank barrier on the handle's stream (signal + wait at `epoch`).
extern "C" int rvgp_halo_barrier(rvgp_handle_t hh, const rvgp_halo_ctx* c, uint64_t epoch) {
    Handle* h = H(hh);
    int rc = halo_check(h, c);
    if (rc) return rc;
    return halo_sync(h, c, epoch);
}

// Y (local rows, ldy) = A X (local rows, ldx) on the row-sharded operator.  Uses epochs epoch0 .. epoch0 + 1.
extern "C" int rvgp_halo_spmm_f64(rvgp_handle_t hh, const rvgp_halo_ctx* c, uint64_t epoch0, const double* X, int64_t ldx,
                                  double* Y, int64_t ldy) {
    Handle* h = H(hh);
    int rc = halo_check(h, c);
    if (rc) return rc;
    if ((rc = halo_sync(h, c, epoch0))) return rc;                 // everyone is done reading the buffers of the last call
    if ((rc = halo_copy_in(h, c, X, ldx, 0))) return rc;
    if ((rc = halo_step(h, c, epoch0 + 1, 0, -1, 1, 1.0, 0.0, 0.0))) return rc;
    return halo_copy_out(h, c, Y, ldy, 1);
}

// Chebyshev filter of degree `degree` of the panel V (local rows) in place; same recurrence as rvgp_cheb_filter_f64.
// Uses epochs epoch0 .. epoch0 + degree.
extern "C" int rvgp_halo_cheb_filter_f64(rvgp_handle_t hh, const rvgp_halo_ctx* c, uint64_t epoch0, double* V, int64_t ldv,
                                         int degree, double lo_spec, double lo_cut, double hi) {
    Handle* h = H(hh);
    int rc = halo_check(h, c);
    if (rc) return rc;
    RVGP_REQUIRE(h, degree >= 0, "cheb_filter: degree >= 0");
    RVGP_REQUIRE(h, hi > lo_cut && lo_cut > lo_spec, "cheb_filter: need lo_spec < lo_cut < hi");
    if (degree == 0) return RVGP_OK;
    if ((rc = halo_sync(h, c, epoch0))) return rc;
    if ((rc = halo_copy_in(h, c, V, ldv, 0))) return rc;
    const double e = 0.5 * (hi - lo_cut), cc = 0.5 * (hi + lo_cut);
    const double sigma1 = e / (lo_spec - cc), tau = 2.0 / sigma1;
    double sigma = sigma1;
    unsigned long long ep = epoch0;
    if ((rc = halo_step(h, c, ++ep, 0, -1, 1, sigma1 / e, -cc * sigma1 / e, 0.0))) return rc;
    int prev = 0, cur = 1;
    for (int i = 2; i <= degree; ++i) {
        const double sn = 1.0 / (tau - sigma);
        const int nxt = 3 - prev - cur;
        if ((rc = halo_step(h, c, ++ep, cur, prev, nxt, 2.0 * sn / e, -2.0 * sn * cc / e, -sigma * sn))) return rc;
        sigma = sn;
        prev = cur;
        cur = nxt;
    }
    return halo_copy_out(h, c, V, ldv, cur);
}
