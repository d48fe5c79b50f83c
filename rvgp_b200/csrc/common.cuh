// Shared helpers for the rvgp_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/rvgp_b200.h"

namespace rvgp {

struct Handle {
    int device;
    cudaStream_t stream;
    int sm_count;
    char last_error[512];
    unsigned long long launches;  // kernels launched through this handle (bench.py: gpu_launches)
    int spmm_lpr;                 // tuning: lanes per block row in the gather SpMM (0 = auto)
    int spmm_v1;                  // tuning: force the 64-bit-load kernel
    const int* rowlist;           // transient: row list of rvgp_bsr_spmm_rows_f64
    int nlist;
    int spmm_stage;               // tuning: stage block values / indices of a CTA's rows in shared memory
    int dgemm_dmma;               // tuning: FP64 tensor (mma.sync m8n8k4) instead of the DFMA register tile
    int mma_gpw;                  // tuning: row groups per warp in the MMA SpMM (0 = default)
    int mma_variant;              // tuning: pipeline / occupancy variant of the native-layout MMA SpMM
    int mma_variant_n2;           // tuning: the same for panels of 32 native columns (2 chunks)
    int mma_prefetch;             // tuning: L2 prefetch distance (warp iterations) of the native-layout MMA SpMM
    int mma_stream_policy;        // tuning: bit0 = no-L1-allocate W loads, bit1 = no-L1-allocate Y stores (MMA SpMM)
    // CUDA graph of the rank-k GP evaluation (rvgp_gp_lowrank_eval_f64): ~110 tiny launches replayed with one cudaGraphLaunch
    cudaGraphExec_t gp_graph;     // nullptr until captured
    const void* gp_graph_key[6];  // k (as pointer-sized int) and the five buffers the captured launches were recorded with
    unsigned long long gp_graph_launches;
    cudaStream_t gp_stream;       // blocking side stream used when the handle follows the legacy default stream (not capturable)
    int gp_graph_off;             // 1: capture failed once, run eagerly from now on (or rvgp_set_option("gp_graph", 0))
    int dgemm_pipe_attr;          // 1 once the pipelined DGEMM's dynamic shared memory limit is raised on this handle's device
};

inline Handle* H(rvgp_handle_t h) { return reinterpret_cast<Handle*>(h); }

inline int set_error(Handle* h, int code, const char* fmt, const char* a = "", const char* b = "") {
    if (h) snprintf(h->last_error, sizeof(h->last_error), fmt, a, b);
    return code;
}

#define RVGP_CUDA_OK(h, expr)                                                               \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return rvgp::set_error((h), RVGP_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

#define RVGP_LAUNCH_OK(h, name)                                                             \
    do {                                                                                    \
        (h)->launches++;                                                                    \
        cudaError_t _e = cudaGetLastError();                                                \
        if (_e != cudaSuccess)                                                              \
            return rvgp::set_error((h), RVGP_ERR_CUDA, "launch %s: %s", name, cudaGetErrorString(_e)); \
    } while (0)

#define RVGP_REQUIRE(h, cond, msg)                                                          \
    do {                                                                                    \
        if (!(cond)) return rvgp::set_error((h), RVGP_ERR_BAD_ARG, "%s (%s)", msg, #cond);  \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace rvgp
