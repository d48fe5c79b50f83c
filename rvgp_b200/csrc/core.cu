// Handle management for librvgp_b200.so (see include/rvgp_b200.h).
#include "common.cuh"

using namespace rvgp;

extern "C" int rvgp_version(void) { return 100; }

extern "C" int rvgp_create(int device, rvgp_handle_t* out) {
    if (!out) return RVGP_ERR_BAD_ARG;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return RVGP_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return RVGP_ERR_CUDA;
    Handle* h = new Handle();
    h->device = device;
    h->stream = nullptr;
    h->launches = 0;
    h->spmm_lpr = 0;
    h->spmm_v1 = 0;
    h->rowlist = nullptr;
    h->nlist = 0;
    h->spmm_stage = 0;   // measured: no gain (tools/profile_stage.py); value loads are not the limiter
    h->dgemm_dmma = 2;   // 0 DFMA register tile, 1 DMMA (17 vs 14 TFLOP/s on the tall-skinny Gram, tools/ncu_dgemm.py), 2 cp.async-pipelined DMMA
    h->mma_gpw = 0;
    // measured on B200 at C4 size (tools/profile_mma.py, profiles/r01_spmm_mma_sweep.txt): 256-thread CTAs x 2, register ring
    // of 2 k-steps, persistent warp-strided schedule, streams evict-first in L2, L2 prefetch one group ahead
    h->mma_stream_policy = 7;
    h->mma_variant = 1;
    h->mma_variant_n2 = 1;   // (2 chunks, ring 2, 256 threads x 3 CTAs/SM): 0.423 vs 0.448 ms for the 64-column L panels (profiles/r02_k9_sweep.txt)
    h->mma_prefetch = 1;
    h->gp_graph = nullptr;
    h->gp_graph_launches = 0;
    h->gp_stream = nullptr;
    h->gp_graph_off = 0;
    h->dgemm_pipe_attr = 0;
    for (int i = 0; i < 6; ++i) h->gp_graph_key[i] = nullptr;
    h->last_error[0] = 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete h; return RVGP_ERR_CUDA; }
    h->sm_count = prop.multiProcessorCount;
    *out = h;
    return RVGP_OK;
}

extern "C" int rvgp_destroy(rvgp_handle_t hh) {
    if (hh) {
        if (H(hh)->gp_graph) cudaGraphExecDestroy(H(hh)->gp_graph);
        if (H(hh)->gp_stream) cudaStreamDestroy(H(hh)->gp_stream);
    }
    delete H(hh);
    return RVGP_OK;
}

extern "C" int rvgp_set_stream(rvgp_handle_t hh, void* s) {
    if (!hh) return RVGP_ERR_BAD_ARG;
    H(hh)->stream = reinterpret_cast<cudaStream_t>(s);
    return RVGP_OK;
}

extern "C" const char* rvgp_last_error(rvgp_handle_t hh) { return hh ? H(hh)->last_error : "null handle"; }
extern "C" int rvgp_sm_count(rvgp_handle_t hh) { return hh ? H(hh)->sm_count : 0; }
extern "C" unsigned long long rvgp_launch_count(rvgp_handle_t hh) { return hh ? H(hh)->launches : 0ull; }

// Tuning knobs (benchmarks / experiments).  "spmm_lpr": lanes per block row in rvgp_bsr_spmm_f64 (0 auto, 8, 16, 32).
extern "C" int rvgp_set_option(rvgp_handle_t hh, const char* key, int value) {
    if (!hh || !key) return RVGP_ERR_BAD_ARG;
    if (strcmp(key, "spmm_lpr") == 0) { H(hh)->spmm_lpr = value; return RVGP_OK; }
    if (strcmp(key, "spmm_v1") == 0) { H(hh)->spmm_v1 = value; return RVGP_OK; }
    if (strcmp(key, "spmm_stage") == 0) { H(hh)->spmm_stage = value; return RVGP_OK; }
    if (strcmp(key, "dgemm_dmma") == 0) { H(hh)->dgemm_dmma = value; return RVGP_OK; }
    if (strcmp(key, "mma_variant") == 0) { H(hh)->mma_variant = value; return RVGP_OK; }
    if (strcmp(key, "mma_variant_n2") == 0) { H(hh)->mma_variant_n2 = value; return RVGP_OK; }
    if (strcmp(key, "mma_prefetch") == 0) { H(hh)->mma_prefetch = value; return RVGP_OK; }
    if (strcmp(key, "mma_gpw") == 0) { H(hh)->mma_gpw = value; return RVGP_OK; }
    if (strcmp(key, "mma_stream_policy") == 0) { H(hh)->mma_stream_policy = value; return RVGP_OK; }
    if (strcmp(key, "gp_graph") == 0) { H(hh)->gp_graph_off = value ? 0 : 1; return RVGP_OK; }
    return set_error(H(hh), RVGP_ERR_BAD_ARG, "unknown option %s%s", key);
}
