// common.cuh - context, error plumbing and small helpers shared by the kernels.
//
// The library is built from three translation units that nvcc compiles side by side (spkdiar.cu: context,
// statistics, batched scoring; spkdiar_gw.cu: the growing-window search; spkdiar_cluster.cu: the clustering
// engines).  Every kernel is `static __global__` (each unit instantiates only what it launches), the shared
// device code lives in the .cuh headers.
#pragma once

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/spkdiar.h"
#include "layout.cuh"

namespace spk {

constexpr int D39 = SPKDIAR_DIM;
using L39 = Layout<D39>;
constexpr int REC = L39::REC;                 // 820 doubles per record
static_assert(REC == SPKDIAR_RECORD, "header / kernel record size disagree");

}  // namespace spk

struct spkdiar_ctx {
    int device = 0;
    int sms = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    char err[512] = {0};
    int64_t launches = 0;
    bool prof = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double prof_ms[SPKDIAR_NPROF] = {0};
    int64_t prof_n[SPKDIAR_NPROF] = {0};
    // device-memory cache: cudaMalloc / cudaFree of the multi-GB statistics arrays cost
    // hundreds of milliseconds per recording; blocks are kept and reused by later calls
    struct Block { void* p; size_t bytes; bool used; };
    std::vector<Block> pool;
    // streams of concurrent growing-window searches (spkdiar_gw_run_multi)
    cudaStream_t aux[SPKDIAR_MAX_RUNS] = {nullptr};
    cudaEvent_t aux_done[SPKDIAR_MAX_RUNS] = {nullptr};
    int naux = 0;
    // spkdiar_ctx_exec: calls run on another stream / with an SM limit (queued behind an asynchronous search)
    cudaStream_t base_stream = nullptr;
    int sms_limit = 0;
    bool gw_busy = false;         // an asynchronous search object is open
};

struct spkdiar_feat {
    spkdiar_ctx* ctx = nullptr;
    const float* x = nullptr;     // (n, dim) frame-major fp32 in HBM
    bool own_x = false;
    int64_t n = 0;
    int32_t dim = 0;
    double* P = nullptr;          // (n + 1) block-local prefix records, lane-paired layout
    double2* C = nullptr;         // (ntiles + 1) double-double block prefix records
    double* shift = nullptr;      // 40 doubles: per-file shift subtracted before accumulation
    double* tile = nullptr;       // per-block totals
    double2* chunk = nullptr;     // per-chunk double-double totals of the block scan
    int64_t ntiles = 0;
    // packed batch of recordings (spkdiar_features_upload_batch): nrec > 0, n = packed rows
    int32_t nrec = 0;
    void* tab = nullptr;          // device RecTab[nrec]
    int64_t max_tiles = 0;        // blocks of the longest recording
    // spkdiar_features_upload_frames: P / C / tile / chunk are allocated and built on first use (P == nullptr
    // until then); cluster records come straight from the frames meanwhile (K5, stats.cuh)
};

namespace spk {

// spkdiar.cu: build the window statistics of a frames-only handle if they are missing (0 = ok)
int ensure_stats(spkdiar_feat* f);
// spkdiar.cu: records of the ranges [seg_a[k], seg_b[k]) (host arrays) into rec (device), from the frames (K5)
int direct_records(spkdiar_feat* f, const int64_t* seg_a, const int64_t* seg_b, int64_t n, double* rec);

inline int set_err(spkdiar_ctx* c, int code, const char* fmt, ...) {
    if (c) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(c->err, sizeof(c->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

#define SPK_CUDA(ctx, call)                                                        \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess)                                                    \
            return spk::set_err((ctx), SPKDIAR_E_CUDA, "%s failed: %s (%s:%d)", #call, \
                                cudaGetErrorString(e__), __FILE__, __LINE__);      \
    } while (0)

// scope timer for one kernel class (only when profiling is on)
struct Prof {
    spkdiar_ctx* c; int k; int64_t n0;
    Prof(spkdiar_ctx* ctx, int klass) : c(ctx), k(klass), n0(ctx->launches) {
        if (c->prof) cudaEventRecord(c->ev0, c->stream);
    }
    ~Prof() {
        if (c->prof) {
            cudaEventRecord(c->ev1, c->stream);
            cudaEventSynchronize(c->ev1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, c->ev0, c->ev1);
            c->prof_ms[k] += ms;
            c->prof_n[k] += c->launches - n0;
        }
    }
};

// Cached device allocation.  Every API call is synchronous on return, so a block freed
// by one call is idle by the time the next call reuses it.
inline cudaError_t pool_alloc(spkdiar_ctx* c, size_t bytes, void** out) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    int best = -1;
    for (int i = 0; i < (int)c->pool.size(); ++i) {
        const spkdiar_ctx::Block& b = c->pool[i];
        if (!b.used && b.bytes >= bytes && b.bytes <= 2 * bytes + (1u << 20) &&
            (best < 0 || b.bytes < c->pool[best].bytes)) best = i;
    }
    if (best >= 0) { c->pool[best].used = true; *out = c->pool[best].p; return cudaSuccess; }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {                       // give cached blocks back and retry once
        cudaGetLastError();
        for (size_t i = 0; i < c->pool.size();) {
            if (!c->pool[i].used) { cudaFree(c->pool[i].p); c->pool.erase(c->pool.begin() + i); }
            else ++i;
        }
        e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return e;
    }
    c->pool.push_back({p, bytes, true});
    *out = p;
    return cudaSuccess;
}
inline void pool_free(spkdiar_ctx* c, const void* p) {
    if (!p) return;
    for (auto& b : c->pool) if (b.p == p) { b.used = false; return; }
}
inline void pool_destroy(spkdiar_ctx* c) {
    for (auto& b : c->pool) cudaFree(b.p);
    c->pool.clear();
}

template <class T>
struct DevBuf {                      // RAII device buffer from the context's cache
    spkdiar_ctx* c = nullptr;
    T* p = nullptr;
    ~DevBuf() { if (p) pool_free(c, p); }
    cudaError_t alloc(spkdiar_ctx* ctx, size_t count) {
        c = ctx;
        return pool_alloc(ctx, count * sizeof(T) + 16, (void**)&p);
    }
};

// grid size of a throughput kernel: enough CTAs for the tasks, at most ctas_per_sm per (allowed) SM
static inline int grid_for(const spkdiar_ctx* c, int64_t tasks, int per_cta, int ctas_per_sm) {
    int64_t want = (tasks + per_cta - 1) / per_cta;
    const int sms = c->sms_limit > 0 ? (c->sms < c->sms_limit ? c->sms : c->sms_limit) : c->sms;
    const int64_t cap = (int64_t)sms * ctas_per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}
static inline int check_metric(spkdiar_ctx* c, int metric) {
    if (metric != SPKDIAR_GLR && metric != SPKDIAR_BIC && metric != SPKDIAR_KL2)
        return set_err(c, SPKDIAR_E_ARG, "unknown metric %d", metric);
    return SPKDIAR_OK;
}
// per-translation-unit kernel attributes (dynamic shared memory sizes), called once by spkdiar_create
cudaError_t gw_configure();
cudaError_t gw_set_dim(int d);          // the feature dimension symbol of each translation unit (ldl.cuh: c_dim)
cudaError_t cluster_set_dim(int d);
cudaError_t cluster_configure();
cudaError_t cluster_batch_configure();

__device__ __forceinline__ double d_nan() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ double d_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

}  // namespace spk
