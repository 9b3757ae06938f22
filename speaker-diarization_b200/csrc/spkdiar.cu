// spkdiar.cu - libspkdiar.so: the C-ABI of include/spkdiar.h over the sm_100a kernels.
//
// One translation unit: context + frame statistics (stats.cuh), batched scoring
// (score.cuh), the persistent growing-window driver (gw.cuh) and the
// agglomerative clustering engine (cluster.cuh).
//
// Build (see __graft_entry__.build()):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo
//        -shared -Xcompiler -fPIC -o libspkdiar.so spkdiar.cu
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <limits>
#include <mutex>
#include <new>
#include <vector>

#include "common.cuh"
#include "stats.cuh"
#include "score.cuh"

using namespace spk;

static thread_local char g_create_err[512] = "";

extern "C" {

int spkdiar_abi_version(void) { return SPKDIAR_ABI_VERSION; }

int spkdiar_create(int device, void* stream, spkdiar_ctx** out) {
    if (!out) return SPKDIAR_E_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        snprintf(g_create_err, sizeof(g_create_err), "no CUDA device: %s",
                 e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return SPKDIAR_E_NODEVICE;
    }
    if (device < 0 || device >= ndev) {
        snprintf(g_create_err, sizeof(g_create_err), "device %d out of range (have %d)", device, ndev);
        return SPKDIAR_E_ARG;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        snprintf(g_create_err, sizeof(g_create_err), "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
        return SPKDIAR_E_CUDA;
    }
    if (prop.major != 10) {
        snprintf(g_create_err, sizeof(g_create_err),
                 "device %d is sm_%d%d; this library carries sm_100a code only (no fallback)",
                 device, prop.major, prop.minor);
        return SPKDIAR_E_NODEVICE;
    }
    spkdiar_ctx* c = new (std::nothrow) spkdiar_ctx();
    if (!c) return SPKDIAR_E_NOMEM;
    c->device = device;
    c->sms = prop.multiProcessorCount;
    if ((e = cudaSetDevice(device)) != cudaSuccess) goto fail;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) goto fail;
        c->own_stream = true;
    }
    if ((e = cudaEventCreate(&c->ev0)) != cudaSuccess) goto fail;
    if ((e = cudaEventCreate(&c->ev1)) != cudaSuccess) goto fail;
    {
        uint8_t row[REC], col[REC];
        fill_lut(row, col);
        if ((e = cudaMemcpyToSymbol(c_row, row, sizeof(row))) != cudaSuccess) goto fail;
        if ((e = cudaMemcpyToSymbol(c_col, col, sizeof(col))) != cudaSuccess) goto fail;
    }
    if ((e = cudaFuncSetAttribute(win_kl2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(SC_WARPS * sizeof(Kl2Scratch)))) != cudaSuccess) goto fail;
    if ((e = cudaFuncSetAttribute(pair_kl2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(SC_WARPS * sizeof(Kl2Scratch)))) != cudaSuccess) goto fail;
    if ((e = cudaFuncSetAttribute(win_terms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SC_SMEM)) != cudaSuccess) goto fail;
    if ((e = cudaFuncSetAttribute(pair_terms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SC_SMEM)) != cudaSuccess) goto fail;
    if ((e = gw_configure()) != cudaSuccess) goto fail;
    if ((e = cluster_configure()) != cudaSuccess) goto fail;
    if ((e = cluster_batch_configure()) != cudaSuccess) goto fail;
    *out = c;
    return SPKDIAR_OK;
fail:
    snprintf(g_create_err, sizeof(g_create_err), "context setup failed: %s", cudaGetErrorString(e));
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return SPKDIAR_E_CUDA;
}

void spkdiar_destroy(spkdiar_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    for (int r = 0; r < c->naux; ++r) { cudaStreamSynchronize(c->aux[r]); cudaEventDestroy(c->aux_done[r]); cudaStreamDestroy(c->aux[r]); }
    pool_destroy(c);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* spkdiar_last_error(const spkdiar_ctx* c) { return c ? c->err : g_create_err; }
int64_t spkdiar_launch_count(const spkdiar_ctx* c) { return c ? c->launches : 0; }
int spkdiar_sm_count(const spkdiar_ctx* c) { return c ? c->sms : 0; }

int spkdiar_profile_enable(spkdiar_ctx* c, int enable) {
    if (!c) return SPKDIAR_E_ARG;
    c->prof = enable != 0;
    for (int k = 0; k < SPKDIAR_NPROF; ++k) { c->prof_ms[k] = 0.0; c->prof_n[k] = 0; }
    return SPKDIAR_OK;
}

int spkdiar_profile_read(const spkdiar_ctx* c, double* ms, int64_t* launches) {
    if (!c) return SPKDIAR_E_ARG;
    for (int k = 0; k < SPKDIAR_NPROF; ++k) {
        if (ms) ms[k] = c->prof_ms[k];
        if (launches) launches[k] = c->prof_n[k];
    }
    return SPKDIAR_OK;
}

// ---- features + K1 ------------------------------------------------------------------------

// One feature dimension per device at a time: the kernels read it from a __constant__ symbol (ldl.cuh: c_dim), which
// is set here in every translation unit whenever the first handle of another dimension is made on a device that
// has no live handle (every call of the library is synchronous, so nothing is in flight then).
static std::mutex g_dim_mu;
static int g_dev_dim[64];          // 0 = still the compiled default (39)
static int g_dev_live[64];
static int device_dim_acquire(spkdiar_ctx* c, int32_t dim) {
    std::lock_guard<std::mutex> lock(g_dim_mu);
    const int dev = c->device & 63;
    const int cur = g_dev_dim[dev] ? g_dev_dim[dev] : D39;
    if (cur != dim) {
        if (g_dev_live[dev] > 0)
            return set_err(c, SPKDIAR_E_UNSUPPORTED, "feature dimension %d while handles of dimension %d are alive on device %d: "
                           "one dimension per device at a time", dim, cur, c->device);
        cudaError_t e = set_dim_symbol(dim);
        if (e == cudaSuccess) e = gw_set_dim(dim);
        if (e == cudaSuccess) e = cluster_set_dim(dim);
        if (e != cudaSuccess) return set_err(c, SPKDIAR_E_CUDA, "setting the feature dimension failed: %s", cudaGetErrorString(e));
        g_dev_dim[dev] = dim;
    }
    g_dev_live[dev] += 1;
    return SPKDIAR_OK;
}
static void device_dim_release(spkdiar_ctx* c) {
    std::lock_guard<std::mutex> lock(g_dim_mu);
    const int dev = c->device & 63;
    if (g_dev_live[dev] > 0) g_dev_live[dev] -= 1;
}

// the arrays of the window statistics (K1): 6,560 B per frame + the block level
static int stats_alloc(spkdiar_feat* f) {
    spkdiar_ctx* c = f->ctx;
    const int32_t nrec = f->nrec;
    cudaError_t e;
    if ((e = pool_alloc(c, (size_t)(f->n + 1) * REC * sizeof(double), (void**)&f->P)) != cudaSuccess ||
        (e = pool_alloc(c, (size_t)(f->ntiles + 1) * REC * sizeof(double2), (void**)&f->C)) != cudaSuccess ||
        (e = pool_alloc(c, (size_t)std::max(nrec, 1) * K1_CHUNKS * REC * sizeof(double2), (void**)&f->chunk)) != cudaSuccess ||
        (e = pool_alloc(c, (size_t)std::max<int64_t>(f->ntiles, 1) * REC * sizeof(double), (void**)&f->tile)) != cudaSuccess) {
        pool_free(c, f->P); pool_free(c, f->C); pool_free(c, f->chunk); pool_free(c, f->tile);
        f->P = nullptr; f->C = nullptr; f->chunk = nullptr; f->tile = nullptr;
        return set_err(c, e == cudaErrorMemoryAllocation ? SPKDIAR_E_NOMEM : SPKDIAR_E_CUDA,
                       "device allocation for the statistics of %lld frames failed: %s", (long long)f->n, cudaGetErrorString(e));
    }
    return SPKDIAR_OK;
}

static int feat_alloc(spkdiar_ctx* c, int64_t n, int32_t dim, spkdiar_feat** out, int32_t nrec = 0, bool with_stats = true) {
    if (!c || !out) return SPKDIAR_E_ARG;
    *out = nullptr;
    if (n < 0) return set_err(c, SPKDIAR_E_ARG, "negative frame count %lld", (long long)n);
    if (dim < 1 || dim > D39)
        return set_err(c, SPKDIAR_E_UNSUPPORTED,
                       "feature dimension %d: the kernels handle 1..%d dimensions (fconfig.cfg has %d) and there is no fallback",
                       dim, D39, D39);
    SPK_CUDA(c, cudaSetDevice(c->device));
    if (int rc = device_dim_acquire(c, dim)) return rc;
    spkdiar_feat* f = new (std::nothrow) spkdiar_feat();
    if (!f) { device_dim_release(c); return set_err(c, SPKDIAR_E_NOMEM, "host allocation failed"); }
    f->ctx = c; f->n = n; f->dim = dim;
    f->ntiles = (n + K1_TILE - 1) / K1_TILE;
    f->nrec = nrec;
    cudaError_t e;
    if ((e = pool_alloc(c, (size_t)std::max(nrec, 1) * K1_XS * sizeof(double), (void**)&f->shift)) != cudaSuccess) {
        device_dim_release(c);
        delete f;
        return set_err(c, e == cudaErrorMemoryAllocation ? SPKDIAR_E_NOMEM : SPKDIAR_E_CUDA,
                       "device allocation for %lld frames failed: %s", (long long)n, cudaGetErrorString(e));
    }
    if (with_stats) {
        if (int rc = stats_alloc(f)) { pool_free(c, f->shift); device_dim_release(c); delete f; return rc; }
    }
    *out = f;
    return SPKDIAR_OK;
}

extern "C++" {
namespace spk {
int ensure_stats(spkdiar_feat* f) {
    if (!f) return SPKDIAR_E_ARG;
    if (f->P) return SPKDIAR_OK;
    if (int rc = stats_alloc(f)) return rc;
    return spkdiar_stats_build(f);
}

int direct_records(spkdiar_feat* f, const int64_t* seg_a, const int64_t* seg_b, int64_t n, double* rec) {
    spkdiar_ctx* c = f->ctx;
    if (n <= 0) return SPKDIAR_OK;
    // tasks of at most K5_SPAN frames; a range that needs several gets partial records and a reduction
    std::vector<int64_t> ta, tb, td, rs, rp, rn;
    int64_t nparts = 0;
    for (int64_t k = 0; k < n; ++k) {
        const int64_t len = seg_b[k] - seg_a[k];
        const int64_t parts = len <= K5_SPAN ? 1 : (len + K5_SPAN - 1) / K5_SPAN;
        if (parts == 1) { ta.push_back(seg_a[k]); tb.push_back(seg_b[k]); td.push_back(k); continue; }
        rs.push_back(k); rp.push_back(nparts); rn.push_back(parts);
        for (int64_t p = 0; p < parts; ++p) {
            ta.push_back(seg_a[k] + p * K5_SPAN);
            tb.push_back(std::min(seg_b[k], seg_a[k] + (p + 1) * K5_SPAN));
            td.push_back(-(nparts + p) - 1);
        }
        nparts += parts;
    }
    const int64_t nt = (int64_t)ta.size(), nr = (int64_t)rs.size();
    std::vector<int64_t> host((size_t)(3 * nt + 3 * nr));
    std::copy(ta.begin(), ta.end(), host.begin());
    std::copy(tb.begin(), tb.end(), host.begin() + nt);
    std::copy(td.begin(), td.end(), host.begin() + 2 * nt);
    std::copy(rs.begin(), rs.end(), host.begin() + 3 * nt);
    std::copy(rp.begin(), rp.end(), host.begin() + 3 * nt + nr);
    std::copy(rn.begin(), rn.end(), host.begin() + 3 * nt + 2 * nr);
    DevBuf<int64_t> dtask; DevBuf<double> part;
    SPK_CUDA(c, dtask.alloc(c, host.size()));
    SPK_CUDA(c, part.alloc(c, (size_t)std::max<int64_t>(nparts, 1) * REC));
    SPK_CUDA(c, cudaMemcpyAsync(dtask.p, host.data(), host.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    {
        Prof p(c, SPKDIAR_PROF_STATS);
        k5_direct<<<(unsigned)nt, K1_THREADS, 0, c->stream>>>(f->x, f->shift, dtask.p, nt, rec, part.p);
        c->launches += 1;
        if (nr > 0) {
            k5_reduce<<<(unsigned)nr, K1_THREADS, 0, c->stream>>>(part.p, dtask.p + 3 * nt, nr, rec);
            c->launches += 1;
        }
    }
    SPK_CUDA(c, cudaGetLastError());
    SPK_CUDA(c, cudaStreamSynchronize(c->stream));          // `host` and the two buffers are temporaries
    return SPKDIAR_OK;
}
}  // namespace spk
}  // extern "C++"

int spkdiar_stats_build(spkdiar_feat* f) {
    if (!f) return SPKDIAR_E_ARG;
    spkdiar_ctx* c = f->ctx;
    SPK_CUDA(c, cudaSetDevice(c->device));
    if (!f->P) { if (int rc = stats_alloc(f)) return rc; }
    {
        Prof p(c, SPKDIAR_PROF_STATS);
        if (f->nrec > 0) {
            const RecTab* tab = (const RecTab*)f->tab;
            const dim3 sg((K1_CHUNKS * REC + 127) / 128, (unsigned)f->nrec);
            k1_shift_batch<<<(unsigned)f->nrec, 1024, 0, c->stream>>>(f->x, tab, f->shift);
            k1_tile_write_batch<<<dim3((unsigned)f->max_tiles, (unsigned)f->nrec), K1_THREADS, 0, c->stream>>>(
                f->x, tab, f->shift, f->tile, f->P);
            k1_chunk_sums_batch<<<sg, 128, 0, c->stream>>>(f->tile, tab, f->chunk);
            k1_chunk_scan_batch<<<sg, 128, 0, c->stream>>>(f->tile, tab, f->chunk, f->C);
            c->launches += 4;
        } else if (f->n == 0) {
            SPK_CUDA(c, cudaMemsetAsync(f->P, 0, REC * sizeof(double), c->stream));
            SPK_CUDA(c, cudaMemsetAsync(f->C, 0, REC * sizeof(double2), c->stream));
        } else {
            k1_shift<<<1, 1024, 0, c->stream>>>(f->x, f->n, f->shift);
            k1_tile_write<<<(unsigned)f->ntiles, K1_THREADS, 0, c->stream>>>(f->x, f->n, f->shift, f->tile, f->P);
            k1_chunk_sums<<<(K1_CHUNKS * REC + 127) / 128, 128, 0, c->stream>>>(f->tile, f->ntiles, f->chunk);
            k1_chunk_scan<<<(K1_CHUNKS * REC + 127) / 128, 128, 0, c->stream>>>(f->tile, f->ntiles, f->chunk, f->C);
            c->launches += 4;
        }
    }
    SPK_CUDA(c, cudaGetLastError());
    SPK_CUDA(c, cudaStreamSynchronize(c->stream));
    return SPKDIAR_OK;
}

static int upload_impl(spkdiar_ctx* c, const float* frames, int64_t n, int32_t dim, spkdiar_feat** out, bool with_stats);

int spkdiar_features_upload(spkdiar_ctx* c, const float* frames, int64_t n, int32_t dim, spkdiar_feat** out) {
    return upload_impl(c, frames, n, dim, out, true);
}
int spkdiar_features_upload_frames(spkdiar_ctx* c, const float* frames, int64_t n, int32_t dim, spkdiar_feat** out) {
    return upload_impl(c, frames, n, dim, out, false);
}

static int upload_impl(spkdiar_ctx* c, const float* frames, int64_t n, int32_t dim, spkdiar_feat** out, bool with_stats) {
    if (!c || !out || (!frames && n > 0)) return c ? set_err(c, SPKDIAR_E_ARG, "null argument") : SPKDIAR_E_ARG;
    int rc = feat_alloc(c, n, dim, out, 0, with_stats);
    if (rc) return rc;
    spkdiar_feat* f = *out;
    float* dx = nullptr;
    cudaError_t e = pool_alloc(c, (size_t)std::max<int64_t>(n, 1) * D39 * sizeof(float), (void**)&dx);
    if (e != cudaSuccess) {
        spkdiar_features_free(f); *out = nullptr;
        return set_err(c, SPKDIAR_E_NOMEM, "device allocation for the frames failed: %s", cudaGetErrorString(e));
    }
    f->x = dx; f->own_x = true;
    {
        Prof p(c, SPKDIAR_PROF_H2D);
        if (n > 0 && dim == D39) e = cudaMemcpyAsync(dx, frames, (size_t)n * dim * sizeof(float), cudaMemcpyHostToDevice, c->stream);
        else if (n > 0) {                                   // fewer dimensions: rows zero-padded to 39 columns
            e = cudaMemsetAsync(dx, 0, (size_t)n * D39 * sizeof(float), c->stream);
            if (e == cudaSuccess)
                e = cudaMemcpy2DAsync(dx, D39 * sizeof(float), frames, (size_t)dim * sizeof(float), (size_t)dim * sizeof(float),
                                      (size_t)n, cudaMemcpyHostToDevice, c->stream);
        }
    }
    if (e != cudaSuccess) {
        spkdiar_features_free(f); *out = nullptr;
        return set_err(c, SPKDIAR_E_CUDA, "feature upload failed: %s", cudaGetErrorString(e));
    }
    if (with_stats) {
        rc = spkdiar_stats_build(f);
    } else {
        // the shift alone: the records of K5 are accumulated around it, like those of K1
        if (n > 0) { k1_shift<<<1, 1024, 0, c->stream>>>(f->x, f->n, f->shift); c->launches += 1; }
        else cudaMemsetAsync(f->shift, 0, K1_XS * sizeof(double), c->stream);
        rc = cudaStreamSynchronize(c->stream) == cudaSuccess ? SPKDIAR_OK : set_err(c, SPKDIAR_E_CUDA, "feature upload failed");
    }
    if (rc) { spkdiar_features_free(f); *out = nullptr; }
    return rc;
}

int spkdiar_features_upload_batch(spkdiar_ctx* c, const float* const* frames, const int64_t* n, int32_t nrec,
                                  int32_t dim, spkdiar_feat** out, int64_t* base_out) {
    if (!c || !out || !frames || !n || !base_out || nrec < 1)
        return c ? set_err(c, SPKDIAR_E_ARG, "null argument / empty batch") : SPKDIAR_E_ARG;
    if (nrec > 65535) return set_err(c, SPKDIAR_E_ARG, "at most 65535 recordings per batch");
    std::vector<RecTab> tab((size_t)nrec);
    int64_t rows = 0, max_tiles = 1;
    for (int32_t r = 0; r < nrec; ++r) {
        if (n[r] < 0 || (n[r] > 0 && !frames[r])) return set_err(c, SPKDIAR_E_ARG, "recording %d: bad frame count / null matrix", r);
        const int64_t tiles = n[r] / K1_TILE + 1;          // the last block is partial or empty (stats.cuh)
        tab[r].base = rows; tab[r].n = n[r];
        base_out[r] = rows;
        rows += tiles * K1_TILE;
        max_tiles = std::max(max_tiles, tiles);
    }
    int rc = feat_alloc(c, rows, dim, out, nrec);
    if (rc) return rc;
    spkdiar_feat* f = *out;
    f->nrec = nrec; f->max_tiles = max_tiles;
    float* dx = nullptr;
    cudaError_t e = pool_alloc(c, (size_t)rows * D39 * sizeof(float), (void**)&dx);
    if (e == cudaSuccess) { f->x = dx; f->own_x = true; e = pool_alloc(c, (size_t)nrec * sizeof(RecTab), &f->tab); }
    if (e != cudaSuccess) {
        spkdiar_features_free(f); *out = nullptr;
        return set_err(c, SPKDIAR_E_NOMEM, "device allocation for the packed frames failed: %s", cudaGetErrorString(e));
    }
    {
        Prof p(c, SPKDIAR_PROF_H2D);
        e = cudaMemsetAsync(dx, 0, (size_t)rows * D39 * sizeof(float), c->stream);      // the padding rows (and columns)
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(f->tab, tab.data(), (size_t)nrec * sizeof(RecTab), cudaMemcpyHostToDevice, c->stream);
        for (int32_t r = 0; r < nrec && e == cudaSuccess; ++r)
            if (n[r] > 0)
                e = dim == D39 ? cudaMemcpyAsync(dx + tab[r].base * D39, frames[r], (size_t)n[r] * dim * sizeof(float),
                                                 cudaMemcpyHostToDevice, c->stream)
                               : cudaMemcpy2DAsync(dx + tab[r].base * D39, D39 * sizeof(float), frames[r], (size_t)dim * sizeof(float),
                                                   (size_t)dim * sizeof(float), (size_t)n[r], cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);       // `tab` is a host temporary
    }
    if (e != cudaSuccess) {
        spkdiar_features_free(f); *out = nullptr;
        return set_err(c, SPKDIAR_E_CUDA, "packed feature upload failed: %s", cudaGetErrorString(e));
    }
    rc = spkdiar_stats_build(f);
    if (rc) { spkdiar_features_free(f); *out = nullptr; }
    return rc;
}

int spkdiar_features_adopt(spkdiar_ctx* c, const float* dev_frames, int64_t n, int32_t dim, spkdiar_feat** out) {
    if (!c || !out || (!dev_frames && n > 0)) return c ? set_err(c, SPKDIAR_E_ARG, "null argument") : SPKDIAR_E_ARG;
    int rc = feat_alloc(c, n, dim, out);
    if (rc) return rc;
    if (dim == D39) {
        (*out)->x = dev_frames;
    } else {                                                // fewer dimensions: an own, zero-padded copy
        float* dx = nullptr;
        cudaError_t e = pool_alloc(c, (size_t)std::max<int64_t>(n, 1) * D39 * sizeof(float), (void**)&dx);
        if (e == cudaSuccess) { (*out)->x = dx; (*out)->own_x = true; e = cudaMemsetAsync(dx, 0, (size_t)std::max<int64_t>(n, 1) * D39 * sizeof(float), c->stream); }
        if (e == cudaSuccess && n > 0)
            e = cudaMemcpy2DAsync(dx, D39 * sizeof(float), dev_frames, (size_t)dim * sizeof(float), (size_t)dim * sizeof(float),
                                  (size_t)n, cudaMemcpyDeviceToDevice, c->stream);
        if (e != cudaSuccess) {
            spkdiar_features_free(*out); *out = nullptr;
            return set_err(c, SPKDIAR_E_CUDA, "padding the adopted frames failed: %s", cudaGetErrorString(e));
        }
    }
    rc = spkdiar_stats_build(*out);
    if (rc) { spkdiar_features_free(*out); *out = nullptr; }
    return rc;
}

int spkdiar_features_free(spkdiar_feat* f) {
    if (!f) return SPKDIAR_OK;
    cudaSetDevice(f->ctx->device);
    cudaStreamSynchronize(f->ctx->stream);
    pool_free(f->ctx, f->P);
    pool_free(f->ctx, f->C);
    pool_free(f->ctx, f->shift);
    pool_free(f->ctx, f->chunk);
    pool_free(f->ctx, f->tile);
    pool_free(f->ctx, f->tab);
    if (f->own_x) pool_free(f->ctx, f->x);
    device_dim_release(f->ctx);
    delete f;
    return SPKDIAR_OK;
}

int64_t spkdiar_features_frames(const spkdiar_feat* f) { return f ? f->n : -1; }

int spkdiar_stats_window(spkdiar_feat* f, int64_t a, int64_t b, double* out819, double* shift39) {
    if (!f || !out819) return SPKDIAR_E_ARG;
    spkdiar_ctx* c = f->ctx;
    if (a < 0 || b < a || b > f->n) return set_err(c, SPKDIAR_E_ARG, "window [%lld,%lld) outside 0..%lld",
                                                  (long long)a, (long long)b, (long long)f->n);
    SPK_CUDA(c, cudaSetDevice(c->device));
    if (int rc = ensure_stats(f)) return rc;
    std::vector<double> ra(REC), rb(REC), ca(2 * REC), cb(2 * REC);
    SPK_CUDA(c, cudaMemcpyAsync(ra.data(), f->P + a * REC, REC * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SPK_CUDA(c, cudaMemcpyAsync(rb.data(), f->P + b * REC, REC * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SPK_CUDA(c, cudaMemcpyAsync(ca.data(), f->C + (a / K1_TILE) * REC, REC * sizeof(double2), cudaMemcpyDeviceToHost, c->stream));
    SPK_CUDA(c, cudaMemcpyAsync(cb.data(), f->C + (b / K1_TILE) * REC, REC * sizeof(double2), cudaMemcpyDeviceToHost, c->stream));
    if (shift39) SPK_CUDA(c, cudaMemcpyAsync(shift39, f->shift, D39 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SPK_CUDA(c, cudaStreamSynchronize(c->stream));
    // same expression as WinSrc::operator() on the device
    auto win = [&](int q) {
        return (rb[q] - ra[q]) + ((cb[2 * q] - ca[2 * q]) + (cb[2 * q + 1] - ca[2 * q + 1]));
    };
    for (int j = 0; j < D39; ++j) out819[j] = win(L39::VEC + j);
    for (int r = 0; r < D39; ++r)
        for (int k = 0; k <= r; ++k) out819[D39 + r * (r + 1) / 2 + k] = win(L39::pos(r, k));
    if ((int64_t)win(L39::CNT) != b - a)
        return set_err(c, SPKDIAR_E_CUDA, "prefix count mismatch: %g vs %lld", win(L39::CNT), (long long)(b - a));
    return SPKDIAR_OK;
}

// ---- K2 ---------------------------------------------------------------------------------

int spkdiar_score_windows(spkdiar_feat* f, const int64_t* a, const int64_t* m, const int64_t* b,
                          int64_t ncand, int metric, double lambda, double* out_d, double* out_terms) {
    if (!f) return SPKDIAR_E_ARG;
    spkdiar_ctx* c = f->ctx;
    if (ncand < 0 || (ncand > 0 && (!a || !m || !b || !out_d))) return set_err(c, SPKDIAR_E_ARG, "null argument");
    if (int rc = check_metric(c, metric)) return rc;
    if (ncand == 0) return SPKDIAR_OK;
    for (int64_t k = 0; k < ncand; ++k)
        if (a[k] < 0 || m[k] < a[k] || b[k] < m[k] || b[k] > f->n)
            return set_err(c, SPKDIAR_E_ARG, "candidate %lld = (%lld,%lld,%lld) outside 0..%lld", (long long)k,
                           (long long)a[k], (long long)m[k], (long long)b[k], (long long)f->n);
    SPK_CUDA(c, cudaSetDevice(c->device));
    if (int rc = ensure_stats(f)) return rc;
    DevBuf<int64_t> idx; DevBuf<double> terms, dout;
    SPK_CUDA(c, idx.alloc(c, 3 * ncand));
    SPK_CUDA(c, terms.alloc(c, 3 * ncand));
    SPK_CUDA(c, dout.alloc(c, ncand));
    int64_t *da = idx.p, *dm = idx.p + ncand, *db = idx.p + 2 * ncand;
    const size_t nb = ncand * sizeof(int64_t);
    SPK_CUDA(c, cudaMemcpyAsync(da, a, nb, cudaMemcpyHostToDevice, c->stream));
    SPK_CUDA(c, cudaMemcpyAsync(dm, m, nb, cudaMemcpyHostToDevice, c->stream));
    SPK_CUDA(c, cudaMemcpyAsync(db, b, nb, cudaMemcpyHostToDevice, c->stream));
    {
        Prof p(c, SPKDIAR_PROF_SCORE);
        if (metric == SPKDIAR_KL2) {
            win_kl2_kernel<<<grid_for(c, ncand, SC_WARPS, 2), SC_THREADS, SC_WARPS * sizeof(Kl2Scratch), c->stream>>>(
                Stats{f->P, f->C}, f->x, da, dm, db, ncand, dout.p, terms.p);
            c->launches += 1;
        } else {
            win_terms_kernel<<<grid_for(c, 3 * ncand, SC_WARPS, 3), SC_THREADS, SC_SMEM, c->stream>>>(
                Stats{f->P, f->C}, da, dm, db, ncand, metric, terms.p);
            win_combine_kernel<<<(unsigned)((ncand + 127) / 128), 128, 0, c->stream>>>(
                da, dm, db, ncand, metric, lambda, terms.p, dout.p);
            c->launches += 2;
        }
    }
    SPK_CUDA(c, cudaGetLastError());
    SPK_CUDA(c, cudaMemcpyAsync(out_d, dout.p, ncand * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (out_terms)
        SPK_CUDA(c, cudaMemcpyAsync(out_terms, terms.p, 3 * ncand * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SPK_CUDA(c, cudaStreamSynchronize(c->stream));
    return SPKDIAR_OK;
}

int spkdiar_score_sets(spkdiar_feat* f, int64_t npairs,
                       const int64_t* off1, const int64_t* a1, const int64_t* b1,
                       const int64_t* off2, const int64_t* a2, const int64_t* b2,
                       int metric, double lambda, double* out_d, double* out_terms) {
    if (!f) return SPKDIAR_E_ARG;
    spkdiar_ctx* c = f->ctx;
    if (npairs < 0 || (npairs > 0 && (!off1 || !a1 || !b1 || !off2 || !a2 || !b2 || !out_d)))
        return set_err(c, SPKDIAR_E_ARG, "null argument");
    if (int rc = check_metric(c, metric)) return rc;
    if (npairs == 0) return SPKDIAR_OK;
    const int64_t nr1 = off1[npairs], nr2 = off2[npairs];
    for (int side = 0; side < 2; ++side) {
        const int64_t* ra = side ? a2 : a1; const int64_t* rb = side ? b2 : b1;
        const int64_t nr = side ? nr2 : nr1;
        for (int64_t r = 0; r < nr; ++r)
            if (ra[r] < 0 || rb[r] < ra[r] || rb[r] > f->n)
                return set_err(c, SPKDIAR_E_ARG, "range %lld of set list %d = [%lld,%lld) outside 0..%lld",
                               (long long)r, side + 1, (long long)ra[r], (long long)rb[r], (long long)f->n);
    }
    SPK_CUDA(c, cudaSetDevice(c->device));
    if (int rc = ensure_stats(f)) return rc;
    DevBuf<int64_t> idx; DevBuf<double> rec, terms, dout;
    const int64_t nidx = 2 * (npairs + 1) + 2 * nr1 + 2 * nr2;
    SPK_CUDA(c, idx.alloc(c, nidx));
    SPK_CUDA(c, rec.alloc(c, 2 * npairs * REC));
    SPK_CUDA(c, terms.alloc(c, 3 * npairs));
    SPK_CUDA(c, dout.alloc(c, npairs));
    int64_t* d_off1 = idx.p; int64_t* d_off2 = d_off1 + npairs + 1;
    int64_t* d_a1 = d_off2 + npairs + 1; int64_t* d_b1 = d_a1 + nr1;
    int64_t* d_a2 = d_b1 + nr1; int64_t* d_b2 = d_a2 + nr2;
    SPK_CUDA(c, cudaMemcpyAsync(d_off1, off1, (npairs + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    SPK_CUDA(c, cudaMemcpyAsync(d_off2, off2, (npairs + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    if (nr1) {
        SPK_CUDA(c, cudaMemcpyAsync(d_a1, a1, nr1 * 8, cudaMemcpyHostToDevice, c->stream));
        SPK_CUDA(c, cudaMemcpyAsync(d_b1, b1, nr1 * 8, cudaMemcpyHostToDevice, c->stream));
    }
    if (nr2) {
        SPK_CUDA(c, cudaMemcpyAsync(d_a2, a2, nr2 * 8, cudaMemcpyHostToDevice, c->stream));
        SPK_CUDA(c, cudaMemcpyAsync(d_b2, b2, nr2 * 8, cudaMemcpyHostToDevice, c->stream));
    }
    double* recX = rec.p; double* recY = rec.p + npairs * REC;
    {
        Prof p(c, SPKDIAR_PROF_SCORE);
        set_records_kernel<<<(unsigned)npairs, 256, 0, c->stream>>>(Stats{f->P, f->C}, d_off1, d_a1, d_b1, npairs, recX);
        set_records_kernel<<<(unsigned)npairs, 256, 0, c->stream>>>(Stats{f->P, f->C}, d_off2, d_a2, d_b2, npairs, recY);
        if (metric == SPKDIAR_KL2) {
            pair_kl2_kernel<<<grid_for(c, npairs, SC_WARPS, 2), SC_THREADS, SC_WARPS * sizeof(Kl2Scratch), c->stream>>>(
                recX, recY, f->x, d_off1, d_a1, d_b1, d_off2, d_a2, d_b2, npairs, dout.p, terms.p);
            c->launches += 3;
        } else {
            pair_terms_kernel<<<grid_for(c, 3 * npairs, SC_WARPS, 3), SC_THREADS, SC_SMEM, c->stream>>>(
                recX, recY, npairs, metric, terms.p);
            pair_combine_kernel<<<(unsigned)((npairs + 127) / 128), 128, 0, c->stream>>>(
                recX, recY, npairs, metric, lambda, terms.p, dout.p);
            c->launches += 4;
        }
    }
    SPK_CUDA(c, cudaGetLastError());
    SPK_CUDA(c, cudaMemcpyAsync(out_d, dout.p, npairs * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (out_terms)
        SPK_CUDA(c, cudaMemcpyAsync(out_terms, terms.p, 3 * npairs * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SPK_CUDA(c, cudaStreamSynchronize(c->stream));
    return SPKDIAR_OK;
}

}  // extern "C"

