// pnb_api.cu -- the C ABI of libphynet_b200.so (see include/phynet_b200.h).
//
// Host-side runtime of the sequence-likelihood path: device arena, per-locus
// partials cache (which (children, lengths) node lives in which HBM slot),
// compilation of a gene tree into the flat op program the pruning kernel walks,
// staging and launches.  The arithmetic lives in pnb_device.cuh.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <new>

#define PNB_DEVICE_KERNELS
#include "pnb_host.hpp"

int pnb_compress_host(const uint8_t* cols, int32_t n, int64_t L, const uint8_t* code_of_byte,
                      int64_t* P_out, uint8_t* codes_out, int64_t ld_codes, int64_t* weights_out,
                      int64_t* first_col_out, int32_t* pattern_of_col_out);
int pnb_compress_device(pnb_ctx* ctx, const uint8_t* cols, int32_t n, int64_t L,
                        const uint8_t* code_of_byte, int64_t* P_out, uint8_t* codes_out,
                        int64_t ld_codes, int64_t* weights_out, int64_t* first_col_out,
                        int32_t* pattern_of_col_out);

namespace pnb {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------- arena
int Arena::alloc(size_t bytes, void** out) {
    bytes = align_up(std::max<size_t>(bytes, 256), 256);
    auto it = free_.find(bytes);
    if (it != free_.end() && !it->second.empty()) {
        *out = it->second.back();
        it->second.pop_back();
        return PNB_OK;
    }
    for (auto& c : chunks_) {
        if (c.cap - c.used >= bytes) {
            *out = c.base + c.used;
            c.used += bytes;
            return PNB_OK;
        }
    }
    size_t chunk = std::max<size_t>(64u << 20, std::min<size_t>(reserved_ / 2, size_t(2) << 30));
    chunk = align_up(std::max(chunk, bytes), 2u << 20);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, chunk);
    if (e != cudaSuccess && chunk > bytes) {  // retry with the exact size
        (void)cudaGetLastError();
        chunk = align_up(bytes, 2u << 20);
        e = cudaMalloc(&p, chunk);
    }
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("device arena: cudaMalloc(%zu bytes) failed: %s (reserved so far %zu bytes)", chunk,
                  cudaGetErrorString(e), reserved_);
        return PNB_ERR_NOMEM;
    }
    reserved_ += chunk;
    chunks_.push_back({static_cast<char*>(p), chunk, bytes});
    *out = p;
    return PNB_OK;
}

void Arena::release(void* p, size_t bytes) {
    if (!p) return;
    bytes = align_up(std::max<size_t>(bytes, 256), 256);
    free_[bytes].push_back(p);
}

void Arena::destroy() {
    for (auto& c : chunks_) cudaFree(c.base);
    chunks_.clear();
    free_.clear();
    reserved_ = 0;
}

static int ensure_device(Buffer& b, size_t bytes) {
    if (b.cap >= bytes) return PNB_OK;
    size_t cap = align_up(std::max(bytes, b.cap * 2), 4096);
    if (b.p) PNB_CUDA_TRY(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    PNB_CUDA_TRY(cudaMalloc(&b.p, cap));
    b.cap = cap;
    return PNB_OK;
}

static int ensure_pinned(Buffer& b, size_t bytes) {
    if (b.cap >= bytes) return PNB_OK;
    size_t cap = align_up(std::max(bytes, b.cap * 2), 4096);
    if (b.p) PNB_CUDA_TRY(cudaFreeHost(b.p));
    b.p = nullptr;
    b.cap = 0;
    PNB_CUDA_TRY(cudaMallocHost(&b.p, cap));
    b.cap = cap;
    return PNB_OK;
}

static inline uint64_t dbits(double x) {
    uint64_t u;
    memcpy(&u, &x, 8);
    return u;
}

}  // namespace pnb

using namespace pnb;

// =====================================================================
// library / context
// =====================================================================
extern "C" int pnb_abi_version(void) { return PNB_ABI_VERSION; }
extern "C" const char* pnb_last_error(void) { return g_err; }

extern "C" int pnb_ctx_create(int device, pnb_ctx** out) {
    PNB_REQUIRE(out != nullptr, PNB_ERR_INVALID, "pnb_ctx_create: out is NULL");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        (void)cudaGetLastError();
        set_error("pnb_ctx_create: no CUDA device available (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return PNB_ERR_CUDA;
    }
    if (device < 0) PNB_CUDA_TRY(cudaGetDevice(&device));
    PNB_REQUIRE(device < n_dev, PNB_ERR_INVALID, "pnb_ctx_create: device %d out of range (%d devices)",
                device, n_dev);
    PNB_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    PNB_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    PNB_REQUIRE(prop.major >= 10, PNB_ERR_UNSUPPORTED,
                "pnb_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                device, prop.major, prop.minor);
    pnb_ctx* ctx = new (std::nothrow) pnb_ctx();
    PNB_REQUIRE(ctx != nullptr, PNB_ERR_NOMEM, "pnb_ctx_create: out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
    cudaError_t se = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) {
        delete ctx;
        set_error("pnb_ctx_create: cudaStreamCreate failed: %s", cudaGetErrorString(se));
        return PNB_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    se = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 8 && se == cudaSuccess; ++i)
        se = cudaEventCreateWithFlags(&ctx->chunk_ready[i], cudaEventDisableTiming);
    if (se == cudaSuccess) se = cudaEventCreateWithFlags(&ctx->prev_done, cudaEventDisableTiming);
    if (se != cudaSuccess) {
        set_error("pnb_ctx_create: stream/event creation failed: %s", cudaGetErrorString(se));
        cudaStreamDestroy(ctx->own_stream);
        delete ctx;
        return PNB_ERR_CUDA;
    }
    se = cudaFuncSetAttribute(pnb_prune_kernel_t<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              ctx->max_smem_optin - 1024);
    if (se == cudaSuccess)
        se = cudaFuncSetAttribute(pnb_prune_kernel_t<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  ctx->max_smem_optin - 1024);
    if (se == cudaSuccess)
        se = cudaFuncSetAttribute(pnb_prune_kernel_t<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  ctx->max_smem_optin - 1024);
    if (se == cudaSuccess)
        se = cudaFuncSetAttribute(pnb_prune_kernel_t<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  ctx->max_smem_optin - 1024);
    if (se != cudaSuccess) {
        cudaStreamDestroy(ctx->own_stream);
        delete ctx;
        set_error("pnb_ctx_create: cannot configure the pruning kernel (%s): was the library built for "
                  "sm_100a?", cudaGetErrorString(se));
        return PNB_ERR_CUDA;
    }
    // host workers for tree compilation: share the box's cores between the ranks of one node
    // Host workers for tree compilation: the node's cores divided between its ranks, minus one per rank
    // for the CUDA driver / NCCL / the caller's own threads (a descheduled worker stalls a whole batch).
    int workers = static_cast<int>(std::thread::hardware_concurrency());
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) {
        const int lw = atoi(e);
        if (lw > 1) workers /= lw;
    }
    workers = std::min(workers - 1, 8);  // leave a core per rank to the driver / NCCL / caller threads
    if (const char* e = getenv("PNB_HOST_THREADS")) workers = atoi(e);
    workers = std::max(1, std::min(workers, 32));
    if (const char* e = getenv("PNB_CHUNK_MIN_JOBS")) ctx->chunk_min_jobs = std::max(2, atoi(e));
    if (const char* e = getenv("PNB_NO_FUSED")) ctx->allow_fused = atoi(e) == 0;
    if (const char* e = getenv("PNB_STORE_STASH")) ctx->store_stash_levels = std::max(0, std::min(atoi(e), 8));
    ctx->scratch.resize(workers);
    if (workers > 1) ctx->pool = new (std::nothrow) WorkerPool(workers);
    *out = ctx;
    return PNB_OK;
}

extern "C" int pnb_ctx_create_host(pnb_ctx** out) {
    PNB_REQUIRE(out != nullptr, PNB_ERR_INVALID, "pnb_ctx_create_host: out is NULL");
    pnb_ctx* ctx = new (std::nothrow) pnb_ctx();
    PNB_REQUIRE(ctx != nullptr, PNB_ERR_NOMEM, "pnb_ctx_create_host: out of host memory");
    ctx->device = -1;  // planning only: no stream, no arena, no kernels
    ctx->scratch.resize(1);
    if (const char* e = getenv("PNB_STORE_STASH")) ctx->store_stash_levels = std::max(0, std::min(atoi(e), 8));
    *out = ctx;
    return PNB_OK;
}

extern "C" int pnb_ctx_destroy(pnb_ctx* ctx) {
    if (!ctx) return PNB_OK;
    if (ctx->device < 0) {
        delete ctx;
        return PNB_OK;
    }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    delete ctx->pool;
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    for (auto& e : ctx->chunk_ready) if (e) cudaEventDestroy(e);
    if (ctx->prev_done) cudaEventDestroy(ctx->prev_done);
    for (auto& b : ctx->h_stage2) if (b.p) cudaFreeHost(b.p);
    for (auto& e : ctx->stage_ev) if (e) cudaEventDestroy(e);
    if (ctx->h_out.p) cudaFreeHost(ctx->h_out.p);
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    for (Buffer* b : {&ctx->d_stage, &ctx->d_P, &ctx->d_tile_sum, &ctx->d_job_done, &ctx->d_out})
        if (b->p) cudaFree(b->p);
    ctx->arena.destroy();
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return PNB_OK;
}

extern "C" int pnb_ctx_set_stream(pnb_ctx* ctx, void* cuda_stream) {
    PNB_REQUIRE(ctx != nullptr && ctx->device >= 0, PNB_ERR_INVALID, "pnb_ctx_set_stream: NULL or host-only context");
    PNB_CUDA_TRY(cudaSetDevice(ctx->device));
    PNB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return PNB_OK;
}

extern "C" int pnb_ctx_synchronize(pnb_ctx* ctx) {
    PNB_REQUIRE(ctx != nullptr && ctx->device >= 0, PNB_ERR_INVALID, "pnb_ctx_synchronize: NULL or host-only context");
    PNB_CUDA_TRY(cudaSetDevice(ctx->device));
    PNB_CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
    PNB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PNB_OK;
}

extern "C" int pnb_ctx_set_profiling(pnb_ctx* ctx, int enable) {
    PNB_REQUIRE(ctx != nullptr && ctx->device >= 0, PNB_ERR_INVALID, "pnb_ctx_set_profiling: NULL or host-only context");
    PNB_CUDA_TRY(cudaSetDevice(ctx->device));
    if (enable && !ctx->ev[0])
        for (auto& e : ctx->ev) PNB_CUDA_TRY(cudaEventCreate(&e));
    ctx->profiling = enable != 0;
    ctx->ev_ka = ctx->ev_prune = false;
    return PNB_OK;
}

extern "C" int pnb_ctx_last_timings(pnb_ctx* ctx, double ms[4]) {
    PNB_REQUIRE(ctx != nullptr && ms != nullptr, PNB_ERR_INVALID, "pnb_ctx_last_timings: NULL argument");
    ms[0] = ctx->t_compile_ms;
    ms[1] = ms[2] = 0.0;
    ms[3] = ctx->t_call_ms;
    if (!ctx->profiling) return PNB_OK;
    PNB_CUDA_TRY(cudaSetDevice(ctx->device));
    float f = 0.f;
    if (ctx->ev_ka) {
        PNB_CUDA_TRY(cudaEventSynchronize(ctx->ev[1]));
        PNB_CUDA_TRY(cudaEventElapsedTime(&f, ctx->ev[0], ctx->ev[1]));
        ms[1] = f;
    }
    if (ctx->ev_prune) {
        PNB_CUDA_TRY(cudaEventSynchronize(ctx->ev[3]));
        PNB_CUDA_TRY(cudaEventElapsedTime(&f, ctx->ev[2], ctx->ev[3]));
        ms[2] = f;
    }
    return PNB_OK;
}

extern "C" int pnb_ctx_last_stats(pnb_ctx* ctx, int64_t stats[8]) {
    PNB_REQUIRE(ctx != nullptr && stats != nullptr, PNB_ERR_INVALID, "pnb_ctx_last_stats: NULL argument");
    memcpy(stats, ctx->stats, sizeof(ctx->stats));
    return PNB_OK;
}

// =====================================================================
// models
// =====================================================================
static int model_new(pnb_ctx* ctx, pnb_model** out) {
    PNB_REQUIRE(ctx != nullptr && out != nullptr, PNB_ERR_INVALID, "pnb_model_create: NULL argument");
    pnb_model* m = new (std::nothrow) pnb_model();
    PNB_REQUIRE(m != nullptr, PNB_ERR_NOMEM, "pnb_model_create: out of host memory");
    memset(&m->mp, 0, sizeof(m->mp));
    m->ctx = ctx;
    m->id = ctx->next_model_id++;
    *out = m;
    return PNB_OK;
}

extern "C" int pnb_model_create_eigen(pnb_ctx* ctx, const double evals[4], const double U[16],
                                      const double Uinv[16], const double pi[4], pnb_model** out) {
    PNB_REQUIRE(evals && U && Uinv && pi, PNB_ERR_INVALID, "pnb_model_create_eigen: NULL array");
    int rc = model_new(ctx, out);
    if (rc) return rc;
    pnb_model* m = *out;
    memcpy(m->mp.evals, evals, 4 * sizeof(double));
    memcpy(m->mp.U, U, 16 * sizeof(double));
    memcpy(m->mp.Uinv, Uinv, 16 * sizeof(double));
    memcpy(m->mp.pi, pi, 4 * sizeof(double));
    m->mp.kind = 0;
    return PNB_OK;
}

extern "C" int pnb_model_create_jc69(pnb_ctx* ctx, pnb_model** out) {
    int rc = model_new(ctx, out);
    if (rc) return rc;
    pnb_model* m = *out;
    for (int i = 0; i < 4; ++i) m->mp.pi[i] = 0.25;
    m->mp.kind = 1;
    return PNB_OK;
}

extern "C" int pnb_model_create_explicit(pnb_ctx* ctx, const double pi[4], pnb_model** out) {
    PNB_REQUIRE(pi != nullptr, PNB_ERR_INVALID, "pnb_model_create_explicit: pi is NULL");
    int rc = model_new(ctx, out);
    if (rc) return rc;
    pnb_model* m = *out;
    memcpy(m->mp.pi, pi, 4 * sizeof(double));
    m->mp.kind = 2;
    return PNB_OK;
}

extern "C" int pnb_model_destroy(pnb_model* m) {
    delete m;
    return PNB_OK;
}

extern "C" int pnb_pmatrix_batch(pnb_ctx* ctx, const pnb_model* m, const double* t, int64_t nb,
                                 double* P_out) {
    PNB_REQUIRE(ctx && m, PNB_ERR_INVALID, "pnb_pmatrix_batch: NULL handle");
    PNB_REQUIRE(ctx->device >= 0, PNB_ERR_STATE, "pnb_pmatrix_batch: host-only planning context has no device");
    PNB_REQUIRE(m->mp.kind != 2, PNB_ERR_INVALID,
                "pnb_pmatrix_batch: explicit models have no device P(t); compute it on the host");
    PNB_REQUIRE(nb >= 0, PNB_ERR_INVALID, "pnb_pmatrix_batch: nb < 0");
    if (nb == 0) return PNB_OK;
    PNB_REQUIRE(t && P_out, PNB_ERR_INVALID, "pnb_pmatrix_batch: NULL array");
    PNB_CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = ensure_device(ctx->d_stage, static_cast<size_t>(nb) * sizeof(double));
    if (rc) return rc;
    rc = ensure_device(ctx->d_P, static_cast<size_t>(nb) * 16 * sizeof(double));
    if (rc) return rc;
    PNB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    PNB_CUDA_TRY(cudaMemcpyAsync(ctx->d_stage.p, t, nb * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const int64_t blocks = (nb + 127) / 128;
    pnb_pmatrix_kernel<<<static_cast<unsigned>(blocks), 128, 0, ctx->stream>>>(
        m->mp, static_cast<const double*>(ctx->d_stage.p), nb, static_cast<double*>(ctx->d_P.p));
    PNB_CUDA_TRY(cudaGetLastError());
    PNB_CUDA_TRY(cudaMemcpyAsync(P_out, ctx->d_P.p, nb * 16 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PNB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PNB_OK;
}

// =====================================================================
// compression
// =====================================================================
extern "C" int pnb_compress(pnb_ctx* ctx, int use_device, const uint8_t* cols, int32_t n, int64_t L,
                            const uint8_t code_of_byte[256], int64_t* P_out, uint8_t* codes_out,
                            int64_t ld_codes, int64_t* weights_out, int64_t* first_col_out,
                            int32_t* pattern_of_col_out) {
    PNB_REQUIRE(n >= 1, PNB_ERR_INVALID, "pnb_compress: empty alignment (n=%d)", n);
    PNB_REQUIRE(L >= 0 && L < (int64_t(1) << 31), PNB_ERR_INVALID, "pnb_compress: L=%lld out of range",
                static_cast<long long>(L));
    PNB_REQUIRE(P_out && code_of_byte, PNB_ERR_INVALID, "pnb_compress: NULL argument");
    if (L == 0) {
        *P_out = 0;
        return PNB_OK;
    }
    PNB_REQUIRE(cols && codes_out && weights_out && first_col_out, PNB_ERR_INVALID,
                "pnb_compress: NULL array");
    PNB_REQUIRE(ld_codes >= 1, PNB_ERR_INVALID, "pnb_compress: ld_codes < 1");
    if (use_device) {
        PNB_REQUIRE(ctx != nullptr, PNB_ERR_INVALID, "pnb_compress: the device path needs a context");
        return pnb_compress_device(ctx, cols, n, L, code_of_byte, P_out, codes_out, ld_codes,
                                   weights_out, first_col_out, pattern_of_col_out);
    }
    return pnb_compress_host(cols, n, L, code_of_byte, P_out, codes_out, ld_codes, weights_out,
                             first_col_out, pattern_of_col_out);
}

// =====================================================================
// locus
// =====================================================================
static void locus_free_partials(pnb_locus* loc) {
    loc->committed_from_tmpl = loc->pending_from_tmpl = false;
    if (loc->d_scale) loc->ctx->arena.release(loc->d_scale, loc->scale_bytes);
    loc->d_scale = nullptr;
    loc->scale_bytes = 0;
    if (loc->d_partials) loc->ctx->arena.release(loc->d_partials, loc->partials_bytes);
    loc->d_partials = nullptr;
    loc->partials_bytes = 0;
    loc->n_slots = 0;
    loc->free_slots.clear();
    loc->committed.clear();
    loc->pending.clear();
    loc->has_pending = false;
}

static void locus_reset_cache(pnb_locus* loc) {
    loc->committed.clear();
    loc->pending.clear();
    loc->has_pending = false;
    loc->committed_from_tmpl = loc->pending_from_tmpl = false;
    loc->free_slots.resize(loc->n_slots);
    for (int s = 0; s < loc->n_slots; ++s) loc->free_slots[s] = loc->n_slots - 1 - s;
}

static int locus_ensure_slots(pnb_locus* loc, int n_internal) {
    const int want = std::max(2 * n_internal, 2);
    if (loc->n_slots >= want) return PNB_OK;
    locus_free_partials(loc);
    if (loc->ctx->device < 0) {  // planning context: slot ids only
        loc->n_slots = want;
        locus_reset_cache(loc);
        return PNB_OK;
    }
    const size_t bytes = static_cast<size_t>(want) * loc->P_pad * 4 * sizeof(double);
    void* p = nullptr;
    int rc = loc->ctx->arena.alloc(bytes, &p);
    if (rc) return rc;
    loc->d_partials = static_cast<double*>(p);
    loc->partials_bytes = bytes;
    loc->n_slots = want;
    // lanes past the last pattern read and write the padding of a slot: make it defined
    if (cudaMemsetAsync(p, 0, bytes, loc->ctx->stream) != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("pnb_eval: cudaMemset of a partials buffer failed");
        return PNB_ERR_CUDA;
    }
    locus_reset_cache(loc);
    return PNB_OK;
}

extern "C" int pnb_locus_create(pnb_ctx* ctx, int32_t n_rows, int64_t P, const uint8_t* codes,
                                int64_t ld_codes, const int64_t* weights, pnb_locus** out) {
    PNB_REQUIRE(ctx && out, PNB_ERR_INVALID, "pnb_locus_create: NULL argument");
    *out = nullptr;
    PNB_REQUIRE(n_rows >= 1, PNB_ERR_INVALID, "pnb_locus_create: n_rows=%d", n_rows);
    PNB_REQUIRE(P >= 0 && P < (int64_t(1) << 30), PNB_ERR_INVALID, "pnb_locus_create: P=%lld out of range",
                static_cast<long long>(P));
    PNB_REQUIRE(P == 0 || (codes && weights && ld_codes >= P), PNB_ERR_INVALID,
                "pnb_locus_create: NULL array or ld_codes < P");
    if (ctx->device >= 0) PNB_CUDA_TRY(cudaSetDevice(ctx->device));
    pnb_locus* loc = new (std::nothrow) pnb_locus();
    PNB_REQUIRE(loc != nullptr, PNB_ERR_NOMEM, "pnb_locus_create: out of host memory");
    loc->ctx = ctx;
    loc->n_rows = n_rows;
    loc->P = P;
    loc->P_pad = static_cast<int64_t>(align_up(static_cast<size_t>(std::max<int64_t>(P, 1)), TILE_PATTERNS));
    if (P > 0 && ctx->device >= 0) {
        loc->codes_bytes = static_cast<size_t>(n_rows) * loc->P_pad;
        loc->weights_bytes = static_cast<size_t>(loc->P_pad) * sizeof(double);
        void* p = nullptr;
        int rc = ctx->arena.alloc(loc->codes_bytes, &p);
        if (rc) { delete loc; return rc; }
        loc->d_codes = static_cast<uint8_t*>(p);
        rc = ctx->arena.alloc(loc->weights_bytes, &p);
        if (rc) { ctx->arena.release(loc->d_codes, loc->codes_bytes); delete loc; return rc; }
        loc->d_weights = static_cast<double*>(p);
        // host-side padded images.  Device tip code: a single state j is stored as j (the kernel
        // reads column j of P directly); anything else as 0x80 | mask.  Padding = 0 (lanes past the last
        // pattern follow the fast path on it; their results are never stored in a sum).
        std::vector<uint8_t> hc(loc->codes_bytes, 0);
        for (int r = 0; r < n_rows; ++r) {
            const uint8_t* src = codes + static_cast<size_t>(r) * ld_codes;
            uint8_t* dst = hc.data() + static_cast<size_t>(r) * loc->P_pad;
            for (int64_t k = 0; k < P; ++k) {
                const uint8_t m = src[k] & 0xF;
                dst[k] = (m == 1) ? 0 : (m == 2) ? 1 : (m == 4) ? 2 : (m == 8) ? 3 : static_cast<uint8_t>(0x80 | m);
            }
        }
        std::vector<double> hw(loc->P_pad, 0.0);
        for (int64_t k = 0; k < P; ++k) hw[k] = static_cast<double>(weights[k]);
        cudaError_t e = cudaMemcpyAsync(loc->d_codes, hc.data(), loc->codes_bytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(loc->d_weights, hw.data(), loc->weights_bytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            set_error("pnb_locus_create: upload failed: %s", cudaGetErrorString(e));
            ctx->arena.release(loc->d_codes, loc->codes_bytes);
            ctx->arena.release(loc->d_weights, loc->weights_bytes);
            delete loc;
            return PNB_ERR_CUDA;
        }
    }
    ctx->n_loci_alive++;
    *out = loc;
    return PNB_OK;
}

extern "C" int pnb_locus_destroy(pnb_locus* loc) {
    if (!loc) return PNB_OK;
    pnb_ctx* ctx = loc->ctx;
    if (ctx->device < 0) {
        ctx->n_loci_alive--;
        delete loc;
        return PNB_OK;
    }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    locus_free_partials(loc);
    for (auto& T : loc->tmpl)
        if (T.d_ops) ctx->arena.release(T.d_ops, static_cast<size_t>(T.d_ops_cap) * sizeof(Op));
    if (loc->d_codes) ctx->arena.release(loc->d_codes, loc->codes_bytes);
    if (loc->d_weights) ctx->arena.release(loc->d_weights, loc->weights_bytes);
    ctx->n_loci_alive--;
    delete loc;
    return PNB_OK;
}

extern "C" int pnb_locus_invalidate(pnb_locus* loc) {
    PNB_REQUIRE(loc != nullptr, PNB_ERR_INVALID, "pnb_locus_invalidate: NULL locus");
    locus_reset_cache(loc);
    return PNB_OK;
}

extern "C" int pnb_locus_read_partial(pnb_locus* loc, int32_t node, double* out) {
    PNB_REQUIRE(loc && out, PNB_ERR_INVALID, "pnb_locus_read_partial: NULL argument");
    const auto& c = loc->committed;
    PNB_REQUIRE(c.valid, PNB_ERR_STATE, "pnb_locus_read_partial: no committed tree");
    PNB_REQUIRE(node >= 0 && node < static_cast<int32_t>(c.node_ref.size()), PNB_ERR_INVALID,
                "pnb_locus_read_partial: node %d out of range", node);
    pnb_ctx* ctx = loc->ctx;
    PNB_CUDA_TRY(cudaSetDevice(ctx->device));
    PNB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const int64_t ref = c.node_ref[node];
    if (ref >= loc->n_rows) {
        const int64_t slot = ref - loc->n_rows;
        PNB_REQUIRE(c.stored[slot], PNB_ERR_STATE,
                    "pnb_locus_read_partial: node %d is a tip-only node; those are recomputed from the tip codes, "
                    "not kept in HBM", node);
        PNB_CUDA_TRY(cudaMemcpy(out, loc->d_partials + slot * loc->P_pad * 4,
                                static_cast<size_t>(loc->P) * 4 * sizeof(double), cudaMemcpyDeviceToHost));
    } else if (ref >= 0) {
        std::vector<uint8_t> row(loc->P);
        PNB_CUDA_TRY(cudaMemcpy(row.data(), loc->d_codes + ref * loc->P_pad, static_cast<size_t>(loc->P),
                                cudaMemcpyDeviceToHost));
        for (int64_t k = 0; k < loc->P; ++k) {
            const unsigned m = (row[k] & 0x80) ? (row[k] & 0xFu) : (1u << row[k]);
            for (int j = 0; j < 4; ++j) out[k * 4 + j] = (m >> j) & 1 ? 1.0 : 0.0;
        }
    } else {
        for (int64_t k = 0; k < loc->P * 4; ++k) out[k] = 1.0;
    }
    return PNB_OK;
}

// =====================================================================
// tree -> op program
// =====================================================================
namespace {

constexpr int64_t REF_ABSENT = -1;   // leaf whose taxon is not in the alignment
constexpr int64_t REF_COMPUTED = -3; // NOSTORE: recomputed node that has no slot


#define CJ_REQUIRE(cond, code, ...)                           \
    do {                                                      \
        if (!(cond)) {                                        \
            snprintf(S.err, sizeof(S.err), __VA_ARGS__);      \
            return (code);                                    \
        }                                                     \
    } while (0)

// Compile one (locus, tree) job into ops / effective branch lengths (and explicit P matrices)
// written at the job's fixed offset; fills loc->pending in store mode.  Thread-safe for
// distinct loci (store mode) and for shared loci in NOSTORE mode (committed cache is read-only).
int compile_job(CompileScratch& S, pnb_locus* loc, const pnb_model* m, int32_t n, const int32_t* parent,
                const double* blen, const int32_t* tip_row, const double* P_edge, double site_rate,
                uint32_t flags, Op* ops, double* t_out, double* Pexp_out, JobOut* jo) {
    const bool nostore = flags & PNB_EVAL_NOSTORE;
    const bool full = flags & PNB_EVAL_FULL;
    auto& cnt = S.child_cnt;
    auto& off = S.child_off;
    auto& kids = S.child;
    auto& post = S.post;
    auto& stk = S.stack;
    auto& need = S.need;
    auto& dirty = S.dirty;
    auto& len = S.len;
    auto& key = S.key;
    auto& ref = S.ref;

    auto& C = loc->committed;
    auto eff_len = [&](int i) -> double {
        double t = blen ? blen[i] : 0.0;  // reference :419-420 (None -> 0) + clamp of p_matrix :241-242
        if (std::isnan(t)) t = 0.0;
        t *= site_rate;
        if (!(t >= 0.0)) t = 0.0;
        return t;
    };
    auto same_topology = [&](const pnb_locus::Template& T) {
        return T.valid && T.n == n && T.has_tip_row == (tip_row != nullptr) &&
               memcmp(T.parent.data(), parent, static_cast<size_t>(n) * sizeof(int32_t)) == 0 &&
               (!tip_row || memcmp(T.tip_row.data(), tip_row, static_cast<size_t>(n) * sizeof(int32_t)) == 0);
    };
    auto refresh_program = [&](const pnb_locus::Template& T) {
        const int n_ops_t = T.jo.n_ops;
        // with a valid device-resident copy of the program only the branch lengths travel
        if (n_ops_t > 0 && !T.d_ops_valid) memcpy(ops, T.ops.data(), static_cast<size_t>(n_ops_t) * sizeof(Op));
        for (int k = 0; k < n_ops_t; ++k) {
            const int c = T.op_child[k];
            t_out[k] = eff_len(c);
            if (Pexp_out) memcpy(Pexp_out + static_cast<size_t>(k) * 16, P_edge + static_cast<size_t>(c) * 16, 128);
        }
    };
    // ---- fastest path: a full, immediately committed re-evaluation of the topology the committed
    // cache already describes (same slots): refresh branch lengths in place, no pending image.
    const bool rescale = (flags & PNB_EVAL_RESCALE) != 0;
    if (!nostore && full && !(flags & PNB_EVAL_PENDING) && loc->committed_from_tmpl && C.valid &&
        C.rescale == rescale && loc->tmpl[0].n_slots == loc->n_slots && same_topology(loc->tmpl[0])) {
        const pnb_locus::Template& T0 = loc->tmpl[0];
        refresh_program(T0);
        const size_t ne = T0.kid_child.size();
        for (size_t e = 0; e < ne; ++e) C.kid_len[e] = dbits(eff_len(T0.kid_child[e]));
        C.model_id = m->id;
        C.logL_valid = false;
        *jo = T0.jo;
        jo->rc = PNB_OK;
        jo->root_clean_cached = false;
        jo->in_place = true;
        jo->tmpl_hit = 0;
        jo->tmpl_new = -1;
        return PNB_OK;
    }
    if (!nostore) {
        // partials of another model, or stored in the other numerical mode, are useless
        if (C.valid && (C.model_id != m->id || C.rescale != rescale)) locus_reset_cache(loc);
        if (full && !(flags & PNB_EVAL_PENDING)) locus_reset_cache(loc);
    }
    const bool can_lookup = C.valid && !full && C.model_id == m->id && C.rescale == rescale;
    // every node will be recomputed and (store mode) slots come off the free list in canonical order
    const bool all_dirty_canonical = nostore ? !can_lookup : !C.valid;
    // ---- fast path: same topology as the last all-dirty compilation of this locus --------------
    pnb_locus::Template& T = loc->tmpl[nostore ? 1 : 0];
    bool read_locked = false;
    if (all_dirty_canonical) {
        int v = loc->tmpl_lock.load(std::memory_order_acquire);
        while (v >= 0 && !loc->tmpl_lock.compare_exchange_weak(v, v + 1, std::memory_order_acquire)) {}
        read_locked = v >= 0;
    }
    struct ReadUnlock {
        std::atomic<int>* l;
        ~ReadUnlock() { if (l) l->fetch_sub(1, std::memory_order_release); }
        void release() { if (l) { l->fetch_sub(1, std::memory_order_release); l = nullptr; } }
    } read_guard{read_locked ? &loc->tmpl_lock : nullptr};
    if (read_locked && (nostore || T.n_slots == loc->n_slots) && same_topology(T)) {
        refresh_program(T);
        if (!nostore) {
            auto& P_ = loc->pending;
            P_ = T.cache;
            const size_t ne = T.kid_child.size();
            P_.kid_len.resize(ne);
            for (size_t e = 0; e < ne; ++e) P_.kid_len[e] = dbits(eff_len(T.kid_child[e]));
            P_.model_id = m->id;
            P_.rescale = rescale;
            P_.valid = true;
            P_.logL_valid = false;
            loc->free_slots.resize(static_cast<size_t>(loc->n_slots - T.slots_used));
            loc->has_pending = true;
            loc->pending_from_tmpl = true;
        }
        *jo = T.jo;
        jo->rc = PNB_OK;
        jo->root_clean_cached = false;
        jo->in_place = false;
        jo->tmpl_hit = nostore ? 1 : 0;
        jo->tmpl_new = -1;
        return PNB_OK;
    }
    read_guard.release();

    cnt.assign(n, 0);
    int root = -1;
    for (int i = 0; i < n; ++i) {
        const int p = parent[i];
        if (p < 0) {
            CJ_REQUIRE(root < 0, PNB_ERR_INVALID, "pnb_eval: tree has more than one root (nodes %d and %d)", root, i);
            root = i;
        } else {
            CJ_REQUIRE(p < n && p != i, PNB_ERR_INVALID, "pnb_eval: parent[%d]=%d out of range", i, p);
            cnt[p]++;
        }
    }
    CJ_REQUIRE(root >= 0, PNB_ERR_INVALID, "pnb_eval: tree has no root");
    off.resize(n + 1);
    off[0] = 0;
    for (int i = 0; i < n; ++i) off[i + 1] = off[i] + cnt[i];
    kids.resize(n);
    {
        auto& fill = S.order;
        fill.assign(off.begin(), off.end() - 1);
        for (int i = 0; i < n; ++i)
            if (parent[i] >= 0) kids[fill[parent[i]]++] = i;
    }
    // post-order (iterative)
    post.clear();
    stk.clear();
    stk.push_back(root);
    stk.push_back(0);
    while (!stk.empty()) {
        const int idx = stk.back();
        const int v = stk[stk.size() - 2];
        if (idx < cnt[v]) {
            stk.back() = idx + 1;
            stk.push_back(kids[off[v] + idx]);
            stk.push_back(0);
        } else {
            post.push_back(v);
            stk.pop_back();
            stk.pop_back();
        }
    }
    CJ_REQUIRE(static_cast<int>(post.size()) == n, PNB_ERR_INVALID,
               "pnb_eval: tree is not connected (reached %zu of %d nodes): cycle or forest", post.size(), n);

    int n_internal = 0;
    for (int i = 0; i < n; ++i) n_internal += cnt[i] > 0;

    if (!nostore) {
        S.rc_alloc = PNB_OK;
        if (loc->n_slots < std::max(2 * n_internal, 2)) return PNB_ERR_STATE + 100;  // caller grows the slots serially
    }
    const int64_t n_rows = loc->n_rows;

    len.resize(n);
    for (int i = 0; i < n; ++i) len[i] = dbits(eff_len(i));

    ref.resize(n);
    dirty.assign(n, 0);
    need.assign(n, 0);
    // Tip-only internal nodes ("cherries") are VIRTUAL: their partial is a product of P columns selected
    // by 1-byte tip codes, cheaper to recompute than to move 32 B/pattern through HBM.  They keep a slot
    // id (identity in the cache keys) but are never written unless they must be held while a sibling is
    // being computed, and never read back: whoever needs them recomputes them.
    auto& virt = S.virt;
    virt.assign(n, 0);
    for (int v = 0; v < n; ++v) {
        if (cnt[v] == 0) continue;
        bool all_tips = true;
        for (int j = 0; j < cnt[v] && all_tips; ++j) all_tips = cnt[kids[off[v] + j]] == 0;
        virt[v] = all_tips;
    }
    auto& P_ = loc->pending;
    if (!nostore) {
        P_.clear();
        P_.kid_begin.assign(loc->n_slots, -1);
        P_.kid_count.assign(loc->n_slots, 0);
        P_.parent_of.assign(n_rows + loc->n_slots, -1);
        P_.stored.assign(loc->n_slots, 0);
        P_.node_ref.assign(n, REF_ABSENT);
        P_.model_id = m->id;
        P_.rescale = rescale;
    }
    int n_dirty = 0;
    bool templ_ok = true;
    S.kid_child.clear();
    S.op_child.clear();
    for (int v : post) {
        const int k = cnt[v];
        if (k == 0) {
            const int r = tip_row ? tip_row[v] : -1;
            CJ_REQUIRE(r < n_rows, PNB_ERR_INVALID, "pnb_eval: tip_row[%d]=%d but the locus has %lld rows", v, r,
                       static_cast<long long>(n_rows));
            ref[v] = r >= 0 ? r : REF_ABSENT;
            if (!nostore) P_.node_ref[v] = ref[v];
            continue;
        }
        key.clear();
        bool any_dirty = false;
        for (int j = 0; j < k; ++j) {
            const int c = kids[off[v] + j];
            key.push_back({ref[c], len[c], c});
            any_dirty |= dirty[c] != 0;
        }
        auto key_less = [](const CompileScratch::Key& a, const CompileScratch::Key& b) {
            return a.ref != b.ref ? a.ref < b.ref : a.len < b.len;
        };
        if (k == 2) {
            if (key_less(key[1], key[0])) std::swap(key[0], key[1]);
        } else if (k > 2) {
            std::sort(key.begin(), key.end(), key_less);
        }
        for (int j = 0; j + 1 < k; ++j) templ_ok &= key[j].ref != key[j + 1].ref;  // order must not depend on lengths
        int hit = -1;
        if (can_lookup && !any_dirty) {
            for (int j = 0; j < k; ++j) {
                if (key[j].ref < 0) continue;
                const int q = C.parent_of[key[j].ref];
                if (q >= 0 && C.kid_count[q] == k) {
                    const int b = C.kid_begin[q];
                    bool same = true;
                    for (int x = 0; x < k && same; ++x)
                        same = C.kid_ref[b + x] == key[x].ref && C.kid_len[b + x] == key[x].len;
                    if (same) hit = q;
                }
                break;  // one probe is enough: every child has exactly one committed parent
            }
        }
        int slot = hit;
        if (hit < 0) {
            dirty[v] = 1;
            n_dirty++;
            if (!nostore) {
                CJ_REQUIRE(!loc->free_slots.empty(), PNB_ERR_STATE, "pnb_eval: internal error: out of partial slots");
                slot = loc->free_slots.back();
                loc->free_slots.pop_back();
            }
        }
        if (nostore) {
            ref[v] = hit >= 0 ? n_rows + hit : REF_COMPUTED;
        } else {
            ref[v] = n_rows + slot;
            P_.node_ref[v] = ref[v];
            P_.stored[slot] = hit >= 0 ? C.stored[hit] : 0;  // set to 1 when this evaluation writes it
            P_.kid_begin[slot] = static_cast<int32_t>(P_.kid_ref.size());
            P_.kid_count[slot] = k;
            for (auto& kv : key) {
                P_.kid_ref.push_back(kv.ref);
                P_.kid_len.push_back(kv.len);
                S.kid_child.push_back(kv.child);
                if (kv.ref >= 0) P_.parent_of[kv.ref] = slot;
            }
        }
        // stash need: recomputed internal kids are visited in order of decreasing need; the last one is
        // forwarded in registers, every earlier one is stashed while its later siblings are computed
        if (dirty[v]) {
            auto& tmp = S.order;
            tmp.clear();
            for (int j = 0; j < k; ++j) {
                const int c = kids[off[v] + j];
                if (cnt[c] > 0 && (dirty[c] || virt[c])) tmp.push_back(need[c]);
            }
            const int kd = static_cast<int>(tmp.size());
            if (kd > 1) std::sort(tmp.begin(), tmp.end(), std::greater<int>());
            int mx = 0;
            for (int i = 0; i < kd; ++i) mx = std::max(mx, i + tmp[i]);
            need[v] = kd == 0 ? 0 : std::max(mx, kd - 1);
        }
    }
    if (!nostore) {
        P_.root_ref = ref[root];
        P_.valid = true;
    }

    // ---- emit ops: DFS over the nodes that are computed in this evaluation ----
    // R(X) = children of X computed now = recomputed stored nodes (largest stash need first) followed by
    // virtual (tip-only) children, clean or not.  The LAST member of R(X) is forwarded in registers; the
    // earlier ones are held in their HBM slot (store mode: written now, re-read from L2) or in the
    // shared-memory stash (NOSTORE).
    int n_ops = 0, n_tip_ops = 0;
    const bool compute_root = cnt[root] > 0 && (dirty[root] || virt[root]);
    const int stash_levels = nostore ? (1 << 30) : loc->ctx->store_stash_levels;
    if (compute_root) {
        auto& fs = S.frames;
        auto& dk = S.dk;
        auto& dk_base = S.dk_base;
        fs.clear();
        dk.clear();
        dk_base.clear();
        // frame.push_idx: >= 0 stash index (NOSTORE); -1 nothing; -2 must be written to its slot
        auto open = [&](int v, int sp, int push_idx) {
            const int base = static_cast<int>(dk.size());
            for (int j = 0; j < cnt[v]; ++j) {
                const int c = kids[off[v] + j];
                if (cnt[c] > 0 && dirty[c] && !virt[c]) dk.push_back(c);
            }
            if (dk.size() - base > 1)
                std::stable_sort(dk.begin() + base, dk.end(), [&](int a, int b) { return need[a] > need[b]; });
            for (int j = 0; j < cnt[v]; ++j) {
                const int c = kids[off[v] + j];
                if (cnt[c] > 0 && virt[c]) dk.push_back(c);
            }
            fs.push_back({v, 0, sp, push_idx, static_cast<int>(dk.size()) - base});
            dk_base.push_back(base);
        };
        open(root, 0, -1);
        while (!fs.empty()) {
            const int base = dk_base.back();
            {
                CompileScratch::Frame& f = fs.back();
                if (f.i < f.kd) {
                    const int d = dk[base + f.i];
                    const int i = f.i++;
                    const int sp = f.sp;
                    const bool is_last = i == f.kd - 1;
                    int disp = -1;
                    if (!is_last) disp = (nostore || sp + i < stash_levels) ? sp + i : -2;
                    open(d, sp + i, disp);
                    continue;
                }
            }
            const CompileScratch::Frame f = fs.back();
            const int v = f.v;
            const int first_op = n_ops;
            auto add = [&](uint32_t kind, int src, int aux, int c) {
                Op& o = ops[n_ops];
                o.flags = kind;
                o.src = src;
                o.dst = 0;
                o.aux = aux;
                double tt;
                memcpy(&tt, &len[c], 8);
                t_out[n_ops] = tt;
                if (Pexp_out) memcpy(Pexp_out + static_cast<size_t>(n_ops) * 16, P_edge + static_cast<size_t>(c) * 16, 128);
                S.op_child.push_back(c);
                n_ops++;
            };
            if (f.kd > 0) add(K_REG, 0, 0, dk[base + f.kd - 1]);
            for (int i = 0; i + 1 < f.kd; ++i) {
                const int d = dk[base + i];
                if (nostore || f.sp + i < stash_levels) add(K_STK, f.sp + i, 0, d);
                else add(K_MEM, static_cast<int>(ref[d] - n_rows), 0, d);
            }
            for (int j = 0; j < cnt[v]; ++j) {
                const int c = kids[off[v] + j];
                if (cnt[c] == 0) {
                    if (ref[c] >= 0) add(K_TIP, static_cast<int>(ref[c]), n_tip_ops++, c);
                    else add(K_TIP, -1, 0, c);
                } else if (!dirty[c] && !virt[c]) {
                    add(K_MEM, static_cast<int>(ref[c] - n_rows), 0, c);  // unchanged stored subtree
                }
            }
            ops[first_op].flags |= F_FIRST;
            Op& last = ops[n_ops - 1];
            last.flags |= F_LAST;
            const bool write_slot = !nostore && ((!virt[v] && dirty[v]) || f.push_idx == -2);
            if (write_slot) {
                last.flags |= F_STORE;
                last.dst = static_cast<int>(ref[v] - n_rows);
                loc->pending.stored[last.dst] = 1;
            }
            if (f.push_idx >= 0) {
                CJ_REQUIRE(f.push_idx < 256, PNB_ERR_UNSUPPORTED, "pnb_eval: stash depth %d exceeds 255", f.push_idx);
                last.flags |= F_PUSH | (static_cast<uint32_t>(f.push_idx) << 8);
            }
            dk.resize(base);
            dk_base.pop_back();
            fs.pop_back();
        }
    }
    // long programs are streamed through shared memory in segments of SEG_MAX_OPS ops: the staging row
    // of a tip op is then its index among the tip ops of ITS segment
    int rows_cap_job = n_tip_ops;
    if (n_ops > SEG_MAX_OPS) {
        rows_cap_job = 0;
        for (int s0 = 0; s0 < n_ops; s0 += SEG_MAX_OPS) {
            int c = 0;
            for (int k = s0; k < std::min(n_ops, s0 + SEG_MAX_OPS); ++k)
                if ((ops[k].flags & F_KIND_MASK) == K_TIP && ops[k].src >= 0) ops[k].aux = c++;
            rows_cap_job = std::max(rows_cap_job, c);
        }
    }
    jo->rows_cap = rows_cap_job;
    jo->n_ops = n_ops;
    jo->n_tip_ops = n_tip_ops;
    jo->n_dirty = n_dirty;
    jo->stk_need = compute_root ? std::min(need[root], stash_levels) : 0;
    if (cnt[root] == 0) {
        jo->root_kind = K_TIP;
        jo->root_src = static_cast<int>(ref[root]);
    } else if (compute_root) {
        jo->root_kind = K_REG;
        jo->root_src = 0;
    } else {
        jo->root_kind = K_MEM;
        jo->root_src = static_cast<int>(ref[root] - n_rows);
    }
    jo->root_clean_cached = cnt[root] > 0 && !dirty[root] && C.valid && C.logL_valid && C.root_ref == ref[root];
    if (!nostore) {
        loc->has_pending = true;  // rollback re-derives the free slots from `committed`
        loc->pending_from_tmpl = false;
    }
    // remember this program: the next all-dirty evaluation of the same topology only refreshes lengths
    jo->tmpl_hit = -1;
    jo->tmpl_new = -1;
    jo->in_place = false;
    int expect0 = 0;
    const bool write_locked = all_dirty_canonical &&
                              loc->tmpl_lock.compare_exchange_strong(expect0, -1, std::memory_order_acquire);
    if (write_locked && templ_ok && n_dirty == n_internal) {
        T.valid = true;
        T.n = n;
        T.n_slots = loc->n_slots;
        T.slots_used = nostore ? 0 : n_dirty;
        T.has_tip_row = tip_row != nullptr;
        T.parent.assign(parent, parent + n);
        if (tip_row) T.tip_row.assign(tip_row, tip_row + n); else T.tip_row.clear();
        T.ops.assign(ops, ops + n_ops);
        T.op_child = S.op_child;
        T.d_ops_valid = false;  // refreshed (stream-ordered) by pnb_eval after this batch's copy
        jo->tmpl_hit = -1;
        jo->in_place = false;
        jo->tmpl_new = nostore ? 1 : 0;
        T.jo = *jo;
        T.jo.tmpl_new = -1;
        if (!nostore) {
            T.cache = loc->pending;
            T.kid_child = S.kid_child;
            loc->pending_from_tmpl = true;
        }
    } else if (write_locked) {
        T.valid = false;
        T.d_ops_valid = false;
    }
    if (write_locked) loc->tmpl_lock.store(0, std::memory_order_release);
    return PNB_OK;
}

void locus_commit(pnb_locus* loc) {
    if (!loc->has_pending) return;
    std::swap(loc->committed, loc->pending);
    loc->pending.clear();
    loc->has_pending = false;
    loc->committed_from_tmpl = loc->pending_from_tmpl;
    loc->pending_from_tmpl = false;
    // free every slot the committed tree does not use
    loc->free_slots.clear();
    for (int s = loc->n_slots - 1; s >= 0; --s)
        if (loc->committed.kid_begin[s] < 0) loc->free_slots.push_back(s);
}

void locus_rollback(pnb_locus* loc) {
    if (!loc->has_pending) return;
    loc->pending.clear();
    loc->has_pending = false;
    loc->free_slots.clear();
    if (loc->committed.valid) {
        for (int s = loc->n_slots - 1; s >= 0; --s)
            if (loc->committed.kid_begin[s] < 0) loc->free_slots.push_back(s);
    } else {
        for (int s = loc->n_slots - 1; s >= 0; --s) loc->free_slots.push_back(s);
    }
}

}  // namespace

// =====================================================================
// evaluation
// =====================================================================
extern "C" int pnb_eval(pnb_ctx* ctx, const pnb_model* m, int64_t M, pnb_locus* const* loci,
                        const int64_t* node_off, const int32_t* parent, const double* blen,
                        const int32_t* tip_row, const double* P_edge, double site_rate, uint32_t flags,
                        double* logL_out) {
    PNB_REQUIRE(ctx && m, PNB_ERR_INVALID, "pnb_eval: NULL handle");
    PNB_REQUIRE(ctx->device >= 0, PNB_ERR_STATE, "pnb_eval: host-only planning context has no device (use pnb_plan)");
    PNB_REQUIRE(m->ctx == ctx, PNB_ERR_INVALID, "pnb_eval: model belongs to another context");
    PNB_REQUIRE(M >= 0, PNB_ERR_INVALID, "pnb_eval: M < 0");
    memset(ctx->stats, 0, sizeof(ctx->stats));
    const auto t_call0 = std::chrono::steady_clock::now();
    ctx->ev_ka = ctx->ev_prune = false;
    if (M == 0) return PNB_OK;
    PNB_REQUIRE(loci && node_off && parent && logL_out, PNB_ERR_INVALID, "pnb_eval: NULL array");
    const bool nostore = flags & PNB_EVAL_NOSTORE;
    const bool out_dev = flags & PNB_EVAL_OUT_DEVICE;
    const bool explicit_P = m->mp.kind == 2;
    PNB_REQUIRE(!explicit_P || P_edge, PNB_ERR_INVALID, "pnb_eval: explicit model needs P_edge");
    PNB_REQUIRE(!(nostore && (flags & PNB_EVAL_PENDING)), PNB_ERR_INVALID,
                "pnb_eval: NOSTORE and PENDING are mutually exclusive");
    PNB_REQUIRE(std::isfinite(site_rate), PNB_ERR_INVALID, "pnb_eval: site_rate is not finite");
    PNB_CUDA_TRY(cudaSetDevice(ctx->device));

    // ---- pass 0: validate, size the staging regions, make sure partial slots exist ---------
    // (parallel over jobs for large batches: at 10^4 loci even this bookkeeping is milliseconds)
    const uint32_t mark = ++ctx->batch_counter;
    const int n_workers = ctx->pool ? ctx->pool->size() : 0;
    auto parallel_for = [&](int64_t n, int64_t min_parallel, const std::function<void(int, int64_t, int64_t)>& fn) {
        if (n >= min_parallel && n_workers > 1) {
            const int64_t grain = std::max<int64_t>(64, n / (n_workers * 4));
            std::atomic<int64_t> next{0};
            ctx->pool->run([&](int w) {
                for (;;) {
                    const int64_t a0 = next.fetch_add(grain);
                    if (a0 >= n) break;
                    fn(w, a0, std::min(n, a0 + grain));
                }
            });
        } else {
            fn(0, 0, n);
        }
    };
    const bool rescale_flag = (flags & PNB_EVAL_RESCALE) != 0;
    for (auto& S : ctx->scratch) { S.first_err_job = -1; S.tiles = 0; S.grow.clear(); }
    parallel_for(M, 512, [&](int w, int64_t j0, int64_t j1) {
        CompileScratch& S = ctx->scratch[w];
        auto fail = [&](int64_t j, int code, const char* msg) {
            if (S.first_err_job < 0) { S.first_err_job = j; S.rc_alloc = code; snprintf(S.err, sizeof(S.err), "%s", msg); }
        };
        for (int64_t j = j0; j < j1; ++j) {
            pnb_locus* loc = loci[j];
            if (loc == nullptr) { fail(j, PNB_ERR_INVALID, "pnb_eval: NULL locus"); continue; }
            if (loc->ctx != ctx) { fail(j, PNB_ERR_INVALID, "pnb_eval: locus belongs to another context"); continue; }
            const int64_t nn = node_off[j + 1] - node_off[j];
            if (nn < 1 || nn >= (1 << 24)) { fail(j, PNB_ERR_INVALID, "pnb_eval: job has an invalid node count"); continue; }
            if (!nostore) {
                if (loc->batch_mark.exchange(mark) == mark) {
                    fail(j, PNB_ERR_INVALID, "pnb_eval: locus appears twice in a storing batch; use PNB_EVAL_NOSTORE");
                    continue;
                }
                if (loc->has_pending) {
                    fail(j, PNB_ERR_STATE, "pnb_eval: locus has a pending evaluation; commit or roll back first");
                    continue;
                }
                // a tree with nn nodes has at most nn - 1 internal nodes; growing slots is rare and serial
                if (loc->n_slots < std::max<int64_t>(2 * (nn - 1), 2) || (rescale_flag && !loc->d_scale)) S.grow.push_back(j);
            } else if (rescale_flag && !loc->d_scale && loc->n_slots > 0) {
                S.grow.push_back(j);  // NOSTORE reading committed partials in RESCALE mode needs the exponent array
            }
            S.tiles += (loc->P + TILE_PATTERNS - 1) / TILE_PATTERNS;
        }
    });
    int64_t total_tiles = 0;
    for (auto& S : ctx->scratch) {
        if (S.first_err_job >= 0) {
            set_error("%s (job %lld)", S.err, static_cast<long long>(S.first_err_job));
            return S.rc_alloc;
        }
        total_tiles += S.tiles;
    }
    const bool rescale = (flags & PNB_EVAL_RESCALE) != 0;
    for (auto& S : ctx->scratch) {
        for (int64_t j : S.grow) {
            pnb_locus* loc = loci[j];
            const int64_t nn = node_off[j + 1] - node_off[j];
            const int32_t* par = parent + node_off[j];
            std::vector<uint8_t> has_kid(nn, 0);
            for (int64_t i = 0; i < nn; ++i)
                if (par[i] >= 0 && par[i] < nn) has_kid[par[i]] = 1;
            int n_int = 0;
            for (int64_t i = 0; i < nn; ++i) n_int += has_kid[i];
            if (!nostore && loc->n_slots < std::max(2 * n_int, 2)) {
                int rc0 = locus_ensure_slots(loc, n_int);
                if (rc0) return rc0;
            }
            if (rescale && !loc->d_scale && loc->n_slots > 0) {
                const size_t bytes = static_cast<size_t>(loc->n_slots) * loc->P_pad * sizeof(int32_t);
                void* pp = nullptr;
                int rc0 = ctx->arena.alloc(bytes, &pp);
                if (rc0) return rc0;
                loc->d_scale = static_cast<int32_t*>(pp);
                loc->scale_bytes = bytes;
                PNB_CUDA_TRY(cudaMemsetAsync(pp, 0, bytes, ctx->stream));
            }
        }
    }
    const int64_t total_nodes = node_off[M] - node_off[0];
    PNB_REQUIRE(total_nodes < (int64_t(1) << 31) && total_tiles < (int64_t(1) << 31), PNB_ERR_UNSUPPORTED,
                "pnb_eval: batch too large (%lld nodes, %lld tiles)", static_cast<long long>(total_nodes),
                static_cast<long long>(total_tiles));
    // job j owns op slots [node_off[j] - node_off[0] - j, ...): a tree with nn nodes has at most nn - 1 edges
    const int64_t ops_bound = total_nodes - M + 1;

    const size_t off_jobs = 0;
    const size_t off_tiles = align_up(off_jobs + static_cast<size_t>(M) * sizeof(JobDesc), 256);
    const size_t off_ops = align_up(off_tiles + static_cast<size_t>(total_tiles) * sizeof(int32_t), 256);
    const size_t off_t = align_up(off_ops + static_cast<size_t>(ops_bound) * sizeof(Op), 256);
    const size_t off_pexp = align_up(off_t + static_cast<size_t>(ops_bound) * sizeof(double), 256);
    const size_t stage_bytes = off_pexp + (explicit_P ? static_cast<size_t>(ops_bound) * 128 : 0);
    int rc;
    // two pinned staging buffers alternate: the one used two calls ago must have finished its H2D copy
    const int sb = ctx->stage_idx;
    ctx->stage_idx ^= 1;
    if (ctx->stage_ev_valid[sb]) PNB_CUDA_TRY(cudaEventSynchronize(ctx->stage_ev[sb]));
    if (ctx->h_stage2[sb].cap < stage_bytes) {
        // growing frees the old pinned buffer: no copy may still read it
        PNB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if ((rc = ensure_pinned(ctx->h_stage2[sb], stage_bytes))) return rc;
    }
    if (ctx->d_stage.cap < stage_bytes) {
        PNB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if ((rc = ensure_device(ctx->d_stage, stage_bytes))) return rc;
    }
    if ((rc = ensure_pinned(ctx->h_out, static_cast<size_t>(M) * 2 * sizeof(double)))) return rc;
    char* hs = static_cast<char*>(ctx->h_stage2[sb].p);
    JobDesc* h_jobs = reinterpret_cast<JobDesc*>(hs + off_jobs);
    int32_t* h_tile_job = reinterpret_cast<int32_t*>(hs + off_tiles);
    Op* h_ops = reinterpret_cast<Op*>(hs + off_ops);
    double* h_t = reinterpret_cast<double*>(hs + off_t);
    double* h_pexp = explicit_P ? reinterpret_cast<double*>(hs + off_pexp) : nullptr;
    double* h_out = static_cast<double*>(ctx->h_out.p);

    const bool auto_commit = !nostore && !(flags & PNB_EVAL_PENDING);
    // ---- passes 1-3, chunk by chunk: compile the chunk's trees (in parallel when large), fill
    // its job descriptors, copy and launch it -- the GPU runs chunk c while the host compiles c+1.
    ctx->job_out.resize(M);
    JobOut* jout = ctx->job_out.data();
    auto compile_range = [&](CompileScratch& S, int64_t j0, int64_t j1) {
        for (int64_t j = j0; j < j1; ++j) {
            pnb_locus* loc = loci[j];
            const int64_t b = node_off[j];
            const int32_t nn = static_cast<int32_t>(node_off[j + 1] - b);
            const int64_t o = b - node_off[0] - j;
            JobOut jo;
            jo.rc = compile_job(S, loc, m, nn, parent + b, blen ? blen + b : nullptr, tip_row ? tip_row + b : nullptr,
                                P_edge ? P_edge + b * 16 : nullptr, site_rate, flags, h_ops + o, h_t + o,
                                h_pexp ? h_pexp + o * 16 : nullptr, &jo);
            if (jo.rc != PNB_OK && S.first_err_job < 0) S.first_err_job = j;
            if (jo.rc == PNB_OK && auto_commit && !jo.in_place) {
                // host bookkeeping only (which slot holds which node); independent of the GPU finishing
                if (jo.root_clean_cached) jo.cached_logL = loc->committed.logL;
                locus_commit(loc);
            }
            jout[j] = jo;
        }
    };
    auto fail_rollback = [&]() {
        if (nostore) return;
        // auto-commit batches have already published their new trees: drop those caches entirely
        for (int64_t q = 0; q < M; ++q) {
            if (auto_commit) locus_reset_cache(loci[q]);
            else locus_rollback(loci[q]);
        }
    };
    // ~1,000 jobs per chunk: below that the extra launches / events cost more than the overlap gains
    const int64_t n_chunks = std::max<int64_t>(1, std::min<int64_t>(8, M / ctx->chunk_min_jobs));
    if ((rc = ensure_device(ctx->d_P, static_cast<size_t>(std::max<int64_t>(ops_bound, 1)) * 128))) return rc;
    if ((rc = ensure_device(ctx->d_tile_sum, static_cast<size_t>(std::max<int64_t>(total_tiles, 1)) * sizeof(double)))) return rc;
    if ((rc = ensure_device(ctx->d_out, static_cast<size_t>(M) * sizeof(double)))) return rc;
    if (ctx->d_job_done.cap < static_cast<size_t>(M) * sizeof(unsigned)) {
        PNB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if ((rc = ensure_device(ctx->d_job_done, static_cast<size_t>(M) * sizeof(unsigned)))) return rc;
        PNB_CUDA_TRY(cudaMemsetAsync(ctx->d_job_done.p, 0, ctx->d_job_done.cap, ctx->stream));
    }
    char* ds = static_cast<char*>(ctx->d_stage.p);
    cudaStream_t st = ctx->stream;
    double* d_out = out_dev ? logL_out : static_cast<double*>(ctx->d_out.p);
    // Latency path for small batches (the single-locus MCMC step): ONE kernel reads the job
    // descriptors / program / branch lengths straight from the pinned staging memory (UVA), computes
    // P(t) itself and writes logL into mapped host memory: no H2D copy, no K-a launch, no D2H copy.
    bool fused = false;
    if (ctx->allow_fused && !explicit_P && n_chunks == 1 && total_tiles <= 64 && stage_bytes <= (64u << 10)) {
        void *da = nullptr, *db = nullptr;
        fused = cudaHostGetDevicePointer(&da, hs, 0) == cudaSuccess && da == static_cast<void*>(hs) &&
                cudaHostGetDevicePointer(&db, h_out, 0) == cudaSuccess && db == static_cast<void*>(h_out);
        if (!fused) (void)cudaGetLastError();
    }
    const char* prog_base = fused ? hs : ds;  // where non-resident op programs are read from
    bool stage_read_by_kernel = false;
    // Chunked batches upload (and run K-a for) chunk c+1 on the copy stream while the main stream prunes
    // chunk c.  The copy stream may only start once the previous evaluation's kernels, which read the
    // single-buffered device staging area and P array, are done.
    const bool two_streams = n_chunks > 1;
    cudaStream_t up = two_streams ? ctx->copy_stream : st;
    if (two_streams) {
        PNB_CUDA_TRY(cudaEventRecord(ctx->prev_done, st));  // everything queued on the main stream so far
        PNB_CUDA_TRY(cudaStreamWaitEvent(up, ctx->prev_done, 0));
    }

    int64_t n_jobs = 0, n_tiles = 0, n_ops_used = 0, updates = 0, n_cached = 0, n_dirty_nodes = 0;
    int64_t h2d = 0, d2h = 0, launches = 0;
    double compile_ms = 0.0;
    auto& job_of_batch = ctx->job_of_batch;
    job_of_batch.assign(M, -1);
    bool memset_done = false;
    for (int64_t c = 0; c < n_chunks; ++c) {
        const int64_t j0 = M * c / n_chunks, j1 = M * (c + 1) / n_chunks;
        const auto t_comp0 = std::chrono::steady_clock::now();
        for (auto& S : ctx->scratch) S.first_err_job = -1;
        if ((node_off[j1] - node_off[j0]) >= 4096 && n_workers > 1) {
            const int64_t grain = std::max<int64_t>(16, (j1 - j0) / (n_workers * 4));
            std::atomic<int64_t> next{j0};
            ctx->pool->run([&](int w) {
                CompileScratch& S = ctx->scratch[w];
                for (;;) {
                    const int64_t a0 = next.fetch_add(grain);
                    if (a0 >= j1) break;
                    compile_range(S, a0, std::min(j1, a0 + grain));
                }
            });
        } else {
            compile_range(ctx->scratch[0], j0, j1);
        }
        for (auto& S : ctx->scratch) {
            if (S.first_err_job >= 0) {
                const int code = jout[S.first_err_job].rc;
                set_error("%s (job %lld)", S.err, static_cast<long long>(S.first_err_job));
                // loci compiled so far (all chunks up to this one) hold pending state: undo it
                cudaStreamSynchronize(st);
                fail_rollback();
                return code >= 100 ? PNB_ERR_STATE : code;
            }
        }
        // job descriptors + tile map of this chunk (serial, cheap); numbering is global
        const int64_t jobs_first = n_jobs, tiles_first = n_tiles;
        int64_t ops_lo = -1, ops_hi = 0;    // range of P matrices / branch lengths used by this chunk
        int64_t prog_lo = -1, prog_hi = 0;  // range of op programs that travel with this chunk
        std::vector<int64_t>& save_list = ctx->save_list;
        save_list.clear();
        int ops_cap = 1, rows_cap = 1, stk_cap = 0;
        for (int64_t j = j0; j < j1; ++j) {
            pnb_locus* loc = loci[j];
            const JobOut& jo = jout[j];
            n_dirty_nodes += jo.n_dirty;
            pnb_locus::TreeCache* tc = nostore ? nullptr : (auto_commit ? &loc->committed : &loc->pending);
            if (loc->P == 0) {
                h_out[j] = 0.0;  // np.dot over zero patterns
                if (tc) { tc->logL = 0.0; tc->logL_valid = true; }
                continue;
            }
            if (jo.root_clean_cached && !out_dev) {
                const double v = auto_commit ? jo.cached_logL : loc->committed.logL;
                h_out[j] = v;
                if (tc) { tc->logL = v; tc->logL_valid = true; }
                n_cached++;
                continue;
            }
            const int64_t o = node_off[j] - node_off[0] - j;
            JobDesc& jd = h_jobs[n_jobs];
            jd.codes = loc->d_codes;
            jd.partials = loc->d_partials;
            jd.weights = loc->d_weights;
            jd.slot_stride = loc->P_pad * 4;
            jd.P = static_cast<int32_t>(loc->P);
            jd.ld_codes = static_cast<int32_t>(loc->P_pad);
            jd.op_off = static_cast<int32_t>(o);
            jd.n_ops = jo.n_ops;
            jd.n_tip_ops = jo.n_tip_ops;
            jd.root_kind = jo.root_kind;
            jd.root_src = jo.root_src;
            jd.tile0 = static_cast<int32_t>(n_tiles);
            jd.n_tiles = static_cast<int32_t>((loc->P + TILE_PATTERNS - 1) / TILE_PATTERNS);
            jd.out_index = static_cast<int32_t>(j);
            jd.scale = loc->d_scale;
            jd.pad0 = 0;
            const pnb_locus::Template* Th = jo.tmpl_hit >= 0 ? &loc->tmpl[jo.tmpl_hit] : nullptr;
            const bool resident = Th && Th->d_ops_valid && jo.n_ops > 0;
            jd.ops = resident ? Th->d_ops : reinterpret_cast<const Op*>(prog_base + off_ops) + o;
            if (!resident && jo.n_ops > 0) {
                if (prog_lo < 0) prog_lo = o;
                prog_hi = std::max(prog_hi, o + jo.n_ops);
            }
            if (jo.tmpl_new >= 0 && jo.n_ops > 0) save_list.push_back(j);
            for (int t = 0; t < jd.n_tiles; ++t) h_tile_job[n_tiles + t] = static_cast<int32_t>(n_jobs);
            job_of_batch[j] = n_jobs;
            n_tiles += jd.n_tiles;
            n_ops_used += jo.n_ops;
            if (jo.n_ops > 0) {
                if (ops_lo < 0) ops_lo = o;
                ops_hi = std::max(ops_hi, o + jo.n_ops);
            }
            updates += static_cast<int64_t>(jo.n_dirty) * loc->P;
            ops_cap = std::max(ops_cap, std::min(jo.n_ops, SEG_MAX_OPS));
            rows_cap = std::max(rows_cap, jo.rows_cap);
            stk_cap = std::max(stk_cap, jo.stk_need);
            n_jobs++;
        }
        compile_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_comp0).count();
        const int64_t cj = n_jobs - jobs_first, ct = n_tiles - tiles_first;
        if (cj == 0) continue;
        const size_t dyn_smem = static_cast<size_t>(ops_cap) * (sizeof(Op) + 128) +
                                static_cast<size_t>(rows_cap) * TILE_PATTERNS +
                                static_cast<size_t>(stk_cap) * TILE_PATTERNS * (rescale ? 36 : 32);
        if (dyn_smem > static_cast<size_t>(ctx->max_smem_optin - 1024)) {
            cudaStreamSynchronize(st);
            fail_rollback();
            set_error("pnb_eval: a tree program needs %zu bytes of shared memory (ops=%d, tip rows=%d, stash=%d); "
                      "limit is %d", dyn_smem, ops_cap, rows_cap, stk_cap, ctx->max_smem_optin - 1024);
            return PNB_ERR_UNSUPPORTED;
        }
        const int64_t n_op_range = ops_lo < 0 ? 0 : ops_hi - ops_lo;
        if (fused) {
            // nothing to copy: the kernel reads the staging buffer in place
        } else if (n_chunks == 1 && stage_bytes <= (256u << 10)) {
            PNB_CUDA_TRY(cudaMemcpyAsync(ds, hs, stage_bytes, cudaMemcpyHostToDevice, up));
            h2d += stage_bytes;
        } else {
            PNB_CUDA_TRY(cudaMemcpyAsync(ds + off_jobs + jobs_first * sizeof(JobDesc), hs + off_jobs + jobs_first * sizeof(JobDesc),
                                         cj * sizeof(JobDesc), cudaMemcpyHostToDevice, up));
            PNB_CUDA_TRY(cudaMemcpyAsync(ds + off_tiles + tiles_first * sizeof(int32_t), hs + off_tiles + tiles_first * sizeof(int32_t),
                                         ct * sizeof(int32_t), cudaMemcpyHostToDevice, up));
            h2d += cj * sizeof(JobDesc) + ct * sizeof(int32_t);
            if (prog_lo >= 0) {
                PNB_CUDA_TRY(cudaMemcpyAsync(ds + off_ops + prog_lo * sizeof(Op), hs + off_ops + prog_lo * sizeof(Op),
                                             (prog_hi - prog_lo) * sizeof(Op), cudaMemcpyHostToDevice, up));
                h2d += (prog_hi - prog_lo) * sizeof(Op);
            }
            if (n_op_range > 0) {
                if (!explicit_P) {
                    PNB_CUDA_TRY(cudaMemcpyAsync(ds + off_t + ops_lo * sizeof(double), hs + off_t + ops_lo * sizeof(double),
                                                 n_op_range * sizeof(double), cudaMemcpyHostToDevice, up));
                    h2d += n_op_range * sizeof(double);
                }
            }
        }
        if (n_op_range > 0 && explicit_P) {
            PNB_CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(ctx->d_P.p) + ops_lo * 128, h_pexp + ops_lo * 16, n_op_range * 128,
                                         cudaMemcpyHostToDevice, up));
            h2d += n_op_range * 128;
        }
        // programs compiled from scratch in this chunk get a device-resident copy next to their locus, so
        // that later evaluations of the same topology send branch lengths only (D2D, stream-ordered)
        for (int64_t j : save_list) {
            pnb_locus* loc = loci[j];
            pnb_locus::Template& Tn = loc->tmpl[jout[j].tmpl_new];
            const int nops = jout[j].n_ops;
            const int64_t o_chk = node_off[j] - node_off[0] - j;
            // several jobs of one locus may have rebuilt the template in this batch: only a job whose
            // program IS the template's final program may publish the device copy
            if (!Tn.valid || Tn.jo.n_ops != nops ||
                memcmp(Tn.ops.data(), h_ops + o_chk, static_cast<size_t>(nops) * sizeof(Op)) != 0) continue;
            if (Tn.d_ops_cap < nops) {
                if (Tn.d_ops) ctx->arena.release(Tn.d_ops, static_cast<size_t>(Tn.d_ops_cap) * sizeof(Op));
                void* pp = nullptr;
                const int cap = std::max(nops, 64);
                if (ctx->arena.alloc(static_cast<size_t>(cap) * sizeof(Op), &pp) != PNB_OK) { Tn.d_ops = nullptr; Tn.d_ops_cap = 0; continue; }
                Tn.d_ops = static_cast<Op*>(pp);
                Tn.d_ops_cap = cap;
            }
            const int64_t o = node_off[j] - node_off[0] - j;
            PNB_CUDA_TRY(cudaMemcpyAsync(Tn.d_ops, prog_base + off_ops + o * sizeof(Op), static_cast<size_t>(nops) * sizeof(Op),
                                         fused ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, up));
            if (fused) h2d += static_cast<size_t>(nops) * sizeof(Op);
            Tn.d_ops_valid = true;
        }
        const bool prof = ctx->profiling && n_chunks == 1;
        if (n_op_range > 0 && !explicit_P && !fused) {
            const int64_t blocks = (n_op_range + 127) / 128;
            if (prof) PNB_CUDA_TRY(cudaEventRecord(ctx->ev[0], up));
            pnb_pmatrix_kernel<<<static_cast<unsigned>(blocks), 128, 0, up>>>(
                m->mp, reinterpret_cast<const double*>(ds + off_t) + ops_lo, n_op_range,
                static_cast<double*>(ctx->d_P.p) + ops_lo * 16);
            PNB_CUDA_TRY(cudaGetLastError());
            if (prof) { PNB_CUDA_TRY(cudaEventRecord(ctx->ev[1], up)); ctx->ev_ka = true; }
            launches++;
        }
        if (two_streams) {
            PNB_CUDA_TRY(cudaEventRecord(ctx->chunk_ready[c], up));
            PNB_CUDA_TRY(cudaStreamWaitEvent(st, ctx->chunk_ready[c], 0));
        }
        PruneLaunch L;
        L.jobs = reinterpret_cast<const JobDesc*>(prog_base + off_jobs);
        L.tile_job = reinterpret_cast<const int32_t*>(prog_base + off_tiles);
        L.Pm = static_cast<const double*>(ctx->d_P.p);
        L.t = reinterpret_cast<const double*>(prog_base + off_t);
        if (fused && !out_dev) L.out = h_out + M;  // mapped pinned memory: the result needs no D2H copy
        L.tile_sum = static_cast<double*>(ctx->d_tile_sum.p);
        L.job_done = static_cast<unsigned int*>(ctx->d_job_done.p);
        if (!(fused && !out_dev)) L.out = d_out;
        L.ops_cap = ops_cap;
        L.rows_cap = rows_cap;
        L.stk_cap = stk_cap;
        L.tile_base = static_cast<int32_t>(tiles_first);
        if (out_dev && !memset_done) {
            // jobs without device work (empty loci) still need their slot defined
            PNB_CUDA_TRY(cudaMemsetAsync(logL_out, 0, M * sizeof(double), st));
            memset_done = true;
        }
        if (ctx->profiling && c == 0) PNB_CUDA_TRY(cudaEventRecord(ctx->ev[2], st));
        const unsigned grid = static_cast<unsigned>(ct);
        if (fused) {
            if (rescale) pnb_prune_kernel_t<true, true><<<grid, TPB, dyn_smem, st>>>(L, m->mp);
            else pnb_prune_kernel_t<true, false><<<grid, TPB, dyn_smem, st>>>(L, m->mp);
            stage_read_by_kernel = true;
        } else {
            if (rescale) pnb_prune_kernel_t<false, true><<<grid, TPB, dyn_smem, st>>>(L, m->mp);
            else pnb_prune_kernel_t<false, false><<<grid, TPB, dyn_smem, st>>>(L, m->mp);
        }
        PNB_CUDA_TRY(cudaGetLastError());
        if (ctx->profiling && c == n_chunks - 1) { PNB_CUDA_TRY(cudaEventRecord(ctx->ev[3], st)); ctx->ev_prune = true; }
        launches++;
    }
    ctx->t_compile_ms = compile_ms;
    if (h2d > 0 || stage_read_by_kernel) {  // the pinned buffer may be rewritten once every reader of it is done
        if (!ctx->stage_ev[sb]) PNB_CUDA_TRY(cudaEventCreateWithFlags(&ctx->stage_ev[sb], cudaEventDisableTiming));
        PNB_CUDA_TRY(cudaEventRecord(ctx->stage_ev[sb], stage_read_by_kernel ? st : up));
        ctx->stage_ev_valid[sb] = true;
    }
    if (n_jobs > 0) {
        if (!out_dev) {
            double* got = h_out + M;  // second half of the pinned result buffer
            if (!fused) {
                PNB_CUDA_TRY(cudaMemcpyAsync(got, d_out, M * sizeof(double), cudaMemcpyDeviceToHost, st));
                d2h += M * sizeof(double);
            }
            cudaError_t e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) {
                fail_rollback();
                set_error("pnb_eval: kernel execution failed: %s", cudaGetErrorString(e));
                return PNB_ERR_CUDA;
            }
            for (int64_t j = 0; j < M; ++j)
                if (job_of_batch[j] >= 0) h_out[j] = got[j];
        }
    } else if (out_dev) {
        // everything was empty or cached: define the output
        PNB_CUDA_TRY(cudaMemsetAsync(logL_out, 0, M * sizeof(double), ctx->stream));
    }

    if (!out_dev) {
        for (int64_t j = 0; j < M; ++j) {
            logL_out[j] = h_out[j];
            if (!nostore && job_of_batch[j] >= 0) {
                pnb_locus::TreeCache& tc = auto_commit ? loci[j]->committed : loci[j]->pending;
                tc.logL = h_out[j];
                tc.logL_valid = true;
            }
        }
    }

    ctx->t_call_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call0).count();
    ctx->stats[0] = updates;
    ctx->stats[1] = n_ops_used;
    ctx->stats[2] = n_tiles;
    ctx->stats[3] = launches;
    ctx->stats[4] = h2d;
    ctx->stats[5] = d2h;
    ctx->stats[6] = n_cached;
    ctx->stats[7] = n_dirty_nodes;
    return PNB_OK;
}

extern "C" int pnb_plan(pnb_ctx* ctx, const pnb_model* m, int64_t M, pnb_locus* const* loci,
                        const int64_t* node_off, const int32_t* parent, const double* blen,
                        const int32_t* tip_row, double site_rate, uint32_t flags, int32_t* job_info,
                        int32_t* ops_out, double* t_out) {
    PNB_REQUIRE(ctx && m && m->ctx == ctx, PNB_ERR_INVALID, "pnb_plan: NULL handle or foreign model");
    PNB_REQUIRE(ctx->device < 0, PNB_ERR_STATE, "pnb_plan: needs a host-only context (pnb_ctx_create_host)");
    PNB_REQUIRE(M >= 0 && (M == 0 || (loci && node_off && parent && job_info && ops_out && t_out)), PNB_ERR_INVALID,
                "pnb_plan: NULL array");
    PNB_REQUIRE(!(flags & PNB_EVAL_OUT_DEVICE), PNB_ERR_INVALID, "pnb_plan: OUT_DEVICE makes no sense without a device");
    const bool nostore = flags & PNB_EVAL_NOSTORE;
    PNB_REQUIRE(!(nostore && (flags & PNB_EVAL_PENDING)), PNB_ERR_INVALID, "pnb_plan: NOSTORE and PENDING are mutually exclusive");
    const bool auto_commit = !nostore && !(flags & PNB_EVAL_PENDING);
    const uint32_t mark = ++ctx->batch_counter;
    CompileScratch& S = ctx->scratch[0];
    static_assert(sizeof(Op) == 4 * sizeof(int32_t), "Op layout");
    for (int64_t j = 0; j < M; ++j) {
        pnb_locus* loc = loci[j];
        PNB_REQUIRE(loc && loc->ctx == ctx, PNB_ERR_INVALID, "pnb_plan: bad locus (job %lld)", static_cast<long long>(j));
        const int64_t b = node_off[j];
        const int64_t nn = node_off[j + 1] - b;
        PNB_REQUIRE(nn >= 1 && nn < (1 << 24), PNB_ERR_INVALID, "pnb_plan: job %lld has an invalid node count", static_cast<long long>(j));
        if (!nostore) {
            PNB_REQUIRE(loc->batch_mark.exchange(mark) != mark, PNB_ERR_INVALID,
                        "pnb_eval: locus appears twice in a storing batch; use PNB_EVAL_NOSTORE (job %lld)", static_cast<long long>(j));
            PNB_REQUIRE(!loc->has_pending, PNB_ERR_STATE,
                        "pnb_eval: locus has a pending evaluation; commit or roll back first (job %lld)", static_cast<long long>(j));
            int n_int = 0;
            std::vector<uint8_t> has_kid(nn, 0);
            for (int64_t i = 0; i < nn; ++i) {
                const int32_t pa = parent[b + i];
                if (pa >= 0 && pa < nn) has_kid[pa] = 1;
            }
            for (int64_t i = 0; i < nn; ++i) n_int += has_kid[i];
            int rc0 = locus_ensure_slots(loc, n_int);
            if (rc0) return rc0;
        }
        const int64_t o = b - node_off[0] - j;
        JobOut jo;
        S.err[0] = 0;
        jo.rc = compile_job(S, loc, m, static_cast<int32_t>(nn), parent + b, blen ? blen + b : nullptr,
                            tip_row ? tip_row + b : nullptr, nullptr, site_rate, flags,
                            reinterpret_cast<Op*>(ops_out) + o, t_out + o, nullptr, &jo);
        if (jo.rc != PNB_OK) {
            set_error("%s (job %lld)", S.err, static_cast<long long>(j));
            for (int64_t q = 0; q <= j; ++q) {
                if (nostore) break;
                if (auto_commit) locus_reset_cache(loci[q]); else locus_rollback(loci[q]);
            }
            return jo.rc >= 100 ? PNB_ERR_STATE : jo.rc;
        }
        // a template hit with a valid resident copy does not rewrite the program: hand it out anyway
        if (jo.tmpl_hit >= 0 && loc->tmpl[jo.tmpl_hit].d_ops_valid && jo.n_ops > 0)
            memcpy(reinterpret_cast<Op*>(ops_out) + o, loc->tmpl[jo.tmpl_hit].ops.data(), static_cast<size_t>(jo.n_ops) * sizeof(Op));
        const bool cached = jo.root_clean_cached;
        if (auto_commit && !jo.in_place) locus_commit(loc);
        int32_t* info = job_info + j * 8;
        info[0] = jo.n_ops; info[1] = jo.n_tip_ops; info[2] = jo.root_kind; info[3] = jo.root_src;
        info[4] = jo.stk_need; info[5] = jo.n_dirty; info[6] = jo.rows_cap; info[7] = cached ? 1 : 0;
    }
    return PNB_OK;
}

extern "C" int pnb_commit(pnb_ctx* ctx, int64_t M, pnb_locus* const* loci) {
    PNB_REQUIRE(ctx && (M == 0 || loci), PNB_ERR_INVALID, "pnb_commit: NULL argument");
    for (int64_t j = 0; j < M; ++j) {
        PNB_REQUIRE(loci[j] && loci[j]->ctx == ctx, PNB_ERR_INVALID, "pnb_commit: bad locus %lld", static_cast<long long>(j));
        locus_commit(loci[j]);
    }
    return PNB_OK;
}

extern "C" int pnb_rollback(pnb_ctx* ctx, int64_t M, pnb_locus* const* loci) {
    PNB_REQUIRE(ctx && (M == 0 || loci), PNB_ERR_INVALID, "pnb_rollback: NULL argument");
    for (int64_t j = 0; j < M; ++j) {
        PNB_REQUIRE(loci[j] && loci[j]->ctx == ctx, PNB_ERR_INVALID, "pnb_rollback: bad locus %lld", static_cast<long long>(j));
        locus_rollback(loci[j]);
    }
    return PNB_OK;
}
