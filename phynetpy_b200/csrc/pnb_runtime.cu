// pnb_runtime.cu -- C ABI of libphynet_b200.so, part 1 (see include/phynet_b200.h): error
// reporting, device arena, contexts, substitution-model handles, pattern-compression dispatch and
// loci (alignment blocks resident in HBM with their partial slots).
// Part 2 = pnb_compile.cu (tree -> op program, partials cache), part 3 = pnb_eval.cu (launches).
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <new>

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

int ensure_device(Buffer& b, size_t bytes) {
    if (b.cap >= bytes) return PNB_OK;
    size_t cap = align_up(std::max(bytes, b.cap * 2), 4096);
    if (b.p) PNB_CUDA_TRY(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    PNB_CUDA_TRY(cudaMalloc(&b.p, cap));
    b.cap = cap;
    return PNB_OK;
}

int ensure_pinned(Buffer& b, size_t bytes) {
    if (b.cap >= bytes) return PNB_OK;
    size_t cap = align_up(std::max(bytes, b.cap * 2), 4096);
    if (b.p) PNB_CUDA_TRY(cudaFreeHost(b.p));
    b.p = nullptr;
    b.cap = 0;
    PNB_CUDA_TRY(cudaMallocHost(&b.p, cap));
    b.cap = cap;
    return PNB_OK;
}


}  // namespace pnb

using namespace pnb;

// =====================================================================
// library / context
// =====================================================================
extern "C" int pnb_abi_version(void) { return PNB_ABI_VERSION; }
extern "C" const char* pnb_last_error(void) { return g_err; }

extern "C" int pnb_device_count(int* count_out) {
    PNB_REQUIRE(count_out != nullptr, PNB_ERR_INVALID, "pnb_device_count: count_out is NULL");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        n = 0;
    }
    *count_out = n;
    return PNB_OK;
}

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
    se = static_cast<cudaError_t>(configure_kernels(ctx));
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
    if (const char* e = getenv("PNB_VIRT_TIPS")) ctx->virt_tips = std::max(1, std::min(atoi(e), 64));
    if (const char* e = getenv("PNB_HOLD_REG")) ctx->hold_reg = pnb::HOLD_REGS && atoi(e) != 0;
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
    if (const char* e = getenv("PNB_VIRT_TIPS")) ctx->virt_tips = std::max(1, std::min(atoi(e), 64));
    if (const char* e = getenv("PNB_HOLD_REG")) ctx->hold_reg = pnb::HOLD_REGS && atoi(e) != 0;
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
    for (Buffer* b : {&ctx->d_stage, &ctx->d_P, &ctx->d_tile_sum, &ctx->d_job_done, &ctx->d_out, &ctx->m_dstage, &ctx->m_scratch,
                      &ctx->m_dout})
        if (b->p) cudaFree(b->p);
    for (Buffer* b : {&ctx->m_hstage, &ctx->m_hout})
        if (b->p) cudaFreeHost(b->p);
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
void pnb::locus_free_partials(pnb_locus* loc) {
    locus_detach_batches(loc);  // resident batches hold the old buffers' addresses
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

void pnb::locus_reset_cache(pnb_locus* loc) {
    loc->lazy_batch = nullptr;
    loc->lazy_job = -1;
    loc->committed.clear();
    loc->pending.clear();
    loc->has_pending = false;
    loc->committed_from_tmpl = loc->pending_from_tmpl = false;
    loc->free_slots.resize(loc->n_slots);
    for (int s = 0; s < loc->n_slots; ++s) loc->free_slots[s] = loc->n_slots - 1 - s;
}

int pnb::locus_ensure_slots(pnb_locus* loc, int n_internal) {
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
    // (resident batches exist on device contexts only)
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    locus_detach_batches(loc);
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
    locus_sync(loc);
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
                    "pnb_locus_read_partial: node %d is the root of a small subtree (tip-only, or at most %d tips); those are "
                    "recomputed from the tip codes, not kept in HBM", node, loc->ctx->virt_tips);
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
