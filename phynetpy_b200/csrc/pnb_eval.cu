// pnb_eval.cu -- C ABI of libphynet_b200.so, part 3: everything that launches a kernel.
//
// pnb_eval = validate -> compile the batch's trees (worker pool, pnb_compile.cu) into the pinned
// staging area -> upload -> K-a -> pruning kernel(s) -> results; in chunks with two streams for large
// batches, as one fused kernel for small ones.  The kernels themselves are in pnb_device.cuh and are
// compiled into this translation unit only.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>

#define PNB_DEVICE_KERNELS
#include "pnb_host.hpp"
#include <mutex>

using namespace pnb;

int pnb::configure_kernels(pnb_ctx* ctx) {
    const int smem = ctx->max_smem_optin - 1024;
    cudaError_t se = cudaFuncSetAttribute(pnb_prune_kernel_t<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (se == cudaSuccess)
        se = cudaFuncSetAttribute(pnb_prune_kernel_t<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (se == cudaSuccess)
        se = cudaFuncSetAttribute(pnb_prune_kernel_t<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (se == cudaSuccess)
        se = cudaFuncSetAttribute(pnb_prune_kernel_t<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    return static_cast<int>(se);
}

int pnb::launch_pmatrix_gather(const ModelParams& mp, const double* d_blen, const int32_t* d_node_of_op, double site_rate,
                               int64_t n_ops, double* d_P, double* d_keep, cudaStream_t st) {
    if (n_ops <= 0) return PNB_OK;
    const int64_t blocks = (n_ops + 127) / 128;
    static const bool pdl = !(getenv("PNB_NO_PDL") && getenv("PNB_NO_PDL")[0] == '1');
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(blocks));
    cfg.blockDim = dim3(128);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    PNB_CUDA_TRY(cudaLaunchKernelEx(&cfg, pnb_pmatrix_gather_kernel, mp, d_blen, d_node_of_op, site_rate, n_ops, d_P, d_keep));
    PNB_CUDA_TRY(cudaGetLastError());
    return PNB_OK;
}

// How a launch of n_tiles tiles is cut into blocks: one block per tile, or -- PNB_SPLIT_TAIL=1 -- the tiles of the last,
// partial wave as two half-tile blocks each when that wave is at most half full (blocks are dispatched in index order
// onto `slots` = SMs x resident blocks per SM).  Written against the idea that an 8.45-wave launch (cfg5 on 8 GPUs)
// loses half a wave to quantisation; measured (gpurun_out/wave_*.json, s3_n8*.json), kernel time is LINEAR in the
// number of tiles with a fixed ~26 us per launch (ramp-up + drain of the resident blocks), a half-tile block lives
// almost as long as a whole one (its latency is the op walk, not the patterns), and the split changes nothing: it
// stays off by default.  The results are the same bits either way (tests/test_gpu_parity.py::test_split_tail_same_bits).
static void split_tail(unsigned n_tiles, size_t dyn_smem, bool fused, bool rescale, int32_t* n_full, unsigned* grid) {
    static std::mutex mu;    // contexts of several devices / threads may launch at once
    static int slots_cache[4] = {0, 0, 0, 0};
    static size_t smem_cache[4] = {~size_t{0}, ~size_t{0}, ~size_t{0}, ~size_t{0}};
    static int sm_count = 0;
    const char* env = getenv("PNB_SPLIT_TAIL");
    const bool enabled = PPT % 2 == 0 && env && env[0] == '1';
    *n_full = static_cast<int32_t>(n_tiles);
    *grid = n_tiles;
    if (!enabled || n_tiles == 0) return;
    std::lock_guard<std::mutex> lock(mu);
    const int v = (fused ? 2 : 0) + (rescale ? 1 : 0);
    if (smem_cache[v] != dyn_smem) {
        int per_sm = 0;
        cudaError_t e;
        if (v == 0) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pnb_prune_kernel_t<false, false>, TPB, dyn_smem);
        else if (v == 1) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pnb_prune_kernel_t<false, true>, TPB, dyn_smem);
        else if (v == 2) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pnb_prune_kernel_t<true, false>, TPB, dyn_smem);
        else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pnb_prune_kernel_t<true, true>, TPB, dyn_smem);
        if (e != cudaSuccess) { (void)cudaGetLastError(); return; }
        if (sm_count == 0) {
            int dev = 0;
            if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
                (void)cudaGetLastError();
                sm_count = 0;
                return;
            }
        }
        slots_cache[v] = per_sm * sm_count;
        smem_cache[v] = dyn_smem;
    }
    const unsigned slots = static_cast<unsigned>(slots_cache[v]);
    if (slots < 2) return;
    const unsigned rem = n_tiles % slots;
    if (rem == 0 || 2 * rem > slots) return;
    *n_full = static_cast<int32_t>(n_tiles - rem);
    *grid = n_tiles + rem;
}

int pnb::launch_prune(const PruneLaunch& L_in, const ModelParams& mp, unsigned n_tiles, size_t dyn_smem, bool rescale,
                      cudaStream_t st) {
    if (n_tiles == 0) return PNB_OK;
    PruneLaunch L = L_in;
    unsigned grid;
    split_tail(n_tiles, dyn_smem, false, rescale, &L.n_full, &grid);
    static const SmallBatch no_small = {};   // n_jobs == 0: unused outside the fused latency path
    // Programmatic dependent launch: this kernel may become resident while the kernel in front of it on the stream (K-a of
    // the same evaluation) is still running; it stages everything K-a does not write and waits for P(t) with
    // griddepcontrol.wait.  PNB_NO_PDL=1: plain stream order.
    static const bool pdl = !(getenv("PNB_NO_PDL") && getenv("PNB_NO_PDL")[0] == '1');
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(TPB);
    cfg.dynamicSmemBytes = dyn_smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    if (rescale) PNB_CUDA_TRY(cudaLaunchKernelEx(&cfg, pnb_prune_kernel_t<false, true>, L, mp, no_small));
    else PNB_CUDA_TRY(cudaLaunchKernelEx(&cfg, pnb_prune_kernel_t<false, false>, L, mp, no_small));
    PNB_CUDA_TRY(cudaGetLastError());
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
    pnb_pmatrix_kernel<false><<<static_cast<unsigned>(blocks), 128, 0, ctx->stream>>>(
        m->mp, static_cast<const double*>(ctx->d_stage.p), nb, static_cast<double*>(ctx->d_P.p));
    PNB_CUDA_TRY(cudaGetLastError());
    PNB_CUDA_TRY(cudaMemcpyAsync(P_out, ctx->d_P.p, nb * 16 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PNB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PNB_OK;
}

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
    // loci owned by a resident batch learn lazily what that batch last computed: settle them before the (parallel) passes
    for (int64_t j = 0; j < M; ++j)
        if (loci[j] && loci[j]->lazy_batch) locus_sync_slow(loci[j]);
    const uint32_t mark = ++ctx->batch_counter;
    for (auto& pr : ctx->deferred_release) ctx->arena.release(pr.first, pr.second);  // stream order protects reuse
    ctx->deferred_release.clear();
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
    auto& tile_first = ctx->tile_first;
    tile_first.assign(M + 1, 0);
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
            const int64_t jt = (loc->P + TILE_PATTERNS - 1) / TILE_PATTERNS;
            S.tiles += jt;
            tile_first[j + 1] = jt;
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
    for (int64_t j = 0; j < M; ++j) tile_first[j + 1] += tile_first[j];  // prefix: job j owns tiles [tile_first[j], tile_first[j+1])
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
    auto& job_of_batch = ctx->job_of_batch;
    job_of_batch.assign(M, -1);
    const char* desc_prog_base = nullptr;  // set once the launch mode is known (before any compile)
    // descriptor + tile map of one job; runs in the compile worker right after the job's compilation.
    // Jobs are numbered by their batch index and tiles by the prefix sum of pass 0, so nothing here
    // depends on other jobs; a job that needs no device work leaves its tiles marked -1.
    auto fill_desc = [&](CompileScratch::ChunkAcc& A, int64_t j, const JobOut& jo) {
        pnb_locus* loc = loci[j];
        A.n_dirty += jo.n_dirty;
        pnb_locus::TreeCache* tc = nostore ? nullptr : (auto_commit ? &loc->committed : &loc->pending);
        const int64_t t0 = tile_first[j], t1 = tile_first[j + 1];
        if (loc->P == 0) {
            h_out[j] = 0.0;  // np.dot over zero patterns
            if (tc) { tc->logL = 0.0; tc->logL_valid = true; }
            return;
        }
        if (jo.root_clean_cached && !out_dev) {
            const double v = auto_commit ? jo.cached_logL : loc->committed.logL;
            h_out[j] = v;
            if (tc) { tc->logL = v; tc->logL_valid = true; }
            A.n_cached++;
            for (int64_t t = t0; t < t1; ++t) h_tile_job[t] = -1;
            return;
        }
        const int64_t o = node_off[j] - node_off[0] - j;
        JobDesc& jd = h_jobs[j];
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
        jd.tile0 = static_cast<int32_t>(t0);
        jd.n_tiles = static_cast<int32_t>(t1 - t0);
        jd.out_index = static_cast<int32_t>(j);
        jd.scale = loc->d_scale;
        jd.pad0 = 0;
        const bool resident = jo.res_ops != nullptr && jo.n_ops > 0;   // decided under the template lock (compile_job)
        jd.ops = resident ? jo.res_ops : reinterpret_cast<const Op*>(desc_prog_base + off_ops) + o;
        if (!resident && jo.n_ops > 0) {
            if (A.prog_lo < 0 || o < A.prog_lo) A.prog_lo = o;
            A.prog_hi = std::max(A.prog_hi, o + jo.n_ops);
        }
        if (jo.tmpl_new >= 0 && jo.n_ops > 0) A.save.push_back(j);
        for (int64_t t = t0; t < t1; ++t) h_tile_job[t] = static_cast<int32_t>(j);
        job_of_batch[j] = j;
        A.n_jobs++;
        A.n_ops_used += jo.n_ops;
        if (jo.n_ops > 0) {
            if (A.ops_lo < 0 || o < A.ops_lo) A.ops_lo = o;
            A.ops_hi = std::max(A.ops_hi, o + jo.n_ops);
        }
        A.updates += static_cast<int64_t>(jo.n_dirty) * loc->P;
        A.ops_cap = std::max(A.ops_cap, std::min(jo.n_ops, SEG_MAX_OPS));
        A.rows_cap = std::max(A.rows_cap, jo.rows_cap);
        A.stk_cap = std::max(A.stk_cap, jo.stk_need);
    };
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
            if (jo.rc == PNB_OK) fill_desc(S.acc, j, jo);
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
    if ((rc = ensure_device(ctx->d_tile_sum, 2 * static_cast<size_t>(std::max<int64_t>(total_tiles, 1)) * sizeof(double)))) return rc;
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
    desc_prog_base = prog_base;
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
    bool memset_done = false;
    for (int64_t c = 0; c < n_chunks; ++c) {
        const int64_t j0 = M * c / n_chunks, j1 = M * (c + 1) / n_chunks;
        const auto t_comp0 = std::chrono::steady_clock::now();
        for (auto& S : ctx->scratch) { S.first_err_job = -1; S.acc.reset(); }
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
        // reduce the workers' descriptor-pass accumulators
        const int64_t jobs_first = j0, tiles_first = tile_first[j0];
        int64_t ops_lo = -1, ops_hi = 0;    // range of P matrices / branch lengths used by this chunk
        int64_t prog_lo = -1, prog_hi = 0;  // range of op programs that travel with this chunk
        std::vector<int64_t>& save_list = ctx->save_list;
        save_list.clear();
        int ops_cap = 1, rows_cap = 1, stk_cap = 0;
        int64_t chunk_jobs = 0;
        for (auto& S : ctx->scratch) {
            const auto& A = S.acc;
            chunk_jobs += A.n_jobs;
            n_ops_used += A.n_ops_used;
            updates += A.updates;
            n_cached += A.n_cached;
            n_dirty_nodes += A.n_dirty;
            if (A.ops_lo >= 0 && (ops_lo < 0 || A.ops_lo < ops_lo)) ops_lo = A.ops_lo;
            ops_hi = std::max(ops_hi, A.ops_hi);
            if (A.prog_lo >= 0 && (prog_lo < 0 || A.prog_lo < prog_lo)) prog_lo = A.prog_lo;
            prog_hi = std::max(prog_hi, A.prog_hi);
            ops_cap = std::max(ops_cap, A.ops_cap);
            rows_cap = std::max(rows_cap, A.rows_cap);
            stk_cap = std::max(stk_cap, A.stk_cap);
            save_list.insert(save_list.end(), A.save.begin(), A.save.end());
        }
        n_jobs += chunk_jobs;
        n_tiles += tile_first[j1] - tile_first[j0];
        compile_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_comp0).count();
        const int64_t cj = j1 - j0, ct = tile_first[j1] - tile_first[j0];  // descriptors / tiles of the chunk (with holes)
        if (chunk_jobs == 0) continue;
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
            // never rewrite a buffer that a job of THIS batch reads in place (an earlier job of the same locus hit the
            // previous template): the rebuilt program gets a fresh buffer, the old one is recycled at the next call
            const bool in_use = Tn.d_ops && Tn.d_ops_used_mark.load(std::memory_order_relaxed) == mark;
            if (Tn.d_ops_cap < nops || in_use) {
                if (Tn.d_ops) {
                    if (in_use) ctx->deferred_release.emplace_back(Tn.d_ops, static_cast<size_t>(Tn.d_ops_cap) * sizeof(Op));
                    else ctx->arena.release(Tn.d_ops, static_cast<size_t>(Tn.d_ops_cap) * sizeof(Op));
                }
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
            pnb_pmatrix_kernel<true><<<static_cast<unsigned>(blocks), 128, 0, up>>>(
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
        L.n_peers = 0;
        L.bar_epoch = 0;
        L.out_len = 0;
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
        unsigned grid;
        split_tail(static_cast<unsigned>(ct), dyn_smem, fused, rescale, &L.n_full, &grid);
        static const SmallBatch no_small = {};
        if (fused) {
            // one or two loci: descriptors, program and lengths go into the kernel parameters (no PCIe reads for them)
            SmallBatch sm = {};
            const int64_t n_ops_all = total_nodes - M;
            if (M <= SMALL_MAX_JOBS && total_tiles <= SMALL_MAX_TILES && n_ops_all <= SMALL_MAX_OPS) {
                sm.n_jobs = static_cast<int32_t>(M);
                const int32_t* tj = reinterpret_cast<const int32_t*>(hs + off_tiles);
                for (int64_t t = 0; t < total_tiles; ++t) sm.tile_job[t] = tj[t];
                const JobDesc* hj = reinterpret_cast<const JobDesc*>(hs + off_jobs);
                const Op* staged = reinterpret_cast<const Op*>(hs + off_ops);
                for (int64_t j = 0; j < M; ++j) {
                    sm.jobs[j] = hj[j];
                    if (sm.jobs[j].ops >= staged && sm.jobs[j].ops < staged + n_ops_all) sm.jobs[j].ops = nullptr;   // staged program -> parameters
                }
                memcpy(sm.ops, staged, static_cast<size_t>(n_ops_all) * sizeof(Op));
                memcpy(sm.t, hs + off_t, static_cast<size_t>(n_ops_all) * sizeof(double));
            }
            if (rescale) pnb_prune_kernel_t<true, true><<<grid, TPB, dyn_smem, st>>>(L, m->mp, sm);
            else pnb_prune_kernel_t<true, false><<<grid, TPB, dyn_smem, st>>>(L, m->mp, sm);
            stage_read_by_kernel = true;
        } else {
            if (rescale) pnb_prune_kernel_t<false, true><<<grid, TPB, dyn_smem, st>>>(L, m->mp, no_small);
            else pnb_prune_kernel_t<false, false><<<grid, TPB, dyn_smem, st>>>(L, m->mp, no_small);
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
