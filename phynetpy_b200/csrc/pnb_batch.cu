// pnb_batch.cu -- C ABI of libphynet_b200.so, part 4: resident batches.
//
// A resident batch is a set of (locus, gene-tree topology) jobs whose compiled op programs, job
// descriptors and tile map stay in HBM between evaluations: a new evaluation of the same topologies
// with other branch lengths (or another substitution model) sends the lengths only -- 8 bytes per node,
// in the caller's node order, host or device -- and launches two kernels.  Nothing is compiled, no
// per-locus host state is touched (the loci learn lazily which tree their partials now describe:
// locus_sync).  This is the steady state of a sharded full evaluation (BASELINE cfg5): the reference
// loops `calc.log_likelihood(gene_trees[i], model)` over all loci for it (src/_mcmc_seq.py:746-750).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <new>

#include "pnb_host.hpp"

using namespace pnb;

struct pnb_batch {
    pnb_ctx* ctx = nullptr;
    uint32_t flags = 0;              // PNB_EVAL_NOSTORE / PNB_EVAL_RESCALE given at creation
    bool dead = false;               // a locus was destroyed or re-allocated underneath: re-create the batch
    int64_t M = 0, total_nodes = 0, total_ops = 0, total_tiles = 0, updates = 0, n_dirty = 0;
    std::vector<pnb_locus*> loci;
    std::vector<int64_t> node_off;
    // per job: image of the committed-tree description its program produces (store mode), and for every
    // (slot, child) entry of that image the job-local node whose branch length it carries
    std::vector<pnb_locus::TreeCache> image;
    std::vector<std::vector<int32_t>> kid_child;
    std::vector<int32_t> slots_used;
    // device-resident program
    Buffer d_jobs, d_tile_job, d_ops, d_node_of_op, d_blen, d_P, d_tile_sum, d_job_done, d_out;
    Buffer h_blen, h_out;            // pinned staging
    int ops_cap = 1, rows_cap = 1, stk_cap = 0;
    size_t dyn_smem = 0;
    int64_t traffic[4] = {0, 0, 0, 0};  // bytes per evaluation, counted from the programs (pnb_batch_traffic)
    int n_peers = 0;                    // fused collective: peers' output vectors (pnb_batch_set_peer_outputs)
    double* peer_out[MAX_PEERS] = {nullptr};
    // in-kernel barrier of the fused collective (pnb_batch_set_peer_barrier): armed for ONE evaluation at a time
    unsigned long long bar_epoch = 0;
    unsigned long long* bar_peer[MAX_PEERS] = {nullptr};
    const unsigned long long* bar_mine = nullptr;
    int32_t bar_rank[MAX_PEERS] = {0};
    int64_t jobs_with_tiles = 0;     // jobs the kernel finishes (a locus without patterns has no tile)
    Buffer d_launch_done;            // one counter
    int32_t* h_bar_timeout = nullptr;   // pinned + mapped
    // chunks of jobs for the pipelined host-lengths path: [c] = first job / node / op / tile of chunk c
    std::vector<int64_t> ck_job, ck_node, ck_op, ck_tile;
    // state of the last evaluation (what locus_sync materialises)
    uint64_t epoch = 0;
    uint64_t model_id = 0;
    double site_rate = 1.0;
    std::vector<double> last_blen;   // node order
    bool last_blen_valid = false;    // false after a device-lengths evaluation (fetched on demand)
    bool last_out_valid = false;
};

namespace {

void batch_free(pnb_batch* b) {
    for (Buffer* q : {&b->d_jobs, &b->d_tile_job, &b->d_ops, &b->d_node_of_op, &b->d_blen, &b->d_P, &b->d_tile_sum,
                      &b->d_job_done, &b->d_out, &b->d_launch_done})
        if (q->p) cudaFree(q->p);
    if (b->h_blen.p) cudaFreeHost(b->h_blen.p);
    if (b->h_out.p) cudaFreeHost(b->h_out.p);
    if (b->h_bar_timeout) cudaFreeHost(b->h_bar_timeout);

    delete b;
}

inline double eff_len_of(double t, double site_rate) {
    if (std::isnan(t)) t = 0.0;  // reference :419-420 (None -> 0)
    t *= site_rate;
    if (!(t >= 0.0)) t = 0.0;    // clamp of p_matrix :241-242
    return t;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// lazy per-locus state
// ---------------------------------------------------------------------------------------------
void pnb::locus_sync_slow(pnb_locus* loc) {
    pnb_batch* b = loc->lazy_batch;
    const int64_t j = loc->lazy_job;
    loc->lazy_batch = nullptr;
    loc->lazy_job = -1;
    if (!b || b->dead || j < 0 || j >= b->M || b->loci[j] != loc) return;
    if (!b->last_blen_valid) {
        // lengths of the last evaluation live on the device only: fetch them once
        cudaSetDevice(b->ctx->device);
        cudaStreamSynchronize(b->ctx->stream);
        b->last_blen.resize(static_cast<size_t>(b->total_nodes));
        if (cudaMemcpy(b->last_blen.data(), b->d_blen.p, static_cast<size_t>(b->total_nodes) * sizeof(double),
                       cudaMemcpyDeviceToHost) != cudaSuccess) {
            (void)cudaGetLastError();
            locus_reset_cache(loc);
            return;
        }
        b->last_blen_valid = true;
    }
    pnb_locus::TreeCache& C = loc->committed;
    C = b->image[j];
    const double* bl = b->last_blen.data() + b->node_off[j];
    const std::vector<int32_t>& kc = b->kid_child[j];
    C.kid_len.resize(kc.size());
    for (size_t e = 0; e < kc.size(); ++e) C.kid_len[e] = dbits(eff_len_of(bl[kc[e]], b->site_rate));
    C.model_id = b->model_id;
    C.valid = true;
    C.logL_valid = false;
    if (b->last_out_valid && b->h_out.p) {
        C.logL = static_cast<const double*>(b->h_out.p)[j];
        C.logL_valid = true;
    }
    loc->pending.clear();
    loc->has_pending = false;
    loc->committed_from_tmpl = loc->pending_from_tmpl = false;
    loc->free_slots.clear();
    for (int s = loc->n_slots - 1; s >= 0; --s)
        if (C.kid_begin[s] < 0) loc->free_slots.push_back(s);
}

// A locus of the batch is going away (destroyed) or its buffers are being re-allocated: the batch's resident
// descriptors hold stale addresses from now on.  The other loci keep what the batch last computed for them.
static void batch_kill(pnb_batch* b, pnb_locus* culprit) {
    if (b->dead) return;
    for (int64_t j = 0; j < b->M; ++j) {
        pnb_locus* q = b->loci[j];
        if (!q) continue;
        if (q->lazy_batch == b) {
            if (q != culprit) locus_sync_slow(q);
            else { q->lazy_batch = nullptr; q->lazy_job = -1; }
        }
        auto& v = q->batches;
        v.erase(std::remove(v.begin(), v.end(), b), v.end());
        b->loci[j] = nullptr;
    }
    b->dead = true;
}

void pnb::locus_detach_batches(pnb_locus* loc) {
    const std::vector<pnb_batch*> owners = loc->batches;  // batch_kill edits loc->batches
    for (pnb_batch* b : owners) batch_kill(b, loc);
    loc->batches.clear();
    loc->lazy_batch = nullptr;
    loc->lazy_job = -1;
}

// ---------------------------------------------------------------------------------------------
extern "C" int pnb_batch_create(pnb_ctx* ctx, const pnb_model* m, int64_t M, pnb_locus* const* loci,
                                const int64_t* node_off, const int32_t* parent, const int32_t* tip_row,
                                uint32_t flags, pnb_batch** out) {
    PNB_REQUIRE(ctx && m && out, PNB_ERR_INVALID, "pnb_batch_create: NULL argument");
    *out = nullptr;
    PNB_REQUIRE(ctx->device >= 0, PNB_ERR_STATE, "pnb_batch_create: host-only planning context has no device");
    PNB_REQUIRE(m->ctx == ctx, PNB_ERR_INVALID, "pnb_batch_create: model belongs to another context");
    PNB_REQUIRE(m->mp.kind != 2, PNB_ERR_UNSUPPORTED,
                "pnb_batch_create: explicit models supply P per edge on the host; use pnb_eval");
    PNB_REQUIRE(M >= 1 && loci && node_off && parent, PNB_ERR_INVALID, "pnb_batch_create: empty batch or NULL array");
    PNB_REQUIRE((flags & ~(PNB_EVAL_NOSTORE | PNB_EVAL_RESCALE)) == 0, PNB_ERR_INVALID,
                "pnb_batch_create: only PNB_EVAL_NOSTORE / PNB_EVAL_RESCALE are creation flags");
    const bool nostore = flags & PNB_EVAL_NOSTORE;
    const bool rescale = flags & PNB_EVAL_RESCALE;
    PNB_CUDA_TRY(cudaSetDevice(ctx->device));
    pnb_batch* b = new (std::nothrow) pnb_batch();
    PNB_REQUIRE(b != nullptr, PNB_ERR_NOMEM, "pnb_batch_create: out of host memory");
    b->ctx = ctx;
    b->flags = flags;
    b->M = M;
    b->loci.assign(loci, loci + M);
    b->node_off.assign(node_off, node_off + M + 1);
    for (int64_t j = M; j >= 0; --j) b->node_off[j] -= b->node_off[0];
    b->total_nodes = b->node_off[M];
    const uint32_t mark = ++ctx->batch_counter;
    auto fail = [&](int code) { batch_free(b); return code; };

    // ---- validate, make sure slots (and exponent arrays) exist --------------------------------
    std::vector<int64_t> tile_first(M + 1, 0);
    for (int64_t j = 0; j < M; ++j) {
        pnb_locus* loc = loci[j];
        if (!loc || loc->ctx != ctx) { set_error("pnb_batch_create: bad locus (job %lld)", (long long)j); return fail(PNB_ERR_INVALID); }
        const int64_t nn = b->node_off[j + 1] - b->node_off[j];
        if (nn < 1 || nn >= (1 << 24)) { set_error("pnb_batch_create: job %lld has an invalid node count", (long long)j); return fail(PNB_ERR_INVALID); }
        if (!nostore) {
            if (loc->batch_mark.exchange(mark) == mark) {
                set_error("pnb_batch_create: locus appears twice in a storing batch; use PNB_EVAL_NOSTORE (job %lld)", (long long)j);
                return fail(PNB_ERR_INVALID);
            }
            locus_sync(loc);
            if (loc->has_pending) {
                set_error("pnb_batch_create: locus has a pending evaluation; commit or roll back first (job %lld)", (long long)j);
                return fail(PNB_ERR_STATE);
            }
            const int32_t* par = parent + node_off[j];
            std::vector<uint8_t> has_kid(nn, 0);
            for (int64_t i = 0; i < nn; ++i)
                if (par[i] >= 0 && par[i] < nn) has_kid[par[i]] = 1;
            int n_int = 0;
            for (int64_t i = 0; i < nn; ++i) n_int += has_kid[i];
            int rc0 = locus_ensure_slots(loc, n_int);
            if (rc0) return fail(rc0);
        }
        if (rescale && !loc->d_scale && loc->n_slots > 0) {
            const size_t bytes = static_cast<size_t>(loc->n_slots) * loc->P_pad * sizeof(int32_t);
            void* pp = nullptr;
            int rc0 = ctx->arena.alloc(bytes, &pp);
            if (rc0) return fail(rc0);
            loc->d_scale = static_cast<int32_t*>(pp);
            loc->scale_bytes = bytes;
            if (cudaMemsetAsync(pp, 0, bytes, ctx->stream) != cudaSuccess) { (void)cudaGetLastError(); return fail(PNB_ERR_CUDA); }
        }
        tile_first[j + 1] = tile_first[j] + (loc->P + TILE_PATTERNS - 1) / TILE_PATTERNS;
    }
    b->total_tiles = tile_first[M];
    if (b->total_nodes >= (int64_t(1) << 31) || b->total_tiles >= (int64_t(1) << 31)) {
        set_error("pnb_batch_create: batch too large (%lld nodes, %lld tiles)", (long long)b->total_nodes, (long long)b->total_tiles);
        return fail(PNB_ERR_UNSUPPORTED);
    }

    // ---- compile every job from a clean slate (no templates: the batch owns its programs) ------
    const int64_t ops_bound = b->total_nodes - M + 1;
    std::vector<Op> h_ops(static_cast<size_t>(std::max<int64_t>(ops_bound, 1)));
    std::vector<double> h_t(h_ops.size());
    std::vector<int32_t> h_node_of_op(h_ops.size(), 0);
    std::vector<JobDesc> h_jobs(static_cast<size_t>(M));
    std::vector<int32_t> h_tile_job(static_cast<size_t>(std::max<int64_t>(b->total_tiles, 1)), -1);
    b->image.resize(nostore ? 0 : M);
    b->kid_child.resize(nostore ? 0 : M);
    b->slots_used.assign(M, 0);
    CompileScratch& S = ctx->scratch[0];
    const uint32_t cflags = PNB_EVAL_FULL | flags | PNB_COMPILE_NO_TEMPLATE;
    int64_t op_cursor = 0;
    for (int64_t j = 0; j < M; ++j) {
        pnb_locus* loc = loci[j];
        const int64_t nb0 = node_off[j];
        const int32_t nn = static_cast<int32_t>(node_off[j + 1] - nb0);
        JobOut jo;
        S.err[0] = 0;
        // compile_job resets the committed cache for a full, immediately committed evaluation and fills
        // loc->pending with the image of what the program computes; lengths are placeholders (zeros)
        jo.rc = compile_job(S, loc, m, nn, parent + nb0, nullptr, tip_row ? tip_row + nb0 : nullptr, nullptr, 1.0, cflags,
                            h_ops.data() + op_cursor, h_t.data() + op_cursor, nullptr, &jo);
        if (jo.rc != PNB_OK) {
            set_error("%s (job %lld)", S.err, (long long)j);
            for (int64_t q = 0; q <= j && !nostore; ++q) locus_reset_cache(loci[q]);
            return fail(jo.rc >= 100 ? PNB_ERR_STATE : jo.rc);
        }
        for (int k = 0; k < jo.n_ops; ++k)
            h_node_of_op[op_cursor + k] = static_cast<int32_t>(b->node_off[j] + S.op_child[k]);
        if (!nostore) {
            b->image[j] = loc->pending;
            b->kid_child[j] = S.kid_child;
            b->slots_used[j] = jo.n_dirty;
            // the locus now belongs to this batch: its committed tree is defined by the batch's evaluations
            locus_reset_cache(loc);
        }
        JobDesc& jd = h_jobs[j];
        memset(&jd, 0, sizeof(jd));
        jd.codes = loc->d_codes;
        jd.partials = loc->d_partials;
        jd.weights = loc->d_weights;
        jd.slot_stride = loc->P_pad * 4;
        jd.P = static_cast<int32_t>(loc->P);
        jd.ld_codes = static_cast<int32_t>(loc->P_pad);
        jd.op_off = static_cast<int32_t>(op_cursor);
        jd.n_ops = jo.n_ops;
        jd.n_tip_ops = jo.n_tip_ops;
        jd.root_kind = jo.root_kind;
        jd.root_src = jo.root_src;
        jd.tile0 = static_cast<int32_t>(tile_first[j]);
        jd.n_tiles = static_cast<int32_t>(tile_first[j + 1] - tile_first[j]);
        jd.out_index = static_cast<int32_t>(j);
        jd.scale = loc->d_scale;
        jd.ops = nullptr;  // patched below once the device array exists
        if (loc->P > 0) {
            for (int64_t t = tile_first[j]; t < tile_first[j + 1]; ++t) h_tile_job[t] = static_cast<int32_t>(j);
            b->jobs_with_tiles++;
        }
        {
            // what the kernel moves for this job, per evaluation: every F_STORE writes one 32-byte partial per (padded)
            // pattern, every K_MEM op reads one (a sibling written earlier in the same launch: normally an L2 hit),
            // every tip op stages one code byte per pattern; each tile also stages the program and its P matrices
            int64_t n_store = 0, n_mem = 0, n_tip = 0;
            for (int k = 0; k < jo.n_ops; ++k) {
                const Op& o = h_ops[op_cursor + k];
                if (o.flags & F_STORE) n_store++;
                if ((o.flags & F_KIND_MASK) == K_MEM) n_mem++;
                if ((o.flags & F_KIND_MASK) == K_TIP && o.src >= 0) n_tip++;
            }
            const int64_t tiles = tile_first[j + 1] - tile_first[j];
            b->traffic[0] += n_store * loc->P_pad * 32;
            b->traffic[1] += n_mem * loc->P_pad * 32;
            b->traffic[2] += n_tip * loc->P_pad;
            b->traffic[3] += loc->P_pad * 8 + tiles * static_cast<int64_t>(jo.n_ops) * (sizeof(Op) + 128);
        }
        b->updates += static_cast<int64_t>(jo.n_dirty) * loc->P;
        b->n_dirty += jo.n_dirty;
        b->ops_cap = std::max(b->ops_cap, std::min(jo.n_ops, SEG_MAX_OPS));
        b->rows_cap = std::max(b->rows_cap, jo.rows_cap);
        b->stk_cap = std::max(b->stk_cap, jo.stk_need);
        op_cursor += jo.n_ops;
    }
    b->total_ops = op_cursor;
    b->dyn_smem = static_cast<size_t>(b->ops_cap) * (sizeof(Op) + 128) + static_cast<size_t>(b->rows_cap) * TILE_PATTERNS +
                  static_cast<size_t>(b->stk_cap) * TILE_PATTERNS * (rescale ? 36 : 32);
    if (b->dyn_smem > static_cast<size_t>(ctx->max_smem_optin - 1024)) {
        set_error("pnb_batch_create: a tree program needs %zu bytes of shared memory; limit is %d", b->dyn_smem,
                  ctx->max_smem_optin - 1024);
        return fail(PNB_ERR_UNSUPPORTED);
    }

    // ---- device-resident copies ---------------------------------------------------------------
    int rc;
    const size_t n_ops_alloc = static_cast<size_t>(std::max<int64_t>(b->total_ops, 1));
    if ((rc = ensure_device(b->d_ops, n_ops_alloc * sizeof(Op))) || (rc = ensure_device(b->d_node_of_op, n_ops_alloc * sizeof(int32_t))) ||
        (rc = ensure_device(b->d_P, n_ops_alloc * 128)) || (rc = ensure_device(b->d_jobs, static_cast<size_t>(M) * sizeof(JobDesc))) ||
        (rc = ensure_device(b->d_tile_job, h_tile_job.size() * sizeof(int32_t))) ||
        (rc = ensure_device(b->d_tile_sum, 2 * h_tile_job.size() * sizeof(double))) ||
        (rc = ensure_device(b->d_job_done, static_cast<size_t>(M) * sizeof(unsigned))) ||
        (rc = ensure_device(b->d_out, static_cast<size_t>(M) * sizeof(double))) ||
        (rc = ensure_device(b->d_blen, static_cast<size_t>(b->total_nodes) * sizeof(double))) ||
        (rc = ensure_pinned(b->h_blen, static_cast<size_t>(b->total_nodes) * sizeof(double))) ||
        (rc = ensure_pinned(b->h_out, static_cast<size_t>(M) * sizeof(double))))
        return fail(rc);
    for (int64_t j = 0; j < M; ++j) h_jobs[j].ops = static_cast<const Op*>(b->d_ops.p) + h_jobs[j].op_off;
    cudaStream_t st = ctx->stream;
    cudaError_t e = cudaMemcpyAsync(b->d_ops.p, h_ops.data(), static_cast<size_t>(b->total_ops) * sizeof(Op), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(b->d_node_of_op.p, h_node_of_op.data(), static_cast<size_t>(b->total_ops) * sizeof(int32_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(b->d_jobs.p, h_jobs.data(), static_cast<size_t>(M) * sizeof(JobDesc), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(b->d_tile_job.p, h_tile_job.data(), h_tile_job.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(b->d_job_done.p, 0, static_cast<size_t>(M) * sizeof(unsigned), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(b->d_out.p, 0, static_cast<size_t>(M) * sizeof(double), st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        set_error("pnb_batch_create: upload failed: %s", cudaGetErrorString(e));
        return fail(PNB_ERR_CUDA);
    }
    memset(b->h_out.p, 0, static_cast<size_t>(M) * sizeof(double));

    // ---- chunks for the pipelined host-lengths path ------------------------------------------------------
    // The upload of chunk c+1 hides behind the pruning of chunk c, so only chunk 0's upload is exposed: chunks grow
    // geometrically (1/16, 3/16, 12/16 of the nodes; 1/4, 3/4 for medium batches) -- few launches (every launch has
    // a partially filled last wave), a small exposed first copy.  Thresholds follow PNB_CHUNK_MIN_JOBS (1024 ->
    // two chunks from 512 jobs / 32768 nodes, three from 4096 jobs / 262144 nodes).
    const int64_t cj = std::max(ctx->chunk_min_jobs / 2, 1), cn = static_cast<int64_t>(ctx->chunk_min_jobs) * 32;
    std::vector<double> cuts;
    if (M >= 8 * cj && b->total_nodes >= 8 * cn) cuts = {1.0 / 16, 4.0 / 16};
    else if (M >= cj && b->total_nodes >= cn) cuts = {1.0 / 4};
    b->ck_job.assign(1, 0);
    for (double f : cuts) {
        const int64_t target = static_cast<int64_t>(b->total_nodes * f);
        int64_t j = std::lower_bound(b->node_off.begin(), b->node_off.end(), target) - b->node_off.begin();
        j = std::min<int64_t>(std::max<int64_t>(j, b->ck_job.back() + 1), M - 1);
        if (j > b->ck_job.back()) b->ck_job.push_back(j);
    }
    b->ck_job.push_back(M);
    for (int64_t j : b->ck_job) {
        b->ck_node.push_back(b->node_off[j]);
        b->ck_op.push_back(j < M ? h_jobs[j].op_off : b->total_ops);
        b->ck_tile.push_back(tile_first[j]);
    }
    for (int64_t j = 0; j < M; ++j)
        if (std::find(loci[j]->batches.begin(), loci[j]->batches.end(), b) == loci[j]->batches.end()) loci[j]->batches.push_back(b);
    *out = b;
    return PNB_OK;
}

extern "C" int pnb_batch_destroy(pnb_batch* b) {
    if (!b) return PNB_OK;
    pnb_ctx* ctx = b->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    if (!b->dead) {
        for (int64_t j = 0; j < b->M; ++j) {
            pnb_locus* loc = b->loci[j];
            if (!loc) continue;
            if (loc->lazy_batch == b) locus_sync_slow(loc);  // the loci keep what the last evaluation computed
            auto& v = loc->batches;
            v.erase(std::remove(v.begin(), v.end(), b), v.end());
        }
    }
    batch_free(b);
    return PNB_OK;
}

extern "C" int pnb_batch_info(const pnb_batch* b, int64_t info[8]) {
    PNB_REQUIRE(b && info, PNB_ERR_INVALID, "pnb_batch_info: NULL argument");
    info[0] = b->M; info[1] = b->total_nodes; info[2] = b->total_ops; info[3] = b->total_tiles;
    info[4] = b->updates; info[5] = b->n_dirty; info[6] = static_cast<int64_t>(b->dyn_smem);
    info[7] = static_cast<int64_t>(b->ck_job.size()) - 1;
    return PNB_OK;
}

extern "C" int pnb_batch_traffic(const pnb_batch* b, int64_t bytes[4]) {
    PNB_REQUIRE(b && bytes, PNB_ERR_INVALID, "pnb_batch_traffic: NULL argument");
    memcpy(bytes, b->traffic, sizeof(b->traffic));
    return PNB_OK;
}

extern "C" int pnb_batch_set_peer_outputs(pnb_batch* b, int32_t n_peers, double* const* peer_out) {
    PNB_REQUIRE(b != nullptr, PNB_ERR_INVALID, "pnb_batch_set_peer_outputs: NULL batch");
    PNB_REQUIRE(n_peers >= 0 && n_peers <= MAX_PEERS, PNB_ERR_INVALID, "pnb_batch_set_peer_outputs: %d peers (at most %d)",
                n_peers, MAX_PEERS);
    PNB_REQUIRE(n_peers == 0 || peer_out, PNB_ERR_INVALID, "pnb_batch_set_peer_outputs: NULL array");
    for (int r = 0; r < n_peers; ++r)
        PNB_REQUIRE(peer_out[r] != nullptr, PNB_ERR_INVALID, "pnb_batch_set_peer_outputs: peer %d is NULL", r);
    for (int r = 0; r < MAX_PEERS; ++r) b->peer_out[r] = r < n_peers ? peer_out[r] : nullptr;
    b->n_peers = n_peers;
    return PNB_OK;
}

extern "C" int pnb_batch_set_peer_barrier(pnb_batch* b, int32_t n_peers, uint64_t* const* peer_flag, const uint64_t* my_flags,
                                          const int32_t* peer_rank, uint64_t epoch) {
    PNB_REQUIRE(b != nullptr, PNB_ERR_INVALID, "pnb_batch_set_peer_barrier: NULL batch");
    if (n_peers == 0 || epoch == 0) {
        b->bar_epoch = 0;
        return PNB_OK;
    }
    PNB_REQUIRE(n_peers > 0 && n_peers <= MAX_PEERS, PNB_ERR_INVALID, "pnb_batch_set_peer_barrier: %d peers (at most %d)", n_peers,
                MAX_PEERS);
    PNB_REQUIRE(peer_flag && my_flags && peer_rank, PNB_ERR_INVALID, "pnb_batch_set_peer_barrier: NULL array");
    PNB_REQUIRE(n_peers == b->n_peers, PNB_ERR_STATE,
                "pnb_batch_set_peer_barrier: %d peers, but pnb_batch_set_peer_outputs named %d (the barrier orders THOSE stores)",
                n_peers, b->n_peers);
    for (int r = 0; r < n_peers; ++r)
        PNB_REQUIRE(peer_flag[r] != nullptr && peer_rank[r] >= 0 && peer_rank[r] < 64, PNB_ERR_INVALID,
                    "pnb_batch_set_peer_barrier: peer %d: NULL flag slot or bad rank", r);
    pnb_ctx* ctx = b->ctx;
    PNB_CUDA_TRY(cudaSetDevice(ctx->device));
    if (!b->d_launch_done.p) {
        int rc = ensure_device(b->d_launch_done, 256);
        if (rc) return rc;
        PNB_CUDA_TRY(cudaMemsetAsync(b->d_launch_done.p, 0, 256, ctx->stream));
    }
    if (!b->h_bar_timeout) {
        void* p = nullptr;
        PNB_CUDA_TRY(cudaHostAlloc(&p, 64, cudaHostAllocMapped));
        memset(p, 0, 64);
        b->h_bar_timeout = static_cast<int32_t*>(p);
    }
    for (int r = 0; r < MAX_PEERS; ++r) {
        b->bar_peer[r] = r < n_peers ? reinterpret_cast<unsigned long long*>(peer_flag[r]) : nullptr;
        b->bar_rank[r] = r < n_peers ? peer_rank[r] : 0;
    }
    b->bar_mine = reinterpret_cast<const unsigned long long*>(my_flags);
    b->bar_epoch = epoch;
    return PNB_OK;
}

extern "C" int pnb_batch_barrier_timed_out(const pnb_batch* b, int32_t* timed_out) {
    PNB_REQUIRE(b && timed_out, PNB_ERR_INVALID, "pnb_batch_barrier_timed_out: NULL argument");
    *timed_out = b->h_bar_timeout ? *static_cast<volatile int32_t*>(b->h_bar_timeout) : 0;
    return PNB_OK;
}

extern "C" int pnb_batch_blen_host(pnb_batch* b, double** pinned_out) {
    PNB_REQUIRE(b && pinned_out, PNB_ERR_INVALID, "pnb_batch_blen_host: NULL argument");
    *pinned_out = static_cast<double*>(b->h_blen.p);
    return PNB_OK;
}

extern "C" int pnb_batch_eval(pnb_batch* b, const pnb_model* m, const double* blen, double site_rate,
                              uint32_t flags, double* logL_out) {
    PNB_REQUIRE(b && m && blen && logL_out, PNB_ERR_INVALID, "pnb_batch_eval: NULL argument");
    PNB_REQUIRE(!b->dead, PNB_ERR_STATE,
                "pnb_batch_eval: a locus of this batch was destroyed or re-allocated; create the batch again");
    pnb_ctx* ctx = b->ctx;
    PNB_REQUIRE(m->ctx == ctx && m->mp.kind != 2, PNB_ERR_INVALID, "pnb_batch_eval: foreign or explicit model");
    PNB_REQUIRE((flags & ~(PNB_EVAL_OUT_DEVICE | PNB_BATCH_BLEN_DEVICE)) == 0, PNB_ERR_INVALID,
                "pnb_batch_eval: only PNB_EVAL_OUT_DEVICE / PNB_BATCH_BLEN_DEVICE are evaluation flags");
    PNB_REQUIRE(std::isfinite(site_rate), PNB_ERR_INVALID, "pnb_batch_eval: site_rate is not finite");
    const auto t_call0 = std::chrono::steady_clock::now();
    const bool out_dev = flags & PNB_EVAL_OUT_DEVICE;
    const bool blen_dev = flags & PNB_BATCH_BLEN_DEVICE;
    const bool nostore = b->flags & PNB_EVAL_NOSTORE;
    const bool rescale = b->flags & PNB_EVAL_RESCALE;
    const int64_t M = b->M;
    PNB_CUDA_TRY(cudaSetDevice(ctx->device));
    memset(ctx->stats, 0, sizeof(ctx->stats));
    ctx->ev_ka = ctx->ev_prune = false;
    if (!nostore) {
        // the loci now describe THIS evaluation; nothing else is written per locus (locus_sync is lazy)
        for (int64_t j = 0; j < M; ++j) {
            pnb_locus* loc = b->loci[j];
            if (loc->has_pending && loc->lazy_batch != b) {
                set_error("pnb_batch_eval: locus has a pending evaluation; commit or roll back first (job %lld)", (long long)j);
                return PNB_ERR_STATE;
            }
        }
        for (int64_t j = 0; j < M; ++j) {
            pnb_locus* loc = b->loci[j];
            loc->lazy_batch = b;
            loc->lazy_job = j;
            loc->has_pending = false;
        }
    }
    cudaStream_t st = ctx->stream;
    PruneLaunch L;
    L.jobs = static_cast<const JobDesc*>(b->d_jobs.p);
    L.tile_job = static_cast<const int32_t*>(b->d_tile_job.p);
    L.Pm = static_cast<const double*>(b->d_P.p);
    L.t = nullptr;
    L.tile_sum = static_cast<double*>(b->d_tile_sum.p);
    L.job_done = static_cast<unsigned int*>(b->d_job_done.p);
    L.out = out_dev ? logL_out : static_cast<double*>(b->d_out.p);
    L.ops_cap = b->ops_cap;
    L.rows_cap = b->rows_cap;
    L.stk_cap = b->stk_cap;
    L.tile_base = 0;
    L.n_peers = out_dev ? b->n_peers : 0;
    for (int r = 0; r < MAX_PEERS; ++r) L.peer_out[r] = b->peer_out[r];
    // the barrier is armed for this evaluation only (the caller owns the epoch counter that lives with the flags)
    L.bar_epoch = (L.n_peers > 0) ? b->bar_epoch : 0;
    b->bar_epoch = 0;
    L.launch_jobs = static_cast<int32_t>(b->jobs_with_tiles);
    L.out_len = static_cast<int32_t>(M);
    if (b->jobs_with_tiles == 0) L.bar_epoch = 0;   // nothing runs, nothing can signal
    L.launch_done = static_cast<unsigned int*>(b->d_launch_done.p);
    L.bar_mine = b->bar_mine;
    L.bar_timeout = nullptr;
    for (int r = 0; r < MAX_PEERS; ++r) { L.bar_peer[r] = b->bar_peer[r]; L.bar_rank[r] = b->bar_rank[r]; }
    if (L.bar_epoch) {
        void* dp = nullptr;
        PNB_CUDA_TRY(cudaHostGetDevicePointer(&dp, b->h_bar_timeout, 0));
        L.bar_timeout = static_cast<int32_t*>(dp);
    }
    int64_t h2d = 0, d2h = 0, launches = 0;
    const double* d_blen = static_cast<const double*>(b->d_blen.p);
    int rc;
    if (blen_dev) {
        // lengths already in HBM (the caller's array is read in place): two launches, nothing else
        if (ctx->profiling) PNB_CUDA_TRY(cudaEventRecord(ctx->ev[0], st));
        // (store mode: the gather also leaves the raw lengths in the batch's own array for a later locus_sync)
        double* keep = (!nostore && blen != d_blen) ? static_cast<double*>(b->d_blen.p) : nullptr;
        if ((rc = launch_pmatrix_gather(m->mp, blen, static_cast<const int32_t*>(b->d_node_of_op.p), site_rate, b->total_ops,
                                        static_cast<double*>(b->d_P.p), keep, st)))
            return rc;
        if (ctx->profiling) { PNB_CUDA_TRY(cudaEventRecord(ctx->ev[1], st)); ctx->ev_ka = true; PNB_CUDA_TRY(cudaEventRecord(ctx->ev[2], st)); }
        if ((rc = launch_prune(L, m->mp, static_cast<unsigned>(b->total_tiles), b->dyn_smem, rescale, st))) return rc;
        if (ctx->profiling) { PNB_CUDA_TRY(cudaEventRecord(ctx->ev[3], st)); ctx->ev_prune = true; }
        launches = 2;
        if (!nostore) b->last_blen_valid = false;
    } else {
        // host lengths: chunk c+1 is staged / uploaded / turned into P(t) on the copy stream while the main
        // stream prunes chunk c
        cudaPointerAttributes attr;
        bool pinned = cudaPointerGetAttributes(&attr, blen) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        if (!pinned) (void)cudaGetLastError();
        const int64_t n_chunks = static_cast<int64_t>(b->ck_job.size()) - 1;
        const bool two = n_chunks > 1;
        cudaStream_t up = two ? ctx->copy_stream : st;
        if (two) {
            PNB_CUDA_TRY(cudaEventRecord(ctx->prev_done, st));
            PNB_CUDA_TRY(cudaStreamWaitEvent(up, ctx->prev_done, 0));
        } else if (!pinned) {
            PNB_CUDA_TRY(cudaStreamSynchronize(st));  // the pinned staging buffer may still be read by the previous upload
        }
        if (two && !pinned) PNB_CUDA_TRY(cudaStreamSynchronize(up));
        double* hb = static_cast<double*>(b->h_blen.p);
        for (int64_t c = 0; c < n_chunks; ++c) {
            const int64_t n0 = b->ck_node[c], n1 = b->ck_node[c + 1];
            const int64_t o0 = b->ck_op[c], o1 = b->ck_op[c + 1];
            const int64_t t0 = b->ck_tile[c], t1 = b->ck_tile[c + 1];
            const double* src = blen + n0;
            if (!pinned && blen != hb) {
                memcpy(hb + n0, blen + n0, static_cast<size_t>(n1 - n0) * sizeof(double));
                src = hb + n0;
            }
            PNB_CUDA_TRY(cudaMemcpyAsync(static_cast<double*>(b->d_blen.p) + n0, src, static_cast<size_t>(n1 - n0) * sizeof(double),
                                         cudaMemcpyHostToDevice, up));
            h2d += (n1 - n0) * static_cast<int64_t>(sizeof(double));
            if (o1 > o0) {
                if (ctx->profiling && c == 0) PNB_CUDA_TRY(cudaEventRecord(ctx->ev[0], up));
                if ((rc = launch_pmatrix_gather(m->mp, d_blen, static_cast<const int32_t*>(b->d_node_of_op.p) + o0, site_rate, o1 - o0,
                                                static_cast<double*>(b->d_P.p) + o0 * 16, nullptr, up)))
                    return rc;
                if (ctx->profiling && c == 0) { PNB_CUDA_TRY(cudaEventRecord(ctx->ev[1], up)); ctx->ev_ka = true; }
                launches++;
            }
            if (two) {
                PNB_CUDA_TRY(cudaEventRecord(ctx->chunk_ready[c], up));
                PNB_CUDA_TRY(cudaStreamWaitEvent(st, ctx->chunk_ready[c], 0));
            }
            if (t1 > t0) {
                if (ctx->profiling && c == 0) PNB_CUDA_TRY(cudaEventRecord(ctx->ev[2], st));
                L.tile_base = static_cast<int32_t>(t0);
                if ((rc = launch_prune(L, m->mp, static_cast<unsigned>(t1 - t0), b->dyn_smem, rescale, st))) return rc;
                if (ctx->profiling && c == n_chunks - 1) { PNB_CUDA_TRY(cudaEventRecord(ctx->ev[3], st)); ctx->ev_prune = true; }
                launches++;
            }
        }
        if (!nostore) {
            b->last_blen.assign(blen, blen + b->total_nodes);
            b->last_blen_valid = true;
        }
    }
    b->epoch++;
    b->model_id = m->id;
    b->site_rate = site_rate;
    b->last_out_valid = false;
    if (!out_dev) {
        PNB_CUDA_TRY(cudaMemcpyAsync(b->h_out.p, b->d_out.p, static_cast<size_t>(M) * sizeof(double), cudaMemcpyDeviceToHost, st));
        d2h += M * static_cast<int64_t>(sizeof(double));
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            set_error("pnb_batch_eval: kernel execution failed: %s", cudaGetErrorString(e));
            b->dead = true;
            return PNB_ERR_CUDA;
        }
        memcpy(logL_out, b->h_out.p, static_cast<size_t>(M) * sizeof(double));
        b->last_out_valid = true;
    }
    ctx->t_compile_ms = 0.0;
    ctx->t_call_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call0).count();
    ctx->stats[0] = b->updates;
    ctx->stats[1] = b->total_ops;
    ctx->stats[2] = b->total_tiles;
    ctx->stats[3] = launches;
    ctx->stats[4] = h2d;
    ctx->stats[5] = d2h;
    ctx->stats[6] = 0;
    ctx->stats[7] = b->n_dirty;
    return PNB_OK;
}
