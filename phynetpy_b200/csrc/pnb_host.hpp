// pnb_host.hpp -- host-side runtime objects behind the C ABI (context, device
// arena, model, locus with its committed partials cache).  Not part of the ABI.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/phynet_b200.h"
#include "pnb_device.cuh"

namespace pnb {

void set_error(const char* fmt, ...);

#define PNB_CUDA_TRY(expr)                                                              \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            pnb::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, \
                           __LINE__, cudaGetErrorString(_e));                           \
            return PNB_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

#define PNB_REQUIRE(cond, code, ...)      \
    do {                                  \
        if (!(cond)) {                    \
            pnb::set_error(__VA_ARGS__);  \
            return (code);                \
        }                                 \
    } while (0)

// Device arena: loci are many and long-lived (10^4 per GPU), so their buffers are
// carved out of a few large cudaMalloc chunks; freed blocks are recycled by size.
class Arena {
  public:
    int alloc(size_t bytes, void** out);
    void release(void* p, size_t bytes);
    void destroy();
    size_t reserved() const { return reserved_; }

  private:
    struct Chunk { char* base; size_t cap; size_t used; };
    std::vector<Chunk> chunks_;
    std::map<size_t, std::vector<void*>> free_;
    size_t reserved_ = 0;
};

// A growable device buffer / pinned host buffer pair used for per-call staging.
struct Buffer {
    void* p = nullptr;
    size_t cap = 0;
};
int ensure_device(Buffer& b, size_t bytes);
int ensure_pinned(Buffer& b, size_t bytes);

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline uint64_t dbits(double x) {
    uint64_t u;
    memcpy(&u, &x, 8);
    return u;
}

// Result of compiling one (locus, tree) job into an op program.
struct JobOut {
    int rc = 0;
    int n_ops = 0;
    int n_tip_ops = 0;
    int rows_cap = 0;           // staged code rows per program segment (max over segments)
    int root_kind = 0;
    int root_src = 0;
    int stk_need = 0;
    int n_dirty = 0;
    bool root_clean_cached = false;
    double cached_logL = 0.0;   // committed logL captured before an in-worker commit
    bool in_place = false;      // the committed cache was refreshed in place: nothing to commit
    int8_t tmpl_hit = -1;       // >= 0: program taken from tmpl[tmpl_hit]; its resident device copy can be used
    int8_t tmpl_new = -1;       // >= 0: tmpl[tmpl_new] was (re)built by this job and needs its device copy
    const Op* res_ops = nullptr; // template hit served from the locus' device-resident program (captured under the
                                 // template read lock; the buffer is not rewritten while this batch references it)
};

// Per-thread scratch of the tree compiler (no allocation in steady state).
struct CompileScratch {
    struct Frame { int v; int i; int sp; int push_idx; int kd; };
    std::vector<int32_t> child_cnt, child_off, child, post, stack, need, order, dk, dk_base;
    std::vector<uint8_t> dirty, virt;
    std::vector<uint64_t> len;
    std::vector<int64_t> ref;
    struct Key { int64_t ref; uint64_t len; int32_t child; };
    std::vector<Key> key;
    std::vector<int32_t> op_child, kid_child;
    std::vector<Frame> frames;
    // per-worker accumulators of one chunk's descriptor pass (reduced by the caller)
    struct ChunkAcc {
        int64_t n_jobs = 0, n_ops_used = 0, updates = 0, n_cached = 0, n_dirty = 0;
        int64_t ops_lo = -1, ops_hi = 0, prog_lo = -1, prog_hi = 0;
        int ops_cap = 1, rows_cap = 1, stk_cap = 0;
        std::vector<int64_t> save;
        void reset() { *this = ChunkAcc(); }
    } acc;
    int64_t first_err_job = -1;
    int rc_alloc = 0;
    int64_t tiles = 0;            // pass-0 partial sums
    std::vector<int64_t> grow;    // jobs whose locus needs more partial slots
    char err[512] = "";
};

// Minimal persistent worker pool: run(f) executes f(worker_index) on every worker (the caller
// is worker 0) and returns when all are done.  Used to compile large batches of trees.  Workers
// spin briefly after a task before sleeping: evaluations arrive back to back in an MCMC loop and a
// futex wake-up (tens of microseconds) would cost more than compiling a few hundred trees.
class WorkerPool {
  public:
    explicit WorkerPool(int n) : n_(n < 1 ? 1 : n) {
        for (int w = 1; w < n_; ++w) threads_.emplace_back([this, w] { loop(w); });
    }
    ~WorkerPool() {
        stop_.store(true, std::memory_order_release);
        gen_.fetch_add(1, std::memory_order_release);
        {
            std::lock_guard<std::mutex> g(m_);
            cv_.notify_all();
        }
        for (auto& t : threads_) t.join();
    }
    int size() const { return n_; }
    void run(const std::function<void(int)>& f) {
        fn_ = &f;
        pending_.store(n_ - 1, std::memory_order_relaxed);
        gen_.fetch_add(1, std::memory_order_release);
        {
            std::lock_guard<std::mutex> g(m_);
            if (sleepers_ > 0) cv_.notify_all();
        }
        f(0);
        int spins = 0;
        while (pending_.load(std::memory_order_acquire) != 0) {
            if (++spins > 4096) std::this_thread::yield();
        }
        fn_ = nullptr;
    }

  private:
    static constexpr int kSpin = 20000;  // ~50 us of polling before a worker goes to sleep
    void loop(int w) {
        uint64_t seen = 0;
        for (;;) {
            int spins = 0;
            while (gen_.load(std::memory_order_acquire) == seen) {
                if (++spins < kSpin) {
#if defined(__x86_64__)
                    __builtin_ia32_pause();
#endif
                } else {
                    std::unique_lock<std::mutex> lk(m_);
                    ++sleepers_;
                    cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
                    --sleepers_;
                }
            }
            seen = gen_.load(std::memory_order_acquire);
            if (stop_.load(std::memory_order_acquire)) return;
            const std::function<void(int)>* f = fn_;
            if (f) (*f)(w);
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }
    int n_;
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_;
    const std::function<void(int)>* volatile fn_ = nullptr;
    std::atomic<uint64_t> gen_{0};
    std::atomic<int> pending_{0};
    std::atomic<bool> stop_{false};
    int sleepers_ = 0;  // guarded by m_
};

}  // namespace pnb

struct pnb_model {
    pnb_ctx* ctx;
    pnb::ModelParams mp;
    uint64_t id;  // unique per model: cached partials are only valid for the model they were built with
};

struct pnb_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // uploads + K-a of chunk c+1 run here while `stream` prunes chunk c
    cudaEvent_t chunk_ready[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t prev_done = nullptr;      // last pruning launch of the previous evaluation (it reads d_stage / d_P)
    bool prev_done_valid = false;
    int sm_count = 0;
    int max_smem_optin = 0;
    uint64_t next_model_id = 1;
    uint32_t batch_counter = 0;
    int chunk_min_jobs = 1024;     // batches of at least this many jobs are launched in chunks
    int store_stash_levels = 0;    // storing evaluations: siblings waiting at stash index < this use shared memory
                                   // instead of an L2 round trip through their slot (PNB_STORE_STASH)
    bool hold_reg = pnb::HOLD_REGS;   // the innermost waiting sibling stays in the kernel's HOLD registers (PNB_HOLD builds)
    int virt_tips = 4;             // subtrees of at most this many tips are recomputed from the tip codes, never stored
                                   // (PNB_VIRT_TIPS; 2 = only tip-only nodes)
    bool allow_fused = true;       // small batches: single-kernel latency path (PNB_NO_FUSED=1 disables)
    bool msnc_configured = false;  // K-m kernels have their dynamic shared-memory limit raised
    pnb::Arena arena;
    pnb::Buffer d_stage;           // device staging of the batch program (stream-ordered reuse)
    pnb::Buffer h_stage2[2];       // pinned host staging, double-buffered so the host can run ahead
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};  // "H2D copy out of h_stage2[i] has completed"
    bool stage_ev_valid[2] = {false, false};
    int stage_idx = 0;
    pnb::Buffer d_P;               // [ops][16] P matrices (K-a output)
    pnb::Buffer d_tile_sum, d_job_done, d_out, h_out;
    // K-m (pnb_msnc.cu) owns its buffers: a gene-tree move runs it next to the pruning path (pnb_eval_move)
    pnb::Buffer m_hstage, m_dstage, m_scratch, m_dout, m_hout;
    int64_t msnc_pending_trees = 0;      // launched, not yet handed out (msnc_finish)
    size_t msnc_pending_status_off = 0;
    int msnc_pending_cap = 0;
    int64_t msnc_stats[8] = {0};
    size_t job_done_zeroed = 0;    // counters [0, job_done_zeroed) are known to be zero
    int64_t stats[8] = {0};
    bool profiling = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // K-a begin/end, prune begin/end
    bool ev_ka = false, ev_prune = false;
    double t_compile_ms = 0.0, t_call_ms = 0.0;
    int64_t n_loci_alive = 0;
    // tree compiler: per-worker scratch, worker pool (lazily sized), per-call job results
    std::vector<pnb::CompileScratch> scratch;
    pnb::WorkerPool* pool = nullptr;
    typedef pnb::JobOut JobOutRaw;
    std::vector<pnb::JobOut> job_out;
    std::vector<int64_t> job_of_batch;
    std::vector<int64_t> save_list;
    std::vector<int64_t> tile_first;   // per job of the batch: its first tile (prefix sum, M + 1 entries)
    // device program buffers replaced while a batch still referenced them: back to the arena at the next call
    std::vector<std::pair<void*, size_t>> deferred_release;
};

struct pnb_locus {
    pnb_ctx* ctx = nullptr;
    int32_t n_rows = 0;
    int64_t P = 0, P_pad = 0;
    uint8_t* d_codes = nullptr;
    double* d_weights = nullptr;
    double* d_partials = nullptr;
    int32_t* d_scale = nullptr;   // [n_slots][P_pad] exponents of the stored partials (RESCALE mode), lazily allocated
    size_t scale_bytes = 0;
    int32_t n_slots = 0;
    size_t codes_bytes = 0, weights_bytes = 0, partials_bytes = 0;

    // One cached tree description: which partial slot holds which (children, lengths) node.
    struct TreeCache {
        bool valid = false;
        bool rescale = false;             // the stored partials carry exponents (PNB_EVAL_RESCALE)
        uint64_t model_id = 0;
        std::vector<int32_t> kid_begin;   // per slot: range in kid_ref/kid_len, -1 = slot unused
        std::vector<int32_t> kid_count;
        std::vector<int64_t> kid_ref;     // sorted (ref, len) pairs; ref = tip row | n_rows + slot | -1 absent
        std::vector<uint64_t> kid_len;    // bit pattern of the effective branch length
        std::vector<int32_t> parent_of;   // per ref (n_rows + n_slots): slot of the node having it as child
        std::vector<int64_t> node_ref;    // per node of the tree as given by the caller
        std::vector<uint8_t> stored;      // per slot: the partial is materialised in HBM
        int64_t root_ref = -2;
        bool logL_valid = false;
        double logL = 0.0;
        void clear() {
            valid = false; logL_valid = false; root_ref = -2;
            kid_begin.clear(); kid_count.clear(); kid_ref.clear(); kid_len.clear();
            parent_of.clear(); node_ref.clear(); stored.clear();
        }
    };
    TreeCache committed, pending;
    bool has_pending = false;
    bool committed_from_tmpl = false, pending_from_tmpl = false;  // cache image == tmpl[0].cache (slots included)

    // Compiled-program template of the last tree that was evaluated with EVERY node recomputed
    // ([0] storing, [1] NOSTORE).  The op program depends on the topology only, so when the same
    // topology comes back with all nodes to recompute (full evaluations, height rescalings, the
    // 5 scale candidates of a coupled move) only the branch lengths are refreshed.
    struct Template {
        bool valid = false;
        int32_t n = 0, n_slots = 0, slots_used = 0;
        bool has_tip_row = false;
        std::vector<int32_t> parent, tip_row, op_child, kid_child;
        std::vector<pnb::Op> ops;
        pnb::JobOut jo;
        TreeCache cache;  // pending-cache image (store mode), kid_len refreshed per evaluation
        pnb::Op* d_ops = nullptr;   // device-resident copy of `ops` (arena): template hits send no program
        int32_t d_ops_cap = 0;      // capacity of d_ops in ops
        bool d_ops_valid = false;   // d_ops holds (or a stream-ordered copy will hold) exactly `ops`
        std::atomic<uint32_t> d_ops_used_mark{0};  // batch (ctx->batch_counter) whose jobs read d_ops in place
        Template() = default;
        Template(const Template&) = delete;
        Template& operator=(const Template&) = delete;
    };
    Template tmpl[2];
    // readers/writer spin state for tmpl[] (a NOSTORE batch may hold the same locus many times and is
    // compiled by several workers): >= 0 number of readers, -1 one writer.  Never waited on: a worker
    // that cannot get the lock simply takes the general path / skips the template update.
    std::atomic<int> tmpl_lock{0};
    std::vector<int32_t> free_slots;
    std::atomic<uint32_t> batch_mark{0};  // duplicate detection inside one pnb_eval batch
    // resident batches (pnb_batch_*) that hold this locus; lazy_batch != NULL: the committed tree is what that
    // batch's last evaluation computed, not yet written into `committed` (locus_sync does it on demand)
    std::vector<struct pnb_batch*> batches;
    struct pnb_batch* lazy_batch = nullptr;
    int64_t lazy_job = -1;
};

namespace pnb {
constexpr uint32_t PNB_COMPILE_NO_TEMPLATE = 0x80000000u;  // internal compile_job flag: no program templates
// pnb_batch.cu: bring a locus' committed-tree description up to date with the last evaluation of the resident
// batch that owns it (lazy: a batch evaluation touches no per-locus host state)
void locus_sync_slow(pnb_locus* loc);
inline void locus_sync(pnb_locus* loc) { if (loc->lazy_batch) locus_sync_slow(loc); }
void locus_detach_batches(pnb_locus* loc);
// pnb_eval.cu: kernel launchers for the other translation units (the kernels are compiled there only)
int launch_pmatrix_gather(const ModelParams& mp, const double* d_blen, const int32_t* d_node_of_op, double site_rate,
                          int64_t n_ops, double* d_P, double* d_keep, cudaStream_t st);
int launch_prune(const PruneLaunch& L, const ModelParams& mp, unsigned grid, size_t dyn_smem, bool rescale,
                 cudaStream_t st);
// pnb_runtime.cu
void locus_free_partials(pnb_locus* loc);
void locus_reset_cache(pnb_locus* loc);
int locus_ensure_slots(pnb_locus* loc, int n_internal);
// pnb_compile.cu
int compile_job(CompileScratch& S, pnb_locus* loc, const pnb_model* m, int32_t n, const int32_t* parent,
                const double* blen, const int32_t* tip_row, const double* P_edge, double site_rate,
                uint32_t flags, Op* ops, double* t_out, double* Pexp_out, JobOut* jo);
void locus_commit(pnb_locus* loc);
void locus_rollback(pnb_locus* loc);
// pnb_eval.cu: opt the pruning kernels into large dynamic shared memory; returns a cudaError_t
int configure_kernels(pnb_ctx* ctx);
}  // namespace pnb
