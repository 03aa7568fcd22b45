// pnb_host.hpp -- host-side runtime objects behind the C ABI (context, device
// arena, model, locus with its committed partials cache).  Not part of the ABI.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
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
    int sm_count = 0;
    int max_smem_optin = 0;
    uint64_t next_model_id = 1;
    uint32_t batch_counter = 0;
    pnb::Arena arena;
    pnb::Buffer h_stage, d_stage;  // pinned host / device staging of the batch program
    pnb::Buffer d_P;               // [ops][16] P matrices (K-a output)
    pnb::Buffer d_tile_sum, d_job_done, d_out, h_out;
    size_t job_done_zeroed = 0;    // counters [0, job_done_zeroed) are known to be zero
    int64_t stats[8] = {0};
    bool profiling = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // K-a begin/end, prune begin/end
    bool ev_ka = false, ev_prune = false;
    double t_compile_ms = 0.0, t_call_ms = 0.0;
    int64_t n_loci_alive = 0;
    // compile scratch (reused across calls)
    std::vector<int32_t> s_child_cnt, s_child_off, s_child, s_post, s_stack, s_ref, s_need, s_order;
    std::vector<uint8_t> s_dirty;
    std::vector<uint64_t> s_len;
    std::vector<std::pair<int64_t, uint64_t>> s_key;
};

struct pnb_locus {
    pnb_ctx* ctx = nullptr;
    int32_t n_rows = 0;
    int64_t P = 0, P_pad = 0;
    uint8_t* d_codes = nullptr;
    double* d_weights = nullptr;
    double* d_partials = nullptr;
    int32_t n_slots = 0;
    size_t codes_bytes = 0, weights_bytes = 0, partials_bytes = 0;

    // One cached tree description: which partial slot holds which (children, lengths) node.
    struct TreeCache {
        bool valid = false;
        uint64_t model_id = 0;
        std::vector<int32_t> kid_begin;   // per slot: range in kid_ref/kid_len, -1 = slot unused
        std::vector<int32_t> kid_count;
        std::vector<int64_t> kid_ref;     // sorted (ref, len) pairs; ref = tip row | n_rows + slot | -1 absent
        std::vector<uint64_t> kid_len;    // bit pattern of the effective branch length
        std::vector<int32_t> parent_of;   // per ref (n_rows + n_slots): slot of the node having it as child
        std::vector<int64_t> node_ref;    // per node of the tree as given by the caller
        int64_t root_ref = -2;
        bool logL_valid = false;
        double logL = 0.0;
        void clear() {
            valid = false; logL_valid = false; root_ref = -2;
            kid_begin.clear(); kid_count.clear(); kid_ref.clear(); kid_len.clear();
            parent_of.clear(); node_ref.clear();
        }
    };
    TreeCache committed, pending;
    bool has_pending = false;
    std::vector<int32_t> free_slots;
    uint32_t batch_mark = 0;  // duplicate detection inside one pnb_eval batch
};
