// pnb_simulate.cu -- K-s: sequence evolution down gene trees on the device (input generator).
//
// Reference: simulate_sequences / _evolve (src/_sim_seq.py:431-504): a root state drawn from pi, then every
// child state drawn from the row P(t)[parent state] of its branch, site by site.  Here: M trees in one launch,
// one thread per (tree, site), nodes visited parents-first, every draw from a COUNTER-BASED generator keyed by
// (seed, stream id of the tree, site, node) -- so the alignment of a tree does not depend on which other trees
// are in the batch, on the GPU count, or on the device at all: phynetpy_b200/simulate.py restates the same
// integer hash and the same comparisons in NumPy and gives the same bytes (tests/test_gpu_chain.py).  The P
// matrices come from the caller (host doubles), so both sides compare against identical thresholds.
#include <algorithm>
#include <vector>

#include "pnb_host.hpp"

using namespace pnb;

namespace {

constexpr int SIM_MAX_NODES = 512;

__host__ __device__ inline uint64_t sim_mix(uint64_t x) {   // splitmix64 finaliser
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

// uniform in [0, 1) with 53 bits, from (seed, stream, site, node)
__host__ __device__ inline double sim_uniform(uint64_t seed, uint64_t stream, uint64_t site, uint64_t node) {
    uint64_t x = sim_mix(seed + 0x9E3779B97F4A7C15ull * (stream + 1));
    x = sim_mix(x ^ (site * 0xD1B54A32D192ED03ull + node * 0x8CB92BA72F3D8DD7ull + 0x2545F4914F6CDD1Dull));
    return static_cast<double>(x >> 11) * (1.0 / 9007199254740992.0);
}

struct SimJob {
    int64_t node0;      // first node of the tree in the batch arrays
    int64_t out0;       // first byte of the tree's block in the output
    uint64_t stream;    // generator stream of the tree
    int32_t n_nodes;
    int32_t pad;
};

// order[node0 + k]: k-th node of the tree, parents before children; tip_rank[node0 + v]: output row of leaf v, -1 otherwise
__global__ void __launch_bounds__(128)
pnb_simulate_kernel(const SimJob* __restrict__ jobs, const int32_t* __restrict__ parent, const int32_t* __restrict__ order,
                    const int32_t* __restrict__ tip_rank, const double* __restrict__ cum /* [nodes][4][4] cumulative rows */,
                    double c0, double c1, double c2, int64_t L, uint64_t seed, uint8_t* __restrict__ out) {
    const SimJob jb = jobs[blockIdx.y];
    const int64_t site = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (site >= L) return;
    uint8_t st[SIM_MAX_NODES];
    const int32_t* par = parent + jb.node0;
    const int32_t* ord = order + jb.node0;
    const int32_t* tr = tip_rank + jb.node0;
    for (int k = 0; k < jb.n_nodes; ++k) {
        const int v = ord[k];
        const double u = sim_uniform(seed, jb.stream, static_cast<uint64_t>(site), static_cast<uint64_t>(v));
        int s;
        if (par[v] < 0) {
            s = (u > c0) + (u > c1) + (u > c2);
        } else {
            const double* row = cum + (jb.node0 + v) * 16 + st[par[v]] * 4;
            s = (u > row[0]) + (u > row[1]) + (u > row[2]);
        }
        st[v] = static_cast<uint8_t>(s);
        const int r = tr[v];
        if (r >= 0) out[jb.out0 + static_cast<int64_t>(r) * L + site] = "ACGT"[s];
    }
}

}  // namespace

// cumulative thresholds of one row-stochastic matrix (sequential sums, the same on host and device)
static inline void cum_rows(const double* P, double* c) {
    for (int i = 0; i < 4; ++i) {
        double s = 0.0;
        for (int j = 0; j < 4; ++j) {
            s += P[i * 4 + j] > 0.0 ? P[i * 4 + j] : 0.0;
            c[i * 4 + j] = s;
        }
    }
}

extern "C" int pnb_simulate_alignments(pnb_ctx* ctx, int64_t M, const int64_t* node_off, const int32_t* parent,
                                       const double* P_edge, const double pi[4], int64_t L, uint64_t seed,
                                       const int64_t* stream_id, uint8_t* out, int64_t* out_off) {
    PNB_REQUIRE(ctx && ctx->device >= 0, PNB_ERR_STATE, "pnb_simulate_alignments: needs a device context");
    PNB_REQUIRE(M >= 0 && L >= 0, PNB_ERR_INVALID, "pnb_simulate_alignments: negative size");
    if (M == 0 || L == 0) return PNB_OK;
    PNB_REQUIRE(node_off && parent && P_edge && pi && stream_id && out && out_off, PNB_ERR_INVALID,
                "pnb_simulate_alignments: NULL array");
    PNB_CUDA_TRY(cudaSetDevice(ctx->device));
    const int64_t total = node_off[M] - node_off[0];
    std::vector<SimJob> jobs(static_cast<size_t>(M));
    std::vector<int32_t> order(static_cast<size_t>(total)), tip_rank(static_cast<size_t>(total), -1), par(static_cast<size_t>(total));
    std::vector<double> cum(static_cast<size_t>(total) * 16);
    std::vector<int32_t> cnt, head, next, stack;
    int64_t out_bytes = 0;
    for (int64_t j = 0; j < M; ++j) {
        const int64_t b = node_off[j] - node_off[0];
        const int64_t nn = node_off[j + 1] - node_off[j];
        PNB_REQUIRE(nn >= 1 && nn <= SIM_MAX_NODES, PNB_ERR_UNSUPPORTED,
                    "pnb_simulate_alignments: tree %lld has %lld nodes (1..%d supported)", (long long)j, (long long)nn, SIM_MAX_NODES);
        const int32_t* p = parent + node_off[j];
        // children lists (in node order), then a parents-first order
        head.assign(nn, -1); next.assign(nn, -1); cnt.assign(nn, 0);
        int root = -1;
        for (int64_t v = nn - 1; v >= 0; --v) {
            par[b + v] = p[v];
            if (p[v] < 0) { PNB_REQUIRE(root < 0, PNB_ERR_INVALID, "pnb_simulate_alignments: tree %lld has two roots", (long long)j); root = static_cast<int>(v); }
            else {
                PNB_REQUIRE(p[v] < nn && p[v] != v, PNB_ERR_INVALID, "pnb_simulate_alignments: parent out of range (tree %lld)", (long long)j);
                next[v] = head[p[v]]; head[p[v]] = static_cast<int32_t>(v); cnt[p[v]]++;
            }
        }
        PNB_REQUIRE(root >= 0, PNB_ERR_INVALID, "pnb_simulate_alignments: tree %lld has no root", (long long)j);
        stack.assign(1, root);
        int64_t k = 0;
        while (!stack.empty()) {
            const int v = stack.back(); stack.pop_back();
            order[b + k++] = v;
            for (int c = head[v]; c >= 0; c = next[c]) stack.push_back(c);
        }
        PNB_REQUIRE(k == nn, PNB_ERR_INVALID, "pnb_simulate_alignments: tree %lld is not connected", (long long)j);
        int rank = 0;
        for (int64_t v = 0; v < nn; ++v)
            if (cnt[v] == 0) tip_rank[b + v] = rank++;
        for (int64_t v = 0; v < nn; ++v) cum_rows(P_edge + (node_off[j] + v) * 16, cum.data() + (b + v) * 16);
        jobs[j] = {b, out_bytes, static_cast<uint64_t>(stream_id[j]), static_cast<int32_t>(nn), 0};
        out_off[j] = out_bytes;
        out_bytes += static_cast<int64_t>(rank) * L;
    }
    out_off[M] = out_bytes;
    double c[3];
    {
        double s = 0.0;
        for (int i = 0; i < 3; ++i) { s += pi[i]; c[i] = s; }
    }
    Buffer d_jobs, d_par, d_order, d_rank, d_cum, d_out;
    auto free_all = [&]() { for (Buffer* q : {&d_jobs, &d_par, &d_order, &d_rank, &d_cum, &d_out}) if (q->p) cudaFree(q->p); };
    int rc;
    if ((rc = ensure_device(d_jobs, jobs.size() * sizeof(SimJob))) || (rc = ensure_device(d_par, par.size() * 4)) ||
        (rc = ensure_device(d_order, order.size() * 4)) || (rc = ensure_device(d_rank, tip_rank.size() * 4)) ||
        (rc = ensure_device(d_cum, cum.size() * 8)) || (rc = ensure_device(d_out, static_cast<size_t>(std::max<int64_t>(out_bytes, 1))))) {
        free_all();
        return rc;
    }
    cudaStream_t st = ctx->stream;
    cudaError_t e = cudaMemcpyAsync(d_jobs.p, jobs.data(), jobs.size() * sizeof(SimJob), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_par.p, par.data(), par.size() * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_order.p, order.data(), order.size() * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_rank.p, tip_rank.data(), tip_rank.size() * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_cum.p, cum.data(), cum.size() * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        // grid.y is limited to 65535: launch the trees in slabs
        for (int64_t j0 = 0; j0 < M && e == cudaSuccess; j0 += 32768) {
            const int64_t mj = std::min<int64_t>(32768, M - j0);
            dim3 grid(static_cast<unsigned>((L + 127) / 128), static_cast<unsigned>(mj));
            pnb_simulate_kernel<<<grid, 128, 0, st>>>(static_cast<const SimJob*>(d_jobs.p) + j0, static_cast<const int32_t*>(d_par.p),
                                                      static_cast<const int32_t*>(d_order.p), static_cast<const int32_t*>(d_rank.p),
                                                      static_cast<const double*>(d_cum.p), c[0], c[1], c[2], L, seed,
                                                      static_cast<uint8_t*>(d_out.p));
            e = cudaGetLastError();
        }
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out.p, static_cast<size_t>(out_bytes), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    free_all();
    if (e != cudaSuccess) {
        set_error("pnb_simulate_alignments: %s", cudaGetErrorString(e));
        return PNB_ERR_CUDA;
    }
    return PNB_OK;
}
