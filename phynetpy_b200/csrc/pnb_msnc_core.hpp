// pnb_msnc_core.hpp -- timed MSNC density log P(g | Psi): data layout, the host compiler of a
// species network into a step program, the host preparation of gene trees (coalescence events,
// species lineage sets), and the per-job dynamic programme shared by the CUDA kernel (K-m).
//
// What it computes is the reference's joint-edge ancestral-configuration DP
// (_msnc_log_density_timed, src/_msnc_density.py:213-389; C twin network_dp_timed_cy,
// src/cython/gt_msc_cy.pyx:458-676) with the event-based branch factor (_apply_branch_timed,
// src/_msnc_density.py:172-210).  How it is organised is different:
//
//   * the network is compiled ONCE into a program of steps (one per network node, children
//     before parents).  Every frontier entry of the reference is a sorted tuple of (edge id,
//     lineage set); all entries alive at one step carry the same edge ids, so here an entry is a
//     fixed-width vector of lineage sets indexed by SLOT, slots being assigned to edges by
//     liveness (like registers).  Nodes are ordered depth-first, widest subtree first, which keeps
//     the number of simultaneously open edges -- the vector width W -- small (the reference's
//     breadth-first order would open one edge per species);
//   * one (gene tree) job is one GPU thread; its frontier lives in a private block of a global
//     scratch buffer, ping-ponged between steps; duplicates are merged by log-sum-exp in
//     first-insertion order (the order of the reference's dict);
//   * lineage sets are NW 64-bit words (NW = 1 up to 64 gene-tree nodes, 2 up to 128).
//
// This header has no CUDA dependency: tests/native compiles it with g++ to run the very same DP
// source on the host in the CPU test-suite (test infrastructure; the product only runs it on the GPU).
#pragma once

#include <stdint.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#if defined(__CUDACC__)
#define PNB_HD __host__ __device__ __forceinline__
#else
#define PNB_HD inline
#endif

namespace pnb {
namespace msnc {

constexpr double LOG_FLOOR = -460.51701859880916;   // log(1e-200), src/_msnc_density.py:543
constexpr int MAX_NW = 2;                            // lineage sets of up to 128 gene-tree nodes
constexpr int KIND_LEAF = 1, KIND_ROOT = 2, KIND_RETIC = 4;

// One network node of the step program.
struct Step {
    int32_t kind;          // KIND_* bits
    int32_t species;       // KIND_LEAF: species id whose lineages enter here (-1: none)
    int32_t n_down;        // slots merged (and closed) at this node
    int32_t down_off;      // -> down_slot[]
    int32_t up_slot[2];    // slot(s) opened above this node ([1] only for a reticulation)
    int32_t pad[2];
    double tau_low;        // height of this node
    double tau_high[2];    // height of the parent(s); unused for the root
    double log_g[2];       // log inheritance probabilities of the two parent edges (reticulation)
};

static_assert(sizeof(Step) == 72 && sizeof(Step) % 8 == 0, "Step is copied to shared memory in 8-byte words");

// Exactly-rounded product / sum: the device must not contract a*b+c into an FMA, so that a
// tree-shaped network gives the reference's bits.
#if defined(__CUDA_ARCH__)
PNB_HD double mul_rn(double a, double b) { return __dmul_rn(a, b); }
PNB_HD double add_rn(double a, double b) { return __dadd_rn(a, b); }
PNB_HD int popc64(uint64_t x) { return __popcll(x); }
#else
PNB_HD double mul_rn(double a, double b) { return a * b; }
PNB_HD double add_rn(double a, double b) { return a + b; }
PNB_HD int popc64(uint64_t x) { return __builtin_popcountll(x); }
#endif

template <int NW>
struct Mask {
    uint64_t w[NW];
    PNB_HD void clear() {
#pragma unroll
        for (int i = 0; i < NW; ++i) w[i] = 0;
    }
    PNB_HD bool test(int b) const { return (w[b >> 6] >> (b & 63)) & 1u; }
    PNB_HD void set(int b) { w[b >> 6] |= (uint64_t{1} << (b & 63)); }
    PNB_HD void reset(int b) { w[b >> 6] &= ~(uint64_t{1} << (b & 63)); }
    PNB_HD int count() const {
        int c = 0;
#pragma unroll
        for (int i = 0; i < NW; ++i) c += popc64(w[i]);
        return c;
    }
    PNB_HD bool zero() const {
        uint64_t o = 0;
#pragma unroll
        for (int i = 0; i < NW; ++i) o |= w[i];
        return o == 0;
    }
    // next subset of `of` in decreasing order: (this - 1) & of, over NW words
    PNB_HD void prev_subset(const Mask& of) {
        bool borrow = true;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            const uint64_t v = w[i];
            w[i] = (borrow ? v - 1 : v) & of.w[i];
            borrow = borrow && (v == 0);
        }
    }
};

// One lineage set through one species branch (reference _apply_branch_timed, single entry).
// ev_n packs parent | child0 << 8 | child1 << 16 (gene-tree node indices).  [j_lo, j_hi) is the range of
// events with tau_low <= t < tau_high -- the events the reference's loop does not skip (`t < tau_low:
// continue`) before it stops (`t >= tau_high: break`); it depends on the step only, so the caller finds it
// once per step instead of re-scanning the event list for every frontier entry and every subset.
template <int NW>
PNB_HD void branch_timed(Mask<NW>& cfg, double& lp_out, double tau_low, double tau_high, bool has_high,
                         double inv_theta, double log_two_over_theta, const double* ev_t,
                         const uint32_t* ev_n, int j_lo, int j_hi) {
    double cur = tau_low, lp = 0.0;
    int u = cfg.count();
    if (u <= 1) {   // nothing can coalesce and the survival term needs two lineages: the factor is exactly 1
        lp_out = 0.0;
        return;
    }
    for (int j = j_lo; j < j_hi; ++j) {
        const uint32_t e = ev_n[j];
        const int c0 = (e >> 8) & 0xff, c1 = (e >> 16) & 0xff;
        if (cfg.test(c0) && cfg.test(c1)) {
            const double t = ev_t[j];
            double b = mul_rn(mul_rn(mul_rn(-(t - cur), static_cast<double>(u)), static_cast<double>(u - 1)), inv_theta);
            b = add_rn(b, log_two_over_theta);
            cur = t;
            cfg.reset(c0);
            cfg.reset(c1);
            cfg.set(e & 0xff);
            --u;
            lp = add_rn(lp, b);
        }
    }
    if (has_high && u > 1)
        lp = add_rn(lp, mul_rn(mul_rn(mul_rn(-(tau_high - cur), static_cast<double>(u)), static_cast<double>(u - 1)), inv_theta));
    lp_out = lp;
}

// Events are sorted by time: first index with t >= tau (n_ev if none).
PNB_HD int first_event_at_or_after(const double* ev_t, int n_ev, double tau) {
    int lo = 0, hi = n_ev;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ev_t[mid] < tau) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

PNB_HD double log_add(double a, double b) {   // _frontier_acc, src/_msnc_density.py:1408-1419
    return (a > b) ? a + log1p(exp(b - a)) : b + log1p(exp(a - b));
}

// status of one job
constexpr int JOB_OK = 0, JOB_OVERFLOW = 1;

// Tree-shaped network (no reticulation: the MSC case): the frontier is always ONE entry, so there is
// nothing to copy, ping-pong or merge -- one vector of W lineage sets updated in place (slot k at
// slots[k * NW * stride + i * stride]) and one running log-density, added in step order exactly as the
// general DP would (lp + branch factor per step).  A closed slot is never read again before it is
// reopened, so it is not cleared.
template <int NW>
PNB_HD double msnc_dp_tree(const Step* steps, int n_steps, const int32_t* down_slot,
                           const double* ev_t, const uint32_t* ev_n, int n_ev,
                           const uint64_t* species_mask, int root_bit, const double* step_theta,
                           uint64_t* slots, size_t stride) {
    double lp = 0.0;
    Mask<NW> at_v;
    at_v.clear();
    for (int s = 0; s < n_steps; ++s) {
        const Step st = steps[s];
        at_v.clear();
        if (st.kind & KIND_LEAF) {
            if (st.species >= 0)
                for (int i = 0; i < NW; ++i) at_v.w[i] = species_mask[static_cast<size_t>(st.species) * NW + i];
        } else {
            for (int d = 0; d < st.n_down; ++d) {
                const size_t slot = static_cast<size_t>(down_slot[st.down_off + d]) * NW;
                for (int i = 0; i < NW; ++i) at_v.w[i] |= slots[(slot + i) * stride];
            }
        }
        const bool is_root = (st.kind & KIND_ROOT) != 0;
        if (at_v.count() > 1) {
            const int j_lo = first_event_at_or_after(ev_t, n_ev, st.tau_low);
            const int j_hi = is_root ? n_ev : first_event_at_or_after(ev_t, n_ev, st.tau_high[0]);
            double blp;
            branch_timed<NW>(at_v, blp, st.tau_low, st.tau_high[0], !is_root, step_theta[4 * s], step_theta[4 * s + 1],
                             ev_t, ev_n, j_lo, j_hi);
            lp = add_rn(lp, blp);
        }
        if (!is_root) {
            const size_t slot = static_cast<size_t>(st.up_slot[0]) * NW;
            for (int i = 0; i < NW; ++i) slots[(slot + i) * stride] = at_v.w[i];
        }
    }
    // the last step is the root (children before parents): its lineage set must be exactly the gene-tree root
    bool hit = true;
    for (int i = 0; i < NW; ++i) {
        const uint64_t want = ((root_bit >> 6) == i) ? (uint64_t{1} << (root_bit & 63)) : 0;
        hit = hit && (at_v.w[i] == want);
    }
    if (!hit) return -INFINITY;
    return (lp <= LOG_FLOOR + 1.0) ? -INFINITY : lp;
}

// Words of frontier storage one job needs: two ping-pong frontiers of cap entries (W*NW words each) and
// one 64-bit hash per entry; plus 2*cap doubles for the log-weights (msnc_lp_words).
PNB_HD size_t msnc_mask_words(int W, int NW, int cap) { return 2 * static_cast<size_t>(cap) * (static_cast<size_t>(W) * NW + 1); }
PNB_HD size_t msnc_lp_words(int cap) { return 2 * static_cast<size_t>(cap); }

// Position-keyed word hash; 0 for an empty lineage set, so closed slots contribute nothing and the hash of
// an entry can be updated incrementally (XOR) when slots are closed and opened.
PNB_HD uint64_t slot_hash(uint32_t word_index, uint64_t w) {
    uint64_t m = w * ((0x9E3779B97F4A7C15ull * (2ull * word_index + 1ull)) | 1ull);
    return m ^ (m >> 29);
}

// The DP of one gene tree (general case: reticulations).  Frontier storage, private to the job: fr_mask
// holds msnc_mask_words() words -- [2][cap] entries of W*NW words, then [2][cap] hashes -- and fr_lp
// 2*cap doubles; element k of either array lives at [k * stride], so the same code runs on a contiguous
// global block (stride 1) and on shared memory interleaved over the threads of a block (stride =
// blockDim.x: conflict-free).  Returns log P(g | Psi), -inf if the gene tree does not embed.
//
// Per (step, entry) the work is: OR the closed slots, run the branch factor(s), write ONE new entry (copy
// of the old one with the closed slots cleared and the opened slot(s) set) and look for an equal entry among
// those already written -- by hash first, word by word only on a hash match.  Equal entries merge by
// log-sum-exp in first-insertion order (the order of the reference's dict).
template <int NW>
PNB_HD double msnc_dp(const Step* steps, int n_steps, const int32_t* down_slot, int W, int cap,
                      const double* ev_t, const uint32_t* ev_n, int n_ev,
                      const uint64_t* species_mask /* [n_species][NW] */, int root_bit,
                      const double* step_theta /* [n_steps][4]: 1/theta and log(2/theta) of the up branch(es) */,
                      uint64_t* fr_mask, double* fr_lp, size_t stride_in, int* status) {
    const uint32_t stride = static_cast<uint32_t>(stride_in);
    const uint32_t entry = static_cast<uint32_t>(W) * NW;
    const uint32_t ucap = static_cast<uint32_t>(cap);
    uint32_t cur_m = 0, new_m = ucap * entry;                    // word offsets of the two frontiers
    uint32_t cur_h = 2 * ucap * entry, new_h = cur_h + ucap;     // ... and of their hashes
    uint32_t cur_lp = 0, new_lp = ucap;
#define PNB_M(off) fr_mask[(off) * stride]
#define PNB_LP(off) fr_lp[(off) * stride]
    for (uint32_t i = 0; i < entry; ++i) PNB_M(cur_m + i) = 0;
    PNB_M(cur_h) = 0;
    PNB_LP(cur_lp) = 0.0;
    int n_cur = 1;
    int root_slot = 0;
    *status = JOB_OK;

    for (int s = 0; s < n_steps; ++s) {
        const Step st = steps[s];
        int n_new = 0;
        const bool is_root = (st.kind & KIND_ROOT) != 0;
        const bool is_retic = (st.kind & KIND_RETIC) != 0;
        // the events inside this node's parent branch(es): a property of the step, shared by all entries;
        // found on first use (a branch entered by fewer than two lineages never looks at events)
        int j_lo = -1, j_hi0 = 0, j_hi1 = 0;
        auto ranges = [&]() {
            if (j_lo >= 0) return;
            j_lo = first_event_at_or_after(ev_t, n_ev, st.tau_low);
            j_hi0 = is_root ? n_ev : first_event_at_or_after(ev_t, n_ev, st.tau_high[0]);
            j_hi1 = is_retic ? first_event_at_or_after(ev_t, n_ev, st.tau_high[1]) : j_lo;
        };
        const double inv_theta0 = step_theta[4 * s], l2t0 = step_theta[4 * s + 1];
        const double inv_theta1 = step_theta[4 * s + 2], l2t1 = step_theta[4 * s + 3];
        // Writes (old entry at src, closed slots cleared, slot a = ma [, slot b = mb]) with hash h and
        // log-weight lp into the new frontier, or merges it into the equal entry already there.
        auto emit = [&](uint32_t src, uint64_t h, int a, const Mask<NW>& ma, int b, const Mask<NW>& mb, double lp) -> bool {
            if (n_new >= cap) return false;
            const uint32_t mine = new_m + static_cast<uint32_t>(n_new) * entry;
            for (uint32_t i = 0; i < entry; ++i) PNB_M(mine + i) = PNB_M(src + i);
            for (int d = 0; d < st.n_down; ++d)
                for (int i = 0; i < NW; ++i) PNB_M(mine + static_cast<uint32_t>(down_slot[st.down_off + d]) * NW + i) = 0;
            for (int i = 0; i < NW; ++i) PNB_M(mine + static_cast<uint32_t>(a) * NW + i) = ma.w[i];
            if (b >= 0)
                for (int i = 0; i < NW; ++i) PNB_M(mine + static_cast<uint32_t>(b) * NW + i) = mb.w[i];
            for (int j = 0; j < n_new; ++j) {
                if (PNB_M(new_h + j) != h) continue;
                const uint32_t other = new_m + static_cast<uint32_t>(j) * entry;
                bool same = true;
                for (uint32_t i = 0; i < entry; ++i) same = same && (PNB_M(other + i) == PNB_M(mine + i));
                if (same) {
                    PNB_LP(new_lp + j) = log_add(PNB_LP(new_lp + j), lp);
                    return true;
                }
            }
            PNB_M(new_h + n_new) = h;
            PNB_LP(new_lp + n_new) = lp;
            ++n_new;
            return true;
        };
        for (int k = 0; k < n_cur; ++k) {
            const uint32_t src = cur_m + static_cast<uint32_t>(k) * entry;
            const double lp = PNB_LP(cur_lp + k);
            uint64_t h = PNB_M(cur_h + k);
            Mask<NW> at_v;
            at_v.clear();
            if (st.kind & KIND_LEAF) {
                if (st.species >= 0)
                    for (int i = 0; i < NW; ++i) at_v.w[i] = species_mask[static_cast<size_t>(st.species) * NW + i];
            } else {
                for (int d = 0; d < st.n_down; ++d) {
                    const uint32_t slot = static_cast<uint32_t>(down_slot[st.down_off + d]) * NW;
                    for (int i = 0; i < NW; ++i) {
                        const uint64_t w = PNB_M(src + slot + i);
                        at_v.w[i] |= w;
                        h ^= slot_hash(slot + i, w);    // the edge is closed: its words leave the hash
                    }
                }
            }
            bool ok = true;
            if (is_retic) {
                const int n_total = at_v.count();
                const uint32_t sa = static_cast<uint32_t>(st.up_slot[0]) * NW, sb = static_cast<uint32_t>(st.up_slot[1]) * NW;
                Mask<NW> sub = at_v;
                while (true) {
                    const int ks = sub.count();
                    const double factor = add_rn(mul_rn(static_cast<double>(ks), st.log_g[0]),
                                                 mul_rn(static_cast<double>(n_total - ks), st.log_g[1]));
                    Mask<NW> top1 = sub, top2;
                    for (int i = 0; i < NW; ++i) top2.w[i] = at_v.w[i] ^ sub.w[i];
                    double lp1 = 0.0, lp2 = 0.0;
                    if (ks > 1) {
                        ranges();
                        branch_timed<NW>(top1, lp1, st.tau_low, st.tau_high[0], true, inv_theta0, l2t0, ev_t, ev_n, j_lo, j_hi0);
                    }
                    if (n_total - ks > 1) {
                        ranges();
                        branch_timed<NW>(top2, lp2, st.tau_low, st.tau_high[1], true, inv_theta1, l2t1, ev_t, ev_n, j_lo, j_hi1);
                    }
                    uint64_t hh = h;
                    for (int i = 0; i < NW; ++i) hh ^= slot_hash(sa + i, top1.w[i]) ^ slot_hash(sb + i, top2.w[i]);
                    // lp + factor + lp1 + lp2, left to right (src/_msnc_density.py:353)
                    ok = emit(src, hh, st.up_slot[0], top1, st.up_slot[1], top2, add_rn(add_rn(add_rn(lp, factor), lp1), lp2));
                    if (!ok || sub.zero()) break;
                    sub.prev_subset(at_v);
                }
            } else {
                double blp = 0.0;
                if (at_v.count() > 1) {
                    ranges();
                    branch_timed<NW>(at_v, blp, st.tau_low, st.tau_high[0], !is_root, inv_theta0, l2t0, ev_t, ev_n, j_lo, j_hi0);
                }
                if (is_root) root_slot = st.up_slot[0];
                const uint32_t sa = static_cast<uint32_t>(st.up_slot[0]) * NW;
                for (int i = 0; i < NW; ++i) h ^= slot_hash(sa + i, at_v.w[i]);
                ok = emit(src, h, st.up_slot[0], at_v, -1, at_v, add_rn(lp, blp));
            }
            if (!ok) {
                *status = JOB_OVERFLOW;
                return nan("");
            }
        }
        uint32_t t = cur_m; cur_m = new_m; new_m = t;
        t = cur_h; cur_h = new_h; new_h = t;
        t = cur_lp; cur_lp = new_lp; new_lp = t;
        n_cur = n_new;
    }

    // the one entry whose root slot holds exactly the gene-tree root (src/_msnc_density.py:374-389)
    for (int k = 0; k < n_cur; ++k) {
        const uint32_t e = cur_m + static_cast<uint32_t>(k) * entry + static_cast<uint32_t>(root_slot) * NW;
        bool hit = true;
        for (int i = 0; i < NW; ++i) {
            const uint64_t want = ((root_bit >> 6) == i) ? (uint64_t{1} << (root_bit & 63)) : 0;
            hit = hit && (PNB_M(e + i) == want);
        }
        if (hit) {
            const double score = PNB_LP(cur_lp + k);
            return (score <= LOG_FLOOR + 1.0) ? -INFINITY : score;   // msnc_log_density_prebuilt :488-490
        }
    }
    return -INFINITY;
#undef PNB_M
#undef PNB_LP
}

// ------------------------------------------------------------------------------------------
// Host side: compile the network, prepare gene trees.
// ------------------------------------------------------------------------------------------
struct NetProgram {
    std::vector<Step> steps;
    std::vector<int32_t> down_slot;
    int W = 1;                      // slots per frontier entry
    int n_species = 0;
    std::vector<int32_t> step_edge;           // [n_steps][2] up edge ids of each step (-1: none / root branch)
    std::vector<int32_t> retic_step;          // steps that are reticulations
    std::vector<int32_t> down_step;           // parallel to down_slot: the step whose node sits below that edge
    std::vector<int32_t> level;               // per step: 0 for a leaf, else 1 + the largest level among its children
    std::vector<std::vector<int32_t>> species_below;   // per reticulation step: species ids below it
};

// Returns "" or an error text.  edge_gamma: NaN = absent (src/_msnc_density.py:311-320 fills it in).
inline std::string compile_network(int n_nodes, int n_edges, const int32_t* edge_src, const int32_t* edge_dst,
                                   const double* edge_gamma, const double* height, const int32_t* node_species,
                                   int n_species, NetProgram& out) {
    if (n_nodes <= 0) return "empty species network";
    std::vector<std::vector<int>> down(n_nodes), up(n_nodes);
    for (int e = 0; e < n_edges; ++e) {
        const int s = edge_src[e], d = edge_dst[e];
        if (s < 0 || s >= n_nodes || d < 0 || d >= n_nodes || s == d) return "edge endpoint out of range";
        down[s].push_back(e);
        up[d].push_back(e);
    }
    int root = -1;
    for (int v = 0; v < n_nodes; ++v) {
        if (up[v].empty()) {
            if (root >= 0) return "species network has more than one root";
            root = v;
        }
        if (up[v].size() > 2) return "a network node has more than two parents";
        if (!(height[v] == height[v])) return "node height is NaN";
        if (down[v].empty()) {
            if (node_species[v] < -1 || node_species[v] >= n_species) return "leaf species id out of range";
        }
    }
    if (root < 0) return "species network has no root (cycle)";
    // subtree "width" for the visiting order: leaves under a node, counted with multiplicity
    std::vector<int> state(n_nodes, 0), order;
    std::vector<int64_t> weight(n_nodes, 0);
    order.reserve(n_nodes);
    {   // iterative DFS, children-first; cycle check through the on-stack state
        std::vector<std::pair<int, int>> stack;
        // pass 1: weights (memoised over the DAG)
        stack.push_back({root, 0});
        state[root] = 1;
        while (!stack.empty()) {
            auto& top = stack.back();
            const int v = top.first;
            if (top.second < static_cast<int>(down[v].size())) {
                const int c = edge_dst[down[v][top.second++]];
                if (state[c] == 1) return "species network has a cycle";
                if (state[c] == 0) {
                    state[c] = 1;
                    stack.push_back({c, 0});
                }
            } else {
                int64_t w = down[v].empty() ? 1 : 0;
                for (int e : down[v]) w += weight[edge_dst[e]];
                weight[v] = w;
                state[v] = 2;
                stack.pop_back();
            }
        }
        for (int v = 0; v < n_nodes; ++v)
            if (state[v] != 2) return "species network is not connected to its root";
        // pass 2: emit order, widest child first
        std::fill(state.begin(), state.end(), 0);
        std::vector<std::vector<int>> kids(n_nodes);
        for (int v = 0; v < n_nodes; ++v) {
            for (int e : down[v]) kids[v].push_back(edge_dst[e]);
            std::stable_sort(kids[v].begin(), kids[v].end(), [&](int a, int b) { return weight[a] > weight[b]; });
        }
        stack.push_back({root, 0});
        state[root] = 1;
        while (!stack.empty()) {
            auto& top = stack.back();
            const int v = top.first;
            if (top.second < static_cast<int>(kids[v].size())) {
                const int c = kids[v][top.second++];
                if (state[c] == 0) {
                    state[c] = 1;
                    stack.push_back({c, 0});
                }
            } else {
                order.push_back(v);
                stack.pop_back();
            }
        }
    }
    // slots by liveness
    out = NetProgram();
    out.n_species = n_species;
    std::vector<int> slot_of_edge(n_edges, -1), free_slots, step_of_node(n_nodes, -1);
    int n_slots = 0;
    auto take = [&]() {
        if (!free_slots.empty()) {
            const int s = *std::min_element(free_slots.begin(), free_slots.end());
            free_slots.erase(std::find(free_slots.begin(), free_slots.end(), s));
            return s;
        }
        return n_slots++;
    };
    std::vector<std::vector<int32_t>> below(n_nodes);   // species ids under each node (sorted, unique)
    for (int v : order) {
        Step st;
        memset(&st, 0, sizeof(st));
        st.species = -1;
        st.up_slot[0] = st.up_slot[1] = -1;
        st.tau_low = height[v];
        st.down_off = static_cast<int32_t>(out.down_slot.size());
        if (down[v].empty()) {
            st.kind |= KIND_LEAF;
            st.species = node_species[v];
            if (st.species >= 0) below[v].push_back(st.species);
        } else {
            for (int e : down[v]) {
                if (slot_of_edge[e] < 0) return "internal error: edge closed before it was opened";
                out.down_slot.push_back(slot_of_edge[e]);
                const int c = edge_dst[e];
                out.down_step.push_back(step_of_node[c]);
                below[v].insert(below[v].end(), below[c].begin(), below[c].end());
            }
            std::sort(below[v].begin(), below[v].end());
            below[v].erase(std::unique(below[v].begin(), below[v].end()), below[v].end());
            st.n_down = static_cast<int32_t>(down[v].size());
        }
        // the slots closed here become free only after the up slots are taken: an entry is copied,
        // its down slots are zeroed, then the up slots are written -- reusing a down slot would be
        // fine too, but keeping them apart makes the program easier to read in tests
        if (up[v].empty()) {
            st.kind |= KIND_ROOT;
            st.up_slot[0] = take();
        } else if (up[v].size() == 2) {
            st.kind |= KIND_RETIC;
            const int e1 = up[v][0], e2 = up[v][1];
            double g1 = edge_gamma[e1], g2 = edge_gamma[e2];
            const bool n1 = !(g1 == g1), n2 = !(g2 == g2);
            if (n1 && n2) g1 = g2 = 0.5;
            else if (n1) g1 = std::max(0.0, 1.0 - g2);
            else if (n2) g2 = std::max(0.0, 1.0 - g1);
            st.log_g[0] = g1 > 0.0 ? log(g1) : LOG_FLOOR;
            st.log_g[1] = g2 > 0.0 ? log(g2) : LOG_FLOOR;
            st.tau_high[0] = height[edge_src[e1]];
            st.tau_high[1] = height[edge_src[e2]];
            st.up_slot[0] = slot_of_edge[e1] = take();
            st.up_slot[1] = slot_of_edge[e2] = take();
            out.retic_step.push_back(static_cast<int32_t>(out.steps.size()));
            out.species_below.push_back(below[v]);
        } else {
            const int e = up[v][0];
            st.tau_high[0] = height[edge_src[e]];
            st.up_slot[0] = slot_of_edge[e] = take();
        }
        for (int e : down[v]) free_slots.push_back(slot_of_edge[e]);
        out.step_edge.push_back(up[v].empty() ? -1 : up[v][0]);
        out.step_edge.push_back(up[v].size() == 2 ? up[v][1] : -1);
        step_of_node[v] = static_cast<int>(out.steps.size());
        int lvl = 0;
        for (int e : down[v]) lvl = std::max(lvl, out.level[step_of_node[edge_dst[e]]] + 1);
        out.level.push_back(lvl);
        out.steps.push_back(st);
    }
    out.W = std::max(1, n_slots);
    return "";
}

// One gene tree -> sorted coalescence events, species lineage sets, root bit.
// (_gene_coalescence_events src/_msnc_density.py:142-169; _node_height src/GraphUtils.py:786-819;
// species_to_bits :262-266.)  Returns nullptr or an error text.  No heap use: at 2*10^4 candidate
// trees per call this runs ~10^5 times per millisecond budget.
struct TreeScratch {   // kept for callers that hold one per worker; the work arrays live on the stack
    int unused = 0;
};

inline const char* prepare_tree(int n, const int32_t* parent, const double* blen, const int32_t* node_species,
                                int n_species, int NW, TreeScratch&, double* ev_t, uint32_t* ev_n, int* n_ev_out,
                                uint64_t* species_mask /* [n_species][NW], zeroed here */, int* root_bit_out) {
    constexpr int MAXN = 64 * MAX_NW;
    if (n > 64 * NW || n > MAXN) return "gene tree has more nodes than the lineage-set width";
    // written branch-free where the data decides (which child comes first, which child is taller): on random
    // trees those branches mispredict every other node and were 3/4 of this function's time
    uint8_t n_kids[MAXN], kid[MAXN][2], remaining[MAXN], queue[MAXN], order[MAXN];
    double h[MAXN];
    int root = -1;
    bool children_first = true;    // parent[i] > i for every non-root node: one forward pass gives the heights
    for (int i = 0; i < n; ++i) {
        n_kids[i] = 0;
        h[i] = 0.0;
    }
    for (int i = 0; i < n; ++i) {
        const int p = parent[i];
        if (p < 0) {
            if (root < 0) root = i;   // the reference takes the first parentless node as root
            continue;
        }
        if (p >= n || p == i) return "gene-tree parent index out of range";
        children_first = children_first & (p > i);
        const unsigned c = n_kids[p];
        kid[p][c < 1u ? c : 1u] = static_cast<uint8_t>(i);    // a third child overwrites [1]; such a node emits no event
        n_kids[p] = static_cast<uint8_t>(c + (c < 255u));
    }
    if (root < 0) return "gene tree has no root";
    for (size_t i = 0; i < static_cast<size_t>(n_species) * NW; ++i) species_mask[i] = 0;
    for (int i = 0; i < n; ++i)
        if (n_kids[i] == 0) {
            const int sp = node_species[i];
            if (sp >= n_species) return "gene-tree leaf species id out of range";
            if (sp >= 0) species_mask[static_cast<size_t>(sp) * NW + (i >> 6)] |= uint64_t{1} << (i & 63);
        }
    if (children_first) {
        for (int v = 0; v < n; ++v) {
            const int p = parent[v];
            if (p < 0) continue;
            const double len = blen[v];
            const double cand = h[v] + ((len == len) ? len : 0.0);   // None counts as 0
            h[p] = (cand > h[p]) ? cand : h[p];
        }
    } else {   // children before parents by Kahn's algorithm
        int tail = 0;
        for (int i = 0; i < n; ++i) {
            remaining[i] = n_kids[i];
            if (n_kids[i] == 0) queue[tail++] = static_cast<uint8_t>(i);
        }
        int head = 0;
        while (head < tail) {
            const int v = queue[head++];
            const int p = parent[v];
            if (p < 0) continue;
            const double len = blen[v];
            const double cand = h[v] + ((len == len) ? len : 0.0);
            h[p] = (cand > h[p]) ? cand : h[p];
            if (--remaining[p] == 0) queue[tail++] = static_cast<uint8_t>(p);
        }
        if (tail != n) return "gene tree has a cycle";
    }
    // events: binary nodes in node order, stably sorted by time (insertion keeps equal times in node order)
    int k = 0;
    for (int v = 0; v < n; ++v) {
        if (n_kids[v] != 2) continue;
        int pos = k;
        while (pos > 0 && h[order[pos - 1]] > h[v]) {
            order[pos] = order[pos - 1];
            --pos;
        }
        order[pos] = static_cast<uint8_t>(v);
        ++k;
    }
    for (int e = 0; e < k; ++e) {
        const int v = order[e];
        ev_t[e] = h[v];
        ev_n[e] = static_cast<uint32_t>(v) | (static_cast<uint32_t>(kid[v][0]) << 8) | (static_cast<uint32_t>(kid[v][1]) << 16);
    }
    *n_ev_out = k;
    *root_bit_out = root;
    return nullptr;
}

// Per call: 1/theta and log(2/theta) of the branch(es) above every step's node.  edge_theta[e] / root_theta
// override the default where they are not NaN (resolve_branch_theta, src/_units.py:90-125; the caller has
// turned label keys into edge ids).  The logarithm is taken here, by the host's libm, as the reference does.
inline void fill_step_theta(const NetProgram& prog, double theta, const double* edge_theta, double root_theta,
                            double* out /* [n_steps][4] */) {
    const size_t n = prog.steps.size();
    for (size_t s = 0; s < n; ++s)
        for (int k = 0; k < 2; ++k) {
            const int e = prog.step_edge[2 * s + k];
            double th = theta;
            if (e >= 0) {
                if (edge_theta && edge_theta[e] == edge_theta[e]) th = edge_theta[e];
            } else if (k == 0 && (prog.steps[s].kind & KIND_ROOT)) {
                if (root_theta == root_theta) th = root_theta;
            }
            const double inv = 1.0 / th;
            out[4 * s + 2 * k] = inv;
            out[4 * s + 2 * k + 1] = log(2.0 * inv);
        }
}

// Upper bound on the frontier size of any job whose species s carries at most alleles[s] gene leaves:
// every reticulation multiplies the frontier by at most 2^(gene leaves below it).
inline double frontier_bound(const NetProgram& prog, const int32_t* alleles) {
    double bound = 1.0;
    for (const auto& sp : prog.species_below) {
        int64_t m = 0;
        for (int32_t s : sp) m += alleles[s];
        bound *= ldexp(1.0, static_cast<int>(std::min<int64_t>(m, 60)));
        if (bound > 1e18) return 1e18;
    }
    return bound;
}
}  // namespace msnc
}  // namespace pnb
