// pnb_msnc.cu -- K-m: batched timed MSNC density log P(g_j | Psi) on the GPU.
//
// Replaces msnc_log_density_prebuilt (src/_msnc_density.py:465-491) called once per locus by
// _SeqLikelihoodEngine.log_likelihood (src/_mcmc_seq.py:750-760) and once per candidate by the
// coupled moves (src/_mcmc_seq.py:1935-1957).  The network is compiled once (pnb_msnc_net_create);
// an evaluation prepares every gene tree on the host workers (heights, sorted coalescence events,
// species lineage sets -- integer / comparison work, O(n log n) per tree), uploads one packed
// buffer and runs ONE kernel, one thread per gene tree (pnb_msnc_core.hpp has the DP).
#include "pnb_host.hpp"
#include "pnb_msnc_core.hpp"

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <functional>

using namespace pnb;
using namespace pnb::msnc;

struct pnb_msnc_net {
    pnb_ctx* ctx = nullptr;
    NetProgram prog;
    Step* d_steps = nullptr;
    int32_t* d_down = nullptr;        // [down_slot (n_down) | down_step (n_down) | level (n_steps)]
    size_t down_bytes = 0;
    int64_t max_frontier = 1 << 16;   // entries per job the scratch may be sized for
    int32_t n_edges = 0;
    std::vector<double> edge_theta;   // per-branch theta overrides (NaN = the call's theta); empty = none
    double root_theta = NAN;
};

namespace {

struct MsncLaunch {
    const Step* steps;
    const int32_t* down_slot;
    int32_t n_steps, W, cap, n_species;
    int64_t n_jobs;
    const int64_t* ev_off;      // [n_jobs + 1] event offsets
    const double* ev_t;
    const uint32_t* ev_n;
    const uint64_t* species_mask;   // [n_jobs][n_species][NW]
    const int32_t* root_bit;        // [n_jobs]; -1 = empty gene tree
    uint64_t* fr_mask;              // [n_jobs][(2 cap + 1) W NW]
    double* fr_lp;                  // [n_jobs][2 cap]
    const double* step_theta;       // [n_steps][4]: 1/theta, log(2/theta) of the branch(es) above each step's node
    double* out;
    int32_t* status;
    int32_t frontier_in_smem;       // 1: the frontier of every thread lives in shared memory, interleaved
    int32_t n_down;                 // entries of down_slot
    int32_t program_words;          // 8-byte words of shared memory holding the staged program (frontiers follow)
    int32_t tree_shaped;            // 1: no reticulation -> msnc_dp_tree
    int32_t stage_words;            // > 0: every thread copies its gene tree's events and species masks into this many
                                    // 8-byte words of shared memory (after the frontiers) before the DP
    int32_t stage_ev_cap;           // events per tree the staging area holds
    int32_t warp_tsize;             // warp kernel: slots of the duplicate-detection table (power of two >= 2 cap)
    int64_t warp_job_words;         // warp kernel: 8-byte words of frontier storage per job (msnc_warp_words)
};

// ONE gene tree (a gene-tree move of the chain): its packed input travels in the kernel parameters -- the root bit, the
// event list, the species lineage sets and the per-step theta terms would otherwise be three dependent reads of
// freshly copied (or host-mapped) memory before the DP can start.
constexpr int MS_MAX_EV = 64, MS_MAX_SP = 32, MS_MAX_THETA = 252;
struct MsncSmall {
    int32_t used, n_ev, root_bit, pad;
    double ev_t[MS_MAX_EV];
    uint64_t sp[MS_MAX_SP];
    double theta[MS_MAX_THETA];
    uint32_t ev_n[MS_MAX_EV];
};

// ------------------------------------------------------------------------------------------------------------
// One WARP per gene tree (reticulate networks).  The DP of pnb_msnc_core.hpp::msnc_dp is a loop over network
// nodes; at each node every frontier entry -- and, at a reticulation, every subset of the lineages that reach it
// -- produces one candidate entry, independently of the others.  Here the candidates of a step are spread over
// the 32 lanes, then merged exactly as the serial DP merges them:
//   A  (reticulation) lanes count the subsets of each entry, a warp scan turns the counts into candidate offsets;
//   B  lane t builds candidate t = (entry k, r-th subset in decreasing order): branch factor(s), new lineage
//      sets, incremental hash, log-weight -- written as entry t of the other frontier buffer;
//   C  equal candidates are found through an open-addressing table keyed by the entry hash (atomicCAS claims a
//      slot for a class of equal entries, atomicMin leaves the FIRST candidate of the class in it);
//   D  the log-weights of later duplicates are folded into the first one in candidate order (one leader lane per
//      class, log-sum-exp in the order the serial DP would have met them), then the first occurrences are
//      compacted in place, in order.
// Entry order, merge order and every floating-point operation are those of msnc_dp, so the two kernels return the
// same bits (tests/test_gpu_msnc.py::test_warp_and_thread_kernels_agree); what changes is that a frontier of T
// candidates costs T / 32 rounds instead of T, and that a batch fills the SMs with 32x more threads.
// Frontier layout (per job, word-major so that consecutive lanes touch consecutive words):
//   mask[2][W NW][cap] | hash[2][cap] | lp[2][cap] | idx int32[cap + 1] | dup int32[cap] | table int32[tsize]
// ------------------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t msnc_warp_words(int W, int NW, int cap, int tsize) {
    const size_t c = static_cast<size_t>(cap);
    return 2 * c * (static_cast<size_t>(W) * NW + 2) + ((c + 2) & ~size_t{1}) / 2 + ((c + 1) & ~size_t{1}) / 2 +
           static_cast<size_t>(tsize) / 2;
}

template <int NW, bool SMEM>
__device__ double msnc_dp_warp(const Step* steps, int n_steps, const int32_t* down_slot, int W, int cap, const double* ev_t,
                               const uint32_t* ev_n, int n_ev, const uint64_t* species_mask, int root_bit,
                               const double* step_theta, uint64_t* buf, int* status) {
    constexpr unsigned FULL = 0xffffffffu;
    __builtin_assume(__isShared(steps));
    __builtin_assume(__isShared(down_slot));
    __builtin_assume(__isShared(step_theta));
    if (SMEM) {   // everything the DP touches is in shared memory: 32-bit addresses, LDS / STS / ATOMS instead of generic accesses
        __builtin_assume(__isShared(buf));
        __builtin_assume(__isShared(ev_t));
        __builtin_assume(__isShared(ev_n));
        __builtin_assume(__isShared(species_mask));
    }
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t entry = static_cast<uint32_t>(W) * NW, ucap = static_cast<uint32_t>(cap);
    // the two frontiers are word offsets into `buf`, swapped after every step (offsets, not pointers: every access stays
    // visibly relative to the one base pointer, which is what lets the SMEM variant use shared-memory instructions)
    double* const bufd = reinterpret_cast<double*>(buf);
    uint32_t m_cur = 0, m_new = entry * ucap;
    uint32_t h_cur = 2 * entry * ucap, h_new = h_cur + ucap;
    uint32_t l_cur = h_new + ucap, l_new = l_cur + ucap;
    int32_t* idx = reinterpret_cast<int32_t*>(buf + (l_new + ucap));
    int32_t* dup = idx + ((ucap + 2) & ~1u);
    int32_t* table = dup + ((ucap + 1) & ~1u);

    if (lane == 0) {
        for (uint32_t i = 0; i < entry; ++i) buf[m_cur + (i) * ucap] = 0;
        buf[h_cur + 0] = 0;
        bufd[l_cur + 0] = 0.0;
    }
    __syncwarp();
    uint32_t n_cur = 1;
    int root_slot = 0;
    *status = JOB_OK;

    for (int s = 0; s < n_steps; ++s) {
        const Step& st = steps[s];
        const int kind = st.kind, species = st.species, n_down = st.n_down, down_off = st.down_off;
        const int up0 = st.up_slot[0], up1 = st.up_slot[1];
        const bool is_root = (kind & KIND_ROOT) != 0, is_retic = (kind & KIND_RETIC) != 0, is_leaf = (kind & KIND_LEAF) != 0;
        const double tau_low = st.tau_low, tau_h0 = st.tau_high[0], tau_h1 = st.tau_high[1];
        const double inv_theta0 = step_theta[4 * s], l2t0 = step_theta[4 * s + 1];
        const double inv_theta1 = step_theta[4 * s + 2], l2t1 = step_theta[4 * s + 3];
        // Short road 1 -- a leaf of the network entered by at most one gene lineage (half of all steps when every species
        // has one allele): no branch factor (exactly 1), nothing closed, nothing that could make two entries equal.  Every
        // entry gets the species' lineage set in the opened slot IN PLACE: no candidate buffer, no duplicate search, no swap.
        if (is_leaf && !is_retic) {
            Mask<NW> a;
            a.clear();
            if (species >= 0)
                for (int i = 0; i < NW; ++i) a.w[i] = species_mask[static_cast<size_t>(species) * NW + i];
            if (a.count() <= 1) {
                const uint32_t sa = static_cast<uint32_t>(up0) * NW;
                uint64_t dh = 0;
                for (int i = 0; i < NW; ++i) dh ^= slot_hash(sa + i, a.w[i]);
                for (uint32_t t = lane; t < n_cur; t += 32) {
                    for (int i = 0; i < NW; ++i) buf[m_cur + (sa + i) * ucap + t] = a.w[i];
                    buf[h_cur + t] ^= dh;
                }
                if (is_root) root_slot = up0;
                __syncwarp();
                continue;
            }
        }
        // events inside this node's parent branch(es): events are sorted by time, so "first index with t >= tau" is the
        // number of events with t < tau -- one ballot per 32 events instead of a binary search per lane
        int j_lo = 0, j_hi0 = 0, j_hi1 = 0;
        for (int base = 0; base < n_ev; base += 32) {
            const int e = base + lane;
            const double t = (e < n_ev) ? ev_t[e] : INFINITY;
            j_lo += __popc(__ballot_sync(FULL, t < tau_low));
            j_hi0 += __popc(__ballot_sync(FULL, t < tau_h0));
            j_hi1 += __popc(__ballot_sync(FULL, t < tau_h1));
        }
        if (is_root) j_hi0 = n_ev;
        if (!is_retic) j_hi1 = j_lo;

        // lineage set reaching the node in entry k; the closed slots leave the hash
        auto arriving = [&](uint32_t k, uint64_t& h) {
            Mask<NW> a;
            a.clear();
            if (is_leaf) {
                if (species >= 0)
                    for (int i = 0; i < NW; ++i) a.w[i] = species_mask[static_cast<size_t>(species) * NW + i];
            } else {
                for (int d = 0; d < n_down; ++d) {
                    const uint32_t slot = static_cast<uint32_t>(down_slot[down_off + d]) * NW;
                    for (int i = 0; i < NW; ++i) {
                        const uint64_t w = buf[m_cur + (slot + i) * ucap + k];
                        a.w[i] |= w;
                        h ^= slot_hash(slot + i, w);
                    }
                }
            }
            return a;
        };

        // Short road 2 -- ONE entry and no reticulation (the DP before the first reticulation, and wherever the frontier has
        // collapsed again): the entry is updated in place by lane 0, as msnc_dp_tree does it.
        if (!is_retic && n_cur == 1) {
            if (lane == 0) {
                uint64_t h = buf[h_cur];
                Mask<NW> ma = arriving(0, h);
                double blp = 0.0;
                if (ma.count() > 1) branch_timed<NW>(ma, blp, tau_low, tau_h0, !is_root, inv_theta0, l2t0, ev_t, ev_n, j_lo, j_hi0);
                for (int d = 0; d < n_down; ++d)
                    for (int i = 0; i < NW; ++i) buf[m_cur + (down_slot[down_off + d] * NW + i) * ucap] = 0;
                const uint32_t sa = static_cast<uint32_t>(up0) * NW;
                for (int i = 0; i < NW; ++i) {
                    buf[m_cur + (sa + i) * ucap] = ma.w[i];
                    h ^= slot_hash(sa + i, ma.w[i]);
                }
                buf[h_cur] = h;
                bufd[l_cur] = add_rn(bufd[l_cur], blp);
            }
            if (is_root) root_slot = up0;
            __syncwarp();
            continue;
        }
        // ---- A: candidate offsets -------------------------------------------------------------------------
        uint32_t T = n_cur;
        if (is_retic) {
            uint32_t running = 0;
            for (uint32_t base = 0; base < n_cur; base += 32) {
                const uint32_t k = base + lane;
                uint32_t n = 0;
                if (k < n_cur) {
                    uint64_t unused = 0;
                    const int c = arriving(k, unused).count();
                    n = (c >= 20) ? ucap + 1 : min(1u << c, ucap + 1);
                }
                uint32_t incl = n;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                if (k < n_cur) idx[k] = static_cast<int32_t>(min(running + incl - n, ucap + 1));
                running = min(running + __shfl_sync(FULL, incl, 31), ucap + 1);
            }
            if (lane == 0) idx[n_cur] = static_cast<int32_t>(running);
            T = running;
        }
        if (T > ucap) {
            *status = JOB_OVERFLOW;
            return nan("");
        }
        __syncwarp();

        // ---- B: one candidate per lane ----------------------------------------------------------------------
        for (uint32_t t = lane; t < T; t += 32) {
            uint32_t k = t, r = 0;
            if (is_retic) {   // the entry whose subsets include candidate t: largest k with idx[k] <= t
                uint32_t lo = 0, hi = n_cur - 1;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi + 1) >> 1;
                    if (static_cast<uint32_t>(idx[mid]) <= t) lo = mid;
                    else hi = mid - 1;
                }
                k = lo;
                r = t - static_cast<uint32_t>(idx[k]);
            }
            const double lp = bufd[l_cur + k];
            uint64_t h = buf[h_cur + k];
            Mask<NW> ma = arriving(k, h), mb;
            mb.clear();
            double lpn;
            if (is_retic) {
                const int n_total = ma.count();
                uint32_t v = ((1u << n_total) - 1u) - r;   // subsets in decreasing order: the r-th one has rank 2^n - 1 - r
                Mask<NW> sub;
                sub.clear();
                for (int w = 0; w < NW; ++w) {
                    uint64_t m = ma.w[w];
                    while (m) {
                        const uint64_t low = m & (~m + 1);
                        if (v & 1u) sub.w[w] |= low;
                        v >>= 1;
                        m ^= low;
                    }
                }
                const int ks = sub.count();
                const double factor = add_rn(mul_rn(static_cast<double>(ks), st.log_g[0]),
                                             mul_rn(static_cast<double>(n_total - ks), st.log_g[1]));
                for (int i = 0; i < NW; ++i) mb.w[i] = ma.w[i] ^ sub.w[i];
                ma = sub;
                double lp1 = 0.0, lp2 = 0.0;
                if (ks > 1) branch_timed<NW>(ma, lp1, tau_low, tau_h0, true, inv_theta0, l2t0, ev_t, ev_n, j_lo, j_hi0);
                if (n_total - ks > 1) branch_timed<NW>(mb, lp2, tau_low, tau_h1, true, inv_theta1, l2t1, ev_t, ev_n, j_lo, j_hi1);
                const uint32_t sa = static_cast<uint32_t>(up0) * NW, sb = static_cast<uint32_t>(up1) * NW;
                for (int i = 0; i < NW; ++i) h ^= slot_hash(sa + i, ma.w[i]) ^ slot_hash(sb + i, mb.w[i]);
                lpn = add_rn(add_rn(add_rn(lp, factor), lp1), lp2);   // left to right (src/_msnc_density.py:353)
            } else {
                double blp = 0.0;
                if (ma.count() > 1) branch_timed<NW>(ma, blp, tau_low, tau_h0, !is_root, inv_theta0, l2t0, ev_t, ev_n, j_lo, j_hi0);
                const uint32_t sa = static_cast<uint32_t>(up0) * NW;
                for (int i = 0; i < NW; ++i) h ^= slot_hash(sa + i, ma.w[i]);
                lpn = add_rn(lp, blp);
            }
            for (uint32_t i = 0; i < entry; ++i) buf[m_new + (i) * ucap + t] = buf[m_cur + (i) * ucap + k];
            for (int d = 0; d < n_down; ++d)
                for (int i = 0; i < NW; ++i) buf[m_new + (down_slot[down_off + d] * NW + i) * ucap + t] = 0;
            for (int i = 0; i < NW; ++i) buf[m_new + (up0 * NW + i) * ucap + t] = ma.w[i];
            if (is_retic)
                for (int i = 0; i < NW; ++i) buf[m_new + (up1 * NW + i) * ucap + t] = mb.w[i];
            buf[h_new + t] = h;
            bufd[l_new + t] = lpn;
        }
        if (is_root) root_slot = up0;
        __syncwarp();

        uint32_t n_new = T;
        if (T > 1) {
            // ---- C: first occurrence of every candidate ---------------------------------------------------
            uint32_t ts = 64;
            while (ts < 2 * T) ts <<= 1;
            for (uint32_t i = lane; i < ts; i += 32) table[i] = -1;
            __syncwarp();
            for (uint32_t t = lane; t < T; t += 32) {
                const uint64_t h = buf[h_new + t];
                uint32_t slot = static_cast<uint32_t>(h ^ (h >> 32)) & (ts - 1);
                for (;;) {
                    const int old = atomicCAS(&table[slot], -1, static_cast<int>(t));
                    if (old == -1) break;   // first of its class to arrive: the slot now belongs to the class
                    bool same = buf[h_new + old] == h;
                    for (uint32_t i = 0; same && i < entry; ++i)
                        same = buf[m_new + (i) * ucap + old] == buf[m_new + (i) * ucap + t];
                    if (same) {
                        atomicMin(&table[slot], static_cast<int>(t));
                        break;
                    }
                    slot = (slot + 1) & (ts - 1);
                }
                idx[t] = static_cast<int32_t>(slot);
            }
            __syncwarp();
            for (uint32_t t = lane; t < T; t += 32) idx[t] = table[idx[t]];
            __syncwarp();
            // ---- D: merge duplicates in candidate order, compact the first occurrences -----------------------
            uint32_t nd = 0;
            for (uint32_t base = 0; base < T; base += 32) {
                const uint32_t t = base + lane;
                const bool isd = t < T && idx[t] != static_cast<int32_t>(t);
                const unsigned b = __ballot_sync(FULL, isd);
                if (isd) dup[nd + __popc(b & lt)] = static_cast<int32_t>(t);
                nd += __popc(b);
            }
            if (nd > 0) {
                __syncwarp();
                for (uint32_t base = 0; base < nd; base += 32) {
                    const int d = (base + lane < nd) ? dup[base + lane] : -1;
                    const int u = (d >= 0) ? idx[d] : -1 - lane;
                    const unsigned grp = __match_any_sync(FULL, u);
                    if (d >= 0 && lane == __ffs(grp) - 1) {
                        double acc = bufd[l_new + u];
                        for (unsigned m = grp; m; m &= m - 1) acc = log_add(acc, bufd[l_new + dup[base + __ffs(m) - 1]]);
                        bufd[l_new + u] = acc;
                    }
                    __syncwarp();
                }
                uint32_t running = 0;
                for (uint32_t base = 0; base < T; base += 32) {
                    const uint32_t t = base + lane;
                    const bool isu = t < T && idx[t] == static_cast<int32_t>(t);
                    const unsigned b = __ballot_sync(FULL, isu);
                    const uint32_t pos = running + __popc(b & lt);
                    {
                        const uint64_t v = isu ? buf[h_new + t] : 0;
                        const double l = isu ? bufd[l_new + t] : 0.0;
                        __syncwarp();
                        if (isu) { buf[h_new + pos] = v; bufd[l_new + pos] = l; }
                    }
                    for (uint32_t i = 0; i < entry; ++i) {
                        uint64_t* row = buf + (m_new + i * ucap);
                        const uint64_t v = isu ? row[t] : 0;
                        __syncwarp();
                        if (isu) row[pos] = v;
                    }
                    __syncwarp();
                    running += __popc(b);
                }
                n_new = running;
            }
        }
        uint32_t tm = m_cur; m_cur = m_new; m_new = tm;
        tm = h_cur; h_cur = h_new; h_new = tm;
        tm = l_cur; l_cur = l_new; l_new = tm;
        n_cur = n_new;
        __syncwarp();
    }

    // the one entry whose root slot holds exactly the gene-tree root (src/_msnc_density.py:374-389)
    for (uint32_t base = 0; base < n_cur; base += 32) {
        const uint32_t k = base + lane;
        bool hit = k < n_cur;
        if (hit)
            for (int i = 0; i < NW; ++i) {
                const uint64_t want = ((root_bit >> 6) == i) ? (uint64_t{1} << (root_bit & 63)) : 0;
                hit = hit && (buf[m_cur + (root_slot * NW + i) * ucap + k] == want);
            }
        const unsigned b = __ballot_sync(FULL, hit);
        if (b) {
            const double score = bufd[l_cur + base + __ffs(b) - 1];
            return (score <= LOG_FLOOR + 1.0) ? -INFINITY : score;   // msnc_log_density_prebuilt :488-490
        }
    }
    return -INFINITY;
}

template <int NW, bool SMEM>
__global__ void __launch_bounds__(128) pnb_msnc_warp_kernel(const __grid_constant__ MsncLaunch L, const __grid_constant__ MsncSmall S) {
    // Dynamic shared memory: [step program | step thetas | down slots] [per warp: frontier storage, if it fits]
    // [per warp: the gene tree's events and species lineage sets, if they fit]
    extern __shared__ uint64_t pnb_msnc_smem[];
    Step* s_steps = reinterpret_cast<Step*>(pnb_msnc_smem);
    double* s_theta = reinterpret_cast<double*>(s_steps + L.n_steps);
    int32_t* s_down = reinterpret_cast<int32_t*>(s_theta + 4 * static_cast<size_t>(L.n_steps));
    {
        const uint64_t* src = reinterpret_cast<const uint64_t*>(L.steps);
        const int words = L.n_steps * static_cast<int>(sizeof(Step) / 8);
        for (int i = threadIdx.x; i < words; i += blockDim.x) pnb_msnc_smem[i] = src[i];
        for (int i = threadIdx.x; i < 4 * L.n_steps; i += blockDim.x) s_theta[i] = S.used ? S.theta[i] : L.step_theta[i];
        for (int i = threadIdx.x; i < L.n_down; i += blockDim.x) s_down[i] = L.down_slot[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int64_t j = static_cast<int64_t>(blockIdx.x) * wpb + wib;
    if (j >= L.n_jobs) return;
    const int root_bit = S.used ? S.root_bit : L.root_bit[j];
    if (root_bit < 0) {   // empty gene tree: density 1 (src/_msnc_density.py:222-223)
        if (lane == 0) { L.out[j] = 0.0; L.status[j] = JOB_OK; }
        return;
    }
    const size_t job_words = static_cast<size_t>(L.warp_job_words);
    uint64_t* buf = L.frontier_in_smem ? pnb_msnc_smem + L.program_words + wib * job_words
                                       : L.fr_mask + static_cast<size_t>(j) * job_words;
    const int64_t e0 = S.used ? 0 : L.ev_off[j];
    const int n_ev = S.used ? S.n_ev : static_cast<int>(L.ev_off[j + 1] - e0);
    const double* ev_t = S.used ? S.ev_t : L.ev_t + e0;
    const uint32_t* ev_n = S.used ? S.ev_n : L.ev_n + e0;
    const uint64_t* sp_mask = S.used ? S.sp : L.species_mask + static_cast<size_t>(j) * L.n_species * NW;
    if (L.stage_words > 0 && n_ev <= L.stage_ev_cap) {
        uint64_t* reg = pnb_msnc_smem + L.program_words + (L.frontier_in_smem ? job_words * wpb : 0) +
                        static_cast<size_t>(wib) * L.stage_words;
        double* stt = reinterpret_cast<double*>(reg);
        uint32_t* sn = reinterpret_cast<uint32_t*>(reg + L.stage_ev_cap);
        uint64_t* ss = reg + L.stage_ev_cap + (L.stage_ev_cap + 1) / 2;
        for (int i = lane; i < n_ev; i += 32) { stt[i] = ev_t[i]; sn[i] = ev_n[i]; }
        for (int i = lane; i < L.n_species * NW; i += 32) ss[i] = sp_mask[i];
        __syncwarp();
        ev_t = stt; ev_n = sn; sp_mask = ss;
    }
    int status;
    const double v = msnc_dp_warp<NW, SMEM>(s_steps, L.n_steps, s_down, L.W, L.cap, ev_t, ev_n, n_ev, sp_mask, root_bit, s_theta,
                                            buf, &status);
    if (lane == 0) { L.out[j] = v; L.status[j] = status; }
}

// ------------------------------------------------------------------------------------------------------------
// One warp per gene tree on a species TREE (the MSC case), for small batches -- a gene-tree move scores one tree, and
// one thread walking the 2 n - 1 network nodes one after the other is 20+ us of dependent instructions.  The frontier
// of a tree-shaped network is a single entry, so the nodes of different subtrees do not interact: lane s takes node
// (step) s, the warp advances LEVEL by level (leaves = 0, a node one above its highest child), every lane of a level
// ORs the lineage sets its children left, runs the same branch factor as msnc_dp_tree and leaves its own set and
// factor in shared memory.  The factors are then added in step order, exactly the additions msnc_dp_tree makes, so the
// result has the same bits.  Depth ~2 log2(n) rounds instead of 2 n - 1.
// Shared memory per warp: out[n_steps][NW] lineage sets | blp[n_steps] factors | counted[n_steps] ("branch entered by < 2
// lineages" factors are flagged and skipped in the sum as msnc_dp_tree skips them) | staged events and species sets.
// ------------------------------------------------------------------------------------------------------------
template <int NW>
__global__ void __launch_bounds__(128) pnb_msnc_tree_warp_kernel(const __grid_constant__ MsncLaunch L, const __grid_constant__ MsncSmall S) {
    extern __shared__ uint64_t pnb_msnc_smem[];
    Step* s_steps = reinterpret_cast<Step*>(pnb_msnc_smem);
    double* s_theta = reinterpret_cast<double*>(s_steps + L.n_steps);
    int32_t* s_down = reinterpret_cast<int32_t*>(s_theta + 4 * static_cast<size_t>(L.n_steps));   // down_slot | down_step | level
    {
        const uint64_t* src = reinterpret_cast<const uint64_t*>(L.steps);
        const int words = L.n_steps * static_cast<int>(sizeof(Step) / 8);
        for (int i = threadIdx.x; i < words; i += blockDim.x) pnb_msnc_smem[i] = src[i];
        for (int i = threadIdx.x; i < 4 * L.n_steps; i += blockDim.x) s_theta[i] = S.used ? S.theta[i] : L.step_theta[i];
        for (int i = threadIdx.x; i < 2 * L.n_down + L.n_steps; i += blockDim.x) s_down[i] = L.down_slot[i];
    }
    __syncthreads();
    const int32_t* s_down_step = s_down + L.n_down;
    const int32_t* s_level = s_down + 2 * L.n_down;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int64_t j = static_cast<int64_t>(blockIdx.x) * wpb + wib;
    if (j >= L.n_jobs) return;
    const int root_bit = S.used ? S.root_bit : L.root_bit[j];
    if (root_bit < 0) {   // empty gene tree: density 1 (src/_msnc_density.py:222-223)
        if (lane == 0) { L.out[j] = 0.0; L.status[j] = JOB_OK; }
        return;
    }
    const int n_steps = L.n_steps;
    uint64_t* reg = pnb_msnc_smem + L.program_words + static_cast<size_t>(wib) * L.warp_job_words;
    uint64_t* out = reg;                                                          // [n_steps][NW]
    double* blp = reinterpret_cast<double*>(reg + static_cast<size_t>(n_steps) * NW);   // [n_steps]
    int32_t* counted = reinterpret_cast<int32_t*>(reg + static_cast<size_t>(n_steps) * (NW + 1));   // [n_steps]: factor enters the sum
    uint64_t* stage = reg + static_cast<size_t>(n_steps) * (NW + 1) + (n_steps + 1) / 2;
    const int64_t e0 = S.used ? 0 : L.ev_off[j];
    const int n_ev = S.used ? S.n_ev : static_cast<int>(L.ev_off[j + 1] - e0);
    double* ev_t = reinterpret_cast<double*>(stage);
    uint32_t* ev_n = reinterpret_cast<uint32_t*>(stage + L.stage_ev_cap);
    uint64_t* sp_mask = stage + L.stage_ev_cap + (L.stage_ev_cap + 1) / 2;
    for (int i = lane; i < n_ev; i += 32) {
        ev_t[i] = S.used ? S.ev_t[i] : L.ev_t[e0 + i];
        ev_n[i] = S.used ? S.ev_n[i] : L.ev_n[e0 + i];
    }
    for (int i = lane; i < L.n_species * NW; i += 32)
        sp_mask[i] = S.used ? S.sp[i] : L.species_mask[static_cast<size_t>(j) * L.n_species * NW + i];
    int max_level = 0;
    for (int s = lane; s < n_steps; s += 32) max_level = max(max_level, s_level[s]);
    max_level = __reduce_max_sync(0xffffffffu, max_level);
    __syncwarp();
    for (int lvl = 0; lvl <= max_level; ++lvl) {
        for (int s = lane; s < n_steps; s += 32) {
            if (s_level[s] != lvl) continue;
            const Step& st = s_steps[s];
            Mask<NW> at_v;
            at_v.clear();
            if (st.kind & KIND_LEAF) {
                if (st.species >= 0)
                    for (int i = 0; i < NW; ++i) at_v.w[i] = sp_mask[static_cast<size_t>(st.species) * NW + i];
            } else {
                for (int d = 0; d < st.n_down; ++d) {
                    const int c = s_down_step[st.down_off + d];
                    for (int i = 0; i < NW; ++i) at_v.w[i] |= out[static_cast<size_t>(c) * NW + i];
                }
            }
            const bool is_root = (st.kind & KIND_ROOT) != 0;
            double b = 0.0;
            const bool two = at_v.count() > 1;
            if (two) {
                const int j_lo = first_event_at_or_after(ev_t, n_ev, st.tau_low);
                const int j_hi = is_root ? n_ev : first_event_at_or_after(ev_t, n_ev, st.tau_high[0]);
                branch_timed<NW>(at_v, b, st.tau_low, st.tau_high[0], !is_root, s_theta[4 * s], s_theta[4 * s + 1], ev_t, ev_n,
                                 j_lo, j_hi);
            }
            for (int i = 0; i < NW; ++i) out[static_cast<size_t>(s) * NW + i] = at_v.w[i];
            blp[s] = b;
            counted[s] = two ? 1 : 0;
        }
        __syncwarp();
    }
    if (lane == 0) {
        double lp = 0.0;
        for (int s = 0; s < n_steps; ++s) {
            if (counted[s]) lp = add_rn(lp, blp[s]);
        }
        // the last step is the root (children before parents): its lineage set must be exactly the gene-tree root
        bool hit = true;
        for (int i = 0; i < NW; ++i) {
            const uint64_t want = ((root_bit >> 6) == i) ? (uint64_t{1} << (root_bit & 63)) : 0;
            hit = hit && (out[static_cast<size_t>(n_steps - 1) * NW + i] == want);
        }
        L.out[j] = !hit ? -INFINITY : ((lp <= LOG_FLOOR + 1.0) ? -INFINITY : lp);
        L.status[j] = JOB_OK;
    }
}

template <int NW, bool SMEM>
__global__ void __launch_bounds__(128) pnb_msnc_kernel(const __grid_constant__ MsncLaunch L, const __grid_constant__ MsncSmall S) {
    // Dynamic shared memory: [step program | step thetas | down slots] [frontiers, if they fit].
    // Every thread walks the same program; staging it once per block turns ~3 dependent global loads per
    // step (L2 / DRAM latency each when the batch is a single gene tree) into shared-memory broadcasts.
    extern __shared__ uint64_t pnb_msnc_smem[];
    Step* s_steps = reinterpret_cast<Step*>(pnb_msnc_smem);
    double* s_theta = reinterpret_cast<double*>(s_steps + L.n_steps);
    int32_t* s_down = reinterpret_cast<int32_t*>(s_theta + 4 * static_cast<size_t>(L.n_steps));
    {
        const uint64_t* src = reinterpret_cast<const uint64_t*>(L.steps);
        const int words = L.n_steps * static_cast<int>(sizeof(Step) / 8);
        for (int i = threadIdx.x; i < words; i += blockDim.x) pnb_msnc_smem[i] = src[i];
        for (int i = threadIdx.x; i < 4 * L.n_steps; i += blockDim.x) s_theta[i] = S.used ? S.theta[i] : L.step_theta[i];
        for (int i = threadIdx.x; i < L.n_down; i += blockDim.x) s_down[i] = L.down_slot[i];
    }
    __syncthreads();
    const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= L.n_jobs) return;
    const int root_bit = S.used ? S.root_bit : L.root_bit[j];
    if (root_bit < 0) {   // empty gene tree: density 1 (src/_msnc_density.py:222-223)
        L.out[j] = 0.0;
        L.status[j] = JOB_OK;
        return;
    }
    const int64_t e0 = S.used ? 0 : L.ev_off[j];
    const size_t mask_words = msnc_mask_words(L.W, NW, L.cap);
    // Frontier storage.  A thread re-reads what it has just written (copy entry, merge slots, compare
    // for duplicates): in global memory every such read is an L2 round trip, in shared memory ~30 cycles.
    // Shared layout: word k of thread t at [k * blockDim.x + t] -- consecutive threads, consecutive banks.
    uint64_t* fr_mask;
    double* fr_lp;
    size_t stride;
    if (L.frontier_in_smem) {
        uint64_t* fr = pnb_msnc_smem + L.program_words;
        fr_mask = fr + threadIdx.x;
        fr_lp = reinterpret_cast<double*>(fr + mask_words * blockDim.x) + threadIdx.x;
        stride = blockDim.x;
    } else {
        fr_mask = L.fr_mask + static_cast<size_t>(j) * mask_words;
        fr_lp = L.fr_lp + static_cast<size_t>(j) * 2 * L.cap;
        stride = 1;
    }
    // The DP looks its events up again and again (two binary searches per step, a scan per entry and branch): from global
    // memory every probe is an L2 round trip, which is most of a single gene tree's latency.  When the block's shared
    // memory has room, each thread first copies its tree's events and species masks next to the frontiers.
    const int n_ev = S.used ? S.n_ev : static_cast<int>(L.ev_off[j + 1] - e0);
    const double* ev_t = S.used ? S.ev_t : L.ev_t + e0;
    const uint32_t* ev_n = S.used ? S.ev_n : L.ev_n + e0;
    const uint64_t* sp_mask = S.used ? S.sp : L.species_mask + static_cast<size_t>(j) * L.n_species * NW;
    if (L.stage_words > 0 && n_ev <= L.stage_ev_cap) {
        uint64_t* reg = pnb_msnc_smem + L.program_words + (L.frontier_in_smem ? (mask_words + 2 * static_cast<size_t>(L.cap)) * blockDim.x : 0) +
                        static_cast<size_t>(threadIdx.x) * L.stage_words;
        double* st = reinterpret_cast<double*>(reg);
        uint32_t* sn = reinterpret_cast<uint32_t*>(reg + L.stage_ev_cap);
        uint64_t* ss = reg + L.stage_ev_cap + (L.stage_ev_cap + 1) / 2;
        for (int i = 0; i < n_ev; ++i) { st[i] = ev_t[i]; sn[i] = ev_n[i]; }
        for (int i = 0; i < L.n_species * NW; ++i) ss[i] = sp_mask[i];
        ev_t = st; ev_n = sn; sp_mask = ss;
    }
    __builtin_assume(__isShared(s_steps));
    __builtin_assume(__isShared(s_down));
    __builtin_assume(__isShared(s_theta));
    if (SMEM) {   // the host picks this variant when the frontier AND the staged events are in shared memory
        __builtin_assume(__isShared(fr_mask));
        __builtin_assume(__isShared(fr_lp));
        __builtin_assume(__isShared(ev_t));
        __builtin_assume(__isShared(ev_n));
        __builtin_assume(__isShared(sp_mask));
    }
    if (L.tree_shaped) {   // no reticulation: single-entry frontier, updated in place
        L.out[j] = msnc_dp_tree<NW>(s_steps, L.n_steps, s_down, ev_t, ev_n, n_ev, sp_mask, root_bit, s_theta, fr_mask, stride);
        L.status[j] = JOB_OK;
        return;
    }
    int status;
    const double v = msnc_dp<NW>(s_steps, L.n_steps, s_down, L.W, L.cap, ev_t, ev_n, n_ev, sp_mask, root_bit, s_theta,
                                 fr_mask, fr_lp, stride, &status);
    L.out[j] = v;
    L.status[j] = status;
}

}  // namespace

extern "C" int pnb_msnc_net_create(pnb_ctx* ctx, int32_t n_nodes, int32_t n_edges, const int32_t* edge_src,
                                   const int32_t* edge_dst, const double* edge_gamma, const double* node_height,
                                   const int32_t* node_species, int32_t n_species, pnb_msnc_net** out) {
    PNB_REQUIRE(ctx && out, PNB_ERR_INVALID, "pnb_msnc_net_create: NULL handle");
    PNB_REQUIRE(n_nodes > 0 && n_edges >= 0 && n_species >= 0, PNB_ERR_INVALID, "pnb_msnc_net_create: bad sizes");
    PNB_REQUIRE(node_height && node_species && (n_edges == 0 || (edge_src && edge_dst && edge_gamma)), PNB_ERR_INVALID,
                "pnb_msnc_net_create: NULL array");
    pnb_msnc_net* net = new (std::nothrow) pnb_msnc_net();
    PNB_REQUIRE(net, PNB_ERR_NOMEM, "pnb_msnc_net_create: out of host memory");
    net->ctx = ctx;
    net->n_edges = n_edges;
    const std::string err = compile_network(n_nodes, n_edges, edge_src, edge_dst, edge_gamma, node_height, node_species,
                                            n_species, net->prog);
    if (!err.empty()) {
        delete net;
        set_error("pnb_msnc_net_create: %s", err.c_str());
        return PNB_ERR_INVALID;
    }
    if (ctx->device >= 0) {
        if (cudaSetDevice(ctx->device) != cudaSuccess) {
            delete net;
            set_error("pnb_msnc_net_create: cudaSetDevice failed");
            return PNB_ERR_CUDA;
        }
        void* p = nullptr;
        const size_t sb = net->prog.steps.size() * sizeof(Step);
        std::vector<int32_t> packed(net->prog.down_slot);
        packed.insert(packed.end(), net->prog.down_step.begin(), net->prog.down_step.end());
        packed.insert(packed.end(), net->prog.level.begin(), net->prog.level.end());
        const size_t db = std::max<size_t>(packed.size(), 1) * sizeof(int32_t);
        net->down_bytes = db;
        int rc = ctx->arena.alloc(sb, &p);
        if (rc) { delete net; return rc; }
        net->d_steps = static_cast<Step*>(p);
        rc = ctx->arena.alloc(db, &p);
        if (rc) { ctx->arena.release(net->d_steps, sb); delete net; return rc; }
        net->d_down = static_cast<int32_t*>(p);
        cudaError_t ce = cudaMemcpyAsync(net->d_steps, net->prog.steps.data(), sb, cudaMemcpyHostToDevice, ctx->stream);
        if (ce == cudaSuccess && !packed.empty())
            ce = cudaMemcpyAsync(net->d_down, packed.data(), packed.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);   // the vectors are pageable
        if (ce != cudaSuccess) {
            ctx->arena.release(net->d_steps, sb);
            ctx->arena.release(net->d_down, db);
            delete net;
            set_error("pnb_msnc_net_create: upload failed: %s", cudaGetErrorString(ce));
            return PNB_ERR_CUDA;
        }
    }
    *out = net;
    return PNB_OK;
}

extern "C" int pnb_msnc_net_destroy(pnb_msnc_net* net) {
    if (!net) return PNB_OK;
    if (net->d_steps) {
        cudaSetDevice(net->ctx->device);
        cudaStreamSynchronize(net->ctx->stream);
        net->ctx->arena.release(net->d_steps, net->prog.steps.size() * sizeof(Step));
        net->ctx->arena.release(net->d_down, net->down_bytes);
    }
    delete net;
    return PNB_OK;
}

extern "C" int pnb_msnc_net_info(const pnb_msnc_net* net, int32_t info[4]) {
    PNB_REQUIRE(net && info, PNB_ERR_INVALID, "pnb_msnc_net_info: NULL argument");
    info[0] = net->prog.W;
    info[1] = static_cast<int32_t>(net->prog.steps.size());
    info[2] = static_cast<int32_t>(net->prog.retic_step.size());
    info[3] = net->prog.n_species;
    return PNB_OK;
}

extern "C" int pnb_msnc_net_set_branch_thetas(pnb_msnc_net* net, const double* edge_theta, double root_theta) {
    PNB_REQUIRE(net, PNB_ERR_INVALID, "pnb_msnc_net_set_branch_thetas: NULL network");
    if (edge_theta) {
        for (int32_t e = 0; e < net->n_edges; ++e)
            PNB_REQUIRE(!(edge_theta[e] == edge_theta[e]) || (edge_theta[e] > 0.0 && std::isfinite(edge_theta[e])), PNB_ERR_INVALID,
                        "pnb_msnc_net_set_branch_thetas: theta of edge %d must be finite and positive", e);
        net->edge_theta.assign(edge_theta, edge_theta + net->n_edges);
    } else {
        net->edge_theta.clear();
    }
    PNB_REQUIRE(!(root_theta == root_theta) || (root_theta > 0.0 && std::isfinite(root_theta)), PNB_ERR_INVALID,
                "pnb_msnc_net_set_branch_thetas: root theta must be finite and positive");
    net->root_theta = root_theta;
    return PNB_OK;
}

extern "C" int pnb_msnc_net_set_max_frontier(pnb_msnc_net* net, int64_t entries) {
    PNB_REQUIRE(net && entries >= 1, PNB_ERR_INVALID, "pnb_msnc_net_set_max_frontier: bad argument");
    net->max_frontier = entries;
    return PNB_OK;
}

// The evaluation in two halves, so that a gene-tree move can put both likelihood factors behind ONE host
// synchronisation (pnb_eval_move): msnc_launch prepares the trees on the host, uploads and launches (everything is
// queued on the context's stream, results travel to the K-m-owned pinned buffer); msnc_finish waits and hands out.
static int msnc_finish(pnb_ctx* ctx, double* logp_out, cudaStream_t st);

static int msnc_launch(pnb_ctx* ctx, const pnb_msnc_net* net, int64_t n_trees, const int64_t* tree_off,
                       const int32_t* parent, const double* blen, const int32_t* node_species, double theta, cudaStream_t st) {
    PNB_REQUIRE(ctx && net, PNB_ERR_INVALID, "pnb_msnc_eval: NULL handle");
    PNB_REQUIRE(ctx->device >= 0, PNB_ERR_STATE, "pnb_msnc_eval: host-only planning context has no device");
    PNB_REQUIRE(net->ctx == ctx, PNB_ERR_INVALID, "pnb_msnc_eval: network belongs to another context");
    PNB_REQUIRE(n_trees >= 0, PNB_ERR_INVALID, "pnb_msnc_eval: n_trees < 0");
    PNB_REQUIRE(theta > 0.0 && std::isfinite(theta), PNB_ERR_INVALID, "pnb_msnc_eval: theta must be positive and finite");
    ctx->msnc_pending_trees = 0;
    if (n_trees == 0) return PNB_OK;
    PNB_REQUIRE(tree_off && parent && blen && node_species, PNB_ERR_INVALID, "pnb_msnc_eval: NULL array");
    PNB_CUDA_TRY(cudaSetDevice(ctx->device));
    const auto t_call0 = std::chrono::steady_clock::now();
    ctx->ev_ka = ctx->ev_prune = false;
    const NetProgram& prog = net->prog;
    const int n_species = std::max(prog.n_species, 1);

    // ---- sizes ---------------------------------------------------------------------------
    int max_n = 0;
    for (int64_t j = 0; j < n_trees; ++j) {
        const int64_t n = tree_off[j + 1] - tree_off[j];
        PNB_REQUIRE(n >= 0, PNB_ERR_INVALID, "pnb_msnc_eval: tree_off is not non-decreasing at tree %lld", static_cast<long long>(j));
        PNB_REQUIRE(n <= 64 * MAX_NW, PNB_ERR_UNSUPPORTED, "pnb_msnc_eval: gene tree %lld has %lld nodes; this build handles %d",
                    static_cast<long long>(j), static_cast<long long>(n), 64 * MAX_NW);
        max_n = std::max(max_n, static_cast<int>(n));
    }
    const int NW = max_n <= 64 ? 1 : 2;
    const int64_t total_nodes = tree_off[n_trees] - tree_off[0];
    // events: a tree of n nodes has at most (n - 1) / 2 binary nodes; job j's events start at
    // (tree_off[j] - tree_off[0]) / 2 rounded down -- a bound, offsets are exact per job below
    const size_t ev_cap = static_cast<size_t>(total_nodes / 2 + n_trees);

    const size_t off_evoff = 0;
    const size_t off_root = align_up(off_evoff + static_cast<size_t>(n_trees + 1) * sizeof(int64_t), 256);
    const size_t off_evt = align_up(off_root + static_cast<size_t>(n_trees) * sizeof(int32_t), 256);
    const size_t off_evn = align_up(off_evt + ev_cap * sizeof(double), 256);
    const size_t off_sp = align_up(off_evn + ev_cap * sizeof(uint32_t), 256);
    const size_t off_theta = align_up(off_sp + static_cast<size_t>(n_trees) * n_species * NW * sizeof(uint64_t), 256);
    const size_t stage_bytes = off_theta + prog.steps.size() * 4 * sizeof(double);

    // K-m owns its staging, scratch and result buffers (the pruning path may be in flight next to it); every
    // evaluation ends in msnc_finish, which waits for the stream, so they are free here
    int rc;
    if ((rc = ensure_pinned(ctx->m_hstage, stage_bytes))) return rc;
    if ((rc = ensure_device(ctx->m_dstage, stage_bytes))) return rc;
    char* hs = static_cast<char*>(ctx->m_hstage.p);
    int64_t* ev_off = reinterpret_cast<int64_t*>(hs + off_evoff);
    int32_t* root_bit = reinterpret_cast<int32_t*>(hs + off_root);
    double* ev_t = reinterpret_cast<double*>(hs + off_evt);
    uint32_t* ev_n = reinterpret_cast<uint32_t*>(hs + off_evn);
    uint64_t* sp_mask = reinterpret_cast<uint64_t*>(hs + off_sp);
    fill_step_theta(prog, theta, net->edge_theta.empty() ? nullptr : net->edge_theta.data(), net->root_theta,
                    reinterpret_cast<double*>(hs + off_theta));

    // ---- host preparation, parallel over trees ---------------------------------------------
    // job j writes its events at slot base (tree_off[j] - tree_off[0]) / 2 + j (disjoint, an upper
    // bound on everything before it); ev_off records [begin, end) so no compaction pass is needed.
    const int n_workers = ctx->pool ? ctx->pool->size() : 1;
    std::vector<TreeScratch> scratch(n_workers);
    std::vector<std::string> errs(n_workers);
    std::vector<int64_t> err_job(n_workers, -1);
    std::vector<int32_t> alleles_max(static_cast<size_t>(n_workers) * n_species, 0);
    std::vector<int64_t> ev_end(n_trees);
    auto body = [&](int w, int64_t a, int64_t b) {
        TreeScratch& sc = scratch[w];
        for (int64_t j = a; j < b; ++j) {
            const int64_t o = tree_off[j];
            const int n = static_cast<int>(tree_off[j + 1] - o);
            const int64_t base = (o - tree_off[0]) / 2 + j;
            ev_off[j] = base;
            uint64_t* mj = sp_mask + static_cast<size_t>(j) * n_species * NW;
            if (n == 0) {
                root_bit[j] = -1;
                ev_end[j] = base;
                for (size_t i = 0; i < static_cast<size_t>(n_species) * NW; ++i) mj[i] = 0;
                continue;
            }
            int n_ev = 0, rb = 0;
            const char* e = prepare_tree(n, parent + o, blen + o, node_species + o, prog.n_species, NW, sc, ev_t + base,
                                         ev_n + base, &n_ev, mj, &rb);
            if (e) {
                if (err_job[w] < 0) { err_job[w] = j; errs[w] = e; }
                root_bit[j] = -1;
                ev_end[j] = base;
                continue;
            }
            root_bit[j] = rb;
            ev_end[j] = base + n_ev;
            int32_t* am = alleles_max.data() + static_cast<size_t>(w) * n_species;
            for (int s = 0; s < prog.n_species; ++s) {
                int c = 0;
                for (int i = 0; i < NW; ++i) c += __builtin_popcountll(mj[static_cast<size_t>(s) * NW + i]);
                if (c > am[s]) am[s] = c;
            }
        }
    };
    if (n_trees >= 256 && n_workers > 1) {
        const int64_t grain = std::max<int64_t>(64, n_trees / (n_workers * 4));
        std::atomic<int64_t> next{0};
        ctx->pool->run([&](int w) {
            for (;;) {
                const int64_t a0 = next.fetch_add(grain);
                if (a0 >= n_trees) break;
                body(w, a0, std::min(n_trees, a0 + grain));
            }
        });
    } else {
        body(0, 0, n_trees);
    }
    for (int w = 0; w < n_workers; ++w)
        PNB_REQUIRE(err_job[w] < 0, PNB_ERR_INVALID, "pnb_msnc_eval: gene tree %lld: %s", static_cast<long long>(err_job[w]),
                    errs[w].c_str());
    // The kernel reads job j's events in [ev_off[j], ev_off[j+1]); with the gaps above that range would
    // include the unused tail of job j's slot range, so events are compacted here (a single memmove pass).
    {
        int64_t w_pos = 0;
        for (int64_t j = 0; j < n_trees; ++j) {
            const int64_t b = ev_off[j], n_ev = ev_end[j] - b;
            if (b != w_pos && n_ev > 0) {
                memmove(ev_t + w_pos, ev_t + b, static_cast<size_t>(n_ev) * sizeof(double));
                memmove(ev_n + w_pos, ev_n + b, static_cast<size_t>(n_ev) * sizeof(uint32_t));
            }
            ev_off[j] = w_pos;
            w_pos += n_ev;
        }
        ev_off[n_trees] = w_pos;
    }
    for (int w = 1; w < n_workers; ++w)
        for (int s = 0; s < n_species; ++s)
            alleles_max[s] = std::max(alleles_max[s], alleles_max[static_cast<size_t>(w) * n_species + s]);

    // ---- frontier capacity and scratch ---------------------------------------------------------
    const double bound = frontier_bound(prog, alleles_max.data());
    const int cap = static_cast<int>(std::min<double>(bound, static_cast<double>(net->max_frontier))) + 1;
    // Reticulate networks: one warp per gene tree (msnc_dp_warp); species trees -- a single-entry frontier, nothing to
    // spread over lanes -- and PNB_MSNC_SERIAL=1 (test knob): one thread per gene tree.
    const char* serial_env = getenv("PNB_MSNC_SERIAL");
    const bool serial_forced = serial_env && serial_env[0] == '1';
    const bool warp_mode = !prog.retic_step.empty() && !serial_forced;
    int tsize = 64;
    while (tsize < 2 * cap) tsize <<= 1;
    const size_t job_mask_words = warp_mode ? msnc_warp_words(prog.W, NW, cap, tsize) : msnc_mask_words(prog.W, NW, cap);
    const size_t per_job = warp_mode ? job_mask_words * sizeof(uint64_t)
                                     : job_mask_words * sizeof(uint64_t) + 2 * static_cast<size_t>(cap) * sizeof(double);
    // Block size: small batches use small blocks so that they spread over the SMs; the frontier goes to
    // shared memory when a block's worth of it fits (it does for every tree-shaped network and for
    // networks whose reticulations sit above few gene lineages), else to a global scratch buffer.
    int threads = n_trees >= int64_t(ctx->sm_count) * 256 ? 128 : (n_trees >= int64_t(ctx->sm_count) * 128 ? 64 : 32);
    if (warp_mode) threads = n_trees >= int64_t(ctx->sm_count) * 8 ? 128 : (n_trees >= int64_t(ctx->sm_count) * 4 ? 64 : 32);
    const int lanes_per_job = warp_mode ? 32 : 1;
    // the step program is staged in shared memory by every block (steps, thetas, down slots)
    const size_t program_bytes = align_up(prog.steps.size() * (sizeof(Step) + 4 * sizeof(double)) +
                                          prog.down_slot.size() * sizeof(int32_t), 16);
    PNB_REQUIRE(program_bytes + 4096 < static_cast<size_t>(ctx->max_smem_optin), PNB_ERR_UNSUPPORTED,
                "pnb_msnc_eval: the species network has too many nodes (%zu) for the staged step program", prog.steps.size());
    const size_t smem_limit = static_cast<size_t>(std::max(ctx->max_smem_optin - 2048, 0)) - program_bytes;
    const char* no_smem = getenv("PNB_MSNC_NO_SMEM");   // test knob: force the global-scratch path
    const bool in_smem = per_job * (warp_mode ? 1 : 32) <= smem_limit && !(no_smem && no_smem[0] == '1');
    if (in_smem)
        while (per_job * (threads / lanes_per_job) > smem_limit) threads /= 2;
    const int jobs_per_block = threads / lanes_per_job;
    // staging of events + species masks (per thread, or per warp): taken when it fits next to the frontiers without
    // shrinking the block
    int max_ev = 0;
    for (int64_t j = 0; j < n_trees; ++j) max_ev = std::max<int>(max_ev, static_cast<int>(ev_off[j + 1] - ev_off[j]));
    int stage_ev_cap = max_ev | 1;                                     // odd: neighbouring threads start on different banks
    size_t stage_words = static_cast<size_t>(stage_ev_cap) + (stage_ev_cap + 1) / 2 + static_cast<size_t>(n_species) * NW;
    stage_words |= 1;
    const size_t stage_words_raw = stage_words;
    const size_t fr_bytes = in_smem ? per_job * jobs_per_block : 0;
    const bool stage = max_ev > 0 && !(getenv("PNB_MSNC_NO_STAGE") && getenv("PNB_MSNC_NO_STAGE")[0] == '1') &&
                       fr_bytes + stage_words * 8 * jobs_per_block <= smem_limit;
    if (!stage) stage_words = 0;
    size_t dyn_smem = program_bytes + fr_bytes + stage_words * 8 * jobs_per_block;
    // Species tree, small batch: one warp per gene tree, network nodes of one level side by side (pnb_msnc_tree_warp_kernel).
    // Everything a warp needs (lineage set and factor per node, staged events) lives in shared memory.
    const size_t tree_warp_words = prog.steps.size() * (NW + 1) + (prog.steps.size() + 1) / 2 + stage_words_raw;
    const size_t tree_program_bytes = align_up(prog.steps.size() * (sizeof(Step) + 4 * sizeof(double)) +
                                               (2 * prog.down_slot.size() + prog.steps.size()) * sizeof(int32_t), 16);
    const bool tree_warp = prog.retic_step.empty() && !serial_forced && in_smem && n_trees <= int64_t(ctx->sm_count) * 64 &&
                           tree_program_bytes + tree_warp_words * 8 + 4096 < static_cast<size_t>(ctx->max_smem_optin);
    int tree_warps_per_block = 1;
    if (tree_warp) {
        tree_warps_per_block = n_trees >= int64_t(ctx->sm_count) * 4 ? 4 : (n_trees >= int64_t(ctx->sm_count) * 2 ? 2 : 1);
        while (tree_warps_per_block > 1 &&
               tree_program_bytes + tree_warp_words * 8 * tree_warps_per_block + 4096 >= static_cast<size_t>(ctx->max_smem_optin))
            tree_warps_per_block /= 2;
        dyn_smem = tree_program_bytes + tree_warp_words * 8 * tree_warps_per_block;
    }
    // jobs per launch: bounded scratch (the frontiers of a batch may not fit at once)
    size_t scratch_budget = size_t{2} << 30;
    if (const char* mb = getenv("PNB_MSNC_SCRATCH_MB")) {   // test knob: small budgets exercise the multi-launch path
        const long v = atol(mb);
        if (v > 0) scratch_budget = static_cast<size_t>(v) << 20;
    }
    const int64_t jobs_per_launch =
        in_smem ? n_trees : std::max<int64_t>(1, std::min<int64_t>(n_trees, static_cast<int64_t>(scratch_budget / per_job)));
    const size_t off_lp = align_up(static_cast<size_t>(jobs_per_launch) * job_mask_words * sizeof(uint64_t), 256);
    if (!in_smem && (rc = ensure_device(ctx->m_scratch, off_lp + (warp_mode ? 0 : static_cast<size_t>(jobs_per_launch) * 2 * cap * sizeof(double))))) return rc;
    if (!ctx->msnc_configured) {   // once per context: opt in to large dynamic shared memory
        const int lim = std::max(ctx->max_smem_optin - 2048, 0);
        PNB_CUDA_TRY(cudaFuncSetAttribute(pnb_msnc_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        PNB_CUDA_TRY(cudaFuncSetAttribute(pnb_msnc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        PNB_CUDA_TRY(cudaFuncSetAttribute(pnb_msnc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        PNB_CUDA_TRY(cudaFuncSetAttribute(pnb_msnc_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        PNB_CUDA_TRY(cudaFuncSetAttribute(pnb_msnc_warp_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        PNB_CUDA_TRY(cudaFuncSetAttribute(pnb_msnc_warp_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        PNB_CUDA_TRY(cudaFuncSetAttribute(pnb_msnc_warp_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        PNB_CUDA_TRY(cudaFuncSetAttribute(pnb_msnc_warp_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        PNB_CUDA_TRY(cudaFuncSetAttribute(pnb_msnc_tree_warp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        PNB_CUDA_TRY(cudaFuncSetAttribute(pnb_msnc_tree_warp_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        ctx->msnc_configured = true;
    }
    const size_t off_status = align_up(static_cast<size_t>(n_trees) * sizeof(double), 256);
    if ((rc = ensure_device(ctx->m_dout, off_status + static_cast<size_t>(n_trees) * sizeof(int32_t)))) return rc;
    if ((rc = ensure_pinned(ctx->m_hout, off_status + static_cast<size_t>(n_trees) * sizeof(int32_t)))) return rc;

    ctx->t_compile_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call0).count();
    // Small batches (a gene-tree move scores ONE tree): the kernel reads the packed input straight from the pinned staging
    // buffer and writes its results into the pinned result buffer (both mapped into the device's address space) -- no
    // H2D copy before the kernel and no D2H copy after it, i.e. two DMA start-ups less on the critical path of a move.
    bool zero_copy = false;
    if (ctx->allow_fused && n_trees <= 64 && stage_bytes <= (64u << 10)) {
        void *da = nullptr, *db = nullptr;
        zero_copy = cudaHostGetDevicePointer(&da, hs, 0) == cudaSuccess && da == static_cast<void*>(hs) &&
                    cudaHostGetDevicePointer(&db, ctx->m_hout.p, 0) == cudaSuccess && db == ctx->m_hout.p;
        if (!zero_copy) (void)cudaGetLastError();
    }
    MsncSmall small = {};
    if (n_trees == 1 && ev_off[1] <= MS_MAX_EV && static_cast<size_t>(n_species) * NW <= MS_MAX_SP &&
        prog.steps.size() * 4 <= MS_MAX_THETA && root_bit[0] >= 0) {
        small.used = 1;
        small.n_ev = static_cast<int32_t>(ev_off[1]);
        small.root_bit = root_bit[0];
        memcpy(small.ev_t, ev_t, static_cast<size_t>(ev_off[1]) * sizeof(double));
        memcpy(small.ev_n, ev_n, static_cast<size_t>(ev_off[1]) * sizeof(uint32_t));
        memcpy(small.sp, sp_mask, static_cast<size_t>(n_species) * NW * sizeof(uint64_t));
        memcpy(small.theta, hs + off_theta, prog.steps.size() * 4 * sizeof(double));
    }
    if (!zero_copy && !small.used) PNB_CUDA_TRY(cudaMemcpyAsync(ctx->m_dstage.p, hs, stage_bytes, cudaMemcpyHostToDevice, st));
    if (ctx->profiling) PNB_CUDA_TRY(cudaEventRecord(ctx->ev[2], st));   // reported as the "prune" slot of last_timings
    char* ds = zero_copy ? hs : static_cast<char*>(ctx->m_dstage.p);
    char* dout = static_cast<char*>(zero_copy ? ctx->m_hout.p : ctx->m_dout.p);
    MsncLaunch L;
    L.steps = net->d_steps;
    L.down_slot = net->d_down;
    L.n_steps = static_cast<int32_t>(prog.steps.size());
    L.W = prog.W;
    L.cap = cap;
    L.n_species = n_species;
    L.step_theta = reinterpret_cast<const double*>(ds + off_theta);
    L.frontier_in_smem = in_smem ? 1 : 0;
    L.tree_shaped = prog.retic_step.empty() ? 1 : 0;
    L.n_down = static_cast<int32_t>(prog.down_slot.size());
    L.program_words = static_cast<int32_t>(program_bytes / 8);
    L.stage_words = static_cast<int32_t>(stage_words);
    L.stage_ev_cap = stage_ev_cap;
    L.warp_tsize = tsize;
    L.warp_job_words = static_cast<int64_t>(tree_warp ? tree_warp_words : job_mask_words);
    if (tree_warp) L.program_words = static_cast<int32_t>(tree_program_bytes / 8);
    L.fr_mask = in_smem ? nullptr : static_cast<uint64_t*>(ctx->m_scratch.p);
    L.fr_lp = in_smem ? nullptr : reinterpret_cast<double*>(static_cast<char*>(ctx->m_scratch.p) + off_lp);
    int64_t launches = 0;
    for (int64_t j0 = 0; j0 < n_trees; j0 += jobs_per_launch) {
        const int64_t nj = std::min(jobs_per_launch, n_trees - j0);
        L.n_jobs = nj;
        L.ev_off = reinterpret_cast<const int64_t*>(ds + off_evoff) + j0;
        L.ev_t = reinterpret_cast<const double*>(ds + off_evt);
        L.ev_n = reinterpret_cast<const uint32_t*>(ds + off_evn);
        L.species_mask = reinterpret_cast<const uint64_t*>(ds + off_sp) + static_cast<size_t>(j0) * n_species * NW;
        L.root_bit = reinterpret_cast<const int32_t*>(ds + off_root) + j0;
        L.out = reinterpret_cast<double*>(dout) + j0;
        L.status = reinterpret_cast<int32_t*>(dout + off_status) + j0;
        unsigned blocks = static_cast<unsigned>((nj + jobs_per_block - 1) / jobs_per_block);
        if (tree_warp) {
            blocks = static_cast<unsigned>((nj + tree_warps_per_block - 1) / tree_warps_per_block);
            if (NW == 1) pnb_msnc_tree_warp_kernel<1><<<blocks, 32 * tree_warps_per_block, dyn_smem, st>>>(L, small);
            else pnb_msnc_tree_warp_kernel<2><<<blocks, 32 * tree_warps_per_block, dyn_smem, st>>>(L, small);
            PNB_CUDA_TRY(cudaGetLastError());
            ++launches;
            continue;
        }
        // all_smem: frontiers AND staged events / lineage sets of every job are in shared memory (every tree has at most
        // stage_ev_cap events by construction) -> the variant compiled with shared-memory addressing throughout
        const bool all_smem = in_smem && stage;
        if (warp_mode) {
            if (NW == 1) { if (all_smem) pnb_msnc_warp_kernel<1, true><<<blocks, threads, dyn_smem, st>>>(L, small); else pnb_msnc_warp_kernel<1, false><<<blocks, threads, dyn_smem, st>>>(L, small); }
            else { if (all_smem) pnb_msnc_warp_kernel<2, true><<<blocks, threads, dyn_smem, st>>>(L, small); else pnb_msnc_warp_kernel<2, false><<<blocks, threads, dyn_smem, st>>>(L, small); }
        } else {
            if (NW == 1) { if (all_smem) pnb_msnc_kernel<1, true><<<blocks, threads, dyn_smem, st>>>(L, small); else pnb_msnc_kernel<1, false><<<blocks, threads, dyn_smem, st>>>(L, small); }
            else { if (all_smem) pnb_msnc_kernel<2, true><<<blocks, threads, dyn_smem, st>>>(L, small); else pnb_msnc_kernel<2, false><<<blocks, threads, dyn_smem, st>>>(L, small); }
        }
        PNB_CUDA_TRY(cudaGetLastError());
        ++launches;
    }
    if (ctx->profiling) { PNB_CUDA_TRY(cudaEventRecord(ctx->ev[3], st)); ctx->ev_prune = true; }
    if (!zero_copy)
        PNB_CUDA_TRY(cudaMemcpyAsync(ctx->m_hout.p, ctx->m_dout.p, off_status + static_cast<size_t>(n_trees) * sizeof(int32_t),
                                     cudaMemcpyDeviceToHost, st));
    ctx->msnc_pending_trees = n_trees;
    ctx->msnc_pending_status_off = off_status;
    ctx->msnc_pending_cap = cap;
    ctx->msnc_stats[0] = n_trees;
    ctx->msnc_stats[1] = ev_off[n_trees];
    ctx->msnc_stats[2] = cap - 1;
    ctx->msnc_stats[3] = launches;
    ctx->msnc_stats[4] = static_cast<int64_t>(stage_bytes);
    ctx->msnc_stats[5] = static_cast<int64_t>(off_status + static_cast<size_t>(n_trees) * sizeof(int32_t));
    ctx->msnc_stats[6] = static_cast<int64_t>(fr_bytes);   // bytes of shared memory per block holding frontiers
    ctx->msnc_stats[7] = threads;
    return PNB_OK;
}

static int msnc_finish(pnb_ctx* ctx, double* logp_out, cudaStream_t st) {
    const int64_t n_trees = ctx->msnc_pending_trees;
    if (n_trees == 0) return PNB_OK;
    ctx->msnc_pending_trees = 0;
    PNB_CUDA_TRY(cudaStreamSynchronize(st));
    const double* ho = static_cast<const double*>(ctx->m_hout.p);
    const int32_t* hstat = reinterpret_cast<const int32_t*>(static_cast<const char*>(ctx->m_hout.p) + ctx->msnc_pending_status_off);
    for (int64_t j = 0; j < n_trees; ++j) {
        PNB_REQUIRE(hstat[j] == JOB_OK, PNB_ERR_UNSUPPORTED,
                    "pnb_msnc_eval: gene tree %lld needs more than %d frontier entries (raise pnb_msnc_net_set_max_frontier)",
                    static_cast<long long>(j), ctx->msnc_pending_cap - 1);
        logp_out[j] = ho[j];
    }
    return PNB_OK;
}

extern "C" int pnb_msnc_eval(pnb_ctx* ctx, const pnb_msnc_net* net, int64_t n_trees, const int64_t* tree_off,
                             const int32_t* parent, const double* blen, const int32_t* node_species, double theta,
                             double* logp_out) {
    PNB_REQUIRE(n_trees <= 0 || logp_out, PNB_ERR_INVALID, "pnb_msnc_eval: NULL array");
    const auto t_call0 = std::chrono::steady_clock::now();
    PNB_REQUIRE(ctx != nullptr, PNB_ERR_INVALID, "pnb_msnc_eval: NULL handle");
    int rc = msnc_launch(ctx, net, n_trees, tree_off, parent, blen, node_species, theta, ctx->stream);
    if (rc) return rc;
    if ((rc = msnc_finish(ctx, logp_out, ctx->stream))) return rc;
    memcpy(ctx->stats, ctx->msnc_stats, sizeof(ctx->stats));
    ctx->t_call_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call0).count();
    return PNB_OK;
}

// One gene-tree move of MCMC_SEQ re-scores both factors of ONE locus (src/_mcmc_seq.py:746-760 with a single dirty
// locus): the pruning of its alignment along the changed path and the MSNC density of its gene tree.  Both kernels are
// queued on the context's stream and the host waits once.
extern "C" int pnb_eval_move(pnb_ctx* ctx, const pnb_model* m, pnb_locus* locus, int32_t n_nodes, const int32_t* parent,
                             const double* blen, const int32_t* tip_row, double site_rate, uint32_t flags,
                             const pnb_msnc_net* net, const int32_t* node_species, double theta, double out[2]) {
    PNB_REQUIRE(ctx && m && locus && net && out, PNB_ERR_INVALID, "pnb_eval_move: NULL argument");
    PNB_REQUIRE(!(flags & (PNB_EVAL_OUT_DEVICE | PNB_EVAL_NOSTORE)), PNB_ERR_INVALID,
                "pnb_eval_move: results go to the host and the evaluation is stored (PENDING or committed)");
    PNB_REQUIRE(n_nodes >= 1, PNB_ERR_INVALID, "pnb_eval_move: empty tree");
    const auto t_call0 = std::chrono::steady_clock::now();
    const int64_t off[2] = {0, n_nodes};
    // the two factors are independent: K-m runs on the context's second stream NEXT TO the pruning kernel
    int rc = msnc_launch(ctx, net, 1, off, parent, blen, node_species, theta, ctx->copy_stream);
    if (rc) return rc;
    pnb_locus* loci[1] = {locus};
    const int rc_f = pnb_eval(ctx, m, 1, loci, off, parent, blen, tip_row, nullptr, site_rate, flags, &out[0]);
    const int rc_m = msnc_finish(ctx, &out[1], ctx->copy_stream);     // always drains, also after a failed pruning call
    if (rc_f) return rc_f;
    if (rc_m) {
        if (flags & PNB_EVAL_PENDING) pnb_rollback(ctx, 1, loci);
        return rc_m;
    }
    ctx->t_call_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call0).count();
    return PNB_OK;
}
