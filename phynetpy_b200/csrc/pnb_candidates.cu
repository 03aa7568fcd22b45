// pnb_candidates.cu -- host-side enumeration of the candidate sets of the coupled moves, for MANY gene trees
// in one call, straight into the CSR arrays pnb_eval / pnb_msnc_eval take.
//
// Replaces, on flat arrays, _candidate_set / _gene_tree_nni_neighbors / _scaled_gene_tree
// (src/_mcmc_seq.py:1724-1861): the tree, its height-preserving NNI neighbours, each at the given height
// scales.  Once the scoring of a coupled move runs on the device (four calls), enumerating ~100 members
// for each of 10^2..10^3 loci in Python is what the move costs; this is integer / copy work, done here
// without tree objects or signature hashing:
//   * an NNI around v changes exactly one clade (v's) and no height: neighbours are told apart by
//     (v, new clade of v), clades being 128-bit sets of leaf node indices;
//   * a scaled member's lengths are max(h[parent]*s - h[child]*s, 0) -- computed exactly like that, so any
//     scale gives the bits _sync_lengths gives on the scaled heights;
//   * two scales of one base coincide only if all internal heights round to the same 12 decimals; only then
//     are scales dropped.
// No device involved (a host-only context, or none, will do).
#include "pnb_host.hpp"

#include <atomic>
#include <cmath>
#include <functional>

using namespace pnb;

namespace {

typedef unsigned __int128 u128;
constexpr int CAND_MAX_NODES = 128;

struct Enumerated {
    int n_bases = 0;
    int n_scales = 0;             // scales kept (after the degenerate-height check)
    int scale_idx[16];            // indices into the caller's scale list
    int16_t base_v[2 * CAND_MAX_NODES], base_s[2 * CAND_MAX_NODES], base_c[2 * CAND_MAX_NODES];   // base 0 = the tree itself
};

// Enumerates the bases (tree + NNI neighbours) of one tree.  Returns nullptr or an error text.
const char* enumerate_tree(int n, const int32_t* parent, const double* h, int n_scales, const double* scales,
                           Enumerated& out) {
    if (n > CAND_MAX_NODES) return "gene tree has more than 128 nodes";
    if (n_scales > 16) return "more than 16 height scales";
    int16_t first_kid[CAND_MAX_NODES], next_sib[CAND_MAX_NODES], n_kids[CAND_MAX_NODES], last_kid[CAND_MAX_NODES];
    for (int i = 0; i < n; ++i) { first_kid[i] = -1; next_sib[i] = -1; n_kids[i] = 0; last_kid[i] = -1; }
    int root = -1;
    for (int i = 0; i < n; ++i) {           // children in ascending node order
        const int p = parent[i];
        if (p < 0) { if (root < 0) root = i; continue; }
        if (p >= n || p == i) return "parent index out of range";
        if (first_kid[p] < 0) first_kid[p] = static_cast<int16_t>(i);
        else next_sib[last_kid[p]] = static_cast<int16_t>(i);
        last_kid[p] = static_cast<int16_t>(i);
        ++n_kids[p];
    }
    if (root < 0) return "tree has no root";
    // clades, children before parents (Kahn)
    u128 clade[CAND_MAX_NODES];
    int16_t remaining[CAND_MAX_NODES], queue[CAND_MAX_NODES];
    int tail = 0;
    for (int i = 0; i < n; ++i) {
        remaining[i] = n_kids[i];
        clade[i] = 0;
        if (n_kids[i] == 0) { clade[i] = static_cast<u128>(1) << i; queue[tail++] = static_cast<int16_t>(i); }
    }
    for (int head = 0; head < tail; ++head) {
        const int v = queue[head], p = parent[v];
        if (p < 0) continue;
        clade[p] |= clade[v];
        if (--remaining[p] == 0) queue[tail++] = static_cast<int16_t>(p);
    }
    if (tail != n) return "tree has a cycle";
    out.n_bases = 1;
    out.base_v[0] = out.base_s[0] = out.base_c[0] = -1;
    u128 seen_mask[2 * CAND_MAX_NODES];
    for (int v = 0; v < n; ++v) {
        if (n_kids[v] < 2 || parent[v] < 0) continue;
        const int p = parent[v];
        if (n_kids[p] != 2) continue;                       // exactly one sibling
        const int s = (first_kid[p] == v) ? next_sib[first_kid[p]] : first_kid[p];
        if (!(h[v] > h[s])) continue;
        for (int c = first_kid[v]; c >= 0; c = next_sib[c]) {
            const u128 nm = (clade[v] & ~clade[c]) | clade[s];
            bool dup = false;
            for (int b = 1; b < out.n_bases && !dup; ++b) dup = (out.base_v[b] == v && seen_mask[b] == nm);
            if (dup) continue;
            if (out.n_bases >= 2 * CAND_MAX_NODES) return "too many NNI neighbours";
            seen_mask[out.n_bases] = nm;
            out.base_v[out.n_bases] = static_cast<int16_t>(v);
            out.base_s[out.n_bases] = static_cast<int16_t>(s);
            out.base_c[out.n_bases] = static_cast<int16_t>(c);
            ++out.n_bases;
        }
    }
    // scales: all kept unless the internal heights are so small that two scales round to the same 12 decimals
    double hmax = 0.0;
    int n_int = 0;
    for (int i = 0; i < n; ++i)
        if (n_kids[i] > 0) { hmax = std::max(hmax, std::fabs(h[i])); ++n_int; }
    double gap = INFINITY;
    for (int a = 0; a < n_scales; ++a)
        for (int b = a + 1; b < n_scales; ++b) gap = std::min(gap, std::fabs(scales[a] - scales[b]));
    out.n_scales = 0;
    if (n_int > 0 && hmax * gap >= 1e-9) {
        for (int k = 0; k < n_scales; ++k) out.scale_idx[out.n_scales++] = k;
    } else {
        for (int k = 0; k < n_scales; ++k) {
            bool dup = false;
            for (int q = 0; q < out.n_scales && !dup; ++q) {
                bool same = true;
                for (int i = 0; i < n && same; ++i)
                    if (n_kids[i] > 0) {
                        const double a = (scales[k] != 1.0) ? h[i] * scales[k] : h[i];
                        const double b = (scales[out.scale_idx[q]] != 1.0) ? h[i] * scales[out.scale_idx[q]] : h[i];
                        same = (std::nearbyint(a * 1e12) / 1e12) == (std::nearbyint(b * 1e12) / 1e12);   // np.round(x, 12)
                    }
                dup = same;
            }
            if (!dup) out.scale_idx[out.n_scales++] = k;
        }
    }
    return nullptr;
}

}  // namespace

extern "C" int pnb_candidates_count(pnb_ctx* ctx, int64_t n_trees, const int64_t* tree_off, const int32_t* parent,
                                    const double* height, int32_t n_scales, const double* scales,
                                    int64_t* n_members_out) {
    (void)ctx;
    PNB_REQUIRE(n_trees >= 0 && n_scales >= 1, PNB_ERR_INVALID, "pnb_candidates_count: bad sizes");
    if (n_trees == 0) return PNB_OK;
    PNB_REQUIRE(tree_off && parent && height && scales && n_members_out, PNB_ERR_INVALID, "pnb_candidates_count: NULL array");
    for (int64_t j = 0; j < n_trees; ++j) {
        Enumerated E;
        const int64_t o = tree_off[j];
        const char* e = enumerate_tree(static_cast<int>(tree_off[j + 1] - o), parent + o, height + o, n_scales, scales, E);
        PNB_REQUIRE(!e, PNB_ERR_INVALID, "pnb_candidates_count: tree %lld: %s", static_cast<long long>(j), e);
        n_members_out[j] = static_cast<int64_t>(E.n_bases) * E.n_scales;
    }
    return PNB_OK;
}

extern "C" int pnb_candidates_fill(pnb_ctx* ctx, int64_t n_trees, const int64_t* tree_off, const int32_t* parent,
                                   const double* blen, const double* height, int32_t n_scales, const double* scales,
                                   const int64_t* member_off, int32_t* parent_out, double* blen_out, double* height_out,
                                   int32_t* changed_node_out, double* scale_out) {
    PNB_REQUIRE(n_trees >= 0 && n_scales >= 1, PNB_ERR_INVALID, "pnb_candidates_fill: bad sizes");
    if (n_trees == 0) return PNB_OK;
    PNB_REQUIRE(tree_off && parent && blen && height && scales && member_off && parent_out && blen_out && height_out,
                PNB_ERR_INVALID, "pnb_candidates_fill: NULL array");
    std::atomic<int64_t> bad{-1};
    const char* bad_text = nullptr;
    // rows of tree j start at row_off(j) = sum_{i<j} members_i * n_i ; computed serially (cheap), filled in parallel
    std::vector<int64_t> row_off(n_trees + 1, 0);
    for (int64_t j = 0; j < n_trees; ++j)
        row_off[j + 1] = row_off[j] + (member_off[j + 1] - member_off[j]) * (tree_off[j + 1] - tree_off[j]);
    auto body = [&](int64_t a, int64_t b) {
        for (int64_t j = a; j < b; ++j) {
            Enumerated E;
            const int64_t o = tree_off[j];
            const int n = static_cast<int>(tree_off[j + 1] - o);
            const int32_t* par = parent + o;
            const double* h = height + o;
            const char* e = enumerate_tree(n, par, h, n_scales, scales, E);
            if (e || static_cast<int64_t>(E.n_bases) * E.n_scales != member_off[j + 1] - member_off[j]) {
                int64_t expect = -1;
                if (bad.compare_exchange_strong(expect, j)) bad_text = e ? e : "member count differs from pnb_candidates_count";
                continue;
            }
            int64_t row = row_off[j];
            int64_t m = member_off[j];
            for (int bi = 0; bi < E.n_bases; ++bi) {
                for (int k = 0; k < E.n_scales; ++k, ++m, row += n) {
                    const double sc = scales[E.scale_idx[k]];
                    int32_t* po = parent_out + row;
                    double* lo = blen_out + row;
                    double* ho = height_out + row;
                    for (int i = 0; i < n; ++i) po[i] = par[i];
                    if (bi > 0) {                                   // s goes under v, c goes under v's parent
                        po[E.base_s[bi]] = E.base_v[bi];
                        po[E.base_c[bi]] = par[E.base_v[bi]];
                    }
                    for (int i = 0; i < n; ++i) ho[i] = (sc != 1.0) ? h[i] * sc : h[i];
                    if (bi == 0 && sc == 1.0) {                     // the tree itself: lengths untouched
                        for (int i = 0; i < n; ++i) lo[i] = blen[o + i];
                    } else {
                        for (int i = 0; i < n; ++i) {
                            const int p = po[i];
                            if (p < 0) { lo[i] = NAN; continue; }
                            const double d = ho[p] - ho[i];
                            lo[i] = d > 0.0 ? d : 0.0;              // _sync_lengths clamps at 0
                        }
                    }
                    if (changed_node_out) changed_node_out[m] = bi > 0 ? E.base_v[bi] : -1;
                    if (scale_out) scale_out[m] = sc;
                }
            }
        }
    };
    WorkerPool* pool = (ctx && n_trees >= 64) ? ctx->pool : nullptr;
    if (pool && pool->size() > 1) {
        const int64_t grain = std::max<int64_t>(8, n_trees / (pool->size() * 4));
        std::atomic<int64_t> next{0};
        pool->run([&](int) {
            for (;;) {
                const int64_t a0 = next.fetch_add(grain);
                if (a0 >= n_trees) break;
                body(a0, std::min(n_trees, a0 + grain));
            }
        });
    } else {
        body(0, n_trees);
    }
    PNB_REQUIRE(bad.load() < 0, PNB_ERR_INVALID, "pnb_candidates_fill: tree %lld: %s", static_cast<long long>(bad.load()), bad_text);
    return PNB_OK;
}
