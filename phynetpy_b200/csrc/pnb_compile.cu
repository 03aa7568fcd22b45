// pnb_compile.cu -- C ABI of libphynet_b200.so, part 2: the host half of an evaluation.
//
// compile_job resolves a gene tree against the locus' committed partials cache (which node already
// lives in which HBM slot), decides what must be recomputed and emits the flat op program the pruning
// kernel walks (register forwarding, stash order, tip-only nodes, segment numbering, per-locus program
// templates).  No CUDA call in this file: pnb_plan runs it without a device for the CPU test-suite.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>

#include "pnb_host.hpp"

using namespace pnb;

// =====================================================================
// tree -> op program
// =====================================================================
namespace {

constexpr int64_t REF_ABSENT = -1;   // leaf whose taxon is not in the alignment
constexpr int64_t REF_COMPUTED = -3; // NOSTORE: recomputed node that has no slot


// explicit (host-supplied) P matrices arrive row-major; the pruning kernel reads P column by column (pnb_device.cuh)
inline void copy_transposed(double* dst, const double* src) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) dst[j * 4 + i] = src[i * 4 + j];
}

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
}  // namespace

int pnb::compile_job(CompileScratch& S, pnb_locus* loc, const pnb_model* m, int32_t n, const int32_t* parent,
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
    const Op* res_ops = nullptr;
    auto refresh_program = [&](pnb_locus::Template& T) {
        const int n_ops_t = T.jo.n_ops;
        // with a valid device-resident copy of the program only the branch lengths travel.  The decision is
        // taken HERE (caller holds the template read lock, or owns the locus in store mode) and travels with
        // the job: a later job of the same batch may rebuild the template, but never rewrites a buffer that
        // carries this batch's mark (pnb_eval gives the rebuilt program a fresh buffer instead).
        if (n_ops_t > 0 && T.d_ops_valid && T.d_ops) {
            res_ops = T.d_ops;
            T.d_ops_used_mark.store(loc->ctx->batch_counter, std::memory_order_relaxed);
        } else if (n_ops_t > 0) {
            memcpy(ops, T.ops.data(), static_cast<size_t>(n_ops_t) * sizeof(Op));
        }
        for (int k = 0; k < n_ops_t; ++k) {
            const int c = T.op_child[k];
            t_out[k] = eff_len(c);
            if (Pexp_out) copy_transposed(Pexp_out + static_cast<size_t>(k) * 16, P_edge + static_cast<size_t>(c) * 16);
        }
    };
    // ---- fastest path: a full, immediately committed re-evaluation of the topology the committed
    // cache already describes (same slots): refresh branch lengths in place, no pending image.
    const bool rescale = (flags & PNB_EVAL_RESCALE) != 0;
    const bool no_tmpl = (flags & PNB_COMPILE_NO_TEMPLATE) != 0;  // resident batches own their programs
    if (!no_tmpl && !nostore && full && !(flags & PNB_EVAL_PENDING) && loc->committed_from_tmpl && C.valid &&
        C.rescale == rescale && loc->tmpl[0].n_slots == loc->n_slots && same_topology(loc->tmpl[0])) {
        pnb_locus::Template& T0 = loc->tmpl[0];
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
        jo->res_ops = res_ops;
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
    if (all_dirty_canonical && !no_tmpl) {
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
        jo->res_ops = res_ops;
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
    {
        // Small subtrees are VIRTUAL: a node with at most `virt_tips` tips below it (and every node whose children are all
        // tips, whatever its arity) is a short product of P columns selected by 1-byte tip codes -- cheaper to recompute
        // from the codes than to move 32 B/pattern through HBM each way.  post-order: children before parents.
        const int vmax = loc->ctx->virt_tips;
        auto& tips = S.order;   // scratch (re-used below)
        tips.assign(n, 0);
        for (int v : post) {
            if (cnt[v] == 0) { tips[v] = 1; continue; }
            int t = 0;
            bool all_tips = true;
            for (int j = 0; j < cnt[v]; ++j) {
                const int c = kids[off[v] + j];
                t += tips[c];
                all_tips &= cnt[c] == 0;
            }
            tips[v] = t;
            virt[v] = all_tips || t <= vmax;
        }
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
        if (dirty[v] || virt[v]) {   // (a clean virtual subtree is recomputed too whenever its parent is)
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
    const bool hold_reg = loc->ctx->hold_reg;
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
            // stored nodes that are recomputed first, virtual (small) subtrees after them, then everything by
            // decreasing stash need (stable): a virtual subtree with internal structure needs stash room too, and
            // need[v] above is computed for exactly this order
            for (int j = 0; j < cnt[v]; ++j) {
                const int c = kids[off[v] + j];
                if (cnt[c] > 0 && dirty[c] && !virt[c]) dk.push_back(c);
            }
            for (int j = 0; j < cnt[v]; ++j) {
                const int c = kids[off[v] + j];
                if (cnt[c] > 0 && virt[c]) dk.push_back(c);
            }
            if (dk.size() - base > 1)
                std::stable_sort(dk.begin() + base, dk.end(), [&](int a, int b) { return need[a] > need[b]; });
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
                    if (!is_last) {
                        // the innermost waiting sibling -- nothing else is held while the last sibling is computed -- stays in
                        // the kernel's HOLD registers: no stash slot, no L2 round trip, and no write at all if it is virtual
                        const bool reg_hold = hold_reg && i == f.kd - 2 && need[dk[base + f.kd - 1]] == 0;
                        disp = reg_hold ? -3 : ((nostore || sp + i < stash_levels) ? sp + i : -2);
                    }
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
                if (Pexp_out) copy_transposed(Pexp_out + static_cast<size_t>(n_ops) * 16, P_edge + static_cast<size_t>(c) * 16);
                S.op_child.push_back(c);
                n_ops++;
            };
            if (f.kd > 0) add(K_REG, 0, 0, dk[base + f.kd - 1]);
            for (int i = 0; i + 1 < f.kd; ++i) {
                const int d = dk[base + i];
                if (hold_reg && i == f.kd - 2 && need[dk[base + f.kd - 1]] == 0) add(K_STK, -1, 0, d);   // from the HOLD registers
                else if (nostore || f.sp + i < stash_levels) add(K_STK, f.sp + i, 0, d);
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
            if (f.push_idx == -3) last.flags |= F_HOLD;
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
    jo->res_ops = nullptr;
    int expect0 = 0;
    const bool write_locked = all_dirty_canonical && !no_tmpl &&
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
        T.jo.res_ops = nullptr;
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

void pnb::locus_commit(pnb_locus* loc) {
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

void pnb::locus_rollback(pnb_locus* loc) {
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
        if (jo.tmpl_hit >= 0 && jo.res_ops && jo.n_ops > 0)
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
