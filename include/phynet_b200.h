/*
 * phynet_b200.h -- C ABI of libphynet_b200.so
 *
 * A from-scratch sm_100a (B200) implementation of ONE hot path of PhyNetPy: the
 * sequence log-likelihood (Felsenstein pruning under GTR-family models) that
 * MCMC_SEQ evaluates per locus per proposal.  The reference has no FFI for this
 * path (it is NumPy + Python recursion); each entry point below therefore cites
 * the reference *Python* interface it replaces (paths are into the reference
 * checkout, file:line).  INTEGRATION.md shows the ctypes binding a reference
 * maintainer would add.
 *
 * Conventions
 *   - every function returns int: 0 = PNB_OK, otherwise an error code; the text
 *     is available from pnb_last_error() (thread-local).  Nothing aborts or
 *     throws across the boundary.
 *   - handles are opaque, owned by the library, freed by the matching *_destroy.
 *   - host arrays are borrowed for the duration of the call only.
 *   - calls on one context are serialised by the caller (the reference is
 *     single-threaded under the GIL: src/_mcmc_seq.py:746-761).
 *   - no torch / Python types: plain pointers and sizes.
 *   - there is NO CPU fallback: without a CUDA device pnb_ctx_create fails.
 *
 * State order A,C,G,T = 0..3 (src/_seq_likelihood.py:104-105); tip codes are
 * 4-bit masks A=1,C=2,G=4,T=8 (IUPAC table src/_seq_likelihood.py:111-117);
 * P[i*4+j] = prob(parent state i -> child state j) (src/_seq_likelihood.py:238).
 */
#ifndef PHYNET_B200_H
#define PHYNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNB_ABI_VERSION 1

#if defined(__GNUC__)
#define PNB_API __attribute__((visibility("default")))
#else
#define PNB_API
#endif

/* status codes */
#define PNB_OK             0
#define PNB_ERR_INVALID    1   /* bad argument (maps to ValueError)            */
#define PNB_ERR_CUDA       2   /* CUDA runtime / launch failure               */
#define PNB_ERR_NOMEM      3   /* host or device allocation failure           */
#define PNB_ERR_STATE      4   /* call sequence error (e.g. pending eval)     */
#define PNB_ERR_UNSUPPORTED 5  /* shape outside this build's limits           */

/* pnb_eval flags */
#define PNB_EVAL_FULL       0x1u  /* ignore cached partials: recompute every internal node        */
#define PNB_EVAL_PENDING    0x2u  /* keep result pending until pnb_commit / pnb_rollback          */
#define PNB_EVAL_NOSTORE    0x4u  /* score only: read cached partials, write none (candidates)    */
#define PNB_EVAL_OUT_DEVICE 0x8u  /* logL_out is a DEVICE pointer; call returns without host sync */
#define PNB_EVAL_RESCALE    0x10u /* beyond the reference: renormalise every node by an exact power of
                                   * two and carry the exponent, so huge trees do not underflow (the
                                   * reference, and the default here, floor such sites at 1e-300,
                                   * src/_seq_likelihood.py:396).  Cached partials of one mode are not
                                   * reused by the other.                                            */

typedef struct pnb_ctx   pnb_ctx;
typedef struct pnb_model pnb_model;
typedef struct pnb_locus pnb_locus;

/* ---- library ---------------------------------------------------------- */
PNB_API int         pnb_abi_version(void);
PNB_API const char *pnb_last_error(void);

/* Number of CUDA devices visible to this process (0 and PNB_OK when there is none).  Used by the
 * device policy of run_parallel_chains (src/_mcmc_seq.py:4144-4175): chain c -> device c % count. */
PNB_API int         pnb_device_count(int *count_out);

/* ---- context: one per process per device ------------------------------ */
/* device < 0 selects the current CUDA device.                             */
PNB_API int pnb_ctx_create(int device, pnb_ctx **out);
PNB_API int pnb_ctx_destroy(pnb_ctx *ctx);
/* Run all work of this context on an existing CUDA stream (cudaStream_t passed
 * as void*; NULL restores the context's own stream).  Lets the caller order our
 * kernels with its NCCL collective on the same stream.                    */
PNB_API int pnb_ctx_set_stream(pnb_ctx *ctx, void *cuda_stream);
PNB_API int pnb_ctx_synchronize(pnb_ctx *ctx);
/* counters of the last pnb_eval: [0]=site-node updates, [1]=ops, [2]=tiles,
 * [3]=kernel launches, [4]=H2D bytes, [5]=D2H bytes, [6]=jobs served from cache,
 * [7]=internal nodes recomputed                                           */
PNB_API int pnb_ctx_last_stats(pnb_ctx *ctx, int64_t stats[8]);
/* Measurement hooks (bench.py): with profiling on, every pnb_eval brackets its two
 * kernels with CUDA events on the context's stream.  pnb_ctx_last_timings waits for
 * the last evaluation and returns milliseconds: [0]=host tree compilation,
 * [1]=K-a kernel, [2]=pruning kernel, [3]=whole call on the host (launch + wait). */
PNB_API int pnb_ctx_set_profiling(pnb_ctx *ctx, int enable);
PNB_API int pnb_ctx_last_timings(pnb_ctx *ctx, double ms[4]);

/* ---- substitution model ------------------------------------------------ */
/* Replaces SubstitutionModel._build_and_decompose's outputs being consumed by
 * p_matrix (src/_seq_likelihood.py:202-244) and GTR._decompose/expt
 * (src/GTR.py:518-581).  The eigen-system is computed by the HOST exactly as the
 * reference does (np.linalg.eigh) and handed over; the device evaluates
 * P(t) = (U * exp(evals*t)) @ Uinv for every branch.                       */
PNB_API int pnb_model_create_eigen(pnb_ctx *ctx, const double evals[4], const double U[16],
                           const double Uinv[16], const double pi[4], pnb_model **out);
/* Closed form of JC69.p_matrix (src/_seq_likelihood.py:259-268).           */
PNB_API int pnb_model_create_jc69(pnb_ctx *ctx, pnb_model **out);
/* A model whose p_matrix is overridden on the host (user subclass): P matrices
 * are supplied per edge to pnb_eval (P_edge argument).                     */
PNB_API int pnb_model_create_explicit(pnb_ctx *ctx, const double pi[4], pnb_model **out);
PNB_API int pnb_model_destroy(pnb_model *m);

/* K-a: batched P(t).  Replaces SubstitutionModel.p_matrix / JC69.p_matrix
 * (src/_seq_likelihood.py:231-244, :259-268) and GTR.expt (src/GTR.py:559-581)
 * called once per branch.  t[nb] host -> P_out[nb*16] host, row-major 4x4,
 * negative t clamped to 0 (src/_seq_likelihood.py:241-242).                */
PNB_API int pnb_pmatrix_batch(pnb_ctx *ctx, const pnb_model *m, const double *t, int64_t nb,
                      double *P_out);

/* ---- K-d: site-pattern compression (integer, bit-exact) ---------------- */
/* Replaces FelsensteinCalculator.__init__ (src/_seq_likelihood.py:344-367).
 * cols: n x L bytes, row-major (row = taxon in dict order), already upper-cased
 * (the reference keys on ch.upper(), :349).  Patterns are numbered in
 * first-occurrence order; weights[k] = number of columns with pattern k.
 * code_of_byte: 256-entry table byte -> 4-bit tip mask (IUPAC, :111-117;
 * unknown -> 0xF, :135-137).
 * Outputs (caller-allocated): P_out (1), codes_out n x ld_codes bytes (first
 * *P_out entries of each row valid), weights_out[L], first_col_out[L] (first
 * *P_out valid), pattern_of_col_out[L] (may be NULL).
 * use_device != 0 runs the CUDA implementation (ctx required), otherwise the
 * host implementation (ctx may be NULL).  Both give identical integers.    */
PNB_API int pnb_compress(pnb_ctx *ctx, int use_device, const uint8_t *cols, int32_t n, int64_t L,
                 const uint8_t code_of_byte[256], int64_t *P_out, uint8_t *codes_out,
                 int64_t ld_codes, int64_t *weights_out, int64_t *first_col_out,
                 int32_t *pattern_of_col_out);

/* ---- locus: one alignment block resident in HBM ------------------------ */
/* Replaces the state FelsensteinCalculator keeps after construction
 * (_weights, _tip_partials: src/_seq_likelihood.py:358-367).  codes: n_rows x P
 * 4-bit masks with row stride ld_codes; weights: integer multiplicities.   */
PNB_API int pnb_locus_create(pnb_ctx *ctx, int32_t n_rows, int64_t P, const uint8_t *codes,
                     int64_t ld_codes, const int64_t *weights, pnb_locus **out);
PNB_API int pnb_locus_destroy(pnb_locus *loc);
/* Forget every cached partial of this locus (next eval recomputes all).    */
PNB_API int pnb_locus_invalidate(pnb_locus *loc);
/* Copy a committed node partial back to the host (tests / debugging):
 * node = index in the tree given to the last committed pnb_eval; out[P*4]. */
PNB_API int pnb_locus_read_partial(pnb_locus *loc, int32_t node, double *out);

/* ---- evaluation -------------------------------------------------------- */
/* Replaces FelsensteinCalculator.log_likelihood + _postorder_partials
 * (src/_seq_likelihood.py:369-426) for M (locus, gene tree) jobs at once, i.e.
 * the loop body of _SeqLikelihoodEngine.log_likelihood (src/_mcmc_seq.py:746-750)
 * and the candidate loops of the coupled moves (src/_mcmc_seq.py:1882, :1957).
 *
 * Trees are flat arrays, concatenated over jobs (CSR): job j owns nodes
 * node_off[j] .. node_off[j+1]-1, indexed locally from 0:
 *   parent[i]  local index of the parent, -1 for the root (exactly one);
 *   blen[i]    length of the edge above node i; NaN = "None" -> 0
 *              (src/_seq_likelihood.py:419-420); the root's is ignored;
 *   tip_row[i] for a leaf (node without children): row of the locus' codes, or
 *              -1 if the label is absent from the alignment -> all ones
 *              (src/_seq_likelihood.py:409-412); ignored for internal nodes.
 * Any out-degree >= 1 is accepted (src/_seq_likelihood.py:415-426).
 * P_edge: NULL, or (explicit models) total_nodes*16 doubles, P of the edge
 * above each node, already including site_rate.
 * logL_out[M]: host doubles (device doubles with PNB_EVAL_OUT_DEVICE).
 *
 * Partials cache: every locus keeps the per-node partials of its last
 * COMMITTED tree in HBM.  An evaluation recomputes only internal nodes whose
 * (children, branch lengths) differ from a committed node -- for a one-branch
 * MCMC move that is the path to the root.  Without PNB_EVAL_PENDING the new
 * tree is committed immediately; with it the caller decides (accept -> commit,
 * reject -> rollback), mirroring clone_caches/restore_caches
 * (src/_mcmc_seq.py:766-772).  A locus may appear more than once in a batch
 * only with PNB_EVAL_NOSTORE.                                              */
PNB_API int pnb_eval(pnb_ctx *ctx, const pnb_model *m, int64_t M, pnb_locus *const *loci,
             const int64_t *node_off, const int32_t *parent, const double *blen,
             const int32_t *tip_row, const double *P_edge, double site_rate,
             uint32_t flags, double *logL_out);
PNB_API int pnb_commit(pnb_ctx *ctx, int64_t M, pnb_locus *const *loci);
PNB_API int pnb_rollback(pnb_ctx *ctx, int64_t M, pnb_locus *const *loci);

/* ---- resident batches: fixed (locus, topology) jobs, lengths-only evaluations ------------------- */
/* The steady state of a sharded full evaluation (BASELINE cfg5; the loop body of
 * _SeqLikelihoodEngine.log_likelihood, src/_mcmc_seq.py:746-750, over ALL loci with unchanged topologies):
 * pnb_batch_create compiles the M jobs once (same arrays as pnb_eval, no lengths) and keeps op programs,
 * job descriptors and the tile map in HBM; pnb_batch_eval then takes ONLY the branch lengths, in the node
 * order of the creation arrays (total_nodes doubles; NaN = None -> 0, src/_seq_likelihood.py:419-420):
 *   - host pointer (pinned memory is uploaded asynchronously in chunks that overlap the pruning of the
 *     previous chunk; pageable memory is staged through the batch's pinned buffer, pnb_batch_blen_host), or
 *   - with PNB_BATCH_BLEN_DEVICE a device pointer that is read in place,
 * computes P(t) per edge on the device and launches the pruning kernel.  No tree is compiled and no per-locus
 * host state is touched per evaluation.  Results are bit-identical to pnb_eval(PNB_EVAL_FULL) on the same
 * trees.  Creation flags: PNB_EVAL_NOSTORE (score only) or 0 (every evaluation is committed: the loci's
 * cached partials describe the last evaluated lengths, so a later dirty-path pnb_eval on any of them starts
 * from there), optionally PNB_EVAL_RESCALE.  Evaluation flags: PNB_EVAL_OUT_DEVICE (logL_out[M] is a
 * device pointer, e.g. this rank's slice of the NCCL buffer; no host synchronisation) and
 * PNB_BATCH_BLEN_DEVICE.  A batch whose locus is destroyed or re-allocated answers PNB_ERR_STATE: create
 * it again.  Explicit (host P) models are not supported here.                                        */
#define PNB_BATCH_BLEN_DEVICE 0x20u
typedef struct pnb_batch pnb_batch;
PNB_API int pnb_batch_create(pnb_ctx *ctx, const pnb_model *m, int64_t M, pnb_locus *const *loci,
                             const int64_t *node_off, const int32_t *parent, const int32_t *tip_row,
                             uint32_t flags, pnb_batch **out);
PNB_API int pnb_batch_destroy(pnb_batch *b);
/* info: [0] jobs, [1] nodes, [2] ops, [3] tiles, [4] site-node updates per evaluation, [5] internal nodes,
 * [6] dynamic shared memory per block, [7] upload chunks                                              */
PNB_API int pnb_batch_info(const pnb_batch *b, int64_t info[8]);
/* Bytes one evaluation moves, counted from the compiled programs: [0] node partials written to HBM (stored
 * nodes x padded patterns x 32 B), [1] node partials read back (siblings written earlier in the same launch:
 * normally L2 hits), [2] tip-code bytes, [3] weights + per-tile staging of programs and P matrices (L2).
 * bench.py reports [0]+[2] (+ weights) as the kernel's compulsory DRAM traffic next to ncu's measured figure. */
PNB_API int pnb_batch_traffic(const pnb_batch *b, int64_t bytes[4]);
/* The batch's pinned staging buffer (total_nodes doubles): lengths written here are uploaded without a copy. */
PNB_API int pnb_batch_blen_host(pnb_batch *b, double **pinned_out);
PNB_API int pnb_batch_eval(pnb_batch *b, const pnb_model *m, const double *blen, double site_rate,
                           uint32_t flags, double *logL_out);
/* Fused collective for loci sharded over the GPUs of one NVSwitch domain (replaces the all-reduce / all-gather
 * that would follow the kernel): with PNB_EVAL_OUT_DEVICE the block that completes a job stores its logL into
 * logL_out[j] AND into peer_out[r][j] of every peer -- peer_out[r] is a device pointer, valid in THIS process, to
 * the same slot of peer r's vector (CUDA IPC / symmetric memory; the caller maps it).  8 bytes per job and peer
 * over NVLink, no collective kernel; the caller orders consumers with one cross-GPU barrier.  n_peers = 0 turns
 * it off.  At most 7 peers.                                                                              */
PNB_API int pnb_batch_set_peer_outputs(pnb_batch *b, int32_t n_peers, double *const *peer_out);
/* The cross-GPU barrier of that fused collective, issued by the kernel itself: the block that finishes the LAST job of
 * the next evaluation publishes `epoch` in peer_flag[r] -- the slot [this rank] of peer r's flag array, peer-mapped
 * like peer_out -- and waits until my_flags[peer_rank[r]] >= epoch for every peer, so when the kernel has completed on
 * this rank every peer's stores into this rank's vector have landed.  Flag arrays are uint64[world] in the same
 * symmetric allocation as the vectors, zero at start; the caller owns the epoch counter (1, 2, 3, ... in the order all
 * ranks run their sharded evaluations) and arms the barrier before EACH evaluation (it is used once).  A peer that
 * does not arrive within ~4 s sets a flag (pnb_batch_barrier_timed_out) instead of hanging the kernel.
 * Replaces: the NCCL collective after the kernel (src/_mcmc_seq.py:746-761 is a plain sum over loci).      */
PNB_API int pnb_batch_set_peer_barrier(pnb_batch *b, int32_t n_peers, uint64_t *const *peer_flag,
                                       const uint64_t *my_flags, const int32_t *peer_rank, uint64_t epoch);
PNB_API int pnb_batch_barrier_timed_out(const pnb_batch *b, int32_t *timed_out);

/* ---- K-m: timed MSNC density log P(g | Psi) ---------------------- */
/* The other factor of the MCMC_SEQ likelihood, sum_i [log P(S_i|g_i) + log P(g_i|Psi)]
 * (src/_mcmc_seq.py:745-761).  A species network handle replaces _NetworkIndex +
 * build_network_msnc_index (src/_msnc_density.py:838-935, :392-417): nodes 0..n_nodes-1, edges
 * parent -> child in the reference's E() order (the in-edge order of a reticulation decides which
 * gamma goes with which parent, :311-322), edge_gamma NaN where the reference has None (filled in as
 * there: both absent -> 0.5 / 0.5, one absent -> complement), node_height = height above the present
 * (_node_height, src/GraphUtils.py:786-819), node_species[v] = species id (0..n_species-1) of leaf
 * v, -1 for internal nodes.  A node with two parents is a reticulation.  Any context will do;
 * on a host-only context the network is only compiled (pnb_msnc_net_info).                     */
typedef struct pnb_msnc_net pnb_msnc_net;
PNB_API int pnb_msnc_net_create(pnb_ctx *ctx, int32_t n_nodes, int32_t n_edges, const int32_t *edge_src,
                                const int32_t *edge_dst, const double *edge_gamma,
                                const double *node_height, const int32_t *node_species,
                                int32_t n_species, pnb_msnc_net **out);
PNB_API int pnb_msnc_net_destroy(pnb_msnc_net *net);
/* info: [0] open-edge slots per frontier entry, [1] steps (= nodes), [2] reticulations, [3] species */
PNB_API int pnb_msnc_net_info(const pnb_msnc_net *net, int32_t info[4]);
/* Fixed per-population thetas (the reference's branch_thetas, resolve_branch_theta src/_units.py:90-125, with
 * the label keys already turned into edge ids by the caller): edge_theta[e] replaces the call's theta on the
 * branch of edge e, root_theta on the branch above the root; NaN = no override.  edge_theta NULL clears all. */
PNB_API int pnb_msnc_net_set_branch_thetas(pnb_msnc_net *net, const double *edge_theta, double root_theta);
/* Frontier entries a single gene tree may need before pnb_msnc_eval gives up with
 * PNB_ERR_UNSUPPORTED (default 65536; the scratch is sized from the network and the batch, this is
 * only the ceiling).                                                                          */
PNB_API int pnb_msnc_net_set_max_frontier(pnb_msnc_net *net, int64_t entries);
/* Batched msnc_log_density_prebuilt (src/_msnc_density.py:465-491) for n_trees gene trees against one
 * network: the per-locus loop of _SeqLikelihoodEngine.log_likelihood (src/_mcmc_seq.py:750-760) and
 * the per-candidate loop of the coupled moves (:1935-1957) in one call.  Trees are CSR-packed like
 * pnb_eval's (tree_off, parent, blen; NaN length = None = 0); node_species[i] = species id of the
 * species a LEAF's allele maps to (species_of, :262-266), -1 if unmapped; ignored for internal nodes.
 * Coalescence times are node heights (max over children of child height + length).  Gene trees of
 * up to 128 nodes.  logp_out[j] = log P(g_j | Psi), -inf when g_j does not embed (:488-490),
 * 0 for an empty tree (:222-223).                                                             */
PNB_API int pnb_msnc_eval(pnb_ctx *ctx, const pnb_msnc_net *net, int64_t n_trees, const int64_t *tree_off,
                          const int32_t *parent, const double *blen, const int32_t *node_species,
                          double theta, double *logp_out);

/* One gene-tree move of the chain (op_gene_node_height / op_gene_tree_nni, src/_mcmc_seq.py:1020-1113) dirties ONE locus:
 * _SeqLikelihoodEngine.log_likelihood then re-scores both of its factors (:746-760).  pnb_eval_move does both with one
 * host synchronisation: pnb_eval of the single (locus, tree) job (flags: PNB_EVAL_PENDING / PNB_EVAL_FULL /
 * PNB_EVAL_RESCALE) and pnb_msnc_eval of the same tree are queued back to back; out[0] = log P(S | g),
 * out[1] = log P(g | Psi).  node_species as for pnb_msnc_eval.                                                   */
PNB_API int pnb_eval_move(pnb_ctx *ctx, const pnb_model *m, pnb_locus *locus, int32_t n_nodes, const int32_t *parent,
                          const double *blen, const int32_t *tip_row, double site_rate, uint32_t flags,
                          const pnb_msnc_net *net, const int32_t *node_species, double theta, double out[2]);

/* ---- candidate sets of the coupled moves (host) -------------------------------- */
/* Replaces _candidate_set / _gene_tree_nni_neighbors / _scaled_gene_tree (src/_mcmc_seq.py:1724-1861) for MANY gene
 * trees at once, writing straight into the CSR arrays pnb_eval / pnb_msnc_eval take.  The candidate set of a tree
 * is: the tree, then its height-preserving NNI neighbours (node v ascending, then v's children ascending: child c
 * and v's only sibling s trade places when height[v] > height[s]), each at every height scale, in that order;
 * member (tree, scale 1) keeps the tree's branch lengths, every other member gets max(h[parent] - h[child], 0) on
 * the scaled heights (_sync_lengths, :418-434).  Trees are CSR-packed (tree_off, parent) with node heights.
 * Two calls: _count gives the members per tree; the caller sizes the outputs (member_off = prefix sum; rows are
 * members x nodes of their tree, concatenated) and _fill writes them.  changed_node_out / scale_out (may be NULL):
 * per member the node whose clade the NNI changed (-1 for the tree itself) and the scale.  ctx may be NULL or
 * host-only (its worker pool, if any, is used).  Up to 128 nodes per tree, up to 16 scales.                    */
PNB_API int pnb_candidates_count(pnb_ctx *ctx, int64_t n_trees, const int64_t *tree_off, const int32_t *parent,
                                 const double *height, int32_t n_scales, const double *scales,
                                 int64_t *n_members_out);
PNB_API int pnb_candidates_fill(pnb_ctx *ctx, int64_t n_trees, const int64_t *tree_off, const int32_t *parent,
                                const double *blen, const double *height, int32_t n_scales, const double *scales,
                                const int64_t *member_off, int32_t *parent_out, double *blen_out,
                                double *height_out, int32_t *changed_node_out, double *scale_out);

/* ---- K-s: synthetic alignments on the device (input generator) -------------------- */
/* Replaces simulate_sequences / _evolve (src/_sim_seq.py:431-504) for M gene trees at once: a root state from pi,
 * every child state from the row P[parent state] of its branch, L sites per tree, one thread per (tree, site).
 * Trees are CSR-packed like pnb_eval's; P_edge: total_nodes x 16 host doubles, row-major P of the edge above each
 * node (the root's is ignored) -- supplied by the caller so that host and device compare against identical
 * thresholds.  Draws come from a counter-based generator keyed by (seed, stream_id[j], site, node): a tree's
 * alignment depends on nothing else (not on the batch, the GPU count or the device; phynetpy_b200/simulate.py
 * reproduces the same bytes in NumPy).  out: host bytes, tree j at out_off[j] (written, M + 1 entries): one row of
 * L ASCII characters A/C/G/T per leaf, leaves in node order.  Up to 512 nodes per tree.                      */
PNB_API int pnb_simulate_alignments(pnb_ctx *ctx, int64_t M, const int64_t *node_off, const int32_t *parent,
                                    const double *P_edge, const double pi[4], int64_t L, uint64_t seed,
                                    const int64_t *stream_id, uint8_t *out, int64_t *out_off);

/* ---- host-only planning (tests of the host runtime; no device involved) -- */
/* A planning context owns no GPU: loci created on it hold no device memory and
 * pnb_eval on it fails.  pnb_plan runs exactly the host half of pnb_eval -- cache
 * resolution against the committed tree, op-program compilation, slot assignment,
 * commit / pending bookkeeping (same flags) -- and returns the programs instead of
 * launching them, so the CPU test-suite can interpret them against the oracle.
 * job_info[j*8..]: n_ops, n_tip_ops, root_kind (0 TIP / 1 MEM / 2 REG), root_src,
 * stash need, nodes recomputed, staged code rows per segment, 1 if served from cache.
 * ops_out: int32[4] per op (flags, src, dst, aux), job j at offset
 * (node_off[j] - node_off[0] - j); t_out: effective branch length per op.     */
PNB_API int pnb_ctx_create_host(pnb_ctx **out);
PNB_API int pnb_plan(pnb_ctx *ctx, const pnb_model *m, int64_t M, pnb_locus *const *loci,
                     const int64_t *node_off, const int32_t *parent, const double *blen,
                     const int32_t *tip_row, double site_rate, uint32_t flags,
                     int32_t *job_info, int32_t *ops_out, double *t_out);

#ifdef __cplusplus
}
#endif
#endif /* PHYNET_B200_H */
