// pnb_device.cuh -- device-side data structures and sm_100a kernels for the
// sequence-likelihood hot path (Felsenstein pruning).  See DESIGN.md for the
// layout in HBM and the roofline of each kernel.
//
//   K-a  pnb_pmatrix_kernel   batched P(t) per branch
//        (reference: SubstitutionModel.p_matrix  src/_seq_likelihood.py:231-244,
//                    JC69.p_matrix               src/_seq_likelihood.py:259-268,
//                    GTR.expt                    src/GTR.py:559-581)
//   K-b/K-c pnb_prune_kernel_t<FUSED, RESCALE>  post-order partial updates + fused
//        root/log/weight reduction; the same kernel runs a full evaluation (all
//        internal nodes) and a dirty-path evaluation (only nodes whose subtree changed)
//        (reference: FelsensteinCalculator._postorder_partials :399-426 and
//                    log_likelihood :391-397)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pnb {

#ifndef PNB_TPB
#define PNB_TPB 128
#endif
#ifndef PNB_PPT
#define PNB_PPT 4
#endif
#ifndef PNB_MIN_BLOCKS
#define PNB_MIN_BLOCKS 4
#endif
constexpr int TPB = PNB_TPB;        // threads per block
constexpr int PPT = PNB_PPT;        // site patterns per thread (independent dependency chains)
constexpr int MAX_PEERS = 7;         // other GPUs of one NVSwitch domain
constexpr int SEG_MAX_OPS = 192;     // a tree program longer than this is streamed through shared memory in segments
constexpr int TILE_PATTERNS = TPB * PPT;  // patterns per tile; FIXED so the summation order of a locus
                                          // never depends on batch composition or GPU count
constexpr int PRUNE_MIN_BLOCKS = PNB_MIN_BLOCKS;  // occupancy target of the pruning kernel (register budget)

// child-contribution source kinds
enum : uint32_t { K_TIP = 0, K_MEM = 1, K_REG = 2, K_STK = 3 };
// op flag bits
constexpr uint32_t F_KIND_MASK = 0x3u;
constexpr uint32_t F_FIRST = 0x4u;   // first child of its node: acc  = v
constexpr uint32_t F_LAST  = 0x8u;   // last child of its node: `acc` is now the complete node partial
constexpr uint32_t F_STORE = 0x10u;  // on F_LAST: write the node partial to HBM slot `dst`
constexpr uint32_t F_PUSH  = 0x20u;  // on F_LAST: stash the node partial in smem stack[(flags>>8)&0xff]
constexpr uint32_t F_HOLD  = 0x40u;  // on F_LAST: keep the node partial in the HOLD registers (read back by K_STK with src < 0)
#ifdef PNB_HOLD
constexpr bool HOLD_REGS = true;     // the kernel has a second register set for ONE waiting sibling (host compiler: hold_reg)
#else
constexpr bool HOLD_REGS = false;
#endif

// One op = one (parent <- child) edge: v = P(edge) . L(child); acc (*)= v.
// Op k of a job uses P matrix k of that job (one-to-one), so the program and
// its P matrices are two contiguous arrays staged by two TMA bulk copies.  A K_REG op consumes the node
// completed just before it and is always the FIRST op of its parent (the host compiler guarantees it).
struct __align__(16) Op {
    uint32_t flags;
    int32_t src;   // TIP: row in the locus code matrix (-1: taxon absent -> all ones)
                   // MEM: partial slot;  STK: stack index
    int32_t dst;   // F_STORE: partial slot to write
    int32_t aux;   // TIP: staging row in shared memory
};

struct __align__(16) JobDesc {
    const uint8_t* codes;   // [n_rows][ld_codes] device tip codes: state j (0..3), or 0x80 | 4-bit mask
    double* partials;       // [n_slots][P_pad][4] per-node partials of this locus
    const double* weights;  // [P_pad] pattern multiplicities (integer valued)
    int64_t slot_stride;    // doubles per slot (= P_pad * 4)
    int32_t P;              // site patterns
    int32_t ld_codes;       // row stride of codes (= P_pad, multiple of TILE_PATTERNS)
    int32_t op_off;         // first op / P matrix of this job in the batch arrays
    int32_t n_ops;
    int32_t n_tip_ops;      // TIP ops with src >= 0 (informational: the producer warp counts per segment)
    int32_t root_kind;      // K_REG: root partial is `acc` after the last op; K_MEM / K_TIP: root_src
    int32_t root_src;
    int32_t tile0;          // first tile of this job
    int32_t n_tiles;
    int32_t out_index;      // where the job's logL goes in the output vector
    const Op* ops;          // the job's op program: in the batch staging area, or the locus' resident copy
    int32_t* scale;         // RESCALE launches: [n_slots][P_pad] binary exponent carried by each stored partial
    int64_t pad0;
};

struct ModelParams {
    double evals[4];
    double U[16];
    double Uinv[16];
    double pi[4];
    int32_t kind;  // 0 eigen, 1 JC69 closed form, 2 explicit (P supplied by the host)
    int32_t pad;
};

struct PruneLaunch {
    const JobDesc* jobs;
    const int32_t* tile_job;   // tile -> job (batch index), -1 for tiles of jobs served from the cache
    const double* Pm;          // batch P matrices [total_ops][16]
    const double* t;           // FUSED launches: effective branch length per op (P(t) is computed in-kernel)
    double* tile_sum;          // [n_tiles][2]: one sum per HALF tile (patterns u < PPT/2, u >= PPT/2 of every thread)
    unsigned int* job_done;    // [batch jobs] arrival counters (self-resetting)
    double* out;               // per-job logL
    int32_t ops_cap;           // smem capacity in ops
    int32_t rows_cap;          // smem capacity in staged code rows
    int32_t stk_cap;           // smem stash entries
    int32_t tile_base;         // first tile of this launch (a batch may be launched in chunks)
    int32_t n_full;            // blocks [0, n_full) take one whole tile each; every later PAIR of blocks shares one tile,
                               // half of its patterns each (the last, partial wave of a launch runs as twice as many
                               // blocks of half the duration; launch_prune decides)
    // fused collective (multi-GPU resident batches): the block that completes a job also stores its logL into the
    // same slot of every peer's vector -- peer-mapped device pointers (NVLink / NVSwitch), 8 bytes per peer and job
    int32_t n_peers;
    double* peer_out[MAX_PEERS];
    // ... and the barrier that orders the peers' consumers after those stores, issued by the block that finishes the
    // evaluation's LAST job (bar_epoch != 0): it publishes bar_epoch in slot [this rank] of every peer's flag array and
    // waits until every peer's epoch has arrived in its own -- no separate barrier launch after the kernel
    int32_t launch_jobs;               // jobs of the whole evaluation that have tiles (it may be launched in chunks)
    int32_t out_len;                   // values in this rank's slice of the vector (out[0 .. out_len))
    unsigned int* launch_done;         // finished jobs of the evaluation (self-resetting)
    unsigned long long bar_epoch;
    unsigned long long* bar_peer[MAX_PEERS];     // slot [this rank] of peer r's flag array (peer-mapped)
    const unsigned long long* bar_mine;          // this rank's flag array, indexed by rank
    int32_t bar_rank[MAX_PEERS];                 // rank of peer r
    int32_t* bar_timeout;                        // mapped host word: set when a peer did not arrive within ~4 s
};

// FUSED launches of one or two loci (the single-locus MCMC step): the tile map, the job descriptors, the op program and
// the branch lengths travel IN THE KERNEL PARAMETERS instead of being read from pinned host memory -- four dependent
// PCIe round trips (~1.5 us each) less on the critical path of the step.
constexpr int SMALL_MAX_TILES = 8, SMALL_MAX_JOBS = 2, SMALL_MAX_OPS = 64;
struct SmallBatch {
    int32_t n_jobs;   // 0: not used (everything is read through PruneLaunch)
    int32_t pad;
    int32_t tile_job[SMALL_MAX_TILES];
    JobDesc jobs[SMALL_MAX_JOBS];
    Op ops[SMALL_MAX_OPS];        // JobDesc.ops == NULL: the job's program is ops[op_off ...]
    double t[SMALL_MAX_OPS];
};

#ifdef PNB_DEVICE_KERNELS  // kernel bodies: compiled into exactly one translation unit (pnb_eval.cu)

// ------------------------------------------------------------------------
// PTX helpers: 256-bit global access, mbarrier, TMA bulk copy (UBLKCP)
// ------------------------------------------------------------------------
__device__ __forceinline__ void ldg256(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}
__device__ __forceinline__ void stg256(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// 1-D TMA bulk copy global -> shared, completion counted on an mbarrier.
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------
// K-a: one thread per branch.
// ------------------------------------------------------------------------
__device__ __forceinline__ void pmatrix_of_t(const ModelParams& mp, double t, double* Pm) {
    if (!(t >= 0.0)) t = 0.0;  // reference clamps t < 0 -> 0 (:241-242); NaN ("None") -> 0 (:420)
    if (mp.kind == 1) {
        // JC69 closed form, same expression order as the reference (:263-265)
        const double e = exp(-4.0 * t / 3.0);
        const double same = 0.25 + 0.75 * e;
        const double diff = 0.25 - 0.25 * e;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) Pm[i * 4 + j] = (i == j) ? same : diff;
        return;
    }
    double ex[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) ex[k] = exp(mp.evals[k] * t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = mp.U[i * 4 + k] * ex[k];  // (U * expd[None,:])  (:244)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double s = a[0] * mp.Uinv[0 * 4 + j];
            s = fma(a[1], mp.Uinv[1 * 4 + j], s);
            s = fma(a[2], mp.Uinv[2 * 4 + j], s);
            s = fma(a[3], mp.Uinv[3 * 4 + j], s);
            Pm[i * 4 + j] = s;
        }
    }
}

// The pruning kernel reads P COLUMN by column (a tip child selects one column of P by its state; an internal
// child's contribution is accumulated column by column, P[:,j] * l_j): K-a writes P transposed for it, so a
// column is 32 contiguous bytes (two LDS.128).  TRANSPOSED = false is the public row-major layout
// (pnb_pmatrix_batch: P[i*4+j] = prob(i -> j)).
template <bool TRANSPOSED>
__device__ __forceinline__ void store_pmatrix(double* o, const double (&Pm)[16]) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        if (TRANSPOSED) stg256(o + a * 4, Pm[0 * 4 + a], Pm[1 * 4 + a], Pm[2 * 4 + a], Pm[3 * 4 + a]);
        else stg256(o + a * 4, Pm[a * 4 + 0], Pm[a * 4 + 1], Pm[a * 4 + 2], Pm[a * 4 + 3]);
    }
}

template <bool TRANSPOSED>
__global__ void __launch_bounds__(128)
pnb_pmatrix_kernel(const __grid_constant__ ModelParams mp, const double* __restrict__ t,
                   int64_t nb, double* __restrict__ P_out) {
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= nb) return;
    double Pm[16];
    pmatrix_of_t(mp, t[g], Pm);
    store_pmatrix<TRANSPOSED>(P_out + g * 16, Pm);
}

// K-a for resident batches: op k takes the length of tree node node_of_op[k] from a node-ordered array
// that lives in HBM (the caller's, or the batch's upload buffer); same arithmetic as the host's
// eff_len (None/NaN -> 0, * site_rate, clamp) followed by pmatrix_of_t, so P is bit-identical to the
// pnb_eval path.  keep != NULL: remember the raw length per node (what a later locus_sync reads).
__global__ void __launch_bounds__(128)
pnb_pmatrix_gather_kernel(const __grid_constant__ ModelParams mp, const double* __restrict__ blen,
                          const int32_t* __restrict__ node_of_op, double site_rate, int64_t n_ops,
                          double* __restrict__ P_out, double* __restrict__ keep) {
    // programmatic dependent launch: the pruning kernel queued behind this one may start its prologue (descriptors, program
    // and tip-code staging -- nothing this kernel writes) now; it waits for P(t) with griddepcontrol.wait
    // (this kernel is itself a dependent launch behind whatever precedes it on the stream -- the previous evaluation's
    // pruning kernel, which still reads the P array this one rewrites: it waits for that kernel's completion first; what
    // the early launch saves is the launch latency)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= n_ops) return;
    const int32_t node = node_of_op[g];
    const double raw = blen[node];
    if (keep) keep[node] = raw;
    double t = raw;
    if (t != t) t = 0.0;
    t *= site_rate;
    double Pm[16];
    pmatrix_of_t(mp, t, Pm);
    store_pmatrix<true>(P_out + g * 16, Pm);
}

// ------------------------------------------------------------------------
// K-b / K-c: pruning.  grid = tiles, block = TPB threads, PPT patterns per thread.
//
// A thread owns PPT patterns of the tile (tid, tid + TPB, ...) and walks the
// job's op program once for all of them: the program, its P matrices and the
// loop control are shared, and the PPT independent dependency chains give the
// FP64 pipe and the LSU something to overlap.
//
// Dynamic shared memory layout (bytes):
//   [0, ops_cap*16)                 Op program of the job
//   [.., + ops_cap*128)             P matrices, one per op (row-major 4x4 f64)
//   [.., + rows_cap*TILE)           tip-code rows of the tile (one byte per pattern)
//   [.., + stk_cap*TILE*32)         (NOSTORE only) partial stash stack
// ------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// fixed-order block reduction; result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < TPB / 32; ++i) s += red[i];
    }
    __syncthreads();
    return s;
}

struct Vec4 { double x0, x1, x2, x3; };

// Shared-memory accesses through explicit 32-bit shared addresses (ld.shared / st.shared): the
// program walk is a chain of small dependent loads, and generic-pointer arithmetic was a third of
// the instruction stream in the first version of this kernel (profiles/prune_r1_v2_cfg2_ncu.json).
template <int OFF>
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ void lds_v2f64(uint32_t a, double& x, double& y) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+%3];" : "=d"(x), "=d"(y) : "r"(a), "n"(OFF));
}
__device__ __forceinline__ void sts_v2f64(uint32_t a, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" :: "r"(a), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void lds_op(uint32_t a, uint32_t& f, int32_t& src, int32_t& dst, int32_t& aux) {
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(f), "=r"(src), "=r"(dst), "=r"(aux) : "r"(a));
}

// v = P . l   (child_partials @ P.T, reference :425).  P is staged TRANSPOSED: column j of P (the 4 values
// P[0..3][j]) is 32 contiguous bytes at pk + 32 j, read by every lane at the same address (broadcast).  The
// accumulation runs over columns, r_i = ((P_i0 l_0 + P_i1 l_1) + P_i2 l_2) + P_i3 l_3 -- the same expression
// and rounding as a row-wise dot product.  ACCUM = multiply the result into `out` instead of overwriting it;
// `out` may alias `l` only when !ACCUM (the register-forwarded child).
template <bool ACCUM, int NP>
__device__ __forceinline__ void matvec(uint32_t pk, const Vec4 (&l)[NP], Vec4 (&out)[NP]) {
    double r[NP][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double c0, c1, c2, c3;
        if (j == 0) { lds_v2f64<0>(pk, c0, c1); lds_v2f64<16>(pk, c2, c3); }
        else if (j == 1) { lds_v2f64<32>(pk, c0, c1); lds_v2f64<48>(pk, c2, c3); }
        else if (j == 2) { lds_v2f64<64>(pk, c0, c1); lds_v2f64<80>(pk, c2, c3); }
        else { lds_v2f64<96>(pk, c0, c1); lds_v2f64<112>(pk, c2, c3); }
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            const double lj = (j == 0) ? l[u].x0 : (j == 1) ? l[u].x1 : (j == 2) ? l[u].x2 : l[u].x3;
            if (j == 0) {
                r[u][0] = c0 * lj; r[u][1] = c1 * lj; r[u][2] = c2 * lj; r[u][3] = c3 * lj;
            } else {
                r[u][0] = fma(c0, lj, r[u][0]); r[u][1] = fma(c1, lj, r[u][1]);
                r[u][2] = fma(c2, lj, r[u][2]); r[u][3] = fma(c3, lj, r[u][3]);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < NP; ++u) {
        if (ACCUM) {
            out[u].x0 *= r[u][0]; out[u].x1 *= r[u][1]; out[u].x2 *= r[u][2]; out[u].x3 *= r[u][3];
        } else {
            out[u].x0 = r[u][0]; out[u].x1 = r[u][1]; out[u].x2 = r[u][2]; out[u].x3 = r[u][3];
        }
    }
}

// contribution of a tip child: column(s) of P selected by the device tip code
//   code < 4        : the single state j = code           -> column j of P
//   code & 0x80     : ambiguity / gap, low 4 bits = mask   -> sum of the allowed columns, in state order
template <bool ACCUM, int NP>
__device__ __forceinline__ void tip_contrib(uint32_t pk, const uint32_t (&code)[NP], Vec4 (&out)[NP]) {
    bool amb = false;
#pragma unroll
    for (int u = 0; u < NP; ++u) amb |= (code[u] & 0x80u) != 0;
    if (!__any_sync(0xffffffffu, amb)) {
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            const uint32_t col = pk + code[u] * 32u;   // column `state` of P: 32 contiguous bytes
            double v0, v1, v2, v3;
            lds_v2f64<0>(col, v0, v1);
            lds_v2f64<16>(col, v2, v3);
            if (ACCUM) { out[u].x0 *= v0; out[u].x1 *= v1; out[u].x2 *= v2; out[u].x3 *= v3; }
            else { out[u].x0 = v0; out[u].x1 = v1; out[u].x2 = v2; out[u].x3 = v3; }
        }
    } else {
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            const uint32_t m = (code[u] & 0x80u) ? (code[u] & 0xFu) : (1u << code[u]);
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (m & (1u << j)) {
                    const uint32_t col = pk + j * 32;
                    s0 += lds_f64<0>(col); s1 += lds_f64<8>(col); s2 += lds_f64<16>(col); s3 += lds_f64<24>(col);
                }
            }
            if (ACCUM) { out[u].x0 *= s0; out[u].x1 *= s1; out[u].x2 *= s2; out[u].x3 *= s3; }
            else { out[u].x0 = s0; out[u].x1 = s1; out[u].x2 = s2; out[u].x3 = s3; }
        }
    }
}

// FUSED = latency path for small batches (a single-locus MCMC step): the kernel reads the job
// descriptors, the program and the branch lengths straight from the caller's pinned staging memory,
// computes P(t) of its ops itself (no K-a launch, no H2D copy) and writes logL to mapped host memory.
// RESCALE = optional numerical mode beyond the reference: every completed node is renormalised by an
// exact power of two and carries the binary exponent of its subtree, so trees of thousands of taxa do
// not underflow (the reference, and the default mode here, floor such sites at 1e-300).
// One block = NP patterns per thread of one tile: NP = PPT (u0 = 0) is a whole tile, NP = PPT / 2 one half of it
// (u0 = 0 or PPT / 2: the thread's patterns tid + (u0 + u) TPB).  Either way a thread's weighted log-likelihoods are
// summed per HALF (u < PPT / 2, u >= PPT / 2) and every half is reduced over the block on its own, so a tile gives the
// same two sums whether one block or two computed it -- results never depend on how a launch was cut into blocks.
template <bool FUSED, bool RESCALE, int NP>
__device__ __forceinline__ void prune_tile(const PruneLaunch& L, const ModelParams& mp, const SmallBatch& S, const int tile,
                                           const int u0) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar_stage;
    __shared__ double red[TPB / 32];
    __shared__ int s_is_last;
    __shared__ int s_final;
    __shared__ int s_rows;

    const int tid = threadIdx.x;
    const bool small = FUSED && S.n_jobs > 0;
    const int job = small ? S.tile_job[tile] : L.tile_job[tile];
    if (job < 0) return;  // tile of a job that needed no device work (served from the cache)
    const JobDesc& jd = small ? S.jobs[job] : L.jobs[job];
    const int n_ops = jd.n_ops;
    const int op_off = jd.op_off;
    const int tile0 = jd.tile0;
    const int P = jd.P;
    const int ld_codes = jd.ld_codes;
    const uint8_t* g_codes = jd.codes;

    Op* s_ops = reinterpret_cast<Op*>(smem_raw);
    double* s_P = reinterpret_cast<double*>(smem_raw + static_cast<size_t>(L.ops_cap) * sizeof(Op));
    unsigned char* s_codes = reinterpret_cast<unsigned char*>(s_P + static_cast<size_t>(L.ops_cap) * 16);
    unsigned char* s_stk = s_codes + static_cast<size_t>(L.rows_cap) * TILE_PATTERNS;

    const int p0 = (tile - tile0) * TILE_PATTERNS;
    const Op* g_ops = jd.ops;

    if (tid == 0) {
        mbar_init(&bar_stage, 1);
        mbar_fence_init();
    }
    __syncthreads();

    // per-thread bookkeeping.  Every array is padded to a multiple of the tile (codes with state 0,
    // partial slots zero-filled), so lanes past the last pattern run the same instruction stream on
    // padding and only the final weighted sum looks at the pattern index.
    const int64_t slot_stride = jd.slot_stride;
    const int pu0 = p0 + u0 * TPB;   // the block's first pattern
    double* part0 = jd.partials + static_cast<int64_t>(pu0 + tid) * 4;
    const bool warp_active = __any_sync(0xffffffffu, (pu0 + tid) < P);
    const uint32_t ops_base = smem_u32(s_ops);
    const uint32_t p_base = smem_u32(s_P);
    const uint32_t codes_addr = smem_u32(s_codes) + u0 * TPB + tid;
    const uint32_t stk_addr = smem_u32(s_stk) + (u0 * TPB + tid) * 16;

    Vec4 acc[NP];  // node under construction; once complete it is also the register-forwarded child
    int acc_e[NP]; // RESCALE: binary exponent carried by `acc` (true partial = acc * 2^acc_e)
#pragma unroll
    for (int u = 0; u < NP; ++u) { acc[u] = {1.0, 1.0, 1.0, 1.0}; acc_e[u] = 0; }
    Vec4 hold[NP];   // HOLD_REGS: one waiting sibling (the innermost: nothing is held while its last sibling is computed)
    int hold_e[NP];
#pragma unroll
    for (int u = 0; u < NP; ++u) { hold[u] = {1.0, 1.0, 1.0, 1.0}; hold_e[u] = 0; }
    int32_t* scale0 = RESCALE ? jd.scale + (pu0 + tid) : nullptr;
    const int64_t scale_stride = slot_stride / 4;  // ints per slot (= P_pad)
    // stash exponents live behind the stash partials: [stk_cap][TILE_PATTERNS] int32
    const uint32_t stk_e_addr = smem_u32(s_stk) + static_cast<uint32_t>(L.stk_cap) * (TILE_PATTERNS * 32) + (u0 * TPB + tid) * 4;

    // The program is walked in segments of at most SEG_MAX_OPS ops (one segment for trees of up to
    // ~97 taxa): warp 0 is the TMA producer of a segment -- its ops, its P matrices and the code rows
    // of its tip ops, all completing on one mbarrier whose phase flips per segment.
    const int seg_cap = (n_ops > SEG_MAX_OPS) ? SEG_MAX_OPS : n_ops;
    uint32_t phase = 0;
    for (int seg = 0; seg < n_ops; seg += seg_cap, phase ^= 1u) {
        const int nseg = min(seg_cap, n_ops - seg);
        if (FUSED) {
            if (tid == 0) s_rows = 0;
            __syncthreads();  // also: every warp is done with the previous segment's shared memory
            int my_rows = 0;
            const bool param_ops = small && g_ops == nullptr;
            for (int k = tid; k < nseg; k += TPB) {
                const Op op = param_ops ? S.ops[op_off + seg + k] : g_ops[seg + k];
                s_ops[k] = op;
                double Pm[16];
                pmatrix_of_t(mp, small ? S.t[op_off + seg + k] : L.t[op_off + seg + k], Pm);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) s_P[k * 16 + j * 4 + i] = Pm[i * 4 + j];   // transposed, as K-a writes it
                if ((op.flags & F_KIND_MASK) == K_TIP && op.src >= 0) {
                    tma_bulk_g2s(s_codes + static_cast<size_t>(op.aux) * TILE_PATTERNS + u0 * TPB,
                                 g_codes + static_cast<int64_t>(op.src) * ld_codes + pu0, NP * TPB, &bar_stage);
                    ++my_rows;
                }
            }
            if (my_rows) atomicAdd(&s_rows, my_rows);
            __syncthreads();
            if (tid == 0) mbar_arrive_expect_tx(&bar_stage, static_cast<uint32_t>(s_rows) * (NP * TPB));
        } else {
        if (seg > 0) __syncthreads();  // every warp is done with the previous segment's shared memory
        if (tid < 32) {
            if (tid == 0) tma_bulk_g2s(s_ops, g_ops + seg, static_cast<uint32_t>(nseg) * sizeof(Op), &bar_stage);
            int my_rows = 0;
            for (int k = tid; k < nseg; k += 32) {
                const Op op = g_ops[seg + k];
                if ((op.flags & F_KIND_MASK) == K_TIP && op.src >= 0) {
                    tma_bulk_g2s(s_codes + static_cast<size_t>(op.aux) * TILE_PATTERNS + u0 * TPB,
                                 g_codes + static_cast<int64_t>(op.src) * ld_codes + pu0, NP * TPB, &bar_stage);
                    ++my_rows;
                }
            }
            const int rows = __reduce_add_sync(0xffffffffu, my_rows);
            if (tid == 0) {
                // P(t) is what the K-a launch in front of this kernel writes: everything above (descriptors, program, tip
                // codes) was staged while K-a may still have been running (programmatic dependent launch, launch_prune);
                // this is where the kernel waits for it.  Without such a launch in front the wait returns at once.
                if (seg == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
                tma_bulk_g2s(s_P, L.Pm + (static_cast<int64_t>(op_off) + seg) * 16, static_cast<uint32_t>(nseg) * 128u,
                             &bar_stage);
            }
            // the pending arrival keeps the phase open until the byte count has been announced
            if (tid == 0)
                mbar_arrive_expect_tx(&bar_stage, static_cast<uint32_t>(nseg) * (sizeof(Op) + 128u) +
                                                      static_cast<uint32_t>(rows) * (NP * TPB));
        }
        }
        mbar_wait(&bar_stage, phase);
        if (!warp_active) continue;

        uint32_t op_addr = ops_base;
        uint32_t pk = p_base;
        uint32_t nflags;
        int32_t nsrc, ndst, naux;
#ifdef PNB_UNIFORM_LD
        { const int4 o4 = __ldg(reinterpret_cast<const int4*>(g_ops + seg)); nflags = o4.x; nsrc = o4.y; ndst = o4.z; naux = o4.w; }
#else
        lds_op(op_addr, nflags, nsrc, ndst, naux);
#endif
        for (int k = 0; k < nseg; ++k, op_addr += 16, pk += 128) {
            // the op one ahead is fetched now: its ~30-cycle shared-memory latency overlaps this op's arithmetic
            // (past the last op this reads the first bytes of the P area, still inside the block's shared memory)
#ifdef PNB_UNIFORM_SHFL
            const uint32_t flags = __shfl_sync(0xffffffffu, nflags, 0);
            const int32_t src = __shfl_sync(0xffffffffu, nsrc, 0), dst = __shfl_sync(0xffffffffu, ndst, 0),
                          aux = __shfl_sync(0xffffffffu, naux, 0);
#else
            uint32_t flags = nflags;
            int32_t dst = ndst;
            const int32_t src = nsrc, aux = naux;
#endif
#ifdef PNB_UNIFORM_LD
            { const int4 o4 = __ldg(reinterpret_cast<const int4*>(g_ops + seg + k + 1)); nflags = o4.x; nsrc = o4.y; ndst = o4.z; naux = o4.w; }
#else
            lds_op(op_addr + 16, nflags, nsrc, ndst, naux);
#endif
            const uint32_t kind = flags & F_KIND_MASK;
            const bool first = (flags & F_FIRST) != 0;
            // A tip child that FOLLOWS inside the same node (the second tip of a cherry, the tip next to the register-forwarded
            // child) is taken in the same iteration: its fields are already in the prefetched registers, so it costs its code
            // bytes, one column of P per pattern and four multiplications -- not another round of op fetch / decode / dispatch /
            // loop control (~50 of a tip op's ~80 instructions).  Same products in the same order, hence the same bits.
            // The loads (the tip's code bytes, the op after it) are issued BEFORE this op's arithmetic, the products after it.
#define PNB_PEEK_TRAILING_TIP()                                                                                      \
            const bool take2 = k + 1 < nseg && (nflags & (F_KIND_MASK | F_FIRST)) == K_TIP;                          \
            uint32_t code2[NP];                                                                                      \
            uint32_t flags2 = 0;                                                                                     \
            int32_t dst2 = 0;                                                                                        \
            if (take2) { /* loads first: the tip's code bytes and the op after it are in flight during this op's arithmetic */ \
                flags2 = nflags;                                                                                     \
                dst2 = ndst;                                                                                         \
                const int32_t src2 = nsrc, aux2 = naux;                                                              \
                lds_op(op_addr + 32, nflags, nsrc, ndst, naux);                                                      \
                _Pragma("unroll") for (int u = 0; u < NP; ++u)                                                       \
                    code2[u] = (src2 >= 0) ? lds_u8(codes_addr + aux2 * TILE_PATTERNS + u * TPB) : 0x8Fu;            \
            }
#define PNB_APPLY_TRAILING_TIP()                                                                                     \
            if (take2) {                                                                                             \
                tip_contrib<true, NP>(pk + 128, code2, acc);                                                         \
                flags = flags2;                                                                                      \
                dst = dst2;                                                                                          \
                ++k;                                                                                                 \
                op_addr += 16;                                                                                       \
                pk += 128;                                                                                           \
            }
            if (kind == K_TIP) {
                uint32_t code[NP];
#pragma unroll
                for (int u = 0; u < NP; ++u)
                    code[u] = (src >= 0) ? lds_u8(codes_addr + aux * TILE_PATTERNS + u * TPB) : 0x8Fu;
#ifndef PNB_NO_TIP_PAIRS
                PNB_PEEK_TRAILING_TIP()
#endif
                if (first) tip_contrib<false, NP>(pk, code, acc);
                else tip_contrib<true, NP>(pk, code, acc);
                if (RESCALE && first) {
#pragma unroll
                    for (int u = 0; u < NP; ++u) acc_e[u] = 0;
                }
#ifndef PNB_NO_TIP_PAIRS
                PNB_APPLY_TRAILING_TIP()
#endif
            } else if (kind == K_REG) {
                // the child is the node completed just before: its partial is still in `acc`
                // (the host always emits the forwarded child as the FIRST op of its parent)
#ifndef PNB_NO_TIP_AFTER_REG
                PNB_PEEK_TRAILING_TIP()
#endif
                matvec<false, NP>(pk, acc, acc);   // reads acc completely before it writes it
#ifndef PNB_NO_TIP_AFTER_REG
                PNB_APPLY_TRAILING_TIP()
#endif
            } else {
                Vec4 l[NP];
                int le[NP];
                if (kind == K_MEM) {
                    const double* g = part0 + static_cast<int64_t>(src) * slot_stride;
#pragma unroll
                    for (int u = 0; u < NP; ++u) ldg256(g + u * TPB * 4, l[u].x0, l[u].x1, l[u].x2, l[u].x3);
                    if (RESCALE) {
                        const int32_t* ge = scale0 + static_cast<int64_t>(src) * scale_stride;
#pragma unroll
                        for (int u = 0; u < NP; ++u) le[u] = ge[u * TPB];
                    }
                } else if (HOLD_REGS && src < 0) {
#pragma unroll
                    for (int u = 0; u < NP; ++u) { l[u] = hold[u]; le[u] = hold_e[u]; }
                } else {
                    const uint32_t a = stk_addr + static_cast<uint32_t>(src) * (TILE_PATTERNS * 32);
#pragma unroll
                    for (int u = 0; u < NP; ++u) {
                        lds_v2f64<0>(a + u * TPB * 16, l[u].x0, l[u].x1);
                        lds_v2f64<TILE_PATTERNS * 16>(a + u * TPB * 16, l[u].x2, l[u].x3);
                    }
                    if (RESCALE) {
                        const uint32_t ae = stk_e_addr + static_cast<uint32_t>(src) * (TILE_PATTERNS * 4);
#pragma unroll
                        for (int u = 0; u < NP; ++u)
                            asm volatile("ld.shared.s32 %0, [%1];" : "=r"(le[u]) : "r"(ae + u * TPB * 4));
                    }
                }
                if (first) matvec<false, NP>(pk, l, acc);
                else matvec<true, NP>(pk, l, acc);
                if (RESCALE) {
#pragma unroll
                    for (int u = 0; u < NP; ++u) acc_e[u] = first ? le[u] : acc_e[u] + le[u];
                }
            }
            if (flags & F_LAST) {
                if (RESCALE) {
#pragma unroll
                    for (int u = 0; u < NP; ++u) {
                        const double mx = fmax(fmax(acc[u].x0, acc[u].x1), fmax(acc[u].x2, acc[u].x3));
                        if (mx > 0.0) {  // exact renormalisation by 2^-e, e = binary exponent of the largest state
                            const int e = ((__double2hiint(mx) >> 20) & 0x7ff) - 1023;
                            const double f = __hiloint2double((1023 - e) << 20, 0);
                            acc[u].x0 *= f; acc[u].x1 *= f; acc[u].x2 *= f; acc[u].x3 *= f;
                            acc_e[u] += e;
                        }
                    }
                }
                if (flags & F_STORE) {
                    double* g = part0 + static_cast<int64_t>(dst) * slot_stride;
#pragma unroll
                    for (int u = 0; u < NP; ++u)
                        stg256(g + u * TPB * 4, acc[u].x0, acc[u].x1, acc[u].x2, acc[u].x3);
                    if (RESCALE) {
                        int32_t* ge = scale0 + static_cast<int64_t>(dst) * scale_stride;
#pragma unroll
                        for (int u = 0; u < NP; ++u) ge[u * TPB] = acc_e[u];
                    }
                }
                if (HOLD_REGS && (flags & F_HOLD)) {
#pragma unroll
                    for (int u = 0; u < NP; ++u) { hold[u] = acc[u]; hold_e[u] = acc_e[u]; }
                }
                if (flags & F_PUSH) {
                    const uint32_t a = stk_addr + ((flags >> 8) & 0xffu) * (TILE_PATTERNS * 32);
#pragma unroll
                    for (int u = 0; u < NP; ++u) {
                        sts_v2f64(a + u * TPB * 16, acc[u].x0, acc[u].x1);
                        sts_v2f64(a + u * TPB * 16 + TILE_PATTERNS * 16, acc[u].x2, acc[u].x3);
                    }
                    if (RESCALE) {
                        const uint32_t ae = stk_e_addr + ((flags >> 8) & 0xffu) * (TILE_PATTERNS * 4);
#pragma unroll
                        for (int u = 0; u < NP; ++u)
                            asm volatile("st.shared.s32 [%0], %1;" :: "r"(ae + u * TPB * 4), "r"(acc_e[u]) : "memory");
                    }
                }
            }
        }
    }

#undef PNB_PEEK_TRAILING_TIP
#undef PNB_APPLY_TRAILING_TIP
    // root: site likelihood, clamp, log, weight  (reference :394-397)
    constexpr int HALF = (PPT % 2 == 0) ? PPT / 2 : PPT;   // patterns per thread and half tile
    constexpr int NH = NP / HALF;                          // halves this block computes (2: a whole tile, 1: one half)
    double lsum[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) lsum[h] = 0.0;
    const int root_kind = jd.root_kind;
    const int root_src = jd.root_src;
    const double* weights = jd.weights;
#pragma unroll
    for (int u = 0; u < NP; ++u) {
        const int pat = pu0 + tid + u * TPB;
        if (pat >= P) continue;
        Vec4 r = acc[u];
        int re = RESCALE ? acc_e[u] : 0;
        if (root_kind == K_MEM) {
            ldg256(part0 + static_cast<int64_t>(root_src) * slot_stride + u * TPB * 4, r.x0, r.x1, r.x2, r.x3);
            if (RESCALE) re = scale0[static_cast<int64_t>(root_src) * scale_stride + u * TPB];
        } else if (root_kind == K_TIP) {
            re = 0;
            const uint32_t c = (root_src >= 0) ? g_codes[static_cast<int64_t>(root_src) * ld_codes + pat] : 0x8Fu;
            const uint32_t m = (c & 0x80u) ? (c & 0xFu) : (1u << c);
            r = {(m & 1u) ? 1.0 : 0.0, (m & 2u) ? 1.0 : 0.0, (m & 4u) ? 1.0 : 0.0, (m & 8u) ? 1.0 : 0.0};
        }
        double site = fma(r.x3, mp.pi[3], fma(r.x2, mp.pi[2], fma(r.x1, mp.pi[1], r.x0 * mp.pi[0])));
        if (RESCALE) {
            // log(site * 2^re); a site that is exactly impossible keeps the reference's floor
            const double ls = (site > 0.0) ? fma(static_cast<double>(re), 0.6931471805599453094, log(site))
                                           : ((site == 0.0) ? -690.77552789821370 : site);
            lsum[u / HALF] = fma(ls, weights[pat], lsum[u / HALF]);
        } else {
            site = (site < 1e-300) ? 1e-300 : site;  // np.maximum(site_like, 1e-300): NaN propagates
            lsum[u / HALF] = fma(log(site), weights[pat], lsum[u / HALF]);
        }
    }

    // fused per-locus reduction: tile sum -> last-arriving tile sums tiles in index order.  (One fence + one ticket per
    // BLOCK: a per-warp version without the block barriers measured 7% slower -- __threadfence invalidates L1.)
    double tsum[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) tsum[h] = block_sum(lsum[h], red);
    const int n_tiles = jd.n_tiles;
    constexpr int HALVES_PER_TILE = PPT / HALF;   // 2 (1 when PPT is odd: tiles are never split)
    if (tid == 0) {
#pragma unroll
        for (int h = 0; h < NH; ++h) __stcg(&L.tile_sum[2 * static_cast<int64_t>(tile) + u0 / HALF + h], tsum[h]);
        if (HALVES_PER_TILE == 1) __stcg(&L.tile_sum[2 * static_cast<int64_t>(tile) + 1], 0.0);
        __threadfence();
        const unsigned int prev = atomicAdd(&L.job_done[job], static_cast<unsigned int>(NH));
        s_is_last = (prev + NH == static_cast<unsigned int>(n_tiles) * HALVES_PER_TILE);
    }
    __syncthreads();
    if (s_is_last) {
        __threadfence();
        double s = 0.0;
        for (int t = tid; t < n_tiles; t += TPB)   // a tile's two halves first, tiles in index order
            s += __ldcg(&L.tile_sum[2 * static_cast<int64_t>(tile0 + t)]) + __ldcg(&L.tile_sum[2 * static_cast<int64_t>(tile0 + t) + 1]);
        const double total = block_sum(s, red);
        if (tid == 0) {
            L.out[jd.out_index] = total;
            L.job_done[job] = 0u;
            s_final = 0;
            if (L.bar_epoch == 0ull) {
                // peers ordered by a barrier the caller launches: every finished job goes to the peers' vectors at once
                for (int r = 0; r < L.n_peers; ++r) L.peer_out[r][jd.out_index] = total;   // stores over NVLink
            } else {
                // release pattern at GPU scope: this job's value happens before its arrival
                __threadfence();
                if (atomicAdd(L.launch_done, 1u) == static_cast<unsigned int>(L.launch_jobs) - 1u) {
                    __threadfence();   // acquire side: every other job's value is visible to this block
                    s_final = 1;
                }
            }
        }
        __syncthreads();
        if (s_final) {
            // The block that sees the evaluation's LAST job arrive is the collective: it pushes this rank's slice of the
            // vector into every peer's copy with coalesced stores over NVLink (out_len x 8 bytes per peer), fences ONCE at
            // system scope, raises its epoch flag at every peer and waits for theirs -- when the kernel has completed on this
            // rank, every peer's slice has landed here.  (Per-job peer stores, as in the branch above, cost the kernel 30 us
            // at 5,000 jobs; a release store per flag would pay the system fence seven times.)
            for (int i0 = 0; i0 < L.out_len; i0 += TPB * 8) {   // 8 independent L2 loads in flight per thread, then the stores
                double v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int i = i0 + q * TPB + tid;
                    v[q] = (i < L.out_len) ? __ldcg(&L.out[i]) : 0.0;
                }
                for (int r = 0; r < L.n_peers; ++r) {
                    double* dst = L.peer_out[r];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int i = i0 + q * TPB + tid;
                        if (i < L.out_len) dst[i] = v[q];
                    }
                }
            }
            __threadfence_system();
            __syncthreads();
            if (tid == 0) {
                *L.launch_done = 0u;
                for (int r = 0; r < L.n_peers; ++r)
                    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(L.bar_peer[r]), "l"(L.bar_epoch) : "memory");
                const long long t0 = clock64();
                for (int r = 0; r < L.n_peers; ++r) {
                    const unsigned long long* f = L.bar_mine + L.bar_rank[r];
                    unsigned long long seen;
                    do {
                        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(f) : "memory");
                        if (seen < L.bar_epoch && clock64() - t0 > 8000000000ll) {   // ~4 s: a peer is gone; say so, do not hang
                            *L.bar_timeout = 1;
                            seen = L.bar_epoch;
                        }
                    } while (seen < L.bar_epoch);
                }
                __threadfence_system();   // acquire side: what the peers pushed before their flags is visible to what follows
            }
        }
    }
}

template <bool FUSED, bool RESCALE>
__global__ void __launch_bounds__(TPB, PRUNE_MIN_BLOCKS)
pnb_prune_kernel_t(const __grid_constant__ PruneLaunch L, const __grid_constant__ ModelParams mp,
                   const __grid_constant__ SmallBatch S) {
    const int b = blockIdx.x;
    if (PPT % 2 != 0 || b < L.n_full) {
        prune_tile<FUSED, RESCALE, PPT>(L, mp, S, L.tile_base + b, 0);
    } else {
        constexpr int HALF = (PPT % 2 == 0) ? PPT / 2 : PPT;
        const int h = b - L.n_full;
        prune_tile<FUSED, RESCALE, HALF>(L, mp, S, L.tile_base + L.n_full + (h >> 1), (h & 1) * HALF);
    }
}

#endif  // PNB_DEVICE_KERNELS

}  // namespace pnb
