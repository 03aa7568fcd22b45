// pnb_device.cuh -- device-side data structures and sm_100a kernels for the
// sequence-likelihood hot path (Felsenstein pruning).  See DESIGN.md for the
// layout in HBM and the roofline of each kernel.
//
//   K-a  pnb_pmatrix_kernel   batched P(t) per branch
//        (reference: SubstitutionModel.p_matrix  src/_seq_likelihood.py:231-244,
//                    JC69.p_matrix               src/_seq_likelihood.py:259-268,
//                    GTR.expt                    src/GTR.py:559-581)
//   K-b/K-c pnb_prune_kernel  post-order partial updates + fused root/log/weight
//        reduction; the same kernel runs a full evaluation (all internal nodes)
//        and a dirty-path evaluation (only nodes whose subtree changed)
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
#define PNB_PPT 2
#endif
#ifndef PNB_MIN_BLOCKS
#define PNB_MIN_BLOCKS 5
#endif
constexpr int TPB = PNB_TPB;        // threads per block
constexpr int PPT = PNB_PPT;        // site patterns per thread (independent dependency chains)
constexpr int TILE_PATTERNS = TPB * PPT;  // patterns per tile; FIXED so the summation order of a locus
                                          // never depends on batch composition or GPU count
constexpr int PRUNE_MIN_BLOCKS = PNB_MIN_BLOCKS;  // occupancy target of the pruning kernel (register budget)

// child-contribution source kinds
enum : uint32_t { K_TIP = 0, K_MEM = 1, K_REG = 2, K_STK = 3 };
// op flag bits
constexpr uint32_t F_KIND_MASK = 0x3u;
constexpr uint32_t F_FIRST = 0x4u;   // first child of its node: acc  = v
constexpr uint32_t F_LAST  = 0x8u;   // last child of its node:  cur  = acc (node complete)
constexpr uint32_t F_STORE = 0x10u;  // on F_LAST: write the node partial to HBM slot `dst`
constexpr uint32_t F_PUSH  = 0x20u;  // on F_LAST: stash the node partial in smem stack[(flags>>8)&0xff]

// One op = one (parent <- child) edge: v = P(edge) . L(child); acc (*)= v.
// Op k of a job uses P matrix k of that job (one-to-one), so the program and
// its P matrices are two contiguous arrays staged by two TMA bulk copies.
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
    int32_t n_tip_ops;      // TIP ops with src >= 0 (rows staged per chunk)
    int32_t root_kind;      // K_REG: root partial is `cur` after the last op; K_MEM / K_TIP: root_src
    int32_t root_src;
    int32_t tile0;          // first tile of this job
    int32_t n_tiles;
    int32_t out_index;      // where the job's logL goes in the output vector
    int32_t pad0, pad1;
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
    const int32_t* tile_job;   // tile -> job
    const Op* ops;             // batch program
    const double* Pm;          // batch P matrices [total_ops][16]
    double* tile_sum;          // [n_tiles]
    unsigned int* job_done;    // [n_jobs] arrival counters (self-resetting)
    double* out;               // per-job logL
    int32_t ops_cap;           // smem capacity in ops
    int32_t rows_cap;          // smem capacity in staged code rows
    int32_t stk_cap;           // smem stash entries
    int32_t tile_base;         // first tile of this launch (a batch may be launched in chunks)
};

#ifdef PNB_DEVICE_KERNELS  // kernel bodies: compiled into exactly one translation unit (pnb_api.cu)

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

__global__ void __launch_bounds__(128)
pnb_pmatrix_kernel(const __grid_constant__ ModelParams mp, const double* __restrict__ t,
                   int64_t nb, double* __restrict__ P_out) {
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= nb) return;
    double Pm[16];
    pmatrix_of_t(mp, t[g], Pm);
    double* o = P_out + g * 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) stg256(o + i * 4, Pm[i * 4 + 0], Pm[i * 4 + 1], Pm[i * 4 + 2], Pm[i * 4 + 3]);
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

// v = P . l   (child_partials @ P.T, reference :425), P row-major in shared memory (broadcast reads)
__device__ __forceinline__ void matvec(const double* __restrict__ Pk, const Vec4 (&l)[PPT], Vec4 (&v)[PPT]) {
    const double2* P2 = reinterpret_cast<const double2*>(Pk);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double2 pa = P2[i * 2 + 0];
        const double2 pb = P2[i * 2 + 1];
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            const double r = fma(pb.y, l[u].x3, fma(pb.x, l[u].x2, fma(pa.y, l[u].x1, pa.x * l[u].x0)));
            if (i == 0) v[u].x0 = r;
            else if (i == 1) v[u].x1 = r;
            else if (i == 2) v[u].x2 = r;
            else v[u].x3 = r;
        }
    }
}

// contribution of a tip child: column(s) of P selected by the device tip code
//   code < 4        : the single state j = code           -> column j of P
//   code & 0x80     : ambiguity / gap, low 4 bits = mask   -> sum of the allowed columns, in state order
__device__ __forceinline__ void tip_contrib(const double* __restrict__ Pk, const uint32_t (&code)[PPT],
                                            Vec4 (&v)[PPT]) {
    bool amb = false;
#pragma unroll
    for (int u = 0; u < PPT; ++u) amb |= (code[u] & 0x80u) != 0;
    if (!__any_sync(0xffffffffu, amb)) {
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            const double* col = Pk + code[u];
            v[u].x0 = col[0]; v[u].x1 = col[4]; v[u].x2 = col[8]; v[u].x3 = col[12];
        }
    } else {
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            const uint32_t m = (code[u] & 0x80u) ? (code[u] & 0xFu) : (1u << code[u]);
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (m & (1u << j)) { s0 += Pk[0 + j]; s1 += Pk[4 + j]; s2 += Pk[8 + j]; s3 += Pk[12 + j]; }
            }
            v[u].x0 = s0; v[u].x1 = s1; v[u].x2 = s2; v[u].x3 = s3;
        }
    }
}

__global__ void __launch_bounds__(TPB, PRUNE_MIN_BLOCKS)
pnb_prune_kernel(const __grid_constant__ PruneLaunch L, const __grid_constant__ ModelParams mp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar_stage;
    __shared__ double red[TPB / 32];
    __shared__ int s_is_last;

    const int tid = threadIdx.x;
    const int tile = blockIdx.x + L.tile_base;
    const int job = L.tile_job[tile];
    const JobDesc& jd = L.jobs[job];
    const int n_ops = jd.n_ops;
    const int op_off = jd.op_off;
    const int n_tip_ops = jd.n_tip_ops;
    const int tile0 = jd.tile0;
    const int P = jd.P;
    const int ld_codes = jd.ld_codes;
    const uint8_t* g_codes = jd.codes;

    Op* s_ops = reinterpret_cast<Op*>(smem_raw);
    double* s_P = reinterpret_cast<double*>(smem_raw + static_cast<size_t>(L.ops_cap) * sizeof(Op));
    unsigned char* s_codes = reinterpret_cast<unsigned char*>(s_P + static_cast<size_t>(L.ops_cap) * 16);
    double2* s_stk = reinterpret_cast<double2*>(s_codes + static_cast<size_t>(L.rows_cap) * TILE_PATTERNS);

    const int p0 = (tile - tile0) * TILE_PATTERNS;
    const Op* g_ops = L.ops + op_off;

    if (tid == 0) {
        mbar_init(&bar_stage, 1);
        mbar_fence_init();
    }
    __syncthreads();

    // warp 0 is the TMA producer: program, P matrices and the tile's code rows, one mbarrier
    if (tid < 32) {
        if (tid == 0) {
            mbar_arrive_expect_tx(&bar_stage, static_cast<uint32_t>(n_ops) * (sizeof(Op) + 128u) +
                                                  static_cast<uint32_t>(n_tip_ops) * TILE_PATTERNS);
            if (n_ops > 0) {
                tma_bulk_g2s(s_ops, g_ops, static_cast<uint32_t>(n_ops) * sizeof(Op), &bar_stage);
                tma_bulk_g2s(s_P, L.Pm + static_cast<int64_t>(op_off) * 16, static_cast<uint32_t>(n_ops) * 128u,
                             &bar_stage);
            }
        }
        __syncwarp();
        for (int k = tid; k < n_ops; k += 32) {
            const Op op = g_ops[k];
            if ((op.flags & F_KIND_MASK) == K_TIP && op.src >= 0)
                tma_bulk_g2s(s_codes + static_cast<size_t>(op.aux) * TILE_PATTERNS,
                             g_codes + static_cast<int64_t>(op.src) * ld_codes + p0, TILE_PATTERNS, &bar_stage);
        }
    }

    // per-thread pattern bookkeeping while the copies are in flight
    const int64_t slot_stride = jd.slot_stride;
    double* part0 = jd.partials + static_cast<int64_t>(p0 + tid) * 4;
    bool active[PPT];
#pragma unroll
    for (int u = 0; u < PPT; ++u) active[u] = (p0 + tid + u * TPB) < P;
    const unsigned char* my_codes = s_codes + tid;

    Vec4 cur[PPT];  // last completed node partial (register forwarding)
    Vec4 acc[PPT];  // node under construction
#pragma unroll
    for (int u = 0; u < PPT; ++u) { cur[u] = {1.0, 1.0, 1.0, 1.0}; acc[u] = {1.0, 1.0, 1.0, 1.0}; }

    mbar_wait(&bar_stage, 0);

    // Warp-uniform guard (tip_contrib votes across the full warp): a warp runs the program if any of
    // its lanes owns a pattern; lanes past the end compute on zero codes and never touch HBM.
    if (__any_sync(0xffffffffu, active[0])) {
        for (int k = 0; k < n_ops; ++k) {
            const Op op = s_ops[k];
            const double* Pk = s_P + k * 16;
            const uint32_t kind = op.flags & F_KIND_MASK;
            Vec4 v[PPT];
            if (kind == K_TIP) {
                uint32_t code[PPT];
#pragma unroll
                for (int u = 0; u < PPT; ++u) {
                    code[u] = 0u;
                    if (active[u]) code[u] = (op.src >= 0) ? my_codes[op.aux * TILE_PATTERNS + u * TPB] : 0x8Fu;
                }
                tip_contrib(Pk, code, v);
            } else if (kind == K_REG) {
                matvec(Pk, cur, v);
            } else {
                Vec4 l[PPT];
                if (kind == K_MEM) {
                    const double* src = part0 + static_cast<int64_t>(op.src) * slot_stride;
#pragma unroll
                    for (int u = 0; u < PPT; ++u) {
                        if (active[u]) ldg256(src + u * TPB * 4, l[u].x0, l[u].x1, l[u].x2, l[u].x3);
                        else l[u] = {1.0, 1.0, 1.0, 1.0};
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < PPT; ++u) {
                        const double2 x = s_stk[(op.src * 2 + 0) * TILE_PATTERNS + u * TPB + tid];
                        const double2 y = s_stk[(op.src * 2 + 1) * TILE_PATTERNS + u * TPB + tid];
                        l[u] = {x.x, x.y, y.x, y.y};
                    }
                }
                matvec(Pk, l, v);
            }
            if (op.flags & F_FIRST) {
#pragma unroll
                for (int u = 0; u < PPT; ++u) acc[u] = v[u];
            } else {
#pragma unroll
                for (int u = 0; u < PPT; ++u) {
                    acc[u].x0 *= v[u].x0; acc[u].x1 *= v[u].x1; acc[u].x2 *= v[u].x2; acc[u].x3 *= v[u].x3;
                }
            }
            if (op.flags & F_LAST) {
#pragma unroll
                for (int u = 0; u < PPT; ++u) cur[u] = acc[u];
                if (op.flags & F_STORE) {
                    double* dst = part0 + static_cast<int64_t>(op.dst) * slot_stride;
#pragma unroll
                    for (int u = 0; u < PPT; ++u)
                        if (active[u]) stg256(dst + u * TPB * 4, cur[u].x0, cur[u].x1, cur[u].x2, cur[u].x3);
                }
                if (op.flags & F_PUSH) {
                    const int si = (op.flags >> 8) & 0xff;
#pragma unroll
                    for (int u = 0; u < PPT; ++u) {
                        s_stk[(si * 2 + 0) * TILE_PATTERNS + u * TPB + tid] = make_double2(cur[u].x0, cur[u].x1);
                        s_stk[(si * 2 + 1) * TILE_PATTERNS + u * TPB + tid] = make_double2(cur[u].x2, cur[u].x3);
                    }
                }
            }
        }
    }

    // root: site likelihood, clamp, log, weight  (reference :394-397)
    double lsum = 0.0;
    const int root_kind = jd.root_kind;
    const int root_src = jd.root_src;
    const double* weights = jd.weights;
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
        if (!active[u]) continue;
        const int pat = p0 + tid + u * TPB;
        Vec4 r = cur[u];
        if (root_kind == K_MEM) {
            ldg256(part0 + static_cast<int64_t>(root_src) * slot_stride + u * TPB * 4, r.x0, r.x1, r.x2, r.x3);
        } else if (root_kind == K_TIP) {
            const uint32_t c = (root_src >= 0) ? g_codes[static_cast<int64_t>(root_src) * ld_codes + pat] : 0x8Fu;
            const uint32_t m = (c & 0x80u) ? (c & 0xFu) : (1u << c);
            r = {(m & 1u) ? 1.0 : 0.0, (m & 2u) ? 1.0 : 0.0, (m & 4u) ? 1.0 : 0.0, (m & 8u) ? 1.0 : 0.0};
        }
        double site = fma(r.x3, mp.pi[3], fma(r.x2, mp.pi[2], fma(r.x1, mp.pi[1], r.x0 * mp.pi[0])));
        site = (site < 1e-300) ? 1e-300 : site;  // np.maximum(site_like, 1e-300): NaN propagates
        lsum = fma(log(site), weights[pat], lsum);
    }

    // fused per-locus reduction: tile sum -> last-arriving tile sums tiles in index order
    const double tsum = block_sum(lsum, red);
    const int n_tiles = jd.n_tiles;
    if (tid == 0) {
        __stcg(&L.tile_sum[tile], tsum);
        __threadfence();
        const unsigned int prev = atomicAdd(&L.job_done[job], 1u);
        s_is_last = (prev == static_cast<unsigned int>(n_tiles) - 1u);
    }
    __syncthreads();
    if (s_is_last) {
        __threadfence();
        double s = 0.0;
        for (int t = tid; t < n_tiles; t += TPB) s += __ldcg(&L.tile_sum[tile0 + t]);
        const double total = block_sum(s, red);
        if (tid == 0) {
            L.out[jd.out_index] = total;
            L.job_done[job] = 0u;
        }
    }
}

#endif  // PNB_DEVICE_KERNELS

}  // namespace pnb
