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

constexpr int TPB = 128;            // threads per block: one thread = one site pattern of a chunk
constexpr int TILE_PATTERNS = 512;  // patterns per tile; FIXED so the summation order of a locus
                                    // never depends on batch composition or GPU count
constexpr int CHUNKS_PER_TILE = TILE_PATTERNS / TPB;

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
    const uint8_t* codes;   // [n_rows][ld_codes] 4-bit tip masks
    double* partials;       // [n_slots][P_pad][4] per-node partials of this locus
    const double* weights;  // [P_pad] pattern multiplicities (integer valued)
    int64_t slot_stride;    // doubles per slot (= P_pad * 4)
    int32_t P;              // site patterns
    int32_t ld_codes;       // row stride of codes (= P_pad, multiple of TPB)
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
    int32_t pad;
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
// K-b / K-c: pruning.  grid = tiles, block = TPB threads.
//
// Dynamic shared memory layout (bytes):
//   [0, ops_cap*16)                      Op program of the job
//   [.., + ops_cap*128)                  P matrices, one per op (row-major 4x4 f64)
//   [.., + 2*rows_cap*TPB)               double-buffered tip-code rows of the chunk
//   [.., + stk_cap*TPB*32)               (NOSTORE only) partial stash stack
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

__global__ void __launch_bounds__(TPB, 8)
pnb_prune_kernel(const __grid_constant__ PruneLaunch L, const __grid_constant__ ModelParams mp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar_prog;
    __shared__ __align__(8) uint64_t bar_codes[2];
    __shared__ double red[TPB / 32];
    __shared__ int s_is_last;

    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int job = L.tile_job[tile];
    const JobDesc jd = L.jobs[job];

    Op* s_ops = reinterpret_cast<Op*>(smem_raw);
    double* s_P = reinterpret_cast<double*>(smem_raw + static_cast<size_t>(L.ops_cap) * sizeof(Op));
    unsigned char* s_codes = reinterpret_cast<unsigned char*>(s_P + static_cast<size_t>(L.ops_cap) * 16);
    double2* s_stk = reinterpret_cast<double2*>(s_codes + static_cast<size_t>(2) * L.rows_cap * TPB);

    const int p0 = (tile - jd.tile0) * TILE_PATTERNS;
    const int remaining = jd.P - p0;
    const int n_chunks = (min(remaining, TILE_PATTERNS) + TPB - 1) / TPB;
    const int n_ops = jd.n_ops;
    const Op* g_ops = L.ops + jd.op_off;

    if (tid == 0) {
        mbar_init(&bar_prog, 1);
        mbar_init(&bar_codes[0], 1);
        mbar_init(&bar_codes[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // warp 0 is the TMA producer: program + P matrices once, code rows per chunk
    auto issue_codes = [&](int chunk) {
        // called by all lanes of warp 0
        const int buf = chunk & 1;
        if (tid == 0) mbar_arrive_expect_tx(&bar_codes[buf], static_cast<uint32_t>(jd.n_tip_ops) * TPB);
        __syncwarp();
        const int64_t pc = p0 + chunk * TPB;
        for (int k = tid; k < n_ops; k += 32) {
            const Op op = g_ops[k];
            if ((op.flags & F_KIND_MASK) == K_TIP && op.src >= 0) {
                tma_bulk_g2s(s_codes + (static_cast<size_t>(buf) * L.rows_cap + op.aux) * TPB,
                             jd.codes + static_cast<int64_t>(op.src) * jd.ld_codes + pc, TPB,
                             &bar_codes[buf]);
            }
        }
    };
    if (tid < 32) {
        if (tid == 0) {
            mbar_arrive_expect_tx(&bar_prog, static_cast<uint32_t>(n_ops) * (sizeof(Op) + 128u));
            if (n_ops > 0) {
                tma_bulk_g2s(s_ops, g_ops, static_cast<uint32_t>(n_ops) * sizeof(Op), &bar_prog);
                tma_bulk_g2s(s_P, L.Pm + static_cast<int64_t>(jd.op_off) * 16,
                             static_cast<uint32_t>(n_ops) * 128u, &bar_prog);
            }
        }
        issue_codes(0);
    }

    double lsum = 0.0;
    mbar_wait(&bar_prog, 0);

    for (int c = 0; c < n_chunks; ++c) {
        if (tid < 32 && c + 1 < n_chunks) issue_codes(c + 1);
        const int buf = c & 1;
        mbar_wait(&bar_codes[buf], (c >> 1) & 1);

        const int pat = p0 + c * TPB + tid;
        const bool active = pat < jd.P;
        const unsigned char* my_codes = s_codes + static_cast<size_t>(buf) * L.rows_cap * TPB + tid;
        double* my_part = jd.partials + static_cast<int64_t>(pat) * 4;

        double c0 = 1.0, c1 = 1.0, c2 = 1.0, c3 = 1.0;  // cur: last completed node partial
        double a0 = 1.0, a1 = 1.0, a2 = 1.0, a3 = 1.0;  // acc: node under construction

        if (active) {
            for (int k = 0; k < n_ops; ++k) {
                const Op op = s_ops[k];
                const double* Pk = s_P + k * 16;
                const uint32_t kind = op.flags & F_KIND_MASK;
                double v0, v1, v2, v3;
                if (kind == K_TIP) {
                    const uint32_t code = (op.src >= 0) ? my_codes[op.aux * TPB] : 0xFu;
                    if (__popc(code) == 1) {
                        const int j = __ffs(code) - 1;
                        v0 = Pk[0 + j]; v1 = Pk[4 + j]; v2 = Pk[8 + j]; v3 = Pk[12 + j];
                    } else {
                        // ambiguity / gap: sum of the allowed columns, in state order
                        v0 = v1 = v2 = v3 = 0.0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (code & (1u << j)) {
                                v0 += Pk[0 + j]; v1 += Pk[4 + j]; v2 += Pk[8 + j]; v3 += Pk[12 + j];
                            }
                        }
                    }
                } else {
                    double l0, l1, l2, l3;
                    if (kind == K_REG) {
                        l0 = c0; l1 = c1; l2 = c2; l3 = c3;
                    } else if (kind == K_MEM) {
                        ldg256(my_part + static_cast<int64_t>(op.src) * jd.slot_stride, l0, l1, l2, l3);
                    } else {
                        const double2 x = s_stk[(op.src * 2 + 0) * TPB + tid];
                        const double2 y = s_stk[(op.src * 2 + 1) * TPB + tid];
                        l0 = x.x; l1 = x.y; l2 = y.x; l3 = y.y;
                    }
                    // child_partials @ P.T  (reference :425): v_i = sum_j P[i][j] * L_j
                    v0 = fma(Pk[3], l3, fma(Pk[2], l2, fma(Pk[1], l1, Pk[0] * l0)));
                    v1 = fma(Pk[7], l3, fma(Pk[6], l2, fma(Pk[5], l1, Pk[4] * l0)));
                    v2 = fma(Pk[11], l3, fma(Pk[10], l2, fma(Pk[9], l1, Pk[8] * l0)));
                    v3 = fma(Pk[15], l3, fma(Pk[14], l2, fma(Pk[13], l1, Pk[12] * l0)));
                }
                if (op.flags & F_FIRST) {
                    a0 = v0; a1 = v1; a2 = v2; a3 = v3;
                } else {
                    a0 *= v0; a1 *= v1; a2 *= v2; a3 *= v3;
                }
                if (op.flags & F_LAST) {
                    c0 = a0; c1 = a1; c2 = a2; c3 = a3;
                    if (op.flags & F_STORE)
                        stg256(my_part + static_cast<int64_t>(op.dst) * jd.slot_stride, c0, c1, c2, c3);
                    if (op.flags & F_PUSH) {
                        const int si = (op.flags >> 8) & 0xff;
                        s_stk[(si * 2 + 0) * TPB + tid] = make_double2(c0, c1);
                        s_stk[(si * 2 + 1) * TPB + tid] = make_double2(c2, c3);
                    }
                }
            }
            // root: site likelihood, clamp, log, weight  (reference :394-397)
            if (jd.root_kind == K_MEM) {
                ldg256(my_part + static_cast<int64_t>(jd.root_src) * jd.slot_stride, c0, c1, c2, c3);
            } else if (jd.root_kind == K_TIP) {
                const uint32_t code = (jd.root_src >= 0)
                    ? jd.codes[static_cast<int64_t>(jd.root_src) * jd.ld_codes + pat] : 0xFu;
                c0 = (code & 1u) ? 1.0 : 0.0; c1 = (code & 2u) ? 1.0 : 0.0;
                c2 = (code & 4u) ? 1.0 : 0.0; c3 = (code & 8u) ? 1.0 : 0.0;
            }
            double site = fma(c3, mp.pi[3], fma(c2, mp.pi[2], fma(c1, mp.pi[1], c0 * mp.pi[0])));
            site = (site < 1e-300) ? 1e-300 : site;  // np.maximum(site_like, 1e-300): NaN propagates
            lsum = fma(log(site), jd.weights[pat], lsum);
        }
        __syncthreads();  // everyone is done with code buffer `buf` before it is refilled
    }

    // fused per-locus reduction: tile sum -> last-arriving tile sums tiles in index order
    const double tsum = block_sum(lsum, red);
    if (tid == 0) {
        __stcg(&L.tile_sum[tile], tsum);
        __threadfence();
        const unsigned int prev = atomicAdd(&L.job_done[job], 1u);
        s_is_last = (prev == static_cast<unsigned int>(jd.n_tiles) - 1u);
    }
    __syncthreads();
    if (s_is_last) {
        __threadfence();
        double s = 0.0;
        for (int t = tid; t < jd.n_tiles; t += TPB) s += __ldcg(&L.tile_sum[jd.tile0 + t]);
        const double total = block_sum(s, red);
        if (tid == 0) {
            L.out[jd.out_index] = total;
            L.job_done[job] = 0u;
        }
    }
}

#endif  // PNB_DEVICE_KERNELS

}  // namespace pnb
