// pnb_compress.cu -- K-d: site-pattern compression (integer work, bit-exact).
//
// Reference: FelsensteinCalculator.__init__, src/_seq_likelihood.py:344-367 --
// a dict keyed by the tuple of upper-cased column characters, so patterns are
// numbered in FIRST-OCCURRENCE order and weights are column counts.  Both
// implementations below reproduce exactly that numbering:
//   host  : open-addressing hash table over column contents, single pass.
//   device: concurrent hash insert where each slot keeps the MINIMUM column
//           index of its content class (atomicMin), then a prefix sum over the
//           "I am the first occurrence" flags gives the pattern ids.
// Hash collisions are resolved by comparing full column contents, never by
// trusting the hash.
#include <algorithm>
#include <climits>
#include <vector>

#include "pnb_host.hpp"

using namespace pnb;

namespace {

__host__ __device__ inline uint64_t mix64(uint64_t h, uint64_t v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 32;
    return h;
}

}  // namespace

// ---------------------------------------------------------------- host
int pnb_compress_host(const uint8_t* cols, int32_t n, int64_t L, const uint8_t* code_of_byte,
                      int64_t* P_out, uint8_t* codes_out, int64_t ld_codes, int64_t* weights_out,
                      int64_t* first_col_out, int32_t* pattern_of_col_out) {
    size_t T = 16;
    while (T < static_cast<size_t>(L) * 2) T <<= 1;
    std::vector<int32_t> table;  // pattern id + 1, 0 = empty
    std::vector<uint8_t> keys;   // P x n contents of the first occurrence of each pattern
    try {
        table.assign(T, 0);
        keys.reserve(static_cast<size_t>(std::min<int64_t>(L, 1 << 16)) * n);
    } catch (...) {
        set_error("pnb_compress: out of host memory");
        return PNB_ERR_NOMEM;
    }
    constexpr int B = 256;  // columns are gathered in blocks so the strided reads stay cache friendly
    std::vector<uint8_t> blk(static_cast<size_t>(B) * n);
    int64_t P = 0;
    for (int64_t c0 = 0; c0 < L; c0 += B) {
        const int nb = static_cast<int>(std::min<int64_t>(B, L - c0));
        for (int r = 0; r < n; ++r) {
            const uint8_t* src = cols + static_cast<size_t>(r) * L + c0;
            for (int b = 0; b < nb; ++b) blk[static_cast<size_t>(b) * n + r] = src[b];
        }
        for (int b = 0; b < nb; ++b) {
            const uint8_t* col = blk.data() + static_cast<size_t>(b) * n;
            uint64_t h = 0x243F6A8885A308D3ull;
            for (int r = 0; r < n; ++r) h = mix64(h, col[r]);
            size_t s = h & (T - 1);
            int64_t pid;
            for (;;) {
                const int32_t e = table[s];
                if (e == 0) {
                    pid = P++;
                    table[s] = static_cast<int32_t>(pid + 1);
                    keys.insert(keys.end(), col, col + n);
                    weights_out[pid] = 1;
                    first_col_out[pid] = c0 + b;
                    if (pid < ld_codes)
                        for (int r = 0; r < n; ++r)
                            codes_out[static_cast<size_t>(r) * ld_codes + pid] = code_of_byte[col[r]];
                    break;
                }
                if (memcmp(keys.data() + static_cast<size_t>(e - 1) * n, col, n) == 0) {
                    pid = e - 1;
                    weights_out[pid] += 1;
                    break;
                }
                s = (s + 1) & (T - 1);
            }
            if (pattern_of_col_out) pattern_of_col_out[c0 + b] = static_cast<int32_t>(pid);
        }
    }
    if (P > ld_codes) {
        set_error("pnb_compress: ld_codes=%lld is smaller than the number of patterns %lld",
                  static_cast<long long>(ld_codes), static_cast<long long>(P));
        return PNB_ERR_INVALID;
    }
    *P_out = P;
    return PNB_OK;
}

// -------------------------------------------------------------- device
namespace {

constexpr int CB = 256;      // threads per block of the column kernels
constexpr int SCAN_B = 1024; // elements per block of the scan

__device__ __forceinline__ bool same_column(const uint8_t* __restrict__ cols, int n, int64_t L, int a, int b) {
    for (int r = 0; r < n; ++r)
        if (cols[static_cast<int64_t>(r) * L + a] != cols[static_cast<int64_t>(r) * L + b]) return false;
    return true;
}

// slot value = smallest column index seen so far among columns with the slot's content
__global__ void __launch_bounds__(CB)
k_insert(const uint8_t* __restrict__ cols, int n, int64_t L, int* table, uint32_t mask, int* slot_of_col) {
    const int64_t c = static_cast<int64_t>(blockIdx.x) * CB + threadIdx.x;
    if (c >= L) return;
    uint64_t h = 0x243F6A8885A308D3ull;
    for (int r = 0; r < n; ++r) h = mix64(h, cols[static_cast<int64_t>(r) * L + c]);
    uint32_t s = static_cast<uint32_t>(h) & mask;
    const int me = static_cast<int>(c);
    for (;;) {
        int cur = *reinterpret_cast<volatile int*>(&table[s]);
        if (cur == INT_MAX) {
            cur = atomicCAS(&table[s], INT_MAX, me);
            if (cur == INT_MAX) break;  // slot claimed for this content
        }
        if (cur == me || same_column(cols, n, L, cur, me)) {
            atomicMin(&table[s], me);
            break;
        }
        s = (s + 1) & mask;
    }
    slot_of_col[c] = static_cast<int>(s);
}

__global__ void __launch_bounds__(CB)
k_count_flag(const int* __restrict__ table, const int* __restrict__ slot_of_col, int64_t L, int* count,
             int* flag) {
    const int64_t c = static_cast<int64_t>(blockIdx.x) * CB + threadIdx.x;
    if (c >= L) return;
    const int s = slot_of_col[c];
    atomicAdd(&count[s], 1);
    flag[c] = (table[s] == static_cast<int>(c)) ? 1 : 0;
}

// per-block exclusive scan of flag -> local prefix, block totals
__global__ void __launch_bounds__(SCAN_B)
k_scan_local(const int* __restrict__ flag, int64_t L, int* prefix, int* block_total) {
    __shared__ int warp_tot[SCAN_B / 32];
    const int64_t c = static_cast<int64_t>(blockIdx.x) * SCAN_B + threadIdx.x;
    const int v = c < L ? flag[c] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
        int w = warp_tot[threadIdx.x];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, w, o);
            if (threadIdx.x >= o) w += y;
        }
        warp_tot[threadIdx.x] = w;  // inclusive
    }
    __syncthreads();
    const int base = (threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0;
    if (c < L) prefix[c] = base + x - v;
    if (threadIdx.x == SCAN_B - 1) block_total[blockIdx.x] = base + x;
}

// single block: exclusive scan of the block totals, in place; writes the grand total
__global__ void __launch_bounds__(1024)
k_scan_totals(int* block_total, int nblocks, int* grand_total) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nblocks; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const int v = i < nblocks ? block_total[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = warp_tot[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            warp_tot[threadIdx.x] = w;
        }
        __syncthreads();
        const int base = carry + ((threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0);
        if (i < nblocks) block_total[i] = base + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = base + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void __launch_bounds__(CB)
k_assign(const uint8_t* __restrict__ cols, int n, int64_t L, const int* __restrict__ slot_of_col,
         const int* __restrict__ flag, const int* __restrict__ prefix, const int* __restrict__ block_off,
         const int* __restrict__ count, const uint8_t* __restrict__ code_of_byte, int* pid_of_slot,
         long long* weights, long long* first_col, uint8_t* codes, int64_t ld) {
    const int64_t c = static_cast<int64_t>(blockIdx.x) * CB + threadIdx.x;
    if (c >= L || !flag[c]) return;
    const int pid = prefix[c] + block_off[c / SCAN_B];
    const int s = slot_of_col[c];
    pid_of_slot[s] = pid;
    weights[pid] = count[s];
    first_col[pid] = c;
    for (int r = 0; r < n; ++r)
        codes[static_cast<int64_t>(r) * ld + pid] = code_of_byte[cols[static_cast<int64_t>(r) * L + c]];
}

__global__ void __launch_bounds__(CB)
k_pattern_of_col(const int* __restrict__ slot_of_col, const int* __restrict__ pid_of_slot, int64_t L, int* out) {
    const int64_t c = static_cast<int64_t>(blockIdx.x) * CB + threadIdx.x;
    if (c < L) out[c] = pid_of_slot[slot_of_col[c]];
}

struct DevTmp {
    std::vector<void*> ptrs;
    ~DevTmp() { for (void* p : ptrs) cudaFree(p); }
    template <typename T>
    cudaError_t get(T** out, size_t count) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(p);
        *out = static_cast<T*>(p);
        return e;
    }
};

}  // namespace

int pnb_compress_device(pnb_ctx* ctx, const uint8_t* cols, int32_t n, int64_t L,
                        const uint8_t* code_of_byte, int64_t* P_out, uint8_t* codes_out,
                        int64_t ld_codes, int64_t* weights_out, int64_t* first_col_out,
                        int32_t* pattern_of_col_out) {
    PNB_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    size_t T = 1024;
    while (T < static_cast<size_t>(L) * 2) T <<= 1;
    const int nscan = static_cast<int>((L + SCAN_B - 1) / SCAN_B);
    DevTmp tmp;
    uint8_t *d_cols, *d_lut, *d_codes;
    int *d_table, *d_slot, *d_count, *d_flag, *d_prefix, *d_btot, *d_total, *d_pid, *d_poc;
    long long *d_w, *d_first;
    PNB_CUDA_TRY(tmp.get(&d_cols, static_cast<size_t>(n) * L));
    PNB_CUDA_TRY(tmp.get(&d_lut, 256));
    PNB_CUDA_TRY(tmp.get(&d_table, T));
    PNB_CUDA_TRY(tmp.get(&d_slot, L));
    PNB_CUDA_TRY(tmp.get(&d_count, T));
    PNB_CUDA_TRY(tmp.get(&d_flag, L));
    PNB_CUDA_TRY(tmp.get(&d_prefix, L));
    PNB_CUDA_TRY(tmp.get(&d_btot, nscan));
    PNB_CUDA_TRY(tmp.get(&d_total, 1));
    PNB_CUDA_TRY(tmp.get(&d_pid, T));
    PNB_CUDA_TRY(cudaMemcpyAsync(d_cols, cols, static_cast<size_t>(n) * L, cudaMemcpyHostToDevice, st));
    PNB_CUDA_TRY(cudaMemcpyAsync(d_lut, code_of_byte, 256, cudaMemcpyHostToDevice, st));
    // INT_MAX = 0x7FFFFFFF is not a byte pattern: fill the table with a tiny kernel-free trick
    // (0x7F7F7F7F would be a valid column index for L > 2.1e9 only, and L < 2^31 is enforced;
    //  still, use the exact value so "empty" cannot collide with a column)
    {
        std::vector<int> init(T, INT_MAX);
        PNB_CUDA_TRY(cudaMemcpyAsync(d_table, init.data(), T * sizeof(int), cudaMemcpyHostToDevice, st));
        PNB_CUDA_TRY(cudaStreamSynchronize(st));
    }
    PNB_CUDA_TRY(cudaMemsetAsync(d_count, 0, T * sizeof(int), st));
    const unsigned gb = static_cast<unsigned>((L + CB - 1) / CB);
    k_insert<<<gb, CB, 0, st>>>(d_cols, n, L, d_table, static_cast<uint32_t>(T - 1), d_slot);
    PNB_CUDA_TRY(cudaGetLastError());
    k_count_flag<<<gb, CB, 0, st>>>(d_table, d_slot, L, d_count, d_flag);
    PNB_CUDA_TRY(cudaGetLastError());
    k_scan_local<<<static_cast<unsigned>(nscan), SCAN_B, 0, st>>>(d_flag, L, d_prefix, d_btot);
    PNB_CUDA_TRY(cudaGetLastError());
    k_scan_totals<<<1, 1024, 0, st>>>(d_btot, nscan, d_total);
    PNB_CUDA_TRY(cudaGetLastError());
    int P32 = 0;
    PNB_CUDA_TRY(cudaMemcpyAsync(&P32, d_total, sizeof(int), cudaMemcpyDeviceToHost, st));
    PNB_CUDA_TRY(cudaStreamSynchronize(st));
    const int64_t P = P32;
    PNB_REQUIRE(P <= ld_codes, PNB_ERR_INVALID, "pnb_compress: ld_codes=%lld is smaller than the number of patterns %lld",
                static_cast<long long>(ld_codes), static_cast<long long>(P));
    PNB_CUDA_TRY(tmp.get(&d_w, P));
    PNB_CUDA_TRY(tmp.get(&d_first, P));
    PNB_CUDA_TRY(tmp.get(&d_codes, static_cast<size_t>(n) * P));
    k_assign<<<gb, CB, 0, st>>>(d_cols, n, L, d_slot, d_flag, d_prefix, d_btot, d_count, d_lut, d_pid, d_w,
                                d_first, d_codes, P);
    PNB_CUDA_TRY(cudaGetLastError());
    if (pattern_of_col_out) {
        PNB_CUDA_TRY(tmp.get(&d_poc, L));
        k_pattern_of_col<<<gb, CB, 0, st>>>(d_slot, d_pid, L, d_poc);
        PNB_CUDA_TRY(cudaGetLastError());
        PNB_CUDA_TRY(cudaMemcpyAsync(pattern_of_col_out, d_poc, L * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");
    PNB_CUDA_TRY(cudaMemcpyAsync(weights_out, d_w, P * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PNB_CUDA_TRY(cudaMemcpyAsync(first_col_out, d_first, P * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PNB_CUDA_TRY(cudaMemcpy2DAsync(codes_out, ld_codes, d_codes, P, P, n, cudaMemcpyDeviceToHost, st));
    PNB_CUDA_TRY(cudaStreamSynchronize(st));
    *P_out = P;
    return PNB_OK;
}
