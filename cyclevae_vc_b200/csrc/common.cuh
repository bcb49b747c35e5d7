// Shared helpers of libcyclevae_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cyclevae_b200.h"

namespace cvb {

void set_error(const char* fmt, ...);

#define CVB_CHECK(expr)                                                                         \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            cvb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

#define CVB_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            cvb::set_error(__VA_ARGS__);  \
            return 2;                     \
        }                                 \
    } while (0)

void count_launch();
#define CVB_LAUNCH_CHECK()              \
    do {                                \
        cvb::count_launch();            \
        CVB_CHECK(cudaGetLastError());  \
    } while (0)

// event-pair profiling of selected launches (gemm.cu)
void prof_begin(cudaStream_t s, int kind);
void prof_end(cudaStream_t s, int kind);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t ceil_div_sz(size_t a, size_t b) { return (a + b - 1) / b; }
static inline size_t round_up_sz(size_t a, size_t b) { return ceil_div_sz(a, b) * b; }

// true when the persistent cluster kernels should be launched WITHOUT the cooperative attribute: CVB_TC_NOCOOP=1, or the
// process runs under Nsight Compute (which cannot replay a cooperative cluster launch: the capture dies at the first
// one).  Co-residency of the whole grid is then guaranteed by the occupancy query on an otherwise idle device only.
bool launch_without_coop();
bool relaxed_polling();   // CVB_TC_POLL=relaxed: relaxed loads + one acquire fence instead of ld.acquire.gpu per poll (A/B; slower)

struct DeviceInfo {
    int n_sm;
    int max_smem_optin;
    int cc_major, cc_minor;
};
int get_device_info(DeviceInfo* out);

// conv geometry of TwoSidedDilConv1d (gru_vae.py:40-51)
static inline int ipow(int b, int e) {
    int r = 1;
    for (int i = 0; i < e; ++i) r *= b;
    return r;
}
static inline int conv_pad(const cvb_net* n) { return (ipow(n->kernel_size, n->n_conv) - 1) / 2; }
static inline int conv_dim(const cvb_net* n) { return n->in_dim * ipow(n->kernel_size, n->n_conv); }
static inline int tot_in_dim(const cvb_net* n) { return conv_dim(n) + n->out_dim; }

// internal GEMM (row-major semantics), gemm.cu
// grad = false: forward-path product (fp16 hi/lo operands on the tensor-core path); grad = true: gradient product
// (bf16 hi/lo).  Small products and CVB_GEMM=cublas go to cuBLAS fp32.
int gemm_rm(cudaStream_t s, bool transA, bool transB, int M, int N, int K, float alpha,
            const float* A, int lda, const float* B, int ldb, float beta, float* C, int ldc, bool grad = false);
bool want_tc_gemm();   // false under CVB_GEMM=cublas
// split-precision tcgen05 GEMM, gemm_tc.cu
bool gemm_tc_eligible(int M, int N, int K);
// operand = virtual im2col of a dilated conv on the flattened padded grid: element (row r, column tap*ci + c) is
// src[(r + tap*dshift)*ci + c] (src already offset to tap 0 of row 0)
struct ConvGather {
    int ci;
    int dshift;
};
int gemm_tc(cudaStream_t s, bool transA, bool transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
            bool beta1, const float* bias, float* C, int ldc, bool f16, const float* B2 = nullptr, int ldb2 = 0, int N1 = 0,
            const ConvGather* gA = nullptr, const ConvGather* gB = nullptr);
// one product of a grouped launch: C[M,N] = alpha op(A) op(B) (+ C if beta1) (+ bias[N]); same operand conventions as gemm_tc
struct GemmDesc {
    bool transA = false, transB = false;
    int M = 0, N = 0, K = 0;
    const float* A = nullptr;
    int lda = 0;
    const float* B = nullptr;
    int ldb = 0;
    const float* A2 = nullptr;   // A stored [K,M] (transA): columns >= M1 come from A2 (lda2)
    int lda2 = 0, M1 = 0;
    const float* B2 = nullptr;   // B stored [K,N]: columns >= N1 come from B2 (ldb2)
    int ldb2 = 0, N1 = 0;
    const ConvGather* gA = nullptr;
    const ConvGather* gB = nullptr;
    float* C = nullptr;
    int ldc = 0;
    const float* bias = nullptr;
    float alpha = 1.f;
    bool beta1 = false;
    bool a_const = false, b_const = false;   // the operand is a PARAMETER: its 16-bit image is kept until weights_changed()
    // optional output map of the fused front-end: product row r = b * map_Tp + t is stored at row t * map_B + b of C
    // (time-major) when t < map_T, dropped otherwise, and multiplied element-wise by mask (same layout as C) if given
    int map_Tp = 0, map_T = 0, map_B = 0;
    const float* mask = nullptr;
    bool f16 = true;             // fp16 hi/lo operands (forward products) or bf16 hi/lo (gradient products)
};
// up to 6 independent products in ONE persistent launch (their tiles are walked back to back; list long-K products first)
int gemm_tc_group(cudaStream_t s, const GemmDesc* d, int n);
void weights_changed();                 // invalidates the cached images of parameter operands
unsigned long long weights_generation();   // bumped by weights_changed()
int reserve_workspace(size_t bytes);    // pre-sizes the operand-image arena of the current device

// elementwise / small kernels, elementwise.cu
int colsum(cudaStream_t s, const float* A, int rows, int cols, int lda, float* out, bool accumulate);
int fill_rows(cudaStream_t s, float* dst, size_t rows, int cols, int ld, const float* bias);
int zero_floats(cudaStream_t s, float* p, size_t n);

#ifdef __CUDACC__
// ---- device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// MUFU-based forms for the tensor-core recurrence kernels: |error| <~ 3e-7 absolute (ex2.approx: 2 ulp near 0)
__device__ __forceinline__ float rcp_approx(float x) {   // MUFU.RCP, 1 ulp, no slow-path subroutine (rcp(inf) = 0)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(2.0f, sigmoid_fast(2.0f * x), -1.0f); }

// 16-byte read-only load pinned in program order (asm volatile): prefetches issued at the top of a recurrence
// step stay there instead of being sunk next to their first use (where they would expose the full HBM latency)
__device__ __forceinline__ float4 ldg_nc_v4_pinned(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// polling without an acquire per poll (ld.acquire.gpu invalidates L1 on every iteration): relaxed loads, ONE acquire
// fence once the value is seen -- the release / relaxed-read + fence pattern of the PTX memory model
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void spin_until_ge(const unsigned* ctr, unsigned target, bool relaxed) {
    if (relaxed) {
        while (ld_relaxed_gpu(ctr) < target) {
        }
        fence_acq_rel_gpu();
    } else {
        while (ld_acquire_gpu(ctr) < target) {
        }
    }
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Arrival counters in CTR_BANKS banks (one 128-byte line each).  128 CTAs adding to ONE address serialise in the L2 atomic
// unit (~27 cycles per same-address atomic, B300_MICROARCH.md "L2-atom multi-CTA"): the last arrival of a round would be
// performed > 3 000 cycles after the first.  CTA c arrives on bank c % CTR_BANKS; a polling warp watches every bank, one
// per lane (the warp barrier behind the polls makes every lane's acquire cumulative for all of them).
constexpr int CTR_BANKS = 8;
__device__ __forceinline__ void banked_arrive(unsigned* base, int c) { red_release_gpu_add(base + 32 * (c & (CTR_BANKS - 1)), 1u); }
__device__ __forceinline__ void banked_wait_warp(const unsigned* base, unsigned per_bank_target, int lane, bool relaxed) {
    if (lane < CTR_BANKS) spin_until_ge(base + 32 * lane, per_bank_target, relaxed);
    __syncwarp();
}

// Grid-wide barrier for a co-resident (cooperatively launched) grid.  `ctr` is zero at kernel
// start; `target` is this thread's running count of expected arrivals (starts at 0).
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& target, unsigned n_cta) {
    __syncthreads();
    target += n_cta;
    if (threadIdx.x == 0) {
        __threadfence();
        red_release_gpu_add(ctr, 1u);
        while (ld_acquire_gpu(ctr) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src, bool valid) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Philox4x32-10 (Salmon et al. 2011), counter-based RNG
struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    __device__ __forceinline__ uint4 operator()(uint64_t ctr, uint32_t stream = 0) const {
        uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = stream, c3 = 0x9E3779B9u;
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            uint32_t n0 = hi1 ^ c1 ^ a, n1 = lo1, n2 = hi0 ^ c3 ^ b, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};
__device__ __forceinline__ float u32_to_unit(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }  // [0,1)
#endif

}  // namespace cvb
