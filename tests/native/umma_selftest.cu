// Self-test of the tcgen05 building blocks (umma.cuh): one CTA computes D[128,N] = A[128,K] * B[N,K]^T
// with bf16 operands staged in shared memory in the K-major no-swizzle core-matrix layout, fp32
// accumulation in TMEM, tcgen05.ld back to registers.  tests/test_gpu_umma.py compares with torch.
//   mode 0: operands converted + arranged by the CTA, core matrices ordered [row-block][k-block]
//   mode 1: same, ordered [k-block][row-block]
//   mode 2: operands pre-arranged in global memory (bf16, mode-0 order) and fetched with
//           cp.async.bulk + mbarrier complete_tx -- the path the recurrence kernel uses for h_{t-1}
#include "common.cuh"
#include "umma.cuh"

namespace cvb {
using namespace umma;

__device__ __forceinline__ uint16_t f2bf(float x) {
    uint32_t u = __float_as_uint(x);
    return (uint16_t)((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);
}

__global__ void __launch_bounds__(128, 1) k_umma_selftest(int mode, int N, int K, const float* __restrict__ A,
                                                          const float* __restrict__ B, float* __restrict__ D) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int M = 128;
    uint16_t* sA = reinterpret_cast<uint16_t*>(smem);
    uint16_t* sB = sA + (size_t)M * K;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)N * K);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KB = K / 8;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t lboA, sboA, lboB, sboB;
    if (mode == 1) {  // [k-block][row-block][8][8]
        lboA = (M / 8) * 128; sboA = 128;
        lboB = (N / 8) * 128; sboB = 128;
    } else {          // [row-block][k-block][8][8]
        lboA = 128; sboA = KB * 128;
        lboB = 128; sboB = KB * 128;
    }
    if (mode == 2 || mode == 3) {
        if (tid == 0) {
            uint32_t bytesA = (uint32_t)M * K * 2, bytesB = (uint32_t)N * K * 2;
            mbar_expect_tx(&bars[0], bytesA + bytesB);
            bulk_g2s(sA, A, bytesA, &bars[0]);
            bulk_g2s(sB, B, bytesB, &bars[0]);
        }
        mbar_wait(&bars[0], 0);
    } else {
        for (int i = tid; i < M * K; i += blockDim.x) {
            int m = i / K, k = i - m * K;
            size_t off = (mode == 1) ? ((size_t)(k / 8) * (M / 8) + m / 8) * 64 + (m % 8) * 8 + (k % 8)
                                     : ((size_t)(m / 8) * KB + k / 8) * 64 + (m % 8) * 8 + (k % 8);
            sA[off] = f2bf(A[i]);
        }
        for (int i = tid; i < N * K; i += blockDim.x) {
            int n = i / K, k = i - n * K;
            size_t off = (mode == 1) ? ((size_t)(k / 8) * (N / 8) + n / 8) * 64 + (n % 8) * 8 + (k % 8)
                                     : ((size_t)(n / 8) * KB + k / 8) * 64 + (n % 8) * 8 + (k % 8);
            sB[off] = f2bf(B[i]);
        }
        fence_proxy_async_smem();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = idesc_bf16_f32(M, N);
        const uint32_t kstepA = (mode == 1) ? 2 * lboA : 2 * 128;  // two k-blocks per MMA
        const uint32_t kstepB = (mode == 1) ? 2 * lboB : 2 * 128;
        for (int k16 = 0; k16 < K / 16; ++k16) {
            uint64_t da = smem_desc(smem_u32(sA) + k16 * kstepA, lboA, sboA);
            uint64_t db = smem_desc(smem_u32(sB) + k16 * kstepB, lboB, sboB);
            if (mode == 3) {   // A staged into TMEM columns [64 + 8*k16, +8) by tcgen05.cp, then a TS-mode MMA
                tmem_cp_128x256b(tmem + 64 + 8 * k16, da);
                mma_bf16_ts(tmem, tmem + 64 + 8 * k16, db, idesc, k16 > 0);
            } else {
                mma_bf16_ss(tmem, da, db, idesc, k16 > 0);
            }
        }
        mma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 8) {
        float v[8];
        tmem_ld_x8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 8; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace cvb

extern "C" int cvb_selftest_umma(int mode, int N, int K, const void* A, const void* B, float* D, void* stream) {
    using namespace cvb;
    CVB_REQUIRE(mode >= 0 && mode <= 3, "mode must be 0..3");
    CVB_REQUIRE(N % 16 == 0 && N >= 16 && N <= 64 && K % 16 == 0 && K >= 16 && K <= 512, "unsupported N=%d K=%d", N, K);
    size_t smem = (size_t)128 * K * 2 + (size_t)N * K * 2 + 64;
    CVB_CHECK(cudaFuncSetAttribute(k_umma_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_umma_selftest<<<1, 128, smem, (cudaStream_t)stream>>>(mode, N, K, (const float*)A, (const float*)B, D);
    CVB_LAUNCH_CHECK();
    return 0;
}
