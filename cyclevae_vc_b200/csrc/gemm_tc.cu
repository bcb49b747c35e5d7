// Split-precision tensor-core GEMM for the throughput-bound dense products of the path (gx = xc W_x^T,
// conv taps, deferred weight gradients of BPTT, dxc = dgi W_x):  C[M,N] (+)= op(A) op(B) (+ bias), fp32 in/out.
//
// fp32 parity on tensor cores: every operand element is split x = hi + lo into two
// 16-bit floats (fp16 for the forward products: 2 x 11 mantissa bits; bf16 for gradients: fp32's exponent
// range), and C = A_hi B_hi + A_lo B_hi + A_hi B_lo with fp32 accumulation in TMEM.
//
//   1. k_split_tiles (one pass per operand, HBM-bound): fp32 row-major (optionally transposed) -> 16-bit hi/lo
//      in "tile order": blocks of 128 rows x 64 k, each block already in the UMMA K-major core-matrix layout
//      [16 row groups][8 k blocks][8 rows][8 k] with the hi block followed by the lo block.  A (block, part)
//      is contiguous, so the GEMM needs no tensor maps: one cp.async.bulk per operand per stage.
//   2. k_gemm_tc (persistent, warp-specialised): w0 bulk-copy producer (3-stage ring, 64 KB per stage),
//      w1 MMA issuer, w2 TMEM allocator, w4-7 epilogue.  B's [hi | lo] blocks are adjacent in shared memory,
//      so ONE tcgen05.mma with N = 256 forms A_hi B_hi (columns 0..127) and A_hi B_lo (columns 128..255) and a
//      second with N = 128 adds A_lo B_hi: 2 MMAs per K step instead of 3.  Two 256-column accumulators
//      alternate so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <mutex>

#include "common.cuh"
#include "umma.cuh"

namespace cvb {
using namespace umma;

constexpr int GT_BM = 128, GT_BN = 128, GT_BK = 64;
constexpr int GT_BLOCK_ELEMS = GT_BM * GT_BK;          // one part of one block
constexpr int GT_STAGE_BYTES = 4 * GT_BLOCK_ELEMS * 2;  // A hi, A lo, B hi, B lo
constexpr int GT_NS = 3;
constexpr int GT_KD = 2;   // chunks (of 64) accumulated inside the tensor core before the slice is added in fp32 registers
constexpr int GT_THREADS = 256;

// ---- operand preparation ---------------------------------------------------------------------------------
// X: row-major [R, K] with leading dimension ld (transposed = false) or [K, R] (transposed = true).
// out: [ceil(R/128)][ceil(K/64)][2 parts][16][8][8][8] 16-bit.
template <bool F16>
__global__ void __launch_bounds__(256) k_split_tiles(const float* __restrict__ X, int ld, int R, int K, int transposed, int KC,
                                                     uint16_t* __restrict__ out, const float* __restrict__ X2, int ld2, int R1,
                                                     int conv_ci, int conv_dshift) {
    // transposed sources may be two matrices side by side: rows [0, R1) from X, rows [R1, R) from X2 (e.g. [xc | y] for dW_ih).
    // conv_ci > 0: the operand is the virtual im2col of a dilated conv on the flattened padded grid (frontend.cu): logical
    // element (row r, column kk = tap*ci + c) lives at X[(r + tap*conv_dshift)*ci + c]; `transposed` then only says which of
    // (r, kk) is the tile's row index.
    __shared__ float T[GT_BM * (GT_BK + 1)];
    const int rt = blockIdx.y, kc = blockIdx.x;
    const int r0 = rt * GT_BM, k0 = kc * GT_BK;
    if (!transposed) {
        for (int i = threadIdx.x; i < GT_BM * GT_BK; i += 256) {
            const int r = i >> 6, k = i & 63;
            const bool ok = (r0 + r < R) && (k0 + k < K);
            float v = 0.f;
            if (ok) {
                if (conv_ci > 0) {
                    const int kk = k0 + k, tap = kk / conv_ci, cch = kk - tap * conv_ci;
                    v = X[((size_t)(r0 + r) + (size_t)tap * conv_dshift) * conv_ci + cch];
                } else {
                    v = X[(size_t)(r0 + r) * ld + k0 + k];
                }
            }
            T[r * (GT_BK + 1) + k] = v;
        }
    } else {
        for (int i = threadIdx.x; i < GT_BM * GT_BK; i += 256) {
            const int k = i >> 7, r = i & 127;
            const bool ok = (r0 + r < R) && (k0 + k < K);
            float v = 0.f;
            if (ok) {
                if (conv_ci > 0) {   // tile row = kk (tap, channel), tile k = grid row
                    const int kk = r0 + r, tap = kk / conv_ci, cch = kk - tap * conv_ci;
                    v = X[((size_t)(k0 + k) + (size_t)tap * conv_dshift) * conv_ci + cch];
                } else {
                    v = (r0 + r < R1) ? X[(size_t)(k0 + k) * ld + r0 + r] : X2[(size_t)(k0 + k) * ld2 + (r0 + r - R1)];
                }
            }
            T[r * (GT_BK + 1) + k] = v;
        }
    }
    __syncthreads();
    uint16_t* blk = out + ((size_t)rt * KC + kc) * (2 * GT_BLOCK_ELEMS);
    for (int p = threadIdx.x; p < GT_BLOCK_ELEMS / 8; p += 256) {
        // 16-byte piece p of the block: row group p/64, k block (p/8)%8, row p%8
        const int r = (p >> 6) * 8 + (p & 7), kb = (p >> 3) & 7;
        const float* src = T + r * (GT_BK + 1) + kb * 8;
        uint16_t h[8], l[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (F16) split_f16(src[q], h[q], l[q]);
            else split_bf16(src[q], h[q], l[q]);
        }
        const uint4 hv = make_uint4((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16),
                                    (uint32_t)h[4] | ((uint32_t)h[5] << 16), (uint32_t)h[6] | ((uint32_t)h[7] << 16));
        const uint4 lv = make_uint4((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16),
                                    (uint32_t)l[4] | ((uint32_t)l[5] << 16), (uint32_t)l[6] | ((uint32_t)l[7] << 16));
        reinterpret_cast<uint4*>(blk)[p] = hv;
        reinterpret_cast<uint4*>(blk + GT_BLOCK_ELEMS)[p] = lv;
    }
}

// ---- the GEMM ------------------------------------------------------------------------------------------------
struct GemmTcArgs {
    const uint16_t* At;
    const uint16_t* Bt;
    float* C;
    const float* bias;
    int M, N, ldc, MT, NTl, KC;
    int beta1;
    int f16;
};

__global__ void __launch_bounds__(GT_THREADS, 1) k_gemm_tc(GemmTcArgs g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + GT_NS * GT_STAGE_BYTES);
    uint64_t* empty = full + GT_NS;
    uint64_t* tmem_full = empty + GT_NS;    // [2]
    uint64_t* tmem_empty = tmem_full + 2;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int n_tiles = g.MT * g.NTl;

    if (threadIdx.x == 0) {
        for (int s = 0; s < GT_NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 128);
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ================= producer =====================================================================
        int s = 0;
        uint32_t ph = 1;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int nt = tile / g.MT, mt = tile - nt * g.MT;
            const uint16_t* a_src = g.At + (size_t)mt * g.KC * (2 * GT_BLOCK_ELEMS);
            const uint16_t* b_src = g.Bt + (size_t)nt * g.KC * (2 * GT_BLOCK_ELEMS);
            for (int kc = 0; kc < g.KC; ++kc) {
                if (lane == 0) {
                    mbar_wait(&empty[s], ph);
                    uint8_t* dst = smem + (size_t)s * GT_STAGE_BYTES;
                    mbar_expect_tx(&full[s], GT_STAGE_BYTES);
                    bulk_g2s(dst, a_src + (size_t)kc * (2 * GT_BLOCK_ELEMS), GT_STAGE_BYTES / 2, &full[s]);
                    bulk_g2s(dst + GT_STAGE_BYTES / 2, b_src + (size_t)kc * (2 * GT_BLOCK_ELEMS), GT_STAGE_BYTES / 2, &full[s]);
                }
                __syncwarp();
                if (++s == GT_NS) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =====================================================================
        // The accumulation chain inside the tensor core rounds toward zero (measured: error grows ~K / 2^22), so a
        // TMEM accumulator only ever sums GT_KD chunks (K = 128); the epilogue warps add the slices in fp32 registers
        // with round-to-nearest while the other accumulator buffer receives the next slice.
        const uint32_t idesc_s = g.f16 ? idesc_f16_f32(128, 256) : idesc_bf16_f32(128, 256);   // B rows [hi | lo]
        const uint32_t idesc_h = g.f16 ? idesc_f16_f32(128, 128) : idesc_bf16_f32(128, 128);   // B hi rows only
        const uint64_t d0 = smem_desc(smem_u32(smem), 128, 1024);
        int s = 0, acc = 0;
        uint32_t ph = 0, acc_ph = 1;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int kc = 0; kc < g.KC; ++kc) {
                const bool slice_start = (kc % GT_KD) == 0;
                const bool slice_end = ((kc + 1) % GT_KD) == 0 || kc == g.KC - 1;
                if (slice_start) {
                    if (lane == 0) mbar_wait(&tmem_empty[acc], acc_ph);
                    __syncwarp();
                }
                if (lane == 0) mbar_wait(&full[s], ph);
                __syncwarp();
                tc_fence_after();
                const uint32_t d_tmem = tmem + (uint32_t)acc * 256u;
                const uint64_t da = d0 + (uint64_t)((uint32_t)s * (GT_STAGE_BYTES >> 4));
                const uint64_t db = da + (uint64_t)(GT_STAGE_BYTES >> 5);
#pragma unroll
                for (int k16 = 0; k16 < GT_BK / 16; ++k16) {
                    mma_bf16_ss_elect(d_tmem, da + 16u * k16, db + 16u * k16, idesc_s, !(slice_start && k16 == 0));
                    mma_bf16_ss_elect(d_tmem, da + (uint32_t)(GT_BLOCK_ELEMS * 2 >> 4) + 16u * k16, db + 16u * k16, idesc_h, true);
                }
                mma_commit_elect(&empty[s]);
                if (++s == GT_NS) {
                    s = 0;
                    ph ^= 1;
                }
                if (slice_end) {
                    mma_commit_elect(&tmem_full[acc]);
                    if (++acc == 2) {
                        acc = 0;
                        acc_ph ^= 1;
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: TMEM lane == row of the tile =========================================
        int acc = 0;
        uint32_t acc_ph = 0;
        const bool vec_ok = (g.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
        const int n_slices = (g.KC + GT_KD - 1) / GT_KD;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int nt = tile / g.MT, mt = tile - nt * g.MT;
            const int row = mt * GT_BM + (warp - 4) * 32 + lane;
            const int n0 = nt * GT_BN;
            float sum[GT_BN];
#pragma unroll
            for (int q = 0; q < GT_BN; ++q) sum[q] = 0.f;
            for (int sl = 0; sl < n_slices; ++sl) {
                mbar_wait(&tmem_full[acc], acc_ph);
                tc_fence_after();
                const uint32_t taddr = tmem + (uint32_t)acc * 256u + ((uint32_t)((warp - 4) * 32) << 16);
#pragma unroll
                for (int c0 = 0; c0 < GT_BN; c0 += 16) {
                    float v[16], v2[16];
                    tmem_ld_x16(taddr + c0, v);
                    tmem_ld_x16(taddr + 128 + c0, v2);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 16; ++q) sum[c0 + q] += v[q] + v2[q];
                }
                tc_fence_before();
                mbar_arrive(&tmem_empty[acc]);
                if (++acc == 2) {
                    acc = 0;
                    acc_ph ^= 1;
                }
            }
            if (row < g.M) {
                float* crow = g.C + (size_t)row * g.ldc + n0;
#pragma unroll
                for (int c0 = 0; c0 < GT_BN; c0 += 4) {
                    if (n0 + c0 < g.N) {
                        float o[4] = {sum[c0], sum[c0 + 1], sum[c0 + 2], sum[c0 + 3]};
                        if (g.bias) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (n0 + c0 + q < g.N) o[q] += g.bias[n0 + c0 + q];
                        }
                        if (vec_ok && n0 + c0 + 4 <= g.N) {
                            float4 w = make_float4(o[0], o[1], o[2], o[3]);
                            if (g.beta1) {
                                const float4 old = *reinterpret_cast<const float4*>(crow + c0);
                                w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w;
                            }
                            *reinterpret_cast<float4*>(crow + c0) = w;
                        } else {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (n0 + c0 + q < g.N) crow[c0 + q] = g.beta1 ? crow[c0 + q] + o[q] : o[q];
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<512>(tmem);
}

// ---- host ------------------------------------------------------------------------------------------------------
struct TcWorkspace {
    uint16_t* buf[2] = {nullptr, nullptr};
    size_t cap[2] = {0, 0};
};
static std::mutex g_ws_mu;
static TcWorkspace g_ws[64];

static int ws_get(int which, size_t elems, uint16_t** out) {
    int dev = 0;
    CVB_CHECK(cudaGetDevice(&dev));
    CVB_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(g_ws_mu);
    TcWorkspace& w = g_ws[dev];
    if (w.cap[which] < elems) {
        if (w.buf[which]) CVB_CHECK(cudaFree(w.buf[which]));   // synchronises: earlier GEMMs are done with it
        w.buf[which] = nullptr;
        w.cap[which] = 0;
        const size_t want = elems + elems / 4;
        CVB_CHECK(cudaMalloc(&w.buf[which], want * sizeof(uint16_t)));
        w.cap[which] = want;
    }
    *out = w.buf[which];
    return 0;
}

// measured on B200 (tools/bench_gemm.py): the two operand passes + the GEMM beat cuBLAS fp32 3-5x from ~3e9 MACs up
// (gx, dW_hh, dW_x, dxc at the training shapes) and lose below ~1.5e9 (conv taps, dW_y, dW_o)
bool gemm_tc_eligible(int M, int N, int K) {
    return M >= 128 && N >= 128 && K >= 128 && (double)M * N * K >= 3.0e9;
}

// C[M,N] = op(A) op(B) (+ C if beta1) (+ bias[N]);  A: [M,K] (lda) or [K,M] if transA;  B: [N,K] (ldb) if transB else [K,N].
int gemm_tc(cudaStream_t s, bool transA, bool transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
            bool beta1, const float* bias, float* C, int ldc, bool f16, const float* B2, int ldb2, int N1, const ConvGather* gA,
            const ConvGather* gB) {
    CVB_REQUIRE(!B2 || !transB, "gemm_tc: a second B source needs B stored [K,N]");
    CVB_REQUIRE(!gA || !transA, "gemm_tc: a conv-gathered A is addressed as [M,K]");
    CVB_REQUIRE(!gB || (!transB && !B2), "gemm_tc: a conv-gathered B is addressed as [K,N]");
    if (!B2) N1 = N;
    const int a_ci = gA ? gA->ci : 0, a_ds = gA ? gA->dshift : 0, b_ci = gB ? gB->ci : 0, b_ds = gB ? gB->dshift : 0;
    const int MT = ceil_div(M, GT_BM), NTl = ceil_div(N, GT_BN), KC = ceil_div(K, GT_BK);
    uint16_t *At, *Bt;
    if (int rc = ws_get(0, (size_t)MT * KC * 2 * GT_BLOCK_ELEMS, &At)) return rc;
    if (int rc = ws_get(1, (size_t)NTl * KC * 2 * GT_BLOCK_ELEMS, &Bt)) return rc;
    // A as [M rows, K]: stored [M,K] when !transA, [K,M] when transA.  B as [N rows, K]: stored [N,K] when transB.
    if (f16) {
        k_split_tiles<true><<<dim3(KC, MT), 256, 0, s>>>(A, lda, M, K, transA ? 1 : 0, KC, At, nullptr, 0, M, a_ci, a_ds);
        CVB_LAUNCH_CHECK();
        k_split_tiles<true><<<dim3(KC, NTl), 256, 0, s>>>(B, ldb, N, K, transB ? 0 : 1, KC, Bt, B2, ldb2, N1, b_ci, b_ds);
        CVB_LAUNCH_CHECK();
    } else {
        k_split_tiles<false><<<dim3(KC, MT), 256, 0, s>>>(A, lda, M, K, transA ? 1 : 0, KC, At, nullptr, 0, M, a_ci, a_ds);
        CVB_LAUNCH_CHECK();
        k_split_tiles<false><<<dim3(KC, NTl), 256, 0, s>>>(B, ldb, N, K, transB ? 0 : 1, KC, Bt, B2, ldb2, N1, b_ci, b_ds);
        CVB_LAUNCH_CHECK();
    }
    DeviceInfo di;
    if (int rc = get_device_info(&di)) return rc;
    const int smem = GT_NS * GT_STAGE_BYTES + 256;
    static bool attr_set[64] = {};   // function attributes are per device
    int dev = 0;
    CVB_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        CVB_CHECK(cudaFuncSetAttribute(k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    GemmTcArgs g;
    g.At = At;
    g.Bt = Bt;
    g.C = C;
    g.bias = bias;
    g.M = M;
    g.N = N;
    g.ldc = ldc;
    g.MT = MT;
    g.NTl = NTl;
    g.KC = KC;
    g.beta1 = beta1 ? 1 : 0;
    g.f16 = f16 ? 1 : 0;
    const int tiles = MT * NTl;
    const int grid = tiles < di.n_sm ? tiles : di.n_sm;
    k_gemm_tc<<<grid, GT_THREADS, smem, s>>>(g);
    CVB_LAUNCH_CHECK();
    return 0;
}

}  // namespace cvb
