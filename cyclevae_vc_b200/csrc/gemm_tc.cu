// Split-precision tensor-core GEMM for the throughput-bound dense products of the path (gx = xc W_x^T, conv taps,
// deferred weight gradients of BPTT, dxc = dgi W_x, y = H W_o^T ...):  C[M,N] (+)= alpha op(A) op(B) (+ bias), fp32 in/out.
//
// fp32 parity on tensor cores: every operand element is split x = hi + lo into two 16-bit
// floats (fp16 for the forward products: 2 x 11 mantissa bits, lo stored scaled by 2^11; bf16 for gradients: fp32's
// exponent range) and C = A_hi B_hi + (A_hi B_lo + A_lo B_hi) with fp32 accumulation in TMEM.
//
//   1. k_split_group (HBM-bound, ONE launch for every operand of a group of products): fp32 operand (row-major,
//      transposed, two sources side by side, or the virtual im2col of a dilated conv) -> 16-bit hi/lo "image" in tile
//      order: blocks of 128 rows x 64 k, each already in the UMMA K-major core-matrix layout
//      [16 row groups][8 k groups][8 rows][8 k], hi plane then lo plane.  No shared memory: a thread reads 8 consecutive k
//      of one row (vector loads; for transposed sources the lanes run along the rows, so every load is one full line)
//      and writes one 16-byte core-matrix row per plane.  Operands that appear more than once in a group are split once
//      (dgi^T feeds both dW_hh and dW_ih); parameter operands keep their image until the parameters change
//      (cvb_weights_changed / cvb_adam_step), so W_x is split once per optimiser step, not once per pass.
//   2. k_gemm_tc (persistent, warp-specialised, up to 6 products per launch): w0 bulk-copy producer (3-stage ring, 64 KB
//      per stage: one cp.async.bulk per operand block), w1 MMA issuer, w2 TMEM allocator, w4-11 epilogue (two warps per TMEM lane quadrant, 64 columns each).  B's [hi | lo]
//      planes are adjacent in shared memory, so ONE tcgen05.mma with N = 256 forms A_hi B_hi (columns 0..127) and
//      A_hi B_lo (columns 128..255) and a second with N = 128 adds A_lo B_hi onto the correction columns.  Two 256-column
//      accumulators alternate; an accumulator only ever sums K = 128 (a tcgen05 accumulation chain truncates toward zero:
//      the error of a 6400-deep chain was measured at 9e-5 relative) -- the epilogue warps add the slices in fp32
//      registers with round-to-nearest while the other accumulator fills.
// Measured and rejected (round 2): splitting inside the GEMM's loader warps (fp32 -> registers -> hi/lo planes in the
// ring).  Correct, but registers cannot hold enough bytes in flight: 256 loader threads x 128 B against ~1500 cycles of
// loaded L2/HBM latency feed 22 B/clk per SM where the tile needs 42; 5800 cycles per stage against 1560 for bulk copies
// of pre-split images (gx 358 us vs 150 us).  The operand pass stays, but as one fused, deduplicated, cached launch.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "umma.cuh"

namespace cvb {
using namespace umma;

constexpr int GT_BM = 128, GT_BN = 128, GT_BK = 64;
constexpr int GT_PLANE_BYTES = GT_BM * GT_BK * 2;        // one 16-bit plane of one operand block
constexpr int GT_BLOCK_BYTES = 2 * GT_PLANE_BYTES;        // hi + lo
constexpr int GT_STAGE_BYTES = 2 * GT_BLOCK_BYTES;        // A block, B block
constexpr int GT_NS = 3;
constexpr int GT_KD = 2;   // chunks (of 64) accumulated inside the tensor core before the slice is added in fp32 registers
constexpr int GT_THREADS = 384;   // w0 producer, w1 MMA issuer, w2 TMEM allocator, w3 idle, w4-11 epilogue
constexpr int GT_EPI_BYTES = 8 * 32 * 32 * 4;   // staging of the 8 epilogue warps: 32 rows x 32 columns each, swizzled
constexpr int GT_MAXP = 6;                                // products per launch
constexpr int GT_MAXO = 2 * GT_MAXP;                      // operands per split launch

// one fp32 operand seen as [rows, K], and where its 16-bit image goes
struct GtOperand {
    const float* p;
    const float* p2;   // rows >= R1 come from p2 (row index - R1), only for kmajor == 0
    uint16_t* img;     // [ceil(rows/128)][KC][hi plane | lo plane]
    int ld, ld2, R1;
    int rows, K, KC;
    int kmajor;        // 1: element (r, k) at p[r*ld + k];  0: at p[k*ld + r]
    int conv_ci;       // > 0: virtual im2col of a dilated conv on the flattened padded grid (frontend.cu):
    int conv_ds;       //      kmajor: (r, kk = tap*ci + c) at p[(r + tap*ds)*ci + c];  !kmajor: (r = kk, k) at p[(k + tap*ds)*ci + c]
    int vec;           // widest aligned vector load of 8 consecutive k (kmajor, no conv): 4, 2 or 1 floats
    int f16;
    int block_end;     // running block count of the split launch up to and including this operand
};
struct SplitArgs {
    GtOperand o[GT_MAXO];
    int n;
};

// ---- operand pass ----------------------------------------------------------------------------------------------
// 256 threads per 128 x 64 block, 4 "items" (8 consecutive k of one row) per thread:
//   K contiguous in memory (kmajor): 8 lanes = the 8 rows of one core matrix (conflict-free / coalesced 16-byte stores,
//     32-byte global segments): row = (lt >> 6) * 8 + (lt & 7) + 16 * it, k group = (lt >> 3) & 7
//   rows contiguous (!kmajor): lanes = consecutive rows (every load of a warp is one 128-byte line): row = lt, k group = it
// with lt = thread & 127 and it = 4 * (thread >> 7) + i.
static __device__ __forceinline__ void gt_zero8(float* v) {
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = 0.f;
}

// generic element fetch of the conv-gather operands (virtual im2col; small products only)
static __device__ __forceinline__ void gt_load8_conv(const GtOperand& o, int r, int k0, float* v) {
    gt_zero8(v);
    if (r >= o.rows || k0 >= o.K) return;
    if (o.kmajor) {
        int tap = k0 / o.conv_ci, c = k0 - tap * o.conv_ci;
        const float* base = o.p + ((size_t)r + (size_t)tap * o.conv_ds) * o.conv_ci;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (k0 + q < o.K) v[q] = __ldg(base + c);
            if (++c == o.conv_ci) {
                c = 0;
                base += (size_t)o.conv_ds * o.conv_ci;
            }
        }
    } else {
        const int tap = r / o.conv_ci, c = r - tap * o.conv_ci;
        const float* base = o.p + ((size_t)k0 + (size_t)tap * o.conv_ds) * o.conv_ci + c;
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (k0 + q < o.K) v[q] = __ldg(base + (size_t)q * o.conv_ci);
    }
}

// v[0..8) -> one 16-byte core-matrix row of the hi plane and one of the lo plane
template <bool F16>
static __device__ __forceinline__ void gt_split_store(const float* v, uint8_t* hi_dst) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (F16) {
            const float a = fminf(fmaxf(v[2 * q], -60000.f), 60000.f), b = fminf(fmaxf(v[2 * q + 1], -60000.f), 60000.f);
            const __half2 hh = __floats2half2_rn(a, b);
            const float2 hf = __half22float2(hh);
            const __half2 ll = __floats2half2_rn((a - hf.x) * F16_LO_SCALE, (b - hf.y) * F16_LO_SCALE);
            h[q] = *reinterpret_cast<const uint32_t*>(&hh);
            l[q] = *reinterpret_cast<const uint32_t*>(&ll);
        } else {
            const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
            const uint32_t hb = *reinterpret_cast<const uint32_t*>(&hh);
            const float r0 = v[2 * q] - __uint_as_float(hb << 16), r1 = v[2 * q + 1] - __uint_as_float(hb & 0xffff0000u);
            const __nv_bfloat162 ll = __floats2bfloat162_rn(r0, r1);
            h[q] = hb;
            l[q] = *reinterpret_cast<const uint32_t*>(&ll);
        }
    }
    *reinterpret_cast<uint4*>(hi_dst) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(hi_dst + GT_PLANE_BYTES) = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(256) k_split_group(const __grid_constant__ SplitArgs a) {
    int oi = 0;
    while ((int)blockIdx.x >= a.o[oi].block_end) ++oi;
    const GtOperand& o = a.o[oi];
    const int local = (int)blockIdx.x - (oi ? a.o[oi - 1].block_end : 0);
    const int rt = local / o.KC, kc = local - rt * o.KC;
    const int lt = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int r0 = rt * GT_BM, k0 = kc * GT_BK;
    float v[32];
    bool kmajor = o.kmajor != 0;
    if (o.conv_ci > 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int it = half * 4 + i;
            const int r = kmajor ? (lt >> 6) * 8 + (lt & 7) + 16 * it : lt;
            const int kg = kmajor ? (lt >> 3) & 7 : it;
            gt_load8_conv(o, r0 + r, k0 + kg * 8, v + 8 * i);
        }
    } else if (kmajor) {
        const int rb = (lt >> 6) * 8 + (lt & 7), kg = (lt >> 3) & 7;
        const int kleft = o.K - k0 - kg * 8;
        const float* p = o.p + (size_t)(r0 + rb) * o.ld + k0 + kg * 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float* d = v + 8 * i;
            const int it = half * 4 + i;
            const float* src = p + (size_t)it * 16 * o.ld;
            const bool row_ok = r0 + rb + 16 * it < o.rows;
            if (row_ok && kleft >= 8 && o.vec == 4) {
                const float4 x = __ldg(reinterpret_cast<const float4*>(src)), y = __ldg(reinterpret_cast<const float4*>(src) + 1);
                d[0] = x.x; d[1] = x.y; d[2] = x.z; d[3] = x.w; d[4] = y.x; d[5] = y.y; d[6] = y.z; d[7] = y.w;
            } else if (row_ok && kleft >= 8 && o.vec == 2) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 x = __ldg(reinterpret_cast<const float2*>(src) + q);
                    d[2 * q] = x.x;
                    d[2 * q + 1] = x.y;
                }
            } else {
                gt_zero8(d);
                if (row_ok) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (q < kleft) d[q] = __ldg(src + q);
                }
            }
        }
    } else {
        const int r = r0 + lt;
        const bool row_ok = r < o.rows;
        const bool second = r >= o.R1;
        const int ld = second ? o.ld2 : o.ld;
        const float* p = (second ? o.p2 + (r - o.R1) : o.p + r) + (size_t)k0 * ld;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float* d = v + 8 * i;
            const int it = half * 4 + i;
            const int kl = o.K - k0 - it * 8;   // warp-uniform
            const float* src = p + (size_t)it * 8 * ld;
            if (row_ok && kl >= 8) {
#pragma unroll
                for (int q = 0; q < 8; ++q) d[q] = __ldg(src + (size_t)q * ld);
            } else {
                gt_zero8(d);
                if (row_ok) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (q < kl) d[q] = __ldg(src + (size_t)q * ld);
                }
            }
        }
    }
    // kmajor: item it -> row group 2*it + (lt >> 6), k group (lt >> 3) & 7, row lt & 7;  !kmajor: row lt, k group it
    uint8_t* blk = reinterpret_cast<uint8_t*>(o.img) + ((size_t)rt * o.KC + kc) * GT_BLOCK_BYTES;
    uint8_t* d = kmajor ? blk + ((half * 8 + (lt >> 6)) * 1024 + ((lt >> 3) & 7) * 128 + (lt & 7) * 16)
                        : blk + ((lt >> 3) * 1024 + half * 4 * 128 + (lt & 7) * 16);
    const int step = kmajor ? 2048 : 128;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (o.f16) gt_split_store<true>(v + 8 * i, d + i * step);
        else gt_split_store<false>(v + 8 * i, d + i * step);
    }
}

// ---- the GEMM ------------------------------------------------------------------------------------------------
struct GtProblem {
    const uint16_t* At;
    const uint16_t* Bt;
    float* C;
    const float* bias;
    float alpha;
    int M, N, ldc, MT, NT, KC;
    int beta1, f16;
    int cm, cn;        // the launch's cluster of cm x cn CTAs covers cm row tiles x cn column tiles of this product
    int MTc, NTc;      // cluster tiles: ceil(MT / cm), ceil(NT / cn)
    int tile_end;      // running CLUSTER-tile count up to and including this product
    int map_Tp, map_T, map_B;   // output row map (GemmDesc), 0 = identity
    const float* mask;
    // split-K (few tiles, deep K): split s sums chunks [s * kc_split, ...) into part[s][M][N]; k_splitk_reduce finishes
    int S, kc_split;
    float* part;
};
struct GemmTcArgs {
    GtProblem p[GT_MAXP];
    int n_prob;
    int cs;            // CTAs per thread-block cluster (1, 2 or 4)
};

// walk of the tile sequence of one persistent CTA, shared by the three roles.  The CTAs take the tile list in
// boustrophedon order (wave w of gridDim tiles forwards, wave w+1 backwards): with the long-K products listed first the
// CTAs that got an extra long tile are the last to be handed a short one.
struct GtWalk {
    int wave, tile, prob, mt, nt, sp, kc0, kc1;   // the tile's chunk range [kc0, kc1) (split sp of a split-K product)
    int ci, cj;        // this CTA's row / column inside the cluster tile (rank = ci * cn + cj)
    bool valid;        // false: the cluster tile hangs over the edge of the product here (operands clamped, nothing stored)
    int ncl, cid, rank;
    __device__ __forceinline__ void start(const GemmTcArgs& g) {
        wave = -1;
        ncl = (int)gridDim.x / g.cs;
        cid = (int)blockIdx.x / g.cs;
        rank = (int)blockIdx.x - cid * g.cs;   // == %cluster_ctarank for a 1-D cluster
    }
    __device__ __forceinline__ bool next(const GemmTcArgs& g) {
        ++wave;
        tile = wave * ncl + ((wave & 1) ? ncl - 1 - cid : cid);
        if (tile >= g.p[g.n_prob - 1].tile_end) return false;
        prob = 0;
        while (tile >= g.p[prob].tile_end) ++prob;
        const GtProblem& P = g.p[prob];
        int local = tile - (prob ? g.p[prob - 1].tile_end : 0);
        const int per = P.MTc * P.NTc;
        sp = local / per;
        local -= sp * per;
        const int ntc = local / P.MTc, mtc = local - ntc * P.MTc;
        ci = rank / P.cn;
        cj = rank - ci * P.cn;
        mt = mtc * P.cm + ci;
        nt = ntc * P.cn + cj;
        valid = mt < P.MT && nt < P.NT;
        if (mt >= P.MT) mt = P.MT - 1;
        if (nt >= P.NT) nt = P.NT - 1;
        kc0 = sp * P.kc_split;
        kc1 = kc0 + P.kc_split < P.KC ? kc0 + P.kc_split : P.KC;
        return true;
    }
};

__global__ void __launch_bounds__(GT_THREADS, 1) k_gemm_tc(const __grid_constant__ GemmTcArgs g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + GT_NS * GT_STAGE_BYTES + GT_EPI_BYTES);
    uint64_t* empty = full + GT_NS;
    uint64_t* tmem_full = empty + GT_NS;    // [2]
    uint64_t* tmem_empty = tmem_full + 2;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < GT_NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], (uint32_t)g.cs);   // every CTA of the cluster releases a stage (its operands are multicast)
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 256);
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    if (g.cs > 1) cluster_sync_all();   // nobody multicasts into a CTA whose barriers are not initialised yet
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ================= producer =====================================================================
        int s = 0;
        uint32_t ph = 1;
        GtWalk w;
        for (w.start(g); w.next(g);) {
            const GtProblem& P = g.p[w.prob];
            const uint8_t* a_src = reinterpret_cast<const uint8_t*>(P.At) + (size_t)w.mt * P.KC * GT_BLOCK_BYTES;
            const uint8_t* b_src = reinterpret_cast<const uint8_t*>(P.Bt) + (size_t)w.nt * P.KC * GT_BLOCK_BYTES;
            const uint16_t row_mask = (uint16_t)(((1u << P.cn) - 1u) << (w.ci * P.cn));
            uint16_t col_mask = 0;
            for (int i = 0; i < P.cm; ++i) col_mask |= (uint16_t)(1u << (i * P.cn + w.cj));
            for (int kc = w.kc0; kc < w.kc1; ++kc) {
                if (lane == 0) {
                    mbar_wait(&empty[s], ph);
                    uint8_t* dst = smem + (size_t)s * GT_STAGE_BYTES;
                    mbar_expect_tx(&full[s], GT_STAGE_BYTES);
                    // the A block is wanted by the cn CTAs of this cluster row, the B block by the cm CTAs of this column:
                    // each pulls one slice and multicasts it (the per-SM request rate, not the delivered bytes, is what
                    // bounds the ingest: tools/bench_mcast.py)
                    const uint8_t* ab = a_src + (size_t)kc * GT_BLOCK_BYTES;
                    const uint8_t* bb = b_src + (size_t)kc * GT_BLOCK_BYTES;
                    if (P.cn == 1) {
                        bulk_g2s(dst, ab, GT_BLOCK_BYTES, &full[s]);
                    } else {
                        const uint32_t sl = GT_BLOCK_BYTES / (uint32_t)P.cn;
                        bulk_g2s_multicast(dst + w.cj * sl, ab + w.cj * sl, sl, &full[s], row_mask);
                    }
                    if (P.cm == 1) {
                        bulk_g2s(dst + GT_BLOCK_BYTES, bb, GT_BLOCK_BYTES, &full[s]);
                    } else {
                        const uint32_t sl = GT_BLOCK_BYTES / (uint32_t)P.cm;
                        bulk_g2s_multicast(dst + GT_BLOCK_BYTES + w.ci * sl, bb + w.ci * sl, sl, &full[s], col_mask);
                    }
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
        const uint64_t d0 = smem_desc(smem_u32(smem), 128, 1024);
        int s = 0, acc = 0;
        uint32_t ph = 0, acc_ph = 1;
        GtWalk w;
        for (w.start(g); w.next(g);) {
            const GtProblem& P = g.p[w.prob];
            const uint32_t idesc_s = P.f16 ? idesc_f16_f32(128, 256) : idesc_bf16_f32(128, 256);   // B rows [hi | lo]
            const uint32_t idesc_h = P.f16 ? idesc_f16_f32(128, 128) : idesc_bf16_f32(128, 128);   // B hi rows only
            const int nkc = w.kc1 - w.kc0;
            for (int kc = 0; kc < nkc; ++kc) {
                const bool slice_start = (kc % GT_KD) == 0;
                const bool slice_end = ((kc + 1) % GT_KD) == 0 || kc == nkc - 1;
                if (slice_start) {
                    if (lane == 0) mbar_wait(&tmem_empty[acc], acc_ph);
                    __syncwarp();
                }
                if (lane == 0) mbar_wait(&full[s], ph);
                __syncwarp();
                tc_fence_after();
                const uint32_t d_tmem = tmem + (uint32_t)acc * 256u;
                const uint64_t da = d0 + (uint64_t)((uint32_t)s * (GT_STAGE_BYTES >> 4));
                const uint64_t db = da + (uint64_t)(GT_BLOCK_BYTES >> 4);
#pragma unroll
                for (int k16 = 0; k16 < GT_BK / 16; ++k16) {
                    mma_bf16_ss_elect(d_tmem, da + 16u * k16, db + 16u * k16, idesc_s, !(slice_start && k16 == 0));
                    mma_bf16_ss_elect(d_tmem + 128u, da + (uint32_t)(GT_PLANE_BYTES >> 4) + 16u * k16, db + 16u * k16, idesc_h, true);
                }
                if (g.cs > 1) mma_commit_multicast_elect(&empty[s], (uint16_t)((1u << g.cs) - 1u));
                else mma_commit_elect(&empty[s]);
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
        // ================= epilogue: 8 warps; TMEM lane == row of the tile; warp (q, hh) owns the rows [32 q, 32 q + 32) and the
        // columns [64 hh, 64 hh + 64) (a warp reads the TMEM lanes of quadrant warp % 4).  Two warps per scheduler: a single
        // warp ran the dependent address / predicate chains of the store phase at ~0.16 instructions per clock.
        const int qd = warp & 3, hh = (warp - 4) >> 2;
        int acc = 0;
        uint32_t acc_ph = 0;
        GtWalk w;
        for (w.start(g); w.next(g);) {
            const GtProblem& P = g.p[w.prob];
            // both cross products sit in the second 128 columns; fp16 lo planes are stored scaled by 2^11 (umma.cuh)
            const float lo_inv = P.f16 ? F16_LO_INV : 1.0f;
            const int n_slices = (w.kc1 - w.kc0 + GT_KD - 1) / GT_KD;
            const int row = w.mt * GT_BM + qd * 32 + lane;
            const int n0 = w.nt * GT_BN + 64 * hh;
            float sum[64];
#pragma unroll
            for (int q = 0; q < 64; ++q) sum[q] = 0.f;
            for (int sl = 0; sl < n_slices; ++sl) {
                mbar_wait(&tmem_full[acc], acc_ph);
                tc_fence_after();
                const uint32_t taddr = tmem + (uint32_t)acc * 256u + ((uint32_t)(qd * 32) << 16) + 64u * (uint32_t)hh;
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    float v[16], v2[16];
                    tmem_ld_x16(taddr + c0, v);
                    tmem_ld_x16(taddr + 128 + c0, v2);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 16; ++q) sum[c0 + q] += fmaf(v2[q], lo_inv, v[q]);
                }
                tc_fence_before();
                mbar_arrive(&tmem_empty[acc]);
                if (++acc == 2) {
                    acc = 0;
                    acc_ph ^= 1;
                }
            }
            // ---- store phase.  During the drain a lane owns one ROW of the tile; storing from there would make every
            // instruction of a warp touch 32 different lines and would chain the bias / mask / old-value loads behind the
            // stores (measured: 25 000 cycles per tile, 90 % of the epilogue warps' time on the short-K forward products).
            // The warp transposes its 32 x 64 block through a swizzled 4 KB shared-memory staging area in two halves of 32
            // columns (16-byte group g of row r sits at g ^ (r & 7): conflict-free in both directions), so that 8 lanes
            // cover 128 contiguous bytes of one output row; the loads of 16 rows are in flight together.  Split-K partial
            // sums take the same path (raw: no alpha / bias / mask / accumulate).
            const bool raw = P.S > 1;
            int my_orow = (w.valid && row < P.M) ? row : -1;
            if (!raw && P.map_Tp && my_orow >= 0) {   // padded-grid row (b, t) -> time-major row t * B + b; padding rows are dropped
                const int bb = row / P.map_Tp, tt = row - bb * P.map_Tp;
                my_orow = tt < P.map_T ? tt * P.map_B + bb : -1;
            }
            float* const obase = raw ? P.part + (size_t)w.sp * P.M * P.N : P.C;
            const int ldo = raw ? P.N : P.ldc;
            const float alpha = raw ? 1.f : P.alpha;
            const float* const bias = raw ? nullptr : P.bias;
            const float* const mask = raw ? nullptr : P.mask;
            const bool acc1 = !raw && P.beta1;
            const bool vec_ok = (ldo & 3) == 0 && ((reinterpret_cast<uintptr_t>(obase) & 15) == 0) &&
                                (!mask || (reinterpret_cast<uintptr_t>(mask) & 15) == 0);
            const bool pair_ok = !vec_ok && (ldo & 1) == 0 && ((reinterpret_cast<uintptr_t>(obase) & 7) == 0) &&
                                 (!mask || (reinterpret_cast<uintptr_t>(mask) & 7) == 0);
            float* const stg = reinterpret_cast<float*>(smem + GT_NS * GT_STAGE_BYTES) + (warp - 4) * 1024;
            const int rsub = lane >> 3, gq = lane & 7;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                if (half == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_float4(sum[4 * j], sum[4 * j + 1], sum[4 * j + 2], sum[4 * j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                            make_float4(sum[32 + 4 * j], sum[32 + 4 * j + 1], sum[32 + 4 * j + 2], sum[32 + 4 * j + 3]);
                }
                __syncwarp();
                const int col = n0 + 32 * half + 4 * gq;
                float bv[4] = {0.f, 0.f, 0.f, 0.f};
                if (bias) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (col + q < P.N) bv[q] = __ldg(bias + col + q);
                }
                const bool full4 = vec_ok && col + 4 <= P.N;
                const bool pair4 = pair_ok && col + 4 <= P.N;
#pragma unroll 1   // rolled on purpose: the unrolled store phase was bound by instruction fetch (stall_no_inst)
                for (int it0 = 0; it0 < 8; it0 += 4) {
                    int orr[4];
                    float4 v[4], mk[4], ov[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = 4 * (it0 + j) + rsub;
                        orr[j] = __shfl_sync(0xffffffffu, my_orow, r);
                        v[j] = *reinterpret_cast<const float4*>(stg + r * 32 + ((gq ^ (r & 7)) << 2));
                        mk[j] = make_float4(1.f, 1.f, 1.f, 1.f);
                        ov[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (full4) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (orr[j] >= 0) {
                                const size_t off = (size_t)orr[j] * ldo + col;
                                if (mask) mk[j] = __ldg(reinterpret_cast<const float4*>(mask + off));
                                if (acc1) ov[j] = *reinterpret_cast<const float4*>(obase + off);
                            }
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (orr[j] >= 0) {
                                float4 o;
                                o.x = (alpha * v[j].x + bv[0]) * mk[j].x + ov[j].x;
                                o.y = (alpha * v[j].y + bv[1]) * mk[j].y + ov[j].y;
                                o.z = (alpha * v[j].z + bv[2]) * mk[j].z + ov[j].z;
                                o.w = (alpha * v[j].w + bv[3]) * mk[j].w + ov[j].w;
                                *reinterpret_cast<float4*>(obase + (size_t)orr[j] * ldo + col) = o;
                            }
                    } else if (pair4) {   // rows that are only 8-byte aligned (ldc = 486, 306: every conv-side output): two float2
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (orr[j] >= 0) {
                                const size_t off = (size_t)orr[j] * ldo + col;
                                if (mask) {
                                    const float2 m0 = __ldg(reinterpret_cast<const float2*>(mask + off)), m1 = __ldg(reinterpret_cast<const float2*>(mask + off) + 1);
                                    mk[j] = make_float4(m0.x, m0.y, m1.x, m1.y);
                                }
                                if (acc1) {
                                    const float2 o0 = *reinterpret_cast<const float2*>(obase + off), o1 = *(reinterpret_cast<const float2*>(obase + off) + 1);
                                    ov[j] = make_float4(o0.x, o0.y, o1.x, o1.y);
                                }
                            }
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (orr[j] >= 0) {
                                float2* d = reinterpret_cast<float2*>(obase + (size_t)orr[j] * ldo + col);
                                d[0] = make_float2((alpha * v[j].x + bv[0]) * mk[j].x + ov[j].x, (alpha * v[j].y + bv[1]) * mk[j].y + ov[j].y);
                                d[1] = make_float2((alpha * v[j].z + bv[2]) * mk[j].z + ov[j].z, (alpha * v[j].w + bv[3]) * mk[j].w + ov[j].w);
                            }
                    } else if (col < P.N) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (orr[j] >= 0) {
                                const size_t off = (size_t)orr[j] * ldo + col;
                                const float vv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
                                float mm[4] = {1.f, 1.f, 1.f, 1.f}, oo[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    if (col + q < P.N) {
                                        if (mask) mm[q] = __ldg(mask + off + q);
                                        if (acc1) oo[q] = obase[off + q];
                                    }
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    if (col + q < P.N) obase[off + q] = (alpha * vv[q] + bv[q]) * mm[q] + oo[q];
                            }
                    }
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (g.cs > 1) cluster_sync_all();   // the peers' commits still arrive on this CTA's barriers until they are done too
    if (warp == 2) tmem_dealloc<512>(tmem);
}

// C = alpha * sum_s part[s] (+ bias) (+ C), splits added in index order (deterministic)
__global__ void k_splitk_reduce(int M, int N, int S, const float* __restrict__ part, float* __restrict__ C, int ldc, float alpha,
                                const float* __restrict__ bias, int beta1) {
    const size_t n = (size_t)M * N;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / N), c = (int)(i - (size_t)r * N);
        float v = 0.f;
        for (int s = 0; s < S; ++s) v += part[(size_t)s * n + i];
        v *= alpha;
        if (bias) v += bias[c];
        float* d = C + (size_t)r * ldc + c;
        *d = beta1 ? *d + v : v;
    }
}

// ---- host ------------------------------------------------------------------------------------------------------
// Operand images live in a per-device arena owned by the library: grow-only, bump-allocated per group of products
// (everything is stream-ordered on the caller's stream, so the next group may overwrite it).  cvb_reserve_workspace
// sizes it up front; growth is a cudaMalloc (illegal while a CUDA graph is being captured: a capturing caller reserves
// or warms up first).  An outgrown chunk is never freed: captured graphs keep replaying into the addresses they saw.
// Images of PARAMETER operands are kept in their own buffers until the parameters change; a buffer that a captured
// graph refers to is pinned to its operand for the life of the process.
struct Arena {
    uint8_t* buf = nullptr;   // the current (largest) chunk
    size_t cap = 0;
};
struct ConstImage {
    const float* p = nullptr;
    int ld = 0, rows = 0, K = 0, kmajor = 0, f16 = 0;
    unsigned long long gen = 0;
    uint16_t* img = nullptr;
    size_t bytes = 0;
    unsigned long long last_use = 0;
    bool pinned = false;
};
static std::mutex g_ws_mu;
static Arena g_arena[64];
static ConstImage g_const[64][16];
static unsigned long long g_weights_gen = 1, g_use_clock = 0;

void weights_changed() {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    ++g_weights_gen;
}
unsigned long long weights_generation() {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    return g_weights_gen;
}

static int arena_reserve(int dev, size_t bytes, bool capturing) {
    Arena& a = g_arena[dev];
    if (a.cap >= bytes) return 0;
    CVB_REQUIRE(!capturing, "the operand-image arena must grow (%zu -> %zu bytes) while a CUDA graph is being captured: run the "
                            "step once before the capture or call cvb_reserve_workspace", a.cap, bytes);
    const size_t want = bytes + bytes / 4 + (1u << 20);
    uint8_t* nb = nullptr;
    CVB_CHECK(cudaMalloc(&nb, want));   // the outgrown chunk stays allocated (captured graphs may still address it)
    a.buf = nb;
    a.cap = want;
    return 0;
}

int reserve_workspace(size_t bytes) {
    int dev = 0;
    CVB_CHECK(cudaGetDevice(&dev));
    CVB_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(g_ws_mu);
    return arena_reserve(dev, bytes, false);
}

// clusters of `cs` CTAs of k_gemm_tc that are co-resident on this device (cached); sets the shared-memory attribute once
static int gemm_tc_clusters(int dev, int cs, int smem, int n_sm, int* out) {
    static int cached[64][5];
    static bool attr_set[64] = {};   // function attributes are per device
    std::lock_guard<std::mutex> lk(g_ws_mu);
    if (!attr_set[dev]) {
        CVB_CHECK(cudaFuncSetAttribute(k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set[dev] = true;
    }
    if (cs == 1) {
        *out = n_sm;
        return 0;
    }
    if (!cached[dev][cs]) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(n_sm / cs * cs);
        cfg.blockDim = dim3(GT_THREADS);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, k_gemm_tc, &cfg) != cudaSuccess) {
            (void)cudaGetLastError();
            ncl = 0;
        }
        if (ncl > n_sm / cs) ncl = n_sm / cs;
        cached[dev][cs] = ncl > 0 ? ncl : -1;
        if (getenv("CVB_DEBUG")) fprintf(stderr, "[cvb] k_gemm_tc: %d co-resident clusters of %d CTAs\n", ncl, cs);
    }
    *out = cached[dev][cs] > 0 ? cached[dev][cs] : 0;
    return 0;
}

bool gemm_tc_eligible(int M, int N, int K) { return M >= 1 && N >= 1 && K >= 16 && (double)M * N * K >= 2.0e5; }

static int widest_vec(const float* p, int ld) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    if ((a & 15) == 0 && (ld & 3) == 0) return 4;
    if ((a & 7) == 0 && (ld & 1) == 0) return 2;
    return 1;
}

static size_t image_bytes(int rows, int K) { return (size_t)ceil_div(rows, GT_BM) * ceil_div(K, GT_BK) * GT_BLOCK_BYTES; }

static bool same_source(const GtOperand& a, const GtOperand& b) {
    return a.p == b.p && a.p2 == b.p2 && a.ld == b.ld && a.ld2 == b.ld2 && a.K == b.K && a.kmajor == b.kmajor && a.conv_ci == b.conv_ci &&
           a.conv_ds == b.conv_ds && a.f16 == b.f16;
}

int gemm_tc_group(cudaStream_t s, const GemmDesc* d, int n) {
    CVB_REQUIRE(n >= 1 && n <= GT_MAXP, "gemm_tc_group: %d products (1..%d)", n, GT_MAXP);
    int dev = 0;
    CVB_CHECK(cudaGetDevice(&dev));
    CVB_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
    GemmTcArgs g;
    memset(&g, 0, sizeof(g));
    g.n_prob = n;
    DeviceInfo di;
    if (int rc = get_device_info(&di)) return rc;
    const int smem = GT_NS * GT_STAGE_BYTES + GT_EPI_BYTES + 256;
    // thread-block clusters (CVB_GEMM_CLUSTER=2|4): the CTAs of a cluster work on neighbouring tiles and multicast the
    // operand blocks they share.  Measured and NOT the default: the ingest of this kernel is capped by the bytes DELIVERED
    // to the SMs (~6.3 KB/clk chip-wide), which multicast does not reduce, and the clusters run in lockstep -- bench step
    // 22.0 ms (no clusters) / 22.4 (pairs) / 22.8 (2 x 2).  tools/bench_mcast.py: multicast only helps a loop that is
    // bound by its own requests in flight (30 -> 42 B/clk per SM with two rounds in flight).
    int cs = 1;
    if (const char* e = getenv("CVB_GEMM_CLUSTER")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) cs = v; }
    int n_cl = 0;
    if (int rc = gemm_tc_clusters(dev, cs, smem, di.n_sm, &n_cl)) return rc;
    if (n_cl < 1) {   // clusters of this size cannot be scheduled here
        cs = 1;
        n_cl = di.n_sm;
    }
    g.cs = cs;
    GtOperand ops[GT_MAXO];       // distinct operands of this group
    bool is_const[GT_MAXO];
    int n_ops = 0, which[GT_MAXP][2];
    int tiles = 0;
    for (int i = 0; i < n; ++i) {
        const GemmDesc& D = d[i];
        CVB_REQUIRE(D.M > 0 && D.N > 0 && D.K > 0 && D.A && D.B && D.C, "gemm_tc: empty product (M=%d N=%d K=%d)", D.M, D.N, D.K);
        CVB_REQUIRE(!D.B2 || !D.transB, "gemm_tc: a second B source needs B stored [K,N]");
        CVB_REQUIRE(!D.A2 || (D.transA && !D.gA), "gemm_tc: a second A source needs A stored [K,M]");
        CVB_REQUIRE(!D.gA || !D.transA, "gemm_tc: a conv-gathered A is addressed as [M,K]");
        CVB_REQUIRE(!D.gB || (!D.transB && !D.B2), "gemm_tc: a conv-gathered B is addressed as [K,N]");
        for (int side = 0; side < 2; ++side) {
            GtOperand o;
            memset(&o, 0, sizeof(o));
            if (side == 0) {
                o.p = D.A; o.p2 = D.A2; o.ld = D.lda; o.ld2 = D.lda2; o.rows = D.M; o.R1 = D.A2 ? D.M1 : D.M;
                o.kmajor = D.transA ? 0 : 1;      // A stored [M,K] unless transA ([K,M])
                o.conv_ci = D.gA ? D.gA->ci : 0; o.conv_ds = D.gA ? D.gA->dshift : 0;
            } else {
                o.p = D.B; o.p2 = D.B2; o.ld = D.ldb; o.ld2 = D.ldb2; o.rows = D.N; o.R1 = D.B2 ? D.N1 : D.N;
                o.kmajor = D.transB ? 1 : 0;      // B as [N rows, K]: stored [N,K] when transB, [K,N] otherwise
                o.conv_ci = D.gB ? D.gB->ci : 0; o.conv_ds = D.gB ? D.gB->dshift : 0;
            }
            o.K = D.K;
            o.KC = ceil_div(D.K, GT_BK);
            o.vec = widest_vec(o.p, o.ld);
            o.f16 = D.f16 ? 1 : 0;
            const bool cst = (side == 0 ? D.a_const : D.b_const) && !o.p2 && !o.conv_ci;
            // the same source with at least as many rows: share the image (a prefix of whole row tiles)
            int found = -1;
            for (int j = 0; j < n_ops && found < 0; ++j)
                if (same_source(ops[j], o) && (ops[j].rows == o.rows || (ops[j].rows > o.rows && o.rows % GT_BM == 0 && !o.p2))) found = j;
            if (found < 0) {
                for (int j = 0; j < n_ops && found < 0; ++j)   // ... or grow an earlier, shorter one
                    if (same_source(ops[j], o) && o.rows > ops[j].rows && ops[j].rows % GT_BM == 0 && !o.p2) {
                        ops[j].rows = o.rows;
                        ops[j].R1 = o.R1;
                        found = j;
                    }
            }
            if (found < 0) {
                found = n_ops;
                is_const[n_ops] = cst;
                ops[n_ops++] = o;
            }
            which[i][side] = found;
        }
        GtProblem& P = g.p[i];
        P.C = D.C;
        P.bias = D.bias;
        P.alpha = D.alpha;
        P.M = D.M;
        P.N = D.N;
        P.ldc = D.ldc;
        P.MT = ceil_div(D.M, GT_BM);
        P.NT = ceil_div(D.N, GT_BN);
        P.KC = ceil_div(D.K, GT_BK);
        P.beta1 = D.beta1 ? 1 : 0;
        P.f16 = D.f16 ? 1 : 0;
        P.map_Tp = D.map_Tp;
        P.map_T = D.map_T;
        P.map_B = D.map_B;
        P.mask = D.mask;
        // cluster shape of this product: fewest padded tiles x requested bytes per CTA (each CTA pulls 1/cn of its A
        // block and 1/cm of its B block; below ~0.52 of the full blocks the tensor pipe binds, not the ingest)
        P.cm = P.cn = 1;
        {
            double best = 1e30;
            for (int cm = 1; cm <= cs; cm *= 2) {
                const int cn = cs / cm;
                const double padded = (double)ceil_div(P.MT, cm) * cm * ceil_div(P.NT, cn) * cn;
                double req = 0.5 / cm + 0.5 / cn;
                if (req < 0.52) req = 0.52;
                if (padded * req < best) {
                    best = padded * req;
                    P.cm = cm;
                    P.cn = cn;
                }
            }
        }
        P.MTc = ceil_div(P.MT, P.cm);
        P.NTc = ceil_div(P.NT, P.cn);
        // split-K: a product with few output tiles and a deep K would keep a handful of clusters busy for its whole length
        P.S = 1;
        P.kc_split = P.KC;
        if (!D.map_Tp && P.MT * P.NT <= 48 && P.KC >= 16) {
            int want = n_cl / (P.MTc * P.NTc);
            if (want > P.KC / 4) want = P.KC / 4;
            if (want > 1) {
                P.kc_split = ceil_div(P.KC, want);
                P.S = ceil_div(P.KC, P.kc_split);
            }
        }
        tiles += P.MTc * P.NTc * P.S;
        P.tile_end = tiles;
    }
    // place the images: cached parameter images in their own buffers, the rest bump-allocated in the arena
    SplitArgs sa;
    memset(&sa, 0, sizeof(sa));
    cudaStreamCaptureStatus cap_st = cudaStreamCaptureStatusNone;
    CVB_CHECK(cudaStreamIsCapturing(s, &cap_st));
    const bool capturing = cap_st != cudaStreamCaptureStatusNone;
    {
        std::lock_guard<std::mutex> lk(g_ws_mu);
        // a parameter image needs a slot that no captured graph owns; without one the operand is split like any other
        for (int j = 0; j < n_ops; ++j) {
            if (!is_const[j]) continue;
            const GtOperand& o = ops[j];
            bool ok = false;
            for (auto& c : g_const[dev])
                if ((c.p == o.p && c.ld == o.ld && c.rows == o.rows && c.K == o.K && c.kmajor == o.kmajor && c.f16 == o.f16) ||
                    (!c.pinned && !(capturing && c.bytes < image_bytes(o.rows, o.K))))
                    ok = true;
            if (!ok) is_const[j] = false;
        }
        size_t need = 0;
        for (int j = 0; j < n_ops; ++j)
            if (!is_const[j]) need += image_bytes(ops[j].rows, ops[j].K);
        for (int i = 0; i < n; ++i)
            if (g.p[i].S > 1) need += round_up_sz((size_t)g.p[i].S * g.p[i].M * g.p[i].N * sizeof(float), 256);
        if (int rc = arena_reserve(dev, need, capturing)) return rc;
        size_t off = 0;
        for (int i = 0; i < n; ++i)
            if (g.p[i].S > 1) {
                g.p[i].part = reinterpret_cast<float*>(g_arena[dev].buf + off);
                off += round_up_sz((size_t)g.p[i].S * g.p[i].M * g.p[i].N * sizeof(float), 256);
            }
        int blocks = 0;
        for (int j = 0; j < n_ops; ++j) {
            GtOperand& o = ops[j];
            const size_t bytes = image_bytes(o.rows, o.K);
            bool fresh = true;
            if (is_const[j]) {
                ConstImage* slot = nullptr;
                for (auto& c : g_const[dev])
                    if (c.p == o.p && c.ld == o.ld && c.rows == o.rows && c.K == o.K && c.kmajor == o.kmajor && c.f16 == o.f16) slot = &c;
                if (!slot) {   // least recently used slot that no captured graph refers to
                    for (auto& c : g_const[dev])
                        if (!c.pinned && !(capturing && c.bytes < bytes) && (!slot || c.last_use < slot->last_use)) slot = &c;
                    CVB_REQUIRE(slot, "internal: no parameter-image slot");
                    if (slot->bytes < bytes) {
                        if (slot->img) CVB_CHECK(cudaFree(slot->img));
                        slot->img = nullptr;
                        slot->bytes = 0;
                        CVB_CHECK(cudaMalloc(&slot->img, bytes));
                        slot->bytes = bytes;
                    }
                    slot->p = o.p; slot->ld = o.ld; slot->rows = o.rows; slot->K = o.K; slot->kmajor = o.kmajor; slot->f16 = o.f16;
                    slot->gen = 0;
                }
                slot->last_use = ++g_use_clock;
                if (capturing) slot->pinned = true;
                o.img = slot->img;
                fresh = slot->gen != g_weights_gen;
                slot->gen = g_weights_gen;
            } else {
                o.img = reinterpret_cast<uint16_t*>(g_arena[dev].buf + off);
                off += bytes;
            }
            if (fresh) {
                GtOperand& t = sa.o[sa.n++];
                t = o;
                blocks += ceil_div(o.rows, GT_BM) * o.KC;
                t.block_end = blocks;
            }
        }
        if (sa.n) {
            k_split_group<<<blocks, 256, 0, s>>>(sa);
            CVB_LAUNCH_CHECK();
        }
    }
    for (int i = 0; i < n; ++i) {
        g.p[i].At = ops[which[i][0]].img;
        g.p[i].Bt = ops[which[i][1]].img;
    }
    int grid = (tiles < n_cl ? tiles : n_cl) * cs;
    if (const char* e = getenv("CVB_GEMM_MAX_CTAS")) { const int v = atoi(e) / cs * cs; if (v > 0 && v < grid) grid = v; }   // tools/overlap_probe.py
    prof_begin(s, CVB_PROF_GEMM);
    if (cs == 1) {
        k_gemm_tc<<<grid, GT_THREADS, smem, s>>>(g);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(GT_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        CVB_CHECK(cudaLaunchKernelEx(&cfg, k_gemm_tc, g));
    }
    CVB_LAUNCH_CHECK();
    for (int i = 0; i < n; ++i)
        if (g.p[i].S > 1) {
            const GtProblem& P = g.p[i];
            const size_t mn = (size_t)P.M * P.N;
            k_splitk_reduce<<<(int)(ceil_div_sz(mn, 256) > 1184 ? 1184 : ceil_div_sz(mn, 256)), 256, 0, s>>>(P.M, P.N, P.S, P.part, P.C, P.ldc, P.alpha,
                                                                                                         P.bias, P.beta1);
            CVB_LAUNCH_CHECK();
        }
    prof_end(s, CVB_PROF_GEMM);
    return 0;
}

// C[M,N] = op(A) op(B) (+ C if beta1) (+ bias[N]);  A: [M,K] (lda) or [K,M] if transA;  B: [N,K] (ldb) if transB else [K,N].
int gemm_tc(cudaStream_t s, bool transA, bool transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
            bool beta1, const float* bias, float* C, int ldc, bool f16, const float* B2, int ldb2, int N1, const ConvGather* gA,
            const ConvGather* gB) {
    GemmDesc d;
    d.transA = transA;
    d.transB = transB;
    d.M = M;
    d.N = N;
    d.K = K;
    d.A = A;
    d.lda = lda;
    d.B = B;
    d.ldb = ldb;
    d.beta1 = beta1;
    d.bias = bias;
    d.C = C;
    d.ldc = ldc;
    d.f16 = f16;
    d.B2 = B2;
    d.ldb2 = ldb2;
    d.N1 = N1;
    d.gA = gA;
    d.gB = gB;
    return gemm_tc_group(s, &d, 1);
}

}  // namespace cvb
