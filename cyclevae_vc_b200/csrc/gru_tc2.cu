// Tensor-core (tcgen05) forward recurrence of the autoregressive GRU (gru_vae.py:364-399), ONE grid-wide exchange per
// recurrent step (gru_tc.cu, the first version, needs two: h_t, then the reduced y_t).
//
// Same 2-D split as gru_tc.cu: clusters of S = 4 CTAs, cluster i owns the hidden units [32i, 32i+32), CTA j owns the
// K-slice [j H/4, (j+1) H/4) of  gh = W_hh h_{t-1}  for all units of the block and FINALISES the units [32i + 8j, +8);
// W_hh rows of (block x K-slice) resident in shared memory as fp16 hi+lo, partial accumulators meet in the finaliser's
// shared memory through bulk DSMEM copies.
//
// What changed is the feedback  y_t = W_o o_t + b_o  ->  W_y y_t  (a reduction over ALL hidden units every step):
//   * producer side: the four CTAs of a cluster swap their o_t (8 units each) through DSMEM, so every CTA holds o_t of the
//     cluster's 32 units as one K = 32 operand; CTA j forms the cluster's partial of the output QUARTER [16j, 16j+16)
//     with one MMA chain (accumulator D3) -- 32 partials per pair instead of 128 -- and ADDS it to a 64-bit FIXED-POINT
//     accumulator in L2 (red.global.add.u64 of round(x 2^36)): integer adds are exact and associative, so the total does
//     not depend on the arrival order (run-to-run deterministic) and no reducer CTA, no second counter hop, no
//     all-to-all traffic is needed (summing the 32 fp32 partials in every consumer was tried first: 20 MB of L2 reads
//     per step, 9 700 cycles);
//   * the accumulators are never cleared: they hold running totals (two slots, by step parity) and every consumer
//     subtracts the total it saw two steps earlier (kept in registers) -- exact in wrap-around integer arithmetic;
//   * two arrival counters, both one hop: H (h_t published, released right after the gates) and Y (the partials of
//     y_t added); the W_hh h chain of the next step starts on H while the partial of y_t is still being formed;
//   * consumer side: once counter Y is complete CTA j reads quarter j of the totals (10 KB), converts, adds b_o,
//     splits to fp16 hi/lo and writes its two k blocks of the y operand; the four quarters are swapped through DSMEM
//     (one bulk copy per peer), and W_y y_t of the own units is the same 64-deep MMA chain as before (accumulator D2).
// Operand buffers that are assembled from four CTAs are K-OUTER: [k block][plane hi|lo][row group][8 rows][8 k], so
// a CTA's contribution is one contiguous piece (descriptor: LBO = k-block stride, SBO = 128).
//
// Roles (384 threads): w0 bulk-copy producer of the h chunks, w1 MMA issuer, w2 TMEM allocator, w4-7 exchange + gates
// (TMEM lane == batch row), w4-11 read the totals of y (one 16-byte operand row per thread), w8 releases counter H.
#include <stdlib.h>

#include "gru_ar.cuh"
#include "umma.cuh"

namespace cvb {
using namespace umma;

constexpr int T2_NT = 384;
constexpr int T2_KC = 64;             // K per ring stage
constexpr int T2_S = 4;               // cluster size
constexpr int T2_UB = 8 * T2_S;       // units per cluster
constexpr int T2_NW = 3 * T2_UB;      // rows of the W_hh operand (r, z, n of the block) = 96
constexpr int T2_OQ = 16;             // outputs per quarter (the output axis is padded to 64)
constexpr float T2_FX = 68719476736.0f;       // 2^36: fixed-point scale of the y accumulators (resolution 1.5e-11, |partial| < 2^27)
constexpr float T2_FX_INV = 1.0f / 68719476736.0f;
// TMEM columns (D1 = [main0 | corrections | main1] as in gru_tc.cu: no accumulation chain longer than K = 128)
constexpr uint32_t T2_COL_M0 = 0;
constexpr uint32_t T2_COL_C = T2_NW;
constexpr uint32_t T2_COL_M1 = 2 * T2_NW;
constexpr uint32_t T2_COL_Y = 288;    // D2: W_y y for the own units, 32 main + 32 correction columns
constexpr uint32_t T2_COL_P = 352;    // D3: cluster partial of the own output quarter, 16 main + 16 correction columns
constexpr uint32_t T2_COL_DUMMY = 480;

struct T2Layout {
    int MB, nch, NS, NCL;
    uint32_t half, stage_bytes, w_chunk_bytes, slot_bytes;
    uint32_t kstr, pstr;   // K-outer operands: bytes between k blocks / between the hi and lo plane of a k block
    uint32_t off_ring, off_ybuf, off_w, off_b2, off_b3, off_a2, off_inbox, off_bias, off_bar, total;
};

__host__ __device__ inline T2Layout t2_layout(int B, int H, int smem_max) {
    T2Layout L;
    L.MB = (B + 7) / 8;
    L.nch = H / T2_KC / T2_S;
    L.NCL = H / T2_UB;
    L.half = (uint32_t)L.MB * 1024u;
    L.stage_bytes = 2u * L.half;
    L.w_chunk_bytes = 2u * (T2_NW / 8) * 1024u;                 // [hi: 12 row groups][lo: 12 row groups] x 1 KB
    L.slot_bytes = (uint32_t)L.MB * 8u * 96u;                   // [rows][24 floats]
    L.pstr = (uint32_t)L.MB * 128u;
    L.kstr = 2u * L.pstr;
    uint32_t inbox = (uint32_t)T2_S * L.slot_bytes;
    inbox = (inbox + 127u) & ~127u;
    const uint32_t fixed = L.stage_bytes + (uint32_t)L.nch * L.w_chunk_bytes + 8192u + 2048u + L.half + inbox + 256u + 256u;
    int ns = ((int)smem_max - (int)fixed) / (int)L.stage_bytes;
    L.NS = ns > 6 ? 6 : ns;
    const uint32_t ring = (uint32_t)(L.NS > 0 ? L.NS : 0) * L.stage_bytes;
    L.off_ring = 0;                       // h chunks; idle between a step's last chunk and the next step's first, when it doubles as
                                          // the staging of the outgoing partial sums
    L.off_ybuf = ring;                    // y_{t-1} operand, K-outer: [8 k blocks][hi | lo][MB][128 B]
    L.off_w = L.off_ybuf + L.stage_bytes;
    L.off_b2 = L.off_w + (uint32_t)L.nch * L.w_chunk_bytes;   // W_y rows of the own units: [hi 4 groups][lo 4 groups] x 1 KB
    L.off_b3 = L.off_b2 + 8192u;          // W_o[own quarter][cluster's units]: [hi 2 groups | lo 2 groups] x 512 B (4 k blocks)
    L.off_a2 = L.off_b3 + 2048u;          // o_t of the cluster's units, K-outer: [4 k blocks][hi | lo][MB][128 B]
    L.off_inbox = L.off_a2 + L.half;
    L.off_bias = L.off_inbox + inbox;     // [24] b_hh of the own units | [16] b_o of the own quarter
    L.off_bar = L.off_bias + 256u;
    L.total = L.off_bar + 256u;
    return L;
}

struct GruTc2Args {
    GruFwdArgs f;
    uint16_t* hx;        // [2 slots][2 parts][H/64 chunks][MB][8 kblk][8 rows][8 k] fp16 (UMMA order) of h_t
    unsigned long long* yacc;   // [2 slots][64 outputs][MB*8 rows] running fixed-point totals of y_t, zero-initialised
    unsigned* ctr;       // CTR_BANKS lines of counter H, then CTR_BANKS lines of counter Y (common.cuh), zero-initialised
    int smem_max;
    int keepalive;
    int relaxed;
    int dbg;            // CVB_TC_DBG bits (experiments): 1 = skip the fixed-point adds (wrong results), 2 = no L2 prefetch of the next frames
    long long* trace;    // optional [T+1][64] clock64 stamps of CTA 0 (CVB_TRACE_FILE_FWD), else null
};

#define T2_TRACE(ev)                                                     \
    do {                                                                 \
        if (a.trace && c == 0) a.trace[(size_t)t * 64 + (ev)] = clock64(); \
    } while (0)

static __device__ __forceinline__ void t2_split8(const float* x, uint4& hi, uint4& lo) {
    uint16_t h[8], l[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) split_f16(x[q], h[q], l[q]);
    hi = make_uint4((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16),
                    (uint32_t)h[4] | ((uint32_t)h[5] << 16), (uint32_t)h[6] | ((uint32_t)h[7] << 16));
    lo = make_uint4((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16),
                    (uint32_t)l[4] | ((uint32_t)l[5] << 16), (uint32_t)l[6] | ((uint32_t)l[7] << 16));
}
// 8-byte load of a running total, pinned in program order (all eight of a thread in flight).  A plain (weak) load is
// enough: the polling thread's ld.acquire.gpu (which also invalidates this SM's L1) and the CTA barrier behind it order it
// after the writers' release.
static __device__ __forceinline__ unsigned long long ld_total_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
static __device__ __forceinline__ void red_add_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(T2_NT, 1) k_gru_fwd_tc2(GruTc2Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const GruFwdArgs& f = a.f;
    const int B = f.B, T = f.T, H = f.H, out = f.out;
    const int G = gridDim.x, c = blockIdx.x;
    const int j = (int)cluster_ctarank();
    const T2Layout L = t2_layout(B, H, a.smem_max);
    const int ci = c / T2_S;                // cluster index
    const int ublk0 = ci * T2_UB;           // first unit of the cluster's block
    const int u0 = ublk0 + 8 * j;           // first of the 8 units this CTA finalises
    const int k0 = j * L.nch * T2_KC;       // first column of W_hh (= unit of h) of this CTA's K-slice
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    uint8_t* ring = smem + L.off_ring;
    float* stage = reinterpret_cast<float*>(ring);                 // [S (to)][MB*8][24], aliases the (idle) ring
    uint8_t* ybuf = smem + L.off_ybuf;
    uint8_t* sW = smem + L.off_w;
    uint8_t* sB2 = smem + L.off_b2;
    uint8_t* sB3 = smem + L.off_b3;
    uint8_t* sA2 = smem + L.off_a2;
    float* inbox = reinterpret_cast<float*>(smem + L.off_inbox);   // [S (from)][MB*8][24]
    float* sBh = reinterpret_cast<float*>(smem + L.off_bias);      // [3][8] b_hh of the own units
    float* sBo = sBh + 32;                                         // [16] b_o of the own output quarter (zero beyond out)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* empty = full + 8;
    uint64_t* accum_full = full + 16;   // D2 (W_y y) complete
    uint64_t* d1_full = full + 17;      // D1 (K-slice of W_hh h) complete
    uint64_t* y_full = full + 18;       // the four quarters of the y operand are in place
    uint64_t* inbox_full = full + 19;
    uint64_t* a2_full = full + 20;      // o_t of the cluster's units is in place
    uint64_t* part_full = full + 21;    // D3 complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 22);
    const size_t hx_part = (size_t)(H / T2_KC) * L.MB * 512;   // elements per part
    unsigned* ctrH = a.ctr;
    unsigned* ctrY = a.ctr + 32 * CTR_BANKS;
    const unsigned per_bank = (unsigned)(G / CTR_BANKS);
    const int n_pairs = B * out;
    const int half1 = (L.nch + 1) / 2;   // first chunk of the K-slice that accumulates into main1
    const bool two_main = half1 < L.nch;
    const size_t RP = (size_t)L.MB * 8;                         // padded rows of the y accumulators
    const size_t yslot = (size_t)64 * RP;                       // elements of one slot

    if (a.trace && c == 0 && threadIdx.x == 0) a.trace[60] = clock64();
    // ---- one-time setup: weights -> fp16 hi/lo in UMMA K-major core-matrix order -----------------
    {
        const int n_items = T2_NW * L.nch * (T2_KC / 8);
        for (int i0 = threadIdx.x; i0 < n_items; i0 += 8 * T2_NT) {
            float4 wv[16];
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int i = i0 + m * T2_NT;
                if (i < n_items) {
                    const int kg = i / T2_NW, n = i - kg * T2_NW;   // n = g*32 + unit of the block
                    const int g = n / T2_UB, ul = n - g * T2_UB;
                    const float4* src = reinterpret_cast<const float4*>(f.Whh + (size_t)(g * H + ublk0 + ul) * H + k0 + kg * 8);
                    wv[2 * m] = __ldg(src);
                    wv[2 * m + 1] = __ldg(src + 1);
                }
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int i = i0 + m * T2_NT;
                if (i < n_items) {
                    const int kg = i / T2_NW, n = i - kg * T2_NW;
                    const int kl = kg * 8;
                    const float w[8] = {wv[2 * m].x, wv[2 * m].y, wv[2 * m].z, wv[2 * m].w, wv[2 * m + 1].x, wv[2 * m + 1].y, wv[2 * m + 1].z, wv[2 * m + 1].w};
                    uint4 hi, lo;
                    t2_split8(w, hi, lo);
                    const uint32_t off = (uint32_t)(kl / T2_KC) * L.w_chunk_bytes + (uint32_t)(n >> 3) * 1024u + (uint32_t)((kl % T2_KC) >> 3) * 128u +
                                         (uint32_t)(n & 7) * 16u;
                    // chunks of the second half of the K walk are stored [lo rows | hi rows] (gru_tc.cu)
                    const bool swapped = (kl / T2_KC) >= half1;
                    *reinterpret_cast<uint4*>(sW + off + (swapped ? (T2_NW / 8) * 1024u : 0u)) = hi;
                    *reinterpret_cast<uint4*>(sW + off + (swapped ? 0u : (T2_NW / 8) * 1024u)) = lo;
                }
            }
        }
        {   // B2[n = g*8+uu][k] = W_y[g*H + u0 + uu][k]; rows 24..31 zero.
            // B3[n = o_local][k = unit of the block] = W_o[16j + o_local][ublk0 + k]: rows [hi 16 | lo 16], 4 k blocks
            float w2[6], w3[2];
#pragma unroll
            for (int m = 0; m < 6; ++m) {
                const int i = threadIdx.x + m * T2_NT, n = i >> 6, k = i & 63;
                w2[m] = (i < 32 * 64 && n < 24 && k < out) ? __ldg(f.Wy + (size_t)((n >> 3) * H + u0 + (n & 7)) * f.ldwy + k) : 0.f;
            }
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const int i = threadIdx.x + m * T2_NT, n = i >> 5, k = i & 31;
                const int o = T2_OQ * j + n;
                w3[m] = (i < T2_OQ * 32 && o < out) ? __ldg(f.Wo + (size_t)o * H + ublk0 + k) : 0.f;
            }
#pragma unroll
            for (int m = 0; m < 6; ++m) {
                const int i = threadIdx.x + m * T2_NT, n = i >> 6, k = i & 63;
                if (i < 32 * 64) {
                    uint16_t hi, lo;
                    split_f16(w2[m], hi, lo);
                    const uint32_t off = (uint32_t)(n >> 3) * 1024u + (uint32_t)(k >> 3) * 128u + (uint32_t)(n & 7) * 16u + (uint32_t)(k & 7) * 2u;
                    *reinterpret_cast<uint16_t*>(sB2 + off) = hi;
                    *reinterpret_cast<uint16_t*>(sB2 + 4096 + off) = lo;
                }
            }
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const int i = threadIdx.x + m * T2_NT, n = i >> 5, k = i & 31;
                if (i < T2_OQ * 32) {
                    uint16_t hi, lo;
                    split_f16(w3[m], hi, lo);
                    const uint32_t off = (uint32_t)(n >> 3) * 512u + (uint32_t)(k >> 3) * 128u + (uint32_t)(n & 7) * 16u + (uint32_t)(k & 7) * 2u;
                    *reinterpret_cast<uint16_t*>(sB3 + off) = hi;
                    *reinterpret_cast<uint16_t*>(sB3 + 1024 + off) = lo;
                }
            }
        }
        for (uint32_t i = threadIdx.x; i < L.half / 16; i += T2_NT) reinterpret_cast<uint4*>(sA2)[i] = make_uint4(0u, 0u, 0u, 0u);
        for (uint32_t i = threadIdx.x; i < L.stage_bytes / 16; i += T2_NT) reinterpret_cast<uint4*>(ybuf)[i] = make_uint4(0u, 0u, 0u, 0u);
        if (threadIdx.x < 24) sBh[threadIdx.x] = f.bhh[(threadIdx.x >> 3) * H + u0 + (threadIdx.x & 7)];
        if (threadIdx.x >= 32 && threadIdx.x < 32 + T2_OQ) {
            const int o = T2_OQ * j + (threadIdx.x - 32);
            sBo[threadIdx.x - 32] = o < out ? f.bo[o] : 0.f;
        }
        if (threadIdx.x == 0) {
            for (int s = 0; s < 8; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            mbar_init(accum_full, 1);
            mbar_init(d1_full, 1);
            mbar_init(y_full, 1);       // one arrive.expect_tx by the owner per step; the three peers' pieces complete_tx
            mbar_init(inbox_full, 1);
            mbar_init(a2_full, 1);      // same protocol as y_full
            mbar_init(part_full, 1);
            mbar_fence_init();
        }
        fence_proxy_async_smem();
        if (warp == 2) tmem_alloc<512>(tmem_slot);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    cluster_sync_all();   // every CTA's barriers are initialised before any peer copies into it
    const uint32_t tmem = *tmem_slot;
    if (a.trace && c == 0 && threadIdx.x == 0) a.trace[61] = clock64();

    if (warp == 0) {
        // ================= producer: K-slice of h_{t-1} chunk by chunk ====================================
        int s = 0;
        uint32_t ph = 1;
        for (int t = 0; t < T; ++t) {
            const uint16_t* src = a.hx + (size_t)(t & 1) * 2 * hx_part + (size_t)(j * L.nch) * L.MB * 512;
            banked_wait_warp(ctrH, per_bank * (unsigned)(t + 1), lane, a.relaxed != 0);   // the writers fenced generic -> async proxy before their release
            if (lane == 0) T2_TRACE(14);
            for (int ch = 0; ch < L.nch; ++ch) {
                if (lane == 0) {
                    mbar_wait(&empty[s], ph);
                    if (ch < 8) T2_TRACE(32 + ch);
                    uint8_t* dst = ring + (size_t)s * L.stage_bytes;
                    mbar_expect_tx(&full[s], 2 * L.half);
                    bulk_g2s(dst, src + (size_t)ch * L.MB * 512, L.half, &full[s]);
                    bulk_g2s(dst + L.half, src + hx_part + (size_t)ch * L.MB * 512, L.half, &full[s]);
                }
                __syncwarp();
                if (++s == L.NS) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (descriptors stay warp-uniform; one elected lane issues) ===========
        const uint32_t idesc1s = idesc_f16_f32(128, 2 * T2_NW), idesc1 = idesc_f16_f32(128, T2_NW);
        const uint32_t idesc2s = idesc_f16_f32(128, 64), idesc2 = idesc_f16_f32(128, 32);
        const uint32_t idesc3s = idesc_f16_f32(128, 2 * T2_OQ), idesc3 = idesc_f16_f32(128, T2_OQ);
        const uint32_t idesc_dummy = idesc_f16_f32(128, 16);
        const uint64_t dA0 = smem_desc(smem_u32(ring), 128, 1024);
        const uint64_t dW0 = smem_desc(smem_u32(sW), 128, 1024);
        const uint64_t dY0 = smem_desc(smem_u32(ybuf), L.kstr, 128);    // K-outer
        const uint64_t dB2 = smem_desc(smem_u32(sB2), 128, 1024);
        const uint64_t dA2 = smem_desc(smem_u32(sA2), L.kstr, 128);     // K-outer
        const uint64_t dB3 = smem_desc(smem_u32(sB3), 128, 512);
        const uint32_t a_step = L.stage_bytes >> 4, half16 = L.half >> 4, w_step = L.w_chunk_bytes >> 4;
        const uint32_t k16o = (2u * L.kstr) >> 4, plo = L.pstr >> 4;    // K-outer operands: one K = 16 step / hi -> lo plane
        bool y_done = false;
        // W_y y_{t-1} of the own units as soon as the y operand is complete (between two h chunks if it lands early)
        auto try_y = [&](int t, bool block) -> bool {
            for (;;) {
                uint32_t ok = (lane == 0) ? (mbar_test_wait(y_full, (uint32_t)t & 1) ? 1u : 0u) : 0u;
                ok = __shfl_sync(0xffffffffu, ok, 0);
                if (ok) break;
                if (!block) return false;
                if (a.keepalive) mma_bf16_ss_elect(tmem + T2_COL_DUMMY, dW0, dB2, idesc_dummy, false);
            }
            if (lane == 0) T2_TRACE(12);
            fence_proxy_async_smem();   // the peers' rows arrived as st.async stores (counted on the barrier): order them before the MMAs' reads
            tc_fence_after();
#pragma unroll
            for (int k16 = 0; k16 < T2_KC / 16; ++k16) {
                mma_bf16_ss_elect(tmem + T2_COL_Y, dY0 + k16o * k16, dB2 + 16u * k16, idesc2s, k16 != 0);
                mma_bf16_ss_elect(tmem + T2_COL_Y + 32u, dY0 + plo + k16o * k16, dB2 + 16u * k16, idesc2, true);
            }
            mma_commit_elect(accum_full);
            return true;
        };
        int s = 0;
        uint32_t ph = 0;
        for (int t = 0; t < T; ++t) {
            y_done = false;
            for (int ch = 0; ch < L.nch; ++ch) {
                for (;;) {
                    uint32_t ok = (lane == 0) ? (mbar_test_wait(&full[s], ph) ? 1u : 0u) : 0u;
                    ok = __shfl_sync(0xffffffffu, ok, 0);
                    if (ok) break;
                    if (!y_done) y_done = try_y(t, false);
                    if (a.keepalive && !(a.dbg & 32))
                        mma_bf16_ss_elect(tmem + T2_COL_DUMMY, dW0, dB2, idesc_dummy, false);
                }
                if (lane == 0 && ch < 8) T2_TRACE(40 + ch);
                tc_fence_after();
                const uint64_t da = dA0 + (uint64_t)((uint32_t)s * a_step);
                const uint64_t db = dW0 + (uint64_t)((uint32_t)ch * w_step);
                const uint32_t w_half16 = (T2_NW / 8) * 1024u >> 4;   // hi rows -> lo rows (or lo -> hi in a swapped chunk)
                if (ch < half1) {   // [hi | lo] rows: main0 and corrections side by side
#pragma unroll
                    for (int k16 = 0; k16 < T2_KC / 16; ++k16) {
                        mma_bf16_ss_elect(tmem + T2_COL_M0, da + 16u * k16, db + 16u * k16, idesc1s, (ch | k16) != 0);
                        mma_bf16_ss_elect(tmem + T2_COL_C, da + half16 + 16u * k16, db + 16u * k16, idesc1, true);
                    }
                } else {            // [lo | hi] rows: corrections and main1 side by side
#pragma unroll
                    for (int k16 = 0; k16 < T2_KC / 16; ++k16) {
                        if (ch == half1 && k16 == 0) {   // main1 starts from zero while the corrections keep accumulating
                            mma_bf16_ss_elect(tmem + T2_COL_C, da, db, idesc1, true);
                            mma_bf16_ss_elect(tmem + T2_COL_M1, da, db + w_half16, idesc1, false);
                        } else {
                            mma_bf16_ss_elect(tmem + T2_COL_C, da + 16u * k16, db + 16u * k16, idesc1s, true);
                        }
                        mma_bf16_ss_elect(tmem + T2_COL_C, da + half16 + 16u * k16, db + w_half16 + 16u * k16, idesc1, true);
                    }
                }
                mma_commit_elect(&empty[s]);
                if (ch == L.nch - 1) mma_commit_elect(d1_full);
                if (lane == 0 && ch < 8) T2_TRACE(48 + ch);
                if (++s == L.NS) {
                    s = 0;
                    ph ^= 1;
                }
            }
            if (!y_done) try_y(t, true);
            // cluster partial of the own output quarter: D3[b][o] = sum_{k < 32} o_t[b][ublk0 + k] W_o[16j + o][ublk0 + k]
            for (;;) {
                uint32_t ok = (lane == 0) ? (mbar_test_wait(a2_full, (uint32_t)t & 1) ? 1u : 0u) : 0u;
                ok = __shfl_sync(0xffffffffu, ok, 0);
                if (ok) break;
                if (a.keepalive) mma_bf16_ss_elect(tmem + T2_COL_DUMMY, dW0, dB2, idesc_dummy, false);
            }
            if (lane == 0) T2_TRACE(13);
            fence_proxy_async_smem();
            tc_fence_after();
#pragma unroll
            for (int k16 = 0; k16 < 2; ++k16) {
                mma_bf16_ss_elect(tmem + T2_COL_P, dA2 + k16o * k16, dB3 + 16u * k16, idesc3s, k16 != 0);
                mma_bf16_ss_elect(tmem + T2_COL_P + T2_OQ, dA2 + plo + k16o * k16, dB3 + 16u * k16, idesc3, true);
            }
            mma_commit_elect(part_full);
        }
    } else if (warp >= 4 && warp < 8) {
        // ================= exchange + gates: TMEM lane == batch row ======================================
        const int b = (warp - 4) * 32 + lane;
        const bool act = b < B;
        const int etid = threadIdx.x - 128;
        const uint32_t inbox_addr = smem_u32(inbox);
        const uint32_t inbox_bar_addr = smem_u32(inbox_full);
        const uint32_t taddr = tmem + ((uint32_t)((warp - 4) * 32) << 16);
        const uint32_t slot_f = L.slot_bytes / 4;
        const bool wact = (warp - 4) * 32 < L.MB * 8;   // this warp holds rows of the staged row groups
        float hreg[8];
        {   // prologue: publish h_in (slot 0) in operand order; y_in needs no partials
#pragma unroll
            for (int q = 0; q < 8; ++q) hreg[q] = act ? f.hs[(size_t)b * H + u0 + q] : 0.f;
            uint4 hh, hl;
            t2_split8(hreg, hh, hl);
            if (act) {
                const size_t off = ((size_t)(u0 >> 6) * L.MB + (b >> 3)) * 512 + (size_t)((u0 & 63) >> 3) * 64 + (size_t)(b & 7) * 8;
                *reinterpret_cast<uint4*>(a.hx + off) = hh;
                *reinterpret_cast<uint4*>(a.hx + hx_part + off) = hl;
            }
            fence_proxy_async_all();
            named_bar_sync(1, 128);
            if (etid == 0) {
                banked_arrive(ctrH, c);
                banked_arrive(ctrY, c);
            }
        }
        // Gate inputs of the step: every lane loads (the rows beyond B re-read row 0 of the frame), so the gate math is branch-free.
        // Their frames were pulled into L2 two steps earlier by warp 8 (bulk prefetch in the window in which no CTA pulls h chunks;
        // see gru_tc2_bwd.cu for what HBM misses at the step top cost).
        float4 gxv[6];
        float4 mk[2] = {make_float4(1.f, 1.f, 1.f, 1.f), make_float4(1.f, 1.f, 1.f, 1.f)};
        auto fetch_inputs = [&](int tt) {
            const size_t rw = (size_t)tt * B + (act ? b : 0);
            const float* gp = f.gx + rw * 3 * H + u0;
#pragma unroll
            for (int gi = 0; gi < 3; ++gi) {
                gxv[2 * gi] = ldg_nc_v4_pinned(gp + (size_t)gi * H);
                gxv[2 * gi + 1] = ldg_nc_v4_pinned(gp + (size_t)gi * H + 4);
            }
            if (f.mask) {
                mk[0] = ldg_nc_v4_pinned(f.mask + rw * H + u0);
                mk[1] = ldg_nc_v4_pinned(f.mask + rw * H + u0 + 4);
            }
        };
        for (int t = 0; t < T; ++t) {
            const size_t row = (size_t)t * B + (act ? b : 0);
            fetch_inputs(t);
            if (etid == 0) T2_TRACE(0);
            if (etid == 0) mbar_expect_tx(inbox_full, (uint32_t)T2_S * L.slot_bytes);
            // ---- exchange of the K-slice partial sums ------------------------------------------------------
            mbar_wait(d1_full, (uint32_t)t & 1);
            if (etid == 0) T2_TRACE(1);
            tc_fence_after();
            // a warp whose 32 rows are all beyond B skips the drain: the 288 columns of D1 are TMEM-read-bandwidth bound
            // (147 KB at 64 B/clk for four warps)
            if (wact) {
#pragma unroll
            for (int g = 0; g < 3; ++g) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    float v[16], v2[16];
                    tmem_ld_x16(taddr + T2_COL_M0 + g * T2_UB + 16 * k, v);
                    tmem_ld_x16(taddr + T2_COL_C + g * T2_UB + 16 * k, v2);
                    if (two_main) {
                        float v3[16];
                        tmem_ld_x16(taddr + T2_COL_M1 + g * T2_UB + 16 * k, v3);
                        tmem_ld_wait();
#pragma unroll
                        for (int q = 0; q < 16; ++q) v[q] += v3[q];
                    } else {
                        tmem_ld_wait();
                    }
#pragma unroll
                    for (int q = 0; q < 16; ++q) v[q] = fmaf(v2[q], F16_LO_INV, v[q]);
                    if (b < L.MB * 8) {
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            float* d = stage + (size_t)(2 * k + h2) * slot_f + b * 24 + g * 8;
                            *reinterpret_cast<float4*>(d) = make_float4(v[8 * h2 + 0], v[8 * h2 + 1], v[8 * h2 + 2], v[8 * h2 + 3]);
                            *reinterpret_cast<float4*>(d + 4) = make_float4(v[8 * h2 + 4], v[8 * h2 + 5], v[8 * h2 + 6], v[8 * h2 + 7]);
                        }
                    }
                }
            }
            }
            fence_proxy_async_smem();
            named_bar_sync(3, 128);
            if (etid < T2_S)
                bulk_s2c(mapa(inbox_addr + (uint32_t)j * L.slot_bytes, (uint32_t)etid), stage + (size_t)etid * slot_f, L.slot_bytes,
                         mapa(inbox_bar_addr, (uint32_t)etid));
            if (etid == 0) T2_TRACE(2);
            // W_y y_{t-1} of the own units: columns [r 8 | z 8 | n 8 | pad 8] (+ correction half at +32)
            float yr[8], yz[8], yn[8];
            mbar_wait(accum_full, (uint32_t)t & 1);
            if (etid == 0) T2_TRACE(4);
            tc_fence_after();
            {
                float c2[8];
                tmem_ld_x8(taddr + T2_COL_Y, yr);
                tmem_ld_x8(taddr + T2_COL_Y + 32, c2);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q) yr[q] = fmaf(c2[q], F16_LO_INV, yr[q]);
                tmem_ld_x8(taddr + T2_COL_Y + 8, yz);
                tmem_ld_x8(taddr + T2_COL_Y + 40, c2);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q) yz[q] = fmaf(c2[q], F16_LO_INV, yz[q]);
                tmem_ld_x8(taddr + T2_COL_Y + 16, yn);
                tmem_ld_x8(taddr + T2_COL_Y + 48, c2);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q) yn[q] = fmaf(c2[q], F16_LO_INV, yn[q]);
            }
            tc_fence_before();
            mbar_wait_cluster(inbox_full, (uint32_t)t & 1);
            if (etid == 0) T2_TRACE(3);
            float ar[8], az[8], an[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) ar[q] = az[q] = an[q] = 0.f;
            {
                const int bs = b < L.MB * 8 ? b : 0;   // rows beyond the staged row groups read row 0 (in bounds)
#pragma unroll
                for (int p = 0; p < T2_S; ++p) {   // fixed order: deterministic
                    const float4* x = reinterpret_cast<const float4*>(inbox + (size_t)p * slot_f + bs * 24);
                    const float4 x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3], x4 = x[4], x5 = x[5];
                    ar[0] += x0.x; ar[1] += x0.y; ar[2] += x0.z; ar[3] += x0.w; ar[4] += x1.x; ar[5] += x1.y; ar[6] += x1.z; ar[7] += x1.w;
                    az[0] += x2.x; az[1] += x2.y; az[2] += x2.z; az[3] += x2.w; az[4] += x3.x; az[5] += x3.y; az[6] += x3.z; az[7] += x3.w;
                    an[0] += x4.x; an[1] += x4.y; an[2] += x4.z; an[3] += x4.w; an[4] += x5.x; an[5] += x5.y; an[6] += x5.z; an[7] += x5.w;
                }
            }
            float rr[8], zz[8], nn[8], gh[8], ov[8];
            {
                const float* gxr = reinterpret_cast<const float*>(&gxv[0]);
                const float* gxz = reinterpret_cast<const float*>(&gxv[2]);
                const float* gxn = reinterpret_cast<const float*>(&gxv[4]);
                const float* mkf = reinterpret_cast<const float*>(&mk[0]);
                float bh[24];
#pragma unroll
                for (int q = 0; q < 6; ++q) *reinterpret_cast<float4*>(bh + 4 * q) = *reinterpret_cast<const float4*>(sBh + 4 * q);
#pragma unroll
                for (int q = 0; q < 8; ++q) {   // straight-line for every lane (gru_tc.cu)
                    rr[q] = sigmoid_fast(gxr[q] + yr[q] + ar[q] + bh[q]);
                    zz[q] = sigmoid_fast(gxz[q] + yz[q] + az[q] + bh[8 + q]);
                    gh[q] = an[q] + bh[16 + q];
                    nn[q] = tanh_fast(gxn[q] + yn[q] + rr[q] * gh[q]);
                    hreg[q] = (1.0f - zz[q]) * nn[q] + zz[q] * hreg[q];
                    ov[q] = hreg[q] * mkf[q];
                }
            }
            if (etid == 0) T2_TRACE(10);
            uint4 oh, ol, hh, hl;
            t2_split8(hreg, hh, hl);
            if (act) {   // publish h_t (fp16 hi/lo, operand order) into the other exchange slot
                uint16_t* hdst = a.hx + (size_t)((t + 1) & 1) * 2 * hx_part;
                const size_t off = ((size_t)(u0 >> 6) * L.MB + (b >> 3)) * 512 + (size_t)((u0 & 63) >> 3) * 64 + (size_t)(b & 7) * 8;
                *reinterpret_cast<uint4*>(hdst + off) = hh;
                *reinterpret_cast<uint4*>(hdst + hx_part + off) = hl;
            }
            fence_proxy_async_global();   // own generic writes of h_t -> visible to the peers' bulk copies (async proxy)
            named_bar_arrive(6, 160);     // warp 8 releases counter H
            if (etid == 0) T2_TRACE(7);
            t2_split8(ov, oh, ol);
            if (act) {   // o_t of the own units: k block j of the cluster's o operand, here and in the three peers (st.async: every
                         // thread sends its row as soon as it has it; the bytes are counted on the receivers' barriers)
                const uint32_t off = (uint32_t)j * L.kstr + (uint32_t)(b >> 3) * 128u + (uint32_t)(b & 7) * 16u;
                *reinterpret_cast<uint4*>(sA2 + off) = oh;
                *reinterpret_cast<uint4*>(sA2 + off + L.pstr) = ol;
                const uint32_t a2_addr = smem_u32(sA2) + off, bar_addr = smem_u32(a2_full);
#pragma unroll
                for (int pp = 1; pp < T2_S; ++pp) {
                    const uint32_t p = (uint32_t)((j + pp) & (T2_S - 1));
                    const uint32_t rb = mapa(bar_addr, p);
                    st_async_v4(mapa(a2_addr, p), oh, rb);
                    st_async_v4(mapa(a2_addr + L.pstr, p), ol, rb);
                }
            }
            fence_proxy_async_smem();   // the own rows (generic stores) -> the tensor core
            named_bar_sync(3, 128);
            if (etid == 0) mbar_expect_tx(a2_full, (uint32_t)(T2_S - 1) * (uint32_t)B * 32u);   // 32 bytes per row and peer
            if (etid == 0) T2_TRACE(6);
            // cluster partial of the own quarter -> fixed-point totals of slot t & 1 (lane == row: coalesced 8-byte adds)
            mbar_wait(part_full, (uint32_t)t & 1);
            if (etid == 0) T2_TRACE(25);
            tc_fence_after();
            {
                float v[16], v2[16];
                tmem_ld_x16(taddr + T2_COL_P, v);
                tmem_ld_x16(taddr + T2_COL_P + T2_OQ, v2);
                tmem_ld_wait();
                if (act) {
                    unsigned long long* d = a.yacc + (size_t)(t & 1) * yslot + (size_t)(T2_OQ * j) * RP + b;
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        if (T2_OQ * j + q < out && !(a.dbg & 1)) red_add_u64(d + (size_t)q * RP, (unsigned long long)__float2ll_rn(fmaf(v2[q], F16_LO_INV, v[q]) * T2_FX));
                }
            }
            tc_fence_before();
            if (etid == 0) T2_TRACE(26);
            named_bar_sync(7, 128);    // every finaliser has added its rows of the partial
            if (etid == 0) banked_arrive(ctrY, c);   // release is cumulative over the barrier: one gpu-scope fence per CTA
            if (etid == 0) T2_TRACE(9);
            if (act) {   // outputs / saved activations that only later kernels read: after both releases (a gpu-scope fence waits for
                         // every store the SM has in flight: issued before the release of counter H they cost it 1 000 cycles)
                float* hd = f.hs + (size_t)(t + 1) * B * H + (size_t)b * H + u0;
                *reinterpret_cast<float4*>(hd) = make_float4(hreg[0], hreg[1], hreg[2], hreg[3]);
                *reinterpret_cast<float4*>(hd + 4) = make_float4(hreg[4], hreg[5], hreg[6], hreg[7]);
                const size_t so = row * H + u0;
                if (f.sv_r) {
                    *reinterpret_cast<float4*>(f.sv_r + so) = make_float4(rr[0], rr[1], rr[2], rr[3]);
                    *reinterpret_cast<float4*>(f.sv_r + so + 4) = make_float4(rr[4], rr[5], rr[6], rr[7]);
                    *reinterpret_cast<float4*>(f.sv_z + so) = make_float4(zz[0], zz[1], zz[2], zz[3]);
                    *reinterpret_cast<float4*>(f.sv_z + so + 4) = make_float4(zz[4], zz[5], zz[6], zz[7]);
                    *reinterpret_cast<float4*>(f.sv_n + so) = make_float4(nn[0], nn[1], nn[2], nn[3]);
                    *reinterpret_cast<float4*>(f.sv_n + so + 4) = make_float4(nn[4], nn[5], nn[6], nn[7]);
                    *reinterpret_cast<float4*>(f.sv_ghn + so) = make_float4(gh[0], gh[1], gh[2], gh[3]);
                    *reinterpret_cast<float4*>(f.sv_ghn + so + 4) = make_float4(gh[4], gh[5], gh[6], gh[7]);
                }
                if (f.sv_o) {
                    *reinterpret_cast<float4*>(f.sv_o + so) = make_float4(ov[0], ov[1], ov[2], ov[3]);
                    *reinterpret_cast<float4*>(f.sv_o + so + 4) = make_float4(ov[4], ov[5], ov[6], ov[7]);
                }
            }
        }
    } else if (warp >= 8) {
        // ================= aux: totals of y_{t-1} -> own quarter of the y operand; w8 releases counter H ===============
        const int rt = threadIdx.x - 256;          // batch row
        const bool y_act = rt < B;
        const int yo = T2_OQ * j;                  // first output of the quarter
        const unsigned long long* ysrc = a.yacc + (size_t)yo * RP + (y_act ? rt : 0);
        unsigned long long p0[16], p1[16];         // the totals this thread saw two steps / one step ago (slot parity)
        float ybias[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            p0[e] = p1[e] = 0ull;
            ybias[e] = (yo + e < out) ? __ldg(f.bo + yo + e) : 0.f;
        }
        for (int t = 0; t <= T; ++t) {
            // step T is the epilogue: y_{T-1} is read once more for the fp32 output (cluster 0 only)
            if (t == T && ci != 0) break;
            if (warp == 8) {
                banked_wait_warp(ctrY, per_bank * (unsigned)(t + 1), lane, a.relaxed != 0);
                if (lane == 0) T2_TRACE(20);
            }
            named_bar_sync(5, 128);
            float yv[16];
            if (t > 0) {
                const unsigned long long* src = ysrc + (size_t)((t + 1) & 1) * yslot;   // slot of step t-1
                unsigned long long cur[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) cur[e] = ld_total_u64(src + (size_t)e * RP);
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    yv[e] = fmaf(__ll2float_rn((long long)(cur[e] - p0[e])), T2_FX_INV, ybias[e]);
                    p0[e] = p1[e];
                    p1[e] = cur[e];
                }
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) yv[e] = (y_act && yo + e < out) ? f.ys[(size_t)rt * out + yo + e] : 0.f;
            }
            if (rt == 0) T2_TRACE(27);
            if (t == T) {
                if (y_act && t > 0 && ci == 0) {   // the fp32 output y_{t-1} (slot t of ys): one cluster stores it
                    float* yd = f.ys + (size_t)t * n_pairs + (size_t)rt * out + yo;
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (yo + e < out) yd[e] = yv[e];
                }
                break;
            }
            if (y_act) {   // two core-matrix rows per plane: this CTA's two k blocks of the y operand, here and in the three peers
                uint4 hi0, lo0, hi1, lo1;
                const uint32_t off = (uint32_t)(2 * j) * L.kstr + (uint32_t)(rt >> 3) * 128u + (uint32_t)(rt & 7) * 16u;
                t2_split8(yv, hi0, lo0);
                t2_split8(yv + 8, hi1, lo1);
                *reinterpret_cast<uint4*>(ybuf + off) = hi0;
                *reinterpret_cast<uint4*>(ybuf + off + L.pstr) = lo0;
                *reinterpret_cast<uint4*>(ybuf + off + L.kstr) = hi1;
                *reinterpret_cast<uint4*>(ybuf + off + L.kstr + L.pstr) = lo1;
                const uint32_t y_addr = smem_u32(ybuf) + off, bar_addr = smem_u32(y_full);
#pragma unroll
                for (int pp = 1; pp < T2_S; ++pp) {
                    const uint32_t p = (uint32_t)((j + pp) & (T2_S - 1));
                    const uint32_t rb = mapa(bar_addr, p);
                    st_async_v4(mapa(y_addr, p), hi0, rb);
                    st_async_v4(mapa(y_addr + L.pstr, p), lo0, rb);
                    st_async_v4(mapa(y_addr + L.kstr, p), hi1, rb);
                    st_async_v4(mapa(y_addr + L.kstr + L.pstr, p), lo1, rb);
                }
            }
            fence_proxy_async_smem();   // the own rows (generic stores) -> the tensor core
            named_bar_sync(5, 128);
            if (rt == 0) {
                mbar_expect_tx(y_full, (uint32_t)(T2_S - 1) * (uint32_t)B * 64u);   // 64 bytes per row and peer
                T2_TRACE(21);
            }
            if (warp == 8) {   // counter H: the finalisers arrive (without waiting) once h_t is published and fenced
                named_bar_sync(6, 160);
                if (lane == 0) banked_arrive(ctrH, c);   // release is cumulative over the barrier
                if (lane == 0) T2_TRACE(22);
                // gate inputs of frame t+2 -> L2: this CTA's 1/G of the gx frame and of the mask frame
                if (t + 2 < T && lane < 2 && !(a.dbg & 2)) {
                    const size_t frame = (size_t)B * H * (lane == 0 ? 3 : 1), slice = frame / (size_t)G;   // a multiple of 16 bytes
                    const float* base = lane == 0 ? f.gx : f.mask;
                    if (base) bulk_prefetch_l2(base + (size_t)(t + 2) * frame + (size_t)c * slice, (uint32_t)(slice * sizeof(float)));
                }
            }
            if (y_act && t > 0 && ci == 0) {   // the fp32 output y_{t-1} (slot t of ys): one cluster stores it, after the release
                float* yd = f.ys + (size_t)t * n_pairs + (size_t)rt * out + yo;
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if (yo + e < out) yd[e] = yv[e];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // no CTA leaves while a peer may still copy into its shared memory
    if (a.trace && c == 0 && threadIdx.x == 0) a.trace[62] = clock64();
    if (warp == 2) tmem_dealloc<512>(tmem);
}

// ---- host side ----------------------------------------------------------------------------------------
static bool t2_shape_ok(int B, int H, int out) {
    return H % (T2_KC * T2_S) == 0 && H >= T2_KC * T2_S && (H / T2_UB) % 4 == 0 && H / T2_UB <= 32 && out >= 1 && out <= 64 && B >= 1 && B <= 128;
}

size_t gru_tc2_scratch_floats(int B, int H) {
    if (H % T2_UB != 0) return 0;
    const size_t MB = (B + 7) / 8;
    const size_t hx = (size_t)2 * 2 * (H / T2_KC) * MB * 512 / 2;   // fp16 elements -> floats
    const size_t yacc = (size_t)2 * 64 * MB * 8 * 2;   // u64 -> floats
    return round_up_sz(hx, 64) + 1024 + round_up_sz(yacc, 64);
}

// are all G/4 clusters co-resident at this shape?  (cached per shape)
static bool t2_runnable(int B, int H, int out, const DeviceInfo& di, T2Layout* Lout) {
    const int G = H / 8;
    if (!t2_shape_ok(B, H, out) || G > di.n_sm) return false;
    struct Entry { int B, H, out, ok; };
    static Entry cache[256];
    static int n_cache = 0;
    int ok = -1;
    for (int i = 0; i < n_cache; ++i)
        if (cache[i].B == B && cache[i].H == H && cache[i].out == out) ok = cache[i].ok;
    T2Layout L = t2_layout(B, H, di.max_smem_optin);
    if (ok < 0) {
        ok = 0;
        if (L.NS >= 2 && (int)L.total <= di.max_smem_optin && (uint32_t)L.NS * L.stage_bytes >= (uint32_t)T2_S * L.slot_bytes &&
            cudaFuncSetAttribute(k_gru_fwd_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total) == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(G);
            cfg.blockDim = dim3(T2_NT);
            cfg.dynamicSmemBytes = L.total;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = T2_S;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int ncl = 0;
            if (cudaOccupancyMaxActiveClusters(&ncl, k_gru_fwd_tc2, &cfg) == cudaSuccess) {
                if (getenv("CVB_DEBUG"))
                    fprintf(stderr, "[cvb] k_gru_fwd_tc2: %d co-resident clusters of %d (need %d), smem %u, ring %d\n", ncl, T2_S, G / T2_S, L.total, L.NS);
                ok = ncl * T2_S >= G ? 1 : 0;
            }
        }
        cudaGetLastError();
        if (n_cache < 256) cache[n_cache++] = Entry{B, H, out, ok};
    }
    if (ok && Lout) *Lout = L;
    return ok != 0;
}

bool gru_tc2_supported(int B, int H, int out, const DeviceInfo& di) { return t2_runnable(B, H, out, di, nullptr); }

int gru_ar_fwd_tc2(GruFwdArgs& f, float* tc_scratch, cudaStream_t s) {
    if (f.T <= 0 || f.B <= 0) return 0;
    DeviceInfo di;
    if (int rc = get_device_info(&di)) return rc;
    T2Layout L;
    CVB_REQUIRE(t2_runnable(f.B, f.H, f.out, di, &L), "gru_ar_fwd_tc2: unsupported shape B=%d H=%d out=%d", f.B, f.H, f.out);
    GruTc2Args a;
    a.f = f;
    const size_t hx_f = round_up_sz((size_t)2 * 2 * (f.H / T2_KC) * L.MB * 512 / 2, 64);
    a.hx = reinterpret_cast<uint16_t*>(tc_scratch);
    a.ctr = reinterpret_cast<unsigned*>(tc_scratch + hx_f);
    a.yacc = reinterpret_cast<unsigned long long*>(tc_scratch + hx_f + 1024);
    a.smem_max = di.max_smem_optin;
    a.keepalive = 1;
    a.relaxed = relaxed_polling() ? 1 : 0;
    a.dbg = 0;
    if (const char* e = getenv("CVB_TC_DBG")) a.dbg = atoi(e);
    if (const char* e = getenv("CVB_TC_KEEPALIVE")) a.keepalive = atoi(e) != 0;
    a.trace = nullptr;
    const char* trace_file = getenv("CVB_TRACE_FILE_FWD");
    const size_t trace_bytes = ((size_t)(f.T + 1) * 64 + 8 * 256 + 48 * 256) * sizeof(long long);
    if (trace_file && trace_file[0]) {
        CVB_CHECK(cudaMalloc(&a.trace, trace_bytes));
        CVB_CHECK(cudaMemsetAsync(a.trace, 0, trace_bytes, s));
    }
    CVB_CHECK(cudaMemsetAsync(a.ctr, 0, (1024 + (size_t)2 * 64 * L.MB * 8 * 2) * sizeof(float), s));   // the counter banks + the totals
    CVB_CHECK(cudaFuncSetAttribute(k_gru_fwd_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(f.H / 8);
    cfg.blockDim = dim3(T2_NT);
    cfg.dynamicSmemBytes = L.total;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = T2_S;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = launch_without_coop() ? 1 : 2;
    prof_begin(s, CVB_PROF_GRU_FWD);
    CVB_CHECK(cudaLaunchKernelEx(&cfg, k_gru_fwd_tc2, a));
    prof_end(s, CVB_PROF_GRU_FWD);
    count_launch();
    if (a.trace) {   // profiling hook only: synchronises
        CVB_CHECK(cudaStreamSynchronize(s));
        long long* h = (long long*)malloc(trace_bytes);
        CVB_CHECK(cudaMemcpy(h, a.trace, trace_bytes, cudaMemcpyDeviceToHost));
        if (FILE* fp = fopen(trace_file, "wb")) {
            fwrite(h, 1, trace_bytes, fp);
            fclose(fp);
        }
        free(h);
        CVB_CHECK(cudaFree(a.trace));
    }
    return 0;
}

}  // namespace cvb
