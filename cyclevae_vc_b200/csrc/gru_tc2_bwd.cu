// Tensor-core (tcgen05) BPTT of the autoregressive GRU (the reverse of gru_vae.py:364-399), ONE
// grid-wide hop per recurrent step on each of its two dependency chains (gru_tc_bwd.cu, the first version, chains two).
//
// Same 2-D split as gru_tc_bwd.cu at S = 4: cluster i owns the hidden units [32i, 32i+32), CTA j owns the K-slice
// [j 3H/4, (j+1) 3H/4) of  dh[b,u] += sum_k dgh_{t+1}[b,k] W_hh[k,u]  for all units of the block and FINALISES the units
// [32i + 8j, +8); W_hh^T of (block x K-slice) resident in shared memory as bf16 hi+lo; partial accumulators meet in the
// finaliser's shared memory through bulk DSMEM copies.
//
// The y feedback (dy_t = dY_t + dgi_{t+1} W_y, a reduction over all of 3H; then dh_t += (dy_t W_o) * m_t) follows
// gru_tc2.cu: the four CTAs of a cluster swap dgi of their units through DSMEM (K = 96 operand, K-outer layout), CTA j
// forms the cluster's partial of the output quarter [16j, 16j+16) with one MMA chain (D3) and ADDS it to 64-bit
// fixed-point running totals in L2 (exact, order-independent: deterministic); when counter Y is complete CTA j reads
// quarter j of the totals, adds the head's dY, splits to bf16 hi/lo and the four quarters of the dy operand are swapped
// through DSMEM; q = dy_t W_o[:, own units] is the same 64-deep MMA chain as before (D2).  Counter H (dgh_t published)
// is released right after the gate math, so the next step's W_hh chain starts while the partial is still being formed.
// The fixed-point scale is chosen per launch from max |dY|, |dh_last| (found by the grid during set-up): 2^38 / max, so
// the resolution is 4e-12 of the largest incoming gradient and totals up to 2^24 times larger fit.
//
// Roles (384 threads): w0 bulk-copy producer, w1 MMA issuer, w2 TMEM allocator, w4-7 exchange + finalise
// (TMEM lane == batch row), w8-11 totals of dy -> operand (one row per thread), w8 releases counter H.
#include <stdlib.h>

#include "gru_ar.cuh"
#include "umma.cuh"

namespace cvb {
using namespace umma;

constexpr int U2_NT = 384;
constexpr int U2_KC = 64;        // K per ring stage
constexpr int U2_S = 4;          // cluster size
constexpr int U2_UB = 8 * U2_S;  // units per cluster
constexpr int U2_OQ = 16;        // outputs per quarter
constexpr int U2_KB3 = 3 * U2_S; // k blocks of the cluster's dgi operand (r, z, n of four CTAs)
constexpr uint32_t U2_COL_Q = 128;     // D2 (q): 2 x 16 columns
constexpr uint32_t U2_COL_DUMMY = 160; // keep-alive scratch, 16 columns
constexpr uint32_t U2_COL_P = 256;     // D3 (cluster partial of the own output quarter): 2 x 16 columns

struct U2Layout {
    int MB, nch, NS;
    uint32_t half, stage_bytes, w_part_bytes, slot_bytes, kstr, pstr;
    uint32_t off_ring, off_w, off_inbox, off_a2, off_b2, off_b3, off_ybuf, off_bar, total;
};

__host__ __device__ inline U2Layout u2_layout(int B, int H, int smem_max) {
    U2Layout L;
    L.MB = (B + 7) / 8;
    L.nch = 3 * H / U2_KC / U2_S;
    L.half = (uint32_t)L.MB * 1024u;
    L.stage_bytes = 2u * L.half;
    L.w_part_bytes = (uint32_t)L.nch * (uint32_t)U2_S * 1024u;
    L.slot_bytes = (uint32_t)L.MB * 8u * 32u;
    L.pstr = (uint32_t)L.MB * 128u;
    L.kstr = 2u * L.pstr;
    const uint32_t inbox = (uint32_t)U2_S * L.slot_bytes;
    const uint32_t fixed = 2u * L.w_part_bytes + inbox + (uint32_t)U2_KB3 * L.kstr + 4096u + 6144u + L.stage_bytes + 256u;
    int ns = ((int)smem_max - (int)fixed) / (int)L.stage_bytes;
    L.NS = ns > 6 ? 6 : ns;
    const uint32_t ring = (uint32_t)(L.NS > 0 ? L.NS : 0) * L.stage_bytes;
    L.off_ring = 0;                      // dgh chunks; idle between a step's last chunk and the next step's first, when it doubles
                                         // as the staging of the outgoing partial sums
    // An MMA with M = 128 reads 16 row groups of every A operand whatever MB is: the K-outer operands are placed so that
    // this over-read (up to 2 KB behind the operand) stays inside the allocation.
    L.off_ybuf = ring;                   // dy_t operand, K-outer: [8 k blocks][hi | lo][MB][128 B]
    L.off_w = L.off_ybuf + L.stage_bytes;
    L.off_inbox = L.off_w + 2u * L.w_part_bytes;
    L.off_a2 = L.off_inbox + inbox;      // dgi of the cluster's units, K-outer: [12 k blocks][hi | lo][MB][128 B]
    L.off_b2 = L.off_a2 + (uint32_t)U2_KB3 * L.kstr;   // [2 parts][2 n blocks][8 kblk][8][8] bf16: W_o^T (own units)
    L.off_b3 = L.off_b2 + 4096u;         // W_y[cluster's gate rows][own quarter]: [hi 2 groups | lo 2 groups] x 12 k blocks x 128 B
    L.off_bar = L.off_b3 + 6144u;
    L.total = L.off_bar + 256u;
    return L;
}

struct GruTc2BwdArgs {
    GruBwdArgs f;
    uint16_t* gxh;    // [2 slots][2 parts][3H/64 chunks][MB][8 kblk][8 rows][8 k] bf16 (UMMA order) of dgh_t
    unsigned long long* yacc;   // [2 slots][64 outputs][MB*8 rows] running fixed-point totals of the feedback, zero-initialised
    unsigned* ctr;    // CTR_BANKS lines of counter H, CTR_BANKS lines of counter Y (common.cuh), then set-up arrivals and max |incoming gradient| (float bits); zero-initialised
    int smem_max;
    int keepalive;
    int relaxed;
    int dbg;            // CVB_TC_DBG bits (experiments): 1 = skip the fixed-point adds (wrong results), 2 = no L2 prefetch of the next frames
    long long* trace;   // optional [T+1][64] clock64 stamps of CTA 0 (CVB_TRACE_FILE), else null
};

#define U2_TRACE(ev)                                                     \
    do {                                                                 \
        if (a.trace && c == 0) a.trace[(size_t)n * 64 + (ev)] = clock64(); \
    } while (0)

static __device__ __forceinline__ uint4 u2_pack(const uint16_t* v) {
    return make_uint4((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16),
                      (uint32_t)v[4] | ((uint32_t)v[5] << 16), (uint32_t)v[6] | ((uint32_t)v[7] << 16));
}
static __device__ __forceinline__ void u2_split8(const float* x, uint4& hi, uint4& lo) {
    uint16_t h[8], l[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) split_bf16(x[q], h[q], l[q]);
    hi = u2_pack(h);
    lo = u2_pack(l);
}
static __device__ __forceinline__ unsigned long long u2_ld_total(const unsigned long long* p) {   // see gru_tc2.cu
    unsigned long long v;
    asm volatile("ld.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
static __device__ __forceinline__ void u2_red_add(unsigned long long* p, unsigned long long v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// fixed-point scale 2^e with max * 2^e in [2^37, 2^38] (max = 0: e = 38): returns the scale and its inverse, both powers of two
static __device__ __forceinline__ void u2_scale(float mx, float& fx, float& fx_inv) {
    int e = 38;
    if (mx > 0.f && mx < 3.0e38f) {
        int ex;
        frexpf(mx, &ex);   // mx = m * 2^ex, 0.5 <= m < 1
        e = 38 - ex;
    }
    e = e > 100 ? 100 : (e < -60 ? -60 : e);
    fx = ldexpf(1.0f, e);
    fx_inv = ldexpf(1.0f, -e);
}

__global__ void __launch_bounds__(U2_NT, 1) k_gru_bwd_tc2(GruTc2BwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const GruBwdArgs& f = a.f;
    const int B = f.B, T = f.T, H = f.H, out = f.out, K3 = 3 * f.H;
    const int G = gridDim.x, c = blockIdx.x;
    constexpr int S = U2_S;
    const int j = (int)cluster_ctarank();
    const U2Layout L = u2_layout(B, H, a.smem_max);
    const int ci = c / S;
    const int ublk0 = ci * U2_UB;         // first unit of the cluster's block
    const int u0 = ublk0 + 8 * j;         // first of the 8 units this CTA finalises
    const int k0 = j * L.nch * U2_KC;     // first row of W_hh (= column of dgh) of this CTA's K-slice
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    uint8_t* ring = smem + L.off_ring;
    float* stage = reinterpret_cast<float*>(ring);                 // [S (to)][MB*8][8], aliases the (idle) ring
    uint8_t* sW = smem + L.off_w;
    float* inbox = reinterpret_cast<float*>(smem + L.off_inbox);   // [S (from)][MB*8][8]
    uint8_t* sA2 = smem + L.off_a2;
    uint8_t* sB2 = smem + L.off_b2;
    uint8_t* sB3 = smem + L.off_b3;
    uint8_t* ybuf = smem + L.off_ybuf;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* empty = full + 8;
    uint64_t* accum_full = full + 16;   // D2 (q) complete
    uint64_t* d1_full = full + 17;      // D1 complete
    uint64_t* y_full = full + 18;       // the four quarters of the dy operand are in place
    uint64_t* inbox_full = full + 19;
    uint64_t* a2_full = full + 20;      // dgi of the cluster's units is in place
    uint64_t* part_full = full + 21;    // D3 complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 22);
    float* s_max = reinterpret_cast<float*>(full + 23);            // [12] per-warp maxima of the set-up scan
    const size_t gx_part = (size_t)(K3 / U2_KC) * L.MB * 512;   // elements per part
    unsigned* ctrH = a.ctr;
    unsigned* ctrY = a.ctr + 32 * CTR_BANKS;
    unsigned* ctrS = a.ctr + 64 * CTR_BANKS;
    unsigned* gmax = ctrS + 1;
    const unsigned per_bank = (unsigned)(G / CTR_BANKS);
    const int n_pairs = B * out;
    const size_t RP = (size_t)L.MB * 8;
    const size_t yslot = (size_t)64 * RP;
    const uint32_t ypiece = 2u * L.kstr;   // a CTA's contribution to the dy operand (2 k blocks)
    const uint32_t opiece = 3u * L.kstr;   // ... to the dgi operand (3 k blocks: r, z, n of its units)

    if (a.trace && c == 0 && threadIdx.x == 0) a.trace[60] = clock64();
    // ---- one-time setup --------------------------------------------------------------------------
    {
        {   // this CTA's share of max |dY|, |dh_last| -> global maximum (non-negative floats order like their bit patterns)
            float m = 0.f;
            const size_t n_dy = (size_t)(T + 1) * n_pairs, n_dh = (size_t)B * H;
            for (size_t i = (size_t)c * U2_NT + threadIdx.x; i < n_dy; i += (size_t)G * U2_NT) m = fmaxf(m, fabsf(__ldg(f.dy_tot + i)));
            for (size_t i = (size_t)c * U2_NT + threadIdx.x; i < n_dh; i += (size_t)G * U2_NT) m = fmaxf(m, fabsf(__ldg(f.dhc + i)));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) s_max[warp] = m;
        }
        // one item = an 8 (k) x 4 (n) block of W_hh^T: 8 float4 reads along the contiguous unit axis, transposed in
        // registers into four 16-byte core-matrix rows (hi) + four (lo)
        const int nq = U2_UB >> 2;
        const int n_items = nq * L.nch * (U2_KC / 8);
        for (int i = threadIdx.x; i < n_items; i += U2_NT) {
            const int kg = i / nq, n4 = (i - kg * nq) * 4;
            const int kl = kg * 8;
            float4 r[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) r[q] = __ldg(reinterpret_cast<const float4*>(f.Whh + (size_t)(k0 + kl + q) * H + ublk0 + n4));
            const uint32_t off = (uint32_t)(kl / U2_KC) * ((uint32_t)S * 2048u) + (uint32_t)((kl % U2_KC) >> 3) * 128u;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float w[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) w[q] = e == 0 ? r[q].x : e == 1 ? r[q].y : e == 2 ? r[q].z : r[q].w;
                uint4 hi, lo;
                u2_split8(w, hi, lo);
                const int nn = n4 + e;
                const uint32_t o2 = off + (uint32_t)(nn >> 3) * 1024u + (uint32_t)(nn & 7) * 16u;
                *reinterpret_cast<uint4*>(sW + o2) = hi;                       // chunk layout: [hi: S blocks][lo: S blocks]
                *reinterpret_cast<uint4*>(sW + (uint32_t)S * 1024u + o2) = lo;
            }
        }
        {   // B2[n][k] = W_o[k][u0 + n] (n < 8, k < out).
            // B3[n = o_local][k = (3 jj + g) * 8 + uu] = W_y[g*H + ublk0 + 8 jj + uu][16 j + o_local]: rows [hi 16 | lo 16], 12 k blocks
            constexpr int M2 = (16 * 64 + U2_NT - 1) / U2_NT, M3 = (U2_OQ * 96 + U2_NT - 1) / U2_NT;
            float w2[M2], w3[M3];
#pragma unroll
            for (int m = 0; m < M2; ++m) {
                const int i = threadIdx.x + m * U2_NT, nn = i >> 6, k = i & 63;
                w2[m] = (i < 16 * 64 && nn < 8 && k < out) ? __ldg(f.Wo + (size_t)k * H + u0 + nn) : 0.f;
            }
#pragma unroll
            for (int m = 0; m < M3; ++m) {
                const int i = threadIdx.x + m * U2_NT, k = i >> 4, nn = i & 15;   // consecutive threads: consecutive outputs (contiguous in W_y)
                const int kb = k >> 3, jj = kb / 3, g = kb - 3 * jj, uu = k & 7;
                const int o = U2_OQ * j + nn;
                w3[m] = (i < U2_OQ * 96 && o < out) ? __ldg(f.Wy + (size_t)(g * H + ublk0 + 8 * jj + uu) * f.ldwy + o) : 0.f;
            }
#pragma unroll
            for (int m = 0; m < M2; ++m) {
                const int i = threadIdx.x + m * U2_NT, nn = i >> 6, k = i & 63;
                if (i < 16 * 64) {
                    uint16_t hi, lo;
                    split_bf16(w2[m], hi, lo);
                    const uint32_t off = (uint32_t)(nn >> 3) * 1024u + (uint32_t)(k >> 3) * 128u + (uint32_t)(nn & 7) * 16u + (uint32_t)(k & 7) * 2u;
                    *reinterpret_cast<uint16_t*>(sB2 + off) = hi;
                    *reinterpret_cast<uint16_t*>(sB2 + 2048 + off) = lo;
                }
            }
#pragma unroll
            for (int m = 0; m < M3; ++m) {
                const int i = threadIdx.x + m * U2_NT, k = i >> 4, nn = i & 15;
                if (i < U2_OQ * 96) {
                    uint16_t hi, lo;
                    split_bf16(w3[m], hi, lo);
                    const uint32_t off = (uint32_t)(nn >> 3) * 1536u + (uint32_t)(k >> 3) * 128u + (uint32_t)(nn & 7) * 16u + (uint32_t)(k & 7) * 2u;
                    *reinterpret_cast<uint16_t*>(sB3 + off) = hi;
                    *reinterpret_cast<uint16_t*>(sB3 + 3072 + off) = lo;
                }
            }
        }
        for (uint32_t i = threadIdx.x; i < (uint32_t)U2_KB3 * L.kstr / 16; i += U2_NT) reinterpret_cast<uint4*>(sA2)[i] = make_uint4(0u, 0u, 0u, 0u);
        for (uint32_t i = threadIdx.x; i < L.stage_bytes / 16; i += U2_NT) reinterpret_cast<uint4*>(ybuf)[i] = make_uint4(0u, 0u, 0u, 0u);
        if (threadIdx.x == 0) {
            for (int s = 0; s < 8; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            mbar_init(accum_full, 1);
            mbar_init(d1_full, 1);
            mbar_init(y_full, 1);       // one arrive.expect_tx by the owner per step; the three peers' pieces complete_tx
            mbar_init(inbox_full, 1);
            mbar_init(a2_full, 1);      // same protocol
            mbar_init(part_full, 1);
            mbar_fence_init();
        }
        fence_proxy_async_smem();
        if (warp == 2) tmem_alloc<512>(tmem_slot);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (threadIdx.x == 0) {
            float m = 0.f;
            for (int w = 0; w < U2_NT / 32; ++w) m = fmaxf(m, s_max[w]);
            atomicMax(gmax, __float_as_uint(m));
            red_release_gpu_add(ctrS, 1u);
        }
    }
    cluster_sync_all();   // every CTA's barriers are initialised before any peer copies into it
    const uint32_t tmem = *tmem_slot;
    if (a.trace && c == 0 && threadIdx.x == 0) a.trace[61] = clock64();

    if (warp == 0) {
        // ================= producer: K-slice of dgh_{t+1} chunk by chunk ================================
        int s = 0;
        uint32_t ph = 1;   // parity to wait on the empty barrier of stage s (first pass: free)
        for (int n = 1; n <= T; ++n) {
            const uint16_t* src = a.gxh + (size_t)((n - 1) & 1) * 2 * gx_part + (size_t)(j * L.nch) * L.MB * 512;
            banked_wait_warp(ctrH, per_bank * (unsigned)n, lane, a.relaxed != 0);
            if (lane == 0) U2_TRACE(14);
            for (int ch = 0; ch < L.nch; ++ch) {
                if (lane == 0) {
                    mbar_wait(&empty[s], ph);
                    if (ch < 8) U2_TRACE(32 + ch);
                    uint8_t* dst = ring + (size_t)s * L.stage_bytes;
                    mbar_expect_tx(&full[s], 2 * L.half);
                    bulk_g2s(dst, src + (size_t)ch * L.MB * 512, L.half, &full[s]);
                    bulk_g2s(dst + L.half, src + gx_part + (size_t)ch * L.MB * 512, L.half, &full[s]);
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
        const uint32_t idesc1 = idesc_bf16_f32(128, U2_UB), idesc1s = idesc_bf16_f32(128, 2 * U2_UB);
        const uint32_t idesc2 = idesc_bf16_f32(128, 16), idesc2s = idesc_bf16_f32(128, 32);
        const uint32_t idesc3 = idesc_bf16_f32(128, U2_OQ), idesc3s = idesc_bf16_f32(128, 2 * U2_OQ);
        const uint64_t dA0 = smem_desc(smem_u32(ring), 128, 1024);
        const uint64_t dW0 = smem_desc(smem_u32(sW), 128, 1024);
        const uint64_t dB2 = smem_desc(smem_u32(sB2), 128, 1024);
        const uint64_t dY0 = smem_desc(smem_u32(ybuf), L.kstr, 128);    // K-outer
        const uint64_t dA2 = smem_desc(smem_u32(sA2), L.kstr, 128);     // K-outer
        const uint64_t dB3 = smem_desc(smem_u32(sB3), 128, 1536);
        const uint32_t a_step = L.stage_bytes >> 4, half16 = L.half >> 4, w_step = (uint32_t)S * 128u;
        const uint32_t k16o = (2u * L.kstr) >> 4, plo = L.pstr >> 4;
        bool y_done = false;
        // q = dy_t W_o[:, own units] as soon as the dy operand is complete (between two dgh chunks if it lands early)
        auto try_y = [&](int n, bool block) -> bool {
            for (;;) {
                uint32_t ok = (lane == 0) ? (mbar_test_wait(y_full, (uint32_t)n & 1) ? 1u : 0u) : 0u;
                ok = __shfl_sync(0xffffffffu, ok, 0);
                if (ok) break;
                if (!block) return false;
                if (a.keepalive) mma_bf16_ss_elect(tmem + U2_COL_DUMMY, dW0, dB2, idesc2, false);
            }
            if (lane == 0) U2_TRACE(12);
            tc_fence_after();
#pragma unroll
            for (int k16 = 0; k16 < U2_KC / 16; ++k16) {
                mma_bf16_ss_elect(tmem + U2_COL_Q, dY0 + k16o * k16, dB2 + 16u * k16, idesc2s, k16 != 0);
                mma_bf16_ss_elect(tmem + U2_COL_Q, dY0 + plo + k16o * k16, dB2 + 16u * k16, idesc2, true);
            }
            mma_commit_elect(accum_full);
            return true;
        };
        int s = 0;
        uint32_t ph = 0;
        for (int n = 0; n <= T; ++n) {
            y_done = n == T;   // the last iteration (t = -1) has no dy
            if (n >= 1) {
                for (int ch = 0; ch < L.nch; ++ch) {
                    for (;;) {
                        uint32_t ok = (lane == 0) ? (mbar_test_wait(&full[s], ph) ? 1u : 0u) : 0u;
                        ok = __shfl_sync(0xffffffffu, ok, 0);
                        if (ok) break;
                        if (!y_done) y_done = try_y(n, false);
                        if (a.keepalive && !(a.dbg & 32))
                            mma_bf16_ss_elect(tmem + U2_COL_DUMMY, dW0, dB2, idesc2, false);
                    }
                    if (lane == 0 && ch < 8) U2_TRACE(40 + ch);
                    tc_fence_after();
                    const uint64_t da = dA0 + (uint64_t)((uint32_t)s * a_step);
                    const uint64_t db = dW0 + (uint64_t)((uint32_t)ch * w_step);
#pragma unroll
                    for (int k16 = 0; k16 < U2_KC / 16; ++k16) {
                        mma_bf16_ss_elect(tmem, da + 16u * k16, db + 16u * k16, idesc1s, (ch | k16) != 0);
                        mma_bf16_ss_elect(tmem, da + half16 + 16u * k16, db + 16u * k16, idesc1, true);
                    }
                    mma_commit_elect(&empty[s]);
                    if (ch == L.nch - 1) mma_commit_elect(d1_full);
                    if (lane == 0 && ch < 8) U2_TRACE(48 + ch);
                    if (++s == L.NS) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
            if (!y_done) try_y(n, true);
            if (n < T) {
                // cluster partial of the own output quarter: D3[b][o] = sum_{k < 96} dgi_cluster[b][k] W_y[k][16j + o]
                for (;;) {
                    uint32_t ok = (lane == 0) ? (mbar_test_wait(a2_full, (uint32_t)n & 1) ? 1u : 0u) : 0u;
                    ok = __shfl_sync(0xffffffffu, ok, 0);
                    if (ok) break;
                    if (a.keepalive) mma_bf16_ss_elect(tmem + U2_COL_DUMMY, dW0, dB2, idesc2, false);
                }
                if (lane == 0) U2_TRACE(13);
                tc_fence_after();
#pragma unroll
                for (int k16 = 0; k16 < U2_KB3 / 2; ++k16) {
                    mma_bf16_ss_elect(tmem + U2_COL_P, dA2 + k16o * k16, dB3 + 16u * k16, idesc3s, k16 != 0);
                    mma_bf16_ss_elect(tmem + U2_COL_P, dA2 + plo + k16o * k16, dB3 + 16u * k16, idesc3, true);
                }
                mma_commit_elect(part_full);
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ================= exchange + finalise: TMEM lane == batch row ===================================
        const int b = (warp - 4) * 32 + lane;
        const bool act = b < B;
        const bool wact = (warp - 4) * 32 < L.MB * 8;   // this warp holds rows of the staged row groups
        const int etid = threadIdx.x - 128;
        const uint32_t inbox_addr = smem_u32(inbox);
        const uint32_t inbox_bar_addr = smem_u32(inbox_full);
        const uint32_t taddr = tmem + ((uint32_t)((warp - 4) * 32) << 16);
        float carry[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) carry[q] = act ? f.dhc[(size_t)b * H + u0 + q] : 0.f;
        float bsum = 0.f;   // lane i: sum over (t, rows of this warp) of value i of [dar 8 | daz 8 | dan 8 | dan*r 8]
        const bool want_db = f.dbih != nullptr || f.dbhh != nullptr;
        float fx = 0.f;     // fixed-point scale of the totals, known once every CTA has finished its set-up scan
        // Saved activations of the step: every lane loads (the rows beyond B re-read row 0 of the frame), so the gate math is
        // branch-free.  Their frames were pulled into L2 two steps earlier by warp 8 (bulk prefetch, in the window in which no
        // CTA pulls dgh chunks): as HBM misses, issued by all CTAs while the chunks of the step are being pulled from L2, these
        // loads cost the chunk chain 2 400 cycles per step (measured by skipping them).
        float4 pr[2], pz[2], pn[2], pg[2], ph[2];
        float4 pm[2] = {make_float4(1.f, 1.f, 1.f, 1.f), make_float4(1.f, 1.f, 1.f, 1.f)};
        auto fetch_saved = [&](int tt) {
            const size_t so = ((size_t)tt * B + (act ? b : 0)) * H + u0;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                pr[q] = ldg_nc_v4_pinned(f.sv_r + so + 4 * q);
                pz[q] = ldg_nc_v4_pinned(f.sv_z + so + 4 * q);
                pn[q] = ldg_nc_v4_pinned(f.sv_n + so + 4 * q);
                pg[q] = ldg_nc_v4_pinned(f.sv_ghn + so + 4 * q);
                ph[q] = ldg_nc_v4_pinned(f.hs + so + 4 * q);   // hs slot t = h_{t-1}
                if (f.mask) pm[q] = ldg_nc_v4_pinned(f.mask + so + 4 * q);
            }
        };
        for (int n = 0; n <= T; ++n) {
            const int t = T - 1 - n;
            const size_t row = (size_t)(t < 0 ? 0 : t) * B + (act ? b : 0);
            if (t >= 0) fetch_saved(t);
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (etid == 0) U2_TRACE(0);
            if (n > 0) {
                if (etid == 0) mbar_expect_tx(inbox_full, (uint32_t)S * L.slot_bytes);
                mbar_wait(d1_full, (uint32_t)(n - 1) & 1);
                if (etid == 0) U2_TRACE(1);
                tc_fence_after();
                if (wact) {
#pragma unroll
                    for (int k = 0; k < U2_UB / 16; ++k) {
                        float v[16], v2[16];
                        tmem_ld_x16(taddr + 16 * k, v);             // A_hi B_hi + A_lo B_hi
                        tmem_ld_x16(taddr + U2_UB + 16 * k, v2);    // A_hi B_lo
                        tmem_ld_wait();
#pragma unroll
                        for (int q = 0; q < 16; ++q) v[q] += v2[q];
                        if (b < L.MB * 8) {
#pragma unroll
                            for (int h2 = 0; h2 < 2; ++h2) {
                                float* d = stage + (size_t)(2 * k + h2) * (L.slot_bytes / 4) + b * 8;
                                *reinterpret_cast<float4*>(d) = make_float4(v[8 * h2 + 0], v[8 * h2 + 1], v[8 * h2 + 2], v[8 * h2 + 3]);
                                *reinterpret_cast<float4*>(d + 4) = make_float4(v[8 * h2 + 4], v[8 * h2 + 5], v[8 * h2 + 6], v[8 * h2 + 7]);
                            }
                        }
                    }
                }
                fence_proxy_async_smem();
                named_bar_sync(3, 128);
                // partial sums of peer p's units -> slot j of p's inbox (bulk DSMEM copy, complete_tx on p's barrier)
                if (etid < S)
                    bulk_s2c(mapa(inbox_addr + (uint32_t)j * L.slot_bytes, (uint32_t)etid), stage + (size_t)etid * (L.slot_bytes / 4),
                             L.slot_bytes, mapa(inbox_bar_addr, (uint32_t)etid));
                if (etid == 0) U2_TRACE(2);
            }
            float qv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (t >= 0) {   // q = dy_t W_o[:, own units]
                mbar_wait(accum_full, (uint32_t)n & 1);
                if (etid == 0) U2_TRACE(4);
                tc_fence_after();
                float q2[8];
                tmem_ld_x8(taddr + U2_COL_Q, qv);
                tmem_ld_x8(taddr + U2_COL_Q + 16, q2);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q) qv[q] += q2[q];
            }
            tc_fence_before();
            if (n > 0) {
                mbar_wait_cluster(inbox_full, (uint32_t)(n - 1) & 1);
                if (etid == 0) U2_TRACE(3);
                {
                    const int bs = b < L.MB * 8 ? b : 0;   // rows beyond the staged row groups read row 0 (in bounds)
#pragma unroll
                    for (int p = 0; p < S; ++p) {
                        const float4 x0 = *reinterpret_cast<const float4*>(inbox + (size_t)p * (L.slot_bytes / 4) + bs * 8);
                        const float4 x1 = *reinterpret_cast<const float4*>(inbox + (size_t)p * (L.slot_bytes / 4) + bs * 8 + 4);
                        acc[0] += x0.x; acc[1] += x0.y; acc[2] += x0.z; acc[3] += x0.w;
                        acc[4] += x1.x; acc[5] += x1.y; acc[6] += x1.z; acc[7] += x1.w;
                    }
                }
            }
            if (t < 0) {
                if (act) {
                    float* d = f.dhc + (size_t)b * H + u0;
                    *reinterpret_cast<float4*>(d) = make_float4(carry[0] + acc[0], carry[1] + acc[1], carry[2] + acc[2], carry[3] + acc[3]);
                    *reinterpret_cast<float4*>(d + 4) = make_float4(carry[4] + acc[4], carry[5] + acc[5], carry[6] + acc[6], carry[7] + acc[7]);
                }
                break;
            }
            float dgr[8], dgz[8], dgn[8], dgnr[8];
            {
                const float* r_ = reinterpret_cast<const float*>(pr);
                const float* z_ = reinterpret_cast<const float*>(pz);
                const float* n_ = reinterpret_cast<const float*>(pn);
                const float* g_ = reinterpret_cast<const float*>(pg);
                const float* h_ = reinterpret_cast<const float*>(ph);
                const float* m_ = reinterpret_cast<const float*>(pm);
#pragma unroll
                for (int q = 0; q < 8; ++q) {   // straight-line for every lane (see gru_tc.cu: per-unit branch regions serialise the chains)
                    const float dh = carry[q] + acc[q] + qv[q] * m_[q];
                    const float r = r_[q], z = z_[q], nn = n_[q];
                    const float dn = dh * (1.0f - z);
                    const float dz = dh * (h_[q] - nn);
                    carry[q] = dh * z;
                    const float dan = dn * (1.0f - nn * nn);
                    dgr[q] = act ? dan * g_[q] * r * (1.0f - r) : 0.f;   // zeros on the rows beyond B: the bias sums below add every lane
                    dgz[q] = act ? dz * z * (1.0f - z) : 0.f;
                    dgn[q] = act ? dan : 0.f;
                    dgnr[q] = act ? dan * r : 0.f;
                }
            }
            if (etid == 0) U2_TRACE(10);
            uint4 hr, lr, hz, lz, hn, ln, hnr, lnr;
            u2_split8(dgr, hr, lr);
            u2_split8(dgz, hz, lz);
            u2_split8(dgnr, hnr, lnr);
            if (act) {
                // publish dgh_t = [dar, daz, dan*r] (bf16 hi/lo, UMMA order) for the next step's contraction
                uint16_t* dst = a.gxh + (size_t)(n & 1) * 2 * gx_part;
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    const int kidx = g * H + u0;
                    const size_t off = ((size_t)(kidx >> 6) * L.MB + (b >> 3)) * 512 + (size_t)((kidx & 63) >> 3) * 64 + (size_t)(b & 7) * 8;
                    *reinterpret_cast<uint4*>(dst + off) = (g == 0) ? hr : (g == 1) ? hz : hnr;
                    *reinterpret_cast<uint4*>(dst + gx_part + off) = (g == 0) ? lr : (g == 1) ? lz : lnr;
                }
            }
            fence_proxy_async_global();   // own generic writes of dgh_t -> visible to the peers' bulk copies (async proxy)
            named_bar_arrive(6, 160);     // warp 8 releases counter H
            if (etid == 0) U2_TRACE(7);
            u2_split8(dgn, hn, ln);
            if (act) {   // dgi of the own units: k blocks 3j (r), 3j+1 (z), 3j+2 (n) of the cluster's operand
                uint8_t* a2 = sA2 + (uint32_t)(3 * j) * L.kstr + (uint32_t)(b >> 3) * 128u + (uint32_t)(b & 7) * 16u;
                *reinterpret_cast<uint4*>(a2) = hr;
                *reinterpret_cast<uint4*>(a2 + L.pstr) = lr;
                *reinterpret_cast<uint4*>(a2 + L.kstr) = hz;
                *reinterpret_cast<uint4*>(a2 + L.kstr + L.pstr) = lz;
                *reinterpret_cast<uint4*>(a2 + 2 * L.kstr) = hn;
                *reinterpret_cast<uint4*>(a2 + 2 * L.kstr + L.pstr) = ln;
            }
            fence_proxy_async_smem();
            named_bar_sync(3, 128);
            if (etid < S - 1) {
                const uint32_t p = (uint32_t)((j + 1 + etid) & (S - 1));
                const uint32_t off = smem_u32(sA2) + (uint32_t)(3 * j) * L.kstr;
                bulk_s2c(mapa(off, p), sA2 + (size_t)(3 * j) * L.kstr, opiece, mapa(smem_u32(a2_full), p));
            } else if (etid == S - 1) {
                mbar_expect_tx(a2_full, (uint32_t)(S - 1) * opiece);
            }
            if (etid == 0) U2_TRACE(6);
            if (fx == 0.f) {   // first step only: every CTA's set-up scan has long finished
                if (etid == 0) spin_until_ge(ctrS, (unsigned)G, false);
                named_bar_sync(3, 128);
                float fi;
                u2_scale(__uint_as_float(*reinterpret_cast<volatile unsigned*>(gmax)), fx, fi);
            }
            // cluster partial of the own quarter -> fixed-point totals of slot n & 1 (lane == row: coalesced 8-byte adds)
            mbar_wait(part_full, (uint32_t)n & 1);
            if (etid == 0) U2_TRACE(25);
            tc_fence_after();
            {
                float v[16], v2[16];
                tmem_ld_x16(taddr + U2_COL_P, v);
                tmem_ld_x16(taddr + U2_COL_P + U2_OQ, v2);
                tmem_ld_wait();
                if (act) {
                    unsigned long long* d = a.yacc + (size_t)(n & 1) * yslot + (size_t)(U2_OQ * j) * RP + b;
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        if (U2_OQ * j + q < out && !(a.dbg & 1)) u2_red_add(d + (size_t)q * RP, (unsigned long long)__float2ll_rn((v[q] + v2[q]) * fx));
                }
            }
            tc_fence_before();
            if (etid == 0) U2_TRACE(26);
            named_bar_sync(7, 128);    // every finaliser has added its rows of the partial
            if (etid == 0) banked_arrive(ctrY, c);   // release is cumulative over the barrier: one gpu-scope fence per CTA
            if (etid == 0) U2_TRACE(9);
            if (act) {   // after both releases: only the products after the kernel read these
                float* gi = f.dgi + row * K3 + u0;
                *reinterpret_cast<float4*>(gi) = make_float4(dgr[0], dgr[1], dgr[2], dgr[3]);
                *reinterpret_cast<float4*>(gi + 4) = make_float4(dgr[4], dgr[5], dgr[6], dgr[7]);
                *reinterpret_cast<float4*>(gi + H) = make_float4(dgz[0], dgz[1], dgz[2], dgz[3]);
                *reinterpret_cast<float4*>(gi + H + 4) = make_float4(dgz[4], dgz[5], dgz[6], dgz[7]);
                *reinterpret_cast<float4*>(gi + 2 * H) = make_float4(dgn[0], dgn[1], dgn[2], dgn[3]);
                *reinterpret_cast<float4*>(gi + 2 * H + 4) = make_float4(dgn[4], dgn[5], dgn[6], dgn[7]);
                float* gn = f.dghn + row * H + u0;
                *reinterpret_cast<float4*>(gn) = make_float4(dgnr[0], dgnr[1], dgnr[2], dgnr[3]);
                *reinterpret_cast<float4*>(gn + 4) = make_float4(dgnr[4], dgnr[5], dgnr[6], dgnr[7]);
            }
            if (want_db) {   // bias gradients: butterfly sums over the rows of the warp, lane i keeps value i (fixed order)
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float v = (i < 8) ? dgr[i & 7] : (i < 16) ? dgz[i & 7] : (i < 24) ? dgn[i & 7] : dgnr[i & 7];
                    const float sres = warp_sum(v);   // inactive rows contribute zeros
                    if (lane == i) bsum += sres;
                }
            }
        }
        if (want_db) {
            // scratch = this CTA's own inbox: its last incoming copies were waited for and summed above
            named_bar_sync(3, 128);   // every finaliser thread is done reading the inbox
            inbox[(warp - 4) * 32 + lane] = bsum;
            named_bar_sync(3, 128);
            if (etid < 32) {
                const float tot = (inbox[etid] + inbox[32 + etid]) + (inbox[64 + etid] + inbox[96 + etid]);
                const int grp = etid >> 3, u = u0 + (etid & 7);
                if (f.dbih && grp < 3) {   // db_ih = sum [dar, daz, dan]
                    float* d = f.dbih + (size_t)grp * H + u;
                    *d = f.db_accumulate ? *d + tot : tot;
                }
                if (f.dbhh && grp != 2) {  // db_hh = sum [dar, daz, dan*r]
                    float* d = f.dbhh + (size_t)(grp == 3 ? 2 : grp) * H + u;
                    *d = f.db_accumulate ? *d + tot : tot;
                }
            }
        }
    } else if (warp >= 8) {
        // ================= aux: totals of the feedback + head's dY -> own quarter of the dy operand; w8 releases H =====
        const int rt = threadIdx.x - 256;          // batch row
        const bool y_act = rt < B;
        const int yo = U2_OQ * j;                  // first output of the quarter
        const unsigned long long* ysrc = a.yacc + (size_t)yo * RP + (y_act ? rt : 0);
        unsigned long long p0[16], p1[16];         // the totals this thread saw two steps / one step ago (slot parity)
#pragma unroll
        for (int e = 0; e < 16; ++e) p0[e] = p1[e] = 0ull;
        float fx_inv = 0.f;
        // Every cluster reads the head's dY of slot t+1 at iteration n, and cluster 0 overwrites that slot with the total: the
        // store is deferred by one iteration (behind the wait for counter Y, i.e. after every CTA's gates of iteration n, which
        // consumed the dy operand built from those reads).  Slot 0 (t = -1) is only read by cluster 0 itself.
        float prevv[16];
        bool have_prev = false;
        for (int n = 0; n <= T; ++n) {
            const int t = T - 1 - n;
            float* dyt = f.dy_tot + (size_t)(t + 1) * n_pairs;
            // the head's dY of this thread's row: fetched before the wait
            float dyv[16];
            const bool need_dy = y_act && (n < T || ci == 0);
#pragma unroll
            for (int e = 0; e < 16; ++e) dyv[e] = (need_dy && yo + e < out) ? __ldcg(dyt + (size_t)rt * out + yo + e) : 0.f;
            if (n > 0) {
                if (warp == 8) {
                    banked_wait_warp(ctrY, per_bank * (unsigned)n, lane, a.relaxed != 0);
                    if (n == 1 && lane == 0) spin_until_ge(ctrS, (unsigned)G, false);
                    if (lane == 0) U2_TRACE(20);
                }
                named_bar_sync(5, 128);
                if (n == 1) {
                    float fxx;
                    u2_scale(__uint_as_float(*reinterpret_cast<volatile unsigned*>(gmax)), fxx, fx_inv);
                }
                const unsigned long long* src = ysrc + (size_t)((n + 1) & 1) * yslot;   // slot of iteration n-1
                unsigned long long cur[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) cur[e] = u2_ld_total(src + (size_t)e * RP);
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    dyv[e] = fmaf(__ll2float_rn((long long)(cur[e] - p0[e])), fx_inv, dyv[e]);
                    p0[e] = p1[e];
                    p1[e] = cur[e];
                }
                if (rt == 0) U2_TRACE(27);
                if (have_prev && y_act) {   // the total of iteration n-1 (slot t+2)
                    float* d = f.dy_tot + (size_t)(t + 2) * n_pairs + (size_t)rt * out + yo;
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (yo + e < out) d[e] = prevv[e];
                }
            }
            if (n > 0 && ci == 0) {
#pragma unroll
                for (int e = 0; e < 16; ++e) prevv[e] = dyv[e];
                have_prev = true;
            }
            if (t >= 0) {
                if (y_act) {   // two core-matrix rows: this CTA's two k blocks of the dy operand
                    uint4 hi, lo;
                    uint8_t* d = ybuf + (uint32_t)(2 * j) * L.kstr + (uint32_t)(rt >> 3) * 128u + (uint32_t)(rt & 7) * 16u;
                    u2_split8(dyv, hi, lo);
                    *reinterpret_cast<uint4*>(d) = hi;
                    *reinterpret_cast<uint4*>(d + L.pstr) = lo;
                    u2_split8(dyv + 8, hi, lo);
                    *reinterpret_cast<uint4*>(d + L.kstr) = hi;
                    *reinterpret_cast<uint4*>(d + L.kstr + L.pstr) = lo;
                }
                fence_proxy_async_smem();
                named_bar_sync(5, 128);
                if (rt < S - 1) {
                    const uint32_t p = (uint32_t)((j + 1 + rt) & (S - 1));
                    const uint32_t off = smem_u32(ybuf) + (uint32_t)(2 * j) * L.kstr;
                    bulk_s2c(mapa(off, p), ybuf + (size_t)(2 * j) * L.kstr, ypiece, mapa(smem_u32(y_full), p));
                } else if (rt == S - 1) {
                    mbar_expect_tx(y_full, (uint32_t)(S - 1) * ypiece);
                    U2_TRACE(21);
                }
                if (warp == 8) {   // counter H: the finalisers arrive (without waiting) once dgh_t is published and fenced
                    named_bar_sync(6, 160);
                    if (lane == 0) banked_arrive(ctrH, c);   // release is cumulative over the barrier
                    if (lane == 0) U2_TRACE(22);
                    // the saved activations of frame t-2 -> L2: this CTA's 1/G of each frame (every CTA reads 32 bytes of every row),
                    // issued now: counter H is not complete yet, so no CTA pulls dgh chunks
                    if (t >= 2 && lane < 6 && !(a.dbg & 2)) {
                        const float* base = lane == 0 ? f.sv_r : lane == 1 ? f.sv_z : lane == 2 ? f.sv_n : lane == 3 ? f.sv_ghn : lane == 4 ? f.hs : f.mask;
                        const size_t frame = (size_t)B * H, slice = frame / (size_t)G;   // B * 8 floats: a multiple of 16 bytes
                        if (base) bulk_prefetch_l2(base + (size_t)(t - 2) * frame + (size_t)c * slice, (uint32_t)(slice * sizeof(float)));
                    }
                }
            }
            if (t < 0) {
                if (n > 0 && y_act && ci == 0) {   // dy_in (slot 0)
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (yo + e < out) dyt[(size_t)rt * out + yo + e] = dyv[e];
                }
                break;
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
static bool u2_shape_ok(int B, int H, int out) {
    return H % (U2_KC * U2_S) == 0 && H >= U2_KC * U2_S && (3 * H / U2_KC) % U2_S == 0 && out >= 1 && out <= 64 && B >= 1 && B <= 128;
}

size_t gru_tc2_bwd_scratch_floats(int B, int H) {
    const size_t MB = (B + 7) / 8;
    const size_t gxh = (size_t)2 * 2 * (3 * H / U2_KC) * MB * 512 / 2;   // bf16 elements -> floats
    const size_t yacc = (size_t)2 * 64 * MB * 8 * 2;                    // u64 -> floats
    return round_up_sz(gxh, 64) + 1024 + round_up_sz(yacc, 64);
}

static bool u2_runnable(int B, int H, int out, const DeviceInfo& di, U2Layout* Lout) {
    const int G = H / 8;
    if (!u2_shape_ok(B, H, out) || G > di.n_sm) return false;
    struct Entry { int B, H, out, ok; };
    static Entry cache[256];
    static int n_cache = 0;
    int ok = -1;
    for (int i = 0; i < n_cache; ++i)
        if (cache[i].B == B && cache[i].H == H && cache[i].out == out) ok = cache[i].ok;
    U2Layout L = u2_layout(B, H, di.max_smem_optin);
    if (ok < 0) {
        ok = 0;
        if (L.NS >= 2 && (int)L.total <= di.max_smem_optin && (uint32_t)L.NS * L.stage_bytes >= (uint32_t)U2_S * L.slot_bytes &&
            cudaFuncSetAttribute(k_gru_bwd_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total) == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(G);
            cfg.blockDim = dim3(U2_NT);
            cfg.dynamicSmemBytes = L.total;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = U2_S;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int ncl = 0;
            if (cudaOccupancyMaxActiveClusters(&ncl, k_gru_bwd_tc2, &cfg) == cudaSuccess) {
                if (getenv("CVB_DEBUG"))
                    fprintf(stderr, "[cvb] k_gru_bwd_tc2: %d co-resident clusters of %d (need %d), smem %u, ring %d\n", ncl, U2_S, G / U2_S, L.total, L.NS);
                ok = ncl * U2_S >= G ? 1 : 0;
            }
        }
        cudaGetLastError();
        if (n_cache < 256) cache[n_cache++] = Entry{B, H, out, ok};
    }
    if (ok && Lout) *Lout = L;
    return ok != 0;
}

bool gru_tc2_bwd_supported(int B, int H, int out, const DeviceInfo& di) { return u2_runnable(B, H, out, di, nullptr); }

int gru_ar_bwd_tc2(GruBwdArgs& f, float* tc_scratch, cudaStream_t s) {
    if (f.T <= 0 || f.B <= 0) return 0;
    DeviceInfo di;
    if (int rc = get_device_info(&di)) return rc;
    U2Layout L;
    CVB_REQUIRE(u2_runnable(f.B, f.H, f.out, di, &L), "gru_ar_bwd_tc2: unsupported shape B=%d H=%d out=%d", f.B, f.H, f.out);
    GruTc2BwdArgs a;
    a.f = f;
    const size_t gxh_f = round_up_sz((size_t)2 * 2 * (3 * f.H / U2_KC) * L.MB * 512 / 2, 64);
    const size_t yacc_f = (size_t)2 * 64 * L.MB * 8 * 2;
    a.gxh = reinterpret_cast<uint16_t*>(tc_scratch);
    a.ctr = reinterpret_cast<unsigned*>(tc_scratch + gxh_f);
    a.yacc = reinterpret_cast<unsigned long long*>(tc_scratch + gxh_f + 1024);
    a.smem_max = di.max_smem_optin;
    a.trace = nullptr;
    a.keepalive = 1;
    a.relaxed = relaxed_polling() ? 1 : 0;
    a.dbg = 0;
    if (const char* e = getenv("CVB_TC_DBG")) a.dbg = atoi(e);
    if (const char* e = getenv("CVB_TC_KEEPALIVE")) a.keepalive = atoi(e) != 0;
    const char* trace_file = getenv("CVB_TRACE_FILE");
    const size_t trace_bytes = (size_t)(f.T + 1) * 64 * sizeof(long long);
    if (trace_file && trace_file[0]) {
        CVB_CHECK(cudaMalloc(&a.trace, trace_bytes));
        CVB_CHECK(cudaMemsetAsync(a.trace, 0, trace_bytes, s));
    }
    CVB_CHECK(cudaMemsetAsync(a.ctr, 0, (1024 + yacc_f) * sizeof(float), s));   // counter banks, maximum and the totals
    CVB_CHECK(cudaFuncSetAttribute(k_gru_bwd_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(f.H / 8);
    cfg.blockDim = dim3(U2_NT);
    cfg.dynamicSmemBytes = L.total;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = U2_S;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = launch_without_coop() ? 1 : 2;
    prof_begin(s, CVB_PROF_GRU_BWD);
    CVB_CHECK(cudaLaunchKernelEx(&cfg, k_gru_bwd_tc2, a));
    prof_end(s, CVB_PROF_GRU_BWD);
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
