// Tensor-core (tcgen05) BPTT of the autoregressive GRU (the reverse of
// gru_vae.py:364-399), one persistent cooperative launch for all T steps.
//
// 2-D work split over thread-block clusters of S CTAs (S = 8 where 16 such clusters are co-resident, else 4):
// cluster i owns the block of 8*S hidden units [8*S*i, 8*S*(i+1)); CTA j of the cluster owns the K-slice
// [j*3H/S, (j+1)*3H/S) of the contraction  dh[b,u] += sum_k dgh_{t+1}[b,k] W_hh[k,u]  for ALL units of the
// block, and FINALISES the 8 units [8*S*i + 8j, +8).  W_hh^T of (block x K-slice) stays in shared memory for the
// whole sequence as bf16 hi+lo (x = hi + lo + O(2^-17 x); three MMAs hi*hi + lo*hi + hi*lo, fp32 accumulate in
// TMEM), so a CTA ingests only 1/S of the all-gathered dgh_{t+1} per step.  The S partial accumulators of a unit
// meet in the finaliser's shared memory through bulk DSMEM copies (cp.async.bulk.shared::cluster, complete_tx on
// the receiver's mbarrier) and are summed in fixed order (deterministic).
//
// The y feedback (dy_t = dY_t + dgi_{t+1} W_y, then dh_t += (dy_t W_o) * m_t) is a reduction over all of 3H.  Both
// of its dense pieces run on the tensor core as well: the finaliser stages dgi of its 8 units as a [128 x 32] bf16
// operand and one MMA chain forms its partial [B,out] (accumulator D3); the per-pair sums over the CTAs are formed
// in fixed order by the "aux" warps of the CTA that owns the pair and published in operand order; every CTA then
// pulls dy_t through its ring and a second MMA chain forms q = dy_t W_o[:, own units] (accumulator D2).
//
// Roles (384 threads): w0 bulk-copy producer, w1 MMA issuer, w2 TMEM allocator, w4-7 exchange + finalise
// (TMEM lane == batch row), w8-11 aux (dy reduction, drain of D3).  Grid-wide ordering is two monotonic
// counters: A (dgh_t / partials published), B (dy_t published).
#include <stdlib.h>

#include "gru_ar.cuh"
#include "umma.cuh"

namespace cvb {
using namespace umma;

constexpr int TB_NT = 384;
constexpr int TB_KC = 64;        // K per ring stage
// TMEM columns.  Every B operand is stored [hi rows | lo rows], so ONE MMA with N = 2 x rows forms A_hi*B_hi (first
// half of the columns) and A_hi*B_lo (second half); a second MMA with N = rows adds A_lo*B_hi to the first half.
// The halves are added in fp32 registers by whoever drains the accumulator (2 instead of 3 MMAs per K step, and
// the small correction term never shares an accumulator chain with the large one).
constexpr uint32_t TB_COL_Q = 128;     // D2 (q): 2 x 16 columns
constexpr uint32_t TB_COL_DUMMY = 160; // keep-alive scratch, 16 columns
constexpr uint32_t TB_COL_P = 256;     // D3 (partial of the y feedback): 2 x 64 columns

struct TbLayout {
    int MB, S, Ublk, nch, NS;
    uint32_t half, stage_bytes, w_part_bytes, slot_bytes;
    int Q;   // pairs (b, o) per reducer CTA, a multiple of 8 (PartWalk, umma.cuh)
    uint32_t off_ring, off_w, off_inbox, off_stage, off_a2, off_b2, off_b3, off_aux, off_bar, total;
};

__host__ __device__ inline TbLayout tb_layout(int B, int H, int S, int G, int out, int smem_max) {
    TbLayout L;
    L.MB = (B + 7) / 8;
    L.S = S;
    L.Ublk = 8 * S;
    L.nch = 3 * H / TB_KC / S;
    L.half = (uint32_t)L.MB * 1024u;
    L.stage_bytes = 2u * L.half;
    L.w_part_bytes = (uint32_t)L.nch * (uint32_t)S * 1024u;
    L.slot_bytes = (uint32_t)L.MB * 8u * 32u;
    const uint32_t inbox = (uint32_t)S * L.slot_bytes;
    L.Q = 8 * ((8 * B + G - 1) / G);
    const uint32_t aux = (((uint32_t)(G * L.Q) * 4u) + 127u) & ~127u;   // a reducer's block of partials [G CTAs][Q pairs] (one bulk copy)
    const uint32_t sps = (uint32_t)(4 * L.Q) * 4u;                      // [4 warp groups][Q] partial sums of the reducers
    const uint32_t fixed = 2u * L.w_part_bytes + 2u * inbox + 16384u + 4096u + 8192u + aux + 256u + sps;
    int ns = ((int)smem_max - (int)fixed) / (int)L.stage_bytes;
    L.NS = ns > 6 ? 6 : ns;
    const uint32_t ring = (uint32_t)(L.NS > 0 ? L.NS : 0) * L.stage_bytes;
    L.off_ring = 0;
    L.off_w = ring;
    L.off_inbox = L.off_w + 2u * L.w_part_bytes;
    L.off_stage = L.off_inbox + inbox;
    L.off_a2 = L.off_stage + inbox;     // [2 parts][16 row groups][4 kblk][8][8] bf16: dgi of the own units
    L.off_b2 = L.off_a2 + 16384u;       // [2 parts][2 n blocks][8 kblk][8][8] bf16: W_o^T (own units)
    L.off_b3 = L.off_b2 + 4096u;        // [2 parts][8 n blocks][4 kblk][8][8] bf16: W_y rows of the own units
    L.off_aux = L.off_b3 + 8192u;
    L.off_bar = L.off_aux + aux;
    L.total = L.off_bar + 256u + sps;    // mbarriers + tmem slot | partial sums of the reducers
    return L;
}

struct GruTcBwdArgs {
    GruBwdArgs f;
    uint16_t* gxh;    // [2 slots][2 parts][3H/64 chunks][MB][8 kblk][8 rows][8 k] bf16 (UMMA order) of dgh_t
    uint16_t* dyx;    // [2 slots][2 parts][MB][8 kblk][8 rows][8 k] bf16 (UMMA order) of dy_t, zero-initialised
    float* part;      // [G reducers][G CTAs][Q] partial sums of the y feedback (PartWalk)
    unsigned* ctr;    // [0] = A, [32] = B, [64] = H (separate 128-B lines), zero-initialised
    int early_h;      // CVB_TC_BWD_EARLYH=1: dgh_t gets its own arrival counter H, released before the partial path of the step
    int S;
    int smem_max;
    int keepalive;      // CVB_TC_KEEPALIVE (default 1): dummy MMAs while the issuer polls
    int relaxed;        // CVB_TC_POLL=relaxed: poll the arrival counters with relaxed loads + one acquire fence instead of ld.acquire
    long long* trace;   // optional [T+1][64] clock64 stamps of CTA 0 (CVB_TRACE_FILE), else null
};

// per-step phase stamps of one CTA (profiling hook; one predictable branch when disabled)
#define TB_TRACE(ev)                                                     \
    do {                                                                 \
        if (a.trace && c == 0) a.trace[(size_t)n * 64 + (ev)] = clock64(); \
    } while (0)

static __device__ __forceinline__ uint4 pack_bf16x8(const uint16_t* v) {
    return make_uint4((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16),
                      (uint32_t)v[4] | ((uint32_t)v[5] << 16), (uint32_t)v[6] | ((uint32_t)v[7] << 16));
}
static __device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
    uint16_t h[8], l[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) split_bf16(x[q], h[q], l[q]);
    hi = pack_bf16x8(h);
    lo = pack_bf16x8(l);
}

// drain 32 columns (outputs [o_lo, o_lo + 32)) of D3 (hi + correction halves) of this warp's 32 TMEM lanes into `part`
// (layout and walk: PartWalk, umma.cuh)
static __device__ __forceinline__ void drain_partial(uint32_t taddr_p, const PartWalk& w, int o_lo, bool row_ok) {
    unsigned long long addr = w.addr;
    int i = w.i;
#pragma unroll
    for (int o0 = o_lo; o0 < o_lo + 32; o0 += 16) {
        float v[16], v2[16];
        tmem_ld_x16(taddr_p + o0, v);
        tmem_ld_x16(taddr_p + 64 + o0, v2);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; q += 4) {
            if (row_ok) st_global_v4(addr, v[q] + v2[q], v[q + 1] + v2[q + 1], v[q + 2] + v2[q + 2], v[q + 3] + v2[q + 3]);
            addr += 16;
            i += 4;
            if (i >= w.Q) {
                i -= w.Q;
                addr += w.wrap;
            }
        }
    }
}

__global__ void __launch_bounds__(TB_NT, 1) k_gru_bwd_tc(GruTcBwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const GruBwdArgs& f = a.f;
    const int B = f.B, T = f.T, H = f.H, out = f.out, K3 = 3 * f.H;
    const int G = gridDim.x, c = blockIdx.x, S = a.S;
    const int j = (int)cluster_ctarank();
    const TbLayout L = tb_layout(B, H, S, G, out, a.smem_max);
    const int ublk0 = (c / S) * L.Ublk;   // first unit of the cluster's block
    const int u0 = ublk0 + 8 * j;         // first of the 8 units this CTA finalises
    const int k0 = j * L.nch * TB_KC;     // first row of W_hh (= column of dgh) of this CTA's K-slice
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
    const int lane = threadIdx.x & 31;
    uint8_t* ring = smem + L.off_ring;
    uint8_t* sW = smem + L.off_w;
    float* inbox = reinterpret_cast<float*>(smem + L.off_inbox);   // [S (from)][MB*8][8]
    float* stage = reinterpret_cast<float*>(smem + L.off_stage);   // [S (to)][MB*8][8]
    uint8_t* sA2 = smem + L.off_a2;
    uint8_t* sB2 = smem + L.off_b2;
    uint8_t* sB3 = smem + L.off_b3;
    float* sRed = reinterpret_cast<float*>(smem + L.off_aux);      // [G][w]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* empty = full + 8;
    uint64_t* accum_full = empty + 8;
    uint64_t* inbox_full = accum_full + 1;
    uint64_t* a2_full = inbox_full + 1;
    uint64_t* part_full = a2_full + 1;
    uint64_t* red_full = part_full + 1;   // the reducers' block of partials has landed in sRed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(red_full + 1);
    const size_t gx_part = (size_t)(K3 / TB_KC) * L.MB * 512;   // elements per part
    const size_t dy_part = (size_t)L.MB * 512;
    unsigned* ctrA = a.ctr;
    unsigned* ctrB = a.ctr + 32;
    unsigned* ctrH = a.ctr + 64;
    const int n_pairs = B * out;

    // ---- one-time setup --------------------------------------------------------------------------
    {
        // one item = an 8 (k) x 4 (n) block of W_hh^T: 8 float4 reads along the contiguous unit axis, transposed in
        // registers into four 16-byte core-matrix rows (hi) + four (lo)
        const int nq = L.Ublk >> 2;
        const int n_items = nq * L.nch * (TB_KC / 8);
        for (int i = threadIdx.x; i < n_items; i += TB_NT) {
            const int kg = i / nq, n4 = (i - kg * nq) * 4;
            const int kl = kg * 8;
            float4 r[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) r[q] = __ldg(reinterpret_cast<const float4*>(f.Whh + (size_t)(k0 + kl + q) * H + ublk0 + n4));
            const uint32_t off = (uint32_t)(kl / TB_KC) * ((uint32_t)S * 2048u) + (uint32_t)((kl % TB_KC) >> 3) * 128u;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float w[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) w[q] = e == 0 ? r[q].x : e == 1 ? r[q].y : e == 2 ? r[q].z : r[q].w;
                uint4 hi, lo;
                split8(w, hi, lo);
                const int n = n4 + e;
                const uint32_t o2 = off + (uint32_t)(n >> 3) * 1024u + (uint32_t)(n & 7) * 16u;
                *reinterpret_cast<uint4*>(sW + o2) = hi;                       // chunk layout: [hi: S blocks][lo: S blocks]
                *reinterpret_cast<uint4*>(sW + (uint32_t)S * 1024u + o2) = lo;
            }
        }
        {   // B2[n][k] = W_o[k][u0 + n];  B3[n = o][k = g*8+uu] = W_y[g*H + u0 + uu][o]   (all loads in flight before the first conversion)
            constexpr int M2 = (16 * 64 + TB_NT - 1) / TB_NT, M3 = (64 * 32 + TB_NT - 1) / TB_NT;
            float w2[M2], w3[M3];
#pragma unroll
            for (int m = 0; m < M2; ++m) {
                const int i = threadIdx.x + m * TB_NT, n = i >> 6, k = i & 63;
                w2[m] = (i < 16 * 64 && n < 8 && k < out) ? __ldg(f.Wo + (size_t)k * H + u0 + n) : 0.f;
            }
#pragma unroll
            for (int m = 0; m < M3; ++m) {
                const int i = threadIdx.x + m * TB_NT, k = i >> 6, n = i & 63;
                w3[m] = (i < 64 * 32 && k < 24 && n < out) ? __ldg(f.Wy + (size_t)((k >> 3) * H + u0 + (k & 7)) * f.ldwy + n) : 0.f;
            }
#pragma unroll
            for (int m = 0; m < M2; ++m) {
                const int i = threadIdx.x + m * TB_NT, n = i >> 6, k = i & 63;
                if (i < 16 * 64) {
                    uint16_t hi, lo;
                    split_bf16(w2[m], hi, lo);
                    const uint32_t off = (uint32_t)(n >> 3) * 1024u + (uint32_t)(k >> 3) * 128u + (uint32_t)(n & 7) * 16u + (uint32_t)(k & 7) * 2u;
                    *reinterpret_cast<uint16_t*>(sB2 + off) = hi;
                    *reinterpret_cast<uint16_t*>(sB2 + 2048 + off) = lo;
                }
            }
#pragma unroll
            for (int m = 0; m < M3; ++m) {
                const int i = threadIdx.x + m * TB_NT, k = i >> 6, n = i & 63;
                if (i < 64 * 32) {
                    uint16_t hi, lo;
                    split_bf16(w3[m], hi, lo);
                    const uint32_t off = (uint32_t)(n >> 3) * 512u + (uint32_t)(k >> 3) * 128u + (uint32_t)(n & 7) * 16u + (uint32_t)(k & 7) * 2u;
                    *reinterpret_cast<uint16_t*>(sB3 + off) = hi;
                    *reinterpret_cast<uint16_t*>(sB3 + 4096 + off) = lo;
                }
            }
        }
        for (int i = threadIdx.x; i < 16384 / 16; i += TB_NT) reinterpret_cast<uint4*>(sA2)[i] = make_uint4(0u, 0u, 0u, 0u);
        if (threadIdx.x == 0) {
            for (int s = 0; s < 8; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            mbar_init(accum_full, 1);
            mbar_init(inbox_full, 1);   // armed with expect_tx(S slots) every step; the peers' bulk copies complete_tx
            mbar_init(a2_full, 128);    // every finaliser thread has staged its dgi row
            mbar_init(part_full, 1);
            mbar_init(red_full, 1);
            mbar_fence_init();
        }
        fence_proxy_async_smem();
        if (warp == 2) tmem_alloc<512>(tmem_slot);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    cluster_sync_all();   // every CTA's inbox barrier is initialised before any peer copies into it
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ================= producer: K-slice of dgh_{t+1} chunk by chunk, then dy_t ====================
        int s = 0;
        uint32_t ph = 1;   // parity to wait on the empty barrier of stage s (first pass: free)
        for (int n = 0; n <= T; ++n) {
            const int nchunks = (n >= 1 ? L.nch : 0) + (n < T ? 1 : 0);
            const uint16_t* src = a.gxh + (size_t)((n - 1) & 1) * 2 * gx_part + (size_t)(j * L.nch) * L.MB * 512;
            const uint16_t* srcy = a.dyx + (size_t)(n & 1) * 2 * dy_part;
            if (n >= 1) {
                if (lane == 0) {
                    spin_until_ge(a.early_h ? ctrH : ctrA, (unsigned)G * (unsigned)n, a.relaxed != 0);
                        TB_TRACE(14);
                }
                __syncwarp();
            }
            for (int ch = 0; ch < nchunks; ++ch) {
                const bool is_dy = (ch == nchunks - 1) && (n < T);
                if (lane == 0) {
                    if (is_dy) {
                        spin_until_ge(ctrB, (unsigned)G * (unsigned)(n + 1), a.relaxed != 0);
                        TB_TRACE(13);
                    }
                    mbar_wait(&empty[s], ph);
                    if (ch < 8) TB_TRACE(32 + ch);
                    uint8_t* dst = ring + (size_t)s * L.stage_bytes;
                    mbar_expect_tx(&full[s], 2 * L.half);
                    const uint16_t* p0 = is_dy ? srcy : src + (size_t)ch * L.MB * 512;
                    const uint16_t* p1 = is_dy ? srcy + dy_part : src + gx_part + (size_t)ch * L.MB * 512;
                    bulk_g2s(dst, p0, L.half, &full[s]);
                    bulk_g2s(dst + L.half, p1, L.half, &full[s]);
                }
                __syncwarp();
                if (++s == L.NS) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (descriptors stay warp-uniform; lane 0 issues) =====================
        const uint32_t idesc1 = idesc_bf16_f32(128, L.Ublk), idesc1s = idesc_bf16_f32(128, 2 * L.Ublk);
        const uint32_t idesc2 = idesc_bf16_f32(128, 16), idesc2s = idesc_bf16_f32(128, 32);
        const uint32_t idesc3 = idesc_bf16_f32(128, 64), idesc3s = idesc_bf16_f32(128, 128);
        const uint64_t dA0 = smem_desc(smem_u32(ring), 128, 1024);
        const uint64_t dW0 = smem_desc(smem_u32(sW), 128, 1024);
        const uint64_t dB2 = smem_desc(smem_u32(sB2), 128, 1024);
        const uint64_t dA2 = smem_desc(smem_u32(sA2), 128, 512);
        const uint64_t dB3 = smem_desc(smem_u32(sB3), 128, 512);
        const uint32_t a_step = L.stage_bytes >> 4, half16 = L.half >> 4, w_step = (uint32_t)S * 128u;
        // poll an mbarrier; while idle keep the tensor pipe warm with a dummy MMA into scratch columns (the first MMA
        // after a few microseconds of idleness was measured to stall ~3400 cycles at issue)
        auto wait_warm = [&](uint64_t* bar, uint32_t parity) {
            for (;;) {
                uint32_t ok = (lane == 0) ? (mbar_test_wait(bar, parity) ? 1u : 0u) : 0u;
                ok = __shfl_sync(0xffffffffu, ok, 0);
                if (ok) break;
                if (a.keepalive) mma_bf16_ss_elect(tmem + TB_COL_DUMMY, dA2, dB2, idesc2, false);
            }
        };
        int s = 0;
        uint32_t ph = 0;
        for (int n = 0; n <= T; ++n) {
            const int nchunks = (n >= 1 ? L.nch : 0) + (n < T ? 1 : 0);
            for (int ch = 0; ch < nchunks; ++ch) {
                const bool is_dy = (ch == nchunks - 1) && (n < T);
                wait_warm(&full[s], ph);
                if (lane == 0 && ch < 8) TB_TRACE(40 + ch);
                tc_fence_after();
                const uint64_t da = dA0 + (uint64_t)((uint32_t)s * a_step);
                const uint64_t db = is_dy ? dB2 : dW0 + (uint64_t)((uint32_t)ch * w_step);
                const uint32_t idesc_s = is_dy ? idesc2s : idesc1s;   // [hi | lo] rows of B
                const uint32_t idesc_h = is_dy ? idesc2 : idesc1;     // hi rows only
                const uint32_t d_tmem = is_dy ? tmem + TB_COL_Q : tmem;
                const bool first = is_dy || ch == 0;
#pragma unroll
                for (int k16 = 0; k16 < TB_KC / 16; ++k16) {
                    mma_bf16_ss_elect(d_tmem, da + 16u * k16, db + 16u * k16, idesc_s, !(first && k16 == 0));
                    mma_bf16_ss_elect(d_tmem, da + half16 + 16u * k16, db + 16u * k16, idesc_h, true);
                }
                mma_commit_elect(&empty[s]);
                if (ch == nchunks - 1) mma_commit_elect(accum_full);
                if (lane == 0 && ch < 8) TB_TRACE(48 + ch);
                if (++s == L.NS) {
                    s = 0;
                    ph ^= 1;
                }
            }
            if (n < T) {
                // partial of the y feedback of this step: D3[b][o] = sum_k dgi_own[b][k] W_y[k][o]
                wait_warm(a2_full, (uint32_t)n & 1);
                tc_fence_after();
#pragma unroll
                for (int k16 = 0; k16 < 2; ++k16) {
                    mma_bf16_ss_elect(tmem + TB_COL_P, dA2 + 16u * k16, dB3 + 16u * k16, idesc3s, k16 != 0);
                    mma_bf16_ss_elect(tmem + TB_COL_P, dA2 + 512u + 16u * k16, dB3 + 16u * k16, idesc3, true);
                }
                mma_commit_elect(part_full);
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ================= exchange + finalise: TMEM lane == batch row ===================================
        const int b = (warp - 4) * 32 + lane;
        const bool act = b < B;
        const int etid = threadIdx.x - 128;
        const uint32_t inbox_addr = smem_u32(inbox);
        const uint32_t inbox_bar_addr = smem_u32(inbox_full);
        const uint32_t taddr = tmem + ((uint32_t)((warp - 4) * 32) << 16);
        float carry[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) carry[q] = act ? f.dhc[(size_t)b * H + u0 + q] : 0.f;
        const PartWalk pw = part_walk(a.part, c, G, L.Q, b, 32);   // this thread drains outputs [32, 64) of its row
        float bsum = 0.f;   // lane i: sum over (t, rows of this warp) of value i of [dar 8 | daz 8 | dan 8 | dan*r 8]
        const bool want_db = f.dbih != nullptr || f.dbhh != nullptr;
        for (int n = 0; n <= T; ++n) {
            const int t = T - 1 - n;
            const size_t row = (size_t)(t < 0 ? 0 : t) * B + (act ? b : 0);
            float4 pr[2], pz[2], pn[2], pg[2], ph[2];
            float4 pm[2] = {make_float4(1.f, 1.f, 1.f, 1.f), make_float4(1.f, 1.f, 1.f, 1.f)};
            if (t >= 0) {   // every lane loads (the rows beyond B re-read row 0 of the frame: `row`), so the gate math below is branch-free
                const size_t so = row * H + u0;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    pr[q] = ldg_nc_v4_pinned(f.sv_r + so + 4 * q);
                    pz[q] = ldg_nc_v4_pinned(f.sv_z + so + 4 * q);
                    pn[q] = ldg_nc_v4_pinned(f.sv_n + so + 4 * q);
                    pg[q] = ldg_nc_v4_pinned(f.sv_ghn + so + 4 * q);
                    ph[q] = ldg_nc_v4_pinned(f.hs + so + 4 * q);   // hs slot t = h_{t-1}
                    if (f.mask) pm[q] = ldg_nc_v4_pinned(f.mask + so + 4 * q);
                }
            }
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (etid == 0) TB_TRACE(0);
            if (n > 0 && etid == 0) mbar_expect_tx(inbox_full, (uint32_t)S * L.slot_bytes);
            mbar_wait(accum_full, (uint32_t)n & 1);
            if (etid == 0) TB_TRACE(1);
            tc_fence_after();
            if (n > 0) {
                for (int k = 0; k < L.Ublk / 16; ++k) {
                    float v[16], v2[16];
                    tmem_ld_x16(taddr + 16 * k, v);             // A_hi B_hi + A_lo B_hi
                    tmem_ld_x16(taddr + L.Ublk + 16 * k, v2);   // A_hi B_lo
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
                fence_proxy_async_smem();
                named_bar_sync(3, 128);
                // partial sums of peer p's units -> slot j of p's inbox (bulk DSMEM copy, complete_tx on p's barrier)
                if (etid < S)
                    bulk_s2c(mapa(inbox_addr + (uint32_t)j * L.slot_bytes, (uint32_t)etid), stage + (size_t)etid * (L.slot_bytes / 4),
                             L.slot_bytes, mapa(inbox_bar_addr, (uint32_t)etid));
                if (etid == 0) TB_TRACE(2);
            }
            float qv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (t >= 0) {   // q = dy_t W_o[:, own units]
                float q2[8];
                tmem_ld_x8(taddr + TB_COL_Q, qv);
                tmem_ld_x8(taddr + TB_COL_Q + 16, q2);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q) qv[q] += q2[q];
            }
            tc_fence_before();
            if (n > 0) {
                mbar_wait_cluster(inbox_full, (uint32_t)(n - 1) & 1);
                if (etid == 0) TB_TRACE(3);
                {
                    const int bs = b < L.MB * 8 ? b : 0;   // rows beyond the staged row groups read row 0 (in bounds)
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
            uint4 hr, lr, hz, lz, hn, ln, hnr, lnr;
            split8(dgr, hr, lr);
            split8(dgz, hz, lz);
            split8(dgn, hn, ln);
            split8(dgnr, hnr, lnr);
            if (act) {
                // dgi of the own units as the A operand of the partial's MMA chain: kblk g = gate g
                uint8_t* a2 = sA2 + (uint32_t)(b >> 3) * 512u + (uint32_t)(b & 7) * 16u;
                *reinterpret_cast<uint4*>(a2) = hr;
                *reinterpret_cast<uint4*>(a2 + 128) = hz;
                *reinterpret_cast<uint4*>(a2 + 256) = hn;
                *reinterpret_cast<uint4*>(a2 + 8192) = lr;
                *reinterpret_cast<uint4*>(a2 + 8192 + 128) = lz;
                *reinterpret_cast<uint4*>(a2 + 8192 + 256) = ln;
            }
            fence_proxy_async_smem();
            mbar_arrive(a2_full);
            if (etid == 0) TB_TRACE(6);
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
            if (etid == 0) TB_TRACE(7);
            fence_proxy_async_global();   // own generic writes of dgh_t -> visible to the peers' bulk copies (async proxy)
            if (a.early_h) named_bar_arrive(6, 160);   // warp 3 releases counter H: the next step's dgh pulls need not wait for the partials
            if (etid == 0) TB_TRACE(8);
            if (out > 32) {   // second half of the partial accumulator (the aux warps drain [0, 32))
                mbar_wait(part_full, (uint32_t)n & 1);
                tc_fence_after();
                drain_partial(taddr + TB_COL_P, pw, 32, act);
                tc_fence_before();
                fence_proxy_async_global();   // the reducers pull the partials with a bulk copy (async proxy)
            }
            named_bar_sync(7, 256);    // every finaliser published; the aux warps have drained D3 into `part`
            if (etid == 0) red_release_gpu_add(ctrA, 1u);   // release is cumulative over the barrier: one gpu-scope fence per CTA
            if (etid == 0) TB_TRACE(9);
            if (act) {   // off the critical path: only the products after the kernel read these
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
            // scratch = this CTA's own inbox: its last incoming copies were waited for and summed above.  (NOT the staging
            // buffer: the peers may still be pulling this CTA's last outgoing partial sums from it.)
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
    } else if (warp == 3) {
        // ================= counter H (early_h): dgh_t is published -> release, one gpu-scope fence, off the finalisers' path
        if (a.early_h) {
            for (int n = 0; n < T; ++n) {
                named_bar_sync(6, 160);
                if (lane == 0) red_release_gpu_add(ctrH, 1u);
            }
        }
    } else if (warp >= 8) {
        // ================= aux: dy reduction + publication, drain of the partial accumulator ============
        const int rt = threadIdx.x - 256;
        const int Q = L.Q;
        const int q_lo = c * Q;
        const int q_n = max(0, min(Q, B * 64 - q_lo));   // whole groups of 8 pairs q = b * 64 + o (PartWalk, umma.cuh)
        const uint32_t taddr = tmem + ((uint32_t)((warp - 8) * 32) << 16) + TB_COL_P;
        const PartWalk pw = part_walk(a.part, c, G, Q, rt, 0);     // this thread drains outputs [0, 32) of row rt
        const uint32_t red_bytes = (uint32_t)(G * Q) * 4u;         // this reducer's block: [G CTAs][Q pairs]
        const float* red_src = a.part + (size_t)c * G * Q;
        float* sPs = reinterpret_cast<float*>(smem + L.off_bar + 256);   // [4 warp groups][Q] partial sums (behind the mbarriers)
        // Reduction geometry, fixed before the steps (no division inside them; same scheme as gru_tc.cu).  Stage 1: a thread
        // sums ONE float4 column (4 consecutive pairs) over every nsub-th CTA; the subsets a warp holds are combined by
        // shuffles in fixed order; each warp group writes its partial row to sPs.  Stage 2 adds the rows in fixed order.
        const int ncol = q_n >> 2;
        const int nwc = ncol > 32 ? 2 : 1;
        const int ncw = ncol > 32 ? 32 : ncol;
        const int spw = (nwc == 1 && ncw > 0) ? 32 / ncw : 1;
        const int nwg = 4 / nwc;
        const int nsub = nwg * spw;
        const int wq = warp - 8, lane_sl = ncw > 0 ? lane / ncw : 0;
        const int col = (wq % nwc) * 32 + (ncw > 0 ? lane - lane_sl * ncw : 0);
        const int sid = (wq / nwc) * spw + lane_sl;
        const bool sum_act = ncol > 0 && lane_sl < spw && col < ncol;
        float* sProw = sPs + (size_t)(wq / nwc) * Q;
        // publication: thread g < ng owns the group of 8 consecutive outputs [po, po + 8) of row pbb
        const int ng = q_n >> 3;
        const int pq0 = q_lo + 8 * rt;
        const int pbb = pq0 >> 6, po = pq0 & 63;
        const bool pub_act = rt < ng;
        for (int n = 0; n <= T; ++n) {
            const int t = T - 1 - n;
            float* dyt = f.dy_tot + (size_t)(t + 1) * n_pairs;
            // the head's dY of this thread's group: fetched before the wait, only this thread ever modifies it
            float dyv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) dyv[e] = pub_act ? __ldcg(dyt + (size_t)pbb * out + min(po + e, out - 1)) : 0.f;
            uint16_t* dyx = a.dyx + (size_t)(n & 1) * 2 * dy_part;
            if (n > 0) {
                if (rt == 0 && q_n > 0) {
                    spin_until_ge(ctrA, (unsigned)G * (unsigned)n, a.relaxed != 0);
                    TB_TRACE(20);
                    // every CTA's partials of this reducer's pairs are one contiguous block: one bulk copy into sRed
                    mbar_expect_tx(red_full, red_bytes);
                    bulk_g2s(sRed, red_src, red_bytes, red_full);
                }
                if (q_n > 0) mbar_wait(red_full, (uint32_t)(n - 1) & 1);
                if (rt == 0) TB_TRACE(27);
                // stage 1
                float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
                if (sum_act) {
                    const float4* p = reinterpret_cast<const float4*>(sRed + (size_t)sid * Q + 4 * col);
                    const int stp = nsub * (Q >> 2);
                    int cc = sid;
                    for (; cc + nsub < G; cc += 2 * nsub, p += 2 * stp) {
                        const float4 x = p[0], y = p[stp];
                        a0.x += x.x; a0.y += x.y; a0.z += x.z; a0.w += x.w;
                        a1.x += y.x; a1.y += y.y; a1.z += y.z; a1.w += y.w;
                    }
                    if (cc < G) {
                        const float4 x = p[0];
                        a0.x += x.x; a0.y += x.y; a0.z += x.z; a0.w += x.w;
                    }
                    a0.x += a1.x; a0.y += a1.y; a0.z += a1.z; a0.w += a1.w;
                }
                for (int k = 1; k < spw; ++k) {   // the warp's other subsets of this column, in fixed order (warp-uniform trip count)
                    const int src = (lane + k * ncw) & 31;
                    const float vx = __shfl_sync(0xffffffffu, a0.x, src), vy = __shfl_sync(0xffffffffu, a0.y, src);
                    const float vz = __shfl_sync(0xffffffffu, a0.z, src), vw = __shfl_sync(0xffffffffu, a0.w, src);
                    if (lane_sl == 0) {
                        a0.x += vx; a0.y += vy; a0.z += vz; a0.w += vw;
                    }
                }
                if (sum_act && lane_sl == 0) *reinterpret_cast<float4*>(sProw + 4 * col) = a0;
                named_bar_sync(2, 128);
                if (pub_act) {   // stage 2: the warp groups' rows in fixed order, added to the head's dY
#pragma unroll
                    for (int e = 0; e < 8; e += 4) {
                        float4 v = *reinterpret_cast<const float4*>(sPs + 8 * rt + e);
                        for (int k = 1; k < nwg; ++k) {
                            const float4 x = *reinterpret_cast<const float4*>(sPs + (size_t)k * Q + 8 * rt + e);
                            v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
                        }
                        dyv[e] += v.x; dyv[e + 1] += v.y; dyv[e + 2] += v.z; dyv[e + 3] += v.w;
                    }
                }
            }
            if (pub_act && t >= 0) {   // dy_t in operand order: one 16-byte core-matrix row per plane
                float pv[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) pv[e] = (po + e < out) ? dyv[e] : 0.f;
                uint4 hi, lo;
                split8(pv, hi, lo);
                uint16_t* dst = dyx + (size_t)(pbb >> 3) * 512 + (size_t)(po >> 3) * 64 + (size_t)(pbb & 7) * 8;
                *reinterpret_cast<uint4*>(dst) = hi;
                *reinterpret_cast<uint4*>(dst + dy_part) = lo;
            }
            if (rt == 0) TB_TRACE(21);
            fence_proxy_async_global();
            named_bar_sync(2, 128);
            if (rt == 0) red_release_gpu_add(ctrB, 1u);
            if (rt == 0) TB_TRACE(22);
            if (n > 0 && pub_act) {   // the fp32 total: stored after the release (only later kernels read it)
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (po + e < out) dyt[(size_t)pbb * out + po + e] = dyv[e];
            }
            if (t < 0) break;
            // drain D3 (partial of the y feedback of step t, this CTA's units) into `part`
            mbar_wait(part_full, (uint32_t)n & 1);
            if (rt == 0) TB_TRACE(25);
            tc_fence_after();
            drain_partial(taddr, pw, 0, rt < B);   // the finaliser warps take [32, 64)
            tc_fence_before();
            fence_proxy_async_global();   // generic stores -> the reducers' bulk copies (async proxy)
            if (rt == 0) TB_TRACE(26);
            named_bar_sync(7, 256);   // a full barrier, not an arrive: thread 0's release after it must cover these warps' stores to `part`
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // no CTA leaves while a peer may still copy into its inbox
    if (warp == 2) tmem_dealloc<512>(tmem);
}

// ---- host side ----------------------------------------------------------------------------------------
bool gru_tc_bwd_shape_ok(int B, int H, int out) {
    return H % 64 == 0 && H >= 64 && out >= 1 && out <= 64 && B >= 1 && B <= 128 && (3 * H / TB_KC) % 4 == 0 && H % 32 == 0;
}

// floats of the partial-sum buffer part[G reducers][G CTAs][Q] (G = CTAs of the launch: H / 8)
static size_t tb_part_floats(int B, int H) {
    const size_t G = (size_t)H / 8;
    const size_t Q = 8 * (((size_t)8 * B + G - 1) / G);
    return round_up_sz(G * G * Q, 64);
}

size_t gru_tc_bwd_scratch_floats(int B, int H) {
    size_t MB = (B + 7) / 8;
    size_t gxh = (size_t)2 * 2 * (3 * H / TB_KC) * MB * 512 / 2;   // bf16 elements -> floats
    size_t dyx = (size_t)2 * 2 * MB * 512 / 2;
    const size_t two_hop = round_up_sz(gxh, 64) + round_up_sz(dyx, 64) + 128 + tb_part_floats(B, H);
    const size_t one_hop = gru_tc2_bwd_scratch_floats(B, H);
    return two_hop > one_hop ? two_hop : one_hop;
}

// picks the cluster size (8 preferred) whose clusters are all co-resident; 0 = not runnable
static int pick_cluster(int B, int H, int out, const DeviceInfo& di, TbLayout* Lout) {
    const int G = H / 8;
    if (!gru_tc_bwd_shape_ok(B, H, out) || G > di.n_sm) return 0;
    // clusters of 8 halve the MMA count per step but need 16 co-resident clusters at H = 1024 (this pool's B200s hold 15);
    // the 4-CTA shape is the one every measurement and parity run of this round used, so 8 is opt-in (CVB_TC_CLUSTER8=1)
    const char* e8 = getenv("CVB_TC_CLUSTER8");
    const int first = (e8 && e8[0] == '1') ? 8 : 4;
    struct Entry { int B, H, out, first, S; };
    static Entry cache[256];   // the row-count probe of cvb_recurrence_max_rows adds up to 16 entries per network shape
    static int n_cache = 0;
    int S = -1;
    for (int i = 0; i < n_cache; ++i)
        if (cache[i].B == B && cache[i].H == H && cache[i].out == out && cache[i].first == first) S = cache[i].S;
    if (S < 0) {
        S = 0;
        for (int cand = first; cand >= 4 && S == 0; cand >>= 1) {
            if (H % (8 * cand) != 0 || (3 * H / TB_KC) % cand != 0) continue;
            TbLayout L = tb_layout(B, H, cand, G, out, di.max_smem_optin);
            if (L.NS < 2 || (int)L.total > di.max_smem_optin) continue;
            if (cudaFuncSetAttribute(k_gru_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total) != cudaSuccess) {
                cudaGetLastError();
                continue;
            }
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(G);
            cfg.blockDim = dim3(TB_NT);
            cfg.dynamicSmemBytes = L.total;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cand;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int ncl = 0;
            if (cudaOccupancyMaxActiveClusters(&ncl, k_gru_bwd_tc, &cfg) != cudaSuccess) {
                cudaGetLastError();
                continue;
            }
            if (getenv("CVB_DEBUG")) fprintf(stderr, "[cvb] k_gru_bwd_tc: cluster %d: %d co-resident clusters (need %d), smem %u, ring %d\n", cand, ncl, G / cand, L.total, L.NS);
            if (ncl * cand >= G) S = cand;
        }
        if (n_cache < 256) cache[n_cache++] = Entry{B, H, out, first, S};
    }
    if (S && Lout) *Lout = tb_layout(B, H, S, G, out, di.max_smem_optin);
    return S;
}

bool gru_tc_bwd_supported(int B, int H, int out, const DeviceInfo& di) { return pick_cluster(B, H, out, di, nullptr) != 0; }

int gru_ar_bwd_tc(GruBwdArgs& f, float* tc_scratch, cudaStream_t s) {
    if (f.T <= 0 || f.B <= 0) return 0;
    DeviceInfo di;
    if (int rc = get_device_info(&di)) return rc;
    {
        // One-hop variant (gru_tc2_bwd.cu): faster up to ~32 rows (B = 8: 16.0k -> 14.8k cycles per step), slower beyond ~48: its
        // dgh pulls (245 KB per CTA and step at B = 80, the same lines read by all 32 clusters) run while other CTAs are still
        // in their tails and take twice as long as in the quiet window this kernel's single barrier gives them (21.9k vs
        // 20.3k cycles at B = 80).  CVB_TC_FEEDBACK=cluster / grid forces one or the other.
        const char* e8 = getenv("CVB_TC_CLUSTER8");
        const char* fb = getenv("CVB_TC_FEEDBACK");
        const bool want2 = fb ? fb[0] == 'c' : f.B <= 32;
        if (want2 && !(e8 && e8[0] == '1') && gru_tc2_bwd_supported(f.B, f.H, f.out, di)) {
            g_tc_hops[1] = 1;
            return gru_ar_bwd_tc2(f, tc_scratch, s);
        }
        g_tc_hops[1] = 2;
    }
    TbLayout L;
    const int S = pick_cluster(f.B, f.H, f.out, di, &L);
    CVB_REQUIRE(S != 0, "gru_ar_bwd_tc: unsupported shape B=%d H=%d out=%d", f.B, f.H, f.out);
    GruTcBwdArgs a;
    a.f = f;
    const size_t gxh_f = round_up_sz((size_t)2 * 2 * (3 * f.H / TB_KC) * L.MB * 512 / 2, 64);
    const size_t dyx_f = round_up_sz((size_t)2 * 2 * L.MB * 512 / 2, 64);
    a.gxh = reinterpret_cast<uint16_t*>(tc_scratch);
    a.dyx = reinterpret_cast<uint16_t*>(tc_scratch + gxh_f);
    a.ctr = reinterpret_cast<unsigned*>(tc_scratch + gxh_f + dyx_f);
    a.part = tc_scratch + gxh_f + dyx_f + 128;
    a.early_h = 0;
    if (const char* e = getenv("CVB_TC_BWD_EARLYH")) a.early_h = atoi(e) != 0;
    a.S = S;
    a.smem_max = di.max_smem_optin;
    a.trace = nullptr;
    a.keepalive = 1;
    a.relaxed = relaxed_polling() ? 1 : 0;
    if (const char* e = getenv("CVB_TC_KEEPALIVE")) a.keepalive = atoi(e) != 0;
    const char* trace_file = getenv("CVB_TRACE_FILE");
    const size_t trace_bytes = (size_t)(f.T + 1) * 64 * sizeof(long long);
    if (trace_file && trace_file[0]) {
        CVB_CHECK(cudaMalloc(&a.trace, trace_bytes));
        CVB_CHECK(cudaMemsetAsync(a.trace, 0, trace_bytes, s));
    }
    CVB_CHECK(cudaMemsetAsync(a.dyx, 0, (dyx_f + 128) * sizeof(float), s));   // dy padding columns + the counters
    CVB_CHECK(cudaFuncSetAttribute(k_gru_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(f.H / 8);
    cfg.blockDim = dim3(TB_NT);
    cfg.dynamicSmemBytes = L.total;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = S;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
    cfg.attrs = at;
    // CVB_TC_NOCOOP=1 drops the cooperative attribute (profilers that cannot replay a cooperative cluster launch; the
    // grid still fits one wave, but co-residency is then only true on an otherwise idle device)
    cfg.numAttrs = launch_without_coop() ? 1 : 2;
    prof_begin(s, CVB_PROF_GRU_BWD);
    CVB_CHECK(cudaLaunchKernelEx(&cfg, k_gru_bwd_tc, a));
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
