// Tensor-core (tcgen05) forward recurrence of the autoregressive GRU (gru_vae.py:364-399), one persistent
// cooperative launch for all T steps.
//
// 2-D work split over thread-block clusters of S = 4 CTAs: cluster i owns the block of 32 hidden units
// [32i, 32i+32); CTA j of the cluster owns the K-slice [j*H/4, (j+1)*H/4) of the contraction
// gh[b, (g,u)] = sum_k h_{t-1}[b,k] W_hh[g*H+u, k]  (g = r, z, n) for ALL units of the block, and FINALISES the 8
// units [32i + 8j, +8).  The W_hh rows of (block x K-slice) stay in shared memory for the whole sequence as fp16
// hi+lo (x = hi + lo + O(2^-22 x)) stored [hi rows | lo rows], so per K step ONE MMA with N = 192 forms
// A_hi B_hi (96 columns) and A_hi B_lo (96 columns) and a second with N = 96 adds A_lo B_hi; fp32 accumulation in
// TMEM, the halves added in registers.  A CTA ingests only 1/4 of the all-gathered h_{t-1} per step.  The four
// partial accumulators of a unit meet in the finaliser's shared memory through bulk DSMEM copies
// (cp.async.bulk.shared::cluster, complete_tx on the receiver's mbarrier) and are summed in fixed order: the
// accumulation chain inside the tensor core is H/64 steps long, the rest is fp32 round-to-nearest.
//
// The y feedback runs on the tensor core as well: W_y y_{t-1} for the own 8 units is one 64-deep MMA chain
// (accumulator D2) on the y_{t-1} chunk every CTA pulls through its ring, and the partial of
// y_t = W_o o_t + b_o over the own units is one MMA (accumulator D3) drained into part[c][o][b]; the per-pair
// sums over the CTAs are formed in fixed order by the "aux" warps of the CTA that owns the pair (no float atomics:
// run-to-run deterministic) and published in operand order for the next step.
//
// Roles (384 threads): w0 bulk-copy producer, w1 MMA issuer, w2 TMEM allocator, w4-7 exchange + gates
// (TMEM lane == batch row), w8-11 aux (drain of D3, y reduction + publication).  Grid-wide ordering is two
// monotonic counters: A (h_t / partials published), B (y_t published).
#include <stdlib.h>

#include "gru_ar.cuh"
#include "umma.cuh"

namespace cvb {
using namespace umma;

constexpr int TF_NT = 384;
constexpr int TF_KC = 64;             // K per ring stage
constexpr int TF_S = 4;               // cluster size
constexpr int TF_UB = 8 * TF_S;       // units per cluster
constexpr int TF_NW = 3 * TF_UB;      // rows of the W_hh operand (r, z, n of the block) = 96
// TMEM columns.  D1 = [main0 | corrections | main1]: the hi x hi products of the first / second half of the K-slice go to
// two accumulators (a tcgen05 accumulation chain truncates toward zero: no fp32 accumulator sums more than K = 128,
// tools/split_error_budget.py), the two cross products (lo planes scaled by 2^11, umma.cuh) share the middle one.
constexpr uint32_t TF_COL_M0 = 0;
constexpr uint32_t TF_COL_C = TF_NW;
constexpr uint32_t TF_COL_M1 = 2 * TF_NW;
constexpr uint32_t TF_COL_Y = 288;    // D2: W_y y for the own units, 32 main + 32 correction columns
constexpr uint32_t TF_COL_P = 352;    // D3: partial of y_t, 64 main + 64 correction columns
constexpr uint32_t TF_COL_DUMMY = 480;
// y_t is pulled by every CTA at the same moment: 128 readers of the same 160 lines queue up in the L2 slices (the pull took
// 1360 cycles on the luckiest SM and 3570 on the unluckiest).  The reducers publish TF_YREP copies; CTA c reads copy c % TF_YREP.
constexpr int TF_YREP = 8;   // upper bound; gru_ar_fwd_tc publishes 2 unless CVB_TC_YREP says otherwise (1 vs 8 measured equal)

struct TfLayout {
    int MB, nch, NS;
    uint32_t half, stage_bytes, w_chunk_bytes, slot_bytes, ybuf_bytes, bias_bytes;
    int Q;   // pairs (b, o) per reducer CTA, a multiple of 8
    uint32_t off_ring, off_ybuf, off_w, off_b2, off_b3, off_a2, off_inbox, off_bias, off_bar, total;
};

__host__ __device__ inline TfLayout tf_layout(int B, int H, int G, int out, int smem_max) {
    TfLayout L;
    L.MB = (B + 7) / 8;
    L.nch = H / TF_KC / TF_S;
    L.half = (uint32_t)L.MB * 1024u;
    L.stage_bytes = 2u * L.half;
    L.w_chunk_bytes = 2u * (TF_NW / 8) * 1024u;                 // [hi: 12 row groups][lo: 12 row groups] x 1 KB
    L.slot_bytes = (uint32_t)L.MB * 8u * 96u;                   // [rows][24 floats]
    uint32_t inbox = (uint32_t)TF_S * L.slot_bytes;
    inbox = (inbox + 127u) & ~127u;
    L.Q = 8 * ((8 * B + G - 1) / G);                             // tf_pairs_per_reducer
    const uint32_t red_bytes = (uint32_t)(G * L.Q) * 4u;         // a reducer's block of partials: staged where the y operand lands
    L.ybuf_bytes = ((L.stage_bytes > red_bytes ? L.stage_bytes : red_bytes) + 1023u) & ~1023u;
    L.bias_bytes = 128u + (uint32_t)(4 * L.Q) * 4u;              // [24] b_hh | [4 warp groups][Q] partial sums of the reducers
    const uint32_t fixed = L.ybuf_bytes + (uint32_t)L.nch * L.w_chunk_bytes + 8192u + 4096u + 8192u + inbox + L.bias_bytes + 256u;
    int ns = ((int)smem_max - (int)fixed) / (int)L.stage_bytes;
    L.NS = ns > 6 ? 6 : ns;
    const uint32_t ring = (uint32_t)(L.NS > 0 ? L.NS : 0) * L.stage_bytes;
    L.off_ring = 0;                       // h chunks only: idle between a step's last h chunk and the next step's first, when it
                                          // doubles as the staging of the outgoing partial sums
    L.off_ybuf = ring;                    // y_{t-1} operand (hi | lo), its own buffer: it lands while the exchange is under way
    L.off_w = L.off_ybuf + L.ybuf_bytes;
    L.off_b2 = L.off_w + (uint32_t)L.nch * L.w_chunk_bytes;   // W_y rows of the own units: [hi 4 groups][lo 4 groups] x 1 KB
    L.off_b3 = L.off_b2 + 8192u;          // W_o columns of the own units: [hi 8 groups][lo 8 groups] x 256 B
    L.off_a2 = L.off_b3 + 4096u;          // o_t of the own units: [2 parts][16 row groups][2 kblk][8][8]
    L.off_inbox = L.off_a2 + 8192u;
    L.off_bias = L.off_inbox + inbox;
    L.off_bar = L.off_bias + L.bias_bytes;
    L.total = L.off_bar + 256u;
    return L;
}

struct GruTcArgs {
    GruFwdArgs f;
    uint16_t* hx;        // [2 slots][2 parts][H/64 chunks][MB][8 kblk][8 rows][8 k] fp16 (UMMA order) of h_t
    uint16_t* yx;        // [TF_YREP replicas][2 slots][2 parts][MB][8 kblk][8 rows][8 k] fp16 of y_t, zero-initialised
    float* part;         // [G reducers][G CTAs][Q] partial sums of y_t (part_walk)
    int yrep;            // replicas of y_t actually published / read (1..TF_YREP, CVB_TC_YREP)
    int ymc;             // CVB_TC_YMC=1: y_t pulled once per cluster and multicast (A/B; default: every CTA pulls its own copy)
    unsigned* ctr;       // [0] = A, [32] = B (separate 128-B lines), zero-initialised
    int smem_max;
    int keepalive;
    int relaxed;         // CVB_TC_POLL=relaxed: poll the arrival counters with relaxed loads + one acquire fence instead of ld.acquire
    long long* trace;    // optional [T+1][64] clock64 stamps of CTA 0 (CVB_TRACE_FILE_FWD), else null
};

static __device__ __forceinline__ long long globaltimer_ns() {
    long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
}
// cross-CTA skew probe: globaltimer of every CTA at a few events of step 40 (trace rows behind the per-step table)
#define TF_SKEW(slot)                                                                                       \
    do {                                                                                                    \
        if (a.trace && t == 40) a.trace[(size_t)(T + 1) * 64 + (size_t)(slot) * 256 + c] = globaltimer_ns(); \
    } while (0)
// per-CTA phase stamps (clock64 of the CTA's own SM: differences within a CTA are exact) of steps 40 and 41
#define TF_PH(slot)                                                                                                          \
    do {                                                                                                                     \
        if (a.trace && (t == 40 || t == 41)) a.trace[(size_t)(T + 1) * 64 + 8 * 256 + (size_t)((t - 40) * 24 + (slot)) * 256 + c] = clock64(); \
    } while (0)
#define TF_TRACE(ev)                                                     \
    do {                                                                 \
        if (a.trace && c == 0) a.trace[(size_t)t * 64 + (ev)] = clock64(); \
    } while (0)

static __device__ __forceinline__ void split8_f16(const float* x, uint4& hi, uint4& lo) {
    uint16_t h[8], l[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) split_f16(x[q], h[q], l[q]);
    hi = make_uint4((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16),
                    (uint32_t)h[4] | ((uint32_t)h[5] << 16), (uint32_t)h[6] | ((uint32_t)h[7] << 16));
    lo = make_uint4((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16),
                    (uint32_t)l[4] | ((uint32_t)l[5] << 16), (uint32_t)l[6] | ((uint32_t)l[7] << 16));
}
// drain 32 columns (outputs [o_lo, o_lo + 32)) of D3 (main + correction halves) of this warp's 32 TMEM lanes into `part`
static __device__ __forceinline__ void drain_partial_y(uint32_t taddr_p, const PartWalk& w, int o_lo, bool row_ok) {
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
            if (row_ok)
                st_global_v4(addr, fmaf(v2[q], F16_LO_INV, v[q]), fmaf(v2[q + 1], F16_LO_INV, v[q + 1]), fmaf(v2[q + 2], F16_LO_INV, v[q + 2]),
                             fmaf(v2[q + 3], F16_LO_INV, v[q + 3]));
            addr += 16;
            i += 4;
            if (i >= w.Q) {
                i -= w.Q;
                addr += w.wrap;
            }
        }
    }
}

__global__ void __launch_bounds__(TF_NT, 1) k_gru_fwd_tc(GruTcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const GruFwdArgs& f = a.f;
    const int B = f.B, T = f.T, H = f.H, out = f.out;
    const int G = gridDim.x, c = blockIdx.x;
    const int j = (int)cluster_ctarank();
    const TfLayout L = tf_layout(B, H, G, out, a.smem_max);
    const int ublk0 = (c / TF_S) * TF_UB;   // first unit of the cluster's block
    const int u0 = ublk0 + 8 * j;           // first of the 8 units this CTA finalises
    const int k0 = j * L.nch * TF_KC;       // first column of W_hh (= unit of h) of this CTA's K-slice
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    uint8_t* ring = smem + L.off_ring;
    float* stage = reinterpret_cast<float*>(ring);                 // [S (to)][MB*8][24], aliases the (idle) ring
    uint8_t* sW = smem + L.off_w;
    uint8_t* sB2 = smem + L.off_b2;
    uint8_t* sB3 = smem + L.off_b3;
    uint8_t* sA2 = smem + L.off_a2;
    float* inbox = reinterpret_cast<float*>(smem + L.off_inbox);   // [S (from)][MB*8][24]
    // [G][w] staging of the partials being reduced: aliases the y operand buffer.  The reducers of round r work between
    // "counter A >= G(r+1)" and their own arrival on counter B; the y copy of step r is issued only after counter B is
    // complete, and its MMAs retire before this CTA's next arrival on counter A -- the two uses never overlap.  (The
    // inbox must NOT be reused: the peers' exchange copies land as soon as THEIR h chunks are consumed.)
    float* sRed = reinterpret_cast<float*>(smem + L.off_ybuf);
    float* sBh = reinterpret_cast<float*>(smem + L.off_bias);      // [3][8] b_hh of the own units
    float* sPs = sBh + 32;                                         // [4 warp groups][Q] partial sums of the reducers
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* empty = full + 8;
    uint64_t* accum_full = empty + 8;   // D2 (W_y y) complete
    uint64_t* d1_full = full + 24;      // D1 (K-slice of W_hh h) complete
    uint64_t* y_full = full + 25;
    uint64_t* y_empty = full + 26;
    uint64_t* red_full = full + 27;     // the reducers' block of partials has landed in sRed
    uint8_t* ybuf = smem + L.off_ybuf;
    uint64_t* inbox_full = accum_full + 1;
    uint64_t* a2_full = inbox_full + 1;
    uint64_t* part_full = a2_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(part_full + 1);
    const size_t hx_part = (size_t)(H / TF_KC) * L.MB * 512;   // elements per part
    const size_t yx_part = (size_t)L.MB * 512;
    const size_t yx_rep = 4 * yx_part;                         // one replica: [2 slots][2 parts]
    unsigned* ctrA = a.ctr;
    unsigned* ctrB = a.ctr + 32;
    const int n_pairs = B * out;
    const int half1 = (L.nch + 1) / 2;   // first chunk of the K-slice that accumulates into main1 (== nch: a single chain of <= 8 MMAs)
    const bool two_main = half1 < L.nch;

    if (a.trace && c == 0 && threadIdx.x == 0) a.trace[60] = clock64();
    // ---- one-time setup: weights -> fp16 hi/lo in UMMA K-major core-matrix order -----------------
    {
        // one item = 8 consecutive k of one operand row = one 16-byte core-matrix row (hi) + one (lo); consecutive
        // threads take consecutive rows, so a quarter warp's stores cover 128 contiguous bytes (no bank conflicts) and
        // every global read is a full 32-byte sector
        const int n_items = TF_NW * L.nch * (TF_KC / 8);
        // batches of 8 items per thread with all 16 loads of a batch in flight before the first conversion (the loop used
        // to expose one HBM round trip per item: 27 us of set-up per launch)
        for (int i0 = threadIdx.x; i0 < n_items; i0 += 8 * TF_NT) {
            float4 wv[16];
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int i = i0 + m * TF_NT;
                if (i < n_items) {
                    const int kg = i / TF_NW, n = i - kg * TF_NW;   // n = g*32 + unit of the block
                    const int g = n / TF_UB, ul = n - g * TF_UB;
                    const float4* src = reinterpret_cast<const float4*>(f.Whh + (size_t)(g * H + ublk0 + ul) * H + k0 + kg * 8);
                    wv[2 * m] = __ldg(src);
                    wv[2 * m + 1] = __ldg(src + 1);
                }
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int i = i0 + m * TF_NT;
                if (i < n_items) {
                    const int kg = i / TF_NW, n = i - kg * TF_NW;
                    const int kl = kg * 8;
                    const float w[8] = {wv[2 * m].x, wv[2 * m].y, wv[2 * m].z, wv[2 * m].w, wv[2 * m + 1].x, wv[2 * m + 1].y, wv[2 * m + 1].z, wv[2 * m + 1].w};
                    uint4 hi, lo;
                    split8_f16(w, hi, lo);
                    const uint32_t off = (uint32_t)(kl / TF_KC) * L.w_chunk_bytes + (uint32_t)(n >> 3) * 1024u + (uint32_t)((kl % TF_KC) >> 3) * 128u +
                                         (uint32_t)(n & 7) * 16u;
                    // chunks of the second half of the K walk are stored [lo rows | hi rows]: their stacked MMA starts at the
                    // correction columns and runs on into main1
                    const bool swapped = (kl / TF_KC) >= half1;
                    *reinterpret_cast<uint4*>(sW + off + (swapped ? (TF_NW / 8) * 1024u : 0u)) = hi;
                    *reinterpret_cast<uint4*>(sW + off + (swapped ? 0u : (TF_NW / 8) * 1024u)) = lo;
                }
            }
        }
        {   // B2[n = g*8+uu][k] = W_y[g*H + u0 + uu][k]; rows 24..31 zero.  B3[n = o][k = uu] = W_o[o][u0 + uu]; k 8..15 zero
            float w2[6], w3[3];   // all loads in flight before the first conversion
#pragma unroll
            for (int m = 0; m < 6; ++m) {
                const int i = threadIdx.x + m * TF_NT, n = i >> 6, k = i & 63;
                w2[m] = (i < 32 * 64 && n < 24 && k < out) ? __ldg(f.Wy + (size_t)((n >> 3) * H + u0 + (n & 7)) * f.ldwy + k) : 0.f;
            }
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const int i = threadIdx.x + m * TF_NT, n = i >> 4, k = i & 15;
                w3[m] = (i < 64 * 16 && k < 8 && n < out) ? __ldg(f.Wo + (size_t)n * H + u0 + k) : 0.f;
            }
#pragma unroll
            for (int m = 0; m < 6; ++m) {
                const int i = threadIdx.x + m * TF_NT, n = i >> 6, k = i & 63;
                if (i < 32 * 64) {
                    uint16_t hi, lo;
                    split_f16(w2[m], hi, lo);
                    const uint32_t off = (uint32_t)(n >> 3) * 1024u + (uint32_t)(k >> 3) * 128u + (uint32_t)(n & 7) * 16u + (uint32_t)(k & 7) * 2u;
                    *reinterpret_cast<uint16_t*>(sB2 + off) = hi;
                    *reinterpret_cast<uint16_t*>(sB2 + 4096 + off) = lo;
                }
            }
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const int i = threadIdx.x + m * TF_NT, n = i >> 4, k = i & 15;
                if (i < 64 * 16) {
                    uint16_t hi, lo;
                    split_f16(w3[m], hi, lo);
                    const uint32_t off = (uint32_t)(n >> 3) * 256u + (uint32_t)(k >> 3) * 128u + (uint32_t)(n & 7) * 16u + (uint32_t)(k & 7) * 2u;
                    *reinterpret_cast<uint16_t*>(sB3 + off) = hi;
                    *reinterpret_cast<uint16_t*>(sB3 + 2048 + off) = lo;
                }
            }
        }
        for (int i = threadIdx.x; i < 8192 / 16; i += TF_NT) reinterpret_cast<uint4*>(sA2)[i] = make_uint4(0u, 0u, 0u, 0u);
        if (threadIdx.x < 24) sBh[threadIdx.x] = f.bhh[(threadIdx.x >> 3) * H + u0 + (threadIdx.x & 7)];
        if (threadIdx.x == 0) {
            for (int s = 0; s < 8; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            mbar_init(accum_full, 1);
            mbar_init(d1_full, 1);
            mbar_init(y_full, 1);
            mbar_init(y_empty, 1);
            mbar_init(inbox_full, 1);   // armed with expect_tx(S slots) every step; the peers' bulk copies complete_tx
            mbar_init(a2_full, 128);
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
    if (a.trace && c == 0 && threadIdx.x == 0) a.trace[61] = clock64();

    if (warp == 0) {
        // ================= producer: K-slice of h_{t-1} chunk by chunk, then y_{t-1} ====================
        int s = 0;
        uint32_t ph = 1;
        for (int t = 0; t < T; ++t) {
            const uint16_t* src = a.hx + (size_t)(t & 1) * 2 * hx_part + (size_t)(j * L.nch) * L.MB * 512;
            const uint16_t* srcy = a.yx + (size_t)((a.ymc ? c / TF_S : c) % a.yrep) * yx_rep + (size_t)(t & 1) * 2 * yx_part;
            if (lane == 0) {
                spin_until_ge(ctrA, (unsigned)G * (unsigned)(t + 1), a.relaxed != 0);   // the writers fenced generic -> async proxy before their release
                TF_TRACE(14);
                TF_SKEW(4);
                TF_PH(11);
            }
            __syncwarp();
            for (int ch = 0; ch < L.nch; ++ch) {
                if (lane == 0) {
                    mbar_wait(&empty[s], ph);
                    if (ch < 8) TF_TRACE(32 + ch);
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
            if (lane == 0) {
                // y_{t-1}: ONE CTA of the cluster pulls it and multicasts it into all four (every CTA reading the same lines at
                // the same moment queued up in the L2 slices: the pull took 1500 cycles on the luckiest SM, 3500 on the
                // unluckiest).  When counter B is complete every CTA's y buffer is free: its MMAs of the previous step
                // retired before that CTA's arrival on A, its reducers' reads (sRed aliases it) before its arrival on B.
                mbar_wait(y_empty, ((uint32_t)t & 1) ^ 1);
                mbar_expect_tx(y_full, 2 * L.half);
                if (!a.ymc || j == 0) {
                    spin_until_ge(ctrB, (unsigned)G * (unsigned)(t + 1), a.relaxed != 0);
                    TF_TRACE(13);
                    TF_SKEW(5);
                    TF_PH(12);
                    if (a.ymc) {
                        bulk_g2s_multicast(ybuf, srcy, L.half, y_full, (uint16_t)((1u << TF_S) - 1u));
                        bulk_g2s_multicast(ybuf + L.half, srcy + yx_part, L.half, y_full, (uint16_t)((1u << TF_S) - 1u));
                    } else {
                        bulk_g2s(ybuf, srcy, L.half, y_full);
                        bulk_g2s(ybuf + L.half, srcy + yx_part, L.half, y_full);
                    }
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ================= MMA issuer (descriptors stay warp-uniform; one elected lane issues) ===========
        const uint32_t idesc1s = idesc_f16_f32(128, 2 * TF_NW), idesc1 = idesc_f16_f32(128, TF_NW);
        const uint32_t idesc2s = idesc_f16_f32(128, 64), idesc2 = idesc_f16_f32(128, 32);
        const uint32_t idesc3s = idesc_f16_f32(128, 128), idesc3 = idesc_f16_f32(128, 64);
        const uint32_t idesc_dummy = idesc_f16_f32(128, 16);
        const uint64_t dA0 = smem_desc(smem_u32(ring), 128, 1024);
        const uint64_t dW0 = smem_desc(smem_u32(sW), 128, 1024);
        const uint64_t dY0 = smem_desc(smem_u32(ybuf), 128, 1024);
        const uint64_t dB2 = smem_desc(smem_u32(sB2), 128, 1024);
        const uint64_t dA2 = smem_desc(smem_u32(sA2), 128, 256);
        const uint64_t dB3 = smem_desc(smem_u32(sB3), 128, 256);
        const uint32_t a_step = L.stage_bytes >> 4, half16 = L.half >> 4, w_step = L.w_chunk_bytes >> 4;
        // poll an mbarrier; while idle keep the tensor pipe warm with a dummy MMA into scratch columns (the first MMA
        // after a few microseconds of idleness was measured to stall ~3400 cycles at issue)
        auto wait_warm = [&](uint64_t* bar, uint32_t parity) {
            for (;;) {
                uint32_t ok = (lane == 0) ? (mbar_test_wait(bar, parity) ? 1u : 0u) : 0u;
                ok = __shfl_sync(0xffffffffu, ok, 0);
                if (ok) break;
                if (a.keepalive) mma_bf16_ss_elect(tmem + TF_COL_DUMMY, dA2, dB2, idesc_dummy, false);
            }
        };
        int s = 0;
        uint32_t ph = 0;
        for (int t = 0; t < T; ++t) {
            for (int ch = 0; ch < L.nch; ++ch) {
                wait_warm(&full[s], ph);
                if (lane == 0 && ch < 8) TF_TRACE(40 + ch);
                tc_fence_after();
                const uint64_t da = dA0 + (uint64_t)((uint32_t)s * a_step);
                const uint64_t db = dW0 + (uint64_t)((uint32_t)ch * w_step);
                const uint32_t w_half16 = (TF_NW / 8) * 1024u >> 4;   // hi rows -> lo rows (or lo -> hi in a swapped chunk)
                if (ch < half1) {   // [hi | lo] rows: main0 and corrections side by side
#pragma unroll
                    for (int k16 = 0; k16 < TF_KC / 16; ++k16) {
                        mma_bf16_ss_elect(tmem + TF_COL_M0, da + 16u * k16, db + 16u * k16, idesc1s, (ch | k16) != 0);
                        mma_bf16_ss_elect(tmem + TF_COL_C, da + half16 + 16u * k16, db + 16u * k16, idesc1, true);
                    }
                } else {            // [lo | hi] rows: corrections and main1 side by side
#pragma unroll
                    for (int k16 = 0; k16 < TF_KC / 16; ++k16) {
                        if (ch == half1 && k16 == 0) {   // main1 starts from zero while the corrections keep accumulating
                            mma_bf16_ss_elect(tmem + TF_COL_C, da, db, idesc1, true);
                            mma_bf16_ss_elect(tmem + TF_COL_M1, da, db + w_half16, idesc1, false);
                        } else {
                            mma_bf16_ss_elect(tmem + TF_COL_C, da + 16u * k16, db + 16u * k16, idesc1s, true);
                        }
                        mma_bf16_ss_elect(tmem + TF_COL_C, da + half16 + 16u * k16, db + w_half16 + 16u * k16, idesc1, true);
                    }
                }
                mma_commit_elect(&empty[s]);
                if (ch == L.nch - 1) mma_commit_elect(d1_full);   // the exchange of the partial sums does not wait for y_{t-1}
                if (lane == 0 && ch < 8) TF_TRACE(48 + ch);
                if (++s == L.NS) {
                    s = 0;
                    ph ^= 1;
                }
            }
            {   // W_y y_{t-1} of the own units
                wait_warm(y_full, (uint32_t)t & 1);
                tc_fence_after();
#pragma unroll
                for (int k16 = 0; k16 < TF_KC / 16; ++k16) {
                    mma_bf16_ss_elect(tmem + TF_COL_Y, dY0 + 16u * k16, dB2 + 16u * k16, idesc2s, k16 != 0);
                    mma_bf16_ss_elect(tmem + TF_COL_Y + 32u, dY0 + half16 + 16u * k16, dB2 + 16u * k16, idesc2, true);
                }
                mma_commit_elect(y_empty);
                mma_commit_elect(accum_full);
            }
            // partial of y_t over the own units: D3[b][o] = sum_uu o_t[b][uu] W_o[o][u0+uu]
            wait_warm(a2_full, (uint32_t)t & 1);
            tc_fence_after();
            mma_bf16_ss_elect(tmem + TF_COL_P, dA2, dB3, idesc3s, false);
            mma_bf16_ss_elect(tmem + TF_COL_P + 64u, dA2 + 256u, dB3, idesc3, true);
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
        float hreg[8];
        const PartWalk pw = part_walk(a.part, c, G, L.Q, b, 32);   // this thread drains outputs [32, 64) of its row
        // prologue: publish h_in (slot 0) in operand order
        {
            const int t = 0;
            (void)t;
#pragma unroll
            for (int q = 0; q < 8; ++q) hreg[q] = act ? f.hs[(size_t)b * H + u0 + q] : 0.f;
            uint4 hh, hl;
            split8_f16(hreg, hh, hl);
            if (act) {
                const size_t off = ((size_t)(u0 >> 6) * L.MB + (b >> 3)) * 512 + (size_t)((u0 & 63) >> 3) * 64 + (size_t)(b & 7) * 8;
                *reinterpret_cast<uint4*>(a.hx + off) = hh;
                *reinterpret_cast<uint4*>(a.hx + hx_part + off) = hl;
            }
            fence_proxy_async_all();
            named_bar_sync(1, 128);
            if (etid == 0) red_release_gpu_add(ctrA, 1u);
        }
        for (int t = 0; t < T; ++t) {
            const size_t row = (size_t)t * B + (act ? b : 0);
            float4 gxv[6];
            float4 mk[2] = {make_float4(1.f, 1.f, 1.f, 1.f), make_float4(1.f, 1.f, 1.f, 1.f)};
            {   // every lane loads (the rows beyond B re-read row 0 of the frame: `row`), so the gate math below is branch-free
                const float* gp = f.gx + row * 3 * H + u0;
#pragma unroll
                for (int gi = 0; gi < 3; ++gi) {
                    gxv[2 * gi] = ldg_nc_v4_pinned(gp + (size_t)gi * H);
                    gxv[2 * gi + 1] = ldg_nc_v4_pinned(gp + (size_t)gi * H + 4);
                }
                if (f.mask) {
                    mk[0] = ldg_nc_v4_pinned(f.mask + row * H + u0);
                    mk[1] = ldg_nc_v4_pinned(f.mask + row * H + u0 + 4);
                }
            }
            if (etid == 0) TF_TRACE(0);
            if (etid == 0) TF_PH(0);
            if (etid == 0) mbar_expect_tx(inbox_full, (uint32_t)TF_S * L.slot_bytes);
            mbar_wait(d1_full, (uint32_t)t & 1);
            if (etid == 0) TF_TRACE(1);
            if (etid == 0) TF_PH(1);
            tc_fence_after();
            // partial sums (main + correction) of the block's units -> the finalisers' inboxes
#pragma unroll
            for (int g = 0; g < 3; ++g) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    float v[16], v2[16];
                    tmem_ld_x16(taddr + TF_COL_M0 + g * TF_UB + 16 * k, v);
                    tmem_ld_x16(taddr + TF_COL_C + g * TF_UB + 16 * k, v2);
                    if (two_main) {
                        float v3[16];
                        tmem_ld_x16(taddr + TF_COL_M1 + g * TF_UB + 16 * k, v3);
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
            fence_proxy_async_smem();
            named_bar_sync(3, 128);
            if (etid < TF_S)
                bulk_s2c(mapa(inbox_addr + (uint32_t)j * L.slot_bytes, (uint32_t)etid), stage + (size_t)etid * slot_f, L.slot_bytes,
                         mapa(inbox_bar_addr, (uint32_t)etid));
            if (etid == 0) TF_TRACE(2);
            if (etid == 0) TF_PH(2);
            // W_y y_{t-1} of the own units: columns [r 8 | z 8 | n 8 | pad 8] (+ correction half at +32)
            float yr[8], yz[8], yn[8];
            mbar_wait(accum_full, (uint32_t)t & 1);
            if (etid == 0) TF_TRACE(4);
            if (etid == 0) TF_PH(3);
            tc_fence_after();
            {
                float c2[8];
                tmem_ld_x8(taddr + TF_COL_Y, yr);
                tmem_ld_x8(taddr + TF_COL_Y + 32, c2);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q) yr[q] = fmaf(c2[q], F16_LO_INV, yr[q]);
                tmem_ld_x8(taddr + TF_COL_Y + 8, yz);
                tmem_ld_x8(taddr + TF_COL_Y + 40, c2);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q) yz[q] = fmaf(c2[q], F16_LO_INV, yz[q]);
                tmem_ld_x8(taddr + TF_COL_Y + 16, yn);
                tmem_ld_x8(taddr + TF_COL_Y + 48, c2);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q) yn[q] = fmaf(c2[q], F16_LO_INV, yn[q]);
            }
            tc_fence_before();
            mbar_wait_cluster(inbox_full, (uint32_t)t & 1);
            if (etid == 0) TF_TRACE(3);
            if (etid == 0) TF_PH(4);
            // Straight-line code for every lane, active row or not (the rows beyond B hold garbage that is never stored): with
            // the math of each unit inside an `if (act)` the compiler kept eight separate branch regions and the eight
            // dependent chains LDS -> ex2 -> rcp -> ex2 -> rcp ran one after the other (2500 cycles per step; per-CTA phase stamps).
            float ar[8], az[8], an[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) ar[q] = az[q] = an[q] = 0.f;
            {
                const int bs = b < L.MB * 8 ? b : 0;   // rows beyond the staged row groups read row 0 (in bounds)
#pragma unroll
                for (int p = 0; p < TF_S; ++p) {   // fixed order: deterministic
                    const float4* x = reinterpret_cast<const float4*>(inbox + (size_t)p * slot_f + bs * 24);
                    const float4 x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3], x4 = x[4], x5 = x[5];
                    ar[0] += x0.x; ar[1] += x0.y; ar[2] += x0.z; ar[3] += x0.w; ar[4] += x1.x; ar[5] += x1.y; ar[6] += x1.z; ar[7] += x1.w;
                    az[0] += x2.x; az[1] += x2.y; az[2] += x2.z; az[3] += x2.w; az[4] += x3.x; az[5] += x3.y; az[6] += x3.z; az[7] += x3.w;
                    an[0] += x4.x; an[1] += x4.y; an[2] += x4.z; an[3] += x4.w; an[4] += x5.x; an[5] += x5.y; an[6] += x5.z; an[7] += x5.w;
                }
            }
            if (etid == 0) TF_TRACE(5);
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
                for (int q = 0; q < 8; ++q) {
                    rr[q] = sigmoid_fast(gxr[q] + yr[q] + ar[q] + bh[q]);
                    zz[q] = sigmoid_fast(gxz[q] + yz[q] + az[q] + bh[8 + q]);
                    gh[q] = an[q] + bh[16 + q];
                    nn[q] = tanh_fast(gxn[q] + yn[q] + rr[q] * gh[q]);
                    hreg[q] = (1.0f - zz[q]) * nn[q] + zz[q] * hreg[q];
                    ov[q] = hreg[q] * mkf[q];
                }
            }
            if (etid == 0) TF_TRACE(10);
            if (etid == 0) TF_PH(5);
            uint4 oh, ol, hh, hl;
            split8_f16(ov, oh, ol);
            split8_f16(hreg, hh, hl);
            if (etid == 0) TF_TRACE(11);
            if (act) {   // o_t of the own units as the A operand of the partial's MMA (k block 0; k block 1 stays zero)
                uint8_t* a2 = sA2 + (uint32_t)(b >> 3) * 256u + (uint32_t)(b & 7) * 16u;
                *reinterpret_cast<uint4*>(a2) = oh;
                *reinterpret_cast<uint4*>(a2 + 4096) = ol;
            }
            fence_proxy_async_smem();
            mbar_arrive(a2_full);
            if (etid == 0) TF_TRACE(6);
            if (act) {   // publish h_t (fp16 hi/lo, operand order) into the other exchange slot
                uint16_t* hdst = a.hx + (size_t)((t + 1) & 1) * 2 * hx_part;
                const size_t off = ((size_t)(u0 >> 6) * L.MB + (b >> 3)) * 512 + (size_t)((u0 & 63) >> 3) * 64 + (size_t)(b & 7) * 8;
                *reinterpret_cast<uint4*>(hdst + off) = hh;
                *reinterpret_cast<uint4*>(hdst + hx_part + off) = hl;
            }
            if (etid == 0) TF_TRACE(7);
            fence_proxy_async_global();   // own generic writes of h_t -> visible to the peers' bulk copies (async proxy)
            if (etid == 0) TF_TRACE(8);
            if (etid == 0) TF_PH(6);
            if (out > 32) {   // second half of the partial accumulator (the aux warps drain [0, 32))
                mbar_wait(part_full, (uint32_t)t & 1);
                if (etid == 0) TF_PH(7);
                tc_fence_after();
                drain_partial_y(taddr + TF_COL_P, pw, 32, act);
                tc_fence_before();
                if (etid == 0) TF_PH(19);
                fence_proxy_async_global();   // the reducers pull the partials with a bulk copy (async proxy)
            }
            if (etid == 0) TF_PH(8);
            named_bar_sync(7, 256);    // every finaliser published; D3 is drained into `part`
            if (etid == 0) TF_PH(9);
            if (etid == 0) TF_SKEW(0);
            if (etid == 0) red_release_gpu_add(ctrA, 1u);   // release is cumulative over the barrier: one gpu-scope fence per CTA
            if (etid == 0) TF_TRACE(9);
            if (etid == 0) TF_PH(10);
            if (etid == 0) TF_SKEW(1);
            if (act) {   // off the critical path: outputs / saved activations that only later kernels read
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
        // ================= aux: drain of D3, y reduction + publication ===================================
        const int rt = threadIdx.x - 256;
        const int Q = L.Q;
        const int q_lo = c * Q;
        const int q_n = max(0, min(Q, B * 64 - q_lo));   // whole groups of 8: Q and 64 are multiples of 8
        const uint32_t taddr = tmem + ((uint32_t)((warp - 8) * 32) << 16) + TF_COL_P;
        const PartWalk pw = part_walk(a.part, c, G, Q, rt, 0);     // this thread drains outputs [0, 32) of row rt
        const uint32_t red_bytes = (uint32_t)(G * Q) * 4u;         // this reducer's block: [G CTAs][Q pairs]
        const float* red_src = a.part + (size_t)c * G * Q;
        const int yrep = a.yrep;
        // Reduction geometry (all of it fixed before the rounds: no division inside them).  Stage 1: a thread sums ONE
        // float4 column (4 consecutive pairs) over every nsub-th CTA; the columns fit one warp when there are <= 32 of
        // them, and then the 32 / ncol subsets a warp holds are combined by shuffles in fixed order; each warp group
        // writes its partial row to sPs.  Stage 2 (the publication threads) adds the rows in fixed order.
        const int ncol = q_n >> 2;                                   // <= 64 (Q <= 256)
        const int nwc = ncol > 32 ? 2 : 1;                           // warps side by side over the columns
        const int ncw = ncol > 32 ? 32 : ncol;
        const int spw = (nwc == 1 && ncw > 0) ? 32 / ncw : 1;        // subsets inside one warp
        const int nwg = 4 / nwc;                                     // warp groups = rows of sPs
        const int nsub = nwg * spw;
        const int wq = warp - 8, lane_sl = ncw > 0 ? lane / ncw : 0;
        const int col = (wq % nwc) * 32 + (ncw > 0 ? lane - lane_sl * ncw : 0);
        const int sid = (wq / nwc) * spw + lane_sl;                  // this thread's subset of the CTAs: sid, sid + nsub, ...
        const bool sum_act = ncol > 0 && lane_sl < spw && col < ncol;
        float* sProw = sPs + (size_t)(wq / nwc) * Q;
        // publication: thread (group g of 8 consecutive outputs of one row, replica rep0 + k * rep_step)
        const int ng = q_n >> 3;
        const int rep0 = ng > 0 ? rt / ng : 0, g = ng > 0 ? rt - rep0 * ng : 0;
        const int rep_step = ng > 0 ? min(yrep, max(1, 128 / ng)) : 1;   // publication threads per group
        const int pq0 = q_lo + 8 * g;
        const int pbb = pq0 >> 6, po = pq0 & 63;                     // row and first output of the group
        const bool pub_act = ng > 0 && rep0 < rep_step;
        float bias[8];   // b_o of this thread's group (zero on the padding outputs)
#pragma unroll
        for (int e = 0; e < 8; ++e) bias[e] = (pub_act && po + e < out) ? __ldg(f.bo + po + e) : 0.f;
        // round 0 publishes y_in; round r >= 1 reduces the partials of step r-1 into y_{r-1}
        for (int round = 0; round <= T; ++round) {
            const int t = round;   // trace row
            if (round > 0) {
                // drain D3 of step round-1 (outputs [0, 32) of this CTA's partial)
                mbar_wait(part_full, (uint32_t)(round - 1) & 1);
                tc_fence_after();
                drain_partial_y(taddr, pw, 0, rt < B);
                tc_fence_before();
                fence_proxy_async_global();   // generic stores -> the reducers' bulk copies (async proxy)
                if (rt == 0) TF_TRACE(26);
                named_bar_sync(7, 256);   // a full barrier, not an arrive: thread 0's release after it must cover these warps' stores to `part`
                if (rt == 0 && q_n > 0) {
                    spin_until_ge(ctrA, (unsigned)G * (unsigned)(round + 1), a.relaxed != 0);
                    TF_TRACE(20);
                    TF_PH(13);
                    // every CTA's partials of this reducer's pairs are one contiguous block: one bulk copy into sRed
                    mbar_expect_tx(red_full, red_bytes);
                    bulk_g2s(sRed, red_src, red_bytes, red_full);
                }
                if (q_n > 0) mbar_wait(red_full, (uint32_t)(round - 1) & 1);
                if (rt == 0) TF_PH(14);
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
                if (rt == 0) TF_TRACE(27);
                named_bar_sync(2, 128);
                if (rt == 0) TF_PH(16);
            }
            float* ydst = f.ys + (size_t)round * n_pairs;
            uint16_t* yx = a.yx + (size_t)(round & 1) * 2 * yx_part;
            float yv[8];
            if (pub_act) {
                if (round > 0) {
                    // stage 2: the warp groups' rows in fixed order
#pragma unroll
                    for (int e = 0; e < 8; e += 4) {
                        float4 v = *reinterpret_cast<const float4*>(sPs + 8 * g + e);
                        for (int k = 1; k < nwg; ++k) {
                            const float4 x = *reinterpret_cast<const float4*>(sPs + (size_t)k * Q + 8 * g + e);
                            v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
                        }
                        yv[e] = v.x; yv[e + 1] = v.y; yv[e + 2] = v.z; yv[e + 3] = v.w;
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) yv[e] = (po + e < out) ? yv[e] + bias[e] : 0.f;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) yv[e] = f.ys[(size_t)pbb * out + min(po + e, out - 1)];
#pragma unroll
                    for (int e = 0; e < 8; ++e) yv[e] = (po + e < out) ? yv[e] : 0.f;
                }
                uint4 hi, lo;
                split8_f16(yv, hi, lo);
                uint16_t* dst = yx + (size_t)(pbb >> 3) * 512 + (size_t)(po >> 3) * 64 + (size_t)(pbb & 7) * 8;
                for (int rep = rep0; rep < yrep; rep += rep_step) {
                    *reinterpret_cast<uint4*>(dst + (size_t)rep * yx_rep) = hi;
                    *reinterpret_cast<uint4*>(dst + (size_t)rep * yx_rep + yx_part) = lo;
                }
            }
            if (rt == 0) TF_TRACE(21);
            if (rt == 0) TF_SKEW(2);
            if (rt == 0) TF_PH(17);
            fence_proxy_async_global();
            named_bar_sync(2, 128);
            if (rt == 0) TF_PH(18);
            if (rt == 0) red_release_gpu_add(ctrB, 1u);
            if (rt == 0) TF_SKEW(3);
            if (rt == 0) TF_PH(15);
            if (round > 0 && pub_act && rep0 == 0) {   // the fp32 outputs: stored after the release (only later kernels read them)
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (po + e < out) ydst[(size_t)pbb * out + po + e] = yv[e];
            }
            if (rt == 0) TF_TRACE(22);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // no CTA leaves while a peer may still copy into its inbox
    if (a.trace && c == 0 && threadIdx.x == 0) a.trace[62] = clock64();
    if (warp == 2) tmem_dealloc<512>(tmem);
}

// ---- host side ----------------------------------------------------------------------------------------
bool gru_tc_shape_ok(int B, int H, int out) {
    return H % (TF_KC * TF_S) == 0 && H >= TF_KC * TF_S && out >= 1 && out <= 64 && B >= 1 && B <= 128;
}

// floats of the partial-sum buffer part[G reducers][G CTAs][Q]
static size_t tf_part_floats(int B, int H) {
    const size_t G = (size_t)H / 8;
    const size_t Q = 8 * (((size_t)8 * B + G - 1) / G);
    return round_up_sz(G * G * Q, 64);
}

size_t gru_tc_scratch_floats(int B, int H) {
    size_t MB = (B + 7) / 8;
    size_t hx = (size_t)2 * 2 * (H / TF_KC) * MB * 512 / 2;   // fp16 elements -> floats
    size_t yx = (size_t)TF_YREP * 2 * 2 * MB * 512 / 2;
    const size_t two_hop = round_up_sz(hx, 64) + round_up_sz(yx, 64) + 64 + tf_part_floats(B, H);
    const size_t one_hop = gru_tc2_scratch_floats(B, H);
    return two_hop > one_hop ? two_hop : one_hop;
}

// CVB_TC_FEEDBACK=grid keeps the two-exchange kernel of this file (A/B); default: the one-exchange kernel of gru_tc2.cu
int g_tc_hops[2] = {0, 0};
bool gru_tc_one_hop() {
    const char* e = getenv("CVB_TC_FEEDBACK");
    return !(e && e[0] == 'g');
}

// are all G/4 clusters co-resident at this shape?  (cached per shape)
static bool fwd_runnable(int B, int H, int out, const DeviceInfo& di, TfLayout* Lout) {
    const int G = H / 8;
    if (!gru_tc_shape_ok(B, H, out) || G > di.n_sm) return false;
    struct Entry { int B, H, out, ok; };
    static Entry cache[256];   // the row-count probe of cvb_recurrence_max_rows adds up to 16 entries per network shape
    static int n_cache = 0;
    int ok = -1;
    for (int i = 0; i < n_cache; ++i)
        if (cache[i].B == B && cache[i].H == H && cache[i].out == out) ok = cache[i].ok;
    TfLayout L = tf_layout(B, H, G, out, di.max_smem_optin);
    if (ok < 0) {
        ok = 0;
        if (L.NS >= 2 && (int)L.total <= di.max_smem_optin && (uint32_t)L.NS * L.stage_bytes >= (uint32_t)TF_S * L.slot_bytes &&
            cudaFuncSetAttribute(k_gru_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total) == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(G);
            cfg.blockDim = dim3(TF_NT);
            cfg.dynamicSmemBytes = L.total;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = TF_S;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int ncl = 0;
            if (cudaOccupancyMaxActiveClusters(&ncl, k_gru_fwd_tc, &cfg) == cudaSuccess) {
                if (getenv("CVB_DEBUG"))
                    fprintf(stderr, "[cvb] k_gru_fwd_tc: %d co-resident clusters of %d (need %d), smem %u, ring %d\n", ncl, TF_S, G / TF_S, L.total, L.NS);
                ok = ncl * TF_S >= G ? 1 : 0;
            }
        }
        cudaGetLastError();
        if (n_cache < 256) cache[n_cache++] = Entry{B, H, out, ok};
    }
    if (ok && Lout) *Lout = L;
    return ok != 0;
}

bool gru_tc_supported(int B, int H, int out, const DeviceInfo& di) { return fwd_runnable(B, H, out, di, nullptr); }

int gru_ar_fwd_tc(GruFwdArgs& f, float* tc_scratch, cudaStream_t s) {
    if (f.T <= 0 || f.B <= 0) return 0;
    DeviceInfo di;
    if (int rc = get_device_info(&di)) return rc;
    if (gru_tc_one_hop() && gru_tc2_supported(f.B, f.H, f.out, di)) {
        g_tc_hops[0] = 1;
        return gru_ar_fwd_tc2(f, tc_scratch, s);
    }
    g_tc_hops[0] = 2;
    TfLayout L;
    CVB_REQUIRE(fwd_runnable(f.B, f.H, f.out, di, &L), "gru_ar_fwd_tc: unsupported shape B=%d H=%d out=%d", f.B, f.H, f.out);
    GruTcArgs a;
    a.f = f;
    const size_t hx_f = round_up_sz((size_t)2 * 2 * (f.H / TF_KC) * L.MB * 512 / 2, 64);
    const size_t yx_f = round_up_sz((size_t)TF_YREP * 2 * 2 * L.MB * 512 / 2, 64);
    a.hx = reinterpret_cast<uint16_t*>(tc_scratch);
    a.yx = reinterpret_cast<uint16_t*>(tc_scratch + hx_f);
    a.ctr = reinterpret_cast<unsigned*>(tc_scratch + hx_f + yx_f);
    a.part = tc_scratch + hx_f + yx_f + 64;
    a.yrep = 2;
    if (const char* e = getenv("CVB_TC_YREP")) a.yrep = max(1, min(TF_YREP, atoi(e)));
    a.ymc = 0;   // measured: 18 440 -> 19 000 cycles per step with the multicast (at cluster size 4 a multicast saves no L2 traffic and adds a hop)
    if (const char* e = getenv("CVB_TC_YMC")) a.ymc = atoi(e) != 0;
    a.smem_max = di.max_smem_optin;
    a.keepalive = 1;
    a.relaxed = relaxed_polling() ? 1 : 0;
    if (const char* e = getenv("CVB_TC_KEEPALIVE")) a.keepalive = atoi(e) != 0;
    a.trace = nullptr;
    const char* trace_file = getenv("CVB_TRACE_FILE_FWD");
    const size_t trace_bytes = ((size_t)(f.T + 1) * 64 + 8 * 256 + 48 * 256) * sizeof(long long);
    if (trace_file && trace_file[0]) {
        CVB_CHECK(cudaMalloc(&a.trace, trace_bytes));
        CVB_CHECK(cudaMemsetAsync(a.trace, 0, trace_bytes, s));
    }
    CVB_CHECK(cudaMemsetAsync(a.yx, 0, (yx_f + 64) * sizeof(float), s));   // y padding columns + both counters
    CVB_CHECK(cudaFuncSetAttribute(k_gru_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(f.H / 8);
    cfg.blockDim = dim3(TF_NT);
    cfg.dynamicSmemBytes = L.total;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = TF_S;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = launch_without_coop() ? 1 : 2;
    prof_begin(s, CVB_PROF_GRU_FWD);
    CVB_CHECK(cudaLaunchKernelEx(&cfg, k_gru_fwd_tc, a));
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
