// Inference-only tensor-core forward recurrence of the autoregressive GRU (gru_vae.py:364-399 with do=False: the
// eval / stage-6 conversion path, decode_*.py:303-323), with the y feedback FOLDED into the recurrent matrix
// (SURVEY.md Appendix A.4):  without dropout o_t = h_t, so
//     W_y y_{t-1} = (W_y W_o) h_{t-1} + W_y b_o =: W_fb h_{t-1} + c_fb      (t >= 1)
// and one step needs only h_{t-1}:
//     a_r = gx_r + b_hr + (W_hr + W_fb,r) h      a_z likewise      gh_n = W_hn h + b_hn      gi_n = gx_n + W_fb,n h
// (c_fb rides in the bias of the caller's gx product; the first step's feedback is the caller's y_in, so
// gx[0] += W_y y_in - c_fb - W_fb h_in takes the folded terms back out there; all y_t come from ONE product
// Y = H W_o^T + b_o after the launch).  Per step there is a single exchange of h_t instead of two grid-wide ones,
// and it is not grid-wide either: arrival is counted per cluster (= per 32 units of h), a K stage of the next step is
// pulled as soon as the cluster(s) producing its units have arrived (see the producer warp).
//
// Same 2-D split over clusters of 4 CTAs as gru_tc.cu: cluster = 32 hidden units, CTA j = K-slice j of the
// contraction for the block's 4 x 32 rows [r' | z' | hn | in'], finaliser of 8 units; operands fp16 hi+lo with B
// stored [hi rows | lo rows] (one MMA with N = 256 and one with N = 128 per K step), fp32 accumulation in TMEM,
// DSMEM bulk-copy exchange of the partial sums, fixed-order fp32 sums, MUFU gates.
// Roles (256 threads): w0 bulk-copy producer, w1 MMA issuer, w2 TMEM allocator, w4-7 exchange + gates.
#include <stdlib.h>

#include "gru_ar.cuh"
#include "umma.cuh"

namespace cvb {
using namespace umma;

constexpr int TE_NT = 256;
constexpr int TE_KC = 64;
constexpr int TE_S = 4;
constexpr int TE_UB = 8 * TE_S;     // units per cluster
constexpr int TE_NW = 4 * TE_UB;    // rows of the folded operand: r', z', hn, in' of the block = 128
// TMEM columns [main0 | corrections | main1]: the hi x hi products of the first / second half of the K walk go to two
// accumulators (a tcgen05 accumulation chain truncates toward zero: none sums more than K = 128), the two cross products
// (lo planes scaled by 2^11, umma.cuh) share the middle one (tools/split_error_budget.py)
constexpr uint32_t TE_COL_M0 = 0, TE_COL_C = TE_NW, TE_COL_M1 = 2 * TE_NW, TE_COL_DUMMY = 3 * TE_NW;

struct TeLayout {
    int MB, nch, NS;
    int KCA, nsub;       // K extent of one ring stage of the h operand (64, or 32 when 64 leaves < 3 stages) and stages per step
    uint32_t half, stage_bytes, w_chunk_bytes, slot_bytes;
    uint32_t off_ring, off_w, off_inbox, off_bias, off_bar, total;
};

__host__ __device__ inline TeLayout te_layout(int B, int H, int smem_max) {
    TeLayout L;
    L.MB = (B + 7) / 8;
    L.nch = H / TE_KC / TE_S;
    L.w_chunk_bytes = 2u * (TE_NW / 8) * 1024u;   // [hi: 16 row groups][lo: 16 row groups] x 1 KB
    L.slot_bytes = (uint32_t)L.MB * 8u * 128u;    // [rows][32 floats]
    // the own partial sums stay in registers (TMEM lane == batch row on both sides): S-1 slots travel
    const uint32_t inbox = ((uint32_t)(TE_S - 1) * L.slot_bytes + 127u) & ~127u;
    const uint32_t fixed = (uint32_t)L.nch * L.w_chunk_bytes + inbox + 128u + 256u;
    L.KCA = 64;
    int ns = 0;
    for (;;) {
        L.half = (uint32_t)L.MB * (uint32_t)L.KCA * 16u;   // [MB row groups][KCA/8 k blocks][8 rows][8 k] fp16
        L.stage_bytes = 2u * L.half;
        ns = ((int)smem_max - (int)fixed) / (int)L.stage_bytes;
        if (ns >= 3 || L.KCA == 32) break;
        L.KCA = 32;
    }
    L.nsub = L.nch * TE_KC / L.KCA;
    L.NS = ns > 6 ? 6 : ns;
    const uint32_t ring = (uint32_t)(L.NS > 0 ? L.NS : 0) * L.stage_bytes;
    L.off_ring = 0;   // idle between a step's last chunk and the next step's first: doubles as the staging of the outgoing sums
    L.off_w = ring;
    L.off_inbox = L.off_w + (uint32_t)L.nch * L.w_chunk_bytes;
    L.off_bias = L.off_inbox + inbox;
    L.off_bar = L.off_bias + 128u;
    L.total = L.off_bar + 256u;
    return L;
}

struct GruTcEvalArgs {
    const float* gx;     // [T,B,3H]: W_x xc + b_ih + (t >= 1: W_y b_o; t == 0: W_y y_in)
    const float* Whh;    // [3H,H]
    const float* Wfb;    // [3H,H] = W_y W_o
    const float* bhh;    // [3H]
    float* hs;           // [T+1,B,H], slot 0 = h_in
    uint16_t* hx;        // [3 slots][2 parts][H/KCA chunks][MB][KCA/8 kblk][8 rows][8 k] fp16 (UMMA order) of h_t
    unsigned* ctr;       // [H/32] arrival counters, one per cluster; zero-initialised
    int B, T, H;
    int smem_max;
    int keepalive;
    int relaxed;         // CVB_TC_POLL=relaxed: poll the arrival counters with relaxed loads + one acquire fence instead of ld.acquire
    int rotate;          // CVB_TC_ROTATE (default 1): per-cluster start stage of the K walk
    long long* trace;    // optional [T+1][64] clock64 stamps of CTA 0 (CVB_TRACE_FILE_EVAL), else null
};

#define TE_TRACE(ev)                                                     \
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

__global__ void __launch_bounds__(TE_NT, 1) k_gru_fwd_tc_eval(GruTcEvalArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int B = a.B, T = a.T, H = a.H;
    const int G = gridDim.x, c = blockIdx.x;
    const int j = (int)cluster_ctarank();
    const TeLayout L = te_layout(B, H, a.smem_max);
    const int ublk0 = (c / TE_S) * TE_UB;
    const int u0 = ublk0 + 8 * j;
    const int k0 = j * L.nch * TE_KC;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    uint8_t* ring = smem + L.off_ring;
    float* stage = reinterpret_cast<float*>(ring);                 // [S-1 (to)][MB*8][32], aliases the (idle) ring
    uint8_t* sW = smem + L.off_w;
    float* inbox = reinterpret_cast<float*>(smem + L.off_inbox);   // [S-1 (from)][MB*8][32]
    float* sBh = reinterpret_cast<float*>(smem + L.off_bias);      // [3][8] b_hh of the own units
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* empty = full + 8;
    uint64_t* d1_full = empty + 8;
    uint64_t* inbox_full = d1_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(inbox_full + 1);
    const size_t hx_part = (size_t)H * L.MB * 8;
    const uint32_t kca_sh = L.KCA == 64 ? 6u : 5u;   // log2(KCA)
    // the 32 CTAs that own the same K-slice read the same stages of h: each cluster starts its walk over the slice at a
    // different stage so that they do not all pull the same L2 lines at the same moment (fixed per cluster: deterministic)
    // (rotation in units of the 64-wide weight chunks: a chunk's stages stay in the same half of the walk)
    const int per64 = TE_KC / L.KCA;
    const int rot = a.rotate ? ((c / TE_S) % L.nch) * per64 : 0;
    const int half1c = (L.nch + 1) / 2;        // walk position (in weight chunks) from which main1 accumulates
    const int half1 = half1c * per64;          // the same in ring stages
    const bool two_main = half1c < L.nch;

    // ---- one-time setup: folded weights -> fp16 hi/lo in UMMA K-major core-matrix order ------------------
    {
        // one item = 8 consecutive k of one operand row = one 16-byte core-matrix row (hi) + one (lo); consecutive
        // threads take consecutive rows: conflict-free 16-byte stores, full-sector global reads
        const int n_items = TE_NW * L.nch * (TE_KC / 8);
        for (int i = threadIdx.x; i < n_items; i += TE_NT) {
            const int kg = i / TE_NW, n = i - kg * TE_NW;   // n = g*32 + unit of the block; g: 0 r', 1 z', 2 hn, 3 in'
            const int g = n / TE_UB, ul = n - g * TE_UB;
            const int kl = kg * 8;
            const size_t at = (size_t)((g < 2 ? g : 2) * H + ublk0 + ul) * H + k0 + kl;
            float w[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) w[q] = 0.f;
            if (g < 3) {
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.Whh + at)), w1 = __ldg(reinterpret_cast<const float4*>(a.Whh + at) + 1);
                w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
            }
            if (g != 2) {
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.Wfb + at)), w1 = __ldg(reinterpret_cast<const float4*>(a.Wfb + at) + 1);
                w[0] += w0.x; w[1] += w0.y; w[2] += w0.z; w[3] += w0.w; w[4] += w1.x; w[5] += w1.y; w[6] += w1.z; w[7] += w1.w;
            }
            uint4 hi, lo;
            split8_f16(w, hi, lo);
            const uint32_t off = (uint32_t)(kl / TE_KC) * L.w_chunk_bytes + (uint32_t)(n >> 3) * 1024u + (uint32_t)((kl % TE_KC) >> 3) * 128u +
                                 (uint32_t)(n & 7) * 16u;
            // chunks walked in the second half are stored [lo rows | hi rows]: their stacked MMA starts at the correction
            // columns and runs on into main1
            int pos = kl / TE_KC - rot / per64;
            if (pos < 0) pos += L.nch;
            const bool swapped = pos >= half1c;
            *reinterpret_cast<uint4*>(sW + off + (swapped ? (TE_NW / 8) * 1024u : 0u)) = hi;
            *reinterpret_cast<uint4*>(sW + off + (swapped ? 0u : (TE_NW / 8) * 1024u)) = lo;
        }
        if (threadIdx.x < 24) sBh[threadIdx.x] = a.bhh[(threadIdx.x >> 3) * H + u0 + (threadIdx.x & 7)];
        if (threadIdx.x == 0) {
            for (int s = 0; s < 8; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            mbar_init(d1_full, 1);
            mbar_init(inbox_full, 1);
            mbar_fence_init();
        }
        fence_proxy_async_smem();
        if (warp == 2) tmem_alloc<512>(tmem_slot);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    cluster_sync_all();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ================= producer: K-slice of h_{t-1} ====================================================
        // One arrival counter per CLUSTER (= 32 units of h): a stage of the K walk is pulled as soon as the cluster(s) that
        // produce its units have published, so the late producers of a step hide behind the MMAs of the early stages.
        // The stages are still issued (and accumulated) in the fixed rotated order: the sums stay deterministic.
        int s = 0, slot = 0;
        uint32_t ph = 1;
        const int ncs = (H / TE_S) / TE_UB;           // clusters that produce this CTA's K-slice
        // lanes [0, ncs) watch the clusters that produce this CTA's K-slice, lane 31 the OWN cluster
        const unsigned* flag = lane == 31 ? a.ctr + c / TE_S : a.ctr + j * ncs + (lane < ncs ? lane : 0);
        const unsigned per_stage = (unsigned)(L.KCA / TE_UB);   // producing clusters per stage: 1 or 2
        const size_t sub_elems = (size_t)L.MB * L.KCA * 8;
        for (int t = 0; t < T; ++t) {
            const uint16_t* src = a.hx + (size_t)slot * 2 * hx_part + (size_t)(j * L.nsub) * sub_elems;
            const unsigned target = (unsigned)TE_S * (unsigned)(t + 1);
            int issued = 0;
            while (issued < L.nsub) {
                // the writers fenced generic -> async proxy before their release
                // The own cluster must have arrived too before the ring is written again: its CTAs arrive after their
                // inboxes completed, i.e. after this CTA's outgoing partial sums (staged in the ring) were delivered, and
                // after they finished reading the inbox the next exchange will overwrite.
                const bool polls = lane < ncs || lane == 31;
                const unsigned v = polls ? (a.relaxed ? ld_relaxed_gpu(flag) : ld_acquire_gpu(flag)) : 0u;
                const unsigned ready = __ballot_sync(0xffffffffu, polls && v >= target);
                if (!(ready >> 31)) continue;
                if (a.relaxed) fence_acq_rel_gpu();
                while (issued < L.nsub) {
                    int che = issued + rot;
                    if (che >= L.nsub) che -= L.nsub;
                    const unsigned need = ((1u << per_stage) - 1u) << ((unsigned)che * per_stage);
                    if ((ready & need) != need) break;
                    if (lane == 0) {
                        if (issued == 0) TE_TRACE(14);
                        mbar_wait(&empty[s], ph);
                        uint8_t* dst = ring + (size_t)s * L.stage_bytes;
                        mbar_expect_tx(&full[s], 2 * L.half);
                        bulk_g2s(dst, src + (size_t)che * sub_elems, L.half, &full[s]);
                        bulk_g2s(dst + L.half, src + hx_part + (size_t)che * sub_elems, L.half, &full[s]);
                    }
                    __syncwarp();
                    ++issued;
                    if (++s == L.NS) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
            if (++slot == 3) slot = 0;
        }
    } else if (warp == 1) {
        // ================= MMA issuer =========================================================================
        const uint32_t idesc_s = idesc_f16_f32(128, 2 * TE_NW), idesc_h = idesc_f16_f32(128, TE_NW), idesc_dummy = idesc_f16_f32(128, 16);
        const uint64_t dA0 = smem_desc(smem_u32(ring), 128, (uint32_t)L.KCA * 16u);
        const uint64_t dW0 = smem_desc(smem_u32(sW), 128, 1024);
        const uint32_t a_step = L.stage_bytes >> 4, half16 = L.half >> 4, w_step = L.w_chunk_bytes >> 4;
        const int kca16 = L.KCA >> 4;
        int s = 0;
        uint32_t ph = 0;
        for (int t = 0; t < T; ++t) {
            for (int ch = 0; ch < L.nsub; ++ch) {
                for (;;) {   // poll; while idle keep the tensor pipe warm (see gru_tc.cu)
                    uint32_t ok = (lane == 0) ? (mbar_test_wait(&full[s], ph) ? 1u : 0u) : 0u;
                    ok = __shfl_sync(0xffffffffu, ok, 0);
                    if (ok) break;
                    if (a.keepalive) mma_bf16_ss_elect(tmem + TE_COL_DUMMY, dW0, dW0, idesc_dummy, false);
                }
                if (lane == 0 && ch < 8) TE_TRACE(40 + ch);
                tc_fence_after();
                const uint64_t da = dA0 + (uint64_t)((uint32_t)s * a_step);
                // weights stay in 64-wide K chunks; stage ch covers k = [ch*KCA, +KCA) of the slice
                int che = ch + rot;
                if (che >= L.nsub) che -= L.nsub;
                const uint32_t kq = (uint32_t)che << kca_sh;
                const uint64_t db = dW0 + (uint64_t)((kq >> 6) * w_step + ((kq & 63u) >> 4) * 16u);
                const uint32_t w_half16 = (TE_NW / 8) * 1024u >> 4;   // hi rows -> lo rows (lo -> hi in a swapped chunk)
                if (ch < half1) {   // [hi | lo] rows: main0 and corrections side by side
                    for (int k16 = 0; k16 < kca16; ++k16) {
                        mma_bf16_ss_elect(tmem + TE_COL_M0, da + 16u * k16, db + 16u * k16, idesc_s, (ch | k16) != 0);
                        mma_bf16_ss_elect(tmem + TE_COL_C, da + half16 + 16u * k16, db + 16u * k16, idesc_h, true);
                    }
                } else {            // [lo | hi] rows: corrections and main1 side by side
                    for (int k16 = 0; k16 < kca16; ++k16) {
                        if (ch == half1 && k16 == 0) {   // main1 starts from zero while the corrections keep accumulating
                            mma_bf16_ss_elect(tmem + TE_COL_C, da, db, idesc_h, true);
                            mma_bf16_ss_elect(tmem + TE_COL_M1, da, db + w_half16, idesc_h, false);
                        } else {
                            mma_bf16_ss_elect(tmem + TE_COL_C, da + 16u * k16, db + 16u * k16, idesc_s, true);
                        }
                        mma_bf16_ss_elect(tmem + TE_COL_C, da + half16 + 16u * k16, db + w_half16 + 16u * k16, idesc_h, true);
                    }
                }
                mma_commit_elect(&empty[s]);
                if (ch == L.nsub - 1) mma_commit_elect(d1_full);
                if (++s == L.NS) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ================= exchange + gates: TMEM lane == batch row ==========================================
        const int b = (warp - 4) * 32 + lane;
        const bool act = b < B;
        const int etid = threadIdx.x - 128;
        const uint32_t inbox_addr = smem_u32(inbox);
        const uint32_t inbox_bar_addr = smem_u32(inbox_full);
        const uint32_t taddr = tmem + ((uint32_t)((warp - 4) * 32) << 16);
        const uint32_t slot_f = L.slot_bytes / 4;
        float hreg[8];
        {   // prologue: publish h_in (slot 0) in operand order
#pragma unroll
            for (int q = 0; q < 8; ++q) hreg[q] = act ? a.hs[(size_t)b * H + u0 + q] : 0.f;
            uint4 hh, hl;
            split8_f16(hreg, hh, hl);
            if (act) {
                const size_t off = (((size_t)(u0 >> kca_sh) * L.MB + (b >> 3)) << (kca_sh + 3)) + (size_t)((u0 & (L.KCA - 1)) >> 3) * 64 + (size_t)(b & 7) * 8;
                *reinterpret_cast<uint4*>(a.hx + off) = hh;
                *reinterpret_cast<uint4*>(a.hx + hx_part + off) = hl;
            }
            fence_proxy_async_global();
            named_bar_sync(1, 128);
            if (etid == 0) red_release_gpu_add(a.ctr + c / TE_S, 1u);
        }
        int wslot = 1;   // h_t goes to slot (t + 1) % 3: a reader may lag its writers by one step, never by two
        for (int t = 0; t < T; ++t) {
            const size_t row = (size_t)t * B + (act ? b : 0);
            float4 gxv[6];
            {   // every lane loads (the rows beyond B re-read row 0 of the frame: `row`), so the gate math below is branch-free
                const float* gp = a.gx + row * 3 * H + u0;
#pragma unroll
                for (int gi = 0; gi < 3; ++gi) {
                    gxv[2 * gi] = ldg_nc_v4_pinned(gp + (size_t)gi * H);
                    gxv[2 * gi + 1] = ldg_nc_v4_pinned(gp + (size_t)gi * H + 4);
                }
            }
            if (etid == 0) TE_TRACE(0);
            if (etid == 0) mbar_expect_tx(inbox_full, (uint32_t)(TE_S - 1) * L.slot_bytes);
            mbar_wait(d1_full, (uint32_t)t & 1);
            if (etid == 0) TE_TRACE(1);
            tc_fence_after();
            // partial sums (main + correction halves) of the block's units -> the finalisers' inboxes; the own ones stay here
            float own[32];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    float v[16], v2[16];
                    tmem_ld_x16(taddr + TE_COL_M0 + g * TE_UB + 16 * k, v);
                    tmem_ld_x16(taddr + TE_COL_C + g * TE_UB + 16 * k, v2);
                    if (two_main) {
                        float v3[16];
                        tmem_ld_x16(taddr + TE_COL_M1 + g * TE_UB + 16 * k, v3);
                        tmem_ld_wait();
#pragma unroll
                        for (int q = 0; q < 16; ++q) v[q] += v3[q];
                    } else {
                        tmem_ld_wait();
                    }
#pragma unroll
                    for (int q = 0; q < 16; ++q) v[q] = fmaf(v2[q], F16_LO_INV, v[q]);
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const int p = 2 * k + h2;   // destination CTA of the cluster
                        if (p == j) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) own[g * 8 + q] = v[8 * h2 + q];
                        } else if (b < L.MB * 8) {
                            // a row is 8 x 16 B; slot q of row b sits at q ^ (b & 7) so that a quarter warp covers all 32 banks
                            float* row = stage + (size_t)(p - (p > j ? 1 : 0)) * slot_f + b * 32;
                            float* d = row + (((2 * g) ^ (b & 7)) << 2);
                            float* d1 = row + (((2 * g + 1) ^ (b & 7)) << 2);
                            *reinterpret_cast<float4*>(d) = make_float4(v[8 * h2 + 0], v[8 * h2 + 1], v[8 * h2 + 2], v[8 * h2 + 3]);
                            *reinterpret_cast<float4*>(d1) = make_float4(v[8 * h2 + 4], v[8 * h2 + 5], v[8 * h2 + 6], v[8 * h2 + 7]);
                        }
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            named_bar_sync(3, 128);
            if (etid < TE_S && etid != j)   // slot order on both sides: the other CTAs by rank
                bulk_s2c(mapa(inbox_addr + (uint32_t)(j - (j > etid ? 1 : 0)) * L.slot_bytes, (uint32_t)etid),
                         stage + (size_t)(etid - (etid > j ? 1 : 0)) * slot_f, L.slot_bytes, mapa(inbox_bar_addr, (uint32_t)etid));
            if (etid == 0) TE_TRACE(2);
            mbar_wait_cluster(inbox_full, (uint32_t)t & 1);
            if (etid == 0) TE_TRACE(3);
            float ar[8], az[8], ahn[8], ain[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                ar[q] = own[q];
                az[q] = own[8 + q];
                ahn[q] = own[16 + q];
                ain[q] = own[24 + q];
            }
            {
                const int bs = b < L.MB * 8 ? b : 0;   // rows beyond the staged row groups read row 0 (in bounds)
#pragma unroll
                for (int p = 0; p < TE_S - 1; ++p) {   // fixed order: deterministic
                    const float4* x = reinterpret_cast<const float4*>(inbox + (size_t)p * slot_f + bs * 32);
                    const int sw = bs & 7;
                    const float4 x0 = x[0 ^ sw], x1 = x[1 ^ sw], x2 = x[2 ^ sw], x3 = x[3 ^ sw], x4 = x[4 ^ sw], x5 = x[5 ^ sw], x6 = x[6 ^ sw],
                                 x7 = x[7 ^ sw];
                    ar[0] += x0.x; ar[1] += x0.y; ar[2] += x0.z; ar[3] += x0.w; ar[4] += x1.x; ar[5] += x1.y; ar[6] += x1.z; ar[7] += x1.w;
                    az[0] += x2.x; az[1] += x2.y; az[2] += x2.z; az[3] += x2.w; az[4] += x3.x; az[5] += x3.y; az[6] += x3.z; az[7] += x3.w;
                    ahn[0] += x4.x; ahn[1] += x4.y; ahn[2] += x4.z; ahn[3] += x4.w; ahn[4] += x5.x; ahn[5] += x5.y; ahn[6] += x5.z; ahn[7] += x5.w;
                    ain[0] += x6.x; ain[1] += x6.y; ain[2] += x6.z; ain[3] += x6.w; ain[4] += x7.x; ain[5] += x7.y; ain[6] += x7.z; ain[7] += x7.w;
                }
            }
            if (etid == 0) TE_TRACE(5);
            {
                const float* gxr = reinterpret_cast<const float*>(&gxv[0]);
                const float* gxz = reinterpret_cast<const float*>(&gxv[2]);
                const float* gxn = reinterpret_cast<const float*>(&gxv[4]);
#pragma unroll
                // straight-line for every lane, active row or not (rows beyond B hold garbage that is never stored): inside an
                // `if (act)` per unit the compiler kept eight branch regions and the eight ex2 -> rcp -> ex2 -> rcp chains ran one
                // after the other (gru_tc.cu: 2500 -> 1240 cycles per step)
                float bh[24];
#pragma unroll
                for (int q = 0; q < 6; ++q) *reinterpret_cast<float4*>(bh + 4 * q) = *reinterpret_cast<const float4*>(sBh + 4 * q);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float r = sigmoid_fast(gxr[q] + ar[q] + bh[q]);
                    const float z = sigmoid_fast(gxz[q] + az[q] + bh[8 + q]);
                    const float n = tanh_fast(gxn[q] + ain[q] + r * (ahn[q] + bh[16 + q]));
                    hreg[q] = (1.0f - z) * n + z * hreg[q];
                }
            }
            if (etid == 0) TE_TRACE(10);
            uint4 hh, hl;
            split8_f16(hreg, hh, hl);
            if (act) {
                uint16_t* hdst = a.hx + (size_t)wslot * 2 * hx_part;
                const size_t off = (((size_t)(u0 >> kca_sh) * L.MB + (b >> 3)) << (kca_sh + 3)) + (size_t)((u0 & (L.KCA - 1)) >> 3) * 64 + (size_t)(b & 7) * 8;
                *reinterpret_cast<uint4*>(hdst + off) = hh;
                *reinterpret_cast<uint4*>(hdst + hx_part + off) = hl;
            }
            if (etid == 0) TE_TRACE(7);
            fence_proxy_async_global();
            if (etid == 0) TE_TRACE(8);
            named_bar_sync(1, 128);
            if (etid == 0) red_release_gpu_add(a.ctr + c / TE_S, 1u);
            if (++wslot == 3) wslot = 0;
            if (etid == 0) TE_TRACE(9);
            // gate inputs of frame t+2 -> L2 (this CTA's 1/G of the frame; every CTA reads 96 bytes of every row): as HBM misses
            // issued by all CTAs at the step top they slow the h pulls of that moment (gru_tc2_bwd.cu)
            if (etid == 32 && t + 2 < T) {
                const size_t frame = (size_t)B * 3 * H, slice = frame / (size_t)G;   // 24 B floats: a multiple of 16 bytes
                bulk_prefetch_l2(a.gx + (size_t)(t + 2) * frame + (size_t)c * slice, (uint32_t)(slice * sizeof(float)));
            }
            if (act) {   // off the critical path: the state trajectory (the y product after the launch reads it)
                float* hd = a.hs + (size_t)(t + 1) * B * H + (size_t)b * H + u0;
                *reinterpret_cast<float4*>(hd) = make_float4(hreg[0], hreg[1], hreg[2], hreg[3]);
                *reinterpret_cast<float4*>(hd + 4) = make_float4(hreg[4], hreg[5], hreg[6], hreg[7]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc<512>(tmem);
}

// ---- host side ----------------------------------------------------------------------------------------
// host-only shape test (no device query): what the scratch sizing keys on -- any out_dim (y comes from a GEMM after the launch)
bool gru_tc_eval_shape_ok(int B, int H) { return H % (TE_KC * TE_S) == 0 && H >= TE_KC * TE_S && B >= 1 && B <= 128 && H <= 3968; }

static bool eval_runnable(int B, int H, const DeviceInfo& di, TeLayout* Lout) {
    const int G = H / 8;
    if (!gru_tc_eval_shape_ok(B, H) || G > di.n_sm) return false;
    struct Entry { int B, H, ok; };
    static Entry cache[256];   // the row-count probe of cvb_recurrence_max_rows adds up to 16 entries per network shape
    static int n_cache = 0;
    int ok = -1;
    for (int i = 0; i < n_cache; ++i)
        if (cache[i].B == B && cache[i].H == H) ok = cache[i].ok;
    TeLayout L = te_layout(B, H, di.max_smem_optin);
    if (ok < 0) {
        ok = 0;
        if (L.NS >= 2 && (int)L.total <= di.max_smem_optin && (uint32_t)L.NS * L.stage_bytes >= (uint32_t)(TE_S - 1) * L.slot_bytes &&
            cudaFuncSetAttribute(k_gru_fwd_tc_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total) == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(G);
            cfg.blockDim = dim3(TE_NT);
            cfg.dynamicSmemBytes = L.total;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = TE_S;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int ncl = 0;
            if (cudaOccupancyMaxActiveClusters(&ncl, k_gru_fwd_tc_eval, &cfg) == cudaSuccess) ok = ncl * TE_S >= G ? 1 : 0;
            if (getenv("CVB_DEBUG")) fprintf(stderr, "[cvb] k_gru_fwd_tc_eval: %d co-resident clusters of %d (need %d), smem %u, ring %d\n", ncl, TE_S, G / TE_S, L.total, L.NS);
        } else if (getenv("CVB_DEBUG")) {
            fprintf(stderr, "[cvb] k_gru_fwd_tc_eval: not runnable at B=%d H=%d: ring %d stages, smem %u of %d\n", B, H, L.NS, L.total, di.max_smem_optin);
        }
        cudaGetLastError();
        if (n_cache < 256) cache[n_cache++] = Entry{B, H, ok};
    }
    if (ok && Lout) *Lout = L;
    return ok != 0;
}

bool gru_tc_eval_supported(int B, int H, int out, const DeviceInfo& di) { return out >= 1 && eval_runnable(B, H, di, nullptr); }

// scratch (floats): W_fb [3H,H] | hx | counter
size_t gru_tc_eval_scratch_floats(int B, int H) {
    size_t MB = (B + 7) / 8;
    size_t hx = (size_t)3 * 2 * H * MB * 8 / 2;
    return round_up_sz((size_t)3 * H * H, 64) + round_up_sz(hx, 64) + 128;
}

__global__ void k_add_rowvec(float* __restrict__ dst, int rows, int cols, float alpha, const float* __restrict__ v) {
    for (int col = blockIdx.x * blockDim.x + threadIdx.x; col < cols; col += gridDim.x * blockDim.x) {
        const float x = alpha * v[col];
        for (int r = blockIdx.y; r < rows; r += gridDim.y) dst[(size_t)r * cols + col] += x;
    }
}

// Weight-only prologue of the folded inference path, BEFORE the input-side product:  W_fb = W_y W_o (scratch),
// cfb[0..3H) = c_fb = W_y b_o,  cfb[3H..6H) = b_ih + c_fb = the bias the caller gives its gx product so that every row
// of gx already carries c_fb (no extra pass over the [T,B,3H] buffer).  cuBLAS fp32, exact products.
int gru_tc_eval_prepare(const GruFwdArgs& f, const float* bih, float* scratch, float* cfb, cudaStream_t s) {
    const int H = f.H, out = f.out;
    if (int rc = gemm_rm(s, false, false, 3 * H, H, out, 1.f, f.Wy, f.ldwy, f.Wo, H, 0.f, scratch, H)) return rc;
    if (int rc = gemm_rm(s, false, false, 3 * H, 1, out, 1.f, f.Wy, f.ldwy, f.bo, 1, 0.f, cfb, 1)) return rc;
    CVB_CHECK(cudaMemcpyAsync(cfb + 3 * H, bih, (size_t)3 * H * sizeof(float), cudaMemcpyDeviceToDevice, s));
    k_add_rowvec<<<dim3(ceil_div_sz((size_t)3 * H, 256), 1), 256, 0, s>>>(cfb + 3 * H, 1, 3 * H, 1.f, cfb);
    CVB_LAUNCH_CHECK();
    return 0;
}

// Inference forward of the recurrence with the folded feedback, after gru_tc_eval_prepare and the caller's
// gx = W_x xc + (b_ih + c_fb); gx[0] is modified in place.  Fills hs[1..T] and ys[1..T].
int gru_ar_fwd_tc_eval(GruFwdArgs& f, float* scratch, const float* cfb, cudaStream_t s) {
    if (f.T <= 0 || f.B <= 0) return 0;
    DeviceInfo di;
    if (int rc = get_device_info(&di)) return rc;
    TeLayout L;
    CVB_REQUIRE(eval_runnable(f.B, f.H, di, &L), "gru_ar_fwd_tc_eval: unsupported shape B=%d H=%d", f.B, f.H);
    const int B = f.B, T = f.T, H = f.H, out = f.out;
    float* Wfb = scratch;
    const size_t wfb_f = round_up_sz((size_t)3 * H * H, 64);
    const size_t hx_f = round_up_sz((size_t)3 * 2 * H * L.MB * 8 / 2, 64);
    // the first step's feedback is the CALLER's y_in, not W_o h_in + b_o:  gx[0] += W_y y_in - c_fb - W_fb h_in  takes
    // the folded terms (the bias above, the product the kernel will add) back out
    float* gx0 = const_cast<float*>(f.gx);
    if (int rc = gemm_rm(s, false, true, B, 3 * H, out, 1.f, f.ys, out, f.Wy, f.ldwy, 1.f, gx0, 3 * H)) return rc;
    if (int rc = gemm_rm(s, false, true, B, 3 * H, H, -1.f, f.hs, H, Wfb, H, 1.f, gx0, 3 * H)) return rc;
    k_add_rowvec<<<dim3(ceil_div_sz((size_t)3 * H, 256), B < 16 ? B : 16), 256, 0, s>>>(gx0, B, 3 * H, -1.f, cfb);
    CVB_LAUNCH_CHECK();
    GruTcEvalArgs a;
    a.gx = f.gx;
    a.Whh = f.Whh;
    a.Wfb = Wfb;
    a.bhh = f.bhh;
    a.hs = f.hs;
    a.hx = reinterpret_cast<uint16_t*>(scratch + wfb_f);
    a.ctr = reinterpret_cast<unsigned*>(scratch + wfb_f + hx_f);
    a.B = B;
    a.T = T;
    a.H = H;
    a.smem_max = di.max_smem_optin;
    a.keepalive = 1;
    a.relaxed = relaxed_polling() ? 1 : 0;
    if (const char* e = getenv("CVB_TC_KEEPALIVE")) a.keepalive = atoi(e) != 0;
    a.rotate = 1;
    if (const char* e = getenv("CVB_TC_ROTATE")) a.rotate = atoi(e) != 0;
    a.trace = nullptr;
    const char* trace_file = getenv("CVB_TRACE_FILE_EVAL");
    const size_t trace_bytes = (size_t)(T + 1) * 64 * sizeof(long long);
    if (trace_file && trace_file[0]) {
        CVB_CHECK(cudaMalloc(&a.trace, trace_bytes));
        CVB_CHECK(cudaMemsetAsync(a.trace, 0, trace_bytes, s));
    }
    CVB_CHECK(cudaMemsetAsync(a.ctr, 0, 128 * sizeof(float), s));
    CVB_CHECK(cudaFuncSetAttribute(k_gru_fwd_tc_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(H / 8);
    cfg.blockDim = dim3(TE_NT);
    cfg.dynamicSmemBytes = L.total;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = TE_S;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = launch_without_coop() ? 1 : 2;
    prof_begin(s, CVB_PROF_GRU_FWD);
    CVB_CHECK(cudaLaunchKernelEx(&cfg, k_gru_fwd_tc_eval, a));
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
    // all outputs at once: ys[1..T] = hs[1..T] W_o^T + b_o
    if (want_tc_gemm() && gemm_tc_eligible(T * B, out, H))
        return gemm_tc(s, false, true, T * B, out, H, f.hs + (size_t)B * H, H, f.Wo, H, false, f.bo, f.ys + (size_t)B * out, out, true);
    if (int rc = fill_rows(s, f.ys + (size_t)B * out, (size_t)T * B, out, out, f.bo)) return rc;
    if (int rc = gemm_rm(s, false, true, T * B, out, H, 1.f, f.hs + (size_t)B * H, H, f.Wo, H, 1.f, f.ys + (size_t)B * out, out)) return rc;
    return 0;
}

}  // namespace cvb
