// Tensor-core (tcgen05) BPTT of the autoregressive GRU (the reverse of
// gru_vae.py:364-399), one persistent cooperative launch for all T steps.
//
// 2-D work split over thread-block clusters of S CTAs (S = 8, or 4): cluster i owns the block of
// 8*S hidden units [8*S*i, 8*S*(i+1)); CTA j of the cluster owns the K-slice [j*3H/S, (j+1)*3H/S) of the
// contraction  dh[b,u] += sum_k dgh_{t+1}[b,k] W_hh[k,u]  for ALL units of the block, and FINALISES the
// 8 units [8*S*i + 8j, +8).  W_hh^T of (block x K-slice) stays in shared memory for the whole sequence as
// bf16 hi+lo (x = hi + lo + O(2^-17 x); three MMAs hi*hi + lo*hi + hi*lo, fp32 accumulate in TMEM), so a
// CTA ingests only 1/S of the all-gathered dgh_{t+1} per step (the 1-D split made every CTA read all of it).
// The S partial accumulators of a unit meet in the finaliser's shared memory through DSMEM stores
// (st.shared::cluster + cluster-scope mbarrier) and are summed in fixed order (deterministic).
//
// The y feedback (dy_t = dY_t + dgi_{t+1} W_y, then dh_t += (dy_t W_o) * m_t) is a reduction over all of 3H:
// each finaliser writes its partial [B,out], the per-pair sums are formed in fixed order by the "aux" warps
// of the CTA that owns the pair, and q = dy_t W_o[:, own units] is recomputed by every CTA's aux warps.
//
// Roles (384 threads): w0 bulk-copy producer (lane 0), w1 MMA issuer (lane 0), w2 TMEM allocator,
// w4-7 exchange + finalise (TMEM lane == batch row), w8-11 aux (dy reduction, q, half of the partial).
// Grid-wide ordering is two monotonic counters: A (dgh_t / partials published), B (dy_t complete).
#include <stdlib.h>

#include "gru_ar.cuh"
#include "umma.cuh"

namespace cvb {
using namespace umma;

constexpr int TB_NT = 384;
constexpr int TB_KC = 64;   // K per ring stage

struct TbLayout {
    int MB, S, Ublk, nch, NS;
    uint32_t half, stage_bytes, w_part_bytes, slot_bytes;
    uint32_t off_ring, off_w, off_inbox, off_stage, off_wo, off_wy, off_aux, off_bar, total;
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
    const int Q = (B * out + G - 1) / G;
    uint32_t aux = (uint32_t)(G * Q) * 4u;                      // staging of the partials being reduced
    const uint32_t aux2 = 128u * 24u * 4u + 128u * 8u * 4u;     // sG [128][24] + sQ [128][8]
    if (aux < aux2) aux = aux2;
    aux = (aux + 127u) & ~127u;
    const uint32_t fixed = 2u * L.w_part_bytes + 2u * inbox + 64u * 8u * 4u + 24u * 64u * 4u + aux + 256u;
    int ns = ((int)smem_max - (int)fixed) / (int)L.stage_bytes;
    L.NS = ns > 6 ? 6 : ns;
    const uint32_t ring = (uint32_t)(L.NS > 0 ? L.NS : 0) * L.stage_bytes;
    L.off_ring = 0;
    L.off_w = ring;
    L.off_inbox = L.off_w + 2u * L.w_part_bytes;
    L.off_stage = L.off_inbox + inbox;
    L.off_wo = L.off_stage + inbox;
    L.off_wy = L.off_wo + 64u * 8u * 4u;
    L.off_aux = L.off_wy + 24u * 64u * 4u;
    L.off_bar = L.off_aux + aux;
    L.total = L.off_bar + 256u;
    return L;
}

struct GruTcBwdArgs {
    GruBwdArgs f;
    uint16_t* gxh;    // [2 slots][2 parts][3H/64 chunks][MB][8 kblk][8 rows][8 k] bf16 (UMMA order) of dgh_t
    unsigned* ctr;    // [0] = A, [32] = B (separate 128-B lines), zero-initialised
    int S;
    int smem_max;
    long long* trace;   // optional [T+1][64] clock64 stamps of CTA 0 (CVB_TRACE_FILE), else null
};

// per-step phase stamps of one CTA (profiling hook; one predictable branch when disabled)
#define TB_TRACE(ev)                                                     \
    do {                                                                 \
        if (a.trace && c == 0) a.trace[(size_t)n * 64 + (ev)] = clock64(); \
    } while (0)

static __device__ __forceinline__ void spin_until(const unsigned* ctr, unsigned target) {
    while (ld_acquire_gpu(ctr) < target) {
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
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* ring = smem + L.off_ring;
    uint8_t* sW = smem + L.off_w;
    float* inbox = reinterpret_cast<float*>(smem + L.off_inbox);   // [S (from)][MB*8][8]
    float* stage = reinterpret_cast<float*>(smem + L.off_stage);   // [S (to)][MB*8][8]
    float* sWo = reinterpret_cast<float*>(smem + L.off_wo);        // [64][8]
    float* sWy = reinterpret_cast<float*>(smem + L.off_wy);        // [24][64]
    float* sG = reinterpret_cast<float*>(smem + L.off_aux);        // [128][24]
    float* sQ = sG + 128 * 24;                                     // [128][8]
    float* sRed = sG;                                              // [G][Q] (aliases sG/sQ, see the step order)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* empty = full + 8;
    uint64_t* accum_full = empty + 8;
    uint64_t* inbox_full = accum_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(inbox_full + 1);
    const size_t gx_part = (size_t)(K3 / TB_KC) * L.MB * 512;   // elements per part
    unsigned* ctrA = a.ctr;
    unsigned* ctrB = a.ctr + 32;
    const int n_pairs = B * out;

    // ---- one-time setup --------------------------------------------------------------------------
    {
        const int Kr = L.nch * TB_KC;
        for (int i = threadIdx.x; i < L.Ublk * Kr; i += TB_NT) {
            const int n = i % L.Ublk, kl = i / L.Ublk;
            const float w = f.Whh[(size_t)(k0 + kl) * H + ublk0 + n];
            uint16_t hi, lo;
            split_bf16(w, hi, lo);
            const uint32_t off = (uint32_t)(kl / TB_KC) * ((uint32_t)S * 1024u) + (uint32_t)(n >> 3) * 1024u +
                                 (uint32_t)((kl % TB_KC) >> 3) * 128u + (uint32_t)(n & 7) * 16u + (uint32_t)(kl & 7) * 2u;
            *reinterpret_cast<uint16_t*>(sW + off) = hi;
            *reinterpret_cast<uint16_t*>(sW + L.w_part_bytes + off) = lo;
        }
        for (int i = threadIdx.x; i < 64 * 8; i += TB_NT) {
            const int o = i >> 3, uu = i & 7;
            sWo[i] = (o < out) ? f.Wo[(size_t)o * H + u0 + uu] : 0.f;
        }
        for (int i = threadIdx.x; i < 24 * 64; i += TB_NT) {
            const int r = i >> 6, o = i & 63;
            const int g = r >> 3, uu = r & 7;
            sWy[i] = (o < out) ? f.Wy[(size_t)(g * H + u0 + uu) * f.ldwy + o] : 0.f;
        }
        if (threadIdx.x == 0) {
            for (int s = 0; s < 8; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            mbar_init(accum_full, 1);
            mbar_init(inbox_full, 1);   // armed with expect_tx(S slots) every step; peers' bulk copies complete_tx
            mbar_fence_init();
        }
        fence_proxy_async_smem();
        if (warp == 2) tmem_alloc<64>(tmem_slot);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    cluster_sync_all();   // every CTA's inbox barrier is initialised before any remote arrive
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ================= producer: this CTA's K-slice of dgh_{t+1}, chunk by chunk ==================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 1;   // parity to wait on the empty barrier of stage s (first pass: free)
            for (int n = 1; n <= T; ++n) {
                const uint16_t* src = a.gxh + (size_t)((n - 1) & 1) * 2 * gx_part + (size_t)(j * L.nch) * L.MB * 512;
                spin_until(ctrA, (unsigned)G * (unsigned)n);
                TB_TRACE(14);
                fence_proxy_async_all();
                for (int ch = 0; ch < L.nch; ++ch) {
                    mbar_wait(&empty[s], ph);
                    if (ch == L.nch - 1) TB_TRACE(15);
                    if (ch < 8) TB_TRACE(32 + ch);
                    uint8_t* dst = ring + (size_t)s * L.stage_bytes;
                    mbar_expect_tx(&full[s], 2 * L.half);
                    bulk_g2s(dst, src + (size_t)ch * L.MB * 512, L.half, &full[s]);
                    bulk_g2s(dst + L.half, src + gx_part + (size_t)ch * L.MB * 512, L.half, &full[s]);
                    if (++s == L.NS) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer ==================================================================
        if (lane == 0) {
            const uint32_t idesc = idesc_bf16_f32(128, L.Ublk);
            int s = 0;
            uint32_t ph = 0;
            for (int n = 1; n <= T; ++n) {
                for (int ch = 0; ch < L.nch; ++ch) {
                    mbar_wait(&full[s], ph);
                    if (ch < 8) TB_TRACE(40 + ch);
                    if (ch == 0) TB_TRACE(16);
                    if (ch == L.nch - 1) TB_TRACE(17);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(ring + (size_t)s * L.stage_bytes);
                    const uint32_t a_lo = a_hi + L.half;
                    const uint32_t b_hi = smem_u32(sW) + (uint32_t)ch * ((uint32_t)S * 1024u);
                    const uint32_t b_lo = b_hi + L.w_part_bytes;
#pragma unroll
                    for (int k16 = 0; k16 < TB_KC / 16; ++k16) {
                        const uint64_t dah = smem_desc(a_hi + k16 * 256, 128, 1024);
                        const uint64_t dal = smem_desc(a_lo + k16 * 256, 128, 1024);
                        const uint64_t dbh = smem_desc(b_hi + k16 * 256, 128, 1024);
                        const uint64_t dbl = smem_desc(b_lo + k16 * 256, 128, 1024);
                        mma_bf16_ss(tmem, dah, dbh, idesc, (ch | k16) != 0);
                        mma_bf16_ss(tmem, dal, dbh, idesc, true);
                        mma_bf16_ss(tmem, dah, dbl, idesc, true);
                    }
                    mma_commit(&empty[s]);
                    if (ch == L.nch - 1) mma_commit(accum_full);
                    if (++s == L.NS) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ================= exchange + finalise: TMEM lane == batch row ===================================
        const int b = (warp - 4) * 32 + lane;
        const bool act = b < B;
        const int etid = threadIdx.x - 128;
        const uint32_t inbox_addr = smem_u32(inbox);
        const uint32_t inbox_bar_addr = smem_u32(inbox_full);
        float carry[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) carry[q] = act ? f.dhc[(size_t)b * H + u0 + q] : 0.f;
        for (int n = 0; n <= T; ++n) {
            const int t = T - 1 - n;
            const size_t row = (size_t)(t < 0 ? 0 : t) * B + (act ? b : 0);
            float4 pr[2], pz[2], pn[2], pg[2], ph[2];
            float4 pm[2] = {make_float4(1.f, 1.f, 1.f, 1.f), make_float4(1.f, 1.f, 1.f, 1.f)};
            if (t >= 0 && act) {
                const size_t so = row * H + u0;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    pr[q] = *reinterpret_cast<const float4*>(f.sv_r + so + 4 * q);
                    pz[q] = *reinterpret_cast<const float4*>(f.sv_z + so + 4 * q);
                    pn[q] = *reinterpret_cast<const float4*>(f.sv_n + so + 4 * q);
                    pg[q] = *reinterpret_cast<const float4*>(f.sv_ghn + so + 4 * q);
                    ph[q] = *reinterpret_cast<const float4*>(f.hs + so + 4 * q);   // hs slot t = h_{t-1}
                    if (f.mask) pm[q] = *reinterpret_cast<const float4*>(f.mask + so + 4 * q);
                }
            }
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (etid == 0) TB_TRACE(0);
            if (n > 0) {
                if (etid == 0) mbar_expect_tx(inbox_full, (uint32_t)S * L.slot_bytes);
                mbar_wait(accum_full, (n - 1) & 1);
                if (etid == 0) TB_TRACE(1);
                tc_fence_after();
                const uint32_t taddr = tmem + ((uint32_t)((warp - 4) * 32) << 16);
                for (int k = 0; k < L.Ublk / 16; ++k) {
                    float v[16];
                    tmem_ld_x16(taddr + 16 * k, v);
                    tmem_ld_wait();
                    if (b < L.MB * 8) {
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            float* d = stage + (size_t)(2 * k + h2) * (L.slot_bytes / 4) + b * 8;
                            *reinterpret_cast<float4*>(d) = make_float4(v[8 * h2 + 0], v[8 * h2 + 1], v[8 * h2 + 2], v[8 * h2 + 3]);
                            *reinterpret_cast<float4*>(d + 4) = make_float4(v[8 * h2 + 4], v[8 * h2 + 5], v[8 * h2 + 6], v[8 * h2 + 7]);
                        }
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();
                named_bar_sync(3, 128);
                // partial sums of peer p's units -> slot j of p's inbox (bulk DSMEM copy, complete_tx on p's barrier)
                if (etid < S)
                    bulk_s2c(mapa(inbox_addr + (uint32_t)j * L.slot_bytes, (uint32_t)etid), stage + (size_t)etid * (L.slot_bytes / 4),
                             L.slot_bytes, mapa(inbox_bar_addr, (uint32_t)etid));
                if (etid == 0) TB_TRACE(2);
                mbar_wait_cluster(inbox_full, (n - 1) & 1);
                if (etid == 0) TB_TRACE(3);
                if (act) {
                    for (int p = 0; p < S; ++p) {
                        const float4 x0 = *reinterpret_cast<const float4*>(inbox + (size_t)p * (L.slot_bytes / 4) + b * 8);
                        const float4 x1 = *reinterpret_cast<const float4*>(inbox + (size_t)p * (L.slot_bytes / 4) + b * 8 + 4);
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
            if (etid == 0) TB_TRACE(4);
            named_bar_sync(5, 256);   // q = dy_t W_o[:, own units] is in sQ
            if (etid == 0) TB_TRACE(5);
            float dgr[8], dgz[8], dgn[8], dgnr[8];
            {
                const float4 q0 = *reinterpret_cast<const float4*>(sQ + b * 8);
                const float4 q1 = *reinterpret_cast<const float4*>(sQ + b * 8 + 4);
                const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                const float* r_ = reinterpret_cast<const float*>(pr);
                const float* z_ = reinterpret_cast<const float*>(pz);
                const float* n_ = reinterpret_cast<const float*>(pn);
                const float* g_ = reinterpret_cast<const float*>(pg);
                const float* h_ = reinterpret_cast<const float*>(ph);
                const float* m_ = reinterpret_cast<const float*>(pm);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (act) {
                        const float dh = carry[q] + acc[q] + qv[q] * m_[q];
                        const float r = r_[q], z = z_[q], nn = n_[q];
                        const float dn = dh * (1.0f - z);
                        const float dz = dh * (h_[q] - nn);
                        carry[q] = dh * z;
                        const float dan = dn * (1.0f - nn * nn);
                        dgr[q] = dan * g_[q] * r * (1.0f - r);
                        dgz[q] = dz * z * (1.0f - z);
                        dgn[q] = dan;
                        dgnr[q] = dan * r;
                    } else {
                        dgr[q] = dgz[q] = dgn[q] = dgnr[q] = 0.f;
                    }
                }
            }
            // dgi of the own units -> aux warps (their half of the partial)
            {
                float* g = sG + b * 24;
                *reinterpret_cast<float4*>(g) = make_float4(dgr[0], dgr[1], dgr[2], dgr[3]);
                *reinterpret_cast<float4*>(g + 4) = make_float4(dgr[4], dgr[5], dgr[6], dgr[7]);
                *reinterpret_cast<float4*>(g + 8) = make_float4(dgz[0], dgz[1], dgz[2], dgz[3]);
                *reinterpret_cast<float4*>(g + 12) = make_float4(dgz[4], dgz[5], dgz[6], dgz[7]);
                *reinterpret_cast<float4*>(g + 16) = make_float4(dgn[0], dgn[1], dgn[2], dgn[3]);
                *reinterpret_cast<float4*>(g + 20) = make_float4(dgn[4], dgn[5], dgn[6], dgn[7]);
            }
            named_bar_arrive(6, 256);
            if (etid == 0) TB_TRACE(6);
            if (act) {
                // publish dgh_t = [dar, daz, dan*r] (bf16 hi/lo, UMMA order) for the next step's contraction
                uint16_t* dst = a.gxh + (size_t)(n & 1) * 2 * gx_part;
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    const float* src = (g == 0) ? dgr : (g == 1) ? dgz : dgnr;
                    uint32_t phi[4], plo[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint16_t h0, l0, h1, l1;
                        split_bf16(src[2 * q], h0, l0);
                        split_bf16(src[2 * q + 1], h1, l1);
                        phi[q] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                        plo[q] = (uint32_t)l0 | ((uint32_t)l1 << 16);
                    }
                    const int kidx = g * H + u0;
                    const size_t off = ((size_t)(kidx >> 6) * L.MB + (b >> 3)) * 512 + (size_t)((kidx & 63) >> 3) * 64 + (size_t)(b & 7) * 8;
                    *reinterpret_cast<uint4*>(dst + off) = make_uint4(phi[0], phi[1], phi[2], phi[3]);
                    *reinterpret_cast<uint4*>(dst + gx_part + off) = make_uint4(plo[0], plo[1], plo[2], plo[3]);
                }
                // partial of dy_{t-1} feedback, outputs [0, 32)
                float* pd = f.part + (size_t)c * n_pairs + b;   // [c][o][b]: lanes write consecutive rows
                const int o_hi = out < 32 ? out : 32;
                for (int o = 0; o < o_hi; o += 4) {
                    float p4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int rr = 0; rr < 24; ++rr) {
                        const float gv = (rr < 8) ? dgr[rr & 7] : (rr < 16) ? dgz[rr & 7] : dgn[rr & 7];
                        const float4 w = *reinterpret_cast<const float4*>(sWy + rr * 64 + o);
                        p4[0] = fmaf(gv, w.x, p4[0]);
                        p4[1] = fmaf(gv, w.y, p4[1]);
                        p4[2] = fmaf(gv, w.z, p4[2]);
                        p4[3] = fmaf(gv, w.w, p4[3]);
                    }
                    for (int q = 0; q < 4 && o + q < o_hi; ++q) pd[(size_t)(o + q) * B] = p4[q];
                }
            }
            if (etid == 0) TB_TRACE(7);
            __threadfence();
            fence_proxy_async_all();
            if (etid == 0) TB_TRACE(8);
            named_bar_sync(7, 256);
            if (etid == 0) red_release_gpu_add(ctrA, 1u);
            if (etid == 0) TB_TRACE(9);
            if (act) {   // off the critical path: only the products after the kernel read these
                // saved for the deferred weight-gradient products
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
        }
    } else if (warp >= 8) {
        // ================= aux: dy reduction, q, outputs [32, 64) of the partial =========================
        const int rt = threadIdx.x - 256;
        const int Q = (n_pairs + G - 1) / G;
        const int q_lo = c * Q;
        const int q_n = max(0, min(Q, n_pairs - q_lo));
        constexpr int LB = 22;   // independent loads in flight per thread
        for (int n = 0; n <= T; ++n) {
            const int t = T - 1 - n;
            if (n > 0) {
                // dy_tot[t+1] += sum over CTAs of the partials of step t+1 (fixed order)
                if (rt == 0) spin_until(ctrA, (unsigned)G * (unsigned)n);
                if (rt == 0) TB_TRACE(20);
                named_bar_sync(2, 128);
                float* dyt = f.dy_tot + (size_t)(t + 1) * n_pairs;
                for (int qb = 0; qb < q_n; qb += 128) {
                    // stage the [G][w] block of partials: thread -> (row r0 of every RP-th CTA, pair qc); no divisions inside
                    const int w = min(128, q_n - qb);
                    const int RP = 128 / w;
                    const int r0 = rt / w, qc = rt - r0 * w;
                    if (r0 < RP) {
                        const float* src = f.part + (size_t)r0 * n_pairs + q_lo + qb + qc;
                        float* dstp = sRed + r0 * w + qc;
                        for (int cc0 = 0; cc0 < G; cc0 += RP * LB) {
                            float v[LB];
#pragma unroll
                            for (int k = 0; k < LB; ++k)
                                if (cc0 + k * RP + r0 < G) v[k] = __ldcg(src + (size_t)(cc0 + k * RP) * n_pairs);
#pragma unroll
                            for (int k = 0; k < LB; ++k)
                                if (cc0 + k * RP + r0 < G) dstp[(cc0 + k * RP) * w] = v[k];
                        }
                    }
                    named_bar_sync(2, 128);
                    if (rt < w) {
                        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                        int cc = 0;
                        for (; cc + 3 < G; cc += 4) {
                            s0 += sRed[(cc + 0) * w + rt];
                            s1 += sRed[(cc + 1) * w + rt];
                            s2 += sRed[(cc + 2) * w + rt];
                            s3 += sRed[(cc + 3) * w + rt];
                        }
                        for (; cc < G; ++cc) s0 += sRed[cc * w + rt];
                        const int qq = q_lo + qb + rt;
                        const int o = qq / B, bb = qq - o * B;   // pair order of `part` is [o][b]
                        float* dyp = dyt + (size_t)bb * out + o;
                        *dyp = __ldcg(dyp) + ((s0 + s1) + (s2 + s3));
                    }
                    if (qb + 128 < q_n) named_bar_sync(2, 128);
                }
                if (rt == 0) TB_TRACE(21);
                __threadfence();
                named_bar_sync(2, 128);
                if (rt == 0) red_release_gpu_add(ctrB, 1u);
                if (rt == 0) TB_TRACE(22);
            }
            if (t < 0) break;
            if (n > 0) {
                if (rt == 0) spin_until(ctrB, (unsigned)G * (unsigned)n);
                if (rt == 0) TB_TRACE(23);
                named_bar_sync(2, 128);
            }
            // q[b][uu] = sum_o dy_t[b][o] W_o[o][u0+uu]
            {
                float qv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (rt < B) {
                    const float* dy = f.dy_tot + (size_t)(t + 1) * n_pairs + (size_t)rt * out;
                    auto fma8 = [&](float d, int o) {
                        const float4 w0 = *reinterpret_cast<const float4*>(sWo + o * 8);
                        const float4 w1 = *reinterpret_cast<const float4*>(sWo + o * 8 + 4);
                        qv[0] = fmaf(d, w0.x, qv[0]); qv[1] = fmaf(d, w0.y, qv[1]);
                        qv[2] = fmaf(d, w0.z, qv[2]); qv[3] = fmaf(d, w0.w, qv[3]);
                        qv[4] = fmaf(d, w1.x, qv[4]); qv[5] = fmaf(d, w1.y, qv[5]);
                        qv[6] = fmaf(d, w1.z, qv[6]); qv[7] = fmaf(d, w1.w, qv[7]);
                    };
                    if ((out & 3) == 0) {   // the row is 16-byte aligned: all loads in flight before the first use
                        float4 d4[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q)
                            if (4 * q < out) d4[q] = __ldcg(reinterpret_cast<const float4*>(dy) + q);
#pragma unroll
                        for (int q = 0; q < 16; ++q)
                            if (4 * q < out) {
                                fma8(d4[q].x, 4 * q);
                                fma8(d4[q].y, 4 * q + 1);
                                fma8(d4[q].z, 4 * q + 2);
                                fma8(d4[q].w, 4 * q + 3);
                            }
                    } else {
                        for (int o0 = 0; o0 < out; o0 += 16) {
                            float d1[16];
#pragma unroll
                            for (int o = 0; o < 16; ++o)
                                if (o0 + o < out) d1[o] = __ldcg(dy + o0 + o);
#pragma unroll
                            for (int o = 0; o < 16; ++o)
                                if (o0 + o < out) fma8(d1[o], o0 + o);
                        }
                    }
                }
                *reinterpret_cast<float4*>(sQ + rt * 8) = make_float4(qv[0], qv[1], qv[2], qv[3]);
                *reinterpret_cast<float4*>(sQ + rt * 8 + 4) = make_float4(qv[4], qv[5], qv[6], qv[7]);
            }
            if (rt == 0) TB_TRACE(24);
            named_bar_arrive(5, 256);
            named_bar_sync(6, 256);   // sG = dgi of the own units
            if (rt == 0) TB_TRACE(25);
            if (rt < B && out > 32) {
                float gv[24];
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    const float4 x = *reinterpret_cast<const float4*>(sG + rt * 24 + 4 * q);
                    gv[4 * q] = x.x; gv[4 * q + 1] = x.y; gv[4 * q + 2] = x.z; gv[4 * q + 3] = x.w;
                }
                float* pd = f.part + (size_t)c * n_pairs + rt;
                for (int o = 32; o < out; o += 4) {
                    float p4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int rr = 0; rr < 24; ++rr) {
                        const float4 w = *reinterpret_cast<const float4*>(sWy + rr * 64 + o);
                        p4[0] = fmaf(gv[rr], w.x, p4[0]);
                        p4[1] = fmaf(gv[rr], w.y, p4[1]);
                        p4[2] = fmaf(gv[rr], w.z, p4[2]);
                        p4[3] = fmaf(gv[rr], w.w, p4[3]);
                    }
                    for (int q = 0; q < 4 && o + q < out; ++q) pd[(size_t)(o + q) * B] = p4[q];
                }
            }
            if (rt == 0) TB_TRACE(26);
            __threadfence();
            named_bar_arrive(7, 256);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // no CTA leaves while a peer may still store into its inbox
    if (warp == 2) tmem_dealloc<64>(tmem);
}

// ---- host side ----------------------------------------------------------------------------------------
bool gru_tc_bwd_shape_ok(int B, int H, int out) {
    return H % 64 == 0 && H >= 64 && out >= 1 && out <= 64 && B >= 1 && B <= 128 && (3 * H / TB_KC) % 4 == 0 && H % 32 == 0;
}

size_t gru_tc_bwd_scratch_floats(int B, int H) {
    size_t MB = (B + 7) / 8;
    size_t gxh = (size_t)2 * 2 * (3 * H / TB_KC) * MB * 512 / 2;   // bf16 elements -> floats
    return round_up_sz(gxh, 64) + 64;
}

// picks the cluster size (8 preferred) whose clusters are all co-resident; 0 = not runnable
static int pick_cluster(int B, int H, int out, const DeviceInfo& di, TbLayout* Lout) {
    const int G = H / 8;
    if (!gru_tc_bwd_shape_ok(B, H, out) || G > di.n_sm) return 0;
    struct Entry { int B, H, out, S; };
    static Entry cache[16];
    static int n_cache = 0;
    int S = -1;
    for (int i = 0; i < n_cache; ++i)
        if (cache[i].B == B && cache[i].H == H && cache[i].out == out) S = cache[i].S;
    if (S < 0) {
        S = 0;
        for (int cand = 8; cand >= 4 && S == 0; cand >>= 1) {
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
        if (n_cache < 16) cache[n_cache++] = Entry{B, H, out, S};
    }
    if (S && Lout) *Lout = tb_layout(B, H, S, G, out, di.max_smem_optin);
    return S;
}

bool gru_tc_bwd_supported(int B, int H, int out, const DeviceInfo& di) { return pick_cluster(B, H, out, di, nullptr) != 0; }

int gru_ar_bwd_tc(GruBwdArgs& f, float* tc_scratch, cudaStream_t s) {
    if (f.T <= 0 || f.B <= 0) return 0;
    DeviceInfo di;
    if (int rc = get_device_info(&di)) return rc;
    TbLayout L;
    const int S = pick_cluster(f.B, f.H, f.out, di, &L);
    CVB_REQUIRE(S != 0, "gru_ar_bwd_tc: unsupported shape B=%d H=%d out=%d", f.B, f.H, f.out);
    GruTcBwdArgs a;
    a.f = f;
    const size_t gxh_f = round_up_sz((size_t)2 * 2 * (3 * f.H / TB_KC) * L.MB * 512 / 2, 64);
    a.gxh = reinterpret_cast<uint16_t*>(tc_scratch);
    a.ctr = reinterpret_cast<unsigned*>(tc_scratch + gxh_f);
    a.S = S;
    a.smem_max = di.max_smem_optin;
    a.trace = nullptr;
    const char* trace_file = getenv("CVB_TRACE_FILE");
    const size_t trace_bytes = (size_t)(f.T + 1) * 64 * sizeof(long long);
    if (trace_file && trace_file[0]) {
        CVB_CHECK(cudaMalloc(&a.trace, trace_bytes));
        CVB_CHECK(cudaMemsetAsync(a.trace, 0, trace_bytes, s));
    }
    CVB_CHECK(cudaMemsetAsync(a.ctr, 0, 64 * sizeof(float), s));
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
    cfg.numAttrs = 2;
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
