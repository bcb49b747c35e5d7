// Tensor-core (tcgen05) variant of the persistent AR-GRU forward recurrence (gru_vae.py:364-399).
//
// Work split: CTA c owns hidden units [8c, 8c+8): their r,z,n rows of W_hh / W_y and their columns of
// W_o stay in shared memory for the whole sequence as bf16 hi+lo pairs (x = hi + lo).
// Per step the gate pre-activations of ALL batch rows for those units are one MMA chain
//     D[128 (batch rows), 32] += A[128, K] * B[32, K]^T,   K = H (h_{t-1} chunks) + 64 (y_{t-1} chunk)
// with D columns [r(8) | z(8) | W_hn h (8) | W_yn y (8)], three MMAs per K-step (hi*hi, lo*hi, hi*lo),
// fp32 accumulation in TMEM.  A (the batch side) is what every CTA must all-gather each step: the
// owners publish h_t as bf16 hi/lo already arranged in the UMMA K-major core-matrix order, so each K
// chunk is ONE contiguous cp.async.bulk into the ring (no tensor maps, no swizzle).
//
// Warp roles (384 threads): w0 bulk-copy producer, w1 MMA issuer, w2 TMEM allocator, w4-7 epilogue
// (TMEM lane = batch row: gates, h_t, saved activations, partial y), w8-11 reducers (fixed-order sum
// of the per-CTA partial y_t = W_o o_t, published in both fp32 and UMMA order).  Grid-wide sync is two
// monotonic counters: A (h_t + partials published), B (y_t published); the y reduction of step t runs
// concurrently with the h-chunk ingest of step t+1 and only the last K chunk waits for it.
#include "gru_ar.cuh"
#include "umma.cuh"

namespace cvb {
using namespace umma;

constexpr int TC_NT = 384;
constexpr int TC_U = 8;
constexpr int TC_N = 32;            // MMA N: r,z,ghn,gin x 8 units
constexpr int TC_KC = 64;           // K per ring stage
constexpr int TC_WH_PART = 65536;   // bytes of one part (hi or lo) of the W_hh operand at H = 1024 (scaled by H/1024)
constexpr int TC_RED_FLOATS = 5120;

struct TcLayout {
    int MB;          // batch row blocks of 8
    int NS;          // ring stages
    int nchunk;      // H / 64
    uint32_t stage_bytes, ring_bytes, wh_part_bytes, off_wh, off_wy, off_wo, off_bh, off_red, off_bar, total;
};

__host__ __device__ inline TcLayout tc_layout(int B, int H, int smem_max) {
    TcLayout L;
    L.MB = (B + 7) / 8;
    L.nchunk = H / TC_KC;
    L.stage_bytes = 2u * L.MB * 1024u;
    L.wh_part_bytes = (uint32_t)L.nchunk * 4096u;
    uint32_t fixed = 2 * L.wh_part_bytes + 8192 + 64 * TC_U * 4 + 128 + TC_RED_FLOATS * 4 + 256;
    int ns = ((int)smem_max - (int)fixed) / (int)L.stage_bytes;
    L.NS = ns > 8 ? 8 : ns;
    L.ring_bytes = (uint32_t)(L.NS > 0 ? L.NS : 0) * L.stage_bytes;
    L.off_wh = L.ring_bytes;
    L.off_wy = L.off_wh + 2 * L.wh_part_bytes;
    L.off_wo = L.off_wy + 8192;
    L.off_bh = L.off_wo + 64 * TC_U * 4;
    L.off_red = L.off_bh + 128;
    L.off_bar = L.off_red + TC_RED_FLOATS * 4;
    L.total = L.off_bar + 256;
    return L;
}

struct GruTcArgs {
    GruFwdArgs f;        // same tensors as the exact kernel
    uint16_t* hx;        // [2 slots][2 parts][nchunk][MB][8 kblk][64] bf16 (UMMA order)
    uint16_t* yx;        // [2 slots][2 parts][MB][8 kblk][64] bf16, zero-initialised
    unsigned* ctr;       // [0] = A, [32] = B (separate 128B lines), zero-initialised
    int smem_max;
};

__device__ __forceinline__ void spin_until(const unsigned* ctr, unsigned target) {
    while (ld_acquire_gpu(ctr) < target) {
    }
}

__global__ void __launch_bounds__(TC_NT, 1) k_gru_fwd_tc(GruTcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const GruFwdArgs& f = a.f;
    const int B = f.B, T = f.T, H = f.H, out = f.out;
    const int G = gridDim.x, c = blockIdx.x, u0 = c * TC_U;
    const TcLayout L = tc_layout(B, H, a.smem_max);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* ring = smem;
    uint8_t* sWh = smem + L.off_wh;
    uint8_t* sWy = smem + L.off_wy;
    float* sWo = reinterpret_cast<float*>(smem + L.off_wo);   // [64][8]
    float* sBh = reinterpret_cast<float*>(smem + L.off_bh);   // [3][8]
    float* sRed = reinterpret_cast<float*>(smem + L.off_red);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* empty = full + 8;
    uint64_t* accum_full = empty + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);
    const size_t hx_part = (size_t)L.nchunk * L.MB * 512;   // elements per part
    const size_t yx_part = (size_t)L.MB * 512;
    unsigned* ctrA = a.ctr;
    unsigned* ctrB = a.ctr + 32;

    // ---- one-time setup: weights -> bf16 hi/lo in UMMA K-major core-matrix order -----------------
    for (int i = threadIdx.x; i < TC_N * H; i += TC_NT) {
        int n = i / H, k = i - n * H;
        int g = n >> 3, uu = n & 7;
        float w = (g < 3) ? f.Whh[(size_t)(g * H + u0 + uu) * H + k] : 0.f;
        uint16_t hi, lo;
        split_f16(w, hi, lo);
        uint32_t off = (uint32_t)(k / TC_KC) * 4096u + (uint32_t)(n >> 3) * 1024u + (uint32_t)((k % TC_KC) >> 3) * 128u + (uint32_t)(n & 7) * 16u +
                       (uint32_t)(k & 7) * 2u;
        *reinterpret_cast<uint16_t*>(sWh + off) = hi;
        *reinterpret_cast<uint16_t*>(sWh + L.wh_part_bytes + off) = lo;
    }
    for (int i = threadIdx.x; i < TC_N * TC_KC; i += TC_NT) {
        int n = i / TC_KC, k = i - n * TC_KC;
        int g = n >> 3, uu = n & 7;
        int row = (g == 0) ? u0 + uu : (g == 1) ? H + u0 + uu : (g == 3) ? 2 * H + u0 + uu : -1;   // block 2 (W_hn h) gets no y term
        float w = (row >= 0 && k < out) ? f.Wy[(size_t)row * f.ldwy + k] : 0.f;
        uint16_t hi, lo;
        split_f16(w, hi, lo);
        uint32_t off = (uint32_t)(n >> 3) * 1024u + (uint32_t)(k >> 3) * 128u + (uint32_t)(n & 7) * 16u + (uint32_t)(k & 7) * 2u;
        *reinterpret_cast<uint16_t*>(sWy + off) = hi;
        *reinterpret_cast<uint16_t*>(sWy + 4096 + off) = lo;
    }
    for (int i = threadIdx.x; i < 64 * TC_U; i += TC_NT) {
        int o = i / TC_U, uu = i - o * TC_U;
        sWo[i] = (o < out) ? f.Wo[(size_t)o * H + u0 + uu] : 0.f;
    }
    if (threadIdx.x < 24) sBh[threadIdx.x] = f.bhh[(threadIdx.x >> 3) * H + u0 + (threadIdx.x & 7)];
    if (threadIdx.x == 0) {
        for (int s = 0; s < 8; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(accum_full, 1);
        mbar_fence_init();
    }
    fence_proxy_async_smem();
    if (warp == 2) tmem_alloc<32>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ================= bulk-copy producer =====================================================
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = 0; t < T; ++t) {
                const uint16_t* hsrc = a.hx + (size_t)(t & 1) * 2 * hx_part;
                const uint16_t* ysrc = a.yx + (size_t)(t & 1) * 2 * yx_part;
                spin_until(ctrA, (unsigned)G * (unsigned)(t + 1));
                fence_proxy_async_all();
                for (int ch = 0; ch <= L.nchunk; ++ch, ++it) {
                    const int s = it % L.NS;
                    mbar_wait(&empty[s], ((it / L.NS) & 1) ^ 1);
                    uint8_t* dst = ring + (size_t)s * L.stage_bytes;
                    const uint32_t half = L.MB * 1024u;
                    mbar_expect_tx(&full[s], 2 * half);
                    if (ch < L.nchunk) {
                        bulk_g2s(dst, hsrc + (size_t)ch * L.MB * 512, half, &full[s]);
                        bulk_g2s(dst + half, hsrc + hx_part + (size_t)ch * L.MB * 512, half, &full[s]);
                    } else {
                        spin_until(ctrB, (unsigned)G * (unsigned)(t + 1));
                        fence_proxy_async_all();
                        bulk_g2s(dst, ysrc, half, &full[s]);
                        bulk_g2s(dst + half, ysrc + yx_part, half, &full[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer ================================================================
        if (lane == 0) {
            const uint32_t idesc = idesc_f16_f32(128, TC_N);
            const uint32_t half = L.MB * 1024u;
            uint32_t it = 0;
            for (int t = 0; t < T; ++t) {
                for (int ch = 0; ch <= L.nchunk; ++ch, ++it) {
                    const int s = it % L.NS;
                    mbar_wait(&full[s], (it / L.NS) & 1);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(ring + (size_t)s * L.stage_bytes);
                    const uint32_t a_lo = a_hi + half;
                    const uint32_t b_hi = (ch < L.nchunk) ? smem_u32(sWh) + (uint32_t)ch * 4096u : smem_u32(sWy);
                    const uint32_t b_lo = (ch < L.nchunk) ? b_hi + L.wh_part_bytes : b_hi + 4096u;
#pragma unroll
                    for (int k16 = 0; k16 < TC_KC / 16; ++k16) {
                        const uint64_t dah = smem_desc(a_hi + k16 * 256, 128, 1024);
                        const uint64_t dal = smem_desc(a_lo + k16 * 256, 128, 1024);
                        const uint64_t dbh = smem_desc(b_hi + k16 * 256, 128, 1024);
                        const uint64_t dbl = smem_desc(b_lo + k16 * 256, 128, 1024);
                        mma_bf16_ss(tmem, dah, dbh, idesc, (ch | k16) != 0);
                        mma_bf16_ss(tmem, dal, dbh, idesc, true);
                        mma_bf16_ss(tmem, dah, dbl, idesc, true);
                    }
                    mma_commit(&empty[s]);
                    if (ch == L.nchunk) mma_commit(accum_full);
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ================= epilogue: TMEM lane = batch row =============================================
        const int b = (warp - 4) * 32 + lane;
        const bool act = b < B;
        const int etid = threadIdx.x - 128;
        float hreg[8];
        // prologue: publish h_in (slot 0) in UMMA order
        {
            uint32_t phi[4], plo[4];
#pragma unroll
            for (int j = 0; j < 8; ++j) hreg[j] = act ? f.hs[(size_t)b * H + u0 + j] : 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint16_t h0, l0, h1, l1;
                split_f16(hreg[2 * j], h0, l0);
                split_f16(hreg[2 * j + 1], h1, l1);
                phi[j] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                plo[j] = (uint32_t)l0 | ((uint32_t)l1 << 16);
            }
            if (act) {
                size_t off = ((size_t)(c >> 3) * L.MB + (b >> 3)) * 512 + (size_t)(c & 7) * 64 + (size_t)(b & 7) * 8;
                *reinterpret_cast<uint4*>(a.hx + off) = make_uint4(phi[0], phi[1], phi[2], phi[3]);
                *reinterpret_cast<uint4*>(a.hx + hx_part + off) = make_uint4(plo[0], plo[1], plo[2], plo[3]);
            }
            __threadfence();
            fence_proxy_async_all();
            named_bar_sync(1, 128);
            if (etid == 0) red_release_gpu_add(ctrA, 1u);
        }
        for (int t = 0; t < T; ++t) {
            const size_t row = (size_t)t * B + (act ? b : 0);
            float4 gxv[6];
            float4 mk[2] = {make_float4(1.f, 1.f, 1.f, 1.f), make_float4(1.f, 1.f, 1.f, 1.f)};
            if (act) {
                const float* g = f.gx + row * 3 * H + u0;
#pragma unroll
                for (int gi = 0; gi < 3; ++gi) {
                    gxv[2 * gi] = *reinterpret_cast<const float4*>(g + (size_t)gi * H);
                    gxv[2 * gi + 1] = *reinterpret_cast<const float4*>(g + (size_t)gi * H + 4);
                }
                if (f.mask) {
                    mk[0] = *reinterpret_cast<const float4*>(f.mask + row * H + u0);
                    mk[1] = *reinterpret_cast<const float4*>(f.mask + row * H + u0 + 4);
                }
            }
            mbar_wait(accum_full, t & 1);
            tc_fence_after();
            float v[32];
            const uint32_t taddr = tmem + ((uint32_t)((warp - 4) * 32) << 16);
            tmem_ld_x16(taddr, v);
            tmem_ld_x16(taddr + 16, v + 16);
            tmem_ld_wait();
            tc_fence_before();
            float ov[8];
            if (act) {
                const float* gxr = reinterpret_cast<const float*>(&gxv[0]);
                const float* gxz = reinterpret_cast<const float*>(&gxv[2]);
                const float* gxn = reinterpret_cast<const float*>(&gxv[4]);
                const float* mkf = reinterpret_cast<const float*>(&mk[0]);
                float rr[8], zz[8], nn[8], gh[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    rr[j] = sigmoidf_(gxr[j] + v[j] + sBh[j]);
                    zz[j] = sigmoidf_(gxz[j] + v[8 + j] + sBh[8 + j]);
                    gh[j] = v[16 + j] + sBh[16 + j];
                    nn[j] = tanhf(gxn[j] + v[24 + j] + rr[j] * gh[j]);
                    hreg[j] = (1.0f - zz[j]) * nn[j] + zz[j] * hreg[j];
                    ov[j] = hreg[j] * mkf[j];
                }
                float* hd = f.hs + (size_t)(t + 1) * B * H + (size_t)b * H + u0;
                *reinterpret_cast<float4*>(hd) = make_float4(hreg[0], hreg[1], hreg[2], hreg[3]);
                *reinterpret_cast<float4*>(hd + 4) = make_float4(hreg[4], hreg[5], hreg[6], hreg[7]);
                if (f.sv_r) {
                    const size_t so = row * H + u0;
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
                    const size_t so = row * H + u0;
                    *reinterpret_cast<float4*>(f.sv_o + so) = make_float4(ov[0], ov[1], ov[2], ov[3]);
                    *reinterpret_cast<float4*>(f.sv_o + so + 4) = make_float4(ov[4], ov[5], ov[6], ov[7]);
                }
                // publish h_t (bf16 hi/lo, UMMA order) into the other exchange slot
                uint32_t phi[4], plo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint16_t h0, l0, h1, l1;
                    split_f16(hreg[2 * j], h0, l0);
                    split_f16(hreg[2 * j + 1], h1, l1);
                    phi[j] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                    plo[j] = (uint32_t)l0 | ((uint32_t)l1 << 16);
                }
                uint16_t* hdst = a.hx + (size_t)((t + 1) & 1) * 2 * hx_part;
                size_t off = ((size_t)(c >> 3) * L.MB + (b >> 3)) * 512 + (size_t)(c & 7) * 64 + (size_t)(b & 7) * 8;
                *reinterpret_cast<uint4*>(hdst + off) = make_uint4(phi[0], phi[1], phi[2], phi[3]);
                *reinterpret_cast<uint4*>(hdst + hx_part + off) = make_uint4(plo[0], plo[1], plo[2], plo[3]);
                // partial y_t = W_o[:, own units] o_t
                float* pd = f.part + ((size_t)c * B + b) * out;
                for (int o = 0; o < out; o += 4) {
                    float p4[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 w0 = *reinterpret_cast<const float4*>(sWo + (o + q) * 8);
                        const float4 w1 = *reinterpret_cast<const float4*>(sWo + (o + q) * 8 + 4);
                        float s = w0.x * ov[0];
                        s = fmaf(w0.y, ov[1], s);
                        s = fmaf(w0.z, ov[2], s);
                        s = fmaf(w0.w, ov[3], s);
                        s = fmaf(w1.x, ov[4], s);
                        s = fmaf(w1.y, ov[5], s);
                        s = fmaf(w1.z, ov[6], s);
                        s = fmaf(w1.w, ov[7], s);
                        p4[q] = s;
                    }
                    if (o + 3 < out) {
                        if ((out & 3) == 0) {
                            *reinterpret_cast<float4*>(pd + o) = make_float4(p4[0], p4[1], p4[2], p4[3]);
                        } else {
                            pd[o] = p4[0]; pd[o + 1] = p4[1]; pd[o + 2] = p4[2]; pd[o + 3] = p4[3];
                        }
                    } else {
                        for (int q = 0; q < 4 && o + q < out; ++q) pd[o + q] = p4[q];
                    }
                }
            }
            __threadfence();
            fence_proxy_async_all();
            named_bar_sync(1, 128);
            if (etid == 0) red_release_gpu_add(ctrA, 1u);
        }
    } else if (warp >= 8) {
        // ================= reducers: y_t = b_o + sum_c partial_c, published fp32 + UMMA order ==========
        const int rtid = threadIdx.x - 256;
        const int n_pairs = B * out;
        const int Q = (n_pairs + G - 1) / G;
        const int q_lo = c * Q;
        const int q_n = max(0, min(Q, n_pairs - q_lo));
        const int QB = max(1, min(128, TC_RED_FLOATS / G));   // pairs per pass through the scratch (one per reducer thread)
        for (int round = 0; round <= T; ++round) {
            // round 0 publishes y_in; round r >= 1 reduces the partials of step r-1 into y_{r-1}
            if (round > 0) {
                if (rtid == 0) spin_until(ctrA, (unsigned)G * (unsigned)(round + 1));
                named_bar_sync(2, 128);
            }
            float* ydst = f.ys + (size_t)round * n_pairs;
            uint16_t* yx = a.yx + (size_t)(round & 1) * 2 * yx_part;
            for (int qb = 0; qb < q_n; qb += QB) {
                const int nq = min(QB, q_n - qb);
                float yv = 0.f;
                if (round > 0) {
                    for (int i = rtid; i < nq * G; i += 128) {
                        int cc = i / nq, q = i - cc * nq;
                        sRed[cc * nq + q] = __ldcg(f.part + (size_t)cc * n_pairs + q_lo + qb + q);
                    }
                    named_bar_sync(2, 128);
                    if (rtid < nq) {
                        float s = 0.f;
                        for (int cc = 0; cc < G; ++cc) s += sRed[cc * nq + rtid];
                        yv = s + f.bo[(q_lo + qb + rtid) % out];
                    }
                } else if (rtid < nq) {
                    yv = f.ys[q_lo + qb + rtid];
                }
                if (rtid < nq) {
                    const int q = q_lo + qb + rtid;
                    const int bb = q / out, o = q - bb * out;
                    if (round > 0) ydst[q] = yv;
                    uint16_t hi, lo;
                    split_f16(yv, hi, lo);
                    size_t off = (size_t)(bb >> 3) * 512 + (size_t)(o >> 3) * 64 + (size_t)(bb & 7) * 8 + (size_t)(o & 7);
                    yx[off] = hi;
                    yx[yx_part + off] = lo;
                }
                named_bar_sync(2, 128);
            }
            __threadfence();
            fence_proxy_async_all();
            named_bar_sync(2, 128);
            if (rtid == 0) red_release_gpu_add(ctrB, 1u);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<32>(tmem);
}

// eligibility of the tensor-core variant
bool gru_tc_shape_ok(int B, int H, int out) { return H % TC_KC == 0 && H >= TC_KC && out <= 64 && B <= 128 && B >= 1; }
bool gru_tc_supported(int B, int H, int out, const DeviceInfo& di) {
    if (!gru_tc_shape_ok(B, H, out) || H / TC_U > di.n_sm) return false;
    TcLayout L = tc_layout(B, H, di.max_smem_optin);
    return L.NS >= 2;
}

size_t gru_tc_scratch_floats(int B, int H) {
    size_t MB = (B + 7) / 8;
    size_t hx = (size_t)2 * 2 * (H / TC_KC) * MB * 512 / 2;   // bf16 elements -> floats
    size_t yx = (size_t)2 * 2 * MB * 512 / 2;
    return round_up_sz(hx, 64) + round_up_sz(yx, 64) + 64;
}

int gru_ar_fwd_tc(GruFwdArgs& f, float* tc_scratch, cudaStream_t s) {
    if (f.T <= 0 || f.B <= 0) return 0;
    DeviceInfo di;
    if (int rc = get_device_info(&di)) return rc;
    CVB_REQUIRE(gru_tc_supported(f.B, f.H, f.out, di), "gru_ar_fwd_tc: unsupported shape B=%d H=%d out=%d", f.B, f.H, f.out);
    TcLayout L = tc_layout(f.B, f.H, di.max_smem_optin);
    GruTcArgs a;
    a.f = f;
    size_t MB = L.MB;
    size_t hx_f = round_up_sz((size_t)2 * 2 * L.nchunk * MB * 512 / 2, 64);
    size_t yx_f = round_up_sz((size_t)2 * 2 * MB * 512 / 2, 64);
    a.hx = reinterpret_cast<uint16_t*>(tc_scratch);
    a.yx = reinterpret_cast<uint16_t*>(tc_scratch + hx_f);
    a.ctr = reinterpret_cast<unsigned*>(tc_scratch + hx_f + yx_f);
    a.smem_max = di.max_smem_optin;
    CVB_CHECK(cudaMemsetAsync(a.yx, 0, (yx_f + 64) * sizeof(float), s));   // y padding columns + both counters
    CVB_CHECK(cudaFuncSetAttribute(k_gru_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    void* params[] = {&a};
    prof_begin(s, CVB_PROF_GRU_FWD);
    CVB_CHECK(cudaLaunchCooperativeKernel((const void*)k_gru_fwd_tc, dim3(f.H / TC_U), dim3(TC_NT), params, L.total, s));
    prof_end(s, CVB_PROF_GRU_FWD);
    count_launch();
    return 0;
}

}  // namespace cvb
