// The autoregressive GRU recurrence of GRU_RNN.forward (gru_vae.py:364-399) and its BPTT
//  as PERSISTENT kernels: one cooperative launch runs all T steps.
//
// "exact" variant (this file): fp32 FMA, any H/out/B.  Every CTA owns U=8 hidden units (their
// r,z,n rows of W_hh and W_y and their columns of W_o) and keeps those weights in shared memory
// for the whole sequence, so the weights are read from HBM once per pass.  Per step the only
// global traffic is the h_{t-1} all-gather (B*H floats per CTA, cp.async double-buffered), the
// gx[t] slice and the saved gate values.  The y_t = W_o o_t feedback is a reduction over all of H:
// each CTA writes its partial, the grid synchronises, a fixed-order tree sums the partials
// (deterministic -- no float atomics), the grid synchronises again.
//
// State layout is time-major: hs [T+1,B,H] (slot 0 = h_in), ys [T+1,B,out] (slot 0 = y_in).
#include "gru_ar.cuh"

namespace cvb {

constexpr int U = 8;        // hidden units per CTA
constexpr int NT = 256;     // threads per CTA (one warp per unit in forward)
constexpr int BT = 128;     // batch tile (4 rows per lane)
// KC = K chunk staged per cp.async group (template parameter: 64, or 16 when the wider staging tile does not fit next to
// a wide output layer, e.g. the encoder at lat_dim 50 / 64); KP = KC + 4 = padded row pitch (conflict-free LDS.128)

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    return fmaf(a.w, b.w, acc);
}

// stage rows [b0, b0+BT) x cols [k0, k0+KC) of src [Brows, K] into dst [BT][KP]; zero-fill outside
template <int KC>
__device__ __forceinline__ void stage_chunk(float* dst, const float* __restrict__ src, int b0, int Brows, int K, int k0,
                                            bool vec4, int nrows) {
    constexpr int KP = KC + 4;
    if (vec4) {
        for (int i = threadIdx.x; i < nrows * (KC / 4); i += NT) {
            int row = i / (KC / 4), q = i - row * (KC / 4);
            int b = b0 + row, k = k0 + 4 * q;
            bool ok = (b < Brows) && (k < K);
            const float* g = ok ? src + (size_t)b * K + k : src;
            cp_async16(dst + row * KP + 4 * q, g, ok);
        }
    } else {
        for (int i = threadIdx.x; i < nrows * KC; i += NT) {
            int row = i / KC, q = i - row * KC;
            int b = b0 + row, k = k0 + q;
            bool ok = (b < Brows) && (k < K);
            const float* g = ok ? src + (size_t)b * K + k : src;
            cp_async4(dst + row * KP + q, g, ok);
        }
    }
}

// deterministic sum of the per-CTA partials for this CTA's share of the (b,o) pairs
template <typename F>
__device__ __forceinline__ void reduce_partials(const float* __restrict__ part, int G, int n_pairs, float* red, F&& emit) {
    int c = blockIdx.x;
    int Q = (n_pairs + G - 1) / G;
    int ql = threadIdx.x & 63, cg = threadIdx.x >> 6;
    for (int qb = 0; qb < Q; qb += 64) {
        int q = c * Q + qb + ql;
        bool ok = (qb + ql < Q) && (q < n_pairs);
        float s = 0.f;
        if (ok) {
#pragma unroll 8
            for (int cc = cg; cc < G; cc += 4) s += __ldcg(part + (size_t)cc * n_pairs + q);
        }
        red[cg * 64 + ql] = s;
        __syncthreads();
        if (ok && cg == 0) emit(q, (red[ql] + red[64 + ql]) + (red[128 + ql] + red[192 + ql]));
        __syncthreads();
    }
}

template <int KC>
__global__ void __launch_bounds__(NT, 1) k_gru_fwd(GruFwdArgs a) {
    constexpr int KP = KC + 4;
    extern __shared__ __align__(16) float smem[];
    const int B = a.B, T = a.T, H = a.H, out = a.out;
    const int Hp = (H + KC - 1) / KC * KC;
    const int outp = out | 1;
    const int G = gridDim.x, c = blockIdx.x, u0 = c * U;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sW = smem;                  // [3U][Hp]
    float* sWy = sW + 3 * U * Hp;      // [3U][out]
    float* sWo = sWy + 3 * U * out;    // [U][out]
    float* sY = sWo + out * U;         // [BT][outp]
    float* sO = sY + BT * outp;        // [U][BT]  (also scratch of the partial reduction)
    float* sH = sO + ((U * BT + 3) & ~3);  // [2][BT][KP]
    // keep sH 16B aligned
    sH = (float*)(((uintptr_t)sH + 15) & ~(uintptr_t)15);
    const bool vec4 = (H % 4 == 0) && ((((uintptr_t)a.hs) & 15) == 0);

    for (int i = threadIdx.x; i < 3 * U * Hp; i += NT) {
        int row = i / Hp, k = i - row * Hp;
        int g = row / U, u = u0 + (row - g * U);
        sW[i] = (u < H && k < H) ? a.Whh[(size_t)(g * H + u) * H + k] : 0.f;
    }
    for (int i = threadIdx.x; i < 3 * U * out; i += NT) {
        int row = i / out, o = i - row * out;
        int g = row / U, u = u0 + (row - g * U);
        sWy[i] = (u < H) ? a.Wy[(size_t)(g * H + u) * a.ldwy + o] : 0.f;
    }
    for (int i = threadIdx.x; i < out * U; i += NT) {
        int uu = i / out, o = i - uu * out;
        sWo[i] = (u0 + uu < H) ? a.Wo[(size_t)o * H + u0 + uu] : 0.f;
    }
    const int u = u0 + w;
    const bool u_ok = u < H;
    const float bhr = u_ok ? a.bhh[u] : 0.f, bhz = u_ok ? a.bhh[H + u] : 0.f, bhn = u_ok ? a.bhh[2 * H + u] : 0.f;
    unsigned bar_target = 0;
    const int nkc = Hp / KC;
    __syncthreads();

    for (int t = 0; t < T; ++t) {
        const float* hprev = a.hs + (size_t)t * B * H;
        const float* yprev = a.ys + (size_t)t * B * out;
        for (int b0 = 0; b0 < B; b0 += BT) {
            const int nb = min(BT, B - b0);
            const int nrows = (nb + 31) & ~31;  // rows the 4-per-lane mapping touches
            __syncthreads();  // previous tile's sO / sY / sH readers are done
            stage_chunk<KC>(sH, hprev, b0, B, H, 0, vec4, nrows);
            cp_async_commit();
            for (int i = threadIdx.x; i < nb * out; i += NT) {
                int bl = i / out, o = i - bl * out;
                sY[bl * outp + o] = __ldcg(yprev + (size_t)(b0 + bl) * out + o);
            }
            float ar[4] = {0.f, 0.f, 0.f, 0.f}, az[4] = {0.f, 0.f, 0.f, 0.f}, an[4] = {0.f, 0.f, 0.f, 0.f};
            for (int kc = 0; kc < nkc; ++kc) {
                if (kc + 1 < nkc) stage_chunk<KC>(sH + ((kc + 1) & 1) * BT * KP, hprev, b0, B, H, (kc + 1) * KC, vec4, nrows);
                cp_async_commit();
                cp_async_wait<1>();
                __syncthreads();
                const float* hb = sH + (kc & 1) * BT * KP;
                const float4* wr = reinterpret_cast<const float4*>(sW + (0 * U + w) * Hp + kc * KC);
                const float4* wz = reinterpret_cast<const float4*>(sW + (1 * U + w) * Hp + kc * KC);
                const float4* wn = reinterpret_cast<const float4*>(sW + (2 * U + w) * Hp + kc * KC);
                // chunk-local accumulators: blocked summation (rounding error ~ sqrt(KC) + sqrt(K/KC) ulp)
                float cr[4] = {0.f, 0.f, 0.f, 0.f}, cz[4] = {0.f, 0.f, 0.f, 0.f}, cn[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
                for (int k4 = 0; k4 < KC / 4; ++k4) {
                    float4 fr = wr[k4], fz = wz[k4], fn = wn[k4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (32 * i < nb) {
                            float4 h4 = *reinterpret_cast<const float4*>(hb + (lane + 32 * i) * KP + 4 * k4);
                            cr[i] = dot4(fr, h4, cr[i]);
                            cz[i] = dot4(fz, h4, cz[i]);
                            cn[i] = dot4(fn, h4, cn[i]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    ar[i] += cr[i];
                    az[i] += cz[i];
                    an[i] += cn[i];
                }
                __syncthreads();
            }
            // feedback W_y y_{t-1}
            float gin[4] = {0.f, 0.f, 0.f, 0.f};
            for (int o = 0; o < out; ++o) {
                float yr = sWy[(0 * U + w) * out + o], yz = sWy[(1 * U + w) * out + o], yn = sWy[(2 * U + w) * out + o];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (32 * i < nb) {
                        float y = sY[(lane + 32 * i) * outp + o];
                        ar[i] = fmaf(yr, y, ar[i]);
                        az[i] = fmaf(yz, y, az[i]);
                        gin[i] = fmaf(yn, y, gin[i]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int bl = lane + 32 * i;
                float o_val = 0.f;
                if (bl < nb && u_ok) {
                    int b = b0 + bl;
                    size_t row = (size_t)t * B + b;
                    const float* g = a.gx + row * 3 * H;
                    float r = sigmoidf_(g[u] + ar[i] + bhr);
                    float z = sigmoidf_(g[H + u] + az[i] + bhz);
                    float ghn = an[i] + bhn;
                    float n = tanhf(g[2 * H + u] + gin[i] + r * ghn);
                    float hp = __ldcg(hprev + (size_t)b * H + u);
                    float h = (1.0f - z) * n + z * hp;
                    a.hs[(size_t)(t + 1) * B * H + (size_t)b * H + u] = h;
                    if (a.sv_r) {
                        a.sv_r[row * H + u] = r;
                        a.sv_z[row * H + u] = z;
                        a.sv_n[row * H + u] = n;
                        a.sv_ghn[row * H + u] = ghn;
                    }
                    o_val = a.mask ? h * a.mask[row * H + u] : h;
                    if (a.sv_o) a.sv_o[row * H + u] = o_val;
                }
                if (bl < BT) sO[w * BT + bl] = o_val;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < nb * out; i += NT) {
                int bl = i / out, o = i - bl * out;
                float s = 0.f;
#pragma unroll
                for (int uu = 0; uu < U; ++uu) s = fmaf(sWo[uu * out + o], sO[uu * BT + bl], s);
                a.part[((size_t)c * B + b0 + bl) * out + o] = s;
            }
        }
        grid_barrier(a.bar, bar_target, G);
        float* ynext = a.ys + (size_t)(t + 1) * B * out;
        reduce_partials(a.part, G, B * out, sO, [&](int q, float s) { ynext[q] = s + a.bo[q % out]; });
        grid_barrier(a.bar, bar_target, G);
    }
}

// BPTT.  Warp w: unit pair (w&3), K half (w>>2); lane: 4 batch rows.
template <int KC>
__global__ void __launch_bounds__(NT, 1) k_gru_bwd(GruBwdArgs a) {
    constexpr int KP = KC + 4;
    extern __shared__ __align__(16) float smem[];
    const int B = a.B, T = a.T, H = a.H, out = a.out;
    const int K3 = 3 * H;
    const int K3p = (K3 + KC - 1) / KC * KC;
    const int outp = out | 1;
    const int G = gridDim.x, c = blockIdx.x, u0 = c * U;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = w & 3, khalf = w >> 2;
    float* sWT = smem;                   // [U][K3p]   W_hh[k][u0+u]
    float* sWy = sWT + U * K3p;          // [3U][out]
    float* sWo = sWy + 3 * U * out;      // [out][U]
    float* sDy = sWo + out * U;          // [BT][outp]
    float* sAcc = sDy + BT * outp;       // [U][BT]   (also scratch of the partial reduction)
    float* sGate = sAcc + U * BT;        // [3U][BT]
    float* sG = sGate + 3 * U * BT;      // [2][BT][KP]
    sG = (float*)(((uintptr_t)sG + 15) & ~(uintptr_t)15);
    const bool vec4 = (K3 % 4 == 0) && ((((uintptr_t)a.gxch) & 15) == 0) && (((size_t)B * K3) % 4 == 0);

    for (int i = threadIdx.x; i < U * K3p; i += NT) {
        int uu = i / K3p, k = i - uu * K3p;
        sWT[i] = (u0 + uu < H && k < K3) ? a.Whh[(size_t)k * H + u0 + uu] : 0.f;
    }
    for (int i = threadIdx.x; i < 3 * U * out; i += NT) {
        int row = i / out, o = i - row * out;
        int g = row / U, u = u0 + (row - g * U);
        sWy[i] = (u < H) ? a.Wy[(size_t)(g * H + u) * a.ldwy + o] : 0.f;
    }
    for (int i = threadIdx.x; i < out * U; i += NT) {
        int o = i / U, u = u0 + (i - o * U);
        sWo[i] = (u < H) ? a.Wo[(size_t)o * H + u] : 0.f;
    }
    unsigned bar_target = 0;
    const int nkc = K3p / KC;
    __syncthreads();

    // t == -1 is the epilogue pass that only finishes dh_in = dh_carry + dgh_0 W_hh
    for (int t = T - 1; t >= -1; --t) {
        const bool have_next = (t < T - 1);  // a dgh_{t+1} exists
        const float* gnext = a.gxch + (size_t)((t + 1) & 1) * B * K3;
        for (int b0 = 0; b0 < B; b0 += BT) {
            const int nb = min(BT, B - b0);
            const int nrows = (nb + 31) & ~31;
            __syncthreads();
            float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
            if (have_next) {
                stage_chunk<KC>(sG, gnext, b0, B, K3, 0, vec4, nrows);
                cp_async_commit();
            }
            if (t >= 0) {
                const float* dy = a.dy_tot + (size_t)(t + 1) * B * out;
                for (int i = threadIdx.x; i < nb * out; i += NT) {
                    int bl = i / out, o = i - bl * out;
                    sDy[bl * outp + o] = __ldcg(dy + (size_t)(b0 + bl) * out + o);
                }
            }
            if (have_next) {
                for (int kc = 0; kc < nkc; ++kc) {
                    if (kc + 1 < nkc) stage_chunk<KC>(sG + ((kc + 1) & 1) * BT * KP, gnext, b0, B, K3, (kc + 1) * KC, vec4, nrows);
                    cp_async_commit();
                    cp_async_wait<1>();
                    __syncthreads();
                    const float* gb = sG + (kc & 1) * BT * KP;
                    const float4* w0 = reinterpret_cast<const float4*>(sWT + (2 * pair) * K3p + kc * KC);
                    const float4* w1 = reinterpret_cast<const float4*>(sWT + (2 * pair + 1) * K3p + kc * KC);
                    float ca[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 4
                    for (int k4 = khalf * (KC / 8); k4 < (khalf + 1) * (KC / 8); ++k4) {
                        float4 f0 = w0[k4], f1 = w1[k4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (32 * i < nb) {
                                float4 g4 = *reinterpret_cast<const float4*>(gb + (lane + 32 * i) * KP + 4 * k4);
                                ca[0][i] = dot4(f0, g4, ca[0][i]);
                                ca[1][i] = dot4(f1, g4, ca[1][i]);
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[0][i] += ca[0][i];
                        acc[1][i] += ca[1][i];
                    }
                    __syncthreads();
                }
                if (khalf == 1) {
#pragma unroll
                    for (int uu = 0; uu < 2; ++uu)
#pragma unroll
                        for (int i = 0; i < 4; ++i) sAcc[(2 * pair + uu) * BT + lane + 32 * i] = acc[uu][i];
                }
            }
            __syncthreads();
            if (khalf == 0) {
#pragma unroll
                for (int uu = 0; uu < 2; ++uu) {
                    const int ul = 2 * pair + uu, u = u0 + ul;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int bl = lane + 32 * i;
                        float gr_ = 0.f, gz_ = 0.f, gn_ = 0.f;
                        if (bl < nb && u < H) {
                            const int b = b0 + bl;
                            float dh = a.dhc[(size_t)b * H + u] + acc[uu][i];
                            if (have_next) dh += sAcc[ul * BT + bl];
                            if (t < 0) {
                                a.dhc[(size_t)b * H + u] = dh;
                            } else {
                                const size_t row = (size_t)t * B + b;
                                float dyw = 0.f;
                                for (int o = 0; o < out; ++o) dyw = fmaf(sDy[bl * outp + o], sWo[o * U + ul], dyw);
                                dh += a.mask ? dyw * a.mask[row * H + u] : dyw;
                                const float r = a.sv_r[row * H + u], z = a.sv_z[row * H + u], n = a.sv_n[row * H + u];
                                const float ghn = a.sv_ghn[row * H + u];
                                const float hp = a.hs[row * H + u];  // hs slot t = h_{t-1}
                                const float dn = dh * (1.0f - z);
                                const float dz = dh * (hp - n);
                                a.dhc[(size_t)b * H + u] = dh * z;
                                const float dan = dn * (1.0f - n * n);
                                const float dar = dan * ghn * r * (1.0f - r);
                                const float daz = dz * z * (1.0f - z);
                                float* gi = a.dgi + row * K3;
                                gi[u] = dar;
                                gi[H + u] = daz;
                                gi[2 * H + u] = dan;
                                a.dghn[row * H + u] = dan * r;
                                float* gx = a.gxch + (size_t)(t & 1) * B * K3 + (size_t)b * K3;
                                gx[u] = dar;
                                gx[H + u] = daz;
                                gx[2 * H + u] = dan * r;
                                gr_ = dar;
                                gz_ = daz;
                                gn_ = dan;
                            }
                        }
                        if (t >= 0) {
                            sGate[(0 * U + ul) * BT + bl] = gr_;
                            sGate[(1 * U + ul) * BT + bl] = gz_;
                            sGate[(2 * U + ul) * BT + bl] = gn_;
                        }
                    }
                }
            }
            if (t >= 0) {
                __syncthreads();
                for (int i = threadIdx.x; i < nb * out; i += NT) {
                    int bl = i / out, o = i - bl * out;
                    float s = 0.f;
#pragma unroll
                    for (int row = 0; row < 3 * U; ++row) s = fmaf(sGate[row * BT + bl], sWy[row * out + o], s);
                    a.part[((size_t)c * B + b0 + bl) * out + o] = s;
                }
            }
        }
        if (t < 0) break;
        grid_barrier(a.bar, bar_target, G);
        float* dyp = a.dy_tot + (size_t)t * B * out;  // grad of ys slot t (= y_{t-1}; slot 0 = y_in)
        reduce_partials(a.part, G, B * out, sAcc, [&](int q, float s) { dyp[q] = __ldcg(dyp + q) + s; });
        grid_barrier(a.bar, bar_target, G);
    }
}

static size_t fwd_smem_bytes(int H, int out, int KC) {
    const size_t KP = KC + 4;
    size_t Hp = (size_t)ceil_div(H, KC) * KC;
    size_t outp = out | 1;
    size_t fl = 3 * U * Hp + 3 * U * out + (size_t)out * U + BT * outp + ((U * BT + 3) & ~3) + 4 + 2 * BT * KP;
    return fl * sizeof(float);
}
static size_t bwd_smem_bytes(int H, int out, int KC) {
    const size_t KP = KC + 4;
    size_t K3p = (size_t)ceil_div(3 * H, KC) * KC;
    size_t outp = out | 1;
    size_t fl = U * K3p + 3 * U * out + (size_t)out * U + BT * outp + U * BT + 3 * U * BT + 4 + 2 * BT * KP;
    return fl * sizeof(float);
}

int gru_exact_grid(int H) { return ceil_div(H, U); }

template <typename Args>
static int launch_coop(void (*kern)(Args), Args& a, int grid, size_t smem, cudaStream_t s, const char* name, int prof_kind) {
    DeviceInfo di;
    if (int rc = get_device_info(&di)) return rc;
    CVB_REQUIRE(smem <= (size_t)di.max_smem_optin,
                "%s: hidden_units=%d out_dim=%d needs %zu B of shared memory per CTA (device max %d)", name, a.H, a.out, smem,
                di.max_smem_optin);
    CVB_REQUIRE(grid <= di.n_sm, "%s: hidden_units=%d needs %d co-resident CTAs (device has %d SMs)", name, a.H, grid, di.n_sm);
    CVB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CVB_CHECK(cudaMemsetAsync(a.bar, 0, 64, s));
    void* params[] = {&a};
    prof_begin(s, prof_kind);
    CVB_CHECK(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(NT), params, smem, s));
    prof_end(s, prof_kind);
    count_launch();
    return 0;
}

// the wide staging tile when it fits next to the resident weights and the output-layer buffers, else the narrow one
static bool fits(size_t bytes) {
    DeviceInfo di;
    return get_device_info(&di) == 0 && bytes <= (size_t)di.max_smem_optin;
}
int gru_ar_fwd_exact(GruFwdArgs& a, cudaStream_t s) {
    if (a.T <= 0 || a.B <= 0) return 0;
    if (fits(fwd_smem_bytes(a.H, a.out, 64)))
        return launch_coop(k_gru_fwd<64>, a, gru_exact_grid(a.H), fwd_smem_bytes(a.H, a.out, 64), s, "gru_ar_fwd", CVB_PROF_GRU_FWD);
    return launch_coop(k_gru_fwd<16>, a, gru_exact_grid(a.H), fwd_smem_bytes(a.H, a.out, 16), s, "gru_ar_fwd", CVB_PROF_GRU_FWD);
}
int gru_ar_bwd_exact(GruBwdArgs& a, cudaStream_t s) {
    if (a.T <= 0 || a.B <= 0) return 0;
    if (fits(bwd_smem_bytes(a.H, a.out, 64)))
        return launch_coop(k_gru_bwd<64>, a, gru_exact_grid(a.H), bwd_smem_bytes(a.H, a.out, 64), s, "gru_ar_bwd", CVB_PROF_GRU_BWD);
    return launch_coop(k_gru_bwd<16>, a, gru_exact_grid(a.H), bwd_smem_bytes(a.H, a.out, 16), s, "gru_ar_bwd", CVB_PROF_GRU_BWD);
}

}  // namespace cvb
