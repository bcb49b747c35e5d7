// Front-end of GRU_RNN.forward: scale_in (gru_vae.py:336) -> TwoSidedDilConv1d (gru_vae.py:36-66)
// -> conv_drop (gru_vae.py:355) and its backward.
//
// Layout: every layer lives on one common "padded grid": Tp = T + 2*pad rows per utterance,
// R = B*Tp rows in all, channels-last.  In padded coordinates every layer is a CENTRED dilated
// conv (layer i output at tau reads layer i-1 at tau + (j-(k-1)/2)*k^i), which is exactly the
// reference's stack of an explicitly padded layer 0 followed by un-padded dilated layers
// (gru_vae.py:47-51) once each layer's output is stored at its receptive-field centre.  On the
// flattened grid a tap is a row-shifted view, so each tap is ONE dense product over all R rows
// (no im2col, no per-utterance batching); rows whose taps cross an utterance boundary hold finite
// values that no valid output ever reads.
//
// COMPOSED path (every shape the tensor-core GEMM takes): the conv stack has no non-linearity, so L layers are ONE
// k^L-tap conv  xc[t] = b_eff + sum_m E[:, m, :] x^[t + m - pad]  (verified 5.5e-7).  Forward is
//   k_pad_scale_all (x -> normalised, zero-padded grid)  ->  ONE product over the virtual im2col of that grid whose
//   epilogue adds b_eff, applies the dropout mask and stores time-major xc
// -- no per-layer grids, no compaction pass.  E and b_eff are derived parameters: composed by tiny products when the
// weights change (once per optimiser step, shared by every pass) and cached.  Backward: dE = dxc'^T im2col(x^), the
// input gradient from G = dxc' E gathered back over the taps, and the chain rule from (dE, db_eff) to the per-layer
// (dW_l, db_l) in parameter space (a few C_l x C_l products).  The layer-by-layer path below remains for shapes that
// stay on cuBLAS (tiny nets, CVB_GEMM=cublas).
#include <mutex>

#include "common.cuh"

namespace cvb {

struct FeGeom {
    int k, L, pad, Tp;
    size_t R;
    int C[6];          // channels of buf_0..buf_L
    size_t buf_off[6]; // float offsets of buf_i in fe_ws
    size_t wr_off[5];  // repacked weights of layer i: [k][C_{i+1}][C_i]
    size_t xc_off;     // xc_tm [T*B, C_L]
    size_t total;
};

static FeGeom fe_geom(const cvb_net* n, int B, int T) {
    FeGeom g;
    g.k = n->kernel_size;
    g.L = n->n_conv;
    g.pad = conv_pad(n);
    g.Tp = T + 2 * g.pad;
    g.R = (size_t)B * g.Tp;
    size_t off = 0;
    for (int i = 0; i <= g.L; ++i) {
        g.C[i] = n->in_dim * ipow(g.k, i);
        g.buf_off[i] = off;
        off += round_up_sz(g.R * g.C[i], 4);
    }
    for (int i = 0; i < g.L; ++i) {
        g.wr_off[i] = off;
        off += round_up_sz((size_t)g.k * g.C[i + 1] * g.C[i], 4);
    }
    g.xc_off = off;
    off += round_up_sz((size_t)B * T * g.C[g.L], 4);
    g.total = off;
    return g;
}

// composed path (below): fe_ws = [normalised padded grid + slack | xc]
static bool fe_composed(const cvb_net* n, int B, int T);
static size_t fec_xp_floats(const cvb_net* n, int B, int T);
static size_t fec_bwd_scratch_floats(const cvb_net* n, int B, int T);

size_t frontend_xc_offset(const cvb_net* n, int B, int T) {
    return fe_composed(n, B, T) ? fec_xp_floats(n, B, T) : fe_geom(n, B, T).xc_off;
}
size_t frontend_ws_floats(const cvb_net* n, int B, int T) {
    if (!fe_composed(n, B, T)) return fe_geom(n, B, T).total;
    return fec_xp_floats(n, B, T) + round_up_sz((size_t)B * T * conv_dim(n), 4);
}
static size_t layers_bwd_scratch_floats(const cvb_net* n, int B, int T);
size_t frontend_bwd_scratch_floats(const cvb_net* n, int B, int T) {
    return fe_composed(n, B, T) ? fec_bwd_scratch_floats(n, B, T) : layers_bwd_scratch_floats(n, B, T);
}
// floats of the backward's gradient grid (d buf_0..d buf_L) + one repacked weight-gradient
static size_t layers_bwd_scratch_floats(const cvb_net* n, int B, int T) {
    FeGeom g = fe_geom(n, B, T);
    size_t mx = 0;
    size_t gmx = 0;   // G = dout Wcat of a tap-fused layer: [rows][k*ci]
    for (int i = 0; i < g.L; ++i) {
        mx = mx > (size_t)g.k * g.C[i + 1] * g.C[i] ? mx : (size_t)g.k * g.C[i + 1] * g.C[i];
        gmx = gmx > g.R * g.k * g.C[i] ? gmx : g.R * g.k * g.C[i];
    }
    return g.wr_off[0] + round_up_sz(mx, 4) + round_up_sz(gmx, 4);
}

// xp[b, pad+t, i] = sum_j Ws[i, j] x[b, t, j] + bs[i]   (or a copy when there is no scale_in)
__global__ void k_pad_scale(int B, int T, int in, int pad, const float* __restrict__ x, const float* __restrict__ Ws,
                            const float* __restrict__ bs, float* __restrict__ xp) {
    extern __shared__ float sW[];  // [in*in + in]
    if (Ws) {
        for (int i = threadIdx.x; i < in * in; i += blockDim.x) sW[i] = Ws[i];
        for (int i = threadIdx.x; i < in; i += blockDim.x) sW[in * in + i] = bs[i];
        __syncthreads();
    }
    int Tp = T + 2 * pad;
    size_t n = (size_t)B * T * in;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx / in;
        int i = (int)(idx - r * in);
        int b = (int)(r / T), t = (int)(r - (size_t)b * T);
        float v;
        if (Ws) {
            const float* xr = x + r * in;
            v = sW[in * in + i];
            for (int j = 0; j < in; ++j) v = fmaf(sW[i * in + j], xr[j], v);
        } else {
            v = x[idx];
        }
        xp[((size_t)b * Tp + pad + t) * in + i] = v;
    }
}

// dx[b,t,j] = sum_i dxp[b,pad+t,i] Ws[i,j]   (or a copy)
__global__ void k_unpad_scale_bwd(int B, int T, int in, int pad, const float* __restrict__ dxp,
                                  const float* __restrict__ Ws, float* __restrict__ dx) {
    extern __shared__ float sW[];
    if (Ws) {
        for (int i = threadIdx.x; i < in * in; i += blockDim.x) sW[i] = Ws[i];
        __syncthreads();
    }
    int Tp = T + 2 * pad;
    size_t n = (size_t)B * T * in;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx / in;
        int j = (int)(idx - r * in);
        int b = (int)(r / T), t = (int)(r - (size_t)b * T);
        const float* g = dxp + ((size_t)b * Tp + pad + t) * in;
        float v;
        if (Ws) {
            v = 0.f;
            for (int i = 0; i < in; ++i) v = fmaf(g[i], sW[i * in + j], v);
        } else {
            v = g[j];
        }
        dx[idx] = v;
    }
}

// W [co][ci][k] -> Wr [k][co][ci]
__global__ void k_repack_w(int co, int ci, int k, const float* __restrict__ W, float* __restrict__ Wr) {
    size_t n = (size_t)co * ci * k;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        int j = (int)(idx / ((size_t)co * ci));
        size_t rem = idx - (size_t)j * co * ci;
        Wr[idx] = W[rem * k + j];
    }
}
// dWr [k][co][ci] -> dW [co][ci][k]
__global__ void k_unrepack_dw(int co, int ci, int k, const float* __restrict__ dWr, float* __restrict__ dW, int accumulate) {
    size_t n = (size_t)co * ci * k;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        size_t rem = idx / k;
        int j = (int)(idx - rem * k);
        float v = dWr[(size_t)j * co * ci + rem];
        dW[idx] = accumulate ? dW[idx] + v : v;
    }
}

// W [co][ci][k] -> Wcat [co][k*ci]  (column tap*ci + c): the B operand of the tap-fused product
__global__ void k_repack_wcat(int co, int ci, int k, const float* __restrict__ W, float* __restrict__ Wc) {
    size_t n = (size_t)co * ci * k;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        size_t o = idx / ((size_t)k * ci);
        int kk = (int)(idx - o * k * ci);
        int j = kk / ci, c = kk - j * ci;
        Wc[idx] = W[(o * ci + c) * k + j];
    }
}
// dWcat [co][k*ci] -> dW [co][ci][k]
__global__ void k_unrepack_dwcat(int co, int ci, int k, const float* __restrict__ dWc, float* __restrict__ dW, int accumulate) {
    size_t n = (size_t)co * ci * k;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        size_t rem = idx / k;   // o*ci + c
        int j = (int)(idx - rem * k);
        size_t o = rem / ci;
        int c = (int)(rem - o * ci);
        float v = dWc[o * k * ci + (size_t)j * ci + c];
        dW[idx] = accumulate ? dW[idx] + v : v;
    }
}
// din[m + r + (j-half)*d][c] += G[r][j*ci + c] over the taps, as a gather (no atomics): grid row p, channel c
__global__ void k_col2im_add(size_t R, size_t rows, size_t m, int ci, int k, int d, const float* __restrict__ G, float* __restrict__ din) {
    const int half = (k - 1) / 2;
    size_t n = R * ci;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        size_t p = idx / ci;
        int c = (int)(idx - p * ci);
        float v = 0.f;
        for (int j = 0; j < k; ++j) {
            long r = (long)p - (long)m - (long)(j - half) * d;
            if (r >= 0 && (size_t)r < rows) v += G[(size_t)r * k * ci + (size_t)j * ci + c];
        }
        din[idx] += v;
    }
}

// the k taps of a layer as ONE tensor-core product over the virtual im2col (K = k*ci) when that is the faster path
static bool conv_fused(size_t rows, int co, int ci, int k) {
    return want_tc_gemm() && gemm_tc_eligible((int)rows, co, ci * k);
}

// xc_tm[t,b,c] = xcp[b,pad+t,c] * mask_tm[t,b,c]
__global__ void k_compact_mask(int B, int T, int C, int pad, const float* __restrict__ xcp,
                               const float* __restrict__ mask, float* __restrict__ xc) {
    int Tp = T + 2 * pad;
    size_t n = (size_t)B * T * C;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx / C;
        int c = (int)(idx - r * C);
        int t = (int)(r / B), b = (int)(r - (size_t)t * B);
        float v = xcp[((size_t)b * Tp + pad + t) * C + c];
        if (mask) v *= mask[idx];
        xc[idx] = v;
    }
}
// dxcp[b,pad+t,c] = dxc_tm[t,b,c] * mask_tm[t,b,c]   (pad rows are pre-zeroed)
__global__ void k_expand_mask(int B, int T, int C, int pad, const float* __restrict__ dxc,
                              const float* __restrict__ mask, float* __restrict__ dxcp) {
    int Tp = T + 2 * pad;
    size_t n = (size_t)B * T * C;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx / C;
        int c = (int)(idx - r * C);
        int t = (int)(r / B), b = (int)(r - (size_t)t * B);
        float v = dxc[idx];
        if (mask) v *= mask[idx];
        dxcp[((size_t)b * Tp + pad + t) * C + c] = v;
    }
}

static inline int grid1d(size_t n) {
    size_t g = ceil_div_sz(n, 256);
    return (int)(g > 148 * 8 ? 148 * 8 : (g < 1 ? 1 : g));
}

// ================================================================================================
// composed path
// ================================================================================================
struct FeC {   // geometry of the composed front-end
    int k, L, in, pad, KT, Tp, CL;
    size_t R, rows;      // padded-grid rows; product rows (windows that stay inside the grid)
    size_t xp_floats;    // fe_ws: the padded grid (+ KT rows of slack)
};
static FeC fec_geom(const cvb_net* n, int B, int T) {
    FeC g;
    g.k = n->kernel_size;
    g.L = n->n_conv;
    g.in = n->in_dim;
    g.KT = ipow(g.k, g.L);
    g.pad = (g.KT - 1) / 2;
    g.Tp = T + 2 * g.pad;
    g.CL = g.in * g.KT;
    g.R = (size_t)B * g.Tp;
    g.rows = g.R - (size_t)(g.KT - 1);
    g.xp_floats = round_up_sz((g.R + g.KT) * g.in, 4);
    return g;
}
static size_t fec_xp_floats(const cvb_net* n, int B, int T) { return fec_geom(n, B, T).xp_floats; }
static bool fe_composed(const cvb_net* n, int B, int T) {
    if (B <= 0 || T <= 0) return false;
    FeC g = fec_geom(n, B, T);
    return want_tc_gemm() && g.R > (size_t)g.KT && gemm_tc_eligible((int)g.rows, g.CL, g.CL);
}

// xp[b, tau, i]: zero in the pad rows, scale_in(x[b, tau - pad]) (or a copy) inside; one thread per (row, 4 channels)
__global__ void k_pad_scale_all(int B, int T, int in, int pad, const float* __restrict__ x, const float* __restrict__ Ws,
                                const float* __restrict__ bs, float* __restrict__ xp) {
    extern __shared__ float sW[];  // [in*in + in]
    if (Ws) {
        for (int i = threadIdx.x; i < in * in; i += blockDim.x) sW[i] = Ws[i];
        for (int i = threadIdx.x; i < in; i += blockDim.x) sW[in * in + i] = bs[i];
        __syncthreads();
    }
    const int Tp = T + 2 * pad;
    const int q4 = (in + 3) / 4;
    const size_t n = (size_t)B * Tp * q4;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t r = idx / q4;
        const int i0 = (int)(idx - r * q4) * 4;
        const int b = (int)(r / Tp), tau = (int)(r - (size_t)b * Tp), t = tau - pad;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (t >= 0 && t < T) {
            const float* xr = x + ((size_t)b * T + t) * in;
            if (Ws) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (i0 + q < in) v[q] = sW[in * in + i0 + q];
                for (int j = 0; j < in; ++j) {
                    const float xv = xr[j];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (i0 + q < in) v[q] = fmaf(sW[(i0 + q) * in + j], xv, v[q]);
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (i0 + q < in) v[q] = xr[i0 + q];
            }
        }
        float* d = xp + r * in + i0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (i0 + q < in) d[q] = v[q];
    }
}

// dxcp[b, t, :] = dxc_tm[t, b, :] * mask_tm[t, b, :] for t < T, 0 for the Tp - T trailing rows of each utterance
__global__ void k_expand_mask_all(int B, int T, int Tp, int C, const float* __restrict__ dxc, const float* __restrict__ mask,
                                  float* __restrict__ dxcp) {
    const size_t n = (size_t)B * Tp * C;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t r = idx / C;
        const int c = (int)(idx - r * C);
        const int b = (int)(r / Tp), t = (int)(r - (size_t)b * Tp);
        float v = 0.f;
        if (t < T) {
            const size_t src = ((size_t)t * B + b) * C + c;
            v = dxc[src];
            if (mask) v *= mask[src];
        }
        dxcp[idx] = v;
    }
}

// bb_next[o] = b_layer[o] + sum_j sum_q Wt[j][o][q] bb[q]     (one warp per output)
__global__ void k_beff(int co, int ci, int k, const float* __restrict__ Wt, const float* __restrict__ b_layer,
                       const float* __restrict__ bb, float* __restrict__ bb_next) {
    const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (o >= co) return;
    float s = 0.f;
    for (int j = 0; j < k; ++j)
        for (int q = lane; q < ci; q += 32) s = fmaf(Wt[((size_t)j * co + o) * ci + q], bb[q], s);
    s = warp_sum(s);
    if (lane == 0) bb_next[o] = b_layer[o] + s;
}
// dbb[q] = sum_j sum_o Wt[j][o][q] dbb_next[o]: 32 columns q per block, the (j, o) rows spread over 32 warps, fixed-order tree
__global__ void __launch_bounds__(1024) k_beff_bwd(int co, int ci, int k, const float* __restrict__ Wt, const float* __restrict__ dbb_next,
                                                   float* __restrict__ dbb) {
    __shared__ float red[32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (q < ci)
        for (int r = w; r < k * co; r += 32) s = fmaf(Wt[(size_t)r * ci + q], dbb_next[r % co], s);
    red[w][lane] = s;
    __syncthreads();
    if (w == 0 && q < ci) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) t += red[i][lane];
        dbb[q] = t;
    }
}
// dWt[j][o][q] += dbb_next[o] * bb[q] for every tap j
__global__ void k_rank1_add(int co, int ci, int k, const float* __restrict__ dbb_next, const float* __restrict__ bb, float* __restrict__ dWt) {
    const size_t n = (size_t)k * co * ci;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const int q = (int)(idx % ci);
        const int o = (int)((idx / ci) % co);
        dWt[idx] += dbb_next[o] * bb[q];
    }
}
// Tm[j][o][n] = E[o][j * Cl + n]: the k column blocks of a C_{l+1} x C_{l+1} matrix stacked as one [k * C_{l+1}, C_l] matrix
__global__ void k_tapmajor(int Cn, int Cl, int k, const float* __restrict__ E, float* __restrict__ Tm) {
    const size_t n = (size_t)k * Cn * Cl;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % Cl);
        const int o = (int)((idx / Cl) % Cn);
        const int j = (int)(idx / ((size_t)Cl * Cn));
        Tm[idx] = E[(size_t)o * Cn + (size_t)j * Cl + c];
    }
}

// dst = src (+ dst)
__global__ void k_copy_acc(size_t n, const float* __restrict__ src, float* __restrict__ dst, int accumulate) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = accumulate ? dst[i] + src[i] : src[i];
}

// ---- derived parameters of the composed conv, cached until the weights change --------------------------------------
// buffer layout (floats): E_1 .. E_L (C_l x C_l, column = tap * in + channel), bb_1 .. bb_L (C_l), Wt_1 .. Wt_{L-1}
// ([k][C_{l+1}][C_l]).  E_1 = wcat(W_0), bb_1 = b_0;  E_{l+1}[:, block j] = Wt_l[j] E_l,  bb_{l+1} = b_l + sum_j Wt_l[j] bb_l.
struct FeDerived {
    const float* key_w[4] = {nullptr, nullptr, nullptr, nullptr};
    int k = 0, L = 0, in = 0;
    unsigned long long gen = 0, last_use = 0;
    float* buf = nullptr;
    size_t floats = 0;
    bool pinned = false;
};
struct FeDLayout {
    size_t E[6], bb[6], Wt[5], total;
};
static FeDLayout fed_layout(int in, int k, int L) {
    FeDLayout D;
    size_t off = 0;
    for (int l = 1; l <= L; ++l) {
        size_t C = (size_t)in * ipow(k, l);
        D.E[l] = off;
        off += round_up_sz(C * C, 4);
    }
    for (int l = 1; l <= L; ++l) {
        D.bb[l] = off;
        off += round_up_sz((size_t)in * ipow(k, l), 4);
    }
    for (int l = 1; l < L; ++l) {
        D.Wt[l] = off;
        off += round_up_sz((size_t)k * in * ipow(k, l + 1) * in * ipow(k, l), 4);
    }
    D.total = off;
    return D;
}
static std::mutex g_fed_mu;
static FeDerived g_fed[64][8];
static unsigned long long g_fed_clock = 0;

// -> *out = buffer holding the derived parameters of `net`, recomputed on stream s if the weights changed
static int fe_derived(const cvb_net* net, cudaStream_t s, const float** out) {
    int dev = 0;
    CVB_CHECK(cudaGetDevice(&dev));
    CVB_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
    const int in = net->in_dim, k = net->kernel_size, L = net->n_conv;
    const FeDLayout D = fed_layout(in, k, L);
    cudaStreamCaptureStatus cap_st = cudaStreamCaptureStatusNone;
    CVB_CHECK(cudaStreamIsCapturing(s, &cap_st));
    const bool capturing = cap_st != cudaStreamCaptureStatusNone;
    const unsigned long long gen = weights_generation();
    std::lock_guard<std::mutex> lk(g_fed_mu);
    FeDerived* e = nullptr;
    for (auto& c : g_fed[dev]) {
        bool same = c.buf && c.k == k && c.L == L && c.in == in;
        for (int i = 0; i < L && same; ++i) same = c.key_w[i] == net->conv_w[i];
        if (same) e = &c;
    }
    if (!e) {
        for (auto& c : g_fed[dev])
            if (!c.pinned && !(capturing && c.floats < D.total) && (!e || c.last_use < e->last_use)) e = &c;
        CVB_REQUIRE(e, "no free slot for the composed conv parameters (8 front-ends are pinned by captured CUDA graphs)");
        if (e->floats < D.total) {
            if (e->buf) CVB_CHECK(cudaFree(e->buf));
            e->buf = nullptr;
            e->floats = 0;
            CVB_CHECK(cudaMalloc(&e->buf, D.total * sizeof(float)));
            e->floats = D.total;
        }
        for (int i = 0; i < 4; ++i) e->key_w[i] = i < L ? net->conv_w[i] : nullptr;
        e->k = k;
        e->L = L;
        e->in = in;
        e->gen = 0;
    }
    e->last_use = ++g_fed_clock;
    if (capturing) e->pinned = true;
    *out = e->buf;
    if (e->gen == gen) return 0;
    e->gen = gen;
    float* buf = e->buf;
    const int C1 = in * k;
    k_repack_wcat<<<grid1d((size_t)C1 * in * k), 256, 0, s>>>(C1, in, k, net->conv_w[0], buf + D.E[1]);
    CVB_LAUNCH_CHECK();
    CVB_CHECK(cudaMemcpyAsync(buf + D.bb[1], net->conv_b[0], (size_t)C1 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    for (int l = 1; l < L; ++l) {
        const int Cl = in * ipow(k, l), Cn = Cl * k;
        float* Wt = buf + D.Wt[l];
        k_repack_w<<<grid1d((size_t)Cn * Cl * k), 256, 0, s>>>(Cn, Cl, k, net->conv_w[l], Wt);
        CVB_LAUNCH_CHECK();
        for (int j0 = 0; j0 < k; j0 += 6) {   // E_{l+1}[:, block j] = Wt[j] E_l, up to 6 taps per launch
            GemmDesc d[6];
            int n = 0;
            for (int j = j0; j < k && n < 6; ++j) {
                GemmDesc& g = d[n++];
                g.M = Cn; g.N = Cl; g.K = Cl;
                g.A = Wt + (size_t)j * Cn * Cl; g.lda = Cl;
                g.B = buf + D.E[l]; g.ldb = Cl;
                g.C = buf + D.E[l + 1] + (size_t)j * Cl; g.ldc = Cn;
            }
            if (gemm_tc_eligible(Cn, Cl, Cl)) {
                if (int rc = gemm_tc_group(s, d, n)) return rc;
            } else {
                for (int i = 0; i < n; ++i)
                    if (int rc = gemm_rm(s, false, false, Cn, Cl, Cl, 1.f, d[i].A, Cl, d[i].B, Cl, 0.f, d[i].C, Cn)) return rc;
            }
        }
        k_beff<<<ceil_div(Cn, 8), 256, 0, s>>>(Cn, Cl, k, Wt, net->conv_b[l], buf + D.bb[l], buf + D.bb[l + 1]);
        CVB_LAUNCH_CHECK();
    }
    return 0;
}

static size_t fec_bwd_scratch_floats(const cvb_net* n, int B, int T) {
    FeC g = fec_geom(n, B, T);
    FeDLayout D = fed_layout(g.in, g.k, g.L);
    // dxcp [R, CL] | G [rows, CL] | dxp [R + KT, in] | dE_1..dE_L, dbb_1..dbb_L, dWt_1..dWt_{L-1} (same layout as the derived parameters)
    return round_up_sz(g.R * g.CL, 4) + round_up_sz(g.rows * g.CL, 4) + g.xp_floats + D.total + round_up_sz((size_t)g.CL * g.CL, 4);
}

static int fec_fwd(const cvb_net* net, int B, int T, const float* x_bm, const float* mask_conv_tm, float* fe_ws, float* xc_tm,
                   cudaStream_t s) {
    const FeC g = fec_geom(net, B, T);
    const FeDLayout D = fed_layout(g.in, g.k, g.L);
    const float* der = nullptr;
    if (int rc = fe_derived(net, s, &der)) return rc;
    const float* Ws = net->has_scale_in ? net->scale_in_w : nullptr;
    const size_t smem = Ws ? (size_t)(g.in * g.in + g.in) * sizeof(float) : 0;
    CVB_REQUIRE(smem <= 48 * 1024, "scale_in matrix too large for shared memory (in_dim=%d)", g.in);
    float* xp = fe_ws;
    k_pad_scale_all<<<grid1d(g.R * ((g.in + 3) / 4)), 256, smem, s>>>(B, T, g.in, g.pad, x_bm, Ws, net->scale_in_b, xp);
    CVB_LAUNCH_CHECK();
    if (int rc = zero_floats(s, xp + g.R * g.in, (size_t)g.KT * g.in)) return rc;   // slack rows behind the last window
    ConvGather ga{g.in, 1};
    GemmDesc d;
    d.transB = true;
    d.M = (int)g.rows; d.N = g.CL; d.K = g.CL;
    d.A = xp; d.lda = g.in; d.gA = &ga;
    d.B = der + D.E[g.L]; d.ldb = g.CL; d.b_const = true;
    d.bias = der + D.bb[g.L];
    d.C = xc_tm; d.ldc = g.CL;
    d.map_Tp = g.Tp; d.map_T = T; d.map_B = B; d.mask = mask_conv_tm;
    return gemm_tc_group(s, &d, 1);
}

static int fec_bwd(const cvb_net* net, int B, int T, const float* x_bm, const float* mask_conv_tm, const float* fe_ws,
                   const float* dxc_tm, float* scratch, float* dx_bm, const cvb_net_grads* gr, cudaStream_t s) {
    const FeC g = fec_geom(net, B, T);
    const FeDLayout D = fed_layout(g.in, g.k, g.L);
    const float* der = nullptr;
    if (int rc = fe_derived(net, s, &der)) return rc;
    const int acc = gr ? gr->accumulate : 0;
    const float* xp = fe_ws;
    float* dxcp = scratch;
    float* G = dxcp + round_up_sz(g.R * g.CL, 4);
    float* dxp = G + round_up_sz(g.rows * g.CL, 4);
    float* dd = dxp + g.xp_floats;   // gradients of the derived parameters, same layout as `der`
    k_expand_mask_all<<<grid1d(g.R * g.CL), 256, 0, s>>>(B, T, g.Tp, g.CL, dxc_tm, mask_conv_tm, dxcp);
    CVB_LAUNCH_CHECK();
    bool want_w = false;
    if (gr)
        for (int i = 0; i < g.L; ++i) want_w = want_w || gr->conv_w[i] || gr->conv_b[i];
    const bool need_dx = dx_bm || (gr && (gr->scale_in_w || gr->scale_in_b));
    GemmDesc d[2];
    int n = 0;
    ConvGather gb{g.in, 1};
    if (want_w) {   // dE_L = dxc'^T im2col(x^)
        GemmDesc& q = d[n++];
        q.transA = true;
        q.M = g.CL; q.N = g.CL; q.K = (int)g.rows;
        q.A = dxcp; q.lda = g.CL;
        q.B = xp; q.ldb = g.in; q.gB = &gb;
        q.C = dd + D.E[g.L]; q.ldc = g.CL; q.f16 = false;
    }
    if (need_dx) {   // G = dxc' E_L: per window row the gradient of its KT x in inputs
        GemmDesc& q = d[n++];
        q.M = (int)g.rows; q.N = g.CL; q.K = g.CL;
        q.A = dxcp; q.lda = g.CL;
        q.B = der + D.E[g.L]; q.ldb = g.CL; q.b_const = true;
        q.C = G; q.ldc = g.CL; q.f16 = false;
    }
    if (n)
        if (int rc = gemm_tc_group(s, d, n)) return rc;
    if (need_dx) {
        if (int rc = zero_floats(s, dxp, g.R * g.in)) return rc;
        k_col2im_add<<<grid1d(g.R * g.in), 256, 0, s>>>(g.R, g.rows, (size_t)g.pad, g.in, g.KT, 1, G, dxp);
        CVB_LAUNCH_CHECK();
    }
    if (want_w) {
        if (int rc = colsum(s, dxcp, (int)g.rows, g.CL, g.CL, dd + D.bb[g.L], false)) return rc;
        // chain rule from (dE_L, dbb_L) down to the layers
        for (int l = g.L - 1; l >= 1; --l) {
            const int Cl = g.in * ipow(g.k, l), Cn = Cl * g.k;
            const float* Wt = der + D.Wt[l];
            float* dWt = dd + D.Wt[l];
            {
                // dWt[j] = dE_{l+1}[:, block j] E_l^T for every tap j, and dE_l = sum_j Wt[j]^T dE_{l+1}[:, block j] as ONE
                // contraction over (j, o) against the tap-major copy of dE_{l+1}: one grouped launch
                float* Tm = dd + D.total;   // [k * Cn, Cl]
                k_tapmajor<<<grid1d((size_t)g.k * Cn * Cl), 256, 0, s>>>(Cn, Cl, g.k, dd + D.E[l + 1], Tm);
                CVB_LAUNCH_CHECK();
                GemmDesc q[6];
                int nq = 0;
                CVB_REQUIRE(g.k <= 5, "kernel_size %d: at most 5 taps per grouped launch", g.k);
                for (int j = 0; j < g.k; ++j) {
                    GemmDesc& a = q[nq++];
                    a.transB = true;
                    a.M = Cn; a.N = Cl; a.K = Cl;
                    a.A = dd + D.E[l + 1] + (size_t)j * Cl; a.lda = Cn;
                    a.B = der + D.E[l]; a.ldb = Cl;
                    a.C = dWt + (size_t)j * Cn * Cl; a.ldc = Cl; a.f16 = false;
                }
                GemmDesc& b = q[nq++];
                b.transA = true;
                b.M = Cl; b.N = Cl; b.K = g.k * Cn;
                b.A = Wt; b.lda = Cl;
                b.B = Tm; b.ldb = Cl;
                b.C = dd + D.E[l]; b.ldc = Cl; b.f16 = false;
                if (int rc = gemm_tc_group(s, q, nq)) return rc;
            }
            k_rank1_add<<<grid1d((size_t)g.k * Cn * Cl), 256, 0, s>>>(Cn, Cl, g.k, dd + D.bb[l + 1], der + D.bb[l], dWt);
            CVB_LAUNCH_CHECK();
            k_beff_bwd<<<ceil_div(Cl, 32), 1024, 0, s>>>(Cn, Cl, g.k, Wt, dd + D.bb[l + 1], dd + D.bb[l]);
            CVB_LAUNCH_CHECK();
            if (gr->conv_w[l]) {
                k_unrepack_dw<<<grid1d((size_t)Cn * Cl * g.k), 256, 0, s>>>(Cn, Cl, g.k, dWt, gr->conv_w[l], acc);
                CVB_LAUNCH_CHECK();
            }
            if (gr->conv_b[l]) {
                k_copy_acc<<<grid1d((size_t)Cn), 256, 0, s>>>((size_t)Cn, dd + D.bb[l + 1], gr->conv_b[l], acc);
                CVB_LAUNCH_CHECK();
            }
        }
        const int C1 = g.in * g.k;
        if (gr->conv_w[0]) {
            k_unrepack_dwcat<<<grid1d((size_t)C1 * g.in * g.k), 256, 0, s>>>(C1, g.in, g.k, dd + D.E[1], gr->conv_w[0], acc);
            CVB_LAUNCH_CHECK();
        }
        if (gr->conv_b[0]) {
            k_copy_acc<<<grid1d((size_t)C1), 256, 0, s>>>((size_t)C1, dd + D.bb[1], gr->conv_b[0], acc);
            CVB_LAUNCH_CHECK();
        }
    }
    const float* Ws = net->has_scale_in ? net->scale_in_w : nullptr;
    const int in_dim = g.in;
    if (dx_bm) {
        size_t smem = Ws ? (size_t)in_dim * in_dim * sizeof(float) : 0;
        k_unpad_scale_bwd<<<grid1d((size_t)B * T * in_dim), 256, smem, s>>>(B, T, in_dim, g.pad, dxp, Ws, dx_bm);
        CVB_LAUNCH_CHECK();
    }
    if (Ws && gr && (gr->scale_in_w || gr->scale_in_b)) {
        // frozen in the trainer (train_*.py:369-370); provided for completeness, one product per utterance
        for (int b = 0; b < B; ++b) {
            const float* dxh = dxp + ((size_t)b * g.Tp + g.pad) * in_dim;
            bool first = (b == 0) && !acc;
            if (gr->scale_in_w)
                if (int rc = gemm_rm(s, true, false, in_dim, in_dim, T, 1.f, dxh, in_dim, x_bm + (size_t)b * T * in_dim, in_dim,
                                     first ? 0.f : 1.f, gr->scale_in_w, in_dim, true))
                    return rc;
            if (gr->scale_in_b)
                if (int rc = colsum(s, dxh, T, in_dim, in_dim, gr->scale_in_b, !first)) return rc;
        }
    }
    return 0;
}

int frontend_fwd(const cvb_net* net, int B, int T, const float* x_bm, const float* mask_conv_tm, float* fe_ws,
                 float* xc_tm, cudaStream_t s) {
    CVB_REQUIRE(net->kernel_size % 2 == 1, "kernel_size must be odd (got %d)", net->kernel_size);
    CVB_REQUIRE(net->n_conv >= 1 && net->n_conv <= 4, "dilation_size (conv layers) must be 1..4 (got %d)", net->n_conv);
    FeGeom g = fe_geom(net, B, T);
    if (g.R == 0 || T == 0) return 0;
    if (fe_composed(net, B, T)) return fec_fwd(net, B, T, x_bm, mask_conv_tm, fe_ws, xc_tm, s);
    // zero the padded grids (pads of buf_0 are the reference's zero padding; the rest keeps
    // never-computed edge rows finite)
    if (int rc = zero_floats(s, fe_ws, g.wr_off[0])) return rc;
    const float* Ws = net->has_scale_in ? net->scale_in_w : nullptr;
    size_t smem = Ws ? (size_t)(net->in_dim * net->in_dim + net->in_dim) * sizeof(float) : 0;
    CVB_REQUIRE(smem <= 48 * 1024, "scale_in matrix too large for shared memory (in_dim=%d)", net->in_dim);
    k_pad_scale<<<grid1d((size_t)B * T * net->in_dim), 256, smem, s>>>(B, T, net->in_dim, g.pad, x_bm, Ws,
                                                                      net->scale_in_b, fe_ws + g.buf_off[0]);
    CVB_LAUNCH_CHECK();
    for (int i = 0; i < g.L; ++i) {
        int ci = g.C[i], co = g.C[i + 1], d = ipow(g.k, i), half = (g.k - 1) / 2;
        size_t m = (size_t)half * d;  // rows skipped at both ends of the flattened grid
        CVB_REQUIRE(g.R > 2 * m, "sequence too short for the receptive field");
        float* Wr = fe_ws + g.wr_off[i];
        const float* in = fe_ws + g.buf_off[i];
        float* out = fe_ws + g.buf_off[i + 1];
        size_t rows = g.R - 2 * m;
        if (conv_fused(rows, co, ci, g.k)) {
            k_repack_wcat<<<grid1d((size_t)co * ci * g.k), 256, 0, s>>>(co, ci, g.k, net->conv_w[i], Wr);
            CVB_LAUNCH_CHECK();
            ConvGather ga{ci, d};
            if (int rc = gemm_tc(s, false, true, (int)rows, co, g.k * ci, in + (m - (size_t)half * d) * ci, ci, Wr, g.k * ci, false,
                                 net->conv_b[i], out + m * co, co, true, nullptr, 0, 0, &ga, nullptr))
                return rc;
            continue;
        }
        k_repack_w<<<grid1d((size_t)co * ci * g.k), 256, 0, s>>>(co, ci, g.k, net->conv_w[i], Wr);
        CVB_LAUNCH_CHECK();
        if (int rc = fill_rows(s, out + m * co, rows, co, co, net->conv_b[i])) return rc;
        for (int j = 0; j < g.k; ++j) {
            long shift = (long)(j - half) * d;
            if (int rc = gemm_rm(s, false, true, (int)rows, co, ci, 1.f, in + (size_t)((long)m + shift) * ci, ci,
                                 Wr + (size_t)j * co * ci, ci, 1.f, out + m * co, co))
                return rc;
        }
    }
    int C = g.C[g.L];
    k_compact_mask<<<grid1d((size_t)B * T * C), 256, 0, s>>>(B, T, C, g.pad, fe_ws + g.buf_off[g.L], mask_conv_tm, xc_tm);
    CVB_LAUNCH_CHECK();
    return 0;
}

// dxc_tm [T,B,C_L] (un-masked gradient of the dropped conv output) -> dx_bm, conv / scale_in grads
int frontend_bwd(const cvb_net* net, int B, int T, const float* x_bm, const float* mask_conv_tm, const float* fe_ws,
                 const float* dxc_tm, float* scratch, float* dx_bm, const cvb_net_grads* gr, cudaStream_t s) {
    FeGeom g = fe_geom(net, B, T);
    if (g.R == 0 || T == 0) return 0;
    if (fe_composed(net, B, T)) return fec_bwd(net, B, T, x_bm, mask_conv_tm, fe_ws, dxc_tm, scratch, dx_bm, gr, s);
    int acc = gr ? gr->accumulate : 0;
    float* dbuf = scratch;  // same offsets as buf_i
    float* dWr = scratch + g.wr_off[0];
    size_t dwr_floats = 0;
    for (int i = 0; i < g.L; ++i) dwr_floats = dwr_floats > (size_t)g.k * g.C[i + 1] * g.C[i] ? dwr_floats : (size_t)g.k * g.C[i + 1] * g.C[i];
    if (int rc = zero_floats(s, dbuf, g.wr_off[0])) return rc;
    int CL = g.C[g.L];
    k_expand_mask<<<grid1d((size_t)B * T * CL), 256, 0, s>>>(B, T, CL, g.pad, dxc_tm, mask_conv_tm, dbuf + g.buf_off[g.L]);
    CVB_LAUNCH_CHECK();
    for (int i = g.L - 1; i >= 0; --i) {
        int ci = g.C[i], co = g.C[i + 1], d = ipow(g.k, i), half = (g.k - 1) / 2;
        size_t m = (size_t)half * d;
        size_t rows = g.R - 2 * m;
        const float* in = fe_ws + g.buf_off[i];
        const float* Wr = fe_ws + g.wr_off[i];
        const float* dout = dbuf + g.buf_off[i + 1] + m * co;
        float* din = dbuf + g.buf_off[i];
        bool want_w = gr && gr->conv_w[i];
        bool need_din = (i > 0) || dx_bm || (gr && (gr->scale_in_w || gr->scale_in_b));
        if (conv_fused(rows, co, ci, g.k)) {   // Wr holds Wcat [co][k*ci] (written by the forward)
            ConvGather gb{ci, d};
            if (want_w) {   // dWcat = dout^T im2col(in)
                if (int rc = gemm_tc(s, true, false, co, g.k * ci, (int)rows, dout, co, in + (m - (size_t)half * d) * ci, ci, false, nullptr,
                                     dWr, g.k * ci, false, nullptr, 0, 0, nullptr, &gb))
                    return rc;
                k_unrepack_dwcat<<<grid1d((size_t)co * ci * g.k), 256, 0, s>>>(co, ci, g.k, dWr, gr->conv_w[i], acc);
                CVB_LAUNCH_CHECK();
            }
            if (need_din) {   // G = dout Wcat, then the taps are gathered back onto the grid
                float* G = dWr + round_up_sz(dwr_floats, 4);
                if (int rc = gemm_tc(s, false, false, (int)rows, g.k * ci, co, dout, co, Wr, g.k * ci, false, nullptr, G, g.k * ci, false))
                    return rc;
                k_col2im_add<<<grid1d(g.R * ci), 256, 0, s>>>(g.R, rows, m, ci, g.k, d, G, din);
                CVB_LAUNCH_CHECK();
            }
        } else {
        for (int j = 0; j < g.k; ++j) {
            long shift = (long)(j - half) * d;
            if (want_w)
                if (int rc = gemm_rm(s, true, false, co, ci, (int)rows, 1.f, dout, co, in + (size_t)((long)m + shift) * ci, ci,
                                     0.f, dWr + (size_t)j * co * ci, ci, true))
                    return rc;
            if (need_din)
                if (int rc = gemm_rm(s, false, false, (int)rows, ci, co, 1.f, dout, co, Wr + (size_t)j * co * ci, ci, 1.f,
                                     din + (size_t)((long)m + shift) * ci, ci, true))
                    return rc;
        }
        if (want_w) {
            k_unrepack_dw<<<grid1d((size_t)co * ci * g.k), 256, 0, s>>>(co, ci, g.k, dWr, gr->conv_w[i], acc);
            CVB_LAUNCH_CHECK();
        }
        }
        if (gr && gr->conv_b[i])
            if (int rc = colsum(s, dout, (int)rows, co, co, gr->conv_b[i], acc != 0)) return rc;
    }
    const float* Ws = net->has_scale_in ? net->scale_in_w : nullptr;
    int in_dim = net->in_dim;
    if (dx_bm) {
        size_t smem = Ws ? (size_t)in_dim * in_dim * sizeof(float) : 0;
        k_unpad_scale_bwd<<<grid1d((size_t)B * T * in_dim), 256, smem, s>>>(B, T, in_dim, g.pad, dbuf + g.buf_off[0], Ws, dx_bm);
        CVB_LAUNCH_CHECK();
    }
    if (Ws && gr && (gr->scale_in_w || gr->scale_in_b)) {
        // frozen in the trainer (train_*.py:369-370); provided for completeness, one product per utterance
        for (int b = 0; b < B; ++b) {
            const float* dxh = dbuf + g.buf_off[0] + ((size_t)b * g.Tp + g.pad) * in_dim;
            bool first = (b == 0) && !acc;
            if (gr->scale_in_w)
                if (int rc = gemm_rm(s, true, false, in_dim, in_dim, T, 1.f, dxh, in_dim, x_bm + (size_t)b * T * in_dim, in_dim,
                                     first ? 0.f : 1.f, gr->scale_in_w, in_dim, true))
                    return rc;
            if (gr->scale_in_b)
                if (int rc = colsum(s, dxh, T, in_dim, in_dim, gr->scale_in_b, !first)) return rc;
        }
    }
    return 0;
}

}  // namespace cvb

extern "C" {
size_t cvb_frontend_ws_floats(const cvb_net* net, int B, int T) { return cvb::frontend_ws_floats(net, B, T); }

int cvb_frontend_fwd(const cvb_net* net, int B, int T, const float* x_bm, const float* mask_conv_tm, float* fe_ws,
                     float* xc_tm, void* stream) {
    return cvb::frontend_fwd(net, B, T, x_bm, mask_conv_tm, fe_ws, xc_tm, (cudaStream_t)stream);
}

size_t cvb_frontend_bwd_ws_floats(const cvb_net* net, int B, int T) { return cvb::frontend_bwd_scratch_floats(net, B, T); }

int cvb_frontend_bwd(const cvb_net* net, int B, int T, const float* x_bm, const float* mask_conv_tm, const float* fe_ws,
                     const float* dxc_tm, float* scratch, float* dx_bm, const cvb_net_grads* grads, void* stream) {
    CVB_REQUIRE(net && x_bm && fe_ws && dxc_tm && scratch, "cvb_frontend_bwd: NULL argument");
    return cvb::frontend_bwd(net, B, T, x_bm, mask_conv_tm, fe_ws, dxc_tm, scratch, dx_bm, grads, (cudaStream_t)stream);
}
}
