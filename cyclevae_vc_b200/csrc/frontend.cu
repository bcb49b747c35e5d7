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

size_t frontend_ws_floats(const cvb_net* n, int B, int T) { return fe_geom(n, B, T).total; }
size_t frontend_xc_offset(const cvb_net* n, int B, int T) { return fe_geom(n, B, T).xc_off; }
// floats of the backward's gradient grid (d buf_0..d buf_L) + one repacked weight-gradient
size_t frontend_bwd_scratch_floats(const cvb_net* n, int B, int T) {
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

int frontend_fwd(const cvb_net* net, int B, int T, const float* x_bm, const float* mask_conv_tm, float* fe_ws,
                 float* xc_tm, cudaStream_t s) {
    CVB_REQUIRE(net->kernel_size % 2 == 1, "kernel_size must be odd (got %d)", net->kernel_size);
    CVB_REQUIRE(net->n_conv >= 1 && net->n_conv <= 4, "dilation_size (conv layers) must be 1..4 (got %d)", net->n_conv);
    FeGeom g = fe_geom(net, B, T);
    if (g.R == 0 || T == 0) return 0;
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
}
