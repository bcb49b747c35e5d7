// GRU_RNN.forward / backward composition behind the C ABI (include/cyclevae_b200.h).
#include <stdlib.h>

#include "gru_ar.cuh"

namespace cvb {

// frontend.cu
size_t frontend_ws_floats(const cvb_net* n, int B, int T);
size_t frontend_xc_offset(const cvb_net* n, int B, int T);
size_t frontend_bwd_scratch_floats(const cvb_net* n, int B, int T);
int frontend_fwd(const cvb_net* net, int B, int T, const float* x_bm, const float* mask_conv_tm, float* fe_ws,
                 float* xc_tm, cudaStream_t s);
int frontend_bwd(const cvb_net* net, int B, int T, const float* x_bm, const float* mask_conv_tm, const float* fe_ws,
                 const float* dxc_tm, float* scratch, float* dx_bm, const cvb_net_grads* gr, cudaStream_t s);

#define LOG_VAR_FLOOR (-13.815510557964274f)  // gru_vae.py:412

static inline size_t r4(size_t n) { return round_up_sz(n, 4); }

struct RecLayout {
    size_t hs, ys, r, z, n, ghn, o, total;
};
static RecLayout rec_layout(const cvb_net* net, int B, int T, bool training, bool has_mask) {
    RecLayout L;
    size_t H = net->hidden, out = net->out_dim, TB = (size_t)T * B;
    size_t off = 0;
    L.hs = off; off += r4((size_t)(T + 1) * B * H);
    L.ys = off; off += r4((size_t)(T + 1) * B * out);
    L.r = L.z = L.n = L.ghn = L.o = 0;
    if (training) {
        L.r = off; off += r4(TB * H);
        L.z = off; off += r4(TB * H);
        L.n = off; off += r4(TB * H);
        L.ghn = off; off += r4(TB * H);
        if (has_mask) { L.o = off; off += r4(TB * H); }
    }
    L.total = off;
    return L;
}

struct FwdScratch {
    size_t gx, part, bar, tc, tc_floats, total;
};
static FwdScratch fwd_scratch(const cvb_net* net, int B, int T) {
    FwdScratch S;
    size_t H = net->hidden, out = net->out_dim, TB = (size_t)T * B;
    size_t off = 0;
    S.gx = off; off += r4(TB * 3 * H);
    S.part = off; off += r4((size_t)gru_exact_grid(net->hidden) * B * out);
    S.bar = off; off += 16;
    S.tc = off;
    // sized by the same shape predicates that pick the kernel: the two-exchange kernel needs out <= 64, the folded
    // inference kernel holds any out_dim (e.g. the encoder at lat_dim 50 / 64, egs/one-to-one/run.sh)
    size_t a = gru_tc_shape_ok(B, net->hidden, net->out_dim) ? gru_tc_scratch_floats(B, net->hidden) : 0;
    size_t e = gru_tc_eval_shape_ok(B, net->hidden) ? gru_tc_eval_scratch_floats(B, net->hidden) + r4((size_t)6 * H) : 0;   // W_fb | hx | ctr | c_fb | b_ih + c_fb
    S.tc_floats = r4(a > e ? a : e);
    off += S.tc_floats;
    S.total = off;
    return S;
}
// CVB_EVAL_FOLD=0 keeps the two-exchange kernel for inference too (A/B testing)
static bool want_fold() {
    const char* e = getenv("CVB_EVAL_FOLD");
    return !(e && e[0] == '0');
}
// CVB_RECURRENCE=exact forces the fp32-FMA kernels; default = tensor-core variant where the shape allows
static bool want_tc() {
    const char* e = getenv("CVB_RECURRENCE");
    return !(e && (e[0] == 'e' || e[0] == 'E'));
}
struct BwdScratch {
    size_t dy_tot, dgi, dghn, gxch, dhc, part, bar, dxc, fe, dtrj, tc, total;
};
static BwdScratch bwd_scratch(const cvb_net* net, int B, int T) {
    BwdScratch S;
    size_t H = net->hidden, out = net->out_dim, TB = (size_t)T * B, C = conv_dim(net);
    size_t off = 0;
    S.dy_tot = off; off += r4((size_t)(T + 1) * B * out);
    S.dgi = off; off += r4(TB * 3 * H);
    S.dghn = off; off += r4(TB * H);
    S.gxch = off; off += r4((size_t)2 * B * 3 * H);
    S.dhc = off; off += r4((size_t)B * H);
    S.part = off; off += r4((size_t)gru_exact_grid(net->hidden) * B * out);
    S.bar = off; off += 16;
    S.dxc = off; off += r4(TB * C);
    S.fe = off; off += r4(frontend_bwd_scratch_floats(net, B, T));
    S.dtrj = off; off += r4(TB * out);
    S.tc = off;
    if (gru_tc_bwd_shape_ok(B, net->hidden, net->out_dim)) off += r4(gru_tc_bwd_scratch_floats(B, net->hidden));
    S.total = off;
    return S;
}

// ---- head: y (tm) -> trj_out (bm) -----------------------------------------------------------
__global__ void k_head_fwd(int B, int T, int out, int mode, int lat, const float* __restrict__ y_tm,
                           const float* __restrict__ Ws, const float* __restrict__ bs, float* __restrict__ trj_bm) {
    extern __shared__ float sW[];  // [out*out + out] in scale_out mode
    if (mode == CVB_HEAD_SCALE_OUT) {
        for (int i = threadIdx.x; i < out * out; i += blockDim.x) sW[i] = Ws[i];
        for (int i = threadIdx.x; i < out; i += blockDim.x) sW[out * out + i] = bs[i];
        __syncthreads();
    }
    size_t n = (size_t)B * T * out;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx / out;
        int i = (int)(idx - r * out);
        int b = (int)(r / T), t = (int)(r - (size_t)b * T);
        const float* y = y_tm + ((size_t)t * B + b) * out;
        float v;
        if (mode == CVB_HEAD_SCALE_OUT) {
            v = sW[out * out + i];
            for (int j = 0; j < out; ++j) v = fmaf(sW[i * out + j], y[j], v);
        } else {
            v = y[i];
            if (mode == CVB_HEAD_CLAMP && i >= lat) v = fmaxf(v, LOG_VAR_FLOOR);
        }
        trj_bm[idx] = v;
    }
}

// d_trj (bm) -> dy (tm slots 1..T of dy_tot; slot 0 zeroed; d_y_last added to slot T)
__global__ void k_head_bwd(int B, int T, int out, int mode, int lat, const float* __restrict__ d_trj_bm,
                           const float* __restrict__ y_tm, const float* __restrict__ Ws,
                           const float* __restrict__ d_y_last, float* __restrict__ dy_tot, float* __restrict__ dtrj_tm) {
    extern __shared__ float sW[];
    if (mode == CVB_HEAD_SCALE_OUT) {
        for (int i = threadIdx.x; i < out * out; i += blockDim.x) sW[i] = Ws[i];
        __syncthreads();
    }
    size_t n = (size_t)(T + 1) * B * out;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx / out;
        int j = (int)(idx - r * out);
        int slot = (int)(r / B), b = (int)(r - (size_t)slot * B);
        float v = 0.f;
        if (slot > 0) {
            int t = slot - 1;
            const float* g = d_trj_bm + ((size_t)b * T + t) * out;
            if (mode == CVB_HEAD_SCALE_OUT) {
                for (int i = 0; i < out; ++i) v = fmaf(g[i], sW[i * out + j], v);
            } else {
                v = g[j];
                if (mode == CVB_HEAD_CLAMP && j >= lat && !(y_tm[((size_t)t * B + b) * out + j] >= LOG_VAR_FLOOR)) v = 0.f;
            }
            if (dtrj_tm) dtrj_tm[((size_t)t * B + b) * out + j] = g[j];
            if (slot == T && d_y_last) v += d_y_last[(size_t)b * out + j];
        }
        dy_tot[idx] = v;
    }
}

static inline int grid1d(size_t n) {
    size_t g = ceil_div_sz(n, 256);
    return (int)(g > 148 * 8 ? 148 * 8 : (g < 1 ? 1 : g));
}

static int check_net(const cvb_net* net) {
    CVB_REQUIRE(net, "net is NULL");
    CVB_REQUIRE(net->in_dim > 0 && net->out_dim > 0 && net->hidden > 0, "bad dims in=%d out=%d hidden=%d", net->in_dim,
                net->out_dim, net->hidden);
    CVB_REQUIRE(net->kernel_size % 2 == 1 && net->kernel_size >= 1, "kernel_size must be odd");
    CVB_REQUIRE(net->n_conv >= 1 && net->n_conv <= 4, "dilation_size (conv layers) must be in 1..4");
    CVB_REQUIRE(net->w_ih && net->w_hh && net->b_ih && net->b_hh && net->out_w && net->out_b, "GRU/out_1 parameter is NULL");
    CVB_REQUIRE((reinterpret_cast<uintptr_t>(net->w_hh) & 15u) == 0, "w_hh must be 16-byte aligned (the kernels read it as float4)");
    for (int i = 0; i < net->n_conv; ++i) CVB_REQUIRE(net->conv_w[i] && net->conv_b[i], "conv parameter %d is NULL", i);
    if (net->has_scale_in) CVB_REQUIRE(net->scale_in_w && net->scale_in_b, "scale_in parameter is NULL");
    if (net->has_scale_out) CVB_REQUIRE(net->scale_out_w && net->scale_out_b, "scale_out parameter is NULL");
    return 0;
}

}  // namespace cvb

using namespace cvb;

extern "C" {

size_t cvb_recurrent_ws_floats(const cvb_net* net, int B, int T, int training, int has_mask) {
    return rec_layout(net, B, T, training != 0, has_mask != 0).total;
}
size_t cvb_scratch_floats(const cvb_net* net, int B, int T, int training) {
    size_t f = fwd_scratch(net, B, T).total;
    if (training) {
        size_t b = bwd_scratch(net, B, T).total;
        if (b > f) f = b;
    }
    return f;
}

// which recurrence kernel the last forward / backward call of this process launched (tests: no silent fallback)
static int g_last_path[2] = {-1, -1};
int cvb_last_recurrence_path(int backward) { return g_last_path[backward ? 1 : 0]; }
int cvb_last_recurrence_hops(int backward) {
    const int p = g_last_path[backward ? 1 : 0];
    return p == CVB_PATH_TC_FOLDED ? 1 : p == CVB_PATH_TC ? g_tc_hops[backward ? 1 : 0] : 0;
}

int cvb_recurrence_max_rows(const cvb_net* net, int mode) {
    // largest batch-row count (multiple of 8, <= 128) one launch of the tensor-core recurrence kernels holds at this
    // network shape; 128 when only the fp32-FMA kernels apply (they tile the batch themselves).
    // mode 0: inference without recurrent dropout (folded kernel), 1: forward + BPTT, 2: forward only, with dropout masks
    if (!net || !want_tc()) return 128;
    DeviceInfo di;
    if (get_device_info(&di)) return 128;
    static int cache_key[64], cache_val[64], n_cache = 0;
    const bool fold = mode == 0 && want_fold();
    const int key = (net->hidden * 131 + net->out_dim) * 4 + (fold ? 3 : mode == 1 ? 1 : 0);
    for (int i = 0; i < n_cache; ++i)
        if (cache_key[i] == key) return cache_val[i];
    int best = 128;
    for (int B = 128; B >= 8; B -= 8) {
        const bool ok = fold ? gru_tc_eval_supported(B, net->hidden, net->out_dim, di)
                             : gru_tc_supported(B, net->hidden, net->out_dim, di) &&
                                   (mode != 1 || gru_tc_bwd_supported(B, net->hidden, net->out_dim, di));
        if (ok) {
            best = B;
            break;
        }
        if (B == 8) best = 128;
    }
    if (n_cache < 64) {
        cache_key[n_cache] = key;
        cache_val[n_cache++] = best;
    }
    return best;
}

int cvb_gru_rnn_forward(const cvb_net* net, int B, int T, const float* x_bm, const float* y_in, const float* h_in,
                        const float* mask_conv_tm, const float* mask_gru_tm, int head_mode, int lat_dim, int training,
                        float* trj_out_bm, float* y_last, float* h_last, float* fe_ws, float* rec_ws, float* scratch,
                        void* stream) {
    if (int rc = check_net(net)) return rc;
    CVB_REQUIRE(B > 0 && T > 0, "empty batch (B=%d, T=%d)", B, T);
    CVB_REQUIRE(x_bm && y_in && trj_out_bm && fe_ws && rec_ws && scratch, "required pointer is NULL");
    CVB_REQUIRE(head_mode != CVB_HEAD_SCALE_OUT || net->has_scale_out, "scale_out head without scale_out parameters");
    CVB_REQUIRE(head_mode != CVB_HEAD_CLAMP || (lat_dim > 0 && lat_dim <= net->out_dim), "lat_dim=%d out of range", lat_dim);
    cudaStream_t s = (cudaStream_t)stream;
    const int H = net->hidden, out = net->out_dim, C = conv_dim(net), TI = tot_in_dim(net);
    const size_t TB = (size_t)T * B;
    RecLayout RL = rec_layout(net, B, T, training != 0, mask_gru_tm != nullptr);
    FwdScratch FS = fwd_scratch(net, B, T);
    float* xc = fe_ws + frontend_xc_offset(net, B, T);
    prof_begin(s, CVB_PROF_FRONTEND);
    if (int rc = frontend_fwd(net, B, T, x_bm, mask_conv_tm, fe_ws, xc, s)) return rc;
    prof_end(s, CVB_PROF_FRONTEND);
    float* gx = scratch + FS.gx;
    DeviceInfo di;
    if (int rc = get_device_info(&di)) return rc;
    GruFwdArgs a;
    a.H = H;
    a.out = out;
    a.Wy = net->w_ih + C;
    a.ldwy = TI;
    a.Wo = net->out_w;
    a.bo = net->out_b;
    // inference (no dropout): feedback folded into the recurrent matrix, all y_t from one product afterwards
    const bool fold = want_tc() && !training && !mask_gru_tm && want_fold() && gru_tc_eval_supported(B, H, out, di);
    float* esc = scratch + FS.tc;
    float* cfb = esc + gru_tc_eval_scratch_floats(B, H);   // c_fb | b_ih + c_fb
    const float* gx_bias = net->b_ih;
    if (fold) {
        CVB_REQUIRE(gru_tc_eval_scratch_floats(B, H) + r4((size_t)6 * H) <= FS.tc_floats,
                    "internal: scratch of the folded inference path not reserved (B=%d H=%d out=%d)", B, H, out);
        if (int rc = gru_tc_eval_prepare(a, net->b_ih, esc, cfb, s)) return rc;
        gx_bias = cfb + 3 * H;
    }
    if (want_tc_gemm() && gemm_tc_eligible((int)TB, 3 * H, C)) {
        // gx = xc W_x^T + b_ih: bias in the GEMM epilogue (no fill pass, no read-modify-write of the 3H-wide rows)
        GemmDesc g;
        g.transB = true; g.M = (int)TB; g.N = 3 * H; g.K = C;
        g.A = xc; g.lda = C; g.B = net->w_ih; g.ldb = TI; g.b_const = true; g.bias = gx_bias; g.C = gx; g.ldc = 3 * H;
        if (int rc = gemm_tc_group(s, &g, 1)) return rc;
    } else {
        if (int rc = fill_rows(s, gx, TB, 3 * H, 3 * H, gx_bias)) return rc;
        if (int rc = gemm_rm(s, false, true, (int)TB, 3 * H, C, 1.f, xc, C, net->w_ih, TI, 1.f, gx, 3 * H)) return rc;
    }
    float* hs = rec_ws + RL.hs;
    float* ys = rec_ws + RL.ys;
    if (h_in)
        CVB_CHECK(cudaMemcpyAsync(hs, h_in, (size_t)B * H * sizeof(float), cudaMemcpyDeviceToDevice, s));
    else if (int rc = zero_floats(s, hs, (size_t)B * H))
        return rc;
    CVB_CHECK(cudaMemcpyAsync(ys, y_in, (size_t)B * out * sizeof(float), cudaMemcpyDeviceToDevice, s));
    a.gx = gx;
    a.Whh = net->w_hh;
    a.bhh = net->b_hh;
    a.mask = mask_gru_tm;
    a.hs = hs;
    a.ys = ys;
    a.sv_r = training ? rec_ws + RL.r : nullptr;
    a.sv_z = training ? rec_ws + RL.z : nullptr;
    a.sv_n = training ? rec_ws + RL.n : nullptr;
    a.sv_ghn = training ? rec_ws + RL.ghn : nullptr;
    a.sv_o = (training && mask_gru_tm) ? rec_ws + RL.o : nullptr;
    a.part = scratch + FS.part;
    a.bar = reinterpret_cast<unsigned*>(scratch + FS.bar);
    a.B = B;
    a.T = T;
    if (fold) {
        if (int rc = gru_ar_fwd_tc_eval(a, esc, cfb, s)) return rc;
        g_last_path[0] = CVB_PATH_TC_FOLDED;
    } else if (want_tc() && gru_tc_supported(B, H, out, di)) {
        CVB_REQUIRE(gru_tc_scratch_floats(B, H) <= FS.tc_floats, "internal: scratch of the tensor-core recurrence not reserved");
        if (int rc = gru_ar_fwd_tc(a, scratch + FS.tc, s)) return rc;
        g_last_path[0] = CVB_PATH_TC;
    } else {
        if (int rc = gru_ar_fwd_exact(a, s)) return rc;
        g_last_path[0] = CVB_PATH_FP32;
    }
    size_t smem = head_mode == CVB_HEAD_SCALE_OUT ? (size_t)(out * out + out) * sizeof(float) : 0;
    CVB_REQUIRE(smem <= 48 * 1024, "scale_out matrix too large (out_dim=%d)", out);
    k_head_fwd<<<grid1d(TB * out), 256, smem, s>>>(B, T, out, head_mode, lat_dim, ys + (size_t)B * out, net->scale_out_w,
                                                   net->scale_out_b, trj_out_bm);
    CVB_LAUNCH_CHECK();
    if (y_last)
        CVB_CHECK(cudaMemcpyAsync(y_last, ys + (size_t)T * B * out, (size_t)B * out * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (h_last)
        CVB_CHECK(cudaMemcpyAsync(h_last, hs + (size_t)T * B * H, (size_t)B * H * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return 0;
}

int cvb_gru_rnn_backward(const cvb_net* net, int B, int T, const float* x_bm, const float* mask_conv_tm,
                         const float* mask_gru_tm, int head_mode, int lat_dim, const float* trj_out_bm,
                         const float* d_trj_out_bm, const float* d_y_last, const float* d_h_last, const float* fe_ws,
                         const float* rec_ws, float* scratch, float* dx_bm, float* dy_in, float* dh_in,
                         const cvb_net_grads* gr, void* stream) {
    (void)trj_out_bm;
    if (int rc = check_net(net)) return rc;
    CVB_REQUIRE(B > 0 && T > 0, "empty batch (B=%d, T=%d)", B, T);
    CVB_REQUIRE(d_trj_out_bm && fe_ws && rec_ws && scratch, "required pointer is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    const int H = net->hidden, out = net->out_dim, C = conv_dim(net), TI = tot_in_dim(net);
    const size_t TB = (size_t)T * B;
    const int iTB = (int)TB;
    RecLayout RL = rec_layout(net, B, T, true, mask_gru_tm != nullptr);
    BwdScratch BS = bwd_scratch(net, B, T);
    const float* hs = rec_ws + RL.hs;
    const float* ys = rec_ws + RL.ys;
    float* dy_tot = scratch + BS.dy_tot;
    float* dgi = scratch + BS.dgi;
    float* dghn = scratch + BS.dghn;
    float* dhc = scratch + BS.dhc;
    const bool want_so = gr && (gr->scale_out_w || gr->scale_out_b) && head_mode == CVB_HEAD_SCALE_OUT;
    float* dtrj_tm = want_so ? scratch + BS.dtrj : nullptr;
    size_t smem = head_mode == CVB_HEAD_SCALE_OUT ? (size_t)out * out * sizeof(float) : 0;
    k_head_bwd<<<grid1d((size_t)(T + 1) * B * out), 256, smem, s>>>(B, T, out, head_mode, lat_dim, d_trj_out_bm,
                                                                    ys + (size_t)B * out, net->scale_out_w, d_y_last, dy_tot,
                                                                    dtrj_tm);
    CVB_LAUNCH_CHECK();
    if (d_h_last)
        CVB_CHECK(cudaMemcpyAsync(dhc, d_h_last, (size_t)B * H * sizeof(float), cudaMemcpyDeviceToDevice, s));
    else if (int rc = zero_floats(s, dhc, (size_t)B * H))
        return rc;
    GruBwdArgs a;
    a.Whh = net->w_hh;
    a.Wy = net->w_ih + C;
    a.ldwy = TI;
    a.Wo = net->out_w;
    a.mask = mask_gru_tm;
    a.hs = hs;
    a.sv_r = rec_ws + RL.r;
    a.sv_z = rec_ws + RL.z;
    a.sv_n = rec_ws + RL.n;
    a.sv_ghn = rec_ws + RL.ghn;
    a.dy_tot = dy_tot;
    a.dgi = dgi;
    a.dghn = dghn;
    a.gxch = scratch + BS.gxch;
    a.dhc = dhc;
    a.part = scratch + BS.part;
    a.bar = reinterpret_cast<unsigned*>(scratch + BS.bar);
    a.B = B;
    a.T = T;
    a.H = H;
    a.out = out;
    bool db_in_kernel = false;
    {
        DeviceInfo di;
        if (int rc = get_device_info(&di)) return rc;
        if (want_tc() && gru_tc_bwd_supported(B, H, out, di)) {
            if (gr) {   // the recurrence sums the GRU bias gradients itself (no colsum passes over dgi / dghn)
                a.dbih = gr->b_ih;
                a.dbhh = gr->b_hh;
                a.db_accumulate = gr->accumulate;
                db_in_kernel = true;
            }
            if (int rc = gru_ar_bwd_tc(a, scratch + BS.tc, s)) return rc;
            g_last_path[1] = CVB_PATH_TC;
        } else {
            if (int rc = gru_ar_bwd_exact(a, s)) return rc;
            g_last_path[1] = CVB_PATH_FP32;
        }
    }
    if (dy_in) CVB_CHECK(cudaMemcpyAsync(dy_in, dy_tot, (size_t)B * out * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (dh_in) CVB_CHECK(cudaMemcpyAsync(dh_in, dhc, (size_t)B * H * sizeof(float), cudaMemcpyDeviceToDevice, s));

    const float beta = (gr && gr->accumulate) ? 1.f : 0.f;
    const bool acc = gr && gr->accumulate;
    const float* xc = fe_ws + frontend_xc_offset(net, B, T);
    bool fe_needed = dx_bm != nullptr;
    if (gr) {
        for (int i = 0; i < net->n_conv; ++i) fe_needed = fe_needed || gr->conv_w[i] || gr->conv_b[i];
        fe_needed = fe_needed || gr->scale_in_w || gr->scale_in_b;
    }
    float* dxc = scratch + BS.dxc;
    const float* o_tm = mask_gru_tm ? rec_ws + RL.o : hs + (size_t)B * H;
    const float* dy1 = dy_tot + (size_t)B * out;
    if (want_tc_gemm() && gemm_tc_eligible(3 * H, H, iTB)) {
        // every product that only needs the recurrence's outputs, in ONE persistent launch (long-K tiles first):
        //   dW_hh = [dgi_r; dgi_z; dan*r]^T hs      dW_ih = dgi^T [xc | y_prev]      dW_o = dy^T o      dxc = dgi W_x
        GemmDesc d[5];
        int n = 0;
        if (gr && gr->w_hh) {   // rows r, z from dgi (the operand image is shared with dW_ih below), rows n from dan*r
            GemmDesc& g = d[n++];
            g.transA = true; g.M = 2 * H; g.N = H; g.K = iTB;
            g.A = dgi; g.lda = 3 * H; g.B = hs; g.ldb = H; g.C = gr->w_hh; g.ldc = H; g.beta1 = acc; g.f16 = false;
            GemmDesc& g2 = d[n++];
            g2.transA = true; g2.M = H; g2.N = H; g2.K = iTB;
            g2.A = dghn; g2.lda = H; g2.B = hs; g2.ldb = H; g2.C = gr->w_hh + (size_t)2 * H * H; g2.ldc = H; g2.beta1 = acc; g2.f16 = false;
        }
        if (gr && gr->w_ih) {
            GemmDesc& g = d[n++];
            g.transA = true; g.M = 3 * H; g.N = TI; g.K = iTB;
            g.A = dgi; g.lda = 3 * H; g.B = xc; g.ldb = C; g.B2 = ys; g.ldb2 = out; g.N1 = C;
            g.C = gr->w_ih; g.ldc = TI; g.beta1 = acc; g.f16 = false;
        }
        if (gr && gr->out_w) {
            GemmDesc& g = d[n++];
            g.transA = true; g.M = out; g.N = H; g.K = iTB;
            g.A = dy1; g.lda = out; g.B = o_tm; g.ldb = H; g.C = gr->out_w; g.ldc = H; g.beta1 = acc; g.f16 = false;
        }
        if (fe_needed) {
            GemmDesc& g = d[n++];
            g.M = iTB; g.N = C; g.K = 3 * H;
            g.A = dgi; g.lda = 3 * H; g.B = net->w_ih; g.ldb = TI; g.b_const = true; g.C = dxc; g.ldc = C; g.f16 = false;
        }
        if (n)
            if (int rc = gemm_tc_group(s, d, n)) return rc;
    } else {
        if (gr && gr->w_hh) {
            if (int rc = gemm_rm(s, true, false, 2 * H, H, iTB, 1.f, dgi, 3 * H, hs, H, beta, gr->w_hh, H, true)) return rc;
            if (int rc = gemm_rm(s, true, false, H, H, iTB, 1.f, dghn, H, hs, H, beta, gr->w_hh + (size_t)2 * H * H, H, true)) return rc;
        }
        if (gr && gr->w_ih) {
            if (int rc = gemm_rm(s, true, false, 3 * H, C, iTB, 1.f, dgi, 3 * H, xc, C, beta, gr->w_ih, TI, true)) return rc;
            if (int rc = gemm_rm(s, true, false, 3 * H, out, iTB, 1.f, dgi, 3 * H, ys, out, beta, gr->w_ih + C, TI, true)) return rc;
        }
        if (gr && gr->out_w)
            if (int rc = gemm_rm(s, true, false, out, H, iTB, 1.f, dy1, out, o_tm, H, beta, gr->out_w, H, true)) return rc;
        if (fe_needed)
            if (int rc = gemm_rm(s, false, false, iTB, C, 3 * H, 1.f, dgi, 3 * H, net->w_ih, TI, 0.f, dxc, C, true)) return rc;
    }
    if (gr) {
        if (gr->b_hh && !db_in_kernel) {
            if (int rc = colsum(s, dgi, iTB, 2 * H, 3 * H, gr->b_hh, acc)) return rc;
            if (int rc = colsum(s, dghn, iTB, H, H, gr->b_hh + 2 * H, acc)) return rc;
        }
        if (gr->b_ih && !db_in_kernel)
            if (int rc = colsum(s, dgi, iTB, 3 * H, 3 * H, gr->b_ih, acc)) return rc;
        if (gr->out_b)
            if (int rc = colsum(s, dy1, iTB, out, out, gr->out_b, acc)) return rc;
        if (want_so) {
            if (gr->scale_out_w)
                if (int rc = gemm_rm(s, true, false, out, out, iTB, 1.f, dtrj_tm, out, ys + (size_t)B * out, out, beta,
                                     gr->scale_out_w, out, true))
                    return rc;
            if (gr->scale_out_b)
                if (int rc = colsum(s, dtrj_tm, iTB, out, out, gr->scale_out_b, acc)) return rc;
        }
    }
    if (fe_needed)
        if (int rc = frontend_bwd(net, B, T, x_bm, mask_conv_tm, fe_ws, dxc, scratch + BS.fe, dx_bm, gr, s)) return rc;
    return 0;
}
}
