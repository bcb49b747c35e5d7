// Argument blocks of the persistent recurrence kernels (gru_ar.cu) shared with api.cu.
#pragma once
#include "common.cuh"

namespace cvb {

struct GruFwdArgs {
    const float* gx;    // [T,B,3H]  = W_x xc' + b_ih
    const float* Whh;   // [3H,H]
    const float* bhh;   // [3H]
    const float* Wy;    // [3H,out] with row stride ldwy (= W_ih[:, C:])
    int ldwy;
    const float* Wo;    // [out,H]
    const float* bo;    // [out]
    const float* mask;  // [T,B,H] or null
    float* hs;          // [T+1,B,H]
    float* ys;          // [T+1,B,out]
    float *sv_r, *sv_z, *sv_n, *sv_ghn, *sv_o;  // [T,B,H] or null
    float* part;        // [G,B,out]
    unsigned* bar;
    int B, T, H, out;
};

struct GruBwdArgs {
    const float* Whh;
    const float* Wy;
    int ldwy;
    const float* Wo;
    const float* mask;
    const float* hs;
    const float *sv_r, *sv_z, *sv_n, *sv_ghn;
    float* dy_tot;  // [T+1,B,out]: in: grad of ys slots from the head; out: total grads (slot 0 = dy_in)
    float* dgi;     // [T,B,3H]
    float* dghn;    // [T,B,H]
    float* gxch;    // [2,B,3H] exchange of dgh_t
    float* dhc;     // [B,H] in: d_h_last (or 0); out: dh_in
    float* part;    // [G,B,out]
    unsigned* bar;
    int B, T, H, out;
    // tensor-core kernel only: bias gradients summed inside the recurrence (db_ih = sum dgi, db_hh = sum dgh); null = not wanted
    float* dbih = nullptr;
    float* dbhh = nullptr;
    int db_accumulate = 0;
};

int gru_exact_grid(int H);
int gru_ar_fwd_exact(GruFwdArgs& a, cudaStream_t s);
int gru_ar_bwd_exact(GruBwdArgs& a, cudaStream_t s);

// tensor-core (tcgen05) variant, gru_tc.cu
bool gru_tc_shape_ok(int B, int H, int out);               // host-only shape test (no device query)
bool gru_tc_supported(int B, int H, int out, const DeviceInfo& di);
size_t gru_tc_scratch_floats(int B, int H);
int gru_ar_fwd_tc(GruFwdArgs& f, float* tc_scratch, cudaStream_t s);

// one-exchange-per-step variant of the training forward kernel (cluster partials of y summed by every consumer), gru_tc2.cu
bool gru_tc2_supported(int B, int H, int out, const DeviceInfo& di);
size_t gru_tc2_scratch_floats(int B, int H);
int gru_ar_fwd_tc2(GruFwdArgs& a, float* tc_scratch, cudaStream_t s);
bool gru_tc2_bwd_supported(int B, int H, int out, const DeviceInfo& di);
size_t gru_tc2_bwd_scratch_floats(int B, int H);
int gru_ar_bwd_tc2(GruBwdArgs& a, float* tc_scratch, cudaStream_t s);
extern int g_tc_hops[2];   // grid-wide exchanges per step of the last tensor-core launch: [0] forward, [1] backward
bool gru_tc_one_hop();   // CVB_TC_FEEDBACK=grid keeps the two-exchange training kernels (A/B); default: the one-exchange kernels

// inference-only forward with the y feedback folded into the recurrent matrix (one exchange per step), gru_tc_eval.cu
bool gru_tc_eval_shape_ok(int B, int H);                    // host-only shape test (any out_dim)
bool gru_tc_eval_supported(int B, int H, int out, const DeviceInfo& di);
size_t gru_tc_eval_scratch_floats(int B, int H);
// prepare (weights only; cfb = 6H floats: c_fb | b_ih + c_fb)  ->  caller's gx product with bias cfb + 3H  ->  launch
int gru_tc_eval_prepare(const GruFwdArgs& f, const float* bih, float* scratch, float* cfb, cudaStream_t s);
int gru_ar_fwd_tc_eval(GruFwdArgs& f, float* scratch, const float* cfb, cudaStream_t s);

// tensor-core BPTT over thread-block clusters, gru_tc_bwd.cu
bool gru_tc_bwd_shape_ok(int B, int H, int out);
bool gru_tc_bwd_supported(int B, int H, int out, const DeviceInfo& di);
size_t gru_tc_bwd_scratch_floats(int B, int H);
int gru_ar_bwd_tc(GruBwdArgs& b, float* tc_scratch, cudaStream_t s);

}  // namespace cvb
