/*
 * cyclevae_b200.h -- C ABI of libcyclevae_b200.so (sm_100a CUDA).
 *
 * The reference (patrickltobing/cyclevae-vc) has no FFI: its boundary for this path is the Python
 * surface of src/nets/gru_vae.py.  These entry points are what a ctypes binding of that module
 * calls (cyclevae_vc_b200/gru_vae.py is that binding; INTEGRATION.md shows the stub).  Each entry
 * point names the reference lines it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 unless stated otherwise; the library
 *     allocates nothing the caller sees: outputs and workspaces are caller-allocated (torch.empty);
 *   - `stream` is a cudaStream_t passed as void*; no hidden synchronisation;
 *   - return value 0 = ok, non-zero = error (message via cvb_last_error()); no exceptions cross;
 *   - "bm" = batch-major [B,T,C] (the reference's tensor layout); "tm" = time-major [T,B,C]
 *     (the internal layout of the recurrence: one contiguous [B,C] block per frame).
 *   - no CPU fallback exists: every call needs a CUDA device.
 */
#ifndef CYCLEVAE_B200_H
#define CYCLEVAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVB_ABI_VERSION 13

/* Shape + parameter pointers of one GRU_RNN instance (gru_vae.py:282-320).  Parameter pointers
 * are the data_ptr()s of the module's own nn.Parameters (same names as the reference's
 * state_dict), so in-place optimizer updates are seen by the next call with no staging step. */
typedef struct cvb_net {
    int32_t in_dim;        /* gru_vae.py:284 */
    int32_t out_dim;       /* gru_vae.py:285 */
    int32_t hidden;        /* hidden_units, gru_vae.py:286 */
    int32_t kernel_size;   /* gru_vae.py:288 */
    int32_t n_conv;        /* "dilation_size" = number of conv layers, gru_vae.py:300 */
    int32_t has_scale_in;  /* gru_vae.py:295-297 */
    int32_t has_scale_out; /* gru_vae.py:316-318 */
    int32_t reserved;
    const float* scale_in_w;  /* [in,in,1]  */
    const float* scale_in_b;  /* [in]       */
    const float* conv_w[4];   /* layer i: [in*k^(i+1), in*k^i, k]  (gru_vae.py:47-51) */
    const float* conv_b[4];   /* layer i: [in*k^(i+1)] */
    const float* w_ih;        /* gru.weight_ih_l0 [3H, in*k^n + out], gate rows r,z,n */
    const float* w_hh;        /* gru.weight_hh_l0 [3H, H]; 16-byte aligned */
    const float* b_ih;        /* [3H] */
    const float* b_hh;        /* [3H] */
    const float* out_w;       /* out_1.weight [out,H,1] */
    const float* out_b;       /* [out] */
    const float* scale_out_w; /* [out,out,1] */
    const float* scale_out_b; /* [out] */
} cvb_net;

/* Gradient destinations, same shapes as the parameters; any pointer may be NULL (= not wanted).
 * Values are OVERWRITTEN when accumulate == 0 and ADDED TO when accumulate != 0. */
typedef struct cvb_net_grads {
    float* scale_in_w;
    float* scale_in_b;
    float* conv_w[4];
    float* conv_b[4];
    float* w_ih;
    float* w_hh;
    float* b_ih;
    float* b_hh;
    float* out_w;
    float* out_b;
    float* scale_out_w;
    float* scale_out_b;
    int32_t accumulate;
    int32_t reserved;
} cvb_net_grads;

/* head modes for cvb_gru_rnn_forward (gru_vae.py:402-412) */
#define CVB_HEAD_NONE 0      /* trj_out = y                         */
#define CVB_HEAD_CLAMP 1     /* encoder, clamp_vae=True (:408-412)  */
#define CVB_HEAD_SCALE_OUT 2 /* decoder, scale_out conv (:402-406)  */

const char* cvb_last_error(void);
int cvb_abi_version(void);
/* number of SMs / max dynamic shared memory of the current device; <0 on error */
int cvb_device_info(int* n_sm, int* max_smem_optin, int* cc_major, int* cc_minor);

/* ---- workspace sizing (host-only arithmetic; callable without a GPU) ------------------------ */
/* floats of the front-end's padded-grid buffers (all layers) + repacked conv weights */
size_t cvb_frontend_ws_floats(const cvb_net* net, int B, int T);
/* floats of the recurrence state kept for backward: hs,ys (+ r,z,n,ghn,o when training) */
size_t cvb_recurrent_ws_floats(const cvb_net* net, int B, int T, int training, int has_mask);
/* floats of scratch not needed after the call (gx, y-partials, barrier words) */
size_t cvb_scratch_floats(const cvb_net* net, int B, int T, int training);

/* Largest batch-row count (<= 128) one launch of the persistent recurrence kernels should be given at this network
 * shape (the tensor-core kernels keep all rows of a step in one MMA tile and their shared-memory ring scales with the
 * rows).  Rows never interact inside GRU_RNN.forward (gru_vae.py:364-399), so the host runs wider batches as
 * independent slices.  mode 0: inference without recurrent dropout (the folded one-exchange kernel), 1: forward + BPTT,
 * 2: forward only, with dropout masks. */
int cvb_recurrence_max_rows(const cvb_net* net, int mode);

/* Which recurrence kernel the last cvb_gru_rnn_forward (backward = 0) / cvb_gru_rnn_backward (backward = 1) call of this
 * process launched; -1 before the first call.  For tests and monitoring: a shape that silently leaves the tensor-core
 * path shows up here. */
#define CVB_PATH_FP32 0      /* fp32-FMA persistent kernels (gru_ar.cu) */
#define CVB_PATH_TC 1        /* tcgen05 training kernels (gru_tc.cu / gru_tc_bwd.cu: two grid-wide exchanges per step; gru_tc2.cu / gru_tc2_bwd.cu: one) */
#define CVB_PATH_TC_FOLDED 2 /* tcgen05 inference kernel, feedback folded, one exchange per step (gru_tc_eval.cu) */
int cvb_last_recurrence_path(int backward);
/* Grid-wide exchanges per recurrent step of the tensor-core kernel that call launched: 1 (gru_tc2.cu, gru_tc2_bwd.cu,
 * gru_tc_eval.cu), 2 (gru_tc.cu, gru_tc_bwd.cu); 0 for the fp32-FMA kernels or before the first call. */
int cvb_last_recurrence_hops(int backward);

/* ---- GRU_RNN.forward (gru_vae.py:322-455; kwargs do / clamp_vae / lat_dim / h_in) ------------
 * x_bm [B,T,in]; y_in [B,out] (the reference's [B,1,out]); h_in [B,H] or NULL (= zeros);
 * mask_conv_tm [T,B,in*k^n] / mask_gru_tm [T,B,H]: dropout masks already scaled by 1/(1-p), or
 * NULL (eval / do=False) -- the replacement of nn.Dropout at :355/:369/:380.
 * Outputs: trj_out_bm [B,T,out], y_last [B,out] (pre-head), h_last [B,H].
 * fe_ws / rec_ws are kept by the caller for cvb_gru_rnn_backward when training != 0. */
int cvb_gru_rnn_forward(const cvb_net* net, int B, int T, const float* x_bm, const float* y_in,
                        const float* h_in, const float* mask_conv_tm, const float* mask_gru_tm,
                        int head_mode, int lat_dim, int training, float* trj_out_bm, float* y_last,
                        float* h_last, float* fe_ws, float* rec_ws, float* scratch, void* stream);

/* BPTT of the above (autograd of gru_vae.py:322-455).
 * d_trj_out_bm [B,T,out], d_y_last [B,out] or NULL, d_h_last [B,H] or NULL.
 * Outputs (any may be NULL): dx_bm [B,T,in], dy_in [B,out], dh_in [B,H]; parameter grads -> grads. */
int cvb_gru_rnn_backward(const cvb_net* net, int B, int T, const float* x_bm,
                         const float* mask_conv_tm, const float* mask_gru_tm, int head_mode,
                         int lat_dim, const float* trj_out_bm, const float* d_trj_out_bm,
                         const float* d_y_last, const float* d_h_last, const float* fe_ws,
                         const float* rec_ws, float* scratch, float* dx_bm, float* dy_in,
                         float* dh_in, const cvb_net_grads* grads, void* stream);

/* ---- pieces, exported for unit tests and for callers that fuse differently ------------------ */
/* scale_in + TwoSidedDilConv1d (+mask) -> xc_tm [T,B,in*k^n]   (gru_vae.py:336,53-66,355) */
int cvb_frontend_fwd(const cvb_net* net, int B, int T, const float* x_bm, const float* mask_conv_tm,
                     float* fe_ws, float* xc_tm, void* stream);

/* backward of cvb_frontend_fwd (TwoSidedDilConv1d.forward stand-alone, gru_vae.py:53-66, and the front-end half of
 * GRU_RNN's BPTT): dxc_tm [T,B,in*k^n] = gradient of the masked conv output -> dx_bm [B,T,in] (may be NULL) and the conv /
 * scale_in gradients named in `grads` (NULL members are skipped; grads may be NULL).  x_bm / mask_conv_tm / fe_ws as
 * given to / left by cvb_frontend_fwd; scratch: cvb_frontend_bwd_ws_floats(net, B, T) floats. */
size_t cvb_frontend_bwd_ws_floats(const cvb_net* net, int B, int T);
int cvb_frontend_bwd(const cvb_net* net, int B, int T, const float* x_bm, const float* mask_conv_tm, const float* fe_ws,
                     const float* dxc_tm, float* scratch, float* dx_bm, const cvb_net_grads* grads, void* stream);

/* ---- sampling_vae_batch + torch.cat((code, z), 2) fused (gru_vae.py:85-98; train_*.py:1302) ---
 * lat_bm [B,T,2*lat] = [mu | log-var]; eps_bm [B,T,lat] or NULL (then N(0,1) is drawn in-kernel
 * from Philox4x32-10 keyed by (seed, offset) and, if eps_out != NULL, stored for backward);
 * code_bm [B,T,n_code] or NULL with n_code == 0.  out_bm [B,T,n_code+lat] = [code | mu+exp(s/2)eps]
 * dev_state: optional device-resident generator state (see cvb_state_advance); NULL = use (seed, offset) as given */
int cvb_reparam_concat_fwd(int B, int T, int lat, int n_code, const float* lat_bm,
                           const float* code_bm, const float* eps_bm, uint64_t seed, uint64_t offset,
                           const uint64_t* dev_state, float* eps_out, float* out_bm, void* stream);
/* d_out_bm [B,T,n_code+lat] -> d_lat_bm [B,T,2*lat] (code gets no gradient) */
int cvb_reparam_concat_bwd(int B, int T, int lat, int n_code, const float* lat_bm,
                           const float* eps_bm, const float* d_out_bm, float* d_lat_bm, void* stream);
/* generic feature concat [B,T,ca] ++ [B,T,cb] (train_*.py:1304,1307) and its split backward */
int cvb_concat2_fwd(int rows, int ca, const float* a, int lda, int cb, const float* b, int ldb,
                    float* out, void* stream);

/* ---- the data-parallel collective: utterance shards, one all-reduce of the flat gradient buffer ----
 * NCCL is bound at run time (dlopen of libnccl.so.2).  One communicator per rank = per GPU = per process (or per host
 * thread that has made its device current).  Rank 0 calls cvb_comm_unique_id and hands the 128 bytes to the other ranks
 * by any side channel; every rank then calls cvb_comm_init on its own device.  cvb_allreduce_sum adds `flat` over the
 * ranks IN PLACE on `stream` -- SUM, not mean: the reference sums per-utterance losses (train_*.py:1403,1408), so summed
 * shard gradients equal the gradient of one process holding every shard's utterances. */
int cvb_comm_unique_id(void* id128);
int cvb_comm_init(void** comm, const void* id128, int rank, int world);
int cvb_allreduce_sum(void* comm, float* flat, size_t n_floats, void* stream);
int cvb_comm_destroy(void* comm);

/* ---- losses -----------------------------------------------------------------------------------
 * loss_vae (gru_vae.py:117-123) per utterance: kl[j] = mean_{t<flen[j]} 0.5*sum_d(e^s+mu^2-s-1).
 * lat_bm [B,T,2*lat]; flens int32 [B] (device); utterances with flen<=0 give 0. */
int cvb_kl_fwd(int B, int T, int lat, const float* lat_bm, const int32_t* flens, float* kl,
               void* stream);
int cvb_kl_bwd(int B, int T, int lat, const float* lat_bm, const int32_t* flens, const float* d_kl,
               float* d_lat_bm, void* stream);
/* TWFSEloss.forward(x, y, L2=False, GV=False) (gru_vae.py:521-534) per utterance:
 * per frame mcd_t = (10/ln10)*sqrt(2)*sum_d|x-y|; out3 [B,3] = (sum, mean, unbiased std) over
 * t<flen[j].  x_bm [B,T,ldx] uses columns [x_off, x_off+D); y likewise. */
int cvb_mcd_l1_fwd(int B, int T, int D, const float* x_bm, int ldx, int x_off, const float* y_bm,
                   int ldy, int y_off, const int32_t* flens, float* out3, void* stream);
/* gradient of d_sum[j]*sum + d_mean[j]*mean w.r.t. x (written to dx_bm [B,T,D], zero beyond flen) */
int cvb_mcd_l1_bwd(int B, int T, int D, const float* x_bm, int ldx, int x_off, const float* y_bm,
                   int ldy, int y_off, const int32_t* flens, const float* d_sum,
                   const float* d_mean, float* dx_bm, void* stream);

/* ---- evaluation metrics on the device (the reference calls the third-party dtw_c on the CPU) ------
 * cvb_mcd_aligned: mean and population std of the frame distance MCD(x_t, y_t) = (10/ln10) sqrt(2 sum_d (x-y)^2) [dB] over
 * n aligned frames (replaces dtw.calc_mcd at train_*.py:1435-1439).  idx_x / idx_y: optional int64 frame indices
 * (the speech-frame gathers of the call sites), NULL = frames 0..n-1.  out2 = {mean, std}.
 * cvb_dtw_mcd: dynamic time warping of org [N,D] onto trg [M,D] (replaces dtw.dtw_org_to_trg at train_*.py:679-688) with
 * the MCD frame distance and the symmetric step pattern {(1,1),(1,0),(0,1)}, ties resolved in that order.
 * path [M] (int32): the last org frame aligned with each target frame; out3 = {mean over target frames of
 * MCD(org[path[j]], trg[j]), number of path steps, accumulated cost}.  ws: cvb_dtw_ws_bytes(N, M) bytes of scratch. */
size_t cvb_dtw_ws_bytes(int N, int M);
int cvb_dtw_mcd(int N, int M, int D, const float* org, int ldo, const float* trg, int ldt, void* ws,
                int32_t* path, float* out3, void* stream);
int cvb_mcd_aligned(int n, int D, const float* x, int ldx, const float* y, int ldy, const int64_t* idx_x,
                    const int64_t* idx_y, float* out2, void* stream);

/* ---- dropout masks (replacement of nn.Dropout's Bernoulli draw, gru_vae.py:303-304,312-313) ---
 * out[i] = (u_i >= p) ? 1/(1-p) : 0 with u from Philox4x32-10(seed, offset + i/4) */
int cvb_dropout_mask(size_t n, float p, uint64_t seed, uint64_t offset, const uint64_t* dev_state, float* out, void* stream);

/* ---- device-resident step state: uint64 dev_state[4] = {Philox seed, Philox counter, optimiser steps taken, reserved}.
 * When a draw is given dev_state it uses seed = dev_state[0] and counters dev_state[1] + offset + i (offset = the
 * draw's fixed base inside one step); cvb_adam_step with dev_state takes its step count from dev_state[2] + 1.
 * cvb_state_advance adds to the counter and the step count ON THE DEVICE, so one optimisation step captured in a CUDA
 * graph draws fresh dropout masks / noise and applies the right bias correction on every replay. */
int cvb_state_advance(uint64_t* dev_state, uint64_t rng_delta, uint64_t step_delta, void* stream);

/* ---- fused Adam over a flat fp32 buffer (caller = train_*.py:377,1420 torch.optim.Adam) ------ */
int cvb_adam_step(size_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                  float lr, float beta1, float beta2, float eps, int step, const uint64_t* dev_state,
                  float grad_scale, void* stream);

/* ---- parameter residency --------------------------------------------------------------------------
 * The dense products keep 16-bit operand images of PARAMETER matrices (W_x ...) across calls.  They are refreshed
 * after cvb_adam_step automatically; a caller that changes parameters any other way (torch.optim, load_state_dict,
 * .cuda()/.cpu() round trips of save_checkpoint, train_*.py:152-167) calls cvb_weights_changed() before the next
 * forward -- the Python module does so whenever a parameter's (data_ptr, version) stamp moved. */
int cvb_weights_changed(void);
/* Pre-size the library's operand-image arena on the current device (bytes).  The arena otherwise grows on demand with
 * cudaFree/cudaMalloc, which synchronises the device and is illegal while a CUDA graph is being captured. */
int cvb_reserve_workspace(size_t bytes);

/* ---- measurement hooks (bench.py's roofline object) -------------------------------------------
 * When enabled, CUDA events are recorded on the launching stream around every launch of the
 * persistent recurrence kernels (kind 0 = forward, 1 = BPTT), around the front-end chain of every forward call
 * (kind 3) and, at level 2, around every dense product (kind 2).  cvb_profile_summary synchronises those events and returns total ms / launches. */
#define CVB_PROF_GRU_FWD 0
#define CVB_PROF_GRU_BWD 1
#define CVB_PROF_GEMM 2
#define CVB_PROF_FRONTEND 3 /* the front-end chain of one forward call (scale_in + conv stack + dropout -> xc) */
int cvb_profile_enable(int level);
int cvb_profile_reset(void);
int cvb_profile_summary(int kind, float* total_ms, int* launches);
/* number of kernels of THIS library launched since process start (library GEMM calls excluded) */
long long cvb_launch_count(void);

/* plain fp32 GEMM used by the path (row-major; C = alpha*op(A)*op(B) + beta*C); exported so the
 * tests can check it in isolation. */
int cvb_gemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
             const float* Bm, int ldb, float beta, float* C, int ldc, void* stream);

/* the split-precision tcgen05 GEMM behind cvb_gemm's large products, exported for tests: C[M,N] = op(A) op(B)
 * (+ C if beta1) (+ bias[N]); operands are split x = hi + lo into fp16 (f16 != 0, forward products) or bf16
 * (gradient products) pairs, three products with fp32 accumulation in TMEM. */
int cvb_gemm_tc(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* Bm, int ldb, int beta1,
                const float* bias, float* C, int ldc, int f16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CYCLEVAE_B200_H */
