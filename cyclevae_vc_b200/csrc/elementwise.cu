// Streaming kernels of the GRU-VAE path: everything here is HBM-bound (one read of the inputs,
// one write of the outputs), 128-bit vectorised where alignment allows, grid-stride over
// multiples of the SM count.  Reductions are warp-shuffle trees with a fixed order
// (deterministic, no atomics).
#include "common.cuh"

namespace cvb {

static inline int grid_for(size_t n, int block, int max_blocks = 148 * 8) {
    size_t g = ceil_div_sz(n, (size_t)block);
    if (g > (size_t)max_blocks) g = max_blocks;
    if (g < 1) g = 1;
    return (int)g;
}

// ---------------------------------------------------------------------------------------------
__global__ void k_zero(float4* __restrict__ p4, size_t n4, float* __restrict__ tail, int ntail) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += st) p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (blockIdx.x == 0 && threadIdx.x < ntail) tail[threadIdx.x] = 0.f;
}
int zero_floats(cudaStream_t s, float* p, size_t n) {
    if (n == 0) return 0;
    if (((uintptr_t)p & 15) == 0) {
        size_t n4 = n / 4;
        k_zero<<<grid_for(n4 ? n4 : 1, 256), 256, 0, s>>>((float4*)p, n4, p + n4 * 4, (int)(n - n4 * 4));
        CVB_LAUNCH_CHECK();
        return 0;
    }
    CVB_CHECK(cudaMemsetAsync(p, 0, n * sizeof(float), s));
    return 0;
}

// dst[r, c] = bias[c]
__global__ void k_fill_rows(float* __restrict__ dst, size_t rows, int cols, int ld, const float* __restrict__ bias) {
    size_t n = rows * (size_t)cols;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i / cols;
        int c = (int)(i - r * cols);
        dst[r * ld + c] = bias[c];
    }
}
int fill_rows(cudaStream_t s, float* dst, size_t rows, int cols, int ld, const float* bias) {
    if (rows == 0 || cols == 0) return 0;
    k_fill_rows<<<grid_for(rows * cols, 256), 256, 0, s>>>(dst, rows, cols, ld, bias);
    CVB_LAUNCH_CHECK();
    return 0;
}

// out[c] (+)= sum_r A[r, c]; block = 8 columns x 128 row-lanes (32-byte row segments, many blocks even for narrow
// matrices), fixed reduction order: lane-strided partial sums, then a sequential sum over the 128 lanes
__global__ void __launch_bounds__(1024) k_colsum(const float* __restrict__ A, int rows, int cols, int lda,
                                                 float* __restrict__ out, int accumulate) {
    __shared__ float red[128][9];
    const int cx = threadIdx.x & 7, ry = threadIdx.x >> 3;
    const int c = blockIdx.x * 8 + cx;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (c < cols) {
        const float* p = A + c;
        int r = ry;
        for (; r + 384 < rows; r += 512) {
            a0 += p[(size_t)r * lda];
            a1 += p[(size_t)(r + 128) * lda];
            a2 += p[(size_t)(r + 256) * lda];
            a3 += p[(size_t)(r + 384) * lda];
        }
        for (; r < rows; r += 128) a0 += p[(size_t)r * lda];
    }
    red[ry][cx] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (ry < 8) {   // 16 lanes per thread, then 8 partials per column
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += red[ry * 16 + i][cx];
        red[ry * 16][cx] = s;
    }
    __syncthreads();
    if (ry == 0 && c < cols) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i * 16][cx];
        out[c] = accumulate ? out[c] + s : s;
    }
}
int colsum(cudaStream_t s, const float* A, int rows, int cols, int lda, float* out, bool accumulate) {
    if (cols <= 0) return 0;
    k_colsum<<<ceil_div(cols, 8), 1024, 0, s>>>(A, rows, cols, lda, out, accumulate ? 1 : 0);
    CVB_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// dropout mask
// `st` (optional): device-resident generator state {seed, counter, ...}: the draw uses seed = st[0] and counters
// st[1] + offset + i, so a captured CUDA graph draws fresh numbers on every replay (cvb_state_advance moves st[1])
__global__ void k_dropout_mask(size_t n, float p, float keep_scale, uint64_t seed, uint64_t offset, const uint64_t* __restrict__ st,
                               float* __restrict__ out) {
    if (st) {
        seed = st[0];
        offset += st[1];
    }
    Philox ph(seed);
    size_t n4 = (n + 3) / 4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        uint4 r = ph(offset + i, 1u);
        float v[4] = {u32_to_unit(r.x) >= p ? keep_scale : 0.f, u32_to_unit(r.y) >= p ? keep_scale : 0.f,
                      u32_to_unit(r.z) >= p ? keep_scale : 0.f, u32_to_unit(r.w) >= p ? keep_scale : 0.f};
        size_t b = i * 4;
        if (b + 3 < n && (((uintptr_t)(out + b)) & 15) == 0) {
            *reinterpret_cast<float4*>(out + b) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            for (int j = 0; j < 4 && b + j < n; ++j) out[b + j] = v[j];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// reparameterise + concat
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
    float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0,1]
    float u2 = u32_to_unit(b);
    float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    z0 = r * c;
    z1 = r * s;
}

__global__ void k_reparam_concat_fwd(size_t rows, int lat, int n_code, const float* __restrict__ latp,
                                     const float* __restrict__ code, const float* __restrict__ eps,
                                     uint64_t seed, uint64_t offset, const uint64_t* __restrict__ st,
                                     float* __restrict__ eps_out, float* __restrict__ out) {
    int W = n_code + lat;
    size_t n = rows * (size_t)W;
    if (st) {
        seed = st[0];
        offset += st[1];
    }
    Philox ph(seed);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i / W;
        int c = (int)(i - r * W);
        float v;
        if (c < n_code) {
            v = code[r * n_code + c];
        } else {
            int d = c - n_code;
            float mu = latp[r * 2 * lat + d], lv = latp[r * 2 * lat + lat + d];
            float e;
            if (eps) {
                e = eps[r * lat + d];
            } else {
                size_t idx = r * lat + d;
                uint4 q = ph(offset + idx / 4, 2u);
                float z0, z1, z2, z3;
                box_muller(q.x, q.y, z0, z1);
                box_muller(q.z, q.w, z2, z3);
                int w = (int)(idx & 3);
                e = w == 0 ? z0 : (w == 1 ? z1 : (w == 2 ? z2 : z3));
                if (eps_out) eps_out[idx] = e;
            }
            v = mu + expf(lv * 0.5f) * e;
        }
        out[i] = v;
    }
}

__global__ void k_reparam_concat_bwd(size_t rows, int lat, int n_code, const float* __restrict__ latp,
                                     const float* __restrict__ eps, const float* __restrict__ d_out,
                                     float* __restrict__ d_lat) {
    int W = n_code + lat;
    size_t n = rows * (size_t)lat;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i / lat;
        int d = (int)(i - r * lat);
        float g = d_out[r * W + n_code + d];
        float lv = latp[r * 2 * lat + lat + d];
        d_lat[r * 2 * lat + d] = g;
        d_lat[r * 2 * lat + lat + d] = g * eps[r * lat + d] * 0.5f * expf(lv * 0.5f);
    }
}

__global__ void k_concat2(size_t rows, int ca, const float* __restrict__ a, int lda, int cb,
                          const float* __restrict__ b, int ldb, float* __restrict__ out) {
    int W = ca + cb;
    size_t n = rows * (size_t)W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i / W;
        int c = (int)(i - r * W);
        out[i] = c < ca ? a[r * lda + c] : b[r * ldb + (c - ca)];
    }
}

// ---------------------------------------------------------------------------------------------
// losses: one CTA per utterance, a warp per frame, shuffle reduction over the feature axis,
// fixed-order block reduction (double for the frame sums so that mean/std match torch's).
__global__ void __launch_bounds__(256) k_kl_fwd(int T, int lat, const float* __restrict__ latp,
                                                const int32_t* __restrict__ flens, float* __restrict__ kl) {
    int j = blockIdx.x;
    int F = min(max(flens[j], 0), T);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* base = latp + (size_t)j * T * 2 * lat;
    double acc = 0.0;
    for (int t = warp; t < F; t += 8) {
        float s = 0.f;
        for (int d = lane; d < lat; d += 32) {
            float mu = base[(size_t)t * 2 * lat + d], lv = base[(size_t)t * 2 * lat + lat + d];
            s += expf(lv) + mu * mu - lv - 1.0f;
        }
        s = warp_sum(s);
        acc += (double)(0.5f * s);
    }
    __shared__ double red[8];
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += red[i];
        kl[j] = F > 0 ? (float)(s / (double)F) : 0.f;
    }
}

__global__ void k_kl_bwd(int B, int T, int lat, const float* __restrict__ latp, const int32_t* __restrict__ flens,
                         const float* __restrict__ d_kl, float* __restrict__ d_lat) {
    size_t n = (size_t)B * T * lat;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i / lat;
        int d = (int)(i - r * lat);
        int j = (int)(r / T), t = (int)(r - (size_t)j * T);
        int F = min(max(flens[j], 0), T);
        float gm = 0.f, gs = 0.f;
        if (t < F) {
            float g = d_kl[j] / (float)F;
            float mu = latp[r * 2 * lat + d], lv = latp[r * 2 * lat + lat + d];
            gm = g * mu;
            gs = g * 0.5f * (expf(lv) - 1.0f);
        }
        d_lat[r * 2 * lat + d] = gm;
        d_lat[r * 2 * lat + lat + d] = gs;
    }
}

#define CVB_MCD_COEF 6.1418514175153268f  // (10/ln10)*sqrt(2), gru_vae.py:525

__global__ void __launch_bounds__(256) k_mcd_fwd(int T, int D, const float* __restrict__ x, int ldx, int x_off,
                                                 const float* __restrict__ y, int ldy, int y_off,
                                                 const int32_t* __restrict__ flens, float* __restrict__ out3) {
    int j = blockIdx.x;
    int F = min(max(flens[j], 0), T);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* xb = x + (size_t)j * T * ldx + x_off;
    const float* yb = y + (size_t)j * T * ldy + y_off;
    double s1 = 0.0, s2 = 0.0;
    for (int t = warp; t < F; t += 8) {
        float s = 0.f;
        for (int d = lane; d < D; d += 32) s += fabsf(xb[(size_t)t * ldx + d] - yb[(size_t)t * ldy + d]);
        s = warp_sum(s);
        float m = CVB_MCD_COEF * s;
        s1 += (double)m;
        s2 += (double)m * (double)m;
    }
    __shared__ double r1[8], r2[8];
    if (lane == 0) {
        r1[warp] = s1;
        r2[warp] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < 8; ++i) {
            a += r1[i];
            b += r2[i];
        }
        double mean = F > 0 ? a / F : 0.0;
        double var = F > 1 ? (b - a * mean) / (double)(F - 1) : 0.0;
        out3[j * 3 + 0] = (float)a;
        out3[j * 3 + 1] = (float)mean;
        out3[j * 3 + 2] = F > 1 ? (float)sqrt(var > 0.0 ? var : 0.0) : nanf("");
    }
}

__global__ void k_mcd_bwd(int B, int T, int D, const float* __restrict__ x, int ldx, int x_off,
                          const float* __restrict__ y, int ldy, int y_off, const int32_t* __restrict__ flens,
                          const float* __restrict__ d_sum, const float* __restrict__ d_mean, float* __restrict__ dx) {
    size_t n = (size_t)B * T * D;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i / D;
        int d = (int)(i - r * D);
        int j = (int)(r / T), t = (int)(r - (size_t)j * T);
        int F = min(max(flens[j], 0), T);
        float g = 0.f;
        if (t < F) {
            float w = (d_sum ? d_sum[j] : 0.f) + (d_mean ? d_mean[j] / (float)F : 0.f);
            float df = x[r * ldx + x_off + d] - y[r * ldy + y_off + d];
            g = CVB_MCD_COEF * w * (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f));
        }
        dx[i] = g;
    }
}

// ---------------------------------------------------------------------------------------------
// `st` (optional): the step count lives on the device (st[2] = steps taken so far), bias corrections computed here
__global__ void k_adam(size_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt,
                       float gscale, const uint64_t* __restrict__ st) {
    if (st) {
        const float t = (float)(st[2] + 1);
        bc1 = 1.0f - powf(b1, t);
        bc2_sqrt = sqrtf(1.0f - powf(b2, t));
    }
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float gi = g[i] * gscale;
        float mi = b1 * m[i] + (1.f - b1) * gi;
        float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] -= (lr / bc1) * (mi / denom);
    }
}

__global__ void k_state_advance(uint64_t* st, uint64_t rng_delta, uint64_t step_delta) {
    st[1] += rng_delta;
    st[2] += step_delta;
}

}  // namespace cvb

using namespace cvb;

extern "C" {

int cvb_dropout_mask(size_t n, float p, uint64_t seed, uint64_t offset, const uint64_t* dev_state, float* out, void* stream) {
    if (n == 0) return 0;
    CVB_REQUIRE(p >= 0.f && p < 1.f, "dropout p=%f out of [0,1)", p);
    k_dropout_mask<<<grid_for((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(n, p, 1.0f / (1.0f - p), seed, offset, dev_state, out);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_state_advance(uint64_t* dev_state, uint64_t rng_delta, uint64_t step_delta, void* stream) {
    CVB_REQUIRE(dev_state, "dev_state is NULL");
    k_state_advance<<<1, 1, 0, (cudaStream_t)stream>>>(dev_state, rng_delta, step_delta);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_reparam_concat_fwd(int B, int T, int lat, int n_code, const float* lat_bm, const float* code_bm,
                           const float* eps_bm, uint64_t seed, uint64_t offset, const uint64_t* dev_state, float* eps_out, float* out_bm,
                           void* stream) {
    CVB_REQUIRE(B >= 0 && T >= 0 && lat > 0 && n_code >= 0, "bad dims");
    size_t rows = (size_t)B * T;
    if (rows == 0) return 0;
    CVB_REQUIRE(n_code == 0 || code_bm, "code_bm is NULL with n_code=%d", n_code);
    k_reparam_concat_fwd<<<grid_for(rows * (n_code + lat), 256), 256, 0, (cudaStream_t)stream>>>(
        rows, lat, n_code, lat_bm, code_bm, eps_bm, seed, offset, dev_state, eps_out, out_bm);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_reparam_concat_bwd(int B, int T, int lat, int n_code, const float* lat_bm, const float* eps_bm,
                           const float* d_out_bm, float* d_lat_bm, void* stream) {
    size_t rows = (size_t)B * T;
    if (rows == 0) return 0;
    CVB_REQUIRE(eps_bm, "eps is required for backward");
    k_reparam_concat_bwd<<<grid_for(rows * lat, 256), 256, 0, (cudaStream_t)stream>>>(rows, lat, n_code, lat_bm, eps_bm,
                                                                                    d_out_bm, d_lat_bm);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_concat2_fwd(int rows, int ca, const float* a, int lda, int cb, const float* b, int ldb, float* out,
                    void* stream) {
    if (rows <= 0 || ca + cb <= 0) return 0;
    k_concat2<<<grid_for((size_t)rows * (ca + cb), 256), 256, 0, (cudaStream_t)stream>>>((size_t)rows, ca, a, lda, cb, b, ldb, out);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_kl_fwd(int B, int T, int lat, const float* lat_bm, const int32_t* flens, float* kl, void* stream) {
    if (B <= 0) return 0;
    k_kl_fwd<<<B, 256, 0, (cudaStream_t)stream>>>(T, lat, lat_bm, flens, kl);
    CVB_LAUNCH_CHECK();
    return 0;
}
int cvb_kl_bwd(int B, int T, int lat, const float* lat_bm, const int32_t* flens, const float* d_kl, float* d_lat_bm,
               void* stream) {
    if ((size_t)B * T == 0) return 0;
    k_kl_bwd<<<grid_for((size_t)B * T * lat, 256), 256, 0, (cudaStream_t)stream>>>(B, T, lat, lat_bm, flens, d_kl, d_lat_bm);
    CVB_LAUNCH_CHECK();
    return 0;
}
int cvb_mcd_l1_fwd(int B, int T, int D, const float* x_bm, int ldx, int x_off, const float* y_bm, int ldy, int y_off,
                   const int32_t* flens, float* out3, void* stream) {
    if (B <= 0) return 0;
    k_mcd_fwd<<<B, 256, 0, (cudaStream_t)stream>>>(T, D, x_bm, ldx, x_off, y_bm, ldy, y_off, flens, out3);
    CVB_LAUNCH_CHECK();
    return 0;
}
int cvb_mcd_l1_bwd(int B, int T, int D, const float* x_bm, int ldx, int x_off, const float* y_bm, int ldy, int y_off,
                   const int32_t* flens, const float* d_sum, const float* d_mean, float* dx_bm, void* stream) {
    if ((size_t)B * T * D == 0) return 0;
    k_mcd_bwd<<<grid_for((size_t)B * T * D, 256), 256, 0, (cudaStream_t)stream>>>(B, T, D, x_bm, ldx, x_off, y_bm, ldy, y_off,
                                                                              flens, d_sum, d_mean, dx_bm);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_adam_step(size_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float lr, float beta1,
                  float beta2, float eps, int step, const uint64_t* dev_state, float grad_scale, void* stream) {
    if (n == 0) return 0;
    CVB_REQUIRE(step >= 1 || dev_state, "adam step must be >= 1");
    float bc1 = 1.0f - powf(beta1, (float)step);
    float bc2 = 1.0f - powf(beta2, (float)step);
    k_adam<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, bc1,
                                                               sqrtf(bc2), grad_scale, dev_state);
    CVB_LAUNCH_CHECK();
    cvb::weights_changed();   // cached 16-bit images of parameter operands (gemm_tc.cu) are stale from here on
    return 0;
}

int cvb_weights_changed(void) {
    cvb::weights_changed();
    return 0;
}
int cvb_reserve_workspace(size_t bytes) { return cvb::reserve_workspace(bytes); }
}
