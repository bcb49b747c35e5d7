// On-device evaluation metrics of the training loop: the mel-cepstral distortion of aligned frames
// (train_*.py:1435-1439, `dtw.calc_mcd`) and the DTW alignment of a converted trajectory onto the target utterance
// (train_*.py:679-688, `dtw.dtw_org_to_trg`), which the reference computes on the CPU with the third-party package
// dtw_c after a device -> host copy of every trajectory (a sync per step).  dtw_c is NOT part of the reference tree
// (unpinned in tools/requirements.txt, source absent): these kernels follow the published definitions -- frame
// distance MCD(x, y) = (10 / ln 10) sqrt(2 sum_d (x_d - y_d)^2) [dB]; DTW with the symmetric step pattern
// {(1,1), (1,0), (0,1)}, unit weights, free of windowing -- and are checked against oracle/dtw_oracle.py (numpy).
#include "common.cuh"

namespace cvb {

#define MCD_K 6.1418514f   // (10 / ln 10) * sqrt(2)

// D[i][j] = MCD(org_i, trg_j); 16 x 16 tile per block
__global__ void __launch_bounds__(256) k_dtw_dist(int N, int M, int Dm, const float* __restrict__ org, int ldo, const float* __restrict__ trg,
                                                  int ldt, float* __restrict__ dist) {
    __shared__ float so[16][65], st[16][65];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
    float acc = 0.f;
    for (int d0 = 0; d0 < Dm; d0 += 64) {
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            const int r = e >> 6, c = e & 63;
            so[r][c] = (i0 + r < N && d0 + c < Dm) ? org[(size_t)(i0 + r) * ldo + d0 + c] : 0.f;
            st[r][c] = (j0 + r < M && d0 + c < Dm) ? trg[(size_t)(j0 + r) * ldt + d0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll 16
        for (int c = 0; c < 64; ++c) {
            const float df = so[ty][c] - st[tx][c];
            acc = fmaf(df, df, acc);
        }
        __syncthreads();
    }
    if (i0 + ty < N && j0 + tx < M) dist[(size_t)(i0 + ty) * M + j0 + tx] = MCD_K * sqrtf(acc);
}

// one CTA: accumulated cost over anti-diagonals (three rolling diagonals in shared memory), step directions to global,
// backtracking by one thread, mean distance along the target axis by all
//   G(i,j) = D(i,j) + min(G(i-1,j-1), G(i-1,j), G(i,j-1)), ties broken in that order (diagonal first)
__global__ void __launch_bounds__(1024) k_dtw_dp(int N, int M, const float* __restrict__ dist, unsigned char* __restrict__ dir,
                                                 int* __restrict__ path, float* __restrict__ out) {
    extern __shared__ float sg[];   // [3][N + 1], diagonal k holds G(i, k - i) at index i
    const int L = N + 1;
    const float INF = 3.0e38f;
    for (int e = threadIdx.x; e < 3 * L; e += blockDim.x) sg[e] = INF;
    __syncthreads();
    for (int k = 0; k <= N + M - 2; ++k) {
        float* cur = sg + (k % 3) * L;
        const float* p1 = sg + ((k + 2) % 3) * L;   // diagonal k - 1
        const float* p2 = sg + ((k + 1) % 3) * L;   // diagonal k - 2
        const int ilo = k - (M - 1) > 0 ? k - (M - 1) : 0, ihi = k < N - 1 ? k : N - 1;
        for (int i = ilo + threadIdx.x; i <= ihi; i += blockDim.x) {
            const int j = k - i;
            float best = INF;
            unsigned char d = 0;
            if (i == 0 && j == 0) {
                best = 0.f;
            } else {
                if (i > 0 && j > 0) best = p2[i - 1];                       // (i-1, j-1)
                if (i > 0 && p1[i - 1] < best) { best = p1[i - 1]; d = 1; } // (i-1, j)
                if (j > 0 && p1[i] < best) { best = p1[i]; d = 2; }         // (i, j-1)
            }
            cur[i] = dist[(size_t)i * M + j] + best;
            dir[(size_t)i * M + j] = d;
        }
        __syncthreads();
        // the diagonal that falls out of the window must read as "outside" when its buffer is reused
        float* old = sg + ((k + 1) % 3) * L;
        for (int e = threadIdx.x; e < L; e += blockDim.x) old[e] = INF;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        for (int j = 0; j < M; ++j) path[j] = -1;
        int i = N - 1, j = M - 1, steps = 0;
        for (;;) {
            if (path[j] < 0) path[j] = i;   // first visit from the end: the LAST org frame aligned with target frame j
            ++steps;
            if (i == 0 && j == 0) break;
            const unsigned char d = dir[(size_t)i * M + j];
            if (d == 0) { --i; --j; }
            else if (d == 1) --i;
            else --j;
        }
        out[1] = (float)steps;
        out[2] = sg[((N + M - 2) % 3) * L + (N - 1)];   // accumulated cost at the end point
    }
    __syncthreads();
    // mean over target frames of D(path[j], j), fixed-order block reduction
    float s = 0.f;
    for (int j = threadIdx.x; j < M; j += blockDim.x) s += dist[(size_t)path[j] * M + j];
    __shared__ float red[32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        out[0] = t / (float)M;
    }
}

// mean / std (population) of the frame MCD of two aligned sequences; optional frame index lists (speech frames)
__global__ void __launch_bounds__(256) k_mcd_aligned(int n, int Dm, const float* __restrict__ x, int ldx, const float* __restrict__ y, int ldy,
                                                     const long long* __restrict__ ix, const long long* __restrict__ iy, float* __restrict__ out) {
    __shared__ double rs[8], rq[8];
    double s = 0.0, q = 0.0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int t = w; t < n; t += 8) {   // one warp per frame
        const float* xr = x + (size_t)(ix ? ix[t] : t) * ldx;
        const float* yr = y + (size_t)(iy ? iy[t] : t) * ldy;
        float a = 0.f;
        for (int d = lane; d < Dm; d += 32) {
            const float df = xr[d] - yr[d];
            a = fmaf(df, df, a);
        }
        a = warp_sum(a);
        const double m = (double)(MCD_K * sqrtf(a));
        s += m;
        q += m * m;
    }
    if (lane == 0) {
        rs[w] = s;
        rq[w] = q;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double S = 0.0, Q = 0.0;
        for (int i = 0; i < 8; ++i) {
            S += rs[i];
            Q += rq[i];
        }
        const double mean = n > 0 ? S / n : 0.0;
        out[0] = (float)mean;
        out[1] = (float)(n > 0 ? sqrt(fmax(Q / n - mean * mean, 0.0)) : 0.0);
    }
}

}  // namespace cvb

using namespace cvb;

extern "C" {

size_t cvb_dtw_ws_bytes(int N, int M) {
    if (N <= 0 || M <= 0) return 0;
    return round_up_sz((size_t)N * M * sizeof(float), 256) + round_up_sz((size_t)N * M, 256);
}

int cvb_dtw_mcd(int N, int M, int D, const float* org, int ldo, const float* trg, int ldt, void* ws, int32_t* path, float* out3,
                void* stream) {
    CVB_REQUIRE(N > 0 && M > 0 && D > 0 && org && trg && ws && path && out3, "cvb_dtw_mcd: bad arguments (N=%d M=%d D=%d)", N, M, D);
    const size_t smem = (size_t)3 * (N + 1) * sizeof(float);
    DeviceInfo di;
    if (int rc = get_device_info(&di)) return rc;
    CVB_REQUIRE(smem <= (size_t)di.max_smem_optin, "cvb_dtw_mcd: %d source frames need %zu B of shared memory (max %d)", N, smem, di.max_smem_optin);
    cudaStream_t s = (cudaStream_t)stream;
    float* dist = reinterpret_cast<float*>(ws);
    unsigned char* dir = reinterpret_cast<unsigned char*>(ws) + round_up_sz((size_t)N * M * sizeof(float), 256);
    k_dtw_dist<<<dim3(ceil_div(M, 16), ceil_div(N, 16)), 256, 0, s>>>(N, M, D, org, ldo, trg, ldt, dist);
    CVB_LAUNCH_CHECK();
    CVB_CHECK(cudaFuncSetAttribute(k_dtw_dp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_dtw_dp<<<1, 1024, smem, s>>>(N, M, dist, dir, path, out3);
    CVB_LAUNCH_CHECK();
    return 0;
}

int cvb_mcd_aligned(int n, int D, const float* x, int ldx, const float* y, int ldy, const int64_t* idx_x, const int64_t* idx_y,
                    float* out2, void* stream) {
    CVB_REQUIRE(n >= 0 && D > 0 && x && y && out2, "cvb_mcd_aligned: bad arguments");
    k_mcd_aligned<<<1, 256, 0, (cudaStream_t)stream>>>(n, D, x, ldx, y, ldy, reinterpret_cast<const long long*>(idx_x),
                                                       reinterpret_cast<const long long*>(idx_y), out2);
    CVB_LAUNCH_CHECK();
    return 0;
}
}
