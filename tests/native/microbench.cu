// Micro-benchmarks of the data paths the persistent recurrence kernels depend on (profiling hooks, not
// on the product path): L2 -> shared-memory ingest by cp.async.bulk vs. by ld.global.v4 + st.shared.
#include "common.cuh"
#include "umma.cuh"

namespace cvb {
using namespace umma;

// mode 0: one thread issues `inflight` bulk copies of `bytes` each per round, waits for all, repeats.
// mode 1: 128 threads copy inflight*bytes per round with 16-byte loads (all loads issued before the stores).
// shared_src != 0: every CTA reads the same region (all-gather pattern); else CTA-private regions.
__global__ void __launch_bounds__(128, 1) k_bench_ingest(int mode, int bytes, int inflight, int shared_src, int iters,
                                                         const uint8_t* __restrict__ src, long long* __restrict__ out_cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    const size_t region = (size_t)bytes * inflight;
    const uint8_t* base = src + (shared_src ? 0 : (size_t)blockIdx.x * region);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    long long t0 = 0;
    for (int it = -2; it < iters; ++it) {
        if (it == 0) t0 = clock64();
        if (mode == 0) {
            if (threadIdx.x == 0) {
                mbar_expect_tx(&bar, (uint32_t)region);
                for (int q = 0; q < inflight; ++q) bulk_g2s(smem + (size_t)q * bytes, base + (size_t)q * bytes, (uint32_t)bytes, &bar);
            }
            mbar_wait(&bar, (uint32_t)(it + 2) & 1);
        } else {
            const int n16 = (int)(region / 16);
            for (int i0 = 0; i0 < n16; i0 += 128 * 8) {
                uint4 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    int i = i0 + k * 128 + threadIdx.x;
                    if (i < n16) v[k] = __ldcg(reinterpret_cast<const uint4*>(base) + i);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    int i = i0 + k * 128 + threadIdx.x;
                    if (i < n16) reinterpret_cast<uint4*>(smem)[i] = v[k];
                }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = clock64() - t0;
}

// All-gather pattern of the recurrence kernels: every round each CTA rewrites its 1/grid slice of the region
// (wmode 0: st.global, 1: st.global.cg, 2: st.global.wt via __stwt), all CTAs synchronise on a global counter,
// then every CTA ingests the whole region (mode 0 bulk copies / mode 1 ld.global.v4).  out[2*cta] = cycles spent
// in the ingest only, out[2*cta+1] = cycles in write+fence+barrier.
__global__ void __launch_bounds__(128, 1) k_bench_allgather(int mode, int wmode, int bytes, int inflight, int iters, uint8_t* __restrict__ buf,
                                                            unsigned* __restrict__ ctr, long long* __restrict__ out_cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    const size_t region = (size_t)bytes * inflight;
    const int G = gridDim.x;
    const size_t slice = region / G;   // multiple of 16
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    long long t_in = 0, t_sync = 0;
    for (int it = 0; it < iters + 2; ++it) {
        uint8_t* reg = buf + (size_t)(it & 1) * region;
        const long long t0 = clock64();
        uint4* dst = reinterpret_cast<uint4*>(reg + (size_t)blockIdx.x * slice);
        for (int i = threadIdx.x; i < (int)(slice / 16); i += 128) {
            uint4 v = make_uint4(it, i, blockIdx.x, 7);
            if (wmode == 1) __stcg(dst + i, v);
            else if (wmode == 2) __stwt(dst + i, v);
            else dst[i] = v;
        }
        __threadfence();
        fence_proxy_async_all();
        __syncthreads();
        if (threadIdx.x == 0) {
            red_release_gpu_add(ctr, 1u);
            while (ld_acquire_gpu(ctr) < (unsigned)G * (unsigned)(it + 1)) {
            }
            fence_proxy_async_all();
        }
        __syncthreads();
        const long long t1 = clock64();
        if (mode == 0) {
            if (threadIdx.x == 0) {
                mbar_expect_tx(&bar, (uint32_t)region);
                for (int q = 0; q < inflight; ++q) bulk_g2s(smem + (size_t)q * bytes, reg + (size_t)q * bytes, (uint32_t)bytes, &bar);
            }
            mbar_wait(&bar, (uint32_t)it & 1);
        } else {
            const int n16 = (int)(region / 16);
            for (int i0 = 0; i0 < n16; i0 += 128 * 8) {
                uint4 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    int i = i0 + k * 128 + threadIdx.x;
                    if (i < n16) v[k] = __ldcg(reinterpret_cast<const uint4*>(reg) + i);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    int i = i0 + k * 128 + threadIdx.x;
                    if (i < n16) reinterpret_cast<uint4*>(smem)[i] = v[k];
                }
            }
            __syncthreads();
        }
        const long long t2 = clock64();
        if (it >= 2) {
            t_in += t2 - t1;
            t_sync += t1 - t0;
        }
    }
    if (threadIdx.x == 0) {
        out_cycles[2 * blockIdx.x] = t_in;
        out_cycles[2 * blockIdx.x + 1] = t_sync;
    }
}

// GEMM-like operand ingest by thread-block clusters: every round a CTA needs a private block (a_bytes) and a block its
// whole cluster shares (b_bytes).  mode 0: every CTA pulls both itself (unicast); mode 1: CTA r pulls slice r of the
// shared block and multicasts it to all CTAs of the cluster.  b_all != 0: all clusters share the same block (one B tile
// read by a whole wave), else one block per cluster.  Two rounds in flight, one cluster barrier per round.
__global__ void __launch_bounds__(128, 1) k_bench_mcast(int mode, int a_bytes, int b_bytes, int b_all, int iters,
                                                        const uint8_t* __restrict__ src, long long* __restrict__ out_cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[2];
    uint32_t cs, rank, cid;
    asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(cs));
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cid));
    rank = cluster_ctarank();
    const int stage_bytes = a_bytes + b_bytes;
    const uint8_t* a_src = src + (size_t)blockIdx.x * a_bytes;
    const uint8_t* b_src = src + ((size_t)32 << 20) + (b_all ? 0 : (size_t)cid * b_bytes);
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    cluster_sync_all();
    long long t0 = 0;
    for (int it = -4; it <= iters; ++it) {
        if (it == 0) t0 = clock64();
        const int r = it + 4;
        if (threadIdx.x == 0) {
            if (it < iters) {
                uint8_t* dst = smem + (size_t)(r & 1) * stage_bytes;
                mbar_expect_tx(&bar[r & 1], (uint32_t)stage_bytes);
                bulk_g2s(dst, a_src, (uint32_t)a_bytes, &bar[r & 1]);
                if (mode == 0 || cs == 1) {
                    bulk_g2s(dst + a_bytes, b_src, (uint32_t)b_bytes, &bar[r & 1]);
                } else {
                    const uint32_t sl = (uint32_t)b_bytes / cs;
                    bulk_g2s_multicast(dst + a_bytes + rank * sl, b_src + rank * sl, sl, &bar[r & 1], (uint16_t)((1u << cs) - 1));
                }
            }
            if (r > 0) mbar_wait(&bar[(r - 1) & 1], (uint32_t)((r - 1) >> 1) & 1);
        }
        __syncthreads();
        cluster_sync_all();
    }
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = clock64() - t0;
}

}  // namespace cvb

extern "C" int cvb_bench_allgather(int grid, int mode, int wmode, int bytes, int inflight, int iters, void* buf, unsigned* ctr,
                                   long long* out_cycles, void* stream) {
    using namespace cvb;
    const size_t smem = (size_t)bytes * inflight;
    CVB_REQUIRE(bytes % 16 == 0 && smem <= 200 * 1024 && (smem / grid) % 16 == 0 && smem % grid == 0, "bad sizes");
    CVB_CHECK(cudaFuncSetAttribute(k_bench_allgather, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CVB_CHECK(cudaMemsetAsync(ctr, 0, 4, (cudaStream_t)stream));
    uint8_t* b = (uint8_t*)buf;
    void* params[] = {&mode, &wmode, &bytes, &inflight, &iters, &b, &ctr, &out_cycles};
    CVB_CHECK(cudaLaunchCooperativeKernel((const void*)k_bench_allgather, dim3(grid), dim3(128), params, smem, (cudaStream_t)stream));
    count_launch();
    return 0;
}

extern "C" int cvb_bench_ingest(int grid, int mode, int bytes, int inflight, int shared_src, int iters, const void* src,
                                long long* out_cycles, void* stream) {
    using namespace cvb;
    const size_t smem = (size_t)bytes * inflight;
    CVB_REQUIRE(bytes % 16 == 0 && smem <= 200 * 1024, "bad sizes");
    CVB_CHECK(cudaFuncSetAttribute(k_bench_ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_bench_ingest<<<grid, 128, smem, (cudaStream_t)stream>>>(mode, bytes, inflight, shared_src, iters, (const uint8_t*)src, out_cycles);
    CVB_LAUNCH_CHECK();
    return 0;
}

extern "C" int cvb_bench_mcast(int grid, int cluster, int mode, int a_bytes, int b_bytes, int b_all, int iters, const void* src,
                               long long* out_cycles, void* stream) {
    using namespace cvb;
    const size_t smem = 2 * (size_t)(a_bytes + b_bytes);
    CVB_REQUIRE(a_bytes % 16 == 0 && b_bytes % (16 * cluster) == 0 && smem <= 200 * 1024 && grid % cluster == 0, "bad sizes");
    CVB_CHECK(cudaFuncSetAttribute(k_bench_mcast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const uint8_t* s8 = (const uint8_t*)src;
    CVB_CHECK(cudaLaunchKernelEx(&cfg, k_bench_mcast, mode, a_bytes, b_bytes, b_all, iters, s8, out_cycles));
    count_launch();
    return 0;
}
