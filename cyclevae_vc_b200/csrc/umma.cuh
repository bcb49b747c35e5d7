// Minimal sm_100a tcgen05 / TMEM / mbarrier / bulk-copy wrappers (inline PTX) used by the
// tensor-core recurrence kernels.  Descriptor bit layouts follow the PTX ISA "tcgen05 matrix
// descriptor" / "instruction descriptor" tables (same fields as cute::UMMA::SmemDescriptor and
// InstrDescriptor in the CUTLASS headers shipped with this image).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cvb {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {   // non-blocking
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---- bulk async copy global -> shared (1-D, contiguous), completion on an mbarrier ---------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// the same copy delivered to the same shared-memory offset of every CTA of the cluster named in cta_mask; each
// destination's mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}
// contiguous range of global memory -> L2 (no destination: a hint, nothing to wait for); bytes a multiple of 16
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (tensor core / bulk copy)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// generic-proxy writes to GLOBAL memory -> visible to async-proxy reads (bulk copies) that are ordered after a later release
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- TMEM -----------------------------------------------------------------------------------------
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- descriptors ------------------------------------------------------------------------------------
// K-major, no swizzle ("interleave"): a core matrix is 8 rows x 16 bytes, stored contiguously (128 B).
//   lbo = byte distance between the two core matrices that are adjacent in K within one MMA (K=16 bf16)
//   sbo = byte distance between core matrices adjacent in M (or N)
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // layout_type (bits 61..63) = 0: SWIZZLE_NONE
}
// kind::f16, A/B = bf16 (K-major both), D = fp32
__device__ __forceinline__ uint32_t idesc_bf16_f32(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16, A/B = fp16 (K-major both), D = fp32
__device__ __forceinline__ uint32_t idesc_f16_f32(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// A operand from TMEM (128 lanes x K/2 32-bit columns, two 16-bit K elements per column), B from shared memory
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_bf16_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// shared memory (matrix descriptor: 128 rows x 256 bits = one K=16 slice of a 16-bit K-major operand) -> 8 TMEM columns
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t dst_tmem, uint64_t src_desc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(dst_tmem), "l"(src_desc) : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b_elect(uint32_t dst_tmem, uint64_t src_desc) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.cp.cta_group::1.128x256b [%0], %1;\n\t}" ::"r"(dst_tmem), "l"(src_desc)
        : "memory");
}
// Warp-convergent variants: every lane executes the statement with warp-uniform operands and one elected
// lane issues (the form that lets ptxas keep descriptors in uniform registers: no R2UR per operand).
__device__ __forceinline__ void mma_bf16_ss_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
        : "memory");
}
// ... arrive(1) on the mbarrier at the same offset in every CTA of the cluster named in cta_mask
__device__ __forceinline__ void mma_commit_multicast_elect(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
        "h"(cta_mask)
        : "memory");
}
// all previously issued MMAs of this thread complete -> arrive(1) on the mbarrier
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: this warp's 32 lanes x N consecutive 32-bit columns -----------------------
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- bf16 split: x = hi + lo (+ O(2^-17 |x|)) -------------------------------------------------------
__device__ __forceinline__ void split_bf16(float x, uint16_t& hi, uint16_t& lo) {
    uint32_t u = __float_as_uint(x);
    uint32_t r = u + 0x7FFFu + ((u >> 16) & 1u);  // round-to-nearest-even to bf16
    hi = (uint16_t)(r >> 16);
    float res = x - __uint_as_float((uint32_t)hi << 16);
    uint32_t v = __float_as_uint(res);
    uint32_t q = v + 0x7FFFu + ((v >> 16) & 1u);
    lo = (uint16_t)(q >> 16);
}

// ---- fp16 split: x = hi + 2^-11 lo (+ O(2^-22 |x|)); 11 + 11 mantissa bits, for bounded operands (|x| < 6e4) ----
// The residual is stored SCALED by 2^11: unscaled it is an fp16 subnormal for every |x| < 0.125 (absolute resolution
// 2^-24 instead of 2^-22 |x|).  The two cross products (hi x lo, lo x hi) therefore live in their own TMEM accumulator
// and are folded in as  main + 2^-11 * corrections  when the accumulators are read (tools/split_error_budget.py).
constexpr float F16_LO_SCALE = 2048.0f;
constexpr float F16_LO_INV = 1.0f / 2048.0f;
__device__ __forceinline__ void split_f16(float x, uint16_t& hi, uint16_t& lo) {
    const float c = fminf(fmaxf(x, -60000.f), 60000.f);
    const __half h = __float2half_rn(c);
    const __half l = __float2half_rn((c - __half2float(h)) * F16_LO_SCALE);
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
}

// ---- thread-block clusters: ranks, distributed shared memory, cluster-scope mbarrier ------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t cluster_addr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// bulk async copy own shared memory -> shared memory of a CTA of the cluster; completion (bytes) on an
// mbarrier of the destination CTA.  dst and bar are shared::cluster addresses (mapa).
__device__ __forceinline__ void bulk_s2c(uint32_t dst_cluster_addr, const void* smem_src, uint32_t bytes, uint32_t cluster_bar_addr) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster_addr),
                 "r"(smem_u32(smem_src)), "r"(bytes), "r"(cluster_bar_addr)
                 : "memory");
}
// 16 bytes from registers straight into the shared memory of a CTA of the cluster; the bytes are counted on an mbarrier of
// the destination CTA (complete_tx).  dst and bar are shared::cluster addresses (mapa).  No local staging, no proxy fence and
// no CTA barrier on the sender's side: every thread sends its own rows as soon as it has them.
__device__ __forceinline__ void st_async_v4(uint32_t dst_cluster_addr, uint4 v, uint32_t cluster_bar_addr) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst_cluster_addr),
                 "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(cluster_bar_addr)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

// ---- partial sums of the y feedback of the training recurrence kernels (gru_tc.cu, gru_tc_bwd.cu) ----------------
// Pairs are numbered q = b * 64 + o (the output axis padded to 64: D3's columns beyond out_dim
// are exact zeros) and reducer CTA r owns the Q = 8 * ceil(8 B / G) pairs [r Q, (r + 1) Q), i.e. whole groups of 8
// consecutive outputs of one row.  CTA c's partial of pair q lives at part[r = q / Q][c][q % Q]: everything reducer r sums
// is ONE contiguous block of G * Q floats (a single bulk copy), a draining thread (fixed b, walking o) writes float4s, and
// the reducer publishes a group as one 16-byte core-matrix row per plane.  (The first version -- pair order [o][b],
// scalar stores addressed as base + index -- compiled to ~20 dependent instructions per store: 2400 cycles per step.)
struct PartWalk {
    unsigned long long addr;   // address of the thread's first float4
    unsigned long long wrap;   // extra bytes when the slot index wraps into the next reducer
    int i, Q;                  // slot of that float4 in the reducer's row
};
static __device__ __forceinline__ PartWalk part_walk(float* part, int c, int G, int Q, int b, int o_first) {
    PartWalk w;
    w.Q = Q;
    const int q0 = b * 64 + o_first;
    const int r0 = q0 / Q;
    w.i = q0 - r0 * Q;
    w.addr = reinterpret_cast<unsigned long long>(part + ((size_t)r0 * G + c) * Q + w.i);
    w.wrap = ((unsigned long long)G * Q - Q) * 4ull;
    asm volatile("" : "+l"(w.addr), "+r"(w.i));   // keep them in registers: no rematerialisation per store
    return w;
}
static __device__ __forceinline__ void st_global_v4(unsigned long long addr, float x, float y, float z, float w) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

}  // namespace umma
}  // namespace cvb
