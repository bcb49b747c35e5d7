/* Test-only entry points of libcvb_test_hooks.so (tests/native/): tcgen05 building-block self-test and the
 * micro-benchmarks behind tools/bench_ingest.py.  NOT part of the product library or of its ABI (include/cyclevae_b200.h). */
#ifndef CVB_TEST_HOOKS_H
#define CVB_TEST_HOOKS_H
#ifdef __cplusplus
extern "C" {
#endif

/* self-test of the tcgen05 / TMEM / bulk-copy building blocks: D[128,N] = A[128,K] * B[N,K]^T with bf16
 * operands (modes 0/1: A,B fp32 row-major, arranged in-kernel; mode 2: A,B bf16 pre-arranged in the
 * K-major core-matrix order [row/8][k/8][8][8] and fetched with cp.async.bulk).  tests only. */
int cvb_selftest_umma(int mode, int N, int K, const void* A, const void* B, float* D, void* stream);

/* micro-benchmark of the L2 -> shared-memory ingest paths of the persistent kernels (profiling hook):
 * mode 0 = cp.async.bulk, mode 1 = ld.global.v4 + st.shared; out_cycles[grid] = SM cycles for `iters` rounds
 * of `inflight` x `bytes`.  tools/bench_ingest.py. */
int cvb_bench_ingest(int grid, int mode, int bytes, int inflight, int shared_src, int iters, const void* src,
                     long long* out_cycles, void* stream);

/* same for the all-gather pattern (every CTA rewrites a slice each round, global barrier, every CTA ingests all);
 * out_cycles[2*cta] = ingest cycles, [2*cta+1] = write+fence+barrier cycles. */
int cvb_bench_allgather(int grid, int mode, int wmode, int bytes, int inflight, int iters, void* buf, unsigned* ctr,
                        long long* out_cycles, void* stream);

#ifdef __cplusplus
}
#endif
#endif
