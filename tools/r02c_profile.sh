#!/bin/bash
# Round-2 (final build: coalesced GEMM epilogue) measurement pass on the GPU box: tests, every bench workload, ncu launch list, ncu --set full of the top kernels.
# Numbers printed under ncu are never bench values; the bench lines come from the plain runs.
set -u
O=gpurun_out
mkdir -p $O
(timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4) > $O/r02c_tests.log 2>&1
timeout -s KILL 400 python bench.py --steps 20 --warmup 5 > $O/r02c_bench_train_b80.json 2> $O/r02c_bench_train_b80.err
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --batch-utt 8 --no-cpu-baseline > $O/r02c_bench_train_b8.json 2>/dev/null
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --batch-utt 1 --no-cpu-baseline > $O/r02c_bench_train_b1.json 2>/dev/null
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --workload spk4 --no-cpu-baseline > $O/r02c_bench_spk4_b8.json 2>/dev/null
timeout -s KILL 400 python bench.py --steps 5 --warmup 3 --workload decode > $O/r02c_bench_decode.json 2> $O/r02c_bench_decode.err
timeout -s KILL 300 python tools/step_timeline.py > $O/r02c_step_timeline.txt 2>&1
# launch list (per-launch times are cold-cache / serialised: compare SHARES)
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02c_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $O/r02c_launches.log 2>&1
# full captures of the top kernels (one launch each)
for k in k_gru_bwd_tc k_gru_fwd_tc2; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"^$k\$" --launch-skip 12 -c 1 -f -o $O/r02c_$k \
      python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > /dev/null 2>&1
done
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc --launch-skip 60 -c 3 -f -o $O/r02c_k_gemm_tc \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > /dev/null 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_split_group --launch-skip 60 -c 2 -f -o $O/r02c_k_split_group \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > /dev/null 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_gru_fwd_tc_eval --launch-skip 8 -c 1 -f -o $O/r02c_k_gru_fwd_tc_eval \
    python bench.py --steps 1 --warmup 3 --workload decode --no-cpu-baseline > /dev/null 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"^k_gru_bwd_tc2$" --launch-skip 12 -c 1 -f -o $O/r02c_k_gru_bwd_tc2 \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --batch-utt 8 > /dev/null 2>&1
timeout -s KILL 200 python tools/trace_recurrence.py 80 80 > $O/r02c_trace.txt 2>&1
timeout -s KILL 200 python tools/trace_recurrence.py 8 80 > $O/r02c_trace_b8.txt 2>&1
cat $O/r02c_tests.log; cut -c1-420 $O/r02c_bench_train_b80.json; for f in train_b8 train_b1 spk4_b8 decode; do cut -c1-200 $O/r02c_bench_$f.json; done
