#!/bin/bash
# Round-2 (final build: 8-warp GEMM epilogue) measurement pass on the GPU box: tests, every bench workload, ncu launch list, ncu --set full of the top kernels.
# Numbers printed under ncu are never bench values; the bench lines come from the plain runs.
set -u
O=gpurun_out
mkdir -p $O
(timeout -s KILL 400 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -4) > $O/r02d_tests.log 2>&1
timeout -s KILL 400 python bench.py --steps 20 --warmup 5 > $O/r02d_bench_train_b80.json 2> $O/r02d_bench_train_b80.err
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --batch-utt 8 --no-cpu-baseline > $O/r02d_bench_train_b8.json 2>/dev/null
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --batch-utt 1 --no-cpu-baseline > $O/r02d_bench_train_b1.json 2>/dev/null
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --workload spk4 --no-cpu-baseline > $O/r02d_bench_spk4_b8.json 2>/dev/null
timeout -s KILL 400 python bench.py --steps 5 --warmup 3 --workload decode > $O/r02d_bench_decode.json 2> $O/r02d_bench_decode.err
timeout -s KILL 300 python tools/step_timeline.py > $O/r02d_step_timeline.txt 2>&1
# launch list (per-launch times are cold-cache / serialised: compare SHARES)
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02d_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $O/r02d_launches.log 2>&1
# full captures of the GEMM (forward products, then the backward groups); the recurrence kernels are unchanged since r02c
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc --launch-skip 60 -c 3 -f -o $O/r02d_k_gemm_tc \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > /dev/null 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc --launch-skip 179 -c 4 -f -o $O/r02d_k_gemm_tc_bwd \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > /dev/null 2>&1
cat $O/r02d_tests.log; cut -c1-420 $O/r02d_bench_train_b80.json; for f in train_b8 train_b1 spk4_b8 decode; do cut -c1-200 $O/r02d_bench_$f.json; done
