timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core_recurrence_vs_exact or bench_shape_training or deterministic or shortest or hu512 or flagship_cyc2" 2>&1 | tail -3
for B in 80 8; do
timeout 120 python tools/trace_recurrence.py $B 80 > gpurun_out/tr_x.txt 2>&1
grep "mean step period" gpurun_out/tr_x.txt
grep -B2 -A40 "trace_fwd_2.bin: 81" gpurun_out/tr_x.txt | grep -v "slot free\|chunk . full\|issued" | head -30
done
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/b_t2f.json 2> gpurun_out/b_t2f.err; python -c "
import json; d=json.load(open('gpurun_out/b_t2f.json')); print(d['value'], d['ms_per_step'], d['roofline']['share_of_step'], d['roofline']['avg_launch_ms'], d['roofline']['kernel'])"
