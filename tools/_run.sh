(timeout -s KILL 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -8) > gpurun_out/t2_tests.log 2>&1; cat gpurun_out/t2_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/b_t2d.json 2> gpurun_out/b_t2d.err; python -c "
import json; d=json.load(open('gpurun_out/b_t2d.json')); print(d['value'], d['ms_per_step'], d['roofline']['share_of_step'], d['roofline']['avg_launch_ms'], d['roofline']['kernel'])"
timeout 300 python bench.py --no-cpu-baseline --batch-utt 8 > gpurun_out/b_t2d_b8.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/b_t2d_b8.json')); print('B8', d['value'], d['ms_per_step'], d['roofline']['share_of_step'])"
timeout 300 python tools/step_timeline.py > gpurun_out/t2_step_timeline.txt 2>&1; tail -40 gpurun_out/t2_step_timeline.txt
