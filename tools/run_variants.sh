python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 400 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; tail -2 gpurun_out/bench_e.err
python -c "
import json; d=json.load(open('gpurun_out/bench_e.json')); print(d['value'], d['ms_per_step']); print(d['kernel_ms']); print({k:(round(v['ms_per_step'],4),round(v['ms_per_pd_launch'],4)) for k,v in d['other_workloads'].items()})"
