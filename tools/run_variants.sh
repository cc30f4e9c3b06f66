python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python bench.py --steps 400 --warmup 10 --no-cpu-baseline > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; tail -2 gpurun_out/bench_e.err
python -c "
import json; d=json.load(open('gpurun_out/bench_e.json')); print(d['value'], d['ms_per_step'], d['e2e']); print(d['kernel_ms']); print({k:(round(v['ms_per_step'],4),round(v['ms_per_pd_launch'],4)) for k,v in d['other_workloads'].items()})"
TRL_STEP_OVERLAP=0 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-others > gpurun_out/bench_e0.json 2> gpurun_out/bench_e0.err; python -c "
import json; d=json.load(open('gpurun_out/bench_e0.json')); print('no overlap', d['value'], d['e2e'])"
