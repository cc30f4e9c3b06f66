python -m pytest tests -m gpu -q -x 2>&1 | tail -2
TRL_PD_SPLIT=0 python -m pytest tests -m gpu -q -x -k "imp_pd or fused" 2>&1 | tail -2
python bench.py --steps 400 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; tail -2 gpurun_out/bench_k.err
python -c "
import json; d=json.load(open('gpurun_out/bench_k.json')); print(d['value'], d['ms_per_step'], d['kernel_ms']); print({k:(round(v['ms_per_step'],4),round(v['ms_per_pd_launch'],4)) for k,v in d['other_workloads'].items()})"
