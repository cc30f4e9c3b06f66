for w in 1 2 4; do
  TRL_PD_WPB=$w python bench.py --steps 400 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_w$w.json 2> gpurun_out/bench_w$w.err; tail -2 gpurun_out/bench_w$w.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_w$w.json')); print('wpb=$w', d['value'], d['ms_per_step'], d['kernel_ms']); print({k:(round(v['ms_per_step'],4),round(v['ms_per_pd_launch'],4)) for k,v in d['other_workloads'].items()})"
done
