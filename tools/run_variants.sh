python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for v in "" _mb5 _mb6; do
  TRL_LIB=$PWD/terrainrlsim_b200/lib/libtrl_b200$v.so python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err; tail -2 gpurun_out/bench_v$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_v$v.json')); print('variant=$v', d['ms_per_step'], d['kernel_ms']); print({k:(round(v['ms_per_step'],4),round(v['ms_per_pd_launch'],4)) for k,v in d['other_workloads'].items()})"
done
