for k in 2 4; do
  sed -i "s/constexpr int kSB = [0-9];/constexpr int kSB = $k;/" terrainrlsim_b200/csrc/trl_kernels.cu
  python -c "from terrainrlsim_b200 import build; build.build(force=True)"
  python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-others > gpurun_out/bench_k$k.json 2> gpurun_out/bench_k$k.err; tail -2 gpurun_out/bench_k$k.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_k$k.json')); print('kSB=$k', d['ms_per_step'], d['kernel_ms'])"
done
