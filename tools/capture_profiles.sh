set -x
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err || exit 1
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_final_ref.json 2> gpurun_out/r02_bench_final_ref.err
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-others --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv $B > gpurun_out/ncu_launch.log 2>&1
for k in k_pd_aba k_kin_tile k_obs_env k_obs_gather; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$k -s 12 -c 1 -o gpurun_out/r02_${k}_final -f $B > gpurun_out/ncu_$k.log 2>&1
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_obs_samples -s 6 -c 1 -o gpurun_out/r02_k_obs_samples_2d_final -f python tools/run_obs.py dog 4096 3 > gpurun_out/ncu_obs2d.log 2>&1
python tools/trace_step.py --graph 2>&1 | tail -6 > gpurun_out/r02_trace_final.txt
ls -la gpurun_out/*final* | head -20
