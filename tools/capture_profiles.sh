# Final captures of a round (run on the GPU box through gpurun): bench line, reference arm, ncu launch list, one
# `ncu --set full` capture per kernel of the fused step, the step's timeline and the phase clocks of k_pd_aba.
set -x
T=${TAG:-r02_final2}
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || exit 1
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-others --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $B > gpurun_out/ncu_launch.log 2>&1
for k in k_pd_aba k_kin_tile k_obs_rec k_obs_feat k_obs_gather; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$k -s 12 -c 1 -o gpurun_out/${T}_${k} -f $B > gpurun_out/ncu_$k.log 2>&1
done
python tools/trace_step.py --graph 2>&1 | tail -6 > gpurun_out/${T}_trace.txt
if [ -f terrainrlsim_b200/lib/libtrl_b200_clk.so ]; then
  TRL_LIB=terrainrlsim_b200/lib/libtrl_b200_clk.so python tools/pd_phases.py biped3d 16384 > gpurun_out/${T}_pd_phases.txt 2>&1
fi
ls -la gpurun_out/${T}* | head -30
