#!/usr/bin/env bash
# ncu evidence for one round (B200_PROFILING.md recipe): launch list of a short bench + full captures of the top kernels.
mkdir -p gpurun_out
R=${1:-r01}
BENCH="python bench.py --steps 2 --warmup 3 --sustain-steps 0 --no-graph --skip-cpu --skip-kernels --skip-train --skip-extras"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$R.csv $BENCH > gpurun_out/ncu_launch_$R.log 2>&1
echo "launch list exit $?"
# the five gemm_tc launches of the second forward: Q|V, key proj, out proj, FFN-up, FFN-down
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 5 -c 5 -f -o gpurun_out/prof_gemm_tc_$R $BENCH > gpurun_out/ncu_gemm_$R.log 2>&1
echo "gemm_tc capture exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 1 -c 1 -f -o gpurun_out/prof_attn_$R $BENCH > gpurun_out/ncu_attn_$R.log 2>&1
echo "attn capture exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"ln_rows_kernel|ln_mean_head|scores_kernel" -s 4 -c 4 -f -o gpurun_out/prof_hbm_$R $BENCH > gpurun_out/ncu_hbm_$R.log 2>&1
echo "hbm kernels capture exit $?"
ls -la gpurun_out | tail -20
