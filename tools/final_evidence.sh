#!/usr/bin/env bash
# End-of-round evidence on one B200 (run under gpurun): tests, bench (both arms), launch lists, ncu of the GEMMs, parity, stress.
R=${1:-r02o}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -1 gpurun_out/${R}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${R}_bench.log 2> gpurun_out/${R}_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference_arm.log 2>> gpurun_out/${R}_bench.err; echo "reference arm exit $?"
PROF_ONLY=1 STEPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_train.csv python tools/prof_train.py > /dev/null 2>&1; echo "train launch list exit $?"
BENCH="python bench.py --steps 2 --warmup 3 --sustain-steps 0 --no-graph --skip-cpu --skip-kernels --skip-train --skip-extras"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches.csv $BENCH > /dev/null 2>&1; echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 5 -c 5 -f -o gpurun_out/prof_gemm_tc_$R $BENCH > /dev/null 2>&1; echo "gemm capture exit $?"
timeout 300 python tools/parity_table.py > gpurun_out/${R}_parity_table.log 2>&1; echo "parity table exit $?"
timeout 300 python tools/stress.py 60 > gpurun_out/${R}_stress.log 2>&1; echo "stress exit $?"; tail -2 gpurun_out/${R}_stress.log
python tools/prof_train.py 2>&1 | tail -2
