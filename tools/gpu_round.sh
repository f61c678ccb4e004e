#!/usr/bin/env bash
# One gpurun round: diagnostics + GPU tests + short bench; logs under gpurun_out/.  Never aborts early.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
STAGE=${1:-all}
run() { echo "=== $*" ; timeout 600 "$@" ; echo "=== exit $?"; }
if [[ $STAGE == all || $STAGE == diag ]]; then run python tools/tc_diag.py > gpurun_out/tc_diag.log 2>&1; fi
if [[ $STAGE == all || $STAGE == kernels ]]; then run python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider > gpurun_out/t_kernels.log 2>&1; fi
if [[ $STAGE == all || $STAGE == tc ]]; then run python -m pytest tests/test_gpu_gemm_tc.py -q -m gpu -p no:cacheprovider > gpurun_out/t_tc.log 2>&1; fi
if [[ $STAGE == all || $STAGE == parity ]]; then run python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider > gpurun_out/t_parity.log 2>&1; fi
if [[ $STAGE == all || $STAGE == bench ]]; then run python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; fi
tail -n 25 gpurun_out/tc_diag.log gpurun_out/t_kernels.log gpurun_out/t_tc.log gpurun_out/t_parity.log gpurun_out/bench.log gpurun_out/bench.err 2>/dev/null | cut -c1-300
