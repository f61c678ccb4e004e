#!/usr/bin/env bash
# One multi-GPU box: PCIe probe per GPU subset, the 2-GPU data-parallel parity test, bench.py at N = 8, 4, 2, 1 (driver launch form).
# Usage (from the repo root on the box): bash tools/gpu_scale.sh <tag>      -> gpurun_out/<tag>_*.log
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
python tools/pcie_probe.py > gpurun_out/${TAG}_pcie_probe.log 2>&1
python -m pytest tests/test_gpu_train.py -q -m gpu -k "two_gpu" --timeout 600 > gpurun_out/${TAG}_two_gpu_test.log 2>&1
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_${N}gpu.log 2> gpurun_out/${TAG}_bench_${N}gpu.err
done
python bench.py --gpus 1 --steps 20 --warmup 5 --skip-extras > gpurun_out/${TAG}_bench_1gpu.log 2> gpurun_out/${TAG}_bench_1gpu.err
for N in 8 4 2 1; do
  python - "$N" "$TAG" <<'PY'
import json, sys
n, tag = sys.argv[1], sys.argv[2]
try:
    line = json.loads([l for l in open(f"gpurun_out/{tag}_bench_{n}gpu.log") if l.startswith("{")][-1])
    e = line["e2e"]
    print(f"N={n} value={line['value']:.0f} e2e={e['value']:.0f} (ceiling {e['h2d_ceiling_slides_per_s']:.0f}, frac {e['frac_of_h2d_ceiling']:.3f}) train={json.dumps(line.get('train_step'))}")
except Exception as exc:
    print(f"N={n}: no line ({exc})")
PY
done
tail -5 gpurun_out/${TAG}_two_gpu_test.log
