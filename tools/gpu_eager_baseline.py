"""The "existing Blackwell path" bar of SURVEY.md §8(d): the reference's own PyTorch op sequence (oracle/torch_port.py, the port
that bench.py times on the host cores) run EAGERLY ON THE B200 in fp32 (TF32 off, then on), one cfg2 slide per call, next to
snuffy_b200 on the same slide.  Measurement infrastructure only.   python tools/gpu_eager_baseline.py > profiles/rXX_gpu_eager.log"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from oracle import torch_port  # noqa: E402
from snuffy_b200 import snuffy  # noqa: E402

dev = torch.device("cuda:0")
c = bench.CFG
model, params = bench.build_model(dev)
for layer in model.b_classifier.encoder.layers:
    layer.return_attn = False
tp = {k: v.detach().to(dev) for k, v in params.items()}
g = torch.Generator(device=dev).manual_seed(1234)
xs = [torch.randn(1, c["n"], c["d"], device=dev, generator=g) for _ in range(8)]          # 8 x 20.5 MB > L2


def timed(fn, iters=40, warm=5):
    for i in range(warm):
        fn(xs[i & 7])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(xs[i & 7])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


with torch.no_grad():
    ref = lambda x: torch_port.forward(x, tp, c["heads"], c["K"], c["r"], c["depth"], c["act"])
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        ms = timed(ref)
        print(f"reference op sequence, torch {torch.__version__} eager on the GPU, fp32 (TF32 {'on' if tf32 else 'off'}): "
              f"{ms:.3f} ms per slide = {1e3 / ms:.0f} slides/s", flush=True)
    torch.backends.cuda.matmul.allow_tf32 = False
    ours = timed(lambda x: model(x))
    print(f"snuffy_b200 MILNet.forward (one slide per call, eager launches): {ours:.3f} ms per slide = {1e3 / ours:.0f} slides/s")
    xb = torch.cat(xs + xs, 0)
    for i in range(3):
        snuffy.forward_bags(model, xb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        snuffy.forward_bags(model, xb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10 / 16
    print(f"snuffy_b200 forward_bags (16 slides per call): {ms:.3f} ms per slide = {1e3 / ms:.0f} slides/s")
    a = ref(xs[0])
    b = model(xs[0])
    print("max |bag logit difference| vs the eager reference sequence:", float((a[1] - b[1]).abs().max()),
          " classes:", float((a[0] - b[0]).abs().max()))
