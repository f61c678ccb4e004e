#!/usr/bin/env python
"""Headline benchmark: slides/sec of the Snuffy aggregator forward on synthetic CAMELYON16-shaped bags
(BASELINE.json configs[1]: 10 000 patches x 512-d, 8 heads, top-k = 200, depth 1, eval mode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = one pass of the hot path over one batch of B slides per GPU.  Prints ONE JSON line (rank 0).
  value     slides/s, inputs resident in HBM, CUDA-event timed, max over ranks (weak scaling: B slides per GPU)
  e2e       slides/s through the public module API with HOST (pinned) bags: H2D of every bag and D2H of the
            predictions are inside the timed region (double-buffered on a copy stream)
  roofline  the dominant kernel (FFN-up tcgen05 GEMM) timed alone with CUDA events vs MEASURED_PEAKS.json
  kernels   the same for the attention / score / head kernels (the metric's "attn HBM GB/s vs peak")
  cpu_baseline  oracle/torch_port.py (CPU port of the reference forward) on this box's host cores, bounded sample
--impl reference runs only that CPU port (the reference is pure PyTorch; /root/reference is not on the box).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(n=10000, d=512, heads=8, K=200, r=0.0, depth=1, act="relu", C=1)
WORKLOAD = "cfg2: CAMELYON16-shaped bags 10000x512 fp32, 8 heads, top-k=200, depth 1, eval forward"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------ clocks sampler (B200_PROFILING.md recipe)
class ClockSampler:
    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ model + data
def synthetic_params(seed=0):
    """Random-init weights of the cfg2 architecture, train.py's init (xavier-normal 2-D weights, zero biases,
    LayerNorm (1, 0); train.py:70,907-909), keyed by the reference's state_dict names.  torch only."""
    c = CFG
    d, dff, C = c["d"], 4 * c["d"], c["C"]
    g = torch.Generator().manual_seed(seed)

    def xavier(out_f, in_f):
        return torch.randn(out_f, in_f, generator=g) * (2.0 / (in_f + out_f)) ** 0.5

    p = {"i_classifier.fc.0.weight": xavier(C, d), "i_classifier.fc.0.bias": torch.zeros(C)}
    for l in range(c["depth"]):
        pre = f"b_classifier.encoder.layers.{l}."
        for i in range(4):
            p[pre + f"self_attn.linears.{i}.weight"] = xavier(d, d)
            p[pre + f"self_attn.linears.{i}.bias"] = torch.zeros(d)
        p[pre + "feed_forward.w_1.weight"], p[pre + "feed_forward.w_1.bias"] = xavier(dff, d), torch.zeros(dff)
        p[pre + "feed_forward.w_2.weight"], p[pre + "feed_forward.w_2.bias"] = xavier(d, dff), torch.zeros(d)
        for s_ in range(2):
            p[pre + f"sublayer.{s_}.norm.weight"], p[pre + f"sublayer.{s_}.norm.bias"] = torch.ones(d), torch.zeros(d)
    p["b_classifier.encoder.norm.weight"], p["b_classifier.encoder.norm.bias"] = torch.ones(d), torch.zeros(d)
    p["b_classifier.linear.weight"], p["b_classifier.linear.bias"] = xavier(C, d), torch.zeros(C)
    return p


def build_model(device):
    import copy
    from snuffy_b200 import snuffy
    c = CFG
    i_cls = snuffy.FCLayer(c["d"], c["C"])
    attn = snuffy.MultiHeadedAttention(c["heads"], c["d"])
    ff = snuffy.PositionwiseFeedForward(c["d"], 4 * c["d"], c["act"], 0.0)
    layer = snuffy.EncoderLayer(c["d"], copy.deepcopy(attn), copy.deepcopy(ff), 0.0, c["K"], c["r"])
    model = snuffy.MILNet(i_cls, snuffy.BClassifier(snuffy.Encoder(layer, c["depth"]), c["C"], c["d"]))
    params = synthetic_params(0)
    model.load_state_dict(params, strict=True)
    return model.to(device).eval(), params


def cuda_time(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3      # seconds per call


def kernel_rooflines(model, batch, peaks, device):
    """Time the individual hot kernels alone (CUDA events on the launching stream) on cfg2-shaped operands."""
    from snuffy_b200 import ops
    c = CFG
    n, d, h, ks = c["n"], c["d"], c["heads"], c["K"]
    rows = batch * n
    layer = model.b_classifier.encoder.layers[0]
    w = layer.layer_weights()
    w.prepare("bf16x3")
    g = torch.Generator(device=device).manual_seed(99)
    x = torch.randn(rows, d, device=device, generator=g)
    # two operand sets so consecutive launches do not re-read the same 164 MB from L2
    xs = [x, torch.randn(rows, d, device=device, generator=g)]
    planes = [ops.ln_rows(t, w.g2, w.be2, want_planes=True)[1] for t in xs]
    dff = w.w1.shape[0]
    res = []
    it = [0]

    def ffn_up():
        it[0] ^= 1
        ops.gemm_tc(planes[it[0]], w.w1_planes, M=rows, N=dff, K=d, passes=3, bias=w.b1, act="relu", want_out=False,
                    want_planes=True)
    t = cuda_time(ffn_up, 10)
    flops = 2.0 * rows * dff * d
    # traffic: dram__bytes_read.sum + dram__bytes_write.sum of this launch (ncu --set full): 168.2 MB read + 597.0 MB written at
    # batch 8 (profiles/r01c_gemm_tc_ncu_full.csv), 332 MB + 1255 MB at batch 16 (profiles/r01k_gemm_tc_ncu_full.csv); the figure
    # reported is the batch-16 capture scaled to the batch in use
    res.append(dict(kernel="gemm_tc<256> FFN-up (LN2(y) W1^T, relu, planes out)", bound="tensor",
                    achieved=flops / t / 1e12, peak=peaks["tf_burst"], unit="TFLOP/s", frac=flops / t / 1e12 / peaks["tf_burst"],
                    traffic=1587.0e6 * batch / 16, launch_ms=t * 1e3, passes=3, issued_tflops=3 * flops / t / 1e12,
                    algorithmic_flops_per_launch=flops, peak_source=peaks["source"] + " bf16 burst"))
    hp = [ops.gemm_tc(p, w.w1_planes, M=rows, N=dff, K=d, passes=3, bias=w.b1, act="relu", want_out=False,
                      want_planes=True)[2] for p in planes]

    def ffn_down():
        it[0] ^= 1
        ops.gemm_tc(hp[it[0]], w.w2_planes, M=rows, N=d, K=dff, passes=3, bias=w.b2, resid=xs[it[0]])
    t = cuda_time(ffn_down, 10)
    res.append(dict(kernel="gemm_tc FFN-down (+bias +residual)", bound="tensor", achieved=flops / t / 1e12,
                    peak=peaks["tf_burst"], unit="TFLOP/s", frac=flops / t / 1e12 / peaks["tf_burst"], traffic=None,
                    launch_ms=t * 1e3, passes=3))
    def qv_proj():
        it[0] ^= 1
        ops.gemm_tc(planes[it[0]], w.wqv_planes, M=rows, N=2 * d, K=d, passes=3, bias=w.bqv, want_out=False, want_planes=True)
    t = cuda_time(qv_proj, 10)
    fl = 2.0 * rows * 2 * d * d
    res.append(dict(kernel="gemm_tc Q|V projection (planes out)", bound="tensor", achieved=fl / t / 1e12,
                    peak=peaks["tf_burst"], unit="TFLOP/s", frac=fl / t / 1e12 / peaks["tf_burst"], traffic=None,
                    launch_ms=t * 1e3, passes=3))
    qvp = [ops.gemm_tc(p, w.wqv_planes, M=rows, N=2 * d, K=d, passes=3, bias=w.bqv, want_out=False, want_planes=True)[2]
           for p in planes]
    kp = torch.randn(batch * ks, d, device=device, generator=g)

    def attn():
        it[0] ^= 1
        ops.sparse_attn_tc(qvp[it[0]], kp, batch, n, ks, h, d, want_probs=False)
    t = cuda_time(attn, 10)
    byts = batch * (2.0 * n * d * 4 + 2.0 * ks * d * 4)
    res.append(dict(kernel="attn_tc (QK^T -> softmax -> P^T V on tcgen05, A not materialised) + fold", bound="hbm",
                    achieved=byts / t / 1e9, peak=peaks["hbm"], unit="GB/s", frac=byts / t / 1e9 / peaks["hbm"],
                    traffic=714.4e6 * batch / 16,          # profiles/r01k_attn_ncu_full.csv (16 slides): 663.2 MB read + 51.2 MB written
                    launch_ms=t * 1e3, algorithmic_bytes_per_launch=byts,
                    useful_tflops=batch * 4.0 * n * ks * d / t / 1e12))
    xs_sel = torch.randn(batch * ks, d, device=device, generator=g)

    def small_proj():
        _, xsp, _ = ops.ln_rows(xs_sel, None, None, apply_ln=False, want_planes=True)
        ops.gemm_tc(xsp, w.wk_planes, M=batch * ks, N=d, K=d, passes=3, bias=w.bk)
    t = cuda_time(small_proj, 10)
    res.append(dict(kernel="key / output projection [B*Ksel, d] x [d, d]: split to planes + gemm_tc", bound="tensor",
                    achieved=2.0 * batch * ks * d * d / t / 1e12, peak=peaks["tf_burst"], unit="TFLOP/s",
                    frac=2.0 * batch * ks * d * d / t / 1e12 / peaks["tf_burst"], traffic=None, launch_ms=t * 1e3))
    wi = model.i_classifier.fc[0]

    def sc():
        it[0] ^= 1
        ops.scores(xs[it[0]], wi.weight.detach(), wi.bias.detach())
    t = cuda_time(sc, 20)
    byts = rows * d * 4.0 + rows * 4.0
    res.append(dict(kernel="scores GEMV", bound="hbm", achieved=byts / t / 1e9, peak=peaks["hbm"], unit="GB/s",
                    frac=byts / t / 1e9 / peaks["hbm"], traffic=None, launch_ms=t * 1e3))
    enc, bc = model.b_classifier.encoder, model.b_classifier

    def head():
        it[0] ^= 1
        ops.ln_mean_head(xs[it[0]].view(batch, n, d), enc.norm.weight.detach(), enc.norm.bias.detach(),
                         bc.linear.weight.detach(), bc.linear.bias.detach())
    t = cuda_time(head, 20)
    byts = rows * d * 4.0
    res.append(dict(kernel="ln_mean_head", bound="hbm", achieved=byts / t / 1e9, peak=peaks["hbm"], unit="GB/s",
                    frac=byts / t / 1e9 / peaks["hbm"], traffic=None, launch_ms=t * 1e3))

    def lnsplit():
        it[0] ^= 1
        ops.ln_rows(xs[it[0]], w.g1, w.be1, want_planes=True)
    t = cuda_time(lnsplit, 20)
    byts = rows * d * 8.0
    res.append(dict(kernel="ln_rows -> split-bf16 planes", bound="hbm", achieved=byts / t / 1e9, peak=peaks["hbm"],
                    unit="GB/s", frac=byts / t / 1e9 / peaks["hbm"], traffic=None, launch_ms=t * 1e3))
    return res


# ------------------------------------------------------------------ training step (configs[4]: DP training loop)
def train_throughput(device, world, steps=10, warm=3, graph=True):
    """Train-mode step of the reference loop (train.py:249-264: one bag per optimizer step per process) at cfg2:
    forward + fused loss + backward (tensor-core products as 3-pass split bf16) + one all-reduce of the flat gradient +
    AdamW.  graph=True: forward .. gradient packing replayed as one CUDA graph.  slides/s over all ranks; CUDA events, max
    over ranks done by the caller."""
    from snuffy_b200 import dp
    model, _ = build_model(device)
    for layer in model.b_classifier.encoder.layers:
        layer.return_attn = False
    trainer = dp.DataParallelTrainer(model, lr=2e-4, betas=(0.5, 0.9), weight_decay=5e-3,     # train.py:58,61,110
                                     cuda_graph=graph)
    c = CFG
    g = torch.Generator(device=device).manual_seed(4321 + int(os.environ.get("RANK", 0)))
    bags = [torch.randn(1, c["n"], c["d"], device=device, generator=g) for _ in range(8)]      # 8 x 20.5 MB > L2
    labels = [torch.full((1, c["C"]), float(i & 1), device=device) for i in range(8)]
    for i in range(warm):
        trainer.train_step(bags[i & 7], labels[i & 7])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = trainer.train_step(bags[i & 7], labels[i & 7])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return ms, steps, float(loss), trainer.flat.numel * 4, getattr(trainer, "_graph_kernels", None)


# ------------------------------------------------------------------ CPU baseline (port of the reference forward)
def cpu_baseline(params, max_seconds=25.0, min_slides=2):
    from oracle import torch_port
    c = CFG
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    tp = {k: v.detach().cpu() for k, v in params.items()}
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(1, c["n"], c["d"], generator=g)
    with torch.no_grad():
        torch_port.forward(x, tp, c["heads"], c["K"], c["r"], c["depth"], c["act"])      # warm-up
        times, t_start = [], time.perf_counter()
        while len(times) < min_slides or (time.perf_counter() - t_start < max_seconds and len(times) < 50):
            t0 = time.perf_counter()
            torch_port.forward(x, tp, c["heads"], c["K"], c["r"], c["depth"], c["act"])
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return dict(value=1.0 / med, unit="slides/s", cores=threads, kind="port",
                sample=f"{len(times)} forwards of one cfg2 slide (median {med * 1e3:.1f} ms) of oracle/torch_port.py, "
                       f"the CPU port of the reference's PyTorch forward, torch {torch.__version__}, {threads} threads")


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (its PyTorch op sequence) on the host cores; rank 0 only."""
    if rank != 0:
        return
    c = CFG
    tp = synthetic_params(0)
    from oracle import torch_port
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(1234)
    slides_per_step = 1                                   # bounded sample of the batch: one slide per step
    x = torch.randn(1, c["n"], c["d"], generator=g)
    with torch.no_grad():
        for _ in range(max(args.warmup, 1)):
            torch_port.forward(x, tp, c["heads"], c["K"], c["r"], c["depth"], c["act"])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            torch_port.forward(x, tp, c["heads"], c["K"], c["r"], c["depth"], c["act"])
        dt = time.perf_counter() - t0
    val = slides_per_step * args.steps / dt
    line = {"impl": "reference", "metric": "slides/sec", "value": val, "unit": "slides/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "slides_per_step": slides_per_step, "device": "host CPU"},
            "cpu_baseline": {"value": val, "unit": "slides/s", "cores": threads, "kind": "port",
                             "sample": f"{args.steps} timed forwards of one cfg2 slide, oracle/torch_port.py "
                                       f"(CPU port of the reference's PyTorch forward), {threads} threads"},
            "e2e": {"value": val, "unit": "slides/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="slides per step per GPU (16 measured best of 4/8/16/32)")
    ap.add_argument("--precision", default=None, choices=[None, "bf16x3", "fp32", "bf16x1"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-kernels", action="store_true")
    ap.add_argument("--skip-train", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from snuffy_b200 import snuffy
    from snuffy_b200._lib import lib
    if args.precision:
        os.environ["SNUFFY_B200_PRECISION"] = args.precision
    precision = os.environ.get("SNUFFY_B200_PRECISION", "bf16x3")
    peaks = load_peaks()
    model, params = build_model(device)
    for layer in model.b_classifier.encoder.layers:
        layer.return_attn = False            # nobody consumes A; parity tests keep it on
    c, B = CFG, args.batch
    n, d = c["n"], c["d"]

    # device-resident synthetic bags: 2 batches x B slides x 20.5 MB, alternated so every step reads > L2 of new data
    g = torch.Generator(device=device).manual_seed(1234 + rank)
    bags = [torch.randn(B, n, d, device=device, generator=g) for _ in range(2)]
    out_slot = [None]

    def step(i):
        with torch.no_grad():
            out_slot[0] = snuffy.forward_bags(model, bags[i & 1])

    graphs = None
    use_graph = not args.no_graph
    step(0); step(1)
    torch.cuda.synchronize()
    if use_graph:
        try:
            graphs = []
            side = torch.cuda.Stream()
            for i in range(2):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    step(i)
                    side.synchronize()
                    with torch.cuda.graph(gr, stream=side):
                        step(i)
                graphs.append(gr)
            torch.cuda.synchronize()
        except Exception as e:                             # still the CUDA path, just launched eagerly
            print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)
            graphs = None
            torch.cuda.synchronize()

    def run(i):
        if graphs is not None:
            graphs[i & 1].replay()
        else:
            step(i)

    l0 = lib.snuffy_launch_count()
    step(0)
    launches_per_step = lib.snuffy_launch_count() - l0
    for i in range(args.warmup):
        run(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for i in range(args.steps):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        ms_total = e0.elapsed_time(e1)
        if ms_total < 1500:                                # keep the sampler alive long enough for a few samples
            t_end = time.time() + 1.0
            j = 0
            while time.time() < t_end:
                run(j); j += 1
                torch.cuda.synchronize()
    clocks = clk.summary()
    if world > 1:
        t = torch.tensor([ms_total], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---------------- e2e: host (pinned) bags -> H2D -> forward -> D2H predictions, double-buffered
    host = [torch.randn(B, n, d).pin_memory() for _ in range(2)]
    dev_in = [torch.empty(B, n, d, device=device) for _ in range(2)]
    host_out = [torch.empty(B, c["C"]).pin_memory() for _ in range(2)]
    host_cls = [torch.empty(B, c["C"]).pin_memory() for _ in range(2)]
    copy_s, comp_s = torch.cuda.Stream(), torch.cuda.Stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    e2e_steps = max(args.steps, 10)                      # copy-bound (PCIe): enough steps to amortise the pipeline fill and drain

    def e2e_loop(steps):
        for ev in freed:
            ev.record(comp_s)
        for i in range(steps):
            s = i & 1
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(freed[s])
                dev_in[s].copy_(host[s], non_blocking=True)
                ready[s].record(copy_s)
            with torch.cuda.stream(comp_s), torch.no_grad():
                comp_s.wait_event(ready[s])
                classes, bag, _ = snuffy.forward_bags(model, dev_in[s])
                host_out[s].copy_(bag, non_blocking=True)
                host_cls[s].copy_(classes.view(B, n, -1)[:, 0, :], non_blocking=True)   # D2H of per-slide outputs
                freed[s].record(comp_s)
        comp_s.synchronize(); copy_s.synchronize()

    e2e_loop(2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_loop(e2e_steps)
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e = {"value": world * B * e2e_steps / e2e_dt, "unit": "slides/s", "h2d_bytes_per_step": B * n * d * 4,
           "d2h_bytes_per_step": 2 * B * c["C"] * 4, "steps": e2e_steps,
           "how": "pinned host bags -> cudaMemcpyAsync on a copy stream -> snuffy.forward_bags -> D2H predictions; "
                  "two buffers, copy of step i+1 overlaps compute of step i; wall clock around the loop"}

    train = None
    if not args.skip_train:
        lc0 = lib.snuffy_launch_count()
        e_ms, e_steps, e_loss, grad_bytes, _ = train_throughput(device, world, graph=False)
        eager_launches = int((lib.snuffy_launch_count() - lc0) / (e_steps + 3))                # 3 warm-up steps
        graph_error = None
        try:
            t_ms, t_steps, t_loss, grad_bytes, graph_kernels = train_throughput(device, world, graph=True)
        except Exception as exc:                            # keep the bench line: report the eager step and say why
            graph_error = repr(exc)[:300]
            t_ms, t_steps, t_loss, graph_kernels = e_ms, e_steps, e_loss, None
        if world > 1:
            t = torch.tensor([t_ms, e_ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ms, e_ms = float(t[0].item()), float(t[1].item())
        train = {"value": world * t_steps / (t_ms * 1e-3), "unit": "slides/s", "ms_per_step": t_ms / t_steps,
                 "bags_per_step_per_gpu": 1, "steps": t_steps, "final_loss": t_loss,
                 "allreduce_bytes_per_step": grad_bytes if world > 1 else 0,
                 "launches_per_step": eager_launches, "kernels_in_graph": graph_kernels,
                 "eager_ms_per_step": e_ms / e_steps,
                 "what": "train.py-style step at cfg2 (train mode, attention dropout 0.1, one bag per optimizer step): forward + "
                         "fused MIL loss + backward (tensor-core products as 3-pass split bf16) + gradient packing replayed as ONE "
                         "CUDA graph (dropout drawn from a device step counter), then one flat-gradient all-reduce + flat AdamW; "
                         "eager_ms_per_step = the same step without the graph (host-launch bound)"}
        if graph_error:
            train["graph_error"] = graph_error
    kernels = [] if (args.skip_kernels or rank != 0) else kernel_rooflines(model, B, peaks, device)
    if world > 1:
        dist.barrier()
    if rank == 0:
        F = c["depth"] * (4 * n * d * d + 16 * n * d * d + 4 * n * c["K"] * d + 4 * c["K"] * d * d) + 2 * n * d
        line = {
            "metric": "slides/sec", "value": value, "unit": "slides/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (tensor-core products as 3-pass split bf16, fp32 accumulate)"
            if precision == "bf16x3" else precision, "data": "synthetic",
            "config": {"workload": WORKLOAD, "slides_per_step_per_gpu": B, "precision": precision,
                       "cuda_graph": graphs is not None, "return_attn": False,
                       "l2": "two alternating input batches of %.0f MB each (> 126 MB L2)" % (B * n * d * 4 / 1e6),
                       "useful_gflop_per_slide": F / 1e9},
            "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps), "gpu_launches_per_step": int(launches_per_step),
            "clocks": clocks,
            "end_to_end_tensor_frac": (value / world) * F / 1e12 / peaks["tf_sustained"],
        }
        if train:
            line["train_step"] = train
        if kernels:
            line["roofline"] = {k: kernels[0][k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")}
            line["roofline"]["kernel"] = kernels[0]["kernel"]
            line["roofline"]["peak_source"] = kernels[0].get("peak_source")
            line["kernels"] = kernels
        if not args.skip_cpu and world == 1:                # the CPU arm is timed beside the GPU number at N = 1 only
            line["cpu_baseline"] = cpu_baseline(params)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
