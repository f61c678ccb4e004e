#!/usr/bin/env python
"""Headline benchmark: slides/sec of the Snuffy aggregator forward on synthetic CAMELYON16-shaped bags
(BASELINE.json configs[1]: 10 000 patches x 512-d, 8 heads, top-k = 200, depth 1, eval mode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = one pass of the hot path over one batch of B slides per GPU.  Prints ONE JSON line (rank 0).
  value         slides/s, inputs resident in HBM, CUDA-event timed over exactly K steps, max over ranks (weak scaling)
  e2e           slides/s through the public module API with HOST (pinned) bags: H2D of every bag and D2H of the instance
                scores + bag logits are inside the timed region; `h2d_ceiling` = the same copies with no compute
  roofline      the dominant kernel (FFN-up tcgen05 GEMM) timed alone with CUDA events vs MEASURED_PEAKS.json
  kernels       the same for the attention / score / head kernels (the metric's "attn HBM GB/s vs peak")
  variants      the same step with random patches (r = 0.5) and with the attention tensor A materialised
  configs       the other BASELINE.json configs (cfg1, cfg3, cfg4 packed) timed once each, vs the 3-pass tensor ceiling
  gpu_eager_reference   the reference's own PyTorch modules run eagerly on this B200 (fp32, TF32 off): the existing GPU path
  cpu_baseline  the reference's CPU forward on this box's host cores, bounded sample
  train_step    (LAST key) configs[4]: one data-parallel training epoch over 512 synthetic slides, one all-reduce per step
--impl reference runs only the CPU arm: the reference's own modules when they are on the box (baseline/_ref, staged by
tools/stage_reference.py; $SNUFFY_REF; /root/reference), else oracle/torch_port.py (its op-for-op port).
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
TRAIN_SLIDES = 512                      # BASELINE.json configs[4]


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
        self.marks = {}

    def _run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:                                             # NVML directly: ~1 ms per sample (nvidia-smi costs ~0.3 s per call)
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[self.index]) if visible and all(t.strip().isdigit() for t in visible.split(",")) else self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            bits = [getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)]
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self._stop.is_set():
                r = get_reasons(h)
                row = [str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx), str(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)]
                row += ["Active" if r & b else "Not Active" for b in bits]
                self.rows.append((time.perf_counter(), row))
                self._stop.wait(0.004)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append((time.perf_counter(), [s.strip() for s in out.split(",")]))
            except Exception:
                pass
            self._stop.wait(0.1)

    def mark(self, name):
        self.marks[name] = time.perf_counter()

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, pw, reasons, in_region = [], [], [], set(), 0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lo, hi = self.marks.get("load_begin", 0.0), self.marks.get("load_end", float("inf"))
        t0, t1 = self.marks.get("timed_begin", 0.0), self.marks.get("timed_end", float("inf"))
        for ts, r in self.rows:
            if not lo <= ts <= hi:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except Exception:
                continue
            in_region += t0 <= ts <= t1
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_region": int(in_region),
                "power_w_max": max(pw) if pw else None,
                "window": "NVML every ~5 ms (nvidia-smi every 0.1 s if NVML is unavailable) while the GPU runs the SAME step back "
                          "to back: an untimed sustain phase (>= 200 steps) followed at once by the K timed steps"}


# ------------------------------------------------------------------ model + data
def synthetic_params(seed=0, cfg=None):
    """Random-init weights of the cfg2 architecture, train.py's init (xavier-normal 2-D weights, zero biases,
    LayerNorm (1, 0); train.py:70,907-909), keyed by the reference's state_dict names.  torch only."""
    c = cfg or CFG
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


def build_from(mod, cfg, device, multiclass=False):
    """The constructor chain of train.py:862-890 / 924-952 on module `mod` (ours or the reference's)."""
    import copy
    c = cfg
    i_cls = mod.FCLayer(c["d"], c["C"])
    attn = mod.MultiHeadedAttention(c["heads"], c["d"])
    ff = mod.PositionwiseFeedForward(c["d"], 4 * c["d"], c["act"], 0.0)
    if multiclass:
        layer = mod.EncoderLayer(c["d"], copy.deepcopy(attn), copy.deepcopy(ff), c["C"], 0.0, c["K"], c["r"])
    else:
        layer = mod.EncoderLayer(c["d"], copy.deepcopy(attn), copy.deepcopy(ff), 0.0, c["K"], c["r"])
    model = mod.MILNet(i_cls, mod.BClassifier(mod.Encoder(layer, c["depth"]), c["C"], c["d"]))
    params = synthetic_params(0, c)
    model.load_state_dict(params, strict=True)
    return model.to(device).eval(), params


def build_model(device, cfg=None, return_attn=False):
    from snuffy_b200 import snuffy
    model, params = build_from(snuffy, cfg or CFG, device)
    for layer in model.b_classifier.encoder.layers:
        layer.return_attn = return_attn          # nobody consumes A; parity tests keep it on
    return model, params


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


def capture(fn):
    """fn() -> CUDA graph replaying it (None if capture is not possible: the step then runs eagerly, still the CUDA path)."""
    try:
        side = torch.cuda.Stream()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            fn()
            side.synchronize()
            with torch.cuda.graph(gr, stream=side):
                fn()
        torch.cuda.synchronize()
        return gr
    except Exception as e:
        print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)
        torch.cuda.synchronize()
        return None


def reference_dir():
    for cand in (os.environ.get("SNUFFY_REF"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "snuffy.py")):
            return cand
    return None


def load_reference_module(name="snuffy"):
    """The reference's own module, imported under a private name from wherever it is on this box (never from snuffy_b200)."""
    ref = reference_dir()
    if ref is None:
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location(f"_reference_{name}", os.path.join(ref, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------ per-kernel rooflines
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
    # traffic: dram__bytes_read.sum + dram__bytes_write.sum of this launch (ncu --set full), captured at batch 16 and scaled to
    # the batch in use (profiles/README.md names the capture the constants come from)
    res.append(dict(kernel="gemm_tc<256> FFN-up (LN2(y) W1^T, relu, planes out)", bound="tensor",
                    achieved=flops / t / 1e12, peak=peaks["tf_burst"], unit="TFLOP/s", frac=flops / t / 1e12 / peaks["tf_burst"],
                    traffic=TRAFFIC["ffn_up"] * batch / 16, launch_ms=t * 1e3, passes=3, issued_tflops=3 * flops / t / 1e12,
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
                    traffic=TRAFFIC["attn"] * batch / 16, launch_ms=t * 1e3, algorithmic_bytes_per_launch=byts,
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
    # DSMIL bag classifier (dsmil.py:72-92) on a cfg2-shaped bag: the q-MLP and the pooling kernel
    from snuffy_b200 import dsmil
    dm = dsmil.MILNet(dsmil.FCLayer(d, 1), dsmil.BClassifier(d, 1, 0.0, True, False)).to(device).eval()
    xd = [t_[:n] for t_ in xs]
    with torch.no_grad():
        def ds():
            it[0] ^= 1
            dm(xd[it[0]])
        t = cuda_time(ds, 20)
        byts = 2.0 * n * d * 4 + 2.0 * n * 128 * 4
        res.append(dict(kernel="dsmil MILNet forward (scores + q-MLP + critical instance + pooling), one bag", bound="hbm",
                        achieved=byts / t / 1e9, peak=peaks["hbm"], unit="GB/s", frac=byts / t / 1e9 / peaks["hbm"],
                        traffic=None, launch_ms=t * 1e3, useful_tflops=2.0 * n * (d * 128 + 128 * 128) / t / 1e12))
        q = torch.randn(n, 128, device=device, generator=g)
        qm = torch.randn(1, 128, device=device, generator=g)
        wf, bf = torch.randn(1, 1, d, device=device, generator=g), torch.zeros(1, device=device)

        def pool():
            it[0] ^= 1
            ops.dsmil_pool(q, qm, xd[it[0]], wf, bf)
        t = cuda_time(pool, 20)
        byts = n * (128 + d) * 4.0 + n * 4.0
        res.append(dict(kernel="dsmil_pool (softmax over N, A^T V, Conv1d head), one bag", bound="hbm", achieved=byts / t / 1e9,
                        peak=peaks["hbm"], unit="GB/s", frac=byts / t / 1e9 / peaks["hbm"], traffic=None, launch_ms=t * 1e3))
    return res


# dram__bytes_read.sum + dram__bytes_write.sum per launch at 16 slides, from the committed `ncu --set full` captures
TRAFFIC = {"ffn_up": 1585.0e6, "attn": 673.8e6}      # profiles/r02g_gemm_tc_ncu_full.csv, r02g_attn_ncu_full.csv


# ------------------------------------------------------------------ other BASELINE configs, timed once each (rank 0, N = 1)
def other_configs(device, peaks):
    import numpy as np
    from snuffy_b200 import snuffy, snuffy_multiclass
    ceiling = peaks["tf_sustained"] / 3.0                                      # useful TFLOP/s of the 3-pass scheme
    out = []

    def flops(n, d, ksel, depth, C):
        return depth * (4 * n * d * d + 16 * n * d * d + 4 * n * ksel * d + 4 * ksel * d * d) + 2 * n * d * C

    def add(name, ms, slides, fl, **kw):
        out.append(dict(name=name, ms=round(ms, 4), slides_per_s=round(slides / ms * 1e3, 1),
                        frac_of_3pass_ceiling=round(fl / (ms * 1e-3) / 1e12 / ceiling, 3), **kw))

    g = torch.Generator(device=device).manual_seed(7)
    with torch.no_grad():
        # cfg1: the reference's own CPU-runnable case, one bag per call (its calling pattern), eager launches and graph replay
        c1 = dict(n=256, d=384, heads=1, K=32, r=0.0, depth=1, act="relu", C=1)
        m1, _ = build_model(device, c1)
        x1 = torch.randn(1, 256, 384, device=device, generator=g)
        eager = cuda_time(lambda: m1(x1), 50) * 1e3
        gr = capture(lambda: m1(x1))
        ms = cuda_time(gr.replay, 50) * 1e3 if gr is not None else eager
        add("cfg1 256x384 h=1 K=32, one bag per call", ms, 1, flops(256, 384, 32, 1, 1), eager_ms=round(eager, 4))
        x64 = torch.randn(64, 256, 384, device=device, generator=g)
        ms = cuda_time(lambda: snuffy.forward_bags(m1, x64), 20) * 1e3
        add("cfg1 x 64 bags per call", ms, 64, 64 * flops(256, 384, 32, 1, 1))
        # cfg3: multiclass, 4 layers, Ksel = 2*ref ~ 392 (snuffy_multiclass.py:130-171)
        c3 = dict(n=6000, d=768, heads=8, K=200, r=0.5, depth=4, act="relu", C=2)
        m3, _ = build_from(snuffy_multiclass, c3, device, multiclass=True)
        for layer in m3.b_classifier.encoder.layers:
            layer.return_attn = False
        for B in (1, 4):
            x3 = torch.randn(B, 6000, 768, device=device, generator=g)
            ms = cuda_time(lambda: m3(x3), 10) * 1e3
            add(f"cfg3 multiclass 6000x768 C=2 L=4, {B} bag(s) per call", ms, B, B * flops(6000, 768, 392, 4, 2))
        # cfg4: 64 bags, N ~ log-uniform[1k, 50k] (seed 7), packed into one launch sequence, K sweep, r = 0.5
        rs = np.random.RandomState(7)
        lens = np.exp(rs.uniform(np.log(1000), np.log(50000), 64)).astype(int)
        for K in (64, 256, 1024):
            c4 = dict(n=0, d=512, heads=8, K=K, r=0.5, depth=1, act="relu", C=1)
            m4, _ = build_model(device, c4)
            ok = lens >= K
            cu = np.concatenate([[0], np.cumsum(lens[ok])])
            xp = torch.randn(int(cu[-1]), 512, device=device, generator=g)
            ms = cuda_time(lambda: snuffy.forward_packed(m4, xp, cu), 3, warm=2) * 1e3
            add(f"cfg4 packed {int(ok.sum())} bags N~logU[1k,50k] ({int(cu[-1])} rows) K={K} r=0.5", ms, int(ok.sum()),
                sum(flops(int(n_), 512, K, 1, 1) for n_ in lens[ok]))
            del xp
    return out


# ------------------------------------------------------------------ training epoch (configs[4]: DP training loop)
def train_epoch_bench(device, world, rank, slides=TRAIN_SLIDES, warm=3):
    """BASELINE.json configs[4]: ONE data-parallel epoch over `slides` synthetic cfg2 slides (slide i -> rank i mod W, one
    bag per optimizer step per rank like train.py:249-264; AdamW lr 2e-4 betas (0.5, 0.9) wd 5e-3: train.py:58,61,110), the
    whole step (forward, fused loss, backward, gradient packing, the flat-gradient all-reduce, AdamW) replayed as one CUDA
    graph.  Also timed: the same step with no collective (a local trainer), and the all-reduce alone."""
    import torch.distributed as dist
    from snuffy_b200 import dp
    from snuffy_b200._lib import lib
    c = CFG
    mine = dp.shard_slides(slides, rank, world)
    bags = []
    for sid in mine + [slides + rank * warm + k for k in range(warm)]:          # the last `warm` bags are warm-up only
        g = torch.Generator(device=device).manual_seed(1234 + sid)
        bags.append((torch.randn(1, c["n"], c["d"], device=device, generator=g),
                     torch.full((1, c["C"]), float(sid & 1), device=device)))
    res = {}

    def run(data_parallel):
        model, _ = build_model(device)
        trainer = dp.DataParallelTrainer(model, lr=2e-4, betas=(0.5, 0.9), weight_decay=5e-3, cuda_graph=True,
                                         data_parallel=data_parallel)
        for k in range(warm):
            trainer.train_step(*bags[len(mine) + k])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(len(mine)):
            loss = trainer.train_step(*bags[k])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1 and data_parallel:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, float(loss), trainer

    l0 = lib.snuffy_launch_count()
    ms, loss, trainer = run(True)
    res["slides_per_s"] = slides / (ms * 1e-3)
    res["ms_per_step"] = ms / len(mine)
    res["steps_per_rank"] = len(mine)
    res["final_loss"] = loss
    res["graph"] = trainer.graph_mode
    res["kernels_per_step"] = getattr(trainer, "_graph_kernels", None)
    res["allreduce_bytes"] = (trainer.flat.numel + 4) * 4 if world > 1 else 0
    if world > 1:
        # the collective alone: eager all-reduces of the same flat buffer, CUDA events, median of 20
        times = []
        for _ in range(25):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            e0.record()
            trainer.flat.allreduce_sum()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        res["allreduce_ms"] = statistics.median(times[5:])
        res["allreduce"] = "one kernel over NVLink peer memory (csrc/comm.cu)" if trainer.flat.peer is not None else "nccl"
        if trainer.flat.peer is not None:                            # NCCL on a buffer of the same size, for comparison
            other, times = torch.zeros_like(trainer.flat._grad_all), []
            for _ in range(25):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                dist.barrier()
                e0.record()
                dist.all_reduce(other)
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            res["allreduce_nccl_ms"] = statistics.median(times[5:])
        del trainer
        ms_local, _, _ = run(False)                                   # the same step with no collective, on this rank alone
        t = torch.tensor([ms_local], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res["local_ms_per_step"] = float(t.item()) / len(mine)
        res["efficiency_vs_local_step"] = res["local_ms_per_step"] / res["ms_per_step"]
    res["library_launches_total"] = int(lib.snuffy_launch_count() - l0)
    return res


# ------------------------------------------------------------------ the reference's own modules
def reference_cpu_forward_fn(params):
    """-> (fn(x [1, N, d] CPU tensor), kind, description): the reference's own MILNet when it is on this box, else its port."""
    c = CFG
    mod = load_reference_module("snuffy")
    if mod is not None:
        mod.device = torch.device("cpu")          # module attribute only decides where index tensors go
        model, _ = build_from(mod, c, "cpu")
        return (lambda x: model(x)), "reference", f"the reference's own snuffy.MILNet ({reference_dir()}), eval mode"
    from oracle import torch_port
    tp = {k: v.detach().cpu() for k, v in params.items()}
    return (lambda x: torch_port.forward(x, tp, c["heads"], c["K"], c["r"], c["depth"], c["act"])), "port", \
        "oracle/torch_port.py, the op-for-op CPU port of the reference's PyTorch forward (the reference is not on this box)"


def cpu_baseline(params, max_seconds=20.0, min_slides=2):
    c = CFG
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    fn, kind, what = reference_cpu_forward_fn(params)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(1, c["n"], c["d"], generator=g)
    with torch.no_grad():
        fn(x)                                                                              # warm-up
        times, t_start = [], time.perf_counter()
        while len(times) < min_slides or (time.perf_counter() - t_start < max_seconds and len(times) < 50):
            t0 = time.perf_counter()
            fn(x)
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return dict(value=1.0 / med, unit="slides/s", cores=threads, kind=kind,
                sample=f"{len(times)} forwards of one cfg2 slide (median {med * 1e3:.1f} ms) of {what}, torch {torch.__version__}, "
                       f"{threads} threads")


def gpu_eager_reference(device, params):
    """the existing GPU path = the reference's own modules run eagerly by PyTorch on this B200, fp32 with
    TF32 off (torch's default), one cfg2 slide per call (its calling pattern, incl. its per-layer host round trip)."""
    c = CFG
    mod = load_reference_module("snuffy")
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        if mod is not None:
            mod.device = device
            model, _ = build_from(mod, c, device)
            fn, kind = (lambda x: model(x)), "reference"
        else:
            from oracle import torch_port
            tp = {k: v.detach().to(device) for k, v in params.items()}
            fn, kind = (lambda x: torch_port.forward(x, tp, c["heads"], c["K"], c["r"], c["depth"], c["act"])), "port"
        g = torch.Generator(device=device).manual_seed(1234)
        xs = [torch.randn(1, c["n"], c["d"], device=device, generator=g) for _ in range(8)]
        it = [0]

        def step():
            it[0] += 1
            fn(xs[it[0] & 7])
        with torch.no_grad():
            s = cuda_time(step, 30, warm=5)
        return dict(value=1.0 / s, unit="slides/s", ms_per_slide=s * 1e3, kind=kind,
                    what="torch eager fp32 (TF32 off) on this GPU, one cfg2 slide per call, 30 timed calls after 5 warm-up")
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path on the host cores; rank 0 only."""
    if rank != 0:
        return
    c = CFG
    tp = synthetic_params(0)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    fn, kind, what = reference_cpu_forward_fn(tp)
    g = torch.Generator().manual_seed(1234)
    slides_per_step = 1                                   # bounded sample of the batch: one slide per step
    x = torch.randn(1, c["n"], c["d"], generator=g)
    with torch.no_grad():
        for _ in range(max(args.warmup, 1)):
            fn(x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn(x)
        dt = time.perf_counter() - t0
    val = slides_per_step * args.steps / dt
    line = {"impl": "reference", "metric": "slides/sec", "value": val, "unit": "slides/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "slides_per_step": slides_per_step, "device": "host CPU"},
            "cpu_baseline": {"value": val, "unit": "slides/s", "cores": threads, "kind": kind,
                             "sample": f"{args.steps} timed forwards of one cfg2 slide, {what}, {threads} threads"},
            "e2e": {"value": val, "unit": "slides/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ main
def main():
    os.environ.setdefault("SNUFFY_B200_PEER_TIMEOUT_S", "60")           # bench runs: a lost rank fails fast (library default 600 s)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="slides per step per GPU (16 measured best of 4/8/16/32)")
    ap.add_argument("--precision", default=None, choices=[None, "bf16x3", "fp32", "bf16x1"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--sustain-steps", type=int, default=400, help="untimed steps run right before the timed region")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-kernels", action="store_true")
    ap.add_argument("--skip-train", action="store_true")
    ap.add_argument("--skip-extras", action="store_true", help="no variants / other configs / GPU eager reference")
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
    c, B = CFG, args.batch
    n, d = c["n"], c["d"]

    # device-resident synthetic bags: 2 batches x B slides x 20.5 MB, alternated so every step reads > L2 of new data
    g = torch.Generator(device=device).manual_seed(1234 + rank)
    bags = [torch.randn(B, n, d, device=device, generator=g) for _ in range(2)]

    def make_runner(mdl):
        def step(i):
            with torch.no_grad():
                snuffy.forward_bags(mdl, bags[i & 1])
        step(0); step(1)
        torch.cuda.synchronize()
        graphs = None
        if not args.no_graph:
            graphs = [capture(lambda i=i: step(i)) for i in range(2)]
            if any(gr is None for gr in graphs):
                graphs = None
        return (lambda i: graphs[i & 1].replay()) if graphs is not None else step, step, graphs is not None

    run, step, graphed = make_runner(model)
    l0 = lib.snuffy_launch_count()
    step(0)
    torch.cuda.synchronize()
    launches_per_step = lib.snuffy_launch_count() - l0
    for i in range(args.warmup):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        clk.mark("load_begin")
        for i in range(args.sustain_steps):              # untimed: the same step back to back, so the clock samples see the
            run(i)                                       # load the timed steps run under (K steps alone last ~50 ms)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clk.mark("timed_begin")
        e0.record()
        for i in range(args.steps):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        clk.mark("timed_end"); clk.mark("load_end")
        ms_total = e0.elapsed_time(e1)
    clocks = clk.summary()
    if world > 1:
        t = torch.tensor([ms_total], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    value = world * B * args.steps / (ms_total * 1e-3)
    # the same K steps from a cool, idle GPU (round 1's protocol): the part is power-capped, so a short burst runs at a higher
    # clock than the sustained figure above; reported for comparison only
    time.sleep(1.5)
    e0.record()
    for i in range(args.steps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    burst_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([burst_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        burst_ms = float(t.item())

    # ---------------- variants of the same step (rank 0 of N = 1 only): what the headline config leaves out
    variants = None
    if not args.skip_extras and world == 1:
        variants = []
        for label, kw in (("r=0.5: 100 top + 100 random patches per slide (device sampler)", dict(r=0.5)),
                          ("return_attn=True: A [B, h, N, Ksel] written to HBM (1.02 GB per 16 slides)", dict(attn=True))):
            mv, _ = build_model(device, dict(CFG, r=kw.get("r", 0.0)), return_attn=kw.get("attn", False))
            rv, _, _ = make_runner(mv)
            it = [0]

            def one():
                it[0] += 1
                rv(it[0])
            s = cuda_time(one, args.steps, warm=3)
            variants.append({"variant": label, "ms_per_step": s * 1e3, "slides_per_s": B / s})
            del mv, rv

    # ---------------- e2e: host (pinned) bags -> H2D -> forward -> D2H of classes + bag logits, double-buffered
    host = [torch.randn(B, n, d).pin_memory() for _ in range(2)]
    dev_in = [torch.empty(B, n, d, device=device) for _ in range(2)]
    host_bag = [torch.empty(B, c["C"]).pin_memory() for _ in range(2)]
    host_cls = [torch.empty(B, n, c["C"]).pin_memory() for _ in range(2)]
    copy_s, comp_s = torch.cuda.Stream(), torch.cuda.Stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    e2e_steps = max(args.steps, 10)                      # copy-bound (PCIe): enough steps to amortise the pipeline fill and drain

    def e2e_loop(steps, compute=True):
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
                if compute:
                    classes, bag, _ = snuffy.forward_bags(model, dev_in[s])
                    host_bag[s].copy_(bag, non_blocking=True)
                    host_cls[s].copy_(classes.view(B, n, -1), non_blocking=True)        # every instance score, like the module returns
                freed[s].record(comp_s)
        comp_s.synchronize(); copy_s.synchronize()

    def timed_loop(compute):
        e2e_loop(2, compute)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e2e_loop(e2e_steps, compute)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    h2d_dt = timed_loop(False)                           # the copies alone: this box's pinned-H2D ceiling at this many ranks
    e2e_dt = timed_loop(True)
    h2d_bytes = B * n * d * 4
    ceiling_gbs = world * e2e_steps * h2d_bytes / h2d_dt / 1e9
    e2e = {"value": world * B * e2e_steps / e2e_dt, "unit": "slides/s", "h2d_bytes_per_step": h2d_bytes,
           "d2h_bytes_per_step": (B * n * c["C"] + B * c["C"]) * 4, "steps": e2e_steps,
           "h2d_GBps": world * e2e_steps * h2d_bytes / e2e_dt / 1e9, "h2d_ceiling_GBps": ceiling_gbs,
           "h2d_ceiling_slides_per_s": ceiling_gbs * 1e9 / (n * d * 4),
           "frac_of_h2d_ceiling": h2d_dt / e2e_dt,
           "bound": "PCIe: the fp32 bag (20.48 MB per slide) has to cross the host link; h2d_ceiling is the same copy loop on "
                    "all ranks at once with no compute (profiles/: pcie probe per GPU subset)",
           "how": "pinned host bags -> cudaMemcpyAsync on a copy stream -> snuffy.forward_bags -> D2H of classes [B, N, 1] and "
                  "bag logits; two buffers, copy of step i+1 overlaps compute of step i; wall clock around the loop, max over ranks"}

    train = None
    if not args.skip_train:
        try:
            train = train_epoch_bench(device, world, rank)
        except Exception as exc:                            # keep the bench line; say why
            train = {"error": repr(exc)[:300]}
    kernels = [] if (args.skip_kernels or rank != 0) else kernel_rooflines(model, B, peaks, device)
    configs = eager_ref = None
    if not args.skip_extras and world == 1:
        configs = other_configs(device, peaks)
        eager_ref = gpu_eager_reference(device, params)
    if world > 1:
        dist.barrier()
    if rank == 0:
        F = c["depth"] * (4 * n * d * d + 16 * n * d * d + 4 * n * c["K"] * d + 4 * c["K"] * d * d) + 2 * n * d
        line = {
            "metric": "slides/sec", "value": value, "unit": "slides/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "value_burst_after_idle": world * B * args.steps / (burst_ms * 1e-3),
            "vs_baseline": None, "dtype": "f32 (tensor-core products as 3-pass split bf16, fp32 accumulate)"
            if precision == "bf16x3" else precision, "data": "synthetic",
            "config": {"workload": WORKLOAD, "slides_per_step_per_gpu": B, "precision": precision,
                       "cuda_graph": graphed, "return_attn": False, "random_patch_share": c["r"],
                       "sustain_steps_before_timed_region": args.sustain_steps,
                       "l2": "two alternating input batches of %.0f MB each (> 126 MB L2)" % (B * n * d * 4 / 1e6),
                       "useful_gflop_per_slide": F / 1e9},
            "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps), "gpu_launches_per_step": int(launches_per_step),
            "clocks": clocks,
            "end_to_end_tensor_frac": (value / world) * F / 1e12 / peaks["tf_sustained"],
            "end_to_end_frac_of_3pass_ceiling": (value / world) * F / 1e12 / (peaks["tf_sustained"] / 3.0),
        }
        if kernels:
            line["roofline"] = {k: kernels[0][k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")}
            line["roofline"]["kernel"] = kernels[0]["kernel"]
            line["roofline"]["peak_source"] = kernels[0].get("peak_source")
            line["kernels"] = kernels
        if variants:
            line["variants"] = variants
        if configs:
            line["configs"] = configs
        if eager_ref:
            line["gpu_eager_reference"] = eager_ref
        if not args.skip_cpu and world == 1:                # the CPU arm is timed beside the GPU number at N = 1 only
            line["cpu_baseline"] = cpu_baseline(params)
        if train:                                           # LAST and compact: it has to survive in the tail of the line
            line["train_step"] = {"workload": f"cfg5: one DP epoch, {TRAIN_SLIDES} cfg2 slides over {world} GPU(s), 1 bag/step/GPU",
                                  **{k: (round(v, 4) if isinstance(v, float) else v) for k, v in train.items()}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
