#!/usr/bin/env python
"""Aggregate pinned host->device ceiling of the box, per GPU subset — the limiter of bench.py's e2e leg (20.48 MB per slide).

    python tools/pcie_probe.py [--mb 328] [--iters 20] > profiles/rXX_pcie_probe.log

For every subset of GPUs below, one process per GPU copies `--mb` MB pinned buffers H2D back to back, all processes start
together (barrier), and the line reports per-GPU and aggregate GB/s (wall clock of the slowest process).  Variants:
  plain      one cudaMemcpyAsync per buffer on one stream (what bench.py does)
  chunk4     the same bytes as 4 copies on 2 streams (copy engines interleave)
  affinity   the process pinned to the CPUs nvidia-smi lists for its GPU before allocating the pinned buffers
Also prints the PCIe path of every GPU (sysfs parents: GPUs that share an upstream switch port share its x16 uplink), the
nvidia-smi topology matrix, and the NUMA layout.  Measurement infrastructure only.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import torch
import torch.multiprocessing as mp


def _worker(rank, gpus, mb, iters, variant, barrier, out):
    dev = gpus[rank]
    torch.cuda.set_device(dev)
    if variant == "affinity":
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(dev)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = [i for i in range(os.cpu_count()) if (words[i // 64] >> (i % 64)) & 1]
            if cpus:
                os.sched_setaffinity(0, cpus)
        except Exception as exc:                                   # report, keep going with the default affinity
            out[f"affinity_error_{dev}"] = repr(exc)[:200]
    n = mb * 1000 * 1000 // 4
    host = [torch.empty(n).pin_memory() for _ in range(2)]
    for h_ in host:
        h_.fill_(1.0)
    devb = [torch.empty(n, device=f"cuda:{dev}") for _ in range(2)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]

    def copies(k):
        if variant == "chunk4":
            q = n // 4
            for c in range(4):
                with torch.cuda.stream(streams[c & 1]):
                    devb[k & 1][c * q:(c + 1) * q].copy_(host[k & 1][c * q:(c + 1) * q], non_blocking=True)
        else:
            with torch.cuda.stream(streams[0]):
                devb[k & 1].copy_(host[k & 1], non_blocking=True)

    for k in range(3):
        copies(k)
    torch.cuda.synchronize(dev)
    barrier.wait()
    t0 = time.perf_counter()
    for k in range(iters):
        copies(k)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    out[dev] = (iters * n * 4 / dt / 1e9, dt)
    barrier.wait()


def run_subset(gpus, mb, iters, variant):
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    out = mgr.dict()
    barrier = ctx.Barrier(len(gpus))
    procs = [ctx.Process(target=_worker, args=(r, gpus, mb, iters, variant, barrier, out)) for r in range(len(gpus))]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    per = {g: round(out[g][0], 2) for g in gpus if g in out}
    slowest = max(out[g][1] for g in gpus if g in out)
    agg = len(per) * iters * (mb * 1000 * 1000 // 4) * 4 / slowest / 1e9
    line = {"gpus": gpus, "variant": variant, "per_gpu_GBps": per, "aggregate_GBps": round(agg, 2),
            "slides_per_s_ceiling": round(agg * 1e9 / (10000 * 512 * 4), 1)}
    errs = {k: v for k, v in out.items() if isinstance(k, str)}
    if errs:
        line["notes"] = errs
    print(json.dumps(line), flush=True)
    return agg


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as exc:
        return f"<{exc}>"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=328)              # bench.py: 16 slides x 20.48 MB per step
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    n = torch.cuda.device_count()
    print(f"# {n} GPUs, {os.cpu_count()} CPUs visible, torch {torch.__version__}")
    print("# nvidia-smi topo -m\n" + sh("nvidia-smi topo -m"))
    print("# NUMA: " + sh("ls -d /sys/devices/system/node/node* | tr '\\n' ' '") + " | lscpu: " +
          sh("lscpu | grep -E 'Model name|Socket|NUMA|^CPU\\(s\\)' | tr -s ' ' | tr '\\n' ';'"))
    print("# MemTotal: " + sh("grep MemTotal /proc/meminfo"))
    ids = sh("nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv,noheader")
    print("# index, bus id, PCIe gen (current), width (current)\n" + ids)
    for line in ids.splitlines():
        try:
            bus = line.split(",")[1].strip().lower()
            bus = bus[4:] if bus.startswith("0000") and len(bus) > 12 else bus
            print(f"# sysfs path of GPU {line.split(',')[0]}: " + sh(f"readlink -f /sys/bus/pci/devices/{bus}"))
        except Exception:
            pass
    subsets = [[0]]
    if n >= 2:
        subsets += [[0, 1]]
    if n >= 4:
        subsets += [[0, 2], [0, 3], [0, 1, 2, 3]]
    if n >= 8:
        subsets += [[0, 4], [4, 5, 6, 7], [0, 2, 4, 6], list(range(8))]
    for g in subsets:
        run_subset(g, a.mb, a.iters, "plain")
    full = list(range(n))
    for variant in ("chunk4", "affinity"):
        run_subset(full, a.mb, a.iters, variant)
    if n >= 2:
        run_subset([0], a.mb, a.iters, "chunk4")


if __name__ == "__main__":
    main()
