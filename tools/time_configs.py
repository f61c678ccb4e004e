"""Forward time of the BASELINE.json configs other than the bench workload (parity-test cases, timed for DESIGN.md):
cfg1 (256x384, h=1, K=32), cfg3 (multiclass 6000x768, C=2, 4 layers), cfg4 extremes (N=1k/K=64 ... N=50k/K=1024)."""
import copy, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from snuffy_b200 import snuffy, snuffy_multiclass


def build(mod, d, h, K, r, depth, C, multiclass):
    i_cls = mod.FCLayer(d, C)
    attn = mod.MultiHeadedAttention(h, d)
    ff = mod.PositionwiseFeedForward(d, 4 * d, "relu", 0.0)
    if multiclass:
        layer = mod.EncoderLayer(d, copy.deepcopy(attn), copy.deepcopy(ff), C, 0.0, K, r)
    else:
        layer = mod.EncoderLayer(d, copy.deepcopy(attn), copy.deepcopy(ff), 0.0, K, r)
    m = mod.MILNet(i_cls, mod.BClassifier(mod.Encoder(layer, depth), C, d)).cuda().eval()
    for p in m.parameters():
        if p.dim() > 1:
            torch.nn.init.xavier_normal_(p)
    for l in m.b_classifier.encoder.layers:
        l.return_attn = False
    return m


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


cases = [("cfg1", snuffy, dict(d=384, h=1, K=32, r=0.0, depth=1, C=1, multiclass=False), 1, 256),
         ("cfg1 x64 bags", snuffy, dict(d=384, h=1, K=32, r=0.0, depth=1, C=1, multiclass=False), 64, 256),
         ("cfg2 r=0.5", snuffy, dict(d=512, h=8, K=200, r=0.5, depth=1, C=1, multiclass=False), 8, 10000),
         ("cfg3", snuffy_multiclass, dict(d=768, h=8, K=200, r=0.5, depth=4, C=2, multiclass=True), 1, 6000),
         ("cfg3 x4 bags", snuffy_multiclass, dict(d=768, h=8, K=200, r=0.5, depth=4, C=2, multiclass=True), 4, 6000),
         ("cfg4 N=1k K=64", snuffy, dict(d=512, h=8, K=64, r=0.5, depth=1, C=1, multiclass=False), 32, 1000),
         ("cfg4 N=50k K=1024", snuffy, dict(d=512, h=8, K=1024, r=0.5, depth=1, C=1, multiclass=False), 1, 50000)]
torch.manual_seed(0)
for name, mod, kw, B, n in cases:
    m = build(mod, **kw)
    x = torch.randn(B, n, kw["d"], device="cuda")
    with torch.no_grad():
        if mod is snuffy and B > 1:
            fn = lambda: snuffy.forward_bags(m, x)
        elif mod is snuffy:
            fn = lambda: m(x)
        else:
            fn = lambda: m(x)
        ms = timeit(fn)
    print(json.dumps({"config": name, "bags": B, "N": n, **{k: v for k, v in kw.items() if k != "multiclass"},
                      "ms_per_call": round(ms, 4), "slides_per_s": round(B / ms * 1e3, 1)}), flush=True)

# cfg4 proper: 64 bags with N ~ log-uniform[1k, 50k] (seed 7), K sweep, packed into one launch sequence vs one call per bag
import numpy as np
rs = np.random.RandomState(7)
lens = np.exp(rs.uniform(np.log(1000), np.log(50000), 64)).astype(int)
cu = np.concatenate([[0], np.cumsum(lens)])
xp = torch.randn(int(cu[-1]), 512, device="cuda")
for K in (64, 256, 1024):
    m = build(snuffy, d=512, h=8, K=K, r=0.5, depth=1, C=1, multiclass=False)
    with torch.no_grad():
        ok = lens >= K
        cu_ok = np.concatenate([[0], np.cumsum(lens[ok])])
        xs = torch.cat([xp[cu[i]:cu[i + 1]] for i in range(64) if ok[i]])
        ms_packed = timeit(lambda: snuffy.forward_packed(m, xs, cu_ok), iters=3)
        ms_loop = timeit(lambda: [m(xp[cu[i]:cu[i + 1]][None]) for i in range(64) if ok[i]], iters=2)
    print(json.dumps({"config": "cfg4 varlen 64 bags N~logU[1k,50k]", "K": K, "bags": int(ok.sum()), "rows": int(cu_ok[-1]),
                      "ms_packed": round(ms_packed, 3), "ms_one_call_per_bag": round(ms_loop, 3),
                      "slides_per_s_packed": round(ok.sum() / ms_packed * 1e3, 1)}), flush=True)

# dsmil: MILNet forward on a cfg2-shaped bag [10000, 512], C = 1, and the pooling kernel alone
from snuffy_b200 import dsmil, ops
for nonlinear, passing_v in ((True, False), (True, True)):
    dm = dsmil.MILNet(dsmil.FCLayer(512, 1), dsmil.BClassifier(512, 1, 0.0, nonlinear, passing_v)).cuda().eval()
    xd = torch.randn(10000, 512, device="cuda")
    with torch.no_grad():
        ms = timeit(lambda: dm(xd))
    print(json.dumps({"config": "dsmil MILNet", "N": 10000, "d": 512, "C": 1, "nonlinear": nonlinear, "passing_v": passing_v,
                      "ms_per_call": round(ms, 4), "slides_per_s": round(1e3 / ms, 1)}), flush=True)
q = torch.randn(10000, 128, device="cuda"); qm = torch.randn(1, 128, device="cuda"); v = torch.randn(10000, 512, device="cuda")
wf = torch.randn(1, 1, 512, device="cuda"); bf = torch.zeros(1, device="cuda")
ms = timeit(lambda: ops.dsmil_pool(q, qm, v, wf, bf), iters=50)
byts = 10000 * (128 + 512) * 4 + 10000 * 4
print(json.dumps({"config": "dsmil_pool kernel alone", "ms": round(ms, 4), "algorithmic_GBps": round(byts / ms / 1e6, 1)}), flush=True)
