"""Per-bag latency of the drop-in module call (reference semantics: one bag per call), eager vs CUDA graph, cfg2."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda", 0)
model, _ = bench.build_model(dev)
for l in model.b_classifier.encoder.layers:
    l.return_attn = False
x = torch.randn(1, 10000, 512, device=dev)
with torch.no_grad():
    for _ in range(5):
        model(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        model(x)
    host = (time.perf_counter() - t0) / 50
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        model(x); s.synchronize()
        with torch.cuda.graph(g, stream=s):
            out = model(x)
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    e0.record()
    for _ in range(50):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    graph_ms = e0.elapsed_time(e1) / 50
print(json.dumps({"cfg2 single bag": {"eager_host_issue_ms": round(host * 1e3, 3), "eager_wall_ms": round(wall * 1e3, 3),
                                       "cuda_graph_ms": round(graph_ms, 3)}}))
