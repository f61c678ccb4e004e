"""Data-parallel MIL training over slides (BASELINE.json configs[4]).

The reference trains one bag per optimizer step in a single process (train.py:249-264, 468-473).  Bags are independent,
so slides shard over ranks with NO data-path collective; the only exchange is ONE all-reduce (sum) of a single flat fp32
gradient buffer per optimizer step (NCCL over NVLink on the GPU box, gloo in the CPU tests), after which every rank
applies the same AdamW update.  World size 1 reproduces the reference trajectory (dropout 0, injected selections).

Everything numeric runs in libsnuffy_b200.so: forward / backward through the drop-in modules, the fused
max-instance + 2 x BCE loss (train.py:828-846), the gradient-norm clip and AdamW over the flat buffers (train.py:809-826,
469-470).  PyTorch provides device memory, the autograd sequencing and ``torch.distributed``.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


# ------------------------------------------------------------------ sharding (host logic, runs anywhere)
def shard_slides(num_slides: int, rank: int, world: int, lengths: Optional[Sequence[int]] = None) -> List[int]:
    """Slide ids owned by `rank`.  Without lengths: slide i -> rank i mod world.  With per-slide patch counts (cost is
    proportional to N): greedy longest-first bins, ties to the lower rank, each rank's list in ascending id order."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    if lengths is None:
        return list(range(rank, num_slides, world))
    if len(lengths) != num_slides:
        raise ValueError("lengths must have one entry per slide")
    load = [0] * world
    owner = [0] * num_slides
    for i in sorted(range(num_slides), key=lambda j: (-int(lengths[j]), j)):
        r = min(range(world), key=lambda q: (load[q], q))
        owner[i] = r
        load[r] += int(lengths[i])
    return [i for i in range(num_slides) if owner[i] == rank]


def steps_per_epoch(num_slides: int, world: int, bags_per_step: int = 1) -> int:
    """Every rank must run the same number of optimizer steps (the all-reduce is collective): ceil over the largest shard."""
    largest = (num_slides + world - 1) // world
    return (largest + bags_per_step - 1) // bags_per_step


# ------------------------------------------------------------------ flat parameter / gradient buffers
class FlatBuffers:
    """All parameters (and their gradients) of a module as views into two flat fp32 buffers, in
    ``named_parameters()`` order: L*(12 d^2 + 13 d) + 2 d + 2 (d C + C) floats (12.6 MB per layer at d = 512)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        sizes = [p.numel() for p in self.params]
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + (s + 3) // 4 * 4)          # keep every view 16-byte aligned
        total = self.offsets[-1]
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, off, s in zip(self.params, self.offsets, sizes):
            view = self.flat_param[off:off + s].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_grad[off:off + s].view(p.shape)                 # autograd accumulates in place

    @property
    def numel(self) -> int:
        return self.flat_param.numel()

    def zero_grad(self) -> None:
        self.flat_grad.zero_()
        for p, off in zip(self.params, self.offsets):                           # re-attach if something set .grad = None
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                p.grad = self.flat_grad[off:off + p.numel()].view(p.shape)

    def detach_grads(self) -> None:
        """Before backward: let autograd hand over each gradient tensor as is (no per-parameter `grad += g` kernel); `pack`
        then gathers them into the flat buffer with one launch."""
        for p in self.params:
            p.grad = None

    def pack(self) -> None:
        """After backward: flat_grad <- the parameters' .grad tensors (zeros where a parameter received none), then the
        .grad attributes become views of the flat buffer again."""
        import ctypes
        from . import _lib
        n = len(self.params)
        srcs, sizes, offs, keep = (ctypes.c_void_p * n)(), (ctypes.c_int64 * n)(), (ctypes.c_int64 * n)(), []
        missing = False
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            g = p.grad
            if g is None:
                missing = True
                srcs[i], sizes[i], offs[i] = None, 0, off
                continue
            if g.dtype != torch.float32 or not g.is_contiguous():
                g = g.float().contiguous()
            keep.append(g)
            srcs[i], sizes[i], offs[i] = g.data_ptr(), g.numel(), off
        if missing:
            self.flat_grad.zero_()
        _lib.check(_lib.lib.snuffy_pack_f32(srcs, sizes, offs, n, self.flat_grad.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "snuffy_pack_f32")
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat_grad[off:off + p.numel()].view(p.shape)

    def allreduce_sum(self, group=None) -> None:
        """THE collective of the path: one all-reduce of the flat gradient buffer."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=group)


# ------------------------------------------------------------------ fused loss (train.py:828-846)
class MilLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, classes, bag, label, w, weight):
        from . import _lib
        lib, check = _lib.lib, _lib.check
        if classes.dim() == 2:                                       # dsmil: [N, C] (train.py:831-832)
            classes3 = classes.unsqueeze(0)
        else:
            classes3 = classes
        classes3 = classes3.detach().contiguous()
        B, N, C = classes3.shape
        dev = classes3.device
        bag2 = bag.detach().contiguous().view(B, C)
        label2 = label.detach().to(device=dev, dtype=torch.float32).contiguous().view(B, C)
        terms = torch.empty(2 * B * C, dtype=torch.float32, device=dev)
        ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        loss = torch.empty(3, dtype=torch.float32, device=dev)
        pred = torch.empty(B, C, dtype=torch.float32, device=dev)
        dclasses = torch.zeros(B, N, C, dtype=torch.float32, device=dev)
        dbag = torch.empty(B, C, dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream().cuda_stream
        check(lib.snuffy_mil_loss(classes3.data_ptr(), bag2.data_ptr(), label2.data_ptr(),
                                  None if weight is None else weight.data_ptr(), B, N, C, float(w), 1.0, terms.data_ptr(),
                                  ticket.data_ptr(), loss.data_ptr(), pred.data_ptr(), dclasses.data_ptr(), dbag.data_ptr(),
                                  stream), "snuffy_mil_loss")
        ctx.save_for_backward(dclasses.view(classes.shape), dbag.view(bag.shape))
        ctx.mark_non_differentiable(pred)
        return loss[0], pred, loss[1:].detach()

    @staticmethod
    def backward(ctx, g, _gp, _gt):
        dclasses, dbag = ctx.saved_tensors
        return dclasses * g, dbag * g, None, None, None


def mil_loss(classes: torch.Tensor, bag: torch.Tensor, label: torch.Tensor, w: float = 0.5,
             weight: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(loss, mixed prediction [B, C], (bag term, max term)) = train.py:831-844 in one launch, differentiable."""
    return MilLossFunction.apply(classes, bag, label, w, weight)


# ------------------------------------------------------------------ AdamW on the flat buffers (train.py:809-826)
class FlatAdamW:
    def __init__(self, flat: FlatBuffers, lr: float = 2e-4, betas: Tuple[float, float] = (0.5, 0.9), eps: float = 1e-8,
                 weight_decay: float = 5e-3, clip_grad: Optional[float] = None):
        self.flat, self.lr, self.betas, self.eps, self.weight_decay, self.clip_grad = flat, lr, betas, eps, weight_decay, clip_grad
        self.exp_avg = torch.zeros_like(flat.flat_param)
        self.exp_avg_sq = torch.zeros_like(flat.flat_param)
        self.step_count = 0
        self._norm = None

    def step(self, grad_scale: float = 1.0) -> None:
        from . import _lib
        lib, check = _lib.lib, _lib.check
        f = self.flat
        stream = torch.cuda.current_stream().cuda_stream
        norm_ptr = None
        if self.clip_grad is not None:
            if self._norm is None:
                self._norm = (torch.empty(lib.snuffy_sumsq_blocks(f.numel), dtype=torch.float32, device=f.flat_grad.device),
                              torch.empty(1, dtype=torch.float32, device=f.flat_grad.device))
            check(lib.snuffy_sumsq(f.flat_grad.data_ptr(), f.numel, self._norm[0].data_ptr(), self._norm[1].data_ptr(), stream),
                  "snuffy_sumsq")
            norm_ptr = self._norm[1].data_ptr()
        self.step_count += 1
        check(lib.snuffy_adamw_flat(f.flat_param.data_ptr(), f.flat_grad.data_ptr(), self.exp_avg.data_ptr(),
                                    self.exp_avg_sq.data_ptr(), f.numel, self.lr, self.betas[0], self.betas[1], self.eps,
                                    self.weight_decay, self.step_count, float(grad_scale), norm_ptr,
                                    float(self.clip_grad or 0.0), stream), "snuffy_adamw_flat")


def invalidate_weight_caches(model: torch.nn.Module) -> None:
    """The optimizer kernel updates parameters through raw pointers (no autograd version bump): drop derived operands."""
    for m in model.modules():
        if hasattr(m, "_wcache"):
            m._wcache = None


# ------------------------------------------------------------------ the trainer
class DataParallelTrainer:
    """One process per GPU.  train_step(bags [B, N, d], labels [B, C]) = forward + fused loss + backward on this rank's
    bags, one all-reduce of the flat gradient, the same AdamW step on every rank.  Returns the local loss (device tensor,
    no host sync; the reference's three `.item()` / `.cpu()` syncs per bag are the caller's choice here).
    cuda_graph=True replays forward + loss + backward + gradient packing as one captured graph per bag shape (dropout masks and
    random patches are drawn from a device-side step counter, so replays differ); the returned loss tensor is then a static
    buffer overwritten by the next step."""

    def __init__(self, model: torch.nn.Module, lr: float = 2e-4, betas=(0.5, 0.9), weight_decay: float = 5e-3,
                 clip_grad: Optional[float] = None, mix_weight: float = 0.5, group=None,
                 forward_fn: Optional[Callable] = None, class_weight: Optional[torch.Tensor] = None,
                 cuda_graph: bool = False):
        self.model = model
        self.cuda_graph = bool(cuda_graph)
        self._graph, self._graph_key = None, None
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.flat = FlatBuffers(model.parameters())
        if self.world > 1:                                            # identical start on every rank
            dist.broadcast(self.flat.flat_param, src=0, group=group)
        self.opt = FlatAdamW(self.flat, lr=lr, betas=betas, weight_decay=weight_decay, clip_grad=clip_grad)
        self.mix_weight = mix_weight
        self.class_weight = class_weight
        if forward_fn is None:
            from . import snuffy
            binary = any(isinstance(m, snuffy.EncoderLayer) for m in model.modules())
            forward_fn = (lambda x: snuffy.forward_bags(model, x)) if binary else model
        self.forward_fn = forward_fn
        self._cached_layers = [m for m in model.modules() if hasattr(m, "_wcache")]
        invalidate_weight_caches(model)

    def _forward_backward(self, bags: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        self.flat.detach_grads()
        classes, bag, _ = self.forward_fn(bags)
        loss, _, _ = mil_loss(classes, bag, labels, self.mix_weight, self.class_weight)
        loss.backward()
        self.flat.pack()                                              # one launch instead of one `grad += g` per parameter
        return loss.detach()

    def train_step(self, bags: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        if not self.model.training:
            self.model.train()
        if self.cuda_graph:
            loss = self._replay(bags, labels)
        else:
            loss = self._forward_backward(bags, labels)
        self.flat.allreduce_sum(self.group)
        self.opt.step(grad_scale=1.0 / self.world)
        for m in self._cached_layers:                                 # the optimizer kernel bypasses autograd's version counters
            m._wcache = None
        return loss

    # ---- cuda_graph=True: forward + loss + backward + gradient packing of one step are ONE graph launch (the eager step is
    # bound by ~100 host-side launches per bag); the all-reduce and the optimizer kernel stay eager behind it.
    def _capture(self, bags: torch.Tensor, labels: torch.Tensor) -> None:
        from . import _lib, engine
        dev = bags.device
        self._static_bags, self._static_labels = bags.clone(), labels.to(device=dev, dtype=torch.float32).clone()
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):                                 # warm-up off the capture: allocator, kernel attributes
            for _ in range(2):
                self._forward_backward(self._static_bags, self._static_labels)
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        for m in self._cached_layers:                                 # derived operands are rebuilt INSIDE the graph, from the
            m._wcache = None                                          # parameters as they are at each replay
        engine._RANDOM.next()                                         # settle the eager stream's (seed, offset) first
        self._rng_counter = torch.tensor([engine._RANDOM._offset], dtype=torch.int64, device=dev)
        graph = torch.cuda.CUDAGraph()
        launches0 = _lib.lib.snuffy_launch_count()
        engine._RANDOM.begin_indirect(self._rng_counter)
        try:
            with torch.cuda.graph(graph):
                self._static_loss = self._forward_backward(self._static_bags, self._static_labels)
                self._draws_per_step = engine._RANDOM.end_indirect()
                _lib.check(_lib.lib.snuffy_rng_advance(self._rng_counter.data_ptr(), self._draws_per_step,
                                                       torch.cuda.current_stream(dev).cuda_stream), "snuffy_rng_advance")
        finally:
            engine._RANDOM.end_indirect()
        for m in self._cached_layers:
            m._wcache = None
        self._graph_kernels = int(_lib.lib.snuffy_launch_count() - launches0)       # library kernels per replay
        self._graph, self._graph_key = graph, (tuple(bags.shape), tuple(labels.shape))

    def _replay(self, bags: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        if self._graph is None or self._graph_key != (tuple(bags.shape), tuple(labels.shape)):
            self._capture(bags, labels)                               # first step, or a new bag shape: (re)capture
        self._static_bags.copy_(bags, non_blocking=True)
        self._static_labels.copy_(labels, non_blocking=True)
        self._graph.replay()
        return self._static_loss

    @torch.no_grad()
    def predict(self, bags: torch.Tensor) -> torch.Tensor:
        """Mixed prediction of train.py:840-844 for evaluation (eval mode, no gradient), from the same fused kernel."""
        self.model.eval()
        classes, bag, _ = self.forward_fn(bags)
        _, pred, _ = mil_loss(classes, bag, torch.zeros_like(bag), self.mix_weight, self.class_weight)
        return pred

    def validate(self, bags, collector=None):
        """The reference's `valid` loop (train.py:334-355) without its per-bag host syncs: `bags` yields
        (slide id, bag [1, N, d] on the device, label [1, C] on the device) — e.g. a store.PinnedPrefetcher.  Per bag:
        eval forward, the fused loss kernel, and (if a patch_outputs.PatchOutputCollector is given) the patch probabilities
        and the mixed prediction appended to its epoch buffers.  Returns (mean loss [device scalar], slide ids); nothing
        is copied to the host here."""
        self.model.eval()
        total, ids = None, []
        with torch.no_grad():
            for sid, bag_x, label in bags:
                classes, bag, _ = self.forward_fn(bag_x)
                loss, pred, _ = mil_loss(classes, bag, label, self.mix_weight, self.class_weight)
                if collector is not None:
                    collector.add(classes, pred)
                total = loss.clone() if total is None else total.add_(loss)
                ids.append(sid)
        if total is None:
            raise ValueError("validate: no bags")
        return total / len(ids), ids
