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

import os
import socket
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


def steps_per_epoch(num_slides: int, world: int, bags_per_step: int = 1, lengths: Optional[Sequence[int]] = None) -> int:
    """Every rank must run the same number of optimizer steps (the all-reduce is collective): the step count of the LARGEST
    shard of `shard_slides` with the same arguments (length-balanced bins can differ a lot in slide count).  Ranks that run
    out of slides join the remaining steps with `DataParallelTrainer.train_step(None, None)` (an idle step)."""
    if lengths is None:
        largest = (num_slides + world - 1) // world
    else:
        largest = max((len(shard_slides(num_slides, r, world, lengths)) for r in range(world)), default=0)
    return (largest + bags_per_step - 1) // bags_per_step


# ------------------------------------------------------------------ flat parameter / gradient buffers
class _RawCudaFloats:
    """Zero-copy view of device memory this package allocated itself (torch.as_tensor reads __cuda_array_interface__)."""

    def __init__(self, ptr: int, numel: int, owner):
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        self._owner = owner


class PeerAllReduce:
    """The gradient all-reduce of the data-parallel step as ONE kernel over NVLink peer memory (csrc/comm.cu): every rank's
    gradient buffer lives in a cudaMalloc'd block its peers map through CUDA IPC; each rank sums its slice of all buffers in
    rank order and writes the sum back into all of them (two-shot), with per-CTA barriers on counters in the same blocks.
    One node, 2..8 ranks, one GPU per rank.  `setup` is collective and returns None on EVERY rank if any rank cannot map its
    peers (the trainer then keeps the NCCL all-reduce)."""

    def __init__(self):
        self.ptr = None

    @classmethod
    def setup(cls, numel: int, device: torch.device, group=None):
        import ctypes
        from . import _lib
        lib = _lib.lib
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        self = cls()
        self.world, self.rank, self.numel, self.group = world, rank, (numel + 3) // 4 * 4, group
        ok, err = 2 <= world <= 8 and device.type == "cuda" and os.environ.get("SNUFFY_B200_PEER_ALLREDUCE", "1") != "0", ""
        counter_bytes = lib.snuffy_comm_counter_bytes()
        self._data_bytes = (self.numel * 4 + 255) // 256 * 256
        total = self._data_bytes + counter_bytes + 256
        handle = b""
        if ok:
            try:
                with torch.cuda.device(device):
                    base = ctypes.c_void_p()
                    _lib.check(lib.snuffy_comm_alloc(total, ctypes.byref(base)), "snuffy_comm_alloc")
                    self.ptr = base.value
                    hb = ctypes.create_string_buffer(lib.snuffy_comm_handle_bytes())
                    _lib.check(lib.snuffy_comm_export(self.ptr, hb), "snuffy_comm_export")
                    handle = hb.raw
            except Exception as exc:                                   # noqa: BLE001 - any failure means "no peer path"
                ok, err = False, repr(exc)
        gathered = [None] * world
        dist.all_gather_object(gathered, (ok, handle, socket.gethostname(), device.index), group=group)
        ok = all(g[0] for g in gathered) and len({g[2] for g in gathered}) == 1 and len({g[3] for g in gathered}) == world
        self._peers = []
        bases = [0] * world
        if ok:
            try:
                with torch.cuda.device(device):
                    for q, g in enumerate(gathered):
                        if q == rank:
                            bases[q] = self.ptr
                            continue
                        mapped = ctypes.c_void_p()
                        _lib.check(lib.snuffy_comm_import(g[1], ctypes.byref(mapped)), "snuffy_comm_import")
                        self._peers.append(mapped.value)
                        bases[q] = mapped.value
            except Exception as exc:                                   # noqa: BLE001
                ok, err = False, repr(exc)
        flags = [None] * world
        dist.all_gather_object(flags, ok, group=group)
        if not all(flags):
            self.close()
            if err and rank == 0:
                print(f"snuffy_b200.dp: peer-memory all-reduce unavailable ({err}); using the NCCL all-reduce", flush=True)
            return None
        self._bufs = (ctypes.c_void_p * world)(*bases)
        self._counters = (ctypes.c_void_p * world)(*[b + self._data_bytes for b in bases])
        self._state = self.ptr + self._data_bytes + counter_bytes
        self.device = device
        self.buffer = torch.as_tensor(_RawCudaFloats(self.ptr, self.numel, self), device=device)
        return self

    def all_reduce(self) -> None:
        from . import _lib
        _lib.check(_lib.lib.snuffy_peer_allreduce(self._bufs, self._counters, self._state, self.rank, self.world, self.numel,
                                                  torch.cuda.current_stream(self.device).cuda_stream), "snuffy_peer_allreduce")

    def __del__(self):
        try:
            self.close()
        except Exception:                                             # noqa: BLE001 - interpreter shutdown
            pass

    def close(self) -> None:
        """Unmap the peers' blocks and free this rank's (call only when no step is in flight: after a synchronize)."""
        from . import _lib
        for m in getattr(self, "_peers", []):
            _lib.lib.snuffy_comm_close(m)
        self._peers = []
        if self.ptr:
            _lib.lib.snuffy_comm_free(self.ptr)
            self.ptr = None


class FlatBuffers:
    """All parameters (and their gradients) of a module as views into two flat fp32 buffers, in
    ``named_parameters()`` order: L*(12 d^2 + 13 d) + 2 d + 2 (d C + C) floats (12.6 MB per layer at d = 512).

    `extra`: further parameters appended after the module's (the learnable loss mix weight, train.py:804,820) — they form
    their own optimizer group.  One more slot behind everything (gradient buffer only) carries the CONTRIBUTOR COUNT of a
    step: 1 on a rank that had a bag, 0 on an idle rank; the all-reduce sums it with the gradient, and the optimizer kernel
    divides by it (uneven shards: `steps_per_epoch`)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], extra: Iterable[torch.Tensor] = (), peer_group=None):
        self.peer = None
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        self.n_model_params = len(self.params)
        self.params += [p for p in extra if p.requires_grad]
        dev = self.params[0].device
        sizes = [p.numel() for p in self.params]
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + (s + 3) // 4 * 4)          # keep every view 16-byte aligned
        total = self.offsets[-1]
        self.model_numel = self.offsets[self.n_model_params]                   # the module's parameters: [0, model_numel)
        self._total = total
        self._param_all = torch.zeros(total + 4, dtype=torch.float32, device=dev)
        if peer_group is not False and dev.type == "cuda" and dist.is_available() and dist.is_initialized() \
                and dist.get_world_size(peer_group) > 1:
            self.peer = PeerAllReduce.setup(total + 4, dev, peer_group)           # the gradient buffer lives in peer-mapped memory
        self._grad_all = self.peer.buffer if self.peer is not None else torch.zeros(total + 4, dtype=torch.float32, device=dev)
        self.flat_param = self._param_all[:total]
        self.flat_grad = self._grad_all[:total]
        self.contributors = self._grad_all[total:total + 1]                    # summed by the all-reduce
        self._one = torch.ones(1, dtype=torch.float32, device=dev)
        for p, off, s in zip(self.params, self.offsets, sizes):
            view = self.flat_param[off:off + s].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_grad[off:off + s].view(p.shape)                 # autograd accumulates in place

    @property
    def numel(self) -> int:
        return self._total

    def zero_grad(self) -> None:
        self._grad_all.zero_()
        for p, off in zip(self.params, self.offsets):                           # re-attach if something set .grad = None
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                p.grad = self.flat_grad[off:off + p.numel()].view(p.shape)

    def detach_grads(self) -> None:
        """Before backward: let autograd hand over each gradient tensor as is (no per-parameter `grad += g` kernel); `pack`
        then gathers them into the flat buffer with one launch."""
        for p in self.params:
            p.grad = None

    def pack(self) -> None:
        """After backward: flat_grad <- the parameters' .grad tensors (zeros where a parameter received none) and the
        contributor slot <- 1, then the .grad attributes become views of the flat buffer again."""
        import ctypes
        from . import _lib
        n = len(self.params) + 1
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
        srcs[n - 1], sizes[n - 1], offs[n - 1] = self._one.data_ptr(), 1, self._total
        if missing:
            self.flat_grad.zero_()
        _lib.check(_lib.lib.snuffy_pack_f32(srcs, sizes, offs, n, self._grad_all.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "snuffy_pack_f32")
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat_grad[off:off + p.numel()].view(p.shape)

    def allreduce_sum(self, group=None) -> None:
        """THE collective of the path: one all-reduce of the flat gradient buffer (and its contributor slot)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if self.peer is not None:
                self.peer.all_reduce()                                          # one kernel over NVLink peer memory (csrc/comm.cu)
            else:
                dist.all_reduce(self._grad_all, op=dist.ReduceOp.SUM, group=group)


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
        # the mix weight is a Python float, or a 0-dim CUDA tensor read on the device (learnable with --soft_average)
        w_is_tensor = torch.is_tensor(w)
        if w_is_tensor and (not w.is_cuda or w.dtype != torch.float32 or w.numel() != 1):
            raise ValueError("mil_loss: a tensor mix weight must be a float32 CUDA tensor with one element")
        dw = torch.empty(1, dtype=torch.float32, device=dev) if w_is_tensor else None
        stream = torch.cuda.current_stream().cuda_stream
        check(lib.snuffy_mil_loss(classes3.data_ptr(), bag2.data_ptr(), label2.data_ptr(),
                                  None if weight is None else weight.data_ptr(), B, N, C, 0.0 if w_is_tensor else float(w),
                                  w.data_ptr() if w_is_tensor else None, 1.0, terms.data_ptr(), ticket.data_ptr(),
                                  loss.data_ptr(), pred.data_ptr(), dclasses.data_ptr(), dbag.data_ptr(),
                                  None if dw is None else dw.data_ptr(), stream), "snuffy_mil_loss")
        ctx.save_for_backward(dclasses.view(classes.shape), dbag.view(bag.shape), dw)
        ctx.w_shape = w.shape if w_is_tensor else None
        ctx.mark_non_differentiable(pred)
        return loss[0], pred, loss[1:].detach()

    @staticmethod
    def backward(ctx, g, _gp, _gt):
        dclasses, dbag, dw = ctx.saved_tensors
        gw = (dw * g).view(ctx.w_shape) if dw is not None and ctx.needs_input_grad[3] else None
        return dclasses * g, dbag * g, None, gw, None


def mil_loss(classes: torch.Tensor, bag: torch.Tensor, label: torch.Tensor, w=0.5,
             weight: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(loss, mixed prediction [B, C], (bag term, max term)) = train.py:831-844 in one launch, differentiable.  `w` is the
    `single_weight_parameter` of train.py:804: a float, or a 0-dim CUDA tensor (its gradient is bag term - max term)."""
    return MilLossFunction.apply(classes, bag, label, w, weight)


# ------------------------------------------------------------------ AdamW on the flat buffers (train.py:809-826)
class FlatAdamW:
    """torch.optim.AdamW over the flat buffers with the reference's two parameter groups (train.py:817-825): the module's
    parameters at `lr`, and the appended extra parameters (the loss mix weight) at `lr * extra_lr_multiplier`, clamped to
    [0, 1] after every step (train.py:852-854).  The step count lives on the device, so `step()` can be captured in a CUDA
    graph; `set_lr` serves schedulers (train.py:182-197) without re-capturing.  The global-norm clip covers the module's
    parameters only, like train.py:469-470."""

    def __init__(self, flat: FlatBuffers, lr: float = 2e-4, betas: Tuple[float, float] = (0.5, 0.9), eps: float = 1e-8,
                 weight_decay: float = 5e-3, clip_grad: Optional[float] = None, extra_lr_multiplier: float = 0.1,
                 extra_clamp: Optional[Tuple[float, float]] = (0.0, 1.0)):
        self.flat, self.lr, self.betas, self.eps, self.weight_decay, self.clip_grad = flat, lr, betas, eps, weight_decay, clip_grad
        self.extra_lr_multiplier, self.extra_clamp = extra_lr_multiplier, extra_clamp
        dev = flat.flat_param.device
        self.exp_avg = torch.zeros_like(flat.flat_param)
        self.exp_avg_sq = torch.zeros_like(flat.flat_param)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.lr_dev = torch.tensor([lr, lr * extra_lr_multiplier], dtype=torch.float32, device=dev)
        self._norm = None

    @property
    def step_count(self) -> int:
        return int(self.step_dev.item())

    def set_lr(self, lr: float) -> None:
        self.lr = lr
        self.lr_dev.copy_(torch.tensor([lr, lr * self.extra_lr_multiplier], dtype=torch.float32), non_blocking=True)

    def step(self, grad_scale: float = 1.0, use_contributors: bool = False) -> None:
        from . import _lib
        lib, check = _lib.lib, _lib.check
        f = self.flat
        stream = torch.cuda.current_stream().cuda_stream
        norm_ptr = None
        n_model = f.model_numel
        if self.clip_grad is not None:
            if self._norm is None:
                self._norm = (torch.empty(lib.snuffy_sumsq_blocks(n_model), dtype=torch.float32, device=f.flat_grad.device),
                              torch.empty(1, dtype=torch.float32, device=f.flat_grad.device))
            check(lib.snuffy_sumsq(f.flat_grad.data_ptr(), n_model, self._norm[0].data_ptr(), self._norm[1].data_ptr(), stream),
                  "snuffy_sumsq")
            norm_ptr = self._norm[1].data_ptr()
        check(lib.snuffy_rng_advance(self.step_dev.data_ptr(), 1, stream), "snuffy_rng_advance")      # step += 1
        contrib = f.contributors.data_ptr() if use_contributors else None
        groups = [(0, n_model, 0, norm_ptr, 1.0, 0.0)]
        if f.numel > n_model:
            lo, hi = self.extra_clamp if self.extra_clamp is not None else (1.0, 0.0)
            groups.append((n_model, f.numel - n_model, 1, None, lo, hi))
        for off, n, g, nptr, lo, hi in groups:
            check(lib.snuffy_adamw_flat_dev(f.flat_param.data_ptr() + 4 * off, f.flat_grad.data_ptr() + 4 * off,
                                            self.exp_avg.data_ptr() + 4 * off, self.exp_avg_sq.data_ptr() + 4 * off, n,
                                            self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                            self.step_dev.data_ptr(), contrib, self.lr_dev.data_ptr() + 4 * g,
                                            float(grad_scale), nptr, float(self.clip_grad or 0.0), lo, hi, stream),
                  "snuffy_adamw_flat_dev")


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

    train_step(None, None) is an IDLE step for a rank whose shard has run out (uneven shards, `steps_per_epoch`): it
    contributes a zero gradient and a contributor count of 0, joins the all-reduce and applies the same update.

    soft_average=True makes the loss mix weight a learnable parameter in its own optimizer group (lr x
    single_weight_lr_multiplier, clamped to [0, 1] after every step: train.py:804, 817-825, 852-854); it is
    `trainer.single_weight_parameter` like in the reference.

    cuda_graph=True replays the whole step — forward, loss, backward, gradient packing, the all-reduce and the optimizer — as
    ONE captured graph per bag shape (dropout masks and random patches are drawn from a device-side step counter, so
    replays differ; the optimizer's step count and learning rate live on the device); the returned loss tensor is then a
    static buffer overwritten by the next step.  If the collective cannot be captured on this build, the step falls back to
    graph(forward .. packing) + eager all-reduce + graph-free optimizer launch (`trainer.graph_mode` says which)."""

    def __init__(self, model: torch.nn.Module, lr: float = 2e-4, betas=(0.5, 0.9), weight_decay: float = 5e-3,
                 clip_grad: Optional[float] = None, mix_weight: float = 0.5, group=None,
                 forward_fn: Optional[Callable] = None, class_weight: Optional[torch.Tensor] = None,
                 cuda_graph: bool = False, soft_average: bool = False, single_weight_lr_multiplier: float = 0.1,
                 data_parallel: bool = True):
        self.model = model
        self.cuda_graph = bool(cuda_graph)
        self._graph, self._graph_key, self.graph_mode = None, None, None
        self.group = group
        # data_parallel=False: a purely local trainer even inside an initialised process group (no broadcast, no all-reduce)
        self.world = dist.get_world_size(group) if data_parallel and dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        dev = next(model.parameters()).device
        # train.py:804: a 0-dim tensor, clamped to [0, 1], trainable only with --soft_average
        self.single_weight_parameter = torch.tensor(float(mix_weight), device=dev).clamp_(0, 1).requires_grad_(bool(soft_average))
        self.flat = FlatBuffers(model.parameters(), extra=[self.single_weight_parameter],
                                peer_group=group if self.world > 1 else False)
        if self.world > 1:                                            # identical start on every rank
            dist.broadcast(self.flat.flat_param, src=0, group=group)
        self.opt = FlatAdamW(self.flat, lr=lr, betas=betas, weight_decay=weight_decay, clip_grad=clip_grad,
                             extra_lr_multiplier=single_weight_lr_multiplier)
        self.class_weight = class_weight
        if forward_fn is None:
            from . import snuffy
            binary = any(isinstance(m, snuffy.EncoderLayer) for m in model.modules())
            forward_fn = (lambda x: snuffy.forward_bags(model, x)) if binary else model
        self.forward_fn = forward_fn
        self._cached_layers = [m for m in model.modules() if hasattr(m, "_wcache")]
        self.allreduce_events = None                                  # (start, end) CUDA events of the last eager all-reduce
        invalidate_weight_caches(model)

    @property
    def mix_weight(self):
        """The loss mix weight w (train.py:836-838): the device tensor when it is learnable, else its value."""
        return self.single_weight_parameter

    def _forward_backward(self, bags: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        self.flat.detach_grads()
        classes, bag, _ = self.forward_fn(bags)
        loss, _, _ = mil_loss(classes, bag, labels, self.single_weight_parameter, self.class_weight)
        loss.backward()
        self.flat.pack()                                              # one launch instead of one `grad += g` per parameter
        return loss.detach()

    def _reduce_and_update(self, timed: bool = False) -> None:
        if timed and self.world > 1:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
            self.flat.allreduce_sum(self.group)
            ev[1].record()
            self.allreduce_events = ev
        elif self.world > 1:
            self.flat.allreduce_sum(self.group)
        self.opt.step(grad_scale=1.0, use_contributors=True)          # the kernel divides by the contributor count

    def train_step(self, bags: Optional[torch.Tensor], labels: Optional[torch.Tensor], time_allreduce: bool = False) -> Optional[torch.Tensor]:
        if not self.model.training:
            self.model.train()
        if bags is None:                                              # idle step: this rank's shard has run out
            self.flat.zero_grad()
            self._reduce_and_update()
            loss = None
        elif self.cuda_graph:
            loss = self._replay(bags, labels)
        else:
            loss = self._forward_backward(bags, labels)
            self._reduce_and_update(time_allreduce)
        for m in self._cached_layers:                                 # the optimizer kernel bypasses autograd's version counters
            m._wcache = None
        return loss

    # ---- cuda_graph=True: the eager step is bound by ~100 host-side launches per bag; a replay is one launch.
    def _capture(self, bags: torch.Tensor, labels: torch.Tensor) -> None:
        from . import _lib, engine
        dev = bags.device
        self._static_bags, self._static_labels = bags.clone(), labels.to(device=dev, dtype=torch.float32).clone()
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        snapshot = (self.flat._param_all.clone(), self.opt.exp_avg.clone(), self.opt.exp_avg_sq.clone(), self.opt.step_dev.clone())
        with torch.cuda.stream(side):                                 # warm-up off the capture: allocator, kernel attributes, NCCL
            for _ in range(2):
                self._forward_backward(self._static_bags, self._static_labels)
                self._reduce_and_update()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        # the warm-up steps must not move the model: restore parameters and optimizer state
        self.flat._param_all.copy_(snapshot[0]); self.opt.exp_avg.copy_(snapshot[1]); self.opt.exp_avg_sq.copy_(snapshot[2])
        self.opt.step_dev.copy_(snapshot[3])
        del snapshot
        engine._RANDOM.next()                                         # settle the eager stream's (seed, offset) first
        self._rng_counter = torch.tensor([engine._RANDOM._offset], dtype=torch.int64, device=dev)

        def capture(whole_step: bool):
            for m in self._cached_layers:                             # derived operands are rebuilt INSIDE the graph, from the
                m._wcache = None                                      # parameters as they are at each replay
            graph = torch.cuda.CUDAGraph()
            launches0 = _lib.lib.snuffy_launch_count()
            engine._RANDOM.begin_indirect(self._rng_counter)
            try:
                with torch.cuda.graph(graph):
                    self._static_loss = self._forward_backward(self._static_bags, self._static_labels)
                    self._draws_per_step = engine._RANDOM.end_indirect()
                    _lib.check(_lib.lib.snuffy_rng_advance(self._rng_counter.data_ptr(), self._draws_per_step,
                                                           torch.cuda.current_stream(dev).cuda_stream), "snuffy_rng_advance")
                    if whole_step:
                        self._reduce_and_update()
            finally:
                engine._RANDOM.end_indirect()
            self._graph_kernels = int(_lib.lib.snuffy_launch_count() - launches0)   # library kernels per replay
            return graph

        try:
            graph, self.graph_mode = capture(True), "whole step (forward .. all-reduce .. AdamW) in one graph"
        except Exception as exc:                                      # e.g. a collective backend that cannot be captured
            if self.world == 1:
                raise
            torch.cuda.synchronize(dev)
            graph = capture(False)
            self.graph_mode = f"forward .. packing in one graph, all-reduce + AdamW eager ({type(exc).__name__})"
        for m in self._cached_layers:
            m._wcache = None
        self._graph, self._graph_key = graph, (tuple(bags.shape), tuple(labels.shape))

    def _replay(self, bags: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        if self._graph is None or self._graph_key != (tuple(bags.shape), tuple(labels.shape)):
            self._capture(bags, labels)                               # first step, or a new bag shape: (re)capture
        self._static_bags.copy_(bags, non_blocking=True)
        self._static_labels.copy_(labels, non_blocking=True)
        self._graph.replay()
        if not self.graph_mode.startswith("whole"):
            self._reduce_and_update()
        return self._static_loss

    @torch.no_grad()
    def predict(self, bags: torch.Tensor) -> torch.Tensor:
        """Mixed prediction of train.py:840-844 for evaluation (eval mode, no gradient), from the same fused kernel."""
        self.model.eval()
        classes, bag, _ = self.forward_fn(bags)
        _, pred, _ = mil_loss(classes, bag, torch.zeros_like(bag), self.single_weight_parameter.detach(), self.class_weight)
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
                loss, pred, _ = mil_loss(classes, bag, label, self.single_weight_parameter.detach(), self.class_weight)
                if collector is not None:
                    collector.add(classes, pred)
                total = loss.clone() if total is None else total.add_(loss)
                ids.append(sid)
        if total is None:
            raise ValueError("validate: no bags")
        return total / len(ids), ids


def train_epoch(trainer: DataParallelTrainer, slides: Sequence, load: Callable, num_slides: Optional[int] = None,
                lengths: Optional[Sequence[int]] = None) -> int:
    """One epoch of the data-parallel loop (the reference's train.py:249-264 with slides sharded over ranks): `slides` is this
    rank's shard (from `shard_slides`), `load(slide_id) -> (bag [1, N, d], label [1, C])` on the device.  Every rank runs
    `steps_per_epoch` optimizer steps; a rank whose shard is shorter joins the tail with idle steps.  Returns the steps run."""
    if num_slides is None:
        steps = len(slides)
        if trainer.world > 1:                                         # agree on the longest shard
            t = torch.tensor([steps], device=trainer.flat.flat_param.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=trainer.group)
            steps = int(t.item())
    else:
        steps = steps_per_epoch(num_slides, trainer.world, 1, lengths)
    for i in range(steps):
        if i < len(slides):
            trainer.train_step(*load(slides[i]))
        else:
            trainer.train_step(None, None)
    return steps
