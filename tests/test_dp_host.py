"""Host-side logic of the data-parallel path on the CPU: slide sharding, the flat gradient buffer and
its single all-reduce over a world_size-2 gloo group.  No kernels run here (the compute path is CUDA-only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from snuffy_b200 import dp


def test_shard_slides_round_robin_covers_every_slide_once():
    for n, w in [(512, 8), (10, 4), (3, 8), (0, 2)]:
        owned = [dp.shard_slides(n, r, w) for r in range(w)]
        assert sorted(i for o in owned for i in o) == list(range(n))
        assert all(o == list(range(r, n, w)) for r, o in enumerate(owned))
        assert dp.steps_per_epoch(n, w) == max([len(o) for o in owned] + [0])
    with pytest.raises(ValueError):
        dp.shard_slides(4, 5, 4)


def test_steps_per_epoch_follows_the_largest_shard():
    """Length-balanced bins can differ a lot in slide count: every rank runs as many steps as the LARGEST shard has slides
    (the shorter ones join with idle steps), so no slide is dropped and no rank misses a collective."""
    lengths = [100, 1, 1, 1, 1, 1]
    shards = [dp.shard_slides(6, r, 2, lengths) for r in range(2)]
    assert sorted(len(s_) for s_ in shards) == [1, 5]
    assert dp.steps_per_epoch(6, 2, lengths=lengths) == 5
    assert dp.steps_per_epoch(6, 2) == 3 and dp.steps_per_epoch(7, 2) == 4 and dp.steps_per_epoch(0, 4) == 0
    assert dp.steps_per_epoch(7, 2, bags_per_step=2) == 2
    rs = np.random.RandomState(1)
    for _ in range(20):
        n, w = int(rs.randint(1, 40)), int(rs.randint(1, 9))
        ln = rs.randint(1, 1000, n)
        assert dp.steps_per_epoch(n, w, lengths=ln) == max(len(dp.shard_slides(n, r, w, ln)) for r in range(w))


def test_shard_slides_length_balanced():
    rs = np.random.RandomState(7)
    lengths = np.exp(rs.uniform(np.log(1000), np.log(50000), 64)).astype(int)      # cfg4: N ~ log-uniform[1k, 50k]
    owned = [dp.shard_slides(64, r, 8, lengths) for r in range(8)]
    assert sorted(i for o in owned for i in o) == list(range(64))
    loads = np.array([lengths[o].sum() for o in owned])
    assert loads.max() - loads.min() <= lengths.max()                              # greedy LPT bound
    assert loads.max() < 1.25 * lengths.sum() / 8
    rr = np.array([lengths[dp.shard_slides(64, r, 8)].sum() for r in range(8)])
    assert loads.max() <= rr.max()


def _tiny_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.LayerNorm(5), torch.nn.Linear(5, 1))


def test_flat_buffers_alias_parameters_and_gradients():
    model = _tiny_model()
    before = [p.detach().clone() for p in model.parameters()]
    flat = dp.FlatBuffers(model.parameters())
    assert flat.numel >= sum(p.numel() for p in model.parameters()) and flat.numel % 4 == 0
    for p, b, off in zip(model.parameters(), before, flat.offsets):
        assert torch.equal(p.detach(), b)
        assert p.data_ptr() == flat.flat_param.data_ptr() + 4 * off and off % 4 == 0
    model(torch.randn(3, 6)).sum().backward()                                      # autograd accumulates into the views
    assert flat.flat_grad.abs().sum() > 0
    for p, off in zip(model.parameters(), flat.offsets):
        assert p.grad.data_ptr() == flat.flat_grad.data_ptr() + 4 * off
    flat.flat_param.mul_(2.0)                                                      # an in-place optimizer kernel
    for p, b in zip(model.parameters(), before):
        assert torch.allclose(p.detach(), 2 * b)
    flat.zero_grad()
    assert flat.flat_grad.abs().sum() == 0 and all(p.grad is not None for p in model.parameters())


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _tiny_model()
        if rank == 1:                                                              # ranks start different ...
            for p in model.parameters():
                p.data.add_(1.0)
        flat = dp.FlatBuffers(model.parameters())
        dist.broadcast(flat.flat_param, src=0)                                     # ... the trainer broadcasts rank 0
        x = torch.full((2, 6), float(rank + 1))
        model(x).sum().backward()
        local = flat.flat_grad.clone()
        flat.contributors.fill_(1.0 if rank == 0 else 0.0)                         # rank 1 plays an idle rank
        flat.allreduce_sum()
        contributors = float(flat.contributors)
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        ok = torch.allclose(flat.flat_grad, sum(gathered)) and flat.flat_grad.abs().sum() > 0
        ref = _tiny_model()
        same_start = all(torch.equal(p.detach(), q.detach()) for p, q in zip(model.parameters(), ref.parameters()))
        mine = dp.shard_slides(9, rank, world)
        out[rank] = (bool(ok), bool(same_start), mine, contributors)
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out[0][0] and out[1][0] and out[0][1] and out[1][1]
    assert sorted(out[0][2] + out[1][2]) == list(range(9))
    assert out[0][3] == out[1][3] == 1.0                       # the contributor slot travels with the gradient


def test_random_stream_keeps_the_indirect_flag_bit_clear():
    """torch's default generator is seeded with a random 64-bit value: bit 63 (the library's "offset is a device counter"
    flag, include/snuffy_b200.h) must never leak into a direct draw."""
    import torch
    from snuffy_b200 import engine
    state = torch.get_rng_state()
    try:
        torch.manual_seed((1 << 63) | 99)
        seed, offset = engine._RANDOM.next()
        assert seed == 99 and offset == 1
        assert engine._RANDOM.next() == (99, 2)
    finally:
        torch.set_rng_state(state)
        engine._RANDOM._seed = None
