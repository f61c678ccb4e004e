"""Training-loop glue and the data-parallel trainer on the GPU (reference caller
train.py:828-846, 468-473, 809-826): fused loss vs torch, flat AdamW vs torch.optim.AdamW, the world-1 trajectory vs the
reference op sequence trained on the CPU in float64, and (when 2 GPUs are visible) NCCL data parallelism."""
import copy
import os
import socket

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import build_snuffy, force_selections, load_golden, load_params, set_precision, snuffy_inputs
from test_gpu_backward import _params64, mil_loss as ref_mil_loss, ref_forward

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,N,C", [(1, 257, 1), (3, 1000, 2), (2, 10, 3)])
def test_fused_mil_loss_matches_torch(B, N, C):
    from snuffy_b200 import dp
    g = torch.Generator(device="cuda").manual_seed(N)
    classes = (torch.randn(B, N, C, device="cuda", generator=g) * 2).requires_grad_(True)
    bag = torch.randn(B, C, device="cuda", generator=g).requires_grad_(True)
    label = (torch.rand(B, C, device="cuda", generator=g) > 0.5).float()
    weight = torch.rand(C, device="cuda", generator=g) + 0.5
    for w, wt in [(0.5, None), (0.3, weight)]:
        classes.grad = bag.grad = None
        loss, pred, terms = dp.mil_loss(classes, bag, label, w, wt)
        (loss * 1.7).backward()
        c64 = classes.detach().double().requires_grad_(True)
        b64 = bag.detach().double().requires_grad_(True)
        wt64 = None if wt is None else wt.double()
        mx = c64.max(dim=1)[0]
        lb = F.binary_cross_entropy_with_logits(b64, label.double(), weight=wt64)
        lm = F.binary_cross_entropy_with_logits(mx, label.double(), weight=wt64)
        ref = w * lb + (1 - w) * lm
        (ref * 1.7).backward()
        assert abs(float(loss) - float(ref)) < 1e-5
        assert abs(float(terms[0]) - float(lb)) < 1e-5 and abs(float(terms[1]) - float(lm)) < 1e-5
        assert (classes.grad.double() - c64.grad).abs().max() < 1e-6
        assert (bag.grad.double() - b64.grad).abs().max() < 1e-6
        ref_pred = (1 - w) * torch.sigmoid(mx) + w * torch.sigmoid(b64)
        assert (pred.double() - ref_pred).abs().max() < 1e-6


@pytest.mark.parametrize("clip", [None, 0.05])
def test_flat_adamw_matches_torch(clip):
    from snuffy_b200 import dp
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(37, 29), torch.nn.LayerNorm(29), torch.nn.Linear(29, 3)).cuda()
    ref = copy.deepcopy(model)
    flat = dp.FlatBuffers(model.parameters())
    opt = dp.FlatAdamW(flat, lr=2e-3, betas=(0.5, 0.9), weight_decay=5e-3, clip_grad=clip)
    ropt = torch.optim.AdamW(ref.parameters(), lr=2e-3, betas=(0.5, 0.9), weight_decay=5e-3)
    g = torch.Generator(device="cuda").manual_seed(1)
    for step in range(5):
        x = torch.randn(16, 37, device="cuda", generator=g)
        flat.zero_grad(); ropt.zero_grad()
        model(x).square().sum().backward()
        ref(x).square().sum().backward()
        if clip is not None:
            torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm=clip)
        opt.step(); ropt.step()
        for p, q in zip(model.parameters(), ref.parameters()):
            assert (p - q).abs().max() < 2e-6, (step, (p - q).abs().max())


def test_world1_trainer_follows_the_reference_trajectory():
    """4 optimizer steps of the reference loop (one bag per step, AdamW lr 2e-4 betas (0.5, 0.9) wd 5e-3: train.py
    defaults 58, 61, 110) on the CPU in float64 vs DataParallelTrainer at world size 1."""
    from snuffy_b200 import dp, snuffy
    z, c = load_golden("bin_rand_gelu")
    params, _ = snuffy_inputs(c)
    model = load_params(build_snuffy(snuffy, c), params)
    for m in model.modules():                       # dropout 0 (train mode would draw the reference's default 0.1)
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    set_precision(model, "fp32")
    rs = np.random.RandomState(5)
    n, d = c["n"], c["d"]
    ksel = len(z["ref32_sel"][0])
    bags = rs.standard_normal((4, 1, n, d)).astype(np.float32)
    labels = np.array([1.0, 0.0, 0.0, 1.0], dtype=np.float32)
    sels = [[rs.permutation(n)[:ksel][None] for _ in range(c["depth"])] for _ in range(4)]
    trainer = dp.DataParallelTrainer(model, lr=2e-3, betas=(0.5, 0.9), weight_decay=5e-3)
    P64 = _params64(params)
    ropt = torch.optim.AdamW(list(P64.values()), lr=2e-3, betas=(0.5, 0.9), weight_decay=5e-3)
    for i in range(4):
        force_selections(model, sels[i])
        loss = trainer.train_step(torch.from_numpy(bags[i]).cuda(), torch.tensor([[labels[i]]]))
        ropt.zero_grad()
        c64, bag64 = ref_forward(torch.from_numpy(bags[i]).double(), P64, c["heads"], c["depth"], c["act"],
                                 [torch.from_numpy(s) for s in sels[i]])
        rl = ref_mil_loss(c64, bag64, torch.tensor([[labels[i]]]).double())
        rl.backward()
        ropt.step()
        assert abs(float(loss) - float(rl)) < 2e-4, (i, float(loss), float(rl))
    for name, p in model.named_parameters():
        assert (p.detach().double().cpu() - P64[name].detach()).abs().max() < 5e-4, name


def test_mil_loss_device_mix_weight_and_nan_scores():
    """The mix weight as a learnable device tensor (train.py:804, --soft_average): value read on the device, gradient =
    bag term - max term.  An all-NaN score column gives a NaN loss like torch.max, not an out-of-bounds write."""
    from snuffy_b200 import dp
    g = torch.Generator(device="cuda").manual_seed(3)
    classes = (torch.randn(2, 300, 2, device="cuda", generator=g) * 2).requires_grad_(True)
    bag = torch.randn(2, 2, device="cuda", generator=g).requires_grad_(True)
    label = (torch.rand(2, 2, device="cuda", generator=g) > 0.5).float()
    w = torch.tensor(0.3, device="cuda", requires_grad=True)
    loss, pred, terms = dp.mil_loss(classes, bag, label, w)
    (loss * 2.0).backward()
    c64, b64 = classes.detach().double().requires_grad_(True), bag.detach().double().requires_grad_(True)
    w64 = torch.tensor(0.3, dtype=torch.float64, requires_grad=True)
    lb = F.binary_cross_entropy_with_logits(b64, label.double().cpu().cuda())
    lm = F.binary_cross_entropy_with_logits(c64.max(dim=1)[0], label.double())
    (2.0 * (w64.cuda() * lb + (1 - w64.cuda()) * lm)).backward()
    assert abs(float(w.grad) - float(w64.grad)) < 1e-5 and w.grad.shape == w.shape
    assert (classes.grad.double() - c64.grad).abs().max() < 1e-6 and (bag.grad.double() - b64.grad).abs().max() < 1e-6
    bad = classes.detach().clone()
    bad[1, :, 0] = float("nan")
    loss, _, _ = dp.mil_loss(bad, bag.detach(), label, 0.5)
    torch.cuda.synchronize()
    assert torch.isnan(loss)
    one_nan = classes.detach().clone()
    one_nan[0, 17, 1] = float("nan")                                   # torch.max propagates a single NaN too
    assert torch.isnan(dp.mil_loss(one_nan, bag.detach(), label, 0.5)[0])


def test_soft_average_trainer_matches_two_group_adamw():
    """--soft_average: the mix weight is a second AdamW parameter group at lr x 0.1, clamped to [0, 1] after each step
    (train.py:804, 817-825, 852-854) — DataParallelTrainer(soft_average=True) vs that loop in float64 on the CPU."""
    from snuffy_b200 import dp, snuffy
    z, c = load_golden("bin_tiny_relu")
    params, _ = snuffy_inputs(c)
    model = load_params(build_snuffy(snuffy, c), params)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    set_precision(model, "fp32")
    rs = np.random.RandomState(6)
    ksel = len(z["ref32_sel"][0])
    lr, mult = 5e-2, 0.9                                                # large steps so that the clamp at 0 is reached
    trainer = dp.DataParallelTrainer(model, lr=lr, betas=(0.5, 0.9), weight_decay=5e-3, mix_weight=0.03, soft_average=True,
                                     single_weight_lr_multiplier=mult)
    assert trainer.single_weight_parameter.requires_grad and trainer.flat.numel > trainer.flat.model_numel
    P64 = _params64(params)
    w64 = torch.tensor(0.03, dtype=torch.float64, requires_grad=True)
    ropt = torch.optim.AdamW([{"params": [w64], "lr": lr * mult}, {"params": list(P64.values())}], lr=lr, betas=(0.5, 0.9),
                             weight_decay=5e-3)
    hit_clamp = False
    for i in range(5):
        x = rs.standard_normal((1, c["n"], c["d"])).astype(np.float32)
        y = 1.0                                                          # bag term > max term or not: w moves either way
        sels = [rs.permutation(c["n"])[:ksel][None] for _ in range(c["depth"])]
        force_selections(model, sels)
        loss = trainer.train_step(torch.from_numpy(x).cuda(), torch.tensor([[y]]))
        ropt.zero_grad()
        c64, bag64 = ref_forward(torch.from_numpy(x).double(), P64, c["heads"], c["depth"], c["act"],
                                 [torch.from_numpy(s_) for s_ in sels])
        yl = torch.tensor([[y]]).double()
        rl = w64 * F.binary_cross_entropy_with_logits(bag64.view(1, -1), yl) + \
            (1 - w64) * F.binary_cross_entropy_with_logits(c64.max(dim=1)[0].view(1, -1), yl)
        rl.backward(); ropt.step()
        before_clamp = float(w64)
        w64.data.clamp_(0, 1)
        hit_clamp |= before_clamp != float(w64)
        assert abs(float(loss) - float(rl)) < 5e-4, (i, float(loss), float(rl))
        assert abs(float(trainer.single_weight_parameter) - float(w64)) < 2e-4, (i, float(trainer.single_weight_parameter), float(w64))
    assert 0.0 <= float(trainer.single_weight_parameter) <= 1.0 and hit_clamp


def test_idle_step_and_uneven_shards_world1():
    """A rank whose shard ran out joins with train_step(None, None): zero gradient, contributor count 0 -> at world size 1
    the update is skipped entirely (nobody contributed); a normal step divides by its own count of 1."""
    from snuffy_b200 import dp, snuffy
    z, c = load_golden("bin_tiny_relu")
    params, x = snuffy_inputs(c)
    model = load_params(build_snuffy(snuffy, c), params)
    trainer = dp.DataParallelTrainer(model, lr=1e-2)
    trainer.train_step(torch.from_numpy(x).cuda(), torch.ones(1, 1))
    assert float(trainer.flat.contributors) == 1.0 and trainer.opt.step_count == 1
    p1 = trainer.flat.flat_param.clone()
    assert trainer.train_step(None, None) is None
    assert float(trainer.flat.contributors) == 0.0
    assert torch.equal(p1, trainer.flat.flat_param)
    assert dp.steps_per_epoch(6, 2, lengths=[100, 1, 1, 1, 1, 1]) == 5 and dp.steps_per_epoch(6, 2) == 3
    steps = dp.train_epoch(trainer, [0, 1], lambda i: (torch.from_numpy(x).cuda(), torch.ones(1, 1)))
    assert steps == 2 and not torch.equal(p1, trainer.flat.flat_param)


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    os.environ.setdefault("SNUFFY_B200_PEER_TIMEOUT_S", "30")              # a lost peer fails the test instead of hanging the box
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from snuffy_b200 import dp, snuffy
        z, c = load_golden("bin_tiny_relu")
        params, _ = snuffy_inputs(c)
        model = load_params(build_snuffy(snuffy, c), params, device=f"cuda:{rank}")
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        set_precision(model, "fp32")
        trainer = dp.DataParallelTrainer(model, lr=1e-3)
        # the exchange itself: the peer-memory kernel (csrc/comm.cu) against NCCL on a random buffer, three calls in a row
        peer_used, peer_err = trainer.flat.peer is not None, 0.0
        if peer_used:
            g = torch.Generator(device=f"cuda:{rank}").manual_seed(100 + rank)
            for _ in range(3):
                x = torch.randn(trainer.flat._grad_all.numel(), device=f"cuda:{rank}", generator=g)
                ref = x.clone()
                dist.all_reduce(ref)
                trainer.flat._grad_all.copy_(x)
                trainer.flat.allreduce_sum()
                peer_err = max(peer_err, float((trainer.flat._grad_all - ref).abs().max()))
            trainer.flat.zero_grad()
        rs = np.random.RandomState(11)
        bags = rs.standard_normal((4, 1, c["n"], c["d"])).astype(np.float32)
        labels = [1.0, 0.0, 1.0, 0.0]
        grads = []
        for step in range(2):
            i = dp.shard_slides(4, rank, world)[step]
            trainer.train_step(torch.from_numpy(bags[i]).cuda(), torch.tensor([[labels[i]]]))
            grads.append(trainer.flat.flat_grad.detach().cpu() / world)      # after the all-reduce: sum over ranks
        # uneven shards: rank 0 has one more bag, rank 1 joins with an idle step -> the update uses rank 0's gradient / 1
        before = trainer.flat.flat_param.detach().clone()
        if rank == 0:
            trainer.train_step(torch.from_numpy(bags[0]).cuda(), torch.tensor([[labels[0]]]))
        else:
            trainer.train_step(None, None)
        contributors = float(trainer.flat.contributors)
        moved = float((trainer.flat.flat_param - before).abs().max())
        out[rank] = (trainer.flat.flat_param.detach().cpu(), grads[0], contributors, moved, peer_used, peer_err)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpu_data_parallel_equals_gradient_averaging():
    """2 ranks x 1 bag per step == one process averaging the two bags' gradients (the DP extension of §8e)."""
    import torch.multiprocessing as mp
    from snuffy_b200 import dp, snuffy
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(2, port, out), nprocs=2, join=True)
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])      # replicas stay bit-identical
    assert out[0][2] == out[1][2] == 1.0 and out[0][3] == out[1][3] > 0                 # idle rank: one contributor, same update
    if os.environ.get("SNUFFY_B200_PEER_ALLREDUCE", "1") != "0":
        assert out[0][4] and out[1][4], "the peer-memory all-reduce was not set up (CUDA IPC between the two GPUs failed)"
        assert out[0][5] < 1e-5 and out[1][5] < 1e-5, (out[0][5], out[1][5])
    z, c = load_golden("bin_tiny_relu")
    params, _ = snuffy_inputs(c)
    model = load_params(build_snuffy(snuffy, c), params)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    set_precision(model, "fp32")
    trainer = dp.DataParallelTrainer(model, lr=1e-3)
    rs = np.random.RandomState(11)
    bags = rs.standard_normal((4, 1, c["n"], c["d"])).astype(np.float32)
    labels = [1.0, 0.0, 1.0, 0.0]
    # both bags of step 0 in ONE process (mean loss over the 2 bags) give the averaged gradient.  Parameters after Adam
    # are not compared: Adam normalises noise-level gradients (e.g. the key-projection bias, identically 0) to +-lr.
    ids = [dp.shard_slides(4, r, 2)[0] for r in range(2)]
    x = torch.from_numpy(np.concatenate([bags[i] for i in ids])).cuda()
    trainer.train_step(x, torch.tensor([[labels[i]] for i in ids]))
    g1, g2 = trainer.flat.flat_grad.cpu(), out[0][1]
    assert (g1 - g2).abs().max() < 1e-5 * max(1.0, float(g2.abs().max())), float((g1 - g2).abs().max())


def test_reference_style_training_loop_with_torch_optimizer():
    """The literal caller pattern of train.py:249-264, 828-846, 468-473 on the drop-in module: torch's BCEWithLogitsLoss on
    (bag, max instance), loss.backward(), torch.optim.AdamW.step() (in-place updates bump the parameter versions, so the
    cached operand planes are rebuilt), three bags -- against the same loop over the reference op sequence in float64."""
    from snuffy_b200 import snuffy
    z, c = load_golden("bin_tiny_relu")
    params, _ = snuffy_inputs(c)
    model = load_params(build_snuffy(snuffy, c), params)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    model.train()
    criterion = torch.nn.BCEWithLogitsLoss()
    w = torch.tensor(0.5, device="cuda")                                      # single_weight_parameter (train.py:804)
    opt = torch.optim.AdamW(model.parameters(), lr=2e-3, betas=(0.5, 0.9), weight_decay=5e-3)
    P64 = _params64(params)
    ropt = torch.optim.AdamW(list(P64.values()), lr=2e-3, betas=(0.5, 0.9), weight_decay=5e-3)
    rs = np.random.RandomState(9)
    ksel = len(z["ref32_sel"][0])
    for step in range(3):
        bag_np = rs.standard_normal((1, c["n"], c["d"])).astype(np.float32)
        label = torch.tensor([[float(step & 1)]])
        sels = [rs.permutation(c["n"])[:ksel][None] for _ in range(c["depth"])]
        force_selections(model, sels)
        ins_prediction, bag_prediction, _ = model(torch.from_numpy(bag_np).cuda())
        max_prediction, _ = torch.max(ins_prediction, 1)
        loss = w * criterion(bag_prediction.view(1, -1), label.cuda().view(1, -1)) + \
            (1 - w) * criterion(max_prediction.view(1, -1), label.cuda().view(1, -1))
        loss.backward()
        opt.step(); opt.zero_grad()
        ropt.zero_grad()
        c64, bag64 = ref_forward(torch.from_numpy(bag_np).double(), P64, c["heads"], c["depth"], c["act"],
                                 [torch.from_numpy(s) for s in sels])
        rl = ref_mil_loss(c64, bag64, label.double())
        rl.backward(); ropt.step()
        assert abs(float(loss) - float(rl)) < 2e-4, (step, float(loss), float(rl))


def _graph_pair(case, dropout, r=None, precision=None):
    from snuffy_b200 import dp, snuffy
    _, c = load_golden(case)
    if r is not None:
        c = dict(c, r=r)
    params, _ = snuffy_inputs(c)
    models = [load_params(build_snuffy(snuffy, c), params) for _ in range(2)]
    for model in models:
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = dropout
        for layer in model.b_classifier.encoder.layers:
            layer.self_attn.dropout.p = dropout
            layer.return_attn = False
        if precision:
            set_precision(model, precision)
    eager = dp.DataParallelTrainer(models[0], lr=1e-3, clip_grad=1.0)
    graph = dp.DataParallelTrainer(models[1], lr=1e-3, clip_grad=1.0, cuda_graph=True)
    return c, eager, graph


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_graph_captured_step_equals_eager_step(precision):
    """cuda_graph=True replays the same kernels: without randomness the two trainers stay bit-identical over several steps
    (new bag contents each step, parameters updated in place between replays)."""
    c, eager, graph = _graph_pair("bin_rand_gelu", dropout=0.0, r=0.0, precision=precision)
    rs = np.random.RandomState(3)
    for step in range(4):
        x = torch.from_numpy(rs.standard_normal((1, c["n"], c["d"])).astype(np.float32)).cuda()
        y = torch.tensor([[float(step & 1)]], device="cuda")
        le, lg = eager.train_step(x, y), graph.train_step(x, y)
        assert torch.equal(le, lg), (step, float(le), float(lg))
        assert torch.equal(eager.flat.flat_grad, graph.flat.flat_grad), step
        assert torch.equal(eager.flat.flat_param, graph.flat.flat_param), step
    # a second bag shape is captured on demand, then the first one replays again
    x2 = torch.from_numpy(rs.standard_normal((1, c["n"] + 37, c["d"])).astype(np.float32)).cuda()
    y = torch.ones(1, 1, device="cuda")
    assert torch.equal(eager.train_step(x2, y), graph.train_step(x2, y))
    assert torch.equal(eager.flat.flat_param, graph.flat.flat_param)
    # evaluation between replays sees the current parameters
    assert torch.equal(eager.predict(x2), graph.predict(x2))
    assert torch.equal(eager.train_step(x2, y), graph.train_step(x2, y))


def test_graph_replays_draw_fresh_masks_and_match_the_eager_draws():
    """Dropout 0.1 + random patches: every replay reads the device step counter, so (a) two replays on identical input and
    parameters differ, (b) forward and backward of a replay agree with each other — checked by running the eager trainer on
    the very (seed, offset) pairs the replay resolved and comparing loss and gradients."""
    from snuffy_b200 import engine
    torch.manual_seed(77)                                              # < 2^32: the indirect draws use seed & 0xFFFFFFFF
    c, eager, graph = _graph_pair("bin_rand_gelu", dropout=0.1, r=0.5, precision="bf16x3")
    for t in (eager, graph):
        t.opt.set_lr(0.0); t.opt.weight_decay = 0.0                     # parameters stay put
    rs = np.random.RandomState(4)
    x = torch.from_numpy(rs.standard_normal((1, c["n"], c["d"])).astype(np.float32)).cuda()
    y = torch.ones(1, 1, device="cuda")
    losses, grads, counters = [], [], []
    for _ in range(3):
        counters.append(int(graph._rng_counter) if graph._graph is not None else None)
        losses.append(float(graph.train_step(x, y)))
        grads.append(graph.flat.flat_grad.clone())
    assert len(set(losses)) == 3, losses
    n = graph._draws_per_step
    assert n >= 3 and int(graph._rng_counter) == counters[1] + 2 * n
    for k in (1, 2):                                                    # replay k used offsets counter + 1 .. counter + n
        engine._RANDOM._seed, engine._RANDOM._offset = torch.initial_seed(), counters[1] + (k - 1) * n
        le = float(eager.train_step(x, y))
        assert le == losses[k], (k, le, losses[k])
        assert torch.equal(eager.flat.flat_grad, grads[k]), k
