"""Patch-level outputs after the hot path: sigmoid of the instance scores, FROC detection tuples and the
threshold filter (train.py:913-916, 312-320, 342-345, 138-141).  CPU: oracle and host parsing against the fixture made
from the reference's own `Snuffy._run_model` / `mp_thresholding`.  GPU: the CUDA kernels against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import patch_outputs_oracle as po
from snuffy_b200 import patch_outputs
from conftest import GOLDEN


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "patch_outputs.npz"))


def _kept(gold, i):
    return [(float(p), int(x), int(y)) for p, x, y in gold[f"kept{i}"]]


def test_oracle_matches_reference_fixture(gold):
    p64 = po.patch_probs(gold["logits"])
    assert p64.shape == gold["probs32"].shape
    assert np.abs(p64 - gold["probs32"]).max() <= 1e-7
    pos = po.parse_positions(list(gold["names"]))
    assert pos == [tuple(r) for r in gold["positions"].tolist()]
    dets = po.detections(gold["probs32"], pos)
    assert dets == [(float(p), int(x), int(y)) for p, x, y in gold["detections"]]
    for i, t in enumerate(gold["thresholds"]):
        assert po.threshold_detections(dets, float(t)) == _kept(gold, i)


def test_host_position_parser_matches_reference_regex(gold):
    got = patch_outputs.parse_positions(list(gold["names"]))
    assert got.dtype == np.int32 and np.array_equal(got, gold["positions"])
    assert patch_outputs.parse_positions(["a/b/12_7.jpeg", "x3y44z5"]).tolist() == [[12, 7], [3, 44]]
    with pytest.raises(ValueError):
        patch_outputs.parse_positions(["only_1_number"[:6]])


@pytest.mark.gpu
def test_patch_probs_and_detections_match_reference_fixture(gold):
    dev = torch.device("cuda:0")
    logits = torch.from_numpy(gold["logits"]).to(dev)
    col = patch_outputs.PatchOutputCollector(total_rows=logits.shape[1] + 5, num_bags=2, num_classes=1, device=dev)
    view = col.add(logits, torch.from_numpy(np.atleast_1d(gold["pred"]).astype(np.float32)).to(dev))
    probs, preds, cu = col.to_host()
    assert cu.tolist() == [0, logits.shape[1]] and view.shape == (logits.shape[1], 1)
    assert np.abs(probs - gold["probs32"]).max() <= 2.4e-7             # 2 ulp at 1.0: torch.sigmoid fp32 vs 1/(1+expf(-z))
    assert np.allclose(preds[0], gold["pred"])
    # saturation ends exactly like the reference: sigmoid(+-100), sigmoid(+-40), sigmoid(+-0)
    assert probs[0, 0] == 0.5 and probs[1, 0] == 0.5 and probs[4, 0] == 1.0 and probs[5, 0] == 0.0
    # detection tuples from the REFERENCE probabilities (so the compaction is compared bit for bit)
    ref_p = torch.from_numpy(gold["probs32"]).to(dev)
    pos = torch.from_numpy(patch_outputs.parse_positions(list(gold["names"]))).to(dev)
    for i, t in enumerate(gold["thresholds"]):
        got = patch_outputs.froc_detections(ref_p, pos, None, float(np.float32(t)) if i != 1 else float(t))
        assert got == [_kept(gold, i)]
    named = patch_outputs.froc_detections(ref_p, pos, col.cu_seqlens(), 0.5, names=["slide"])
    assert named == {"slide": _kept(gold, 0)}


@pytest.mark.gpu
@pytest.mark.parametrize("lens", [[1], [255, 256, 257], [0, 10, 0, 1000, 3], [10000, 1, 50000]])
def test_froc_detections_ragged_slides_against_oracle(lens):
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(sum(lens) + len(lens))
    T = sum(lens)
    logits = (rs.standard_normal(T) * 2).astype(np.float32)
    pos = rs.randint(0, 1000, size=(T, 2)).astype(np.int32)
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    from snuffy_b200 import ops
    probs = ops.patch_probs(torch.from_numpy(logits).to(dev))
    p_h = probs.cpu().numpy()
    assert np.abs(p_h - po.patch_probs(logits).reshape(-1)).max() <= 2.4e-7
    thr = float(np.float32(0.6))
    got = patch_outputs.froc_detections(probs, torch.from_numpy(pos).to(dev), torch.from_numpy(cu).to(dev), thr)
    assert len(got) == len(lens)
    for b in range(len(lens)):
        s, e = cu[b], cu[b + 1]
        want = po.threshold_detections(po.detections(p_h[s:e], [tuple(r) for r in pos[s:e].tolist()]), thr)
        assert got[b] == want
    # two-column probabilities: column 0 is the one the binary caller reads
    two = torch.stack([probs, 1 - probs], dim=1).contiguous()
    assert patch_outputs.froc_detections(two, torch.from_numpy(pos).to(dev), torch.from_numpy(cu).to(dev), thr) == got


@pytest.mark.gpu
def test_patch_outputs_reject_bad_arguments():
    dev = torch.device("cuda:0")
    from snuffy_b200 import ops
    p = torch.rand(10, device=dev)
    with pytest.raises(ValueError):
        ops.froc_detections(p, torch.zeros(10, 2, dtype=torch.int64, device=dev), 0.5)
    with pytest.raises(ValueError):
        ops.patch_probs(p, out=torch.empty(9, device=dev))
    with pytest.raises(RuntimeError):
        ops.patch_probs(torch.rand(4))
    col = patch_outputs.PatchOutputCollector(4, 1, 1, dev)
    with pytest.raises(ValueError):
        col.add(torch.zeros(1, 5, 1, device=dev))


@pytest.mark.gpu
def test_validate_loop_matches_the_reference_valid_loop(tmp_path):
    """store -> pinned prefetcher -> trainer.validate -> collector, against train.py:334-355 replayed bag by bag with torch."""
    from helpers import build_snuffy, load_golden, load_params, snuffy_inputs
    from snuffy_b200 import dp, snuffy, store
    rs = np.random.RandomState(4)
    _, c = load_golden("bin_tiny_relu")
    lens = [64, 200, 90]
    bags = [rs.standard_normal((n, c["d"])).astype(np.float32) for n in lens]
    labels = [0.0, 1.0, 1.0]
    pos_names = [[f"{rs.randint(0, 99)}_{rs.randint(0, 99)}.jpeg" for _ in range(n)] for n in lens]
    store.write_store(str(tmp_path / "v"), bags, labels, positions=pos_names)
    st = store.BagStore(str(tmp_path / "v"))
    params, _ = snuffy_inputs(c)
    model = load_params(build_snuffy(snuffy, c), params)
    trainer = dp.DataParallelTrainer(model, mix_weight=0.5)
    col = patch_outputs.PatchOutputCollector(sum(lens), len(lens), 1, "cuda")
    order = [1, 2, 0]
    loss, ids = trainer.validate(store.PinnedPrefetcher(st, order, "cuda"), col)
    probs, preds, cu = col.to_host()
    assert ids == order and cu.tolist() == [0, 200, 290, 354]
    crit = torch.nn.BCEWithLogitsLoss()
    total = 0.0
    with torch.no_grad():
        for k, i in enumerate(order):
            ins, bag, _ = model(torch.from_numpy(bags[i])[None].cuda())
            mx, _ = torch.max(ins, 1)
            y = torch.tensor([[labels[i]]], device="cuda")
            total += float(0.5 * crit(bag.view(1, -1), y) + 0.5 * crit(mx.view(1, -1), y))            # train.py:836-838
            want_pred = 0.5 * torch.sigmoid(mx) + 0.5 * torch.sigmoid(bag)                              # train.py:840-844
            assert abs(float(want_pred) - preds[k, 0]) <= 1e-6
            assert np.abs(torch.sigmoid(ins.view(-1, 1)).cpu().numpy() - probs[cu[k]:cu[k + 1]]).max() <= 1e-6
    assert abs(float(loss) - total / 3) <= 1e-5
    pos = np.concatenate([patch_outputs.parse_positions(st.index["positions"][i]) for i in order])
    thr = float(np.median(probs))
    dets = patch_outputs.froc_detections(col.probs, torch.from_numpy(pos).cuda(), col.cu_seqlens(), thr,
                                         names=[st.name(i) for i in order])
    for k, i in enumerate(order):
        want = po.threshold_detections(po.detections(probs[cu[k]:cu[k + 1]], po.parse_positions(pos_names[i])), thr)
        assert dets[st.name(i)] == want
