"""Binary bag store: round trip, conversion from the reference's CSV layout (utils.py:138-183,
compute_feats.py:256-266) on the CPU; the pinned prefetcher on the GPU."""
import json
import os

import numpy as np
import pytest
import torch

from snuffy_b200 import store


def _bags(rs, lens, d):
    return [rs.standard_normal((n, d)).astype(np.float32) for n in lens]


def test_store_round_trip(tmp_path):
    rs = np.random.RandomState(0)
    bags = _bags(rs, [5, 1, 300], 16)
    labels = [1.0, 0.0, [1.0]]
    path = str(tmp_path / "s")
    store.write_store(path, bags, labels, names=["a", "b", "c"], patch_labels=[None, np.zeros(1), None])
    st = store.BagStore(path)
    assert len(st) == 3 and st.d == 16 and list(st.lengths) == [5, 1, 300]
    for i, b in enumerate(bags):
        assert np.array_equal(st.bag(i), b) and st.label(i).shape == (1,)
    assert st.name(2) == "c" and st.index["patch_labels"][0] is None
    assert os.path.getsize(path + ".bin") == 306 * 16 * 4
    with open(path + ".bin", "ab") as f:
        f.write(b"\0\0\0\0")
    with pytest.raises(ValueError):
        store.BagStore(path)                                           # size no longer matches the index
    with pytest.raises(ValueError):
        store.write_store(path, [bags[0], bags[1][:, :8]], [0, 1])     # ragged feature size


def test_csv_to_store_follows_the_reference_layout(tmp_path):
    pd = pytest.importorskip("pandas")
    rs = np.random.RandomState(1)
    rows = []
    originals = []
    for i, n in enumerate([7, 3]):
        feats = np.round(rs.standard_normal((n, 6)).astype(np.float32), 4)
        df = pd.DataFrame(feats, dtype=np.float32)
        if i == 0:                                                     # camelyon16-style patch labels / positions
            df["label"] = np.arange(n) % 2
            df["position"] = [f"({k}, {k})" for k in range(n)]
        p = tmp_path / f"slide_{i}.csv"
        df.to_csv(p, index=False, float_format="%.4f")                 # compute_feats.py:266
        rows.append((str(p), i))
        originals.append(feats)
    listing = tmp_path / "bags.csv"
    pd.DataFrame(rows).to_csv(listing, index=False)
    n = store.csv_to_store(str(listing), str(tmp_path / "conv"), num_classes=1)
    st = store.BagStore(str(tmp_path / "conv"))
    assert n == 2 and st.d == 6
    for i in range(2):
        assert np.allclose(st.bag(i), originals[i], atol=1e-4) and st.label(i)[0] == float(i)
    assert st.index["patch_labels"][0] == [0, 1, 0, 1, 0, 1, 0] and st.index["patch_labels"][1] is None
    assert st.name(0) == "slide_0"
    # multiclass labels are one-hot like utils.py:170-175
    store.csv_to_store(str(listing), str(tmp_path / "conv2"), num_classes=2)
    assert store.BagStore(str(tmp_path / "conv2")).label(1).tolist() == [0.0, 1.0]


@pytest.mark.gpu
def test_pinned_prefetcher_feeds_the_model(tmp_path):
    from helpers import build_snuffy, load_golden, load_params, snuffy_inputs
    from snuffy_b200 import snuffy
    rs = np.random.RandomState(2)
    _, c = load_golden("bin_tiny_relu")
    bags = _bags(rs, [64, 200, 90, 128], c["d"])
    store.write_store(str(tmp_path / "s"), bags, [0, 1, 1, 0])
    st = store.BagStore(str(tmp_path / "s"))
    params, _ = snuffy_inputs(c)
    model = load_params(build_snuffy(snuffy, c), params)
    order = [2, 0, 3, 1]
    pf = store.PinnedPrefetcher(st, order, "cuda")
    seen = []
    with torch.no_grad():
        for i, x, y in pf:
            assert x.shape == (1, len(bags[i]), c["d"]) and torch.equal(x[0].cpu(), torch.from_numpy(bags[i]))
            assert float(y) == [0, 1, 1, 0][i]
            _, bag, _ = model(x)
            ref = model(torch.from_numpy(bags[i])[None].cuda())[1]
            assert torch.equal(bag, ref)
            seen.append(i)
    assert seen == order and pf.h2d_bytes == sum(len(bags[i]) for i in order) * c["d"] * 4


@pytest.mark.gpu
def test_pinned_prefetcher_reuses_slots_without_corruption(tmp_path):
    """Many more bags than slots, a slow consumer kernel in between and an early exit: every bag arrives intact (a buffer is
    never overwritten while its copy or the consumer's work on it is pending) and the staging thread shuts down."""
    import threading
    rs = np.random.RandomState(5)
    lens = [int(n) for n in rs.randint(1000, 6000, 14)]
    bags = _bags(rs, lens, 64)
    store.write_store(str(tmp_path / "m"), bags, [i & 1 for i in range(len(bags))])
    st = store.BagStore(str(tmp_path / "m"))
    order = list(rs.permutation(len(bags)))
    pf = store.PinnedPrefetcher(st, order, "cuda", slots=2)
    big = torch.randn(4096, 4096, device="cuda")
    sums, seen = [], []
    for i, x, y in pf:
        _ = big @ big                                     # keeps the compute stream busy while later bags are staged
        sums.append(x.double().sum())                     # enqueued behind it: reads the slot after the matmul
        seen.append(i)
    torch.cuda.synchronize()
    assert seen == order
    for i, s_ in zip(seen, sums):
        assert abs(float(s_) - float(bags[i].astype(np.float64).sum())) < 1e-6 * max(1.0, abs(float(s_)))
    before = threading.active_count()
    for k, (i, x, y) in enumerate(store.PinnedPrefetcher(st, order, "cuda")):
        if k == 2:
            break
    assert threading.active_count() <= before
