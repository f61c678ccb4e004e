"""CPU oracle for the Snuffy / DSMIL MIL-aggregator hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``snuffy_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker.

What this is
------------
A numpy restatement of the math the reference evaluates on its hot path,
written from the equations in SURVEY.md Appendix A (not from the reference
source).  Every function is a pure function of explicit weights (a dict keyed
by the reference's ``state_dict`` names), an explicit selection ``S`` (or an
injected sampler for the random patches) and runs in float64 by default so it
can arbitrate between two fp32 implementations.

Reference anchors (file:line in jafarinia/snuffy @ 4b5b918):
  scores                 snuffy.py:39-41
  selection              snuffy.py:126-147      (multiclass: snuffy_multiclass.py:130-157)
  LN1 / attention block  snuffy.py:100-108, 183-205, 160-168
  scatter                snuffy.py:152-155      (multiclass: 164-168)
  FFN block              snuffy.py:109-110, 224-225
  final LN, mean, head   snuffy.py:82-86, 68-71
  model glue             snuffy.py:234-238
  DSMIL bag classifier   dsmil.py:72-92, 101-106
  loss glue              train.py:828-846

Parity pinning
--------------
The reference ships no tests, fixtures or golden vectors, so
this oracle is pinned against outputs of the reference itself: the committed
fixtures ``tests/golden/*.npz`` were produced by ``oracle/make_golden.py``,
which imports the unmodified reference modules from ``/root/reference`` in the
build container.  ``tests/test_oracle_golden.py`` checks this restatement
against every fixture.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

try:  # scipy is present in the image; keep a slow exact fallback anyway
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf, otypes=[np.float64])

Array = np.ndarray
LN_EPS = 1e-5  # nn.LayerNorm default used at snuffy.py:80,97


# --------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------
@dataclass
class SnuffyConfig:
    """Hyper-parameters of one aggregator (train.py:56-57, 98-108)."""

    d: int                      # feats_size
    heads: int                  # num_heads
    big_lambda: int             # K
    random_patch_share: float   # r
    depth: int = 1
    mlp_multiplier: int = 4
    activation: str = "relu"
    num_classes: int = 1

    @property
    def k_top(self) -> int:
        # Python doubles, exactly as written at snuffy.py:124,129
        return math.ceil(self.big_lambda * (1.0 - self.random_patch_share))

    def k_rand(self, n: int) -> int:
        # snuffy.py:137-140
        return min(int(self.big_lambda * self.random_patch_share), max(0, n - self.k_top))


# --------------------------------------------------------------------------
# elementary pieces
# --------------------------------------------------------------------------
def linear(x: Array, w: Array, b: Optional[Array]) -> Array:
    y = x @ w.T
    if b is not None:
        y = y + b
    return y


def layer_norm(x: Array, gamma: Array, beta: Array, eps: float = LN_EPS) -> Array:
    mu = x.mean(axis=-1, keepdims=True)
    xc = x - mu
    var = (xc * xc).mean(axis=-1, keepdims=True)  # biased, as torch
    return xc / np.sqrt(var + eps) * gamma + beta


def softmax(x: Array, axis: int) -> Array:
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=axis, keepdims=True)


def activation_fn(name: str) -> Callable[[Array], Array]:
    """relu / gelu(erf) / leakyrelu(0.01) / selu — snuffy.py:216-222."""
    if name == "relu":
        return lambda t: np.maximum(t, 0)
    if name == "gelu":
        return lambda t: 0.5 * t * (1.0 + _erf(t / math.sqrt(2.0)))
    if name == "leakyrelu":
        return lambda t: np.where(t >= 0, t, 0.01 * t)
    if name == "selu":
        alpha = 1.6732632423543772848170429916717
        scale = 1.0507009873554804934193349852946
        return lambda t: scale * np.where(t > 0, t, alpha * np.expm1(np.minimum(t, 0)))
    raise KeyError(name)  # same exception class the reference raises (dict lookup)


def argsort_desc_stable(v: Array) -> Array:
    """Descending order, ties broken by LOWER index first."""
    v = np.asarray(v)
    # lexsort: last key is primary. primary = -value (ascending), secondary = index
    return np.lexsort((np.arange(v.shape[0]), -v.astype(np.float64)))


# --------------------------------------------------------------------------
# selection
# --------------------------------------------------------------------------
def select_top(c: Array, k_top: int) -> Array:
    """T = first k_top entries of the descending order of the scores."""
    order = argsort_desc_stable(c.reshape(-1))
    return order[:k_top].astype(np.int64)


def remaining_ascending(n: int, taken: Array) -> Array:
    """{0..n-1} minus `taken`, ascending (what python's set-diff of small ints yields)."""
    mask = np.ones(n, dtype=bool)
    mask[np.asarray(taken, dtype=np.int64)] = False
    return np.nonzero(mask)[0].astype(np.int64)


def numpy_global_rng_sampler(remaining: Array, k: int) -> Array:
    """The reference's sampler: one draw from NumPy's GLOBAL RNG (snuffy.py:141-143)."""
    return np.random.choice(remaining, k, replace=False).astype(np.int64)


def select_binary(c: Array, n: int, cfg: SnuffyConfig,
                  sampler: Optional[Callable[[Array, int], Array]] = None,
                  top: Optional[Array] = None) -> Tuple[Array, Array]:
    """(T, R) for the binary path."""
    t = select_top(c, cfg.k_top) if top is None else np.asarray(top, dtype=np.int64)
    kr = cfg.k_rand(n)
    if kr == 0:
        return t, np.zeros(0, dtype=np.int64)
    if sampler is None:
        raise ValueError("random patches requested: pass `sampler` or explicit indices")
    return t, sampler(remaining_ascending(n, t), kr)


def select_multiclass(c: Array, cfg: SnuffyConfig,
                      sampler: Optional[Callable[[Array, int], Array]] = None
                      ) -> Tuple[Array, Array]:
    """(T[B,ref], R[B,ref]) for snuffy_multiclass.py:130-157.

    c: [B, N, C].  Per bag: top-k_top per class, flattened, unique (ascending);
    ref = min_b |U_b| then min(ref, N-ref); T_b = U_b[:ref]; R_b drawn from the
    complement of the FULL U_b; ref random patches regardless of r.
    """
    bsz, n, ncls = c.shape
    uniq: List[Array] = []
    for b in range(bsz):
        per_class = [argsort_desc_stable(c[b, :, j])[:cfg.k_top] for j in range(ncls)]
        uniq.append(np.unique(np.concatenate(per_class)).astype(np.int64))
    ref = min(len(u) for u in uniq)
    ref = min(ref, n - ref)
    t = np.stack([u[:ref] for u in uniq]) if ref > 0 else np.zeros((bsz, 0), np.int64)
    if ref == 0:
        return t, np.zeros((bsz, 0), np.int64)
    if sampler is None:
        raise ValueError("multiclass path always draws random patches: pass `sampler`")
    r = np.stack([sampler(remaining_ascending(n, uniq[b]), ref) for b in range(bsz)])
    return t, r.astype(np.int64)


# --------------------------------------------------------------------------
# one encoder layer on one bag
# --------------------------------------------------------------------------
def _layer_params(params: Dict[str, Array], l: int, dtype) -> Dict[str, Array]:
    p = f"b_classifier.encoder.layers.{l}."
    g = lambda k: np.asarray(params[p + k], dtype=dtype)
    return {
        "wq": g("self_attn.linears.0.weight"), "bq": g("self_attn.linears.0.bias"),
        "wk": g("self_attn.linears.1.weight"), "bk": g("self_attn.linears.1.bias"),
        "wv": g("self_attn.linears.2.weight"), "bv": g("self_attn.linears.2.bias"),
        "wo": g("self_attn.linears.3.weight"), "bo": g("self_attn.linears.3.bias"),
        "w1": g("feed_forward.w_1.weight"), "b1": g("feed_forward.w_1.bias"),
        "w2": g("feed_forward.w_2.weight"), "b2": g("feed_forward.w_2.bias"),
        "g1": g("sublayer.0.norm.weight"), "be1": g("sublayer.0.norm.bias"),
        "g2": g("sublayer.1.norm.weight"), "be2": g("sublayer.1.norm.bias"),
    }


def sparse_attention(q: Array, kp: Array, v: Array, heads: int,
                     keep_mask: Optional[Array] = None, p_drop: float = 0.0
                     ) -> Tuple[Array, Array]:
    """snuffy.py:160-168 + head split/concat 187-201.

    q, v: [N, d]; kp: [Ksel, d].  Returns O [Ksel, d] and P [h, N, Ksel].
    softmax over the Ksel axis per query row; O_j = P_j^T V_j.
    """
    n, d = q.shape
    ks = kp.shape[0]
    dk = d // heads
    # the reference divides an fp32 tensor by the python double sqrt(dk)
    scale = q.dtype.type(math.sqrt(dk))
    out = np.empty((ks, d), dtype=q.dtype)
    probs = np.empty((heads, n, ks), dtype=q.dtype)
    for j in range(heads):
        sl = slice(j * dk, (j + 1) * dk)
        s = (q[:, sl] @ kp[:, sl].T) / scale
        p = softmax(s, axis=-1)
        if keep_mask is not None:
            p = p * keep_mask[j] / (1.0 - p_drop)
        probs[j] = p
        out[:, sl] = p.T @ v[:, sl]
    return out, probs


def encoder_layer(x: Array, sel: Array, lp: Dict[str, Array], cfg: SnuffyConfig
                  ) -> Tuple[Array, Array, Dict[str, Array]]:
    """One layer on one bag x [N,d] with explicit selection `sel` (T ++ R).

    Returns (x_next, P, intermediates).  Eval mode (all dropouts identity).
    """
    act = activation_fn(cfg.activation)
    x_sel = x[sel]                                   # raw rows (App. B-1)
    u = layer_norm(x, lp["g1"], lp["be1"])
    q = linear(u, lp["wq"], lp["bq"])
    kp = linear(x_sel, lp["wk"], lp["bk"])
    v = linear(u, lp["wv"], lp["bv"])
    o, p = sparse_attention(q, kp, v, cfg.heads)
    z = linear(o, lp["wo"], lp["bo"])
    x_sel_new = x_sel + z
    y = x.copy()
    y[sel] = x_sel_new
    hdn = act(linear(layer_norm(y, lp["g2"], lp["be2"]), lp["w1"], lp["b1"]))
    x_next = y + linear(hdn, lp["w2"], lp["b2"])
    return x_next, p, {"q": q, "kp": kp, "v": v, "o": o, "x_sel_new": x_sel_new, "y": y}


# --------------------------------------------------------------------------
# whole-model forwards
# --------------------------------------------------------------------------
def instance_scores(x: Array, params: Dict[str, Array], dtype=np.float64) -> Array:
    w = np.asarray(params["i_classifier.fc.0.weight"], dtype=dtype)
    b = np.asarray(params["i_classifier.fc.0.bias"], dtype=dtype)
    return linear(np.asarray(x, dtype=dtype), w, b)


def bag_head(x: Array, params: Dict[str, Array], dtype=np.float64) -> Array:
    """final LN -> mean over ALL N tokens -> linear (snuffy.py:86, 71)."""
    z = layer_norm(x, np.asarray(params["b_classifier.encoder.norm.weight"], dtype=dtype),
                   np.asarray(params["b_classifier.encoder.norm.bias"], dtype=dtype))
    pooled = z.mean(axis=-2)
    return linear(pooled, np.asarray(params["b_classifier.linear.weight"], dtype=dtype),
                  np.asarray(params["b_classifier.linear.bias"], dtype=dtype))


def snuffy_forward(x: Array, params: Dict[str, Array], cfg: SnuffyConfig,
                   selections: Optional[Sequence[Array]] = None,
                   sampler: Optional[Callable[[Array, int], Array]] = None,
                   scores: Optional[Array] = None,
                   dtype=np.float64, keep_layers: bool = False) -> Dict[str, object]:
    """Binary Snuffy MILNet.forward (snuffy.py:234-238) on x [1,N,d] or [N,d].

    selections: optional list (len depth) of int arrays S_l = T ++ R_l.  When
    absent, T comes from the scores and R from `sampler` (called once per
    layer, like the reference).  `scores` optionally overrides c for the
    selection only (to decouple index parity from fp32 rounding of c).
    """
    x = np.asarray(x)
    if x.ndim == 3:
        if x.shape[0] != 1:
            raise ValueError("binary snuffy path is batch-1 only (snuffy.py:129-131)")
        x = x[0]
    if cfg.num_classes != 1:
        raise ValueError("binary snuffy path is single-class only (snuffy.py:129-131)")
    x = x.astype(dtype)
    n = x.shape[0]
    c = instance_scores(x, params, dtype)
    c_sel = c if scores is None else np.asarray(scores).reshape(n, 1)
    top = select_top(c_sel, cfg.k_top)
    used: List[Array] = []
    layers_out: List[Array] = []
    p = None
    for l in range(cfg.depth):
        if selections is not None:
            s = np.asarray(selections[l], dtype=np.int64)
        else:
            t, r = select_binary(c_sel, n, cfg, sampler, top=top)
            s = np.concatenate([t, r])
        used.append(s)
        x, p, _ = encoder_layer(x, s, _layer_params(params, l, dtype), cfg)
        if keep_layers:
            layers_out.append(x.copy())
    bag = bag_head(x, params, dtype)
    return {
        "classes": c[None],            # [1,N,1]
        "bag": bag[None],              # [1,1]
        "attn": p[None] if p is not None else None,   # [1,h,N,Ksel] (last layer)
        "selections": used,
        "layers": layers_out,
    }


def snuffy_multiclass_forward(x: Array, params: Dict[str, Array], cfg: SnuffyConfig,
                              selections: Optional[Sequence[Array]] = None,
                              sampler: Optional[Callable[[Array, int], Array]] = None,
                              dtype=np.float64) -> Dict[str, object]:
    """snuffy_multiclass MILNet.forward (snuffy_multiclass.py:249-253) on x [B,N,d].

    selections: optional list (len depth) of int arrays [B, Ksel].
    """
    x = np.asarray(x).astype(dtype)
    bsz, n, _ = x.shape
    c = instance_scores(x, params, dtype)            # [B,N,C]
    used: List[Array] = []
    attn = None
    for l in range(cfg.depth):
        if selections is not None:
            s = np.asarray(selections[l], dtype=np.int64)
        else:
            t, r = select_multiclass(c, cfg, sampler)
            s = np.concatenate([t, r], axis=1)
        used.append(s)
        lp = _layer_params(params, l, dtype)
        nxt = np.empty_like(x)
        probs = []
        for b in range(bsz):
            nxt[b], p, _ = encoder_layer(x[b], s[b], lp, cfg)
            probs.append(p)
        x = nxt
        attn = np.stack(probs)
    bag = bag_head(x, params, dtype)                 # [B,C]
    return {"classes": c, "bag": bag, "attn": attn, "selections": used}


# --------------------------------------------------------------------------
# DSMIL (dsmil.py:72-92, 101-106)
# --------------------------------------------------------------------------
def dsmil_q(feats: Array, params: Dict[str, Array], nonlinear: bool, dtype) -> Array:
    g = lambda k: np.asarray(params[k], dtype=dtype)
    if nonlinear:
        h = np.maximum(linear(feats, g("b_classifier.q.0.weight"), g("b_classifier.q.0.bias")), 0)
        return np.tanh(linear(h, g("b_classifier.q.2.weight"), g("b_classifier.q.2.bias")))
    return linear(feats, g("b_classifier.q.weight"), g("b_classifier.q.bias"))


def dsmil_forward(x: Array, params: Dict[str, Array], nonlinear: bool = True,
                  passing_v: bool = False, dtype=np.float64) -> Dict[str, Array]:
    """dsmil.MILNet.forward on x [N,d] (any leading shape is flattened to rows)."""
    g = lambda k: np.asarray(params[k], dtype=dtype)
    d = params["i_classifier.fc.0.weight"].shape[1]
    feats = np.asarray(x).reshape(-1, d).astype(dtype)
    c = linear(feats, g("i_classifier.fc.0.weight"), g("i_classifier.fc.0.bias"))   # [N,C]
    if passing_v:  # Dropout(eval)=id -> Linear -> ReLU
        v = np.maximum(linear(feats, g("b_classifier.v.1.weight"), g("b_classifier.v.1.bias")), 0)
    else:
        v = feats
    q = dsmil_q(feats, params, nonlinear, dtype)                                   # [N,128]
    ncls = c.shape[1]
    crit = np.array([argsort_desc_stable(c[:, j])[0] for j in range(ncls)])        # arg-max rows
    q_max = dsmil_q(feats[crit], params, nonlinear, dtype)                         # [C,128]
    # the reference divides by an fp32-rounded sqrt(128) (dsmil.py:85)
    scale = dtype(np.float32(np.sqrt(np.float32(q.shape[1]))))
    a = softmax((q @ q_max.T) / scale, axis=0)                                     # over instances
    bmat = a.T @ v                                                                 # [C,d]
    w = g("b_classifier.fcc.weight")                                               # [C,C,d]
    logits = np.einsum("ocd,cd->o", w, bmat) + g("b_classifier.fcc.bias")
    return {"classes": c, "bag": logits[None], "attn": a, "B": bmat[None], "critical": crit}


# --------------------------------------------------------------------------
# loss glue (train.py:828-846) — defines what the backward must match
# --------------------------------------------------------------------------
def bce_with_logits(z: Array, y: Array, weight: Optional[Array] = None) -> float:
    """mean_i weight_i * BCE(z_i, y_i).

    Quirk kept from the caller: train.py:245-246 builds
    ``nn.BCEWithLogitsLoss(pos_weight)`` POSITIONALLY, so the tensor lands in the
    ``weight`` slot (a plain per-class rescale of the loss), not ``pos_weight``.
    """
    z = np.asarray(z, dtype=np.float64).reshape(-1)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    wt = np.ones_like(z) if weight is None else np.asarray(weight, np.float64).reshape(-1)
    log_sig = -np.logaddexp(0.0, -z)
    log_one_minus = -np.logaddexp(0.0, z)
    return float(np.mean(-wt * (y * log_sig + (1 - y) * log_one_minus)))


def mil_loss(classes: Array, bag: Array, label: Array, w: float = 0.5,
             weight: Optional[Array] = None) -> Tuple[float, Array]:
    """loss = w*BCE(bag) + (1-w)*BCE(max over instances); prediction mix."""
    cls = np.asarray(classes, dtype=np.float64)
    mx = cls.max(axis=0) if cls.ndim == 2 else cls.max(axis=1)
    loss = w * bce_with_logits(bag, label, weight) + (1 - w) * bce_with_logits(mx, label, weight)
    sig = lambda t: 1.0 / (1.0 + np.exp(-np.asarray(t, np.float64)))
    pred = (1 - w) * sig(mx).reshape(-1) + w * sig(bag).reshape(-1)
    return loss, pred
