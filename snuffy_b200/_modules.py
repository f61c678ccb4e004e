"""Shared nn.Module building blocks of the drop-in ``snuffy`` / ``snuffy_multiclass`` modules.

Same class names, constructor signatures, attribute names and ``state_dict`` keys as the reference
(snuffy.py:34-238, snuffy_multiclass.py:34-253) so ``train.py`` / ``roi.py`` / published checkpoints work
unchanged; parameters live in real ``nn.Linear`` / ``nn.LayerNorm`` children (utils.py:69-120 re-initialises
them with ``.apply``).  The forward passes do not call PyTorch math: they dispatch to the sm_100a kernels in
``libsnuffy_b200.so`` via :mod:`snuffy_b200.engine`.
"""
from __future__ import annotations

import copy
from typing import Optional

import torch
import torch.nn as nn

from . import engine, ops
from .engine import LayerWeights

_ACTIVATIONS = {"relu": nn.ReLU, "gelu": nn.GELU, "leakyrelu": nn.LeakyReLU, "selu": nn.SELU}


def _require_cuda(x: torch.Tensor, who: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"{who}: input is on {x.device}; snuffy_b200 runs on a CUDA device only "
                           "(no CPU fallback). Move the module and the bag to cuda.")


class FCLayer(nn.Module):
    """Instance scorer (snuffy.py:34-41): returns (feats unchanged, c = Linear(feats))."""

    def __init__(self, in_size, out_size=1):
        super().__init__()
        self.in_size = in_size           # dsmil.py:31 keeps it; harmless for snuffy
        self.fc = nn.Sequential(nn.Linear(in_size, out_size))

    def forward(self, feats):
        _require_cuda(feats, "FCLayer")
        lin = self.fc[0]
        from .autograd import scores_fn
        return feats, scores_fn(feats, lin.weight, lin.bias)


class IClassifier(nn.Module):
    """Backbone + linear instance classifier (snuffy.py:44-54).  Only the Linear is ours."""

    def __init__(self, feature_extractor, feature_size, output_class):
        super().__init__()
        self.in_size = feature_size
        self.feature_extractor = feature_extractor
        self.fc = nn.Linear(feature_size, output_class)

    def forward(self, x):
        feats = self.feature_extractor(x)
        feats = feats.view(feats.shape[0], -1)
        _require_cuda(feats, "IClassifier")
        from .autograd import scores_fn
        return feats, scores_fn(feats, self.fc.weight, self.fc.bias)


def clones(module, N):
    "Produce N identical layers."
    return nn.ModuleList([copy.deepcopy(module) for _ in range(N)])


def _draw(p: float):
    return (float(p),) + engine._RANDOM.next() if p > 0.0 else (0.0, 0, 0)


def _linear(lin: nn.Linear, x2: torch.Tensor, act: str = "none", drop=(0.0, 0, 0)) -> torch.Tensor:
    """act(x2 lin.weight^T + lin.bias) on the exact-fp32 kernel, differentiable when anything asks for a gradient."""
    if torch.is_grad_enabled() and (x2.requires_grad or lin.weight.requires_grad):
        from .backward import LinearFunction
        return LinearFunction.apply(x2, lin.weight, lin.bias, act, drop)
    return ops.linear_f32(x2.detach(), lin.weight.detach(), None if lin.bias is None else lin.bias.detach(), act=act, drop=drop)


def _sparse_attn(q2, v2, k2, nb, n, k, h, drop):
    if torch.is_grad_enabled() and (q2.requires_grad or v2.requires_grad or k2.requires_grad):
        from .backward import SparseAttnFunction
        return SparseAttnFunction.apply(q2, v2, k2, nb, n, k, h, drop)
    o, probs, _ = ops.sparse_attn(q2.detach(), v2.detach(), k2.detach(), nb, n, k, h, want_probs=True, dropout_p=drop[0],
                                  seed=drop[1], offset=drop[2])
    return o, probs


def attention(query, key, value, dropout=None):
    """'Scaled dot product attention' with the reference's transposed aggregation (snuffy.py:160-168).

    query/value [nb, h, N, dk], key [nb, h, K, dk] -> (P^T V [nb, h, K, dk], P [nb, h, N, K]).
    API-compat helper (the encoder calls the fused kernel directly); layout shuffles only, math in CUDA.  Gradients flow to
    query / key / value through the first output; P is returned without a gradient (no caller differentiates it, App. B-9)."""
    _require_cuda(query, "attention")
    nb, h, n, dk = query.shape
    k = key.shape[2]
    p_drop = float(dropout.p) if dropout is not None and dropout.training else 0.0
    q2 = query.transpose(1, 2).reshape(nb * n, h * dk)
    v2 = value.transpose(1, 2).reshape(nb * n, h * dk)
    k2 = key.transpose(1, 2).reshape(nb * k, h * dk)
    o, probs = _sparse_attn(q2, v2, k2, nb, n, k, h, _draw(p_drop))
    return o.view(nb, k, h, dk).transpose(1, 2), probs


class MultiHeadedAttention(nn.Module):
    """snuffy.py:171-205.  Holds the four d x d projections; dropout default 0.1 like the reference."""

    def __init__(self, h, d_model, dropout=0.1):
        super().__init__()
        assert d_model % h == 0
        self.d_big_lambda = d_model // h
        self.h = h
        self.linears = clones(nn.Linear(d_model, d_model), 4)
        self.attn = None
        self.dropout = nn.Dropout(p=dropout)

    def forward(self, query, key, value):
        """Stand-alone call form (the fused encoder layer reads `linears` directly): query / value [nb, N, d], key [nb, K, d]
        -> (W3 . concat_heads(P^T V) [nb, K, d], P [nb, h, N, K])."""
        _require_cuda(query, "MultiHeadedAttention")
        nb = query.size(0)
        d = self.h * self.d_big_lambda
        proj = [_linear(lin, t.reshape(-1, d)) for lin, t in zip(self.linears, (query, key, value))]
        n, k = query.shape[1], key.shape[1]
        o, self.attn = _sparse_attn(proj[0], proj[2], proj[1], nb, n, k, self.h,
                                    _draw(float(self.dropout.p) if self.training else 0.0))
        return _linear(self.linears[3], o).view(nb, k, d), self.attn


class PositionwiseFeedForward(nn.Module):
    """snuffy.py:208-225: w_2(dropout(act(w_1 x)))."""

    def __init__(self, d_model, d_ff, activation, dropout=0.1):
        super().__init__()
        self.w_1 = nn.Linear(d_model, d_ff)
        self.w_2 = nn.Linear(d_ff, d_model)
        self.dropout = nn.Dropout(dropout)
        activation_dictionary = {name: cls() for name, cls in _ACTIVATIONS.items()}
        self.activation = activation_dictionary[activation]          # KeyError for unknown names, like the reference
        self.activation_name = activation

    def forward(self, x):
        _require_cuda(x, "PositionwiseFeedForward")
        shape = x.shape
        hdn = _linear(self.w_1, x.reshape(-1, shape[-1]), self.activation_name,
                      _draw(float(self.dropout.p) if self.training else 0.0))
        out = _linear(self.w_2, hdn)
        return out.view(*shape[:-1], out.shape[-1])


class SublayerConnection(nn.Module):
    """Pre-norm residual wrapper (snuffy.py:89-110).  Parameter holder for LN1/LN2: the fused encoder layer reads ``norm`` and
    ``dropout`` from here.  ``forward`` keeps the reference's stand-alone call form in both modes."""

    def __init__(self, size, dropout):
        super().__init__()
        self.norm = nn.LayerNorm(size)
        self.dropout = nn.Dropout(dropout)

    def layer_norm(self, x):
        from .autograd import layer_norm_fn
        return layer_norm_fn(x, self.norm.weight, self.norm.bias)

    def _residual(self, x, y):
        drop = _draw(float(self.dropout.p) if self.training else 0.0)
        if torch.is_grad_enabled() and (x.requires_grad or y.requires_grad):
            from .backward import ResidualDropoutFunction
            return ResidualDropoutFunction.apply(x, y, drop)
        return ops.residual_dropout(x.detach(), y.detach(), drop)

    def forward(self, x, sublayer, c=None, top_big_lambda_indices=None, random_indices=None, *rest, mode=None):
        """Apply residual connection to any sublayer with the same size.  Both reference call forms are accepted:
        snuffy.py:100  (x, sublayer, c, top_indices, random_indices, mode) and snuffy_multiclass.py:103
        (x, sublayer, c, top_indices [B, ref], random_indices [B, ref], num_batch, feats_size, num_classes, k, mode)."""
        _require_cuda(x, "SublayerConnection")
        if rest:
            mode = rest[-1]
        if mode == "ff":
            return self._residual(x, sublayer(self.layer_norm(x)))
        if mode != "attn":
            raise ValueError(f"SublayerConnection: mode must be 'attn' or 'ff', got {mode!r}")
        B = x.shape[0]
        idx = top_big_lambda_indices.reshape(B, -1)
        if random_indices is not None:
            idx = torch.cat((idx, random_indices.reshape(B, -1)), dim=1)
        idx = idx.to(device=x.device, dtype=torch.int64).contiguous()
        if torch.is_grad_enabled() and x.requires_grad:
            from .backward import GatherRowsFunction
            top_big_lambda = GatherRowsFunction.apply(x, idx)
        else:
            top_big_lambda = ops.gather_rows(x.detach().contiguous(), idx)
        multiheadedattn = sublayer(self.layer_norm(x))
        return self._residual(top_big_lambda, multiheadedattn[0]), multiheadedattn[1]


class Encoder(nn.Module):
    "Core encoder is a stack of N layers (snuffy.py:74-86)"

    def __init__(self, layer, N):
        super().__init__()
        self.layers = clones(layer, N)
        self.norm = nn.LayerNorm(layer.size)

    def run_layers(self, x, c):
        """All layers WITHOUT the final LayerNorm (BClassifier fuses it with the mean-pool + head)."""
        attn = None
        state = {}
        hint = getattr(self, "_zplanes_hint", None)           # (data_ptr of x, planes) left by MILNet.forward's fused scorer
        self._zplanes_hint = None
        if hint is not None and hint[0] == x.data_ptr():
            state["zplanes"] = hint[1]
        for i, layer in enumerate(self.layers):
            x, attn = layer(x, c, i, _state=state)
            state.pop("zplanes", None)
        return x, attn

    def forward(self, x, c):
        "Pass the input through each layer in turn; returns (LN_f(x), A of the last layer)."
        _require_cuda(x, "Encoder")
        x, attn = self.run_layers(x, c)
        from .autograd import layer_norm_fn
        return layer_norm_fn(x, self.norm.weight, self.norm.bias), attn


class BClassifier(nn.Module):
    """Bag classifier (snuffy.py:62-71): encoder -> mean over ALL tokens -> linear."""

    def __init__(self, encoder, num_classes, input_size: int):
        super().__init__()
        self.encoder = encoder
        self.linear = nn.Linear(input_size, num_classes)

    def forward(self, x, c):
        _require_cuda(x, "BClassifier")
        x, attentions = self.encoder.run_layers(x, c)
        from .autograd import ln_mean_head_fn
        bag = ln_mean_head_fn(x, self.encoder.norm.weight, self.encoder.norm.bias, self.linear.weight, self.linear.bias)
        return bag, attentions


class MILNet(nn.Module):
    """snuffy.py:228-238: returns (classes, prediction_bag, A)."""

    def __init__(self, i_classifier, b_classifier):
        super().__init__()
        self.i_classifier = i_classifier
        self.b_classifier = b_classifier

    def forward(self, x):
        if x.dim() == 3 and engine.fused_scores_available(self, x):
            # inference: ONE pass over the bag yields the instance scores and layer 0's normalised operand planes
            lin = self.i_classifier.fc[0]
            feats = x.contiguous()
            classes, planes = ops.scores_ln_planes(feats, lin.weight.detach(), lin.bias.detach() if lin.bias is not None else None)
            self.b_classifier.encoder._zplanes_hint = (feats.data_ptr(), planes)
        else:
            feats, classes = self.i_classifier(x)
        prediction_bag, A = self.b_classifier(feats, classes)
        return classes, prediction_bag, A


class EncoderLayerBase(nn.Module):
    """Fused encoder layer: selection -> gather -> LN1+Q|V GEMM -> key proj -> sparse attention -> out proj +
    residual -> LN2+FFN over all tokens (selected rows read through row_map, x never cloned)."""

    multiclass = False

    #: "bf16x3" (tcgen05, 3-pass split, default) | "fp32" (SIMT, exact fp32) | "bf16x1"; None = env/default
    precision: Optional[str] = None
    #: "device" (Philox sampler on the GPU) | "numpy" (the reference's NumPy global-RNG stream, with host trip)
    random_mode: str = "device"
    #: materialise the [B, h, N, Ksel] attention tensor the reference returns (nobody consumes it, App. B-9)
    return_attn: bool = True

    def _init_common(self, size, self_attn, feed_forward, dropout, big_lambda, random_patch_share):
        self.self_attn = self_attn
        self.feed_forward = feed_forward
        self.sublayer = clones(SublayerConnection(size, dropout), 2)
        self.size = size
        self.big_lambda = big_lambda
        self.random_patch_share = random_patch_share
        self.top_big_lambda_share = 1.0 - random_patch_share
        self._wcache = None
        self.forced_selection = None      # tests / parity: explicit S [B, Ksel] for this layer

    # ---- parameters as kernel operands (cached until any of them is updated in place or replaced)
    def _params(self):
        a, f, s = self.self_attn.linears, self.feed_forward, self.sublayer
        return (a[0].weight, a[0].bias, a[1].weight, a[1].bias, a[2].weight, a[2].bias, a[3].weight, a[3].bias,
                f.w_1.weight, f.w_1.bias, f.w_2.weight, f.w_2.bias, s[0].norm.weight, s[0].norm.bias,
                s[1].norm.weight, s[1].norm.bias)

    def layer_weights(self) -> LayerWeights:
        """Kernel operands derived from the parameters (fused Q|V weight, LN-folded and transposed planes).  In train mode
        they are rebuilt on every forward (parameters change every step, and an update through `p.data` or a raw pointer
        does not bump `p._version`); in eval mode they are cached until a parameter is replaced or updated in place through
        autograd-visible ops.  After editing `p.data` of an eval-mode model call `snuffy_b200.invalidate_weight_caches`."""
        ps = self._params()
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self.training or self._wcache is None or self._wcache[0] != key:
            self._wcache = (key, LayerWeights(*[p.detach() for p in ps]))
        return self._wcache[1]

    def __deepcopy__(self, memo):
        cache, self._wcache = self._wcache, None              # derived operands are not copied
        try:
            cls = self.__class__
            new = cls.__new__(cls)
            memo[id(self)] = new
            for k, v in self.__dict__.items():
                setattr(new, k, copy.deepcopy(v, memo))
        finally:
            self._wcache = cache
        return new

    def _effective_precision(self) -> str:
        return self.precision or engine.default_precision()

    def _select(self, x, c, state):
        raise NotImplementedError

    def forward(self, x, c, current_layer=None, _state=None):
        _require_cuda(x, type(self).__name__)
        if x.dim() != 3:
            raise ValueError(f"EncoderLayer expects x [B, N, d], got {tuple(x.shape)}")
        state = _state if _state is not None else {}
        from .autograd import encoder_layer_fn
        # Nothing before the attention kernel but the key projection needs the selection: on the tensor-core path it is drawn
        # on the second stream (forked here; engine.encoder_layer_forward continues there with the gather, the row map and the
        # key projection and joins before the attention), beside LayerNorm 1 and the Q|V projection of all N rows.
        side, on_side = engine.selection_stream("fp32" if self.forced_selection is not None else self._effective_precision(),
                                                x.device)
        with on_side:
            sel = self.forced_selection if self.forced_selection is not None else self._select(x, c, state)
            if sel.dim() == 1:
                sel = sel.unsqueeze(0)
            sel = sel.to(device=x.device, dtype=torch.int64).contiguous()
        engine._SEL_PENDING = side
        try:
            return encoder_layer_fn(self, x, sel, zplanes=state.get("zplanes"))
        finally:
            engine.join_pending_selection(x.device)           # no-op when the layer joined it itself
