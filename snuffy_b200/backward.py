"""Backward passes of the drop-in modules (train.py:259 ``loss.backward()``).

Every ``torch.autograd.Function`` here runs its forward AND backward in ``libsnuffy_b200.so``: PyTorch's autograd
engine only sequences them and owns the buffers.  The gradient math follows SURVEY.md Appendix A differentiated by
hand; ``tests/test_gpu_backward.py`` checks it against autograd of the CPU port of the reference.

Per encoder layer (snuffy.py:126-157), with y = x-with-rows-S-replaced, g the upstream gradient of x_next:
  FFN       x_next = y + D2(W2 a + b2), a = Dff(act(h)), h = W1 LN2(y) + b1
  attention xs_new = xs + D1(Wo O + bo), O_j = D_attn(softmax_keys(Q_j Kp_j^T / sqrt(dk)))^T V_j,
            Q|V = Wqv LN1(x) + bqv (all N rows), Kp = Wk xs + bk, xs = x[S] raw rows
  scatter   y[S] = xs_new, y[not S] = x
"""
from __future__ import annotations

import torch

from . import engine, ops


#: attention backward: "fused" (default: one tcgen05 kernel, csrc/attn_bwd_tc.cu, where the shape allows), "tc" (dense tcgen05
#: GEMMs over block-diagonal head operands + row kernels), "simt" (fp32 SIMT products)
ATTN_BWD = __import__("os").environ.get("SNUFFY_B200_ATTN_BWD", "fused")
ATTN_BWD_TC = ATTN_BWD != "simt"


#: the weight- and bias-gradient products are off the critical chain (nothing in the backward pass reads them): issued on a
#: second stream they fill the SMs the dX products leave idle in their last, partial wave (one bag = 79 row tiles: 158 tiles
#: on 148 SMs) and the small ones no longer serialise between the big ones.  "0" = everything on one stream.
DW_SIDE_STREAM = __import__("os").environ.get("SNUFFY_B200_DW_SIDE_STREAM", "1") != "0"
_DW_STREAMS = {}


def _dw_stream(device: torch.device) -> torch.cuda.Stream:
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _DW_STREAMS:
        _DW_STREAMS[key] = torch.cuda.Stream(device=device)
    return _DW_STREAMS[key]


def _flat(t: torch.Tensor, d: int) -> torch.Tensor:
    return t.contiguous().view(-1, d)


# ------------------------------------------------------------------ instance scores (snuffy.py:39-41)
class ScoresFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, weight, bias):
        ctx.save_for_backward(feats, weight)
        ctx.has_bias = bias is not None
        return ops.scores(feats.detach(), weight.detach(), None if bias is None else bias.detach())

    @staticmethod
    def backward(ctx, dc):
        feats, weight = ctx.saved_tensors
        d = feats.shape[-1]
        C = weight.shape[0]
        x2 = _flat(feats.detach(), d)
        dc2 = dc.contiguous().view(-1, C)
        d_feats = d_w = d_b = None
        if ctx.needs_input_grad[1]:
            d_w = ops.colsum(x2, dc2)                                    # [C, d]: only the arg-max rows are non-zero
        if ctx.has_bias and ctx.needs_input_grad[2]:
            d_b = ops.colsum(dc2).view(C)
        if ctx.needs_input_grad[0]:
            d_feats = ops.matmul_nn(dc2, weight.detach()).view(feats.shape)
        return d_feats, d_w, d_b


# ------------------------------------------------------------------ LayerNorm (Encoder.forward's final norm, snuffy.py:86)
class LayerNormFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta):
        d = x.shape[-1]
        out, _, stats = ops.ln_rows(_flat(x.detach(), d), gamma.detach(), beta.detach(), want_f32=True, want_stats=True)
        ctx.save_for_backward(x, gamma, stats)
        return out.view(x.shape)

    @staticmethod
    def backward(ctx, g):
        x, gamma, stats = ctx.saved_tensors
        d = x.shape[-1]
        dx, dg, db = ops.ln_rows_bwd(_flat(x.detach(), d), stats, gamma.detach(), dy=_flat(g, d),
                                     want_dx=ctx.needs_input_grad[0])
        return (dx.view(x.shape) if dx is not None else None), dg, db


# ------------------------------------------------------------------ final LN + mean + head (snuffy.py:86,71)
class LnMeanHeadFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, w_head, b_head):
        bag, stats, pooled = ops.ln_mean_head(x.detach(), gamma.detach(), beta.detach(), w_head.detach(),
                                              None if b_head is None else b_head.detach(), want_stats=True, want_pooled=True)
        ctx.save_for_backward(x, gamma, w_head, stats, pooled)
        ctx.has_bias = b_head is not None
        return bag

    @staticmethod
    def backward(ctx, dbag):
        x, gamma, w_head, stats, pooled = ctx.saved_tensors
        B, N, d = x.shape
        C = w_head.shape[0]
        dbag = dbag.contiguous().view(B, C)
        d_wh = ops.matmul_tn(dbag, pooled)                               # [C, d]
        d_bh = ops.colsum(dbag).view(C) if ctx.has_bias else None
        dpooled = ops.matmul_nn(dbag, w_head.detach())                   # [B, d]; every row of bag b receives dpooled[b] / N
        dx, dg, db = ops.ln_rows_bwd(_flat(x.detach(), d), stats, gamma.detach(), dy_bcast=dpooled, rows_per_bag=N,
                                     bscale=1.0 / N, want_dx=ctx.needs_input_grad[0])
        return (dx.view(B, N, d) if dx is not None else None), dg, db, d_wh, d_bh


# ------------------------------------------------------------------ one encoder layer (snuffy.py:126-157)
class EncoderLayerFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, layer, x, sel, *params):
        B, N, d = x.shape
        training = layer.training
        w = layer.layer_weights()
        xc = x.detach().contiguous().view(B * N, d)
        x_next, probs, tape = engine.encoder_layer_forward(
            xc, B, N, sel, w, layer.self_attn.h, layer.feed_forward.activation_name, layer._effective_precision(),
            want_probs=layer.return_attn, save=True,
            attn_dropout=float(layer.self_attn.dropout.p) if training else 0.0,
            enc_dropout=float(layer.sublayer[0].dropout.p) if training else 0.0,
            ff_dropout=float(layer.feed_forward.dropout.p) if training else 0.0)
        ctx.tape, ctx.w = tape, w
        ctx.meta = (B, N, d, layer.self_attn.h, layer.feed_forward.activation_name, layer._effective_precision())
        x_next = x_next.view(B, N, d)
        if probs is None:
            return x_next, None
        ctx.mark_non_differentiable(probs)       # nobody differentiates through A
        return x_next, probs

    @staticmethod
    def backward(ctx, g_out, _g_probs=None):
        t, w = ctx.tape, ctx.w
        B, N, d, heads, act, precision = ctx.meta
        rows, ksel = B * N, t.sel.shape[1]
        dff = w.w1.shape[0]
        need = ctx.needs_input_grad
        g = _flat(g_out, d)
        # The six contractions over all N rows run on tcgen05 (split-bf16, like the forward) unless the layer is in the
        # exact-fp32 mode; the [Ksel, d]-sized ones and the attention backward stay fp32 SIMT.
        tc = precision != "fp32" and engine.tc_supported(d) and dff % 8 == 0
        passes = 1 if precision == "bf16x1" else 3
        if tc:
            if not w.prepare_train():
                w.prepare(precision)
                w.prepare_backward()
            rc_d, rc_ff = ops._block_n(d), ops._block_n(dff)
        by_rows = tc and engine.DW_BY_ROWS and t.a_planes is not None
        side = _dw_stream(g.device) if by_rows and DW_SIDE_STREAM else None
        keep = []                 # operands of side-stream products stay referenced until the join (the allocator would hand
                                  # their memory to later main-stream kernels the moment they are dropped)

        def off_chain(fn, *operands):
            if side is None:
                return fn()
            keep.extend(operands)
            side.wait_stream(torch.cuda.current_stream(g.device))
            with torch.cuda.stream(side):
                return fn()

        def dx_gemm(dy, w_f32, wt_planes, n_in):                # dX = dY . W        (W = nn.Linear.weight [out, in])
            if not tc:
                return ops.matmul_nn(dy, w_f32)
            _, ap, _ = ops.ln_rows(dy, None, None, apply_ln=False, want_planes=True)
            out, _, _ = ops.gemm_tc(ap, wt_planes, M=dy.shape[0], N=n_in, K=dy.shape[1], passes=passes)
            return out

        # ---- feed-forward sub-layer
        gf = ops.act_bwd(None, g, drop=t.drop_enc2)[0] if t.drop_enc2[0] > 0 else g
        d_b2 = off_chain(lambda: ops.colsum(gf).view(-1), gf)
        dh_planes = None
        if tc:
            # dh = (gf W2) * act'(h_pre) * mask in the epilogue of the product, as the next product's operand planes
            _, gfp, _ = ops.ln_rows(gf, None, None, apply_ln=False, want_planes=True)
            if by_rows and ops.gemm_tc_splitk_rows_supported(d, dff) and ops.gemm_tc_splitk_rows_supported(dff, d):
                # weight gradients contract over the rows: straight from the row planes both sides already exist as (the
                # forward's GEMM operands, this pass's dX operands); db1 = column sums of dh from the same epilogue, so dh is
                # never written as fp32
                dh = None
                if t.h_pre is None:       # ReLU: the gate is read off the forward's activated planes (no h_pre, no dropout draw)
                    _, dh_planes, d_b1 = ops.gemm_tc_relugrad(gfp, w.w2t_planes, t.a_planes, t.drop_ff[0], M=rows, N=dff, K=d,
                                                              passes=passes, want_colsum=True)
                else:
                    _, dh_planes, d_b1 = ops.gemm_tc_actgrad(gfp, w.w2t_planes, t.h_pre, act, M=rows, N=dff, K=d, passes=passes,
                                                             drop=t.drop_ff, want_out=False, want_colsum=True)
                d_w2 = off_chain(lambda: ops.gemm_tc_splitk_rows(gfp, t.a_planes, M=d, N=dff, R=rows, passes=passes),
                                 gfp, t.a_planes)                                             # gf^T . dropout(act(h_pre))
                d_w1 = off_chain(lambda: ops.gemm_tc_splitk_rows(dh_planes, t.u2_planes, M=dff, N=d, R=rows, passes=passes),
                                 dh_planes, t.u2_planes)                                      # dh^T . LN2(y)
            else:
                dh, dh_planes = ops.gemm_tc_actgrad(gfp, w.w2t_planes, t.h_pre, act, M=rows, N=dff, K=d, passes=passes,
                                                    drop=t.drop_ff)
                d_w2 = ops.gemm_tc_splitk(ops.planes_t(gf, 128), ops.planes_t(t.h_pre, rc_ff, mode=2, act=act, drop=t.drop_ff),
                                          M=d, N=dff, K=rows, passes=passes)
                d_w1 = ops.gemm_tc_splitk(ops.planes_t(dh, 128),
                                          ops.planes_t(t.x_in, rc_d, mode=1, stats=t.ln2_stats, gamma=w.g2, beta=w.be2,
                                                       row_map=t.row_map, alt=t.xs_new),
                                          M=dff, N=d, K=rows, passes=passes)
                d_b1 = ops.colsum(dh).view(-1)
            del gfp
        else:
            da = dx_gemm(gf, w.w2, w.w2t_planes, dff)                              # [rows, dff]
            dh, a = ops.act_bwd(t.h_pre, da, act, t.drop_ff, want_dh=True, want_a=True)
            del da
            d_w2 = ops.matmul_tn(gf, a)                                            # [d, dff]
            del a
            u2, _, _ = ops.ln_rows(t.x_in, w.g2, w.be2, row_map=t.row_map, alt=t.xs_new, want_f32=True)
            d_w1 = ops.matmul_tn(dh, u2)                                           # [dff, d]
            del u2
            d_b1 = ops.colsum(dh).view(-1)
        if dh_planes is not None:
            du2, _, _ = ops.gemm_tc(dh_planes, w.w1t_planes, M=rows, N=d, K=dff, passes=passes)
        else:
            du2 = dx_gemm(dh, w.w1, w.w1t_planes, d)                               # [rows, d]
        del dh, dh_planes
        # the dgamma / dbeta partial sums are folded off the chain
        dy, fold2 = ops.ln_rows_bwd(t.x_in, t.ln2_stats, w.g2, dy=du2, row_map=t.row_map, alt=t.xs_new, add=g, defer_fold=True)
        d_g2, d_be2 = off_chain(fold2, fold2.partials)
        del du2

        # ---- attention sub-layer: the selected rows of y are xs_new = xs + D1(Wo O + bo)
        dxs_new = ops.gather_rows(dy.view(B, N, d), t.sel).view(B * ksel, d)
        dz = ops.act_bwd(None, dxs_new, drop=t.drop_enc1)[0] if t.drop_enc1[0] > 0 else dxs_new
        d_bo = off_chain(lambda: ops.colsum(dz).view(-1), dz)
        if tc:
            d_wo = off_chain(lambda: ops.gemm_tc_splitk(ops.planes_t(dz, 128), ops.planes_t(t.o, rc_d), M=d, N=d, K=B * ksel,
                                                        passes=passes), dz, t.o)
        else:
            d_wo = ops.matmul_tn(dz, t.o)
        d_o = dx_gemm(dz, w.wo, w.wot_planes, d)                                   # [B*Ksel, d]
        q, v = t.qv[:, :d], t.qv[:, d:]
        if tc and t.qvp is not None and ATTN_BWD == "fused" and ops.sparse_attn_bwd_fused_supported(B, N, ksel, heads, d) \
                and (t.drop[0] == 0 or t.attn_mask is not None):
            dq, dv, dkp, dqv = ops.sparse_attn_bwd_fused(t.qvp, t.kp, d_o, t.attn_stats, B, N, ksel, heads, d, t.drop[0],
                                                         t.attn_mask)
        elif tc and t.qvp is not None and ATTN_BWD_TC and ops.sparse_attn_bwd_tc_supported(B, N, ksel, heads, d):
            dq, dv, dkp, dqv = ops.sparse_attn_bwd_tc(t.qvp, t.qv, t.kp, d_o, t.attn_stats, B, N, ksel, heads, d, t.drop, passes)
        else:
            dq, dv, dkp, dqv = ops.sparse_attn_bwd(q, v, t.kp, d_o, t.attn_stats, B, N, ksel, heads, t.drop)
        d_bk = off_chain(lambda: ops.colsum(dkp).view(-1), dkp)                    # == 0 up to rounding (App. B-16)
        if tc:
            d_wk = off_chain(lambda: ops.gemm_tc_splitk(ops.planes_t(dkp, 128), ops.planes_t(t.xs, rc_d), M=d, N=d, K=B * ksel,
                                                        passes=passes), dkp, t.xs)
        else:
            d_wk = ops.matmul_tn(dkp, t.xs)
        d_bqv = off_chain(lambda: ops.colsum(dqv).view(-1), dqv)
        d_bq, d_bv = d_bqv[:d], d_bqv[d:]
        if tc:
            _, dqvp, _ = ops.ln_rows(dqv, None, None, apply_ln=False, want_planes=True)
            if by_rows and ops.gemm_tc_splitk_rows_supported(2 * d, d):
                d_wqv = off_chain(lambda: ops.gemm_tc_splitk_rows(dqvp, t.u1_planes, M=2 * d, N=d, R=rows, passes=passes),
                                  dqvp, t.u1_planes)                                          # [dQ | dV]^T . LN1(x)
            else:
                d_wqv = ops.gemm_tc_splitk(ops.planes_t(dqv, 128),
                                           ops.planes_t(t.x_in, rc_d, mode=1, stats=t.ln1_stats, gamma=w.g1, beta=w.be1),
                                           M=2 * d, N=d, K=rows, passes=passes)
            d_wq, d_wv = d_wqv[:d], d_wqv[d:]
            du1, _, _ = ops.gemm_tc(dqvp, w.wqvt_planes, M=rows, N=d, K=2 * d, passes=passes)   # dQ Wq + dV Wv
            del dqvp
        else:
            u1, _, _ = ops.ln_rows(t.x_in, w.g1, w.be1, want_f32=True)
            d_wq, d_wv = ops.matmul_tn(dq, u1), ops.matmul_tn(dv, u1)
            del u1
            du1 = ops.matmul_nn(dv, w.wv, resid=ops.matmul_nn(dq, w.wq))           # [rows, d]
        dx, fold1 = ops.ln_rows_bwd(t.x_in, t.ln1_stats, w.g1, dy=du1, add=dy, want_dx=need[1], defer_fold=True)
        d_g1, d_be1 = off_chain(fold1, fold1.partials)
        if dx is not None:
            # raw selected rows also feed the key projection: dx[S] += dKp Wk  (the xs residual is already in `add`)
            ops.scatter_add_rows(dx.view(B, N, d), t.sel, dx_gemm(dkp, w.wk, w.wkt_planes, d))
            dx = dx.view(B, N, d)
        if side is not None and keep:
            torch.cuda.current_stream(g.device).wait_stream(side)   # join: the weight gradients are complete from here on
        keep.clear()
        ctx.tape = None
        return (None, dx, None, d_wq, d_bq, d_wk, d_bk, d_wv, d_bv, d_wo, d_bo, d_w1, d_b1, d_w2, d_b2,
                d_g1, d_be1, d_g2, d_be2)


# ------------------------------------------------------------------ stand-alone pieces of the class surface
# MultiHeadedAttention.forward (snuffy.py:183-205), PositionwiseFeedForward.forward (224-225), attention (160-168) and
# SublayerConnection.forward (100-110) called on their own (the encoder layer uses the fused path above): exact-fp32 kernels.
class LinearFunction(torch.autograd.Function):
    """y = dropout(act(x W^T + b)) for x [rows, K], W [N, K]."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, drop):
        x2 = x.detach().contiguous()
        need_pre = act != "none" and any(ctx.needs_input_grad[:3])
        res = ops.gemm_f32(x2, weight.detach(), M=x2.shape[0], N=weight.shape[0], K=x2.shape[1],
                           bias=None if bias is None else bias.detach(), act=act, want_preact=need_pre, drop=drop)
        out, pre = res if need_pre else (res, None)
        ctx.save_for_backward(x2, weight.detach(), pre)
        ctx.act, ctx.drop, ctx.has_bias = act, drop, bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        x2, weight, pre = ctx.saved_tensors
        g = g.contiguous()
        if pre is not None or ctx.drop[0] > 0:
            g = ops.act_bwd(pre, g, ctx.act if pre is not None else "none", ctx.drop)[0]
        dx = ops.matmul_nn(g, weight) if ctx.needs_input_grad[0] else None
        dw = ops.matmul_tn(g, x2) if ctx.needs_input_grad[1] else None
        db = ops.colsum(g).view(-1) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db, None, None


class SparseAttnFunction(torch.autograd.Function):
    """(O [B*Ksel, d], P [B, h, N, Ksel]) of snuffy.py:160-168 for q, v [B*N, d], kp [B*Ksel, d]; P carries no gradient."""

    @staticmethod
    def forward(ctx, q, v, kp, B, N, Ksel, h, drop):
        q2, v2, k2 = q.detach().contiguous(), v.detach().contiguous(), kp.detach().contiguous()
        o, probs, stats = ops.sparse_attn(q2, v2, k2, B, N, Ksel, h, want_probs=True, want_stats=True, dropout_p=drop[0],
                                          seed=drop[1], offset=drop[2])
        ctx.save_for_backward(q2, v2, k2, stats)
        ctx.meta = (B, N, Ksel, h, drop)
        ctx.mark_non_differentiable(probs)
        return o, probs

    @staticmethod
    def backward(ctx, d_o, _dp=None):
        q2, v2, k2, stats = ctx.saved_tensors
        B, N, Ksel, h, drop = ctx.meta
        dq, dv, dkp, _ = ops.sparse_attn_bwd(q2, v2, k2, d_o.contiguous(), stats, B, N, Ksel, h, drop)
        return dq.contiguous(), dv.contiguous(), dkp, None, None, None, None, None


class GatherRowsFunction(torch.autograd.Function):
    """x [B, N, d], idx [B, K] -> x[b, idx[b]] (snuffy.py:103-106); backward scatters the rows back."""

    @staticmethod
    def forward(ctx, x, idx):
        ctx.save_for_backward(idx)
        ctx.shape = x.shape
        return ops.gather_rows(x.detach().contiguous(), idx)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        dx = torch.zeros(ctx.shape, dtype=torch.float32, device=g.device)
        ops.scatter_add_rows(dx, idx, g.contiguous().view(-1, ctx.shape[-1]))
        return dx, None


class ResidualDropoutFunction(torch.autograd.Function):
    """x + dropout(y)  (snuffy.py:108,110)."""

    @staticmethod
    def forward(ctx, x, y, drop):
        ctx.drop = drop
        return ops.residual_dropout(x.detach().contiguous(), y.detach().contiguous(), drop)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        dy = ops.act_bwd(None, g, drop=ctx.drop)[0] if ctx.drop[0] > 0 else g
        return g, dy, None


# ------------------------------------------------------------------ DSMIL bag classifier (dsmil.py:72-92)
def _dsmil_params(mod):
    import torch.nn as nn
    if isinstance(mod.q, nn.Sequential):
        qp = [mod.q[0].weight, mod.q[0].bias, mod.q[2].weight, mod.q[2].bias]
    else:
        qp = [mod.q.weight, mod.q.bias]
    vp = [mod.v[1].weight, mod.v[1].bias] if isinstance(mod.v, nn.Sequential) else []
    return qp, vp, [mod.fcc.weight, mod.fcc.bias]


class _QMlp:
    """q(.) of dsmil.py:56-60 with saved pre-activations, and its backward."""

    def __init__(self, qp):
        self.p = [t.detach() for t in qp]
        self.nonlinear = len(qp) == 4

    def forward(self, inp):
        if not self.nonlinear:
            return ops.linear_f32(inp, self.p[0], self.p[1]), (inp,)
        M = inp.shape[0]
        a1, h1 = ops.gemm_f32(inp, self.p[0], M=M, N=self.p[0].shape[0], K=inp.shape[1], bias=self.p[1], act="relu",
                              want_preact=True)
        out, h2 = ops.gemm_f32(a1, self.p[2], M=M, N=self.p[2].shape[0], K=a1.shape[1], bias=self.p[3], act="tanh",
                               want_preact=True)
        return out, (inp, h1, a1, h2)

    def backward(self, saved, d_out, want_dinp):
        """-> (grads of the q parameters in order, d_inp or None)"""
        if not self.nonlinear:
            (inp,) = saved
            d_inp = ops.matmul_nn(d_out, self.p[0]) if want_dinp else None
            return [ops.matmul_tn(d_out, inp), ops.colsum(d_out).view(-1)], d_inp
        inp, h1, a1, h2 = saved
        dh2, _ = ops.act_bwd(h2, d_out, "tanh")
        dh1, _ = ops.act_bwd(h1, ops.matmul_nn(dh2, self.p[2]), "relu")
        d_inp = ops.matmul_nn(dh1, self.p[0]) if want_dinp else None
        return [ops.matmul_tn(dh1, inp), ops.colsum(dh1).view(-1), ops.matmul_tn(dh2, a1), ops.colsum(dh2).view(-1)], d_inp


class DsmilBClassifierFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, feats, c, *params):
        import torch.nn as nn
        qp, vp, fp = _dsmil_params(mod)
        feats_d = feats.detach().contiguous()
        n, d = feats_d.shape
        ncls = c.shape[1]
        drop = (0.0, 0, 0)
        v_saved = None
        if vp:
            p = float(mod.v[0].p) if mod.training else 0.0
            if p > 0:
                drop = (p,) + engine._RANDOM.next()
            a0 = ops.act_bwd(None, feats_d, drop=drop)[0] if p > 0 else feats_d      # nn.Dropout on the input (dsmil.py:64)
            v, hv = ops.gemm_f32(a0, vp[0].detach(), M=n, N=d, K=d, bias=vp[1].detach(), act="relu", want_preact=True)
            v_saved = (a0, hv)
        else:
            v = feats_d
        mlp = _QMlp(qp)
        q, q_saved = mlp.forward(feats_d)
        crit = ops.select_topk(c.detach().contiguous().view(1, n, ncls), 1).view(1, ncls)   # dsmil.py:78-81
        q_max = ops.gather_rows(q.view(1, n, q.shape[1]), crit).view(ncls, q.shape[1])    # = q(feats[crit]): row crit of Q
        qm_saved = None
        a, bm, logits = ops.dsmil_pool(q, q_max, v, fp[0].detach(), fp[1].detach())
        ctx.mlp, ctx.saved = mlp, (feats_d, v, v_saved, q, q_saved, q_max, qm_saved, crit, a, bm, drop)
        ctx.nq, ctx.nv = len(qp), len(vp)
        ctx.vw = vp[0].detach() if vp else None
        ctx.fw = fp[0].detach()
        ctx.mark_non_differentiable(crit)
        return logits.view(1, -1), a, bm.view(1, ncls, d)

    @staticmethod
    def backward(ctx, d_logits, d_a_up, d_b_up):
        feats, v, v_saved, q, q_saved, q_max, qm_saved, crit, a, bm, drop = ctx.saved
        n, d = feats.shape
        ncls = a.shape[1]
        dq_dim = q.shape[1]
        want_dfeats = ctx.needs_input_grad[1]
        fw = ctx.fw.view(ncls, ncls * d)
        dl = (d_logits if d_logits is not None else torch.zeros(1, ncls, device=feats.device)).contiguous().view(1, ncls)
        d_fw = ops.matmul_tn(dl, bm.view(1, ncls * d)).view(ncls, ncls, d)           # Conv1d(C, C, kernel = d) weight
        d_fb = dl.view(ncls).clone()
        d_bm = ops.matmul_nn(dl, fw).view(ncls, d)
        if d_b_up is not None:
            d_bm = d_bm + d_b_up.view(ncls, d)
        d_v = ops.matmul_nn(a, d_bm)                                                 # [N, d]
        d_a = ops.matmul_nt(v, d_bm)                                                 # [N, C]
        if d_a_up is not None:
            d_a = d_a + d_a_up
        d_s = ops.softmax_cols_bwd(a, d_a, float(torch.sqrt(torch.tensor(dq_dim, dtype=torch.float32))))
        d_q = ops.matmul_nn(d_s, q_max)                                              # [N, 128]
        d_qmax = ops.matmul_tn(d_s, q)                                               # [C, 128]
        ops.scatter_add_rows(d_q.view(1, n, dq_dim), crit, d_qmax)                   # q_max is row crit of Q
        q_grads, d_feats = ctx.mlp.backward(q_saved, d_q, want_dfeats)
        v_grads = []
        if ctx.nv:
            a0, hv = v_saved
            dhv, _ = ops.act_bwd(hv, d_v, "relu")
            v_grads = [ops.matmul_tn(dhv, a0), ops.colsum(dhv).view(-1)]
            if want_dfeats:
                dfv = ops.matmul_nn(dhv, ctx.vw)
                d_feats = d_feats + (ops.act_bwd(None, dfv, drop=drop)[0] if drop[0] > 0 else dfv)
        elif want_dfeats:
            d_feats = d_feats + d_v
        ctx.saved = None
        return (None, d_feats if want_dfeats else None, None, *q_grads, *v_grads, d_fw, d_fb)


def dsmil_bclassifier_fn(mod, feats, c):
    qp, vp, fp = _dsmil_params(mod)
    return DsmilBClassifierFunction.apply(mod, feats, c, *qp, *vp, *fp)
