"""GPU diagnostic for the tcgen05 GEMM (run under gpurun): checks snuffy_gemm_tc against float64 numpy on a
ladder of shapes and, when a result is wrong, tests alternative readings of the shared-memory descriptor
(LBO/SBO swapped) against the observed output so one run tells which field is mis-encoded.

    python tools/tc_diag.py            # prints one line per case; exit code 1 if any case fails
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from snuffy_b200 import ops  # noqa: E402
from snuffy_b200._lib import lib  # noqa: E402


def bf16_round(a):
    t = torch.from_numpy(a.astype(np.float32)).to(torch.bfloat16).to(torch.float32)
    return t.numpy().astype(np.float64)


def decode_planes(p, rows, K):
    """planes buffer -> (hi, lo) float64 [rows, Kpad] following csrc/common.cuh."""
    buf = p.buf.float().cpu().numpy().astype(np.float64)
    rc = p.rc
    kbn = (K + 31) // 32
    rt = (rows + rc - 1) // rc
    out = []
    for plane in range(2):
        a = buf[plane * p.stride:(plane + 1) * p.stride].reshape(rt, kbn, 4, rc, 8)
        # [rt, kb, kg, rr, e] -> [rt, rr, kb, kg, e]
        a = a.transpose(0, 3, 1, 2, 4).reshape(rt * rc, kbn * 32)
        out.append(a[:rows])
    return out


def alt_read(mat, rc, swap):
    """What the tensor core would see if LBO and SBO were swapped: element (r, k) of a [rc, 32] chunk is fetched from
    byte (r%8)*16 + (r//8)*LBO + (k//8)*SBO with the two strides exchanged."""
    if not swap:
        return mat
    rows, kp = mat.shape
    out = np.zeros_like(mat)
    for r0 in range(0, rows, rc):
        for k0 in range(0, kp, 32):
            chunk = np.zeros((4, rc, 8))
            blk = mat[r0:r0 + rc, k0:k0 + 32]
            rr = blk.shape[0]
            chunk[:, :rr, :] = blk.reshape(rr, 4, 8).transpose(1, 0, 2)
            flat = chunk.reshape(-1)            # element offset = kg*rc*8 + r*8 + e
            lbo, sbo = rc * 8, 64               # in elements (rc*16 B, 128 B)
            for ks in range(2):
                base = ks * 2 * lbo
                for r in range(min(rc, rows - r0)):
                    for kh in range(2):
                        off = base + (r % 8) * 8 + (r // 8) * lbo + kh * sbo     # swapped roles
                        if off + 8 <= flat.size:
                            out[r0 + r, k0 + ks * 16 + kh * 8:k0 + ks * 16 + kh * 8 + 8] = flat[off:off + 8]
    return out


def run_case(M, N, K, passes, seed=0, act="none", with_bias=False, with_resid=False, planes_out=False):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = torch.randn(M, K, generator=g)
    b = torch.randn(N, K, generator=g) * 0.1
    bias = torch.randn(N, generator=g) if with_bias else None
    resid = torch.randn(M, N, generator=g) if with_resid else None
    ad, bd = a.cuda(), b.cuda()
    _, ap, _ = ops.ln_rows(ad, None, None, apply_ln=False, want_planes=True)
    bp = ops.weight_planes(bd)
    out, _, op = ops.gemm_tc(ap, bp, M=M, N=N, K=K, passes=passes, bias=None if bias is None else bias.cuda(),
                             act=act, resid=None if resid is None else resid.cuda(), want_out=True,
                             want_planes=planes_out)
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)
    a64, b64 = a.numpy().astype(np.float64), b.numpy().astype(np.float64)
    if passes == 1:
        ref = bf16_round(a.numpy()) @ bf16_round(b.numpy()).T
    else:
        ref = a64 @ b64.T
    if bias is not None:
        ref = ref + bias.numpy()
    if act == "relu":
        ref = np.maximum(ref, 0)
    ref_act = ref.copy()
    if resid is not None:
        ref = ref + resid.numpy()
    scale = np.abs(ref).max() + 1e-30
    err = np.abs(got - ref).max() / scale
    tol = 2e-2 if passes == 1 else 2e-5
    ok = bool(err < tol) and np.isfinite(got).all()
    line = f"M={M:6d} N={N:5d} K={K:5d} passes={passes} act={act:4s} bias={int(with_bias)} resid={int(with_resid)} " \
           f"rel_err={err:.3e} {'OK' if ok else 'FAIL'}"
    if planes_out and op is not None:
        hi, lo = decode_planes(op, M, N)
        perr = np.abs((hi + lo)[:, :N] - ref_act).max() / (np.abs(ref_act).max() + 1e-30)
        pad_ok = np.all((hi + lo)[:, N:] == 0)
        line += f" planes_rel_err={perr:.3e} kpad_zero={pad_ok}"
        ok = ok and perr < 1e-4 and pad_ok
    print(line, flush=True)
    if not ok and passes == 1:
        ah, _ = decode_planes(ap, M, K)
        bh, _ = decode_planes(bp, N, K)
        base = ah @ bh.T
        print(f"    host decode of planes vs ref: {np.abs(base - ref).max() / scale:.3e}")
        for sa in (False, True):
            for sb in (False, True):
                alt = alt_read(ah, 128, sa) @ alt_read(bh, bp.rc, sb).T
                print(f"    hypothesis swapA={int(sa)} swapB={int(sb)}: rel diff to GPU {np.abs(alt - got).max() / scale:.3e}")
        bad = np.argwhere(np.abs(got - ref) > tol * scale)
        print(f"    {len(bad)} bad of {got.size}; first {bad[:6].tolist()}; rows hit {np.unique(bad[:, 0])[:12].tolist()} "
              f"cols hit {np.unique(bad[:, 1])[:12].tolist()}")
    return ok


def main():
    torch.manual_seed(0)
    print("device:", torch.cuda.get_device_name(0), "SMs:", lib.snuffy_sm_count(), flush=True)
    ok = True
    # plane layout round trip first (pure data movement)
    x = torch.randn(200, 96).cuda()
    _, p, _ = ops.ln_rows(x, None, None, apply_ln=False, want_planes=True)
    torch.cuda.synchronize()
    hi, lo = decode_planes(p, 200, 96)
    e = np.abs(hi + lo - x.cpu().numpy()).max()
    print(f"plane round trip: max err {e:.3e} (expect < 1e-4)", flush=True)
    ok &= e < 1e-4
    cases = [
        dict(M=128, N=128, K=32, passes=1),
        dict(M=128, N=128, K=64, passes=1),
        dict(M=128, N=256, K=32, passes=1),
        dict(M=128, N=128, K=32, passes=3),
        dict(M=256, N=512, K=512, passes=3),
        dict(M=300, N=384, K=96, passes=3, with_bias=True, act="relu"),
        dict(M=1000, N=2048, K=512, passes=3, with_bias=True, act="relu", planes_out=True),
        dict(M=1000, N=512, K=2048, passes=3, with_bias=True, with_resid=True),
        dict(M=10000, N=1024, K=512, passes=3, with_bias=True),
        dict(M=777, N=48, K=48, passes=3, with_bias=True, planes_out=True),
    ]
    for c in cases:
        ok &= run_case(**c)
    print("ALL OK" if ok else "SOME FAILED", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
