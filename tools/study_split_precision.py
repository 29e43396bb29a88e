#!/usr/bin/env python
"""CPU study (NumPy emulation, no GPU): how accurate would cheaper operand splits be for the two big GEMMs of the C3
gradient, against today's bf16x3 scheme and the 1e-4 parity budget?  Emulated per scheme: operand rounding only
(products and sums in float64 — accumulation effects are the same for all schemes and are handled by split-K).

  bf16x3            a_hi b_hi + a_lo b_hi + a_hi b_lo, hi/lo bf16                      (3 bf16-rate passes; today)
  fp16x3            the same with fp16 parts                                           (3 passes)
  fp16 + 2x mxfp8   a_h16 b_h16 + a_l8 b_h8 + a_h8 b_l8: hi part fp16 (11 bits), the two correction products with
                    e4m3 operands block-scaled by powers of two over 32 K-elements (kind::mxf8f6f4, 2x the bf16 rate)
                                                                                       (1 + 1/2 + 1/2 = 2 pass equivalents)
  fp16 + 2x e4m3 / e5m2, one scale per operand
                    the same three products, the correction operands in plain FP8 (kind::f8f6f4, no block scales): each
                    operand pre-scaled by ONE power of two by the kernel that writes it          (2 pass equivalents)
  fp16 + 2x mxfp4   correction operands e2m1 with a power-of-two scale per 32 K-elements (kind::mxf4, 4x the bf16 rate)
                                                                                       (1 + 1/4 + 1/4 = 1.5)
  int8 x 2 slices   Ozaki-style fixed point: per row (A) / column (B) scale, x = s/127 (hi + lo/254) with int8 hi, lo;
                    kind::i8 MMAs (2x the bf16 rate, exact int32 accumulation): hh + hl + lh = 1.5 pass equivalents and
                    2 bytes per operand element instead of 4; the hl + lh products carry the weight 1/254 and need their
                    own accumulator
  fp16x2 (one-sided) a_h16 (b_h16 + b_l16): only the B operand is split                (2 passes)
  bf16 / fp16 x1    single pass                                                        (1 pass)

Measured: norm-wise relative error of Z1 = X W1 (forward GEMM) and of dW1 = X^T dZ1 (the 60000-row reduction, here at
N rows) for U[0,1) data, N(0, w_scale) weights and a softmax-CE dZ1; worst case over `chains` weight draws.
Usage: python tools/study_split_precision.py [N=8192] [chains=3]
       python tools/study_split_precision.py converged [N=4096] [steps=400]     (heavy-tailed deltas of a trained net)
"""
import sys

import numpy as np


def round_bits(x, bits):
    """round to `bits` significant binary digits (no exponent limits)"""
    m, e = np.frexp(x)
    return np.ldexp(np.round(m * (1 << bits)) / (1 << bits), e)


def bf16(x):
    return round_bits(x, 8)


def fp16(x, scale=1.0):
    """fp16 with its exponent range (normal >= 2^-14, subnormal step 2^-24); `scale` applied before and undone after"""
    y = x * scale
    q = round_bits(y, 11)
    sub = np.abs(y) < 2.0 ** -14
    q = np.where(sub, np.round(y * 2.0 ** 24) / 2.0 ** 24, q)
    return q / scale


def mxfp8(x, axis):
    """e4m3 (4 significant bits, max 448, min normal 2^-6, subnormal step 2^-9) with one power-of-two scale per 32
    consecutive elements along `axis` (the K dimension of the MMA)"""
    x = np.moveaxis(x, axis, -1)
    K = x.shape[-1]
    pad = (-K) % 32
    xp = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(0, pad)])
    blk = xp.reshape(xp.shape[:-1] + (-1, 32))
    amax = np.abs(blk).max(axis=-1, keepdims=True)
    e = np.where(amax > 0, np.ceil(np.log2(np.maximum(amax, 1e-300) / 448.0)), 0.0)
    s = 2.0 ** e
    y = blk / s
    q = round_bits(y, 4)
    sub = np.abs(y) < 2.0 ** -6
    q = np.where(sub, np.round(y * 2.0 ** 9) / 2.0 ** 9, q)
    out = (q * s).reshape(xp.shape)[..., :K]
    return np.moveaxis(out, -1, axis)


def fp8_tensor(x, sig_bits=4, min_normal_exp=-6, sub_step_exp=-9, top=256.0):
    """e4m3 (default; e5m2: sig_bits=3, min_normal_exp=-14, sub_step_exp=-16) with ONE power-of-two scale for the whole
    operand, chosen so that its largest magnitude lands in (top/2, top] — what plain kind::f8f6f4 MMAs can use when the
    producing kernel pre-scales the operand (no scale-factor tiles in TMEM); the product of the two operand scales is
    then a constant that the main fp16 pass carries too and the epilogue divides out"""
    amax = float(np.abs(x).max())
    s = 2.0 ** np.floor(np.log2(top / amax)) if amax > 0 else 1.0
    y = x * s
    q = round_bits(y, sig_bits)
    sub = np.abs(y) < 2.0 ** min_normal_exp
    q = np.where(sub, np.round(y * 2.0 ** -sub_step_exp) / 2.0 ** -sub_step_exp, q)
    return q / s


def mxfp4(x, axis):
    """e2m1 (2 significant bits, max 6, min normal 1, subnormal step 0.5) with one power-of-two scale per 32 consecutive
    K-elements (kind::mxf4, 4x the bf16 rate)"""
    x = np.moveaxis(x, axis, -1)
    K = x.shape[-1]
    pad = (-K) % 32
    xp = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(0, pad)])
    blk = xp.reshape(xp.shape[:-1] + (-1, 32))
    amax = np.abs(blk).max(axis=-1, keepdims=True)
    e = np.where(amax > 0, np.ceil(np.log2(np.maximum(amax, 1e-300) / 6.0)), 0.0)
    s = 2.0 ** e
    y = blk / s
    q = round_bits(y, 2)
    q = np.where(np.abs(y) < 1.0, np.round(y * 2.0) / 2.0, q)
    out = (q * s).reshape(xp.shape)[..., :K]
    return np.moveaxis(out, -1, axis)


def int8_slices(x, k_axis, n_slices=2):
    """Ozaki-style fixed-point slices: one scale per index of the NON-contracted axis (it factors out of the dot
    product), x / s * 127 = hi + lo / 254 (+ ...), every slice an int8 in [-127, 127]; returns the float64 values of the
    slices with their weights applied (the integer products themselves are exact in the int32 accumulator)"""
    s = np.abs(x).max(axis=k_axis, keepdims=True)
    s = np.where(s > 0, s, 1.0)
    r = x / s * 127.0
    out, w = [], 1.0
    for _ in range(n_slices):
        q = np.clip(np.round(r), -127, 127)
        out.append(q * w * s / 127.0)
        r = (r - q) * 254.0
        w /= 254.0
    return out


def schemes(A, B, k_axis_a, k_axis_b, scale_b=1.0, only=None):
    """products A @ B under each scheme (or just the scheme `only`); k_axis_*: which axis of the operand is the
    contraction axis"""
    a16, b16 = fp16(A), fp16(B, scale_b)
    e5 = dict(sig_bits=3, min_normal_exp=-14, sub_step_exp=-16, top=32768.0)

    def bf16x3():
        ah, bh = bf16(A), bf16(B)
        return ah @ bh + bf16(A - ah) @ bh + ah @ bf16(B - bh)

    table = {
        "bf16x3 (today)": bf16x3,
        "bf16 x1": lambda: bf16(A) @ bf16(B),
        "fp16x3": lambda: a16 @ b16 + fp16(A - a16) @ b16 + a16 @ fp16(B - b16, scale_b * 2.0 ** 11),
        "fp16 x1": lambda: a16 @ b16,
        "fp16x2 (B split only)": lambda: a16 @ b16 + a16 @ fp16(B - b16, scale_b * 2.0 ** 11),
        "fp16 + 2x mxfp8 (2 pass-equivalents)": lambda: (a16 @ b16 + mxfp8(A - a16, k_axis_a) @ mxfp8(B, k_axis_b)
                                                         + mxfp8(A, k_axis_a) @ mxfp8(B - b16, k_axis_b)),
        "fp16 + 2x e4m3, one scale per operand (2)": lambda: (a16 @ b16 + fp8_tensor(A - a16) @ fp8_tensor(B)
                                                              + fp8_tensor(A) @ fp8_tensor(B - b16)),
        "fp16 + 2x e5m2, one scale per operand (2)": lambda: (a16 @ b16 + fp8_tensor(A - a16, **e5) @ fp8_tensor(B, **e5)
                                                              + fp8_tensor(A, **e5) @ fp8_tensor(B - b16, **e5)),
        "int8 x 2 slices, hh + hl + lh (1.5, 2 B/element)": lambda: (lambda a, b: a[0] @ b[0] + a[0] @ b[1] + a[1] @ b[0])(
            int8_slices(A, k_axis_a), int8_slices(B, k_axis_b)),
        "int8: A 2 slices, B 3 slices, 4 products (2, 2 + 3 B)": lambda: (lambda a, b: a[0] @ (b[0] + b[1] + b[2]) + a[1] @ b[0])(
            int8_slices(A, k_axis_a), int8_slices(B, k_axis_b, 3)),
        "int8 x 2 slices, all four products (2)": lambda: (lambda a, b: (a[0] + a[1]) @ (b[0] + b[1]))(
            int8_slices(A, k_axis_a), int8_slices(B, k_axis_b)),
        "fp16 + 2x mxfp4 (1.5 pass-equivalents)": lambda: (a16 @ b16 + mxfp4(A - a16, k_axis_a) @ mxfp4(B, k_axis_b)
                                                           + mxfp4(A, k_axis_a) @ mxfp4(B - b16, k_axis_b)),
    }
    if only is not None:
        return {only: table[only]()}
    return {k: f() for k, f in table.items()}


def product(A, B, name, k_axis_a, k_axis_b, scale_b=1.0):
    return schemes(A, B, k_axis_a, k_axis_b, scale_b, only=name)[name]


def full_gradient(X, y, W1, b1, W2, b2, name, mask=None):
    """loss and flat gradient [dW1, db1, dW2, db2] of the mean sparse-CE loss with the three big GEMMs under `name`
    (None: exact float64); layer 2 stays exact, as in the fused epilogue (fp32 there).  `mask`: relu'(z1) taken from the
    exact forward pass — relu' is discontinuous, so ANY rounding of z1 (float32 itself included) flips the units that sit
    within that rounding of zero and the flips, not the scheme, then dominate the comparison (the GPU parity tests move
    the test points off the kinks for the same reason)"""
    N, C = X.shape[0], W2.shape[1]
    mm = (lambda A, B, ka, kb, sb=1.0: A @ B) if name is None else (lambda A, B, ka, kb, sb=1.0: product(A, B, name, ka, kb, sb))
    Z1 = mm(X, W1, 1, 0) + b1
    A1 = np.maximum(Z1, 0)
    Z2 = A1 @ W2 + b2
    m = Z2.max(1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(Z2 - m).sum(1))
    loss = float((lse - Z2[np.arange(N), y]).mean())
    P = np.exp(Z2 - lse[:, None])
    dZ2 = (P - np.eye(C)[y]) / N
    relu_mask = (Z1 > 0) if mask is None else mask
    A1 = A1 * relu_mask if mask is not None else A1
    dZ1 = (dZ2 @ W2.T) * relu_mask
    sb = float(2 ** int(np.ceil(np.log2(N))))
    dW1 = mm(X.T.copy(), dZ1, 1, 0, sb)
    dW2 = mm(A1.T.copy(), dZ2, 1, 0, sb)
    return loss, np.concatenate([dW1.ravel(), dZ1.sum(0), dW2.ravel(), dZ2.sum(0)]), relu_mask


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def converged_case(N=4096, steps=400):
    """Heavy-tailed deltas: a student trained on teacher labels until most rows are confidently right, so that a few
    rows carry dZ1.  Prints the dW1 = X^T dZ1 error of the schemes whose operand scale is per tensor / per row — the
    case the random-weight study cannot show — and, for the int8 slices, what an a-priori scale (2 max_c |W2[h, c]| / N,
    known before the deltas exist) costs against the true row maximum."""
    D, H, C = 784, 256, 10
    rng = np.random.default_rng(1)
    X = rng.random((N, D))
    y = (X @ rng.normal(0, 1, (D, C))).argmax(1)
    W1, b1 = rng.normal(0, 0.05, (D, H)), np.zeros(H)
    W2, b2 = rng.normal(0, 0.05, (H, C)), np.zeros(C)
    m = [np.zeros_like(v) for v in (W1, b1, W2, b2)]
    v2 = [np.zeros_like(v) for v in (W1, b1, W2, b2)]
    for t in range(1, steps + 1):
        A1 = np.maximum(X @ W1 + b1, 0)
        Z2 = A1 @ W2 + b2
        P = np.exp(Z2 - Z2.max(1, keepdims=True))
        P /= P.sum(1, keepdims=True)
        dZ2 = (P - np.eye(C)[y]) / N
        dZ1 = (dZ2 @ W2.T) * (A1 > 0)
        grads = (X.T @ dZ1, dZ1.sum(0), A1.T @ dZ2, dZ2.sum(0))
        for i, (w, g) in enumerate(zip((W1, b1, W2, b2), grads)):
            m[i] = 0.9 * m[i] + 0.1 * g
            v2[i] = 0.999 * v2[i] + 0.001 * g * g
            w -= 3e-3 * (m[i] / (1 - 0.9 ** t)) / (np.sqrt(v2[i] / (1 - 0.999 ** t)) + 1e-8)
    acc = float((Z2.argmax(1) == y).mean())
    rowmax = np.abs(dZ1).max(0)
    live = rowmax > 0                                             # dead relu units have an all-zero dZ1 column
    peak = rowmax[live] / np.sqrt((dZ1[:, live] ** 2).mean(0))
    print("converged case: N = %d, %d Adam steps, train accuracy %.3f, loss %.4f; %d of %d hidden units live; |dZ1| max / rms "
          "per live unit: median %.0f, max %.0f" % (N, steps, acc, float(-np.log(P[np.arange(N), y]).mean()), int(live.sum()), H,
                                                    float(np.median(peak)), float(peak.max())))
    want = X.T @ dZ1
    sb = float(2 ** int(np.ceil(np.log2(N))))
    for name in ("bf16x3 (today)", "fp16 + 2x e4m3, one scale per operand (2)", "fp16 + 2x mxfp8 (2 pass-equivalents)",
                 "int8 x 2 slices, hh + hl + lh (1.5, 2 B/element)"):
        print("  dW1 = X^T dZ1    %-50s %.2e" % (name, rel(product(X.T.copy(), dZ1, name, 1, 0, sb), want)))
    # int8 slices of dZ1 with the a-priori bound instead of the true row maximum
    bound = 2.0 * np.abs(W2).max(1) / N                           # |dZ1[n, h]| <= sum_c |dZ2[n, c]| |W2[h, c]| <= 2 max_c |W2[h, c]| / N
    assert (rowmax <= bound * (1 + 1e-12)).all()
    xs = int8_slices(X.T.copy(), 1)

    def slices_with_scale(x, s, n=2):
        r, out, w = x / s * 127.0, [], 1.0
        for _ in range(n):
            q = np.clip(np.round(r), -127, 127)
            out.append(q * w * s / 127.0)
            r, w = (r - q) * 254.0, w / 254.0
        return out
    for n in (2, 3):
        d = slices_with_scale(dZ1, bound[None, :], n)
        got = xs[0] @ d[0] + xs[0] @ d[1] + xs[1] @ d[0] + (xs[0] @ d[2] + xs[1] @ d[1] if n == 3 else 0.0)
        print("  dW1 = X^T dZ1    int8 slices, a-priori scale (bound / true max: median %.1f), %d dZ1 slices   %.2e"
              % (float(np.median(bound[live] / rowmax[live])), n, rel(got, want)))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "converged":
        return converged_case(*(int(a) for a in sys.argv[2:4]))
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    chains = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    D, H, C = 784, 256, 10
    rng = np.random.default_rng(0)
    X = rng.random((N, D))
    y = rng.integers(0, C, N)
    worst = {}
    for w_scale in (0.05, 1.0):
        for _ in range(chains):
            W1, b1 = rng.normal(0, w_scale, (D, H)), rng.normal(0, w_scale, H)
            W2, b2 = rng.normal(0, w_scale, (H, C)), rng.normal(0, w_scale, C)
            Z1 = X @ W1
            for name, z in schemes(X, W1, 1, 0).items():
                worst[("Z1 = X W1", w_scale, name)] = max(worst.get(("Z1 = X W1", w_scale, name), 0), rel(z, Z1))
            A1 = np.maximum(Z1 + b1, 0)
            Z2 = A1 @ W2 + b2
            P = np.exp(Z2 - Z2.max(1, keepdims=True))
            P /= P.sum(1, keepdims=True)
            dZ2 = (P - np.eye(C)[y]) / N
            dZ1 = (dZ2 @ W2.T) * (Z1 + b1 > 0)
            dW1 = X.T @ dZ1
            # dZ1 carries the 1/N of the mean loss: the fp16 parts are taken of N * dZ1 (a power-of-two scale in practice)
            for name, g in schemes(X.T.copy(), dZ1, 1, 0, scale_b=float(2 ** int(np.ceil(np.log2(N))))).items():
                worst[("dW1 = X^T dZ1", w_scale, name)] = max(worst.get(("dW1 = X^T dZ1", w_scale, name), 0), rel(g, dW1))
    # whole gradient with all three GEMMs under one scheme (forward error propagates into the deltas and relu masks)
    names = ["bf16x3 (today)", "fp16x3", "fp16 + 2x mxfp8 (2 pass-equivalents)", "fp16 + 2x e4m3, one scale per operand (2)",
             "fp16 + 2x e5m2, one scale per operand (2)", "fp16 + 2x mxfp4 (1.5 pass-equivalents)", "int8 x 2 slices, hh + hl + lh (1.5, 2 B/element)",
             "int8 x 2 slices, all four products (2)", "int8: A 2 slices, B 3 slices, 4 products (2, 2 + 3 B)", "fp16x2 (B split only)",
             "fp16 x1", "bf16 x1"]
    for w_scale in (0.05, 1.0):
        for _ in range(chains):
            W1, b1 = rng.normal(0, w_scale, (D, H)), rng.normal(0, w_scale, H)
            W2, b2 = rng.normal(0, w_scale, (H, C)), rng.normal(0, w_scale, C)
            l0, g0, mask0 = full_gradient(X, y, W1, b1, W2, b2, None)
            for name in names:
                l, g, mask = full_gradient(X, y, W1, b1, W2, b2, name)
                key = ("gradient, relu flips", w_scale, name)        # fraction of hidden activations whose mask flipped
                worst[key] = max(worst.get(key, 0), float((mask != mask0).mean()))
                l, g, _ = full_gradient(X, y, W1, b1, W2, b2, name, mask0)
                key = ("whole gradient", w_scale, name)
                worst[key] = max(worst.get(key, 0), rel(g, g0))
                key = ("loss", w_scale, name)
                worst[key] = max(worst.get(key, 0), abs(l - l0) / abs(l0))
    print("rows N = %d, %d weight draws per scale; norm-wise relative error (worst case); parity budget 1e-4" % (N, chains))
    for (what, ws, name), v in sorted(worst.items(), key=lambda kv: (kv[0][0], kv[0][1], kv[1])):
        print("  %-16s weights ~ N(0, %-4g)  %-54s %.2e" % (what, ws, name, v))


if __name__ == "__main__":
    main()
