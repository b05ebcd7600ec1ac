"""NumPy-only restatement ("port") of what the reference computes on the hot path — TEST INFRASTRUCTURE.

No autograd here: the backward passes are written out by hand from the reference's VJP rules, so this is an
oracle that is independent of the tracing machinery.  Each function cites the reference rule it restates.
Pinned against tests/golden/*.npz (generated from the real reference) by tests/test_oracle_port.py; that is
what makes it trustworthy — parity of the device path is then checked against BOTH the live reference and this port.
"""

from __future__ import annotations

import numpy as np
from scipy.special import logsumexp


def mlp_objective_and_grad(params, X, T, L2_reg):
    """-log_posterior of examples/neural_net.py:28-48 and its gradient.

    forward : h_{l+1} = tanh(h_l W_l + b_l); logits = last pre-activation; logp = logits - logsumexp(logits)
    backward: dot VJPs (numpy_vjps.py:615-651): dW = h^T G, dh = G W^T; add VJP with unbroadcast (:76-78,883-892):
              db = sum_rows G; tanh VJP (:178): G / cosh(a)^2; logsumexp VJP (scipy/special.py:123-131):
              g * exp(x - ans); prior: d/dtheta (L2 * theta.theta) = 2 L2 theta (dot VJP, 1-D case)."""
    acts, pre = [X], []
    h = X
    for W, b in params:
        a = np.dot(h, W) + b
        pre.append(a)
        h = np.tanh(a)
        acts.append(h)
    logits = pre[-1]
    lse = logsumexp(logits, axis=1, keepdims=True)
    logp = logits - lse
    flat = np.concatenate([np.ravel(p) for wb in params for p in wb])
    value = -(-L2_reg * np.dot(flat, flat) + np.sum(logp * T))
    # objective = L2 * theta.theta - sum(logp * T)
    G = -T                                                     # d objective / d logp
    G = G - np.exp(logits - lse) * np.sum(G, axis=1, keepdims=True)   # through logits - logsumexp(logits)
    grads = [None] * len(params)
    for l in range(len(params) - 1, -1, -1):
        W, b = params[l]
        h_in = acts[l]
        grads[l] = (np.dot(h_in.T, G) + 2.0 * L2_reg * W, np.sum(G, axis=0) + 2.0 * L2_reg * b)
        if l > 0:
            G = np.dot(G, W.T) / np.cosh(pre[l - 1]) ** 2
    return value, grads


def tanh_derivatives(x, max_order=4):
    """Closed forms of d^k/dx^k tanh(x), k = 0..4, in t = tanh(x) (the function of examples/tanh.py:22-23 is tanh
    written with exponentials).  Mathematically equal to the nested elementwise_grad tower, not bit-identical."""
    t = np.tanh(x)
    s = 1.0 - t * t
    out = [t, s, -2.0 * t * s, s * (6.0 * t * t - 2.0), s * t * (16.0 - 24.0 * t * t)]
    return out[: max_order + 1]


def conv_valid_nchw(A, B):
    """autograd.scipy.signal.convolve(A, B, axes=([2,3],[2,3]), dot_axes=([1],[0]), mode='valid')
    (autograd/scipy/signal.py:10-65): true convolution, i.e. correlation with the flipped kernel."""
    N, C, H, W = A.shape
    _, F, kh, kw = B.shape
    Ho, Wo = H - kh + 1, W - kw + 1
    Bf = B[:, :, ::-1, ::-1]
    out = np.zeros((N, F, Ho, Wo))
    for i in range(kh):
        for j in range(kw):
            out += np.einsum("nchw,cf->nfhw", A[:, :, i:i + Ho, j:j + Wo], Bf[:, :, i, j])
    return out


def untake_pairs(shape, rows, cols, g):
    """VJP of f[rows, cols] (ArrayBox.__getitem__ with index arrays): np.add.at with duplicates accumulated
    (numpy_vjps.py:941-954)."""
    out = np.zeros(shape)
    np.add.at(out, (rows, cols), g)
    return out


def logsumexp_weighted_and_grad(x, b, axis=None):
    """log(sum(b * exp(x))) over ``axis`` and the gradient of sum(result**2) wrt x, written out by hand.

    forward : shift by the maximum (scipy/special/_logsumexp.py); backward: autograd/scipy/special.py:123-131 —
    d lse / d x = b * exp(x - lse) broadcast back over the reduced axes; the square's VJP is 2 * lse."""
    m = np.max(x, axis=axis, keepdims=True)
    s = np.sum(b * np.exp(x - m), axis=axis, keepdims=True)
    lse = np.log(s) + m
    grad = 2.0 * lse * b * np.exp(x - lse)
    return (lse.reshape(()) if axis is None else np.squeeze(lse, axis=axis)), grad


def convolve_direct(A, B, axes=None, dot_axes=((), ()), mode="full"):
    """autograd.scipy.signal.convolve (scipy/signal.py:10-76) as explicit sums: no strided views, no einsum.

    out[ia..., ib..., p...] = sum_{d, k} A[ia, d, k] * B[ib, d, p - k]   (true convolution along the convolved axes,
    contraction over the dot axes, untouched axes of A then of B first — the reference's output layout), restricted to
    the window that ``mode`` keeps: 'full' every p with overlap, 'valid' the p where the smaller operand lies fully
    inside the larger, 'same' the central part of 'full' with A's extent."""
    A, B = np.asarray(A, dtype=float), np.asarray(B, dtype=float)
    if axes is None:
        axes = (list(range(A.ndim)), list(range(A.ndim)))
    ca, cb = [list(a) for a in axes]
    da, db = [list(a) for a in dot_axes]
    ia = [i for i in range(A.ndim) if i not in ca and i not in da]
    ib = [i for i in range(B.ndim) if i not in cb and i not in db]
    At = np.transpose(A, ia + da + ca)
    Bt = np.transpose(B, ib + db + cb)
    na, nb = [A.shape[i] for i in ca], [B.shape[i] for i in cb]
    full_shape = [x + y - 1 for x, y in zip(na, nb)]
    out = np.zeros([A.shape[i] for i in ia] + [B.shape[i] for i in ib] + full_shape)
    nia, nib, nd = len(ia), len(ib), len(da)
    for ka in np.ndindex(*na):
        a_slab = At[(Ellipsis,) + ka]                                   # [ia..., d...]
        for kb in np.ndindex(*nb):
            b_slab = Bt[(Ellipsis,) + kb]                               # [ib..., d...]
            contrib = np.tensordot(a_slab, b_slab, axes=(list(range(nia, nia + nd)), list(range(nib, nib + nd))))
            out[(Ellipsis,) + tuple(i + j for i, j in zip(ka, kb))] += contrib
    if mode == "full":
        return out
    sl = [slice(None)] * (nia + nib)
    for x, y in zip(na, nb):
        if mode == "valid":
            lo, n = min(x, y) - 1, abs(x - y) + 1
        else:                                                            # 'same': A's extent, centred like the reference
            lo, n = (y - 1) // 2, x
        sl.append(slice(lo, lo + n))
    return out[tuple(sl)]
