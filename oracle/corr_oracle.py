"""numpy restatement of the reference's RAFT correlation path -- TEST INFRASTRUCTURE ONLY.

This file is the parity checker for the CUDA path.  It is never imported by the
product package (see ``oracle/__init__.py``).

Every function cites the reference lines (relative to ``/root/reference``) it
restates.  The arithmetic that the reference delegates to PyTorch
(``torch.matmul``, ``F.avg_pool2d(k=2, s=2)``, ``F.grid_sample(bilinear, zeros,
align_corners=True)``; torch>=1.6 per ``environment.yml:6``, 2.11.0 in this image)
is restated here in closed form:

    out[b, l*K*K + i*K + j, h, w] =
        bilinear_zero_pad(V_l[b, h, w], x = cx / 2**l + (i - r), y = cy / 2**l + (j - r))

with ``K = 2r+1`` and ``(cx, cy) = coords[b, :, h, w]``: the FIRST window index
``i`` moves along x (raft/core/corr.py:66-78 builds ``delta`` with ``meshgrid(dy,
dx, 'ij')`` and adds it to an (x, y) centroid).

All functions take/return numpy arrays and run in the dtype given by ``dtype``
(float32 mirrors the reference, float64 gives the "true" answer used to size
tolerances).  Pyramid levels are ``[B, N, H_l, W_l]`` with ``N = H*W`` source
pixels, i.e. the reference's ``[B*N, 1, H_l, W_l]`` (corr.py:34-35) un-flattened.
"""
from __future__ import annotations

import math

import numpy as np


# --------------------------------------------------------------------------- shapes
def level_shapes(H: int, W: int, num_levels: int):
    """Sizes of the pyramid levels: cascaded floor halving (F.avg_pool2d(2, stride=2),
    ceil_mode=False; raft/core/corr.py:41-43)."""
    out = []
    h, w = H, W
    for _ in range(num_levels):
        out.append((h, w))
        h, w = h // 2, w // 2
    return out


# --------------------------------------------------------------------------- encoder head
def fnet_head(x, weight, bias=None, dtype=np.float32):
    """The feature encoders' output head: ``fmap = conv2(x)`` with a 1x1 kernel (raft/core/extractor.py:231
    BasicEncoder, :341 SmallEncoder; ``conv2 = nn.Conv2d(Cin, D, kernel_size=1)``, :172/:287), split per frame
    (``torch.split(x, [B, B], dim=0)``, :239/:349) and cast to fp32 (raft/core/raft.py:114-115).

    ``x``: ``[2B, Cin, H, W]``, ``weight``: ``[D, Cin(, 1, 1)]``, ``bias``: ``[D]`` or None.
    Returns ``(fmap1, fmap2)``, each ``[B, D, H, W]``.
    """
    B2, Cin, H, W = x.shape
    w = np.asarray(weight).reshape(weight.shape[0], Cin).astype(dtype)
    f = np.einsum("dc,bchw->bdhw", w, np.asarray(x).astype(dtype), optimize=True)
    if bias is not None:
        f = f + np.asarray(bias).astype(dtype)[None, :, None, None]
    return f[: B2 // 2], f[B2 // 2:]


# --------------------------------------------------------------------------- build
def corr_volume(fmap1, fmap2, dtype=np.float32):
    """All-pairs volume ``V0[b,h1,w1,0,h2,w2] = <f1[b,:,h1,w1], f2[b,:,h2,w2]> / sqrt(D)``.

    Restates ``CorrBlock.corr`` (raft/core/corr.py:93-116): view [B,D,HW], matmul of the
    transposed first map with the second, view [B,H,W,1,H,W], multiply by the fp32
    scalar ``1/sqrt(D)`` (corr.py:116 builds it as ``1.0 / torch.sqrt(torch.tensor(dim).float())``).
    """
    B, D, H, W = fmap1.shape
    a = np.ascontiguousarray(fmap1.reshape(B, D, H * W)).astype(dtype)
    b = np.ascontiguousarray(fmap2.reshape(B, D, H * W)).astype(dtype)
    v = np.matmul(a.transpose(0, 2, 1), b)
    if dtype == np.float32:
        scale = np.float32(1.0) / np.sqrt(np.float32(D))
    else:
        scale = dtype(1.0 / math.sqrt(D))
    v = v * scale
    return v.reshape(B, H, W, 1, H, W)


def avg_pool2(v):
    """2x2 mean, stride 2, floor mode (odd trailing row/col dropped) over the last two
    dims -- ``F.avg_pool2d(corr, 2, stride=2)`` at raft/core/corr.py:42."""
    H, W = v.shape[-2:]
    h, w = H // 2, W // 2
    c = v[..., : 2 * h, : 2 * w]
    s = c[..., 0::2, 0::2] + c[..., 0::2, 1::2] + c[..., 1::2, 0::2] + c[..., 1::2, 1::2]
    return s * v.dtype.type(0.25)


def corr_pyramid(fmap1, fmap2, num_levels=4, dtype=np.float32):
    """``CorrBlock.__init__`` (raft/core/corr.py:18-43): volume, fold the source pixels into the
    leading dim, then ``num_levels-1`` cascaded poolings over the TARGET dims.
    Returns ``[V_0, ..., V_{L-1}]`` with ``V_l`` of shape ``[B, N, H_l, W_l]``."""
    B, D, H, W = fmap1.shape
    v = corr_volume(fmap1, fmap2, dtype).reshape(B, H * W, H, W)
    pyr = [v]
    for _ in range(num_levels - 1):
        v = avg_pool2(v)
        pyr.append(v)
    return pyr


# --------------------------------------------------------------------------- lookup
def _lattice(V, x, y, r):
    """Gather the (K+1)x(K+1) zero-padded lattice patch around floor(x), floor(y).

    V: [B,N,Hl,Wl]; x, y: [B,N].  Returns P[B,N,row,col], fx, fy, rows, cols, mask.
    Zero padding per corner is ``F.grid_sample``'s default ``padding_mode='zeros'``
    reached through ``bilinear_sampler`` (raft/core/utils/utils.py:73-78)."""
    B, N, Hl, Wl = V.shape
    K1 = 2 * r + 2
    x0 = np.floor(x)
    y0 = np.floor(y)
    fx = (x - x0).astype(V.dtype)
    fy = (y - y0).astype(V.dtype)
    ar = np.arange(K1) - r
    cols = x0.astype(np.int64)[..., None] + ar  # [B,N,K1]
    rows = y0.astype(np.int64)[..., None] + ar
    cm = (cols >= 0) & (cols < Wl)
    rm = (rows >= 0) & (rows < Hl)
    cc = np.clip(cols, 0, Wl - 1)
    rc = np.clip(rows, 0, Hl - 1)
    bi = np.arange(B)[:, None, None, None]
    ni = np.arange(N)[None, :, None, None]
    P = V[bi, ni, rc[:, :, :, None], cc[:, :, None, :]]
    mask = rm[:, :, :, None] & cm[:, :, None, :]
    P = np.where(mask, P, V.dtype.type(0))
    return P, fx, fy, rc, cc, mask


def lookup(pyramid, coords, radius=4):
    """``CorrBlock.__call__`` (raft/core/corr.py:45-91) + ``bilinear_sampler``
    (raft/core/utils/utils.py:57-85).

    pyramid: list of ``[B,N,H_l,W_l]``; coords: ``[B,2,H,W]`` (channel 0 = x, 1 = y, level-0
    pixel units, corr.py:57).  Returns ``[B, L*K*K, H, W]``.
    """
    r = radius
    K = 2 * r + 1
    B, _, H, W = coords.shape
    N = H * W
    dt = pyramid[0].dtype
    cx = coords[:, 0].reshape(B, N).astype(dt)
    cy = coords[:, 1].reshape(B, N).astype(dt)
    outs = []
    for l, V in enumerate(pyramid):
        x = cx / dt.type(2 ** l)  # corr.py:71 centroid / 2**i
        y = cy / dt.type(2 ** l)
        P, fx, fy, *_ = _lattice(V, x, y, r)
        fx = fx[:, :, None, None]
        fy = fy[:, :, None, None]
        one = dt.type(1)
        o = ((one - fy) * (one - fx) * P[:, :, :-1, :-1] + (one - fy) * fx * P[:, :, :-1, 1:]
             + fy * (one - fx) * P[:, :, 1:, :-1] + fy * fx * P[:, :, 1:, 1:])  # [B,N,j(y),i(x)]
        outs.append(o.transpose(0, 1, 3, 2).reshape(B, N, K * K))  # channel = i*K + j
    out = np.concatenate(outs, axis=-1)  # corr.py:90
    return np.ascontiguousarray(out.transpose(0, 2, 1)).reshape(B, len(pyramid) * K * K, H, W)


def lookup_convc1(pyramid, coords, weight, bias=None, radius=4, relu=True):
    """the lookup followed by its only consumer, the motion encoder's first 1x1
    convolution and ReLU -- ``cor = F.relu(self.convc1(corr))`` (raft/core/update.py:135 SmallMotionEncoder,
    :170 BasicMotionEncoder; ``convc1 = nn.Conv2d(cor_planes, 96 | 256, 1)``, :116 / :158).

    weight ``[Cout, L*K*K(, 1, 1)]``, bias ``[Cout]`` or None.  Returns ``[B, Cout, H, W]``.
    """
    feat = lookup(pyramid, coords, radius)  # [B, Cin, H, W]
    dt = feat.dtype
    w = np.asarray(weight).reshape(weight.shape[0], -1).astype(dt)
    out = np.einsum("oc,bchw->bohw", w, feat)
    if bias is not None:
        out = out + np.asarray(bias).astype(dt)[None, :, None, None]
    return np.maximum(out, dt.type(0)) if relu else out


def lookup_grad(pyramid, coords, grad_out, radius=4, want_coords=False):
    """Adjoint of :func:`lookup` w.r.t. the pyramid (and optionally the coordinates).

    Restates what PyTorch autograd does for raft/core/corr.py:81 / utils.py:78
    (``grid_sampler_2d_backward``): ``dV_l[row,col] += w * g`` for the in-bounds corners of
    every tap, and ``dcoords = sum g * dP/dx / 2**l``.
    Returns ``([dV_0..dV_{L-1}], dcoords or None)``.
    """
    r = radius
    K = 2 * r + 1
    B, _, H, W = coords.shape
    N = H * W
    L = len(pyramid)
    dt = pyramid[0].dtype
    cx = coords[:, 0].reshape(B, N).astype(dt)
    cy = coords[:, 1].reshape(B, N).astype(dt)
    g_all = grad_out.reshape(B, L, K, K, N).astype(dt)  # [B,l,i,j,N]
    dpyr = []
    dco = np.zeros((B, 2, N), dt) if want_coords else None
    bi = np.arange(B)[:, None, None, None]
    ni = np.arange(N)[None, :, None, None]
    one = dt.type(1)
    for l, V in enumerate(pyramid):
        x = cx / dt.type(2 ** l)
        y = cy / dt.type(2 ** l)
        P, fx, fy, rc, cc, mask = _lattice(V, x, y, r)
        g = g_all[:, l].transpose(0, 3, 2, 1)  # [B,N,j,i]
        fxe = fx[:, :, None, None]
        fye = fy[:, :, None, None]
        dP = np.zeros_like(P)
        dP[:, :, :-1, :-1] += (one - fye) * (one - fxe) * g
        dP[:, :, :-1, 1:] += (one - fye) * fxe * g
        dP[:, :, 1:, :-1] += fye * (one - fxe) * g
        dP[:, :, 1:, 1:] += fye * fxe * g
        dP = np.where(mask, dP, dt.type(0))
        dV = np.zeros_like(V)
        np.add.at(dV, (bi, ni, rc[:, :, :, None], cc[:, :, None, :]), dP)
        dpyr.append(dV)
        if want_coords:
            ddx = (one - fye) * (P[:, :, :-1, 1:] - P[:, :, :-1, :-1]) + fye * (P[:, :, 1:, 1:] - P[:, :, 1:, :-1])
            ddy = (one - fxe) * (P[:, :, 1:, :-1] - P[:, :, :-1, :-1]) + fxe * (P[:, :, 1:, 1:] - P[:, :, :-1, 1:])
            dco[:, 0] += (g * ddx).sum(axis=(2, 3)) / dt.type(2 ** l)
            dco[:, 1] += (g * ddy).sum(axis=(2, 3)) / dt.type(2 ** l)
    if want_coords:
        dco = dco.reshape(B, 2, H, W)
    return dpyr, dco


def avg_pool2_adjoint(g, H, W):
    """Adjoint of :func:`avg_pool2`: each pooled gradient is spread as g/4 over its 2x2
    children; rows/cols dropped by floor mode get zero (``avg_pool2d_backward``)."""
    out = np.zeros(g.shape[:-2] + (H, W), g.dtype)
    h, w = g.shape[-2:]
    q = g * g.dtype.type(0.25)
    out[..., 0:2 * h:2, 0:2 * w:2] = q
    out[..., 0:2 * h:2, 1:2 * w:2] = q
    out[..., 1:2 * h:2, 0:2 * w:2] = q
    out[..., 1:2 * h:2, 1:2 * w:2] = q
    return out


def build_grad(dpyramid, fmap1, fmap2):
    """Autograd of raft/core/corr.py:31-43,107-116: fold the pooled-level gradients down to
    level 0, then ``dF1 = F2 . dV0^T / sqrt(D)``, ``dF2 = F1 . dV0 / sqrt(D)``."""
    dt = dpyramid[0].dtype
    B, D, H, W = fmap1.shape
    g = dpyramid[-1]
    for l in range(len(dpyramid) - 2, -1, -1):
        Hl, Wl = dpyramid[l].shape[-2:]
        g = dpyramid[l] + avg_pool2_adjoint(g, Hl, Wl)
    g0 = g.reshape(B, H * W, H * W)
    if dt == np.float32:
        scale = np.float32(1.0) / np.sqrt(np.float32(D))
    else:
        scale = dt.type(1.0 / math.sqrt(D))
    a = fmap1.reshape(B, D, H * W).astype(dt)
    b = fmap2.reshape(B, D, H * W).astype(dt)
    df1 = np.matmul(b, g0.transpose(0, 2, 1)) * scale
    df2 = np.matmul(a, g0) * scale
    return df1.reshape(B, D, H, W), df2.reshape(B, D, H, W)


# --------------------------------------------------------------------------- on-the-fly variant
def alt_feature_pyramid(fmap2, num_levels=4):
    """Pooled target feature maps read by ``AlternateCorrBlock.__call__``: ``pyramid[l][1]`` for
    ``l < num_levels`` (raft/core/corr.py:136-140,159).  NCHW in, list of NCHW out."""
    out = [fmap2]
    f = fmap2
    for _ in range(num_levels - 1):
        f = avg_pool2(f)
        out.append(f)
    return out


def alt_lookup(fmap1, fmap2, coords, num_levels=4, radius=4, dtype=np.float32):
    """``AlternateCorrBlock.__call__`` (raft/core/corr.py:142-167) with the native forward it
    calls, ``corr_forward_kernel`` (raft/alt_cuda_corr/correlation_kernel.cu:18-119), restated:

    for every source pixel and lattice point ``(iy, ix) in [0, K]^2`` the dot product
    ``s = <f1[p], f2_l[floor(y)-r+iy, floor(x)-r+ix]>`` (0 outside the map, :80-84) is splatted
    into up to four outputs with weights ``dy*dx, dy*(1-dx), (1-dy)*dx, (1-dy)*(1-dx)`` at channel
    ``iy + K*ix`` (:92-114); levels are stacked and divided by sqrt(D) (corr.py:165-167).
    """
    r = radius
    K = 2 * r + 1
    B, D, H, W = fmap1.shape
    N = H * W
    f1 = fmap1.reshape(B, D, N).transpose(0, 2, 1).astype(dtype)  # [B,N,D]  (corr.py:158 NHWC)
    cx = coords[:, 0].reshape(B, N).astype(dtype)
    cy = coords[:, 1].reshape(B, N).astype(dtype)
    outs = []
    one = dtype(1)
    for l, f2l in enumerate(alt_feature_pyramid(fmap2.astype(dtype), num_levels)):
        Hl, Wl = f2l.shape[-2:]
        f2n = f2l.transpose(0, 2, 3, 1)  # NHWC (corr.py:159)
        x = cx / dtype(2 ** l)  # corr.py:161
        y = cy / dtype(2 ** l)
        x0 = np.floor(x)
        y0 = np.floor(y)
        dx = (x - x0)[:, :, None, None]
        dy = (y - y0)[:, :, None, None]
        ar = np.arange(K + 1) - r
        cols = x0.astype(np.int64)[..., None] + ar
        rows = y0.astype(np.int64)[..., None] + ar
        mask = ((rows >= 0) & (rows < Hl))[:, :, :, None] & ((cols >= 0) & (cols < Wl))[:, :, None, :]
        rc = np.clip(rows, 0, Hl - 1)
        cc = np.clip(cols, 0, Wl - 1)
        bi = np.arange(B)[:, None, None, None]
        g = f2n[bi, rc[:, :, :, None], cc[:, :, None, :]]  # [B,N,K+1,K+1,D]
        s = np.einsum("bnd,bnyxd->bnyx", f1, g)
        s = np.where(mask, s, dtype(0))
        o = np.zeros((B, N, K, K), dtype)  # [.., iy, ix]
        o += s[:, :, 1:, 1:] * dy * dx  # nw  -> (iy-1, ix-1)
        o += s[:, :, 1:, :-1] * dy * (one - dx)  # ne  -> (iy-1, ix)
        o += s[:, :, :-1, 1:] * (one - dy) * dx  # sw  -> (iy, ix-1)
        o += s[:, :, :-1, :-1] * (one - dy) * (one - dx)  # se -> (iy, ix)
        outs.append(o.transpose(0, 1, 3, 2).reshape(B, N, K * K))  # channel iy + K*ix
    # corr.py:165-166: stack levels on dim 1, reshape to [B, L*K*K, H, W]
    out = np.stack([o.transpose(0, 2, 1) for o in outs], axis=1).reshape(B, num_levels * K * K, H, W)
    if dtype == np.float32:
        return out / np.sqrt(np.float32(D))
    return out / dtype(math.sqrt(D))


def alt_grad(fmap1, fmap2, coords, grad_out, num_levels=4, radius=4, dtype=np.float32):
    """Gradient of :func:`alt_lookup` w.r.t. ``fmap1`` and ``fmap2`` (NCHW).

    Restates ``corr_backward_kernel`` (raft/alt_cuda_corr/correlation_kernel.cu:122-256):
    ``g = sum_{4 outputs} w * corr_grad`` per lattice point (:213-223), ``f1_grad += g*f2``,
    ``f2_grad += g*f1`` (:225-238), coords gradient identically zero (:307) -- plus what the
    reference leaves to Python: the 1/sqrt(D) factor (corr.py:167) and the adjoint of the
    fmap2 poolings (corr.py:138-139).
    """
    r = radius
    K = 2 * r + 1
    B, D, H, W = fmap1.shape
    N = H * W
    f1 = fmap1.reshape(B, D, N).transpose(0, 2, 1).astype(dtype)
    cx = coords[:, 0].reshape(B, N).astype(dtype)
    cy = coords[:, 1].reshape(B, N).astype(dtype)
    scale = dtype(1.0 / math.sqrt(D))
    go = grad_out.reshape(B, num_levels, K, K, N).astype(dtype) * scale  # [B,l,ix,iy,N]
    df1 = np.zeros((B, N, D), dtype)
    pyr = alt_feature_pyramid(fmap2.astype(dtype), num_levels)
    dpyr = []
    one = dtype(1)
    for l, f2l in enumerate(pyr):
        Hl, Wl = f2l.shape[-2:]
        f2n = f2l.transpose(0, 2, 3, 1)
        x = cx / dtype(2 ** l)
        y = cy / dtype(2 ** l)
        x0 = np.floor(x)
        y0 = np.floor(y)
        dx = (x - x0)[:, :, None, None]
        dy = (y - y0)[:, :, None, None]
        ar = np.arange(K + 1) - r
        cols = x0.astype(np.int64)[..., None] + ar
        rows = y0.astype(np.int64)[..., None] + ar
        mask = ((rows >= 0) & (rows < Hl))[:, :, :, None] & ((cols >= 0) & (cols < Wl))[:, :, None, :]
        rc = np.clip(rows, 0, Hl - 1)
        cc = np.clip(cols, 0, Wl - 1)
        gl = go[:, l].transpose(0, 3, 2, 1)  # [B,N,iy,ix]
        gs = np.zeros((B, N, K + 1, K + 1), dtype)
        gs[:, :, 1:, 1:] += gl * dy * dx
        gs[:, :, 1:, :-1] += gl * dy * (one - dx)
        gs[:, :, :-1, 1:] += gl * (one - dy) * dx
        gs[:, :, :-1, :-1] += gl * (one - dy) * (one - dx)
        gs = np.where(mask, gs, dtype(0))
        bi = np.arange(B)[:, None, None, None]
        gath = f2n[bi, rc[:, :, :, None], cc[:, :, None, :]]  # [B,N,K+1,K+1,D]
        df1 += np.einsum("bnyx,bnyxd->bnd", gs, gath)
        d2 = np.zeros((B, Hl, Wl, D), dtype)
        contrib = gs[..., None] * f1[:, :, None, None, :]
        np.add.at(d2, (bi, rc[:, :, :, None], cc[:, :, None, :]), contrib)
        dpyr.append(d2.transpose(0, 3, 1, 2))
    g = dpyr[-1]
    for l in range(num_levels - 2, -1, -1):
        Hl, Wl = dpyr[l].shape[-2:]
        g = dpyr[l] + avg_pool2_adjoint(g, Hl, Wl)
    return df1.transpose(0, 2, 1).reshape(B, D, H, W), g


# --------------------------------------------------------------------------- convex upsampling
def _neighbours(flow):
    """``F.unfold(8 * flow, [3, 3], padding=1)`` as ``[B, C, 9, H, W]`` (k = ky*3 + kx, zero padding)."""
    B, C, H, W = flow.shape
    pad = np.zeros((B, C, H + 2, W + 2), flow.dtype)
    pad[:, :, 1:-1, 1:-1] = flow.dtype.type(8) * flow
    return np.stack([pad[:, :, ky:ky + H, kx:kx + W] for ky in range(3) for kx in range(3)], axis=2)


def upsample_flow(flow, mask):
    """``RAFT.upsample_flow`` (raft/core/raft.py:70-81): convex combination of the 3x3 neighbourhood of ``8 * flow``
    with ``softmax`` weights, 8x8 sub-pixels per feature pixel.  flow ``[B,C,H,W]``, mask ``[B,576,H,W]`` ->
    ``[B,C,8H,8W]``."""
    B, C, H, W = flow.shape
    m = mask.reshape(B, 1, 9, 8, 8, H, W).astype(flow.dtype)       # raft.py:73
    m = np.exp(m - m.max(axis=2, keepdims=True))
    m = m / m.sum(axis=2, keepdims=True)                           # raft.py:74
    nb = _neighbours(flow).reshape(B, C, 9, 1, 1, H, W)            # raft.py:76-77
    up = (m * nb).sum(axis=2)                                      # raft.py:79  [B,C,8,8,H,W]
    return np.ascontiguousarray(up.transpose(0, 1, 4, 2, 5, 3)).reshape(B, C, 8 * H, 8 * W)  # raft.py:80-81


def upsample_flow_grad(flow, mask, grad_out):
    """Adjoint of :func:`upsample_flow`: returns ``(dflow [B,C,H,W], dmask [B,576,H,W])``."""
    B, C, H, W = flow.shape
    dt = flow.dtype
    m = mask.reshape(B, 1, 9, 8, 8, H, W).astype(dt)
    p = np.exp(m - m.max(axis=2, keepdims=True))
    p = p / p.sum(axis=2, keepdims=True)
    nb = _neighbours(flow).reshape(B, C, 9, 1, 1, H, W)
    g = grad_out.reshape(B, C, H, 8, W, 8).transpose(0, 1, 3, 5, 2, 4)[:, :, None]  # [B,C,1,8,8,H,W]
    s = (g * nb).sum(axis=1, keepdims=True)                        # d out / d p_k   [B,1,9,8,8,H,W]
    dmask = p * (s - (p * s).sum(axis=2, keepdims=True))
    dnb = (g * p).sum(axis=(3, 4))                                 # [B,C,9,H,W]
    dpad = np.zeros((B, C, H + 2, W + 2), dt)
    for k in range(9):
        ky, kx = divmod(k, 3)
        dpad[:, :, ky:ky + H, kx:kx + W] += dnb[:, :, k]
    return dt.type(8) * dpad[:, :, 1:-1, 1:-1], dmask.reshape(B, 576, H, W)


# --------------------------------------------------------------------------- helpers
def coords_grid(B, H, W, dtype=np.float32):
    """``coords_grid`` (raft/core/utils/utils.py:88-102): channel 0 = x, channel 1 = y."""
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    g = np.stack([xs, ys], axis=0).astype(dtype)
    return np.broadcast_to(g[None], (B, 2, H, W)).copy()
