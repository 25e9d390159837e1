"""Host-side mirror of the reference's ``raft/core/corr.py``: same classes, constructor and call
signatures, attributes and autograd behaviour -- all arithmetic in libraftcorr_b200 (sm_100a).

    CorrBlock(fmap1, fmap2, num_levels=4, radius=4)        raft/core/corr.py:18
    CorrBlock.__call__(coords) -> [B, L*(2r+1)^2, H, W]    raft/core/corr.py:45
    CorrBlock.corr(fmap1, fmap2) -> [B, H, W, 1, H, W]     raft/core/corr.py:93
    AlternateCorrBlock(...) / __call__                      raft/core/corr.py:124,142

PyTorch is used for device memory (caching allocator), streams and autograd bookkeeping only.
CPU tensors are rejected: there is no CPU or eager fallback.
"""
from __future__ import annotations

import ctypes
import math
import os

import torch

from . import _lib
from ._lib import RaftCorrError, check

_MODES = {"fp32": _lib.BUILD_FP32_SIMT, "tf32": _lib.BUILD_TF32_TC, "f16": _lib.BUILD_F16_TC, "f16x3": _lib.BUILD_F16X3_TC}
_default_mode = os.environ.get("RAFTCORR_BUILD_MODE", "f16")
_F16_MIN_PIXELS = 4096  # B*H*W below which the DEFAULT mode builds with TF32 operands instead of block-scaled fp16
_LAYOUTS = {"rowmajor": 0, "pairs": 1, "interleaved": 1, "quads": 2}  # raftcorr_pyramid_layout::interleaved


def set_default_build_mode(mode: str) -> None:
    """``"tf32"``: tcgen05 tensor-core build (TF32 operands, fp32 accumulate; <=1e-3 of max|V|).
    ``"f16"``: the same kernel on block-scaled fp16 operands (TF32's 11 significant bits, one power-of-two scale per
    operand tile, fp32 accumulate; same error, half the operand traffic) -- volume build only, everything else
    (backward, on-the-fly class, fused encoder head) runs as ``"tf32"``.
    ``"f16x3"``: fp32-class accuracy on the tensor cores (three-term fp16 split, ~1e-6 of max|V|; forward only).
    ``"fp32"``: exact fp32 FFMA build on CUDA cores."""
    global _default_mode
    if mode not in _MODES:
        raise ValueError(f"unknown build mode {mode!r} (expected one of {sorted(_MODES)})")
    _default_mode = mode


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _raw_stream(index):
    # the current stream of device `index` as an integer handle, without building a torch.cuda.Stream object
    # (the lookup is called 12-32 times per volume; at batch 1 its kernel takes ~11 us, so host microseconds count)
    return torch._C._cuda_getCurrentRawStream(index)


def _check_fmaps(fmap1, fmap2, who):
    for name, t in (("fmap1", fmap1), ("fmap2", fmap2)):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{who}: {name} must be a torch.Tensor")
        if not t.is_cuda:
            # same contract as the reference extension's CHECK_CUDA (raft/alt_cuda_corr/correlation.cpp:19)
            raise RuntimeError(f"{who}: {name} must be a CUDA tensor (this package has no CPU path)")
        if t.dtype != torch.float32:
            raise RuntimeError(f"{who}: {name} must be float32 (RAFT casts with .float(), raft/core/raft.py:114-115)")
        if t.dim() != 4:
            raise RuntimeError(f"{who}: {name} must be [B, C, H, W]")
    if fmap1.shape != fmap2.shape or fmap1.device != fmap2.device:
        raise RuntimeError(f"{who}: fmap1 and fmap2 must have the same shape and device")


def _check_coords(coords, B, H, W, device, who):
    if not (isinstance(coords, torch.Tensor) and coords.is_cuda and coords.dtype == torch.float32):
        raise RuntimeError(f"{who}: coords must be a float32 CUDA tensor")
    if tuple(coords.shape) != (B, 2, H, W):
        raise RuntimeError(f"{who}: coords must be [B, 2, H, W] = {(B, 2, H, W)}, got {tuple(coords.shape)}")
    if coords.device != device:
        raise RuntimeError(f"{who}: coords is on {coords.device}, the correlation block on {device}")


# ----------------------------------------------------------------------------------------------
# all-pairs volume
# ----------------------------------------------------------------------------------------------
class _BuildFn(torch.autograd.Function):
    """fmaps -> flat pyramid buffer (raft/core/corr.py:31-43)."""

    @staticmethod
    def forward(ctx, fmap1, fmap2, block):
        ctx.block = block
        ctx.save_for_backward(fmap1, fmap2)
        return block._build(fmap1, fmap2)

    @staticmethod
    def backward(ctx, dpyr):
        fmap1, fmap2 = ctx.saved_tensors
        block = ctx.block
        dpyr, folded = block._take_grad_acc(dpyr)
        if dpyr is None:
            return None, None, None
        df1 = torch.empty_like(fmap1)
        df2 = torch.empty_like(fmap2)
        block._build_bwd(dpyr, folded, fmap1, fmap2, df1, df2)
        return df1, df2, None


class _HeadBuildFn(torch.autograd.Function):
    """(x, conv2.weight, conv2.bias) -> flat pyramid buffer: extractor.py:231,239 + raft.py:114-119 in one kernel pair.
    Backward: recompute fmap = conv2(x) (the forward never stored it), volume adjoint -> dfmap, then conv2's own
    adjoint through torch autograd."""

    @staticmethod
    def forward(ctx, x, w2, bias, block):
        ctx.block = block
        ctx.save_for_backward(x, w2, bias)
        return block._head_build(x, w2.contiguous(), bias)

    @staticmethod
    def backward(ctx, dpyr):
        x, w2, bias = ctx.saved_tensors
        block = ctx.block
        dpyr, folded = block._take_grad_acc(dpyr)
        if dpyr is None:
            return None, None, None, None
        B, D = block._B, block._D
        dev = x.device
        with torch.enable_grad():
            xg = x.detach().requires_grad_(ctx.needs_input_grad[0])
            wg = w2.detach().requires_grad_(ctx.needs_input_grad[1])
            bg = bias.detach().requires_grad_(ctx.needs_input_grad[2]) if bias is not None else None
            f = torch.nn.functional.conv2d(xg, wg.reshape(D, -1, 1, 1), bg)
        fd = f.detach()
        df = torch.empty_like(fd)
        block._build_bwd(dpyr, folded, fd[:B], fd[B:], df[:B], df[B:])
        wanted = [t for t in (xg, wg, bg) if t is not None and t.requires_grad]
        grads = list(torch.autograd.grad(f, wanted, df)) if wanted else []
        out = [grads.pop(0) if (t is not None and t.requires_grad) else None for t in (xg, wg, bg)]
        return out[0], out[1], out[2], None


class _LookupFn(torch.autograd.Function):
    """(pyramid, coords) -> motion features (raft/core/corr.py:45-91)."""

    @staticmethod
    def forward(ctx, pyr, coords, block):
        ctx.block = block
        ctx.need_dcoords = coords.requires_grad
        ctx.save_for_backward(coords, pyr if coords.requires_grad else None)
        return block._lookup(pyr, coords)

    @staticmethod
    def backward(ctx, gout):
        coords, pyr = ctx.saved_tensors
        dpyr_ret, dcoords = _lookup_backward(ctx.block, gout, coords, pyr, ctx.needs_input_grad[0],
                                             ctx.need_dcoords and ctx.needs_input_grad[1])
        return dpyr_ret, dcoords, None


def _lookup_backward(block, gout, coords, pyr, want_pyr, want_dcoords):
    """Adjoint of one lookup: scatter ``gout`` into the block's gradient pyramid (and optionally d/dcoords)."""
    lay = block._layout
    dev = coords.device
    dpyr_ret = None
    if want_pyr and not want_dcoords and block._defer_ok:
        # DEFERRED: the lookups of a backward pass only queue (grad_out, coords); the build backward -- which depends
        # on every lookup, so it runs last -- scatters all of them in ONE launch that accumulates a source pixel's
        # gradient image in shared memory, folds the pooled levels on chip and writes level 0 once
        # (raftcorr_lookup_bwd_batched): no atomics, no zero-filled gradient pyramid, no fold passes.
        acc, first = block._open_grad_acc(dense=False)
        if not block._acc_dense:
            block._queue_lookup_grad(gout.contiguous(), coords)
            return (acc if first else None), None
        if first:
            dpyr_ret = acc
    elif want_pyr:
        # One persistent gradient pyramid for ALL refinement iterations: every lookup backward
        # scatter-adds into it; it is handed to autograd exactly once (by the first lookup whose
        # backward runs) and the build backward -- which depends on every lookup -- only runs after
        # the last one has added its share.  The reference instead materialises iters x L dense,
        # zero-filled level-sized tensors (grid_sampler_2d_backward) and sums them.
        acc, first = block._open_grad_acc()
        if first:
            dpyr_ret = acc
    else:
        acc = None
    dcoords = torch.empty_like(coords) if want_dcoords else None
    if acc is None and dcoords is None:
        return None, None
    if acc is None:  # coords gradient only: scatter into a scratch pyramid
        acc = torch.zeros(lay.total_floats, dtype=torch.float32, device=dev)
    gout = gout.contiguous()  # bound to a local: the buffer must outlive the launch below
    check(_lib.lib().raftcorr_lookup_bwd(_ptr(gout), _ptr(coords), _ptr(pyr), _ptr(acc),
                                         ctypes.byref(lay), block.radius, _ptr(dcoords), dev.index, _stream(dev)),
          "raftcorr_lookup_bwd")
    return dpyr_ret, dcoords


class _PyramidLevelsFn(torch.autograd.Function):
    """pyramid buffer -> the reference's ``corr_pyramid`` list (raft/core/corr.py:35-43) as independent tensors.  The
    backward adds each level's gradient into the block's gradient pyramid (row-major gradient layout, whatever layout
    the forward buffer has), under the same hand-over-once rule as the lookups."""

    @staticmethod
    def forward(ctx, pyr, block):
        ctx.block = block
        return tuple(t.clone() for t in block._levels(pyr))

    @staticmethod
    def backward(ctx, *glevels):
        block = ctx.block
        if all(g is None for g in glevels):
            return None, None
        acc, first = block._open_grad_acc()
        glay = block._grad_layout
        BN = block._B * block._H * block._W
        for l, g in enumerate(glevels):
            if g is None:
                continue
            h, w, p = glay.h[l], glay.w[l], glay.pitch[l]
            acc[glay.offset[l]: glay.offset[l] + BN * h * p].view(BN, 1, h, p)[..., :w] += g
        return (acc if first else None), None


class _LookupConvFn(torch.autograd.Function):
    """``relu(conv1x1(lookup(coords)))`` with autograd: the forward is the fused kernel; the backward RECOMPUTES the
    lookup (50 us) instead of keeping the ``[B, L*K*K, H, W]`` tensor alive per refinement iteration (73 MB each at
    config 2 -- the reference's autograd graph holds 12 of them), forms the weight / bias gradients with cuBLAS and
    scatters ``W^T g`` into the block's gradient pyramid with the ordinary lookup adjoint."""

    @staticmethod
    def forward(ctx, pyr, coords, w2, bias, block, relu):
        out = block._lookup_convc1(pyr.detach(), coords, w2.detach(), bias.detach() if bias is not None else None, relu)
        ctx.block, ctx.relu = block, relu
        ctx.save_for_backward(pyr, coords, w2, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, g):
        pyr, coords, w2, out = ctx.saved_tensors
        block = ctx.block
        g = g.contiguous()
        if ctx.relu:
            g = g * (out > 0)
        dw = db = None
        if ctx.needs_input_grad[2]:
            feat = block._lookup(pyr.detach(), coords)
            dw = torch.einsum("bohw,bchw->oc", g, feat)
            del feat
        if ctx.needs_input_grad[3]:
            db = g.sum(dim=(0, 2, 3))
        dpyr_ret = None
        if ctx.needs_input_grad[0]:
            dfeat = torch.einsum("oc,bohw->bchw", w2, g)
            dpyr_ret, _ = _lookup_backward(block, dfeat, coords, pyr, True, False)
        return dpyr_ret, None, dw, db, None, None


class _CorrCore:
    """Everything of a :class:`CorrBlock` except the pyramid tensor: shapes, layouts, the lookup plan, the gradient
    accumulator of a backward pass and the kernel calls (which take the pyramid as an argument).  The autograd nodes keep
    THIS object alive, not the CorrBlock: a node that referenced the block would close a reference cycle
    (block -> pyramid tensor -> grad_fn -> ctx -> block) and every training step's pyramid (0.4 - 2 GB) would live until
    Python's cycle collector happened to run -- measured as two cudaMalloc calls per step and 4.5 ms of host time in an
    eager config-3 step before this split (scripts/eager_profile.py)."""

    def __init__(self, B, D, H, W, device, num_levels, radius, build_mode):
        self.num_levels = int(num_levels)
        self.radius = int(radius)
        if not 1 <= self.radius <= 4:
            raise RaftCorrError(f"CorrBlock: radius must be in [1, 4] (got {radius})")
        mode = build_mode or _default_mode
        if mode not in _MODES:
            raise ValueError(f"unknown build mode {mode!r}")
        if build_mode is None and mode == "f16" and B * H * W < _F16_MIN_PIXELS:
            # small problems (batch-1 inference, evaluate.py:114-120) are launch-bound: the extra operand-scaling launch of
            # the f16 build costs more than its MMAs save (config 1: 0.153 vs 0.195 ms per 12-iteration step, eager)
            mode = "tf32"
        if D > 256 and mode in ("f16", "f16x3"):  # the fp16 operand pass keeps a 64-pixel x D block in registers
            mode = "tf32" if mode == "f16" else "fp32"
        self._mode = _MODES[mode]
        self._B, self._D, self._H, self._W = B, D, H, W
        self._device = device
        # the tensor-core builds write row QUADS interleaved (a lookup window is 3-4 runs of 160 bytes at any x: fewest
        # 64-byte DRAM fills, exact TMA boxes); gradient pyramids are row-major inside the library whatever this says, and
        # never larger.  RAFTCORR_PYRAMID_LAYOUT = quads (default) | pairs (round-1 layout) | rowmajor: A/B switches
        # (rowmajor is needed by the cp.async lookup and the CTA-pair build experiments)
        il = 0 if mode == "fp32" else _LAYOUTS[os.environ.get("RAFTCORR_PYRAMID_LAYOUT", "quads")]
        self._layout = _lib.make_layout(B, H, W, self.num_levels, interleaved=il)
        # gradient pyramids are GEMM operands of the build backward: always plain rows, own level offsets
        self._grad_layout = _lib.make_layout(B, H, W, self.num_levels, interleaved=0) if il else self._layout
        self._grad_acc = None
        self._grad_acc_pass = -1
        self._acc_dense = True      # the accumulator is zero-filled and takes scatter-adds / dense level gradients
        self._acc_folded = False    # its level 0 holds folded partial sums of already flushed lookups
        self._pending = []          # (grad_out, coords) of lookups whose adjoint has not been launched yet
        # batched lookup adjoint: needs a source pixel's gradient pyramid in shared memory (RAFTCORR_LOOKUP_BWD=
        # immediate keeps the per-iteration red.global.add kernel: A/B switch)
        self._defer_ok = (os.environ.get("RAFTCORR_LOOKUP_BWD", "batched") != "immediate" and
                          bool(_lib.lib().raftcorr_lookup_bwd_batched_supported(ctypes.byref(self._layout), self.radius)))
        self._plan = None
        self._plan_ptr = 0
        K = 2 * self.radius + 1
        self._out_shape = (B, self.num_levels * K * K, H, W)
        self._coords_shape = (B, 2, H, W)
        self._dev_index = device.index if device.index is not None else torch.cuda.current_device()
        self._lookup_fn = _lib.lib().raftcorr_lookup_fwd_planned

    def _head_build(self, x, w2, bias):
        lay, dev = self._layout, self._device
        L = _lib.lib()
        Cin = x.shape[1]
        pyr = torch.empty(lay.total_floats, dtype=torch.float32, device=dev)
        ws_bytes = L.raftcorr_fnet_head_build_workspace_bytes(self._B, Cin, self._D, self._H, self._W)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        bias = bias.contiguous() if bias is not None else None  # local: must outlive the launch
        check(L.raftcorr_fnet_head_build_fwd(_ptr(x), _ptr(w2), _ptr(bias),
                                             Cin, self._D, _ptr(pyr), ctypes.byref(lay), _ptr(ws), ws_bytes, dev.index,
                                             _stream(dev)), "raftcorr_fnet_head_build_fwd")
        return pyr

    # -- the single gradient pyramid of a backward pass ---------------------------------------
    def _open_grad_acc(self, dense=True):
        """The block's gradient pyramid (row-major gradient layout) and whether the caller is the FIRST contributor of
        this backward pass.  The first contributor returns it to autograd as d(pyramid) -- exactly once; everybody else
        contributes in place and returns None.  The build backward depends on every contributor, so it runs last.
        ``dense=False`` (queued lookups only): the buffer is not even zero-filled -- the batched adjoint overwrites level
        0 and nothing reads the other levels.  A dense contributor (``corr_pyramid`` gradients, a lookup that also wants
        d/dcoords) turns it into the zero-filled scatter-add accumulator, after flushing what was queued."""
        # a pass that never reached the build backward (exception, interrupt) must not leak its partial sums into the
        # next one: the accumulator is keyed to the autograd engine's graph-task id
        this_pass = torch._C._current_graph_task_id()
        first = self._grad_acc is None or self._grad_acc_pass != this_pass
        if first:
            # sized like the forward buffer (autograd checks the shape); the row-major gradient layout is never larger
            alloc = torch.zeros if dense else torch.empty
            self._grad_acc = alloc(self._layout.total_floats, dtype=torch.float32, device=self._device)
            self._grad_acc_pass = this_pass
            self._acc_dense, self._acc_folded, self._pending = dense, False, []
        elif dense and not self._acc_dense:
            self._flush_lookup_grads(final=False)
            if self._acc_folded:  # level 0 holds folded partial sums: keep them, zero the pooled levels
                if self.num_levels > 1:
                    self._grad_acc[self._grad_layout.offset[1]:].zero_()
            else:
                self._grad_acc.zero_()
            self._acc_dense = True
        return self._grad_acc, first

    def _queue_lookup_grad(self, gout, coords):
        if len(self._pending) == _lib.LOOKUP_BWD_MAX_BATCH:  # bounds the memory held by queued grad_out tensors
            self._flush_lookup_grads(final=False)
        self._pending.append((gout, coords))

    def _flush_lookup_grads(self, final):
        """Launch the batched adjoint of the queued lookups into level 0 of the accumulator."""
        if not self._pending:
            return
        n = len(self._pending)
        L = _lib.lib()
        gptr = (ctypes.c_void_p * n)(*[g.data_ptr() for g, _ in self._pending])
        cptr = (ctypes.c_void_p * n)(*[c.data_ptr() for _, c in self._pending])
        dev = self._device
        rnd = 1 if (final and L.raftcorr_build_bwd_wants_tf32(self._D, self._mode)) else 0
        check(L.raftcorr_lookup_bwd_batched(gptr, cptr, n, _ptr(self._grad_acc), ctypes.byref(self._layout), self.radius,
                                            1 if self._acc_folded else 0, rnd, dev.index, _stream(dev)),
              "raftcorr_lookup_bwd_batched")
        self._acc_folded = True
        self._pending = []  # the launch is stream-ordered: the caching allocator keeps the buffers valid for it

    def _take_grad_acc(self, dpyr):
        """Called by the build backward with whatever autograd delivered for the pyramid buffer.  Returns the gradient
        pyramid and whether its level 0 is already folded (batched lookup adjoint) or the pooled levels still have to be
        folded into it (scatter-add accumulator)."""
        acc = self._grad_acc
        if acc is not None and self._grad_acc_pass != torch._C._current_graph_task_id():
            acc = None  # left over from an aborted pass
        if dpyr is None:
            self._grad_acc, self._pending = None, []
            return None, False
        if acc is None or dpyr.data_ptr() != acc.data_ptr():
            self._grad_acc, self._pending = None, []
            # every contributor in this module (lookups, corr_pyramid, CorrBlock.corr, lookup_convc1) adds into the
            # accumulator and hands it to autograd once, so anything else is a gradient some foreign op attached to the
            # PRIVATE flat buffer, whose layout (row quads interleaved, padded rows) it cannot know
            raise RaftCorrError("CorrBlock backward: unexpected gradient for the internal pyramid buffer; differentiate "
                                "through corr_fn(coords), corr_fn.corr_pyramid or CorrBlock.corr instead")
        folded = not self._acc_dense
        if folded:
            self._flush_lookup_grads(final=True)  # never empty here: a flush always leaves the newest lookup queued
        self._grad_acc, self._pending = None, []  # a second backward pass (retain_graph) starts a fresh accumulator
        return dpyr, folded

    def _build_bwd(self, dpyr, folded, fmap1, fmap2, df1, df2):
        L = _lib.lib()
        dev = self._device
        ws_bytes = L.raftcorr_build_bwd_workspace_bytes(self._B, self._D, self._H, self._W, self._mode)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
        fn = L.raftcorr_build_bwd_folded if folded else L.raftcorr_build_bwd
        check(fn(_ptr(dpyr), ctypes.byref(self._layout), _ptr(fmap1), _ptr(fmap2), self._D, _ptr(df1), _ptr(df2),
                 self._mode, _ptr(ws), ws_bytes, dev.index, _stream(dev)),
              "raftcorr_build_bwd_folded" if folded else "raftcorr_build_bwd")

    # -- kernels ------------------------------------------------------------------------------
    def _build(self, fmap1, fmap2):
        lay, dev = self._layout, self._device
        L = _lib.lib()
        pyr = torch.empty(lay.total_floats, dtype=torch.float32, device=dev)
        ws_bytes = L.raftcorr_build_fwd_workspace_bytes(self._B, self._D, self._H, self._W, self._mode)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
        check(L.raftcorr_build_fwd(_ptr(fmap1), _ptr(fmap2), self._D, _ptr(pyr), ctypes.byref(lay), self._mode,
                                   _ptr(ws), ws_bytes, dev.index, _stream(dev)), "raftcorr_build_fwd")
        return pyr

    def _lookup(self, pyr, coords):
        dev = self._device
        out = torch.empty(self._out_shape, dtype=torch.float32, device=dev)
        # per-pyramid setup (tensor maps, launch geometry) once; every refinement iteration reuses it.  Plain integers
        # go through ctypes' c_void_p argtypes as they are: no wrapper objects on the per-iteration path.
        rc_ = self._lookup_fn(self._ensure_plan(pyr), coords.data_ptr(), out.data_ptr(), self._dev_index,
                              _raw_stream(self._dev_index))
        if rc_:
            check(rc_, "raftcorr_lookup_fwd_planned")
        return out

    def _ensure_plan(self, pyr):
        if self._plan is None or self._plan_ptr != pyr.data_ptr():
            self._plan = ctypes.create_string_buffer(_lib.LOOKUP_PLAN_BYTES)
            check(_lib.lib().raftcorr_lookup_plan_init(self._plan, _ptr(pyr), ctypes.byref(self._layout), self.radius,
                                                       self._device.index), "raftcorr_lookup_plan_init")
            self._plan_ptr = pyr.data_ptr()
        return self._plan

    def _lookup_convc1(self, pyr, coords, w2, bias, relu):
        cout, cin = w2.shape
        dev = self._device
        L = _lib.lib()
        if bias is not None:
            bias = bias.contiguous()
        # packed (TF32-rounded, k-blocked) copy of the weight, redone only when the parameter changes
        key = (w2.data_ptr(), w2._version, cout)
        if getattr(self, "_convw_key", None) != key:
            nbytes = int(L.raftcorr_convc1_weight_bytes(self.num_levels, self.radius, cout))
            packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            wc = w2.contiguous()
            check(L.raftcorr_convc1_pack_weight(_ptr(wc), self.num_levels, self.radius, cout, _ptr(packed), dev.index,
                                                _stream(dev)), "raftcorr_convc1_pack_weight")
            self._convw_key, self._convw = key, packed
        out = torch.empty((self._B, cout, self._H, self._W), dtype=torch.float32, device=dev)
        check(L.raftcorr_lookup_convc1_fwd(self._ensure_plan(pyr), _ptr(coords), _ptr(self._convw), _ptr(bias), cout,
                                           1 if relu else 0, _ptr(out), dev.index, _stream(dev)),
              "raftcorr_lookup_convc1_fwd")
        return out

    def _levels(self, pyr):
        lay = self._layout
        N = self._H * self._W
        out = []
        for l in range(self.num_levels):
            h, w, p = lay.h[l], lay.w[l], lay.pitch[l]
            if lay.interleaved:
                rg = 4 if lay.interleaved == 2 else 2  # rows per interleaved group: layout 1 = pairs, 2 = quads (raftcorr.h)
                hp = (h + rg - 1) // rg
                flat = pyr[lay.offset[l]: lay.offset[l] + self._B * N * hp * rg * p]
                lvl = flat.view(self._B * N, hp, p, rg).permute(0, 1, 3, 2).reshape(self._B * N, 1, rg * hp, p)
                out.append(lvl[:, :, :h, :w])
            else:
                flat = pyr[lay.offset[l]: lay.offset[l] + self._B * N * h * p]
                out.append(flat.view(self._B * N, 1, h, p)[..., :w])
        return out


class CorrBlock:
    """All-pairs correlation volume + pyramid + lookup; drop-in for raft/core/corr.py:13-116.

    ``build_mode`` (extension; RAFT never passes it): ``"f16"`` (default: block-scaled fp16 operands, TF32's 11
    significant bits, fp32 accumulate) | ``"tf32"`` | ``"fp32"``; default from :func:`set_default_build_mode` /
    ``RAFTCORR_BUILD_MODE``.
    """

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, build_mode=None):
        _check_fmaps(fmap1, fmap2, "CorrBlock")
        B, D, H, W = fmap1.shape
        core = self._adopt(_CorrCore(B, D, H, W, fmap1.device, num_levels, radius, build_mode))
        fmap1 = fmap1.contiguous()
        fmap2 = fmap2.contiguous()
        if torch.is_grad_enabled() and (fmap1.requires_grad or fmap2.requires_grad):
            self._pyr = _BuildFn.apply(fmap1, fmap2, core)
        else:
            self._pyr = core._build(fmap1, fmap2)

    def _adopt(self, core):
        """The block's immutable configuration is mirrored on the facade (``__call__`` runs 12-32 times per volume: no
        delegation on that path); mutable state (plan, gradient accumulator) lives in the core only."""
        self._core = core
        for k in ("num_levels", "radius", "_mode", "_B", "_D", "_H", "_W", "_device", "_layout", "_grad_layout", "_defer_ok",
                  "_out_shape", "_coords_shape", "_dev_index"):
            setattr(self, k, getattr(core, k))
        return core

    def __getattr__(self, name):  # anything else private (plan, accumulator state, kernel wrappers) is the core's
        core = self.__dict__.get("_core")
        if core is None or name.startswith("__"):
            raise AttributeError(name)
        return getattr(core, name)

    @classmethod
    def from_encoder_head(cls, x, weight, bias=None, num_levels=4, radius=4):
        """``CorrBlock(*split(conv2(x)))`` in one call -- the feature encoder's output head fused into the build.

        Replaces the tail of the reference encoders, ``x = self.conv2(x)`` + ``torch.split`` (raft/core/extractor.py:
        231,239 BasicEncoder; :341,349 SmallEncoder), the ``.float()`` casts and the ``CorrBlock(fmap1, fmap2)``
        constructor (raft/core/raft.py:114-119): ``x`` is the head's INPUT ``[2B, Cin, H, W]`` (the output of
        ``layer3`` for the concatenated frames, frame-1 images first), ``weight``/``bias`` are ``conv2``'s parameters
        (``[D, Cin, 1, 1]`` or ``[D, Cin]``; ``[D]``).  The 1x1 convolution runs as a tcgen05 TF32 product (fp32
        accumulate -- the precision cuDNN uses for it by default) that writes the build's pixel-major operand buffers
        directly; the NCHW feature maps are never materialised.  TF32 / block-scaled fp16 build; when the configured
        default mode is ``"fp32"`` or ``"f16x3"`` the head runs as ``conv2`` in torch + the ordinary constructor in that
        mode.  Differentiable w.r.t. ``x``,
        ``weight`` and ``bias`` (the backward recomputes ``conv2(x)`` and chains the volume's adjoint through it).
        Valid when the encoder's dropout is inactive (eval mode or ``dropout=0``, extractor.py:234-235).
        Shapes the fused kernel does not take (``Cin % 4``, ``H*W % 4``, ``D > 256``) go through ``conv2`` + the
        ordinary constructor -- still this library's kernels, never a CPU path.
        """
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 4 and x.shape[0] % 2 == 0):
            raise RuntimeError("CorrBlock.from_encoder_head: x must be a CUDA tensor [2B, Cin, H, W]")
        if not (isinstance(weight, torch.Tensor) and weight.is_cuda and weight.dtype == torch.float32):
            raise RuntimeError("CorrBlock.from_encoder_head: weight must be a float32 CUDA tensor")
        if weight.dim() == 4 and tuple(weight.shape[2:]) != (1, 1):
            raise RuntimeError("CorrBlock.from_encoder_head: weight must be a 1x1 convolution kernel")
        B2, Cin, H, W = x.shape
        if weight.dim() not in (2, 4) or weight.shape[1] != Cin:
            raise RuntimeError(f"CorrBlock.from_encoder_head: weight must be [D, {Cin}(, 1, 1)], got {tuple(weight.shape)}")
        D = int(weight.shape[0])
        if bias is not None and not (bias.is_cuda and bias.dtype == torch.float32 and tuple(bias.shape) == (D,)):
            raise RuntimeError("CorrBlock.from_encoder_head: bias must be a float32 CUDA tensor of shape [D]")
        x = x.float().contiguous()  # raft.py:114-115 casts after the head; the fused product wants fp32 input
        w2 = weight.reshape(D, Cin)
        if _default_mode in ("fp32", "f16x3"):
            # the user asked for fp32-class volume accuracy (set_default_build_mode / RAFTCORR_BUILD_MODE): the fused head
            # is a TF32 product, so run conv2 in torch and the ordinary constructor in the configured mode
            f = torch.nn.functional.conv2d(x, w2.reshape(D, Cin, 1, 1), bias)
            return cls(f[: B2 // 2], f[B2 // 2:], num_levels=num_levels, radius=radius, build_mode=_default_mode)
        if not _lib.lib().raftcorr_fnet_head_supported(Cin, D, H, W):
            f = torch.nn.functional.conv2d(x, w2.reshape(D, Cin, 1, 1), bias)
            return cls(f[: B2 // 2], f[B2 // 2:], num_levels=num_levels, radius=radius, build_mode="tf32")
        self = cls.__new__(cls)
        core = self._adopt(_CorrCore(B2 // 2, D, H, W, x.device, num_levels, radius, "tf32"))
        needs_grad = x.requires_grad or w2.requires_grad or (bias is not None and bias.requires_grad)
        if torch.is_grad_enabled() and needs_grad:
            self._pyr = _HeadBuildFn.apply(x, w2, bias, core)
        else:
            self._pyr = core._head_build(x, w2.contiguous(), bias)
        return self

    # -- reference surface ----------------------------------------------------------------------
    def __call__(self, coords):
        if not (type(coords) is torch.Tensor and coords.shape == self._coords_shape and coords.dtype is torch.float32
                and coords.device == self._device and coords.is_contiguous()):
            _check_coords(coords, self._B, self._H, self._W, self._device, "CorrBlock.__call__")  # raises with the reason
            coords = coords.contiguous()
        if torch.is_grad_enabled() and (self._pyr.requires_grad or coords.requires_grad):
            return _LookupFn.apply(self._pyr, coords, self._core)
        return self._core._lookup(self._pyr, coords)

    def lookup_convc1(self, coords, weight, bias=None, relu=True):
        """``relu(conv1x1(self(coords), weight, bias))`` in one kernel -- the lookup fused with its only consumer,
        the motion encoder's first convolution (``F.relu(self.convc1(corr))``, raft/core/update.py:135,170; the
        call it replaces in the refinement loop is raft/core/raft.py:144 + the first line of the encoder).

        ``weight``: ``convc1.weight`` ``[Cout, L*(2r+1)^2, 1, 1]`` (or 2-D), ``bias``: ``convc1.bias`` or None.
        Returns ``[B, Cout, H, W]`` fp32.  The ``[B, L*(2r+1)^2, H, W]`` lookup tensor is never written: its samples
        become TF32 operand tiles in shared memory and the product runs on the tensor cores with fp32 accumulation
        (the precision cuDNN uses for this convolution by default).  Differentiable w.r.t. the feature maps (through
        the pyramid), ``weight`` and ``bias``; the backward recomputes the lookup instead of saving it (coords get no
        gradient: raft.py:143 detaches them).  Cout: multiple of 16 up to 128 or of 32 up to 256 (RAFT-small 96,
        RAFT-basic 256); at most 4 pyramid levels.
        """
        _check_coords(coords, self._B, self._H, self._W, self._device, "CorrBlock.lookup_convc1")
        K = 2 * self.radius + 1
        cin = self.num_levels * K * K
        if not (isinstance(weight, torch.Tensor) and weight.is_cuda and weight.dtype == torch.float32):
            raise RuntimeError("CorrBlock.lookup_convc1: weight must be a float32 CUDA tensor")
        if weight.dim() == 4 and tuple(weight.shape[2:]) != (1, 1):
            raise RuntimeError("CorrBlock.lookup_convc1: weight must be a 1x1 convolution kernel")
        if weight.dim() not in (2, 4) or weight.shape[1] != cin:
            raise RuntimeError(f"CorrBlock.lookup_convc1: weight must be [Cout, {cin}(, 1, 1)], got {tuple(weight.shape)}")
        if self.num_levels > 4:
            raise RaftCorrError("CorrBlock.lookup_convc1: at most 4 pyramid levels")
        cout = int(weight.shape[0])
        if bias is not None:
            if not (bias.is_cuda and bias.dtype == torch.float32 and tuple(bias.shape) == (cout,)):
                raise RuntimeError("CorrBlock.lookup_convc1: bias must be a float32 CUDA tensor of shape [Cout]")
        coords = coords.detach().contiguous()
        w2 = weight.reshape(cout, cin)
        if torch.is_grad_enabled() and (weight.requires_grad or self._pyr.requires_grad or
                                        (bias is not None and bias.requires_grad)):
            return _LookupConvFn.apply(self._pyr, coords, w2, bias, self._core, bool(relu))
        return self._core._lookup_convc1(self._pyr.detach(), coords, w2.detach(), bias.detach() if bias is not None else None, relu)

    @property
    def corr_pyramid(self):
        """The reference's list of ``[B*H*W, 1, H_l, W_l]`` tensors (corr.py:35-43).  fp32 build mode: strided
        views into the single pyramid buffer (rows are padded to a multiple of 8 floats).  Tensor-core build modes:
        the buffer stores row quads interleaved, so the levels are de-interleaved copies (nothing in RAFT reads this
        attribute; the lookup works on the buffer)."""
        if torch.is_grad_enabled() and self._pyr.requires_grad:
            # differentiable like the reference's list (corr.py:35-43): the levels' gradients are written into the
            # block's single row-major gradient pyramid, next to the lookups' contributions
            return list(_PyramidLevelsFn.apply(self._pyr, self._core))
        return self._core._levels(self._pyr.detach())

    @staticmethod
    def corr(fmap1, fmap2, build_mode=None):
        """Level-0 volume ``[B, H, W, 1, H, W]`` = <f1, f2> / sqrt(D)  (raft/core/corr.py:93-116)."""
        blk = CorrBlock(fmap1, fmap2, num_levels=1, radius=1, build_mode=build_mode)
        B, H, W = blk._B, blk._H, blk._W
        return blk.corr_pyramid[0].reshape(B, H, W, 1, H, W)  # differentiable (corr_pyramid)

# ----------------------------------------------------------------------------------------------
# on-the-fly variant
# ----------------------------------------------------------------------------------------------
def _alt_sizes(B, D, H, W, L):
    total = _lib.lib().raftcorr_alt_level_offset(B, D, H, W, L)
    return int(total)


class _AltState:
    """What the autograd nodes of an :class:`AlternateCorrBlock` need in their backward: shapes and the gradient
    accumulators of a pass -- no tensor with a grad_fn, so that keeping it in ``ctx`` closes no reference cycle through the
    block's NHWC feature tensors (see :class:`_CorrCore`)."""

    def __init__(self, B, D, H, W, num_levels, radius, device):
        self._B, self._D, self._H, self._W = B, D, H, W
        self.num_levels, self.radius, self._device = num_levels, radius, device
        self._grad_acc = (None, None)
        self._grad_acc_pass = -1


class _AltPrepFn(torch.autograd.Function):
    """NCHW fmaps -> (fmap1 NHWC, pooled fmap2 pyramid NHWC)   (raft/core/corr.py:136-140,158-159)."""

    @staticmethod
    def forward(ctx, fmap1, fmap2, block):
        ctx.block = block._state
        return block._prep(fmap1, fmap2)

    @staticmethod
    def backward(ctx, df1n, df2p):
        block = ctx.block
        B, D, H, W, L = block._B, block._D, block._H, block._W, block.num_levels
        dev = block._device
        acc1, acc2 = block._grad_acc
        if block._grad_acc_pass != torch._C._current_graph_task_id():
            acc1 = acc2 = None  # left over from an aborted backward pass
        block._grad_acc = (None, None)
        if df1n is None:
            df1n = torch.zeros((B, H, W, D), dtype=torch.float32, device=dev)
        if df2p is None:
            df2p = torch.zeros(_alt_sizes(B, D, H, W, L), dtype=torch.float32, device=dev)
        elif acc2 is None or df2p.data_ptr() != acc2.data_ptr():
            df2p = df2p.contiguous().clone()  # grad_finish folds in place
        df1 = torch.empty((B, D, H, W), dtype=torch.float32, device=dev)
        df2 = torch.empty((B, D, H, W), dtype=torch.float32, device=dev)
        df1n = df1n.contiguous()  # bound to a local: must outlive the launch
        check(_lib.lib().raftcorr_alt_grad_finish(_ptr(df1n), _ptr(df2p), B, D, H, W, L, _ptr(df1),
                                                  _ptr(df2), dev.index, _stream(dev)), "raftcorr_alt_grad_finish")
        return df1, df2, None


class _AltHeadPrepFn(torch.autograd.Function):
    """(x, conv2.weight, conv2.bias) -> (fmap1 NHWC, pooled fmap2 pyramid NHWC): the encoder head fused in front of
    AlternateCorrBlock.__init__.  Backward: fold + transpose the NHWC gradients (raftcorr_alt_grad_finish), then
    conv2's own adjoint through torch autograd."""

    @staticmethod
    def forward(ctx, x, w2, bias, block):
        ctx.block = block._state
        ctx.save_for_backward(x, w2, bias)
        return block._head_prep(x, w2.contiguous(), bias)

    @staticmethod
    def backward(ctx, df1n, df2p):
        x, w2, bias = ctx.saved_tensors
        df1, df2, _ = _AltPrepFn.backward(ctx, df1n, df2p)
        D = ctx.block._D
        with torch.enable_grad():
            xg = x.detach().requires_grad_(ctx.needs_input_grad[0])
            wg = w2.detach().requires_grad_(ctx.needs_input_grad[1])
            bg = bias.detach().requires_grad_(ctx.needs_input_grad[2]) if bias is not None else None
            f = torch.nn.functional.conv2d(xg, wg.reshape(D, -1, 1, 1), bg)
        wanted = [t for t in (xg, wg, bg) if t is not None and t.requires_grad]
        grads = list(torch.autograd.grad(f, wanted, torch.cat([df1, df2]))) if wanted else []
        out = [grads.pop(0) if (t is not None and t.requires_grad) else None for t in (xg, wg, bg)]
        return out[0], out[1], out[2], None


class _AltLookupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f1n, f2p, coords, block):
        ctx.block = block._state
        ctx.save_for_backward(f1n, f2p, coords)
        return block._lookup(f1n, f2p, coords)

    @staticmethod
    def backward(ctx, gout):
        f1n, f2p, coords = ctx.saved_tensors
        block = ctx.block
        B, D, H, W, L = block._B, block._D, block._H, block._W, block.num_levels
        dev = block._device
        acc1, acc2 = block._grad_acc
        this_pass = torch._C._current_graph_task_id()
        first = acc1 is None or block._grad_acc_pass != this_pass
        if first:  # same single-accumulator scheme as _LookupFn.backward, keyed to the engine's graph task
            acc1 = torch.zeros((B, H, W, D), dtype=torch.float32, device=dev)
            acc2 = torch.zeros(_alt_sizes(B, D, H, W, L), dtype=torch.float32, device=dev)
            block._grad_acc = (acc1, acc2)
            block._grad_acc_pass = this_pass
        gout = gout.contiguous()  # bound to a local: must outlive the launch
        check(_lib.lib().raftcorr_alt_lookup_bwd(_ptr(f1n), _ptr(f2p), _ptr(coords), _ptr(gout), B, D, H,
                                                 W, L, block.radius, _ptr(acc1), _ptr(acc2), dev.index, _stream(dev)),
              "raftcorr_alt_lookup_bwd")
        # coords: no gradient, exactly like the reference kernel (correlation_kernel.cu:307 returns zeros
        # and raft.py:143 detaches coords anyway)
        return (acc1, acc2, None, None) if first else (None, None, None, None)


class AlternateCorrBlock:
    """Memory-efficient on-the-fly correlation; drop-in for raft/core/corr.py:119-167.

    Unlike the reference (whose ``alt_cuda_corr.backward`` is never called from Python) this class is
    differentiable w.r.t. both feature maps.

    ``build_mode`` (extension; RAFT never passes it): ``"tf32"`` evaluates blocks of source pixels against the
    bounding box of their windows on the tcgen05 tensor cores (needs ``D % 32 == 0``), ``"fp32"`` computes
    exact fp32 dot products on the CUDA cores.  Default: :func:`set_default_build_mode` /
    ``RAFTCORR_BUILD_MODE`` when the feature dim allows it, else ``"fp32"``.
    """

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, build_mode=None):
        _check_fmaps(fmap1, fmap2, "AlternateCorrBlock")
        mode = build_mode or (_default_mode if fmap1.shape[1] % 32 == 0 else "fp32")
        if mode not in _MODES:
            raise ValueError(f"unknown build mode {mode!r} (expected one of {sorted(_MODES)})")
        if mode == "f16":  # block-scaled fp16 operands exist for the volume build only
            mode = "tf32" if fmap1.shape[1] % 32 == 0 else "fp32"
        if mode == "f16x3":  # the accurate on-the-fly path is the exact-fp32 kernel
            mode = "fp32"
        if mode == "tf32" and fmap1.shape[1] % 32 != 0:
            raise RaftCorrError("AlternateCorrBlock: build_mode='tf32' needs a feature dim that is a multiple of 32")
        self._mode = _MODES[mode]
        self.num_levels = int(num_levels)
        self.radius = int(radius)
        if not 1 <= self.radius <= 4:
            raise RaftCorrError(f"AlternateCorrBlock: radius must be in [1, 4] (got {radius})")
        if not 1 <= self.num_levels <= 4:
            raise RaftCorrError(f"AlternateCorrBlock: num_levels must be in [1, 4] (got {num_levels})")
        B, D, H, W = fmap1.shape
        if (H >> (self.num_levels - 1)) < 1 or (W >> (self.num_levels - 1)) < 1:
            raise RaftCorrError("AlternateCorrBlock: feature maps too small for the number of levels")
        self._B, self._D, self._H, self._W = B, D, H, W
        self._device = fmap1.device
        self._fmaps = (fmap1, fmap2)
        self._state = _AltState(B, D, H, W, self.num_levels, self.radius, fmap1.device)
        fmap1 = fmap1.contiguous()
        fmap2 = fmap2.contiguous()
        if torch.is_grad_enabled() and (fmap1.requires_grad or fmap2.requires_grad):
            self._f1n, self._f2p = _AltPrepFn.apply(fmap1, fmap2, self)
        else:
            self._f1n, self._f2p = self._prep(fmap1, fmap2)

    @classmethod
    def from_encoder_head(cls, x, weight, bias=None, num_levels=4, radius=4):
        """``AlternateCorrBlock(*split(conv2(x)))`` in one call: the encoders' 1x1 output convolution (raft/core/
        extractor.py:231,239 / :341,349) writes the NHWC, TF32-rounded feature tensors of the on-the-fly path directly
        (see :meth:`CorrBlock.from_encoder_head`).  TF32 mode, ``D % 32 == 0``; other shapes run ``conv2`` + the
        ordinary constructor.  Differentiable w.r.t. ``x``, ``weight``, ``bias``."""
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 4 and x.shape[0] % 2 == 0):
            raise RuntimeError("AlternateCorrBlock.from_encoder_head: x must be a CUDA tensor [2B, Cin, H, W]")
        if not (isinstance(weight, torch.Tensor) and weight.is_cuda and weight.dtype == torch.float32 and
                weight.dim() in (2, 4) and weight.shape[1] == x.shape[1] and tuple(weight.shape[2:]) in ((), (1, 1))):
            raise RuntimeError("AlternateCorrBlock.from_encoder_head: weight must be a float32 CUDA tensor [D, Cin(, 1, 1)]")
        B2, Cin, H, W = x.shape
        D = int(weight.shape[0])
        if bias is not None and not (bias.is_cuda and bias.dtype == torch.float32 and tuple(bias.shape) == (D,)):
            raise RuntimeError("AlternateCorrBlock.from_encoder_head: bias must be a float32 CUDA tensor of shape [D]")
        x = x.float().contiguous()
        w2 = weight.reshape(D, Cin)
        if D % 32 != 0 or not _lib.lib().raftcorr_fnet_head_supported(Cin, D, H, W):
            f = torch.nn.functional.conv2d(x, w2.reshape(D, Cin, 1, 1), bias)
            return cls(f[: B2 // 2], f[B2 // 2:], num_levels=num_levels, radius=radius)
        self = cls.__new__(cls)
        self._mode = _MODES["tf32"]
        self.num_levels, self.radius = int(num_levels), int(radius)
        if not 1 <= self.radius <= 4 or not 1 <= self.num_levels <= 4:
            raise RaftCorrError("AlternateCorrBlock: radius and num_levels must be in [1, 4]")
        if (H >> (self.num_levels - 1)) < 1 or (W >> (self.num_levels - 1)) < 1:
            raise RaftCorrError("AlternateCorrBlock: feature maps too small for the number of levels")
        self._B, self._D, self._H, self._W = B2 // 2, D, H, W
        self._device = x.device
        self._fmaps = None
        self._head = (x, w2, bias)
        self._state = _AltState(B2 // 2, D, H, W, self.num_levels, self.radius, x.device)
        needs_grad = x.requires_grad or w2.requires_grad or (bias is not None and bias.requires_grad)
        if torch.is_grad_enabled() and needs_grad:
            self._f1n, self._f2p = _AltHeadPrepFn.apply(x, w2, bias, self)
        else:
            self._f1n, self._f2p = self._head_prep(x, w2.contiguous(), bias)
        return self

    def _head_prep(self, x, w2, bias):
        B, D, H, W, L = self._B, self._D, self._H, self._W, self.num_levels
        dev = self._device
        lib = _lib.lib()
        f1n = torch.empty((B, H, W, D), dtype=torch.float32, device=dev)
        f2p = torch.empty(_alt_sizes(B, D, H, W, L), dtype=torch.float32, device=dev)
        wsb = int(lib.raftcorr_fnet_head_alt_workspace_bytes(x.shape[1], D))
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        bias = bias.contiguous() if bias is not None else None  # local: must outlive the launch
        check(lib.raftcorr_fnet_head_alt_pyramid(_ptr(x), _ptr(w2), _ptr(bias), B,
                                                 x.shape[1], D, H, W, L, _ptr(f1n), _ptr(f2p), _ptr(ws), wsb, dev.index,
                                                 _stream(dev)), "raftcorr_fnet_head_alt_pyramid")
        ws_bytes = int(lib.raftcorr_alt_lookup_workspace_bytes(B, H, W, L, self._mode))
        self._ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        return f1n, f2p

    def _prep(self, fmap1, fmap2):
        B, D, H, W, L = self._B, self._D, self._H, self._W, self.num_levels
        dev = self._device
        f1n = torch.empty((B, H, W, D), dtype=torch.float32, device=dev)
        f2p = torch.empty(_alt_sizes(B, D, H, W, L), dtype=torch.float32, device=dev)
        check(_lib.lib().raftcorr_alt_pyramid(_ptr(fmap1), _ptr(fmap2), B, D, H, W, L, self._mode, _ptr(f1n),
                                              _ptr(f2p), dev.index, _stream(dev)), "raftcorr_alt_pyramid")
        ws_bytes = int(_lib.lib().raftcorr_alt_lookup_workspace_bytes(B, H, W, L, self._mode))
        self._ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        return f1n, f2p

    def _ensure_f16(self, f1n, f2p):
        """fp16 copies of the NHWC tensors (one power-of-two scale per image and tensor), made once per block: the
        tensor-core lookup is bound by its operand bytes.  ``RAFTCORR_ALT_F16=0`` keeps TF32 operands (A/B)."""
        if getattr(self, "_f16", None) is None:
            use = (self._mode == _lib.BUILD_TF32_TC and self._D % 64 == 0 and
                   os.environ.get("RAFTCORR_ALT_F16", "1") != "0")
            if not use:
                self._f16 = False
            else:
                B, D, H, W, L = self._B, self._D, self._H, self._W, self.num_levels
                lib = _lib.lib()
                buf = torch.empty(int(lib.raftcorr_alt_f16_bytes(B, D, H, W, L)), dtype=torch.uint8, device=self._device)
                check(lib.raftcorr_alt_f16_prepare(_ptr(f1n), _ptr(f2p), B, D, H, W, L, _ptr(buf), self._device.index,
                                                   _stream(self._device)), "raftcorr_alt_f16_prepare")
                self._f16 = buf
        return self._f16

    def _lookup(self, f1n, f2p, coords):
        B, D, H, W, L = self._B, self._D, self._H, self._W, self.num_levels
        dev = self._device
        K = 2 * self.radius + 1
        out = torch.empty((B, L * K * K, H, W), dtype=torch.float32, device=dev)
        f16 = self._ensure_f16(f1n, f2p)
        if f16 is not False:
            check(_lib.lib().raftcorr_alt_lookup_fwd_f16(_ptr(f16), _ptr(coords), B, D, H, W, L, self.radius, _ptr(out),
                                                         _ptr(self._ws), self._ws.numel(), dev.index, _stream(dev)),
                  "raftcorr_alt_lookup_fwd_f16")
            return out
        check(_lib.lib().raftcorr_alt_lookup_fwd(_ptr(f1n), _ptr(f2p), _ptr(coords), B, D, H, W, L, self.radius,
                                                 self._mode, _ptr(out), _ptr(self._ws), self._ws.numel(), dev.index,
                                                 _stream(dev)), "raftcorr_alt_lookup_fwd")
        return out

    def __call__(self, coords):
        _check_coords(coords, self._B, self._H, self._W, self._device, "AlternateCorrBlock.__call__")
        coords = coords.contiguous()
        if torch.is_grad_enabled() and (self._f1n.requires_grad or self._f2p.requires_grad):
            return _AltLookupFn.apply(self._f1n, self._f2p, coords, self)
        return self._lookup(self._f1n, self._f2p, coords)

    @property
    def pyramid(self):
        """The reference's ``[(fmap1_l, fmap2_l)]`` list, ``num_levels + 1`` entries, NCHW
        (corr.py:136-140).  Nothing in RAFT reads it; materialised on demand from the inputs."""
        import torch.nn.functional as F

        if self._fmaps is None:  # built from the encoder head: the NCHW maps were never materialised
            x, w2, bias = self._head
            f = F.conv2d(x, w2.reshape(self._D, -1, 1, 1), bias)
            self._fmaps = (f[: self._B], f[self._B:])
        f1, f2 = self._fmaps
        out = [(f1, f2)]
        for _ in range(self.num_levels):
            f1 = F.avg_pool2d(f1, 2, stride=2)
            f2 = F.avg_pool2d(f2, 2, stride=2)
            out.append((f1, f2))
        return out
