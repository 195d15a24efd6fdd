"""Autograd bindings of the fused CUDA kernels (libnicer_b200.so).

Every Function here is once-differentiable *by construction*: the SDF network returns its own analytic
gradient d sdf/dx as a regular output and its backward contains the second-order terms, so no
``create_graph`` double backward is needed (the reference: model/base_networks.py:195-221 +
hashencoder/hashgrid.py:54-134).

Layouts: points are [P,3] row-major; wide per-point tensors are kept feature-major ([rows, P]) inside the
library and exposed to PyTorch as transposed *views* ([P, rows]) so no copy is made between kernels.
"""
import ctypes as C
import os as _os
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import ColorNetT, GridT, LossT, OaJobT, SdfNetT, WnJobT


def lib():
    return _lib.lib()


def check(rc, what=""):
    return _lib.check(rc, what)


def ptr(t, dtype=torch.float32, name="tensor"):
    return _lib.ptr(t, dtype, name)


def stream():
    return _lib.stream()


HIDDEN = 64
F_SDF_ONLY, F_ACCUMULATE, F_NO_FEAT = 1, 2, 4

# ---------------------------------------------------------------------------------------------------------------------
# The grid-gradient scatter (atomics-bound, csrc/grid_scatter.cu) can run on a side stream (NICER_SCATTER_OVERLAP=1).
# SDF nets join before their backward returns (their tables collect a second gradient from the eikonal pass, which
# autograd adds on the main stream); the color grid, used once per forward, joins when the whole backward pass is over
# (an autograd-engine callback), unless its .grad already holds a tensor that AccumulateGrad would add to right away.
# Measured on B200 (demo_2 mapping step): 12.9 ms with the side stream vs 12.2 ms on one stream -- the tcgen05 backward
# kernels take the whole register file of an SM (256 threads x 255 registers), so scatter blocks and tensor-core CTAs
# cannot share an SM and only delay each other.  Off by default.
_SIDE_STREAMS = {}
_PENDING_JOINS = []      # (side stream, tensors the side-stream kernel still reads)


def _scatter_stream(dev):
    if dev.type != "cuda" or _os.environ.get("NICER_SCATTER_OVERLAP", "0") != "1":
        return None
    s = _SIDE_STREAMS.get(dev.index)
    if s is None:
        s = _SIDE_STREAMS[dev.index] = torch.cuda.Stream(device=dev)
    return s


def _sptr(s):
    return C.c_void_p(s.cuda_stream) if s is not None else None


def _flush_joins():
    while _PENDING_JOINS:
        side, _keep = _PENDING_JOINS.pop()
        torch.cuda.current_stream(side.device).wait_stream(side)


def _join(side, keep, defer):
    if side is None:
        return
    if defer:
        if not _PENDING_JOINS:
            torch.autograd.Variable._execution_engine.queue_callback(_flush_joins)
        _PENDING_JOINS.append((side, keep))
    else:
        torch.cuda.current_stream(side.device).wait_stream(side)


@dataclass(frozen=True)
class GridMeta:
    L: int
    C: int
    H: int
    S: float            # log2(per_level_scale)
    divide_factor: float = 1.0


@dataclass(frozen=True)
class SdfMeta:
    grid: GridMeta
    multires: int
    n_hidden: int
    d_out: int

    @property
    def d_in(self):
        return 3 + 6 * self.multires + self.grid.L * self.grid.C


@dataclass(frozen=True)
class ColorMeta:
    grid: GridMeta      # grid.L == 0: no color grid
    multires_view: int
    feature: int
    n_hidden: int
    detached: bool

    @property
    def d_in(self):
        return 3 + (3 + 6 * self.multires_view) + 3 + self.feature + self.grid.L * self.grid.C


def _grid_struct(g, table, offsets):
    s = GridT()
    s.table = ptr(table, name="embeddings") if table is not None else None
    s.offsets = ptr(offsets, torch.int32, "offsets") if offsets is not None else None
    s.L, s.C, s.H, s.S, s.divide_factor = g.L, g.C, g.H, float(g.S), float(g.divide_factor)
    return s


def _sdf_struct(meta, table, offsets, wb):
    n = meta.n_hidden
    assert len(wb) == 2 * (n + 1)
    s = SdfNetT()
    s.grid = _grid_struct(meta.grid, table, offsets)
    s.multires, s.n_hidden, s.d_out = meta.multires, n, meta.d_out
    for l in range(n + 1):
        s.W[l] = ptr(wb[2 * l], name=f"W{l}").value
        s.b[l] = ptr(wb[2 * l + 1], name=f"b{l}").value
    return s


def _color_struct(meta, table, offsets, wb):
    n = meta.n_hidden
    assert len(wb) == 2 * (n + 1)
    s = ColorNetT()
    s.grid = _grid_struct(meta.grid, table, offsets)
    s.multires_view, s.feature, s.n_hidden, s.grid_detached = meta.multires_view, meta.feature, n, int(meta.detached)
    for l in range(n + 1):
        s.W[l] = ptr(wb[2 * l], name=f"W{l}").value
        s.b[l] = ptr(wb[2 * l + 1], name=f"b{l}").value
    return s


def _fm(t):
    """[P, rows] tensor -> contiguous feature-major [rows, P] (no copy when t is a transposed view)."""
    tt = t.t()
    return tt if tt.is_contiguous() else tt.contiguous()


def _zeros_like_many(tensors):
    """Zero tensors shaped like ``tensors`` carved out of ONE flat buffer (one fill kernel instead of one per tensor)."""
    sizes = [t.numel() for t in tensors]
    flat = torch.zeros(sum(sizes), device=tensors[0].device, dtype=tensors[0].dtype)
    out, off = [], 0
    for t, n in zip(tensors, sizes):
        out.append(flat[off:off + n].view(t.shape))
        off += n
    return out


def _c(t):
    """Contiguous version of t (None stays None).  The result must be bound to a name that outlives the library call:
    a temporary passed as ``ptr(_c(t))`` would be freed before the (host-emulated) kernel reads it."""
    if t is None:
        return None
    return t if t.is_contiguous() else t.contiguous()


# Weight / bias gradients are plain [M,P]x[P,N] fp32 GEMMs over the sample dimension (what the reference's autograd
# hands to cuBLAS); nicer_outer_accum is the in-house split-K kernel.
import os as _os
# measured on B200 (demo_2 mapping step): cuBLAS sgemm is ~2x slower than nicer_outer_accum on these M=64, K=4e5
# shapes, so the in-house kernel is the default; NICER_WGRAD=cublas switches for comparison
_WGRAD_LIBRARY_MIN_P = 0 if _os.environ.get("NICER_WGRAD", "") == "cublas" else (1 << 62)


def outer_accum(A, B, Cmat, bias=None, col0=0):
    """Cmat[M, col0:col0+N] += A[M,P] @ B[N,P]^T ; bias[M] += A.sum(1).  A, B feature-major (row stride = P)."""
    M, P = A.shape
    N = B.shape[0]
    if col0:
        cptr = C.c_void_p(_lib.require(Cmat, torch.float32, "C").data_ptr() + 4 * col0)
        check(lib().nicer_outer_accum(ptr(A), A.stride(0), M, ptr(B), B.stride(0), N, P, cptr, Cmat.stride(0),
                                      ptr(bias) if bias is not None else None, stream()), "nicer_outer_accum")
        return
    if A.is_cuda and P >= _WGRAD_LIBRARY_MIN_P:
        Cmat.addmm_(A, B.t())
        if bias is not None:
            bias.add_(A.sum(dim=1))
        return
    check(lib().nicer_outer_accum(ptr(A), A.stride(0), M, ptr(B), B.stride(0), N, P, ptr(Cmat), Cmat.stride(0),
                                  ptr(bias) if bias is not None else None, stream()), "nicer_outer_accum")


def _rows(t, name):
    """[rows, P] operand of a contraction: rows may be strided (a column range of a wider buffer), samples must be contiguous."""
    if t.dim() != 2 or t.stride(1) != 1:
        raise RuntimeError(f"{name} must be a [rows, samples] tensor with contiguous samples")
    _lib.require(t[0], torch.float32, name)      # device / dtype guard on one (contiguous) row
    return t


class OuterAccumBatch:
    """Collects the weight / bias gradient contractions of one network backward (all over the same P samples) and issues
    them through nicer_outer_accum_batch: one kernel launch per 8 jobs instead of one per job."""

    def __init__(self):
        self.jobs, self.keep = [], []

    def add(self, A, B, Cmat, bias=None, col0=0):
        if A.is_cuda and _WGRAD_LIBRARY_MIN_P == 0:        # NICER_WGRAD=cublas comparison path
            outer_accum(A, B, Cmat, bias, col0)
            return
        j = OaJobT()
        j.A, j.lda, j.M = _rows(A, "A").data_ptr(), A.stride(0), A.shape[0]
        j.B, j.ldb, j.N = _rows(B, "B").data_ptr(), B.stride(0), B.shape[0]
        j.C, j.ldc = _lib.require(Cmat, torch.float32, "C").data_ptr() + 4 * col0, Cmat.stride(0)
        j.bias = _lib.require(bias, torch.float32, "bias").data_ptr() if bias is not None else None
        self.jobs.append(j)
        self.keep.append((A, B, Cmat, bias))
        self.P = A.shape[1]

    def flush(self):
        if not self.jobs:
            return
        arr = (OaJobT * len(self.jobs))(*self.jobs)
        check(lib().nicer_outer_accum_batch(arr, len(self.jobs), self.P, stream()), "nicer_outer_accum_batch")
        _lib.launch_count += (len(self.jobs) + 7) // 8 - 1
        self.jobs, self.keep = [], []


# --------------------------------------------------------------------------------------------- weight norm
class WeightNormFn(torch.autograd.Function):
    """(v_0, g_0, v_1, g_1, ...) -> (W_0, W_1, ...): the weight_norm reparametrisation of every Linear of one network
    (torch._weight_norm(v, g, 0), base_networks.py:149-153) in one launch per direction."""

    @staticmethod
    def _jobs(vs, gs, ws=None, norms=None, dws=None, dvs=None, dgs=None):
        arr = (WnJobT * len(vs))()
        for i, (v, g) in enumerate(zip(vs, gs)):
            j = arr[i]
            j.v, j.g = ptr(v, name="weight_v").value, ptr(g, name="weight_g").value
            j.rows, j.cols = v.shape[0], v.numel() // v.shape[0]
            for name, lst in (("w", ws), ("norm", norms), ("dw", dws), ("dv", dvs), ("dg", dgs)):
                if lst is not None:
                    setattr(j, name, ptr(lst[i], name=name).value)
        return arr

    @staticmethod
    def forward(ctx, *vg):
        vs = [_c(t.detach()) for t in vg[0::2]]
        gs = [_c(t.detach()) for t in vg[1::2]]
        ws = [torch.empty_like(v) for v in vs]
        norms = [torch.empty(v.shape[0], device=v.device) for v in vs]
        check(lib().nicer_weight_norm(WeightNormFn._jobs(vs, gs, ws=ws, norms=norms), len(vs), stream()), "nicer_weight_norm")
        ctx.save_for_backward(*vs, *gs, *norms)
        ctx.n = len(vs)
        return tuple(ws)

    @staticmethod
    def backward(ctx, *dws):
        n = ctx.n
        saved = ctx.saved_tensors
        vs, gs, norms = saved[:n], saved[n:2 * n], saved[2 * n:]
        dws = [_c(d) if d is not None else torch.zeros_like(v) for d, v in zip(dws, vs)]
        dvs = [torch.empty_like(v) for v in vs]
        dgs = [torch.empty_like(g) for g in gs]
        check(lib().nicer_weight_norm_backward(WeightNormFn._jobs(vs, gs, norms=norms, dws=dws, dvs=dvs, dgs=dgs), n, stream()),
              "nicer_weight_norm_backward")
        out = []
        for dv, dg in zip(dvs, dgs):
            out += [dv, dg]
        return tuple(out)


# --------------------------------------------------------------------------------------------- SDF network
def _sdf_forward_impl(ctx, x, Pf, table, offsets, meta, want_feat, wb):
    """x [P,3]: all points; the first Pf of them carry features (Pf == P for the plain call).  Returns (sdf [P], feat_fm [64,Pf] |
    None, grad [P,3]) and saves what the backward needs on ctx."""
    wb = tuple(_c(t.detach()) for t in wb)
    table_d = table.detach()
    P = x.shape[0]
    n = meta.n_hidden
    dev = x.device
    sdf = torch.empty(P, device=dev)
    feat_fm = torch.empty(HIDDEN, Pf, device=dev) if want_feat else None
    grad = torch.empty(P, 3, device=dev)
    Z = torch.empty(n * HIDDEN, P, device=dev)
    R = torch.empty((n - 1) * HIDDEN, P, device=dev) if n > 1 else None
    DYDX = torch.empty(meta.grid.L * 3 * meta.grid.C, P, device=dev)
    H0 = torch.empty(meta.d_in, P, device=dev)      # network input, kept for the layer-0 weight gradient
    net = _sdf_struct(meta, table_d, offsets, wb)
    flags = 0 if want_feat else F_NO_FEAT
    check(lib().nicer_sdf_forward(C.byref(net), ptr(x), P, Pf, flags, ptr(sdf), ptr(feat_fm), ptr(grad), ptr(Z),
                                  ptr(R), ptr(DYDX), ptr(H0), stream()), "nicer_sdf_forward")
    ctx.meta, ctx.want_feat, ctx.Pf = meta, want_feat, Pf
    ctx.save_for_backward(x, table_d, offsets, Z, R, DYDX, H0, *wb)
    ctx.set_materialize_grads(False)
    return sdf, feat_fm, grad


def _sdf_backward_impl(ctx, g_sdf, g_feat, gg, need_x, need_table, need_w):
    """g_sdf [Pf,1] | None, g_feat [Pf,F] | None (upstream gradients of the first Pf points), gg [P,3] | None.
    Returns (grad_x [P,3] | None, grad_table | None, [dW_0, db_0, ...] | None)."""
    x, table, offsets, Z, R, DYDX, H0, *wb = ctx.saved_tensors
    meta, Pf = ctx.meta, ctx.Pf
    n, P, dev = meta.n_hidden, x.shape[0], x.device
    nfeat = meta.d_out - 1
    gs = _c(g_sdf.reshape(Pf)) if g_sdf is not None else None
    gf = None
    if g_feat is not None and ctx.want_feat:
        gf = _fm(g_feat)
        if nfeat < HIDDEN:
            pad = torch.zeros(HIDDEN, Pf, device=dev)
            pad[:nfeat] = gf
            gf = pad
    grad_x = torch.zeros(P, 3, device=dev) if need_x else None
    # pose-only tracking (detached parameters): no table scatter, no weight-gradient contractions
    grad_table = torch.zeros_like(table) if need_table else None
    ZB = torch.empty(n * HIDDEN, P, device=dev)
    QB, AB, TAN = torch.empty_like(ZB), torch.empty_like(ZB), torch.empty_like(ZB)
    T0 = torch.empty(meta.d_in, P, device=dev)
    GY = torch.empty(2 * meta.grid.L * meta.grid.C, P, device=dev)     # MLP-backward -> grid-scatter hand-over
    net = _sdf_struct(meta, table, offsets, wb)
    side = _scatter_stream(dev)
    zeros = _zeros_like_many(list(wb)) if need_w else None        # dW_0, db_0, dW_1, ... in one buffer
    # the tangent kernel adds sum_p tan_n (second-order part of the sdf row of dW_n) straight into dW_n[0, :]
    tan_sum = zeros[2 * n][0] if need_w else None
    check(lib().nicer_sdf_backward(C.byref(net), ptr(x), P, Pf, ptr(Z), ptr(R), ptr(DYDX), ptr(H0), ptr(gs), ptr(gf), ptr(gg),
                                   ptr(grad_x), ptr(grad_table), ptr(ZB), ptr(QB), ptr(AB), ptr(TAN),
                                   ptr(T0), ptr(tan_sum), ptr(GY), stream(), _sptr(side)), "nicer_sdf_backward")
    if not need_w:
        _join(side, None, defer=False)
        return grad_x, grad_table, None
    grads = []
    oa = OuterAccumBatch()
    oa_f = OuterAccumBatch()        # contractions over the first Pf points only (upstream sdf / feature gradients)
    for l in range(n + 1):
        dW, db = zeros[2 * l], zeros[2 * l + 1]
        if l == 0:
            oa.add(ZB[:HIDDEN], H0, dW, db)
            oa.add(QB[:HIDDEN], T0, dW)
        elif l < n:
            oa.add(ZB[l * HIDDEN:(l + 1) * HIDDEN], AB[(l - 1) * HIDDEN:l * HIDDEN], dW, db)
            oa.add(QB[l * HIDDEN:(l + 1) * HIDDEN], TAN[(l - 1) * HIDDEN:l * HIDDEN], dW)
        else:
            a_n = AB[(n - 1) * HIDDEN:, :Pf]        # rows of stride P, the first Pf samples of each
            if gs is not None:
                oa_f.add(gs.view(1, Pf), a_n, dW[:1], db[:1])
            if gf is not None and nfeat > 0:
                oa_f.add(gf[:nfeat], a_n, dW[1:], db[1:])
        grads += [dW, db]
    oa.flush()
    oa_f.flush()
    _join(side, None, defer=False)       # the scatter overlapped with the weight-gradient GEMMs above
    return grad_x, grad_table, grads


class SdfNetFn(torch.autograd.Function):
    """(x, table, W0,b0,...,Wn,bn) -> (sdf [P,1], feat [P,F] (view of [F,P]), grad [P,3])."""

    @staticmethod
    def forward(ctx, x, table, offsets, meta, want_feat, *wb):
        x = _c(x.detach())
        P = x.shape[0]
        sdf, feat_fm, grad = _sdf_forward_impl(ctx, x, P, table, offsets, meta, want_feat, wb)
        nfeat = meta.d_out - 1
        feat = feat_fm[:nfeat].t() if want_feat else torch.zeros(P, 0, device=x.device)
        return sdf.view(P, 1), feat, grad

    @staticmethod
    def backward(ctx, g_sdf, g_feat, g_grad):
        n_wb = len(ctx.saved_tensors) - 7
        gg = _c(g_grad) if g_grad is not None else None
        grad_x, grad_table, grads = _sdf_backward_impl(ctx, g_sdf, g_feat, gg, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                                       any(ctx.needs_input_grad[5:]))
        return (grad_x, grad_table, None, None, None, *(grads if grads is not None else [None] * n_wb))


class SdfNetPairFn(torch.autograd.Function):
    """(x [P1,3], x2 [P2,3], table, ...) -> (sdf [P1,1], feat [P1,F], grad [P1,3], grad2 [P2,3]): the main-pass points and a
    second set that only needs d sdf/dx (the eikonal samples, network.py:313-336) in ONE set of launches -- 13 instead of 14 waves of
    128-point tile pairs at the demo_2 shape, one staging of the weights, one weight-gradient batch, no second accumulation of
    every parameter gradient.  The second set takes no part in the feature head (nicer_sdf_forward's P_feat)."""

    @staticmethod
    def forward(ctx, x, x2, table, offsets, meta, want_feat, *wb):
        P1, P2 = x.shape[0], x2.shape[0]
        xa = torch.cat([x.detach(), x2.detach()], 0)
        sdf, feat_fm, grad = _sdf_forward_impl(ctx, xa, P1, table, offsets, meta, want_feat, wb)
        ctx.P1 = P1
        nfeat = meta.d_out - 1
        feat = feat_fm[:nfeat].t() if want_feat else torch.zeros(P1, 0, device=xa.device)
        return sdf[:P1].view(P1, 1), feat, grad[:P1], grad[P1:]

    @staticmethod
    def backward(ctx, g_sdf, g_feat, g_grad, g_grad2):
        n_wb = len(ctx.saved_tensors) - 7
        x = ctx.saved_tensors[0]
        P, P1 = x.shape[0], ctx.P1
        gg = None
        if g_grad is not None or g_grad2 is not None:
            gg = torch.empty(P, 3, device=x.device)
            if g_grad is not None:
                gg[:P1] = g_grad
            else:
                gg[:P1].zero_()
            if g_grad2 is not None:
                gg[P1:] = g_grad2
            else:
                gg[P1:].zero_()
        need_x = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        grad_x, grad_table, grads = _sdf_backward_impl(ctx, g_sdf, g_feat, gg, need_x, ctx.needs_input_grad[2],
                                                       any(ctx.needs_input_grad[6:]))
        gx = grad_x[:P1] if (grad_x is not None and ctx.needs_input_grad[0]) else None
        gx2 = grad_x[P1:] if (grad_x is not None and ctx.needs_input_grad[1]) else None
        return (gx, gx2, grad_table, None, None, None, *(grads if grads is not None else [None] * n_wb))


def sdf_values(x, nets, out=None):
    """No-grad SDF of the sum of networks (get_sdf_vals; the sampler's 640-sample pass).
    nets: list of (meta, table, offsets, wb)."""
    x = _c(x.detach())
    P = x.shape[0]
    sdf = out if out is not None else torch.empty(P, device=x.device)
    # scratch for the grid features of one net at a time (the gathers run in their own high-occupancy kernel)
    rows = max(meta.grid.L * meta.grid.C for meta, _t, _o, _w in nets)
    feat_ws = torch.empty(rows, P, device=x.device) if x.is_cuda else None
    for i, (meta, table, offsets, wb) in enumerate(nets):
        wb = tuple(_c(t.detach()) for t in wb)
        net = _sdf_struct(meta, table.detach(), offsets, wb)
        flags = F_SDF_ONLY | (F_ACCUMULATE if i > 0 else 0)
        check(lib().nicer_sdf_forward(C.byref(net), ptr(x), P, 0, flags, ptr(sdf), None, None, None, None, None,
                                      ptr(feat_ws), stream()), "nicer_sdf_forward")
    return sdf.view(P, 1)


# --------------------------------------------------------------------------------------------- color network
class ColorNetFn(torch.autograd.Function):
    """(x, view, normals, feat [P,F], table, W..,b..) -> rgb [P,3]."""

    @staticmethod
    def forward(ctx, x, view, normals, feat, table, offsets, meta, *wb):
        x, view, normals = _c(x.detach()), _c(view.detach()), _c(normals.detach())
        feat_fm = _fm(feat.detach())
        wb = tuple(_c(t.detach()) for t in wb)
        table_d = table.detach() if table is not None else None
        P, dev, n = x.shape[0], x.device, meta.n_hidden
        has_grid = table_d is not None
        want_dx = has_grid and (not meta.detached) and ctx.needs_input_grad[0]
        rgb = torch.empty(P, 3, device=dev)
        A_fm = torch.empty(n * HIDDEN, P, device=dev)
        DYDX = torch.empty(meta.grid.L * 3 * meta.grid.C, P, device=dev) if want_dx else None
        # network input kept for the layer-0 weight gradient; the kernel writes every row but the feature block
        # [33, 33+F), which is feat_fm itself
        H0 = torch.empty(meta.d_in, P, device=dev)
        net = _color_struct(meta, table_d, offsets, wb)
        check(lib().nicer_color_forward(C.byref(net), ptr(x), ptr(view), ptr(normals), ptr(feat_fm), P, ptr(rgb),
                                        ptr(A_fm), ptr(DYDX), ptr(H0), stream()), "nicer_color_forward")
        ctx.meta, ctx.has_grid = meta, has_grid
        ctx.table_leaf = table if (table is not None and table.is_leaf) else None
        ctx.save_for_backward(x, view, normals, feat_fm, table_d, offsets, rgb, A_fm, DYDX, H0, *wb)
        return rgb

    @staticmethod
    def backward(ctx, g_rgb):
        x, view, normals, feat_fm, table, offsets, rgb, A_fm, DYDX, H0, *wb = ctx.saved_tensors
        meta = ctx.meta
        P, dev, n = x.shape[0], x.device, meta.n_hidden
        g_rgb = _c(g_rgb)
        grad_x = torch.zeros(P, 3, device=dev) if ctx.needs_input_grad[0] else None
        grad_view = torch.empty(P, 3, device=dev) if ctx.needs_input_grad[1] else None
        grad_normals = torch.empty(P, 3, device=dev)
        grad_feat_fm = torch.empty(meta.feature, P, device=dev)
        has_gy = ctx.has_grid and not meta.detached
        scatter = has_gy and ctx.needs_input_grad[4]
        need_w = any(ctx.needs_input_grad[7:])
        # (zeroing this 1.06 GB buffer early on a side stream was measured: no gain under graph replay, 7.57 vs 7.55 ms)
        grad_table = torch.zeros_like(table) if scatter else None
        ZB = torch.empty(n * HIDDEN, P, device=dev)
        OB = torch.empty(3, P, device=dev)
        GY = torch.empty(meta.grid.L * meta.grid.C, P, device=dev) if has_gy else None
        net = _color_struct(meta, table, offsets, wb)
        side = _scatter_stream(dev) if scatter else None
        check(lib().nicer_color_backward(C.byref(net), ptr(x), ptr(view), ptr(normals), ptr(feat_fm), P, ptr(rgb),
                                         ptr(A_fm), ptr(DYDX), ptr(g_rgb), ptr(grad_x), ptr(grad_view),
                                         ptr(grad_normals), ptr(grad_feat_fm), ptr(grad_table), ptr(ZB), ptr(OB),
                                         ptr(GY), stream(), _sptr(side)), "nicer_color_backward")
        if not need_w:
            _join(side, None, defer=False)
            return (grad_x, grad_view, grad_normals, grad_feat_fm.t(), grad_table, None, None, *([None] * len(wb)))
        grads = []
        oa = OuterAccumBatch()
        zeros = _zeros_like_many(list(wb))
        for l in range(n + 1):
            W = wb[2 * l]
            dW, db = zeros[2 * l], zeros[2 * l + 1]
            if l == 0:
                # input = [x, PE(view), normals (33) | feat (F) | grid]: the feature block reads feat_fm in place, the
                # forward kernel only materialised the 33 + L*C other rows of H0
                nf = meta.feature
                oa.add(ZB[:HIDDEN], H0[:33], dW, db)
                oa.add(ZB[:HIDDEN], feat_fm, dW, None, col0=33)
                if meta.d_in > 33 + nf:
                    oa.add(ZB[:HIDDEN], H0[33 + nf:], dW, None, col0=33 + nf)
            elif l < n:
                oa.add(ZB[l * HIDDEN:(l + 1) * HIDDEN], A_fm[(l - 1) * HIDDEN:l * HIDDEN], dW, db)
            else:
                oa.add(OB, A_fm[(n - 1) * HIDDEN:], dW, db)
            grads += [dW, db]
        oa.flush()
        # the color grid is used once per forward: its gradient is only read after the backward pass, so the scatter may
        # keep running under the SDF backward kernels -- unless AccumulateGrad is about to add to an existing .grad
        leaf = ctx.table_leaf
        _join(side, (GY, x, grad_table), defer=leaf is not None and leaf.grad is None)
        return (grad_x, grad_view, grad_normals, grad_feat_fm.t(), grad_table, None, None, *grads)


# --------------------------------------------------------------------------------------------- compositing
class CompositeFn(torch.autograd.Function):
    """(sdf [P,1], x [P,3], z [R,S], rgb [P,3], grad [P,3], voxels) ->
    (weights [R,S], rgb_values [R,3], depth [R,1], normal_map [R,3] (before rotation))."""

    @staticmethod
    def forward(ctx, sdf, x, z, rgb, grad, voxels):
        R, S = z.shape
        sdf_, x_, z_, rgb_, grad_ = (_c(t.detach()) for t in (sdf.reshape(-1), x, z, rgb, grad))
        vox = _c(voxels.detach())
        dev = z.device
        weights = torch.empty(R, S, device=dev)
        rgb_out = torch.empty(R, 3, device=dev)
        depth = torch.empty(R, device=dev)
        normal = torch.empty(R, 3, device=dev)
        wsum = torch.empty(R, device=dev)
        check(lib().nicer_composite_forward(ptr(sdf_), ptr(x_), ptr(z_), ptr(rgb_), ptr(grad_), ptr(vox), vox.shape[0],
                                            R, S, ptr(weights), ptr(rgb_out), ptr(depth), ptr(normal), ptr(wsum),
                                            stream()), "nicer_composite_forward")
        ctx.save_for_backward(sdf_, x_, z_, rgb_, grad_, vox, weights, depth, wsum)
        ctx.set_materialize_grads(False)
        ctx.sdf_shape = sdf.shape
        return weights, rgb_out, depth.view(R, 1), normal

    @staticmethod
    def backward(ctx, g_w, g_rgb_out, g_depth, g_normal):
        sdf_, x_, z_, rgb_, grad_, vox, weights, depth, wsum = ctx.saved_tensors
        R, S = z_.shape
        dev = z_.device
        g_sdf = torch.empty(R * S, device=dev)
        g_rgb = torch.empty(R * S, 3, device=dev)
        g_grad = torch.empty(R * S, 3, device=dev)
        gd = _c(g_depth.reshape(R)) if g_depth is not None else None
        g_rgb_out, g_normal, g_w = _c(g_rgb_out), _c(g_normal), _c(g_w)
        check(lib().nicer_composite_backward(ptr(sdf_), ptr(x_), ptr(z_), ptr(rgb_), ptr(grad_), ptr(vox),
                                             vox.shape[0], R, S, ptr(weights), ptr(depth), ptr(wsum), ptr(g_rgb_out),
                                             ptr(gd), ptr(g_normal), ptr(g_w), ptr(g_sdf), ptr(g_rgb), ptr(g_grad),
                                             stream()), "nicer_composite_backward")
        return g_sdf.view(ctx.sdf_shape), None, None, g_rgb, g_grad, None


def sampler_weights(sdf, x, z, voxels):
    R, S = z.shape
    w = torch.empty(R, S, device=z.device)
    vox, sdf, x, z = _c(voxels), _c(sdf.reshape(-1)), _c(x), _c(z)
    check(lib().nicer_sampler_weights(ptr(sdf), ptr(x), ptr(z), ptr(vox), vox.shape[0], R, S,
                                      ptr(w), stream()), "nicer_sampler_weights")
    return w


def sampler_uniform(cam_loc, ray_dirs, near, far_cap, bound, use_cube, rnd, N):
    """Coarse depths z [R,N], far [R,1] and the sample points [R*N,3] in one kernel (UniformSampler.get_z_vals)."""
    cam_loc, ray_dirs, rnd = _c(cam_loc), _c(ray_dirs), _c(rnd)
    R, dev = ray_dirs.shape[0], ray_dirs.device
    z = torch.empty(R, N, device=dev)
    far = torch.empty(R, 1, device=dev)
    points = torch.empty(R * N, 3, device=dev)
    check(lib().nicer_sampler_uniform(ptr(cam_loc), ptr(ray_dirs), float(near), float(far_cap), float(bound), int(bool(use_cube)),
                                      ptr(rnd), R, N, ptr(z), ptr(far), ptr(points), stream()), "nicer_sampler_uniform")
    return z, far, points


def sampler_resample(sdf, points, z, voxels, N, sel, near, far, eik_idx):
    """weights -> pdf -> cdf -> inverse CDF -> merge with near / far / z[:, sel] -> sort -> eikonal pick, one warp per ray
    (ImportantSampler.get_z_vals after the SDF pass).  Returns (z_out [R, N+2+len(sel)], z_eik [R,1])."""
    R, U = z.shape
    dev = z.device
    n_extra = 0 if sel is None else int(sel.shape[0])
    z_out = torch.empty(R, N + 2 + n_extra, device=dev)
    z_eik = torch.empty(R, 1, device=dev)
    sdf, points, z, vox, far = _c(sdf.reshape(-1)), _c(points), _c(z), _c(voxels), _c(far.reshape(-1))
    sel = _c(sel.to(torch.int64)) if sel is not None else None
    eik_idx = _c(eik_idx.to(torch.int64))
    check(lib().nicer_sampler_resample(ptr(sdf), ptr(points), ptr(z), ptr(vox), vox.shape[0], R, U, N,
                                       ptr(sel, torch.int64, "sel"), n_extra, float(near), ptr(far),
                                       ptr(eik_idx, torch.int64, "eik_idx"), ptr(z_out), ptr(z_eik), None, stream()),
          "nicer_sampler_resample")
    return z_out, z_eik


def voxel_count(x, voxels):
    """In-place: voxels[cell(x)] += 1 for points with all |x_i| <= 0.99 (SLAMNetwork.update_voxels)."""
    x = _c(x.detach())
    check(lib().nicer_voxel_count(ptr(x), x.shape[0], ptr(voxels), voxels.shape[0], stream()), "nicer_voxel_count")


# --------------------------------------------------------------------------------------------- drop-in hash op
class _HashEncode(torch.autograd.Function):
    """hashencoder/hashgrid.py:13-69 on top of the C-ABI drop-in op (reference layouts)."""

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False):
        inputs, embeddings, offsets = _c(inputs), _c(embeddings), _c(offsets)
        B, D = inputs.shape
        L, Cc = offsets.shape[0] - 1, embeddings.shape[1]
        S, H = float(np.log2(per_level_scale)), int(base_resolution)
        outputs = torch.empty(L, B, Cc, device=inputs.device)
        dy_dx = torch.empty(B, L * D * Cc, device=inputs.device) if calc_grad_inputs else torch.empty(1, device=inputs.device)
        check(lib().nicer_hash_encode_forward(ptr(inputs), ptr(embeddings), ptr(offsets, torch.int32), ptr(outputs), B,
                                              D, Cc, L, S, H, int(calc_grad_inputs), ptr(dy_dx), stream()),
              "nicer_hash_encode_forward")
        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.dims = (B, D, Cc, L, S, H, bool(calc_grad_inputs))
        return outputs.permute(1, 0, 2).reshape(B, L * Cc)

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, Cc, L, S, H, cgi = ctx.dims
        grad = grad.view(B, L, Cc).permute(1, 0, 2).contiguous()
        gi, ge = _HashEncodeBackward.apply(grad, inputs, embeddings, offsets, dy_dx, ctx.dims)
        return (gi if cgi else None), ge, None, None, None, None


class _HashEncodeBackward(torch.autograd.Function):
    """hashencoder/hashgrid.py:79-134: first backward as a differentiable op; its backward is K4+K5."""

    @staticmethod
    def forward(ctx, grad, inputs, embeddings, offsets, dy_dx, dims):
        B, D, Cc, L, S, H, cgi = dims
        grad_inputs = torch.zeros_like(inputs)
        grad_embeddings = torch.zeros_like(embeddings)
        check(lib().nicer_hash_encode_backward(ptr(grad), ptr(inputs), ptr(embeddings), ptr(offsets, torch.int32),
                                               ptr(grad_embeddings), B, D, Cc, L, S, H, int(cgi), ptr(dy_dx),
                                               ptr(grad_inputs), stream()), "nicer_hash_encode_backward")
        ctx.save_for_backward(grad, inputs, embeddings, offsets, dy_dx)
        ctx.dims = dims
        return grad_inputs, grad_embeddings

    @staticmethod
    def backward(ctx, ggi, _gge):
        grad, inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, Cc, L, S, H, cgi = ctx.dims
        if not cgi:
            raise RuntimeError("hash_encode: second backward requires calc_grad_inputs (dy_dx was not computed)")
        grad_grad = torch.zeros_like(grad)
        grad2_embeddings = torch.zeros_like(embeddings)
        ggi = _c(ggi)
        check(lib().nicer_hash_encode_second_backward(ptr(grad), ptr(inputs), ptr(embeddings),
                                                      ptr(offsets, torch.int32), B, D, Cc, L, S, H, int(cgi), ptr(dy_dx),
                                                      ptr(ggi), ptr(grad_grad), ptr(grad2_embeddings), stream()),
              "nicer_hash_encode_second_backward")
        return grad_grad, None, grad2_embeddings, None, None, None


hash_encode = _HashEncode.apply


# --------------------------------------------------------------------------------------------- camera / rays
class PoseFromCam7Fn(torch.autograd.Function):
    """cam7 [B,7] (quat wxyz un-normalised, translation) -> c2w [B,4,4]  (get_camera_from_tensor, general.py:52-100)."""

    @staticmethod
    def forward(ctx, cam7):
        cam7 = _c(cam7.detach())
        B = cam7.shape[0]
        pose = torch.empty(B, 4, 4, device=cam7.device)
        check(lib().nicer_pose_from_cam7(ptr(cam7), B, ptr(pose), stream()), "nicer_pose_from_cam7")
        ctx.save_for_backward(cam7)
        return pose

    @staticmethod
    def backward(ctx, g_pose):
        (cam7,) = ctx.saved_tensors
        g = torch.empty_like(cam7)
        g_pose = _c(g_pose)
        check(lib().nicer_pose_from_cam7_backward(ptr(cam7), ptr(g_pose), cam7.shape[0], ptr(g), stream()),
              "nicer_pose_from_cam7_backward")
        return g


class Inv4x4Fn(torch.autograd.Function):
    """A [B,4,4] -> A^-1: the w2c matrices of the flow / warp blocks (torch.inverse(pose), network.py:157,171), one kernel per
    direction instead of the ~15 launches of the LU path and its matmul backward."""

    @staticmethod
    def forward(ctx, A):
        A = _c(A.detach())
        Ai = torch.empty_like(A)
        check(lib().nicer_inv4x4(ptr(A), A.shape[0], ptr(Ai), stream()), "nicer_inv4x4")
        ctx.save_for_backward(Ai)
        return Ai

    @staticmethod
    def backward(ctx, G):
        (Ai,) = ctx.saved_tensors
        G = _c(G)
        GA = torch.empty_like(Ai)
        check(lib().nicer_inv4x4_backward(ptr(Ai), ptr(G), Ai.shape[0], ptr(GA), stream()), "nicer_inv4x4_backward")
        return GA


def inv4x4(A):
    return Inv4x4Fn.apply(A.reshape(-1, 4, 4)).reshape(A.shape)


class CameraRaysFn(torch.autograd.Function):
    """(uv [B,N,2], pose [B,4,4], K [B,4,4]) -> (ray_dirs [B,N,3], cam_loc [B,3])  (get_camera_params,
    rend_util.py:68-93).  Differentiable w.r.t. the pose only, like the reference's use of it."""

    @staticmethod
    def forward(ctx, uv, pose, K):
        uv, pose, K = _c(uv.detach()), _c(pose.detach()), _c(K.detach())
        B, N = uv.shape[0], uv.shape[1]
        dirs = torch.empty(B, N, 3, device=uv.device)
        loc = torch.empty(B, 3, device=uv.device)
        check(lib().nicer_camera_rays(ptr(uv), ptr(pose), ptr(K), B, N, ptr(dirs), ptr(loc), stream()), "nicer_camera_rays")
        ctx.save_for_backward(uv, pose, K)
        ctx.set_materialize_grads(False)
        return dirs, loc

    @staticmethod
    def backward(ctx, g_dirs, g_loc):
        uv, pose, K = ctx.saved_tensors
        B, N = uv.shape[0], uv.shape[1]
        if g_dirs is None:
            g_dirs = torch.zeros(B, N, 3, device=uv.device)
        g_pose = torch.empty(B, 4, 4, device=uv.device)
        g_dirs, g_loc = _c(g_dirs), _c(g_loc)
        check(lib().nicer_camera_rays_backward(ptr(uv), ptr(pose), ptr(K), B, N, ptr(g_dirs), ptr(g_loc), ptr(g_pose), stream()),
              "nicer_camera_rays_backward")
        return None, g_pose, None


class RayPointsFn(torch.autograd.Function):
    """(cam_loc [R,3], ray_dirs [R,3], z [R,S]) -> (points [R*S,3], dirs_flat [R*S,3])  (network.py:112-117)."""

    @staticmethod
    def forward(ctx, cam_loc, ray_dirs, z):
        cam_loc, ray_dirs, z = _c(cam_loc.detach()), _c(ray_dirs.detach()), _c(z.detach())
        R, S = z.shape
        points = torch.empty(R * S, 3, device=z.device)
        dirs_flat = torch.empty(R * S, 3, device=z.device)
        check(lib().nicer_ray_points(ptr(cam_loc), ptr(ray_dirs), ptr(z), R, S, ptr(points), ptr(dirs_flat), stream()),
              "nicer_ray_points")
        ctx.save_for_backward(z)
        ctx.set_materialize_grads(False)
        return points, dirs_flat

    @staticmethod
    def backward(ctx, g_points, g_dirs_flat):
        (z,) = ctx.saved_tensors
        R, S = z.shape
        g_loc = torch.empty(R, 3, device=z.device)
        g_dirs = torch.empty(R, 3, device=z.device)
        g_points, g_dirs_flat = _c(g_points), _c(g_dirs_flat)
        check(lib().nicer_ray_points_backward(ptr(z), R, S, ptr(g_points), ptr(g_dirs_flat), ptr(g_loc),
                                              ptr(g_dirs), stream()), "nicer_ray_points_backward")
        return g_loc, g_dirs, None


# --------------------------------------------------------------------------------------------- loss terms
LOSS_RGB, LOSS_DEPTH, LOSS_GT_DEPTH, LOSS_NORMAL_L1, LOSS_NORMAL_COS, LOSS_EIKONAL, LOSS_SMOOTH, LOSS_SUM, LOSS_TERMS = range(9)


class SlamLossFn(torch.autograd.Function):
    """The ray / eikonal-point terms of SLAMLoss (loss.py:113-233) in three kernels.

    forward(rgb_pred, depth_pred, normal_pred, grad_theta, grad_theta_nei, consts) -> (weighted sum of the terms,
    terms [LOSS_TERMS] (unweighted, not differentiable)).  ``consts``: dict with sdf, mask_gt, rgb_gt, depth_gt, gt_depth,
    gt_depth_valid, normal_gt (tensors or None), B, N, depth_mask_all and the weights w_*; a None prediction switches
    its term off.  The gradients are computed in the forward kernels (they need nothing from upstream but a scale)."""

    @staticmethod
    def forward(ctx, rgb_pred, depth_pred, normal_pred, grad_theta, grad_theta_nei, k):
        ref = next(t for t in (rgb_pred, depth_pred, normal_pred, grad_theta) if t is not None)
        dev = ref.device

        def cz(t):
            return None if t is None else _c(t.detach().float())
        sdf = cz(k["sdf"])
        R, S = sdf.shape
        a = LossT()
        a.R, a.S, a.B, a.N = R, S, k["B"], k["N"]
        a.G = 0 if grad_theta is None else grad_theta.shape[0]
        a.depth_mask_all = int(bool(k.get("depth_mask_all", False)))
        keep = {"sdf": sdf}
        for name, t in (("mask_gt", k.get("mask_gt")), ("rgb_pred", rgb_pred), ("rgb_gt", k.get("rgb_gt")),
                        ("depth_pred", depth_pred), ("depth_gt", k.get("depth_gt")), ("gt_depth", k.get("gt_depth")),
                        ("gt_depth_valid", k.get("gt_depth_valid")), ("normal_pred", normal_pred),
                        ("normal_gt", k.get("normal_gt")), ("grad_theta", grad_theta), ("grad_theta_nei", grad_theta_nei)):
            keep[name] = cz(t)
        for name, t in keep.items():
            setattr(a, name, t.data_ptr() if t is not None else None)
            if t is not None:
                _lib.require(t, torch.float32, name)
        for w in ("w_rgb", "w_depth", "w_gt_depth", "w_normal_l1", "w_normal_cos", "w_eik", "w_smooth"):
            setattr(a, w, float(k.get(w, 0.0)))
        grads = {}
        for name, t, need in (("g_rgb", rgb_pred, 0), ("g_depth", depth_pred, 1), ("g_normal", normal_pred, 2),
                              ("g_theta", grad_theta, 3), ("g_theta_nei", grad_theta_nei, 4)):
            g = torch.empty(t.shape, device=dev) if (t is not None and ctx.needs_input_grad[need]) else None
            grads[name] = g
            setattr(a, name, g.data_ptr() if g is not None else None)
        acc = torch.empty(8 + 5 * a.B, dtype=torch.float64, device=dev)
        maskf = torch.empty(R, device=dev)
        terms = torch.empty(LOSS_TERMS, device=dev)
        check(lib().nicer_slam_loss(C.byref(a), C.c_void_p(acc.data_ptr()), ptr(maskf), ptr(terms), stream()), "nicer_slam_loss")
        _lib.launch_count += 2
        ctx.grads = grads
        out_terms = terms.detach()
        ctx.mark_non_differentiable(out_terms)
        return terms[LOSS_SUM].clone(), out_terms

    @staticmethod
    def backward(ctx, g_sum, _g_terms):
        g = ctx.grads
        return tuple((g[n] * g_sum if g[n] is not None else None)
                     for n in ("g_rgb", "g_depth", "g_normal", "g_theta", "g_theta_nei")) + (None,)


# --------------------------------------------------------------------------------------------- warp sampling
class WarpSampleFn(torch.autograd.Function):
    """(depth [B,N], dirs_p [B,N*pp,3], loc_p [B,3], w2c [B,4,4], K [B,4,4], full_rgb [B,H,W,3]) ->
    (sampled [B,B,N,pp,3], in-image mask [B,B,N,pp] bool)   -- the projection + grid_sample core of the warp block
    (network.py:167-279); differentiable w.r.t. depth, the patch rays and the world-to-camera matrices."""

    @staticmethod
    def forward(ctx, depth, dirs_p, loc_p, w2c, K, full_rgb, pp):
        depth, dirs_p, loc_p = _c(depth.detach().float()), _c(dirs_p.detach()), _c(loc_p.detach())
        w2c, K, img = _c(w2c.detach()), _c(K.detach()), _c(full_rgb.detach())
        B, N = depth.shape[0], depth.shape[1]
        H, W = img.shape[1], img.shape[2]
        E = B * N * pp
        sampled = torch.empty(B, E, 3, device=depth.device)
        mask = torch.empty(B, E, dtype=torch.uint8, device=depth.device)
        check(lib().nicer_warp_sample(ptr(depth), ptr(dirs_p), ptr(loc_p), ptr(w2c), ptr(K), ptr(img), B, N, pp, H, W,
                                      ptr(sampled), C.c_void_p(mask.data_ptr()), stream()), "nicer_warp_sample")
        ctx.save_for_backward(depth, dirs_p, loc_p, w2c, K, img)
        ctx.dims = (B, N, pp, H, W)
        m = mask.bool().reshape(B, B, N, pp)
        ctx.mark_non_differentiable(m)
        return sampled.reshape(B, B, N, pp, 3), m

    @staticmethod
    def backward(ctx, g_sampled, _g_mask):
        depth, dirs_p, loc_p, w2c, K, img = ctx.saved_tensors
        B, N, pp, H, W = ctx.dims
        dev = depth.device
        g_depth, g_loc, g_w2c = _zeros_like_many([depth, loc_p, w2c])
        g_dirs = torch.empty_like(dirs_p)
        g_sampled = _c(g_sampled)
        check(lib().nicer_warp_sample_backward(ptr(depth), ptr(dirs_p), ptr(loc_p), ptr(w2c), ptr(K), ptr(img), B, N, pp, H, W,
                                               ptr(g_sampled), ptr(g_depth), ptr(g_dirs), ptr(g_loc), ptr(g_w2c), stream()),
              "nicer_warp_sample_backward")
        return g_depth, g_dirs, g_loc, g_w2c, None, None, None


def warp_gt(uv_patch, full_rgb, full_depth):
    """Ground truth of the warp block (network.py:226-246): uv_patch [B,M,2], full_rgb [B,H,W,3], full_depth [B,H,W,1] ->
    gt_rgb [B,M,3], gt_depth [B,M,1] (1 outside the image), inside [B,M] bool.  No gradient (ground truth)."""
    uvp, img, dep = _c(uv_patch.detach().float()), _c(full_rgb.detach().float()), _c(full_depth.detach().float())
    B, M = uvp.shape[0], uvp.shape[1]
    H, W = img.shape[1], img.shape[2]
    gt_rgb = torch.empty(B, M, 3, device=uvp.device)
    gt_depth = torch.empty(B, M, 1, device=uvp.device)
    inside = torch.empty(B, M, dtype=torch.bool, device=uvp.device)
    check(lib().nicer_warp_gt(ptr(uvp), ptr(img), ptr(dep), B, M, H, W, ptr(gt_rgb), ptr(gt_depth),
                              C.c_void_p(inside.data_ptr()), stream()), "nicer_warp_gt")
    return gt_rgb, gt_depth, inside


class MaskedL1MeanFn(torch.autograd.Function):
    """mean |a - b| over the selected entries (torch.abs(a[mask] - b[mask]).mean(), loss.py:93-104,145-152): a [..., inner],
    mask one flag per `inner` values, b either shaped like a or a trailing block of it (leading-dimension broadcast).
    One kernel per direction; entries that are masked out never enter the sum (NaN / Inf there are dropped, as indexing does)."""

    @staticmethod
    def forward(ctx, a, b, mask, inner):
        a_, b_ = _c(a.detach().float()), _c(b.detach().float())
        m = _c(mask.detach())
        if m.dtype != torch.bool:
            m = m != 0
        n_mask = m.numel()
        if a_.numel() != n_mask * inner or a_.numel() % b_.numel() != 0:
            raise ValueError(f"masked_l1_mean: a {tuple(a.shape)}, b {tuple(b.shape)}, mask {tuple(mask.shape)}, inner {inner}")
        out = torch.empty(2, device=a_.device)
        ws = torch.empty((lib().nicer_masked_l1_mean_workspace() + 7) // 8, dtype=torch.float64, device=a_.device)
        check(lib().nicer_masked_l1_mean(ptr(a_), ptr(b_), C.c_void_p(m.data_ptr()), n_mask, inner, b_.numel(),
                                         C.c_void_p(ws.data_ptr()), ptr(out), stream()), "nicer_masked_l1_mean")
        ctx.save_for_backward(a_, b_, m, out)
        ctx.inner, ctx.shape = inner, a.shape
        return out[0]

    @staticmethod
    def backward(ctx, g):
        a_, b_, m, out = ctx.saved_tensors
        ga = torch.empty_like(a_)
        g = _c(g.reshape(1).float())
        check(lib().nicer_masked_l1_mean_backward(ptr(a_), ptr(b_), C.c_void_p(m.data_ptr()), m.numel(), ctx.inner, b_.numel(),
                                                  ptr(out), ptr(g), ptr(ga), stream()), "nicer_masked_l1_mean_backward")
        return ga.reshape(ctx.shape), None, None, None


def masked_l1_mean(a, b, mask):
    """a [..., c] or [...]; mask [...] (same leading shape, one flag per trailing c values) or shaped like a; b like a, possibly
    expanded over leading dimensions (stride 0)."""
    inner = a.numel() // mask.numel()
    # an expanded target (the warp ground truth is the same for every target frame) is passed as its un-expanded block
    lead = 0
    while lead < b.dim() and b.stride(lead) == 0 and b.shape[lead] > 1:
        lead += 1
    if lead:
        b = b[(0,) * lead]
    return MaskedL1MeanFn.apply(a, b, mask, inner)


# --------------------------------------------------------------------------------------------- flow projection
class FlowProjectFn(torch.autograd.Function):
    """(depth [B*n], dirs [B*n,3], loc [B,3], w2c [E,4,4], K [E,4,4], uv [B,n,2], idii [E]) -> flow [E,n,2]
    (network.py:153-165); differentiable w.r.t. depth, dirs, loc and w2c."""

    @staticmethod
    def forward(ctx, depth, dirs, loc, w2c, K, uv, idii):
        depth, dirs, loc = _c(depth.detach().reshape(-1).float()), _c(dirs.detach().reshape(-1, 3)), _c(loc.detach())
        w2c, K, uv, idii = _c(w2c.detach()), _c(K.detach()), _c(uv.detach()), _c(idii.to(torch.int64))
        E, n = idii.shape[0], uv.shape[1]
        flow = torch.empty(E, n, 2, device=depth.device)
        check(lib().nicer_flow_project(ptr(depth), ptr(dirs), ptr(loc), ptr(w2c), ptr(K), ptr(uv), ptr(idii, torch.int64, "idii"), E, n,
                                       ptr(flow), stream()), "nicer_flow_project")
        ctx.save_for_backward(depth, dirs, loc, w2c, K, idii)
        ctx.dims = (E, n)
        return flow

    @staticmethod
    def backward(ctx, g_flow):
        depth, dirs, loc, w2c, K, idii = ctx.saved_tensors
        E, n = ctx.dims
        g_flow = _c(g_flow)
        g_depth, g_dirs, g_loc, g_w2c = _zeros_like_many([depth, dirs, loc, w2c])
        check(lib().nicer_flow_project_backward(ptr(depth), ptr(dirs), ptr(loc), ptr(w2c), ptr(K), ptr(idii, torch.int64, "idii"), E, n,
                                                ptr(g_flow), ptr(g_depth), ptr(g_dirs), ptr(g_loc), ptr(g_w2c), stream()),
              "nicer_flow_project_backward")
        return g_depth, g_dirs, g_loc, g_w2c, None, None, None
