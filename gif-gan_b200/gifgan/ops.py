"""Operator layer -- the drop-in for /root/reference/models/recurrent_z/ops.py.

Same function names, argument order and defaults as the reference (`conv2d`,
`deconv2d`, `conv3d`, `linear`, `batch_norm`, `lrelu`, `add_noise`, `get_std`,
`conv_cond_concat`), same variable names under the same scopes,
NHWC activations -- but every op is a `torch.autograd.Function` whose forward and
backward call the hand-written sm_100a kernels of libgifgan.so through the C ABI
(`_cabi.py`).  TensorFlow's graph/variable machinery is replaced by:

  * `VariableStore` + `variable_scope` / `get_variable`: named fp32 master variables,
    re-packed into flat parameter/gradient/Adam buffers per optimiser group so that
    one fused Adam launch and one gradient all-reduce bucket cover a whole var_list;
  * "symbolic" graph construction = calling the ops on `device='meta'` tensors: shapes
    propagate, variables get created, nothing is computed (what `build_model` does);
  * filter gradients are written by the wgrad kernels straight into the flat gradient
    buffer (accumulating, like TF's gradient aggregation when a variable is used
    twice); autograd only routes activation gradients.

Extensions over the reference signatures (keyword-only, default = reference
behaviour): `act=` fuses the activation that follows into the producing kernel's
epilogue; `groups=` on batch_norm normalises row blocks independently.

Precision: `set_precision('fp32')` -- everything fp32, SIMT FFMA kernels (1e-4 parity
mode); `set_precision('bf16')` -- activations bf16, fp32 master weights / statistics /
accumulation, tcgen05 tensor-core kernels where the channel counts allow.
"""
from __future__ import annotations

import contextlib
import os
import ctypes
from collections import OrderedDict

import numpy as np
import torch

from . import _cabi as cabi
from ._cabi import ACT, ConvDesc, check, dt, ptr, stream

# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------
_PRECISION = "fp32"
TC_DEFAULT = True       # tcgen05 kernels validated on B200 (tools/tc_probe.py, tests/test_gpu_tc.py)
_USE_TC = TC_DEFAULT


def set_precision(p: str, tensor_cores=None):
    """'fp32' | 'bf16'.  tensor_cores: use the tcgen05 kernels in bf16 mode (None -> TC_DEFAULT)."""
    global _PRECISION, _USE_TC
    if p not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    _PRECISION, _USE_TC = p, (TC_DEFAULT if tensor_cores is None else bool(tensor_cores))
    if _PRECISION == "bf16" and _USE_TC:
        ensure_workspace()


_WORKSPACE = None


def ensure_workspace():
    """Hand the library its split-K workspace (include/gifgan.h: gg_set_workspace) once per process: a zero-filled device buffer
    that lives as long as the process, allocated here -- outside any CUDA-graph capture -- because the library never allocates
    device memory.  GG_TC_WORKSPACE_MB overrides the size (0 = none: the small-M layers then run unsplit)."""
    global _WORKSPACE
    if _WORKSPACE is not None or not torch.cuda.is_available():
        return
    L = cabi.lib()
    mb = os.environ.get("GG_TC_WORKSPACE_MB")
    nbytes = int(L.gg_workspace_bytes()) if mb is None else int(float(mb) * (1 << 20))
    if nbytes <= 0:
        _WORKSPACE = False
        return
    _WORKSPACE = torch.zeros(nbytes + 1024, dtype=torch.uint8, device="cuda")
    base = (_WORKSPACE.data_ptr() + 1023) // 1024 * 1024
    check(L.gg_set_workspace(base, nbytes), "gg_set_workspace")


def get_precision() -> str:
    return _PRECISION


def act_dtype():
    return torch.float32 if _PRECISION == "fp32" else torch.bfloat16


def same_pad(n: int, k: int, s: int):
    """TF 'SAME' (SURVEY App. A.1): out = ceil(n/s), pad_lo = total//2."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2, total - total // 2


# --------------------------------------------------------------------------------------
# variables
# --------------------------------------------------------------------------------------
class Var:
    """A named fp32 variable (tf.get_variable).  `.data` / `.grad` may be views into the
    store's flat buffers after `VariableStore.finalize`."""

    def __init__(self, name, data, trainable=True, filter_taps=None):
        self.name, self.data, self.trainable = name, data, trainable
        self.grad = None
        self.filter_taps = filter_taps      # not None for conv/deconv/conv3d filters [taps..., C, K]
        self.version = 0
        self._bf16 = None          # bf16 copy (a view into the store's shadow buffer after finalize)
        self._bf16_version = -1

    @property
    def shape(self):
        return tuple(self.data.shape)

    def numel(self):
        return self.data.numel()

    def bf16(self):
        """bf16 copy of the variable, same layout -- the filter operand of the tensor-core kernels (conv_up reads it K-major,
        conv_down MN-major, the class-concatenated conv_up one box per class: no transposed / re-packed copies exist).
        After VariableStore.finalize it is a view into the store's bf16 shadow of the flat parameter buffer, which the
        fused Adam launch rewrites together with the fp32 masters (AdamOptimizer.apply), so a train step carries no
        re-pack launch; anything else that changes the variable bumps `version` and the copy is re-cast at its next use."""
        if self._bf16 is None:
            self._bf16 = torch.empty(self.data.shape, dtype=torch.bfloat16, device=self.data.device)
            self._bf16_version = -1
        if self._bf16_version != self.version:
            check(cabi.lib().gg_cast(ptr(self.data), cabi.GG_F32, ptr(self._bf16), cabi.GG_BF16, self.data.numel(), stream()), "gg_cast")
            self._bf16_version = self.version
        return self._bf16

    def packs_stale(self):
        return self._bf16 is not None and self._bf16_version != self.version

    def refresh_packs(self):
        """Re-cast the bf16 copy if it exists and is stale."""
        if self._bf16 is not None:
            self.bf16()

    def invalidate_packed(self):
        """Force the bf16 copy to be rebuilt at its next use."""
        self._bf16_version = -1


class VariableStore:
    def __init__(self, device=None, seed=0):
        if device is None:
            device = "cuda" if torch.cuda.is_available() else "cpu"   # cpu: inventory/checkpoint logic only
        self.device = torch.device(device)
        self.vars: "OrderedDict[str, Var]" = OrderedDict()
        self.rng = np.random.RandomState(seed)
        self._scope: list[str] = []
        self.flat = None           # dict(params=, grads=, m=, v=) after finalize
        self.ranges = {}           # group name -> (begin, end) element range in the flat buffers

    # -- scopes ---------------------------------------------------------------------
    def scope_name(self):
        return "".join(s + "/" for s in self._scope)

    @contextlib.contextmanager
    def absolute_scope(self, prefix: str):
        """Enter the scope `prefix` ("a/b/") regardless of the caller's current scope: a model remembers the scope
        it was built under, so its methods find their variables wherever they are called from."""
        saved = self._scope
        self._scope = [p for p in prefix.split("/") if p]
        _ACTIVE.append(self)              # variable_scope / get_variable inside resolve to THIS store
        try:
            yield self
        finally:
            _ACTIVE.pop()
            self._scope = saved

    # -- creation -------------------------------------------------------------------
    def get_variable(self, name, shape, initializer, trainable=True, filter_taps=None) -> Var:
        full = self.scope_name() + name
        v = self.vars.get(full)
        if v is not None:
            if tuple(v.shape) != tuple(int(s) for s in shape):
                raise ValueError(f"variable {full} exists with shape {v.shape}, requested {tuple(shape)}")
            return v
        if self.flat is not None:
            raise RuntimeError(f"variable {full} created after VariableStore.finalize()")
        arr = initializer(self.rng, tuple(int(s) for s in shape))
        data = torch.as_tensor(np.asarray(arr, dtype=np.float32)).to(self.device).contiguous()
        v = Var(full, data, trainable, filter_taps)
        self.vars[full] = v
        return v

    # -- flat buffers -----------------------------------------------------------------
    def finalize(self, groups: "OrderedDict[str, list[Var]]"):
        """Re-pack variables into one flat fp32 buffer; each group (an optimiser's
        var_list) occupies a contiguous, 16-byte aligned range so that Adam and the
        gradient all-reduce run once per group.  Adam slots m, v mirror it."""
        ordered, seen = [], set()
        self.ranges = {}
        off = 0
        for gname, vs in groups.items():
            beg = off
            for v in vs:
                if v.name in seen:
                    raise ValueError(f"{v.name} is in two optimiser groups")
                seen.add(v.name)
                ordered.append((v, off))
                off += (v.numel() + 63) // 64 * 64       # 64-element granules: every variable starts on a 128-byte line in the bf16
                                                         # shadow (TMA box rows of the filter operand = whole lines) and on 256 bytes
                                                         # in the fp32 buffers (the filter-gradient TMA reduce-adds)
            self.ranges[gname] = (beg, off)
        for v in self.vars.values():
            if v.name not in seen:
                ordered.append((v, off))
                off += (v.numel() + 63) // 64 * 64
        total = off
        dev = self.device
        params = torch.zeros(total, dtype=torch.float32, device=dev)
        shadow = torch.zeros(total, dtype=torch.bfloat16, device=dev)    # bf16 copy of `params`, kept current by Adam (Var.bf16)
        grads = torch.zeros(total, dtype=torch.float32, device=dev)
        trainable_end = max([e for _, e in self.ranges.values()], default=0)
        m = torch.zeros(trainable_end, dtype=torch.float32, device=dev)
        v2 = torch.zeros(trainable_end, dtype=torch.float32, device=dev)
        for v, o in ordered:
            n = v.numel()
            params[o:o + n].copy_(v.data.reshape(-1))
            shape = v.data.shape
            v.data = params[o:o + n].view(shape)
            v.grad = grads[o:o + n].view(shape)
            v.offset = o
            v.version += 1
            v._bf16, v._bf16_version = shadow[o:o + n].view(shape), -1
        self.flat = dict(params=params, grads=grads, m=m, v=v2)
        self.shadow = shadow

    # -- checkpoint (keys = TF variable names) -------------------------
    def state_dict(self):
        return OrderedDict((k, v.data.detach().clone().cpu()) for k, v in self.vars.items())

    def load_state_dict(self, sd, strict=True, prefix=""):
        missing = []
        with torch.no_grad():
            for k, v in self.vars.items():
                if not k.startswith(prefix):
                    continue
                key = k[len(prefix):]
                if key in sd:
                    v.data.copy_(torch.as_tensor(np.asarray(sd[key]), dtype=torch.float32).reshape(v.data.shape))
                    v.version += 1
                elif strict:
                    missing.append(key)
        if missing:
            raise KeyError(f"missing variables in checkpoint: {missing[:5]}...")


_STORE = VariableStore.__new__(VariableStore)  # replaced by reset_default_store()
_STORE_READY = False
_ACTIVE: list = []        # stack of stores entered with use_store() / VariableStore.absolute_scope(); empty -> the default store


def reset_default_store(device=None, seed=0) -> VariableStore:
    """Start a fresh 'graph' (tf.reset_default_graph)."""
    global _STORE, _STORE_READY
    _STORE = VariableStore(device, seed)
    _STORE_READY = True
    return _STORE


def default_store() -> VariableStore:
    if not _STORE_READY:
        reset_default_store()
    return _STORE


def current_store() -> VariableStore:
    """The store that variable_scope / the op constructors create variables in: the innermost use_store() /
    absolute_scope() region, else the default store."""
    return _ACTIVE[-1] if _ACTIVE else default_store()


@contextlib.contextmanager
def use_store(store: VariableStore):
    """Make `store` the one the operator layer creates its variables in (a model built with `store=`)."""
    _ACTIVE.append(store)
    try:
        yield store
    finally:
        _ACTIVE.pop()


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    """tf.variable_scope(name).  Reuse is automatic: get_variable returns the existing
    variable of that name (the reference always calls reuse_variables() before a second use)."""
    st = current_store()
    st._scope.append(name)
    try:
        yield st
    finally:
        st._scope.pop()


def truncated_normal_initializer(stddev=0.02):
    def init(rs, shape):
        out = rs.normal(0.0, 1.0, size=shape)
        bad = np.abs(out) > 2.0
        while bad.any():
            out[bad] = rs.normal(0.0, 1.0, size=int(bad.sum()))
            bad = np.abs(out) > 2.0
        return out * stddev
    return init


def random_normal_initializer(stddev=0.02):
    return lambda rs, shape: rs.normal(0.0, stddev, size=shape)


def constant_initializer(value=0.0):
    return lambda rs, shape: np.full(shape, value, dtype=np.float32)


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def _is_meta(t):
    return t.device.type == "meta"


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: tensor on {t.device}; gif-gan_b200 has no CPU fallback (use CUDA, or 'meta' to build)")


def _wtensor(var: Var, requires_grad: bool):
    """The autograd handle of a variable for ONE op call: a fresh detached alias of its data with requires_grad set by the
    current var_list (see `trainable`).  A fresh leaf per call, not the stored tensor itself: a leaf's AccumulateGrad node is
    cached on the tensor together with the stream it was created on, and it stays alive as long as ANY graph that used the
    tensor does -- e.g. the loss tensor of an earlier eager update that the caller still holds.  Re-using such a leaf while a
    CUDA graph is being captured made the autograd engine synchronise the capture stream with that (uncaptured) stream:
    cudaErrorStreamCaptureIsolation.  The ops write filter gradients themselves (Var.grad) and return None to autograd."""
    t = var.data.detach()
    if requires_grad:
        t.requires_grad_(True)
    return t


_TRAINABLE: set = set()
DEBUG_TAP = None     # tools/diag_parity.py sets this to a dict to capture per-layer tensors


@contextlib.contextmanager
def trainable(var_list):
    """Variables that receive gradients in this region -- the `var_list` of
    tf.train.AdamOptimizer(...).minimize(loss, var_list=...) (model.py:153-156)."""
    global _TRAINABLE
    old = _TRAINABLE
    _TRAINABLE = {v.name for v in var_list}
    try:
        yield
    finally:
        _TRAINABLE = old


# ---- zeroed scratch for batch-norm reductions -----------------------------------------------------
# Every batch-norm layer needs a zeroed fp64 [groups, 2, C] accumulator in forward (fused statistics) and in backward.
# Inside `with stats_arena():` (the models' update functions) they are bump-allocated from ONE buffer that is zeroed by a
# single memset when the region is entered -- 3 memsets per train step instead of ~32 fill / memset nodes in the graph.
_ARENA = None


class _StatsArena:
    def __init__(self, device, nbytes=1 << 20):
        self.buf = torch.zeros(nbytes // 8, dtype=torch.float64, device=device)
        self.off = 0

    def take(self, n):
        n8 = (n + 1) // 2 * 2                       # 16-byte granules
        if self.off + n8 > self.buf.numel():
            return None
        t = self.buf[self.off:self.off + n]
        self.off += n8
        return t


_ARENAS = {}


@contextlib.contextmanager
def stats_arena():
    global _ARENA
    old = _ARENA
    if torch.cuda.is_available():
        dev = torch.cuda.current_device()
        a = _ARENAS.get(dev)
        if a is None:
            a = _ARENAS[dev] = _StatsArena(torch.device("cuda", dev))
        a.buf.zero_()
        a.off = 0
        _ARENA = a
    try:
        yield
    finally:
        _ARENA = old


def _zeroed_f64(n, device):
    """(tensor, prezeroed_by_arena)"""
    if _ARENA is not None:
        t = _ARENA.take(n)
        if t is not None:
            return t, True
    return torch.zeros(n, dtype=torch.float64, device=device), False


def _wants_grad(var: Var) -> bool:
    return var.trainable and var.name in _TRAINABLE and torch.is_grad_enabled()


def _tc_ok(C_, K_, *tensors) -> bool:
    return (_PRECISION == "bf16" and _USE_TC and C_ % 64 == 0 and K_ % 64 == 0
            and all(t.dtype == torch.bfloat16 for t in tensors))


class _Geom:
    """Geometry of one strided-conv relation (gg_conv_desc without dtypes)."""
    __slots__ = ("N", "D", "H", "W", "C", "Do", "Ho", "Wo", "K", "k", "s", "p")

    def __init__(self, N, large_sp, C_, small_sp, K_, k, s, p):
        self.N, (self.D, self.H, self.W), self.C = N, large_sp, C_
        (self.Do, self.Ho, self.Wo), self.K = small_sp, K_
        self.k, self.s, self.p = k, s, p

    def desc(self, large_dt, small_dt, act=None, act_param=0.2, tc=False) -> ConvDesc:
        return ConvDesc(self.N, self.D, self.H, self.W, self.C, self.Do, self.Ho, self.Wo, self.K,
                        self.k[0], self.k[1], self.k[2], self.s[0], self.s[1], self.s[2],
                        self.p[0], self.p[1], self.p[2], large_dt, small_dt, ACT[act], float(act_param),
                        cabi.CONV_TENSOR_CORE if tc else 0)

    def large_shape(self, ndim):
        return (self.N, self.H, self.W, self.C) if ndim == 4 else (self.N, self.D, self.H, self.W, self.C)

    def small_shape(self, ndim):
        return (self.N, self.Ho, self.Wo, self.K) if ndim == 4 else (self.N, self.Do, self.Ho, self.Wo, self.K)


class _BnInfo:
    """What a train-mode fused batch norm leaves on its output tensor (`y._gg_bn`) so that the conv that consumes y can, in
    ITS backward pass, produce the batch norm's backward reductions together with dy (gg_conv_dgrad_bnbwd)."""
    __slots__ = ("pre", "mean", "rstd", "gamma", "beta", "act", "act_param", "groups", "Cc", "bwd_sums", "bwd_dy", "bwd_version")

    def __init__(self):
        self.pre = self.bwd_sums = self.bwd_dy = None
        self.bwd_version = -1


FUSE_LINEAR_BN = os.environ.get("GG_FUSE_LINEAR_BN", "0") == "1"    # opt-in: thin linear + batch norm + activation as one kernel per direction (csrc/linbn.cu: measured break-even)
FUSE_BN_BWD = os.environ.get("GG_FUSE_BN_BWD", "1") != "0"    # A/B switch: batch-norm backward reductions in the dgrad epilogue


def _run_dgrad_bnbwd(g: _Geom, up, src, wvar, out, bnb: _BnInfo, tc=True):
    """dgrad launch that also accumulates the consumer batch norm's (sum g, sum g*xhat); returns True when it was fused."""
    L = cabi.lib()
    large, small = (out, src) if up else (src, out)
    d = g.desc(dt(large), dt(small), None, 0.0, tc)
    sums = _zeroed_f64(L.gg_bn_workspace_bytes(bnb.Cc, bnb.groups) // 8, out.device)[0]
    fused = ctypes.c_int32(0)
    check(L.gg_conv_dgrad_bnbwd(ctypes.byref(d), 1 if up else 0, ptr(src), ptr(wvar.bf16() if tc else wvar.data), ptr(out), ptr(bnb.pre), ptr(bnb.mean), ptr(bnb.rstd),
                                ptr(bnb.gamma), ptr(bnb.beta), ACT[bnb.act], float(bnb.act_param), bnb.groups, ptr(sums), ctypes.byref(fused),
                                stream()), "gg_conv_dgrad_bnbwd")
    if fused.value:
        bnb.bwd_sums, bnb.bwd_dy, bnb.bwd_version = sums, out, out._version
    return bool(fused.value)


def _bnb_usable(bnb, out_shape, out_dtype):
    # same memory layout is what matters: [B, s*s*C] of a linear feeding a batch norm over C channels IS [B, s, s, C]
    n = 1
    for e in out_shape:
        n *= int(e)
    return (FUSE_BN_BWD and bnb is not None and bnb.pre is not None and bnb.pre.numel() == n
            and bnb.pre.dtype == torch.float32 and out_shape[-1] == bnb.Cc)


def reshape(x, shape):
    """x.reshape(shape) that keeps the batch-norm hand-over of a fused node's output (`_gg_bn`), so that the conv consuming the
    reshaped tensor still produces that batch norm's backward reductions in its dgrad launch (model.py:305: the generator's
    [B, 8192] projection viewed as [B, 4, 4, 512])."""
    y = x.reshape(shape)
    info = getattr(x, "_gg_bn", None)
    if info is not None and y.shape[-1] % info.Cc == 0:      # channel = last index % Cc in both views (same memory)
        y._gg_bn = info
    return y


def _run_down(g: _Geom, large, wvar: Var, bias, out_dtype, act, act_param, ndim, out=None, stats=None, groups=1, bnb=None):
    """stats: optional fp64 [groups, 2, K] accumulator for the batch-norm statistics of the output.
    bnb: the _BnInfo of the batch norm whose output gradient this launch produces (dgrad of a deconv)."""
    small = out if out is not None else torch.empty(g.small_shape(ndim), dtype=out_dtype, device=large.device)
    tc = _tc_ok(g.C, g.K, large)
    # image-side layer in bf16 mode (g_h4's dgrad: fp32 image gradient in, bf16 dy of g_bn3 out): the warp-MMA kernel fuses too
    c3 = (not tc) and g.C == 3 and _PRECISION == "bf16" and small.dtype == torch.bfloat16 and large.dtype == torch.float32 and bnb is not None and bnb.groups == 1
    if (tc or c3) and bias is None and act is None and stats is None and _bnb_usable(bnb, small.shape, small.dtype):
        _run_dgrad_bnbwd(g, False, large, wvar, small, bnb, tc=tc)
        return small
    d = g.desc(dt(large), dt(small), act, act_param, tc)
    w = wvar.bf16() if tc else wvar.data
    if stats is not None:
        check(cabi.lib().gg_conv_down_stats(ctypes.byref(d), ptr(large), ptr(w), ptr(bias), ptr(small), ptr(stats), groups, stream()),
              "gg_conv_down_stats")
    else:
        check(cabi.lib().gg_conv_down(ctypes.byref(d), ptr(large), ptr(w), ptr(bias), ptr(small), stream()), "gg_conv_down")
    return small


def _run_up(g: _Geom, small, wvar: Var, bias, out_dtype, act, act_param, ndim, out=None, stats=None, groups=1, bnb=None):
    large = out if out is not None else torch.empty(g.large_shape(ndim), dtype=out_dtype, device=small.device)
    tc = _tc_ok(g.C, g.K, small)
    if tc and bias is None and act is None and stats is None and _bnb_usable(bnb, large.shape, large.dtype):
        _run_dgrad_bnbwd(g, True, small, wvar, large, bnb)
        return large
    d = g.desc(dt(large), dt(small), act, act_param, tc)
    w = wvar.bf16() if tc else wvar.data
    if stats is not None:
        check(cabi.lib().gg_conv_up_stats(ctypes.byref(d), ptr(small), ptr(w), ptr(bias), ptr(large), ptr(stats), groups, stream()),
              "gg_conv_up_stats")
    else:
        check(cabi.lib().gg_conv_up(ctypes.byref(d), ptr(small), ptr(w), ptr(bias), ptr(large), stream()), "gg_conv_up")
    return large


# Filter gradients are leaves of the backward pass: nothing downstream of them runs before the optimiser.  They are
# issued on a side stream so that they overlap with the activation-gradient chain (under CUDA-graph capture the
# fork/join becomes parallel branches of the graph).  join_side() is the join point (optimiser / all-reduce).
def refresh_packs(store):
    """Bring the bf16 filter copies up to date.  A captured train step assumes current copies on entry -- Adam rewrites the
    bf16 shadow of its group inside the step -- so this runs after anything that changes weights behind the optimisers'
    back: the restore around graph capture, a checkpoint load.  With a finalized store ONE cast over the flat buffer
    refreshes every variable; otherwise one cast per stale variable."""
    stale = [v for v in store.vars.values() if v.packs_stale()]
    if not stale:
        return 0
    sh = getattr(store, "shadow", None)
    if sh is not None and store.flat is not None:
        check(cabi.lib().gg_cast(ptr(store.flat["params"]), cabi.GG_F32, ptr(sh), cabi.GG_BF16, sh.numel(), stream()), "gg_cast")
        for v in store.vars.values():
            if v._bf16 is not None:
                v._bf16_version = v.version
    else:
        for v in stale:
            v.refresh_packs()
    return len(stale)


OVERLAP_WGRAD = False      # enabled inside `with overlap_wgrad():` (the model's update functions)
GRAD_READY_HOOK = None     # data parallel: called as hook(var, producer_stream) after a filter gradient was enqueued (dp.py)
_SIDE = {}


@contextlib.contextmanager
def overlap_wgrad(enable=True):
    """Region in which filter-gradient kernels go to the side stream; joined on exit."""
    global OVERLAP_WGRAD
    old = OVERLAP_WGRAD
    OVERLAP_WGRAD = bool(enable) and torch.cuda.is_available()
    try:
        yield
    finally:
        OVERLAP_WGRAD = old
        join_side()


def _side_stream():
    dev = torch.cuda.current_device()
    s = _SIDE.get(dev)
    if s is None:
        s = _SIDE[dev] = dict(stream=torch.cuda.Stream(), dirty=False)
    return s


class _on_side:
    """with _on_side(tensors...): work enqueued inside runs on the side stream, ordered after everything already
    enqueued on the current stream."""

    def __init__(self, *tensors):
        self.tensors = tensors

    def __enter__(self):
        if not OVERLAP_WGRAD:
            return self
        s = _side_stream()
        s["stream"].wait_stream(torch.cuda.current_stream())
        s["dirty"] = True
        for t in self.tensors:
            if t is not None:
                t.record_stream(s["stream"])
        self.ctx = torch.cuda.stream(s["stream"])
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if OVERLAP_WGRAD:
            self.ctx.__exit__(*exc)
        return False


def join_side():
    """Make the current stream wait for all side-stream work (call before consuming filter gradients)."""
    s = _SIDE.get(torch.cuda.current_device()) if torch.cuda.is_available() else None
    if s is not None and s["dirty"]:
        torch.cuda.current_stream().wait_stream(s["stream"])
        s["dirty"] = False


def _run_wgrad(g: _Geom, large, small, wvar: Var, bvar: Var = None):
    """bvar: also produce the bias gradient of a `down` conv (sum of `small` over its grid) -- one launch on the image-side layer."""
    tc = _tc_ok(g.C, g.K, large, small)
    d = g.desc(dt(large), dt(small), None, 0.0, tc)
    with _on_side(large, small):
        if bvar is not None:
            check(cabi.lib().gg_conv_wgrad_bias(ctypes.byref(d), ptr(large), ptr(small), ptr(wvar.grad), ptr(bvar.grad), stream()),
                  "gg_conv_wgrad_bias")
        else:
            check(cabi.lib().gg_conv_wgrad(ctypes.byref(d), ptr(large), ptr(small), ptr(wvar.grad), stream()), "gg_conv_wgrad")
    if GRAD_READY_HOOK is not None:
        GRAD_READY_HOOK(wvar, _side_stream()["stream"] if OVERLAP_WGRAD else None)


def _act_bwd(y, dy, act, act_param):
    dpre = torch.empty_like(y)
    check(cabi.lib().gg_act_bwd(ptr(y), dt(y), ptr(dy), dt(dy), ptr(dpre), dt(dpre), y.numel(), ACT[act], float(act_param), stream()),
          "gg_act_bwd")
    return dpre


def _bias_grad(dpre, bvar: Var):
    Cc = dpre.shape[-1]
    check(cabi.lib().gg_bias_grad(ptr(dpre), dt(dpre), ptr(bvar.grad), dpre.numel() // Cc, Cc, stream()), "gg_bias_grad")


class _ConvRel(torch.autograd.Function):
    """conv2d / conv3d (direction 'down') and deconv2d (direction 'up') with fused bias and
    optional fused activation.  Backward: activation', bias gradient, filter gradient into the
    flat gradient buffer, input gradient through the opposite-direction kernel."""

    @staticmethod
    def forward(ctx, x, w, b, wvar, bvar, geom, direction, act, act_param, out_dtype, out, in_bn=None):
        ndim = x.dim()
        ctx.in_bn = in_bn
        run = _run_down if direction == "down" else _run_up
        y = run(geom, x, wvar, b, out_dtype, act, act_param, ndim, out)
        if out is not None:
            ctx.mark_dirty(out)
        ctx.wvar, ctx.bvar, ctx.geom, ctx.direction, ctx.act, ctx.act_param, ctx.ndim = wvar, bvar, geom, direction, act, act_param, ndim
        ctx.x_dtype = x.dtype
        ctx.save_for_backward(x, y if act else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        dy = dy.contiguous()
        g = ctx.geom
        need_b = ctx.bvar is not None and ctx.needs_input_grad[2]
        if (FUSE_ACT_BIAS and ctx.act and need_b and y.shape[-1] <= 4 and y.dtype == torch.float32 and dy.dtype == torch.float32):
            # image-side deconv (g_h4 + tanh): activation' and the 3-channel bias gradient in one pass over the image gradient
            dpre = torch.empty_like(y)
            Cc = y.shape[-1]
            check(cabi.lib().gg_act_bwd_bias(ptr(y), dt(y), ptr(dy), dt(dy), ptr(dpre), dt(dpre), y.numel() // Cc, Cc, ACT[ctx.act],
                                             float(ctx.act_param), ptr(ctx.bvar.grad), stream()), "gg_act_bwd_bias")
            need_b = False
        else:
            dpre = _act_bwd(y, dy, ctx.act, ctx.act_param) if ctx.act else dy
        # image-side conv (d_h0_conv) in bf16 mode: the filter-gradient launch carries the bias gradient
        wb = (FUSE_WGRAD_BIAS and need_b and ctx.needs_input_grad[1] and ctx.direction == "down" and g.C == 3 and _PRECISION == "bf16"
              and dpre.dtype == torch.bfloat16 and x.dtype == torch.float32)
        if need_b and not wb:
            _bias_grad(dpre, ctx.bvar)
        if ctx.needs_input_grad[1]:
            if ctx.direction == "down":
                _run_wgrad(g, x, dpre, ctx.wvar, ctx.bvar if wb else None)
            else:
                _run_wgrad(g, dpre, x, ctx.wvar)
        dx = None
        if ctx.needs_input_grad[0]:
            run = _run_up if ctx.direction == "down" else _run_down
            dx = run(g, dpre, ctx.wvar, None, ctx.x_dtype, None, 0.0, ctx.ndim, bnb=ctx.in_bn)
        return dx, None, None, None, None, None, None, None, None, None, None, None


def _conv_common(input_, wvar, bvar, geom, direction, act, act_param, out_dtype, out=None):
    if _is_meta(input_):
        shp = geom.small_shape(input_.dim()) if direction == "down" else geom.large_shape(input_.dim())
        return torch.empty(shp, dtype=out_dtype, device="meta")
    _require_cuda(input_, "conv")
    x = input_.contiguous()
    w = _wtensor(wvar, _wants_grad(wvar))
    b = _wtensor(bvar, _wants_grad(bvar)) if bvar is not None else None
    if out is not None:
        want = geom.small_shape(x.dim()) if direction == "down" else geom.large_shape(x.dim())
        if tuple(out.shape) != tuple(want) or out.dtype != out_dtype or not out.is_contiguous():
            raise ValueError(f"out= must be a contiguous {out_dtype} tensor of shape {want}")
    return _ConvRel.apply(x, w, b, wvar, bvar, geom, direction, act, act_param, out_dtype, out, getattr(input_, "_gg_bn", None))


def _out_dtype(act, out_dtype):
    if out_dtype is not None:
        return out_dtype
    return act_dtype()


# --------------------------------------------------------------------------------------
# the reference operator API (ops.py)
# --------------------------------------------------------------------------------------
def conv2d(input_, output_dim, k_h=5, k_w=5, d_h=2, d_w=2, stddev=0.02, name="conv2d", *, act=None, act_param=0.2,
           out_dtype=None, bias=True, bn=None, train=True, groups=1):
    """ops.py:51-62 -- tf.nn.conv2d(input_, w, [1,d_h,d_w,1], 'SAME') + bias_add.
    Variables `name/w` [k_h,k_w,Cin,Cout] (truncated normal) and `name/biases` (zeros)."""
    B, H, W, Cin = input_.shape
    with variable_scope(name) as st:
        wvar = st.get_variable("w", [k_h, k_w, Cin, output_dim], truncated_normal_initializer(stddev), filter_taps=k_h * k_w)
        bvar = st.get_variable("biases", [output_dim], constant_initializer(0.0)) if bias else None
    Ho, ph, _ = same_pad(H, k_h, d_h)
    Wo, pw, _ = same_pad(W, k_w, d_w)
    geom = _Geom(B, (1, H, W), Cin, (1, Ho, Wo), output_dim, (1, k_h, k_w), (1, d_h, d_w), (0, ph, pw))
    if bn is not None:
        return _fused_bn(input_, _ConvProducer(geom, "down", wvar, bvar, 4), bn, train, act, act_param, _out_dtype(act, out_dtype), groups)
    return _conv_common(input_, wvar, bvar, geom, "down", act, act_param, _out_dtype(act, out_dtype))


def conv3d(input_, output_dim, k_d=3, k_h=3, k_w=3, d_d=2, d_h=2, d_w=2, stddev=0.02, name="conv3d", *, act=None,
           act_param=0.2, out_dtype=None, bn=None, train=True, groups=1):
    """ops.py:64-75 -- tf.nn.conv3d(input_, w, [1,d_d,d_h,d_w,1], 'SAME') + bias_add."""
    B, D, H, W, Cin = input_.shape
    with variable_scope(name) as st:
        wvar = st.get_variable("w", [k_d, k_h, k_w, Cin, output_dim], truncated_normal_initializer(stddev), filter_taps=k_d * k_h * k_w)
        bvar = st.get_variable("biases", [output_dim], constant_initializer(0.0))
    Do, pd, _ = same_pad(D, k_d, d_d)
    Ho, ph, _ = same_pad(H, k_h, d_h)
    Wo, pw, _ = same_pad(W, k_w, d_w)
    geom = _Geom(B, (D, H, W), Cin, (Do, Ho, Wo), output_dim, (k_d, k_h, k_w), (d_d, d_h, d_w), (pd, ph, pw))
    if bn is not None:
        return _fused_bn(input_, _ConvProducer(geom, "down", wvar, bvar, 5), bn, train, act, act_param, _out_dtype(act, out_dtype), groups)
    return _conv_common(input_, wvar, bvar, geom, "down", act, act_param, _out_dtype(act, out_dtype))


def deconv2d(input_, output_shape, k_h=5, k_w=5, d_h=2, d_w=2, stddev=0.02, name="deconv2d", with_w=False, *, act=None,
             act_param=0.2, out_dtype=None, bias=True, out=None, bn=None, train=True, groups=1):
    """ops.py:77-100 -- tf.nn.conv2d_transpose(input_, w[k_h,k_w,Cout,Cin], output_shape, [1,d_h,d_w,1]) + bias_add.
    Variables `name/w` (normal) and `name/biases` (zeros).  with_w=True also returns the Var objects."""
    B, h, w_, Cin = input_.shape
    _, Ho, Wo, Cout = [int(s) for s in output_shape]
    with variable_scope(name) as st:
        wvar = st.get_variable("w", [k_h, k_w, Cout, Cin], random_normal_initializer(stddev), filter_taps=k_h * k_w)
        bvar = st.get_variable("biases", [Cout], constant_initializer(0.0)) if bias else None
    oh, ph, _ = same_pad(Ho, k_h, d_h)
    ow, pw, _ = same_pad(Wo, k_w, d_w)
    if (oh, ow) != (h, w_):
        raise ValueError(f"deconv2d: output_shape {output_shape} inconsistent with input {tuple(input_.shape)}")
    geom = _Geom(B, (1, Ho, Wo), Cout, (1, h, w_), Cin, (1, k_h, k_w), (1, d_h, d_w), (0, ph, pw))
    if bn is not None:
        y = _fused_bn(input_, _ConvProducer(geom, "up", wvar, bvar, 4), bn, train, act, act_param, _out_dtype(act, out_dtype), groups)
    else:
        y = _conv_common(input_, wvar, bvar, geom, "up", act, act_param, _out_dtype(act, out_dtype), out)
    return (y, wvar, bvar) if with_w else y


# ---- variable-level entry points (rnn_test scripts build tf.Variables by hand and call tf.nn.* directly) ----
def conv2d_v(input_, wvar: Var, bvar=None, d_h=2, d_w=2, *, act=None, act_param=0.2, out_dtype=None, bn=None, train=True, groups=1):
    """tf.nn.conv2d(input_, w, [1,d_h,d_w,1], 'SAME') (+ bias) on an explicit filter Var [kh,kw,Cin,Cout]
    (recurrent_DCGAN.py:189,255,275)."""
    B, H, W, Cin = input_.shape
    k_h, k_w, _, Cout = wvar.shape
    Ho, ph, _ = same_pad(H, k_h, d_h)
    Wo, pw, _ = same_pad(W, k_w, d_w)
    geom = _Geom(B, (1, H, W), Cin, (1, Ho, Wo), Cout, (1, k_h, k_w), (1, d_h, d_w), (0, ph, pw))
    if bn is not None:
        return _fused_bn(input_, _ConvProducer(geom, "down", wvar, bvar, 4), bn, train, act, act_param, _out_dtype(act, out_dtype), groups)
    return _conv_common(input_, wvar, bvar, geom, "down", act, act_param, _out_dtype(act, out_dtype))


def deconv2d_v(input_, wvar: Var, output_shape, bvar=None, d_h=2, d_w=2, *, act=None, act_param=0.2, out_dtype=None, out=None):
    """tf.nn.conv2d_transpose(input_, w[kh,kw,Cout,Cin], output_shape, [1,d_h,d_w,1], 'SAME') (recurrent_DCGAN.py:223)."""
    B, h, w_, Cin = input_.shape
    _, Ho, Wo, Cout = [int(s) for s in output_shape]
    k_h, k_w = wvar.shape[0], wvar.shape[1]
    oh, ph, _ = same_pad(Ho, k_h, d_h)
    ow, pw, _ = same_pad(Wo, k_w, d_w)
    if (oh, ow) != (h, w_):
        raise ValueError(f"deconv2d: output_shape {output_shape} inconsistent with input {tuple(input_.shape)}")
    geom = _Geom(B, (1, Ho, Wo), Cout, (1, h, w_), Cin, (1, k_h, k_w), (1, d_h, d_w), (0, ph, pw))
    return _conv_common(input_, wvar, bvar, geom, "up", act, act_param, _out_dtype(act, out_dtype), out)


def linear_v(input_, mvar: Var, bvar=None, *, act=None, act_param=0.2, out_dtype=None):
    """tf.matmul(input_, W) + b on explicit Vars (recurrent_DCGAN.py:215,262,266)."""
    od = out_dtype if out_dtype is not None else (torch.float32 if mvar.shape[1] <= 4 else act_dtype())
    if _is_meta(input_):
        return torch.empty((input_.shape[0], mvar.shape[1]), dtype=od, device="meta")
    _require_cuda(input_, "linear")
    b = _wtensor(bvar, _wants_grad(bvar)) if bvar is not None else None
    return _Linear.apply(input_.contiguous(), _wtensor(mvar, _wants_grad(mvar)), b, mvar, bvar, act, act_param, od)


def bn_act(x, bn, *, act=None, act_param=0.2, out_dtype=None, train=True, groups=1):
    """batch norm + activation on an existing tensor (decoder order BN -> ReLU -> deconv, recurrent_DCGAN.py:218-223)."""
    return bn(x, train=train, act=act, act_param=act_param, out_dtype=out_dtype, groups=groups)


TC_LINEAR = os.environ.get("GG_TC_LINEAR", "1") != "0"    # A/B switch: wide linears on the tcgen05 kernels (a linear is a one-tap conv)
TC_LINEAR_MIN_ROWS = 64


def _lin_geom(rows, in_dim, out_dim) -> _Geom:
    """tf.matmul(x[rows, in], Matrix[in, out]) as the strided-conv relation with ONE tap on a 1x1x1 grid per row: the Matrix IS
    the filter w[tap = 0][C = in][K = out], forward = `down`, input gradient = `up`, Matrix gradient = `wgrad`."""
    return _Geom(rows, (1, 1, 1), in_dim, (1, 1, 1), out_dim, (1, 1, 1), (1, 1, 1), (0, 0, 0))


def _lin_tc(rows, in_dim, out_dim, *tensors) -> bool:
    """Wide linears (recurrent_z's gvideo_1/2: 512 x 512 at clips x frames rows, z_model_lib.py:160-161) go to the tcgen05 pixel GEMM;
    thin ones (in_dim 100/101, out_dim 1 / 100 / 400) have no 64-channel rows for a TMA box and stay on the SIMT kernels."""
    return TC_LINEAR and rows >= TC_LINEAR_MIN_ROWS and _tc_ok(in_dim, out_dim, *tensors)


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, wvar, bvar, act, act_param, out_dtype):
        rows, in_dim = x.shape
        out_dim = w.shape[1]
        if _lin_tc(rows, in_dim, out_dim, x):
            y = _run_down(_lin_geom(rows, in_dim, out_dim), x.view(rows, 1, 1, in_dim), wvar, b, out_dtype, act, act_param, 4).view(rows, out_dim)
        else:
            y = torch.empty((rows, out_dim), dtype=out_dtype, device=x.device)
            check(cabi.lib().gg_linear_fwd(ptr(x), dt(x), ptr(w), ptr(b), ptr(y), dt(y), rows, in_dim, out_dim, ACT[act], float(act_param),
                                           stream()), "gg_linear_fwd")
        ctx.wvar, ctx.bvar, ctx.act, ctx.act_param = wvar, bvar, act, act_param
        ctx.save_for_backward(x, y if act else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        dy = dy.contiguous()
        dpre = _act_bwd(y, dy, ctx.act, ctx.act_param) if ctx.act else dy
        rows, in_dim = x.shape
        out_dim = dpre.shape[1]
        L = cabi.lib()
        wg, bg = ctx.needs_input_grad[1], ctx.bvar is not None and ctx.needs_input_grad[2]
        g = _lin_geom(rows, in_dim, out_dim)
        if wg and _lin_tc(rows, in_dim, out_dim, x, dpre):
            _run_wgrad(g, x.view(rows, 1, 1, in_dim), dpre.view(rows, 1, 1, out_dim), ctx.wvar)
            wg = False
            if bg:
                _bias_grad(dpre, ctx.bvar)
                bg = False
        if wg or bg:
            check(L.gg_linear_wgrad(ptr(x), dt(x), ptr(dpre), dt(dpre), ptr(ctx.wvar.grad) if wg else None,
                                    ptr(ctx.bvar.grad) if bg else None, rows, in_dim, out_dim, stream()), "gg_linear_wgrad")
        dx = None
        if ctx.needs_input_grad[0]:
            if x.dtype == torch.bfloat16 and _lin_tc(rows, in_dim, out_dim, dpre):
                dx = _run_up(g, dpre.view(rows, 1, 1, out_dim), ctx.wvar, None, x.dtype, None, 0.0, 4).view(rows, in_dim)
            else:
                dx = torch.empty_like(x)
                check(L.gg_linear_dgrad(ptr(dpre), dt(dpre), ptr(ctx.wvar.data), ptr(dx), dt(dx), rows, in_dim, out_dim, stream()),
                      "gg_linear_dgrad")
        return dx, None, None, None, None, None, None, None


EARLY_ADAM_TICK = os.environ.get("GG_EARLY_ADAM_TICK", "0") == "1"  # opt-in: Adam's step-counter tick on the side stream, under the forward pass (measured: 1.3596 vs 1.3532 ms/step -- slower)
FUSE_ACT_BIAS = os.environ.get("GG_FUSE_ACT_BIAS", "1") != "0"      # A/B switch: g_h4's tanh' + bias gradient as one launch
FUSE_WGRAD_BIAS = os.environ.get("GG_FUSE_WGRAD_BIAS", "1") != "0"  # A/B switch: d_h0_conv's bias gradient inside its filter-gradient launch
ZERO_ON_SIDE = os.environ.get("GG_ZERO_ON_SIDE", "1") != "0"        # A/B switch: gradient zero-fill under the forward pass
FUSE_LOSS_HEAD = os.environ.get("GG_FUSE_LOSS_HEAD", "1") != "0"    # A/B switch: d_h3_lin + cross-entropy means as two launches
_LH_TICKET = {}


class _LossHead(torch.autograd.Function):
    """linear(h, 1) + sum_i weight_i * reduce_mean(sigmoid_cross_entropy_with_logits(logits[a_i:b_i], target_i)) as ONE node:
    one forward launch (gg_loss_head_fwd: logits, loss parts, d loss / d logits) and one backward launch (gg_loss_head_bwd:
    Matrix and bias gradients, dh, and the backward reductions of the batch norm that produced h).  Returns (parts, logits);
    like _SigmoidCESum, `parts` must be the root of backward() (upstream gradient taken to be 1); logits carry no gradient."""

    @staticmethod
    def forward(ctx, h, w, b, mvar, bvar, segs, in_bn):
        L = cabi.lib()
        rows, in_dim = h.shape
        logits = torch.empty((rows, 1), dtype=torch.float32, device=h.device)
        parts = torch.empty(len(segs) + 1, dtype=torch.float32, device=h.device)
        need = any(ctx.needs_input_grad[:3])
        dl = torch.empty(rows, dtype=torch.float32, device=h.device) if need else None
        tk = _LH_TICKET.get(h.device)
        if tk is None:          # zero-filled once; the kernel hands it back zeroed
            tk = _LH_TICKET[h.device] = torch.zeros(4, dtype=torch.int32, device=h.device)
        I, F = ctypes.c_int32 * len(segs), ctypes.c_float * len(segs)
        check(L.gg_loss_head_fwd(ptr(h), dt(h), ptr(w), ptr(b), rows, in_dim, I(*[s[0] for s in segs]), I(*[s[1] for s in segs]),
                                 F(*[s[2] for s in segs]), F(*[s[3] for s in segs]), len(segs), ptr(logits), ptr(parts), ptr(dl), ptr(tk),
                                 stream()), "gg_loss_head_fwd")
        ctx.mvar, ctx.bvar, ctx.in_bn, ctx.dl = mvar, bvar, in_bn, dl
        ctx.save_for_backward(h)
        ctx.mark_non_differentiable(logits)
        ctx.set_materialize_grads(False)      # no zero-filled gradient tensor for the logits output (a fill launch per backward)
        return parts, logits

    @staticmethod
    def backward(ctx, g_parts, g_logits):
        (h,) = ctx.saved_tensors
        L = cabi.lib()
        rows, in_dim = h.shape
        wg, bg, xg = ctx.needs_input_grad[1], ctx.bvar is not None and ctx.needs_input_grad[2], ctx.needs_input_grad[0]
        dh = torch.empty_like(h) if xg else None
        bnb = ctx.in_bn
        use_bn = (xg and FUSE_BN_BWD and bnb is not None and bnb.pre is not None and bnb.pre.numel() == h.numel()
                  and bnb.pre.dtype == torch.float32 and in_dim % bnb.Cc == 0)
        sums = _zeroed_f64(L.gg_bn_workspace_bytes(bnb.Cc, bnb.groups) // 8, h.device)[0] if use_bn else None
        fused = ctypes.c_int32(0)
        check(L.gg_loss_head_bwd(ptr(h), dt(h), ptr(ctx.dl), ptr(ctx.mvar.data), rows, in_dim, ptr(ctx.mvar.grad) if wg else None,
                                 ptr(ctx.bvar.grad) if bg else None, ptr(dh), ptr(bnb.pre) if use_bn else None,
                                 ptr(bnb.mean) if use_bn else None, ptr(bnb.rstd) if use_bn else None, ptr(bnb.gamma) if use_bn else None,
                                 ptr(bnb.beta) if use_bn else None, ACT[bnb.act] if use_bn else 0, float(bnb.act_param) if use_bn else 0.0,
                                 bnb.groups if use_bn else 1, bnb.Cc if use_bn else 0, ptr(sums), ctypes.byref(fused), stream()),
              "gg_loss_head_bwd")
        if use_bn and fused.value:
            bnb.bwd_sums, bnb.bwd_dy, bnb.bwd_version = sums, dh, dh._version
        return dh, None, None, None, None, None, None


def linear(input_, output_size, scope=None, stddev=0.02, bias_start=0.0, with_w=False, *, act=None, act_param=0.2, out_dtype=None,
           bn=None, bn_channels=None, train=True, groups=1, ce_segments=None):
    """ops.py:106-117 -- tf.matmul(input_, Matrix) + bias; variables `scope/Matrix`, `scope/bias`.
    `bn=` fuses the batch norm (+activation) that follows; `bn_channels` is the channel count it normalises
    when the reference reshapes the output to [-1, h, w, bn_channels] first (model.py:306-307).
    `ce_segments=[(begin, end, target, weight), ...]` (output_size 1: the discriminator's last layer, model.py:277) also
    evaluates the sigmoid cross-entropy means of model.py:121-131 over those row ranges in the same launch; the logits are
    returned as usual and `sigmoid_cross_entropy_loss(logits, ce_segments)` hands the fused result over."""
    rows, in_dim = input_.shape
    with variable_scope(scope or "Linear") as st:
        mvar = st.get_variable("Matrix", [in_dim, output_size], random_normal_initializer(stddev))
        bvar = st.get_variable("bias", [output_size], constant_initializer(bias_start))
    od = out_dtype if out_dtype is not None else (torch.float32 if output_size <= 4 else act_dtype())
    if (ce_segments is not None and FUSE_LOSS_HEAD and output_size == 1 and bn is None and act is None and od == torch.float32
            and not _is_meta(input_) and cabi.lib().gg_loss_head_ok(rows, in_dim, len(ce_segments))):
        _require_cuda(input_, "linear")
        segs = tuple((int(a), int(b), float(t), float(w)) for a, b, t, w in ce_segments)
        parts, y = _LossHead.apply(input_.contiguous(), _wtensor(mvar, _wants_grad(mvar)), _wtensor(bvar, _wants_grad(bvar)), mvar, bvar,
                                   segs, getattr(input_, "_gg_bn", None))
        y._gg_ce = (segs, parts)
        return (y, mvar, bvar) if with_w else y
    if bn is not None:
        y = _fused_bn(input_, _LinearProducer(mvar, bvar, rows, output_size), bn, train, act, act_param, od, groups, bn_channels or output_size)
        return (y, mvar, bvar) if with_w else y
    if _is_meta(input_):
        y = torch.empty((rows, output_size), dtype=od, device="meta")
    else:
        _require_cuda(input_, "linear")
        y = _Linear.apply(input_.contiguous(), _wtensor(mvar, _wants_grad(mvar)), _wtensor(bvar, _wants_grad(bvar)), mvar, bvar, act,
                          act_param, od)
    return (y, mvar, bvar) if with_w else y


# ---- batch norm ----------------------------------------------------------------------
class _BatchNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, bn, train, act, act_param, out_dtype, groups):
        Cc = x.shape[-1]
        rows = x.numel() // Cc
        L = cabi.lib()
        y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
        save_mean = torch.empty((groups, Cc), dtype=torch.float32, device=x.device)
        save_rstd = torch.empty((groups, Cc), dtype=torch.float32, device=x.device)
        mm = bn.moving_mean.data if bn.moving_mean is not None else None
        mv = bn.moving_variance.data if bn.moving_variance is not None else None
        if train:
            nbytes = L.gg_bn_workspace_bytes(Cc, groups)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            check(L.gg_bn_fwd_train(ptr(x), dt(x), ptr(y), dt(y), rows, Cc, groups, ptr(gamma), ptr(beta), ptr(mm), ptr(mv),
                                    ptr(save_mean), ptr(save_rstd), bn.epsilon, bn.momentum, ACT[act], float(act_param),
                                    ptr(ws), nbytes, stream()), "gg_bn_fwd_train")
        else:
            check(L.gg_bn_infer_stats(ptr(mm), ptr(mv), bn.epsilon, Cc, ptr(save_mean), ptr(save_rstd), stream()), "gg_bn_infer_stats")
            check(L.gg_bn_fwd_infer(ptr(x), dt(x), ptr(y), dt(y), rows, Cc, ptr(gamma), ptr(beta), ptr(mm), ptr(mv), bn.epsilon,
                                    ACT[act], float(act_param), stream()), "gg_bn_fwd_infer")
        ctx.bn, ctx.train, ctx.act, ctx.act_param, ctx.groups = bn, train, act, act_param, groups
        ctx.save_for_backward(x, gamma, beta, save_mean, save_rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, save_mean, save_rstd = ctx.saved_tensors
        dy = dy.contiguous()
        bn = ctx.bn
        Cc = x.shape[-1]
        rows = x.numel() // Cc
        L = cabi.lib()
        gg_, bg_ = gamma is not None and ctx.needs_input_grad[1], beta is not None and ctx.needs_input_grad[2]
        dx = torch.empty_like(x)
        nbytes = L.gg_bn_workspace_bytes(Cc, ctx.groups)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        check(L.gg_bn_bwd(ptr(x), dt(x), ptr(dy), dt(dy), ptr(dx), dt(dx), rows, Cc, ctx.groups, ptr(gamma), ptr(beta),
                          ptr(save_mean), ptr(save_rstd), ptr(bn.gamma.grad) if gg_ else None, ptr(bn.beta.grad) if bg_ else None,
                          ACT[ctx.act], float(ctx.act_param), 1 if ctx.train else 0, ptr(ws), nbytes, stream()), "gg_bn_bwd")
        return dx, None, None, None, None, None, None, None, None


class batch_norm(object):
    """ops.py:10-24 -- tf.contrib.layers.batch_norm(decay=momentum, updates_collections=None,
    epsilon, scale=True, is_training=train): variables `name/{beta,gamma,moving_mean,moving_variance}`.
    Train mode updates the EMAs on every call (App. A.4)."""

    def __init__(self, epsilon=1e-5, momentum=0.9, name="batch_norm", affine=True, ema=True):
        self.epsilon, self.momentum, self.name = epsilon, momentum, name
        self.affine, self.ema = affine, ema
        self.beta = self.gamma = self.moving_mean = self.moving_variance = None

    def _vars(self, Cc):
        if self.affine and self.beta is None:
            with variable_scope(self.name) as st:
                self.beta = st.get_variable("beta", [Cc], constant_initializer(0.0))
                self.gamma = st.get_variable("gamma", [Cc], constant_initializer(1.0))
        if self.ema and self.moving_mean is None:
            with variable_scope(self.name) as st:
                self.moving_mean = st.get_variable("moving_mean", [Cc], constant_initializer(0.0), trainable=False)
                self.moving_variance = st.get_variable("moving_variance", [Cc], constant_initializer(1.0), trainable=False)

    def __call__(self, x, train=True, *, act=None, act_param=0.2, out_dtype=None, groups=1):
        self._vars(x.shape[-1])
        od = _out_dtype(act, out_dtype)
        if _is_meta(x):
            return torch.empty(x.shape, dtype=od, device="meta")
        _require_cuda(x, "batch_norm")
        if not train and not self.ema:
            raise ValueError("inference-mode batch_norm needs moving statistics")
        g = _wtensor(self.gamma, _wants_grad(self.gamma)) if self.affine else None
        b = _wtensor(self.beta, _wants_grad(self.beta)) if self.affine else None
        return _BatchNorm.apply(x.contiguous(), g, b, self, bool(train), act, act_param, od, groups)


# ---- fused producer + batch norm + activation -------------------------------------------------
class _ConvProducer:
    """A conv2d / conv3d ('down') or deconv2d ('up') feeding a batch norm."""

    def __init__(self, geom, direction, wvar, bvar, ndim):
        self.geom, self.direction, self.wvar, self.bvar, self.ndim = geom, direction, wvar, bvar, ndim

    def out_shape(self):
        return self.geom.small_shape(self.ndim) if self.direction == "down" else self.geom.large_shape(self.ndim)

    fuses_stats = True      # batch statistics come out of the conv call (GEMM epilogue on the tensor-core path)

    def fwd(self, x, b, stats=None, groups=1):
        run = _run_down if self.direction == "down" else _run_up
        return run(self.geom, x, self.wvar, b, torch.float32, None, 0.0, self.ndim, stats=stats, groups=groups)

    def wgrad(self, x, dpre):
        if self.direction == "down":
            _run_wgrad(self.geom, x, dpre, self.wvar)
        else:
            _run_wgrad(self.geom, dpre, x, self.wvar)

    def dgrad(self, dpre, x_dtype, bnb=None):
        run = _run_up if self.direction == "down" else _run_down
        return run(self.geom, dpre, self.wvar, None, x_dtype, None, 0.0, self.ndim, bnb=bnb)


class _LinearProducer:
    def __init__(self, mvar, bvar, rows, out_dim):
        self.wvar, self.bvar, self.rows, self.out_dim = mvar, bvar, rows, out_dim

    fuses_stats = True      # gg_linear_fwd_stats: batch statistics from the same launch (thin path) or a pass inside the call

    def out_shape(self):
        return (self.rows, self.out_dim)

    def fwd(self, x, b, stats=None, groups=1, Cc=None):
        rows, in_dim = x.shape
        if _lin_tc(rows, in_dim, self.out_dim, x) and (stats is None or (Cc or self.out_dim) == self.out_dim):
            # one-tap tcgen05 GEMM; the batch statistics (channel = column) come out of its epilogue like a conv's
            return _run_down(_lin_geom(rows, in_dim, self.out_dim), x.view(rows, 1, 1, in_dim), self.wvar, b, torch.float32, None, 0.0, 4,
                             stats=stats, groups=groups).view(rows, self.out_dim)
        y = torch.empty((rows, self.out_dim), dtype=torch.float32, device=x.device)
        if stats is not None:
            check(cabi.lib().gg_linear_fwd_stats(ptr(x), dt(x), ptr(self.wvar.data), ptr(b), ptr(y), rows, in_dim, self.out_dim,
                                                 Cc or self.out_dim, groups, ptr(stats), stream()), "gg_linear_fwd_stats")
            return y
        check(cabi.lib().gg_linear_fwd(ptr(x), dt(x), ptr(self.wvar.data), ptr(b), ptr(y), dt(y), rows, in_dim, self.out_dim, 0, 0.0,
                                       stream()), "gg_linear_fwd")
        return y

    def wgrad(self, x, dpre):
        rows, in_dim = x.shape
        if _lin_tc(rows, in_dim, self.out_dim, x, dpre):
            _run_wgrad(_lin_geom(rows, in_dim, self.out_dim), x.view(rows, 1, 1, in_dim), dpre.view(rows, 1, 1, self.out_dim), self.wvar)
            return
        check(cabi.lib().gg_linear_wgrad(ptr(x), dt(x), ptr(dpre), dt(dpre), ptr(self.wvar.grad), None, rows, in_dim, self.out_dim,
                                         stream()), "gg_linear_wgrad")

    def dgrad(self, dpre, x_dtype, bnb=None):
        rows = dpre.shape[0]
        in_dim = self.wvar.data.shape[0]
        if x_dtype == torch.bfloat16 and _lin_tc(rows, in_dim, self.out_dim, dpre):
            # the launch also carries the backward reductions of the batch norm that produced x (bnb), like a conv's dgrad
            return _run_up(_lin_geom(rows, in_dim, self.out_dim), dpre.view(rows, 1, 1, self.out_dim), self.wvar, None, x_dtype, None, 0.0, 4,
                           bnb=bnb).view(rows, in_dim)
        dx = torch.empty((rows, in_dim), dtype=x_dtype, device=dpre.device)
        check(cabi.lib().gg_linear_dgrad(ptr(dpre), dt(dpre), ptr(self.wvar.data), ptr(dx), dt(dx), rows, in_dim, self.out_dim, stream()),
              "gg_linear_dgrad")
        return dx


class _FusedBN(torch.autograd.Function):
    """producer (conv / deconv / conv3d / linear, + bias) -> batch norm -> activation as ONE autograd node.
    The pre-normalisation tensor stays fp32 inside the node in both precisions (the batch-norm backward
    subtracts nearly equal quantities; bf16 there costs several % of gradient accuracy), while everything that
    crosses the node boundary -- and every GEMM operand -- is in the activation dtype.
    A bias that feeds a TRAIN-mode batch norm has an exactly zero gradient (the mean subtraction cancels it);
    it is left at zero rather than filled with rounding noise."""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, prod, bn, train, act, act_param, out_dtype, groups, Cc, in_bn=None, info=None):
        L = cabi.lib()
        ctx.in_bn, ctx.info = in_bn, info
        ctx.linbn = False
        if (FUSE_LINEAR_BN and train and isinstance(prod, _LinearProducer) and x.dim() == 2
                and L.gg_linear_bn_ok(x.shape[0], x.shape[1], prod.out_dim, Cc, groups)):
            # thin linear + batch statistics + normalise + activation in ONE launch (a CTA owns whole channels)
            rows, in_dim = x.shape
            pre = torch.empty((rows, prod.out_dim), dtype=torch.float32, device=x.device)
            y = torch.empty((rows, prod.out_dim), dtype=out_dtype, device=x.device)
            save_mean = torch.empty((1, Cc), dtype=torch.float32, device=x.device)
            save_rstd = torch.empty((1, Cc), dtype=torch.float32, device=x.device)
            mm = bn.moving_mean.data if bn.moving_mean is not None else None
            mv = bn.moving_variance.data if bn.moving_variance is not None else None
            check(L.gg_linear_bn_fwd(ptr(x), dt(x), ptr(prod.wvar.data), ptr(b), ptr(gamma), ptr(beta), ptr(mm), ptr(mv), ptr(pre), ptr(y), dt(y),
                                     ptr(save_mean), ptr(save_rstd), rows, in_dim, prod.out_dim, Cc, bn.epsilon, bn.momentum, ACT[act],
                                     float(act_param), stream()), "gg_linear_bn_fwd")
            ctx.prod, ctx.bn, ctx.train, ctx.act, ctx.act_param, ctx.groups, ctx.Cc = prod, bn, train, act, act_param, groups, Cc
            ctx.x_dtype, ctx.linbn = x.dtype, True
            ctx.save_for_backward(x, pre, gamma, beta, save_mean, save_rstd)
            if DEBUG_TAP is not None:
                DEBUG_TAP.setdefault("fwd", []).append((prod.wvar.name, pre, y))
            return y
        is_lin = isinstance(prod, _LinearProducer)
        fused_stats = train and prod.fuses_stats and (Cc == prod.out_shape()[-1] or (is_lin and x.dim() == 2 and x.shape[0] % groups == 0))
        stats = _zeroed_f64(L.gg_bn_workspace_bytes(Cc, groups) // 8, x.device)[0] if fused_stats else None   # [replicas][groups][2][C]
        pre = prod.fwd(x, b, stats=stats, groups=groups, Cc=Cc) if is_lin else prod.fwd(x, b, stats=stats, groups=groups)
        rows = pre.numel() // Cc
        y = torch.empty(pre.shape, dtype=out_dtype, device=x.device)
        save_mean = torch.empty((groups, Cc), dtype=torch.float32, device=x.device)
        save_rstd = torch.empty((groups, Cc), dtype=torch.float32, device=x.device)
        mm = bn.moving_mean.data if bn.moving_mean is not None else None
        mv = bn.moving_variance.data if bn.moving_variance is not None else None
        if fused_stats:
            check(L.gg_bn_fwd_train_stats(ptr(pre), dt(pre), ptr(y), dt(y), rows, Cc, groups, ptr(gamma), ptr(beta), ptr(mm), ptr(mv),
                                          ptr(save_mean), ptr(save_rstd), bn.epsilon, bn.momentum, ACT[act], float(act_param),
                                          ptr(stats), stream()), "gg_bn_fwd_train_stats")
        elif train:
            nbytes = L.gg_bn_workspace_bytes(Cc, groups)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            check(L.gg_bn_fwd_train(ptr(pre), dt(pre), ptr(y), dt(y), rows, Cc, groups, ptr(gamma), ptr(beta), ptr(mm), ptr(mv),
                                    ptr(save_mean), ptr(save_rstd), bn.epsilon, bn.momentum, ACT[act], float(act_param),
                                    ptr(ws), nbytes, stream()), "gg_bn_fwd_train")
        else:
            check(L.gg_bn_infer_stats(ptr(mm), ptr(mv), bn.epsilon, Cc, ptr(save_mean), ptr(save_rstd), stream()), "gg_bn_infer_stats")
            check(L.gg_bn_fwd_infer(ptr(pre), dt(pre), ptr(y), dt(y), rows, Cc, ptr(gamma), ptr(beta), ptr(mm), ptr(mv), bn.epsilon,
                                    ACT[act], float(act_param), stream()), "gg_bn_fwd_infer")
        ctx.prod, ctx.bn, ctx.train, ctx.act, ctx.act_param, ctx.groups, ctx.Cc = prod, bn, train, act, act_param, groups, Cc
        ctx.x_dtype = x.dtype
        ctx.save_for_backward(x, pre, gamma, beta, save_mean, save_rstd)
        if info is not None and train:
            info.pre, info.mean, info.rstd, info.gamma, info.beta = pre, save_mean, save_rstd, gamma, beta
            info.act, info.act_param, info.groups, info.Cc = act, act_param, groups, Cc
        if DEBUG_TAP is not None:
            DEBUG_TAP.setdefault("fwd", []).append((prod.wvar.name, pre, y))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, pre, gamma, beta, save_mean, save_rstd = ctx.saved_tensors
        dy = dy.contiguous()
        L = cabi.lib()
        bn, prod, Cc = ctx.bn, ctx.prod, ctx.Cc
        rows = pre.numel() // Cc
        need_w, need_b = ctx.needs_input_grad[1], prod.bvar is not None and ctx.needs_input_grad[2]
        need_g, need_be = gamma is not None and ctx.needs_input_grad[3], beta is not None and ctx.needs_input_grad[4]
        if ctx.linbn and need_w and not ctx.needs_input_grad[0] and DEBUG_TAP is None:
            # batch-norm backward + filter gradient of the thin linear in ONE launch (dpre never leaves the CTA)
            r_, in_dim = x.shape
            check(L.gg_linear_bn_bwd(ptr(x), dt(x), ptr(pre), ptr(dy), dt(dy), ptr(gamma), ptr(beta), ptr(save_mean), ptr(save_rstd),
                                     ptr(prod.wvar.grad), ptr(bn.gamma.grad) if need_g else None, ptr(bn.beta.grad) if need_be else None,
                                     r_, in_dim, prod.out_dim, Cc, ACT[ctx.act], float(ctx.act_param), 1 if act_dtype() == torch.bfloat16 else 0,
                                     stream()), "gg_linear_bn_bwd")
            return (None,) * 15
        dpre = torch.empty(pre.shape, dtype=act_dtype(), device=pre.device)     # GEMM operand precision
        nbytes = L.gg_bn_workspace_bytes(Cc, ctx.groups)
        info = ctx.info
        if (ctx.train and info is not None and info.bwd_sums is not None and dy.data_ptr() == info.bwd_dy.data_ptr()
                and dy._version == info.bwd_version and dy.dtype == info.bwd_dy.dtype and dy.numel() == info.bwd_dy.numel()):
            # the dgrad launch that wrote dy (the ONLY contribution to it: an accumulated gradient is another tensor or another
            # version) already reduced (sum g, sum g*xhat) in its epilogue: apply pass only
            ws, mode = info.bwd_sums, 3
        else:
            ws, prezeroed = _zeroed_f64(nbytes // 8, pre.device) if ctx.train else (torch.empty(nbytes // 8, dtype=torch.float64, device=pre.device), False)
            mode = (2 if prezeroed else 1) if ctx.train else 0
        if info is not None:
            info.bwd_sums = info.bwd_dy = info.pre = None
        check(L.gg_bn_bwd(ptr(pre), dt(pre), ptr(dy), dt(dy), ptr(dpre), dt(dpre), rows, Cc, ctx.groups, ptr(gamma), ptr(beta),
                          ptr(save_mean), ptr(save_rstd), ptr(bn.gamma.grad) if need_g else None, ptr(bn.beta.grad) if need_be else None,
                          ACT[ctx.act], float(ctx.act_param), mode, ptr(ws), nbytes, stream()), "gg_bn_bwd")
        if DEBUG_TAP is not None:
            DEBUG_TAP.setdefault("bwd", []).append((prod.wvar.name, dy, dpre))
        if need_b and not ctx.train:
            _bias_grad(dpre.reshape(-1, prod.bvar.data.numel()), prod.bvar)
        if need_w:
            prod.wgrad(x, dpre)
        dx = prod.dgrad(dpre, ctx.x_dtype, bnb=ctx.in_bn) if ctx.needs_input_grad[0] else None
        return (dx,) + (None,) * 14


def _fused_bn(input_, prod, bn, train, act, act_param, out_dtype, groups, channels=None):
    shape = prod.out_shape()
    Cc = channels if channels is not None else shape[-1]
    bn._vars(Cc)
    if _is_meta(input_):
        return torch.empty(shape, dtype=out_dtype, device="meta")
    _require_cuda(input_, "fused batch norm")
    if not train and not bn.ema:
        raise ValueError("inference-mode batch_norm needs moving statistics")
    w = _wtensor(prod.wvar, _wants_grad(prod.wvar))
    b = _wtensor(prod.bvar, _wants_grad(prod.bvar)) if prod.bvar is not None else None
    g = _wtensor(bn.gamma, _wants_grad(bn.gamma)) if bn.affine else None
    be = _wtensor(bn.beta, _wants_grad(bn.beta)) if bn.affine else None
    info = _BnInfo() if (train and torch.is_grad_enabled() and FUSE_BN_BWD) else None
    y = _FusedBN.apply(input_.contiguous(), w, b, g, be, prod, bn, bool(train), act, act_param, out_dtype, groups, Cc,
                       getattr(input_, "_gg_bn", None), info)
    if info is not None and y.requires_grad:
        y._gg_bn = info          # picked up by the conv that consumes y (its dgrad produces this batch norm's backward reductions)
    return y


# ---- standalone activations ------------------------------------------------------------
class _Act(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act, act_param, out_dtype):
        y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
        check(cabi.lib().gg_act_fwd(ptr(x), dt(x), ptr(y), dt(y), x.numel(), ACT[act], float(act_param), stream()), "gg_act_fwd")
        ctx.act, ctx.act_param, ctx.x_dtype = act, act_param, x.dtype
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty(y.shape, dtype=ctx.x_dtype, device=y.device)
        check(cabi.lib().gg_act_bwd(ptr(y), dt(y), ptr(dy), dt(dy), ptr(dx), dt(dx), y.numel(), ACT[ctx.act], float(ctx.act_param),
                                    stream()), "gg_act_bwd")
        return dx, None, None, None


def _activation(x, act, act_param=0.0, out_dtype=None):
    od = out_dtype if out_dtype is not None else x.dtype
    if _is_meta(x):
        return torch.empty(x.shape, dtype=od, device="meta")
    _require_cuda(x, act)
    return _Act.apply(x.contiguous(), act, act_param, od)


def lrelu(x, leak=0.2, name="lrelu"):
    """ops.py:103-104 -- tf.maximum(x, leak*x) (gradient 1 at x == 0)."""
    return _activation(x, "lrelu", leak)


def relu(x):
    return _activation(x, "relu")


def tanh(x, out_dtype=None):
    return _activation(x, "tanh", 0.0, out_dtype)


def sigmoid(x, out_dtype=None):
    return _activation(x, "sigmoid", 0.0, out_dtype)


# ---- misc reference ops --------------------------------------------------------------------
def add_noise(inpt, stddev):
    """ops.py:119-123.  Identity at stddev 0 (the default everywhere, z_model.py:50-51).  For
    stddev > 0 Gaussian noise is added with torch's generator (TF's RNG stream cannot be reproduced)."""
    if stddev == 0.0:
        return inpt
    if _is_meta(inpt):
        return inpt
    return inpt + torch.randn(inpt.shape, device=inpt.device, dtype=torch.float32).to(inpt.dtype) * stddev


def get_std(inpt):
    """ops.py:125-128 -- sqrt(mean(var over axis 0)); diagnostic, no gradient."""
    if _is_meta(inpt):
        return torch.empty((), dtype=torch.float32, device="meta")
    _require_cuda(inpt, "get_std")
    x = inpt.detach().contiguous()
    B = x.shape[0]
    F_ = x.numel() // B
    out = torch.empty(1, dtype=torch.float32, device=x.device)
    ws = torch.empty(2 * F_ * 8, dtype=torch.uint8, device=x.device)
    check(cabi.lib().gg_get_std(ptr(x), dt(x), B, F_, ptr(out), ptr(ws), ws.numel(), stream()), "gg_get_std")
    return out[0]


def frames_to_input(frames_u8, size, swap_rb=True, out=None):
    """The tail of the reference's frame decode on the GPU (z_model_lib.py:339-346, utils.py:57-63):
    `transform(cv2.cvtColor(cv2.resize(im, (size, size), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2RGB), is_crop=False)` for a
    whole batch of decoded uint8 frames [n, H0, W0, 3] already in device memory -> float32 [n, size, size, 3] in [-1, 1], bit for
    bit what the host code produces (gg_frames_to_input).  `out=` writes into a caller buffer (a slice of the step's static input)."""
    if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[-1] != 3:
        raise ValueError("frames_to_input: expects uint8 frames [n, H, W, 3]")
    n, H0, W0, _ = frames_u8.shape
    if _is_meta(frames_u8):
        return torch.empty((n, size, size, 3), dtype=torch.float32, device="meta")
    _require_cuda(frames_u8, "frames_to_input")
    fr = frames_u8.contiguous()
    if out is None:
        out = torch.empty((n, size, size, 3), dtype=torch.float32, device=fr.device)
    elif tuple(out.shape) != (n, size, size, 3) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError(f"out= must be a contiguous float32 tensor of shape {(n, size, size, 3)}")
    check(cabi.lib().gg_frames_to_input(ptr(fr), n, H0, W0, H0 * W0 * 3, W0 * 3, ptr(out), size, size, 1 if swap_rb else 0, stream()),
          "gg_frames_to_input")
    return out


def conv_cond_concat(x, y):
    """ops.py:45-49 -- concat y broadcast over H, W on the channel axis (MNIST branch; tensor plumbing)."""
    B, H, W, _ = x.shape
    return torch.cat([x, y.to(x.dtype).expand(B, H, W, y.shape[3])], dim=3)


class _SigmoidCESum(torch.autograd.Function):
    """sum over row segments of  weight * reduce_mean(sigmoid_cross_entropy_with_logits(logits[a:b], target))
    with the backward fused into the forward launch (model.py:121-131).  Must be the root of backward()
    (the upstream gradient is taken to be 1)."""

    @staticmethod
    def forward(ctx, logits, segments):
        L = cabi.lib()
        logits = logits.contiguous()
        parts = torch.empty(len(segments) + 1, dtype=torch.float32, device=logits.device)   # [total, part_0, ...]
        need = ctx.needs_input_grad[0]
        dl = torch.empty_like(logits) if need else None
        flat = logits.reshape(-1)
        dflat = dl.reshape(-1) if need else None
        for i, (a, b, tg, wt) in enumerate(segments):
            check(L.gg_sigmoid_ce(ptr(flat[a:b]), b - a, float(tg), float(wt), ptr(parts[i + 1:i + 2]), 0,
                                  ptr(dflat[a:b]) if need else None, stream()), "gg_sigmoid_ce")
            check(L.gg_axpby(ptr(parts[i + 1:i + 2]), 1.0, ptr(parts[0:1]), 0.0 if i == 0 else 1.0, 1, stream()), "gg_axpby")
        ctx.dl = dl
        return parts

    @staticmethod
    def backward(ctx, g_parts):
        return ctx.dl, None


def sigmoid_cross_entropy_loss(logits, segments=None, target=None):
    """Returns a float32 vector [total, part_0, part_1, ...]; total = sum_i w_i * mean(CE(logits[a_i:b_i], t_i)).
    `segments` = [(begin, end, target, weight), ...] over the rows of `logits`; or pass a single `target`."""
    n = logits.shape[0]
    if segments is None:
        segments = [(0, n, target, 1.0)]
    if _is_meta(logits):
        return torch.empty(len(segments) + 1, dtype=torch.float32, device="meta")
    if logits.dtype != torch.float32:
        raise TypeError("logits must be float32")
    ce = getattr(logits, "_gg_ce", None)
    if ce is not None and ce[0] == tuple((int(a), int(b), float(t), float(w)) for a, b, t, w in segments):
        return ce[1]            # linear(..., ce_segments=...) already evaluated exactly these means (gg_loss_head_fwd)
    return _SigmoidCESum.apply(logits, segments)


_DIST_WS = {}


class _DistanceLoss(torch.autograd.Function):
    """w_l2 * mean((a - target)^2) + w_l1 * mean(|a - target|) with the gradient written by the same launch
    (z_space_finder.py:258-292).  Like _SigmoidCESum it must be a root of backward() (upstream gradient 1)."""

    @staticmethod
    def forward(ctx, a, target, w_l2, w_l1):
        L = cabi.lib()
        a = a.contiguous()
        ws = _DIST_WS.get(a.device)
        if ws is None:          # zero-filled once; the kernel hands it back zeroed
            ws = _DIST_WS[a.device] = torch.zeros(L.gg_distance_loss_workspace_bytes(), dtype=torch.uint8, device=a.device)
        loss = torch.empty(1, dtype=torch.float32, device=a.device)
        da = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        check(L.gg_distance_loss(ptr(a), dt(a), ptr(target), a.numel(), float(w_l2), float(w_l1), ptr(loss), 0, ptr(da), ptr(ws),
                                 ws.numel(), stream()), "gg_distance_loss")
        ctx.da = da
        return loss

    @staticmethod
    def backward(ctx, g):
        return ctx.da, None, None, None


def distance_loss(a, target, w_l2=1.0, w_l1=0.0):
    """tf.reduce_mean(tf.reduce_mean(w_l2*tf.square(a - target) + w_l1*tf.abs(a - target), [1,2,3])) as a float32 [1]
    tensor.  `target` is a constant float32 tensor of a's shape (a placeholder in the reference)."""
    if _is_meta(a):
        return torch.empty(1, dtype=torch.float32, device="meta")
    _require_cuda(a, "distance_loss")
    if target.dtype != torch.float32 or tuple(target.shape) != tuple(a.shape) or not target.is_contiguous():
        raise ValueError("distance_loss: target must be a contiguous float32 tensor of the input's shape")
    return _DistanceLoss.apply(a, target, w_l2, w_l1)


def binary_cross_entropy(preds, targets, name=None):
    """ops.py:27-43 (unused by the reference's models; kept for API completeness, torch plumbing)."""
    eps = 1e-12
    return torch.mean(-(targets * torch.log(preds + eps) + (1. - targets) * torch.log(1. - preds + eps)))


# ---- optimiser -----------------------------------------------------------------------------
class AdamOptimizer:
    """tf.train.AdamOptimizer(learning_rate, beta1).minimize(loss, var_list) (model.py:153-156),
    TF semantics: one fused launch over the group's flat range."""

    def __init__(self, store: VariableStore, group: str, learning_rate=2e-4, beta1=0.5, beta2=0.999, epsilon=1e-8):
        self.store, self.group = store, group
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon
        self.t = 0
        self.var_list = None
        self.state = torch.zeros(4, dtype=torch.int32, device=store.device)   # [t, lr_t bits, ticket, -] (device-side, graph-safe)

    def range(self):
        """Element range of this optimiser's var_list in the flat buffers.  `group` may be a tuple of
        adjacent groups (e.g. video vars + image vars when --train_img_disc is set)."""
        if isinstance(self.group, (tuple, list)):
            rs = [self.store.ranges[g] for g in self.group]
            for (_, e0), (b1, _) in zip(rs, rs[1:]):
                if e0 != b1:
                    raise ValueError(f"optimiser groups {self.group} are not adjacent in the flat buffer")
            return rs[0][0], rs[-1][1]
        return self.store.ranges[self.group]

    def zero_grad(self, overlap=False, tick=False):
        """Zero the group's gradient range.  overlap=True issues the fill on the side stream (GG_ZERO_ON_SIDE=0 disables), so
        it runs under the update's forward pass -- one-wave launches that leave SMs and all of HBM idle -- instead of in front
        of it; the caller must `join_side()` before the first gradient kernel (the filter gradients queue up behind the fill
        on the side stream by themselves).  tick=True (the caller WILL call apply() for this update) also advances Adam's
        device-side step counter there (gg_adam_tick), so apply() is the Adam launch alone."""
        b, e = self.range()
        if overlap and ZERO_ON_SIDE and torch.cuda.is_available():
            s = _side_stream()
            s["stream"].wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s["stream"]):
                self.store.flat["grads"][b:e].zero_()
                if tick and EARLY_ADAM_TICK:
                    check(cabi.lib().gg_adam_tick(ptr(self.state), self.lr, self.b1, self.b2, stream()), "gg_adam_tick")
                    self._ticked = True
            s["dirty"] = True
            return
        self.store.flat["grads"][b:e].zero_()

    def lr_t(self, t):
        return self.lr * float(np.sqrt(1.0 - self.b2 ** t)) / (1.0 - self.b1 ** t)

    def apply(self, grad_scale=1.0):
        """One fused Adam launch over the group.  The step counter t lives on the device
        (self.state) so the call can be captured in a CUDA graph and replayed."""
        self.t += 1
        b, e = self.range()
        f = self.store.flat
        sh = getattr(self.store, "shadow", None)
        if getattr(self, "_ticked", False):       # zero_grad(tick=True) advanced the step counter under the forward pass
            self._ticked = False
            check(cabi.lib().gg_adam_apply(ptr(f["params"][b:e]), ptr(sh[b:e]) if sh is not None else None, ptr(f["grads"][b:e]), ptr(f["m"][b:e]),
                                           ptr(f["v"][b:e]), e - b, ptr(self.state), self.b1, self.b2, self.eps, grad_scale, stream()),
                  "gg_adam_apply")
        else:
            check(cabi.lib().gg_adam_graph(ptr(f["params"][b:e]), ptr(sh[b:e]) if sh is not None else None, ptr(f["grads"][b:e]), ptr(f["m"][b:e]),
                                           ptr(f["v"][b:e]), e - b, ptr(self.state), self.lr, self.b1, self.b2, self.eps, grad_scale, stream()),
                  "gg_adam_graph")
        for v in self.var_list or []:
            v.version += 1
            if sh is not None and v._bf16 is not None:
                v._bf16_version = v.version          # the same launch rewrote the bf16 shadow of the whole range
