"""Define-by-run graph engine over the C ABI (``include/dl4ds_b200.h``).

The reference hands its graphs (dl4ds/models/*.py) to Keras, which executes and differentiates
them.  Here a model is a Python function over a :class:`Ctx`; every op launches hand-written CUDA
kernels through ctypes and records its backward closure on a tape.  PyTorch supplies device
storage and the stream only -- no torch operator touches activations or weights on this path.

Storage conventions
  * activations: fp32 NHWC, possibly a channel slice of a wider buffer (:class:`Var` keeps the base
    buffer, channel offset and pitch) so Concatenate never needs a transposition;
  * parameters: one flat fp32 arena (:class:`Arena`) holding theta / grad / Adam m / v; parameter
    gradients ACCUMULATE into ``grad`` (zeroed once per step) which is what sums the gradient of
    shared layers (blocks.py:415,421-422,528-531) over their applications and lets the
    data-parallel all-reduce be a single collective on one buffer.
"""
import ctypes
import os
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from ._lib import ACT, MATH, W_FLIP_T, W_HWIO, W_PREPACKED, call

# losses.py:5-151: every entry of LOSS_FUNCTIONS as a weighted sum of kernel terms (pixel terms first)
LOSS_TERMS = {
    'mae': (('mae', 1.0),), 'mse': (('mse', 1.0),),
    'dssim': (('dssim', 1.0),),
    'dssim_mae': (('mae', 0.2), ('dssim', 0.8)),
    'dssim_mse': (('mse', 0.2), ('dssim', 0.8)),
    'dssim_mae_mse': (('mae', 0.2), ('mse', 0.2), ('dssim', 0.6)),
    'msdssim': (('msdssim', 1.0),),
    'msdssim_mae': (('mae', 0.2), ('msdssim', 0.8)),
    'msdssim_mae_mse': (('mae', 0.2), ('mse', 0.2), ('msdssim', 0.6)),
}
MSSSIM_POWER_FACTORS = (0.0448, 0.2856, 0.3001, 0.2363)      # losses.py:128
DROPOUT_KIND = {None: 0, 'vanilla': 0, 'mcdrop': 0, 'gaussian': 1, 'mcgaussiandrop': 1, 'spatial': 2,
                'mcspatialdrop': 2, 'droppath': 3}           # blocks.py:680-706 (+ DropPath, :106-129) -> dl4ds_dropout variant
RESIZE_METHOD = {'bilinear': 0, 'nearest': 1, 'bicubic': 2}     # DL4DS_RESIZE_*
LOSS_ACCUMULATE = 16                                          # DL4DS_LOSS_ACCUMULATE, OR-ed into `kind`
_PF_HOST = (ctypes.c_float * len(MSSSIM_POWER_FACTORS))(*MSSSIM_POWER_FACTORS)
_SKIP_WGRAD = os.environ.get('DL4DS_SKIP_WGRAD', '0') == '1'



def _stream():
    return torch.cuda.current_stream().cuda_stream


def same_pads(n, k, s):
    """TF 'SAME' padding: out = ceil(n/s); total = max((out-1)*s + k - n, 0); before = total//2."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2


class Arena:
    """Flat fp32 parameter arena: theta, grad and the two Adam slots, with named views."""

    def __init__(self, spec, device):
        self.spec = OrderedDict((k, tuple(int(s) for s in v)) for k, v in spec.items())
        self.offsets = OrderedDict()
        n = 0
        for name, shape in self.spec.items():
            self.offsets[name] = n
            n += int(np.prod(shape))
            n = (n + 3) // 4 * 4            # keep every tensor 16-byte aligned
        self.n = max(n, 4)
        self.device = torch.device(device)
        self.theta = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros_like(self.theta)
        self.m = torch.zeros_like(self.theta)
        self.v = torch.zeros_like(self.theta)
        self.t = 0                          # optimizer iterations

    def n_params(self):
        return sum(int(np.prod(s)) for s in self.spec.values())

    def _view(self, flat, name):
        shape = self.spec[name]
        o = self.offsets[name]
        return flat[o:o + int(np.prod(shape))].view(shape)

    def param(self, name):
        return self._view(self.theta, name)

    def gradient(self, name):
        return self._view(self.grad, name)

    def load(self, weights):
        """Copy a {name: array-like} dict (Keras layouts) into theta."""
        for name in self.spec:
            w = weights[name]
            w = torch.as_tensor(np.asarray(w, dtype=np.float32) if not torch.is_tensor(w) else w)
            assert tuple(w.shape) == self.spec[name], (name, tuple(w.shape), self.spec[name])
            self.param(name).copy_(w.to(torch.float32))

    def state_dict(self):
        return OrderedDict((n, self.param(n).detach().cpu().numpy().copy()) for n in self.spec)

    def grads(self):
        return OrderedDict((n, self.gradient(n).detach().cpu().numpy().copy()) for n in self.spec)

    def zero_grad(self):
        self.grad.zero_()

    # dropout masks: DEVICE uint64[2] {seed, step} read by dl4ds_dropout, bumped by dl4ds_rng_advance
    RNG_SEED = 0x5DEECE66D

    def rng_state(self):
        if getattr(self, '_rng', None) is None:
            rank = 0
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    rank = dist.get_rank()          # independent masks per replica, as with Horovod
            except Exception:
                rank = 0
            self._rng = torch.tensor([self.RNG_SEED + 7919 * rank, 0], dtype=torch.int64, device=self.device)
        return self._rng

    def seed_rng(self, seed, step=0):
        self._rng = torch.tensor([int(seed), int(step)], dtype=torch.int64, device=self.device)


class Var:
    """fp32 NHWC activation: channels [off, off+C) of ``buf`` (N,H,W,ld)."""
    __slots__ = ('buf', 'off', 'C', 'grad', 'requires_grad', 'n_uses', 'n_contrib', 'act_src', 'premasked')

    def __init__(self, buf, off=0, C=None, requires_grad=True):
        assert buf.dim() == 4 and buf.is_contiguous() and buf.dtype == torch.float32
        self.buf = buf
        self.off = off
        self.C = buf.shape[3] - off if C is None else C
        self.grad = None
        self.requires_grad = requires_grad
        # bookkeeping for the fused epilogue-backward (Ctx.conv): consumers registered in the forward pass, gradient
        # contributions received in the backward pass, the producing convolution's (activation, bias-gradient name), and
        # whether .grad already holds d(pre-activation) (the last consumer's dgrad applied act' and summed the bias grad)
        self.n_uses = 0
        self.n_contrib = 0
        self.act_src = None
        self.premasked = False

    @property
    def N(self):
        return self.buf.shape[0]

    @property
    def H(self):
        return self.buf.shape[1]

    @property
    def W(self):
        return self.buf.shape[2]

    @property
    def ld(self):
        return self.buf.shape[3]

    @property
    def ptr(self):
        return self.buf.data_ptr() + 4 * self.off

    @property
    def npix(self):
        return self.N * self.H * self.W

    @property
    def t(self):
        return self.buf[..., self.off:self.off + self.C]

    def slice(self, off, C):
        return Var(self.buf, self.off + off, C, self.requires_grad)

    def like(self, C=None, pad=False):
        """A new tensor of this shape.  ``pad``: a channel count that is no multiple of 4 (98 = 48 + 48 + 2 after a
        concatenation with static / localized channels) gets a pitch rounded up to 4 floats, so rows and 4-channel
        aligned slices stay 16-byte aligned and the vector / TMA kernels apply (the Var is then a slice of its buffer)."""
        C = self.C if C is None else C
        ld = padded_channels(C) if pad else C
        return Var(torch.empty((self.N, self.H, self.W, ld), dtype=torch.float32, device=self.buf.device), 0, C)


def padded_channels(C):
    return (C + 3) // 4 * 4 if (C % 4 and C > 4) else C


def new_var(N, H, W, C, device, zero=False):
    f = torch.zeros if zero else torch.empty
    return Var(f((N, H, W, C), dtype=torch.float32, device=device))


class PackPlan:
    """Every tensor-core weight image of a model packed by ONE kernel (``dl4ds_conv2d_pack_multi``): built from the
    model's pack cache after an eager pass has discovered which (parameter, pass, shape) images the graph uses.  A
    captured step launches it first; the convolutions then find their images prepacked."""

    def __init__(self, arena, pack_cache):
        import numpy as np
        lib = _lib.load()
        self.keys = set()
        recs = np.zeros((max(len(pack_cache), 1), 64), dtype=np.uint8)
        total = 0
        n = 0
        self._keep = []
        for key, ws in pack_cache.items():
            wname, wmode, k, cin, cout, math = key
            w = arena.param(wname)
            row = recs[n]
            units = lib.dl4ds_conv2d_pack_desc(w.data_ptr(), wmode, k, k, cin, cout, math, ws.data_ptr(),
                                               row.ctypes.data)
            if units < 0:
                raise _lib.Dl4dsError('conv2d_pack_desc: %s' % _lib.last_error())
            row[56:64].view(np.int64)[0] = total
            total += int(units)
            n += 1
            self.keys.add(key)
            self._keep.append(ws)
        self.n, self.total = n, total
        self.descs = torch.from_numpy(recs[:max(n, 1)].copy()).to(arena.device)

    def run(self):
        if self.n:
            call('dl4ds_conv2d_pack_multi', self.descs.data_ptr(), self.n, self.total, _stream())
        return 1 if self.n else 0


class Ctx:
    """One recorded forward pass (and its backward)."""
    TIMER_REPS = 4

    def __init__(self, arena, math='fp32', training=True):
        self.arena = arena
        self.math = MATH[math] if isinstance(math, str) else math
        self.training = training
        self.tape = []
        self.device = arena.device
        self.names_used = []
        self.launches = 0
        self.param_grads = True     # False: backward propagates to inputs only (cGAN G-through-D pass)
        self.timers = None          # {label: [(start_event, end_event), ...]} when profiling (bench.py)
        # tensor-core weight images: with a persistent cache ({(param, wmode, shape): uint8 tensor}, owned by the
        # model) every (layer, pass) is packed once per Ctx -- shared layers are not re-packed per application --
        # and, when `pack_stream` is set (the optimizer step's capture), on a side stream forked at the start of
        # the step, so the ~30 small pack kernels leave the critical path of the captured graph.
        self.pack_cache = None
        self.pack_stream = None
        self.prepacked = None       # keys of a PackPlan that already ran on this stream: their images are up to date
        self._packed = set()
        self._pack_forked = False
        # weight-gradient kernels on a side stream (the captured optimizer step): a wgrad only feeds the gradient arena,
        # so it can overlap the dgrad chain that the rest of the backward pass waits on.  Both kernel families are
        # persistent one-CTA-per-SM grids with 3-5 tiles per CTA on the 32x32 layers: run alone, a quarter of every
        # launch is ramp and tail; two streams let the block scheduler fill one kernel's tail with the other's CTAs.
        # Only convolutions whose dz buffer has no later in-place writer qualify (no residual hand-over); the buffers a
        # side-stream kernel reads are kept alive until the join in backward().
        self.wgrad_stream = None
        self.first_use = {}
        self.bwd_hooks = {}         # {tape index: callable}: run once backward has executed every entry >= that index
        self._side_keep = []
        self._side_used = False
        self._rng_advanced = False  # dropout: the arena's RNG step is bumped once per Ctx, before the first mask
        self._dropout_calls = 0     # ... and every dropout application gets its own layer id

    # ---------------------------------------------------------------- helpers
    def _call(self, name, *args):
        self.launches += 1
        return call(name, *args)

    def _timed(self, label, name, *args):
        """``_call`` bracketed by CUDA events on the launching stream when ``self.timers`` is set
        (per-kernel durations for bench.py's roofline line; never active inside graph capture)."""
        if self.timers is None:
            return self._call(name, *args)
        # profiling pass: one launch for the result, then TIMER_REPS back-to-back launches of the same call
        # between two events -- the queue stays full, so the CPU-side launch latency (tens of microseconds per
        # ctypes call in eager mode) does not leak into the per-launch duration.  The repeats may accumulate into
        # gradient buffers: the caller restores the optimizer state after a profiling step.
        rc = self._call(name, *args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(self.TIMER_REPS):
            self._call(name, *args)
        e1.record()
        self.launches -= self.TIMER_REPS
        self.timers.setdefault(label, []).append((e0, e1))
        return rc

    def input(self, tensor, requires_grad=False):
        """Wrap an NHWC fp32 CUDA tensor as a graph input."""
        assert tensor.is_cuda and tensor.dtype == torch.float32 and tensor.dim() == 4
        return Var(tensor.contiguous(), requires_grad=requires_grad)

    def _conv_ws(self, N, H, W, Cin, Ho, Wo, Cout, k, stride, up, r):
        """Packed-weight workspace for the tensor-core convolution path (None when not needed)."""
        if self.math == 0:
            return None
        nb = _lib.load().dl4ds_conv2d_fwd_workspace_bytes(N, H, W, Cin, Ho, Wo, Cout, k, k, stride, up, r, self.math)
        if nb <= 0:
            return None
        return torch.empty(nb, dtype=torch.uint8, device=self.device)

    def _packed_ws(self, wname, w, wmode, k, Cin, Cout, ws_query):
        """(workspace tensor or None, wmode flags) for a tensor-core convolution of parameter ``wname``.
        ``ws_query()`` returns the workspace size in bytes (0: no tensor-core kernel for this shape)."""
        if self.math == 0:
            return None, wmode
        nb = ws_query()
        if nb <= 0:
            return None, wmode
        if self.pack_cache is None or wname is None:
            return torch.empty(nb, dtype=torch.uint8, device=self.device), wmode
        key = (wname, wmode, k, Cin, Cout, self.math)
        ws = self.pack_cache.get(key)
        fresh = ws is None or ws.numel() < nb
        if fresh:
            ws = self.pack_cache[key] = torch.empty(nb, dtype=torch.uint8, device=self.device)
        if not fresh and self.prepacked is not None and key in self.prepacked:
            return ws, wmode | W_PREPACKED
        if key not in self._packed:
            self._packed.add(key)
            main = torch.cuda.current_stream()
            side = self.pack_stream
            if side is not None:
                if not self._pack_forked:
                    side.wait_stream(main)          # fork once: the packs only depend on theta
                    self._pack_forked = True
                with torch.cuda.stream(side):
                    self._call('dl4ds_conv2d_pack', w.data_ptr(), wmode, k, k, Cin, Cout, self.math, ws.data_ptr(),
                               side.cuda_stream)
                    ev = torch.cuda.Event()
                    ev.record(side)
                main.wait_event(ev)
            else:
                self._call('dl4ds_conv2d_pack', w.data_ptr(), wmode, k, k, Cin, Cout, self.math, ws.data_ptr(),
                           _stream())
        return ws, wmode | W_PREPACKED

    def join_pack_stream(self):
        """Join the side stream back into the current one (required before a stream capture ends)."""
        if self.pack_stream is not None and self._pack_forked:
            torch.cuda.current_stream().wait_stream(self.pack_stream)

    def _conv_raw(self, xptr, xld, wptr, bptr, resptr, resld, yptr, yld, N, H, W, Cin, Ho, Wo, Cout,
                  k, stride, up, pt, pl, wmode, act, r, beta):
        """dl4ds_conv2d_fwd on raw pointers (ConvLSTM time slices), workspace handled here."""
        ws = self._conv_ws(N, H, W, Cin, Ho, Wo, Cout, k, stride, up, r)
        self._call('dl4ds_conv2d_fwd', xptr, xld, wptr, bptr, resptr, resld, yptr, yld, N, H, W, Cin, Ho, Wo,
                   Cout, k, k, stride, up, pt, pl, wmode, act, r, beta, self.math,
                   ws.data_ptr() if ws is not None else None, _stream())

    def _p(self, name):
        self.names_used.append(name)
        self.first_use.setdefault(name, len(self.tape))     # tape position of the op that first reads the parameter
        return self.arena.param(name)

    def _g(self, name):
        return self.arena.gradient(name)

    def _record(self, fn):
        if self.training:
            self.tape.append(fn)

    def _use(self, *vs):
        """Register the calling op as a consumer of each Var (forward pass).  Every op does this for its
        differentiable inputs: a convolution may fuse the producer's epilogue-backward into its dgrad only when it is
        the LAST of ``n_uses`` contributors to the input's gradient (``_contrib`` raises if the counts disagree)."""
        for v in vs:
            if v is not None:
                v.n_uses += 1

    def _contrib(self, var):
        var.n_contrib += 1
        if var.premasked or (var.act_src is not None and var.n_contrib > var.n_uses):
            raise RuntimeError('gradient bookkeeping: a contribution arrived after the fused epilogue-backward was '
                               'applied (an op consumed the tensor without Ctx._use)')

    def _acc(self, var, writer):
        """Accumulate a gradient into ``var``.  writer(dst_var, beta) must write (beta=0) or
        add (beta=1) the gradient; ops that cannot accumulate are wrapped by ``_acc_via_tmp``."""
        if not var.requires_grad:
            return
        self._contrib(var)
        if var.grad is None:
            var.grad = var.like(pad=True)
            writer(var.grad, 0)
        else:
            writer(var.grad, 1)

    def _acc_via_tmp(self, var, writer):
        """writer(dst_var) writes the full gradient; summed into var.grad if one exists."""
        if not var.requires_grad:
            return
        self._contrib(var)
        if var.grad is None:
            var.grad = var.like()
            writer(var.grad)
        else:
            tmp = var.like()
            writer(tmp)
            self._copy(tmp, var.grad, accumulate=1)

    def _copy(self, src, dst, accumulate=0):
        assert src.C == dst.C and src.npix == dst.npix
        self._call('dl4ds_copy_channels', src.ptr, src.ld, dst.ptr, dst.ld, src.npix, src.C, accumulate,
                   _stream())

    def _dense(self, gvar):
        """Return ``gvar`` itself if it is a dense tensor, else a dense copy (ops whose kernels
        take no pitch use this on incoming gradients that are channel slices of a concat)."""
        if gvar.off == 0 and gvar.ld == gvar.C:
            return gvar
        d = gvar.like()
        self._copy(gvar, d)
        return d

    def _give_grad(self, var, gvar, adopt=True):
        """Hand ``gvar`` (a Var holding d loss / d var) to ``var``.  A gradient buffer has exactly one
        owner (backward closures scale and accumulate in place): ``adopt`` transfers ownership when
        ``var`` has no gradient yet; otherwise the values are copied / added.  Returns True if the
        buffer was adopted."""
        if not var.requires_grad:
            return False
        self._contrib(var)
        if var.grad is None:
            if adopt:
                var.grad = gvar
                return True
            var.grad = var.like()
            self._copy(gvar, var.grad, accumulate=0)
        else:
            self._copy(gvar, var.grad, accumulate=1)
        return False

    # ---------------------------------------------------------------- convolution family
    def conv(self, x, name, cout, k=3, act=None, bias=True, stride=1, padding='same', res=None,
             d2s=1, out=None, dense=False):
        """Keras Conv2D (+bias +activation [+residual add before the activation] [+depth_to_space])
        -- blocks.py:49-61,91,97,208,299,427; sp_postups.py:134,156."""
        if act == 'gelu':       # not expressible through its output: linear epilogue, then the standalone kernel
            return self.gelu(self.conv(x, name, cout, k=k, act=None, bias=bias, stride=stride, padding=padding,
                                       res=res, d2s=d2s, out=out, dense=dense))
        w = self._p(name + '/kernel')
        wshape = (x.C, cout) if dense else (k, k, x.C, cout)
        assert tuple(w.shape) == wshape, (name, tuple(w.shape), wshape)
        b = self._p(name + '/bias') if bias else None
        if padding == 'same':
            Ho, pt = same_pads(x.H, k, stride)
            Wo, pl = same_pads(x.W, k, stride)
        else:
            Ho, pt = (x.H - k) // stride + 1, 0
            Wo, pl = (x.W - k) // stride + 1, 0
        r = d2s if d2s > 1 else 1
        if out is None:
            out = new_var(x.N, Ho * r, Wo * r, cout // (r * r), self.device)
        else:
            assert (out.N, out.H, out.W, out.C) == (x.N, Ho * r, Wo * r, cout // (r * r))
        if res is not None:
            assert (res.N, res.H, res.W, res.C) == (x.N, Ho, Wo, cout)
        a = ACT[act]
        lib = _lib.load()
        self._use(x, res)
        if r == 1 and out.off == 0 and out.ld == out.C and (a != 0 or bias):
            out.act_src = (a, (name + '/bias') if bias else None)     # what a consumer's fused dgrad needs
        ws, wm = self._packed_ws(name + '/kernel', w, W_HWIO, k, x.C, cout,
                                 lambda: lib.dl4ds_conv2d_fwd_workspace_bytes(x.N, x.H, x.W, x.C, Ho, Wo, cout, k, k,
                                                                              stride, 1, r, self.math))
        self._timed('%s:fwd@%dx%d' % (name, x.H, x.W),
                    'dl4ds_conv2d_fwd', x.ptr, x.ld, w.data_ptr(), b.data_ptr() if bias else None,
                   res.ptr if res is not None else None, res.ld if res is not None else 0,
                   out.ptr, out.ld, x.N, x.H, x.W, x.C, Ho, Wo, cout, k, k, stride, 1, pt, pl,
                   wm, a, r, 0, self.math, ws.data_ptr() if ws is not None else None, _stream())

        def bwd():
            dy = out.grad
            if dy is None:
                return
            # epilogue backward: dz = dy * act'(y) (un-shuffled if d2s), dbias += sum dz
            pg = self.param_grads
            if out.premasked:
                dz = dy         # the last consumer's dgrad already stored d(pre-activation) and summed the bias gradient
            elif a == 0 and r == 1:
                dz = dy
                if bias and pg:
                    # a pure column sum that only feeds the gradient arena: like the weight gradient it leaves the
                    # dgrad chain for the side stream (same condition: dz is not handed on to a residual input)
                    import contextlib
                    with (self._on_side([dy.buf]) if res is None else contextlib.nullcontext()):
                        self._call('dl4ds_bias_act_bwd', dy.ptr, dy.ld, None, 0, None, 0,
                                   self._g(name + '/bias').data_ptr(), x.N, Ho, Wo, cout, 0, 1, _stream())
            else:
                dz = new_var(x.N, Ho, Wo, cout, self.device) if r > 1 else dy
                self._call('dl4ds_bias_act_bwd', dy.ptr, dy.ld, out.ptr, out.ld, dz.ptr, dz.ld,
                           self._g(name + '/bias').data_ptr() if (bias and pg) else None,
                           x.N, Ho, Wo, cout, a, r, _stream())
            # weight gradient (accumulating)
            if pg:
                self._wgrad(x, dz, self._g(name + '/kernel'), k, stride, pt, pl,
                            label='%s:wgrad@%dx%d' % (name, x.H, x.W), side=res is None)
            # input gradient
            if x.requires_grad:
                ws_q = lambda: lib.dl4ds_conv2d_fwd_workspace_bytes(x.N, Ho, Wo, cout, x.H, x.W, x.C, k, k, 1, stride, 1,
                                                                    self.math)
                # fused epilogue-backward of x's producer: this dgrad is the last of x's gradient contributions
                fuse = (x.act_src is not None and x.n_contrib == x.n_uses - 1 and stride == 1 and self.math != 0 and
                        (Ho, Wo) == (x.H, x.W) and x.off == 0 and x.ld == x.C and
                        dz.ptr % 16 == 0 and dz.ld % 4 == 0 and
                        (x.grad is None or (x.grad.ptr % 16 == 0 and x.grad.ld % 4 == 0)) and
                        lib.dl4ds_conv2d_dgrad_fused_supported(x.N, x.H, x.W, cout, x.C, k, k, self.math) == 1)
                if fuse:
                    ws2, wm2 = self._packed_ws(name + '/kernel', w, W_FLIP_T, k, cout, x.C, ws_q)
                    fuse = ws2 is not None or (x.C == 8 and cout == 8)      # (the 8-channel warp-level kernel packs nothing)
                if fuse:
                    pa, pbias = x.act_src
                    self._contrib(x)
                    beta = 1
                    if x.grad is None:
                        x.grad, beta = x.like(), 0
                    dbp = self._g(pbias).data_ptr() if (pbias is not None and pg) else None
                    self._timed('%s:dgrad@%dx%d' % (name, x.H, x.W),
                                'dl4ds_conv2d_dgrad_fused', dz.ptr, dz.ld, w.data_ptr(), x.grad.ptr, x.grad.ld,
                                x.ptr if pa != 0 else None, x.ld, pa, dbp, x.N, x.H, x.W, cout, x.C, k, k,
                                k - 1 - pt, k - 1 - pl, wm2, beta, self.math,
                                ws2.data_ptr() if ws2 is not None else None, _stream())
                    x.premasked = True
                else:
                    def wr(dst, beta):
                        ws2, wm2 = self._packed_ws(name + '/kernel', w, W_FLIP_T, k, cout, x.C, ws_q)
                        self._timed('%s:dgrad@%dx%d' % (name, x.H, x.W),
                                   'dl4ds_conv2d_fwd', dz.ptr, dz.ld, w.data_ptr(), None, None, 0,
                                   dst.ptr, dst.ld, x.N, Ho, Wo, cout, x.H, x.W, x.C, k, k, 1, stride,
                                   k - 1 - pt, k - 1 - pl, wm2, 0, 1, beta, self.math,
                                   ws2.data_ptr() if ws2 is not None else None, _stream())
                    self._acc(x, wr)
            if res is not None:     # d(res) = dz; handed over last (stream order keeps reads before
                self._give_grad(res, dz)   # any later in-place update by the new owner)
            out.grad = None
            out.n_contrib = 0
            out.premasked = False
        self._record(bwd)
        return out

    def conv_d2s_pointwise(self, x, name1, cm, name2, co, act=None, r=2, k=3):
        """The last x`r` stage of SubpixelConvolutionBlock -- Conv2D(r*r*cm, k, linear) + depth_to_space(r),
        blocks.py:421-427 -- COMPOSED with the 1x1 convolution (+bias +activation) that consumes it
        (TransitionLast, sp_postups.py:205 / blocks.py:299).  Both maps are linear, so the pair equals one
        k x k convolution to r*r*co channels + depth_to_space whose weights are W1 (x) W2
        (``dl4ds_spc_pointwise_compose``); the cm-channel HR tensor is never written.  The backward pass
        differentiates the composed layer and maps (dW_eff, db_eff) back onto the four original parameter
        gradients with the exact chain rule (``dl4ds_spc_pointwise_chain``).  Results equal the unfused
        graph up to fp32 re-association."""
        if act == 'gelu':
            return self.gelu(self.conv_d2s_pointwise(x, name1, cm, name2, co, act=None, r=r, k=k))
        self._use(x)
        w1, b1 = self._p(name1 + '/kernel'), self._p(name1 + '/bias')
        w2, b2 = self._p(name2 + '/kernel'), self._p(name2 + '/bias')
        R2 = r * r
        assert tuple(w1.shape) == (k, k, x.C, R2 * cm) and tuple(w2.shape) == (1, 1, cm, co), (w1.shape, w2.shape)
        dev = self.device
        rows, ce = k * k * x.C, R2 * co
        weff = torch.empty((k, k, x.C, ce), dtype=torch.float32, device=dev)
        beff = torch.empty(ce, dtype=torch.float32, device=dev)
        self._call('dl4ds_spc_pointwise_compose', w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                   weff.data_ptr(), beff.data_ptr(), rows, cm, co, r, _stream())
        Ho, pt = same_pads(x.H, k, 1)
        Wo, pl = same_pads(x.W, k, 1)
        out = new_var(x.N, Ho * r, Wo * r, co, dev)
        a = ACT[act]
        label = '%s*%s' % (name1, name2)
        ws = self._conv_ws(x.N, x.H, x.W, x.C, Ho, Wo, ce, k, 1, 1, r)
        self._timed('%s:fwd@%dx%d' % (label, x.H, x.W),
                    'dl4ds_conv2d_fwd', x.ptr, x.ld, weff.data_ptr(), beff.data_ptr(), None, 0, out.ptr, out.ld,
                    x.N, x.H, x.W, x.C, Ho, Wo, ce, k, k, 1, 1, pt, pl, W_HWIO, a, r, 0, self.math,
                    ws.data_ptr() if ws is not None else None, _stream())

        def bwd():
            dy = out.grad
            if dy is None:
                return
            pg = self.param_grads
            dz = new_var(x.N, Ho, Wo, ce, dev)
            dbeff = torch.zeros(ce, dtype=torch.float32, device=dev)
            self._call('dl4ds_bias_act_bwd', dy.ptr, dy.ld, out.ptr, out.ld, dz.ptr, dz.ld,
                       dbeff.data_ptr() if pg else None, x.N, Ho, Wo, ce, a, r, _stream())
            if pg:
                dweff = torch.zeros_like(weff)
                self.launches += 1

                def chain():
                    self._call('dl4ds_spc_pointwise_chain', w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                               dweff.data_ptr(), dbeff.data_ptr(), self._g(name1 + '/kernel').data_ptr(),
                               self._g(name1 + '/bias').data_ptr(), self._g(name2 + '/kernel').data_ptr(),
                               self._g(name2 + '/bias').data_ptr(), rows, cm, co, r, _stream())
                self._wgrad(x, dz, dweff, k, 1, pt, pl, label='%s:wgrad@%dx%d' % (label, x.H, x.W), side=True,
                            then=chain, keep=(dbeff, weff))
            if x.requires_grad:
                def wr(dst, beta):
                    ws2 = self._conv_ws(x.N, Ho, Wo, ce, x.H, x.W, x.C, k, 1, 1, 1)
                    self._timed('%s:dgrad@%dx%d' % (label, x.H, x.W),
                                'dl4ds_conv2d_fwd', dz.ptr, dz.ld, weff.data_ptr(), None, None, 0,
                                dst.ptr, dst.ld, x.N, Ho, Wo, ce, x.H, x.W, x.C, k, k, 1, 1,
                                k - 1 - pt, k - 1 - pl, W_FLIP_T, 0, 1, beta, self.math,
                                ws2.data_ptr() if ws2 is not None else None, _stream())
                self._acc(x, wr)
            out.grad = None
        self._record(bwd)
        return out

    def _on_side(self, keep):
        """Context manager: the weight-gradient side stream, forked after everything queued on the current stream so
        far; ``keep`` lists the tensors the side-stream kernels read or write (held until the join)."""
        import contextlib
        side = self.wgrad_stream
        if side is None or self.timers is not None:
            return contextlib.nullcontext()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        side.wait_event(ev)
        self._side_keep.extend(keep)
        self._side_used = True
        return torch.cuda.stream(side)

    def _join_side(self):
        if self._side_used:
            torch.cuda.current_stream().wait_stream(self.wgrad_stream)
            self._side_used = False
        self._side_keep = []

    def _wgrad(self, P, Q, dw, k, stride, pt, pl, label='wgrad', side=False, then=None, keep=()):
        """``side``: launch on the weight-gradient side stream when one is set (see __init__); ``then``: a callable
        issued right after the wgrad on the same stream (the chain-rule kernels of composed layers)."""
        if _SKIP_WGRAD:     # timing experiment only (DL4DS_SKIP_WGRAD=1): how long is the step without weight gradients
            return
        ws_bytes = _lib.load().dl4ds_conv2d_wgrad_workspace_bytes(P.N, Q.H, Q.W, P.C, Q.C, k, k, self.math)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device) if ws_bytes > 0 else None

        def go():
            self._timed(label, 'dl4ds_conv2d_wgrad', P.ptr, P.ld, Q.ptr, Q.ld, dw.data_ptr(),
                        P.N, P.H, P.W, P.C, Q.H, Q.W, Q.C, k, k, stride, pt, pl,
                        ws.data_ptr() if ws is not None else None, self.math, _stream())
            if then is not None:
                then()
        if side:
            with self._on_side([P.buf, Q.buf, ws, dw] + list(keep)):
                go()
        else:
            go()

    def conv_transpose(self, x, name, cout, k, stride, act=None):
        """Keras Conv2DTranspose(cout, k, strides=stride, padding='same', use_bias=False)
        -- blocks.py:508-516.  out = stride * in; kernel layout (kh,kw,Cout,Cin)."""
        if act == 'gelu':
            return self.gelu(self.conv_transpose(x, name, cout, k, stride, act=None))
        w = self._p(name + '/kernel')
        assert tuple(w.shape) == (k, k, cout, x.C), (name, tuple(w.shape))
        self._use(x)
        Ho, Wo = x.H * stride, x.W * stride
        _, pt = same_pads(Ho, k, stride)
        _, pl = same_pads(Wo, k, stride)
        if stride > 1 and pt == pl and os.environ.get('DL4DS_CONVT_DIRECT', '0') != '1':
            return self._conv_transpose_d2s(x, name, w, cout, k, stride, pt, act)
        out = new_var(x.N, Ho, Wo, cout, self.device)
        a = ACT[act]
        self._call('dl4ds_conv2d_fwd', x.ptr, x.ld, w.data_ptr(), None, None, 0, out.ptr, out.ld,
                   x.N, x.H, x.W, x.C, Ho, Wo, cout, k, k, 1, stride, k - 1 - pt, k - 1 - pl,
                   W_FLIP_T, a, 1, 0, self.math, None, _stream())

        def bwd():
            dy = out.grad
            if dy is None:
                return
            dz = dy
            if a != 0:
                self._call('dl4ds_bias_act_bwd', dy.ptr, dy.ld, out.ptr, out.ld, dz.ptr, dz.ld, None,
                           x.N, Ho, Wo, cout, a, 1, _stream())
            # dw[k][cout][cin] += sum dz[s*i + k - p][cout] * x[i][cin]
            if self.param_grads:
                self._wgrad(dz, x, self._g(name + '/kernel'), k, stride, pt, pl)
            if x.requires_grad:
                def wr(dst, beta):
                    self._call('dl4ds_conv2d_fwd', dz.ptr, dz.ld, w.data_ptr(), None, None, 0,
                               dst.ptr, dst.ld, x.N, Ho, Wo, cout, x.H, x.W, x.C, k, k, stride, 1,
                               pt, pl, W_HWIO, 0, 1, beta, self.math, None, _stream())
                self._acc(x, wr)
            out.grad = None
        self._record(bwd)
        return out

    def _conv_transpose_d2s(self, x, name, w, cout, k, s, pt, act):
        """Conv2DTranspose as a stride-1 Kp x Kp convolution to s*s*cout channels + depth_to_space(s)
        (``dl4ds_convt_rearrange``): every output phase (dy, dx) of a fractionally strided convolution is an
        ordinary convolution on the input grid, so the layer runs on the tensor-core kernels (forward with the
        fused depth_to_space store, dgrad, stacked-taps wgrad) instead of the CUDA-core fractional-stride path.
        The weight gradient of the rearranged image is scattered back onto the Keras-layout kernel (each element
        has exactly one image).  Same function; only the summation order differs."""
        dev = self.device
        pad = k - 1 - pt
        offs = sorted({(d + kh - pad) // s for d in range(s) for kh in range(k) if (d + kh - pad) % s == 0})
        off_min, Kp = offs[0], offs[-1] - offs[0] + 1
        ce = s * s * cout
        wp = torch.empty((Kp, Kp, x.C, ce), dtype=torch.float32, device=dev)
        self._call('dl4ds_convt_rearrange', w.data_ptr(), wp.data_ptr(), k, s, pad, off_min, Kp, cout, x.C, 0, _stream())
        out = new_var(x.N, x.H * s, x.W * s, cout, dev)
        a = ACT[act]
        pp = -off_min                                       # top / left padding of the equivalent convolution
        ws = self._conv_ws(x.N, x.H, x.W, x.C, x.H, x.W, ce, Kp, 1, 1, s)
        self._timed('%s:fwd@%dx%d' % (name, x.H, x.W),
                    'dl4ds_conv2d_fwd', x.ptr, x.ld, wp.data_ptr(), None, None, 0, out.ptr, out.ld,
                    x.N, x.H, x.W, x.C, x.H, x.W, ce, Kp, Kp, 1, 1, pp, pp, W_HWIO, a, s, 0, self.math,
                    ws.data_ptr() if ws is not None else None, _stream())

        def bwd():
            dy = out.grad
            if dy is None:
                return
            dz = new_var(x.N, x.H, x.W, ce, dev)
            self._call('dl4ds_bias_act_bwd', dy.ptr, dy.ld, out.ptr, out.ld, dz.ptr, dz.ld, None,
                       x.N, x.H, x.W, ce, a, s, _stream())
            if self.param_grads:
                dwp = torch.zeros_like(wp)

                def scatter():
                    self._call('dl4ds_convt_rearrange', self._g(name + '/kernel').data_ptr(), dwp.data_ptr(), k, s,
                               pad, off_min, Kp, cout, x.C, 1, _stream())
                self._wgrad(x, dz, dwp, Kp, 1, pp, pp, label='%s:wgrad@%dx%d' % (name, x.H, x.W), side=True,
                            then=scatter)
            if x.requires_grad:
                def wr(dst, beta):
                    ws2 = self._conv_ws(x.N, x.H, x.W, ce, x.H, x.W, x.C, Kp, 1, 1, 1)
                    self._timed('%s:dgrad@%dx%d' % (name, x.H, x.W),
                                'dl4ds_conv2d_fwd', dz.ptr, dz.ld, wp.data_ptr(), None, None, 0, dst.ptr, dst.ld,
                                x.N, x.H, x.W, ce, x.H, x.W, x.C, Kp, Kp, 1, 1, Kp - 1 - pp, Kp - 1 - pp,
                                W_FLIP_T, 0, 1, beta, self.math, ws2.data_ptr() if ws2 is not None else None, _stream())
                self._acc(x, wr)
            out.grad = None
        self._record(bwd)
        return out

    def dense(self, x, name, cout, act=None):
        """Keras Dense acts on the last axis: a 1x1 convolution with a (Cin, Cout) kernel -- on the pooled
        (B,1,1,Cin) vector in the discriminator (discriminator.py:78-79), on every pixel in ConvNextBlock
        (blocks.py:150,155)."""
        return self.conv(x, name, cout, k=1, act=act, dense=True)

    # ---------------------------------------------------------------- element-wise / structural
    def add(self, a, b, act=None):
        """Add (+ optional activation) -- sp_postups.py:164, blocks.py:228-229."""
        assert (a.N, a.H, a.W, a.C) == (b.N, b.H, b.W, b.C)
        if act == 'gelu':
            return self.gelu(self.add(a, b))
        self._use(a, b)
        out = a.like()
        code = ACT[act]
        self._call('dl4ds_add', a.ptr, a.ld, b.ptr, b.ld, out.ptr, out.ld, a.npix, a.C, code, _stream())

        def bwd():
            dy = out.grad
            if dy is None:
                return
            if code != 0:
                self._call('dl4ds_bias_act_bwd', dy.ptr, dy.ld, out.ptr, out.ld, dy.ptr, dy.ld, None,
                           a.N, a.H, a.W, a.C, code, 1, _stream())
            adopted = self._give_grad(a, dy)
            self._give_grad(b, dy, adopt=not adopted)
            out.grad = None
        self._record(bwd)
        return out

    def concat(self, parts):
        """Concatenate along channels -- blocks.py:276, sp_postups.py:186,201."""
        p0 = parts[0]
        self._use(*parts)
        ctot = sum(p.C for p in parts)
        out = Var(torch.empty((p0.N, p0.H, p0.W, padded_channels(ctot)), dtype=torch.float32, device=self.device), 0, ctot)
        offs = []
        o = 0
        for p in parts:
            assert (p.N, p.H, p.W) == (p0.N, p0.H, p0.W)
            self._copy(p, out.slice(o, p.C))
            offs.append(o)
            o += p.C

        def bwd():
            dy = out.grad
            if dy is None:
                return
            for p, o in zip(parts, offs):
                g = dy.slice(o, p.C)
                if (g.off % 4 or g.ld % 4) and p.C % 8 == 0 and p.C >= 16:
                    # a wide part that starts at an odd channel (48 channels behind 48 + 2 in cfg3's tail): every kernel
                    # downstream of a 16-byte-misaligned gradient is the generic CUDA-core one (r02c3b: 846 us weight
                    # gradient + 505 us dgrad against ~130 + 100 us on the tensor cores); one re-aligning copy instead
                    d = p.like()
                    self._copy(g, d)
                    g = d
                self._give_grad(p, g)
            out.grad = None
        self._record(bwd)
        return out

    def act(self, x, act):
        """Standalone activation -- blocks.py:391,397."""
        if act == 'gelu':
            return self.gelu(x)
        code = ACT[act]
        if code == 0:
            return x
        self._use(x)
        out = x.like()
        self._call('dl4ds_act_fwd', x.ptr, x.ld, out.ptr, out.ld, x.npix, x.C, code, _stream())

        def bwd():
            dy = out.grad
            if dy is None:
                return
            self._call('dl4ds_bias_act_bwd', dy.ptr, dy.ld, out.ptr, out.ld, dy.ptr, dy.ld, None,
                       x.N, x.H, x.W, x.C, code, 1, _stream())
            self._give_grad(x, dy)
            out.grad = None
        self._record(bwd)
        return out

    def norm(self, x, name, kind, act=None, eps=1e-3):
        """BatchNormalization() / LayerNormalization() followed by the activation, fused -- blocks.py:63-71 and
        their uses :92-101,216-224,263-272.  Keras defaults (axis -1, epsilon 1e-3, BN momentum 0.99).  BN in
        training mode normalises with the batch statistics and updates ``moving_mean`` / ``moving_variance`` (they
        live in the parameter arena with a gradient that stays zero, so Adam leaves them alone); in inference mode
        it uses the moving ones."""
        if kind not in ('bn', 'ln'):
            raise ValueError('Normalization not supported, got %s' % (kind,))           # blocks.py:64-65
        if act == 'gelu':
            return self.gelu(self.norm(x, name, kind, act=None, eps=eps))
        self._use(x)
        code, C, n_pix = ACT[act], x.C, x.npix
        gamma, beta = self._p(name + '/gamma'), self._p(name + '/beta')
        out = x.like()
        if kind == 'ln':
            self._call('dl4ds_layernorm_fwd', x.ptr, x.ld, gamma.data_ptr(), beta.data_ptr(), float(eps), out.ptr,
                       out.ld, n_pix, C, code, _stream())
            stats = None
        else:
            mm, mv = self._p(name + '/moving_mean'), self._p(name + '/moving_variance')
            if self.training:
                stats = torch.empty(4 * C, dtype=torch.float32, device=self.device)   # mean | var | scratch (2C)
                mean_p, var_p = stats.data_ptr(), stats.data_ptr() + 4 * C
                self._call('dl4ds_batchnorm_stats', x.ptr, x.ld, n_pix, C, mean_p, var_p, mm.data_ptr(),
                           mv.data_ptr(), 0.99, stats.data_ptr() + 8 * C, _stream())
                self.launches += 3
            else:
                stats = None
                mean_p, var_p = mm.data_ptr(), mv.data_ptr()
            self._call('dl4ds_norm_apply', x.ptr, x.ld, mean_p, var_p, gamma.data_ptr(), beta.data_ptr(), float(eps),
                       out.ptr, out.ld, n_pix, C, code, _stream())

        def bwd():
            dy = out.grad
            if dy is None:
                return
            pg = self.param_grads
            dg = self._g(name + '/gamma').data_ptr() if pg else None
            db = self._g(name + '/beta').data_ptr() if pg else None
            if kind == 'ln':
                def wr(dst):
                    self._call('dl4ds_layernorm_bwd', dy.ptr, dy.ld, x.ptr, x.ld, out.ptr, out.ld, gamma.data_ptr(),
                               float(eps), dst.ptr if dst is not None else None, dst.ld if dst is not None else 0,
                               dg, db, n_pix, C, code, _stream())
                if x.requires_grad:
                    self._acc_via_tmp(x, wr)
                elif pg:
                    wr(None)
            else:
                def wr(dst):
                    self._call('dl4ds_batchnorm_bwd', dy.ptr, dy.ld, x.ptr, x.ld, out.ptr, out.ld, stats.data_ptr(),
                               stats.data_ptr() + 4 * C, gamma.data_ptr(), float(eps), dst.ptr, dst.ld, dg, db,
                               stats.data_ptr() + 8 * C, n_pix, C, code, _stream())
                    self.launches += 2
                if x.requires_grad:
                    self._acc_via_tmp(x, wr)
                elif pg:
                    wr(x.like())
            out.grad = None
        self._record(bwd)
        return out

    def depthwise_conv(self, x, name, k=7):
        """DepthwiseConv2D(kernel_size=k, padding='same', depth_multiplier=1) with bias -- blocks.py:147-148."""
        self._use(x)
        wt, bias = self._p(name + '/depthwise_kernel'), self._p(name + '/bias')
        out = x.like()
        self._call('dl4ds_depthwise_conv_fwd', x.ptr, x.ld, wt.data_ptr(), bias.data_ptr(), out.ptr, out.ld,
                   x.N, x.H, x.W, x.C, k, 0, 0, _stream())

        def bwd():
            dy = out.grad
            if dy is None:
                return
            if self.param_grads:
                self._call('dl4ds_bias_act_bwd', dy.ptr, dy.ld, None, 0, None, 0, self._g(name + '/bias').data_ptr(),
                           x.N, x.H, x.W, x.C, 0, 1, _stream())
                self._call('dl4ds_depthwise_conv_wgrad', x.ptr, x.ld, dy.ptr, dy.ld,
                           self._g(name + '/depthwise_kernel').data_ptr(), x.N, x.H, x.W, x.C, k, _stream())
            if x.requires_grad:
                def wr(dst, beta):
                    self._call('dl4ds_depthwise_conv_fwd', dy.ptr, dy.ld, wt.data_ptr(), None, dst.ptr, dst.ld,
                               x.N, x.H, x.W, x.C, k, 1, beta, _stream())
                self._acc(x, wr)
            out.grad = None
        self._record(bwd)
        return out

    def gelu(self, x):
        """Activation('gelu'), exact erf form -- ConvNextBlock's default activation, blocks.py:143,153."""
        self._use(x)
        xd = self._dense(x)
        out = xd.like()
        n = xd.npix * xd.C
        self._call('dl4ds_gelu_fwd', xd.ptr, out.ptr, n, _stream())

        def bwd():
            dy = out.grad
            if dy is None:
                return
            dy = self._dense(dy)

            def wr(dst):
                self._call('dl4ds_gelu_bwd', xd.ptr, dy.ptr, dst.ptr, n, _stream())
            self._acc_via_tmp(x, wr)
            out.grad = None
        self._record(bwd)
        return out

    def channel_scale(self, x, name, init_value):
        """ConvNextBlock's layer scale (blocks.py:166-169,178-179): y = gamma * x with a trainable per-channel ``gamma``
        (initialised to ``layer_scale_init_value``; parameter ``<name>`` of shape (C,))."""
        self._use(x)
        gamma = self._p(name)
        assert tuple(gamma.shape) == (x.C,)
        out = x.like()
        self._call('dl4ds_channel_scale_fwd', x.ptr, x.ld, gamma.data_ptr(), out.ptr, out.ld, x.npix, x.C, _stream())

        def bwd():
            dy = out.grad
            if dy is None:
                return
            dg = self._g(name).data_ptr() if self.param_grads else None

            def wr(dst):
                self._call('dl4ds_channel_scale_bwd', dy.ptr, dy.ld, x.ptr, x.ld, gamma.data_ptr(),
                           dst.ptr if dst is not None else None, dst.ld if dst is not None else 0, dg, x.npix, x.C,
                           _stream())
            if x.requires_grad:
                self._acc_via_tmp(x, wr)
            elif dg is not None:
                wr(None)
            out.grad = None
        self._record(bwd)
        return out

    def dropout(self, x, rate, variant=None, n_samples=None):
        """get_dropout_layer(rate, variant)(x) -- blocks.py:680-706: Dropout / GaussianDropout / SpatialDropout2D,
        active in training mode; the 'mc*' variants also in inference (blocks.py:662-677).  Masks: dl4ds_dropout
        (Philox, seed and step in device memory); the backward pass regenerates the same mask.  ``n_samples``: batch
        size B when ``x`` holds time-major frames (T*B,H,W,C), so that the spatial variant draws once per sample and
        channel over (T,H,W) like SpatialDropout3D (``dim=3``, blocks.py:692-693,701-702)."""
        if variant is not None and variant not in DROPOUT_KIND:
            raise ValueError('`dropout_variant` must be None or one of %s, got %s' % (sorted(k for k in DROPOUT_KIND if k), variant))
        if not rate or rate <= 0:
            return x
        if not (self.training or (variant or '').startswith('mc')):
            return x
        self._use(x)
        state = self.arena.rng_state()
        if not self._rng_advanced:
            self._call('dl4ds_rng_advance', state.data_ptr(), _stream())
            self._rng_advanced = True
        self._dropout_calls += 1
        lid, kind = self._dropout_calls, DROPOUT_KIND[variant]
        ns = int(n_samples) if n_samples else x.N
        out = x.like()
        self._call('dl4ds_dropout', x.ptr, x.ld, out.ptr, out.ld, x.npix, x.H * x.W, ns, x.C, float(rate), kind,
                   state.data_ptr(), lid, _stream())

        def bwd():
            dy = out.grad
            if dy is None:
                return

            def wr(dst):
                self._call('dl4ds_dropout', dy.ptr, dy.ld, dst.ptr, dst.ld, x.npix, x.H * x.W, ns, x.C, float(rate),
                           kind, state.data_ptr(), lid, _stream())
            self._acc_via_tmp(x, wr)
            out.grad = None
        self._record(bwd)
        return out

    def channel_attention(self, x, name, r=4, groups=None):
        """ChannelAttention2D -- blocks.py:537-593.  ``groups`` = (n_groups, pix_per_group, inner)
        overrides the default per-image pooling (used for the 5-D (T,H) pooling quirk)."""
        self._use(x)
        C = x.C
        Cr = int(C / r)
        w1, b1 = self._p(name + '/conv1/kernel'), self._p(name + '/conv1/bias')
        w2, b2 = self._p(name + '/conv2/kernel'), self._p(name + '/conv2/bias')
        ng, ppg, inner = groups if groups is not None else (x.N, x.H * x.W, 1)
        assert ng * ppg == x.npix
        dev = self.device
        pooled = torch.empty(ng * C, dtype=torch.float32, device=dev)
        hidden = torch.empty(ng * Cr, dtype=torch.float32, device=dev)
        scale = torch.empty(ng * C, dtype=torch.float32, device=dev)
        out = x.like()
        self.launches += 3
        self._call('dl4ds_channel_attention_fwd', x.ptr, x.ld, out.ptr, out.ld, w1.data_ptr(),
                   b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), pooled.data_ptr(), hidden.data_ptr(),
                   scale.data_ptr(), ng, ppg, inner, C, Cr, _stream())

        def bwd():
            dy = out.grad
            if dy is None:
                return
            dsum = torch.empty(ng * C, dtype=torch.float32, device=dev)
            if self.param_grads:
                gp = [self._g(name + s) for s in ('/conv1/kernel', '/conv1/bias', '/conv2/kernel', '/conv2/bias')]
            else:       # parameter gradients go to scratch
                gp = [torch.zeros(n_, dtype=torch.float32, device=dev) for n_ in (C * Cr, Cr, Cr * C, C)]

            def wr(dst):
                self.launches += 3
                self._call('dl4ds_channel_attention_bwd', x.ptr, x.ld, dy.ptr, dy.ld, dst.ptr, dst.ld,
                           w1.data_ptr(), w2.data_ptr(), pooled.data_ptr(), hidden.data_ptr(),
                           scale.data_ptr(), dsum.data_ptr(),
                           gp[0].data_ptr(), gp[1].data_ptr(), gp[2].data_ptr(), gp[3].data_ptr(),
                           ng, ppg, inner, C, Cr, _stream())
            self._acc_via_tmp(x, wr)
            out.grad = None
        self._record(bwd)
        return out

    def local_conv(self, x, name, filters):
        """LocallyConnected2D(filters, (1,1), implementation=3) -- blocks.py:322-328."""
        self._use(x)
        w, b = self._p(name + '/kernel'), self._p(name + '/bias')
        assert tuple(w.shape) == (x.H, x.W, x.C, filters)
        out = x.like(filters)
        self._call('dl4ds_local_conv1x1_fwd', x.ptr, x.ld, w.data_ptr(), b.data_ptr(), out.ptr, out.ld,
                   x.N, x.H, x.W, x.C, filters, _stream())

        def bwd():
            dy = out.grad
            if dy is None:
                return

            if self.param_grads:
                gk, gb = self._g(name + '/kernel'), self._g(name + '/bias')
            else:
                gk, gb = torch.zeros_like(w), torch.zeros_like(b)

            def wr(dst):
                self._call('dl4ds_local_conv1x1_bwd', x.ptr, x.ld, dy.ptr, dy.ld, w.data_ptr(),
                           dst.ptr, dst.ld, gk.data_ptr(), gb.data_ptr(), x.N, x.H, x.W, x.C, filters,
                           _stream())
            self._acc_via_tmp(x, wr)
            out.grad = None
        self._record(bwd)
        return out

    def resize_bilinear(self, x, Ho, Wo):
        """keras Resizing(Ho, Wo, 'bilinear') -- blocks.py:489, discriminator.py:62."""
        self._use(x)
        out = new_var(x.N, Ho, Wo, x.C, self.device)
        self._call('dl4ds_resize_bilinear_fwd', x.ptr, x.ld, out.ptr, out.ld, x.N, x.H, x.W, x.C, Ho, Wo,
                   _stream())

        def bwd():
            dy = out.grad
            if dy is None or not x.requires_grad:
                return
            self._contrib(x)
            if x.grad is None:
                x.grad = Var(torch.zeros((x.N, x.H, x.W, x.C), dtype=torch.float32, device=self.device))
            self._call('dl4ds_resize_bilinear_bwd', dy.ptr, dy.ld, x.grad.ptr, x.grad.ld, x.N, x.H, x.W,
                       x.C, Ho, Wo, _stream())
            out.grad = None
        self._record(bwd)
        return out

    def resize(self, x, Ho, Wo, method='bilinear'):
        """keras Resizing(Ho, Wo, interpolation=method) -- blocks.py:457-491 (`rc_interpolation`): bilinear, nearest,
        bicubic (tf.image.resize without antialiasing, half-pixel centres)."""
        if method == 'bilinear':
            return self.resize_bilinear(x, Ho, Wo)
        from .resize_tables import TAP_METHODS
        if method in TAP_METHODS:
            return self._resize_taps(x, Ho, Wo, method)
        if method not in RESIZE_METHOD:
            raise NotImplementedError('interpolation=%r is not one of keras Resizing\'s methods' % (method,))
        code = RESIZE_METHOD[method]
        self._use(x)
        out = new_var(x.N, Ho, Wo, x.C, self.device)
        self._call('dl4ds_resize_fwd', x.ptr, x.ld, out.ptr, out.ld, x.N, x.H, x.W, x.C, Ho, Wo, code, _stream())

        def bwd():
            dy = out.grad
            if dy is None or not x.requires_grad:
                return
            self._contrib(x)
            if x.grad is None:
                x.grad = Var(torch.zeros((x.N, x.H, x.W, x.C), dtype=torch.float32, device=self.device))
            self._call('dl4ds_resize_bwd', dy.ptr, dy.ld, x.grad.ptr, x.grad.ld, x.N, x.H, x.W, x.C, Ho, Wo, code,
                       _stream())
            out.grad = None
        self._record(bwd)
        return out

    def _tap_tables(self, n_in, n_out, method, transposed):
        """Device (indices, weights, K) of tf.image.resize's 1-D operator n_in -> n_out (or of its transpose),
        cached on the arena."""
        from .resize_tables import matrix_to_taps, tf_resize_matrix
        cache = self.arena.__dict__.setdefault('_resize_taps', {})
        key = (n_in, n_out, method, transposed)
        if key not in cache:
            r = tf_resize_matrix(n_in, n_out, method)
            idx, w = matrix_to_taps(r.T.copy() if transposed else r)
            cache[key] = (torch.as_tensor(idx).to(self.device), torch.as_tensor(w).to(self.device), idx.shape[1])
        return cache[key]

    def _resize_taps(self, x, Ho, Wo, method):
        """keras Resizing with interpolation in area / lanczos3 / lanczos5 / gaussian / mitchellcubic
        (blocks.py:457-491): a separable linear map, run as tap tables by ``dl4ds_resample_taps``; the backward pass
        is the same kernel with the transposed tables."""
        self._use(x)
        xd = self._dense(x)
        out = new_var(x.N, Ho, Wo, x.C, self.device)
        iy, wy, ky = self._tap_tables(x.H, Ho, method, False)
        ix, wx, kx = self._tap_tables(x.W, Wo, method, False)
        self._call('dl4ds_resample_taps', xd.ptr, out.ptr, x.N, x.H, x.W, x.C, Ho, Wo, iy.data_ptr(), wy.data_ptr(), ky,
                   ix.data_ptr(), wx.data_ptr(), kx, out.ld, 0, _stream())

        def bwd():
            dy = out.grad
            if dy is None or not x.requires_grad:
                return
            dyd = self._dense(dy)
            ty, tw, tk = self._tap_tables(x.H, Ho, method, True)
            sx, sw, sk = self._tap_tables(x.W, Wo, method, True)

            def wr(dst):
                self._call('dl4ds_resample_taps', dyd.ptr, dst.ptr, x.N, Ho, Wo, x.C, x.H, x.W, ty.data_ptr(),
                           tw.data_ptr(), tk, sx.data_ptr(), sw.data_ptr(), sk, dst.ld, 0, _stream())
            self._acc_via_tmp(x, wr)
            out.grad = None
        self._record(bwd)
        return out

    def maxpool2(self, x):
        """MaxPooling2D((2,2)) -- blocks.py:613."""
        self._use(x)
        out = new_var(x.N, x.H // 2, x.W // 2, x.C, self.device)
        self._call('dl4ds_maxpool2_fwd', x.ptr, x.ld, out.ptr, out.ld, x.N, x.H, x.W, x.C, _stream())

        def bwd():
            dy = out.grad
            if dy is None:
                return

            def wr(dst):
                self._call('dl4ds_maxpool2_bwd', x.ptr, x.ld, dy.ptr, dy.ld, dst.ptr, dst.ld, x.N, x.H,
                           x.W, x.C, _stream())
            self._acc_via_tmp(x, wr)
            out.grad = None
        self._record(bwd)
        return out

    def pad_to(self, x, H, W):
        """ZeroPadding2D bottom/right to (H, W) -- PadConcat, blocks.py:639-655."""
        if (x.H, x.W) == (H, W):
            return x
        self._use(x)
        out = new_var(x.N, H, W, x.C, self.device)
        self._call('dl4ds_pad_bottom_right', x.ptr, x.ld, out.ptr, out.ld, x.N, x.H, x.W, H, W, x.C,
                   _stream())

        def bwd():
            dy = out.grad
            if dy is None:
                return

            def wr(dst):   # adjoint = crop
                self._call('dl4ds_pad_bottom_right', dy.ptr, dy.ld, dst.ptr, dst.ld, x.N, H, W, x.H, x.W,
                           x.C, _stream())
            self._acc_via_tmp(x, wr)
            out.grad = None
        self._record(bwd)
        return out

    def permute_frames(self, x, A, B):
        """(A,B,frame) -> (B,A,frame) over the leading (frame) dimension of a (A*B,H,W,C) Var."""
        assert x.N == A * B and x.ld == x.C
        self._use(x)
        out = x.like()
        fe = x.H * x.W * x.C
        self._call('dl4ds_permute_frames', x.ptr, out.ptr, A, B, fe, _stream())

        def bwd():
            dy = out.grad
            if dy is None:
                return
            dy = self._dense(dy)

            def wr(dst):
                self._call('dl4ds_permute_frames', dy.ptr, dst.ptr, B, A, fe, _stream())
            self._acc_via_tmp(x, wr)
            out.grad = None
        self._record(bwd)
        return out

    def group_mean(self, x, n_groups=None):
        """GlobalAveragePooling2D -- discriminator.py:76.  (N,H,W,C) -> (N,1,1,C).  ``n_groups`` = B on BATCH-major
        frames (B*T,H,W,C) pools every sample's T consecutive frames together: GlobalAveragePooling3D over (T,H,W),
        discriminator.py:74 -> (B,1,1,C)."""
        self._use(x)
        ng = x.N if n_groups is None else int(n_groups)
        assert x.N % ng == 0
        ppg = (x.N // ng) * x.H * x.W
        out = new_var(ng, 1, 1, x.C, self.device)
        self.launches += 1
        self._call('dl4ds_group_mean_fwd', x.ptr, x.ld, out.ptr, ng, ppg, x.C, _stream())

        def bwd():
            dy = out.grad
            if dy is None:
                return
            dy = self._dense(dy)

            def wr(dst):
                self._call('dl4ds_group_mean_bwd', dy.ptr, dst.ptr, dst.ld, ng, ppg, x.C, _stream())
            self._acc_via_tmp(x, wr)
            out.grad = None
        self._record(bwd)
        return out

    def repeat_frames(self, x, T):
        """tf.repeat(tf.expand_dims(s, 1), T, axis=1) (spt_postups.py:139-140) in the time-major
        frame layout: (B,H,W,C) -> (T*B,H,W,C), T stacked copies."""
        B_ = x.N
        for _ in range(T):
            self._use(x)            # T gradient contributions come back
        out = new_var(T * B_, x.H, x.W, x.C, self.device)
        for t in range(T):
            self._copy(x, Var(out.buf[t * B_:(t + 1) * B_]))

        def bwd():
            dy = out.grad
            if dy is None:
                return
            for t in range(T):
                self._give_grad(x, Var(dy.buf[t * B_:(t + 1) * B_], dy.off, dy.C), adopt=False)
            out.grad = None
        self._record(bwd)
        return out

    def mul_mask(self, x, mask):
        """x * mask (inverted-dropout keep mask, discriminator.py:77).  mask: Var/tensor (N,1,1,C)."""
        if isinstance(mask, Var):
            assert mask.off == 0 and mask.ld == mask.C
            mask = mask.buf
        assert x.ld == x.C and mask.numel() == x.npix * x.C
        self._use(x)
        out = x.like()
        self._call('dl4ds_mul', x.ptr, mask.data_ptr(), out.ptr, x.npix * x.C, _stream())

        def bwd():
            dy = out.grad
            if dy is None:
                return
            dy = self._dense(dy)

            def wr(dst):
                self._call('dl4ds_mul', dy.ptr, mask.data_ptr(), dst.ptr, x.npix * x.C, _stream())
            self._acc_via_tmp(x, wr)
            out.grad = None
        self._record(bwd)
        return out

    def convlstm(self, x, name, filters, k, T):
        """ConvLSTM2D(filters, k, return_sequences=True, padding='same') -- blocks.py:350-355
        (Keras 2.x: tanh / hard_sigmoid, gates i,f,c,o, zero initial state, recurrent conv 'same'
        without bias).  ``x``: TIME-MAJOR frames (T*B, H, W, C); returns (T*B, H, W, filters)."""
        self._use(x)
        TB, H, W, C = x.N, x.H, x.W, x.C
        B = TB // T
        F4 = 4 * filters
        wx, wh, bias = self._p(name + '/kernel'), self._p(name + '/recurrent_kernel'), self._p(name + '/bias')
        assert tuple(wx.shape) == (k, k, C, F4) and tuple(wh.shape) == (k, k, filters, F4)
        dev = self.device
        pad = k // 2
        npx = B * H * W
        # input convolution for all T at once (one GEMM): z (T*B, H, W, 4F)
        z = new_var(TB, H, W, F4, dev)
        self._conv_raw(x.ptr, x.ld, wx.data_ptr(), bias.data_ptr(), None, 0, z.ptr, z.ld,
                       TB, H, W, C, H, W, F4, k, 1, 1, pad, pad, W_HWIO, 0, 1, 0)
        out = new_var(TB, H, W, filters, dev)
        cs = torch.empty((TB, H, W, filters), dtype=torch.float32, device=dev)
        gates = torch.empty((TB, H, W, F4), dtype=torch.float32, device=dev)

        def step(buf, t):
            return buf[t * B:(t + 1) * B]

        for t in range(T):
            zt = step(z.buf, t)
            if t > 0:   # z_t += conv(h_{t-1}, Wh)
                self._conv_raw(step(out.buf, t - 1).data_ptr(), filters, wh.data_ptr(), None,
                               None, 0, zt.data_ptr(), F4, B, H, W, filters, H, W, F4, k, 1, 1, pad, pad,
                               W_HWIO, 0, 1, 1)
            self._call('dl4ds_convlstm_gates_fwd', zt.data_ptr(), step(cs, t - 1).data_ptr() if t > 0 else None,
                       step(cs, t).data_ptr(), step(out.buf, t).data_ptr(), filters,
                       step(gates, t).data_ptr(), npx, filters, _stream())

        def bwd():
            dy = out.grad
            if dy is None:
                return
            pg = self.param_grads
            if not pg and not x.requires_grad:      # cGAN generator pass through D: nothing to compute here
                out.grad = None
                return
            dy = self._dense(dy)
            dh = dy.buf      # owned; dh_{t-1} is accumulated in place
            dz = new_var(TB, H, W, F4, dev)
            dc = [torch.empty((B, H, W, filters), dtype=torch.float32, device=dev) for _ in range(2)]
            gwh = self._g(name + '/recurrent_kernel')
            for t in range(T - 1, -1, -1):
                dzt = step(dz.buf, t)
                self._call('dl4ds_convlstm_gates_bwd', step(gates, t).data_ptr(),
                           step(cs, t - 1).data_ptr() if t > 0 else None, step(cs, t).data_ptr(),
                           step(dh, t).data_ptr(), filters,
                           dc[(t + 1) % 2].data_ptr() if t < T - 1 else None,
                           dzt.data_ptr(), dc[t % 2].data_ptr(), npx, filters, _stream())
                if t > 0:
                    # dh_{t-1} += dgrad(dz_t, Wh);  dWh += wgrad(h_{t-1}, dz_t)
                    self._conv_raw(dzt.data_ptr(), F4, wh.data_ptr(), None, None, 0,
                                   step(dh, t - 1).data_ptr(), filters, B, H, W, F4, H, W, filters, k, 1, 1,
                                   k - 1 - pad, k - 1 - pad, W_FLIP_T, 0, 1, 1)
                    if pg:      # (side stream: dz_t and h_{t-1} have no later writer)
                        self._wgrad(Var(step(out.buf, t - 1)), Var(dzt), gwh, k, 1, pad, pad, side=True,
                                    keep=[out.buf, dz.buf])
            # the input convolution's bias / weight / input gradients, all T in one shot
            if pg:
                self._call('dl4ds_bias_act_bwd', dz.ptr, dz.ld, None, 0, None, 0,
                           self._g(name + '/bias').data_ptr(), TB, H, W, F4, 0, 1, _stream())
                self._wgrad(x, dz, self._g(name + '/kernel'), k, 1, pad, pad, side=True)
            if x.requires_grad:
                def wr(dst, beta):
                    self._conv_raw(dz.ptr, dz.ld, wx.data_ptr(), None, None, 0, dst.ptr,
                                   dst.ld, TB, H, W, F4, H, W, C, k, 1, 1, k - 1 - pad, k - 1 - pad,
                                   W_FLIP_T, 0, 1, beta)
                self._acc(x, wr)
            out.grad = None
        self._record(bwd)
        return out

    # ---------------------------------------------------------------- losses
    def pixel_loss(self, y_pred, y_true, kind='mae', scale=1.0, loss_buf=None):
        """losses.mae / losses.mse (losses.py:5-20): loss_buf[0] += scale*mean; seeds y_pred.grad."""
        return self.loss(y_pred, y_true, kind, scale=scale, loss_buf=loss_buf)

    def loss(self, y_pred, y_true, name='mae', scale=1.0, loss_buf=None):
        """Any of LOSS_FUNCTIONS (losses.py:5-151, looked up by utils.checkarg_loss, utils.py:139-171):
        loss_buf[0] += scale * loss(y_true, y_pred); in training mode seeds y_pred.grad with scale * d loss / d y_pred.
        The weighted mixes (losses.py:62-93,134-151) run their pixel terms first, then the SSIM term accumulates
        into the same gradient buffer."""
        assert y_pred.ld == y_pred.C and y_true.ld == y_true.C
        self._use(y_pred)
        if name not in LOSS_TERMS:
            raise ValueError('unknown loss %r' % (name,))
        n = y_pred.npix * y_pred.C
        if loss_buf is None:
            loss_buf = torch.zeros(1, dtype=torch.float32, device=self.device)
        dy = y_pred.like() if self.training else None
        dyp = dy.ptr if dy is not None else None
        acc = 0
        for term, weight in LOSS_TERMS[name]:
            if term in ('mae', 'mse'):
                self._call('dl4ds_pixel_loss', y_pred.ptr, y_true.ptr, loss_buf.data_ptr(), dyp, n,
                           {'mae': 0, 'mse': 1}[term] | (LOSS_ACCUMULATE if acc else 0), float(scale * weight),
                           _stream())
            else:
                n_scales = 1 if term == 'dssim' else len(MSSSIM_POWER_FACTORS)
                nws = _lib.load().dl4ds_ssim_loss_workspace_floats(y_pred.N, y_pred.H, y_pred.W, y_pred.C, n_scales)
                if nws < 0:
                    raise _lib.Dl4dsError('ssim_loss: %s' % _lib.last_error())
                ws = torch.empty(nws, dtype=torch.float32, device=self.device)
                # several kernels behind one entry point: range (2), per scale maps (+2 poolings), combine,
                # per scale backward, fixup
                self.launches += 2 + n_scales + 2 * (n_scales - 1) + 1 + ((n_scales + 1) if dy is not None else 0) - 1
                self._call('dl4ds_ssim_loss', y_pred.ptr, y_true.ptr, y_pred.N, y_pred.H, y_pred.W, y_pred.C,
                           n_scales, _PF_HOST, float(scale * weight), loss_buf.data_ptr(), dyp, acc, ws.data_ptr(),
                           _stream())
            acc = 1
        if dy is not None:
            self._give_grad(y_pred, dy)
        return loss_buf

    def bce_loss(self, p, target, scale=1.0, loss_buf=None, seed=True):
        """BinaryCrossentropy(from_logits=False) vs a constant target -- cgan.py:546-552,567-571."""
        assert p.ld == p.C
        n = p.npix * p.C
        if loss_buf is None:
            loss_buf = torch.zeros(1, dtype=torch.float32, device=self.device)
        dp = None
        acc = 0
        if self.training and seed:
            if p.grad is None:
                p.grad = p.like()
            else:
                acc = 1
            dp = p.grad
        self._call('dl4ds_bce_loss', p.ptr, float(target), loss_buf.data_ptr(),
                   dp.ptr if dp is not None else None, n, float(scale), acc, _stream())
        return loss_buf

    # ---------------------------------------------------------------- backward
    def backward(self, keep_tape=False):
        """Run the recorded backward closures.  ``keep_tape`` allows a second pass with other seeds
        (cGAN: D(fake) is differentiated once for the discriminator weights, once for the generator)."""
        for i in range(len(self.tape) - 1, -1, -1):
            self.tape[i]()
            hook = self.bwd_hooks.get(i)
            if hook is not None:
                hook()
        self._join_side()
        if not keep_tape:
            self.tape = []


def adam_step(arena, lr, beta_1=0.9, beta_2=0.999, eps=1e-7, grad_scale=1.0, lr_t_dev=None):
    """tf.keras Adam on the whole arena (supervised.py:353, cgan.py:277-278)."""
    arena.t += 1
    if lr_t_dev is not None:
        call('dl4ds_adam_step_dev', arena.theta.data_ptr(), arena.grad.data_ptr(), arena.m.data_ptr(),
             arena.v.data_ptr(), arena.n, lr_t_dev.data_ptr(), float(beta_1), float(beta_2), float(eps),
             float(grad_scale), _stream())
    else:
        call('dl4ds_adam_step', arena.theta.data_ptr(), arena.grad.data_ptr(), arena.m.data_ptr(),
             arena.v.data_ptr(), arena.n, float(lr), float(beta_1), float(beta_2), float(eps), arena.t,
             float(grad_scale), _stream())
