"""Device-side training / evaluation steps: what Keras' ``train_function`` + Horovod's
``DistributedOptimizer`` do per batch in the reference (supervised.py:363-369,396-406), as ONE
replayable CUDA graph of hand-written kernels:

    zero grad arena -> forward -> pixel loss (+ gradient seed) -> backward -> [all-reduce] -> Adam

Static shapes (fixed per-rank batch) make the whole step capturable; the host only refreshes the
input buffers and the scalar ``lr_t`` (Adam bias correction x PiecewiseConstantDecay) between
replays.  Data parallelism shards the batch only: every rank runs the same graph on its own
samples; the flat gradient arena is summed over ranks with a single NCCL all-reduce and the
1/world_size of Horovod's average is folded into the Adam kernel's ``grad_scale``.
"""
import math

import torch

from . import _lib
from .engine import Ctx, Var, _stream, adam_step


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def allreduce_sum_(flat, dist=None):
    """In-place sum of a flat gradient buffer over the data-parallel group (the exchange step of
    hvd.DistributedOptimizer / DistributedGradientTape, supervised.py:365, cgan.py:610-611).  The
    average's 1/world is NOT applied here: it is folded into the Adam kernel's ``grad_scale``
    (``world_grad_scale``).  NCCL for CUDA buffers, gloo for the CPU tests."""
    d = dist if dist is not None else _dist()
    if d is not None and d.get_world_size() > 1:
        from . import comm
        if flat.is_cuda and comm.ensure(d):
            comm.allreduce_sum_(flat)              # dl4ds_comm_allreduce_sum on the current stream (capturable)
        else:
            d.all_reduce(flat, op=d.ReduceOp.SUM)
    return flat


def world_grad_scale(dist=None):
    d = dist if dist is not None else _dist()
    return 1.0 / d.get_world_size() if d is not None else 1.0


class LRSchedule:
    """float, or PiecewiseConstantDecay([boundary], [v0, v1]) (supervised.py:336-353): v0 while the
    optimizer iteration count <= boundary, else v1."""

    def __init__(self, learning_rate, lr_decay_after=1e5, scale=1.0):
        if isinstance(learning_rate, (tuple, list)) and len(learning_rate) > 1:
            self.values = (float(learning_rate[0]) * scale, float(learning_rate[1]) * scale)
            self.boundary = lr_decay_after
        else:
            if isinstance(learning_rate, (tuple, list)):
                learning_rate = learning_rate[0]
            self.values = (float(learning_rate) * scale,) * 2
            self.boundary = float('inf')

    def __call__(self, iterations):
        return self.values[0] if iterations <= self.boundary else self.values[1]


class SupervisedStep:
    """One optimizer step of ``model`` on fixed-shape device batches."""

    def __init__(self, model, batch_shapes, target_shape, loss='mae', lr=1e-3, lr_decay_after=1e5,
                 beta_1=0.9, beta_2=0.999, eps=1e-7, math=None, use_graph=True, lr_scale=1.0):
        self.model = model
        self.arena = model.arena
        dev = self.arena.device
        self.loss_kind = loss
        self.schedule = lr if isinstance(lr, LRSchedule) else LRSchedule(lr, lr_decay_after, lr_scale)
        self.b1, self.b2, self.eps = beta_1, beta_2, eps
        self.math = math or model.math
        self.inputs = [torch.zeros(s, dtype=torch.float32, device=dev) for s in batch_shapes]
        self.target = torch.zeros(target_shape, dtype=torch.float32, device=dev)
        self.loss_buf = torch.zeros(1, dtype=torch.float32, device=dev)
        self.lr_t_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        # lr_t staging: a small ring of pinned scalars, each guarded by an event recorded after its H2D copy, so a
        # host that runs a step ahead (the pipelined fit loops) never overwrites a value still waiting to be copied
        self._lr_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(4)] if dev.type == 'cuda' else None
        self._lr_events = [None] * 4
        self._lr_slot = 0
        import os
        # weight gradients on a side stream (engine.Ctx.wgrad_stream); DL4DS_WGRAD_STREAM=0 keeps one stream
        self.wgrad_stream = (torch.cuda.Stream(device=dev) if dev.type == 'cuda' and
                             os.environ.get('DL4DS_WGRAD_STREAM', '1') != '0' else None)
        self.use_graph = use_graph
        self.graph_fb = None        # the whole step: zero-grad + forward + loss + backward + all-reduce + Adam
        self.graph_opt = None       # (kept for callers that test for it: aliases graph_fb)
        self.launches_per_step = 0
        self.world = 1
        d = _dist()
        if d is not None:
            self.world = d.get_world_size()
        # DL4DS_COMM_BUCKETS=1: two gradient buckets, the upper one all-reduced on a communication stream under the rest
        # of the backward pass.  Measured on 2 B200 (round 2): 2.119-2.125 ms against 2.126-2.130 ms per step with the
        # single all-reduce -- what stays exposed is the latency of the LAST collective, not its size -- so the single
        # all-reduce remains the default.
        self.comm_stream = (torch.cuda.Stream(device=dev) if self.world > 1 and dev.type == 'cuda' and
                            os.environ.get('DL4DS_COMM_BUCKETS', '0') == '1' else None)
        self._early_off = 0

    # the recorded work -------------------------------------------------------------------------
    def _fwd_bwd(self, timers=None, plan=None):
        """``plan``: a PackPlan -- all tensor-core weight images in one launch at the head of the step (the captured
        graph); without it every convolution packs its own image on first use (eager passes)."""
        self.arena.grad.zero_()
        self.loss_buf.zero_()
        n_pack = plan.run() if plan is not None else 0
        ctx, out = self.model.forward(self.inputs, training=True, math=self.math, timers=timers,
                                      prepacked=plan.keys if plan is not None else None,
                                      wgrad_stream=self.wgrad_stream)
        ctx.pixel_loss(out, ctx.input(self.target), self.loss_kind, loss_buf=self.loss_buf)
        self._early_off = 0
        if self.comm_stream is not None and timers is None:
            self._arm_early_allreduce(ctx)
        ctx.backward()
        return ctx.launches + 2 + n_pack     # + the two memsets

    def _arm_early_allreduce(self, ctx):
        """Two gradient buckets: the upper half of the arena (parameters first used late in the forward pass: their
        gradients are final early in the backward pass) is all-reduced on a communication stream while the backward
        pass of the lower half still runs; ``_allreduce`` then only has the lower half left."""
        a = self.arena
        names = [n for n in a.spec if n in ctx.first_use]
        split = next((n for n in names if a.offsets[n] >= a.n // 2), None)
        if split is None or a.offsets[split] == 0:
            return
        upper = [n for n in names if a.offsets[n] >= a.offsets[split]]
        lower = [n for n in names if a.offsets[n] < a.offsets[split]]
        mark = min(ctx.first_use[n] for n in upper)
        if lower and max(ctx.first_use[n] for n in lower) >= mark:
            return          # creation order and first-use order disagree (not the case for the builders): one bucket
        off = a.offsets[split]

        def hook():
            main = torch.cuda.current_stream()
            self.comm_stream.wait_stream(main)
            if ctx._side_used:
                self.comm_stream.wait_stream(ctx.wgrad_stream)
            with torch.cuda.stream(self.comm_stream):
                allreduce_sum_(a.grad[off:])
            self._early_off = off
        ctx.bwd_hooks[mark] = hook

    def _opt(self):
        _lib.call('dl4ds_adam_step_dev', self.arena.theta.data_ptr(), self.arena.grad.data_ptr(),
                  self.arena.m.data_ptr(), self.arena.v.data_ptr(), self.arena.n,
                  self.lr_t_dev.data_ptr(), float(self.b1), float(self.b2), float(self.eps),
                  1.0 / self.world, _stream())
        return 1

    def _allreduce(self):
        if self.comm_stream is not None and getattr(self, '_early_off', 0) > 0:
            main = torch.cuda.current_stream()
            self.comm_stream.wait_stream(main)
            with torch.cuda.stream(self.comm_stream):
                allreduce_sum_(self.arena.grad[:self._early_off])
            main.wait_stream(self.comm_stream)
        else:
            allreduce_sum_(self.arena.grad)

    def _set_lr_t(self):
        self.arena.t += 1
        t = self.arena.t
        lr = self.schedule(t - 1)     # Keras evaluates the schedule at `iterations` before increment
        lr_t = lr * math.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
        if self._lr_host is not None:
            k = self._lr_slot
            self._lr_slot = (k + 1) % len(self._lr_host)
            if self._lr_events[k] is not None:
                self._lr_events[k].synchronize()
            self._lr_host[k][0] = lr_t
            self.lr_t_dev.copy_(self._lr_host[k], non_blocking=True)
            if self._lr_events[k] is None:
                self._lr_events[k] = torch.cuda.Event()
            self._lr_events[k].record()
        else:
            self.lr_t_dev.fill_(lr_t)

    def capture(self):
        """Warm up eagerly (module load, allocator), restore the optimizer state, then capture."""
        a = self.arena
        snap = (a.theta.clone(), a.m.clone(), a.v.clone(), a.t)
        self._set_lr_t()
        n1 = self._fwd_bwd()
        self._allreduce()
        n2 = self._opt()
        self.launches_per_step = n1 + n2
        torch.cuda.synchronize()
        a.theta.copy_(snap[0]); a.m.copy_(snap[1]); a.v.copy_(snap[2]); a.t = snap[3]
        if not self.use_graph:
            return self
        from .engine import PackPlan
        self.pack_plan = PackPlan(a, getattr(self.model, '_pack_cache', {}))      # images the eager pass above used
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.graph_fb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_fb, stream=s):
                # ONE graph: the NCCL all-reduce (dl4ds_comm_allreduce_sum on this stream) is captured between the
                # backward pass and Adam, so a replay has no host gap and no second graph launch
                self.launches_per_step = self._fwd_bwd(plan=self.pack_plan) + n2
                self._allreduce()
                self._opt()
            self.graph_opt = self.graph_fb
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        return self

    # per-batch entry points --------------------------------------------------------------------
    def load_batch(self, inputs, target):
        """Copy a batch (CUDA tensors, or pinned host tensors -> async H2D) into the static buffers."""
        for dst, src in zip(self.inputs, inputs):
            dst.copy_(src, non_blocking=True)
        self.target.copy_(target, non_blocking=True)

    def run(self):
        """One optimizer step on the batch in the static buffers.  Returns the device loss scalar
        (valid after the stream is synchronised / ``.item()``)."""
        if self.model.arena is not self.arena:
            raise RuntimeError('the model arena was replaced (Model.to() onto another device) after this step was '
                               'built: its captured graphs point into the old arena -- build a new SupervisedStep')
        self._set_lr_t()
        if self.graph_fb is not None:
            self.graph_fb.replay()
        else:
            self._fwd_bwd()
            self._allreduce()
            self._opt()
        return self.loss_buf

    def run_profiled(self, timers):
        """One eager (un-graphed) forward + backward with CUDA events around ``Ctx.TIMER_REPS`` repeats of every
        convolution-family launch; ``timers`` collects {label: [(start, end)]} (elapsed / TIMER_REPS = one
        launch).  The repeated launches over-accumulate gradients, so no optimizer step is taken and the state is
        left untouched.  Used by bench.py's roofline line."""
        self._fwd_bwd(timers)
        torch.cuda.synchronize()
        return self.loss_buf

    def broadcast_from_rank0(self):
        """hvd BroadcastGlobalVariablesCallback(0) (supervised.py:369) + optimizer slots."""
        d = _dist()
        if d is not None:
            from . import comm
            for t in (self.arena.theta, self.arena.m, self.arena.v):
                if t.is_cuda and comm.ensure(d):
                    comm.broadcast_(t, 0)
                else:
                    d.broadcast(t, src=0)


class EvalStep:
    """Forward + loss on fixed-shape device batches (Keras ``evaluate`` / validation)."""

    def __init__(self, model, loss='mae', math=None):
        self.model = model
        self.loss_kind = loss
        self.math = math or model.math

    def run(self, inputs, target):
        dev = self.model.arena.device
        ctx = Ctx(self.model.arena, self.math, training=False)
        vs = [ctx.input(x) for x in inputs]
        out = self.model.fn(ctx, vs)
        buf = torch.zeros(1, dtype=torch.float32, device=dev)
        ctx.pixel_loss(out, ctx.input(target), self.loss_kind, loss_buf=buf)
        return buf


def coarsen_on_device(hr, scale):
    """(N,H,W,C) CUDA fp32 -> (N,H/s,W/s,C): exact s x s block mean == cv2 INTER_AREA at an integer
    factor (utils.py:376-384), the HR->LR step of dataloader.py:204-208 done on the device."""
    assert hr.is_cuda and hr.dtype == torch.float32 and hr.is_contiguous()
    n, h, w, c = hr.shape
    lr = torch.empty((n, h // scale, w // scale, c), dtype=torch.float32, device=hr.device)
    _lib.call('dl4ds_avgpool_coarsen', hr.data_ptr(), lr.data_ptr(), n, h, w, c, int(scale), _stream())
    return lr
