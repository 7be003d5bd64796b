"""Shared helpers for the parity tests: run a graph function both through the CUDA engine
(dl4ds_b200.engine.Ctx via the C ABI) and through the oracle (oracle/torch_ref.py, torch-CPU),
and compare outputs, input gradients and parameter gradients."""
from collections import OrderedDict

import numpy as np
import torch

from dl4ds_b200.engine import Arena, Ctx, Var
from dl4ds_b200.spec import SpecCtx
from oracle import torch_ref as R


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def trace_spec(fn, shapes):
    sc = SpecCtx()
    fn(sc, [sc.input(s) for s in shapes])
    return sc.spec


def run_engine(fn, spec, weights, inputs, device, math='fp32', seed_grad=None, input_grads=True):
    """fn(ctx, [Var]) -> Var.  Returns (out ndarray, {param grads}, [input grads])."""
    arena = Arena(spec, device)
    arena.load(weights)
    ctx = Ctx(arena, math, training=True)
    vs = [ctx.input(torch.as_tensor(x).to(device).float().contiguous(), requires_grad=input_grads)
          for x in inputs]
    out = fn(ctx, vs)
    y = out.t.detach().cpu().numpy().copy()
    if seed_grad is not None:
        out.grad = Var(torch.as_tensor(seed_grad).to(device).float().contiguous())
        ctx.backward()
        torch.cuda.synchronize()
        igr = [v.grad.t.detach().cpu().numpy().copy() if v.grad is not None else None for v in vs]
        return y, arena.grads(), igr
    torch.cuda.synchronize()
    return y, None, None


def run_oracle(ofn, weights, inputs, seed_grad=None):
    """ofn(Params, [torch NHWC tensors]) -> NHWC tensor (torch CPU)."""
    ws = OrderedDict((k, torch.as_tensor(np.asarray(v)).clone().requires_grad_(True)) for k, v in weights.items())
    xs = [torch.as_tensor(np.asarray(x)).clone().requires_grad_(True) for x in inputs]
    y = ofn(R.Params(ws), xs)
    if seed_grad is not None:
        y.backward(torch.as_tensor(seed_grad))
        pg = OrderedDict((k, (w.grad.numpy() if w.grad is not None else np.zeros(tuple(w.shape), np.float32)))
                         for k, w in ws.items())
        ig = [x.grad.numpy() if x.grad is not None else None for x in xs]
        return y.detach().numpy(), pg, ig
    return y.detach().numpy(), None, None


def compare(fn, ofn, shapes, device, seed=0, math='fp32', tol=2e-5, gtol=2e-4, bias_scale=0.1,
            input_grads=True, scale_inputs=1.0, skip_grads=()):
    """Full fwd/bwd parity of an engine graph `fn` against the oracle graph `ofn`."""
    spec = trace_spec(fn, shapes)
    weights = R.init_weights(spec, seed=seed, bias_scale=bias_scale)
    rng = np.random.default_rng(seed + 1)
    inputs = [(scale_inputs * rng.standard_normal(s)).astype(np.float32) for s in shapes]
    y_ref, pg_ref, ig_ref = None, None, None
    y0, _, _ = run_oracle(ofn, weights, inputs)
    seed_grad = rng.standard_normal(y0.shape).astype(np.float32)
    y_ref, pg_ref, ig_ref = run_oracle(ofn, weights, inputs, seed_grad)
    y, pg, ig = run_engine(fn, spec, {k: v.numpy() for k, v in weights.items()}, inputs, device, math,
                           seed_grad, input_grads)
    assert y.shape == y_ref.shape, (y.shape, y_ref.shape)
    e = rel_err(y, y_ref)
    assert e <= tol, 'forward rel err %g > %g' % (e, tol)
    for k in spec:
        if (skip_grads(k) if callable(skip_grads) else k in skip_grads):      # analytically zero gradients (a bias in front of a batch norm): noise / noise
            continue
        e = rel_err(pg[k], pg_ref[k])
        assert e <= gtol, 'grad %s rel err %g > %g' % (k, e, gtol)
    if input_grads:
        for a, b in zip(ig, ig_ref):
            if b is None:
                continue
            e = rel_err(a, b)
            assert e <= gtol, 'input grad rel err %g > %g' % (e, gtol)
    return y, y_ref


def assert_adam_weights_close(new, ref, lr, steps=1, tight=5e-5, frac=2e-3):
    """Weights after `steps` Adam steps against the oracle's.  Adam's first steps move a weight by ~lr*sign(g), so
    a gradient element whose sign is decided by summation-order noise (fp32 atomics of the split-K kernels vary
    from run to run) legitimately shifts that weight by up to 2*lr per step: almost all weights must agree to
    `tight`; at most a fraction `frac` may differ, and by no more than 2*lr*steps (+ tight)."""
    n_bad, n_all, worst = 0, 0, 0.0
    for k in ref:
        d = np.abs(np.asarray(new[k], np.float64) - np.asarray(ref[k], np.float64))
        n_bad += int((d > tight).sum())
        n_all += d.size
        worst = max(worst, float(d.max()))
    assert worst <= 2.0 * lr * steps + tight, worst
    assert n_bad <= max(2, int(frac * n_all)), (n_bad, n_all, worst)
