"""Run one convolution forward (+ dgrad) at full size in a math mode and compare with tf32x3.
usage: python scratch/f16_shapes.py N H W Cin Cout k [d2s]"""
import sys
import torch
sys.path.insert(0, '.')
from dl4ds_b200.engine import Arena, Ctx, Var
from dl4ds_b200.spec import SpecCtx
N, H, W, Cin, Cout, k = [int(v) for v in sys.argv[1:7]]
d2s = int(sys.argv[7]) if len(sys.argv) > 7 else 1
dev = torch.device('cuda')
fn = lambda c, xs: c.conv(xs[0], 'cv', Cout, k=k, d2s=d2s, act='relu' if d2s == 1 else None)
sc = SpecCtx(); fn(sc, [sc.input((N, H, W, Cin))])
arena = Arena(sc.spec, dev); arena.theta.normal_(0, 0.05)
torch.manual_seed(0)
x = torch.randn((N, H, W, Cin), device=dev)
outs = {}
for math in ('tf32x3', 'f16x3'):
    ctx = Ctx(arena, math, training=True)
    xv = ctx.input(x, requires_grad=True)
    y = fn(ctx, [xv])
    g = torch.randn_like(y.buf)
    torch.manual_seed(1)
    y.grad = Var(torch.randn(y.buf.shape, device=dev))
    ctx.backward()
    torch.cuda.synchronize()
    outs[math] = (y.buf.clone(), xv.grad.buf.clone())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx2 = Ctx(arena, math, training=False); ctx2.pack_cache = {}
    v = ctx2.input(x)
    for _ in range(3): fn(ctx2, [v])
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): fn(ctx2, [v])
    e1.record(); torch.cuda.synchronize()
    print(math, 'fwd %.1f us/launch (eager, incl. pack + launch gaps)' % (e0.elapsed_time(e1) * 100), flush=True)
for i, nm in enumerate(('fwd', 'dgrad')):
    a, b = outs['f16x3'][i], outs['tf32x3'][i]
    print(nm, 'max |f16x3 - tf32x3| / max|.| = %.2e' % ((a - b).abs().max() / b.abs().max()).item())
