"""Time individual conv passes (fwd / dgrad / wgrad) of the headline layers with CUDA events.
usage: python scratch/time_layers.py [math]"""
import sys
import torch
sys.path.insert(0, '.')
from dl4ds_b200.engine import Arena, Ctx, Var
from dl4ds_b200.spec import SpecCtx

math = sys.argv[1] if len(sys.argv) > 1 else 'tf32x3'
dev = torch.device('cuda')


def run(tag, shape, cout, k, d2s=1, reps=5):
    fn = lambda c, xs: c.conv(xs[0], 'cv', cout, k=k, d2s=d2s)
    sc = SpecCtx()
    fn(sc, [sc.input(shape)])
    arena = Arena(sc.spec, dev)
    arena.theta.normal_(0, 0.05)
    x = torch.randn(shape, device=dev)
    acc = {}
    for i in range(reps):
        ctx = Ctx(arena, math, training=True)
        ctx.timers = {}
        xv = ctx.input(x, requires_grad=True)
        out = fn(ctx, [xv])
        out.grad = Var(torch.randn_like(out.buf))
        ctx.backward()
        torch.cuda.synchronize()
        if i >= 2:
            for kk, v in ctx.timers.items():
                acc.setdefault(kk.split(':')[1].split('@')[0], []).append(sum(a.elapsed_time(b) for a, b in v))
    macs = shape[0] * shape[1] * shape[2] * k * k * shape[3] * cout
    print('%-28s' % tag, ' '.join('%s %.1f us (%.0f TF/s)' % (kk, 1e3 * sum(v) / len(v), 2 * macs / (sum(v) / len(v) * 1e-3) / 1e12)
                                  for kk, v in acc.items()), flush=True)


run('SPC 48->192 @64 d2s', (64, 64, 64, 48), 192, 3, 2)
run('SPC 48->192 @32 d2s', (64, 32, 32, 48), 192, 3, 2)
run('bb 48->48 @32', (64, 32, 32, 48), 48, 3)
run('bb 40->40 @32', (64, 32, 32, 40), 40, 3)
run('bb 32->32 @32', (64, 32, 32, 32), 32, 3)
run('bb 24->24 @32', (64, 32, 32, 24), 24, 3)
run('bb 16->16 @32', (64, 32, 32, 16), 16, 3)
run('bb 8->48 1x1 @32', (64, 32, 32, 8), 48, 1)
run('TL 48->8 1x1 @128', (64, 128, 128, 48), 8, 1)
run('tail 8->8 @128', (64, 128, 128, 8), 8, 3)
