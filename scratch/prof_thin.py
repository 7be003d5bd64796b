"""HR-tail layer (8 -> 8, 3x3, 128x128, batch 64) forward / wgrad / dgrad in isolation for ncu --set full."""
import sys
import torch
sys.path.insert(0, '.')
from dl4ds_b200.engine import Arena, Ctx, Var
from dl4ds_b200.spec import SpecCtx
dev = torch.device('cuda')
fn = lambda c, xs: c.conv(xs[0], 'cv', 8, k=3, act='relu')
sc = SpecCtx(); fn(sc, [sc.input((64, 128, 128, 8))])
arena = Arena(sc.spec, dev); arena.theta.normal_(0, 0.05)
x = torch.randn((64, 128, 128, 8), device=dev)
for _ in range(2):
    ctx = Ctx(arena, 'tf32x3', training=True)
    xv = ctx.input(x, requires_grad=True)
    out = fn(ctx, [xv]); out.grad = Var(torch.randn_like(out.buf)); ctx.backward()
torch.cuda.synchronize(); print('done')
