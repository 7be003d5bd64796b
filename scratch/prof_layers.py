"""Run the headline layers in isolation (for ncu --set full): SPC conv 48->192 + depth_to_space at
64x64 (batch 64) forward / dgrad / wgrad, and one backbone layer 48->48 at 32x32."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from dl4ds_b200.engine import Arena, Ctx, Var
from dl4ds_b200.spec import SpecCtx

math = sys.argv[1] if len(sys.argv) > 1 else 'tf32x3'
dev = torch.device('cuda')


def run(shape, cout, d2s, reps):
    fn = lambda c, xs: c.conv(xs[0], 'cv', cout, k=3, d2s=d2s)
    sc = SpecCtx()
    fn(sc, [sc.input(shape)])
    arena = Arena(sc.spec, dev)
    arena.theta.normal_(0, 0.05)
    x = torch.randn(shape, device=dev)
    for _ in range(reps):
        ctx = Ctx(arena, math, training=True)
        xv = ctx.input(x, requires_grad=True)
        out = fn(ctx, [xv])
        out.grad = Var(torch.randn_like(out.buf))
        ctx.backward()
    torch.cuda.synchronize()


run((64, 64, 64, 48), 32, 2, 2)       # composed last sub-pixel stage x TransitionLast (48 -> 4*8, depth_to_space)
run((64, 32, 32, 48), 192, 2, 2)      # first sub-pixel stage
run((64, 32, 32, 48), 48, 1, 2)       # widest backbone layer
run((64, 128, 128, 8), 8, 1, 2)       # HR tail
print('done')
