"""Which kernel serves a 48 -> 48 3x3 convolution / its dgrad / wgrad at 16 x 128 x 128 when a tensor is a channel slice
of a wider (pitch 100) buffer -- run under `ncu --metrics gpu__time_duration.sum --csv` and read the kernel names."""
import sys
import torch
sys.path.insert(0, '.')
from dl4ds_b200 import engine
from dl4ds_b200.engine import Ctx, Var, Arena
N, H, W, C = 16, 128, 128, 48
spec = {'cv/kernel': (3, 3, C, C), 'cv/bias': (C,)}
arena = Arena(spec, 'cuda')
arena.theta.normal_()
for tag, ld, off in (('dense', 48, 0), ('slice of 100 @48', 100, 48), ('slice of 100 @0', 100, 0)):
    c = Ctx(arena, 'tf32x3', training=True)
    xb = torch.randn(N, H, W, C, device='cuda')
    x = Var(xb, requires_grad=True)
    torch.cuda.synchronize(); print('==', tag, flush=True)
    y = c.conv(x, 'cv', C, act='relu')
    gb = torch.randn(N, H, W, ld, device='cuda')
    y.grad = Var(gb, off, C)
    torch.cuda.nvtx.range_push(tag)
    c.backward()
    torch.cuda.nvtx.range_pop()
    torch.cuda.synchronize()
