"""Pipeline timeline of CTA 0 of conv_tc_halo_kernel from its clock64 stamps.
usage: python scratch/halo_stamps.py N H W Cin Cout k [d2s]"""
import sys
import torch
sys.path.insert(0, '.')
from dl4ds_b200 import _lib
from dl4ds_b200.engine import Arena, Ctx
from dl4ds_b200.spec import SpecCtx
N, H, W, Cin, Cout, k = [int(v) for v in sys.argv[1:7]]
d2s = int(sys.argv[7]) if len(sys.argv) > 7 else 1
import os
MATH = os.environ.get('MATH', 'tf32x3')
dev = torch.device('cuda')
lib = _lib.load()
fn = lambda c, xs: c.conv(xs[0], 'cv', Cout, k=k, d2s=d2s)
sc = SpecCtx(); fn(sc, [sc.input((N, H, W, Cin))])
arena = Arena(sc.spec, dev); arena.theta.normal_(0, 0.05)
x = torch.randn((N, H, W, Cin), device=dev)
cache = {}
dbg = torch.zeros(64 * 16 + 64, dtype=torch.int64, device=dev)
for rep in range(3):
    dbg.zero_()
    lib.dl4ds_debug_set_buffer(dbg.data_ptr())
    ctx = Ctx(arena, MATH, training=False); ctx.pack_cache = cache
    fn(ctx, [ctx.input(x)])
    torch.cuda.synchronize()
lib.dl4ds_debug_set_buffer(None)
t = dbg.cpu()[:1024].view(64, 16)
t0 = int(t[t > 0].min())
print('entry %d prologue_done %d cta_done %d (clk rel. to first stamp)' % tuple(int(t[0, j]) - t0 for j in (12, 13, 14)))
import time
ctx = Ctx(arena, MATH, training=False); ctx.pack_cache = cache
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
v = ctx.input(x)
for _ in range(3): fn(ctx, [v])
e0.record()
for _ in range(20): fn(ctx, [v])
e1.record(); torch.cuda.synchronize()
print('kernel (back-to-back launches incl. gaps): %.2f us' % (e0.elapsed_time(e1) * 50))
names = ['P:start', 'P:loaded', 'P:fenced', 'P:acquired', 'M:a_ready', 'M:issued', 'M:commit', 'P:stored', 'M:tempty', 'M:tfull', 'E:tfull', 'E:done']
print('item ' + ' '.join('%10s' % n for n in names))
for it in range(40):
    if int(t[it].max()) == 0:
        break
    print('%4d ' % it + ' '.join('%10d' % (int(t[it, j]) - t0 if int(t[it, j]) else -1) for j in range(12)))
