"""Forward-only timing of single tensor-core layers (CUDA events, 20 back-to-back launches after warm-up).
usage: python scratch/time_fwd.py [math]"""
import os, sys
import torch
sys.path.insert(0, '.')
from dl4ds_b200.engine import Arena, Ctx
from dl4ds_b200.spec import SpecCtx

math = sys.argv[1] if len(sys.argv) > 1 else 'tf32x3'
dev = torch.device('cuda')
CASES = [('composed 48->32 @64 d2s', (64, 64, 64, 48), 32, 3, 2), ('SPC1 48->192 @32 d2s', (64, 32, 32, 48), 192, 3, 2),
         ('SPC1 dgrad 192->48 @32', (64, 32, 32, 192), 48, 3, 1), ('bb 48->48 @32', (64, 32, 32, 48), 48, 3, 1),
         ('bb 32->32 @32', (64, 32, 32, 32), 32, 3, 1), ('bb 16->16 @32', (64, 32, 32, 16), 16, 3, 1)]
out = []
for tag, shape, cout, k, d2s in CASES:
    fn = lambda c, xs: c.conv(xs[0], 'cv', cout, k=k, d2s=d2s)
    sc = SpecCtx(); fn(sc, [sc.input(shape)])
    arena = Arena(sc.spec, dev); arena.theta.normal_(0, 0.05)
    x = torch.randn(shape, device=dev)
    cache = {}
    def once():
        ctx = Ctx(arena, math, training=False); ctx.pack_cache = cache
        fn(ctx, [ctx.input(x)])
    for _ in range(3): once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(10): once()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    macs = shape[0] * shape[1] * shape[2] * k * k * shape[3] * cout
    out.append('%s %.1f us (%.0f TF/s)' % (tag, us, 2 * macs / us / 1e6))
print(os.environ.get('TAG', ''), ' | '.join(out), flush=True)
