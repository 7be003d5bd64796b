"""Where does the whole-network gradient error of the literal cfg2 come from?  Compare, against an fp64 evaluation of
the oracle graph: the oracle in fp32 (torch CPU), the CUDA path in exact fp32 (CUDA cores) and in tf32x3."""
import sys
from collections import OrderedDict
import numpy as np, torch
sys.path.insert(0, '.')
from dl4ds_b200 import nets
from oracle import torch_ref as R
from tests.util import rel_err, run_engine, trace_spec

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
act = sys.argv[2] if len(sys.argv) > 2 else 'relu'
m = nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (32, 32), math='tf32x3', activation=act)
shapes = [(B, 32, 32, 1)]
spec = trace_spec(m.fn, shapes)
weights = R.init_weights(spec, seed=3, bias_scale=0.1)
rng = np.random.default_rng(4)
x = rng.standard_normal(shapes[0]).astype(np.float32)
sg = (rng.standard_normal((B, 128, 128, 1)) / (B * 128 * 128)).astype(np.float32)

def oracle(dtype):
    ws = OrderedDict((k, torch.as_tensor(np.asarray(v)).to(dtype).clone().requires_grad_(True)) for k, v in weights.items())
    xs = [torch.as_tensor(x).to(dtype)]
    y = R.net_postupsampling(R.Params(ws, dtype=dtype), xs, 'resnet', 'spc', 4, activation=act)
    y.backward(torch.as_tensor(sg).to(dtype))
    return y.detach().numpy(), OrderedDict((k, w.grad.numpy()) for k, w in ws.items())

y64, g64 = oracle(torch.float64)
y32, g32 = oracle(torch.float32)
res = {'oracle fp32': (y32, g32)}
for math in ('fp32', 'tf32x3'):
    y, pg, _ = run_engine(m.fn, spec, {k: v.numpy() for k, v in weights.items()}, [x], 'cuda', math, sg, input_grads=False)
    res['cuda ' + math] = (y, pg)
y2, pg2, _ = run_engine(m.fn, spec, {k: v.numpy() for k, v in weights.items()}, [x], 'cuda', 'tf32x3', sg, input_grads=False)
res['cuda tf32x3 (2nd run)'] = (y2, pg2)
print('batch %d, activation %s: error against the fp64 oracle (forward: of max|y|; gradients: worst tensor, of its max)' % (B, act))
for tag, (y, pg) in res.items():
    worst = max((rel_err(pg[k], g64[k]), k) for k in g64)
    errs = sorted(rel_err(pg[k], g64[k]) for k in g64)
    print('%-24s fwd %.2e  grad worst %.2e (%s)  median %.2e' % (tag, rel_err(y, y64), worst[0], worst[1], errs[len(errs) // 2]))
a, b = res['cuda tf32x3'][1], res['cuda tf32x3 (2nd run)'][1]
print('tf32x3 run-to-run: worst %.2e' % max(rel_err(a[k], b[k]) for k in a))
a, b = res['cuda tf32x3'][1], res['oracle fp32'][1]
print('tf32x3 vs oracle fp32: worst %.2e' % max(rel_err(a[k], b[k]) for k in a))
if len(sys.argv) > 3:
    print('%-46s %10s %10s %10s' % ('tensor', 'oracle32', 'cuda fp32', 'tf32x3'))
    for k in g64:
        print('%-46s %10.2e %10.2e %10.2e' % (k, rel_err(res['oracle fp32'][1][k], g64[k]), rel_err(res['cuda fp32'][1][k], g64[k]),
                                              rel_err(res['cuda tf32x3'][1][k], g64[k])))
