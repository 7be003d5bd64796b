"""Loss trajectory of the headline step on fixed data: run under DL4DS_PDL=0 / 1 and compare."""
import sys, os
sys.path.insert(0, '.')
import numpy as np, torch
from dl4ds_b200 import nets
from dl4ds_b200.step import SupervisedStep
B = 64
torch.manual_seed(0)
m = nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (32, 32), math='tf32x3').to('cuda').init_weights(0)
st = SupervisedStep(m, [(B, 32, 32, 1)], (B, 128, 128, 1), use_graph=True).capture()
xs = [torch.randn((B, 32, 32, 1), device='cuda') for _ in range(4)]
ys = [torch.randn((B, 128, 128, 1), device='cuda') for _ in range(4)]
out = []
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
    st.load_batch([xs[i % 4]], ys[i % 4])
    out.append(float(st.run().item()))
print(' '.join('%.7f' % v for v in out))
w = m.arena.theta.double()
print('theta checksum %.10e' % float((w * w).sum()))
