import sys, os
sys.path.insert(0, '.')
import numpy as np, torch
from dl4ds_b200 import nets
from dl4ds_b200.step import SupervisedStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (32, 32), math='f16x3').to('cuda').init_weights(0)
st = SupervisedStep(m, [(B, 32, 32, 1)], (B, 128, 128, 1), use_graph=len(sys.argv) > 2)
st.inputs[0].normal_(); st.target.normal_()
print('eager step', flush=True)
st.capture()
print('captured', flush=True)
for i in range(3):
    print('loss', float(st.run().item()), flush=True)
