"""A few optimizer steps of the convnext + 4x SPC model at the headline size, for an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <count> --csv ... python scratch/prof_convnext.py`)."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from dl4ds_b200 import SupervisedTrainer   # noqa: E402

B = 64
hr = np.random.default_rng(1234).standard_normal((2 * B, 128, 128, 1), dtype=np.float32)
np.random.seed(0)
tr = SupervisedTrainer('convnext', 'spc', hr, hr[:B], hr[:B], scale=4, batch_size=B, epochs=1, learning_rate=1e-3,
                       verbose=False, math='tf32x3', seed=1, normalization='ln', activation='gelu')
tr.setup_datagen()
tr.setup_model()
x, y = tr.ds_train[0]
for _ in range(4):
    tr.train_on_batch(x, y[0])
torch.cuda.synchronize()
print('launches per step', tr.train_step.launches_per_step)
import time
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st = tr.train_step
e0.record()
for _ in range(20):
    st.run()
e1.record(); torch.cuda.synchronize()
print('convnext + ln + gelu, 4x SPC, batch 64: %.3f ms per captured step' % (e0.elapsed_time(e1) / 20))
