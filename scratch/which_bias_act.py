"""Which layers of the headline step still launch dl4ds_bias_act_bwd (i.e. have no fused epilogue-backward)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from dl4ds_b200 import training, engine
B, HW = 64, 128
rng = np.random.default_rng(0)
hr = rng.standard_normal((2 * B, HW, HW, 1), dtype=np.float32)
tr = training.SupervisedTrainer('resnet', 'spc', hr, hr[:B], hr[:B], scale=4, batch_size=B, loss='mae', epochs=1,
                                learning_rate=1e-3, device='GPU', verbose=False, save=False, show_plot=False, math='tf32x3', seed=0)
tr.setup_model()
step = tr.train_step
orig = engine.Ctx._call
def patched(self, name, *args):
    if name in ('dl4ds_bias_act_bwd',):
        ints = [a for a in args if isinstance(a, int) and abs(a) < (1 << 20)]
        import sys as _s
        fr = _s._getframe(1)
        print(name, ints, fr.f_code.co_name, fr.f_locals.get('name'), 'res' if fr.f_locals.get('res') is not None else '', flush=True)
    return orig(self, name, *args)
engine.Ctx._call = patched
step._fwd_bwd()
torch.cuda.synchronize()
