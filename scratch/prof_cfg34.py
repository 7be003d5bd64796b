"""One eager optimizer step of BASELINE config 3 or 4 between cudaProfilerStart/Stop for an ncu launch list:
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv ... python scratch/prof_cfg34.py cfg3"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from dl4ds_b200 import training
which = sys.argv[1] if len(sys.argv) > 1 else 'cfg3'
rng = np.random.default_rng(4321)
if which == 'cfg3':
    B = 16
    hr = rng.standard_normal((2 * B, 128, 128, 1), dtype=np.float32)
    preds = [rng.standard_normal((2 * B, 128, 128, 1), dtype=np.float32) for _ in range(3)]
    static = rng.standard_normal((128, 128)).astype(np.float32)
    tr = training.SupervisedTrainer('densenet', 'dc', hr, hr[:B], hr[:B], predictors_train=preds,
                                    predictors_val=[q[:B] for q in preds], predictors_test=[q[:B] for q in preds],
                                    static_vars=[static], scale=8, batch_size=B, epochs=1, learning_rate=1e-3,
                                    verbose=False, save=False, math='tf32x3', seed=1, attention=True, localcon_layer=True)
else:
    B, T = 8, 6
    hr = rng.standard_normal((2 * B + T, 128, 128, 1), dtype=np.float32)
    tr = training.SupervisedTrainer('resnet', 'rc', hr, hr[:B + T], hr[:B + T], scale=4, time_window=T, batch_size=B,
                                    epochs=1, learning_rate=1e-3, verbose=False, save=False, math='tf32x3', seed=1)
tr.setup_datagen()
tr.setup_model()
st = tr.train_step
for _ in range(3):
    st.run()
torch.cuda.synchronize()
if len(sys.argv) > 2 and sys.argv[2] == 'trace':     # per-call device time (events, eager) with the calling layer
    from dl4ds_b200 import engine
    orig = engine.Ctx._call
    rows = []
    def patched(self, name, *args):
        fr = sys._getframe(1)
        if fr.f_code.co_name in ('_timed', 'go'):
            fr = fr.f_back
        if fr.f_code.co_name in ('go', '_wgrad', 'wr'):
            fr = fr.f_back
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = orig(self, name, *args)
        e1.record()
        ints = [a for a in args if isinstance(a, int) and abs(a) < (1 << 20)]
        rows.append((e0, e1, name, ints, fr.f_code.co_name, fr.f_locals.get('name') or fr.f_locals.get('label')))
        return rc
    engine.Ctx._call = patched
    st.wgrad_stream = None
    st._fwd_bwd()
    torch.cuda.synchronize()
    out = sorted(((a.elapsed_time(b) * 1e3, n, i, f, l) for a, b, n, i, f, l in rows), key=lambda r: -r[0])
    print('total %.0f us in %d calls' % (sum(r[0] for r in out), len(out)))
    for r in out[:28]:
        print('%8.1f us  %-24s %-14s %-34s %s' % (r[0], r[1], r[3], r[4], r[2]))
    sys.exit(0)
torch.cuda.profiler.start()
st._fwd_bwd()          # the recorded work, eagerly (the graph's kernels, serialised)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
