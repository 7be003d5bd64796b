"""In-graph kernel timeline of the headline training step via torch.profiler (CUPTI): per-kernel warm durations,
total busy time vs wall span (= launch gaps inside the captured graph)."""
import sys, collections
import numpy as np
import torch
sys.path.insert(0, '.')
from dl4ds_b200 import training
B, HW = 64, 128
rng = np.random.default_rng(0)
hr = rng.standard_normal((2 * B, HW, HW, 1), dtype=np.float32)
tr = training.SupervisedTrainer('resnet', 'spc', hr, hr[:B], hr[:B], scale=4, batch_size=B, loss='mae', epochs=1,
                                learning_rate=1e-3, device='GPU', verbose=False, save=False, show_plot=False, math='tf32x3', seed=0)
tr.setup_model()
step = tr.train_step
for _ in range(5):
    step.run()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step.run()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
busy = sum(e.time_range.end - e.time_range.start for e in evs)
print('kernels %d, span %.1f us, busy (sum of durations) %.1f us, per step span %.1f busy %.1f' % (len(evs), t1 - t0, busy, (t1 - t0) / 3, busy / 3))
agg = collections.defaultdict(lambda: [0, 0.0])
for e in evs:
    k = e.name.replace('dl4ds::', '').replace('void ', '')[:60]
    agg[k][0] += 1; agg[k][1] += e.time_range.end - e.time_range.start
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print('%-62s %4d %9.1f us/step %5.1f%%' % (k, v[0] // 3, v[1] / 3, 100 * v[1] / busy))
# gaps
gaps = []
for a, b in zip(evs[:-1], evs[1:]):
    gaps.append(b.time_range.start - a.time_range.end)
gaps = np.array(gaps)
print('gaps: mean %.2f us, median %.2f, sum of positive gaps per step %.1f us, overlaps (negative) per step %.1f us' % (gaps.mean(), np.median(gaps), gaps[gaps > 0].sum() / 3, -gaps[gaps < 0].sum() / 3))
