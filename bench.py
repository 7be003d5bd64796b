#!/usr/bin/env python
"""Headline benchmark: HR pixels per second of one SUPERVISED TRAINING STEP (forward + MAE +
backward + [gradient all-reduce] + Adam) of the residual-backbone 4x sub-pixel model on synthetic
32 -> 128 single-channel tiles, batch 64 per GPU (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--math fp32|tf32x3|tf32]

N > 1 is launched by ``python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N``:
one rank per GPU, batch sharded (weak scaling, per-GPU batch fixed as in the reference,
training/base.py:115-116), ONE gradient all-reduce per step.  Rank 0 prints ONE JSON line.

  value      device-resident inputs (a pool of pre-staged batches in HBM), the captured CUDA graph
             of the whole step replayed K times, CUDA-event timed, max over ranks.
  e2e        the same step driven through the public trainer API (`SupervisedTrainer.train_on_batch`)
             with HOST numpy batches: pinned staging + H2D of LR/HR every step and a D2H read of
             the loss inside the timed region.
  roofline   the dominant kernel of the step (largest summed duration), timed with CUDA events
             around its launches on the launching stream in extra eager steps right after the
             timed region; algorithmic FLOPs = 2*MACs of those launches.
  cpu_baseline / --impl reference
             the oracle (oracle/torch_ref.py: torch-CPU fp32 restatement of the reference's TF/Keras
             graph + Keras Adam) on the host cores.  TF itself is not installable offline.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'HR-pixels/sec, supervised training step (4x residual SPC, 32->128)'
UNIT = 'HR-px/s'
LR_HW, SCALE, BATCH = 32, 4, 64
HR_HW = LR_HW * SCALE
WORKLOAD = ('SupervisedTrainer step: resnet backbone (n_filters=8, n_blocks=6), 4x SPC, 32->128, '
            '1 channel, batch 64 per GPU, MAE, Adam(1e-3)')


def _config(n_gpus, extra=None):
    cfg = {'workload': WORKLOAD, 'global_batch': BATCH * n_gpus, 'per_gpu_batch': BATCH,
           'lr_hw': LR_HW, 'hr_hw': HR_HW, 'scale': SCALE, 'parallelism': 'dp%d' % n_gpus,
           'l2_policy': 'working set >> L2: ~1 GB of activations per step and 4 rotating device '
                        'batches; no explicit flush'}
    if extra:
        cfg.update(extra)
    return cfg


# --------------------------------------------------------------------------------------------
# CPU leg (oracle): used for cpu_baseline and for --impl reference
# --------------------------------------------------------------------------------------------
def cpu_oracle_rate(batch, steps, warmup, threads, with_batch_creation=False):
    """HR-px/s and seconds per step of the oracle's training step on the host cores.  ``with_batch_creation``: every
    timed step also builds its (LR, HR) batch from the HR array the way the reference does per step on the host
    (``create_batch_hr_lr``: per-sample slicing + cv2 INTER_AREA coarsening, dataloader.py:297-360) -- through
    dl4ds_b200.dataloader's restatement, which tests/test_datapath.py pins bit-for-bit to the reference's own output
    (/root/reference is not present on the GPU box)."""
    import numpy as np
    import torch
    from oracle import torch_ref as R
    from dl4ds_b200 import nets
    torch.set_num_threads(threads)
    m = nets.net_postupsampling('resnet', 'spc', SCALE, 1, 0, (LR_HW, LR_HW))
    w = R.init_weights(m.spec, seed=0)
    opt = R.TFAdam(list(w), lr=1e-3)
    fwd = lambda p, xs: R.net_postupsampling(p, xs, 'resnet', 'spc', SCALE)
    rng = np.random.default_rng(1234)
    hr = rng.standard_normal((batch, HR_HW, HR_HW, 1), dtype=np.float32)
    lr = hr.reshape(batch, LR_HW, SCALE, LR_HW, SCALE, 1).mean(axis=(2, 4)).astype(np.float32)
    hr_t, lr_t = torch.from_numpy(hr), torch.from_numpy(lr)

    def one():
        if with_batch_creation:
            from dl4ds_b200.dataloader import create_batch_hr_lr
            (lr_b,), (hr_b,) = create_batch_hr_lr(np.arange(batch), 0, hr, None, upsampling='spc', scale=SCALE,
                                                  batch_size=batch)
            R.supervised_step(fwd, w, opt, [torch.from_numpy(np.ascontiguousarray(lr_b))],
                              torch.from_numpy(np.ascontiguousarray(hr_b)))
        else:
            R.supervised_step(fwd, w, opt, [lr_t], hr_t)
    for _ in range(warmup):
        one()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return batch * HR_HW * HR_HW / sec, sec


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_b = 16
    rate, sec = cpu_oracle_rate(sample_b, args.steps, args.warmup, threads)
    rate_e2e, sec_e2e = cpu_oracle_rate(sample_b, max(2, args.steps // 2), 1, threads, with_batch_creation=True)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': _config(args.gpus),
        'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': 'full training step (fwd+MAE+bwd+Adam) on a %d-sample slice of the '
                                   'batch-64 workload per step; torch-CPU fp32 restatement of the '
                                   'reference graph (TF/Keras not installable offline)' % sample_b},
        'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'cpu_end_to_end': {'value': rate_e2e, 'unit': UNIT, 'ms_per_step': sec_e2e * 1e3,
                           'what': 'the same step including the per-step host batch creation (create_batch_hr_lr: slicing + '
                                   'cv2 INTER_AREA), BASELINE.md section 3 "end-to-end" row'},
        'gpu_launches': 0,
    }
    _emit(line)


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return self
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons = [], 0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        rows = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        for _, line in rows:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx or None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# --------------------------------------------------------------------------------------------
# GPU leg
# --------------------------------------------------------------------------------------------
def run_native(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from dl4ds_b200 import _lib, training

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback on the product path)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    _lib.load()

    # the reference's public entry point; data_* are only used for shapes here (bench drives the
    # step directly so that the timed region is exactly K steps)
    rng = np.random.default_rng(1234 + rank)
    n_pool = 4
    hr_host = rng.standard_normal((n_pool * BATCH, HR_HW, HR_HW, 1), dtype=np.float32)
    trainer = training.SupervisedTrainer(
        'resnet', 'spc', hr_host, hr_host[:BATCH], hr_host[:BATCH], scale=SCALE, batch_size=BATCH,
        loss='mae', epochs=1, learning_rate=1e-3, device='GPU', verbose=False, save=False,
        show_plot=False, math=args.math, seed=0)
    trainer.setup_model()
    step = trainer.train_step          # SupervisedStep (captured CUDA graph)
    model = trainer.model

    # ---------------- device-resident pool (value) ----------------
    hr_dev = torch.from_numpy(hr_host).to(dev)
    from dl4ds_b200.step import coarsen_on_device
    lr_dev = coarsen_on_device(hr_dev, SCALE)
    pool = [(lr_dev[i * BATCH:(i + 1) * BATCH], hr_dev[i * BATCH:(i + 1) * BATCH]) for i in range(n_pool)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(i):
        lr_b, hr_b = pool[i % n_pool]
        step.load_batch([lr_b], hr_b)      # D2D into the graph's static buffers
        return step.run()

    for i in range(args.warmup):
        one_step(i)
    barrier()
    sampler = ClockSampler(local).start() if rank == 0 else None
    time.sleep(0.25)
    t_wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = one_step(i)
    e1.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    final_loss = float(loss.item())

    # ---------------- e2e through the trainer API with host batches ----------------
    lr_host = lr_dev.cpu().numpy()
    host_pool = [(lr_host[i * BATCH:(i + 1) * BATCH], hr_host[i * BATCH:(i + 1) * BATCH]) for i in range(n_pool)]
    def host_batches(n):
        for i in range(n):
            yield [host_pool[i % n_pool][0]], host_pool[i % n_pool][1]
    for lval in trainer.train_on_batches(host_batches(max(3, args.warmup))):
        pass
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_e2e0 = time.perf_counter()
    f0.record()
    n_loss = 0
    for lval in trainer.train_on_batches(host_batches(args.steps)):     # one float loss per step (D2H)
        n_loss += 1
    f1.record()
    barrier()
    assert n_loss == args.steps
    ms_e2e = max(f0.elapsed_time(f1), (time.perf_counter() - t_e2e0) * 1e3)    # device events vs host wall clock: the larger
    t_wall2 = time.perf_counter()
    if sampler is not None:
        sampler.stop()

    # ---------------- forward only: Predictor path (host LR in, host HR out) ----------------
    lr_all = np.concatenate([hp[0] for hp in host_pool], axis=0)          # n_pool batches of LR tiles
    model.predict([lr_all], batch_size=BATCH)                              # warm-up + graph capture
    barrier()
    t_p0 = time.perf_counter()
    n_pred_rounds = 4
    for _ in range(n_pred_rounds):
        y_pred = model.predict([lr_all], batch_size=BATCH)
    torch.cuda.synchronize()
    pred_s = (time.perf_counter() - t_p0) / n_pred_rounds
    pred_px = y_pred.shape[0] * HR_HW * HR_HW

    # ---------------- per-kernel durations (roofline) ----------------
    timers = {}
    n_prof = 3
    for i in range(n_prof):
        step.load_batch([pool[i % n_pool][0]], pool[i % n_pool][1])
        step.run_profiled(timers)
    torch.cuda.synchronize()
    from dl4ds_b200.engine import Ctx
    reps = Ctx.TIMER_REPS       # every timed interval holds `reps` back-to-back launches of the same kernel
    per_label = {k: (sum(a.elapsed_time(b) for a, b in v) / (n_prof * reps), len(v) // n_prof) for k, v in timers.items()}

    # max over ranks
    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    peaks = {}
    peaks_src = 'fallback'
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
        peaks_src = 'measured'
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    bf16_peak = float(peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops', 1590.0)))
    tf32_peak = bf16_peak / 2.0         # kind::tf32 issues at half the kind::f16 rate
    # --math f16x3: the forward / input-gradient launches issue kind::f16 MMAs (the full bf16-class rate); the weight
    # gradients stay on kind::tf32
    fwd_peak = bf16_peak if args.math == 'f16x3' else tf32_peak

    # ---------------- BASELINE configs 3 / 4 / 5 (after the headline's timed regions; every rank takes part) ----------------
    n_launch_headline = int(step.launches_per_step)
    extra = {}
    if args.configs:
        del trainer.train_step, step
        torch.cuda.empty_cache()
        try:
            extra = run_extra_configs(args, dev, rank, world, tf32_peak)
        except Exception as e:      # the headline line must survive a failure of an extra configuration
            extra = {'error': '%s: %s' % (type(e).__name__, e)}

    if rank == 0:

        px_per_step = BATCH * HR_HW * HR_HW * world
        value = px_per_step * args.steps / (ms * 1e-3)
        e2e = px_per_step * args.steps / (ms_e2e * 1e-3)

        # dominant conv-family kernel group: label = '<layer>:<pass>@HxW'
        macs = trainer.layer_macs(BATCH)            # {label: MACs of ONE launch}
        top = max(((k, v) for k, v in per_label.items() if macs.get(k, 0) > 0), key=lambda kv: kv[1][0])
        label, (ms_lab, n_launch) = top
        flops_launch = 2.0 * macs.get(label, 0)
        avg_ms = ms_lab / max(n_launch, 1)
        achieved = flops_launch / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
        conv_ms = sum(v[0] for v in per_label.values())
        step_ms = ms / args.steps
        algo_flops_step = 2.0 * trainer.train_macs_per_sample() * BATCH
        # operator composition (last sub-pixel stage x TransitionLast) executes fewer MACs than the reference graph
        # specifies: fwd, dgrad and wgrad each save `macs_saved_per_sample`
        exec_flops_step = algo_flops_step - 2.0 * 3 * model.macs_saved_per_sample * BATCH
        traffic = None
        try:        # per-launch DRAM bytes of the dominant kernel from the committed ncu --set full capture
            with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
                traffic = json.load(f).get(label)
        except Exception:
            pass
        def label_peak(lab):
            """Dense peak of the MMA kind that serves a launch: kind::f16 for the weight gradients conv_tc_wgrad3 takes
            (3x3 / 5x5, 32 <= Cin, Cout <= 64, W % 32 == 0, taps * pad16(Cout) <= 512: conv_tc_wgrad3.cu) and, with
            --math f16x3, for the forward / input-gradient launches; kind::tf32 (half that rate) otherwise."""
            lay, rest = lab.rsplit(':', 1)
            pas, hw = rest.split('@')
            if pas != 'wgrad':
                return fwd_peak
            try:
                w_ = int(hw.split('x')[1])
                if '*' in lay:
                    a_, b_ = lay.split('*')
                    k_, _, cin, _ = model.spec[a_ + '/kernel']
                    cout = 4 * model.spec[b_ + '/kernel'][3]
                else:
                    k_, _, cin, cout = model.spec[lay + '/kernel']
                taps = k_ * k_
                if k_ in (3, 5) and 32 <= cin <= 64 and 32 <= cout <= 64 and w_ % 32 == 0 and taps * ((cout + 15) // 16 * 16) <= 512:
                    return bf16_peak
            except Exception:
                pass
            return tf32_peak
        lab_peak = label_peak(label)
        roofline = {
            'bound': 'tensor', 'achieved': achieved, 'peak': lab_peak, 'unit': 'TFLOP/s',
            'frac': achieved / lab_peak, 'traffic': traffic,
            'kernel': label, 'avg_launch_ms': avg_ms, 'launches_per_step': n_launch,
            'share_of_step': ms_lab / max(sum(v[0] for v in per_label.values()), 1e-9),
            'peak_source': '%s: bf16_tflops_sustained%s' % (peaks_src, '/2 (tcgen05 kind::tf32 issue rate)'
                                                            if lab_peak == tf32_peak else ' (tcgen05 kind::f16 issue rate)'),
            'math': args.math,
            'step_conv_roofline_frac': (algo_flops_step / (step_ms * 1e-3) / 1e12) / tf32_peak,
            'step_algorithmic_tflops': algo_flops_step / (step_ms * 1e-3) / 1e12,
            'step_executed_tflops': exec_flops_step / (step_ms * 1e-3) / 1e12,
            'note': 'achieved = executed FLOPs of the named launch (2*MACs; tf32x3 issues 2-3 MMAs per MAC on top); '
                    'step_algorithmic_* counts the reference graph (220 GFLOP/step), step_executed_* what the '
                    'composed sub-pixel x TransitionLast kernels really run; `peak` is the dense peak of the MMA kind of the named '
                    'launch (conv_tc_wgrad3: kind::f16 = the bf16 figure; kind::tf32 = half of it), step_conv_roofline_frac is '
                    'quoted against the kind::tf32 peak',
            'conv_family_ms_per_step_eager': conv_ms,
            'hbm_peak_gbs': hbm_peak,
        }
        # per-FUNCTION view next to the per-layer one: every launch of a pass type summed (the tensor-core layers of the
        # forward / input-gradient passes run conv_tc_halo_kernel, those of the weight-gradient pass conv_tc_wgrad2_kernel;
        # the 1- and 8-channel layers of the same passes run the thin / pointwise kernels and are included in the sums)
        groups = {}
        for k, (ms_k, n_k) in per_label.items():
            pas = k.rsplit(':', 1)[1].split('@')[0]
            g = 'wgrad (conv_tc_wgrad3 / wgrad2 kernels + thin)' if pas == 'wgrad' else 'fwd+dgrad (conv_tc_halo_kernel + thin)'
            a_ = groups.setdefault(g, [0.0, 0.0, 0, 0.0])
            a_[0] += ms_k
            a_[1] += 2.0 * macs.get(k, 0) * n_k
            a_[2] += n_k
            a_[3] += 2.0 * macs.get(k, 0) * n_k / (label_peak(k) * 1e12)      # seconds at the peak of each launch's MMA kind
        roofline['per_function'] = {
            g: {'ms_per_step': v[0], 'launches_per_step': v[2], 'achieved': v[1] / (v[0] * 1e-3) / 1e12 if v[0] > 0 else 0.0,
                'frac': v[3] / (v[0] * 1e-3) if v[0] > 0 else 0.0,
                'share_of_conv_family': v[0] / max(conv_ms, 1e-9)}
            for g, v in groups.items()}

        # bounded CPU sample: 10-30 s of host work
        threads = os.cpu_count() or 1
        cpu_rate, cpu_sec = cpu_oracle_rate(4, 8, 2, threads) if not args.no_cpu else (None, None)
        cpu_e2e, _ = cpu_oracle_rate(4, 4, 1, threads, with_batch_creation=True) if not args.no_cpu else (None, None)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': step_ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': _config(world, {'math': args.math}),
            'e2e': {'value': e2e, 'unit': UNIT,
                    'h2d_bytes_per_step': int(BATCH * (LR_HW * LR_HW + HR_HW * HR_HW) * 4),
                    'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / args.steps,
                    'api': 'SupervisedTrainer.train_on_batches(host numpy (LR, HR) batches) -> one float loss per step '
                           '(the fit loop of SupervisedTrainer.run: pinned staging + H2D of batch i+1 overlap step i)'},
            'predict': {'value': pred_px * world / pred_s, 'unit': UNIT, 'ms_per_batch': pred_s / n_pool * 1e3,
                        'api': 'Model.predict(host LR array, batch_size=64) -> host HR array (forward only, captured '
                               'graph per chunk, H2D + D2H included; what Predictor.run calls, inference.py:238)'},
            'gpu_launches': int(n_launch_headline * args.steps),
            'launches_per_step': n_launch_headline,
            'roofline': roofline,
            'cpu_baseline': None if cpu_rate is None else {
                'value': cpu_rate, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                'sample': '8 timed training steps at batch 4 (configs[0]) of the torch-CPU fp32 '
                          'restatement of the reference graph; TF/Keras not installable offline',
                'end_to_end_value': cpu_e2e,
                'end_to_end_what': 'the same step including the per-step host batch creation (create_batch_hr_lr)'},
            'clocks': sampler.summary(t_wall0, t_wall2) if sampler is not None else None,
            'final_loss': final_loss,
            'top_kernels_ms_per_step': dict(sorted(((k, round(v[0], 4)) for k, v in per_label.items()),
                                                   key=lambda kv: -kv[1])[:8]),
            'kernels_ms_per_step': dict(sorted(((k, round(v[0], 4)) for k, v in per_label.items()), key=lambda kv: -kv[1])),
            'configs': extra,
        }
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
# BASELINE.json configs[2..4] at their per-GPU sizes (weak scaling: the stated global batches 128 / 32 / 32 are
# 16 / 8 / 4 per GPU at the stated 8 / 4 / 8 GPUs).  Run after the headline's timed region; the headline keys of the
# JSON line are unaffected.  Every entry: device-resident captured-graph step time (CUDA events, max over ranks),
# the same through the trainer API with host batches (e2e), algorithmic FLOPs (2 * MACs of the reference graph:
# fwd + dgrad + wgrad) and its fraction of the kind::tf32 tensor peak.
# --------------------------------------------------------------------------------------------
def _max_over_ranks(vals, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def _time_steps(fn, n, warm, dev, world):
    """ms per step of fn(i): warm-up, barrier, CUDA events around n calls, max over ranks."""
    import torch
    import torch.distributed as dist
    for i in range(warm):
        fn(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    if world > 1:
        dist.barrier()
    ms = max(e0.elapsed_time(e1), 0.0)
    ms, wall = _max_over_ranks([ms, wall], dev, world)
    return ms / n, max(ms, wall) / n


def _supervised_config(tag, workload, trainer, hr_px_per_sample, batch, args, dev, world, tf32_peak, stated_gpus):
    import numpy as np
    import torch
    trainer.setup_datagen()
    trainer.setup_model()
    st = trainer.train_step
    batches = [trainer.ds_train[i % len(trainer.ds_train)] for i in range(2)]
    fold = trainer.model_is_spatiotemporal

    def fix(x):          # App. B #2: spatio-temporal LR batches without predictors come out without a channel axis
        x = [np.asarray(a, np.float32) for a in x]
        if fold and x[0].ndim == 4:
            x[0] = x[0][..., None]
        return x
    batches = [(fix(x), np.asarray(y[0], np.float32)) for x, y in batches]
    # device-resident pool: stage both batches once in the layout of the static buffers
    pool = []
    for x, y in batches:
        trainer._stage(x, y)
        torch.cuda.synchronize()
        pool.append(([t.clone() for t in st.inputs], st.target.clone()))

    def dev_step(i):
        ins, tgt = pool[i % len(pool)]
        st.load_batch(ins, tgt)
        return st.run()
    n, warm = max(10, args.steps // 2), max(3, args.warmup)
    ms_dev, _ = _time_steps(dev_step, n, warm, dev, world)

    def host_epoch(k):
        for _ in trainer.train_on_batches((batches[i % len(batches)] for i in range(k))):
            pass
    host_epoch(warm)
    _, ms_e2e = _time_steps(lambda i: None if i else host_epoch(n), 1, 0, dev, world)
    ms_e2e /= n
    px = batch * hr_px_per_sample * world
    flops = 2.0 * trainer.train_macs_per_sample() * batch
    h2d = sum(int(np.prod(t.shape)) for t in st.inputs) * 4 + int(np.prod(st.target.shape)) * 4
    return {
        'workload': workload, 'per_gpu_batch': batch, 'n_gpus': world, 'stated_gpus': stated_gpus,
        'ms_per_step': ms_dev, 'value': px / (ms_dev * 1e-3), 'unit': UNIT,
        'e2e': {'value': px / (ms_e2e * 1e-3), 'unit': UNIT, 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': 4, 'api': 'SupervisedTrainer.train_on_batches(host numpy batches)'},
        'params': trainer.model.count_params(), 'launches_per_step': int(st.launches_per_step),
        'roofline': {'bound': 'tensor', 'unit': 'TFLOP/s', 'peak': tf32_peak,
                     'achieved': flops / (ms_dev * 1e-3) / 1e12, 'frac': flops / (ms_dev * 1e-3) / 1e12 / tf32_peak,
                     'algorithmic_gflop_per_step': flops / 1e9,
                     'note': 'whole-step algorithmic FLOPs (2 * MACs of the reference graph, fwd + dgrad + wgrad) / step time'},
    }


def run_extra_configs(args, dev, rank, world, tf32_peak):
    import numpy as np
    import torch
    from dl4ds_b200 import nets, training
    from dl4ds_b200.training import cgan
    out = {}
    rng = np.random.default_rng(4321 + rank)
    math = args.math
    want = set(args.configs.split(','))
    # ---- cfg3: densenet + channel attention + LCB, 8x deconvolution, 3 predictors + 1 static, LR 16 -> HR 128
    if 'cfg3' in want:
        B = 16
        hr = rng.standard_normal((2 * B, 128, 128, 1), dtype=np.float32)
        preds = [rng.standard_normal((2 * B, 128, 128, 1), dtype=np.float32) for _ in range(3)]
        static = rng.standard_normal((128, 128)).astype(np.float32)
        tr = training.SupervisedTrainer('densenet', 'dc', hr, hr[:B], hr[:B], predictors_train=preds,
                                        predictors_val=[q[:B] for q in preds], predictors_test=[q[:B] for q in preds],
                                        static_vars=[static], scale=8, batch_size=B, epochs=1, learning_rate=1e-3,
                                        verbose=False, save=False, math=math, seed=1, attention=True, localcon_layer=True)
        out['cfg3'] = _supervised_config(
            'cfg3', 'SupervisedTrainer step: densenet + ChannelAttention + LCB, 8x deconv, 3 predictors + 1 static, '
                    'LR 16 -> HR 128, batch 16 per GPU (128 over 8 GPUs), MAE, Adam', tr, 128 * 128, B, args, dev, world,
            tf32_peak, 8)
        del tr
        torch.cuda.empty_cache()
    # ---- cfg4: recurrent (ConvLSTM) resnet, 4x resize-convolution, T = 6, 32 -> 128
    if 'cfg4' in want:
        B, T = 8, 6
        hr = rng.standard_normal((2 * B + T, 128, 128, 1), dtype=np.float32)
        tr = training.SupervisedTrainer('resnet', 'rc', hr, hr[:B + T], hr[:B + T], scale=4, time_window=T, batch_size=B,
                                        epochs=1, learning_rate=1e-3, verbose=False, save=False, math=math, seed=1)
        out['cfg4'] = _supervised_config(
            'cfg4', 'SupervisedTrainer step: recurrent ConvLSTM resnet, 4x resize-conv, seq_len 6, 32 -> 128, '
                    'batch 8 per GPU (32 over 4 GPUs), MAE, Adam', tr, T * 128 * 128, B, args, dev, world, tf32_peak, 4)
        del tr
        torch.cuda.empty_cache()
    # ---- cfg5: CGANTrainer step, U-Net generator (pin) + residual discriminator, 256 x 256
    if 'cfg5' in want:
        import torch.distributed as dist
        B, hw = 4, 256
        hr = rng.standard_normal((2 * B, hw, hw, 1), dtype=np.float32)
        static = rng.standard_normal((hw, hw)).astype(np.float32)
        G = nets.unet_pin('unet', 2, 1, (hw, hw), 1, 8, 6, math=math).to(dev).init_weights(seed=1)
        D = nets.residual_discriminator(2, 'pin', False, 4, (hw, hw), n_filters=8, n_res_blocks=4, math=math).to(dev).init_weights(seed=2)
        st_b = np.broadcast_to(static[None, :, :, None], (B, hw, hw, 1)).astype(np.float32).copy()
        lrs = [np.concatenate([hr[i * B:(i + 1) * B], st_b], axis=-1).astype(np.float32) for i in range(2)]
        hrs = [hr[i * B:(i + 1) * B] for i in range(2)]
        d = dist if world > 1 else None
        step = cgan.CGANStep(G, D, lrs[0].shape, hrs[0].shape, st_b.shape, dist=d).capture()
        step.run(lrs[0], hrs[0], st_b, first_batch=True)
        lrd = [torch.from_numpy(a).to(dev) for a in lrs]
        hrd = [torch.from_numpy(a).to(dev) for a in hrs]
        std = torch.from_numpy(st_b).to(dev)
        n, warm = max(10, args.steps // 2), max(3, args.warmup)

        def dev_step(i):        # device-resident arrays; the loss read (D2H of 4 floats) stays, as train_step returns floats
            step.run(lrd[i % 2], hrd[i % 2], std)
        ms_dev, _ = _time_steps(dev_step, n, warm, dev, world)
        _, ms_e2e = _time_steps(lambda i: step.run(lrs[i % 2], hrs[i % 2], st_b), n, warm, dev, world)
        FG, FGd = G.macs_per_sample, G.macs_dgrad_per_sample
        FD, FDd = D.macs_per_sample, D.macs_dgrad_per_sample
        # G: fwd + dgrad + wgrad; D: two forward passes, weight + input gradients through both, one more input-gradient
        # pass through D(fake) for the generator loss (the reference's two tapes, cgan.py:587-611)
        macs = (2 * FG + FGd) + (2 * FD + 2 * (FD + FDd) + FDd)
        flops = 2.0 * macs * B
        px = B * hw * hw * world
        out['cfg5'] = {
            'workload': 'CGANTrainer train_step: U-Net generator (pin, n_filters 8, n_blocks 6) + residual discriminator, '
                        '256 x 256, 1 static variable, batch 4 per GPU (32 over 8 GPUs), MAE + BCE, two Adam(beta_1 0.5)',
            'per_gpu_batch': B, 'n_gpus': world, 'stated_gpus': 8,
            'ms_per_step': ms_dev, 'value': px / (ms_dev * 1e-3), 'unit': UNIT,
            'e2e': {'value': px / (ms_e2e * 1e-3), 'unit': UNIT, 'ms_per_step': ms_e2e,
                    'h2d_bytes_per_step': int(lrs[0].nbytes + hrs[0].nbytes + st_b.nbytes), 'd2h_bytes_per_step': 16,
                    'api': 'CGANStep.run(host numpy LR / HR / static arrays) -> 4 float losses (what CGANTrainer.run loops over)'},
            'params': G.count_params() + D.count_params(),
            'roofline': {'bound': 'tensor', 'unit': 'TFLOP/s', 'peak': tf32_peak,
                         'achieved': flops / (ms_dev * 1e-3) / 1e12, 'frac': flops / (ms_dev * 1e-3) / 1e12 / tf32_peak,
                         'algorithmic_gflop_per_step': flops / 1e9,
                         'note': 'whole-step algorithmic FLOPs (generator fwd+bwd, discriminator 2 fwd + 2 bwd + 1 dgrad) / step time'},
        }
    return out


_REAL_STDOUT = None


def _stdout_to_stderr():
    """Keep stdout for the ONE JSON line: anything a library writes to fd 1 during the run (NCCL prints its
    version banner there under torchrun) goes to stderr instead."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)


def _emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--math', default=os.environ.get('DL4DS_MATH', 'auto'),
                    choices=['auto', 'fp32', 'tf32x3', 'tf32', 'f16x3'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline sample')
    ap.add_argument('--configs', default='cfg3,cfg4,cfg5',
                    help="BASELINE configs measured after the headline ('' = none)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    _stdout_to_stderr()
    if args.impl == 'reference':
        run_reference(args)
    else:
        if args.math == 'auto':
            args.math = 'tf32x3'       # tensor cores with the 3-term split: fp32-level parity
        run_native(args)


if __name__ == '__main__':
    main()
