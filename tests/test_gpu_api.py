"""GPU tests of the reference-facing API (SupervisedTrainer / CGANTrainer / Predictor) and of the
training steps against the oracle (oracle/torch_ref.py)."""
import numpy as np
import pytest
import torch

from dl4ds_b200 import CGANTrainer, Predictor, SupervisedTrainer, nets
from dl4ds_b200.training import cgan
from oracle import torch_ref as R

pytestmark = pytest.mark.gpu


def _data(n, hw, seed=0):
    return np.random.default_rng(seed).standard_normal((n, hw, hw, 1)).astype(np.float32)


@pytest.mark.parametrize('math', ['fp32', 'tf32x3'])
def test_supervised_step_sequence_matches_oracle(cuda, math):
    """Five optimizer steps (fwd + MAE + bwd + Keras-Adam with PiecewiseConstantDecay) through
    SupervisedTrainer.train_on_batch == the oracle's supervised_step, loss by loss."""
    np.random.seed(0)                      # the DataGenerator permutation is drawn from numpy's global RNG
    hr = _data(24, 64)
    tr = SupervisedTrainer('resnet', 'spc', hr, hr[:8], hr[:8], scale=4, batch_size=8, epochs=1,
                           learning_rate=(1e-3, 1e-4), lr_decay_after=3, verbose=False, math=math, seed=7,
                           n_blocks=2)
    tr.setup_datagen()
    tr.setup_model()
    w = {k: torch.from_numpy(v.copy()) for k, v in tr.model.get_weights().items()}
    opt = R.TFAdam(list(w), lr=R.piecewise_constant(3, 1e-3, 1e-4))
    fwd = lambda p, xs: R.net_postupsampling(p, xs, 'resnet', 'spc', 4, n_blocks=2)
    for i in range(5):
        (lr,), (y,) = tr.ds_train[i % len(tr.ds_train)]
        loss = tr.train_on_batch([lr], y)
        ref, _ = R.supervised_step(fwd, w, opt, [torch.from_numpy(lr)], torch.from_numpy(y))
        assert abs(loss - ref) <= 2e-4 * max(1.0, abs(ref)), (i, loss, ref)
    from tests.util import assert_adam_weights_close
    # (in the tensor-core mode ReLU masks flip on ~1e-6 pre-activation differences: a looser agreement bound)
    assert_adam_weights_close(tr.model.get_weights(), {k: v.numpy() for k, v in w.items()}, lr=1e-3, steps=5,
                              tight=5e-5 if math == 'fp32' else 5e-4, frac=5e-3 if math == 'fp32' else 5e-2)


def test_supervised_run_and_predictor(cuda, tmp_path):
    hr = _data(40, 32, 1)
    tr = SupervisedTrainer('resnet', 'spc', hr[:24], hr[24:32], hr[32:], scale=4, batch_size=8, epochs=3,
                           learning_rate=1e-3, verbose=False, seed=3, n_blocks=2, save=True,
                           save_path=str(tmp_path))
    tr.run()
    h = tr.fithist.history
    assert len(h['loss']) == 3 and len(h['val_loss']) == 3 and np.isfinite(tr.test_loss)
    assert h['loss'][-1] < h['loss'][0]                       # it learns
    assert tr.model.name == 'resnet_spc'
    out = Predictor(tr, hr[32:], scale=4, array_in_hr=True, batch_size=3).run()
    assert out.shape == (8, 32, 32, 1) and out.dtype == np.float32
    # same result as one big forward
    lr = hr[32:].reshape(8, 8, 4, 8, 4, 1).mean(axis=(2, 4)).astype(np.float32)
    full = tr.model.predict([lr], batch_size=8)
    assert np.abs(out - full).max() <= 1e-6
    out2, lr2 = Predictor(tr, lr, scale=4, array_in_hr=False, return_lr=True).run()
    assert out2.shape == (8, 32, 32, 1) and lr2.shape == (8, 8, 8, 1)


def test_cgan_step_matches_oracle(cuda):
    """train_step (cgan.py:575-639): losses and BOTH weight updates after one step == oracle."""
    rng = np.random.default_rng(11)
    B, hw = 3, 16
    G = nets.unet_pin('unet', 2, 1, (hw, hw), 1, 4, 2, math='fp32').to(cuda)
    D = nets.residual_discriminator(2, 'pin', False, 4, (hw, hw), n_filters=4, n_res_blocks=1, math='fp32').to(cuda)
    gw = R.init_weights(G.spec, seed=1, bias_scale=0.05)
    dw = R.init_weights(D.spec, seed=2, bias_scale=0.05)
    G.set_weights({k: v.numpy() for k, v in gw.items()})
    D.set_weights({k: v.numpy() for k, v in dw.items()})
    lr = rng.standard_normal((B, hw, hw, 2)).astype(np.float32)
    hr = rng.standard_normal((B, hw, hw, 1)).astype(np.float32)
    st = rng.standard_normal((B, hw, hw, 1)).astype(np.float32)
    nfeat = D.spec['dense1/kernel'][0]
    masks = [(rng.random((B, 1, 1, nfeat)) < 0.6).astype(np.float32) / 0.6 for _ in range(2)]
    losses = cgan.train_step(lr, hr, G, D, cgan.Adam(2e-4, beta_1=0.5), cgan.Adam(2e-4, beta_1=0.5),
                             gen_pxloss_function='mae', static_array=st, dropout_masks=masks)
    gen_fn = lambda p, xs: R.unet_pin(p, xs, 4, 2)
    disc_fn = lambda p, xs, m: R.residual_discriminator(p, xs, 'pin', 4, (hw, hw), n_filters=4, n_res_blocks=1,
                                                        dropout_mask=m)
    gopt, dopt = R.TFAdam(list(gw), lr=2e-4, beta_1=0.5), R.TFAdam(list(dw), lr=2e-4, beta_1=0.5)
    ref, _, _ = R.cgan_step(gen_fn, disc_fn, gw, dw, gopt, dopt, torch.from_numpy(lr), torch.from_numpy(hr),
                            torch.from_numpy(st), mask_real=torch.from_numpy(masks[0].reshape(B, nfeat)),
                            mask_fake=torch.from_numpy(masks[1].reshape(B, nfeat)))
    for a, b in zip(losses, ref):
        assert abs(a - b) <= 1e-4 * max(1.0, abs(b)), (losses, ref)
    from tests.util import assert_adam_weights_close
    for model, w in ((G, gw), (D, dw)):
        assert_adam_weights_close(model.get_weights(), {k: v.numpy() for k, v in w.items()}, lr=2e-4, steps=1,
                                  tight=2e-5)


@pytest.mark.parametrize('graph', [False, True])
def test_cgan_step_spatiotemporal_matches_oracle(cuda, graph):
    """train_step on 5-D arrays (cgan.py:575-639 with a recurrent generator and the spatio-temporal discriminator):
    losses and both weight updates after one step == oracle; eagerly and as the captured CGANStep."""
    rng = np.random.default_rng(12)
    B, T, hw = 2, 3, 8
    G = nets.recnet_postupsampling('resnet', 'rc', 4, 1, 1, (hw, hw), T, n_filters=4, n_blocks=1, math='fp32').to(cuda)
    D = nets.residual_discriminator(1, 'rc', True, 4, (hw, hw), n_filters=4, n_res_blocks=1, math='fp32',
                                    time_window=T).to(cuda)
    gw = R.init_weights(G.spec, seed=1, bias_scale=0.05)
    dw = R.init_weights(D.spec, seed=2, bias_scale=0.05)
    G.set_weights({k: v.numpy() for k, v in gw.items()})
    D.set_weights({k: v.numpy() for k, v in dw.items()})
    lr = rng.standard_normal((B, T, hw, hw, 1)).astype(np.float32)
    hr = rng.standard_normal((B, T, 4 * hw, 4 * hw, 1)).astype(np.float32)
    st = rng.standard_normal((B, 4 * hw, 4 * hw, 1)).astype(np.float32)
    nfeat = D.spec['dense1/kernel'][0]
    masks = [(rng.random((B, 1, 1, nfeat)) < 0.6).astype(np.float32) / 0.6 for _ in range(2)]
    if graph:
        step = cgan.CGANStep(G, D, lr.shape, hr.shape, st.shape, learning_rates=(2e-4, 2e-4)).capture()
        losses = step.run(lr, hr, st, dropout_masks=masks)
    else:
        losses = cgan.train_step(lr, hr, G, D, cgan.Adam(2e-4, beta_1=0.5), cgan.Adam(2e-4, beta_1=0.5),
                                 gen_pxloss_function='mae', static_array=st, dropout_masks=masks)
    gen_fn = lambda p, xs: R.recnet_postupsampling(p, xs, 'resnet', 'rc', 4, T, n_filters=4, n_blocks=1)
    disc_fn = lambda p, xs, m: R.residual_discriminator(p, xs, 'rc', 4, (hw, hw), n_filters=4, n_res_blocks=1,
                                                        dropout_mask=m, is_spatiotemporal=True)
    gopt, dopt = R.TFAdam(list(gw), lr=2e-4, beta_1=0.5), R.TFAdam(list(dw), lr=2e-4, beta_1=0.5)
    ref, _, _ = R.cgan_step(gen_fn, disc_fn, gw, dw, gopt, dopt, torch.from_numpy(lr), torch.from_numpy(hr),
                            torch.from_numpy(st), mask_real=torch.from_numpy(masks[0].reshape(B, nfeat)),
                            mask_fake=torch.from_numpy(masks[1].reshape(B, nfeat)))
    for a, b in zip(losses, ref):
        assert abs(a - b) <= 1e-4 * max(1.0, abs(b)), (losses, ref)
    from tests.util import assert_adam_weights_close
    for model, w in ((G, gw), (D, dw)):
        assert_adam_weights_close(model.get_weights(), {k: v.numpy() for k, v in w.items()}, lr=2e-4, steps=1,
                                  tight=2e-5)


def test_cgan_trainer_spatiotemporal_runs(cuda, tmp_path):
    """CGANTrainer with time_window > 1 (cgan.py:179-185,208-231): recurrent generator + 5-D discriminator."""
    hr = _data(14, 32, 5)
    static = np.random.default_rng(6).standard_normal((32, 32)).astype(np.float32)
    tr = CGANTrainer('resnet', 'rc', hr[:10], hr[10:], scale=4, batch_size=2, epochs=2, static_vars=[static],
                     generator_params=dict(n_filters=4, n_blocks=1),
                     discriminator_params=dict(n_filters=4, n_res_blocks=1), verbose=False, seed=2, time_window=2,
                     save_loss_history=False, save_path=str(tmp_path))
    tr.run()
    assert len(tr.gentotal) == 2 and all(np.isfinite(v) for v in tr.gentotal + tr.disc)
    assert np.isfinite(tr.test_loss) and tr.generator.name == 'recresnet_rc'


def test_cgan_trainer_runs(cuda, tmp_path):
    hr = _data(12, 32, 5)
    static = np.random.default_rng(6).standard_normal((32, 32)).astype(np.float32)
    tr = CGANTrainer('unet', 'pin', hr[:8], hr[8:], scale=4, batch_size=4, epochs=2, static_vars=[static],
                     generator_params=dict(n_filters=4, n_blocks=2, n_channels_out=1),
                     discriminator_params=dict(n_filters=4, n_res_blocks=1), verbose=False, seed=2, time_window=None,
                     save_loss_history=False, save_path=str(tmp_path))
    tr.run()
    assert len(tr.gentotal) == 2 and all(np.isfinite(v) for v in tr.gentotal + tr.disc)
    assert np.isfinite(tr.test_loss) and tr.generator.name == 'unet_pin'


# ------------------------------------------------------------------------------------------ device-resident data path
@pytest.mark.parametrize('patch', [None, 32])
def test_device_data_generator_matches_host(cuda, patch):
    """DeviceDataGenerator (gather/crop + block-mean coarsening on the GPU) == the host DataGenerator (numpy
    slicing + cv2 INTER_AREA, itself pinned bit-for-bit to the reference by tests/test_datapath.py) for the same
    numpy RNG state: same permutation, same crops, same batches."""
    from dl4ds_b200.dataloader import DataGenerator, DeviceDataGenerator
    hr = np.random.default_rng(5).standard_normal((20, 48, 64, 2)).astype(np.float32)
    kw = dict(backbone='resnet', upsampling='spc', scale=4, batch_size=6, patch_size=patch)
    np.random.seed(11)
    host = DataGenerator(hr, None, **kw)
    hb = [host[i] for i in range(len(host))]
    np.random.seed(11)
    dev = DeviceDataGenerator(hr, None, device=cuda, **kw)
    assert len(dev) == len(host) == 3
    for i in range(len(dev)):
        (lr_d,), (hr_d,) = dev[i]
        (lr_h,), (hr_h,) = hb[i]
        assert hr_d.shape == hr_h.shape and lr_d.shape == lr_h.shape
        assert np.array_equal(hr_d, hr_h)
        assert np.abs(lr_d - lr_h).max() <= 1e-6 * max(1.0, np.abs(lr_h).max())


@pytest.mark.parametrize('case', ['pred_hr+static', 'pred_lr', 'static+patch'])
def test_device_data_generator_with_predictors_and_static(cuda, case):
    """Predictors (on the HR or the LR grid) and static variables assembled on the GPU into their channel slices
    of the LR batch (+ the static HR planes as the aux input) == the host generator (BASELINE config 3's data path:
    3 predictors + 1 static field)."""
    from dl4ds_b200.dataloader import DataGenerator, DeviceDataGenerator
    rng = np.random.default_rng(6)
    hr = rng.standard_normal((14, 32, 48, 1)).astype(np.float32)
    kw = dict(backbone='densenet', upsampling='dc', scale=4, batch_size=4)
    if case == 'pred_hr+static':
        kw.update(predictors=[rng.standard_normal((14, 32, 48, 1)).astype(np.float32) for _ in range(3)],
                  static_vars=[rng.standard_normal((32, 48)).astype(np.float32)])
    elif case == 'pred_lr':
        kw.update(predictors=[rng.standard_normal((14, 8, 12, 2)).astype(np.float32)])
    else:
        kw.update(static_vars=[rng.standard_normal((32, 48)).astype(np.float32),
                               rng.standard_normal((32, 48, 1)).astype(np.float32)], patch_size=16)
    assert DeviceDataGenerator.supported(hr, None, 'dc', 4, kw.get('patch_size'), None, kw.get('static_vars'),
                                         kw.get('predictors'), 'inter_area')
    np.random.seed(12)
    host = DataGenerator(hr, None, **kw)
    hb = [host[i] for i in range(len(host))]
    np.random.seed(12)
    dev = DeviceDataGenerator(hr, None, device=cuda, **kw)
    for i in range(len(dev)):
        xd, (hr_d,) = dev[i]
        xh, (hr_h,) = hb[i]
        assert len(xd) == len(xh) and np.array_equal(hr_d, hr_h)
        for a, b in zip(xd, xh):
            assert a.shape == b.shape, (a.shape, b.shape)
            assert np.abs(a - b).max() <= 1e-6 * max(1.0, np.abs(b).max())
    # outside the device domain: the reference's patch + predictors branch
    assert not DeviceDataGenerator.supported(hr, None, 'dc', 4, 16, None, None, kw.get('predictors') or [hr],
                                             'inter_area')


def test_supervised_run_cfg3_data_on_device(cuda):
    """BASELINE config 3 shape (densenet + attention + LCB, 8x deconvolution, 3 predictors + 1 static) trained from
    the device-resident data path: the same loss history as from the host generator."""
    rng = np.random.default_rng(8)
    hr = rng.standard_normal((24, 64, 64, 1)).astype(np.float32)
    preds = [rng.standard_normal((24, 64, 64, 1)).astype(np.float32) for _ in range(3)]
    static = rng.standard_normal((64, 64)).astype(np.float32)
    hist = []
    for on_dev in (False, True):
        np.random.seed(4)
        tr = SupervisedTrainer('densenet', 'dc', hr[:16], hr[16:], hr[16:], predictors_train=[p[:16] for p in preds],
                               predictors_val=[p[16:] for p in preds], predictors_test=[p[16:] for p in preds],
                               static_vars=[static.copy()], scale=8, batch_size=4, epochs=2, learning_rate=1e-3,
                               verbose=False, seed=3, n_blocks=2, attention=True, localcon_layer=True, math='fp32',
                               data_on_device=on_dev)
        tr.run()
        from dl4ds_b200.dataloader import DeviceDataGenerator
        assert isinstance(tr.ds_train, DeviceDataGenerator) == on_dev
        hist.append(tr.fithist.history['loss'])
    assert np.allclose(hist[0], hist[1], rtol=2e-4), hist


def test_supervised_run_data_on_device(cuda):
    """SupervisedTrainer(data_on_device=True): the same losses as the host data path, epoch by epoch."""
    hr = _data(40, 32, 2)
    hist = []
    for on_dev in (False, True):
        np.random.seed(4)
        tr = SupervisedTrainer('resnet', 'spc', hr[:24], hr[24:32], hr[32:], scale=4, batch_size=8, epochs=2,
                               learning_rate=1e-3, verbose=False, seed=3, n_blocks=2, data_on_device=on_dev)
        tr.run()
        from dl4ds_b200.dataloader import DeviceDataGenerator
        assert isinstance(tr.ds_train, DeviceDataGenerator) == on_dev
        hist.append(tr.fithist.history['loss'])
    assert np.allclose(hist[0], hist[1], rtol=2e-4), hist


def test_cgan_graph_step_equals_eager_step(cuda):
    """CGANStep (captured CUDA graphs, device-side Adam step sizes) == train_step (eager launches) over three steps
    with the same dropout masks: same four losses, same generator and discriminator weights."""
    rng = np.random.default_rng(21)
    B, hw = 4, 32
    lr = rng.standard_normal((B, hw, hw, 2)).astype(np.float32)
    hr = rng.standard_normal((B, hw, hw, 1)).astype(np.float32)
    st = rng.standard_normal((B, hw, hw, 1)).astype(np.float32)
    results = []
    for graph in (False, True):
        G = nets.unet_pin('unet', 2, 1, (hw, hw), 1, 8, 2, math='tf32x3').to(cuda).init_weights(3)
        D = nets.residual_discriminator(2, 'pin', False, 4, (hw, hw), n_filters=8, n_res_blocks=1, math='tf32x3').to(cuda).init_weights(4)
        nfeat = D.spec['dense1/kernel'][0]
        mrng = np.random.default_rng(5)
        step = cgan.CGANStep(G, D, lr.shape, hr.shape, st.shape).capture() if graph else None
        go, do = cgan.Adam(2e-4, beta_1=0.5), cgan.Adam(2e-4, beta_1=0.5)
        hist = []
        for i in range(3):
            masks = [(mrng.random((B, 1, 1, nfeat)) < 0.6).astype(np.float32) / 0.6 for _ in range(2)]
            if graph:
                hist.append(step.run(lr, hr, st, dropout_masks=masks))
            else:
                hist.append(cgan.train_step(lr, hr, G, D, go, do, gen_pxloss_function='mae', static_array=st,
                                            dropout_masks=masks))
        results.append((np.array(hist), G.get_weights(), D.get_weights()))
    (h0, g0, d0), (h1, g1, d1) = results
    assert np.allclose(h0, h1, rtol=2e-4, atol=1e-6), (h0, h1)
    for a, b in ((g0, g1), (d0, d1)):
        worst = max(float(np.abs(a[k] - b[k]).max()) for k in a)
        assert worst <= 1e-3, worst      # Adam's first steps move weights by ~lr per step; sign flips of ~0 gradients


def test_predict_graphed_equals_eager(cuda):
    """Model.predict: full chunks through the captured forward graph (+ an eager trailing chunk) == one eager
    forward over everything; still correct after the weights change (the graph re-packs the weight images)."""
    m = nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (16, 16), n_blocks=2, math='tf32x3').to(cuda).init_weights(5)
    lr = np.random.default_rng(9).standard_normal((22, 16, 16, 1)).astype(np.float32)
    for rnd in range(2):
        ref = m([lr]).cpu().numpy()                       # eager, one batch of 22
        out = m.predict([lr], batch_size=4)               # 5 graphed chunks + 1 eager chunk of 2
        assert out.shape == (22, 64, 64, 1)
        assert np.abs(out - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
        m.arena.theta.mul_(1.5)                           # change the parameters, predict again


def test_resume_from_checkpoint_continues_the_adam_trajectory(cuda, tmp_path):
    """train k steps, save_checkpoint, load into a fresh model, resume through
    ``trained_model`` -- the resumed losses equal uninterrupted training (Adam m / v / t survive ``Model.to``)."""
    hr = _data(32, 32, 21)
    lr = hr.reshape(32, 8, 4, 8, 4, 1).mean(axis=(2, 4)).astype(np.float32)
    kw = dict(scale=4, batch_size=8, epochs=1, learning_rate=(1e-3, 1e-4), lr_decay_after=4, verbose=False,
              math='tf32x3', n_blocks=2)
    tr = SupervisedTrainer('resnet', 'spc', hr, hr[:8], hr[:8], seed=5, **kw)
    tr.setup_model()
    for i in range(3):
        tr.train_on_batch([lr[8 * i:8 * i + 8]], hr[8 * i:8 * i + 8])
    path = str(tmp_path / 'ck.npz')
    tr.model.save_checkpoint(path)
    ref = [tr.train_on_batch([lr[8 * (i % 4):8 * (i % 4) + 8]], hr[8 * (i % 4):8 * (i % 4) + 8]) for i in range(3, 6)]
    m2 = nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (8, 8), n_blocks=2, math='tf32x3')
    m2.to('cuda').load_checkpoint(path)            # index-less 'cuda', as load_checkpoint / set_weights do
    assert m2.arena.t == 3
    tr2 = SupervisedTrainer('resnet', 'spc', hr, hr[:8], hr[:8], trained_model=m2, trained_epochs=0, seed=99, **kw)
    tr2.setup_model()
    assert tr2.model.arena.t == 3 and float(tr2.model.arena.m.abs().max()) > 0     # slots survived to(cuda:0)
    got = [tr2.train_on_batch([lr[8 * (i % 4):8 * (i % 4) + 8]], hr[8 * (i % 4):8 * (i % 4) + 8]) for i in range(3, 6)]
    assert np.allclose(got, ref, rtol=2e-4), (got, ref)
    # Predictor on the trained model must not replace the arena the trainer's captured step updates
    arena = tr2.model.arena
    Predictor(tr2, hr[:8], scale=4, array_in_hr=True, batch_size=8).run()
    assert tr2.model.arena is arena
    tr2.train_on_batch([lr[:8]], hr[:8])


_DDG_CASES = [
    # upsampling, scale, patch, time_window, interpolation, C, predictors ('hr' | 'lr' | None), static, exact
    ('pin', 4, None, None, 'inter_area', 1, None, True, True),          # BASELINE config 5's data path
    ('pin', 4, 24, None, 'inter_area', 2, 'hr', True, True),
    ('pin', 4, None, None, 'bicubic', 1, 'lr', False, False),
    ('pin', 2, 16, None, 'bilinear', 2, None, True, False),
    ('pin', 4, None, 3, 'inter_area', 1, None, True, True),             # recnet_pin
    ('pin', 4, None, 3, 'bicubic', 2, 'hr', False, False),
    ('rc', 4, None, 3, 'inter_area', 1, None, True, True),              # BASELINE config 4's data path
    ('spc', 4, 16, 2, 'inter_area', 2, None, False, True),
    ('spc', 4, None, 3, 'inter_area', 1, 'hr', True, True),
    ('spc', 4, None, None, 'bilinear', 2, 'hr', True, False),
    ('spc', 4, 16, None, 'bicubic', 1, None, False, False),
    ('spc', 5, None, None, 'inter_area', 1, None, False, False),        # 48/5, 64/5: cv2's non-integer area path
    ('dc', 2, None, None, 'lanczos', 1, 'lr', False, False),
    ('spc', 4, None, None, 'nearest', 1, None, True, True),
]


@pytest.mark.parametrize('case', _DDG_CASES, ids=lambda c: '-'.join(str(v) for v in c))
def test_device_data_generator_pin_time_window_interpolations(cuda, case):
    """DeviceDataGenerator for `pin` pairs, `time_window` windows and every interpolation (SURVEY 8f row 1,
    dataloader.py:108-222, utils.py:341-401) against the host DataGenerator (bit-exact vs the reference's own
    outputs, tests/test_datapath.py) under the same numpy RNG state.  `exact`: bit-for-bit (block means, pure
    gathers / replication); otherwise the same cv2 coefficients in another summation order: 2e-6 of max|x|."""
    from dl4ds_b200.dataloader import DataGenerator, DeviceDataGenerator
    ups, scale, patch, tw, interp, C, pred, static, exact = case
    rng = np.random.default_rng(5)
    n, H, W = 14, 48, 64
    hr = rng.standard_normal((n, H, W, C)).astype(np.float32)
    preds = None
    if pred == 'hr':
        preds = [rng.standard_normal((n, H, W, 1)).astype(np.float32) for _ in range(2)]
    elif pred == 'lr':
        preds = [rng.standard_normal((n, int(H / scale), int(W / scale), 1)).astype(np.float32) for _ in range(2)]
    st = [rng.standard_normal((H, W)).astype(np.float32)] if static else None
    kw = dict(backbone='resnet', upsampling=ups, scale=scale, batch_size=4, patch_size=patch, time_window=tw,
              static_vars=st, predictors=preds, interpolation=interp)
    assert DeviceDataGenerator.supported(hr, None, ups, scale, patch, tw, st, preds, interp)
    np.random.seed(11)
    host = DataGenerator(hr, None, **kw)
    hb = [host[i] for i in range(len(host))]
    np.random.seed(11)
    dev = DeviceDataGenerator(hr, None, device=cuda, **kw)
    assert len(dev) == len(host) and len(dev) >= 2
    for i in range(len(dev)):
        din, (hr_d,) = dev[i]
        hin, (hr_h,) = hb[i]
        np.testing.assert_array_equal(hr_d, hr_h)
        assert len(din) == len(hin)
        for a, b in zip(din, hin):
            b = np.asarray(b, np.float32).reshape(a.shape)      # App. B #2: a single-channel 5-D LR batch lost its axis
            if exact:
                np.testing.assert_array_equal(a, b)
            else:
                assert np.abs(a - b).max() <= 2e-6 * np.abs(b).max(), np.abs(a - b).max()


def test_device_data_generator_trains_cfg4_cfg5_shapes(cuda):
    """SupervisedTrainer(data_on_device=True) for a recurrent post-upsampling model (time_window) and a U-Net `pin`
    model: same loss history as the host generator under the same numpy RNG state."""
    hr = _data(22, 32, 5)
    static = np.random.default_rng(6).standard_normal((32, 32)).astype(np.float32)
    for backbone, ups, tw, extra in (('resnet', 'rc', 3, dict(n_filters=4, n_blocks=1)),
                                     ('unet', 'pin', None, dict(n_filters=4, n_blocks=2, n_channels_out=1))):
        hist = []
        for on_dev in (False, True):
            np.random.seed(3)
            tr = SupervisedTrainer(backbone, ups, hr[:14], hr[14:18], hr[18:], scale=4, batch_size=4, epochs=2,
                                   time_window=tw, static_vars=[static], verbose=False, show_plot=False, seed=4,
                                   data_on_device=on_dev, math='fp32', **extra)
            tr.run()
            from dl4ds_b200.dataloader import DeviceDataGenerator
            assert isinstance(tr.ds_train, DeviceDataGenerator) == on_dev
            hist.append(tr.fithist.history['loss'])
        np.testing.assert_allclose(hist[0], hist[1], rtol=2e-5)



@pytest.mark.parametrize('case', [('spc', 4, None, None, 'inter_area', True), ('pin', 4, None, None, 'bicubic', True),
                                  ('pin', 4, 24, None, 'inter_area', False), ('rc', 4, None, 3, 'inter_area', True)],
                         ids=lambda c: '-'.join(str(v) for v in c))
def test_device_data_generator_explicit_lr(cuda, case):
    """Explicit pairs (`array_lr` given, the MOS case of dataloader.py:108-116,157-222): the LR member is gathered as
    is (post-upsampling) or interpolated onto the HR grid once (`pin`), against the host DataGenerator."""
    from dl4ds_b200.dataloader import DataGenerator, DeviceDataGenerator
    ups, scale, patch, tw, interp, static = case
    rng = np.random.default_rng(8)
    n, H, W, C = 12, 48, 64, 2
    hr = rng.standard_normal((n, H, W, C)).astype(np.float32)
    lr = rng.standard_normal((n, H // scale, W // scale, C)).astype(np.float32)
    st = [rng.standard_normal((H, W)).astype(np.float32)] if static else None
    kw = dict(backbone='resnet', upsampling=ups, scale=scale, batch_size=4, patch_size=patch, time_window=tw,
              static_vars=st, interpolation=interp)
    assert DeviceDataGenerator.supported(hr, lr, ups, scale, patch, tw, st, None, interp)
    np.random.seed(5)
    host = DataGenerator(hr, lr, **kw)
    hb = [host[i] for i in range(len(host))]
    np.random.seed(5)
    dev = DeviceDataGenerator(hr, lr, device=cuda, **kw)
    for i in range(len(dev)):
        din, (hr_d,) = dev[i]
        hin, (hr_h,) = hb[i]
        np.testing.assert_array_equal(hr_d, hr_h)
        for a, b in zip(din, hin):
            b = np.asarray(b, np.float32).reshape(a.shape)
            assert np.abs(a - b).max() <= 2e-6 * np.abs(b).max(), np.abs(a - b).max()
