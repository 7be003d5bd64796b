"""Oracle parity on the LITERAL BASELINE.json configurations, in the default math mode of the trainers and of
bench.py (``tf32x3``, tensor cores, 3-term split) -- VERDICT r01 item 1.

  cfg2  resnet + 4x SPC, 32 -> 128, 1 channel, n_blocks 6, batch 64, composed SPC x TransitionLast path: forward,
        every parameter-tensor gradient, and 5 Adam steps through SupervisedTrainer vs oracle.supervised_step
        (models/sp_postups.py:95-217, training/supervised.py:336-353,396-406)
  cfg3  densenet + attention + LCB, 8x deconvolution, LR 16 -> HR 128, 5 LR channels + 1 HR aux, batch 16 (per GPU)
  cfg4  recurrent resnet + 4x resize-convolution, T = 6, 32 -> 128, batch 8 (per GPU)  (models/spt_postups.py:12-163)
  cfg5  one cGAN train_step, U-Net generator (pin) + residual discriminator, 256 x 256, batch 4 (per GPU), given
        dropout masks (training/cgan.py:575-639)
  smooth the cfg2 backbone + sub-pixel block with tanh activations and a linear head (no ReLU anywhere): 2e-4 on every
        gradient through the same ~20 tensor-core layers; and the full cfg2 graph with tanh blocks, whose tail keeps
        TransitionLast's ReLU (App. B #9), at 3e-3

Why the ReLU graphs get GRAD_TOL = 1.5e-2 (scratch/parity_diag.py, profiles/r02_parity_diag.txt): against an fp64
evaluation of the oracle graph the tf32x3 forward is within 4.4e-6 of max|y| (fp32 CUDA cores: 6e-7; the tensor core
accumulates 432+ products per output with its own rounding).  A pre-activation that close to zero comes out on the other
side of a ReLU for a few elements per million; a weight gradient is a sum over ~10^6 pixels of random-sign terms, so m
flipped terms move it by ~sqrt(m / n) of its size: 1e-3 .. 1e-2 of the tensor's max (ResidualBlock5/conv1/kernel: 1.0e-2
at batch 8, 6.8e-3 at batch 64), identical from run to run.  The tail layers behind the last ReLU agree to 4e-6, the
smooth graph to 2e-4, and the 5-step Adam trajectories to 2e-4 in the loss.

The oracle (oracle/torch_ref.py, torch-CPU fp32) is test infrastructure; it takes 0.5 - 20 s per case here.
"""
import numpy as np
import pytest
import torch

from dl4ds_b200 import SupervisedTrainer, nets
from dl4ds_b200.training import cgan
from oracle import torch_ref as R
from tests.util import assert_adam_weights_close, compare, rel_err, run_engine, run_oracle, trace_spec

pytestmark = pytest.mark.gpu
MATH = 'tf32x3'
FWD_TOL = 5e-5          # of max |y|
GRAD_TOL = 1.5e-2       # of each gradient tensor's max, ReLU graphs (mask flips, see the module docstring)
SMOOTH_GRAD_TOL = 2e-4


def _report(tag, y, y_ref, pg, pg_ref):
    worst = max(((rel_err(pg[k], pg_ref[k]), k) for k in pg_ref), default=(0.0, ''))
    print('[%s] fwd rel err %.2e; worst gradient rel err %.2e (%s) over %d tensors'
          % (tag, rel_err(y, y_ref), worst[0], worst[1], len(pg_ref)))


def test_cfg2_forward_and_all_gradients_batch64(cuda):
    """BASELINE configs[1] exactly as bench.py times it: batch 64, n_blocks 6, composed last sub-pixel stage."""
    m = nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (32, 32), math=MATH, fuse_spc_transition=True)
    assert m.count_params() == 204405
    ofn = lambda p, xs: R.net_postupsampling(p, xs, 'resnet', 'spc', 4)
    shapes = [(64, 32, 32, 1)]
    spec = trace_spec(m.fn, shapes)
    assert len(spec) == 56
    weights = R.init_weights(spec, seed=3, bias_scale=0.1)
    rng = np.random.default_rng(4)
    x = rng.standard_normal(shapes[0]).astype(np.float32)
    seed_grad = (rng.standard_normal((64, 128, 128, 1)) / (64 * 128 * 128)).astype(np.float32)   # MAE-sized seed
    y_ref, pg_ref, _ = run_oracle(ofn, weights, [x], seed_grad)
    y, pg, _ = run_engine(m.fn, spec, {k: v.numpy() for k, v in weights.items()}, [x], cuda, MATH, seed_grad,
                          input_grads=False)
    _report('cfg2 b64', y, y_ref, pg, pg_ref)
    assert rel_err(y, y_ref) <= FWD_TOL
    for k in spec:
        assert rel_err(pg[k], pg_ref[k]) <= GRAD_TOL, (k, rel_err(pg[k], pg_ref[k]))


def test_cfg2_five_adam_steps_batch64(cuda):
    """Five optimizer steps of the literal cfg2 through SupervisedTrainer.train_on_batch (captured CUDA graphs,
    tf32x3, composed SPC path) == the oracle's supervised_step, loss by loss, then the weights."""
    np.random.seed(0)
    hr = np.random.default_rng(5).standard_normal((128, 128, 128, 1)).astype(np.float32)
    tr = SupervisedTrainer('resnet', 'spc', hr, hr[:64], hr[:64], scale=4, batch_size=64, epochs=1,
                           learning_rate=(1e-3, 1e-4), lr_decay_after=3, verbose=False, math=MATH, seed=7)
    tr.setup_datagen()
    tr.setup_model()
    assert tr.model.count_params() == 204405
    w = {k: torch.from_numpy(v.copy()) for k, v in tr.model.get_weights().items()}
    opt = R.TFAdam(list(w), lr=R.piecewise_constant(3, 1e-3, 1e-4))
    fwd = lambda p, xs: R.net_postupsampling(p, xs, 'resnet', 'spc', 4)
    for i in range(5):
        (lr,), (y,) = tr.ds_train[i % len(tr.ds_train)]
        loss = tr.train_on_batch([lr], y)
        ref, _ = R.supervised_step(fwd, w, opt, [torch.from_numpy(lr)], torch.from_numpy(y))
        print('[cfg2 adam] step %d loss %.7f oracle %.7f' % (i, loss, ref))
        assert abs(loss - ref) <= 2e-4 * max(1.0, abs(ref)), (i, loss, ref)
    assert_adam_weights_close(tr.model.get_weights(), {k: v.numpy() for k, v in w.items()}, lr=1e-3, steps=5,
                              tight=5e-4, frac=5e-2)


def test_cfg2_smooth_graph_holds_2e4(cuda):
    """cfg2 backbone (tanh blocks) + the 4x sub-pixel block + a linear 3x3 head, batch 16: no ReLU in the graph, so
    nothing amplifies the forward rounding: every gradient within 2e-4 of its tensor's max in tf32x3."""
    from dl4ds_b200 import blocks as B

    def fn(c, xs):
        x, nf = nets._backbone(c, xs[0], 'resnet', 8, 6, False, 'tanh')
        x = B.subpixel_block(c, 'SubpixelConvolution', x, 4, nf)
        return c.conv(x, 'head', 8, k=3)

    def ofn(p, xs):
        x, nf = R._backbone(p, R._nchw(xs[0]), 'resnet', 8, 6, False, 'tanh')
        x = R.subpixel_block(p, 'SubpixelConvolution', x, 4, nf)
        return R._nhwc(R._conv(p, 'head', x, 8, k=3))
    shapes = [(16, 32, 32, 1)]
    spec = trace_spec(fn, shapes)
    weights = R.init_weights(spec, seed=8, bias_scale=0.1)
    rng = np.random.default_rng(9)
    x = rng.standard_normal(shapes[0]).astype(np.float32)
    seed_grad = rng.standard_normal((16, 128, 128, 8)).astype(np.float32)
    y_ref, pg_ref, _ = run_oracle(ofn, weights, [x], seed_grad)
    y, pg, _ = run_engine(fn, spec, {k: v.numpy() for k, v in weights.items()}, [x], cuda, MATH, seed_grad,
                          input_grads=False)
    _report('cfg2 smooth', y, y_ref, pg, pg_ref)
    assert rel_err(y, y_ref) <= 2e-5
    for k in spec:
        assert rel_err(pg[k], pg_ref[k]) <= SMOOTH_GRAD_TOL, (k, rel_err(pg[k], pg_ref[k]))


def test_cfg2_tanh_blocks_full_graph(cuda):
    """The full cfg2 graph with tanh block activations (batch 16).  TransitionLast keeps its ReLU (App. B #9) and the
    attention MLP its own, so a few mask flips remain: 3e-3."""
    m = nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (32, 32), math=MATH, activation='tanh')
    ofn = lambda p, xs: R.net_postupsampling(p, xs, 'resnet', 'spc', 4, activation='tanh')
    shapes = [(16, 32, 32, 1)]
    spec = trace_spec(m.fn, shapes)
    weights = R.init_weights(spec, seed=8, bias_scale=0.1)
    rng = np.random.default_rng(9)
    x = rng.standard_normal(shapes[0]).astype(np.float32)
    seed_grad = rng.standard_normal((16, 128, 128, 1)).astype(np.float32)
    y_ref, pg_ref, _ = run_oracle(ofn, weights, [x], seed_grad)
    y, pg, _ = run_engine(m.fn, spec, {k: v.numpy() for k, v in weights.items()}, [x], cuda, MATH, seed_grad,
                          input_grads=False)
    _report('cfg2 tanh', y, y_ref, pg, pg_ref)
    assert rel_err(y, y_ref) <= 2e-5
    for k in spec:
        assert rel_err(pg[k], pg_ref[k]) <= 3e-3, (k, rel_err(pg[k], pg_ref[k]))


def test_cfg3_densenet_attention_lcb_dc8_batch16(cuda):
    """BASELINE configs[2] at its per-GPU size: LR 16 -> HR 128, 5 LR channels (1 + 1 static + 3 predictors) + 1 HR aux
    channel, densenet + channel attention + LocalizedConvBlock, 8x deconvolution, batch 16 (128 over 8 GPUs)."""
    m = nets.net_postupsampling('densenet', 'dc', 8, 5, 1, (16, 16), attention=True, localcon_layer=True, math=MATH)
    assert m.count_params() == 596026
    ofn = lambda p, xs: R.net_postupsampling(p, xs, 'densenet', 'dc', 8, attention=True, localcon_layer=True)
    compare(m.fn, ofn, [(16, 16, 16, 5), (16, 128, 128, 1)], cuda, math=MATH, tol=FWD_TOL, gtol=GRAD_TOL,
            input_grads=False)


def test_cfg4_recurrent_resnet_rc_T6_batch8(cuda):
    """BASELINE configs[3] at its per-GPU size: ConvLSTM residual backbone, 4x resize-convolution, T = 6, 32 -> 128,
    batch 8 (32 over 4 GPUs); LR frames carry the channel axis the reference evidently intends (App. B #2)."""
    T, Bz = 6, 8
    m = nets.recnet_postupsampling('resnet', 'rc', 4, 1, 0, (32, 32), T, n_blocks=4, math=MATH)
    assert m.count_params() == 83385

    def ofn(p, xs):
        x = xs[0]
        x5 = x.reshape(T, Bz, *x.shape[1:]).permute(1, 0, 2, 3, 4)      # time-major frames -> (B,T,h,w,C)
        y5 = R.recnet_postupsampling(p, [x5], 'resnet', 'rc', 4, T, n_blocks=4)
        return y5.permute(1, 0, 2, 3, 4).reshape(T * Bz, *y5.shape[2:])
    compare(m.fn, ofn, [(T * Bz, 32, 32, 1)], cuda, math=MATH, tol=FWD_TOL, gtol=GRAD_TOL, input_grads=False)


def test_cfg5_cgan_step_256_batch4(cuda):
    """BASELINE configs[4] at its per-GPU size: one cGAN train_step (U-Net generator n_filters 8 / n_blocks 6 on
    256 x 256 pre-upsampled input + 1 static channel, residual discriminator), batch 4 (32 over 8 GPUs): the four
    losses and BOTH Adam(beta_1 = 0.5) updates vs the oracle, with the Dropout(0.4) masks given."""
    rng = np.random.default_rng(11)
    B, hw = 4, 256
    G = nets.unet_pin('unet', 2, 1, (hw, hw), 1, 8, 6, math=MATH).to(cuda)
    D = nets.residual_discriminator(2, 'pin', False, 4, (hw, hw), n_filters=8, n_res_blocks=4, math=MATH).to(cuda)
    assert G.count_params() == 5705365 and D.count_params() == 15961
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
    gen_fn = lambda p, xs: R.unet_pin(p, xs, 8, 6)
    disc_fn = lambda p, xs, mk: R.residual_discriminator(p, xs, 'pin', 4, (hw, hw), n_filters=8, n_res_blocks=4,
                                                         dropout_mask=mk)
    gopt, dopt = R.TFAdam(list(gw), lr=2e-4, beta_1=0.5), R.TFAdam(list(dw), lr=2e-4, beta_1=0.5)
    ref, _, _ = R.cgan_step(gen_fn, disc_fn, gw, dw, gopt, dopt, torch.from_numpy(lr), torch.from_numpy(hr),
                            torch.from_numpy(st), mask_real=torch.from_numpy(masks[0].reshape(B, nfeat)),
                            mask_fake=torch.from_numpy(masks[1].reshape(B, nfeat)))
    print('[cfg5] losses', losses, 'oracle', [float(r) for r in ref])
    for a, b in zip(losses, ref):
        assert abs(a - b) <= 2e-4 * max(1.0, abs(b)), (losses, ref)
    for model, w in ((G, gw), (D, dw)):
        assert_adam_weights_close(model.get_weights(), {k: v.numpy() for k, v in w.items()}, lr=2e-4, steps=1,
                                  tight=5e-5, frac=5e-2)
