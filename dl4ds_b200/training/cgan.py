"""CGANTrainer, train_step, generator_loss, discriminator_loss -- dl4ds/training/cgan.py:30-639 on
the CUDA engine.

One ``train_step`` (cgan.py:575-639) = generator forward, discriminator on (LR, HR) and on
(LR, generated), the four losses, two gradient computations and two Adam(beta_1=0.5) updates.  The
reference differentiates D(fake) twice with its two tapes (once for the discriminator weights, once
through to the generator); here the recorded D(fake) tape is replayed with the two seeds:
    pass 1 (D loss):  seed dBCE(0,p_fake)/dp -> discriminator weight gradients (input gradient dropped)
    pass 2 (G loss):  seed dBCE(1,p_fake)/dp -> input gradient only (``param_grads=False``) -> generator
which is mathematically identical.  Dropout(0.4) inside the discriminator is active during training
(``training=True``); its keep masks are drawn per call, as Keras does, and applied by ``dl4ds_mul``.
Data parallelism: both gradient arenas are all-reduced (sum) and averaged inside Adam
(``hvd.DistributedGradientTape``, cgan.py:608-611); after the first step theta/m/v of both models
are broadcast from rank 0 (cgan.py:626-637).  The cGAN learning rate is NOT scaled by world size.

Reference bug fixed: ``run`` passes an undefined ``aux_hr`` to ``train_step`` when
``static_vars`` is None (cgan.py:346); here ``static_array=None`` is passed.
"""
import os

import numpy as np

from .. import nets
from ..dataloader import create_batch_hr_lr
from ..utils import POSTUPSAMPLING_METHODS, Timing
from .base import Trainer

DROPOUT_RATE = 0.4      # discriminator.py:77
LAMBDA = 100.0          # cgan.py:526


class Adam:
    """tf.keras.optimizers.Adam(lr, beta_1) bound to a model's flat arena (theta / m / v / t)."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon

    def apply(self, model, grad_scale=1.0):
        from ..engine import adam_step
        adam_step(model.arena, self.learning_rate, self.beta_1, self.beta_2, self.epsilon, grad_scale)


def _bce_np(target, p, eps=1e-7):
    p = np.clip(np.asarray(p, np.float64), eps, 1.0 - eps)
    return float(-np.mean(target * np.log(p + eps) + (1.0 - target) * np.log(1.0 - p + eps)))


def _host_loss(name, y_true, y_pred):
    """Pixel losses in numpy fp64; the SSIM family (losses.py:23-151) through the device kernels."""
    name = getattr(name, '__name__', name)
    if name in ('mae', 'mse'):
        d = np.asarray(y_true, np.float64) - np.asarray(y_pred, np.float64)
        return float(np.mean(np.abs(d))) if name == 'mae' else float(np.mean(d * d))
    from .. import losses
    return losses._value(name, y_true, y_pred)


def generator_loss(disc_generated_output, gen_output, target, gen_pxloss_function, lambda_scaling_factor=100):
    """cgan.py:525-553 on host arrays (API parity; the training step computes these on the device)."""
    gan = _bce_np(1.0, disc_generated_output)
    px = _host_loss(gen_pxloss_function or 'mae', target, gen_output)
    return gan + lambda_scaling_factor * px, gan, px


def discriminator_loss(disc_real_output, disc_generated_output):
    """cgan.py:556-572 on host arrays."""
    return _bce_np(1.0, disc_real_output) + _bce_np(0.0, disc_generated_output)


def _draw_keep_masks(D, masks):
    """Fill ``masks`` (CUDA tensors (B,1,1,F)) with fresh inverted-dropout keep masks of rate DROPOUT_RATE
    (discriminator.py:77 under ``training=True``: an independent mask per D call): ``dl4ds_dropout`` (Philox keyed by
    the discriminator arena's device-resident {seed, step}) applied to a tensor of ones; no torch operator."""
    import torch
    from .. import _lib
    state = D.arena.rng_state()
    st = torch.cuda.current_stream().cuda_stream
    _lib.call('dl4ds_rng_advance', state.data_ptr(), st)
    for i, m in enumerate(masks):
        ones = getattr(D, '_mask_ones', None)
        if ones is None or ones.shape != m.shape or ones.device != m.device:
            ones = D._mask_ones = torch.ones_like(m)
        B, F = m.shape[0], m.shape[-1]
        _lib.call('dl4ds_dropout', ones.data_ptr(), F, m.data_ptr(), F, B, 1, B, F, float(DROPOUT_RATE), 0,
                  state.data_ptr(), 1000 + i, st)
    return masks


def _dev_inputs(model, arrays, dev):
    ins, _, _ = model._prep_inputs(arrays, dev)
    return ins


def _ctx(model, plan):
    """A training Ctx on ``model``'s arena that shares the model's tensor-core weight-image cache (``plan``: the
    engine.PackPlan that has just refreshed those images on this stream, or None: first use packs)."""
    from ..engine import Ctx
    c = Ctx(model.arena, model.math, training=True)
    if not hasattr(model, '_pack_cache'):
        model._pack_cache = {}
    c.pack_cache = model._pack_cache
    c.prepacked = plan.keys if plan is not None else None
    return c


def _fwd_bwd(G, D, lr_t, hr_t, st_t, mk, losses, gen_pxloss_function, plans=None, side=None, shadow_grad=None,
             wgrad_stream=None):
    """Device work of one cGAN step up to the gradients (cgan.py:587-611): zero both gradient arenas, generator
    forward, D(real), D(fake), the four loss terms into ``losses`` (gan, lambda*px, d_real, d_fake) and the two
    backward passes.  Pure kernel launches: ``train_step`` runs it eagerly on the current stream, ``CGANStep``
    captures it into a CUDA graph with ``plans`` = (generator, discriminator) PackPlans (every tensor-core weight
    image packed by one launch per model) and ``side`` = a second stream for the D(real) branch: its forward and
    backward depend on nothing the generator produces, and at the small per-GPU batches of the cGAN configurations
    (BASELINE config 5: 4 samples) no kernel of either branch fills the GPU.  The branch accumulates its weight
    gradients into ``shadow_grad`` (same layout as the discriminator's gradient arena), added at the join.
    ``wgrad_stream``: the weight-gradient kernels of the other two backward passes leave the dgrad chain for a third
    stream (engine.Ctx.wgrad_stream)."""
    import copy
    import torch
    from .. import _lib
    from ..engine import Var
    G.arena.zero_grad(); D.arena.zero_grad()
    losses.zero_()
    pg, pd = plans if plans is not None else (None, None)
    for pl in (pg, pd):
        if pl is not None:
            pl.run()
    main = torch.cuda.current_stream()

    def d_real(arena):
        cr = _ctx(D, pd)
        cr.arena = arena
        p_real = D.fn(cr, [cr.input(lr_t), cr.input(hr_t), cr.input(mk[0])])
        cr.bce_loss(p_real, 1.0, loss_buf=losses[2:3])
        cr.backward()

    two_streams = side is not None and shadow_grad is not None and plans is not None
    if two_streams:
        shadow = copy.copy(D.arena)
        shadow.grad = shadow_grad
        side.wait_stream(main)
        with torch.cuda.stream(side):
            shadow_grad.zero_()
            d_real(shadow)
    # ---- forward: generator, D(real), D(fake)
    cg = _ctx(G, pg)
    cg.wgrad_stream = wgrad_stream if two_streams else None
    gin = [cg.input(lr_t)] + ([cg.input(st_t)] if st_t is not None else [])
    gen = G.fn(cg, gin)
    if not two_streams:
        d_real(D.arena)
    cf = _ctx(D, pd)
    cf.wgrad_stream = wgrad_stream if two_streams else None
    gen_in = Var(gen.buf, gen.off, gen.C, requires_grad=True)
    p_fake = D.fn(cf, [cf.input(lr_t), gen_in, cf.input(mk[1])])

    # ---- discriminator loss and weight gradients
    cf.bce_loss(p_fake, 0.0, loss_buf=losses[3:4])
    cf.backward(keep_tape=True)
    gen_in.grad = None                      # d(D loss)/d(gen) is not used by either optimizer
    # ---- generator loss: through D(fake) to the generated field, then through G
    cf.param_grads = False
    cf.bce_loss(p_fake, 1.0, loss_buf=losses[0:1])
    cf.backward()
    if gen_in.grad is not None:
        cg._use(gen)                        # the discriminator is a second consumer of the generated field
        cg._give_grad(gen, gen_in.grad)
    cg.pixel_loss(gen, cg.input(hr_t), gen_pxloss_function or 'mae', scale=LAMBDA, loss_buf=losses[1:2])
    cg.backward()
    if two_streams:
        main.wait_stream(side)
        g = D.arena.grad
        _lib.call('dl4ds_axpby', 1.0, shadow_grad.data_ptr(), 1.0, g.data_ptr(), g.numel(), main.cuda_stream)


class CGANStep:
    """One cGAN optimisation step on fixed-shape device batches as replayable CUDA graphs (what ``train_step``
    does eagerly, ~450 launches whose Python/ctypes issue cost exceeds the GPU time at small per-GPU batches):
    graph 1 = ``_fwd_bwd``; [NCCL all-reduce of both gradient arenas]; graph 2 = the two Adam(beta_1=0.5) updates
    with the bias-corrected step sizes read from device scalars.  The host refreshes the static input buffers, the
    two dropout keep masks (drawn per step, discriminator.py:77 with ``training=True``) and the step sizes."""

    def __init__(self, generator, discriminator, lr_shape, hr_shape, static_shape=None, gen_pxloss_function='mae',
                 learning_rates=(2e-4, 2e-4), beta_1=0.5, beta_2=0.999, eps=1e-7, dist=None):
        import torch
        self.G, self.D = generator, discriminator
        dev = self.G.arena.device
        z = lambda shp: torch.zeros(tuple(shp), dtype=torch.float32, device=dev)
        self.lr, self.hr = z(lr_shape), z(hr_shape)
        self.st = z(static_shape) if static_shape is not None else None
        # spatio-temporal step (5-D arrays, cgan.py:575-639 with recnet generators): the (B,T,H,W,C) staging buffers
        # are re-ordered into the engine's time-major frames (T*B,H,W,C) by dl4ds_permute_frames inside the graph
        self.frames = None
        if len(lr_shape) == 5:
            fr = lambda shp: z((shp[0] * shp[1],) + tuple(shp[2:]))
            self.frames = (fr(lr_shape), fr(hr_shape))
        nfeat = self.D.spec['dense1/kernel'][0]
        self.masks = [z((lr_shape[0], 1, 1, nfeat)) for _ in range(2)]
        self.losses = z((4,))
        self.pxloss = gen_pxloss_function
        self.lrs = learning_rates
        self.b1, self.b2, self.eps = beta_1, beta_2, eps
        self.dist = dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.lr_t = [z((1,)), z((1,))]
        self._lr_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        self.graph_fb = self.graph_opt = None
        # captured step only: one pack launch per model and the D(real) branch on a second stream (see _fwd_bwd);
        # a discriminator with batch normalisation updates its moving statistics in both passes, in program order
        self.plans = None
        n_streams = int(os.environ.get('DL4DS_CGAN_STREAMS', '3'))
        two = n_streams > 1 and not any('moving_mean' in k for k in self.D.spec)
        self.side = torch.cuda.Stream(device=dev) if two and dev.type == 'cuda' else None
        self.wside = torch.cuda.Stream(device=dev) if self.side is not None and n_streams > 2 else None
        self.shadow_grad = torch.zeros_like(self.D.arena.grad) if self.side is not None else None

    def _opt(self):
        import torch
        from .. import _lib
        st = torch.cuda.current_stream().cuda_stream
        for m, lr_t in ((self.G, self.lr_t[0]), (self.D, self.lr_t[1])):
            a = m.arena
            _lib.call('dl4ds_adam_step_dev', a.theta.data_ptr(), a.grad.data_ptr(), a.m.data_ptr(), a.v.data_ptr(), a.n,
                      lr_t.data_ptr(), float(self.b1), float(self.b2), float(self.eps), 1.0 / self.world, st)

    def _device_step(self):
        """Everything the captured graph holds up to the gradients."""
        import torch
        from .. import _lib
        lr, hr = self.lr, self.hr
        if self.frames is not None:
            st = torch.cuda.current_stream().cuda_stream
            for src, dst in zip((self.lr, self.hr), self.frames):
                b, t = src.shape[0], src.shape[1]
                _lib.call('dl4ds_permute_frames', src.data_ptr(), dst.data_ptr(), b, t, src[0, 0].numel(), st)
            lr, hr = self.frames
        _fwd_bwd(self.G, self.D, lr, hr, self.st, self.masks, self.losses, self.pxloss, plans=self.plans,
                 side=self.side, shadow_grad=self.shadow_grad, wgrad_stream=self.wside)

    def _exchange(self):
        """hvd.DistributedGradientTape (cgan.py:608-611): both gradient arenas summed over the ranks on this stream."""
        if self.dist is not None and self.world > 1:
            from ..step import allreduce_sum_
            allreduce_sum_(self.G.arena.grad, self.dist)
            allreduce_sum_(self.D.arena.grad, self.dist)

    def capture(self):
        import torch
        snap = [(m.arena.theta.clone(), m.arena.m.clone(), m.arena.v.clone(), m.arena.t) for m in (self.G, self.D)]
        self._device_step()                                                                         # warm-up (eager)
        self._exchange()
        self._opt()
        torch.cuda.synchronize()
        for m, (th, mm, vv, t) in zip((self.G, self.D), snap):
            m.arena.theta.copy_(th); m.arena.m.copy_(mm); m.arena.v.copy_(vv); m.arena.t = t
        from ..engine import PackPlan
        self.plans = tuple(PackPlan(m.arena, getattr(m, '_pack_cache', {})) for m in (self.G, self.D))   # images the eager pass used
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.graph_fb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_fb, stream=s):
                # one graph: both backward passes, the two NCCL all-reduces (dl4ds_comm_*, captured), both Adam updates
                self._device_step()
                self._exchange()
                self._opt()
            self.graph_opt = self.graph_fb
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        return self

    def run(self, lr_array, hr_array, static_array=None, dropout_masks=None, first_batch=False):
        """One step on host (or device) arrays; returns (gen_total, gen_gan, gen_px, disc) floats like ``train_step``."""
        import math
        import torch
        f32 = lambda a: torch.as_tensor(np.asarray(a, np.float32) if not torch.is_tensor(a) else a)
        self.lr.copy_(f32(lr_array), non_blocking=True)
        self.hr.copy_(f32(hr_array), non_blocking=True)
        if self.st is not None:
            self.st.copy_(f32(static_array), non_blocking=True)
        keep = 1.0 - DROPOUT_RATE
        if dropout_masks is not None:
            for i, m in enumerate(self.masks):
                m.copy_(f32(dropout_masks[i]).reshape(m.shape))
        else:
            _draw_keep_masks(self.D, self.masks)
        for i, mdl in enumerate((self.G, self.D)):
            mdl.arena.t += 1
            t = mdl.arena.t
            self._lr_host[i] = self.lrs[i] * math.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
        self.lr_t[0].copy_(self._lr_host[0:1], non_blocking=True)
        self.lr_t[1].copy_(self._lr_host[1:2], non_blocking=True)
        self.graph_fb.replay()
        d = self.dist
        if d is not None and self.world > 1 and first_batch:
            for mdl in (self.G, self.D):
                for t in (mdl.arena.theta, mdl.arena.m, mdl.arena.v):
                    d.broadcast(t, src=0)
        gan, pxs, dr, df = [float(v) for v in self.losses.cpu().numpy()]
        return gan + pxs, gan, pxs / LAMBDA, dr + df


def train_step(lr_array, hr_array, generator, discriminator, generator_optimizer, discriminator_optimizer,
               epoch=0, gen_pxloss_function='mae', summary_writer=None, first_batch=False, static_array=None,
               dropout_masks=None, dist=None):
    """One cGAN optimisation step (cgan.py:575-639).  Arrays are host numpy / CUDA tensors (NHWC).
    ``dropout_masks`` = (mask_real, mask_fake) of shape (B,1,1,F) overrides the random keep masks
    (tests).  Returns (gen_total_loss, gen_gan_loss, gen_px_loss, disc_loss) as floats."""
    import torch
    from ..engine import Ctx, Var
    G, D = generator, discriminator
    G.to('cuda'); D.to('cuda')
    dev = G.arena.device
    f32 = lambda a: torch.as_tensor(np.asarray(a, np.float32) if not torch.is_tensor(a) else a).to(dev, torch.float32).contiguous()
    lr_t, hr_t = f32(lr_array), f32(hr_array)
    B = lr_t.shape[0]
    if lr_t.dim() == 5:         # spatio-temporal models: (B,T,H,W,C) -> the engine's time-major frames (T*B,H,W,C)
        tm = lambda x: x.transpose(0, 1).reshape(x.shape[0] * x.shape[1], *x.shape[2:]).contiguous()
        lr_t, hr_t = tm(lr_t), tm(hr_t)
    st_t = f32(static_array) if static_array is not None else None
    nfeat = D.spec['dense1/kernel'][0]
    if dropout_masks is None:
        mk = _draw_keep_masks(D, [torch.empty((B, 1, 1, nfeat), dtype=torch.float32, device=dev) for _ in range(2)])
    else:
        mk = [f32(m) for m in dropout_masks]
    world = dist.get_world_size() if dist is not None else 1

    losses = torch.zeros(4, dtype=torch.float32, device=dev)        # gan, px, d_real, d_fake
    _fwd_bwd(G, D, lr_t, hr_t, st_t, mk, losses, gen_pxloss_function)

    # ---- exchange + Adam
    if dist is not None and world > 1:
        dist.all_reduce(G.arena.grad, op=dist.ReduceOp.SUM)
        dist.all_reduce(D.arena.grad, op=dist.ReduceOp.SUM)
    generator_optimizer.apply(G, 1.0 / world)
    discriminator_optimizer.apply(D, 1.0 / world)
    if dist is not None and world > 1 and first_batch:
        for m in (G, D):
            for t in (m.arena.theta, m.arena.m, m.arena.v):
                dist.broadcast(t, src=0)
    gan, pxs, dr, df = [float(v) for v in losses.cpu().numpy()]
    px = pxs / LAMBDA
    return gan + pxs, gan, px, dr + df


class CGANTrainer(Trainer):
    """Procedure for training the conditional adversarial models (same signature as cgan.py:33-64;
    ``math`` and ``seed`` are additions)."""

    def __init__(self, backbone, upsampling, data_train, data_test, data_train_lr=None, data_test_lr=None,
                 predictors_train=None, predictors_test=None, scale=5, patch_size=None, time_window=True,
                 loss='mae', epochs=60, batch_size=16, learning_rates=(2e-4, 2e-4), device='GPU',
                 gpu_memory_growth=True, model_list=None, steps_per_epoch=None, interpolation='inter_area',
                 static_vars=None, checkpoints_frequency=0, save=False, save_path=None, save_logs=False,
                 save_loss_history=True, generator_params={}, discriminator_params={}, verbose=True,
                 math='tf32x3', seed=None):
        # `time_window=True` is the reference's default (sic); True is not > 1, i.e. a spatial model
        super().__init__(backbone=backbone, upsampling=upsampling, data_train=data_train,
                         data_train_lr=data_train_lr, time_window=time_window, loss=loss, batch_size=batch_size,
                         patch_size=patch_size, scale=scale, device=device, gpu_memory_growth=gpu_memory_growth,
                         verbose=verbose, model_list=model_list, save=save, save_path=save_path, show_plot=False)
        self.data_test = data_test
        self.data_test_lr = data_test_lr
        self.predictors_train = predictors_train
        if self.predictors_train is not None and not isinstance(self.predictors_train, list):
            raise TypeError('`predictors_train` must be a list of ndarrays')
        self.predictors_test = predictors_test
        if self.predictors_test is not None and not isinstance(self.predictors_test, list):
            raise TypeError('`predictors_test` must be a list of ndarrays')
        self.epochs = epochs
        self.learning_rates = learning_rates
        self.steps_per_epoch = steps_per_epoch
        self.interpolation = interpolation
        self.static_vars = static_vars
        if self.static_vars is not None:
            for i in range(len(self.static_vars)):
                self.static_vars[i] = getattr(self.static_vars[i], 'values', self.static_vars[i])
        self.checkpoints_frequency = checkpoints_frequency
        self.save_loss_history = save_loss_history
        self.save_logs = save_logs
        self.generator_params = dict(generator_params)
        self.discriminator_params = dict(discriminator_params)
        self.gentotal, self.gengan, self.gen_pxloss, self.disc = [], [], [], []
        if self.time_window is not None and not self.model_is_spatiotemporal:
            self.time_window = None
        if self.model_is_spatiotemporal and self.time_window is None:
            raise ValueError('The argument `time_window` must be a postive integer for spatio-temporal models')
        self.math = math
        self.seed = seed
        self.use_graph = True        # replay the step as CUDA graphs (CGANStep); False = eager train_step per batch
        self._step = None

    def setup_model(self):
        """cgan.py:173-262."""
        n_channels = self.data_train.shape[-1]
        n_aux_channels = 0
        st = self.model_is_spatiotemporal
        if st:                      # cgan.py:179-185: static variables only feed the HR aux branch
            if self.predictors_train is not None:
                n_channels += len(self.predictors_train)
            if self.static_vars is not None:
                n_aux_channels += len(self.static_vars)
        else:
            if self.static_vars is not None:
                n_channels += len(self.static_vars)
                n_aux_channels = len(self.static_vars)
            if self.predictors_train is not None:
                n_channels += len(self.predictors_train)
        if self.patch_size is None:
            lr_h, lr_w = int(self.data_train.shape[1] / self.scale), int(self.data_train.shape[2] / self.scale)
            hr_h, hr_w = int(self.data_train.shape[1]), int(self.data_train.shape[2])
        else:
            lr_h = lr_w = int(self.patch_size / self.scale)
            hr_h = hr_w = int(self.patch_size)
        gp = self.generator_params
        if self.upsampling in POSTUPSAMPLING_METHODS and st:
            self.generator = nets.recnet_postupsampling(
                backbone_block=self.backbone, upsampling=self.upsampling, scale=self.scale, n_channels=n_channels,
                n_aux_channels=n_aux_channels, lr_size=(lr_h, lr_w), time_window=self.time_window, math=self.math, **gp)
            d_size = (lr_h, lr_w)
        elif self.upsampling in POSTUPSAMPLING_METHODS:
            self.generator = nets.net_postupsampling(
                backbone_block=self.backbone, upsampling=self.upsampling, scale=self.scale,
                n_channels=n_channels, n_aux_channels=n_aux_channels, lr_size=(lr_h, lr_w), math=self.math, **gp)
            d_size = (lr_h, lr_w)
        elif st:
            self.generator = nets.recnet_pin(backbone_block=self.backbone, n_channels=n_channels,
                                             n_aux_channels=n_aux_channels, hr_size=(hr_h, hr_w),
                                             time_window=self.time_window, math=self.math, **gp)
            d_size = (hr_h, hr_w)
        elif self.backbone == 'unet':
            self.generator = nets.unet_pin(backbone_block=self.backbone, n_channels=n_channels,
                                           n_aux_channels=n_aux_channels, hr_size=(hr_h, hr_w),
                                           math=self.math, **gp)
            d_size = (hr_h, hr_w)
        else:
            self.generator = nets.net_pin(backbone_block=self.backbone, n_channels=n_channels,
                                          n_aux_channels=n_aux_channels, hr_size=(hr_h, hr_w),
                                          math=self.math, **gp)
            d_size = (hr_h, hr_w)
        # the reference passes lr_size=(lr_height, lr_width) also for 'pin', where the discriminator
        # inputs live on the HR grid (the Keras Input shapes are then only nominal); here the grid
        # the tensors really have is passed so that shape inference is exact
        self.discriminator = nets.residual_discriminator(
            n_channels=n_channels, scale=self.scale, upsampling=self.upsampling,
            is_spatiotemporal=st, lr_size=d_size, math=self.math,
            time_window=self.time_window if st else None, **self.discriminator_params)
        seed = self.seed if self.seed is not None else int(np.random.randint(0, 2 ** 31 - 2))
        dev = self.dp.torch_device
        self.generator.to(dev).init_weights(seed)
        self.discriminator.to(dev).init_weights(seed + 1)
        if self.verbose == 1 and self.running_on_first_worker:
            self.generator.summary()
            self.discriminator.summary()

    def run(self):
        """cgan.py:264-444."""
        self.timing = Timing(self.verbose)
        self.setup_model()
        if isinstance(self.learning_rates, (tuple, list)) and len(self.learning_rates) > 1:
            genlr, dislr = self.learning_rates
        else:
            if isinstance(self.learning_rates, (tuple, list)):
                self.learning_rates = self.learning_rates[0]
            genlr = dislr = self.learning_rates
        self.generator_optimizer = Adam(genlr, beta_1=0.5)
        self.discriminator_optimizer = Adam(dislr, beta_1=0.5)
        if self.predictors_train is not None:
            self.predictors_train = np.concatenate([getattr(p, 'values', p) for p in self.predictors_train], axis=-1)
        self.n = self.data_train.shape[0] - (self.time_window or 0)
        self.indices_train = np.random.permutation(np.arange(self.n))
        if self.steps_per_epoch is None:
            self.steps_per_epoch = int(self.n / self.batch_size)
        self.data_train = getattr(self.data_train, 'values', self.data_train)
        self.data_train_lr = getattr(self.data_train_lr, 'values', self.data_train_lr)
        chatty = bool(self.verbose) and self.running_on_first_worker
        losses = (float('nan'),) * 4
        for epoch in range(self.epochs):
            for i in range(self.steps_per_epoch):
                res = create_batch_hr_lr(
                    self.indices_train, i, self.data_train, self.data_train_lr, upsampling=self.upsampling,
                    scale=self.scale, batch_size=self.batch_size, patch_size=self.patch_size,
                    time_window=self.time_window, static_vars=self.static_vars,
                    predictors=self.predictors_train, interpolation=self.interpolation, time_metadata=None)
                if self.static_vars is not None:
                    [lr_array, aux_hr], [hr_array] = res
                else:
                    [lr_array], [hr_array] = res
                    aux_hr = None
                if self.model_is_spatiotemporal and np.ndim(lr_array) == 4:
                    lr_array = lr_array[..., None]      # the single-variable LR batch lost its channel axis
                if self._step is None and self.use_graph:
                    # static shapes (fixed batch / patch size): capture the step once, replay it per batch
                    self._step = CGANStep(
                        self.generator, self.discriminator, np.shape(lr_array), np.shape(hr_array),
                        np.shape(aux_hr) if aux_hr is not None else None, gen_pxloss_function=self.lossf,
                        learning_rates=(self.generator_optimizer.learning_rate,
                                        self.discriminator_optimizer.learning_rate), dist=self.dp.dist).capture()
                if self._step is not None:
                    losses = self._step.run(lr_array, hr_array, aux_hr, first_batch=(epoch == 0 and i == 0))
                else:
                    losses = train_step(
                        lr_array, hr_array, generator=self.generator, discriminator=self.discriminator,
                        generator_optimizer=self.generator_optimizer,
                        discriminator_optimizer=self.discriminator_optimizer, epoch=epoch,
                        gen_pxloss_function=self.lossf, summary_writer=None,
                        first_batch=(epoch == 0 and i == 0), static_array=aux_hr, dist=self.dp.dist)
            self.gentotal.append(losses[0]); self.gengan.append(losses[1])
            self.gen_pxloss.append(losses[2]); self.disc.append(losses[3])
            if chatty:
                print('Epoch %d/%d - gen_total_loss: %.4f - gen_crosentr_loss: %.4f - gen_px_loss: %.4f - '
                      'disc_loss: %.4f' % ((epoch + 1, self.epochs) + tuple(losses)))
            if self.checkpoints_frequency > 0 and self.running_on_first_worker and \
                    (epoch + 1) % self.checkpoints_frequency == 0:
                ck = os.path.join(self.savecheckpoint_path, 'checkpoints')
                os.makedirs(ck, exist_ok=True)
                # tf.train.Checkpoint(generator, discriminator, both optimizers), cgan.py:288-292,375: weights + Adam slots
                self.generator.save_checkpoint(os.path.join(ck, 'generator_epoch%d.npz' % (epoch + 1)))
                self.discriminator.save_checkpoint(os.path.join(ck, 'discriminator_epoch%d.npz' % (epoch + 1)))
        if self.save_loss_history and self.running_on_first_worker:
            np.save(self.save_path + './losses.npy',
                    np.array((self.gentotal, self.gengan, self.gen_pxloss, self.disc)))
        self.timing.checktime()
        # ---- loss on the test set (first worker)
        if self.predictors_test is not None:
            self.predictors_test = np.concatenate([getattr(p, 'values', p) for p in self.predictors_test], axis=-1)
        self.data_test = getattr(self.data_test, 'values', self.data_test)
        self.data_test_lr = getattr(self.data_test_lr, 'values', self.data_test_lr)
        self.n_test = self.data_test.shape[0] - (self.time_window or 0)
        self.indices_test = np.random.permutation(np.arange(self.n_test))
        if self.running_on_first_worker:
            res = create_batch_hr_lr(
                self.indices_test, 0, self.data_test, self.data_test_lr, upsampling=self.upsampling,
                scale=self.scale, batch_size=self.n_test, patch_size=self.patch_size,
                time_window=self.time_window, static_vars=self.static_vars, predictors=self.predictors_test,
                interpolation=self.interpolation, time_metadata=None)
            if self.static_vars is not None:
                [lr_array, aux_hr], [hr_array] = res
                input_test = [lr_array, aux_hr]
            else:
                [lr_array], [hr_array] = res
                input_test = [lr_array]
            if self.model_is_spatiotemporal and np.ndim(input_test[0]) == 4:
                input_test[0] = input_test[0][..., None]
            y_pred = self.generator.predict(input_test, batch_size=self.batch_size)
            self.test_loss = _host_loss(self.lossf, hr_array, y_pred)
            if self.verbose:
                print('\n%s on the test set: %s' % (self.lossf, self.test_loss))
        self.timing.runtime()
        self.save_results(self.generator, folder_prefix='cgan_')


def load_checkpoint(checkpoint_dir, checkpoint_number, backbone, upsampling, scale, input_height_width,
                    n_static_vars=0, n_predictors=0, time_window=None, n_blocks=(20, 4), n_filters=(8, 32),
                    attention=False, localcon_layer=False, math='tf32x3', device='cuda'):
    """load_checkpoint -- cgan.py:447-522 (same arguments): rebuild generator and discriminator and restore the
    epoch-``checkpoint_number`` files written by ``CGANTrainer(checkpoints_frequency=...)``, optimizer slots
    included.  Returns ``(generator, generator_optimizer, discriminator, discriminator_optimizer)``; the Adam
    objects are bound to the restored arenas (theta / m / v / iteration count)."""
    n_channels, n_aux = 1, 0
    if n_static_vars > 0:
        n_channels += n_static_vars
        n_aux += n_static_vars
    if n_predictors > 0:
        n_channels += n_predictors
    st = time_window is not None and time_window > 1
    if upsampling in POSTUPSAMPLING_METHODS:
        if st:
            generator = nets.recnet_postupsampling(
                backbone_block=backbone, upsampling=upsampling, scale=scale, n_channels=n_channels,
                n_aux_channels=n_aux, n_filters=n_filters[0], n_blocks=n_blocks[0], lr_size=input_height_width,
                n_channels_out=1, time_window=time_window, attention=attention, localcon_layer=localcon_layer, math=math)
        else:
            generator = nets.net_postupsampling(
                backbone_block=backbone, upsampling=upsampling, scale=scale, n_channels=n_channels,
                n_aux_channels=n_aux, n_filters=n_filters[0], n_blocks=n_blocks[0], lr_size=input_height_width,
                n_channels_out=1, attention=attention, localcon_layer=localcon_layer, math=math)
    elif upsampling == 'pin':
        if st:
            generator = nets.recnet_pin(backbone_block=backbone, n_channels=n_channels, n_aux_channels=n_aux,
                                        hr_size=input_height_width, time_window=time_window, n_filters=n_filters[0],
                                        n_blocks=n_blocks[0], n_channels_out=1, attention=attention,
                                        localcon_layer=localcon_layer, math=math)
        else:
            build = nets.unet_pin if backbone == 'unet' else nets.net_pin
            generator = build(backbone_block=backbone, n_channels=n_channels, n_aux_channels=n_aux,
                              hr_size=input_height_width, n_filters=n_filters[0], n_blocks=n_blocks[0], n_channels_out=1,
                              attention=attention, localcon_layer=localcon_layer, math=math)
    else:
        raise ValueError('`upsampling` not recognized')
    discriminator = nets.residual_discriminator(
        n_channels=n_channels, upsampling=upsampling, is_spatiotemporal=st, scale=scale, lr_size=input_height_width,
        n_filters=n_filters[1], n_res_blocks=n_blocks[1], attention=attention, math=math,
        time_window=time_window if st else None)
    ck = os.path.join(checkpoint_dir, 'checkpoints')
    if not os.path.isdir(ck):
        ck = checkpoint_dir
    generator.to(device).load_checkpoint(os.path.join(ck, 'generator_epoch%d.npz' % checkpoint_number))
    discriminator.to(device).load_checkpoint(os.path.join(ck, 'discriminator_epoch%d.npz' % checkpoint_number))
    return generator, Adam(2e-4, beta_1=0.5), discriminator, Adam(2e-4, beta_1=0.5)
