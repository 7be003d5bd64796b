"""SupervisedTrainer -- dl4ds/training/supervised.py:28-416 with the Keras/Horovod runtime
replaced by the CUDA engine:

  ``model.compile`` + ``model.fit``  ->  one captured CUDA graph per optimizer step
                                         (``step.SupervisedStep``: zero-grad, forward, MAE/MSE +
                                         gradient seed, backward, Adam) replayed per batch;
  ``hvd.DistributedOptimizer``       ->  one NCCL all-reduce(sum) of the flat gradient arena per
                                         step, 1/size folded into the Adam kernel;
  ``BroadcastGlobalVariablesCallback(0)`` -> one broadcast of theta / m / v from rank 0;
  LR x ``hvd.size()``                ->  ``lr_scale = world size`` (supervised.py:338-352).

Batches come from the same ``DataGenerator`` (numpy/cv2 semantics of the reference); they are
staged through pinned host buffers and copied H2D asynchronously.
"""
import os

import numpy as np

from .. import nets
from ..dataloader import DataGenerator
from ..utils import POSTUPSAMPLING_METHODS, Timing
from .base import Trainer


class History:
    """Stand-in for ``keras.callbacks.History`` (``.history`` dict of per-epoch lists)."""

    def __init__(self):
        self.history = {'loss': [], 'val_loss': []}
        self.epoch = []


class SupervisedTrainer(Trainer):
    """Procedure for training the supervised residual models (same signature as the reference,
    supervised.py:31-72; ``math`` and ``seed`` are additions: convolution math mode
    'fp32' | 'tf32x3' | 'tf32' and the seed of the Glorot initialiser)."""

    def __init__(self, backbone, upsampling, data_train, data_val, data_test, data_train_lr=None,
                 data_val_lr=None, data_test_lr=None, predictors_train=None, predictors_val=None,
                 predictors_test=None, static_vars=None, scale=5, interpolation='inter_area',
                 patch_size=None, time_window=None, batch_size=64, loss='mae', epochs=60,
                 steps_per_epoch=None, test_steps=None, validation_steps=None, device='GPU',
                 gpu_memory_growth=True, use_multiprocessing=False, model_list=None,
                 learning_rate=(1e-3, 1e-4), lr_decay_after=1e5, early_stopping=False, patience=6,
                 min_delta=0, show_plot=True, save=False, save_path=None, save_bestmodel=False,
                 trained_model=None, trained_epochs=0, verbose=True, math='tf32x3', seed=None,
                 data_on_device=False, **architecture_params):
        super().__init__(backbone=backbone, upsampling=upsampling, data_train=data_train,
                         data_train_lr=data_train_lr, time_window=time_window, loss=loss,
                         batch_size=batch_size, patch_size=patch_size, scale=scale, device=device,
                         gpu_memory_growth=gpu_memory_growth, use_multiprocessing=use_multiprocessing,
                         verbose=verbose, model_list=model_list, save=save, save_path=save_path,
                         show_plot=show_plot)
        self.data_val = data_val
        self.data_test = data_test
        self.data_val_lr = data_val_lr
        self.data_test_lr = data_test_lr
        self.predictors_train = predictors_train
        if self.predictors_train is not None and not isinstance(self.predictors_train, list):
            raise TypeError('`predictors_train` must be a list of ndarrays')
        self.predictors_test = predictors_test
        if self.predictors_test is not None and not isinstance(self.predictors_test, list):
            raise TypeError('`predictors_test` must be a list of ndarrays')
        self.predictors_val = predictors_val
        if self.predictors_val is not None and not isinstance(self.predictors_val, list):
            raise TypeError('`predictors_val` must be a list of ndarrays')
        self.static_vars = static_vars
        if self.static_vars is not None:
            for i in range(len(self.static_vars)):
                self.static_vars[i] = getattr(self.static_vars[i], 'values', self.static_vars[i])
        self.interpolation = interpolation
        self.epochs = epochs
        self.steps_per_epoch = steps_per_epoch
        self.validation_steps = validation_steps
        self.test_steps = test_steps
        self.learning_rate = learning_rate
        self.lr_decay_after = lr_decay_after
        self.early_stopping = early_stopping
        self.patience = patience
        self.min_delta = min_delta
        self.architecture_params = architecture_params
        self.trained_model = trained_model
        self.trained_epochs = trained_epochs
        self.save_bestmodel = save_bestmodel
        self.math = math
        self.seed = seed
        self.data_on_device = data_on_device    # addition: keep the training array in HBM (DeviceDataGenerator)
        self.model = None
        self.train_step = None
        self._pinned = None

    # ------------------------------------------------------------------------------------------
    def setup_datagen(self):
        """supervised.py:220-240."""
        p = dict(backbone=self.backbone, upsampling=self.upsampling, scale=self.scale,
                 batch_size=self.global_batch_size, static_vars=self.static_vars,
                 patch_size=self.patch_size, interpolation=self.interpolation,
                 time_window=self.time_window)
        from ..dataloader import DeviceDataGenerator
        if self.data_on_device and DeviceDataGenerator.supported(
                getattr(self.data_train, 'values', self.data_train), self.data_train_lr, self.upsampling, self.scale,
                self.patch_size, self.time_window, self.static_vars, self.predictors_train, self.interpolation):
            self.ds_train = DeviceDataGenerator(self.data_train, self.data_train_lr, device=self.dp.torch_device,
                                                predictors=self.predictors_train, **p)
        else:
            self.ds_train = DataGenerator(self.data_train, self.data_train_lr, predictors=self.predictors_train, **p)
        self.ds_val = DataGenerator(self.data_val, self.data_val_lr, predictors=self.predictors_val, **p)
        self.ds_test = DataGenerator(self.data_test, self.data_test_lr, predictors=self.predictors_test, **p)

    def setup_model(self):
        """supervised.py:242-325: channel bookkeeping + builder dispatch, then the device-side
        optimizer step (what ``compile`` prepares in Keras)."""
        n_channels = self.data_train.shape[-1]
        n_aux_channels = 0
        if self.model_is_spatiotemporal:
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

        ap = dict(self.architecture_params)
        if self.trained_model is None:
            if self.upsampling in POSTUPSAMPLING_METHODS:
                if self.model_is_spatiotemporal:
                    self.model = nets.recnet_postupsampling(
                        backbone_block=self.backbone, upsampling=self.upsampling, scale=self.scale,
                        n_channels=n_channels, n_aux_channels=n_aux_channels, lr_size=(lr_h, lr_w),
                        time_window=self.time_window, math=self.math, **ap)
                else:
                    self.model = nets.net_postupsampling(
                        backbone_block=self.backbone, upsampling=self.upsampling, scale=self.scale,
                        lr_size=(lr_h, lr_w), n_channels=n_channels, n_aux_channels=n_aux_channels,
                        math=self.math, **ap)
            elif self.upsampling == 'pin':
                if self.model_is_spatiotemporal:          # supervised.py:300-307
                    self.model = nets.recnet_pin(
                        backbone_block=self.backbone, n_channels=n_channels, n_aux_channels=n_aux_channels,
                        hr_size=(hr_h, hr_w), time_window=self.time_window, math=self.math, **ap)
                elif self.backbone == 'unet':
                    self.model = nets.unet_pin(
                        backbone_block=self.backbone, n_channels=n_channels, n_aux_channels=n_aux_channels,
                        hr_size=(hr_h, hr_w), math=self.math, **ap)
                else:
                    self.model = nets.net_pin(
                        backbone_block=self.backbone, n_channels=n_channels, n_aux_channels=n_aux_channels,
                        hr_size=(hr_h, hr_w), math=self.math, **ap)
            self.model.to(self.dp.torch_device)
            self.model.init_weights(seed=self.seed if self.seed is not None
                                    else int(np.random.randint(0, 2 ** 31 - 1)))
            if self.verbose == 1 and self.running_on_first_worker:
                self.model.summary()
        else:
            self.model = self.trained_model
            self.model.to(self.dp.torch_device)
            print('Loading pre-trained model')
        if getattr(self.model.arena, '_rng', None) is None:
            # dropout masks: seeded from the trainer seed (OS entropy when unseeded) + the rank, so replicas and
            # unseeded runs draw independent masks; a resumed model keeps the RNG state its checkpoint carried
            base = self.seed if self.seed is not None else int.from_bytes(os.urandom(6), 'little')
            self.model.arena.seed_rng((int(base) * 2654435761 + 7919 * self.dp.rank) & 0x7FFFFFFFFFFFFFFF)
        self._compile()

    def _batch_shapes(self):
        B = self.global_batch_size
        T = self.time_window if self.model_is_spatiotemporal else None
        shapes = []
        for s in self.model.input_shapes:
            if len(s) == 4:
                shapes.append((s[0] * B,) + tuple(s[1:]))       # time-major frames (T*B,H,W,C)
            else:
                shapes.append((B,) + tuple(s))
        oh, ow, oc = self.model.output_shape
        tgt = ((T * B) if T else B, oh, ow, oc)
        return shapes, tgt

    def _compile(self):
        """Adam + schedule (supervised.py:336-353) and the captured step graph."""
        from ..step import EvalStep, LRSchedule, SupervisedStep
        sched = LRSchedule(self.learning_rate, self.lr_decay_after, scale=float(self.dp.size))
        shapes, tgt = self._batch_shapes()
        self.train_step = SupervisedStep(self.model, shapes, tgt, loss=self.lossf, lr=sched, math=self.math)
        self.train_step.broadcast_from_rank0()
        self.train_step.capture()
        self.eval_step = EvalStep(self.model, loss=self.lossf, math=self.math)
        self.optimizer = self.train_step

    # ------------------------------------------------------------------------------------------
    def _stage(self, inputs, target):
        """Host numpy batch -> pinned staging buffers (allocated once) -> async H2D into the step's
        static device buffers.  Spatio-temporal batches (B,T,...) become time-major frames."""
        import torch
        st = self.train_step
        T = self.time_window if self.model_is_spatiotemporal else None
        if self._pinned is None:
            self._pinned = ([torch.empty(t.shape, dtype=torch.float32).pin_memory() for t in st.inputs],
                            torch.empty(st.target.shape, dtype=torch.float32).pin_memory())
        pins, ptgt = self._pinned

        def put(dst, src, fold):
            a = np.asarray(src, dtype=np.float32)
            if fold:
                a = np.swapaxes(a, 0, 1).reshape(dst.shape)
            dst.numpy()[...] = a.reshape(dst.shape)
        for dst, src, s in zip(pins, inputs, self.model.input_shapes):
            put(dst, src, len(s) == 4)
        put(ptgt, target, T is not None)
        st.load_batch(pins, ptgt)

    def train_on_batch(self, inputs, target):
        """keras ``Model.train_on_batch``: one optimizer step on a HOST batch ``([lr(, aux)], hr)``;
        returns the batch loss as a Python float (device -> host read)."""
        self._stage(inputs, target)
        return float(self.train_step.run().item())

    def train_on_batches(self, batches):
        """Pipelined ``fit`` loop body: iterate ``batches`` (host ``([lr(, aux)], hr)`` pairs) and yield one
        float loss per batch, in order.  While the GPU runs step i the host already converts batch i+1 into the
        second pinned staging set and enqueues its H2D copy on a copy stream; the loss of step i is read back
        (async D2H + event) after step i+1 has been queued.  Every batch still crosses PCIe and every loss still
        comes back to the host -- only the waiting overlaps.  (Keras ``fit`` overlaps the same way through its
        ``Sequence`` worker, supervised.py:397-409.)"""
        import torch
        st = self.train_step
        dev = self.dp.torch_device
        T = self.time_window if self.model_is_spatiotemporal else None
        if getattr(self, '_pipe', None) is None:
            mk = lambda t: torch.empty(t.shape, dtype=torch.float32)
            self._pipe = {
                'pin': [([mk(t).pin_memory() for t in st.inputs], mk(st.target).pin_memory()) for _ in range(2)],
                'dev': [([torch.empty_like(t) for t in st.inputs], torch.empty_like(st.target)) for _ in range(2)],
                'loss': [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)],
                'h2d': [torch.cuda.Event() for _ in range(2)],
                'done': [torch.cuda.Event() for _ in range(2)],
                'copy': torch.cuda.Stream(device=dev),
            }
        P = self._pipe
        main = torch.cuda.current_stream(dev)

        def put(dst, src, fold):
            a = np.asarray(src, dtype=np.float32)
            if fold:
                a = np.swapaxes(a, 0, 1)
            dst.numpy()[...] = a.reshape(dst.shape)

        def stage(i, inputs, target):
            k = i & 1
            P['h2d'][k].synchronize()                       # the slot's previous H2D copy has left the pinned buffers
            pins, ptgt = P['pin'][k]
            for dst, src, shp in zip(pins, inputs, self.model.input_shapes):
                put(dst, src, len(shp) == 4)
            put(ptgt, target, T is not None)
            devs, dtgt = P['dev'][k]
            with torch.cuda.stream(P['copy']):
                P['copy'].wait_event(P['done'][k])          # step i-2 no longer reads this device slot
                for d, s_ in zip(devs, pins):
                    d.copy_(s_, non_blocking=True)
                dtgt.copy_(ptgt, non_blocking=True)
                P['h2d'][k].record(P['copy'])

        def launch(i):
            k = i & 1
            main.wait_event(P['h2d'][k])
            devs, dtgt = P['dev'][k]
            st.load_batch(devs, dtgt)                        # D2D into the captured graph's static buffers
            loss = st.run()
            P['loss'][k].copy_(loss, non_blocking=True)
            P['done'][k].record(main)

        it = iter(batches)
        i = 0
        pending = None
        for inputs, target in it:
            stage(i, inputs, target)
            launch(i)
            if pending is not None:
                P['done'][pending & 1].synchronize()
                yield float(P['loss'][pending & 1][0])
            pending = i
            i += 1
        if pending is not None:
            P['done'][pending & 1].synchronize()
            yield float(P['loss'][pending & 1][0])

    def train_on_device_batches(self, gen, order):
        """The fit loop over a :class:`DeviceDataGenerator`: batch ``order[s]`` is gathered / coarsened on the
        GPU straight into the step's static buffers; yields one float loss per step (read back one step late)."""
        import torch
        st = self.train_step
        slots = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        evs = [torch.cuda.Event() for _ in range(2)]
        pending = None
        for i, idx in enumerate(order):
            gen.fill(int(idx), st.inputs[0], st.target, st.inputs[1] if len(st.inputs) > 1 else None)
            loss = st.run()
            slots[i & 1].copy_(loss, non_blocking=True)
            evs[i & 1].record()
            if pending is not None:
                evs[pending & 1].synchronize()
                yield float(slots[pending & 1][0])
            pending = i
        if pending is not None:
            evs[pending & 1].synchronize()
            yield float(slots[pending & 1][0])

    def test_on_batch(self, inputs, target):
        import torch
        dev = self.dp.torch_device
        ins, _, _ = self.model._prep_inputs(inputs, dev)
        tgt = torch.as_tensor(np.asarray(target, dtype=np.float32)).to(dev)
        if tgt.dim() == 5:
            tgt = tgt.transpose(0, 1).reshape(-1, *tgt.shape[2:]).contiguous()
        return float(self.eval_step.run(ins, tgt).item())

    def evaluate(self, ds, steps=None):
        n = len(ds) if steps is None else min(steps, len(ds))
        if n == 0:
            return float('nan')
        tot = 0.0
        for i in range(n):
            x, y = ds[i]
            tot += self.test_on_batch(x, y[0])
        return tot / n

    # ------------------------------------------------------------------------------------------
    def run(self):
        """Compiling, training and saving the model (supervised.py:328-416)."""
        self.timing = Timing(self.verbose)
        self.setup_datagen()
        self.setup_model()
        steps = self.steps_per_epoch
        if steps is not None and self.dp.size > 1:
            steps = steps // self.dp.size
        n_train = len(self.ds_train) if steps is None else steps
        self.fithist = History()
        best, wait = float('inf'), 0
        chatty = bool(self.verbose) and self.running_on_first_worker
        for epoch in range(self.trained_epochs, self.epochs):
            order = np.random.permutation(len(self.ds_train))      # Keras fit(shuffle=True) on a Sequence
            def epoch_batches():
                for s in range(n_train):
                    x, y = self.ds_train[int(order[s % len(order)])]
                    yield x, y[0]
            tot = 0.0
            from ..dataloader import DeviceDataGenerator
            if isinstance(self.ds_train, DeviceDataGenerator):
                losses = self.train_on_device_batches(self.ds_train, [order[s % len(order)] for s in range(n_train)])
            else:
                losses = self.train_on_batches(epoch_batches())
            for lval in losses:
                tot += lval
            loss = tot / max(n_train, 1)
            val = self.evaluate(self.ds_val, self.validation_steps)
            if self.dp.size > 1:
                # every rank draws its own validation batches: average val_loss over the ranks so that the
                # early-stopping / best-model decisions below are identical everywhere (a rank that left the epoch
                # loop alone would leave the others waiting in the next gradient all-reduce)
                val = self.dp.allreduce_mean_scalar(val)
            self.fithist.history['loss'].append(loss)
            self.fithist.history['val_loss'].append(val)
            self.fithist.epoch.append(epoch)
            if chatty:
                print('Epoch %d/%d - loss: %.4f - val_loss: %.4f' % (epoch + 1, self.epochs, loss, val))
            if self.save_bestmodel and self.running_on_first_worker and val <= min(self.fithist.history['val_loss']):
                # ModelCheckpoint(save_best_only=True, monitor='val_loss'), supervised.py:380-390 -- with the
                # optimizer slots, so that `trained_model` + `trained_epochs` resume the same Adam trajectory
                os.makedirs(self.savecheckpoint_path, exist_ok=True)
                self.model.save_checkpoint(os.path.join(self.savecheckpoint_path, 'best_model.npz'))
            if self.early_stopping:
                if val < best - self.min_delta:
                    best, wait = val, 0
                else:
                    wait += 1
                    if wait >= self.patience:
                        if chatty:
                            print('Epoch %d: early stopping' % (epoch + 1))
                        break
        if self.running_on_first_worker:
            self.test_loss = self.evaluate(self.ds_test, self.test_steps)
            if self.verbose:
                print('\nScore on the test set: %s' % self.test_loss)
            self.timing.runtime()
        self.save_results(self.model)

    # ------------------------------------------------------------------------------------------
    # accounting used by bench.py
    def layer_macs(self, batch):
        """{'<layer>:<pass>@HxW': MACs of ONE launch at `batch` samples} for fwd / dgrad / wgrad."""
        out = {}
        for key, macs in self.model.layer_macs_per_sample.items():
            name, hw = key.rsplit('@', 1)
            for p in ('fwd', 'dgrad', 'wgrad'):
                out['%s:%s@%s' % (name, p, hw)] = macs * batch
        return out

    def train_macs_per_sample(self):
        """forward + input-gradient (all layers but those fed by a graph input) + weight-gradient."""
        return 2 * self.model.macs_per_sample + self.model.macs_dgrad_per_sample
