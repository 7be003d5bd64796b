"""Network graph builders + the ``Model`` object the trainers / ``Predictor`` hold.

Each builder keeps the signature and the model naming (``<backbone>_<upsampling>``,
``rec<backbone>_<upsampling>``, ``discriminator``) of its reference counterpart in
``dl4ds/models/`` and returns a :class:`Model` whose graph function runs over the CUDA engine
(``engine.Ctx``) or, for shape/parameter inference, over ``spec.SpecCtx``.

Scope (SURVEY.md sections 8a and 8f row 3): backbones convnet / resnet / densenet / unet / convnext (and the
recurrent ConvLSTM networks), ``normalization`` None / 'bn' / 'ln', every ``dropout_variant``, activations in
{None, relu, sigmoid, tanh, gelu}, every ``rc_interpolation`` of Keras ``Resizing``.  Anything else raises
``NotImplementedError`` (or the reference's own ``ValueError``) -- there is no fallback path.
"""
import math
from collections import OrderedDict

import numpy as np
import torch

from . import blocks as B
from .engine import Arena, Ctx
from .spec import SpecCtx
from .utils import checkarg_dropout_variant

BACKBONES = ('convnet', 'resnet', 'densenet', 'unet', 'convnext')
POSTUPSAMPLING_METHODS = ('spc', 'rc', 'dc')
RC_INTERPOLATIONS = ('bilinear', 'nearest', 'bicubic', 'area', 'lanczos3', 'lanczos5', 'gaussian',
                     'mitchellcubic')                          # Keras Resizing's eight (blocks.py:463-465)


def _check_common(activation, output_activation, normalization, dropout_rate, backbone_block=None,
                  dropout_variant=None, dropout_built=True):
    for a in (activation, output_activation):
        if a not in B.SUPPORTED_ACTIVATIONS:
            raise NotImplementedError('activation %r is outside the B200 hot path' % (a,))
    if normalization not in (None, 'bn', 'ln'):
        raise ValueError('Normalization not supported, got %s' % (normalization,))      # blocks.py:64-65
    checkarg_dropout_variant(dropout_variant)
    if dropout_rate and not dropout_built:
        raise NotImplementedError('dropout_rate>0 in the recurrent (ConvLSTM) networks is not built')
    if backbone_block is not None and backbone_block not in BACKBONES:
        raise NotImplementedError('backbone %r is outside the B200 hot path' % (backbone_block,))


class _InputSpec:
    """Stand-in for ``keras.Model.input``: ``Predictor`` only reads ``.shape`` (inference.py:173)."""

    def __init__(self, shape):
        self.shape = tuple(shape)


class Model:
    """A built network: graph function + flat parameter arena.

    Exposes what the reference's callers use on a ``tf.keras.Model``: ``name``, ``input.shape``,
    ``predict(inputs, batch_size, verbose)``, ``count_params()``, ``summary()``, ``get_weights`` /
    ``set_weights`` by name, ``save_weights`` / ``load_weights`` (npz).
    """

    def __init__(self, name, fn, input_shapes, time_window=None, math='fp32'):
        self.name = name
        self.fn = fn                                # fn(ctx, [Var...]) -> Var
        self.input_shapes = [tuple(s) for s in input_shapes]   # without batch dim, None allowed
        self.time_window = time_window
        self.math = math
        self.input = _InputSpec((None,) + self.input_shapes[0])
        sc = SpecCtx()
        ins = [sc.input(self._spec_shape(s)) for s in self.input_shapes]
        out = fn(sc, ins)
        self.spec = sc.spec
        self.const_init = dict(getattr(sc, 'const_init', {}))     # parameters with a constant initial value (layer scale)
        n0 = ins[0].N if len(self.input_shapes[0]) == 3 else 1    # traced batch (T folded into N)
        self.macs_per_sample = sc.macs // n0 if len(self.input_shapes[0]) == 3 else sc.macs
        self.macs_dgrad_per_sample = sc.macs_dgrad
        self.layer_macs_per_sample = dict(sc.layer_macs)
        self.macs_saved_per_sample = sc.macs_saved       # forward MACs of the reference graph removed by composition
        self.output_shape = out.shape[1:]
        self.arena = None

    def _spec_shape(self, s):
        """Concrete (N,H,W,C) used for tracing; spatio-temporal inputs (T,H,W,C) fold T into N."""
        if len(s) == 4:
            t, h, w, ch = s
            return ((t or self.time_window or 1), h or 32, w or 32, ch)
        h, w, ch = s
        return (1, h or 32, w or 32, ch)

    # -- parameters -------------------------------------------------------------------------
    def count_params(self):
        return sum(int(np.prod(s)) for s in self.spec.values())

    def summary(self, print_fn=print):
        print_fn('Model: "%s"' % self.name)
        groups = OrderedDict()
        for n, s in self.spec.items():
            groups.setdefault(n.split('/')[0], 0)
            groups[n.split('/')[0]] += int(np.prod(s))
        for g, n in groups.items():
            print_fn('  %-40s %12d' % (g, n))
        print_fn('Total params: %d' % self.count_params())

    @staticmethod
    def _resolve_device(device):
        """torch.device with an explicit index: 'cuda' means the current device ('cuda' != 'cuda:0' for torch)."""
        d = torch.device(device)
        if d.type == 'cuda' and d.index is None:
            d = torch.device('cuda', torch.cuda.current_device() if torch.cuda.is_available() else 0)
        return d

    def to(self, device='cuda'):
        """Place the parameter arena on ``device``.  A no-op when it already lives on that physical device; a real
        move carries the whole optimizer state (theta, Adam m / v, iteration count, dropout RNG state) and drops
        every captured CUDA graph / packed weight image, which hold raw pointers into the old arena."""
        dev = self._resolve_device(device)
        if self.arena is None:
            self.arena = Arena(self.spec, dev)
        elif self._resolve_device(self.arena.device) != dev:
            old = self.arena
            new = Arena(self.spec, dev)
            for name in ('theta', 'm', 'v'):
                getattr(new, name).copy_(getattr(old, name))
            new.t = old.t
            if getattr(old, '_rng', None) is not None:
                new._rng = old._rng.to(dev)
            self.arena = new
            self._predict_graphs = {}
            self._pack_cache = {}
            self.arena_generation = getattr(self, 'arena_generation', 0) + 1   # captured steps re-capture on a change
        return self

    def init_weights(self, seed=0):
        """Keras default initialisers from a seeded numpy generator: glorot_uniform kernels
        (fans include the receptive field), zero biases, ConvLSTM forget-gate bias 1
        (unit_forget_bias), orthogonal-free: recurrent kernels also glorot (seeded synthetic runs;
        trained weights are loaded by name)."""
        rng = np.random.default_rng(seed)
        w = OrderedDict()
        for name, shape in self.spec.items():
            if name in self.const_init:                              # ConvNextBlock layer scale: layer_scale_init_value
                w[name] = np.full(shape, self.const_init[name], np.float32)
            elif name.endswith(('/gamma', '/moving_variance')):      # BatchNormalization / LayerNormalization: ones
                w[name] = np.ones(shape, np.float32)
            elif name.endswith(('/beta', '/moving_mean')):
                w[name] = np.zeros(shape, np.float32)
            elif name.endswith('/bias'):
                a = np.zeros(shape, np.float32)
                if 'convlstm' in name:
                    f = shape[0] // 4
                    a[f:2 * f] = 1.0
                w[name] = a
            elif name.endswith('localconv/kernel'):
                h, wd, cin, f = shape
                lim = math.sqrt(6.0 / (h * wd * cin + h * wd * f))
                w[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
            else:
                rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
                lim = math.sqrt(6.0 / (shape[-2] * rf + shape[-1] * rf))
                w[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        self.set_weights(w)
        return self

    def set_weights(self, weights):
        self.to(self.arena.device if self.arena is not None else 'cuda')
        self.arena.load(weights)

    def get_weights(self):
        return self.arena.state_dict()

    def save_weights(self, path):
        np.savez(path, **{k.replace('/', '|'): v for k, v in self.get_weights().items()})

    def load_weights(self, path):
        with np.load(path) as z:
            self.set_weights({k.replace('|', '/'): z[k] for k in z.files})

    save = save_weights

    def load_keras_weights(self, source):
        """Weights of a reference-trained Keras model, exported on the reference side with
        ``keras_import.REFERENCE_EXPORT_SNIPPET`` (structural variable names); see keras_import.py."""
        from .keras_import import load_keras_weights
        load_keras_weights(self, source)
        return self

    def save_checkpoint(self, path):
        """Weights AND optimizer state of the flat arena (theta, Adam m / v per parameter, iteration count) in one
        ``.npz`` -- what ``tf.train.Checkpoint(generator=..., generator_optimizer=...)`` stores in the reference
        (cgan.py:288-292,375): resuming from it continues the Adam trajectory exactly."""
        a = self.arena
        out = {'__iterations__': np.asarray(a.t, np.int64), '__model_name__': np.asarray(self.name)}
        if getattr(a, '_rng', None) is not None:
            out['__rng__'] = a._rng.detach().cpu().numpy()          # dropout Philox {seed, step}
        for n in a.spec:
            key = n.replace('/', '|')
            out[key] = a.param(n).detach().cpu().numpy()
            out['__adam_m__' + key] = a._view(a.m, n).detach().cpu().numpy()
            out['__adam_v__' + key] = a._view(a.v, n).detach().cpu().numpy()
        np.savez(path, **out)

    def load_checkpoint(self, path):
        """Restore ``save_checkpoint`` output (a weights-only ``save_weights`` file also loads: slots reset)."""
        import torch
        with np.load(path) as z:
            files = set(z.files)
            self.set_weights({n: z[n.replace('/', '|')] for n in self.spec})
            a = self.arena
            full = '__iterations__' in files
            a.t = int(z['__iterations__']) if full else 0
            if '__rng__' in files:
                a.seed_rng(int(z['__rng__'][0]), int(z['__rng__'][1]))
            for n in a.spec:
                key = n.replace('/', '|')
                for flat, tag in ((a.m, '__adam_m__'), (a.v, '__adam_v__')):
                    view = a._view(flat, n)
                    if full and tag + key in files:
                        view.copy_(torch.as_tensor(z[tag + key]).to(view.device))
                    else:
                        view.zero_()
        return self

    # -- execution --------------------------------------------------------------------------
    def _prep_inputs(self, inputs, device):
        """numpy / torch NHWC (or NTHWC) -> list of CUDA fp32 tensors; returns (tensors, B, T)."""
        if not isinstance(inputs, (list, tuple)):
            inputs = [inputs]
        out = []
        bsz, T = None, None
        for i, (x, s) in enumerate(zip(inputs, self.input_shapes)):
            x = torch.as_tensor(x)
            x = x.to(device=device, dtype=torch.float32)
            if len(s) == 4:                       # (B,T,H,W,C) -> time-major frames (T*B,H,W,C)
                assert x.dim() == 5, 'expected a 5-D spatio-temporal input'
                bsz, T = x.shape[0], x.shape[1]
                x = x.transpose(0, 1).reshape(T * bsz, *x.shape[2:])
            else:
                assert x.dim() == 4, 'expected a 4-D NHWC input'
                if bsz is None:
                    bsz = x.shape[0]
            out.append(x.contiguous())
        return out, bsz, T

    def forward(self, inputs, training=False, math=None, timers=None, pack_stream=None, prepacked=None,
                wgrad_stream=None):
        """Run the graph on CUDA tensors prepared by ``_prep_inputs``.  Returns (ctx, out Var).  ``prepacked``: keys
        of a :class:`engine.PackPlan` that has just run on this stream."""
        ctx = Ctx(self.arena, math or self.math, training=training)
        ctx.timers = timers
        if not hasattr(self, '_pack_cache'):
            self._pack_cache = {}
        ctx.pack_cache = self._pack_cache
        ctx.pack_stream = pack_stream
        ctx.prepacked = prepacked
        ctx.wgrad_stream = wgrad_stream
        vs = [ctx.input(x) for x in inputs]
        out = self.fn(ctx, vs)
        return ctx, out

    def _finish_output(self, out_t, bsz, T):
        """(T*B,H,W,C) time-major frames -> (B,T,H,W,C) for spatio-temporal models."""
        if T is None:
            return out_t
        return out_t.reshape(T, bsz, *out_t.shape[1:]).transpose(0, 1).contiguous()

    def __call__(self, inputs, training=False):
        dev = self.arena.device
        ins, bsz, T = self._prep_inputs(inputs, dev)
        _, out = self.forward(ins, training=False)
        return self._finish_output(out.t.contiguous(), bsz, T)

    def predict(self, inputs, batch_size=32, verbose=0):
        """keras ``Model.predict`` (inference.py:238): forward in chunks of ``batch_size``.  Spatial models run every
        full chunk through ONE captured CUDA graph of the forward pass (static input buffers, weight images re-packed
        from the live parameters inside the graph), with pinned staging: the host converts chunk i+1 and copies chunk
        i-1's result out while chunk i computes.  A trailing partial chunk (and spatio-temporal models) run eagerly."""
        self.to('cuda')
        if not isinstance(inputs, (list, tuple)):
            inputs = [inputs]
        n = len(inputs[0])
        outs = []
        start = 0
        spatial = all(len(s_) == 3 for s_ in self.input_shapes) and all(np.ndim(x) == 4 for x in inputs)
        if spatial and n >= batch_size and self.arena.device.type == 'cuda':
            outs, start = self._predict_graphed(inputs, batch_size, verbose)
        for s in range(start, n, batch_size):
            chunk = [x[s:s + batch_size] for x in inputs]
            y = self(chunk)
            outs.append(y.cpu().numpy())
            if verbose:
                print('%d/%d' % (min(s + batch_size, n), n))
        return np.concatenate(outs, axis=0)

    def _predict_graphed(self, inputs, batch_size, verbose):
        """All full chunks through a captured forward graph; returns (list of output arrays, samples consumed)."""
        dev = self.arena.device
        shapes = tuple((batch_size,) + tuple(np.shape(x)[1:]) for x in inputs)
        cache = self.__dict__.setdefault('_predict_graphs', {})
        key = (shapes, self.math)
        if key not in cache:
            xin = [torch.zeros(shp, dtype=torch.float32, device=dev) for shp in shapes]
            _, o = self.forward(xin, training=False)                     # warm-up (module load, allocator)
            torch.cuda.synchronize()
            stream = torch.cuda.Stream()
            stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(stream):
                g = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream()
                with torch.cuda.graph(g, stream=stream):
                    ctx, o = self.forward(xin, training=False, pack_stream=side)
                    ctx.join_pack_stream()
                    y = o.t.contiguous()
            torch.cuda.current_stream().wait_stream(stream)
            pin_in = [[torch.empty(shp, dtype=torch.float32).pin_memory() for shp in shapes] for _ in range(2)]
            pin_out = [torch.empty(tuple(y.shape), dtype=torch.float32).pin_memory() for _ in range(2)]
            cache[key] = (g, xin, y, pin_in, pin_out, [torch.cuda.Event() for _ in range(2)])
        g, xin, y, pin_in, pin_out, evs = cache[key]
        n_full = (len(inputs[0]) // batch_size) * batch_size
        outs = []
        for i, s in enumerate(range(0, n_full, batch_size)):
            k = i & 1
            if i >= 2:
                evs[k].synchronize()                                     # slot k's previous result has landed
                outs.append(pin_out[k].numpy().copy())
            for dst, x in zip(pin_in[k], inputs):
                dst.numpy()[...] = np.asarray(x[s:s + batch_size], dtype=np.float32)
            for d, src in zip(xin, pin_in[k]):
                d.copy_(src, non_blocking=True)
            g.replay()
            pin_out[k].copy_(y, non_blocking=True)
            evs[k].record()
            if verbose:
                print('%d/%d' % (s + batch_size, len(inputs[0])))
        nchunks = n_full // batch_size
        for i in range(max(0, nchunks - 2), nchunks):                    # drain the last (up to) two slots in order
            evs[i & 1].synchronize()
            outs.append(pin_out[i & 1].numpy().copy())
        return outs, n_full


# ---------------------------------------------------------------------------------------------
# shared sections
# ---------------------------------------------------------------------------------------------
def _backbone(c, x_in, backbone_block, n_filters, n_blocks, attention, activation, normalization=None,
              dropout_rate=0, dropout_variant=None):
    """Stem + N blocks + last conv + long skip -- sp_postups.py:132-168, sp_preups.py:116-151."""
    init_n_filters = n_filters
    if backbone_block == 'convnext':      # sp_postups.py:120-131, sp_preups.py:104-115: 7x7 stem, no last conv
        x = b = c.conv(x_in, 'stem', n_filters, k=7)
        for i in range(n_blocks):
            n_filters = init_n_filters * (i + 1)
            b = B.convnext_block(c, 'ConvNextBlock%d' % (i + 1), b, n_filters, activation, normalization,
                                 use_1x1conv=(i != 0))
        x = B.transition_block(c, 'TransitionSkip', x, n_filters, activation)
        return c.add(x, b), n_filters
    x = b = c.conv(x_in, 'stem', n_filters)
    for i in range(n_blocks):
        n_filters = init_n_filters * (i + 1)
        if backbone_block == 'convnet':
            b = B.conv_block(c, 'ConvBlock%d' % (i + 1), b, n_filters, activation, attention,
                             normalization=normalization, dropout_rate=dropout_rate, dropout_variant=dropout_variant)
        elif backbone_block == 'resnet':
            b = B.residual_block(c, 'ResidualBlock%d' % (i + 1), b, n_filters, activation, attention,
                                 use_1x1conv=(i != 0), normalization=normalization, dropout_rate=dropout_rate,
                                 dropout_variant=dropout_variant)
        elif backbone_block == 'densenet':
            b = B.dense_block(c, 'DenseBlock%d' % (i + 1), b, n_filters, activation, attention,
                              normalization=normalization, dropout_rate=dropout_rate, dropout_variant=dropout_variant)
            b = B.transition_block(c, 'Transition%d' % (i + 1), b, b.C // 2)
    b = c.conv(b, 'backbone_last', n_filters, act=activation)
    b = c.dropout(b, dropout_rate, dropout_variant)                     # sp_postups.py:158
    if backbone_block == 'convnet':
        x = b
    elif backbone_block == 'resnet':
        x = B.transition_block(c, 'TransitionSkip', x, n_filters, activation)
        x = c.add(x, b)
    elif backbone_block == 'densenet':
        x = c.concat([x, b])
        x = B.transition_block(c, 'TransitionBackboneLast', x, n_filters, activation)
    return x, n_filters


def _tail(c, x, s_in, init_n_filters, n_filters_aux, n_channels_out, activation, output_activation,
          localcon_layer, transition_done=False, normalization=None, convnext=False, dropout_rate=0):
    """LCB, HR aux branch, TransitionLast, ConvBlock(att), ConvBlock(out)
    -- sp_postups.py:184-212, sp_preups.py:155-183,291-309.  ``transition_done``: TransitionLast was already
    applied by the caller (composed with the last sub-pixel stage)."""
    nz = normalization
    ks = 7 if convnext else 3             # sp_postups.py:121,133,207-212: the tail ConvBlocks use the backbone's `ks`
    if transition_done:
        x = B.conv_block(c, 'ConvBlock_tail', x, init_n_filters, activation=None, attention=True, normalization=nz,
                         ks1=ks, ks2=ks, dropout_rate=dropout_rate)
        return B.conv_block(c, 'ConvBlock_out', x, n_channels_out, activation=output_activation, normalization=nz,
                            ks1=ks, ks2=ks)
    if localcon_layer:
        lws = B.localized_conv_block(c, 'LocalizedConvBlock', x, 2)
        x = c.concat([x, lws])
    if s_in is not None:
        if convnext:                      # sp_postups.py:191-195
            s = B.convnext_block(c, 'ConvNextBlock_aux', s_in, n_filters_aux, activation, nz, use_1x1conv=True)
        else:
            s = B.conv_block(c, 'ConvBlock_aux', s_in, n_filters_aux, activation=activation, normalization=nz)
        x = c.concat([x, s])
    x = B.transition_block(c, 'TransitionLast', x, init_n_filters)      # default relu
    # (the reference passes the rate but not the variant here: plain Dropout, sp_postups.py:206-208)
    x = B.conv_block(c, 'ConvBlock_tail', x, init_n_filters, activation=None, attention=True, normalization=nz,
                     ks1=ks, ks2=ks, dropout_rate=dropout_rate)
    x = B.conv_block(c, 'ConvBlock_out', x, n_channels_out, activation=output_activation, normalization=nz,
                     ks1=ks, ks2=ks)
    return x


# ---------------------------------------------------------------------------------------------
# builders (signatures follow dl4ds/models/*.py)
# ---------------------------------------------------------------------------------------------
def net_postupsampling(backbone_block, upsampling, scale, n_channels, n_aux_channels, lr_size,
                       n_channels_out=1, n_filters=8, n_blocks=6, dropout_rate=0,
                       dropout_variant=None, normalization=None, attention=False, activation='relu',
                       output_activation=None, rc_interpolation='bilinear', localcon_layer=False,
                       math='fp32', fuse_spc_transition=True):
    """net_postupsampling -- sp_postups.py:14-217.  ``fuse_spc_transition`` (an addition): compose the last
    sub-pixel stage with TransitionLast when nothing sits between them (same function, same parameters)."""
    _check_common(activation, output_activation, normalization, dropout_rate, backbone_block, dropout_variant)
    if upsampling not in POSTUPSAMPLING_METHODS:
        raise ValueError('`upsampling` must be one of %s' % (POSTUPSAMPLING_METHODS,))
    if upsampling == 'rc' and rc_interpolation not in RC_INTERPOLATIONS:
        raise NotImplementedError('rc_interpolation=%r is not built (%s are)' % (rc_interpolation, RC_INTERPOLATIONS))
    h_lr, w_lr = lr_size
    aux = n_aux_channels > 0

    def fn(c, inputs):
        x, nf = _backbone(c, inputs[0], backbone_block, n_filters, n_blocks, attention, activation, normalization,
                          dropout_rate, dropout_variant)
        fused = False
        if upsampling == 'spc':
            if fuse_spc_transition and not aux and not localcon_layer:
                # nothing sits between the sub-pixel block and TransitionLast (sp_postups.py:172-205):
                # compose the last x2 stage with the 1x1 convolution
                x = B.subpixel_transition(c, 'SubpixelConvolution', x, scale, nf, 'TransitionLast', n_filters)
                fused = True
            else:
                x = B.subpixel_block(c, 'SubpixelConvolution', x, scale, nf)
        elif upsampling == 'rc':
            x = B.resize_conv_block(c, 'ResizeConvolution', x, scale, nf, rc_interpolation)
        else:
            x = B.transition_block(c, 'TransitionDC', x, n_filters, activation)
            x = B.deconv_block(c, 'Deconvolution', x, scale, nf, activation)
        return _tail(c, x, inputs[1] if aux else None, n_filters, nf, n_channels_out, activation,
                     output_activation, localcon_layer, transition_done=fused, normalization=normalization,
                     convnext=(backbone_block == 'convnext'), dropout_rate=dropout_rate)

    ups_total = scale
    if upsampling == 'dc' and scale == 4:
        ups_total = 16          # reference fall-through, blocks.py:525-533
    shapes = [(h_lr, w_lr, n_channels)]
    if aux:
        shapes.append((int(h_lr * ups_total), int(w_lr * ups_total), n_aux_channels))
    return Model(backbone_block + '_' + upsampling, fn, shapes, math=math)


def net_pin(backbone_block, n_channels, n_aux_channels, hr_size, n_channels_out=1, n_filters=8,
            n_blocks=6, dropout_rate=0, dropout_variant=None, normalization=None, attention=False,
            activation='relu', output_activation=None, localcon_layer=False, math='fp32'):
    """net_pin -- sp_preups.py:13-189."""
    _check_common(activation, output_activation, normalization, dropout_rate, backbone_block, dropout_variant)
    aux = n_aux_channels > 0

    def fn(c, inputs):
        x, nf = _backbone(c, inputs[0], backbone_block, n_filters, n_blocks, attention, activation, normalization,
                          dropout_rate, dropout_variant)
        return _tail(c, x, inputs[1] if aux else None, n_filters, nf, n_channels_out, activation,
                     output_activation, localcon_layer, normalization=normalization,
                     convnext=(backbone_block == 'convnext'), dropout_rate=dropout_rate)

    shapes = [(hr_size[0], hr_size[1], n_channels)]
    if aux:
        shapes.append((hr_size[0], hr_size[1], n_aux_channels))
    return Model(backbone_block + '_pin', fn, shapes, math=math)


def _check_nblocks(shape, power):
    """_check_nblocks -- sp_preups.py:318-324."""
    while shape[0] // 2 ** power < 2 or shape[1] // 2 ** power < 2:
        print('`n_blocks` is too large, cannot downsample %d times given the input grid size. '
              'Setting `n_blocks` to %d' % (power, power - 1))
        power -= 1
    return power


def unet_pin(backbone_block, n_channels, n_aux_channels, hr_size, n_channels_out, n_filters, n_blocks,
             activation='relu', dropout_rate=0, dropout_variant=None, normalization=None,
             attention=False, decoder_upsampling='rc', rc_interpolation='bilinear',
             output_activation=None, width_cap=256, localcon_layer=False, math='fp32'):
    """unet_pin -- sp_preups.py:192-315."""
    _check_common(activation, output_activation, normalization, dropout_rate, backbone_block, dropout_variant)
    if decoder_upsampling not in POSTUPSAMPLING_METHODS:
        raise ValueError('`decoder_upsampling` must be one of %s' % (POSTUPSAMPLING_METHODS,))
    if decoder_upsampling == 'rc' and rc_interpolation not in RC_INTERPOLATIONS:
        raise NotImplementedError('rc_interpolation=%r is not built (%s are)' % (rc_interpolation, RC_INTERPOLATIONS))
    n_blocks = _check_nblocks(hr_size, n_blocks)
    aux = n_aux_channels > 0

    def fn(c, inputs):
        x = inputs[0]
        nf = n_filters
        skips, flist = [], []
        for i in range(n_blocks):       # EncoderBlock: ConvBlock then 2x2 max-pool (blocks.py:602-618)
            y = B.conv_block(c, 'EncoderBlock%d' % (i + 1), x, nf, activation, attention, normalization=normalization)
            skips.append(y)
            x = c.maxpool2(y)
            flist.append(nf)
            nf = min(width_cap, nf * 2)
        x = B.conv_block(c, 'Bottleneck', x, nf, activation, dropout_rate=dropout_rate,
                         dropout_variant=dropout_variant)       # sp_preups.py:265-268 (encoder blocks: rate 0, :255)
        for j, skip in enumerate(reversed(skips)):
            nf = flist[::-1][j]
            if decoder_upsampling == 'spc':
                x = B.subpixel_block(c, 'SubpixelConvolution%d' % (j + 1), x, 2, nf)
            elif decoder_upsampling == 'rc':
                x = B.resize_conv_block(c, 'ResizeConvolution%d' % (j + 1), x, 2, nf, rc_interpolation)
            else:
                x = B.deconv_block(c, 'Deconvolution%d' % (j + 1), x, 2, nf, activation)
            x = B.pad_concat(c, x, skip)
            x = B.conv_block(c, 'DecoderConvBlock%d' % (j + 1), x, nf, activation, attention,
                             normalization=normalization)
        x = c.dropout(x, dropout_rate, dropout_variant)                 # sp_preups.py:287
        return _tail(c, x, inputs[1] if aux else None, n_filters, nf, n_channels_out, activation,
                     output_activation, localcon_layer, normalization=normalization, dropout_rate=dropout_rate)

    shapes = [(hr_size[0], hr_size[1], n_channels)]
    if aux:
        shapes.append((hr_size[0], hr_size[1], n_aux_channels))
    return Model(backbone_block + '_pin', fn, shapes, math=math)


def recnet_postupsampling(backbone_block, upsampling, scale, n_channels, n_aux_channels, lr_size,
                          time_window, n_channels_out=1, n_filters=8, n_blocks=4, dropout_rate=0,
                          dropout_variant=None, normalization=None, attention=False,
                          activation='relu', output_activation=None, rc_interpolation='bilinear',
                          localcon_layer=False, math='fp32'):
    """recnet_postupsampling -- spt_postups.py:12-163.  Inputs (B,T,h,w,C) [+ (B,H,W,n_aux)];
    output (B,T,H,W,n_channels_out).  Internally frames are time-major (T*B,H,W,C)."""
    _check_common(activation, output_activation, normalization, dropout_rate, backbone_block, dropout_variant)
    if backbone_block == 'unet':
        raise ValueError('unet backbone is not compatible with post-upsampling')
    if upsampling == 'rc' and rc_interpolation not in RC_INTERPOLATIONS:
        raise NotImplementedError('rc_interpolation=%r is not built (%s are)' % (rc_interpolation, RC_INTERPOLATIONS))
    if upsampling not in POSTUPSAMPLING_METHODS + ('pin',):
        raise ValueError('`upsampling` must be one of %s' % (POSTUPSAMPLING_METHODS,))
    T = int(time_window)
    aux = n_aux_channels > 0
    h_lr, w_lr = lr_size

    def fn(c, inputs):
        x_in = inputs[0]
        bsz = x_in.N // T
        nz = normalization
        x = b = B.recurrent_conv_block(c, 'RecurrentConvBlock1', x_in, n_filters, T, activation, nz)
        for i in range(n_blocks):
            b = B.recurrent_conv_block(c, 'RecurrentConvBlock%d' % (i + 2), b, n_filters, T, activation, nz,
                                       dropout_rate, dropout_variant)
        b = c.dropout(b, dropout_rate, dropout_variant, n_samples=bsz)      # dim=3 layer, spt_postups.py:113
        if backbone_block == 'convnet':
            x = b
        elif backbone_block == 'resnet':
            x = c.add(x, b)
        else:
            x = c.concat([x, b])
        nf_ups = x.C
        # TimeDistributed(upsampler): frames are independent images
        if upsampling == 'spc':
            x = B.subpixel_block(c, 'SubpixelConvolution', x, scale, nf_ups)
        elif upsampling == 'rc':
            x = B.resize_conv_block(c, 'ResizeConvolution', x, scale, nf_ups, rc_interpolation)
        elif upsampling == 'dc':
            x = B.deconv_block(c, 'Deconvolution', x, scale, nf_ups, None)    # no activation passed
        # ('pin': the samples already live on the HR grid, spt_preups.py:100-118)
        if aux:
            # ConvBlock on the static HR field, then tf.repeat over time (spt_postups.py:135-141):
            # the static branch is time- and (per-sample) batch-dependent only through s_in
            s = B.conv_block(c, 'ConvBlock_aux', inputs[1], n_filters, activation, attention)
            s = c.repeat_frames(s, T)
            x = c.concat([x, s])
        if localcon_layer:
            lws = B.localized_conv_block(c, 'LocalizedConvBlock', x, 2)
            x = c.concat([x, lws])
        # spt_postups.py:150 halves the channels; spt_preups.py:133 goes straight to n_filters
        x = B.transition_block(c, 'TransitionLast', x, n_filters if upsampling == 'pin' else x.C // 2)
        # ConvBlock(n_filters, activation=None, attention=True) on the 5-D tensor: the attention's
        # reduce_mean over axes [1,2] pools (T,H) and keeps W (blocks.py:587)
        # (with dropout_rate > 0: plain Dropout in front of each convolution -- the variant is not passed, :152-153;
        #  with a normalisation: bias-free convolutions, each followed by the normalisation)
        y = c.dropout(x, dropout_rate)
        y = c.conv(y, 'ConvBlock_tail/conv1', n_filters, bias=nz is None)
        if nz:
            y = c.norm(y, 'ConvBlock_tail/norm1', nz)
        y = c.dropout(y, dropout_rate)
        y = c.conv(y, 'ConvBlock_tail/conv2', n_filters, bias=nz is None)
        if nz:
            y = c.norm(y, 'ConvBlock_tail/norm2', nz)
        yb = c.permute_frames(y, T, bsz)                       # batch-major (B*T,H,W,C)
        yb = c.channel_attention(yb, 'ConvBlock_tail/att', groups=(bsz * y.W, T * y.H, y.W))
        y = c.permute_frames(yb, bsz, T)                       # back to time-major
        return B.conv_block(c, 'ConvBlock_out', y, n_channels_out, activation=output_activation, normalization=nz)

    shapes = [(T, h_lr, w_lr, n_channels)]
    if aux:
        shapes.append((int(h_lr * scale), int(w_lr * scale), n_aux_channels))
    return Model('rec' + backbone_block + '_' + upsampling, fn, shapes, time_window=T, math=math)


def recnet_pin(backbone_block, n_channels, n_aux_channels, hr_size, time_window, n_channels_out=1, n_filters=8,
               n_blocks=6, normalization=None, dropout_rate=0, dropout_variant=None, attention=False,
               activation='relu', output_activation=None, localcon_layer=False, math='fp32'):
    """recnet_pin -- spt_preups.py:12-163: the recurrent (ConvLSTM) network on samples that were interpolated to
    the HR grid beforehand.  Same graph as ``recnet_postupsampling`` without the upsampler; TransitionLast maps to
    ``n_filters``.  Inputs (B,T,H,W,C) [+ (B,H,W,n_aux)]; output (B,T,H,W,n_channels_out)."""
    m = recnet_postupsampling(backbone_block, 'pin', 1, n_channels, n_aux_channels, hr_size, time_window,
                              n_channels_out=n_channels_out, n_filters=n_filters, n_blocks=n_blocks,
                              dropout_rate=dropout_rate, dropout_variant=dropout_variant, normalization=normalization,
                              attention=attention, activation=activation, output_activation=output_activation,
                              localcon_layer=localcon_layer, math=math)
    return m


def residual_discriminator(n_channels, upsampling, is_spatiotemporal, scale, lr_size, n_filters=8,
                           n_res_blocks=4, normalization=None, activation='relu', attention=False,
                           math='fp32', time_window=None):
    """residual_discriminator -- discriminator.py:11-81.  ResidualBlocks always use relu (the ``activation``
    argument only reaches the recurrent stem, :31-33,36-38).  Dropout(0.4) on the pooled features is applied with a
    caller-supplied keep mask as third input (training) or skipped.

    Spatio-temporal variant (``is_spatiotemporal``, discriminator.py:25-33,42-47,73-74): inputs (B,T,h,w,C) and
    (B,T,H,W,1), held as time-major frames; the LR branch starts with RecurrentConvBlock(normalization='ln'), every
    Conv2D / ResidualBlock acts per frame (Keras Conv2D on a 5-D tensor folds the leading axes), and
    GlobalAveragePooling3D pools (T,H,W).  ``time_window`` (an addition) fixes T for shape inference; the reference's
    Input has T = None."""
    if normalization not in (None, 'bn', 'ln'):
        raise ValueError('Normalization not supported, got %s' % (normalization,))
    if is_spatiotemporal and attention:
        raise NotImplementedError('ChannelAttention2D on the 5-D tensors of the spatio-temporal discriminator is not built')
    if is_spatiotemporal and activation not in B.SUPPORTED_ACTIVATIONS:
        raise NotImplementedError('activation %r is outside the B200 hot path' % (activation,))
    T = int(time_window) if (is_spatiotemporal and time_window) else None
    if is_spatiotemporal and not T:
        raise ValueError('the spatio-temporal discriminator needs `time_window`')

    def fn(c, inputs):
        x_in, x_ref = inputs[0], inputs[1]
        mask = inputs[2] if len(inputs) > 2 else None
        if is_spatiotemporal:
            x1 = b = B.recurrent_conv_block(c, 'branch1_recurrent', x_in, n_filters, T, activation, 'ln')
        else:
            x1 = b = c.conv(x_in, 'branch1_stem', n_filters)
        for i in range(n_res_blocks):
            b = B.residual_block(c, 'ResidualBlock%d_branch1' % (i + 1), b, n_filters, 'relu', attention,
                                 normalization=normalization)
        x1 = c.conv(b, 'branch1_last', n_filters, res=x1)
        x2 = cc = c.conv(x_ref, 'branch2_stem', n_filters)
        for i in range(n_res_blocks):
            cc = B.residual_block(c, 'ResidualBlock%d_branch2' % (i + 1), cc, n_filters, 'relu', attention,
                                  normalization=normalization)
        if upsampling in POSTUPSAMPLING_METHODS:
            if scale == 5:
                cc = c.conv(cc, 'branch2_down1', n_filters, stride=2, padding='valid')
                x2 = c.conv(cc, 'branch2_down2', n_filters, stride=2, padding='valid')
                x2 = c.pad_to(x2, x2.H - 1, x2.W - 1)            # Cropping2D(((0,1),(0,1)))
            elif scale == 4:
                cc = c.conv(cc, 'branch2_down1', n_filters, stride=2)
                x2 = c.conv(cc, 'branch2_down2', n_filters, stride=2)
            else:
                x2 = c.resize_bilinear(cc, lr_size[0], lr_size[1])
        else:
            x2 = c.conv(cc, 'branch2_last', n_filters, res=x2)
        x = c.concat([x1, x2])
        x = B.residual_block(c, 'ResidualBlock_merged', x, x.C, 'relu', attention, normalization=normalization)
        if is_spatiotemporal:
            bsz = x.N // T
            x = c.group_mean(c.permute_frames(x, T, bsz), n_groups=bsz)     # GlobalAveragePooling3D over (T,H,W)
        else:
            x = c.group_mean(x)
        if mask is not None:
            x = c.mul_mask(x, mask)
        x = c.dense(x, 'dense1', 32, act='sigmoid')
        return c.dense(x, 'dense2', 1, act='sigmoid')

    if upsampling in POSTUPSAMPLING_METHODS:
        in_hw = (lr_size[0], lr_size[1])
        ref_hw = (int(lr_size[0] * scale), int(lr_size[1] * scale))
    else:
        in_hw = ref_hw = (lr_size[0], lr_size[1])
    if is_spatiotemporal:
        return Model('discriminator', fn, [(T,) + in_hw + (n_channels,), (T,) + ref_hw + (1,)], time_window=T,
                     math=math)
    return Model('discriminator', fn, [in_hw + (n_channels,), ref_hw + (1,)], math=math)
