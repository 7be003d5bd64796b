"""Keras weights -> flat arena importer: serve / fine-tune models that the reference trained.

The reference saves ``tf.keras`` models (``model.save(path, save_format='tf')``, training/base.py:177; best-model
checkpoints supervised.py:380-390; ``tf.train.Checkpoint`` cgan.py:288-292,370-382).  Neither TensorFlow nor h5py
exists in this image, so the hand-over format is a plain ``.npz`` written on the REFERENCE side by
:data:`REFERENCE_EXPORT_SNIPPET` (20 lines, run where TF lives): one entry per Keras variable, keyed by its
*structural* name ``<top-level layer name>/<attribute path>/<variable>`` -- attribute names (``conv1``, ``att``,
``conv2x`` ...) are fixed by dl4ds/models/blocks.py, unlike Keras' auto-numbered variable names.

Resolution rules (each cites the reference line that fixes it):
  * layers the builders name explicitly keep their name: ``ResidualBlock3``, ``DenseBlock2``, ``Transition2``,
    ``TransitionLast``, ``ConvBlock_aux``, ``Bottleneck``, ``DecoderConvBlock1``, ``EncoderBlock1``,
    ``SubpixelConvolution[n]`` / ``ResizeConvolution[n]`` / ``Deconvolution[n]``, ``RecurrentConvBlock[n]``,
    ``ResidualBlock2_branch1`` (sp_postups.py:129-205, sp_preups.py:113-304, blocks.py:346,411,475,505,608,
    discriminator.py:36-51);
  * anonymous layers get Keras' auto name ``<snake_case class>[_<counter>]``; counters grow in construction order,
    which is the order of the builder's Python code and therefore of this package's parameter table: the k-th
    anonymous ``conv2d`` of the file is the k-th anonymous convolution here (``stem``, ``backbone_last`` ...;
    sp_postups.py:134,156), likewise ``conv_block`` (``ConvBlock_tail``, ``ConvBlock_out``; :207-212),
    ``transition_block`` (``TransitionSkip``; :163), ``localized_conv_block`` (:185), ``residual_block``
    (``ResidualBlock_merged``; discriminator.py:70), ``dense`` (:78-79);
  * ``TimeDistributed`` wrappers of the recurrent networks (``upsampling_<m>``, ``localized_conv_block``;
    spt_postups.py:131,147) contribute a ``layer/`` path element that is dropped;
  * attribute aliases: ``EncoderBlock.conv`` is the ConvBlock (blocks.py:609) -> dropped;
    ``DeconvolutionBlock.conv2dtranspose1 / 2 / (none)`` (blocks.py:508-516) -> ``deconv_1of2_scale_x2`` /
    ``deconv_2of2_scale_x2`` / ``deconv_scale_x<s>``; DenseBlock re-assigns ``conv1`` / ``conv2`` (blocks.py:246-259,
    SURVEY App. B #19): attribute names resolve to the layers that are really called, the first pair has no variables.
Every array is shape-checked against the parameter table; a missing, surplus or mis-shaped entry raises.
"""
import re
from collections import OrderedDict

import numpy as np

REFERENCE_EXPORT_SNIPPET = '''
# run on the reference side (TensorFlow + dl4ds installed); `model` = trainer.model / trainer.generator / a loaded SavedModel
import numpy as np, tensorflow as tf

def _walk(layer, prefix, out):
    subs = [(k, v) for k, v in vars(layer).items() if isinstance(v, tf.keras.layers.Layer) and not k.startswith('_')]
    own = {id(w) for _, s in subs for w in s.weights}
    for w in layer.weights:
        if id(w) not in own:
            out['|'.join(prefix + [w.name.split('/')[-1].split(':')[0]])] = w.numpy()
    for k, s in subs:
        _walk(s, prefix + [k], out)

def export_structural(model, path):
    out = {}
    for layer in model.layers:
        if layer.weights:
            _walk(layer, [layer.name], out)
    np.savez(path, **out)
'''

# our top-level parameter group -> Keras auto-name stem of the anonymous layer it mirrors
_ANON_TYPES = (
    (re.compile(r'^(stem|backbone_last|branch[12]_(stem|last|down[12]))$'), 'conv2d'),
    (re.compile(r'^dense[12]$'), 'dense'),
    (re.compile(r'^ConvBlock_(tail|out)$'), 'conv_block'),
    (re.compile(r'^TransitionSkip$'), 'transition_block'),
    (re.compile(r'^LocalizedConvBlock$'), 'localized_conv_block'),
    (re.compile(r'^ResidualBlock_merged$'), 'residual_block'),
    (re.compile(r'^branch1_recurrent$'), 'RecurrentConvBlock'),
)
_TD_UPSAMPLERS = {'upsampling_spc': 'SubpixelConvolution', 'upsampling_rc': 'ResizeConvolution',
                  'upsampling_dc': 'Deconvolution'}
_AUTO = re.compile(r'^(?P<stem>[a-z][a-z0-9_]*?)(?:_(?P<n>\d+))?$')


def _anon_type(group):
    for rx, t in _ANON_TYPES:
        if rx.match(group):
            return t
    return None


def _norm_rest(top, rest, scale_hint=None):
    """Keras attribute path (list) below top-level layer ``top`` -> this package's sub-path."""
    rest = [r for r in rest if r != 'layer']                       # TimeDistributed(...).layer
    if top.startswith('EncoderBlock') and rest and rest[0] == 'conv':
        rest = rest[1:]                                            # EncoderBlock.conv = the ConvBlock (blocks.py:609)
    if top.startswith('Deconvolution') and rest:
        alias = {'conv2dtranspose1': 'deconv_1of2_scale_x2', 'conv2dtranspose2': 'deconv_2of2_scale_x2'}
        if rest[0] in alias:
            rest[0] = alias[rest[0]]
        elif rest[0] == 'conv2dtranspose':
            rest[0] = 'deconv_scale_x%s' % (scale_hint if scale_hint is not None else '')
    return rest


def resolve(spec, keras_keys):
    """{our parameter name: keras structural key}.  ``spec``: the model's ordered parameter table; ``keras_keys``:
    iterable of structural names ('/' or '|' separated)."""
    ours = OrderedDict()                                    # group -> [param names]
    for name in spec:
        ours.setdefault(name.split('/')[0], []).append(name)
    keras = OrderedDict()                                   # top-level keras layer -> {tuple(rest): key}
    for key in keras_keys:
        parts = key.replace('|', '/').split('/')
        keras.setdefault(parts[0], {})[tuple(parts[1:])] = key
    # --- top-level matching
    top_of = {}
    for g in ours:
        if g in keras:
            top_of[g] = g
    for td, blk in _TD_UPSAMPLERS.items():                  # recurrent networks: TimeDistributed(upsampler)
        if td in keras and blk in ours and blk not in top_of:
            top_of[blk] = td
    by_type = {}
    for top in keras:
        if top in top_of.values():
            continue
        m = _AUTO.match(top)
        if m:
            by_type.setdefault(m.group('stem'), []).append((int(m.group('n') or 0), top))
    for lst in by_type.values():
        lst.sort()
    anon = {}
    for g in ours:
        if g in top_of:
            continue
        t = _anon_type(g)
        if t is None:
            raise KeyError('no Keras layer named %r in the file and no anonymous-layer rule for it' % g)
        if t == 'RecurrentConvBlock':                       # discriminator.py:31: name_suffix '' -> explicit name
            if t not in keras:
                raise KeyError('missing Keras layer %r for %r' % (t, g))
            top_of[g] = t
            continue
        anon.setdefault(t, []).append(g)
    for t, groups in anon.items():
        have = by_type.get(t, [])
        if len(have) != len(groups):
            raise KeyError('anonymous %r layers: the file has %d (%s), the model needs %d (%s)'
                           % (t, len(have), [h[1] for h in have], len(groups), groups))
        for g, (_, top) in zip(groups, have):
            top_of[g] = top
    # --- per-variable matching
    out = OrderedDict()
    used = set()
    for g, names in ours.items():
        top = top_of[g]
        table = {}
        for rest, key in keras[top].items():
            hint = None
            if g.startswith('Deconvolution'):
                for n in names:
                    m = re.search(r'deconv_scale_x(\d+)', n)
                    if m:
                        hint = m.group(1)
            table[tuple(_norm_rest(g, list(rest), hint))] = key
        for n in names:
            rest = tuple(n.split('/')[1:])
            if rest not in table:
                raise KeyError('parameter %r: no variable %r under Keras layer %r (has %s)'
                               % (n, '/'.join(rest), top, sorted('/'.join(r) for r in table)))
            out[n] = table[rest]
            used.add(table[rest])
    extra = [k for k in keras_keys if k not in used]
    if extra:
        raise KeyError('variables in the file that the model does not have: %s' % extra[:8])
    return out


def load_keras_weights(model, source):
    """Load a structural-name ``.npz`` (or dict) written by :data:`REFERENCE_EXPORT_SNIPPET` into ``model``."""
    if isinstance(source, (str, bytes)) or hasattr(source, 'read'):
        with np.load(source) as z:
            arrays = {k: z[k] for k in z.files}
    else:
        arrays = dict(source)
    mapping = resolve(model.spec, list(arrays))
    weights = OrderedDict()
    for name, key in mapping.items():
        a = np.asarray(arrays[key], np.float32)
        want = tuple(model.spec[name])
        if a.shape != want:
            if a.size == int(np.prod(want)) and name.endswith('localconv/kernel'):
                # LocallyConnected2D(implementation=3) keeps only the non-zero taps as a flat vector ordered by the sorted
                # index tuples (in_row, in_col, in_ch, out_row, out_col, filter); for a 1x1 kernel that is (H,W,Cin,F)
                a = a.reshape(want)
            else:
                raise ValueError('%s <- %s: shape %s, expected %s' % (name, key, a.shape, want))
        weights[name] = a
    model.set_weights(weights)
    return mapping


def export_structural(model, keras_counters=None):
    """The inverse map, for tests and for handing weights BACK to a Keras model: {structural key: array} with the
    names the reference-side snippet would produce for this architecture (anonymous layers numbered from
    ``keras_counters[type]``, default 0, as a fresh Keras session would)."""
    w = model.get_weights()
    counters = dict(keras_counters or {})
    tops = {}
    out = OrderedDict()
    rec = model.name.startswith('rec')
    for name, arr in w.items():
        parts = name.split('/')
        g = parts[0]
        if g not in tops:
            t = _anon_type(g)
            if t == 'RecurrentConvBlock':
                tops[g] = t
            elif t is not None:
                n = counters.get(t, 0)
                counters[t] = n + 1
                tops[g] = t if n == 0 else '%s_%d' % (t, n)
            elif rec and g in _TD_UPSAMPLERS.values():
                tops[g] = [k for k, v in _TD_UPSAMPLERS.items() if v == g][0]
            else:
                tops[g] = g
        rest = parts[1:]
        if g.startswith('EncoderBlock'):
            rest = ['conv'] + rest
        if g.startswith('Deconvolution'):
            alias = {'deconv_1of2_scale_x2': 'conv2dtranspose1', 'deconv_2of2_scale_x2': 'conv2dtranspose2'}
            rest[0] = alias.get(rest[0], 'conv2dtranspose')
        if tops[g].startswith('upsampling_') or (rec and g == 'LocalizedConvBlock'):
            rest = ['layer'] + rest
        out['|'.join([tops[g]] + rest)] = arr
    return out
