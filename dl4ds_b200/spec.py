"""Shape-only twin of :class:`dl4ds_b200.engine.Ctx`.

Running a model function over a :class:`SpecCtx` performs shape inference on the host and yields
the ordered parameter table ``{name: shape}`` (Keras layouts, creation order) plus MAC counts --
what ``model.summary()`` / ``count_params()`` report in the reference.  No GPU is needed.
"""
from collections import OrderedDict

from .engine import same_pads


class SVar:
    __slots__ = ('N', 'H', 'W', 'C', 'requires_grad')

    def __init__(self, N, H, W, C, requires_grad=True):
        self.N, self.H, self.W, self.C = int(N), int(H), int(W), int(C)
        self.requires_grad = requires_grad

    @property
    def shape(self):
        return (self.N, self.H, self.W, self.C)


class SpecCtx:
    def __init__(self):
        self.spec = OrderedDict()
        self.macs = 0            # multiply-accumulates of one forward pass (conv / dense / local)
        self.macs_dgrad = 0      # ... of the input-gradient passes backward actually runs
        self.layer_macs = {}     # '<layer>@HxW' -> MACs of that convolution application (as executed)
        self.macs_saved = 0      # reference-graph MACs (fwd) that operator composition does not execute
        self.training = False
        self.n_dropout = 0

    def _reg(self, name, shape):
        shape = tuple(int(s) for s in shape)
        if name in self.spec:
            assert self.spec[name] == shape, (name, self.spec[name], shape)   # shared layer
        else:
            self.spec[name] = shape

    def input(self, shape, requires_grad=False):
        return SVar(*shape, requires_grad=requires_grad)

    def _count(self, name, x, macs):
        self.macs += macs
        if x.requires_grad:
            self.macs_dgrad += macs
        self.layer_macs['%s@%dx%d' % (name, x.H, x.W)] = macs

    def conv(self, x, name, cout, k=3, act=None, bias=True, stride=1, padding='same', res=None,
             d2s=1, out=None, dense=False):
        self._reg(name + '/kernel', (x.C, cout) if dense else (k, k, x.C, cout))
        if bias:
            self._reg(name + '/bias', (cout,))
        if padding == 'same':
            Ho, _ = same_pads(x.H, k, stride)
            Wo, _ = same_pads(x.W, k, stride)
        else:
            Ho, Wo = (x.H - k) // stride + 1, (x.W - k) // stride + 1
        self._count(name, x, x.N * Ho * Wo * k * k * x.C * cout)
        r = d2s if d2s > 1 else 1
        return SVar(x.N, Ho * r, Wo * r, cout // (r * r))

    def conv_d2s_pointwise(self, x, name1, cm, name2, co, act=None, r=2, k=3):
        """Shape / parameter / MAC accounting of the composed SubpixelConvolution stage + 1x1 conv.  ``macs``
        keeps counting the REFERENCE graph (both layers, what Keras executes); ``layer_macs`` and
        ``macs_saved`` record what the composed kernel really executes."""
        self._reg(name1 + '/kernel', (k, k, x.C, r * r * cm))
        self._reg(name1 + '/bias', (r * r * cm,))
        self._reg(name2 + '/kernel', (1, 1, cm, co))
        self._reg(name2 + '/bias', (co,))
        m_conv = x.N * x.H * x.W * k * k * x.C * r * r * cm
        m_pw = x.N * x.H * r * x.W * r * cm * co
        m_exec = x.N * x.H * x.W * k * k * x.C * r * r * co
        self.macs += m_conv + m_pw
        self.macs_dgrad += (m_conv if x.requires_grad else 0) + m_pw
        self.macs_saved += m_conv + m_pw - m_exec
        self.layer_macs['%s*%s@%dx%d' % (name1, name2, x.H, x.W)] = m_exec
        return SVar(x.N, x.H * r, x.W * r, co)

    def conv_transpose(self, x, name, cout, k, stride, act=None):
        self._reg(name + '/kernel', (k, k, cout, x.C))
        self._count(name, x, x.N * x.H * x.W * k * k * x.C * cout)
        return SVar(x.N, x.H * stride, x.W * stride, cout)

    def dense(self, x, name, cout, act=None):
        return self.conv(x, name, cout, k=1, act=act, dense=True)

    def add(self, a, b, act=None):
        assert a.shape == b.shape, (a.shape, b.shape)
        return SVar(*a.shape)

    def concat(self, parts):
        p0 = parts[0]
        for p in parts:
            assert (p.N, p.H, p.W) == (p0.N, p0.H, p0.W), [q.shape for q in parts]
        return SVar(p0.N, p0.H, p0.W, sum(p.C for p in parts))

    def act(self, x, act):
        return x

    def norm(self, x, name, kind, act=None, eps=1e-3):
        if kind not in ('bn', 'ln'):
            raise ValueError('Normalization not supported, got %s' % (kind,))
        self._reg(name + '/gamma', (x.C,))
        self._reg(name + '/beta', (x.C,))
        if kind == 'bn':
            self._reg(name + '/moving_mean', (x.C,))
            self._reg(name + '/moving_variance', (x.C,))
        return SVar(*x.shape)

    def depthwise_conv(self, x, name, k=7):
        self._reg(name + '/depthwise_kernel', (k, k, x.C, 1))
        self._reg(name + '/bias', (x.C,))
        self._count(name, x, x.N * x.H * x.W * k * k * x.C)
        return SVar(*x.shape)

    def gelu(self, x):
        return x

    def channel_scale(self, x, name, init_value):
        self._reg(name, (x.C,))
        self.const_init = getattr(self, 'const_init', {})
        self.const_init[name] = float(init_value)
        return SVar(*x.shape)

    def dropout(self, x, rate, variant=None, n_samples=None):
        if rate and rate > 0:
            self.n_dropout += 1        # applications in graph order = the layer ids of the mask generator
        return x

    def channel_attention(self, x, name, r=4, groups=None):
        cr = int(x.C / r)
        self._reg(name + '/conv1/kernel', (1, 1, x.C, cr))
        self._reg(name + '/conv1/bias', (cr,))
        self._reg(name + '/conv2/kernel', (1, 1, cr, x.C))
        self._reg(name + '/conv2/bias', (x.C,))
        ng = groups[0] if groups is not None else x.N
        self.macs += ng * 2 * x.C * cr
        return SVar(*x.shape)

    def local_conv(self, x, name, filters):
        self._reg(name + '/kernel', (x.H, x.W, x.C, filters))
        self._reg(name + '/bias', (x.H, x.W, filters))
        self.macs += x.N * x.H * x.W * x.C * filters
        return SVar(x.N, x.H, x.W, filters)

    def resize_bilinear(self, x, Ho, Wo):
        return SVar(x.N, Ho, Wo, x.C)

    def resize(self, x, Ho, Wo, method='bilinear'):
        return SVar(x.N, Ho, Wo, x.C)

    def maxpool2(self, x):
        return SVar(x.N, x.H // 2, x.W // 2, x.C)

    def pad_to(self, x, H, W):
        return SVar(x.N, H, W, x.C)

    def permute_frames(self, x, A, B):
        return x

    def repeat_frames(self, x, T):
        return SVar(x.N * T, x.H, x.W, x.C)

    def group_mean(self, x, n_groups=None):
        return SVar(x.N if n_groups is None else n_groups, 1, 1, x.C)

    def mul_mask(self, x, mask):
        return x

    def convlstm(self, x, name, filters, k, T):
        self._reg(name + '/kernel', (k, k, x.C, 4 * filters))
        self._reg(name + '/recurrent_kernel', (k, k, filters, 4 * filters))
        self._reg(name + '/bias', (4 * filters,))
        self.macs += x.N * x.H * x.W * k * k * (x.C + filters) * 4 * filters
        return SVar(x.N, x.H, x.W, filters)
