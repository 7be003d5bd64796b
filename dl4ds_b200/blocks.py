"""Graph blocks of the DL4DS hot path as functions over an engine context (``Ctx`` or ``SpecCtx``).

Mirrors the layer classes of the reference's ``dl4ds/models/blocks.py`` (cited per function) for the
configurations on the north-star path and its "next" rows (``normalization`` None, 'bn' or 'ln'; every
``dropout_variant``; model defaults are None / 0, sp_postups.py:26-28).  Parameter names follow ``<layer>/<sublayer>/{kernel,bias}``.
Where the reference fuses nothing, the CUDA path fuses bias + activation (+ residual add)
(+ depth_to_space) into the convolution epilogue.
"""

SUPPORTED_ACTIVATIONS = (None, 'linear', 'relu', 'sigmoid', 'tanh', 'gelu')


def conv_block(c, name, x, filters, activation='relu', attention=False, ks1=3, ks2=3, normalization=None,
               dropout_rate=0, dropout_variant=None):
    """ConvBlock.call -- blocks.py:87-103.  With ``normalization`` ('bn' | 'ln') the convolutions carry no bias
    (:37,44,52,58) and each is followed by the normalisation with the activation fused into it; with
    ``dropout_rate`` > 0 a dropout layer sits in front of each convolution (:89,95-96)."""
    x = c.dropout(x, dropout_rate, dropout_variant)
    if normalization is None:
        y = c.conv(x, name + '/conv1', filters, k=ks1, act=activation)
        y = c.dropout(y, dropout_rate, dropout_variant)
        y = c.conv(y, name + '/conv2', filters, k=ks2, act=activation)
    else:
        y = c.conv(x, name + '/conv1', filters, k=ks1, bias=False)
        y = c.norm(y, name + '/norm1', normalization, act=activation)
        y = c.dropout(y, dropout_rate, dropout_variant)
        y = c.conv(y, name + '/conv2', filters, k=ks2, bias=False)
        y = c.norm(y, name + '/norm2', normalization, act=activation)
    if attention:
        y = c.channel_attention(y, name + '/att')
    return y


def residual_block(c, name, x, filters, activation='relu', attention=False, use_1x1conv=False, normalization=None,
                   dropout_rate=0, dropout_variant=None):
    """ResidualBlock.call -- blocks.py:210-230: act([norm2](conv2([drop](act([norm1](conv1([drop](X))))))) [*att]
    + [conv1x1](X)); the skip takes the un-dropped X."""
    xd = c.dropout(x, dropout_rate, dropout_variant)
    if normalization is not None:
        y = c.conv(xd, name + '/conv1', filters, bias=False)
        y = c.norm(y, name + '/norm1', normalization, act=activation)
        y = c.dropout(y, dropout_rate, dropout_variant)
        y = c.conv(y, name + '/conv2', filters, bias=False)
        y = c.norm(y, name + '/norm2', normalization)
        if attention:
            y = c.channel_attention(y, name + '/att')
        skip = c.conv(x, name + '/conv1x1', filters, k=1) if use_1x1conv else x
        return c.add(y, skip, act=activation)
    y = c.conv(xd, name + '/conv1', filters, act=activation)
    y = c.dropout(y, dropout_rate, dropout_variant)
    skip = c.conv(x, name + '/conv1x1', filters, k=1) if use_1x1conv else x
    if attention:
        y = c.conv(y, name + '/conv2', filters)
        y = c.channel_attention(y, name + '/att')
        return c.add(y, skip, act=activation)
    # residual add + activation fused into conv2's epilogue
    return c.conv(y, name + '/conv2', filters, act=activation, res=skip)


def dense_block(c, name, x, filters, activation='relu', attention=False, normalization=None, dropout_rate=0,
                dropout_variant=None):
    """DenseBlock.call -- blocks.py:262-277.  The pre-activation of X is computed and discarded by
    the reference (:263-267): Y = conv3x3([drop](act([norm2](conv1x1(X))))) [*att]; out = concat([Y, X]).  With a
    normalisation, norm1 still runs on X (its parameters exist, BN moving statistics follow X) and its output is
    dropped -- as is dropout1's; DenseBlock's own conv1 / conv2 keep their bias (:249-259)."""
    if normalization is None:
        y = c.conv(x, name + '/conv1', 4 * filters, k=1, act=activation)
    else:
        c.norm(x, name + '/norm1', normalization, act=activation)
        y = c.conv(x, name + '/conv1', 4 * filters, k=1)
        y = c.norm(y, name + '/norm2', normalization, act=activation)
    y = c.dropout(y, dropout_rate, dropout_variant)
    y = c.conv(y, name + '/conv2', filters, k=3)
    if attention:
        y = c.channel_attention(y, name + '/att')
    return c.concat([y, x])


def convnext_block(c, name, x, filters, activation='gelu', normalization='ln', use_1x1conv=False, drop_path=0.,
                   layer_scale_init_value=0):
    """ConvNextBlock.call -- blocks.py:170-184: 7x7 depthwise conv -> norm (LN with epsilon 1e-6, or BN) -> Dense(4F)
    -> activation -> Dense(F) [-> * gamma, the layer scale, when ``layer_scale_init_value`` > 0 (:166-169,178-179)]
    [-> DropPath(drop_path): whole samples of the branch dropped in training, :106-129,183], added to the
    ([1x1-projected]) input.  The builders pass neither option (sp_postups.py:125-129): defaults 0."""
    if normalization not in ('bn', 'ln'):
        # the reference only creates `self.norm` for 'bn' / 'ln' (:158-164) and calls it unconditionally (:173)
        raise ValueError("ConvNextBlock needs normalization 'bn' or 'ln' (the reference fails in call() with "
                         "%r: blocks.py:158-164,173)" % (normalization,))
    y = c.depthwise_conv(x, name + '/dwconv', 7)
    y = c.norm(y, name + '/norm', normalization, eps=1e-6 if normalization == 'ln' else 1e-3)
    y = c.dense(y, name + '/pwconv1', 4 * filters, act=activation)
    y = c.dense(y, name + '/pwconv2', filters)
    if layer_scale_init_value > 0:
        y = c.channel_scale(y, name + '/gamma', layer_scale_init_value)
    skip = c.conv(x, name + '/conv1x1', filters, k=1) if use_1x1conv else x
    if drop_path and drop_path > 0:
        y = c.dropout(y, drop_path, 'droppath')
    return c.add(skip, y)


def transition_block(c, name, x, filters, activation='relu'):
    """TransitionBlock.call without BN: 1x1 conv then activation -- blocks.py:306-308."""
    return c.conv(x, name + '/conv', filters, k=1, act=activation)


def localized_conv_block(c, name, x, filters=2):
    """LocalizedConvBlock.call -- blocks.py:330-333."""
    y = transition_block(c, name + '/transition', x, filters)
    return c.local_conv(y, name + '/localconv', filters)


def subpixel_block(c, name, x, scale, n_filters):
    """SubpixelConvolutionBlock.call -- blocks.py:433-454.  One shared ``conv2x`` layer serves every
    x2 stage (:415,421-422); depth_to_space is fused into the convolution's store."""
    plan = {2: [2], 4: [2, 2], 8: [2, 2, 2], 10: [2, 5], 20: [2, 2, 5]}.get(scale, [scale])
    for f in plan:
        lname = {2: 'conv2x', 5: 'conv5x'}.get(f, 'conv')
        x = c.conv(x, name + '/' + lname, n_filters * f * f, d2s=f)
    return x


def subpixel_transition(c, name, x, scale, n_filters, tname, t_filters, activation='relu'):
    """SubpixelConvolutionBlock (blocks.py:433-454) followed directly by a TransitionBlock (blocks.py:306-308):
    every x2 stage but the last runs as in :func:`subpixel_block`; the last one is composed with the
    transition's 1x1 convolution (``Ctx.conv_d2s_pointwise``), which removes the widest HR tensor of the graph.
    Falls back to the two separate blocks when the last stage is not x2."""
    plan = {2: [2], 4: [2, 2], 8: [2, 2, 2], 10: [2, 5], 20: [2, 2, 5]}.get(scale, [scale])
    if plan[-1] != 2 or n_filters % 4 or t_filters % 4:
        return transition_block(c, tname, subpixel_block(c, name, x, scale, n_filters), t_filters, activation)
    for f in plan[:-1]:
        lname = {2: 'conv2x', 5: 'conv5x'}.get(f, 'conv')
        x = c.conv(x, name + '/' + lname, n_filters * f * f, d2s=f)
    return c.conv_d2s_pointwise(x, name + '/conv2x', n_filters, tname + '/conv', t_filters, act=activation, r=2)


def resize_conv_block(c, name, x, scale, n_filters, interpolation='bilinear'):
    """ResizeConvolutionBlock.call -- blocks.py:485-491 (interpolation: bilinear, nearest or bicubic)."""
    y = c.resize(x, int(x.H * scale), int(x.W * scale), interpolation)
    return c.conv(y, name + '/conv', n_filters)


def deconv_block(c, name, x, scale, n_filters, output_activation=None):
    """DeconvolutionBlock.call -- blocks.py:522-534, reproducing its if/if/else fall-through:
    scale 8 = T1, T2, T2 (T2 shared); scale 4 additionally runs the stride-4 transpose (x16 total);
    any other scale is a single stride-`scale` transpose."""
    def t(lname, v, stride, a):
        return c.conv_transpose(v, name + '/' + lname, n_filters, 9, stride, act=a)
    if scale == 4:
        x = t('deconv_1of2_scale_x2', x, 2, None)
        x = t('deconv_2of2_scale_x2', x, 2, output_activation)
    if scale == 8:
        x = t('deconv_1of2_scale_x2', x, 2, None)
        x = t('deconv_2of2_scale_x2', x, 2, output_activation)
        x = t('deconv_2of2_scale_x2', x, 2, output_activation)
    else:
        x = t('deconv_scale_x' + str(scale), x, scale, output_activation)
    return x


def recurrent_conv_block(c, name, x, filters, T, activation='relu', normalization=None, dropout_rate=0,
                         dropout_variant=None):
    """RecurrentConvBlock.call -- blocks.py:380-398: [drop] -> ConvLSTM2D 5x5 -> [norm1] -> act -> [drop] ->
    ConvLSTM2D 3x3 -> [norm2] -> act.  ``x`` holds time-major frames (T*B, H, W, C): batch norm over (B,T,H,W) and
    layer norm over C are the same reductions on the folded tensor; the dropout layers are the ``dim=3`` ones
    (:371-374: the spatial variant draws per sample and channel over (T,H,W))."""
    bsz = x.N // T
    y = c.dropout(x, dropout_rate, dropout_variant, n_samples=bsz)
    y = c.convlstm(y, name + '/convlstm1', filters, 5, T)
    y = c.norm(y, name + '/norm1', normalization, act=activation) if normalization else c.act(y, activation)
    y = c.dropout(y, dropout_rate, dropout_variant, n_samples=bsz)
    y = c.convlstm(y, name + '/convlstm2', filters, 3, T)
    y = c.norm(y, name + '/norm2', normalization, act=activation) if normalization else c.act(y, activation)
    return y


def pad_concat(c, t1, t2):
    """PadConcat.call -- blocks.py:629-656 (zero-pad the smaller one at the bottom / right)."""
    H, W = max(t1.H, t2.H), max(t1.W, t2.W)
    return c.concat([c.pad_to(t1, H, W), c.pad_to(t2, H, W)])
