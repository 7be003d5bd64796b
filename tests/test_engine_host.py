"""Host-side bookkeeping of the define-by-run engine that needs no GPU: the activation descriptor (a Var is a channel
slice of an NHWC buffer) and the 16-byte pitch rule for concatenations / accumulated gradients."""
import torch

from dl4ds_b200.engine import Var, padded_channels


def test_padded_channels():
    assert [padded_channels(c) for c in (1, 2, 3, 4, 5, 8, 9, 48, 50, 98, 100)] == [1, 2, 3, 4, 8, 8, 12, 48, 52, 100, 100]


def test_var_like_keeps_channels_and_pads_pitch():
    x = Var(torch.zeros(2, 4, 4, 98))
    assert (x.C, x.ld, x.off) == (98, 98, 0)
    g = x.like(pad=True)
    assert (g.N, g.H, g.W, g.C) == (2, 4, 4, 98) and g.ld == 100 and g.off == 0
    assert x.like().ld == 98                          # dense by default: ops that take no pitch rely on it
    s = g.slice(48, 48)                               # a 4-channel-aligned slice of a padded buffer stays 16-byte aligned
    assert (s.C, s.off, s.ld) == (48, 48, 100) and (s.off * 4) % 16 == 0
    assert Var(torch.zeros(1, 2, 2, 2)).like(pad=True).ld == 2      # narrow tensors are left alone
