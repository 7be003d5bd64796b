"""Checkpoint / resume: weights + Adam slots + iteration count round-trip through
``Model.save_checkpoint`` / ``load_checkpoint`` and through ``training.load_checkpoint`` (dl4ds cgan.py:447-522).
CPU-only: no kernel is launched (the arena is plain storage)."""
import numpy as np
import torch

from dl4ds_b200 import nets
from dl4ds_b200.training import load_checkpoint


def _fill(model, seed):
    g = torch.Generator().manual_seed(seed)
    a = model.arena
    for t in (a.theta, a.m, a.v):
        t.copy_(torch.randn(t.shape, generator=g))
    a.t = 17 + seed


def test_model_checkpoint_roundtrip(tmp_path):
    m = nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (8, 8), n_blocks=2).to('cpu')
    _fill(m, 1)
    path = str(tmp_path / 'ck.npz')
    m.save_checkpoint(path)
    m2 = nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (8, 8), n_blocks=2).to('cpu')
    m2.load_checkpoint(path)
    for name in m.spec:
        assert np.array_equal(m.arena.param(name).numpy(), m2.arena.param(name).numpy())
        assert np.array_equal(m.arena._view(m.arena.m, name).numpy(), m2.arena._view(m2.arena.m, name).numpy())
        assert np.array_equal(m.arena._view(m.arena.v, name).numpy(), m2.arena._view(m2.arena.v, name).numpy())
    assert m2.arena.t == m.arena.t == 18
    # a weights-only file loads too and resets the optimizer
    wpath = str(tmp_path / 'w.npz')
    m.save_weights(wpath)
    m2.load_checkpoint(wpath)
    assert m2.arena.t == 0 and float(m2.arena.m.abs().max()) == 0.0
    assert np.array_equal(m.get_weights()['ConvBlock_out/conv1/kernel'], m2.get_weights()['ConvBlock_out/conv1/kernel'])


def test_cgan_load_checkpoint(tmp_path):
    """Files written the way CGANTrainer(checkpoints_frequency=k) writes them restore into freshly built models."""
    gen = nets.net_postupsampling('resnet', 'spc', 2, 1, 0, (8, 8), n_filters=8, n_blocks=2).to('cpu')
    disc = nets.residual_discriminator(1, 'spc', False, 2, (8, 8), n_filters=8, n_res_blocks=1).to('cpu')
    _fill(gen, 2)
    _fill(disc, 3)
    ck = tmp_path / 'checkpoints'
    ck.mkdir()
    gen.save_checkpoint(str(ck / 'generator_epoch4.npz'))
    disc.save_checkpoint(str(ck / 'discriminator_epoch4.npz'))
    g2, gopt, d2, dopt = load_checkpoint(str(tmp_path), 4, 'resnet', 'spc', 2, (8, 8), n_blocks=(2, 1), n_filters=(8, 8),
                                         device='cpu')
    assert g2.name == 'resnet_spc' and d2.name == 'discriminator'
    for a, b in ((g2, gen), (d2, disc)):
        for name in b.spec:     # (the flat arenas also hold alignment padding, which is not part of the state)
            assert torch.equal(a.arena.param(name), b.arena.param(name))
            assert torch.equal(a.arena._view(a.arena.m, name), b.arena._view(b.arena.m, name))
            assert torch.equal(a.arena._view(a.arena.v, name), b.arena._view(b.arena.v, name))
    assert g2.arena.t == 19 and d2.arena.t == 20
    assert gopt.beta_1 == 0.5 and dopt.learning_rate == 2e-4
