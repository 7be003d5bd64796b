"""TEST INFRASTRUCTURE ONLY -- generate the committed fixtures under tests/golden/ by running the
UNMODIFIED reference (imported from /root/reference under stubs, oracle/ref_datapath.py) on
seeded inputs.  Run in the builder container:  ``python -m oracle.make_golden``.

Fixtures (inputs are re-generated from the seed by the tests, only outputs are stored):
  datapath_<case>.npz : outputs of dl4ds.create_batch_hr_lr for the BASELINE.json config shapes
  utils_misc.npz      : spatiotemporal_to_spatial_samples, resize_array variants, crop_array
"""
import os

import numpy as np

from .ref_datapath import load_reference

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def make_inputs(case):
    """Seeded inputs per case -- the tests call this very function to rebuild them."""
    rng = np.random.default_rng({'cfg1': 1234, 'cfg3': 1235, 'cfg5': 1236, 'cfg4': 1237,
                                 'patch': 1238, 'lrgiven': 1239}[case])
    if case == 'cfg1':      # resnet/spc x4, 32->128, single channel
        return dict(array=rng.standard_normal((6, 128, 128, 1), dtype=np.float32), array_lr=None,
                    upsampling='spc', scale=4, batch_size=2, index=1)
    if case == 'cfg3':      # dc x8, 1 static + 3 predictors
        return dict(array=rng.standard_normal((4, 64, 64, 1), dtype=np.float32), array_lr=None,
                    upsampling='dc', scale=8, batch_size=2, index=0,
                    static_vars=[rng.standard_normal((64, 64)).astype(np.float32)],
                    predictors=np.concatenate(
                        [rng.standard_normal((4, 64, 64, 1), dtype=np.float32) for _ in range(3)], -1))
    if case == 'cfg5':      # pin x4, 1 static (coarsen then re-interpolate to HR)
        return dict(array=rng.standard_normal((4, 64, 64, 1), dtype=np.float32), array_lr=None,
                    upsampling='pin', scale=4, batch_size=2, index=1,
                    static_vars=[rng.standard_normal((64, 64)).astype(np.float32)],
                    interpolation='bicubic')
    if case == 'cfg4':      # spatio-temporal rc x4, time_window 3, one predictor (App. B #2)
        return dict(array=rng.standard_normal((8, 32, 32, 1), dtype=np.float32), array_lr=None,
                    upsampling='rc', scale=4, batch_size=2, index=0, time_window=3,
                    predictors=rng.standard_normal((8, 32, 32, 1), dtype=np.float32))
    if case == 'patch':     # random 16x16 crops (np.random.seed(7) right before the call)
        return dict(array=rng.standard_normal((4, 48, 48, 1), dtype=np.float32), array_lr=None,
                    upsampling='spc', scale=4, batch_size=3, index=0, patch_size=16)
    if case == 'lrgiven':   # explicit LR array given
        return dict(array=rng.standard_normal((4, 32, 32, 1), dtype=np.float32),
                    array_lr=rng.standard_normal((4, 8, 8, 1), dtype=np.float32),
                    upsampling='spc', scale=4, batch_size=2, index=1)
    raise KeyError(case)


CASES = ['cfg1', 'cfg3', 'cfg5', 'cfg4', 'patch', 'lrgiven']


def main():
    ref = load_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    for case in CASES:
        kw = make_inputs(case)
        n = kw['array'].shape[0] - (kw.get('time_window') or 0)
        idx = np.arange(n)[::-1].copy()           # a fixed non-trivial permutation
        np.random.seed(7)
        res = ref.create_batch_hr_lr(idx, kw.pop('index'), **kw)
        ins, outs = res
        save = {'lr': ins[0], 'hr': outs[0]}
        if len(ins) > 1:
            save['aux'] = ins[1]
        np.savez_compressed(os.path.join(GOLDEN, 'datapath_%s.npz' % case), **save)
        print(case, {k: (v.shape, str(v.dtype)) for k, v in save.items()})

    rng = np.random.default_rng(99)
    a5 = rng.standard_normal((5, 3, 4, 4, 1)).astype(np.float32)
    misc = {'st2s': ref.spatiotemporal_to_spatial_samples(a5, 3)}
    img = rng.standard_normal((3, 8, 12, 2)).astype(np.float32)
    for interp in ref.INTERPOLATION_METHODS:
        misc['resize_up_' + interp] = ref.resize_array(img, (24, 16), interp, squeezed=False)
        misc['resize_dn_' + interp] = ref.resize_array(img, (6, 4), interp, squeezed=False)
    misc['crop'] = ref.crop_array(img, 4, yx=(2, 3))
    np.savez_compressed(os.path.join(GOLDEN, 'utils_misc.npz'), **misc)
    print('utils_misc', sorted(misc))


if __name__ == '__main__':
    main()
