"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- import the UNMODIFIED reference package
from /root/reference with TensorFlow/Keras/xarray/plotting stubbed out, so that its numpy/cv2
data path (create_batch_hr_lr, create_pair_hr_lr, crop_array, resize_array, checkarg_*,
spatiotemporal_to_spatial_samples) runs as the oracle for batch construction.

Only usable where /root/reference exists (the builder container).  It is used by
``oracle/make_golden.py`` to generate the committed fixtures under ``tests/golden/``; nothing
run on the GPU box imports this module.
"""
import sys
import types
from unittest.mock import MagicMock

REF_ROOT = '/root/reference'

_STUBS = [
    'tensorflow', 'tensorflow.keras', 'tensorflow.keras.layers', 'tensorflow.keras.models',
    'tensorflow.keras.callbacks', 'tensorflow.keras.backend', 'tensorflow.keras.optimizers',
    'tensorflow.keras.optimizers.schedules', 'tensorflow.keras.utils', 'keras', 'xarray',
    'ecubevis', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.axes', 'matplotlib.figure',
    'seaborn', 'horovod', 'horovod.tensorflow', 'horovod.tensorflow.keras',
]


def load_reference():
    """Return the reference ``dl4ds`` module (v1.8.0) imported under stubs."""
    if 'dl4ds' in sys.modules and getattr(sys.modules['dl4ds'], '__version__', None):
        return sys.modules['dl4ds']
    for name in _STUBS:
        if name.startswith('horovod'):
            continue   # leave horovod missing -> reference takes its has_horovod=False path
        m = MagicMock()
        m.__path__ = []
        m.__name__ = name
        sys.modules[name] = m

    class _Base:
        def __init__(self, *a, **k):
            pass

    tf = sys.modules['tensorflow']
    tf.keras = sys.modules['tensorflow.keras']
    tf.keras.layers = sys.modules['tensorflow.keras.layers']
    tf.keras.utils = sys.modules['tensorflow.keras.utils']
    tf.keras.layers.Layer = type('Layer', (_Base,), {})
    tf.keras.utils.Sequence = type('Sequence', (_Base,), {})
    for n in ('Dropout', 'GaussianDropout', 'SpatialDropout2D', 'SpatialDropout3D'):
        setattr(sys.modules['tensorflow.keras.layers'], n, type(n, (_Base,), {}))
    sys.modules['tensorflow.keras.callbacks'].History = type('History', (_Base,), {})
    sys.modules['xarray'].DataArray = type('DataArray', (), {})
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import dl4ds  # noqa: E402
    assert dl4ds.__version__ == '1.8.0'
    return dl4ds
