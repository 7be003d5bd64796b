"""dl4ds_b200 -- B200-native implementation of the DL4DS convolutional super-resolution hot path
(hand-written sm_100a CUDA behind a C ABI).  The names exported here mirror
``dl4ds/__init__.py:7-45`` of the reference for the part of the package that is on the hot path."""
__version__ = '0.1.0'

from .utils import (BACKBONE_BLOCKS, DROPOUT_VARIANTS, INTERPOLATION_METHODS, LOSS_FUNCTIONS,  # noqa: F401
                    POSTUPSAMPLING_METHODS, UPSAMPLING_METHODS)
from .dataloader import DataGenerator, create_batch_hr_lr, create_pair_hr_lr  # noqa: F401
from .nets import (net_pin, net_postupsampling, recnet_pin, recnet_postupsampling,  # noqa: F401
                   residual_discriminator, unet_pin)
from .training import CGANTrainer, SupervisedTrainer, Trainer  # noqa: F401
from .inference import Predictor, predict  # noqa: F401
from . import losses  # noqa: F401
from .metrics import compute_correlation, compute_metrics, compute_rmse  # noqa: F401
