"""dl4ds_b200 -- B200-native implementation of the DL4DS convolutional super-resolution hot path
(hand-written sm_100a CUDA behind a C ABI)."""
__version__ = '0.1.0'
