"""ctypes binding of ``libdl4ds_b200.so`` (C ABI declared in ``include/dl4ds_b200.h``).

The library is built in-tree by ``dl4ds_b200/csrc/Makefile`` (``__graft_entry__.build()``).  There is
no CPU fallback: if the shared object is missing, or a call returns a negative status, a
``RuntimeError`` carrying ``dl4ds_last_error()`` is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libdl4ds_b200.so')

ACT = {None: 0, 'linear': 0, 'relu': 1, 'sigmoid': 2, 'tanh': 3}
MATH_FP32, MATH_TF32X3, MATH_TF32 = 0, 1, 2
MATH_F16X3 = 3
MATH = {'fp32': MATH_FP32, 'tf32x3': MATH_TF32X3, 'tf32': MATH_TF32, 'f16x3': MATH_F16X3}
W_HWIO, W_FLIP_T, W_PREPACKED = 0, 1, 4

_P, _I, _L, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
_CODES = {'p': _P, 'i': _I, 'l': _L, 'f': _F}

# name -> (restype code, argument codes); order follows include/dl4ds_b200.h
SIGNATURES = {
    'dl4ds_last_error': ('s', ''),
    'dl4ds_version': ('i', ''),
    'dl4ds_device_is_sm100': ('i', ''),
    'dl4ds_tc_launch_count': ('l', ''),
    'dl4ds_debug_set_buffer': ('i', 'p'),
    'dl4ds_conv2d_fwd_workspace_bytes': ('l', 'iiiiiiiiiiiii'),
    'dl4ds_conv2d_pack': ('i', 'piiiiiipp'),
    'dl4ds_conv2d_pack_desc': ('l', 'piiiiiipp'),
    'dl4ds_conv2d_pack_multi': ('i', 'pilp'),
    'dl4ds_conv2d_fwd': ('i', 'pipppipiiiiiiiiiiiiiiiiiiipp'),
    'dl4ds_conv2d_dgrad_fused_supported': ('i', 'iiiiiiii'),
    'dl4ds_conv2d_dgrad_fused': ('i', 'pippipiipiiiiiiiiiiiipp'),
    'dl4ds_conv2d_wgrad_workspace_bytes': ('l', 'iiiiiiii'),
    'dl4ds_conv2d_wgrad': ('i', 'pipipiiiiiiiiiiiipip'),
    'dl4ds_spc_pointwise_compose': ('i', 'ppppppiiiip'),
    'dl4ds_spc_pointwise_chain': ('i', 'pppppppppiiiip'),
    'dl4ds_bias_act_bwd': ('i', 'pipipipiiiiiip'),
    'dl4ds_add': ('i', 'pipipiliip'),
    'dl4ds_copy_channels': ('i', 'pipiliip'),
    'dl4ds_channel_attention_fwd': ('i', 'pipipppppppiliiip'),
    'dl4ds_channel_attention_bwd': ('i', 'pipipippppppppppiliiip'),
    'dl4ds_pixel_loss': ('i', 'pppplifp'),
    'dl4ds_ssim_loss_workspace_floats': ('l', 'iiiii'),
    'dl4ds_ssim_loss': ('i', 'ppiiiiipfppipp'),
    'dl4ds_ssim_index': ('i', 'ppiiiifppp'),
    'dl4ds_metrics_moments': ('i', 'ppilppp'),
    'dl4ds_batchnorm_stats': ('i', 'pilippppfpp'),
    'dl4ds_norm_apply': ('i', 'pippppfpiliip'),
    'dl4ds_batchnorm_bwd': ('i', 'pipipipppfpipppliip'),
    'dl4ds_layernorm_fwd': ('i', 'pippfpiliip'),
    'dl4ds_layernorm_bwd': ('i', 'pipipipfpippliip'),
    'dl4ds_depthwise_conv_fwd': ('i', 'pipppiiiiiiiip'),
    'dl4ds_depthwise_conv_wgrad': ('i', 'pipipiiiiip'),
    'dl4ds_gelu_fwd': ('i', 'pplp'),
    'dl4ds_gelu_bwd': ('i', 'ppplp'),
    'dl4ds_channel_scale_fwd': ('i', 'pippilip'),
    'dl4ds_channel_scale_bwd': ('i', 'pipippiplip'),
    'dl4ds_dropout': ('i', 'pipilliifipip'),
    'dl4ds_rng_advance': ('i', 'pp'),
    'dl4ds_adam_step': ('i', 'pppplffffifp'),
    'dl4ds_adam_step_dev': ('i', 'pppplpffffp'),
    'dl4ds_convt_rearrange': ('i', 'ppiiiiiiiip'),
    'dl4ds_gather_crop': ('i', 'pppppiiiiiiiip'),
    'dl4ds_avgpool_coarsen': ('i', 'ppiiiiip'),
    'dl4ds_resample_taps': ('i', 'ppiiiiiippippiiip'),
    'dl4ds_resize_bilinear_fwd': ('i', 'pipiiiiiiip'),
    'dl4ds_resize_bilinear_bwd': ('i', 'pipiiiiiiip'),
    'dl4ds_resize_fwd': ('i', 'pipiiiiiiiip'),
    'dl4ds_resize_bwd': ('i', 'pipiiiiiiiip'),
    'dl4ds_maxpool2_fwd': ('i', 'pipiiiiip'),
    'dl4ds_maxpool2_bwd': ('i', 'pipipiiiiip'),
    'dl4ds_local_conv1x1_fwd': ('i', 'pipppiiiiiip'),
    'dl4ds_local_conv1x1_bwd': ('i', 'pipippippiiiiip'),
    'dl4ds_convlstm_gates_fwd': ('i', 'ppppiplip'),
    'dl4ds_convlstm_gates_bwd': ('i', 'ppppippplip'),
    'dl4ds_act_fwd': ('i', 'pipiliip'),
    'dl4ds_group_mean_fwd': ('i', 'pipilip'),
    'dl4ds_group_mean_bwd': ('i', 'ppiilip'),
    'dl4ds_mul': ('i', 'ppplp'),
    'dl4ds_bce_loss': ('i', 'pfpplfip'),
    'dl4ds_permute_frames': ('i', 'ppiilp'),
    'dl4ds_pad_bottom_right': ('i', 'pipiiiiiiip'),
    'dl4ds_axpby': ('i', 'fpfplp'),
    'dl4ds_comm_unique_id_bytes': ('i', ''),
    'dl4ds_comm_nccl_version': ('i', ''),
    'dl4ds_comm_get_unique_id': ('i', 'p'),
    'dl4ds_comm_init_rank': ('i', 'pii'),
    'dl4ds_comm_size': ('i', ''),
    'dl4ds_comm_rank': ('i', ''),
    'dl4ds_comm_allreduce_sum': ('i', 'plp'),
    'dl4ds_comm_broadcast': ('i', 'plip'),
    'dl4ds_comm_destroy': ('i', ''),
}

_lib = None


class Dl4dsError(RuntimeError):
    pass


_TRACE = os.environ.get("DL4DS_TRACE", "0") == "1"


def load():
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Dl4dsError(
            'libdl4ds_b200.so not found at %s -- build it with `python -c "import __graft_entry__ as g; '
            'g.build()"` or `make -C dl4ds_b200/csrc`; there is no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = ctypes.c_char_p if res == 's' else _CODES[res]
        fn.argtypes = [_CODES[c] for c in args]
    _lib = lib
    return lib


def last_error():
    return load().dl4ds_last_error().decode('utf-8', 'replace')


def call(name, *args):
    """Call a status-returning entry point; raise on a negative status."""
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise Dl4dsError('%s failed (%d): %s' % (name, rc, last_error()))
    if _TRACE:          # DL4DS_TRACE=1: name every launch and wait for it (locating a kernel that never returns)
        import sys
        import torch
        print('[dl4ds] %s %s' % (name, ' '.join(str(a) for a in args if isinstance(a, (int, float)) and abs(a) < 1 << 20)),
              file=sys.stderr, flush=True)
        if not torch.cuda.is_current_stream_capturing():
            torch.cuda.synchronize()
    return rc
