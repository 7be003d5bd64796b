"""Base trainer: argument validation, data-parallel context, global batch, first-worker flag
(dl4ds/training/base.py:24-187).  Horovod's role (init, local-rank GPU pinning, rank/size) is
played by ``torch.distributed`` over NCCL: one process per GPU launched by torchrun; the trainer
joins the process group if the launcher's environment (RANK / WORLD_SIZE / LOCAL_RANK) is present.
"""
import os
from abc import ABC, abstractmethod

import numpy as np

from ..utils import check_compatibility_upsbackb, checkarg_loss


def _is_array(a):
    return isinstance(a, np.ndarray) or (hasattr(a, 'values') and hasattr(a, 'dims'))   # xr.DataArray


class DataParallel:
    """What ``hvd.init() / hvd.rank() / hvd.size() / hvd.local_rank()`` provide (base.py:97-107)."""

    def __init__(self, device='GPU'):
        import torch
        self.rank, self.size, self.local_rank = 0, 1, 0
        self.dist = None
        if 'RANK' in os.environ and int(os.environ.get('WORLD_SIZE', '1')) > 1:
            import torch.distributed as dist
            self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
            if not dist.is_initialized():
                os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
                if device == 'GPU' and torch.cuda.is_available():
                    torch.cuda.set_device(self.local_rank)
                    dist.init_process_group('nccl', device_id=torch.device('cuda', self.local_rank))
                else:
                    dist.init_process_group('gloo')
            self.dist = dist
            self.rank, self.size = dist.get_rank(), dist.get_world_size()
        if device == 'GPU':
            if not torch.cuda.is_available():
                raise RuntimeError("device='GPU' needs a CUDA device: the B200 path has no CPU fallback")
            torch.cuda.set_device(self.local_rank)
            self.torch_device = torch.device('cuda', self.local_rank)
        else:
            raise ValueError("device not recognized" if device != 'CPU' else
                             "device='CPU' is not available: dl4ds_b200 runs the hot path on CUDA only")

    def allreduce_mean_scalar(self, value):
        """Mean of a host scalar over the ranks (hvd.callbacks.MetricAverageCallback's role for val_loss)."""
        if self.dist is None or self.size == 1:
            return float(value)
        import torch
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.torch_device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item()) / self.size


class Trainer(ABC):
    """Trainer -- training/base.py:24-152 (same arguments, checks and exception types)."""

    def __init__(self, backbone, upsampling, data_train, data_train_lr=None, time_window=None,
                 loss='mae', batch_size=64, patch_size=None, scale=4, device='GPU',
                 gpu_memory_growth=True, use_multiprocessing=False, verbose=True, model_list=None,
                 save=True, save_path=None, show_plot=False):
        self.data_train = data_train
        if not _is_array(self.data_train):
            raise TypeError('`data_train` object must be of np.ndarray or xr.DataArray type')
        if not self.data_train.ndim > 3:
            raise ValueError('`data_train` must be at least 4D [samples, lat, lon, variables]')
        self.data_train_lr = data_train_lr
        if self.data_train_lr is not None:
            if not _is_array(self.data_train_lr):
                raise TypeError('`data_train_lr` must be a np.ndarray or xr.DataArray object')
            if self.data_train_lr.shape[0] != self.data_train.shape[0]:
                raise ValueError('`data_train_lr` and `data_train` must contain the same number of '
                                 'samples (equal 1st dim lenght)')
            if not self.data_train_lr.ndim > 3:
                raise ValueError('`data_train_lr` must be at least 4D [samples, lat, lon, variables]')
        self.backbone, self.upsampling = check_compatibility_upsbackb(backbone, upsampling, time_window)
        self.time_window = time_window
        self.model_is_spatiotemporal = self.time_window is not None and self.time_window > 1
        self.batch_size = batch_size
        self.patch_size = patch_size
        self.loss = loss
        self.scale = scale
        self.device = device
        self.gpu_memory_growth = gpu_memory_growth
        self.use_multiprocessing = use_multiprocessing
        self.verbose = verbose
        self.model_list = model_list
        self.save = save
        self.save_path = save_path
        if self.save_path is None:
            self.save_path = './'
        elif not self.save_path.endswith('/'):
            self.save_path += '/'
        self.savecheckpoint_path = self.save_path
        self.show_plot = show_plot

        # one process per GPU (Horovod pins one visible GPU per process => n_devices == 1)
        self.dp = DataParallel(self.device)
        n_devices = 1
        self.global_batch_size = self.batch_size * n_devices
        if self.verbose in [1, 2]:
            print('Number of devices: {} (data-parallel world size {})'.format(n_devices, self.dp.size))
            print('Global batch size: {}'.format(self.global_batch_size * self.dp.size))
        self.running_on_first_worker = self.dp.rank == 0

        imsize = self.patch_size if self.patch_size is not None else self.data_train.shape[-2]
        if self.scale is not None:
            if imsize % self.scale != 0:
                raise ValueError('The image size must be divisible by `scale` (remainder must be zero). '
                                 'Crop the images or set `patch_size` accordingly')
            if self.data_train_lr is not None:
                scale_from_data = self.data_train.shape[1] / self.data_train_lr.shape[1]
                if not int(scale_from_data) == int(self.scale):
                    raise ValueError('Wrong `scale` value, check `data_train` and `data_train_lr` grid sizes')
        self.lossf = checkarg_loss(self.loss)

    @abstractmethod
    def run(self):
        pass

    @abstractmethod
    def setup_model(self):
        pass

    def save_results(self, model_to_save=None, folder_prefix=None):
        """base.py:162-187: model weights, running time and test score (first worker only).  The
        model is stored as ``<save_path>/<backbone>_<upsampling>/weights.npz`` (named Keras-layout
        arrays) instead of a TF SavedModel; the learning-curve plot is out of scope."""
        if not self.save:
            return
        if model_to_save is None:
            model_to_save = self.model
        prefix = folder_prefix or ''
        self.model_save_path = self.save_path + prefix + self.backbone + '_' + self.upsampling + '/'
        if self.running_on_first_worker:
            os.makedirs(self.model_save_path, exist_ok=True)
            model_to_save.save(self.model_save_path + 'weights.npz')
            np.savetxt(self.save_path + 'running_time.txt', [self.timing.running_time], fmt='%s')
            np.savetxt(self.save_path + 'test_loss.txt', [self.test_loss], fmt='%0.6f')
