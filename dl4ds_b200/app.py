#!/usr/bin/env python
"""absl.FLAGS command-line app -- dl4ds/app.py:1-304, same flags, same phases (train / test / metrics):

    python -m dl4ds_b200.app --flagfile=params.cfg
    torchrun --nproc-per-node 8 -m dl4ds_b200.app --flagfile=params.cfg      (one process per GPU, NCCL)

``--data_module`` is a Python file that defines ``data_train / data_val / data_test`` (+ ``*_lr``),
``predictors_train / _val / _test``, ``static_vars``, ``inference_data``, ``inference_predictors``,
``inference_scaler``, ``gt_holdout_dataset`` and ``gt_mask`` (app.py:113-118,173-298).  Differences from the
reference: Horovod's rank comes from the torch.distributed environment (training/base.py); the downscaled product is
written as netCDF only when xarray is installed, else as ``y_hat.npy``; activations outside the CUDA path
(elu, crelu, leaky_relu, selu) are rejected by the model builders; two extra flags, ``--math`` and
``--data_on_device``, expose this package's additions.
"""
import importlib.util
import os

import numpy as np
from absl import app, flags

import dl4ds_b200 as dds
from dl4ds_b200.nets import BACKBONES
from dl4ds_b200.utils import DROPOUT_VARIANTS, INTERPOLATION_METHODS, LOSS_FUNCTIONS, UPSAMPLING_METHODS

FLAGS = flags.FLAGS

# ---- experiment (app.py:38-47)
flags.DEFINE_bool('train', True, 'Training a model')
flags.DEFINE_bool('test', True, 'Testing the trained model on holdout data')
flags.DEFINE_bool('metrics', True, 'Running verification metrics on the downscaled arrays')
flags.DEFINE_bool('debug', False, 'If True a debug training run (2 epochs with 6 steps) is executed')
flags.DEFINE_enum('trainer', 'SupervisedTrainer', ['SupervisedTrainer', 'CGANTrainer'], 'Trainer')
flags.DEFINE_enum('paired_samples', 'implicit', ['implicit', 'explicit'],
                  'Type of learning: implicit (PerfectProg) or explicit (MOS)')
flags.DEFINE_string('data_module', None, 'Python module where the data pre-processing is done')
# ---- model (app.py:49-66)
flags.DEFINE_enum('backbone', 'resnet', list(BACKBONES), 'Backbone section')
flags.DEFINE_enum('upsampling', 'spc', list(UPSAMPLING_METHODS), 'Upsampling method')
flags.DEFINE_integer('time_window', None, 'Time window for training spatio-temporal models')
flags.DEFINE_integer('n_filters', 8, 'Number of convolutional filters for the first convolutional block')
flags.DEFINE_integer('n_blocks', 6, 'Number of convolutional blocks')
flags.DEFINE_integer('n_disc_filters', 32, 'Number of convolutional filters per block in the discriminator')
flags.DEFINE_integer('n_disc_blocks', 4, 'Number of residual blocks for discriminator network')
flags.DEFINE_enum('normalization', None, ['bn', 'ln'], 'Normalization')
flags.DEFINE_float('dropout_rate', 0.2, 'Dropout rate')
flags.DEFINE_enum('dropout_variant', 'vanilla', list(DROPOUT_VARIANTS), 'Dropout variants')
flags.DEFINE_bool('attention', False, 'Attention block in convolutional layers')
_ACTS = ['elu', 'relu', 'gelu', 'crelu', 'leaky_relu', 'selu', 'sigmoid', 'tanh']
flags.DEFINE_enum('activation', 'relu', _ACTS, 'Activation used in intermediate convolutional blocks')
flags.DEFINE_enum('output_activation', None, _ACTS, 'Activation used in the last convolutional block')
flags.DEFINE_bool('localcon_layer', False, 'Locally connected convolutional layer')
flags.DEFINE_enum('decoder_upsampling', 'rc', list(UPSAMPLING_METHODS), 'Upsampling in decoder blocks (unet backbone)')
flags.DEFINE_enum('rc_interpolation', 'bilinear', list(INTERPOLATION_METHODS) + ['area', 'lanczos3', 'lanczos5', 'gaussian', 'mitchellcubic'],
                  'Interpolation used in resize convolution upsampling')
# ---- training procedure (app.py:68-90)
flags.DEFINE_enum('device', 'GPU', ['GPU', 'CPU'], 'Device to be used: GPU (CPU is rejected: there is no CPU path)')
flags.DEFINE_bool('save', True, 'Saving to disk the trained model (last epoch), metrics, run info, etc')
flags.DEFINE_string('save_path', './dl4ds_results/', 'Path for saving results to disk')
flags.DEFINE_integer('scale', 2, 'Scaling factor, positive integer')
flags.DEFINE_integer('epochs', 100, 'Number of training epochs')
flags.DEFINE_enum('loss', 'mae', list(LOSS_FUNCTIONS), 'Loss function')
flags.DEFINE_enum('interpolation', 'inter_area', list(INTERPOLATION_METHODS), 'Interpolation method')
flags.DEFINE_integer('patch_size', None, 'Patch size in number of px/gridpoints')
flags.DEFINE_integer('batch_size', 32, 'Batch size (of samples) used during training')
flags.DEFINE_multi_float('learning_rate', 1e-3, 'Learning rate')
flags.DEFINE_bool('gpu_memory_growth', True, 'Accepted for compatibility (no effect)')
flags.DEFINE_bool('use_multiprocessing', True, 'Accepted for compatibility (no effect)')
flags.DEFINE_float('lr_decay_after', 1e5, 'Steps to tweak the learning rate using the PiecewiseConstantDecay scheduler')
flags.DEFINE_bool('early_stopping', False, 'Early stopping')
flags.DEFINE_integer('patience', 6, 'Patience in number of epochs w/o improvement for early stopping')
flags.DEFINE_float('min_delta', 0.0, 'Minimum delta improvement for early stopping')
flags.DEFINE_bool('show_plot', False, 'Show the learning curve plot on finish')
flags.DEFINE_bool('save_bestmodel', True, 'SupervisedTrainer - save the epoch with the best val_loss')
flags.DEFINE_bool('verbose', True, 'Verbosity')
flags.DEFINE_integer('checkpoints_frequency', 2, 'CGANTrainer - frequency for saving checkpoints')
# ---- inference / test (app.py:92-94)
flags.DEFINE_bool('inference_array_in_hr', False, 'Whether the inference array is in high resolution')
flags.DEFINE_string('inference_save_fname', None, 'Filename for saving the inference array')
# ---- additions
flags.DEFINE_enum('math', 'tf32x3', ['fp32', 'tf32x3', 'tf32'], 'Convolution arithmetic of the CUDA path')
flags.DEFINE_bool('data_on_device', True, 'Keep the training array in HBM and build batches on the GPU when possible')


def _first_worker():
    return int(os.environ.get('RANK', '0')) == 0


def architecture_params():
    """app.py:120-170: which builder arguments each (time_window, upsampling) combination receives."""
    common = dict(n_filters=FLAGS.n_filters, normalization=FLAGS.normalization, dropout_rate=FLAGS.dropout_rate,
                  dropout_variant=FLAGS.dropout_variant, attention=FLAGS.attention, activation=FLAGS.activation,
                  localcon_layer=FLAGS.localcon_layer, output_activation=FLAGS.output_activation)
    if FLAGS.time_window is None:
        p = dict(common, n_blocks=FLAGS.n_blocks)
        if FLAGS.upsampling == 'pin':
            if FLAGS.backbone == 'unet':
                p['decoder_upsampling'] = FLAGS.decoder_upsampling
                p['rc_interpolation'] = FLAGS.rc_interpolation
        else:
            p['rc_interpolation'] = FLAGS.rc_interpolation
    elif FLAGS.upsampling == 'pin':
        p = dict(common, n_blocks=FLAGS.n_blocks)
    else:
        p = dict(common, rc_interpolation=FLAGS.rc_interpolation)        # n_blocks not passed (app.py:160-170)
    return p


def dl4ds(argv):
    """DL4DS absl.FLAGS-based command line app (app.py:98-301)."""
    first = _first_worker()
    if first:
        print('<<<<<<<<<<<<<<<<<<<<<<<<<<<<<<<<<<<<< DL4DS >>>>>>>>>>>>>>>>>>>>>>>>>>>>>>>>>>>>\n')
    if FLAGS.debug:
        epochs = 2
        steps_per_epoch = test_steps = validation_steps = 6
    else:
        epochs = FLAGS.epochs
        steps_per_epoch = test_steps = validation_steps = None
    if first:
        print('<<<<<<<<<<<<<<<<<<<<<<<<<<<<<<<<< Loading data >>>>>>>>>>>>>>>>>>>>>>>>>>>>>>>>>\n')
    if FLAGS.data_module is None:
        raise ValueError('`data_module` flag must be provided (path to the data preprocessing module)')
    spec = importlib.util.spec_from_file_location('module.name', FLAGS.data_module)
    DATA = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(DATA)
    get = lambda name: getattr(DATA, name, None)
    explicit = FLAGS.paired_samples == 'explicit'
    arch = architecture_params()
    lr = FLAGS.learning_rate if len(FLAGS.learning_rate) > 1 else FLAGS.learning_rate[0]
    trainer, y_hat = None, None

    if FLAGS.train:
        if first:
            print('\n<<<<<<<<<<<<<<<<<<<<<<<<<<<<< DL4DS Training phase >>>>>>>>>>>>>>>>>>>>>>>>>>>>>\n')
        if FLAGS.trainer == 'SupervisedTrainer':
            trainer = dds.SupervisedTrainer(
                backbone=FLAGS.backbone, upsampling=FLAGS.upsampling, data_train=DATA.data_train,
                data_val=DATA.data_val, data_test=DATA.data_test,
                data_train_lr=get('data_train_lr') if explicit else None,
                data_val_lr=get('data_val_lr') if explicit else None,
                data_test_lr=get('data_test_lr') if explicit else None,
                predictors_train=get('predictors_train'), predictors_val=get('predictors_val'),
                predictors_test=get('predictors_test'), static_vars=get('static_vars'), scale=FLAGS.scale,
                interpolation=FLAGS.interpolation, patch_size=FLAGS.patch_size, time_window=FLAGS.time_window,
                batch_size=FLAGS.batch_size, loss=FLAGS.loss, epochs=epochs, steps_per_epoch=steps_per_epoch,
                validation_steps=validation_steps, test_steps=test_steps, device=FLAGS.device,
                gpu_memory_growth=FLAGS.gpu_memory_growth, use_multiprocessing=FLAGS.use_multiprocessing,
                learning_rate=lr, lr_decay_after=FLAGS.lr_decay_after, early_stopping=FLAGS.early_stopping,
                patience=FLAGS.patience, min_delta=FLAGS.min_delta, show_plot=FLAGS.show_plot, save=FLAGS.save,
                save_path=FLAGS.save_path, save_bestmodel=FLAGS.save_bestmodel, trained_model=None,
                trained_epochs=0, verbose=FLAGS.verbose, math=FLAGS.math, data_on_device=FLAGS.data_on_device,
                **arch)
        else:
            disc = dict(n_filters=FLAGS.n_disc_filters, n_res_blocks=FLAGS.n_disc_blocks,
                        normalization=FLAGS.normalization, activation=FLAGS.activation, attention=FLAGS.attention)
            trainer = dds.CGANTrainer(
                backbone=FLAGS.backbone, upsampling=FLAGS.upsampling, data_train=DATA.data_train,
                data_test=DATA.data_test, data_train_lr=get('data_train_lr') if explicit else None,
                data_test_lr=get('data_test_lr') if explicit else None, predictors_train=get('predictors_train'),
                predictors_test=get('predictors_test'), scale=FLAGS.scale, patch_size=FLAGS.patch_size,
                time_window=FLAGS.time_window, loss=FLAGS.loss, epochs=epochs, batch_size=FLAGS.batch_size,
                learning_rates=FLAGS.learning_rate, device=FLAGS.device, gpu_memory_growth=FLAGS.gpu_memory_growth,
                steps_per_epoch=steps_per_epoch, interpolation=FLAGS.interpolation, static_vars=get('static_vars'),
                checkpoints_frequency=FLAGS.checkpoints_frequency, save=FLAGS.save, save_path=FLAGS.save_path,
                save_logs=False, save_loss_history=FLAGS.save, verbose=FLAGS.verbose, generator_params=arch,
                discriminator_params=disc, math=FLAGS.math)
        trainer.run()

    if FLAGS.test:
        if first:
            print('\n<<<<<<<<<<<<<<<<<<<<<<<<<<<<<<<<< DL4DS Test phase >>>>>>>>>>>>>>>>>>>>>>>>>>>>>\n')
        if trainer is None:
            raise ValueError('--test needs a trained model: run with --train as the reference does (app.py:262)')
        if first:
            predictor = dds.Predictor(
                trainer=trainer, array=DATA.inference_data, array_in_hr=FLAGS.inference_array_in_hr,
                scale=FLAGS.scale, interpolation=FLAGS.interpolation, predictors=get('inference_predictors'),
                static_vars=get('static_vars'), time_window=FLAGS.time_window, batch_size=FLAGS.batch_size,
                scaler=get('inference_scaler'), save_path=FLAGS.save_path, save_fname=FLAGS.inference_save_fname,
                device=FLAGS.device)
            y_hat = predictor.run()
            if FLAGS.save_path is not None:
                os.makedirs(FLAGS.save_path, exist_ok=True)
                gt = get('gt_holdout_dataset')
                try:
                    import xarray as xr
                    xr.DataArray(data=np.squeeze(y_hat), dims=('time', 'lat', 'lon'),
                                 coords={'time': gt.time, 'lon': gt.lon, 'lat': gt.lat}
                                 ).to_netcdf('%sy_hat.nc' % FLAGS.save_path)
                except (ImportError, AttributeError):
                    np.save(os.path.join(FLAGS.save_path, 'y_hat.npy'), y_hat)

    if FLAGS.metrics:
        if first:
            print('\n<<<<<<<<<<<<<<<<<<<<<<<<< DL4DS Metrics computation phase >>>>>>>>>>>>>>>>>>>>>>\n')
        if y_hat is None and first:
            raise ValueError('--metrics needs the downscaled array: run with --test')
        if first:
            if FLAGS.save_path is not None:
                os.makedirs(FLAGS.save_path, exist_ok=True)
            dds.compute_metrics(y_test=DATA.gt_holdout_dataset, y_test_hat=y_hat, dpi=300, plot_size_px=1200,
                                mask=get('gt_mask'), save_path=FLAGS.save_path, n_jobs=-1, verbose=FLAGS.verbose)


def main():
    app.run(dl4ds)


if __name__ == '__main__':
    main()
