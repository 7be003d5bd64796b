"""Trainers with the reference's API surface (dl4ds/training/__init__.py): ``Trainer``,
``SupervisedTrainer``, ``CGANTrainer`` (+ the cGAN step functions)."""
from .base import Trainer
from .cgan import CGANTrainer, discriminator_loss, generator_loss, load_checkpoint, train_step
from .supervised import SupervisedTrainer

__all__ = ['Trainer', 'SupervisedTrainer', 'CGANTrainer', 'train_step', 'generator_loss', 'discriminator_loss',
           'load_checkpoint']
