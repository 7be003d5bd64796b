"""Trainers with the reference's API surface (dl4ds/training/__init__.py): ``Trainer``,
``SupervisedTrainer``, ``CGANTrainer``."""
from .base import Trainer
from .supervised import SupervisedTrainer

__all__ = ['Trainer', 'SupervisedTrainer']
