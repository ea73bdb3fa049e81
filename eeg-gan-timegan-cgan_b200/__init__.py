"""timegan_b200: B200-native (sm_100a) TimeGAN training step behind the reference's module/function API.

Drop-in for timeGAN/timegan_model.py, train_timegan.py, main.py and generate_long_synth.py of
Jeniya1378/eeg-gan-timegan-cgan.  Importing the package loads libtimegan_b200.so (C ABI in
include/timegan_b200.h) and fails loudly if it has not been built; there is no CPU fallback.
Import as `timegan_b200` (see ../timegan_b200/__init__.py).
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is missing)
from .timegan_model import (TimeGAN, Embedder, Recovery, Generator, Supervisor, Discriminator, GRUStack,  # noqa: F401
                            FusedGRU, init_weights_)
from .optim import FusedAdam  # noqa: F401

__version__ = "0.1.0"
