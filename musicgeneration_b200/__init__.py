"""B200-native drop-in for the MusicTransformer hot path of SJTMusicTeam/MusicGeneration.

Same module surface as mg/model/MusicTransformer (layers / network / criterion / metrics /
utils / config); all arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/mt_b200.h`` (``libmt_b200.so``, bound with ctypes in ``_lib.py``).  CUDA only: there
is no CPU fallback, and a missing extension raises at first use.
"""
from . import (_lib, config, criterion, data, engine, layers, metrics, network, ops, optim,  # noqa: F401
               parallel, sequence, utils)
from .criterion import CustomSchedule, SmoothCrossEntropyLoss  # noqa: F401
from .engine import Mask  # noqa: F401
from .layers import (DynamicPositionEmbedding, Encoder, EncoderLayer,  # noqa: F401
                     RelativeGlobalAttention)
from .network import MusicTransformer  # noqa: F401

__version__ = "0.1.0"
