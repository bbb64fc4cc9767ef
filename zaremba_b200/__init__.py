"""zaremba_b200: B200-native (sm_100a) implementation of the LSTM-LM hot path of
ahmetumutdurmus/zaremba behind the reference's `model.Model` interface.

    from zaremba_b200 import Model            # drop-in for the reference's model.py
    from zaremba_b200 import Trainer          # fused train / eval step (main.py:109-117, :86-95)
"""
from .model import Model, Embed, LSTM, Linear  # noqa: F401
from .trainer import Trainer, minibatch  # noqa: F401
from . import ensemble, parallel  # noqa: F401

__all__ = ["Model", "Embed", "LSTM", "Linear", "Trainer", "minibatch"]
