"""B200-native drop-in for the ISBFSAR `modules/ar` one-shot open-set scoring path."""
from .params import TRXConfig  # noqa: F401
from .model import TRXOS  # noqa: F401
from .ar import ActionRecognizer  # noqa: F401
from .decode import HeatmapDecoder  # noqa: F401

__all__ = ["TRXConfig", "TRXOS", "ActionRecognizer", "HeatmapDecoder"]
