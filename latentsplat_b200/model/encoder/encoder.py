"""Re-export (reference layout: src/model/encoder/encoder.py); the class lives in ..interfaces."""
from ..interfaces import Encoder  # noqa: F401
