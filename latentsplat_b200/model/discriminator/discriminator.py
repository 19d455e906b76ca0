"""Re-export (reference layout: src/model/discriminator/discriminator.py); the class lives in ..interfaces."""
from ..interfaces import Discriminator  # noqa: F401
