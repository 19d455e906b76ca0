"""Re-export (reference layout: src/model/autoencoder/autoencoder.py); the class lives in ..interfaces."""
from ..interfaces import Autoencoder  # noqa: F401
