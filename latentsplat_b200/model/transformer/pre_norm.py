"""Re-export (reference layout: src/model/transformer/pre_norm.py); the class lives in .transformer."""
from .transformer import PreNorm  # noqa: F401
