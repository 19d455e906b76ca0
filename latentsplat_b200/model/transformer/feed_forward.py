"""Re-export (reference layout: src/model/transformer/feed_forward.py); the class lives in .transformer."""
from .transformer import FeedForward  # noqa: F401
