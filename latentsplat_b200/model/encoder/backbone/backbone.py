"""Re-export (reference layout: src/model/encoder/backbone/backbone.py); the class lives in ...interfaces."""
from ...interfaces import Backbone  # noqa: F401
