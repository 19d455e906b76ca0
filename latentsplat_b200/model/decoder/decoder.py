"""Re-export (reference layout: src/model/decoder/decoder.py); the classes live in ..interfaces."""
from ..interfaces import Decoder, DecoderOutput, DepthRenderingMode  # noqa: F401
