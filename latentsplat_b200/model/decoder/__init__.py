"""Decoder factory (/root/reference/src/model/decoder/__init__.py:12-17)."""
from .decoder import Decoder, DecoderOutput
from .decoder_splatting_cuda import DecoderSplattingCUDA, DecoderSplattingCUDACfg

DECODERS = {"splatting_cuda": DecoderSplattingCUDA}
DecoderCfg = DecoderSplattingCUDACfg


def get_decoder(decoder_cfg: DecoderCfg, background_color, variational: bool) -> Decoder:
    return DECODERS[decoder_cfg.name](decoder_cfg, background_color, variational)
