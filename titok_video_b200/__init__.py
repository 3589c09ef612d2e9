"""titok_video_b200 -- B200-native (sm_100a) implementation of the TiTok-Video tokenizer hot path.

Drop-in for the reference's `model.titok.TiTok`, `model.base.blocks.TiTokEncoder/TiTokDecoder`,
`model.quantizer.fsq.FSQ` and `train_utils.codebook_logging.CodebookLogger`; all device work runs in the
hand-written CUDA kernels of libtitok_b200.so (see include/titok_b200.h). No CPU fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA extension is missing)
from .model.titok import TiTok
from .model.base.blocks import TiTokEncoder, TiTokDecoder
from .model.quantizer.fsq import FSQ
from .model.quantizer.vq import VectorQuantizer
from .train_utils.codebook_logging import CodebookLogger
from .config import load_config, AttrDict

__all__ = ["TiTok", "TiTokEncoder", "TiTokDecoder", "FSQ", "VectorQuantizer", "CodebookLogger", "load_config", "AttrDict"]
__version__ = "0.1.0"
