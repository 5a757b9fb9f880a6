"""chatterbox_b200: B200-native (sm_100a) inference engine for Chatterbox's two-stage generate path.

Drop-in surface (reference file:line in each docstring):
    ChatterboxTTS.from_pretrained / from_local / generate     (reference src/chatterbox/tts.py:106-272)
    T3.inference / T3.inference_turbo                         (reference src/chatterbox/models/t3/t3.py:225-390, 392-468)
    ChatterboxTurboTTS.from_local / generate                  (reference src/chatterbox/tts_turbo.py:104-321)
    S3Gen.inference / flow_inference / hift_inference         (reference src/chatterbox/models/s3gen/s3gen.py:300-362)
All arithmetic runs in libcbx.so (hand-written CUDA behind the C ABI in include/cbx.h); importing this package
does not need a GPU, constructing an Engine does.
"""
from ._lib import CbxError, LIB_PATH  # noqa: F401
from .engine import Engine, PackedLayout  # noqa: F401
from .t3 import T3, T3Cond  # noqa: F401
from .s3gen import S3Gen  # noqa: F401
from .tts import ChatterboxTTS, ChatterboxMultilingualTTS, Conditionals, punc_norm  # noqa: F401
from .tts_turbo import ChatterboxTurboTTS  # noqa: F401

__all__ = ["Engine", "T3", "T3Cond", "S3Gen", "ChatterboxTTS", "ChatterboxMultilingualTTS", "ChatterboxTurboTTS", "Conditionals", "punc_norm", "CbxError"]
