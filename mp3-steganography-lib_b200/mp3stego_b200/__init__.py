"""mp3stego_b200 -- B200 (sm_100a) implementation of mp3stego's per-granule codec hot path behind the
reference's own Python surface (mp3stego/__init__.py:1-4 exports Decoder, Encoder, Steganography).
There is no CPU fallback: every codec call goes through libmp3stego_b200.so and fails loudly without it."""
from mp3stego_b200 import _lib  # noqa: F401
from mp3stego_b200.decoder import Decoder, MP3Parser  # noqa: F401
from mp3stego_b200.encoder import Encoder, MP3Encoder  # noqa: F401
from mp3stego_b200.steganography import Steganography  # noqa: F401
from mp3stego_b200.wavio import WavReader  # noqa: F401

__all__ = ["Decoder", "Encoder", "Steganography", "MP3Parser", "MP3Encoder", "WavReader"]
