"""mp3stego_b200 -- B200 (sm_100a) implementation of mp3stego's per-granule codec hot path behind the
reference's own Python surface (mp3stego/__init__.py:1-4 exports Decoder, Encoder, Steganography)."""
from mp3stego_b200 import _lib  # noqa: F401

__all__ = ["_lib"]
