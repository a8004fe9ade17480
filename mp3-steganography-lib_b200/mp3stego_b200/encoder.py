"""Encoder-side mirror of the reference's Python surface (encoder/encoder.py, encoder/MP3_Encoder.py): same class
names, arguments, return values and failure behaviour, with MP3Encoder.encode replaced by one call into the CUDA
library (include/mp3stego_b200.h)."""
import os
import sys

import numpy as np

from mp3stego_b200 import _lib
from mp3stego_b200.wavio import WavReader


class MP3Encoder:
    """MP3Encoder(wav_file: WavReader, hide_str='') with encode / write_mp3_file / print_info / hide_str_offset."""

    def __init__(self, wav_file: WavReader, hide_str: str = "", device: int = 0):
        self.__wav_file = wav_file
        self.__hide_str = hide_str
        self.__hide_str_offset = 0
        self.__out_buffer = b""
        self.__device = device

    def print_info(self):
        w = self.__wav_file
        mode = "stereo" if w.num_of_channels > 1 else "mono"
        print(f"MPEG-I layer III, {mode} Psychoacoustic Model: Shine")
        print(f"Bitrate: {w.bitrate} kbps ", end="")
        print(f"De-emphasis: none\t{'Original' if w.original else ''}\t{'(C)' if w.copyright else ''}")
        print(f"Encoding \"{w.file_path}\" to \"{w.file_path[:-3]}mp3\"\n")

    def encode(self):
        w = self.__wav_file
        n = w.num_of_samples
        if w.num_of_channels != 2:
            # WAV_Reader.py:109,163-164 steps every channel cursor by 2: mono input runs off the buffer (SURVEY A.E1)
            raise IndexError("index out of bounds: the reference encoder only works on 16-bit stereo input")
        if n % 1152 or len(w.buffer) < 2 * n:
            # the extra pass for a partial frame reads past the sample buffer (MP3_Encoder.py:611-614, :756-757; SURVEY A.E2)
            raise IndexError("index out of bounds: sample count is not a multiple of 1152")
        h = _lib.default_handle(self.__device)
        pcm = np.ascontiguousarray(w.buffer[: 2 * n], dtype=np.int16)
        res = h.encode(pcm, [n], w.samplerate, w.bitrate, payloads=[self.__hide_str] if self.__hide_str else None)
        self.__out_buffer = bytes(res["mp3"][: int(res["out_len"][0])])
        self.__hide_str_offset = int(res["hide_str_offset"][0])

    def write_mp3_file(self, output_file: str):
        with open(output_file, "wb") as f:
            f.write(self.__out_buffer)

    @property
    def hide_str_offset(self):
        return self.__hide_str_offset


class Encoder:
    """Encoder(file_path, output_file_path, bitrate=320, hide_str='').encode(quiet) -> too_long."""

    def __init__(self, file_path: str, output_file_path: str, bitrate: int = 320, hide_str: str = "", device: int = 0):
        self.__file_path = file_path
        self.__output_file_path = output_file_path
        if not os.path.exists(self.__file_path):
            sys.exit(f"File {self.__file_path} not found.")
        self.__wav_file = WavReader(self.__file_path, bitrate)
        self.__hide_str = hide_str
        self.__encoder = MP3Encoder(self.__wav_file, hide_str=hide_str, device=device)

    def encode(self, quiet: bool = True) -> bool:
        if not quiet:
            self.__encoder.print_info()
        self.__encoder.encode()
        self.__encoder.write_mp3_file(self.__output_file_path)
        too_long = self.__encoder.hide_str_offset < len(self.__hide_str) - 1   # encoder.py:49-51
        if not quiet:
            if too_long:
                print("File too short for this message length, your message has been trimmed.")
            print(f"MP3 file created on {self.__output_file_path}")
        return too_long
