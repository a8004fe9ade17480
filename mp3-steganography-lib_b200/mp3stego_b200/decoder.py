"""Decoder-side mirror of the reference's Python surface (decoder/decoder.py, decoder/MP3_Parser.py,
decoder/ID3_Parser.py): same class names, arguments, return values, temp-WAV / reveal-text side effects and failure messages, with
the frame loop of MP3Parser.parse_file replaced by one call into the CUDA library (include/mp3stego_b200.h).  Not mirrored: the
ID3 frame listing the reference writes to ./METADATA.txt when quiet=False (decoder.py:38-57; host-only text)."""
import os
import sys
import time

import numpy as np

from mp3stego_b200 import _lib
from mp3stego_b200.wavio import write_wav


def id3_offset(data) -> int:
    """Audio start as Decoder.__init__ computes it (decoder.py:29-33): the ID3v2 tag is honoured only when the four
    low flag bits are clear (ID3_Parser.py:129-135); offset = synchsafe size + 10 (+10 more with a footer, :121-125)."""
    if len(data) >= 10 and data[0] == 0x49 and data[1] == 0x44 and data[2] == 0x33:
        flags = int(data[5])                     # int(): `data` may be a numpy uint8 array, whose arithmetic wraps at 256
        if flags & 0x0F:
            return 0
        size = 0
        for i in range(4):
            size = (size << 7) + int(data[6 + i])   # util.char_to_int (decoder/util.py:6-19)
        return size + (20 if flags & 0x10 else 10)
    return 0


def parse_reveal(output_bits: str) -> str:
    """decoder.py:90-105: 8-bit groups -> chars, '<len>#' prefix (a non-numeric prefix means length 0), slice."""
    output_str = "".join(chr(int("".join(x), 2)) for x in zip(*[iter(output_bits)] * 8))
    message_len_str = ""
    for ch in output_str:
        if ch == "#":
            break
        message_len_str += ch
    try:
        message_len = int(message_len_str)
    except Exception:
        message_len = 0
        message_len_str = ""
    if (len(message_len_str) + 1 + message_len) > len(output_str):
        return output_str[len(message_len_str) + 1:]
    return output_str[len(message_len_str) + 1: len(message_len_str) + 1 + message_len]


class MP3Parser:
    """MP3Parser(file_data, offset, wav_file_path) with parse_file / write_to_wav / get_bitrate / output_bits."""

    def __init__(self, file_data, offset: int, wav_file_path: str, device: int = 0, exact: bool = True):
        self.__file_data = np.frombuffer(bytes(file_data), dtype=np.uint8) if not isinstance(file_data, np.ndarray) else file_data
        self.__offset = offset
        self.__wav_file_path = wav_file_path
        self.__device = device
        self.__exact = exact
        buf = self.__file_data[offset:]
        self.__valid = bool(buf[0] == 0xFF and buf[1] >= 0xE0)   # IndexError on an empty buffer, as in the reference
        self.__pcm16 = np.zeros((0, 2), np.int16)
        self.__sampling_rate = None
        self.__bit_rate = None
        self.output_bits = ""

    def parse_file(self) -> int:
        if not self.__valid:
            return 0
        h = _lib.default_handle(self.__device)
        sc = h.decode_scan(self.__file_data, [0, len(self.__file_data)], [self.__offset])
        _lib.raise_for_status(int(sc["status"][0]))
        _, bits = h.decode_reveal()
        self.output_bits = bits[0]
        pcm, _ = h.decode_run(exact=self.__exact)
        ch = max(int(sc["channels"][0]), 1)
        self.__pcm16 = pcm.reshape(-1, ch)
        self.__sampling_rate = int(sc["sample_rate"][0])
        self.__bit_rate = int(sc["bitrate"][0])
        return int(sc["n_frames"][0])

    def write_to_wav(self):
        if self.__sampling_rate is None:
            raise AttributeError("'Frame' object has no attribute 'sampling_rate'")   # what the reference does on an unparsed file
        write_wav(self.__wav_file_path, self.__sampling_rate, self.__pcm16)

    def get_bitrate(self) -> int:
        if self.__bit_rate is None:
            raise AttributeError("'Frame' object has no attribute 'bit_rate'")
        return self.__bit_rate

    @property
    def pcm16(self) -> np.ndarray:
        """int16 [rows, channels] exactly as write_to_wav stores them (extension: the reference keeps float64)."""
        return self.__pcm16


class Decoder:
    """Decoder(file_path, output_file_path).decode(quiet, reveal, txt_file_path) -> kbps; delete_wav_file()."""

    def __init__(self, file_path: str, output_file_path: str, device: int = 0, exact: bool = True):
        self.__file_path = file_path
        self.__output_file_path = output_file_path
        if not os.path.exists(self.__file_path):
            sys.exit(f"File {self.__file_path} not found.")
        with open(self.__file_path, "rb") as f:
            self.__data = np.frombuffer(f.read(), dtype=np.uint8)
        self.__offset = id3_offset(self.__data)
        self.__parser = MP3Parser(self.__data, self.__offset, self.__output_file_path, device=device, exact=exact)

    def decode(self, quiet: bool = True, reveal: bool = False, txt_file_path: str = "") -> int:
        start = time.time()
        num_of_parsed_frames = self.__parser.parse_file()
        parsing_time = time.time() - start
        if not quiet:
            print("\nParsed", num_of_parsed_frames, "frames in", parsing_time, "seconds.")
        self.__parser.write_to_wav()
        if not quiet:
            print(f"Wav file created on {self.__output_file_path}")
        if reveal:
            if txt_file_path[-4:] != ".txt":
                sys.exit("txt_file_path must be txt file.")
            with open(txt_file_path, "wb") as f:
                f.write(bytes(parse_reveal(self.__parser.output_bits), "utf-8"))
        return self.__parser.get_bitrate() // 1000

    def delete_wav_file(self):
        if os.path.exists(self.__output_file_path):
            os.remove(self.__output_file_path)
