"""Host-side WAV I/O with the reference's semantics.

write_wav   what scipy.io.wavfile.write emits for an int16 array (MP3Parser.write_to_wav, decoder/MP3_Parser.py:87-91)
WavReader   encoder/WAV_Reader.py:20-164 restated: RIFF header sniffing by substring search in the first 128 bytes,
            the same sys.exit messages, and the same (quirky) sample buffer: int16 words from the end of the data
            header to the end of the file, at most 2 * num_of_samples * channels of them.
"""
import struct
import sys

import numpy as np

MPEG1_L3_BITRATES = (32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320)  # encoder/util.py:24-31, version 3 row
SAMPLERATES = (44100, 48000, 32000)                                                   # encoder/util.py:36-48, MPEG-1 rows


def write_wav(path: str, sample_rate: int, pcm16: np.ndarray) -> None:
    pcm16 = np.ascontiguousarray(pcm16, dtype="<i2")
    nch = 1 if pcm16.ndim == 1 else pcm16.shape[1]
    data = pcm16.tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, 1, nch, sample_rate, sample_rate * nch * 2, nch * 2, 16) + b"data" + struct.pack("<I", len(data))
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(data)


class WavReader:
    """Same constructor, properties and failure messages as the reference's WavReader."""

    def __init__(self, file_path: str, bit_rate: int = 320):
        self.__file_path = file_path
        self.__bitrate = bit_rate
        with open(file_path, "rb") as f:
            raw = f.read()
        buffer = raw[:128]
        idx = buffer.find(b"RIFF")
        if idx == -1:
            sys.exit("Bad WAVE file.")
        if buffer.find(b"WAVE") == -1:
            sys.exit("Bad WAVE file.")
        idx = buffer.find(b"fmt ")
        if idx == -1:
            sys.exit("Bad WAVE file.")
        idx += 4
        if struct.unpack("<I", buffer[idx:idx + 4])[0] != 16:
            sys.exit("Unsupported WAVE file, compression used instead of PCM.")
        idx += 4
        if struct.unpack("<H", buffer[idx:idx + 2])[0] != 1:
            sys.exit("Unsupported WAVE file, compression used instead of PCM.")
        idx += 2
        self.__num_of_ch = struct.unpack("<H", buffer[idx:idx + 2])[0]
        self.__mpeg_mode = 0 if self.__num_of_ch > 1 else 3          # util.MODES: STEREO 0, MONO 3
        idx += 2
        self.__samplerate = struct.unpack("<I", buffer[idx:idx + 4])[0]
        if self.__samplerate not in (32000, 44100, 48000):
            sys.exit("Unsupported sampling frequency.")
        idx += 4 + 4 + 2
        self.__bits_per_sample = struct.unpack("<H", buffer[idx:idx + 2])[0]
        if self.__bits_per_sample not in (8, 16, 32):
            sys.exit("Unsupported WAVE file, samples not int8, int16 or int32 type.")
        idx = buffer.find(b"data")
        if idx == -1:
            sys.exit("Bad WAVE file.")
        idx += 4
        sub_chunk2_size = struct.unpack("<I", buffer[idx:idx + 4])[0]
        self.__num_of_samples = int(sub_chunk2_size * 8 / self.__bits_per_sample / self.__num_of_ch)
        body = raw[idx + 4:]
        count = min(len(body) // 2, self.__num_of_samples * self.__num_of_ch * 2)   # np.fromfile(..., 'int16', count)
        self.__buffer = np.frombuffer(body, dtype="<i2", count=count)
        self.__buffer_pos = {0: 0, 1: 1}
        self.__copyright = 0
        self.__original = 1
        self.__emphasis = 0
        if bit_rate not in MPEG1_L3_BITRATES:
            sys.exit("Unsupported bitrate configuration.")
        if self.__samplerate not in SAMPLERATES:
            sys.exit("Unsupported samplerate configuration.")

    mpeg_mode = property(lambda self: self.__mpeg_mode)
    bitrate = property(lambda self: self.__bitrate)
    emphasis = property(lambda self: self.__emphasis)
    copyright = property(lambda self: self.__copyright)
    original = property(lambda self: self.__original)
    samplerate = property(lambda self: self.__samplerate)
    num_of_channels = property(lambda self: self.__num_of_ch)
    file_path = property(lambda self: self.__file_path)
    num_of_samples = property(lambda self: self.__num_of_samples)
    buffer = property(lambda self: self.__buffer)

    def get_buffer_pos(self, ch):
        return self.__buffer_pos[ch]

    def set_buffer_pos(self, ch, offset):
        self.__buffer_pos[ch] += offset
