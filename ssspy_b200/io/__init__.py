"""Canonical 16-byte-fmt PCM WAV reader / writer (host mirror of ssspy/io/__init__.py: wavread :8-63,
wavwrite :66-121), the byte-level step either side of the separation path.

Pure host code (bytes in, NumPy out); the behaviour follows the reference: little-endian RIFF/WAVE only, the ``fmt ``
chunk must be the 16-byte PCM one and be followed directly by ``data``; samples are scaled by ``2 ** (bits - 1)``;
floating-point input is written as 16-bit PCM by truncation towards zero of ``x * 32768`` (``astype``), 8- and
16-bit integer input as it is; at most two channels on write.  Written from the RIFF layout:

    0  "RIFF"  u32 file_size - 8   "WAVE"
    12 "fmt "  u32 16   u16 format=1   u16 channels   u32 rate   u32 byte_rate   u16 block_align   u16 bits
    36 "data"  u32 n_bytes   samples (interleaved, little endian)
"""
import struct

import numpy as np

__all__ = ["wavread", "wavwrite"]

_HEAD = struct.Struct("<4sI4s")          # RIFF, size, WAVE
_FMT = struct.Struct("<HHIIHH")          # format, channels, rate, byte_rate, block_align, bits


def _expect(tag, want):
    if tag != want:
        raise NotImplementedError("Not support {!r}.".format(tag))


def wavread(path, frame_offset=0, num_frames=None, return_2d=None, channels_first=None):
    """Returns ``(waveform, sample_rate)``: float64 in [-1, 1); shape (n_frames,) for mono unless ``return_2d``,
    (n_frames, n_channels) for multichannel, transposed when ``channels_first``."""
    with open(path, "rb") as f:
        riff, _, wave = _HEAD.unpack(f.read(_HEAD.size))
        _expect(riff, b"RIFF")
        _expect(wave, b"WAVE")
        _expect(f.read(4), b"fmt ")
        (fmt_size,) = struct.unpack("<I", f.read(4))
        if fmt_size != _FMT.size:
            raise NotImplementedError("Invalid header is detected.")
        fmt, n_channels, sample_rate, byte_rate, block_align, bits = _FMT.unpack(f.read(_FMT.size))
        if fmt != 1:
            raise NotImplementedError("Invalid header {} is detected.".format(fmt))
        if bits * sample_rate * n_channels != 8 * byte_rate:
            raise ValueError("Invalid header is detected.")
        _expect(f.read(4), b"data")
        (n_bytes,) = struct.unpack("<I", f.read(4))
        width = block_align // n_channels          # bytes per sample
        max_frame = n_bytes // block_align
        if num_frames is None:
            n_samples = n_bytes // width - n_channels * frame_offset
            end_frame = max_frame
        elif num_frames >= 0:
            n_samples = n_channels * num_frames
            end_frame = frame_offset + num_frames
        else:
            raise ValueError("Invalid num_frames={} is given. Set nonnegative integer.".format(num_frames))
        if end_frame > max_frame:
            raise ValueError("num_frames={} exceeds maximum frame {}.".format(num_frames, max_frame))
        f.seek(block_align * frame_offset, 1)
        raw = f.read(n_samples * width)
    data = np.frombuffer(raw, dtype="<i{}".format(width), count=n_samples)
    if n_channels > 1 or return_2d:
        data = data.reshape(-1, n_channels)
        if channels_first:
            data = data.transpose(1, 0)
    return data / 2 ** (8 * width - 1), sample_rate


def wavwrite(path, waveform, sample_rate, channels_first=None):
    """Writes ``waveform`` (1-D, or 2-D with one or two channels) as PCM."""
    assert path[-4:] == ".wav", "Only wav file is supported."
    if waveform.ndim == 1:
        frames, n_channels = waveform, 1
    elif waveform.ndim == 2:
        frames = waveform.transpose(1, 0) if channels_first else waveform
        n_channels = frames.shape[1]
        if not 1 <= n_channels <= 2:
            raise ValueError("{}channel-input is not supported.".format(n_channels))
    else:
        raise ValueError("waveform.ndim should be less or equal to 2, but given {}.".format(waveform.ndim))
    if frames.dtype.kind == "f":
        bits = 16
        frames = (frames * 2 ** (bits - 1)).astype("<i2")
    elif frames.dtype == np.dtype("i1"):
        bits = 8
    elif frames.dtype == np.dtype("i2"):
        bits = 16
    else:
        raise ValueError("Invalid dtype={} is detected.".format(frames.dtype))
    byte_rate = bits * sample_rate * n_channels // 8
    block_align = byte_rate // sample_rate
    payload = np.ascontiguousarray(frames).tobytes()
    body = (b"WAVE" + b"fmt " + struct.pack("<I", _FMT.size)
            + _FMT.pack(1, n_channels, sample_rate, byte_rate, block_align, bits)
            + b"data" + struct.pack("<I", len(payload)) + payload)
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)
