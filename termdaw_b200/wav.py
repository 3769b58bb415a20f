"""RIFF/WAVE container I/O for the render path's two file boundaries.

termdaw decodes samples with hound (`sample.rs:231-274`) and encodes the render with
`hound::WavWriter` (`state.rs:508-531`).  hound 3.5 is not vendored with the reference, so this file
follows the WAV standard and hound's documented behaviour:

* reader: `samples::<i32>()` yields sign-extended integers for 8 (stored unsigned, offset 128), 16, 24
  and 32-bit PCM, which termdaw casts `as f32` UNSCALED (`sample.rs:266-273`); IEEE-float files yield
  f32 samples as they are;
* writer: 16-byte PCM `fmt ` chunk (44-byte header) for <= 16-bit, WAVE_FORMAT_EXTENSIBLE (40-byte
  `fmt ` chunk) above 16 bit; the data chunk is exactly the bytes `tdg_render` returns.

File decoding/encoding is host I/O; all arithmetic on the decoded stream runs on the device
(`tdg_sample_add`).
"""
from __future__ import annotations

import struct
from typing import Tuple

import numpy as np

from .state import DecodedWav

KSDATAFORMAT_SUBTYPE_PCM = bytes([0x01, 0x00, 0x00, 0x00, 0x00, 0x00, 0x10, 0x00,
                                  0x80, 0x00, 0x00, 0xAA, 0x00, 0x38, 0x9B, 0x71])
KSDATAFORMAT_SUBTYPE_FLOAT = bytes([0x03]) + KSDATAFORMAT_SUBTYPE_PCM[1:]


class WavError(ValueError):
    pass


def _chunks(buf: bytes):
    if len(buf) < 12 or buf[:4] != b"RIFF" or buf[8:12] != b"WAVE":
        raise WavError("not a RIFF/WAVE file")
    pos = 12
    while pos + 8 <= len(buf):
        cid = buf[pos:pos + 4]
        (size,) = struct.unpack_from("<I", buf, pos + 4)
        body = buf[pos + 8: pos + 8 + size]
        yield cid, body
        pos += 8 + size + (size & 1)


def decode_wav(buf: bytes) -> DecodedWav:
    fmt = data = None
    for cid, body in _chunks(buf):
        if cid == b"fmt " and fmt is None:
            fmt = body
        elif cid == b"data" and data is None:
            data = body
    if fmt is None or data is None or len(fmt) < 16:
        raise WavError("missing fmt or data chunk")
    tag, channels, rate, _, block_align, bits = struct.unpack_from("<HHIIHH", fmt, 0)
    if tag == 0xFFFE:                                    # WAVE_FORMAT_EXTENSIBLE
        if len(fmt) < 40:
            raise WavError("short extensible fmt chunk")
        sub = fmt[24:40]
        tag = 3 if sub == KSDATAFORMAT_SUBTYPE_FLOAT else 1
    if channels == 0:
        raise WavError("zero channels")
    nbytes = bits // 8
    n = len(data) // nbytes if nbytes else 0
    raw = np.frombuffer(data, dtype=np.uint8, count=n * nbytes)
    if tag == 3:
        if bits != 32:
            raise WavError("only 32-bit float WAV is supported")
        lin = raw.view("<f4").astype(np.float32)
    elif tag == 1:
        if bits == 8:
            lin = (raw.astype(np.int32) - 128).astype(np.float32)
        elif bits == 16:
            lin = raw.view("<i2").astype(np.float32)
        elif bits == 24:
            t = raw.reshape(-1, 3).astype(np.int32)
            v = t[:, 0] | (t[:, 1] << 8) | (t[:, 2] << 16)
            lin = np.where(v & 0x800000, v - (1 << 24), v).astype(np.float32)
        elif bits == 32:
            lin = raw.view("<i4").astype(np.float32)     # i32 `as f32` rounds to nearest, like numpy
        else:
            raise WavError(f"unsupported bit depth {bits}")
    else:
        raise WavError(f"unsupported WAV format tag {tag}")
    return DecodedWav(lin, int(channels), int(rate), int(bits))


def read_wav(path: str) -> DecodedWav:
    with open(path, "rb") as f:
        return decode_wav(f.read())


def wav_header(data_bytes: int, sample_rate: int, bitdepth: int, channels: int = 2) -> bytes:
    block = channels * (bitdepth // 8)
    if bitdepth <= 16:
        fmt = struct.pack("<HHIIHH", 1, channels, sample_rate, sample_rate * block, block, bitdepth)
    else:
        mask = {1: 0x4, 2: 0x3}.get(channels, 0)
        fmt = struct.pack("<HHIIHHHHI", 0xFFFE, channels, sample_rate, sample_rate * block, block, bitdepth,
                          22, bitdepth, mask) + KSDATAFORMAT_SUBTYPE_PCM
    riff_size = 4 + (8 + len(fmt)) + (8 + data_bytes)
    return b"RIFF" + struct.pack("<I", riff_size) + b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + \
        b"data" + struct.pack("<I", data_bytes)


def encode_wav(pcm, sample_rate: int, bitdepth: int, channels: int = 2) -> bytes:
    body = bytes(memoryview(np.ascontiguousarray(pcm, dtype=np.uint8)))
    return wav_header(len(body), sample_rate, bitdepth, channels) + body


def write_wav(path: str, pcm, sample_rate: int, bitdepth: int, channels: int = 2) -> None:
    """`pcm` is the interleaved little-endian data chunk as returned by `Context.render`."""
    with open(path, "wb") as f:
        f.write(encode_wav(pcm, sample_rate, bitdepth, channels))


def pcm_from_float(l: np.ndarray, r: np.ndarray, bitdepth: int) -> np.ndarray:
    """Host-side packer used only to SYNTHESISE input sample files for tests (never on the render
    path): float [-1, 1] -> interleaved PCM bytes."""
    inter = np.empty(l.size * 2, np.float64)
    inter[0::2], inter[1::2] = l, r
    if bitdepth == 8:
        return (np.clip(np.round(inter * 127.0), -128, 127).astype(np.int16) + 128).astype(np.uint8)
    if bitdepth == 16:
        return np.clip(np.round(inter * 32767.0), -32768, 32767).astype("<i2").view(np.uint8)
    if bitdepth == 24:
        v = np.clip(np.round(inter * 8388607.0), -8388608, 8388607).astype(np.int32)
        out = np.empty((v.size, 3), np.uint8)
        out[:, 0], out[:, 1], out[:, 2] = v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF
        return out.reshape(-1)
    if bitdepth == 32:
        return np.clip(np.round(inter * 2147483647.0), -2147483648, 2147483647).astype("<i4").view(np.uint8)
    raise WavError(f"unsupported bit depth {bitdepth}")


def frames_and_channels(w: DecodedWav) -> Tuple[int, int]:
    return w.linear.size // max(1, w.channels), w.channels
