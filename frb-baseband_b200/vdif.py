"""VDIF frame format helpers (host side).

The reference never parses VDIF itself; it shells out to `vdif_print_headers`
(/root/reference/base2fil.sh:130-147, /root/reference/extract_baseband_chunk.py:36-70) and
lets digifil read the split 2-channel files jive5ab writes
(/root/reference/spif2file.sh:178-186: 8000 B payload, 2 bits/sample, one R/L pair per
file).  This module holds the same facts as code: header field layout (VDIF 1.1), the frame
geometry base2fil derives (/root/reference/base2fil.sh:395-402) and a writer used by the
synthetic generator.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

FILL_WORD = 0x11223344
HEADER_BYTES = 32
LEGACY_HEADER_BYTES = 16


@dataclass
class FrameInfo:
    seconds: int
    frame_nr: int
    ref_epoch: int
    frame_bytes: int
    header_bytes: int
    nchan: int
    nbit: int
    legacy: int
    invalid: int
    thread: int
    station: int
    is_complex: int

    @property
    def payload_bytes(self) -> int:
        return self.frame_bytes - self.header_bytes

    @property
    def samples_per_frame(self) -> int:
        """time samples per frame (all channels share a time sample)"""
        return self.payload_bytes * 8 // (self.nbit * self.nchan)


def parse_header(buf) -> FrameInfo:
    w = np.frombuffer(bytes(buf[:16]), dtype="<u4")
    w0, w1, w2, w3 = (int(v) for v in w)
    legacy = (w0 >> 30) & 1
    return FrameInfo(
        seconds=w0 & 0x3FFFFFFF, frame_nr=w1 & 0xFFFFFF, ref_epoch=(w1 >> 24) & 0x3F,
        frame_bytes=(w2 & 0xFFFFFF) * 8, header_bytes=LEGACY_HEADER_BYTES if legacy else HEADER_BYTES,
        nchan=1 << ((w2 >> 24) & 0x1F), nbit=((w3 >> 26) & 0x1F) + 1, legacy=legacy,
        invalid=(w0 >> 31) & 1, thread=(w3 >> 16) & 0x3FF, station=w3 & 0xFFFF,
        is_complex=(w3 >> 31) & 1)


def make_headers(nframes: int, *, frames_per_sec: int, payload_bytes: int = 8000, nbit: int = 2,
                 log2_nchan: int = 1, ref_epoch: int = 40, sec0: int = 0, frame0: int = 0,
                 station: int = 0x4566, thread: int = 0) -> np.ndarray:
    """[nframes, 8] uint32 header words for a contiguous run of frames."""
    idx = np.arange(nframes, dtype=np.int64) + frame0
    sec = sec0 + idx // frames_per_sec
    fnr = idx % frames_per_sec
    h = np.zeros((nframes, 8), dtype="<u4")
    h[:, 0] = sec & 0x3FFFFFFF
    h[:, 1] = (fnr & 0xFFFFFF) | (ref_epoch << 24)
    h[:, 2] = ((payload_bytes + HEADER_BYTES) // 8) | (log2_nchan << 24)
    h[:, 3] = station | (thread << 16) | ((nbit - 1) << 26)
    return h


MARK5B_SYNC = 0xABADDEED
MARK5B_HEADER_BYTES = 16
MARK5B_PAYLOAD_BYTES = 10000


def _bcd(v: np.ndarray, ndigits: int) -> np.ndarray:
    out = np.zeros_like(v, dtype=np.uint32)
    for d in range(ndigits):
        out |= ((v // 10 ** d) % 10).astype(np.uint32) << np.uint32(4 * d)
    return out


def make_mark5b_headers(nframes: int, *, frames_per_sec: int, mjd: int = 60000, sec0: int = 0, frame0: int = 0) -> np.ndarray:
    """[nframes, 4] uint32 Mark5B disk frame headers: sync word, frame number within the second (bits 0..14), BCD
    time code JJJSSSSS (MJD mod 1000, second of day), BCD fraction of the second .SSSS (the CRC half is left 0).
    /root/reference/spif2file.sh:105-108: 16-byte header, 10000-byte payload."""
    idx = np.arange(nframes, dtype=np.int64) + frame0
    sec = sec0 + idx // frames_per_sec
    fnr = idx % frames_per_sec
    day = (mjd + sec // 86400) % 1000
    h = np.zeros((nframes, 4), dtype="<u4")
    h[:, 0] = MARK5B_SYNC
    h[:, 1] = fnr & 0x7FFF
    h[:, 2] = (_bcd(day, 3) << np.uint32(20)) | _bcd(sec % 86400, 5)
    h[:, 3] = _bcd(fnr * 10000 // frames_per_sec, 4) << np.uint32(16)
    return h


def epoch_mjd(ref_epoch: int) -> int:
    """MJD at 00:00 UTC of the start of a VDIF reference epoch."""
    year = 2000 + ref_epoch // 2
    month = 1 if ref_epoch % 2 == 0 else 7
    a = (14 - month) // 12
    y = year + 4800 - a
    m = month + 12 * a - 3
    jdn = 1 + (153 * m + 2) // 5 + 365 * y + y // 4 - y // 100 + y // 400 - 32045
    return jdn - 2400001


def frame_mjd(info: FrameInfo, frames_per_sec: float) -> float:
    return epoch_mjd(info.ref_epoch) + (info.seconds + info.frame_nr / frames_per_sec) / 86400.0


def frames_per_second(bw_mhz: float, info: FrameInfo) -> float:
    """Real Nyquist sampling at 2*|BW| Msamp/s per channel."""
    return 2.0 * abs(bw_mhz) * 1e6 / info.samples_per_frame


def nsec_in_file(file_size: int, frame_bytes: int, frames_per_sec_per_band: int) -> int:
    """Integer seconds held by a split file, as `bc` computes it in
    /root/reference/base2fil.sh:395-402 (integer division at every step)."""
    return file_size // frame_bytes // frames_per_sec_per_band
