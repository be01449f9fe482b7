"""SIGPROC filterbank header writer / reader.

digifil writes this header at the top of every per-IF .fil (/root/reference/process_vdif.py:143-145)
and `splice` re-emits the first file's header with nchans summed
(/root/reference/base2fil.sh:422).  The reference reads two fields back through SIGPROC's
`header` tool: "Source Name" and "Number of channels" (/root/reference/dm_utils.py:108-125).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

#: stock SIGPROC telescope ids (aliases.c); the pharaofranz fork's additions are unknown
#: -> unknown names map to 0
TELESCOPE_IDS = {
    "arecibo": 1, "ooty": 2, "nancay": 3, "parkes": 4, "jodrell": 5, "gbt": 6, "gmrt": 7,
    "effelsberg": 8, "ata": 9, "srt": 10, "lofar": 11, "vla": 12, "chime": 20, "meerkat": 64,
}

_INT_KEYS = ("telescope_id", "machine_id", "data_type", "barycentric", "pulsarcentric", "nbits", "nsamples",
             "nchans", "nifs", "nbeams", "ibeam")
_DBL_KEYS = ("az_start", "za_start", "src_raj", "src_dej", "tstart", "tsamp", "fch1", "foff", "refdm", "period")
_STR_KEYS = ("rawdatafile", "source_name")


def _s(b: str) -> bytes:
    e = b.encode()
    return struct.pack("<i", len(e)) + e


def sexagesimal_to_sigproc(v: str | None) -> float:
    """'hh:mm:ss.ss' / 'dd:mm:ss.ss' -> hhmmss.ss as a double (SIGPROC src_raj / src_dej)."""
    if v is None:
        return 0.0
    v = str(v).strip()
    sign = -1.0 if v.startswith("-") else 1.0
    parts = v.lstrip("+-").split(":")
    parts += ["0"] * (3 - len(parts))
    return sign * (abs(float(parts[0])) * 10000.0 + float(parts[1]) * 100.0 + float(parts[2]))


@dataclass
class FilHeader:
    source_name: str = "unknown"
    rawdatafile: str = ""
    telescope_id: int = 0
    machine_id: int = 0
    data_type: int = 1
    barycentric: int = 0
    pulsarcentric: int = 0
    az_start: float = 0.0
    za_start: float = 0.0
    src_raj: float = 0.0
    src_dej: float = 0.0
    tstart: float = 0.0
    tsamp: float = 0.0
    nbits: int = 8
    fch1: float = 0.0
    foff: float = 0.0
    nchans: int = 0
    nifs: int = 1
    refdm: float | None = None
    extra: dict = field(default_factory=dict)

    def pack(self) -> bytes:
        out = [_s("HEADER_START")]
        for k in ("telescope_id", "machine_id", "data_type"):
            out.append(_s(k) + struct.pack("<i", int(getattr(self, k))))
        out.append(_s("rawdatafile") + _s(self.rawdatafile))
        out.append(_s("source_name") + _s(self.source_name))
        for k in ("barycentric", "pulsarcentric"):
            out.append(_s(k) + struct.pack("<i", int(getattr(self, k))))
        for k in ("az_start", "za_start", "src_raj", "src_dej", "tstart", "tsamp"):
            out.append(_s(k) + struct.pack("<d", float(getattr(self, k))))
        out.append(_s("nbits") + struct.pack("<i", int(self.nbits)))
        for k in ("fch1", "foff"):
            out.append(_s(k) + struct.pack("<d", float(getattr(self, k))))
        out.append(_s("nchans") + struct.pack("<i", int(self.nchans)))
        out.append(_s("nifs") + struct.pack("<i", int(self.nifs)))
        if self.refdm is not None:
            out.append(_s("refdm") + struct.pack("<d", float(self.refdm)))
        out.append(_s("HEADER_END"))
        return b"".join(out)


def read_header(buf: bytes) -> tuple[FilHeader, int]:
    """Parse a SIGPROC header; returns (header, byte offset of the first sample)."""
    pos = 0

    def rstr():
        nonlocal pos
        (n,) = struct.unpack_from("<i", buf, pos)
        pos += 4
        if n < 0 or n > 80:
            raise ValueError("not a SIGPROC header")
        s = buf[pos:pos + n].decode()
        pos += n
        return s

    if rstr() != "HEADER_START":
        raise ValueError("missing HEADER_START")
    h = FilHeader()
    while True:
        k = rstr()
        if k == "HEADER_END":
            break
        if k in _INT_KEYS:
            (v,) = struct.unpack_from("<i", buf, pos)
            pos += 4
        elif k in _DBL_KEYS:
            (v,) = struct.unpack_from("<d", buf, pos)
            pos += 8
        elif k in _STR_KEYS:
            v = rstr()
        else:
            raise ValueError(f"unknown SIGPROC keyword {k}")
        if hasattr(h, k):
            setattr(h, k, v)
        else:
            h.extra[k] = v
    return h, pos


def read_fil(path: str):
    """(header, data[rows, nifs, nchans]) of a filterbank file; 8/16/32-bit samples."""
    raw = open(path, "rb").read()
    h, off = read_header(raw)
    dt = {8: np.uint8, 16: np.uint16, 32: np.float32}[h.nbits]
    d = np.frombuffer(raw, dtype=dt, offset=off)
    per = h.nifs * h.nchans
    return h, d[: d.size // per * per].reshape(-1, h.nifs, h.nchans)


def header_report(h: FilHeader) -> str:
    """The two lines of SIGPROC `header` output the reference parses (/root/reference/dm_utils.py:108-125)."""
    return (f"Source Name                      : {h.source_name}\n"
            f"Number of channels               : {h.nchans}\n")
