"""Source name -> dispersion measure, and the two header look-ups the reference does on a filterbank.

Host-side mirror of /root/reference/dm_utils.py: `get_dm` first consults the
operators' list of repeaters / localised FRBs (`dm_utils.py:12-93`), then asks `psrcat` (`:97-101`) and remembers
in `isPulsar` that the source is a pulsar; unknown sources give None (callers then search up to DM 1500,
`submit_job.py:52-54`).  FRB 20121102A ("R1") is deliberately absent upstream as well: BASELINE config 4 passes
DM 560 explicitly.  `get_src` / `get_nchan` read "Source Name" / "Number of channels" (`dm_utils.py:108-125`)
straight from the SIGPROC header instead of parsing the output of SIGPROC's `header` program.

The catalogue is kept as "DM : names sharing it" lines so that aliases of one source stay together.
"""
from __future__ import annotations

import subprocess

from . import sigproc

isPulsar = False   # set by get_dm when psrcat, not the catalogue, supplied the DM (dm_utils.py:6,99)

_CATALOGUE = """
    88.0 : FRB20200120 M81 M81R R200120
   103.0 : R4
     159 : FRB20181030E
   173.1 : FRB20190111A
   183.0 : NR4
   187.0 : NR2
   190.0 : R2
   195.8 : R15
   202.3 : FRB20200223B
   206.0 : FRB211212
     220 : FRB20220912A R220912
   220.0 : R54
   222.0 : R25
   223.7 : R17
     225 : FRB20190118A
  234.83 : FRB211127
   241.0 : LS63 LSI61 LSI63
   251.0 : NR7
   251.9 : FRB210807
   277.0 : NR1
     288 : FRB20181226B
   290.0 : R70
   301.7 : R14
   302.5 : FRB20190124C
     306 : FRB20190202A
   309.6 : R9
   323.2 : FRB20201114A
   325.0 : R34
   332.7 : BSGR SGR SGR1935
   338.7 : FRB190608
   349.7 : R3
   363.5 : R6
   365.0 : R47
     371 : FRB20180915A
   384.8 : FRB210320
   394.2 : R16
   400.0 : R24
   413.0 : R67
   415.0 : R68
   424.9 : R10
   443.0 : NR5
     444 : FRB20190518C
   444.0 : R7
   450.0 : R5
   460.2 : R11
   488.7 : FRB20190915D
   490.0 : R19
   504.1 : FRB190714A
   510.0 : R74
   517.0 : FRB180301 FRB20180901A R180301
   523.6 : FRB20191013D
   552.7 : R13
   578.9 : R12
   580.7 : FRB20181224E
   583.0 : FRB220105
   597.0 : NR6
   625.0 : R48
     690 : FRB20190122C
   714.0 : R21
   730.0 : FRB210117
   764.0 : NR3
   977.9 : R200616
  1202.0 : F19 FRB190520 R190520
  1281.5 : R8
    1349 : FRB20190103C
  1379.0 : FRB190417 R18
  1705.0 : R65
  1785.3 : FRB210407
"""


def _parse(text: str) -> dict:
    out = {}
    for line in text.strip().splitlines():
        value, names = line.split(":")
        v = value.strip()
        dm = float(v) if "." in v else int(v)          # the reference mixes ints and floats; keep them as written
        for name in names.split():
            out[name] = dm
    return out


FRB_DMS = _parse(_CATALOGUE)


def get_dm(src: str, *, psrcat: str = "psrcat"):
    """DM of `src` in pc/cc, or None when neither the catalogue nor psrcat knows it."""
    global isPulsar
    if src in FRB_DMS:
        return FRB_DMS[src]
    try:
        out = subprocess.check_output([psrcat, "-c", "dm", "-o", "short", "-nohead", "-nonumber", src],
                                      stderr=subprocess.DEVNULL)
        dm = float(out)
    except (OSError, subprocess.CalledProcessError, ValueError):
        return None
    isPulsar = True
    return dm


def _header(fil_file: str) -> sigproc.FilHeader:
    with open(fil_file, "rb") as f:
        return sigproc.read_header(f.read(4096))[0]


def get_src(fil_file: str) -> str:
    return _header(fil_file).source_name.strip()


def get_nchan(fil_file: str) -> int:
    return int(_header(fil_file).nchans)
