"""frb.conf reader.

The reference's config is a bash file that base2fil.sh sources on top of its defaults
(/root/reference/base2fil.sh:207-233; schema /root/reference/frb.conf:1-54).  Values may be
bash arrays (`scans=( 001 002 )`, /root/reference/create_config.py:435-448) and may use
${USER}/${HOME}, so the file is evaluated by bash itself -- the same interpreter the
pipeline uses -- and the resulting variables are read back.
"""
from __future__ import annotations

import os
import shlex
import subprocess
from dataclasses import dataclass, field

#: defaults base2fil.sh sets before sourcing the config (/root/reference/base2fil.sh:207-230)
DEFAULTS = {
    "workdir_odd_base": "/scratch0/${USER}/", "workdir_even_base": "/scratch1/${USER}/",
    "outdir_base": "/data1/${USER}/", "fifodir_base": "/tmp/${USER}/", "vbsdir_base": "${HOME}/vbs_data/",
    "start": "0", "pol": "2", "digifil_nthreads": "1", "flipIF": "0", "njobs_parallel": "20",
    "submit2fetch": "0", "nbit": "8", "isMark5b": "0", "keepVDIF": "0", "flagFile": "", "keepBP": "0",
    "split_vdif_only": "0", "online_process": "0", "nbits": "2",
}
REQUIRED = ("experiment", "target", "scans", "skips", "lengths", "scannames", "bw", "nif", "freqLSB_0", "station",
            "nchan", "tscrunch")
ARRAYS = ("scans", "skips", "lengths", "scannames")
_ALL = tuple(DEFAULTS) + REQUIRED + ("frame_size",)


@dataclass
class FrbConf:
    values: dict = field(default_factory=dict)

    def __getattr__(self, k):
        try:
            return self.__dict__["values"][k]
        except KeyError:
            raise AttributeError(k)

    # derived quantities, as base2fil.sh computes them
    @property
    def datarate(self) -> int:
        """Mbps = bw*nif*nbits*4 (2 pol x 2 Nyquist), /root/reference/base2fil.sh:251"""
        return int(float(self.bw) * int(self.nif) * int(self.nbits) * 4)

    @property
    def freqUSB_0(self) -> float:
        return float(self.freqLSB_0) + float(self.bw)          # base2fil.sh:254

    def if_plan(self):
        """[(IF number, centre MHz, 'l'|'u')] for IF 1..nif: odd IFs LSB stepping 2*bw from
        freqLSB_0, even IFs USB stepping 2*bw from freqLSB_0+bw (base2fil.sh:54,65,267-268,407-414)."""
        bw = float(self.bw)
        out = []
        for i in range(1, int(self.nif) + 1):
            if i % 2:
                out.append((i, float(self.freqLSB_0) + (i - 1) // 2 * 2 * bw, "l"))
            else:
                out.append((i, self.freqUSB_0 + (i - 2) // 2 * 2 * bw, "u"))
        return out

    def target_args(self) -> list[str]:
        """`target` is expanded unquoted by base2fil (it may hold '--ra .. --dec ..', frb.conf:3-6)"""
        return shlex.split(str(self.target))


def read_conf(path: str, env: dict | None = None) -> FrbConf:
    """Evaluate the config like base2fil.sh does and return its variables."""
    script = ["set -a"]
    for k, v in DEFAULTS.items():
        script.append(f'{k}="{v}"')
    script.append('source "$1"')
    for k in _ALL:
        script.append(f'if declare -p {k} >/dev/null 2>&1; then if [[ "$(declare -p {k})" == "declare -a"* ]]; '
                      f'then printf "%s\\0A\\0" {k}; printf "%s\\0" "${{{k}[@]}}"; printf "\\0"; '
                      f'else printf "%s\\0S\\0%s\\0\\0" {k} "${{{k}}}"; fi; fi')
    e = dict(os.environ)
    e.setdefault("USER", "user")
    e.setdefault("HOME", "/tmp")
    if env:
        e.update(env)
    out = subprocess.run(["bash", "-c", "\n".join(script), "read_conf", path], capture_output=True, env=e)
    if out.returncode != 0:
        raise ValueError(f"cannot source {path}: {out.stderr.decode(errors='replace')}")
    toks = out.stdout.split(b"\0")
    vals: dict = {}
    i = 0
    while i + 1 < len(toks):
        name, kind = toks[i].decode(), toks[i + 1].decode()
        if not name:
            break
        i += 2
        if kind == "S":              # name, S, value, terminator
            vals[name] = toks[i].decode()
            i += 2
            continue
        items = []
        while i < len(toks) and toks[i] != b"":
            items.append(toks[i].decode())
            i += 1
        i += 1                       # terminator
        vals[name] = items
    for k in ARRAYS:                 # a scalar is a one-element array to bash's "${x[@]}"
        if k in vals and not isinstance(vals[k], list):
            vals[k] = vals[k].split() if vals[k] else []
    missing = [k for k in REQUIRED if k not in vals or vals[k] in ("", [])]
    if missing:
        raise ValueError(f"{path}: required keys not set: {', '.join(missing)}")
    return FrbConf(vals)
