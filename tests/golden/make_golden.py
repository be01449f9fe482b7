#!/usr/bin/env python3
"""Generate golden fixtures by importing the reference's own Python (only possible in the build
container: /root/reference does not exist on the GPU box).  Outputs are committed next to this
script; tests/test_host_dropin.py compares our drop-in against them.

    python tests/golden/make_golden.py
"""
import importlib.util
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/process_vdif.py"

spec = importlib.util.spec_from_file_location("ref_process_vdif", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

cases = [
    dict(psr="R3", freq=1254.0, pol=2, usb=False, ra="01:58:00.7502", dec="65:43:00.3152", bw=32.0, telescope="effelsberg"),
    dict(psr="B0355+54", freq=1608.0, pol=4, usb=True, ra="03:58:53.7", dec="+54:13:13.7", bw=16.0, telescope="ONSALA85"),
    dict(psr="FRB190520", freq=4926.49, pol=0, usb=True, ra="16:02:04.27", dec="-11:17:17.3", bw=64.0, telescope="srt", npol=2),
]
out = []
with tempfile.TemporaryDirectory() as d:
    for k, c in enumerate(cases):
        fn = os.path.join(d, f"ek048c_ef_no0{k:03d}_IF{k + 1}.vdif")
        hdr = ref.make_hdr(filename=fn, **c)
        text = open(hdr).read().replace(d, "<DIR>")
        out.append({"args": c, "basename": os.path.basename(fn), "hdr_name": os.path.basename(hdr), "text": text})

# argparse defaults of the reference CLI (flag -> default), the contract base2fil.sh relies on
sys.argv = ["process_vdif", "SRC", "/x/y.vdif", "-l"]
ns = vars(ref.options())
json.dump({"make_hdr": out, "cli_defaults": {k: v for k, v in sorted(ns.items())}}, open(os.path.join(HERE, "process_vdif_reference.json"), "w"), indent=1)
print("wrote", os.path.join(HERE, "process_vdif_reference.json"))


# ----------------------------------------------------------------------------------------------
# submit_job.py chooser: run the reference's own main() with its
# external calls (psrcat, create_config.py, base2fil) captured instead of executed.
def submit_job_golden():
    import types
    sys.path.insert(0, "/root/reference")
    spec = importlib.util.spec_from_file_location("ref_submit_job", "/root/reference/submit_job.py")
    sj = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sj)
    cases = []
    grid = [
        # dm, pulsar period (None = FRB branch), fref, IF, nIF
        (560.0, None, 1286.0, 32.0, 8), (332.7, None, 1286.0, 32.0, 8), (None, None, 1286.0, 32.0, 8),
        (1785.3, None, 1286.0, 32.0, 8), (87.75, None, 1624.0, 16.0, 8), (3000.0, None, 346.0, 16.0, 4),
        (100.0, None, 4926.0, 64.0, 4), (1202.0, None, 1286.0, 32.0, 16), (8000.0, None, 1286.0, 32.0, 8),
        (26.76, 0.714519699726, 1286.0, 32.0, 8), (57.14, 0.156384121559, 1624.0, 16.0, 8),
        (2.97, 1.2922413, 1286.0, 32.0, 8), (56.8, 0.033, 1286.0, 32.0, 8),
    ]
    for dm, period, fref, IF, nif in grid:
        captured = []
        sj.dm.get_dm = lambda src, _dm=dm: _dm
        sj.dm.isPulsar = period is not None
        sj.os.system = lambda cmd: captured.append(cmd) or 0
        sj.check_output = lambda cmd, shell=True, _p=period: f"{_p:.12f}\n".encode()
        args = types.SimpleNamespace(vex="/vex/ek048c.vix", expname=None, telescope="ef", source="SRC", scannum="012",
                                     fref=fref, IF=IF, nIF=nif)
        sj.main(args)
        cc = captured[0]
        tail = " --search" if period is None else " --pol 4"
        assert cc.endswith(tail)
        body = cc[: -len(tail)]                       # the reference appends the command to itself (:119-122)
        assert body[: len(body) // 2] == body[len(body) // 2:]
        first = body[: len(body) // 2].split()
        get = lambda flag: first[first.index(flag) + 1]
        cases.append({"dm": dm, "period_s": period, "fref": fref, "IF": IF, "nIF": nif,
                      "tscrunch": int(get("-d")), "nchan_if": int(get("-n")), "flag_file": get("-F"),
                      "config_file": get("-o"), "first_argv": first, "tail": tail.split(), "submit": captured[1]})
    json.dump({"submit_job": cases}, open(os.path.join(HERE, "submit_job_reference.json"), "w"), indent=1)
    print("wrote", os.path.join(HERE, "submit_job_reference.json"))


submit_job_golden()
