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
