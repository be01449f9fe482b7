"""The C-ABI library loads and exports every symbol include/b2f.h declares (no compute
calls here: this file runs without a GPU)."""
import ctypes as C
import os
import re

import pytest

from frb_baseband_b200 import _lib
from frb_baseband_b200.plan import Plan, PlanConfig, pol_mode_from_reference, reference_freq_res

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b2f.h")).read()
    return sorted(set(re.findall(r"\b(b2f_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    l = _lib.lib()
    names = header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(l, n), f"{n} declared in include/b2f.h but not exported by libb2f.so"
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)


def test_version_and_struct_size():
    l = _lib.lib()
    assert l.b2f_version() == 100
    # a wrong struct_size must be refused before anything touches the GPU
    p = _lib.Params()
    p.struct_size = 4
    h = C.c_void_p()
    assert l.b2f_plan_create(C.byref(p), C.byref(h)) == _lib.EINVAL
    assert b"struct_size" in l.b2f_last_error()


def test_parameter_validation_mirrors_process_vdif():
    # process_vdif.py:153-155 nbit whitelist, :175-176 pol whitelist
    with pytest.raises(_lib.B2FError) as e:
        Plan(PlanConfig(nchan=128, bw_mhz=[-32.0], out_nbit=4))
    assert e.value.code == _lib.EINVAL and "nbit" in e.value.message
    with pytest.raises(ValueError):
        pol_mode_from_reference(5)
    assert reference_freq_res(128) == 512 and reference_freq_res(512) == 1024    # process_vdif.py:162


def test_no_cpu_fallback():
    """Without a device the product path fails loudly (B2F_ECUDA), it never computes on the CPU."""
    l = _lib.lib()
    if l.b2f_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(_lib.B2FError) as e:
        Plan(PlanConfig(nchan=128, bw_mhz=[-32.0]))
    assert e.value.code == _lib.ECUDA and "no CPU fallback" in e.value.message


def test_sigproc_header_matches_python_writer():
    """b2f_sigproc_header (pure host code inside libb2f) == sigproc.FilHeader.pack byte for byte."""
    import ctypes as C
    from frb_baseband_b200 import sigproc
    for refdm in (None, 560.0):
        h = _lib.FilHeaderC()
        h.struct_size = C.sizeof(h)
        h.source_name, h.rawdatafile, h.telescope_id = b"R3", b"ek048c_ef_no0001_IF8.vdif", 8
        h.src_raj, h.src_dej, h.tstart_mjd, h.tsamp_s = 15800.7502, 654300.3152, 58849.123456789, 6.4e-5
        h.nbits, h.fch1_mhz, h.foff_mhz, h.nchans, h.nifs = 8, 1509.875, -0.25, 1024, 4
        h.write_refdm, h.refdm = int(refdm is not None), refdm or 0.0
        n = C.c_size_t()
        _lib.check(_lib.lib().b2f_sigproc_header(C.byref(h), None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        _lib.check(_lib.lib().b2f_sigproc_header(C.byref(h), buf, n.value, C.byref(n)))
        py = sigproc.FilHeader(source_name="R3", rawdatafile="ek048c_ef_no0001_IF8.vdif", telescope_id=8,
                               src_raj=15800.7502, src_dej=654300.3152, tstart=58849.123456789, tsamp=6.4e-5, nbits=8,
                               fch1=1509.875, foff=-0.25, nchans=1024, nifs=4, refdm=refdm).pack()
        assert buf.raw == py
        hh, off = sigproc.read_header(buf.raw)
        assert off == n.value and hh.nchans == 1024 and hh.source_name == "R3"
    assert _lib.lib().b2f_sigproc_header(C.byref(h), buf, 8, C.byref(n)) == _lib.EINVAL


def test_header_is_plain_c_and_links(tmp_path):
    """include/b2f.h compiles as strict C99 and a C program links against libb2f.so: the boundary a digifil- or
    splice-like tool written in C would use (examples/b2f_scan.c is mode B of INTEGRATION.md from C)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "b2f_scan")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                           os.path.join(root, "examples", "b2f_scan.c"), "-o", exe, "-L", libdir, "-lb2f",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe, "--version"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("libb2f 0.1.0")
    bad = subprocess.run([exe, str(tmp_path / "o.fil")], capture_output=True, text=True)
    assert bad.returncode == 2 and "usage" in bad.stderr


def test_ctypes_structs_match_the_c_layout(tmp_path):
    """sizeof of every struct of include/b2f.h as gcc lays it out == the ctypes mirror in _lib.py."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "b2f.h"\nint main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", '
                   "sizeof(b2f_params), sizeof(b2f_geometry), sizeof(b2f_counters), sizeof(b2f_fil_header), "
                   "sizeof(b2f_scan_io), sizeof(b2f_scan_result)); return 0; }\n")
    exe = str(tmp_path / "sizes")
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe])
    got = [int(x) for x in subprocess.check_output([exe], text=True).split()]
    want = [C.sizeof(t) for t in (_lib.Params, _lib.Geometry, _lib.Counters, _lib.FilHeaderC, _lib.ScanIO, _lib.ScanResult)]
    assert got == want
