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
