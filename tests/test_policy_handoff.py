"""Job policy and downstream hand-off (N3) against fixtures generated from the reference's own
Python (tests/golden/make_golden.py imports /root/reference/submit_job.py)."""
import json
import os

import pytest

from frb_baseband_b200 import handoff, job_policy

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "submit_job_reference.json")))["submit_job"]


@pytest.mark.parametrize("g", GOLD, ids=lambda g: f"dm{g['dm']}-p{g['period_s']}-{g['nIF']}x{g['IF']}")
def test_chooser_matches_reference(g):
    if g["period_s"] is None:
        p = job_policy.search_plan(g["dm"], g["fref"], g["IF"], g["nIF"])
    else:
        p = job_policy.pulsar_plan(g["dm"], g["period_s"], g["fref"], g["IF"], g["nIF"])
    assert (p.nchan_if, p.tscrunch) == (g["nchan_if"], g["tscrunch"])
    assert p.flag_file("ef") == g["flag_file"]
    argv = job_policy.create_config_argv(p, vex="/vex/ek048c.vix", source="SRC", telescope="ef", scan="012",
                                         config_file=g["config_file"], search=g["period_s"] is None)
    assert argv == g["first_argv"] + g["tail"]


def test_policy_properties():
    # time resolution after tscrunch never exceeds the wanted one, and the smearing rule is monotonic in DM
    last = 0
    for dm in (50, 100, 200, 400, 800, 1600):
        p = job_policy.search_plan(dm, 1286.0, 32.0, 8)
        assert p.nchan_if >= last and p.tscrunch * p.nchan_if / 32.0 <= p.t_res_us
        last = p.nchan_if
    g = job_policy.gpu_plan(560.0, 1286.0, 32.0, 8)
    assert g.coherent and (g.nchan_if, g.tscrunch) == (128, 16)          # BASELINE config 2/4: 64 us at 128 channels
    assert job_policy.gpu_plan(None, 1286.0, 32.0, 8) == job_policy.search_plan(None, 1286.0, 32.0, 8)


def test_handoff_strings(tmp_path):
    assert handoff.fetch_message("/d/x.fil", "") == "/d/x.fil "            # base2fil.sh:121 with an empty flag file
    assert handoff.fetch_argv("/d/x.fil", "/f/flag")[2:] == ["-q", "stage01_queue", "-m", "/d/x.fil /f/flag"]
    assert handoff.target_name("B0329+54 --ra 03:32:59 --dec +54:34:43") == "B0329+54"
    assert handoff.target_name("R3") == "R3" and handoff.fold_commands("BSGR", "Ef", "x.fil", 2) == []
    c = handoff.fold_commands("B0329+54 --ra 1", "Ef", "/o/a.fil", 4)
    assert c[0] == "dspsr -E B0329+54.psrcat.par -L 10 -A -k Ef -d1 /o/a.fil -O /o/a.fil -t 8"          # base2fil.sh:474
    assert c[2] == "mv pgplot.ps /o/a.fil.ps" and c[-1] == "mv pgplot.ps /o/a.fil_fullPol.ps" and len(c) == 5
    assert len(handoff.fold_commands("B0329+54", "Ef", "/o/a.fil", 2)) == 3
    assert handoff.split_vdif_globs("ek", "ef", "003", "/s0/ek", "/s1/ek") == ["/s1/ek/ek_ef_no0003_IF*.vdif",
                                                                               "/s0/ek/ek_ef_no0003_IF*.vdif"]


def test_after_scan_order_and_failure(tmp_path):
    fil = tmp_path / "a.fil"
    v = [tmp_path / f"e_ef_no0001_IF{i}.vdif" for i in (1, 2)]
    for f in v:
        f.write_bytes(b"x")
    with pytest.raises(FileNotFoundError):                                   # nothing is removed if the splice failed
        handoff.after_scan(str(fil), keep_vdif=False, vdif_globs=[str(tmp_path / "*.vdif")])
    assert all(f.exists() for f in v)
    fil.write_bytes(b"fil")
    sent = []
    r = handoff.after_scan(str(fil), flag_file="/f", submit2fetch=True, keep_vdif=True, send=lambda a: sent.append(a) or 0,
                           vdif_globs=[str(tmp_path / "*.vdif")], log=lambda m: None)
    assert all(f.exists() for f in v) and r["submitted"] == f"{fil} /f" and len(sent) == 1
    with pytest.raises(RuntimeError):
        handoff.after_scan(str(fil), submit2fetch=True, keep_vdif=False, send=lambda a: 1,
                           vdif_globs=[str(tmp_path / "*.vdif")], log=lambda m: None)
    assert not any(f.exists() for f in v)                                    # removal precedes the submit, as in the shell


def test_dm_catalogue_matches_reference(tmp_path):
    """source_dm.get_dm == the table of the reference's dm_utils.get_dm (fixture: tests/golden/dm_utils_reference.json,
    extracted from /root/reference/dm_utils.py:12-93), same int/float types; R1 absent; unknown -> None."""
    from frb_baseband_b200 import sigproc, source_dm
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "dm_utils_reference.json")))["table"]
    assert source_dm.FRB_DMS == gold and all(type(source_dm.FRB_DMS[k]) is type(v) for k, v in gold.items())
    assert source_dm.get_dm("R3") == 349.7 and source_dm.get_dm("R1", psrcat="/nonexistent/psrcat") is None
    assert source_dm.isPulsar is False
    fake = tmp_path / "psrcat"                                  # a stand-in psrcat that knows one pulsar
    fake.write_text("#!/bin/sh\n[ \"$7\" = B0329+54 ] && echo ' 26.7641 ' || exit 1\n")
    fake.chmod(0o755)
    assert source_dm.get_dm("B0329+54", psrcat=str(fake)) == pytest.approx(26.7641) and source_dm.isPulsar is True
    assert source_dm.get_dm("J0000+0000", psrcat=str(fake)) is None
    fil = tmp_path / "x.fil"
    fil.write_bytes(sigproc.FilHeader(source_name="R3", nchans=1024, nbits=8, nifs=1).pack() + bytes(2048))
    assert source_dm.get_src(str(fil)) == "R3" and source_dm.get_nchan(str(fil)) == 1024          # dm_utils.py:108-125
