set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench18_C2.json 2> gpurun_out/bench18_C2.err; echo "bench rc=$?"
B2F_PATH=legacy timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e > gpurun_out/bench18_C2_legacy.json 2> gpurun_out/bench18_C2_legacy.err
for c in C3 C5; do timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-e2e > gpurun_out/bench18_$c.json 2> gpurun_out/bench18_$c.err; done
python - <<'PY'
import json
for n in ("C2","C2_legacy","C3","C5"):
    try:
        d=json.loads(open(f"gpurun_out/bench18_{n}.json").read().strip().splitlines()[-1])
        print(n, d["config"]["channeliser_path"], "ms/step", round(d["ms_per_step"],2), "GB/s", round(d["value"],1), "rt", round(d["rt_factor"],1), "e2e", d["e2e"] and round(d["e2e"]["rt_factor"],1), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["parity_check"]["ok"], d["parity_check"].get("max_rel"), "roofline", round(d["roofline"]["frac"],3), round(d["roofline"]["whole_step"]["frac"],3))
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/bench18_{n}.err").read()[-1500:])
PY
python tools/prof_run.py 2 > gpurun_out/plain18.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"kt_row_tiles|kf_fused|k0t_transpose" -c 3 -o gpurun_out/r02_default_final python tools/prof_run.py 2 > gpurun_out/ncu18.log 2>&1
echo "ncu rc=$?"
