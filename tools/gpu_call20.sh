set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for c in C2 C1 C3 C4 C5; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/r02f_bench_$c.json 2> gpurun_out/r02f_bench_$c.err; echo "bench $c rc=$?"
done
for v in legacy fused; do
  B2F_PATH=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seconds 20 > gpurun_out/r02f_bench_C2_20s_$v.json 2> gpurun_out/r02f_bench_C2_20s_$v.err
done
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seconds 20 > gpurun_out/r02f_bench_C2_20s_default.json 2> gpurun_out/r02f_bench_C2_20s_default.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02f_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_")[1], "ms/step", round(d["ms_per_step"],2), "GB/s", round(d["value"],1), "rt", round(d["rt_factor"],1), "e2e", d["e2e"] and round(d["e2e"]["rt_factor"],1), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["parity_check"].get("ok"), d["parity_check"].get("max_rel"), "roof", round(d["roofline"]["frac"],3), round(d["roofline"]["whole_step"]["frac"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "cpu", d.get("cpu_baseline",{}).get("value"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-800:])
PY
python tools/prof_run.py 2 > gpurun_out/plain20.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"kt_row_tiles|kf_fused|k0t_transpose" -c 3 -o gpurun_out/r02_default_final python tools/prof_run.py 2 > gpurun_out/ncu20.log 2>&1
echo "ncu rc=$?"
python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/plain_bench20.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"b2f::" -c 800 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/ncu_launch20.log 2>&1
echo "ncu launches rc=$?"
tail -1 gpurun_out/plain_bench20.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d['ms_per_step'])"
