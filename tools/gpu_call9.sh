set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
for c in C1 C2 C3 C4 C5; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err; echo "bench $c rc=$?"
done
for c in C1 C2 C3; do
  B2F_PATH=fused timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-e2e > gpurun_out/r02_bench_${c}_fused.json 2> gpurun_out/r02_bench_${c}_fused.err; echo "bench $c fused rc=$?"
done
B2F_PATH=split timeout 600 python bench.py --config C2 --steps 3 --warmup 3 --no-e2e > gpurun_out/r02_bench_C2_split.json 2> gpurun_out/r02_bench_C2_split.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_")[1], "ms/step", round(d["ms_per_step"],2), "GB/s", round(d["value"],1), "rt", round(d["rt_factor"],1), "e2e", d["e2e"] and round(d["e2e"]["rt_factor"],1), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["parity_check"].get("ok"), d["parity_check"].get("max_rel"), "roof", round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-800:])
PY
B2F_PATH=fused python tools/prof_run.py 2 > gpurun_out/plain_fused9.log 2>&1 &&
B2F_PATH=fused ncu --set full --clock-control none --import-source on -k regex:"kf_fused" -c 1 -o gpurun_out/r02_fused_final python tools/prof_run.py 2 > gpurun_out/ncu_fused9.log 2>&1
echo "ncu fused rc=$?"
B2F_PATH=split python tools/prof_run.py 2 > gpurun_out/plain_split9.log 2>&1 &&
B2F_PATH=split ncu --set full --clock-control none --import-source on -k regex:"kf_fused|kr_row_pass|k0t_transpose" -c 3 -o gpurun_out/r02_split_final python tools/prof_run.py 2 > gpurun_out/ncu_split9.log 2>&1
echo "ncu split rc=$?"
python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/plain_bench9.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/ncu_launch9.log 2>&1
echo "ncu launches rc=$?"
