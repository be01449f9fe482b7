set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_knobs.py tests/test_gpu_paths.py -x -q -m gpu 2>&1 | tail -4
for lag in 2 1; do
  B2F_ROW_LAG=$lag B2F_PATH=fused timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seconds 20 > gpurun_out/bench10_fused_lag$lag.json 2> gpurun_out/bench10_fused_lag$lag.err
done
B2F_PATH=split timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seconds 20 > gpurun_out/bench10_split.json 2> gpurun_out/bench10_split.err
timeout 300 python bench.py --config C4 --steps 2 --warmup 3 --no-e2e --seconds 20 > gpurun_out/bench10_C4.json 2> gpurun_out/bench10_C4.err
python - <<'PY'
import json
for n in ("fused_lag2","fused_lag1","split","C4"):
    try:
        d=json.loads(open(f"gpurun_out/bench10_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],2), "rt", round(d["rt_factor"],1), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["parity_check"], "roofline", round(d["roofline"]["frac"],3))
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/bench10_{n}.err").read()[-1500:])
PY
B2F_PATH=fused python tools/prof_run.py 2 > gpurun_out/plain_fused10.log 2>&1 &&
B2F_PATH=fused ncu --set full --clock-control none --import-source on -k regex:"kf_fused" -c 1 -o gpurun_out/r02_fused_final python tools/prof_run.py 2 > gpurun_out/ncu_fused10.log 2>&1
echo "ncu fused rc=$?"
B2F_ROW_LAG=1 B2F_PATH=fused ncu --set full --clock-control none -k regex:"kf_fused" -c 1 -o gpurun_out/r02_fused_lag1 python tools/prof_run.py 2 > gpurun_out/ncu_fused10b.log 2>&1
B2F_PATH=split ncu --set full --clock-control none --import-source on -k regex:"kf_fused|kr_row_pass|k0t_transpose" -c 3 -o gpurun_out/r02_split_final python tools/prof_run.py 2 > gpurun_out/ncu_split10.log 2>&1
echo "ncu split rc=$?"
