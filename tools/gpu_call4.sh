set -x
cd $GRAFT_REPO_ROOT
timeout 600 python tools/diag_fused.py 4 > gpurun_out/diag4.log 2>&1; echo "diag rc=$?"
grep -E "===|rows:|FAILED" gpurun_out/diag4.log
for lag in 2 1; do
  B2F_ROW_LAG=$lag B2F_PATH=fused timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seconds 20 > gpurun_out/bench4_fused_lag$lag.json 2> gpurun_out/bench4_fused_lag$lag.err; echo "bench fused lag $lag rc=$?"
done
B2F_PATH=split timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seconds 20 > gpurun_out/bench4_split.json 2> gpurun_out/bench4_split.err
python - <<'PY'
import json
for n in ("fused_lag2","fused_lag1","split"):
    try:
        d=json.loads(open(f"gpurun_out/bench4_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],2), "rt", round(d["rt_factor"],1), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["clocks"]["sm_mhz"], d["parity_check"], "roofline", round(d["roofline"]["frac"],3))
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/bench4_{n}.err").read()[-1500:])
PY
timeout 900 python -m pytest tests/test_gpu_paths.py tests/test_gpu_parity.py -x -q -m gpu -k "paths or fused or rescale_boundary or column_pass_stages or float_spectra or requantised" 2>&1 | tail -15
