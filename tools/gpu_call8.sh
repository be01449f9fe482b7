set -x
cd $GRAFT_REPO_ROOT
for v in split legacy fused; do
  B2F_PATH=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seconds 20 > gpurun_out/bench8_$v.json 2> gpurun_out/bench8_$v.err; echo "bench $v rc=$?"
done
python - <<'PY'
import json
for n in ("split","legacy","fused"):
    try:
        d=json.loads(open(f"gpurun_out/bench8_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],2), "rt", round(d["rt_factor"],1), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["clocks"]["sm_mhz"], d["parity_check"]["ok"], "roofline", round(d["roofline"]["frac"],3))
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/bench8_{n}.err").read()[-1500:])
PY
timeout 900 python -m pytest tests/test_gpu_knobs.py tests/test_gpu_paths.py -x -q -m gpu 2>&1 | tail -15
B2F_PATH=split timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "column_pass_stages or float_spectra or pol_modes or other_row" 2>&1 | tail -5
