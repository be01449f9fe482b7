set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_paths.py tests/test_gpu_knobs.py -x -q -m gpu 2>&1 | tail -6
B2F_PATH=split timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "column_pass_stages or float_spectra or pol_modes or multi_if or requantised" 2>&1 | tail -4
for v in split legacy fused; do
  B2F_PATH=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seconds 20 > gpurun_out/bench14_$v.json 2> gpurun_out/bench14_$v.err; echo "bench $v rc=$?"
done
B2F_NO_TILE_ROWS=1 B2F_PATH=split timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seconds 20 > gpurun_out/bench14_split_kr.json 2> gpurun_out/bench14_split_kr.err
python - <<'PY'
import json
for n in ("split","split_kr","legacy","fused"):
    try:
        d=json.loads(open(f"gpurun_out/bench14_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],2), "rt", round(d["rt_factor"],1), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["parity_check"]["ok"], "roofline", round(d["roofline"]["frac"],3))
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/bench14_{n}.err").read()[-1500:])
PY
