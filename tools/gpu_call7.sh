set -x
cd $GRAFT_REPO_ROOT
timeout 600 python tools/diag_fused.py 4 > gpurun_out/diag7.log 2>&1; echo "diag rc=$?"
grep -E "===|rows:|FAILED" gpurun_out/diag7.log
for v in split legacy fused; do
  B2F_PATH=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seconds 20 > gpurun_out/bench7_$v.json 2> gpurun_out/bench7_$v.err; echo "bench $v rc=$?"
done
python - <<'PY'
import json
for n in ("split","legacy","fused"):
    try:
        d=json.loads(open(f"gpurun_out/bench7_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],2), "rt", round(d["rt_factor"],1), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["clocks"]["sm_mhz"], d["parity_check"]["ok"], "roofline", round(d["roofline"]["frac"],3))
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/bench7_{n}.err").read()[-1500:])
PY
B2F_PATH=split timeout 900 python -m pytest tests/test_gpu_paths.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
