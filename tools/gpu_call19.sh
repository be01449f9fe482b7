set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_paths.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e > gpurun_out/bench19_C2.json 2> gpurun_out/bench19_C2.err; echo "bench rc=$?"
timeout 600 python bench.py --config C3 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench19_C3.json 2> gpurun_out/bench19_C3.err
python - <<'PY'
import json
for n in ("C2","C3"):
    try:
        d=json.loads(open(f"gpurun_out/bench19_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],2), "GB/s", round(d["value"],1), "rt", round(d["rt_factor"],1), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["parity_check"]["ok"], "roofline", round(d["roofline"]["frac"],3), round(d["roofline"]["whole_step"]["frac"],3), d["clocks"])
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/bench19_{n}.err").read()[-1500:])
PY
