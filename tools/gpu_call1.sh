set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python tools/diag_fused.py > gpurun_out/diag1.log 2>&1; echo "diag rc=$?"
tail -60 gpurun_out/diag1.log
B2F_PATH=fused timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seconds 20 > gpurun_out/bench_fused1.json 2> gpurun_out/bench_fused1.err; echo "bench fused rc=$?"
B2F_PATH=split timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seconds 20 > gpurun_out/bench_split1.json 2> gpurun_out/bench_split1.err; echo "bench split rc=$?"
B2F_PATH=legacy timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seconds 20 > gpurun_out/bench_legacy1.json 2> gpurun_out/bench_legacy1.err; echo "bench legacy rc=$?"
python - <<'PY'
import json
for n in ("fused","split","legacy"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{n}1.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],2), "rt", round(d["rt_factor"],1), d["kernel_ms_per_step"], d["clocks"])
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/bench_{n}1.err").read()[-1500:])
PY
