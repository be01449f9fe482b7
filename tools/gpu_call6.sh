set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tools/fused_prof.py 3 > gpurun_out/fused_prof6.log 2>&1; echo "prof rc=$?"
cat gpurun_out/fused_prof6.log
timeout 600 python tools/diag_fused.py 4 > gpurun_out/diag6.log 2>&1; echo "diag rc=$?"
grep -E "===|rows:|FAILED" gpurun_out/diag6.log
for ns in 3 4; do
  B2F_RING_SLOTS=$ns B2F_PATH=fused timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seconds 20 > gpurun_out/bench6_fused_ns$ns.json 2> gpurun_out/bench6_fused_ns$ns.err; echo "bench fused ns $ns rc=$?"
done
B2F_ROW_LAG=3 B2F_PATH=fused timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seconds 20 > gpurun_out/bench6_fused_lag3.json 2> gpurun_out/bench6_fused_lag3.err
python - <<'PY'
import json
for n in ("fused_ns3","fused_ns4","fused_lag3"):
    try:
        d=json.loads(open(f"gpurun_out/bench6_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],2), "rt", round(d["rt_factor"],1), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["clocks"]["sm_mhz"], d["parity_check"]["ok"], "roofline", round(d["roofline"]["frac"],3))
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/bench6_{n}.err").read()[-1500:])
PY
timeout 900 python -m pytest tests/test_gpu_paths.py -x -q -m gpu 2>&1 | tail -5
