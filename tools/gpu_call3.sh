set -x
cd $GRAFT_REPO_ROOT
timeout 600 python tools/diag_fused.py 4 > gpurun_out/diag3.log 2>&1; echo "diag rc=$?"
grep -E "===|rows:|counters|FAILED" gpurun_out/diag3.log
for v in fused split; do
  B2F_PATH=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seconds 20 > gpurun_out/bench_${v}3.json 2> gpurun_out/bench_${v}3.err; echo "bench $v rc=$?"
  B2F_LIB=$PWD/frb-baseband_b200/libb2f_t8.so B2F_PATH=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seconds 20 > gpurun_out/bench_${v}3_t8.json 2> gpurun_out/bench_${v}3_t8.err; echo "bench $v t8 rc=$?"
done
B2F_RING_SLOTS=3 B2F_PATH=fused timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seconds 20 > gpurun_out/bench_fused3_ns3.json 2> gpurun_out/bench_fused3_ns3.err
python - <<'PY'
import json
for n in ("fused3","fused3_t8","split3","split3_t8","fused3_ns3"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],2), "rt", round(d["rt_factor"],1), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["clocks"]["sm_mhz"])
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/bench_{n}.err").read()[-1500:])
PY
