set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_dropin.py -x -q -m gpu -k "peer_splice or one_plan_two_scans" 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/h2d_topology.py > gpurun_out/r02_h2d_topology_n2.json 2> gpurun_out/h2d_n2.err; echo "h2d n2 rc=$?"
cat gpurun_out/r02_h2d_topology_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_C2_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_C2_n2.json").read().strip().splitlines()[-1])
    print("n2 weak ms", round(d["ms_per_step"],2), "rt", round(d["rt_factor"],1), "e2e rt", round(d["e2e"]["rt_factor"],1), "strong", d.get("strong"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_n2.err").read()[-2000:])
PY
for r in 1 2 4 8; do timeout 300 python tools/bench_runner.py --seconds 20 --readers $r >> gpurun_out/r02_runner_readers.jsonl 2>> gpurun_out/runner.err; done
cat gpurun_out/r02_runner_readers.jsonl
