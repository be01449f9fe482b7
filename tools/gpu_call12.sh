set -x
cd $GRAFT_REPO_ROOT
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" > gpurun_out/r02_lscpu.txt 2>&1
for n in 8 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/h2d_topology.py > gpurun_out/r02_h2d_topology_n$n.json 2> gpurun_out/h2d_n$n.err; echo "h2d n$n rc=$?"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_bench_C2_n8.json 2> gpurun_out/bench_n8.err; echo "bench C2 n8 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --config C5 --steps 3 --warmup 3 --no-e2e > gpurun_out/r02_bench_C5_n8.json 2> gpurun_out/bench_c5_n8.err; echo "bench C5 n8 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --config C4 --steps 2 --warmup 3 --no-e2e --seconds 20 > gpurun_out/r02_bench_C4_n8.json 2> gpurun_out/bench_c4_n8.err; echo "bench C4 n8 rc=$?"
python - <<'PY'
import json
for n in ("C2_n8","C5_n8","C4_n8"):
    try:
        d=json.loads(open(f"gpurun_out/r02_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "weak ms", round(d["ms_per_step"],2), "GB/s", round(d["value"],1), "rt", round(d["rt_factor"],1), "e2e", d["e2e"] and (round(d["e2e"]["rt_factor"],1), round(d["e2e"]["h2d_copy_GBps_per_gpu"],1)), "strong", d.get("strong"))
    except Exception as e:
        print(n, "ERR", e)
for n in (8,4):
    try:
        d=json.loads(open(f"gpurun_out/r02_h2d_topology_n{n}.json").read().strip().splitlines()[-1])
        print("h2d", n, d["aggregate_GBps"], [(r["rank"], r["numa_node"], round(r["unbound"],1), round(r.get("bound_to_gpu_numa_node",0),1)) for r in d["per_rank"]], d["host"])
    except Exception as e:
        print("h2d", n, "ERR", e)
PY
cat gpurun_out/r02_lscpu.txt
