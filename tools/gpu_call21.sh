set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02f_bench_C2_n8.json 2> gpurun_out/r02f_bench_n8.err; echo "bench C2 n8 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --config C5 --steps 3 --warmup 3 --no-e2e > gpurun_out/r02f_bench_C5_n8.json 2> gpurun_out/r02f_bench_c5_n8.err; echo "bench C5 n8 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r02f_bench_C2_n4.json 2> gpurun_out/r02f_bench_n4.err; echo "bench C2 n4 rc=$?"
python - <<'PY'
import json
for n in ("C2_n8","C5_n8","C2_n4"):
    try:
        d=json.loads(open(f"gpurun_out/r02f_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "weak ms", round(d["ms_per_step"],2), "GB/s", round(d["value"],1), "rt", round(d["rt_factor"],1), "e2e", d["e2e"] and (round(d["e2e"]["rt_factor"],1), round(d["e2e"]["h2d_copy_GBps_per_gpu"],1)), "strong", d.get("strong") and {k: (round(v,3) if isinstance(v,float) else v) for k,v in d["strong"].items() if k!="note"})
    except Exception as e:
        print(n, "ERR", e)
PY
