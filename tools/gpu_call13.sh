set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_abi.py -x -q -m gpu 2>&1 | tail -4
for r in 2 4; do timeout 300 python tools/bench_runner.py --seconds 20 --readers $r >> gpurun_out/r02_runner_after.jsonl 2>> gpurun_out/runner13.err; done
timeout 300 python tools/bench_runner.py --seconds 60 --readers 4 >> gpurun_out/r02_runner_after.jsonl 2>> gpurun_out/runner13.err
cat gpurun_out/r02_runner_after.jsonl
python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/plain_bench13.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"b2f::" -c 700 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/ncu_launch13.log 2>&1
echo "ncu launches rc=$?"
tail -3 gpurun_out/plain_bench13.log | cut -c1-600
