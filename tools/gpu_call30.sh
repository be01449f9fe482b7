set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "generic" 2>&1 | tail -5
python tools/bench_configs.py generic > gpurun_out/r02_generic_after.jsonl 2> gpurun_out/r02_generic_after.err
cat gpurun_out/r02_generic_after.jsonl; tail -3 gpurun_out/r02_generic_after.err
