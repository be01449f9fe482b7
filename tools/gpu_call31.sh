set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_paths.py tests/test_gpu_knobs.py tests/test_gpu_parity.py -x -q -m gpu -k "not generic and not tscrunch and not property and not dedisp" 2>&1 | tail -5
python tools/bench_configs.py "C2 (" "C1 (4" > gpurun_out/r02_k0t_pairstores.jsonl 2> gpurun_out/r02_k0t.err
cat gpurun_out/r02_k0t_pairstores.jsonl; tail -3 gpurun_out/r02_k0t.err
