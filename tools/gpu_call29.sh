set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_parity.py::test_generic_compile_time_geometry --deselect tests/test_gpu_parity.py::test_any_tscrunch 2>&1 | tail -5
python bench.py --config C4 --steps 3 --warmup 3 > gpurun_out/r02_bench_C4_n1_v2.json 2> gpurun_out/bench_c4.err; tail -c 1500 gpurun_out/r02_bench_C4_n1_v2.json
