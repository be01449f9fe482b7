set -x
cd $GRAFT_REPO_ROOT
for cf in 1024 2048 4096; do B2F_DEDISP_CHUNK_FRAMES=$cf python tools/bench_configs.py "C4 (" 2>/dev/null | tail -1; done > gpurun_out/r02_c4_chunk_sweep.jsonl
cat gpurun_out/r02_c4_chunk_sweep.jsonl
python bench.py --config P1 --steps 3 --warmup 3 > gpurun_out/r02_bench_P1_n1.json 2> gpurun_out/bench_p1.err; tail -c 2500 gpurun_out/r02_bench_P1_n1.json; tail -5 gpurun_out/bench_p1.err
