set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "generic or unsupported" 2>&1 | tail -15
python tools/bench_configs.py generic > gpurun_out/r02_generic_after.jsonl 2> gpurun_out/r02_generic_after.err
cat gpurun_out/r02_generic_after.jsonl; tail -3 gpurun_out/r02_generic_after.err
python tools/prof_generic.py 1024 2 8 6 > gpurun_out/prof_generic_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kgt_ -c 2 -o gpurun_out/r02_generic_kgt2 python tools/prof_generic.py 1024 2 8 6 > gpurun_out/ncu_generic.log 2>&1
tail -n 3 gpurun_out/prof_generic_plain.log gpurun_out/ncu_generic.log
