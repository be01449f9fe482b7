set -x
cd $GRAFT_REPO_ROOT
python tools/bench_configs.py generic > gpurun_out/r02_generic_before.jsonl 2> gpurun_out/r02_generic_before.err
cat gpurun_out/r02_generic_before.jsonl; tail -3 gpurun_out/r02_generic_before.err
python tools/prof_generic.py 1024 2 8 6 > gpurun_out/prof_generic_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kg_ -c 4 -o gpurun_out/r02_generic_before python tools/prof_generic.py 1024 2 8 6 > gpurun_out/ncu_generic.log 2>&1
tail -3 gpurun_out/prof_generic_plain.log gpurun_out/ncu_generic.log
