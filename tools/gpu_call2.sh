set -x
cd $GRAFT_REPO_ROOT
B2F_PATH=split python tools/prof_run.py 2 > gpurun_out/plain_split.log 2>&1 &&
B2F_PATH=split ncu --set full --clock-control none --import-source on -k regex:"kf_fused|k0t_transpose|kt_sum" -c 4 -o gpurun_out/r02_split python tools/prof_run.py 2 > gpurun_out/ncu_split.log 2>&1
echo "ncu split rc=$?"
B2F_PATH=fused python tools/prof_run.py 2 > gpurun_out/plain_fused.log 2>&1 &&
B2F_PATH=fused ncu --set full --clock-control none --import-source on -k regex:"kf_fused" -c 1 -o gpurun_out/r02_fused python tools/prof_run.py 2 > gpurun_out/ncu_fused.log 2>&1
echo "ncu fused rc=$?"
tail -5 gpurun_out/ncu_split.log gpurun_out/ncu_fused.log
ls -la gpurun_out/*.ncu-rep
