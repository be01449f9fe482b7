set -x
cd $GRAFT_REPO_ROOT
B2F_PATH=split python tools/prof_run.py 2 > gpurun_out/plain_split15.log 2>&1 &&
B2F_PATH=split ncu --set full --clock-control none --import-source on -k regex:"kt_row_tiles" -c 1 -o gpurun_out/r02_kt python tools/prof_run.py 2 > gpurun_out/ncu_kt15.log 2>&1
echo "ncu rc=$?"
