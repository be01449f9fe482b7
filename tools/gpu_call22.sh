set -x
cd $GRAFT_REPO_ROOT
python tools/sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1 && timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/r02_memcheck.log 2>&1
echo "memcheck rc=$?"
tail -15 gpurun_out/r02_memcheck.log
cat gpurun_out/sanitize_plain.log
