set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tools/fused_prof.py 3 > gpurun_out/fused_prof5.log 2>&1; echo "prof rc=$?"
cat gpurun_out/fused_prof5.log
timeout 600 python tools/diag_fused.py 4 > gpurun_out/diag5.log 2>&1; echo "diag rc=$?"
grep -E "===|rows:|FAILED" gpurun_out/diag5.log
timeout 900 python -m pytest tests/test_gpu_paths.py tests/test_gpu_parity.py -x -q -m gpu -k "paths or fused or rescale_boundary or column_pass_stages" 2>&1 | tail -15
