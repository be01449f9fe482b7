set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "compile_time or any_tscrunch or unsupported" 2>&1 | tail -15
