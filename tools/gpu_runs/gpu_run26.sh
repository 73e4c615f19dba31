set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r26; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
ls -la $O
