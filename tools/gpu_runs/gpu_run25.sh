set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r25; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; tail -9 $O/bench.err
ls -la $O
