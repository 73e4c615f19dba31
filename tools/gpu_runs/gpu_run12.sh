set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r12; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
timeout 300 python tools/conv_bench.py > $O/conv_bench.log 2>&1; tail -32 $O/conv_bench.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; tail -8 $O/bench.err
ls -la $O
