set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r7; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; tail -8 $O/bench.err
E2F_NO_OVERLAP=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads > $O/bench_no_overlap.json 2> $O/bench_no_overlap.err; tail -3 $O/bench_no_overlap.err
ls -la $O
