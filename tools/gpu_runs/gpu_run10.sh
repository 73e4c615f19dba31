set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r10; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "kxn" 2>&1 | tail -25 > $O/pytest_kxn.log; tail -6 $O/pytest_kxn.log
timeout 300 python tools/kxn_bench.py > $O/kxn_bench.log 2>&1; cat $O/kxn_bench.log
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; tail -8 $O/bench.err
E2F_KXN=0 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads > $O/bench_no_kxn.json 2> $O/bench_no_kxn.err; tail -3 $O/bench_no_kxn.err
ls -la $O
