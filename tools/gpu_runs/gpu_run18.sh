set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r18; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -15 > $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 120 python tools/dcn_bench.py > $O/dcn_bench.log 2>&1; cat $O/dcn_bench.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-extra-workloads > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dcn_kernel -s 3 -c 1 -o $O/ncu_full_dcn python tools/dcn_bench.py > $O/ncu_full_dcn.log 2>&1
ls -la $O
