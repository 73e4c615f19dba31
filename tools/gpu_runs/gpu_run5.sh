set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r5; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "t2t or fold or deform or dcn or propagation" 2>&1 | tail -15 > $O/pytest_t2t.log; tail -5 $O/pytest_t2t.log
timeout 300 python tools/t2t_bench.py 64 fused > $O/t2t_bench.log 2>&1; cat $O/t2t_bench.log
timeout 300 python tools/t2t_bench_hq.py > $O/t2t_bench_hq.log 2>&1; cat $O/t2t_bench_hq.log
timeout 300 python tools/dcn_bench.py > $O/dcn_bench.log 2>&1; cat $O/dcn_bench.log
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; tail -8 $O/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:t2t_ffn_mid -s 3 -c 1 -o $O/ncu_ffn_mid python tools/t2t_bench.py 64 fused > $O/ncu_ffn_mid.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcn_kernel -s 3 -c 1 -o $O/ncu_dcn python tools/dcn_bench.py > $O/ncu_dcn.log 2>&1
ls -la $O
