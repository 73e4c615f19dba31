set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r4; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -8 $O/bench.err; head -c 300 $O/bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:t2t_fold733 -s 3 -c 1 -o $O/ncu_fold733_fused python tools/t2t_bench.py 64 fused > $O/ncu_fold733.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:t2t_fold733 -s 3 -c 1 -o $O/ncu_fold733_fused_hq python tools/t2t_bench_hq.py > $O/ncu_fold733_hq.log 2>&1
ls -la $O
