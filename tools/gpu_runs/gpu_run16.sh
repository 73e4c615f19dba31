set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r16; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 120 python tools/t2t_bench.py 64 fused > $O/t2t.log 2>&1
timeout 120 python tools/t2t_bench.py 8 fused >> $O/t2t.log 2>&1; cat $O/t2t.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err
ls -la $O
