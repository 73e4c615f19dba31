set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r14; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
for i in 1 2; do timeout 120 python tools/t2t_bench.py 64 fused; done > $O/t2t_lb3.log 2>&1; cat $O/t2t_lb3.log
timeout 120 python tools/t2t_bench.py 8 fused >> $O/t2t_lb3.log 2>&1
cp e2fgvi_b200/libe2fgvi_b200.so /tmp/main.so; cp e2fgvi_b200/alt_lb2.so e2fgvi_b200/libe2fgvi_b200.so
for i in 1 2; do timeout 120 python tools/t2t_bench.py 64 fused; done > $O/t2t_lb2.log 2>&1; cat $O/t2t_lb2.log
timeout 120 python tools/t2t_bench.py 8 fused >> $O/t2t_lb2.log 2>&1
cp /tmp/main.so e2fgvi_b200/libe2fgvi_b200.so
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra-workloads > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:t2t_fold733 -s 3 -c 1 -o $O/ncu_full_ffnmid python tools/t2t_bench.py 64 fused > $O/ncu_full_ffnmid.log 2>&1
ls -la $O
