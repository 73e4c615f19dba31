set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/pytest_gpu.log; tail -12 $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -6 $O/bench.err; head -c 300 $O/bench.json
timeout 300 python tools/conv_bench.py > $O/conv_bench.log 2>&1; tail -2 $O/conv_bench.log
timeout 300 python tools/linear_bench.py > $O/linear_bench.log 2>&1; cat $O/linear_bench.log
timeout 300 python tools/torch_ops_trace.py > $O/torch_ops_trace.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file $O/launches_b1.csv python tools/profile_step.py --clips 1 > $O/prof_b1.log 2>&1
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file $O/launches_b8.csv python tools/profile_step.py --clips 8 > $O/prof_b8.log 2>&1
ls -la $O
