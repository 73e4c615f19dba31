set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1/pytest_gpu.log; cat gpurun_out/r1/pytest_gpu.log | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1/bench.json 2> gpurun_out/r1/bench.err; tail -3 gpurun_out/r1/bench.err; head -c 600 gpurun_out/r1/bench.json
timeout 300 python tools/conv_bench.py > gpurun_out/r1/conv_bench.log 2>&1; tail -3 gpurun_out/r1/conv_bench.log
timeout 300 python tools/linear_bench.py > gpurun_out/r1/linear_bench.log 2>&1
timeout 300 python tools/torch_ops_trace.py > gpurun_out/r1/torch_ops_trace.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r1/launches_b1.csv python tools/profile_step.py --clips 1 > gpurun_out/r1/prof_b1.log 2>&1
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r1/launches_b8.csv python tools/profile_step.py --clips 8 > gpurun_out/r1/prof_b8.log 2>&1
ls -la gpurun_out/r1
