set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r24; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/nvidia_smi.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -9 $O/bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>> $O/bench.err
timeout 300 python tools/conv_bench.py > $O/conv_bench.log 2>&1
timeout 300 python tools/linear_bench.py > $O/linear_bench.log 2>&1
timeout 300 python tools/kxn_bench.py > $O/kxn_bench.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file $O/launches_b1.csv python tools/profile_step.py --clips 1 > $O/prof_b1.log 2>&1
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file $O/launches_b8.csv python tools/profile_step.py --clips 8 > $O/prof_b8.log 2>&1
for spec in "dcn_kernel 0 1 dcn" "t2t_fold733 0 1 ffnmid" "upsample2x 0 2 upsample" "prop_prologue 0 1 prologue"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$1 -s $2 -c $3 -o $O/ncu_full_$4 python tools/profile_step.py --clips 8 > $O/ncu_full_$4.log 2>&1
done
timeout 200 python tools/t2t_bench.py 64 fused > $O/t2t_bench.log 2>&1
timeout 200 python tools/dcn_bench.py > $O/dcn_bench.log 2>&1
ls -la $O
