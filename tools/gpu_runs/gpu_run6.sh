set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r6; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
timeout 300 python tools/t2t_bench.py 64 fused > $O/t2t_bench.log 2>&1; cat $O/t2t_bench.log
timeout 300 python tools/t2t_bench_hq.py >> $O/t2t_bench.log 2>&1; tail -1 $O/t2t_bench.log
timeout 300 python tools/dcn_bench.py > $O/dcn_bench.log 2>&1; cat $O/dcn_bench.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -8 $O/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2>> $O/bench.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file $O/launches_b1.csv python tools/profile_step.py --clips 1 > $O/prof_b1.log 2>&1
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file $O/launches_b8.csv python tools/profile_step.py --clips 8 > $O/prof_b8.log 2>&1
for spec in "conv3x3_kernel 45 3 conv" "conv3x3_halo 0 2 halo" "linear_kernel 2 3 linear" "focal_attn 0 1 attn" "dcn_kernel 0 1 dcn" "t2t_fold733 0 1 ffnmid" "layernorm_pool 0 1 lnpool" "prop_prologue 0 1 prologue"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$1 -s $2 -c $3 -o $O/ncu_full_$4 python tools/profile_step.py --clips 8 > $O/ncu_full_$4.log 2>&1
done
ls -la $O
