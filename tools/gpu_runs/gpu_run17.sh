set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r17; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_pdl.json 2> $O/bench_pdl.err; tail -3 $O/bench_pdl.err
E2F_NO_PDL=1 timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_nopdl.json 2> $O/bench_nopdl.err; tail -3 $O/bench_nopdl.err
timeout 900 python bench.py --steps 20 --warmup 3 --no-extra-workloads > $O/bench_pdl2.json 2> $O/bench_pdl2.err
ls -la $O
