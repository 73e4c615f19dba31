set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r20; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -k "two_devices" 2>&1 | tail -5 > $O/pytest_two_devices.log; cat $O/pytest_two_devices.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_4gpu.json 2> $O/bench_4gpu.err; tail -12 $O/bench_4gpu.err; tail -c 600 $O/bench_4gpu.json
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads > $O/bench_1gpu.json 2> $O/bench_1gpu.err; tail -2 $O/bench_1gpu.err
ls -la $O
