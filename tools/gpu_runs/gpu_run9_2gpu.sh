set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r9; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_2gpu.json 2> $O/bench_2gpu.err; tail -12 $O/bench_2gpu.err; head -c 400 $O/bench_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $O/bench_2gpu_ref.json 2> $O/bench_2gpu_ref.err; head -c 300 $O/bench_2gpu_ref.json
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads > $O/bench_1gpu.json 2> $O/bench_1gpu.err; tail -2 $O/bench_1gpu.err
ls -la $O
