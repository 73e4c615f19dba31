set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r23; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $R --master-port 29571 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-extra-workloads > $O/bench_8gpu_peer.json 2> $O/bench_8gpu_peer.err; tail -3 $O/bench_8gpu_peer.err; tail -c 400 $O/bench_8gpu_peer.json
E2F_STITCH=nccl timeout 300 $R --master-port 29572 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-extra-workloads > $O/bench_8gpu_nccl.json 2> $O/bench_8gpu_nccl.err; tail -3 $O/bench_8gpu_nccl.err
ls -la $O
