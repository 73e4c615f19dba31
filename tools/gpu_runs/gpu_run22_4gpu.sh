set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r22; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $R --master-port 29561 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads > $O/bench_4gpu_peer.json 2> $O/bench_4gpu_peer.err; tail -3 $O/bench_4gpu_peer.err
E2F_STITCH=nccl timeout 600 $R --master-port 29562 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads > $O/bench_4gpu_nccl.json 2> $O/bench_4gpu_nccl.err; tail -3 $O/bench_4gpu_nccl.err
timeout 600 $R --master-port 29563 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads > $O/bench_4gpu_peer2.json 2> $O/bench_4gpu_peer2.err; tail -3 $O/bench_4gpu_peer2.err
E2F_STITCH=nccl timeout 600 $R --master-port 29564 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads > $O/bench_4gpu_nccl2.json 2> $O/bench_4gpu_nccl2.err; tail -3 $O/bench_4gpu_nccl2.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads > $O/bench_1gpu.json 2> $O/bench_1gpu.err; tail -2 $O/bench_1gpu.err
ls -la $O
