set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r21; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv
timeout 300 python -m pytest tests/test_clips_dist.py -m gpu -q -x 2>&1 | tail -15 > $O/pytest_peer.log; cat $O/pytest_peer.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_2gpu_peer.json 2> $O/bench_2gpu_peer.err; tail -8 $O/bench_2gpu_peer.err; tail -c 300 $O/bench_2gpu_peer.json
E2F_STITCH=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29554 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads > $O/bench_2gpu_nccl.json 2> $O/bench_2gpu_nccl.err; tail -3 $O/bench_2gpu_nccl.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads > $O/bench_2gpu_peer2.json 2> $O/bench_2gpu_peer2.err; tail -3 $O/bench_2gpu_peer2.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads > $O/bench_1gpu.json 2> $O/bench_1gpu.err; tail -2 $O/bench_1gpu.err
ls -la $O
