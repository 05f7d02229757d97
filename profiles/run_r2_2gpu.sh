# round 2: two-GPU sanity of the bench contract (weak scaling, no step-path collective)
cd "$(dirname "$0")/.."
O=gpurun_out/r2g
mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; tail -c 300 $O/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/bench_2gpu_reference.json 2> $O/bench_2gpu_reference.err
cat $O/bench_2gpu.json | cut -c1-400
