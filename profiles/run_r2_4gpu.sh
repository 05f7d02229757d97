# round 2: four-GPU bench line (weak scaling, no step-path collective)
cd "$(dirname "$0")/.."
O=gpurun_out/r2g
mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_4gpu.json 2> $O/bench_4gpu.err; tail -c 300 $O/bench_4gpu.err
cat $O/bench_4gpu.json | cut -c1-300
