# round 2, final build at N = 8: bench line (C2 weak scaling, C4 strong with parity vs one GPU, C5 shards, e2e per-rank uploads)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"
tail -c 600 gpurun_out/bench_n8.err
