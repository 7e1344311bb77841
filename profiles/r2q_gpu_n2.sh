# round 2, final build at N = 2: bench line (C2 weak scaling, C4 strong with parity vs one GPU, C5 shards), NCCL + peer-memory read-out
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/bench_n2.err
