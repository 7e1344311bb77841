# round 2, build with the next tile's flag loaded ahead + bulk L2 prefetch as defaults: variants, tests, bench, ncu
mkdir -p gpurun_out
run() {
  env $3 OPM_B200_JIT_DEFINES="$1" python bench.py --no-cpu-baseline --no-c4 --no-c5 --steps 30 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('[$1] [$2] [$3]', round(d['value']/1e9,2), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms_per_launch']*d['roofline']['launches_per_step'],4), round(d['roofline']['frac'],4))"
}
{
run "" default X=0
run "-DOPM_TILE_PREFETCH_DIST=2" "" X=0
run "" "block 1024" OPM_B200_JIT_BLOCK=1024
run "" "block 768" OPM_B200_JIT_BLOCK=768
run "" "block 640" OPM_B200_JIT_BLOCK=640
run "-DOPM_TILE_NO_BULK_PREFETCH -DOPM_TILE_NO_FLAG_AHEAD" "old" X=0
run "" default-again X=0
} 2>&1 | tee gpurun_out/r2n_variants.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-c4 --no-c5 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:opm_jit_trace -s 6 -c 1 -o gpurun_out/jit_trace -f python bench.py --steps 3 --warmup 3 --no-c4 --no-c5 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -4 gpurun_out/pytest_gpu.log
