# round 2: prefetch by program (5+ surfaces), flag ahead everywhere: C4 / C3 with and without, then tests, bench, ncu
mkdir -p gpurun_out
{
echo "== C4 N=1 probe, default (no bulk prefetch for 3 surfaces, flag ahead)"; python profiles/r2_c4_n1_probe.py 2>&1 | head -3
echo "== C4 N=1 probe, -DOPM_TILE_NO_FLAG_AHEAD"; OPM_B200_JIT_DEFINES="-DOPM_TILE_NO_FLAG_AHEAD" python profiles/r2_c4_n1_probe.py 2>&1 | head -3
echo "== C4 N=1 probe, OPM_B200_JIT_PREFETCH=1"; OPM_B200_JIT_PREFETCH=1 python profiles/r2_c4_n1_probe.py 2>&1 | head -3
echo "== C3 default"; python profiles/r2_c3_run.py 2>&1 | tail -3
echo "== C3 -DOPM_TILE_NO_FLAG_AHEAD"; OPM_B200_JIT_DEFINES="-DOPM_TILE_NO_FLAG_AHEAD" python profiles/r2_c3_run.py 2>&1 | tail -3
echo "== C3 OPM_B200_JIT_PREFETCH=1"; OPM_B200_JIT_PREFETCH=1 python profiles/r2_c3_run.py 2>&1 | tail -3
} > gpurun_out/r2o_variants.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-c4 --no-c5 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:opm_jit_trace -s 6 -c 1 -o gpurun_out/jit_trace -f python bench.py --steps 3 --warmup 3 --no-c4 --no-c5 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
cat gpurun_out/r2o_variants.txt; tail -4 gpurun_out/pytest_gpu.log
