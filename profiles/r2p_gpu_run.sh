# round 2: host-time profile of the ghost-focus analysis, C4 probe of the final heuristics, GPU tests
mkdir -p gpurun_out
{
echo "== C3 host-time profile"; OPM_B200_GF_PROFILE=1 python profiles/r2_c3_run.py 2>&1 | tail -8
echo "== C3 without the side stream"; OPM_B200_GF_SIDE_STREAM=0 python profiles/r2_c3_run.py 2>&1 | tail -3
echo "== C4 N=1 probe, default"; python profiles/r2_c4_n1_probe.py 2>&1 | head -3
} > gpurun_out/r2p_variants.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
cat gpurun_out/r2p_variants.txt; tail -4 gpurun_out/pytest_gpu.log
