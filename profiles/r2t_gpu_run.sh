# round 2: ghost focus with lazy clones (a re-injected bundle is written by its first launch, no copy pass)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
{
echo "== C3 default (batched checks, lazy clones)"; python profiles/r2_c3_run.py 2>&1 | tail -1
echo "== C3 OPM_B200_GF_LAZY_CLONE=0"; OPM_B200_GF_LAZY_CLONE=0 python profiles/r2_c3_run.py 2>&1 | tail -1
echo "== C3 default again"; OPM_B200_GF_PROFILE=1 python profiles/r2_c3_run.py 2>&1 | tail -2
} > gpurun_out/r2t_variants.txt 2>&1
cat gpurun_out/r2t_variants.txt; tail -30 gpurun_out/pytest_gpu.log | grep -v "^\.\+" | tail -25
