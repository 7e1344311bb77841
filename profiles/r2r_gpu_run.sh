# round 2: ghost-focus checks deferred into batches (three launches for all of an analysis' checks) against one at a time
mkdir -p gpurun_out
{
echo "== C3 batched checks (default)"; OPM_B200_GF_PROFILE=1 python profiles/r2_c3_run.py 2>&1 | tail -4
echo "== C3 one check at a time (OPM_B200_GF_BATCH=0)"; OPM_B200_GF_BATCH=0 python profiles/r2_c3_run.py 2>&1 | tail -3
} > gpurun_out/r2r_variants.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
cat gpurun_out/r2r_variants.txt; tail -30 gpurun_out/pytest_gpu.log | grep -v "^\.\+" | tail -25
