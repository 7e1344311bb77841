# round 2: ghost focus, check batches on the side stream: checks per batch (0 = each check by itself)
mkdir -p gpurun_out
{
for b in 0 8 16 32 64 256; do
echo "== C3 lazy clones, OPM_B200_GF_BATCH=$b"; OPM_B200_GF_BATCH=$b python profiles/r2_c3_run.py 2>&1 | tail -1
done
echo "== again 0 / 16 / 256"
for b in 0 16 256; do OPM_B200_GF_BATCH=$b python profiles/r2_c3_run.py 2>&1 | tail -1; done
} > gpurun_out/r2v_variants.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
cut -c1-200 gpurun_out/r2v_variants.txt; tail -30 gpurun_out/pytest_gpu.log | grep -v "^\.\+" | tail -25
