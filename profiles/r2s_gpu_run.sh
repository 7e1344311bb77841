# round 2: ghost-focus analysis, 14 analyses per variant
mkdir -p gpurun_out
{
for rep in 1 2; do
echo "== C3 batched checks (default)"; python profiles/r2_c3_run.py 2>&1 | tail -1
echo "== C3 one check at a time, side stream (OPM_B200_GF_BATCH=0)"; OPM_B200_GF_BATCH=0 python profiles/r2_c3_run.py 2>&1 | tail -1
done
echo "== C3 one check at a time, main stream"; OPM_B200_GF_BATCH=0 OPM_B200_GF_SIDE_STREAM=0 python profiles/r2_c3_run.py 2>&1 | tail -1
echo "== host profile, batched"; OPM_B200_GF_PROFILE=1 python profiles/r2_c3_run.py 6 2>&1 | tail -3
} > gpurun_out/r2s_variants.txt 2>&1
cat gpurun_out/r2s_variants.txt
