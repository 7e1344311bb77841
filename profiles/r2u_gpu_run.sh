# round 2: ghost focus, variants measured after the specialised kernels are there (jit_wait)
mkdir -p gpurun_out
{
for rep in 1 2; do
echo "== C3 default (batched checks, lazy clones)"; python profiles/r2_c3_run.py 2>&1 | tail -1
echo "== C3 OPM_B200_GF_LAZY_CLONE=0"; OPM_B200_GF_LAZY_CLONE=0 python profiles/r2_c3_run.py 2>&1 | tail -1
echo "== C3 OPM_B200_GF_LAZY_CLONE=0 OPM_B200_GF_BATCH=0 (start of this session)"; OPM_B200_GF_LAZY_CLONE=0 OPM_B200_GF_BATCH=0 python profiles/r2_c3_run.py 2>&1 | tail -1
done
echo "== host profile"; OPM_B200_GF_PROFILE=1 python profiles/r2_c3_run.py 8 2>&1 | tail -2
} > gpurun_out/r2u_variants.txt 2>&1
cat gpurun_out/r2u_variants.txt
