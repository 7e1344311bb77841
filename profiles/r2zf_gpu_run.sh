mkdir -p gpurun_out
{
for rep in 1 2; do echo "== C3 default"; python profiles/r2_c3_run.py 2>&1 | tail -1; done
echo "== C3 OPM_B200_GF_FUSE_LENS=0"; OPM_B200_GF_FUSE_LENS=0 python profiles/r2_c3_run.py 2>&1 | tail -1
} > gpurun_out/r2zf_variants.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/c3_launches.csv python profiles/r2_c3_run.py 6 > gpurun_out/ncu_c3.log 2>&1; echo "ncu rc=$?"
cut -c1-230 gpurun_out/r2zf_variants.txt
