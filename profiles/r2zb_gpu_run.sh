# round 2: launch list of the final ghost-focus analysis (C3)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/c3_launches.csv python profiles/r2_c3_run.py 6 > gpurun_out/ncu_c3.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/ncu_c3.log | cut -c1-200
