# DRAM traffic of the specialised trace kernel with different cache operators on its stores
for v in "" "-DOPM_STORE_PLAIN" "-DOPM_STORE_CG"; do
  export OPM_B200_JIT_DEFINES="$v"
  echo "== stores: ${v:-default (.cs)}"
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('kernel ms', round(d['roofline']['kernel_ms_per_launch'],3), 'step ms', round(d['ms_per_step'],3))"
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_data_ecc.sum,gpu__time_duration.sum --clock-control none -k regex:opm_jit_trace -s 2 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "dram__|ecc|gpu__time"
done
