# does the size of the contiguous chunk a CTA writes change the ECC read traffic?
for v in "512 1" "256 2" "64 10"; do
  set -- $v
  export OPM_B200_JIT_BLOCK=$1 OPM_B200_JIT_MIN_CTAS=$2
  echo "== CTA $1 x $2"
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_data_ecc.sum,gpu__time_duration.sum --clock-control none -k regex:opm_jit_trace -s 2 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "dram__|ecc|gpu__time"
done
