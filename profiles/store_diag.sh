# diagnosis: which element width is behind lts__t_sectors_data_ecc?  (results of these builds are not valid rays)
for v in "-DOPM_DIAG_NO_U8" "-DOPM_DIAG_NO_U32" "-DOPM_DIAG_NO_U8 -DOPM_DIAG_NO_U32"; do
  export OPM_B200_JIT_DEFINES="$v"
  echo "== $v"
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_data_ecc.sum --clock-control none -k regex:opm_jit_trace -s 2 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "dram__|ecc"
done
