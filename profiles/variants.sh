for v in default b128_m4 b128_m5 b128_m6 b256_m3 b512_m1 b64_m8; do
  if [ $v = default ]; then unset OPM_B200_LIB; else export OPM_B200_LIB=$PWD/lib_variants/$v/libopossum_b200.so; fi
  python bench.py --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e9,2), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_launch']*2,3))"
done
