# CTA shapes of the interpreter kernel (libraries built with make OUT=lib_variants/x TRACE_FLAGS='-DOPM_TRACE_BLOCK=.. -DOPM_TRACE_MIN_CTAS=..'):
# C2 bench with the specialised kernels off, and the C3 ghost-focus analysis (always interpreted)
for v in default b512_m1 b768_m1 b1024_m1 b384_m2; do
  if [ $v = default ]; then unset OPM_B200_LIB; else export OPM_B200_LIB=$PWD/lib_variants/$v/libopossum_b200.so; fi
  OPM_B200_JIT=0 python bench.py --no-cpu-baseline --steps 10 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'C2 interpreted', round(d['value']/1e9,2), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_launch']*d['roofline']['launches_per_step'],3))"
  python profiles/run_configs.py c3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'C3', round(d['seconds']*1e3,3), 'ms')"
done
