# CTA shapes of the run-time specialised kernel: kernel ms per step (2 launches) of bench.py
for v in "256 2" "256 3" "128 4" "128 5" "128 6" "128 8" "512 1" "64 10"; do
  set -- $v
  OPM_B200_JIT_BLOCK=$1 OPM_B200_JIT_MIN_CTAS=$2 python bench.py --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 x $2', round(d['value']/1e9,2), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_launch']*2,3))"
done
