for v in "768 1 0" "768 1 1" "512 2 0" "512 2 1"; do
  set -- $v
  OPM_B200_JIT_BLOCK=$1 OPM_B200_JIT_MIN_CTAS=$2 OPM_B200_JIT_PIPE=$3 python bench.py --no-cpu-baseline --steps 20 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 x $2 pipe=$3', round(d['value']/1e9,2), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_launch']*d['roofline']['launches_per_step'],3), d['roofline']['jit'])"
done
