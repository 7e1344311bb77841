# CTA shapes of the specialised kernel after the translation-only specialisation (fewer live values): value 1e9/s, step ms, kernel ms
for v in "1024 1" "768 1" "896 1" "640 1" "512 2" "384 2"; do
  set -- $v
  OPM_B200_JIT_BLOCK=$1 OPM_B200_JIT_MIN_CTAS=$2 python bench.py --no-cpu-baseline --steps 20 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 x $2', round(d['value']/1e9,2), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_launch']*d['roofline']['launches_per_step'],3))"
done
