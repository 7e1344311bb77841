# CTA shapes of the specialised kernel on the short broadband chain (C5, 10^8 rays) and the C4 lens + detector
for v in "1024 1" "768 1" "384 2" "256 3" "128 5" "512 2"; do
  set -- $v
  OPM_B200_JIT_BLOCK=$1 OPM_B200_JIT_MIN_CTAS=$2 python profiles/run_configs.py c4 c5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$1 x $2', d['config'][:12], round(d['seconds']*1e3,3), 'ms')"
done
