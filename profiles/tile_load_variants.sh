# tile prologue of the specialised kernel on C2: value 1e9/s, step ms, kernel ms
for v in "" "-DOPM_TILE_INDEPENDENT_LOADS" "-DOPM_TILE_INDEPENDENT_LOADS -DOPM_TILE_PREFETCH" "-DOPM_TILE_PREFETCH"; do
  OPM_B200_JIT_DEFINES="$v" python bench.py --no-cpu-baseline --steps 20 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('[$v]', round(d['value']/1e9,2), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_launch']*d['roofline']['launches_per_step'],3))"
done
