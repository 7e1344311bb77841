# tile prologue variants of the specialised kernel on C2 (round 2, final kernel): value 1e9/s, step ms, kernel ms
mkdir -p gpurun_out
run() {
  OPM_B200_JIT_DEFINES="$1" python bench.py --no-cpu-baseline --no-c4 --no-c5 --steps 30 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('[$1] [$2]', round(d['value']/1e9,2), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms_per_launch']*d['roofline']['launches_per_step'],4), d['roofline']['frac'], d.get('parity_vs_oracle',{}).get('ok'))"
}
{
run "" base
run "-DOPM_TILE_INDEPENDENT_LOADS" ""
run "-DOPM_TILE_FLAG_AHEAD" ""
run "-DOPM_TILE_BULK_PREFETCH" ""
run "-DOPM_TILE_FLAG_AHEAD -DOPM_TILE_BULK_PREFETCH" ""
run "-DOPM_TILE_INDEPENDENT_LOADS -DOPM_TILE_BULK_PREFETCH" ""
run "-DOPM_TILE_PREFETCH" ""
run "" base-again
} 2>&1 | tee gpurun_out/tile_variants.txt
for v in "-DOPM_TILE_FLAG_AHEAD" "-DOPM_TILE_FLAG_AHEAD -DOPM_TILE_BULK_PREFETCH"; do
  echo "== pytest with [$v]" >> gpurun_out/tile_variants.txt
  OPM_B200_JIT_DEFINES="$v" timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/tile_variants.txt
done
tail -8 gpurun_out/tile_variants.txt
