# round 2: bulk prefetch per warp instead of per CTA (probe)
mkdir -p gpurun_out
run() {
  env $3 OPM_B200_JIT_DEFINES="$1" python bench.py --no-cpu-baseline --no-c4 --no-c5 --steps 30 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('[$1] [$2] [$3]', round(d['value']/1e9,2), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms_per_launch']*d['roofline']['launches_per_step'],4), round(d['roofline']['frac'],4))"
}
{
run "" default X=0
run "-DOPM_TILE_BULK_PREFETCH_PER_WARP" "per warp" X=0
run "-DOPM_TILE_BULK_PREFETCH_PER_WARP -DOPM_TILE_PREFETCH_DIST=2" "per warp, 2 tiles ahead" X=0
run "" default-again X=0
run "-DOPM_TILE_BULK_PREFETCH_PER_WARP" "per warp again" X=0
} 2>&1 | tee gpurun_out/r2y_variants.txt
