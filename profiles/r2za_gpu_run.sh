# round 2: ghost focus, bounce-level keys delivered by the next launch's prologue (no 8-byte copies on the stream)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
{
for rep in 1 2; do
echo "== C3 default"; python profiles/r2_c3_run.py 2>&1 | tail -1
done
echo "== C2 bench kernel x3"
for rep in 1 2 3; do
python bench.py --no-cpu-baseline --no-c4 --no-c5 --steps 30 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms_per_launch'],4), round(d['roofline']['frac'],4))"
done
} > gpurun_out/r2za_variants.txt 2>&1
cut -c1-230 gpurun_out/r2za_variants.txt; tail -40 gpurun_out/pytest_gpu.log | grep -v "^\.\+" | tail -30
