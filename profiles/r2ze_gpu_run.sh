# round 2: ghost focus, deferred checks of up to four hit maps (re-injected bundles), lens fusion for them too
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
{
for rep in 1 2; do echo "== C3 default"; python profiles/r2_c3_run.py 2>&1 | tail -1; done
} > gpurun_out/r2ze_variants.txt 2>&1
cut -c1-230 gpurun_out/r2ze_variants.txt; tail -40 gpurun_out/pytest_gpu.log | grep -v "^\.\+" | tail -30
