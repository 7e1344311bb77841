# round 2: bench with the report reads of the e2e / document-level steps pipelined one step behind (like `value`)
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 800 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'doc', d['e2e']['document_level']['value'], d['e2e']['document_level']['ms_per_step'], 'frac', d['roofline']['frac'], d['parity_vs_oracle']['ok'])
PY
