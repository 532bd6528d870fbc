cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest10.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r02_pytest10.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke4.log 2>&1; echo "smoke rc=$?"; tail -n 6 gpurun_out/r02_smoke4.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err; echo "bench rc=$?"
python - <<'P'
import json
d=[json.loads(l) for l in open('gpurun_out/r02_bench5.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['ms_per_step_instrumented'], d['e2e']['ms_per_step'], d['kernels_ms_per_step'], d['roofline']['frac'], d['roofline']['issued_frac'], d['roofline']['traffic'], d['clocks'])
print(d['batched_solve']['value'], d['batched_solve']['kernels_ms'], d.get('cpu_baseline',{}).get('value'))
P
