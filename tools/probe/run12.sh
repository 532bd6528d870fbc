cd $GRAFT_REPO_ROOT
timeout 300 python tools/exp_features.py 303104 8192 2>&1 | tail -n 1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest8.log 2>&1; echo "pytest rc=$?"
tail -n 5 gpurun_out/r02_pytest8.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke3.log 2>&1; echo "smoke rc=$?"; tail -n 6 gpurun_out/r02_smoke3.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err; echo "bench rc=$?"
python - <<'P'
import json
d=[json.loads(l) for l in open('gpurun_out/r02_bench4.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['kernels_ms_per_step'], d['roofline']['frac'], d['roofline']['issued_frac'], d['clocks'])
print(d['batched_solve']['value'], d['batched_solve']['kernels_ms'], d.get('cpu_baseline'))
print({k:(v if not isinstance(v,dict) else {a:b for a,b in v.items() if a in ('ms_per_call','solves_per_s','precompute_ms','solve_ms','solves_ms_total','precompute_points_per_s')}) for k,v in (d.get('strong') or {}).items()})
P
