set -x
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -s > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?"
tail -n 6 gpurun_out/r02_pytest3.log
timeout 300 python tools/time_solve.py 1048576 2>&1 | tail -n 1
timeout 300 python tools/time_solve.py 160000 2>&1 | tail -n 1
FPS_SOLVE_GEMV=0 timeout 300 python tools/time_solve.py 160000 2>&1 | tail -n 1
timeout 600 python bench.py --steps 10 --warmup 3 --no-strong --no-cpu > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; echo "bench rc=$?"
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r02_bench3.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['kernels_ms_per_step'])"
