set -x
timeout 900 python -m pytest tests/test_gpu_exact.py -m gpu -q --maxfail=30 > gpurun_out/r02_pytest7.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r02_pytest7.log
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-strong --no-cpu --batch 0 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('run$i', d['value'], d['ms_per_step'], d['kernels_ms_per_step'], d['roofline']['frac'], d['roofline']['issued_frac'])"
done
