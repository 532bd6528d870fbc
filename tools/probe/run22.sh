cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_exact.py -m gpu -x -q > gpurun_out/r02_pytest11.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r02_pytest11.log
for i in 1 2; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-strong --no-cpu 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
b=d['batched_solve']; print(b['value'], b['ms_per_call'], b['kernels_ms'], b['issued_int8_tops'])"
done
