cd $GRAFT_REPO_ROOT
FPS_RECON_SOLO=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_exact.py -m gpu -x -q > gpurun_out/r02_pytest12.log 2>&1; echo "pytest rc=$?"
tail -n 6 gpurun_out/r02_pytest12.log
for v in 1 0 1 0; do
FPS_RECON_SOLO=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-strong --no-cpu 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
b=d['batched_solve']; print('solo=$v', b['value'], b['ms_per_call'], b['kernels_ms'])"
done
