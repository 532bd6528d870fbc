set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_exact.py -m gpu -q --maxfail=30 > gpurun_out/r02_pytest4.log 2>&1; echo "pytest rc=$?"
tail -n 6 gpurun_out/r02_pytest4.log
timeout 300 python tools/time_solve.py 1048576 2>&1 | tail -n 1
FPS_STREAM=0 timeout 300 python tools/time_solve.py 1048576 2>&1 | tail -n 1
timeout 300 python tools/time_solve.py 160000 2>&1 | tail -n 1
timeout 300 python tools/time_solve.py 16777 2>&1 | tail -n 1
