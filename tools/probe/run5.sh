set -x
timeout 900 python -m pytest tests/test_gpu_exact.py tests/test_gpu_parity.py -m gpu -q --maxfail=30 -k "chol or golden or kernels or mutate" > gpurun_out/r02_pytest5.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r02_pytest5.log
timeout 300 python tools/time_solve.py 1048576 2>&1 | tail -n 1
timeout 300 python tools/time_solve.py 160000 2>&1 | tail -n 1
timeout 1500 bash tools/profile_r02.sh 2>&1 | tail -n 30
