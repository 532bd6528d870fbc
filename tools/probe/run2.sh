set -x
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -s > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?"
tail -n 8 gpurun_out/r02_pytest2.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-strong > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err; echo "bench rc=$?"
tail -n 3 gpurun_out/r02_bench2.err
FPS_TC_STATS=1 timeout 300 python tools/time_features.py 303104 2>&1 | tail -n 4
timeout 300 python tools/time_solve.py 1048576 2>&1 | tail -n 1
timeout 300 python tools/time_solve.py 160000 2>&1 | tail -n 1
