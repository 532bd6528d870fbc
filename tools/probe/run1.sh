# first GPU pass of round 2: tests, default bench without the strong block, quick strong block, smoke
set -x
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,memory.total --format=csv
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -s > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest1.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-strong > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; echo "bench rc=$?"
timeout 900 python bench_configs.py --config all --quick > gpurun_out/r02_strong_quick.json 2> gpurun_out/r02_strong_quick.err; echo "strong rc=$?"
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke1.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/r02_bench1.err gpurun_out/r02_strong_quick.err gpurun_out/r02_smoke1.log
