cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/r02_pytest_multi1gpu.log 2>&1; echo "pytest rc=$?"
grep -E "world|passed|failed|Error" gpurun_out/r02_pytest_multi1gpu.log | tail -n 12
