set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > gpurun_out/r02_pytest_multi.log 2>&1; echo "pytest multi rc=$?"
tail -n 5 gpurun_out/r02_pytest_multi.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke_2gpu.log 2>&1; echo "smoke rc=$?"
tail -n 4 gpurun_out/r02_smoke_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --quick > gpurun_out/r02_bench_2gpu_quick.json 2> gpurun_out/r02_bench_2gpu_quick.err; echo "bench2 rc=$?"
tail -n 5 gpurun_out/r02_bench_2gpu_quick.err
