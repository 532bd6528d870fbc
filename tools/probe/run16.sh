cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/r02_pytest_multi2gpu.log 2>&1; echo "pytest rc=$?"
grep -E "world|passed|failed|Error" gpurun_out/r02_pytest_multi2gpu.log | tail -n 12
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke_2gpu_b.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r02_smoke_2gpu_b.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --quick > gpurun_out/r02_bench_2gpu_b.json 2> gpurun_out/r02_bench_2gpu_b.err; echo "bench2 rc=$?"
python - <<'P'
import json
d=[json.loads(l) for l in open('gpurun_out/r02_bench_2gpu_b.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['multi_gpu_parity'])
P
