cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_4gpu_c.json 2> gpurun_out/r02_bench_4gpu_c.err; echo "bench4 rc=$?"
tail -n 3 gpurun_out/r02_bench_4gpu_c.err
python - <<'P'
import json
d=[json.loads(l) for l in open('gpurun_out/r02_bench_4gpu_c.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['batched_solve']['value'], d['multi_gpu_parity']['worst_u_rel_l2_batched'], d['multi_gpu_parity']['pass'])
s=d['strong']
for k,v in s.items():
    if isinstance(v,dict): print(k, {a:b for a,b in v.items() if a not in ('workload','accuracy_vs_fd','collectives')})
P
