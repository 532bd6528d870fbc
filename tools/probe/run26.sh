cd $GRAFT_REPO_ROOT
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/probe/dist_solve_time.py 2097152 2>&1 | grep "^rank"
