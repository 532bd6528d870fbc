cd $GRAFT_REPO_ROOT
timeout 300 python tools/check_gram_i8.py 2>&1 | grep -v snapped | tail -9
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gi_|gram_i8' --csv --log-file gpurun_out/gram_i8_launches.csv python tools/_gram_time.py > /dev/null 2>&1
tail -5 gpurun_out/gram_i8_launches.csv | awk -F'","' '{print $5, $NF}'
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_i8.json 2> gpurun_out/bench_i8.err; tail -c 3000 gpurun_out/bench_i8.json
