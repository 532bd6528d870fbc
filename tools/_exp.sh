cd $GRAFT_REPO_ROOT
timeout 300 python tools/check_gram_i8.py 2>&1 | grep -v snapped | tail -3
timeout 300 python tools/check_project_i8.py 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gi_|gram_i8' --csv --log-file gpurun_out/gram_i8_launches.csv python tools/time_gram.py > /dev/null 2>&1
tail -4 gpurun_out/gram_i8_launches.csv | awk -F'","' '{print $5, $NF}'
timeout 600 python -m pytest tests -x -q -m gpu -k "int8 or gram" 2>&1 | tail -2
