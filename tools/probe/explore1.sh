set -x
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
python tools/probe/int8_peak.py 2>&1 | tail -3
for seg in 3 6 11; do
  FPS_TC_SEG=$seg python tools/calibrate_kappa.py 2>&1 | tail -1
  FPS_TC_SEG=$seg FPS_TC_STATS=1 python tools/time_features.py 303104 2>&1 | tail -4
done
python tools/time_solve.py 1048576 2>&1 | tail -1
