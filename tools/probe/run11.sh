cd $GRAFT_REPO_ROOT
run() { echo "== $*"; env "$@" REPS=10 timeout 300 python tools/exp_features.py 606208 4096 2>&1 | tail -n 1; }
run FPS_TC_SEG=3
for i in 1 2 3; do
run FPS_TC_SEG=3
run FPS_TC_SEG=4
run FPS_TC_SEG_HID=3
done
run FPS_K0_LD=416
run FPS_TC_SEG=4
