cd $GRAFT_REPO_ROOT
run() { echo "== $*"; env "$@" timeout 300 python tools/exp_features.py 2>&1 | tail -n 1; }
run FPS_TC_STATS=0
run FPS_TC_SEG_HID=3
run FPS_TC_SEG_HID=3 FPS_TC_SEG_L0=2
run FPS_TC_SEG_L0=3 FPS_TC_SEG_HID=3 FPS_TC_SEG_LAST=4
run FPS_TC_SEG=3
run FPS_TC_SEG_LAST=6
