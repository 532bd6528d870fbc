set -x
cd $GRAFT_REPO_ROOT
run() { echo "== $*"; env "$@" timeout 300 python tools/exp_features.py 2>&1 | tail -n 2; }
run FPS_TC_STATS=1
run FPS_TC_STATS=1 FPS_K0_LD=416
run FPS_TC_STATS=1 FPS_TC_SEG=4
run FPS_TC_STATS=1 FPS_TC_SEG=6
run FPS_TC_STATS=1 FPS_TC_SEG=2
