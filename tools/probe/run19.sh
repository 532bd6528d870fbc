cd $GRAFT_REPO_ROOT
run() { echo "== $*"; env "$@" REPS=10 timeout 300 python tools/exp_features.py 606208 4096 2>&1 | tail -n 1; }
run CHUNK=0
run CHUNK=18944
run CHUNK=9472
run CHUNK=4736
run CHUNK=0
run CHUNK=75776
