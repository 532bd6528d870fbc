cd $GRAFT_REPO_ROOT
run() { echo "== $*"; env "$@" REPS=8 timeout 300 python tools/exp_features.py 1212416 4096 2>&1 | tail -n 1; }
for i in 1 2; do
run CHUNK=37888
run CHUNK=75776
run CHUNK=151552
run CHUNK=303104
done
