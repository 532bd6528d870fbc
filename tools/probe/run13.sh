cd $GRAFT_REPO_ROOT
b() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-strong --no-cpu --batch 0 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print(d['ms_per_step'], d['ms_per_step_instrumented'], d['e2e']['ms_per_step'], d['clocks'])"; }
b A=1
b FPS_BENCH_NO_CLOCKS=1
b FPS_BENCH_CLOCK_MS=500
b A=1
b FPS_BENCH_NO_CLOCKS=1
b FPS_BENCH_CLOCK_MS=500
