set -x
CMD="python tools/prof_step.py"
$CMD > gpurun_out/r02_prof_plain.log 2>&1 || { tail -n 20 gpurun_out/r02_prof_plain.log; exit 1; }
cap() {
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      -f -o gpurun_out/r02_prof_$1 $CMD > gpurun_out/r02_ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
cap hidden 'layer_tc_pair_kernel<\(int\)4, \(bool\)0, \(bool\)1' 12 1
cap last 'layer_tc_pair_kernel<\(int\)4, \(bool\)1, \(bool\)1' 2 1
cap layer0 'layer_tc_pair_kernel<\(int\)4, \(bool\)0, \(bool\)1' 10 1
ls -la gpurun_out/r02_prof_*.ncu-rep
for fuse in 1 0; do
FPS_FUSE=$fuse timeout 300 python bench.py --steps 8 --warmup 3 --no-strong --no-cpu --batch 0 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('FUSE=$fuse', d['value'], d['ms_per_step'], d['kernels_ms_per_step'])"
done
