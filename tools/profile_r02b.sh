#!/bin/bash
# ncu evidence of the layer kernel after the round-2b changes (aligned first-layer pitch, 256-wide segments in the first and
# last layer): launch list + one --set full capture per layer-kernel role.  Every ncu run follows a plain run of the same
# command in the same call (B200_PROFILING.md).
cd $GRAFT_REPO_ROOT
CMD="python tools/prof_step.py"
$CMD > gpurun_out/r02b_prof_plain.log 2>&1 || { echo "plain run failed"; tail -n 20 gpurun_out/r02b_prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches.csv $CMD > gpurun_out/r02b_ncu_launches.log 2>&1
cap() {  # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      -f -o gpurun_out/r02b_prof_$1 $CMD > gpurun_out/r02b_ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
cap hidden 'layer_tc_pair_kernel<\(int\)4, \(bool\)0, \(bool\)1' 12 1
cap last 'layer_tc_pair_kernel<\(int\)4, \(bool\)1, \(bool\)1' 2 1
cap layer0 'layer_tc_pair_kernel<\(int\)4, \(bool\)0, \(bool\)1' 10 1
ls -la gpurun_out/r02b*.ncu-rep
