#!/bin/bash
# ncu evidence of round 2 (run under gpurun, one GPU): launch list + one --set full capture per kernel of the path.
# Every ncu run follows a plain run of the same command in the same call (B200_PROFILING.md).
set -x
CMD="python tools/prof_step.py"
$CMD > gpurun_out/r02_prof_plain.log 2>&1 || { echo "plain run failed"; tail -n 20 gpurun_out/r02_prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
cap() {  # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      -f -o gpurun_out/r02_prof_$1 $CMD > gpurun_out/r02_ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
cap hidden 'layer_tc_pair_kernel<\(int\)4, \(bool\)0, \(bool\)1' 12 1
cap last 'layer_tc_pair_kernel<\(int\)4, \(bool\)1, \(bool\)1' 2 1
cap layer0 'layer_tc_pair_kernel<\(int\)4, \(bool\)0, \(bool\)1' 10 1
cap gram 'gram_i8_kernel' 1 2
cap recon_i8 'recon_i8_kernel' 2 1
cap stream 'proj_stream_kernel|recon_stream_kernel' 0 3
cap chol 'chol_fused_kernel|tri_gemv_kernel' 1 3
ls -la gpurun_out/*.ncu-rep
