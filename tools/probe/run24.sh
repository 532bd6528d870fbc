cd $GRAFT_REPO_ROOT
CMD="python tools/prof_step.py"
$CMD > gpurun_out/r02c_prof_plain.log 2>&1 || { echo "plain run failed"; tail -n 20 gpurun_out/r02c_prof_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:recon_i8_solo_kernel" -s 2 -c 1 -f -o gpurun_out/r02c_prof_recon_solo $CMD > gpurun_out/r02c_ncu_recon_solo.log 2>&1; echo "rc=$?"
ls -la gpurun_out/r02c*.ncu-rep
