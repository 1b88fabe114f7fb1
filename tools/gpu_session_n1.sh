set -x
python tools/prof_ops.py --case leff_dec0_256 --reps 5 > gpurun_out/r2n1_prof.log 2>&1; cat gpurun_out/r2n1_prof.log
ncu --set full --clock-control none --import-source on -k regex:leff_fc2 -c 1 -o gpurun_out/r2n1_leff_fc2 python tools/prof_ops.py --case leff_dec0_256 --reps 1 > gpurun_out/r2n1_ncu.log 2>&1
tail -n 2 gpurun_out/r2n1_ncu.log
