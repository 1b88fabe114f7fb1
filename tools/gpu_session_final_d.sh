set -x
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:fbanet -c 1300 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --other-configs none --no-cpu-baseline > gpurun_out/r2f_ncu_bench.log 2>&1
grep -c head_conv gpurun_out/r2f_launches.csv
ncu --set full --clock-control none --import-source on -k regex:leff_mlp -c 1 -o gpurun_out/r2f_leff_mlp python tools/prof_ops.py --case mlp16_dec1_128 --reps 1 > gpurun_out/r2f_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:leff_fc2 -c 1 -o gpurun_out/r2f_leff_fc2 python tools/prof_ops.py --case leff_dec0_256 --reps 1 > gpurun_out/r2f_ncu8.log 2>&1
ls -la gpurun_out/r2f_leff*
