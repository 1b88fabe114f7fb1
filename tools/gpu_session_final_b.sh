set -x
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:fbanet -c 1300 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --other-configs none --no-cpu-baseline > gpurun_out/r2f_ncu_bench.log 2>&1
grep -c head_conv gpurun_out/r2f_launches.csv
