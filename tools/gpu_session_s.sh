set -x
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --other-configs none --no-cpu-baseline > gpurun_out/r2s_ncu_bench.log 2>&1
tail -2 gpurun_out/r2s_ncu_bench.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:faf_fuse -c 1 -o gpurun_out/r2s_faf_fuse python tools/prof_ops.py --case faf_fuse_160 --reps 1 > gpurun_out/r2s_ncu_faf.log 2>&1
tail -2 gpurun_out/r2s_ncu_faf.log
