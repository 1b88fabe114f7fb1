set -x
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --other-configs none --no-cpu-baseline > gpurun_out/r2f_ncu_bench.log 2>&1
tail -n 1 gpurun_out/r2f_ncu_bench.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:leff_mlp -c 1 -o gpurun_out/r2f_leff_mlp python tools/prof_ops.py --case mlp_dec1_128 --reps 1 > gpurun_out/r2f_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -c 1 -o gpurun_out/r2f_body3x3 python tools/prof_conv.py --case body3x3_64_64 --reps 1 > gpurun_out/r2f_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dh16 -c 1 -o gpurun_out/r2f_attn_dh16 python tools/prof_ops.py --case attn_dec1_128x8_s5 --reps 1 > gpurun_out/r2f_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -c 1 -o gpurun_out/r2f_attn_tc python tools/prof_ops.py --case attn_enc0_64x1_s5 --reps 1 > gpurun_out/r2f_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:head_conv -c 1 -o gpurun_out/r2f_head python tools/prof_ops.py --case head_160 --reps 1 > gpurun_out/r2f_ncu5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:warp_planar4 -c 1 -o gpurun_out/r2f_warp python tools/prof_ops.py --case warp_160 --reps 1 > gpurun_out/r2f_ncu6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -c 1 -o gpurun_out/r2f_qkv python tools/prof_conv.py --case qkv_128_384 --reps 1 > gpurun_out/r2f_ncu7.log 2>&1
ls -la gpurun_out/r2f_*
