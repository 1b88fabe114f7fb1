set -x
python tools/prof_ops.py --case warp_160 --reps 10 > gpurun_out/r2v_prof.log 2>&1
FBANET_WARP_PLANAR4=1 python tools/prof_ops.py --case warp_160 --reps 10 >> gpurun_out/r2v_prof.log 2>&1
cat gpurun_out/r2v_prof.log
ncu --set full --clock-control none --import-source on -k regex:warp_rows -c 1 -o gpurun_out/r2v_warp_rows python tools/prof_ops.py --case warp_160 --reps 1 > gpurun_out/r2v_ncu1.log 2>&1
FBANET_WARP_PLANAR4=1 ncu --set full --clock-control none --import-source on -k regex:warp_planar4 -c 1 -o gpurun_out/r2v_warp_p4 python tools/prof_ops.py --case warp_160 --reps 1 > gpurun_out/r2v_ncu2.log 2>&1
tail -2 gpurun_out/r2v_ncu1.log gpurun_out/r2v_ncu2.log
