set -x
python -m pytest tests/test_gpu_ops.py -q -k "warp or flow" 2>&1 | tail -4 > gpurun_out/r2u_pytest_warp.log
tail -3 gpurun_out/r2u_pytest_warp.log
python tools/prof_ops.py --case warp_160,flow_160 --reps 10 > gpurun_out/r2u_prof.log 2>&1
echo "== old kernels" >> gpurun_out/r2u_prof.log
FBANET_WARP_PLANAR4=1 FBANET_FLOW_ROWS=0 python tools/prof_ops.py --case warp_160,flow_160 --reps 10 >> gpurun_out/r2u_prof.log 2>&1
cat gpurun_out/r2u_prof.log
python -m pytest tests/test_gpu_model.py -q -s 2>&1 | grep -E "seed|passed|failed|rror" > gpurun_out/r2u_pytest_model.log
cat gpurun_out/r2u_pytest_model.log
