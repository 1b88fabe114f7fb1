set -x
python -m pytest tests/test_gpu_ops.py -q -k "warp or flow" 2>&1 | tail -4 > gpurun_out/r2w_pytest_warp.log
tail -3 gpurun_out/r2w_pytest_warp.log
python tools/prof_ops.py --case warp_160,flow_160 --reps 10 > gpurun_out/r2w_prof.log 2>&1
echo "== previous kernels" >> gpurun_out/r2w_prof.log
(cd .ab_base && python tools/prof_ops.py --case warp_160,flow_160 --reps 10) >> gpurun_out/r2w_prof.log 2>&1
cat gpurun_out/r2w_prof.log
