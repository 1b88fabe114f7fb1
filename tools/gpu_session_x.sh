set -x
python -m pytest tests/test_gpu_ops.py -q -x -k "head_conv or warp or homograph" 2>&1 | tail -6 > gpurun_out/r2x_pytest.log
cat gpurun_out/r2x_pytest.log
python -m pytest tests/test_abi.py tests/test_gpu_model.py -q -x -k "not psnr" 2>&1 | tail -4 >> gpurun_out/r2x_pytest.log
tail -4 gpurun_out/r2x_pytest.log
python tools/prof_ops.py --case head_160,head_warp_160,warp_160 --reps 10 > gpurun_out/r2x_prof.log 2>&1
echo "== previous kernels" >> gpurun_out/r2x_prof.log
(cd .ab_base && python tools/prof_ops.py --case head_160,warp_160 --reps 10) >> gpurun_out/r2x_prof.log 2>&1
cat gpurun_out/r2x_prof.log
