set -x
python -m pytest tests/test_gpu_tcgen05.py -q -x -k "fused_leff_dwconv_fc2" 2>&1 | tail -3 > gpurun_out/r2fc2_pytest.log; cat gpurun_out/r2fc2_pytest.log
python tools/prof_ops.py --case leff_dec0_256,leff_dec1_128,leff_enc0_64 --reps 10 > gpurun_out/r2fc2_prof.log 2>&1
echo "== FBANET_LEFF_ITEM=0" >> gpurun_out/r2fc2_prof.log
FBANET_LEFF_ITEM=0 python tools/prof_ops.py --case leff_dec0_256,leff_dec1_128,leff_enc0_64 --reps 10 >> gpurun_out/r2fc2_prof.log 2>&1
cat gpurun_out/r2fc2_prof.log
python tools/prof_conv.py --case fc1_256_1024,qkv_256_768,proj_256_256 --reps 10
