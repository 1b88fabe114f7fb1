set -x
python -m pytest tests/test_gpu_tcgen05.py -q -x -k "leff_mlp" 2>&1 | tail -3 > gpurun_out/r2o_pytest.log; cat gpurun_out/r2o_pytest.log
python tools/prof_ops.py --case mlp_dec1_128,mlp_enc1_128,mlp_enc0_64 --reps 10 > gpurun_out/r2o_prof.log 2>&1
echo "== FBANET_LEFF_ORDER=0 (fixed order)" >> gpurun_out/r2o_prof.log
FBANET_LEFF_ORDER=0 python tools/prof_ops.py --case mlp_dec1_128,mlp_enc1_128,mlp_enc0_64 --reps 10 >> gpurun_out/r2o_prof.log 2>&1
cat gpurun_out/r2o_prof.log
