set -x
python -m pytest tests/test_gpu_tcgen05.py -q -k "leff_mlp" 2>&1 | tail -5 > gpurun_out/r2h_pytest.log; cat gpurun_out/r2h_pytest.log
python tools/prof_ops.py --case mlp16_dec1_128,mlp16_enc1_128,mlp16_enc0_64,mlp_dec1_128,mlp_enc1_128,mlp_enc0_64 --reps 10 > gpurun_out/r2h_prof.log 2>&1
cat gpurun_out/r2h_prof.log
python -m pytest tests/test_gpu_model.py -q -s -k "psnr_delta or batch_invariance or small_model" 2>&1 | grep -E "seed|passed|failed|rror" > gpurun_out/r2h_pytest_model.log
cat gpurun_out/r2h_pytest_model.log
