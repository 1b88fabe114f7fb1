set -x
python -m pytest tests/test_gpu_ops.py -q -k "window_attention or torch_library" 2>&1 | tail -4 > gpurun_out/r2t_pytest_attn.log
tail -3 gpurun_out/r2t_pytest_attn.log
python -m pytest tests/test_gpu_tcgen05.py -q -k "leff_mlp" 2>&1 | tail -4 > gpurun_out/r2t_pytest_mlp.log
tail -3 gpurun_out/r2t_pytest_mlp.log
for P in 0 1 2 3; do
  echo "== FBANET_ATTN_POLY=$P" >> gpurun_out/r2t_prof.log
  FBANET_ATTN_POLY=$P python tools/prof_ops.py --case attn_dec1_128x8_s5,attn_dec1_128x8_s0,attn_dec0_256x16_s5,attn_bott_256x16_s5 --reps 5 >> gpurun_out/r2t_prof.log 2>&1
done
echo "== mlp new" >> gpurun_out/r2t_prof.log
python tools/prof_ops.py --case mlp_dec1_128,mlp_enc1_128,mlp_enc0_64 --reps 5 >> gpurun_out/r2t_prof.log 2>&1
echo "== mlp old (.ab_base)" >> gpurun_out/r2t_prof.log
(cd .ab_base && python tools/prof_ops.py --case mlp_dec1_128,mlp_enc1_128,mlp_enc0_64 --reps 5) >> gpurun_out/r2t_prof.log 2>&1
echo "== mlp new again" >> gpurun_out/r2t_prof.log
python tools/prof_ops.py --case mlp_dec1_128 --reps 5 >> gpurun_out/r2t_prof.log 2>&1
cat gpurun_out/r2t_prof.log
python -m pytest tests/test_gpu_model.py -q -k "psnr_delta or batch_invariance" -s 2>&1 | grep -E "seed|passed|failed" > gpurun_out/r2t_pytest_model.log
cat gpurun_out/r2t_pytest_model.log
