set -x
for BN in 256 192 128 64; do
  echo "== FBANET_TC_BN=$BN" >> gpurun_out/r2bn_prof.log
  FBANET_TC_BN=$BN python tools/prof_conv.py --case fc1_256_1024,qkv_256_768,qkv_128_384,fc1_128_512,qkv_64_192,proj_256_256 --reps 10 >> gpurun_out/r2bn_prof.log 2>&1
done
cat gpurun_out/r2bn_prof.log
