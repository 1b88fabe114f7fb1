set -x
python -m pytest tests/test_gpu_tcgen05.py -q -x -k "layernorm_in" 2>&1 | tail -4 > gpurun_out/r2z_pytest.log
python -m pytest tests/test_gpu_ops.py -q -x -k "window_attention" 2>&1 | tail -3 >> gpurun_out/r2z_pytest.log
cat gpurun_out/r2z_pytest.log
python tools/prof_ops.py --case attn_dec1_128x8_s5,attn_dec0_256x16_s5,attn_bott_256x16_s5 --reps 10 > gpurun_out/r2z_prof.log 2>&1
echo "== previous" >> gpurun_out/r2z_prof.log
(cd .ab_base && python tools/prof_ops.py --case attn_dec1_128x8_s5,attn_dec0_256x16_s5,attn_bott_256x16_s5 --reps 10) >> gpurun_out/r2z_prof.log 2>&1
cat gpurun_out/r2z_prof.log
python tools/prof_e2e_sched.py > gpurun_out/r2z_e2e_sched.log 2>&1
cat gpurun_out/r2z_e2e_sched.log
python - > gpurun_out/r2z_cfg3.log 2>&1 <<'PY'
import os, sys, json, torch
sys.path.insert(0, '.')
import bench
dev = torch.device('cuda:0'); stream = torch.cuda.Stream()
for fuse in ("1", "0"):
    os.environ["FBANET_FUSE_WARP"] = fuse
    r = bench.other_cfg3(dev, stream, 0, 1, 10, "bf16")
    print("FBANET_FUSE_WARP", fuse, json.dumps({k: r[k] for k in ("ms_per_step", "bursts_per_s", "warp_fused_into_head_conv", "head_conv_ms", "warp_ms")}))
PY
cat gpurun_out/r2z_cfg3.log
