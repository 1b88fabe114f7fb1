set -x
python -m pytest tests/test_gpu_ops.py -q -k "window_attention" 2>&1 | tail -3
python tools/prof_ops.py --case attn_enc0_64x1_s5,attn_enc0_64x1_s0,attn_enc1_128x2_s5 --reps 10
(cd .ab_base && python tools/prof_ops.py --case attn_enc0_64x1_s5,attn_enc0_64x1_s0,attn_enc1_128x2_s5 --reps 10)
