set -x
python -m pytest tests/test_gpu_ops.py -q -k "head_conv or homograph" 2>&1 | tail -3
python tools/prof_ops.py --case head_160,head_warp_160 --reps 20
(cd .ab_base && python tools/prof_ops.py --case head_160 --reps 20)
