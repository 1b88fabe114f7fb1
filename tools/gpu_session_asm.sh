set -x
python -m pytest tests/test_gpu_ops.py -q -k "assemble" 2>&1 | tail -3
python -m pytest tests/test_gpu_model.py -q -k "within_1e3 or every_stage or batch_invariance" 2>&1 | tail -3
python tools/prof_ops.py --case assemble_160 --reps 20
FBANET_ASSEMBLE_QUAD=0 python tools/prof_ops.py --case assemble_160 --reps 20
