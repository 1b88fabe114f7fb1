set -x
python -m pytest tests/test_gpu_tcgen05.py -q -k "fp16 or leff" 2>&1 | tail -5 > gpurun_out/r2i_pytest.log; cat gpurun_out/r2i_pytest.log
python tools/prof_ops.py --case leff16_dec0_256,leff_dec0_256 --reps 10 > gpurun_out/r2i_prof.log 2>&1; cat gpurun_out/r2i_prof.log
python -m pytest tests/test_gpu_model.py -q -s -k "psnr_delta or batch_invariance or small_model or within_1e3" 2>&1 | grep -E "seed|passed|failed|rror" > gpurun_out/r2i_pytest_model.log
cat gpurun_out/r2i_pytest_model.log
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown --other-configs none --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
grep "breakdown" gpurun_out/r2i_bench.err | grep -E "one step|leff_fc2|256->1024"
