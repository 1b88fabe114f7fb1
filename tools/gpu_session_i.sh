set -x
python -m pytest tests -m gpu -q -s -k "psnr_delta or narrow_io or bf16_embed64" 2>&1 | grep -E "psnr delta|passed|failed|worst|Error|error|assert" | head -40 > gpurun_out/r2i_pytest_sel.log
cat gpurun_out/r2i_pytest_sel.log
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2i_pytest.log
tail -3 gpurun_out/r2i_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
tail -c 3000 gpurun_out/r2i_bench.json
