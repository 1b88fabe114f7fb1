set -x
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_full_pytest.log
cat gpurun_out/r2_full_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_full_smoke.log 2>&1; tail -3 gpurun_out/r2_full_smoke.log
