export PYTHONUNBUFFERED=1
python -m pytest tests -m gpu -x -q > gpurun_out/c5_pytest.log 2>&1; tail -8 gpurun_out/c5_pytest.log
