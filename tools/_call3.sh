export PYTHONUNBUFFERED=1
python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.log 2>&1; tail -8 gpurun_out/c3_pytest.log
python tools/run_one.py ecdsa+secp256k1 4 > gpurun_out/c3_runs.log 2>&1
python tools/run_tiled.py 8 3 >> gpurun_out/c3_runs.log 2>&1
cat gpurun_out/c3_runs.log
