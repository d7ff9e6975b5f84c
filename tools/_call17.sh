export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c17_pytest.log 2>&1; tail -6 gpurun_out/c17_pytest.log
python tools/run_one.py ecdsa+secp256k1 4 2>&1 | tail -2
python tools/run_one.py ecdsa 3 2>&1 | tail -1
timeout 600 python tools/stress.py 60 2>&1 | tail -3
