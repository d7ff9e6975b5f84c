export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c14_pytest.log 2>&1; tail -6 gpurun_out/c14_pytest.log
python tools/run_one.py ecdsa+secp256k1 5 2>&1 | tail -4
python tools/run_one.py ecdsa 3 2>&1 | tail -2
timeout 600 python tools/stress.py 100 2>&1 | tail -4
