export PYTHONUNBUFFERED=1
ECNE_HOST_PROF=1 python tools/file_to_verdict.py ecdsa+secp256k1 4 2>&1 | grep "rep\|ecne dev" | tail -6
timeout 600 python -m pytest tests/test_gpu_abstraction.py -x -q 2>&1 | tail -2
