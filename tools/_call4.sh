export PYTHONUNBUFFERED=1
python tools/multi_check.py 1 root/trivial_mult tornado/merkleTree root/bigmult86_3 > gpurun_out/c4_multi1.log 2>&1
cat gpurun_out/c4_multi1.log | cut -c1-700
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multi.py > gpurun_out/c4_pytest.log 2>&1; tail -8 gpurun_out/c4_pytest.log
