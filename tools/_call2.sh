export PYTHONUNBUFFERED=1
python -m pytest tests -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1; tail -5 gpurun_out/c2_pytest.log
for so in libecne_b200.so libecne_b200_variant.so; do
  export ECNE_ENGINE_SO=$PWD/ecneproject_b200/$so
  echo "== $so" >> gpurun_out/c2_runs.log
  python tools/run_one.py ecdsa+secp256k1 4 >> gpurun_out/c2_runs.log 2>&1
  python tools/run_tiled.py 8 3 >> gpurun_out/c2_runs.log 2>&1
  python tools/run_tiled.py 16 3 >> gpurun_out/c2_runs.log 2>&1
done
cat gpurun_out/c2_runs.log
