export PYTHONUNBUFFERED=1
export ECNE_ENGINE_SO=$PWD/ecneproject_b200/libecne_b200_prof.so ECNE_DEBUG_PROF=3
ECNE_GPUS=2 timeout 300 python tools/run_tiled.py 8 2 > gpurun_out/c9_tiled8_n2.log 2>&1
grep -v "^\[round\]" gpurun_out/c9_tiled8_n2.log | tail -45 | cut -c1-300
ECNE_GPUS=2 timeout 300 python tools/run_one.py ecdsa+secp256k1 2 > gpurun_out/c9_ecdsa_n2.log 2>&1
grep "dense \|rep" gpurun_out/c9_ecdsa_n2.log | tail -24 | cut -c1-300
