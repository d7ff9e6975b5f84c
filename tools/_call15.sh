export PYTHONUNBUFFERED=1
export ECNE_ENGINE_SO=$PWD/ecneproject_b200/libecne_b200_prof.so ECNE_DEBUG_PROF=3
python tools/run_one.py ecdsa+secp256k1 2 > gpurun_out/c15_ecdsa_prof.log 2>&1
grep "prof\]\|p2scan\] outer [1-5]:\|^\[dense\|rep1" gpurun_out/c15_ecdsa_prof.log | tail -30 | cut -c1-330
