export PYTHONUNBUFFERED=1
for so in libecne_b200.so libecne_b200_variant.so; do
  export ECNE_ENGINE_SO=$PWD/ecneproject_b200/$so
  echo "== $so"
  python tools/run_one.py ecdsa+secp256k1 4 2>&1 | tail -2
  python tools/run_tiled.py 16 2 2>&1 | tail -1
done
