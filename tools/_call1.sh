set -x
export PYTHONUNBUFFERED=1
python tools/run_tiled.py 8 3 > gpurun_out/c1_tiled8_default.log 2>&1
ECNE_ENGINE_SO=$PWD/ecneproject_b200/libecne_b200_prof.so ECNE_DEBUG_PROF=3 python tools/run_tiled.py 8 2 > gpurun_out/c1_tiled8_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve -c 1 -f -o gpurun_out/c1_tiled8_ncu python tools/run_tiled.py 8 1 > gpurun_out/c1_ncu.log 2>&1
tail -3 gpurun_out/c1_tiled8_default.log
grep -c round gpurun_out/c1_tiled8_prof.log
