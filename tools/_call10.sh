export PYTHONUNBUFFERED=1
timeout 300 python tools/multi_check.py 2 > gpurun_out/c10_multi2.log 2>&1; echo "multi_check rc=$?"; tail -2 gpurun_out/c10_multi2.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --tile 8 > gpurun_out/c10_bench2.json 2> gpurun_out/c10_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
for f in ["gpurun_out/c10_bench2.json"]:
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "ms/step", j["ms_per_step"], "value", j["value"], "evals", j["config"]["evals_per_step"], "e2e ms", j["e2e"]["ms_per_step"], "match", j["config"]["matches_golden"])
    t=j["roofline_tiled"]; print(" tiled ms", t["ms_solve"], "dense ms", t["dense_rounds"]["ms"], "frac", t["dense_rounds"]["frac"], "ok", t["bitmap_is_base_repeated"])
PY
export ECNE_ENGINE_SO=$PWD/ecneproject_b200/libecne_b200_prof.so ECNE_DEBUG_PROF=3
ECNE_GPUS=2 timeout 300 python tools/run_tiled.py 8 2 > gpurun_out/c10_tiled8_n2.log 2>&1
grep "dense\|tile8" gpurun_out/c10_tiled8_n2.log | tail -20 | cut -c1-300
timeout 300 python tools/run_tiled.py 8 2 > gpurun_out/c10_tiled8_n1.log 2>&1
grep "dense\|tile8" gpurun_out/c10_tiled8_n1.log | tail -20 | cut -c1-300
