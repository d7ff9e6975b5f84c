export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c11_pytest.log 2>&1; tail -4 gpurun_out/c11_pytest.log
timeout 300 python tools/multi_check.py 2 > gpurun_out/c11_multi2.log 2>&1; echo "multi_check rc=$?"; grep "FAIL\|configs" gpurun_out/c11_multi2.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --tile 8 > gpurun_out/c11_bench2.json 2> gpurun_out/c11_bench2.err; echo "bench2 rc=$?"; tail -3 gpurun_out/c11_bench2.err
python - <<'PY'
import json
for f in ["gpurun_out/c11_bench2.json"]:
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "ms/step", j["ms_per_step"], "value", j["value"], "evals", j["config"]["evals_per_step"], "e2e ms", j["e2e"]["ms_per_step"], "match", j["config"]["matches_golden"], "sharded", j["config"]["sharded"])
    print(" forced", j["config"]["sharded_forced"])
    t=j["roofline_tiled"]; print(" tiled ms", t["ms_solve"], "dense ms", t["dense_rounds"]["ms"], "frac", t["dense_rounds"]["frac"], "ok", t["bitmap_is_base_repeated"])
PY
