export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_abstraction.py -x -q > gpurun_out/c12_pytest.log 2>&1; tail -15 gpurun_out/c12_pytest.log
ECNE_HOST_PROF=1 timeout 300 python - > gpurun_out/c12_time.log 2>&1 <<'PY'
import time, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
from ecneproject_b200 import api, fixtures
from configs import CONFIGS
cfg = CONFIGS["ecdsa+secp256k1"]
for rep in range(3):
    t0 = time.perf_counter()
    main = api.readR1CS(fixtures.path(cfg["main"])); sub = api.readR1CS(fixtures.path(cfg["trusted"][0]))
    t1 = time.perf_counter()
    da = api.DeviceAbstraction(main)
    t2 = time.perf_counter()
    da.apply(cfg["trusted_names"][0], sub)
    t3 = time.perf_counter()
    h = da.upload(False)
    t4 = time.perf_counter()
    import ctypes as C
    res = api.SolveResult(main.n_vars)
    st = api._engine().ecne_solve_resident(h, C.byref(res.c))
    t5 = time.perf_counter()
    api._engine().ecne_free_resident(h); da.free()
    print(f"rep{rep}: read {1e3*(t1-t0):.1f} ms | begin (H2D unreduced) {1e3*(t2-t1):.1f} | apply {1e3*(t3-t2):.1f} | upload (classify) {1e3*(t4-t3):.1f} | solve {1e3*(t5-t4):.1f} | verdict {res.c.verdict} st {st} | file->verdict {1e3*(t5-t0):.1f} ms", flush=True)
PY
tail -12 gpurun_out/c12_time.log
