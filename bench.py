#!/usr/bin/env python3
"""bench.py — headline benchmark: constraint-evals/s on ecdsa.r1cs (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank/GPU)
    python bench.py --impl reference --steps K --warmup W    (the CPU path on the host cores)

One step = one full run of the hot path (SolveConstraintsSymbolic) on the workload
`ecdsa.r1cs + trusted secp256k1.r1cs` (BASELINE.json configs[4]; 1 092 639 rows -> 694 264 after
abstraction; the configuration the `metric` is quoted on — it fits one GPU).

  value   constraint-evals/s (rows visited by the solve kernel / time), inputs resident in HBM
          (ecne_upload once, ecne_solve_resident per step), timed with CUDA events on the engine's own
          stream (ecne_result.ms_device: reset + the persistent solve kernel + verdict + D2H of the
          bitmaps), max over ranks.  L2 is flushed between timed steps.
  e2e     the same metric through the reference-facing C-ABI call ecne_solve() with HOST buffers:
          H2D of the CSR from pinned memory + classify + the solve + D2H of the bitmaps, per step.
  roofline  the solve kernel (k_solve, one launch per step): algorithmic bytes of the compact sweep
          layout (32 B row record + 1 B per term) x rows it visited / its CUDA-event
          time.  roofline.dense_rounds is the same for the dense Jacobi rounds alone (in-kernel
          clock64), the only part of the solve that streams rows.
  N > 1   the rows are sharded by stored-term balance (ecne_shard_rows), wire state is replicated and
          the update records of the sharded (dense) rounds are exchanged over NVLink inside the solve
          kernel: the SAME problem is split, so "scaling" is "strong" (at every N, N = 1 included).
          Every rank hashes its `unique` bitmap and compares it with the committed oracle golden
          (config.sha_unique / config.matches_golden); a mismatch on any rank fails the run.
          scale_tiled: the same for the K-times tiled workload (the bandwidth regime).
  cpu_baseline  oracle/ (a single-threaded C++ port of the reference's Julia) on the same workload,
          timed on this box's host, rank 0, N=1 only.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = {"main": "ecdsa.r1cs", "trusted": ["secp256k1.r1cs"], "trusted_names": ["Secp256k1AddUnequal"]}
WORKLOAD_NAME = "ecdsa.r1cs + trusted secp256k1.r1cs (Secp256k1AddUnequal), via solveWithTrustedFunctions"
METRIC = "constraint_evals_per_sec_ecdsa"
UNIT = "constraint-evals/s"


PREP = {}


def load_problem():
    """readR1CS + abstraction through the native host fast path (include/ecne_host.h); outside every timed
    region — its wall time is reported as config.host_prep_seconds for context."""
    from ecneproject_b200 import api, fixtures
    t0 = time.perf_counter()
    reduced, specials, main = api.prepare(fixtures.path(WORKLOAD["main"]),
                                          [fixtures.path(t) for t in WORKLOAD["trusted"]],
                                          WORKLOAD["trusted_names"])
    PREP["read_and_abstraction"] = time.perf_counter() - t0
    return reduced, specials, main


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def golden_sha():
    """SHA-256 of the packed `unique` bitmap the oracle produces on the headline workload (committed pin)."""
    try:
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_goldens.json")))
        return g["ecdsa+secp256k1"]["sha_unique"]
    except Exception:
        return None


def measured_traffic(key):
    """DRAM bytes per k_solve launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu
    capture of this round, profiles/r02_traffic.json (written by tools/ncu_traffic.py from the .ncu-rep)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        return t[key]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for i, n in enumerate(names):
                if f[5 + i].lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


def algorithmic_bytes_per_eval(n_rows, nnz_nonzero):
    # sweep layout: one 32-byte row record (flags + <= 6 inline wire ids) per swept row
    # + 1 gathered state byte per non-zero term.  Bytes that are not moved are not credited: a row the
    # live mask has retired is neither counted as an eval nor as bytes.
    return 32.0 + nnz_nonzero / float(n_rows)


def tile_problem(np, reduced, specials, main, K):
    """Block-diagonal tiling of the workload K times (SURVEY.md §8d "S-K"): wire ids offset by
    k*(n_vars-1), wire 1 shared.  The working set (K x 22 MB of row records) leaves the L2."""
    V = main.n_vars
    seg, col, coef = reduced.seg_ptr.astype(np.int64), reduced.col.astype(np.int64), reduced.coef

    class T:
        pass
    t = T()
    t.n_rows = reduced.n_rows * K
    t.nnz = reduced.nnz * K
    t.seg_ptr = np.concatenate([[0]] + [seg[1:] + k * seg[-1] for k in range(K)]).astype(np.uint64)
    t.col = np.concatenate([np.where(col == 1, 1, col + k * (V - 1)) for k in range(K)]).astype(np.uint32)
    t.coef = np.tile(coef, (K, 1))
    nv = 1 + (V - 1) * K

    def off(a, k):
        a = np.asarray(a, dtype=np.int64)
        return np.where(a == 1, 1, a + k * (V - 1))
    known = np.unique(np.concatenate([off(main.known, k) for k in range(K)]))
    targets = np.concatenate([off(main.targets, k) for k in range(K)])
    sp = []
    base = specials.as_list()
    for k in range(K):
        for n, i, o in base:
            sp.append((n, off(i, k).tolist(), off(o, k).tolist()))
    return t, sp, known, targets, nv


def run_reference(args):
    """--impl reference: the reference's CPU path.  Julia is not in the image, so this is the
    oracle port (kind 'port'), single-threaded like the reference, on the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    reduced, specials, main = load_problem()
    lib = oracle_lib.lib()
    # bounded sample: a full solve is ~4.5 s of one host core; fall back to the first outer round (the
    # queue loop over every row + the three whole-set sweeps) when K + W full solves would take minutes
    full = (args.steps + args.warmup) <= 40
    lib.ecne_oracle_set_max_outer(0 if full else 1)
    sample = ("full solve (28 outer rounds, 59.9 M evals)" if full else
              "first outer round only (queue loop + the three sweeps)")
    times, evals = [], 0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = oracle_lib.solve(reduced, specials, main.known, main.targets, main.n_vars, False, full_state=False)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
            evals = int(res.c.constraint_evals)
    lib.ecne_oracle_set_max_outer(0)
    ms = 1e3 * sum(times) / len(times)
    value = evals / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u256 (4x64-bit limbs, BN254 scalar field)",
        "data": "real circuit (reference fixture ecdsa.r1cs), no synthetic data needed",
        "config": {"workload": WORKLOAD_NAME, "rows": reduced.n_rows, "wires": main.n_vars,
                   "evals_per_step": evals},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count(), "cpu_model": cpu_model()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tile", type=int, default=16,
                    help="also time a K-times tiled copy of the workload (0 = skip); 16 = SURVEY.md §8d's S16: 11.1 M rows, "
                         "355 MB of row records, far beyond the L2")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    from ecneproject_b200 import api, _abi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the engine has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local)
    dist = None
    lib = _abi.engine_lib()
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on stdout when a communicator is created: keep stdout for the one
        # JSON line of the contract
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        import torch.distributed as dist
        from ecneproject_b200 import dist as edist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        edist.init_from_torch(local)   # ecne_init + the engine's own communicator / peer mappings
    else:
        st = lib.ecne_init(local)
        if st != 0:
            raise RuntimeError(lib.ecne_last_error().decode())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    reduced, specials, main = load_problem()

    # ---- pinned host copies of the inputs for the e2e leg ---------------------------------------
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t

    class PinnedR1CS:
        pass

    pr = PinnedR1CS()
    pr._keep = [pinned(reduced.seg_ptr.view(np.int64)), pinned(reduced.col.view(np.int32)),
                pinned(reduced.coef.reshape(-1).view(np.int64))]
    pr.n_rows, pr.nnz = reduced.n_rows, reduced.nnz
    pr.seg_ptr = pr._keep[0].numpy().view(np.uint64)
    pr.col = pr._keep[1].numpy().view(np.uint32)
    pr.coef = pr._keep[2].numpy().view(np.uint64).reshape(-1, 4)
    ph = api.ProblemHandle(pr, specials, main.known, main.targets, main.n_vars, False)
    # ProblemHandle copies non-contiguous inputs only; make sure the pinned buffers are what it points at
    assert ph.keep[0].ctypes.data == pr.col.ctypes.data
    h2d_bytes_full = pr.seg_ptr.nbytes + pr.col.nbytes + pr.coef.nbytes + 4 * (len(main.known) + len(main.targets))
    # The e2e leg hands the rows over in the COMPACT form of include/ecne_abi.h (ABI 3): 32-bit offsets, one class byte
    # per term (0, 1, p - 1, other) and 32-byte limbs only for the other values — what a host that flattens
    # Dict{Int, GFElem} can emit as cheaply as the full form.  Pinned like the full form; the full form is timed too.
    t0c = time.perf_counter()
    cls_, other_, term_ = api.compact_coef(reduced.coef)
    compact_prep_s = time.perf_counter() - t0c
    pc = PinnedR1CS()
    pc._keep = [pinned(np.asarray(reduced.seg_ptr).astype(np.uint32).view(np.int32)), pinned(cls_),
                pinned(other_.reshape(-1).view(np.int64)), pinned(term_.view(np.int32))]
    ph_c = api.ProblemHandle(pr, specials, main.known, main.targets, main.n_vars, False)
    ph_c.c.seg_ptr = None
    ph_c.c.coef = None
    ph_c.c.seg_ptr32 = C.cast(pc._keep[0].data_ptr(), _abi.u32p)
    ph_c.c.coef_class = C.cast(pc._keep[1].data_ptr(), _abi.u8p)
    ph_c.c.coef_other = C.cast(pc._keep[2].data_ptr(), _abi.u64p)
    ph_c.c.coef_other_term = C.cast(pc._keep[3].data_ptr(), _abi.u32p)
    ph_c.c.n_coef_other = len(term_)
    h2d_bytes = (4 * (3 * reduced.n_rows + 1) + pr.col.nbytes + cls_.nbytes + other_.nbytes + term_.nbytes +
                 4 * (len(main.known) + len(main.targets)))
    res = api.SolveResult(main.n_vars, full_state=False)
    d2h_bytes = 2 * res.unique_bits.nbytes + 4 * 8

    # ---- resident leg ------------------------------------------------------------------------------
    handle = C.c_void_p()
    st = lib.ecne_upload(C.byref(ph.c), C.byref(handle))
    if st != 0:
        raise RuntimeError(lib.ecne_last_error().decode())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def l2_flush():
        flush.fill_(rank + 1)

    def step_resident():
        st = lib.ecne_solve_resident(handle, C.byref(res.c))
        if st != 0:
            raise RuntimeError(lib.ecne_last_error().decode())
        return res.c

    # clocks / throttle reasons are sampled from the first warm-up step to the end of the e2e leg (the timed
    # regions are tens of milliseconds: too short for nvidia-smi's sampling period on their own)
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        l2_flush()
        barrier()
        step_resident()
    t_wall = t_dev = 0.0
    sweep_ms = solve_ms = 0.0
    launches = 0
    dense_cycles = dense_evals = dense_rounds = 0
    for _ in range(args.steps):
        l2_flush()
        barrier()
        t0 = time.perf_counter()
        c = step_resident()
        torch.cuda.synchronize()
        t_wall += time.perf_counter() - t0
        t_dev += c.ms_device          # CUDA events on the engine's stream, around the whole call
        sweep_ms += c.ms_sweep
        solve_ms += c.ms_solve
        launches += int(c.sweep_launches)
        evals, rounds, outer = int(c.constraint_evals), int(c.inner_rounds), int(c.outer_rounds)
        rule_evals = int(c.rule_evals)
        dense_cycles += int(c.dense_cycles)
        dense_evals, dense_rounds = int(c.dense_evals), int(c.dense_rounds)
    barrier()
    verdict = bool(res.c.verdict)
    n_unique = int(res.c.n_unique)
    # ---- parity, on EVERY rank: the determined-variable set against the committed oracle pin --------------
    import hashlib
    sha_unique = hashlib.sha256(res.unique_bytes()).hexdigest()
    want_sha = golden_sha()
    matches = int(want_sha is not None and sha_unique == want_sha and verdict)
    if dist is not None:
        m = torch.tensor([matches], device="cuda", dtype=torch.int64)
        dist.all_reduce(m, op=dist.ReduceOp.MIN)
        matches_all = int(m[0])
    else:
        matches_all = matches

    # ---- N > 1: the same workload with the dense sweeps FORCED onto row ranges ("shard_min_rows" = 0) ------------
    # By default a problem of this size is solved by every GPU in full (a sharded round costs a cross-GPU barrier
    # that a 694 k-row sweep does not earn back); this leg runs the sharded data path on the named workload anyway,
    # so that its bit-parity and its cost are on the driver's record.
    sharded_forced = None
    sharded_default = bool(res.c.sharded)
    if world > 1 and not sharded_default:
        assert lib.ecne_set_option(b"shard_min_rows", 0) == 0
        h_s = C.c_void_p()
        st = lib.ecne_upload(C.byref(ph.c), C.byref(h_s))
        if st != 0:
            raise RuntimeError(lib.ecne_last_error().decode())
        res_s = api.SolveResult(main.n_vars, full_state=False)
        ts = 0.0
        for i in range(2 + 5):
            l2_flush()
            barrier()
            st = lib.ecne_solve_resident(h_s, C.byref(res_s.c))
            if st != 0:
                raise RuntimeError(lib.ecne_last_error().decode())
            if i >= 2:
                ts += res_s.c.ms_device
        barrier()
        lib.ecne_free_resident(h_s)
        assert lib.ecne_set_option(b"shard_min_rows", 2000000) == 0
        ok_s = int(hashlib.sha256(res_s.unique_bytes()).hexdigest() == want_sha and bool(res_s.c.sharded))
        tt = torch.tensor([ts / 5], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ee = torch.tensor([int(res_s.c.constraint_evals), -ok_s], device="cuda", dtype=torch.int64)
        dist.all_reduce(ee, op=dist.ReduceOp.SUM)
        sharded_forced = {"ms_per_step": float(tt[0]), "evals_per_step": int(ee[0]),
                          "matches_golden_on_every_rank": int(ee[1]) == -world,
                          "sha_unique": hashlib.sha256(res_s.unique_bytes()).hexdigest(),
                          "how": "ecne_set_option(shard_min_rows, 0): dense sweeps split by row range, records exchanged over NVLink"}
        matches_all = int(matches_all and sharded_forced["matches_golden_on_every_rank"])

    # ---- e2e leg: host buffers in, host buffers out -------------------------------------------------
    res2 = api.SolveResult(main.n_vars, full_state=False)
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_leg(handle_c):
        for _ in range(2):
            barrier()
            lib.ecne_solve(C.byref(handle_c), C.byref(res2.c))
        tt_ = 0.0
        for _ in range(e2e_steps):
            l2_flush()
            barrier()
            t0 = time.perf_counter()
            st = lib.ecne_solve(C.byref(handle_c), C.byref(res2.c))
            torch.cuda.synchronize()
            tt_ += time.perf_counter() - t0
            if st != 0:
                raise RuntimeError(lib.ecne_last_error().decode())
        barrier()
        assert res2.unique_bits.tobytes() == res.unique_bits.tobytes()
        return tt_, res2.c.ms_h2d, res2.c.ms_classify

    t_e2e_full, h2d_ms_full, classify_ms_full = e2e_leg(ph.c)     # full 32-byte coefficients, 64-bit offsets
    t_e2e, h2d_ms, classify_ms = e2e_leg(ph_c.c)                  # compact form: the e2e number of the line
    # ---- file -> verdict with abstraction() on the device (N = 1): parse on the host cores, upload the UNREDUCED
    # system once, abstract + classify where it lies, solve (include/ecne_abi.h "abstraction() on the device")
    f2v = None
    if world == 1:
        from ecneproject_b200 import fixtures
        f2v_t, f2v_parts = [], None
        m_ = subs_ = da = res3 = None
        for _ in range(4):
            m_ = subs_ = da = res3 = None   # (the circuits of the run before are released outside the timed region)
            import gc
            gc.collect()
            t0 = time.perf_counter()
            m_, subs_ = api.read_and_prepare(fixtures.path(WORKLOAD["main"]), [fixtures.path(t) for t in WORKLOAD["trusted"]],
                                             WORKLOAD["trusted_names"])
            t1 = time.perf_counter()
            da = api.DeviceAbstraction(m_)
            for nm_, sub_ in subs_:
                da.apply(nm_, sub_)
            t2 = time.perf_counter()
            h_ = da.upload(False)
            res3 = api.SolveResult(m_.n_vars, full_state=False)
            st = lib.ecne_solve_resident(h_, C.byref(res3.c))
            t3 = time.perf_counter()
            lib.ecne_free_resident(h_)
            da.free()
            if st != 0:
                raise RuntimeError(lib.ecne_last_error().decode())
            assert res3.unique_bits.tobytes() == res.unique_bits.tobytes()
            f2v_t.append(t3 - t0)
            f2v_parts = {"read_s": t1 - t0, "upload_and_abstraction_on_device_s": t2 - t1, "classify_and_solve_s": t3 - t2}
        f2v = {"seconds": sum(f2v_t[1:]) / len(f2v_t[1:]), "runs": f2v_t, "last_run": f2v_parts,
               "how": "readR1CS on the host cores (the trusted circuits read and prepared on a second thread meanwhile), ecne_abstract_begin/apply/upload (abstraction and classification on the "
                      "GPU, the reduced system never crosses PCIe), ecne_solve_resident; bitmap equal to the resident leg's"}
    clocks = sampler.stop()
    e2e_ms = 1e3 * t_e2e / e2e_steps
    e2e_ms_full = 1e3 * t_e2e_full / e2e_steps
    lib.ecne_free_resident(handle)

    ms_step = t_dev / args.steps
    ms_wall = 1e3 * t_wall / args.steps
    total_evals = evals
    if dist is not None:   # max over ranks of the times, sum over ranks of the rows each rank visited
        t = torch.tensor([ms_step, e2e_ms, ms_wall, e2e_ms_full], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms, ms_wall, e2e_ms_full = float(t[0]), float(t[1]), float(t[2]), float(t[3])
        e = torch.tensor([evals], device="cuda", dtype=torch.int64)
        dist.all_reduce(e, op=dist.ReduceOp.SUM)
        total_evals = int(e[0])
    value = total_evals / (ms_step / 1e3)
    e2e_value = total_evals / (e2e_ms / 1e3)

    # ---- roofline of the solve kernel ---------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "6650 GB/s (of fallback)"
    nnz_nonzero = int(np.count_nonzero(reduced.coef.any(axis=1)))
    b_eval = algorithmic_bytes_per_eval(reduced.n_rows, nnz_nonzero)
    sweep_ms_step = sweep_ms / args.steps
    achieved = b_eval * evals / (sweep_ms_step / 1e3) / 1e9 if sweep_ms_step > 0 else 0.0
    sm_mhz = clocks.get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    dense_ms = dense_cycles / args.steps / (sm_mhz * 1e3) if sm_mhz else 0.0
    dense_ach = b_eval * dense_evals / (dense_ms / 1e3) / 1e9 if dense_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_solve (the whole fixpoint, one persistent cooperative launch)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of the one k_solve launch of a solve, read from the
                # ncu capture of this round (profiles/r02_traffic.json <- tools/ncu_traffic.py); None if absent
                "traffic": (measured_traffic("ecdsa") or {}).get("dram_bytes") if world == 1 else None,
                "traffic_source": (measured_traffic("ecdsa") or {}).get("source") if world == 1 else None,
                "peak_source": peak_src, "bytes_per_eval": b_eval, "evals_in_kernel_per_step": evals,
                "jacobi_rounds_per_step": rounds, "kernel_ms_per_step": sweep_ms_step, "launches_per_step": 1,
                "dense_rounds": {"rounds": dense_rounds, "evals": dense_evals, "ms": dense_ms,
                                 "achieved": dense_ach, "frac": dense_ach / peak,
                                 "how": "clock64 in block 0 around the dense rounds / SM clock sampled by nvidia-smi"},
                "note": "the named workload is a dependency chain: %d dependent Jacobi rounds, all but %d of them "
                        "driven by <= 100 update records; the kernel is bound by the latency of a round, not by "
                        "HBM (DESIGN.md §5); see roofline_tiled for the bandwidth regime" % (rounds, dense_rounds)}

    # ---- the same kernel on a tiled copy whose working set does not fit the 126 MB L2 ---------------
    tiled = None
    if args.tile and args.tile > 1:
        K = args.tile
        t, sp_t, known_t, targets_t, nv_t = tile_problem(np, reduced, specials, main, K)
        ph_t = api.ProblemHandle(t, sp_t, known_t, targets_t, nv_t, False)
        res_t = api.SolveResult(nv_t, full_state=False)
        h_t = C.c_void_p()
        st = lib.ecne_upload(C.byref(ph_t.c), C.byref(h_t))
        if st != 0:
            raise RuntimeError(lib.ecne_last_error().decode())
        for _ in range(2):
            barrier()
            lib.ecne_solve_resident(h_t, C.byref(res_t.c))
        tsw = tso = 0.0
        reps = 3
        for _ in range(reps):
            l2_flush()
            barrier()
            st = lib.ecne_solve_resident(h_t, C.byref(res_t.c))
            if st != 0:
                raise RuntimeError(lib.ecne_last_error().decode())
            tsw += res_t.c.ms_sweep
            tso += res_t.c.ms_solve
        barrier()
        lib.ecne_free_resident(h_t)
        ev_t, dev_t = int(res_t.c.constraint_evals), int(res_t.c.dense_evals)
        d_ms_t = int(res_t.c.dense_cycles) / (sm_mhz * 1e3) if sm_mhz else 0.0
        # the tiled bitmap is the base bitmap repeated: wire 1 shared, wires 2.. of copy k behind those of copy k-1
        base_bits = np.unpackbits(np.frombuffer(res.unique_bytes(), dtype=np.uint8), bitorder="little")[:main.n_vars]
        want_bits = np.concatenate([base_bits[:1]] + [base_bits[1:]] * K)
        got_bits = np.unpackbits(np.frombuffer(res_t.unique_bytes(), dtype=np.uint8), bitorder="little")[:nv_t]
        ok_t = int(bool(np.array_equal(want_bits, got_bits)) and bool(res_t.c.verdict) == verdict)
        tsw, tso = tsw / reps, tso / reps
        if dist is not None:   # slowest rank's times, all ranks' row visits, every rank's bitmap check
            tt = torch.tensor([tsw, tso, d_ms_t], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            tsw, tso, d_ms_t = float(tt[0]), float(tt[1]), float(tt[2])
            ee = torch.tensor([ev_t, dev_t, -ok_t], device="cuda", dtype=torch.int64)
            dist.all_reduce(ee, op=dist.ReduceOp.SUM)
            ev_t, dev_t, ok_t = int(ee[0]), int(ee[1]), int(int(ee[2]) == -world)
        ach_t = b_eval * ev_t / (tsw / 1e3) / 1e9
        d_ach_t = b_eval * dev_t / (d_ms_t / 1e3) / 1e9 if d_ms_t > 0 else 0.0
        tr_t = measured_traffic("tiled%d" % K) if world == 1 else None
        tiled = {"tile": K, "n_gpus": world, "sharded": bool(res_t.c.sharded), "rows": t.n_rows, "row_record_bytes": 32 * t.n_rows,
                 "evals_per_step": ev_t,
                 "value": ev_t / (tso / 1e3), "unit": UNIT, "ms_solve": tso,
                 "kernel_ms_per_step": tsw, "achieved": ach_t, "peak": peak * world, "frac": ach_t / (peak * world),
                 "traffic": (tr_t or {}).get("dram_bytes"), "traffic_source": (tr_t or {}).get("source"),
                 "dense_rounds": {"rounds": int(res_t.c.dense_rounds), "evals": dev_t,
                                  "ms": d_ms_t, "achieved": d_ach_t, "frac": d_ach_t / (peak * world),
                                  "how": "clock64 in block 0 of the slowest rank around the dense rounds, barrier "
                                         "and cross-GPU exchange included; rows visited summed over the ranks"},
                 "bitmap_is_base_repeated": bool(ok_t)}
        del ph_t, res_t, t

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        # the SAME problem at every N (rows sharded over the GPUs): strong scaling, N = 1 included
        "scaling": "strong", "vs_baseline": None,
        "dtype": "u256 (4x64-bit limbs, BN254 scalar field; sweep works on u8/u32 state)",
        "data": "real circuit (reference fixture ecdsa.r1cs), no synthetic data needed",
        "config": {"workload": WORKLOAD_NAME, "rows": reduced.n_rows, "rows_before_abstraction": main.n_rows,
                   "wires": main.n_vars, "nnz": nnz_nonzero, "evals_per_step": total_evals,
                   "outer_rounds": outer, "jacobi_rounds": rounds, "verdict": verdict, "n_unique": n_unique,
                   # parity, every rank: SHA-256 of the packed `unique` bitmap against the committed oracle pin
                   "sha_unique": sha_unique, "golden_sha_unique": want_sha, "matches_golden": bool(matches_all),
                   "evals_note": "row visits actually executed: a row the live mask / the open-row bitmap has retired "
                                 "is neither visited nor counted (the linear-system sweep of round 1 counted every live "
                                 "row in every outer round: 15.0 M visits for the same verdict; two rounds that used to be "
                                 "dense sweeps are frontier-driven now) - evals/s FALLS when the engine avoids work, the time "
                                 "per verdict is the like-for-like figure",
                   # the same verdict's work as this engine counted it at the end of round 1 (15 031 701 row visits, FIFO
                   # reference: 59 920 653) over today's time: what `value` would read had the schedule not changed
                   "value_at_round1_evals": 15031701 / (ms_step / 1e3),
                   "rule_evals_per_step": rule_evals,
                   "l2": "flushed between timed steps (256 MB fill)",
                   "parallelism": "1 GPU" if world == 1 else (
                       f"{world} GPUs: rows sharded by stored-term balance, wire state replicated, the dense sweeps' update "
                       "records exchanged over NVLink inside the solve kernel (DESIGN.md §7)" if sharded_default else
                       f"{world} GPUs, {reduced.n_rows} rows < shard_min_rows: every GPU solves the whole problem, no exchange "
                       "(sharding this workload is slower than one GPU, see config.sharded_forced; the tiled leg is sharded)"),
                   "sharded": sharded_default, "sharded_forced": sharded_forced,
                   "timing": "CUDA events on the engine's stream around each whole call (ecne_result.ms_device), max over ranks",
                   "wall_ms_per_step": ms_wall, "device_ms_solve_per_step": solve_ms / args.steps,
                   "host_prep_seconds": PREP.get("read_and_abstraction"),
                   "seconds_to_verdict_resident": ms_wall / 1e3},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms, "ms_h2d": h2d_ms,
                "ms_classify": classify_ms, "seconds_to_verdict": e2e_ms / 1e3,
                "input_form": "compact (include/ecne_abi.h, ABI 3): 32-bit offsets, wire ids, one class byte per stored term "
                              "{0, 1, p-1, other} + 32-byte limbs of the other values; expanded on the device into the same "
                              "arrays the full form is copied into",
                "compact_prep_seconds_untimed": compact_prep_s,
                "upload": ("every rank copies 1/%d of the rows over its own PCIe link, ncclAllGather of the slices over NVLink "
                           "(h2d_bytes_per_step is the sum over the ranks: the problem once)" % world) if world > 1 else "one GPU",
                # the same call with the full form (64-bit offsets, 32-byte limbs for every stored term)
                "full_form": {"ms_per_step": e2e_ms_full, "value": total_evals / (e2e_ms_full / 1e3),
                              "h2d_bytes_per_step": h2d_bytes_full, "ms_h2d": h2d_ms_full, "ms_classify": classify_ms_full},
                # the whole user-visible pipeline from the .r1cs files: with abstraction() on the device (file_to_verdict),
                # and with the host library's abstraction followed by ecne_solve (…_host_abstraction)
                "seconds_file_to_verdict": f2v["seconds"] if f2v else None,
                "file_to_verdict": f2v,
                "seconds_file_to_verdict_host_abstraction": (PREP.get("read_and_abstraction") or 0.0) + e2e_ms / 1e3,
                "timing": "host clock around ecne_solve() (pinned host buffers in, host bitmaps out), max over ranks"},
        "gpu_launches": launches,
        "roofline": roofline,
        "clocks": clocks,
    }
    if tiled is not None:
        line["roofline_tiled"] = tiled
        if world > 1:
            line["scale_tiled"] = tiled
    # ---- CPU baseline (rank 0, N=1 only) ------------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        t0 = time.perf_counter()
        o = oracle_lib.solve(reduced, specials, main.known, main.targets, main.n_vars, False, full_state=False)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": int(o.c.constraint_evals) / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "one full solve of the same workload by oracle/ (C++ port of the Julia), "
                                          f"{int(o.c.constraint_evals)} evals in {dt:.2f} s",
                                "seconds_to_verdict": dt, "host_cpus": os.cpu_count(), "cpu_model": cpu_model(),
                                "matches_gpu_bitmap": o.unique_bits.tobytes() == res.unique_bits.tobytes()}
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    if not matches_all:
        sys.stderr.write("bench.py: the `unique` bitmap of some rank differs from the committed oracle golden\n")
        return 3
    if tiled is not None and not tiled["bitmap_is_base_repeated"]:
        sys.stderr.write("bench.py: the tiled workload's bitmap is not the base bitmap repeated\n")
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
