"""Shared test helpers: big-int <-> limb conversion, tiny hand-built problems, a pure-Python parser."""
import struct

import numpy as np

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def to_limbs(vals):
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        v %= 1 << 256
        for k in range(4):
            out[i, k] = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


def from_limbs(arr):
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    return [sum(int(r[k]) << (64 * k) for k in range(4)) for r in arr]


class MiniR1CS:
    """A hand-built constraint system in the ABI layout.  rows: list of (A, B, C) dicts wire->int."""

    def __init__(self, rows, n_vars, known, targets):
        seg, col, coef = [0], [], []
        for forms in rows:
            for form in forms:
                items = list(form.items())
                if not items:
                    items = [(1, 0)]  # the parser stores an explicit zero on key 1 (ParseR1CS.jl:113-115)
                for w, c in items:
                    col.append(w)
                    coef.append(c % P)
                seg.append(len(col))
        self.n_rows = len(rows)
        self.n_vars = n_vars
        self.nnz = len(col)
        self.seg_ptr = np.asarray(seg, dtype=np.uint64)
        self.col = np.asarray(col, dtype=np.uint32)
        self.coef = to_limbs(coef) if coef else np.zeros((0, 4), dtype=np.uint64)
        self.known = np.asarray(known, dtype=np.uint32)
        self.targets = np.asarray(targets, dtype=np.uint32)


def py_read_r1cs(path):
    """Independent restatement of ParseR1CS.readR1CS (ParseR1CS.jl:50-124) in plain Python."""
    b = open(path, "rb").read()
    assert struct.unpack_from("<I", b, 4)[0] == 1
    nsec = struct.unpack_from("<I", b, 8)[0]
    assert nsec == 3
    cur, starts = 12, {}
    for _ in range(nsec):
        ty = struct.unpack_from("<I", b, cur)[0]
        sz = struct.unpack_from("<Q", b, cur + 4)[0]
        starts[ty] = cur
        cur += 12 + sz
    s1 = starts[1] + 12
    fs = struct.unpack_from("<I", b, s1)[0]
    s1 += 4 + fs
    n_wires, pub_out, pub_in, prv_in = struct.unpack_from("<IIII", b, s1)
    n_cons = struct.unpack_from("<I", b, s1 + 24)[0]
    s2 = starts[2] + 12
    rows = []
    for _ in range(n_cons):
        forms = []
        for _ in range(3):
            n = struct.unpack_from("<I", b, s2)[0]
            s2 += 4
            d = {}
            for _ in range(n):
                idx = struct.unpack_from("<I", b, s2)[0]
                d[idx + 1] = int.from_bytes(b[s2 + 4:s2 + 36], "little") % P
                s2 += 36
            if n == 0:
                d[1] = 0
            forms.append(d)
        rows.append(forms)
    known = [1] + list(range(2 + pub_out, 2 + pub_out + pub_in + prv_in))
    targets = list(range(2, 2 + pub_out))
    return rows, known, targets, n_wires + 1
