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


def dsu_system(link, zero_out=True, extra_rows=()):
    """BigMultModP inputs 5..13, BigLessThan inputs 15..20 (+ no output when zero_out).  `link` makes
    inputs[4:9] of the first and inputs[1:6] of the second the same sets in three different ways."""
    rows = [({2: 1}, {3: 1}, {4: 1})]
    for k in range(6):
        a, b = 8 + k, 15 + k                         # constraint_i[2][k+3], constraint_j[2][k]  (1-based k = 1..6)
        if link == "xy":
            rows.append(({}, {}, {a: 1, b: -1}))         # a - b = 0
        elif link == "chain":
            rows.append(({}, {}, {a: -1, 22 + k: 1}))    # a = t, t = b through a third wire
            rows.append(({}, {}, {22 + k: 1, b: -1}))
        elif link == "const":
            rows.append(({}, {}, {1: -(7 + k), a: 1}))   # a = 7 + k and 3 b = 3 (7 + k): both equal the same constant
            rows.append(({}, {}, {1: -3 * (7 + k), b: 3}))
        elif link == "none":
            rows.append(({}, {}, {a: 1, b: -1, 29: 0}))  # three stored keys: not looked at by the construction
    rows += list(extra_rows)
    m = MiniR1CS(rows, n_vars=30, known=[1, 2, 3], targets=[4])
    sp = [("BigMultModP", list(range(5, 14)), [14]), ("BigLessThan", list(range(15, 21)), [] if zero_out else [21])]
    return m, sp
