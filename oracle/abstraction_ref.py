"""oracle/abstraction_ref.py — TEST INFRASTRUCTURE, not product: a plain-Python restatement of the reference's
`abstraction()` (/root/reference/src/R1CSConstraintSolver.jl:237-395) with its helpers `hash_r1cs_equation`
(:228-235) and `checkNonZeroValues` (:205-226), on per-row dictionaries exactly like the Julia (`R1CSEquation.a/b/c`
= wire -> coefficient).  It shares no code and no data structure with csrc/host_r1cs.cpp (CSR, flat signatures) and
is what tests/test_host_side.py checks `ecne_abstraction` against.  Only tests/ may import it.

Parity status: Julia cannot run here, so this restatement is pinned only by what the reference asserts about the
pass — the match counts of bench/bench_abstraction.jl:16,24 and the verdicts of the abstracted configurations
(test/runtests.jl:25,30,35, examples/ecdsa_secp_abstraction.jl:4).  One choice is unpinned in the reference itself:
`sort(collect(appearance_map), by=signature)` (:334-335) is a stable sort of a Dict's iteration order, so wires with
IDENTICAL signatures come out in hash order; here (and in the host library) ties are broken by wire id.
"""


def hash_r1cs_equation(row):
    """:228-235 — the three sorted coefficient lists concatenated, zeros dropped; the list itself stands for its hash."""
    a, b, c = row
    l = sorted(a.values()) + sorted(b.values()) + sorted(c.values())
    return tuple(x for x in l if x != 0)


def check_nonzero_values(m1, m2):
    """:205-226 — same multiset of non-zero values."""
    def counter(m):
        out = {}
        for v in m.values():
            if v != 0:
                out[v] = out.get(v, 0) + 1
        return out
    return counter(m1) == counter(m2)


class KeyErrorAt(KeyError):
    pass


def abstraction(function_name, constraints, known_inputs, sub_equation, known_outputs):
    """:237-395.  constraints / sub_equation: lists of (A, B, C) dicts.  Returns (special_cons, red_cons)."""
    n = len(sub_equation)
    hashed_constraints = [hash_r1cs_equation(x) for x in constraints]          # :252
    hashed_sub_equation = [hash_r1cs_equation(x) for x in sub_equation]        # :253
    candidates = []
    for i in range(len(constraints) - n + 1):                                  # :259-270 (first n-1 rows only)
        if all(hashed_constraints[i + j] == hashed_sub_equation[j] for j in range(n - 1)):
            candidates.append(i)
    appearance_map_orig = {}                                                   # :276-292
    counter = 1
    for j in range(n):
        for eq in sub_equation[j]:
            for wire, coef in eq.items():
                if coef != 0:
                    appearance_map_orig.setdefault(wire, []).append((counter, coef))
            counter += 1
    by_sig = lambda kv: (kv[1], kv[0])   # signature, then wire id (the unpinned tie order, see the header)
    l2 = sorted(appearance_map_orig.items(), key=by_sig)
    matches = []
    for i in candidates:                                                       # :293-352
        works = True
        appearance_map_cur = {}
        app_counter = 0
        for j in range(n):
            for f in range(3):
                app_counter += 1
                eq1, eq2 = constraints[i + j][f], sub_equation[j][f]
                if not check_nonzero_values(eq1, eq2):                         # :301-304
                    works = False
                    break
                for wire, coef in eq1.items():
                    if coef != 0:
                        appearance_map_cur.setdefault(wire, []).append((app_counter, coef))
            if not works:
                break
        if not works:
            continue
        l1 = sorted(appearance_map_cur.items(), key=by_sig)                    # :334-335
        if len(l1) != len(l2):
            continue
        if any(l1[k][1] != l2[k][1] for k in range(len(l1))):                  # :339-347
            continue
        matches.append((i, {l2[k][0]: l1[k][0] for k in range(len(l1))}))      # :351
    red_cons, special_cons = [], []
    cur_idx, i = 0, 0
    while i < len(constraints):                                                # :368-388
        if cur_idx >= len(matches) or i != matches[cur_idx][0]:
            red_cons.append(constraints[i])   # a match that starts inside a consumed window stalls cur_idx for good (:370)
            i += 1
        else:
            m = matches[cur_idx][1]
            try:
                special_cons.append((function_name, [m[x] for x in known_inputs if x != 1], [m[x] for x in known_outputs]))
            except KeyError as e:                                              # :381-382
                raise KeyErrorAt(*e.args)
            i += n
            cur_idx += 1
    return special_cons, red_cons
