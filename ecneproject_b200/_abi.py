"""ctypes mirror of include/ecne_abi.h and include/ecne_host.h (field order must match exactly)."""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
u64p = C.POINTER(C.c_uint64)

ECNE_OK = 0
ECNE_E_BADARG = -1
ECNE_E_DIVZERO = -2
ECNE_E_BOUNDS = -3
ECNE_E_NODSU = -4
ECNE_E_CUDA = -5
ECNE_E_NCCL = -6
ECNE_E_UNSUPPORTED = -7
ECNE_E_NOCONVERGE = -8
ECNE_E_INTERNAL = -9
ECNE_E_KEYERROR = -10
ECNE_E_IO = -11
ECNE_E_ASSERT = -12

SPECIAL_GENERIC, SPECIAL_BIGMULTMODP, SPECIAL_BIGLESSTHAN = 0, 1, 2


class Problem(C.Structure):
    _fields_ = [
        ("n_rows", C.c_uint64), ("n_vars", C.c_uint64),
        ("seg_ptr", u64p), ("col", u32p), ("coef", u64p),
        ("known", u32p), ("n_known", C.c_uint64),
        ("targets", u32p), ("n_targets", C.c_uint64),
        ("n_specials", C.c_uint64), ("sp_kind", i32p),
        ("sp_in_ptr", u64p), ("sp_in", u32p), ("sp_out_ptr", u64p), ("sp_out", u32p),
        ("secp_solve", C.c_int32), ("debug", C.c_int32),
        # compact form of the coefficients (coef == NULL selects it) and 32-bit offsets (seg_ptr == NULL)
        ("coef_class", u8p), ("coef_other", u64p), ("coef_other_term", u32p), ("n_coef_other", C.c_uint64),
        ("seg_ptr32", u32p),
    ]


class Result(C.Structure):
    _fields_ = [
        ("verdict", C.c_int32), ("status", C.c_int32),
        ("unique_bits", u64p), ("known_bits", u64p),
        ("lb", u64p), ("ub", u64p), ("nvalues", u8p), ("values", u64p), ("abz", i32p),
        ("n_unique_nontrivial", C.c_uint64), ("n_nontrivial", C.c_uint64),
        ("n_targets_unique", C.c_uint64), ("n_unique", C.c_uint64),
        ("outer_rounds", C.c_uint64), ("inner_rounds", C.c_uint64),
        ("constraint_evals", C.c_uint64), ("sweep_launches", C.c_uint64),
        ("ms_h2d", C.c_double), ("ms_classify", C.c_double), ("ms_solve", C.c_double),
        ("ms_d2h", C.c_double), ("ms_exchange", C.c_double), ("ms_total", C.c_double),
        ("ms_sweep", C.c_double), ("rule_evals", C.c_uint64),
        ("dense_rounds", C.c_uint64), ("dense_evals", C.c_uint64), ("dense_cycles", C.c_uint64),
        ("ms_device", C.c_double), ("gpus_used", C.c_uint64), ("sharded", C.c_uint64),
    ]


class R1CSStruct(C.Structure):
    _fields_ = [
        ("n_rows", C.c_uint64), ("n_vars", C.c_uint64), ("nnz", C.c_uint64),
        ("seg_ptr", u64p), ("col", u32p), ("coef", u64p),
        ("known", u32p), ("n_known", C.c_uint64),
        ("targets", u32p), ("n_targets", C.c_uint64),
        ("n_pub_out", C.c_uint32), ("n_pub_in", C.c_uint32), ("n_prv_in", C.c_uint32),
        ("field_size", C.c_uint32), ("n_labels", C.c_uint64),
        ("coef_class", u8p), ("coef_other", u64p), ("coef_other_term", u32p), ("n_coef_other", C.c_uint64),
        ("seg_ptr32", u32p),
    ]


class Report(C.Structure):
    """ecne_report_t (include/ecne_abi.h): the rows and wires of the "Bad Constraints" listing."""
    _fields_ = [
        ("bad_row_bits", u64p), ("cap_wires", C.c_uint64), ("wire", u32p), ("flags", u8p),
        ("lb", u64p), ("ub", u64p), ("nvalues", u8p), ("values", u64p),
        ("n_bad_rows", C.c_uint64), ("n_wires", C.c_uint64),
    ]


class SpecialsStruct(C.Structure):
    _fields_ = [
        ("n", C.c_uint64), ("kind", i32p), ("in_ptr", u64p), ("in_", u32p),
        ("out_ptr", u64p), ("out", u32p),
        ("cap_n", C.c_uint64), ("cap_in", C.c_uint64), ("cap_out", C.c_uint64),
    ]


# every symbol include/ecne_abi.h declares (tests check the built library exports all of them)
ENGINE_SYMBOLS = [
    "ecne_version", "ecne_abi_layout", "ecne_init", "ecne_init_multi", "ecne_shutdown", "ecne_last_error", "ecne_solve",
    "ecne_upload", "ecne_solve_resident", "ecne_free_resident", "ecne_report_resident",
    "ecne_dist_unique_id", "ecne_dist_init", "ecne_dist_rank", "ecne_dist_world", "ecne_shard_rows",
    "ecne_set_option", "ecne_fr_batch",
    "ecne_abstract_begin", "ecne_abstract_apply", "ecne_abstract_sizes", "ecne_abstract_export",
    "ecne_abstract_upload", "ecne_abstract_free",
    "ecne_abstract_prepare", "ecne_abstract_apply_prepared", "ecne_abstract_prepared_free",
]
HOST_SYMBOLS = [
    "ecne_read_r1cs", "ecne_read_r1cs_opts", "ecne_read_r1cs_mem", "ecne_r1cs_free", "ecne_specials_new",
    "ecne_specials_free", "ecne_abstraction", "ecne_host_last_error", "ecne_compact_coef", "ecne_host_trim",
]

_host = None
_engine = None


def host_lib():
    """libecne_host.so (pure C++)."""
    global _host
    if _host is None:
        path = os.path.join(PKG, "libecne_host.so")
        if not os.path.exists(path):
            raise RuntimeError(
                "libecne_host.so is not built; run `python -m ecneproject_b200.build` (or "
                "__graft_entry__.build())")
        lib = C.CDLL(path)
        R1p = C.POINTER(R1CSStruct)
        Sp = C.POINTER(SpecialsStruct)
        lib.ecne_read_r1cs.argtypes = [C.c_char_p, C.POINTER(R1p)]
        lib.ecne_read_r1cs.restype = C.c_int
        lib.ecne_read_r1cs_opts.argtypes = [C.c_char_p, C.c_uint, C.POINTER(R1p)]
        lib.ecne_read_r1cs_opts.restype = C.c_int
        lib.ecne_read_r1cs_mem.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(R1p)]
        lib.ecne_read_r1cs_mem.restype = C.c_int
        lib.ecne_r1cs_free.argtypes = [R1p]
        lib.ecne_r1cs_free.restype = None
        lib.ecne_specials_new.argtypes = []
        lib.ecne_specials_new.restype = Sp
        lib.ecne_specials_free.argtypes = [Sp]
        lib.ecne_specials_free.restype = None
        lib.ecne_abstraction.argtypes = [C.c_int32, R1p, R1p, C.POINTER(R1p), Sp, u64p]
        lib.ecne_abstraction.restype = C.c_int
        lib.ecne_compact_coef.argtypes = [u64p, C.c_uint64, u8p, u64p, u32p, u64p]
        lib.ecne_compact_coef.restype = C.c_int
        lib.ecne_host_last_error.argtypes = []
        lib.ecne_host_last_error.restype = C.c_char_p
        _host = lib
    return _host


ABI_VERSION = 3


def layout_table():
    """{sizeof, n fields, offsets...} of the three ctypes mirrors, in the order ecne_abi_layout() reports them."""
    t = []
    for st in (Problem, Result, Report):
        t += [C.sizeof(st), len(st._fields_)] + [getattr(st, f[0]).offset for f in st._fields_]
    return t


def check_layout(lib):
    """One field-order slip in a mirror is a silent ABI break: compare with what the library was compiled with."""
    if lib.ecne_version() != ABI_VERSION:
        raise RuntimeError(f"libecne_b200.so has ABI version {lib.ecne_version()}, this mirror is for {ABI_VERSION}")
    n = lib.ecne_abi_layout(None, 0)
    buf = (C.c_uint32 * n)()
    lib.ecne_abi_layout(buf, n)
    if list(buf) != layout_table():
        raise RuntimeError("ctypes mirror of include/ecne_abi.h does not match the library's struct layout: "
                           f"library {list(buf)} mirror {layout_table()}")


def engine_lib():
    """libecne_b200.so — the CUDA engine.  There is no fallback: missing library is an error."""
    global _engine
    if _engine is None:
        path = os.environ.get("ECNE_ENGINE_SO", os.path.join(PKG, "libecne_b200.so"))
        if not os.path.exists(path):
            raise RuntimeError(
                "libecne_b200.so (the sm_100a CUDA engine) is not built and there is no CPU "
                "fallback; run `python -m ecneproject_b200.build`")
        lib = C.CDLL(path)
        Pp = C.POINTER(Problem)
        Rp = C.POINTER(Result)
        lib.ecne_version.restype = C.c_int
        lib.ecne_init.argtypes = [C.c_int]
        lib.ecne_init.restype = C.c_int
        lib.ecne_init_multi.argtypes = [C.c_int]
        lib.ecne_init_multi.restype = C.c_int
        lib.ecne_shutdown.restype = None
        lib.ecne_last_error.restype = C.c_char_p
        lib.ecne_solve.argtypes = [Pp, Rp]
        lib.ecne_solve.restype = C.c_int
        lib.ecne_upload.argtypes = [Pp, C.POINTER(C.c_void_p)]
        lib.ecne_upload.restype = C.c_int
        lib.ecne_solve_resident.argtypes = [C.c_void_p, Rp]
        lib.ecne_solve_resident.restype = C.c_int
        lib.ecne_free_resident.argtypes = [C.c_void_p]
        lib.ecne_free_resident.restype = None
        lib.ecne_report_resident.argtypes = [C.c_void_p, C.POINTER(Report)]
        lib.ecne_report_resident.restype = C.c_int
        lib.ecne_dist_unique_id.argtypes = [u8p]
        lib.ecne_dist_unique_id.restype = C.c_int
        lib.ecne_dist_init.argtypes = [C.c_int, C.c_int, u8p]
        lib.ecne_dist_init.restype = C.c_int
        lib.ecne_dist_rank.restype = C.c_int
        lib.ecne_dist_world.restype = C.c_int
        lib.ecne_shard_rows.argtypes = [Pp, C.c_int, C.c_int, u64p, u64p]
        lib.ecne_shard_rows.restype = C.c_int
        lib.ecne_set_option.argtypes = [C.c_char_p, C.c_int64]
        lib.ecne_set_option.restype = C.c_int
        lib.ecne_fr_batch.argtypes = [C.c_int, C.c_uint64, u64p, u64p, u64p]
        lib.ecne_fr_batch.restype = C.c_int
        lib.ecne_abstract_begin.argtypes = [Pp, C.POINTER(C.c_void_p)]
        lib.ecne_abstract_begin.restype = C.c_int
        lib.ecne_abstract_apply.argtypes = [C.c_void_p, C.c_int32, Pp, u64p]
        lib.ecne_abstract_apply.restype = C.c_int
        lib.ecne_abstract_prepare.argtypes = [Pp, C.POINTER(C.c_void_p)]
        lib.ecne_abstract_prepare.restype = C.c_int
        lib.ecne_abstract_apply_prepared.argtypes = [C.c_void_p, C.c_int32, Pp, C.c_void_p, u64p]
        lib.ecne_abstract_apply_prepared.restype = C.c_int
        lib.ecne_abstract_prepared_free.argtypes = [C.c_void_p]
        lib.ecne_abstract_prepared_free.restype = None
        lib.ecne_abstract_sizes.argtypes = [C.c_void_p, u64p]
        lib.ecne_abstract_sizes.restype = C.c_int
        lib.ecne_abstract_export.argtypes = [C.c_void_p, u64p, u32p, u64p, i32p, u64p, u32p, u64p, u32p]
        lib.ecne_abstract_export.restype = C.c_int
        lib.ecne_abstract_upload.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
        lib.ecne_abstract_upload.restype = C.c_int
        lib.ecne_abstract_free.argtypes = [C.c_void_p]
        lib.ecne_abstract_free.restype = None
        lib.ecne_abi_layout.argtypes = [u32p, C.c_uint32]
        lib.ecne_abi_layout.restype = C.c_int
        check_layout(lib)
        _engine = lib
    return _engine
