/*
 * ecne_abi.h — C ABI of libecne_b200.so, the B200-native R1CS soundness-propagation engine.
 *
 * Drop-in boundary: the ONE call
 *     result = SolveConstraintsSymbolic(reduced, specials, knowns_main, debug, outs_main,
 *                                       num_variables, input_sym, secp_solve)
 * at /root/reference/src/R1CSConstraintSolver.jl:552 (signature :583-592, returns Bool :1645).
 * The reference has no FFI of its own; a Julia maintainer keeps readR1CS (ParseR1CS.jl:50-124),
 * abstraction (:237-395) and solveWithTrustedFunctions (:502-581) untouched and replaces the body
 * of that one call with `ccall((:ecne_solve, "libecne_b200"), ...)` — see INTEGRATION.md.
 *
 * Conventions
 *   - plain C, caller owns every buffer, the library never keeps a host pointer after return;
 *   - wire ids are the reference's 1-based keys (ParseR1CS.jl:111 stores `wire+1`), wire 1 is the
 *     constant one;  n_vars == num_variables == nWires+1 (ParseR1CS.jl:123);
 *   - field elements are 4 little-endian uint64 limbs of the canonical integer in [0, p), p = the
 *     BN254 scalar prime (R1CSConstraintSolver.jl:21-24) — exactly GFElem.d;
 *   - a constraint row is three sparse linear forms A, B, C (ParseR1CS.jl:13-25).  They are passed
 *     as ONE term array in the .r1cs on-disk order (ParseR1CS.jl:100-117): segment 3*i+0/1/2 is
 *     A/B/C of row i and covers terms seg_ptr[3*i+s] .. seg_ptr[3*i+s+1]-1.  EVERY stored key of
 *     the Julia DefaultDict is passed, explicit zeros included (the parser stores `[1]=F(0)` for an
 *     empty form, ParseR1CS.jl:113-115; :641, :1001, :1083, :1512 look at stored zeros);
 *   - all entry points return 0 on success or a negative ecne_status; ecne_last_error() gives text.
 *     There is no CPU fallback: without a usable CUDA device every compute entry fails with
 *     ECNE_E_CUDA.
 */
#ifndef ECNE_ABI_H
#define ECNE_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECNE_ABI_VERSION 3

/* Status codes.  The negative ones mirror the exception classes the reference can raise on this
 * path (SURVEY.md §5 / §8b "Errors"); the Julia shim rethrows them. */
typedef enum ecne_status {
  ECNE_OK = 0,
  ECNE_E_BADARG = -1,      /* AssertionError-class: malformed problem                       */
  ECNE_E_DIVZERO = -2,     /* DivideError: divexact(_, 0) at :919-920 or :1467                */
  ECNE_E_BOUNDS = -3,      /* BoundsError: variable_states[-1] at :916, short lists at :762   */
  ECNE_E_NODSU = -4,       /* UndefVarError(:dsu) at :762 when secp_solve == false            */
  ECNE_E_CUDA = -5,        /* CUDA runtime failure / no device / extension not built          */
  ECNE_E_NCCL = -6,        /* multi-GPU exchange failure                                      */
  ECNE_E_UNSUPPORTED = -7, /* a linear-system group larger than ECNE_P2_KBIG triggered        */
  ECNE_E_NOCONVERGE = -8,  /* round guard hit (the reference would loop forever)              */
  ECNE_E_INTERNAL = -9
} ecne_status;

/* Special-constraint kinds: only these two names are looked at by the solver (:751, :755). */
#define ECNE_SPECIAL_GENERIC 0
#define ECNE_SPECIAL_BIGMULTMODP 1
#define ECNE_SPECIAL_BIGLESSTHAN 2

/* Largest k for which the k x k "slow_det" of the linear-system sweep (:1389-1400) is evaluated. */
#define ECNE_P2_KMAX 8   /* up to here: the reference's own enumeration of the k! permutations (:1389-1400)          */
#define ECNE_P2_KBIG 16  /* up to here: the same odd-permutation sum as (permanent - determinant) / 2, a block per group */

typedef struct ecne_problem {
  uint64_t n_rows;         /* length(constraints)                                   (:584) */
  uint64_t n_vars;         /* num_variables                                         (:589) */
  const uint64_t* seg_ptr; /* [3*n_rows + 1] offsets into col/coef                         */
  const uint32_t* col;     /* [nnz] 1-based wire id of each stored term                    */
  const uint64_t* coef;    /* [nnz*4] canonical limbs, zeros allowed                       */
  const uint32_t* known;   /* known_variables (:586), contains wire 1   (ParseR1CS.jl:123) */
  uint64_t n_known;
  const uint32_t* targets; /* target_variables (:588)                                      */
  uint64_t n_targets;
  uint64_t n_specials;        /* special_constraints (:585): (name, inputs, outputs)       */
  const int32_t* sp_kind;     /* [n_specials] ECNE_SPECIAL_*                               */
  const uint64_t* sp_in_ptr;  /* [n_specials + 1]                                          */
  const uint32_t* sp_in;      /* inputs, 1-based wires                                     */
  const uint64_t* sp_out_ptr; /* [n_specials + 1]                                          */
  const uint32_t* sp_out;     /* outputs                                                   */
  int32_t secp_solve;         /* :591                                                      */
  int32_t debug;              /* :587 — accepted, ignored (printing stays in the host)     */
  /* Optional COMPACT form of the coefficients (ABI version 3; `coef == NULL` selects it).  The coefficients of a
   * circom circuit are almost all 0, 1 or p-1 (ecdsa.r1cs after abstraction: 2.62 M of 2.83 M stored terms), so the
   * host that flattens `Dict{Int, GFElem}` (`R1CSEquation`, ParseR1CS.jl:9-13) can hand over one class byte per term
   * and 32-byte limbs only for the others: 38 MB instead of 118 MB cross PCIe for the headline workload, and the
   * library expands them into the same device arrays the full form is copied into (csrc/setup.cu upload_rows).
   * The host library (include/ecne_host.h) has a helper that makes this form from a full `coef` array. */
  const uint8_t* coef_class;       /* [nnz] 0: stored zero, 1: one, 2: p - 1, 3: the value is in coef_other    */
  const uint64_t* coef_other;      /* [n_coef_other*4] canonical limbs of the class-3 terms, in term order   */
  const uint32_t* coef_other_term; /* [n_coef_other] index (into col / coef_class) of each of them, ascending */
  uint64_t n_coef_other;
  const uint32_t* seg_ptr32;       /* optional 32-bit copy of seg_ptr (`seg_ptr == NULL` selects it)           */
} ecne_problem_t;

typedef struct ecne_result {
  int32_t verdict;       /* function_good (:1594-1597, :1645): 1 sound, 0 potentially unsound */
  int32_t status;        /* copy of the return code                                         */
  /* Per-wire final VariableState (:135-160).  unique_bits/known_bits are required; the rest may
   * be NULL.  Bit (w-1)&63 of word (w-1)>>6 belongs to wire w. */
  uint64_t* unique_bits; /* [(n_vars+63)/64]  .unique                                       */
  uint64_t* known_bits;  /* [(n_vars+63)/64]  .is_known                                     */
  uint64_t* lb;          /* [n_vars*4]        .lb.d                                          */
  uint64_t* ub;          /* [n_vars*4]        .ub.d                                          */
  uint8_t* nvalues;      /* [n_vars]          length(.values)                                */
  uint64_t* values;      /* [n_vars*8]        .values[1..2].d                                */
  int32_t* abz;          /* [n_vars]          .abz (1-based wire or -1)                      */
  /* counters (:1558-1597 prints the first four) */
  uint64_t n_unique_nontrivial; /* "Solved for X variables ..."                              */
  uint64_t n_nontrivial;        /* "... out of Y total variables"                            */
  uint64_t n_targets_unique;
  uint64_t n_unique;            /* popcount(unique_bits)                                     */
  uint64_t outer_rounds;        /* iterations of the `while true` at :706                    */
  uint64_t inner_rounds;        /* Jacobi rounds of the single-row rule sweep                */
  uint64_t constraint_evals;    /* rows visited by sweep/phase kernels        */
  uint64_t sweep_launches;      /* kernels launched by this call (reset, ONE solve, verdict) */
  double ms_h2d, ms_classify, ms_solve, ms_d2h, ms_exchange, ms_total;
  double ms_sweep;              /* device time inside the single-row sweep kernel only       */
  uint64_t rule_evals;          /* of constraint_evals: rows whose rule set was run in full       */
  /* the sweep kernel's two kinds of Jacobi round: dense rounds sweep every row that can still fire,
   * the others are driven by the previous round's update records */
  uint64_t dense_rounds;        /* number of dense rounds                                          */
  uint64_t dense_evals;         /* rows visited by them                                            */
  uint64_t dense_cycles;        /* SM cycles (clock64, block 0) spent in them, barrier included    */
  double ms_device;             /* CUDA-event time of the whole call on the engine's stream: reset,
                                   all rounds, verdict kernels and the D2H of the bitmaps           */
  uint64_t gpus_used;           /* GPUs the solve kernel ran on (1, or the world of a multi-GPU run) */
  uint64_t sharded;             /* 1: the dense sweeps were split over those GPUs; 0: one GPU, or every GPU
                                   solved the whole problem (fewer rows than "shard_min_rows")         */
} ecne_result_t;

/* ---- life cycle ------------------------------------------------------------------------- */
int ecne_version(void);
/* Layout self-check for foreign-language mirrors of the structs above (Python ctypes, a Julia `struct`): for each
 * of ecne_problem_t, ecne_result_t, ecne_report_t, in this order, the words {sizeof, number of fields,
 * offsetof(field) ... in declaration order}.  Returns the length of the table and writes at most `cap` words.
 * A binding compares it with its own offsets once at load time (ecneproject_b200/_abi.py). */
int ecne_abi_layout(uint32_t* out, uint32_t cap);
/* Bind this process to CUDA device `device` (one GPU, or one process per GPU of a sharded run: see
 * ecne_dist_init below).  Idempotent. */
int ecne_init(int device);
/* SURVEY.md §8b "Threading": ONE process, one host thread, the GPUs 0 .. n_gpus-1 of one box — what the
 * single-threaded Julia caller of :552 needs to reach every GPU.  The library opens the devices, enables peer
 * access between them (no IPC handles, no NCCL, no second process) and from then on ecne_upload / ecne_solve /
 * ecne_solve_resident shard the rows over all of them exactly as a one-process-per-GPU run does (same kernels,
 * same exchange protocol, results bit-identical to one GPU).  n_gpus == 1 is ecne_init(0).  Idempotent. */
int ecne_init_multi(int n_gpus);
void ecne_shutdown(void);
const char* ecne_last_error(void);

/* ---- the drop-in call: host buffers in, host buffers out (H2D + all rounds + D2H) -------- */
int ecne_solve(const ecne_problem_t* problem, ecne_result_t* result);

/* ---- resident variant: upload + classify once, solve many times (bench `value` leg) ------ */
typedef struct ecne_resident ecne_resident_t;
int ecne_upload(const ecne_problem_t* problem, ecne_resident_t** out);
int ecne_solve_resident(ecne_resident_t* r, ecne_result_t* result);
void ecne_free_resident(ecne_resident_t* r);

/* ---- abstraction() on the device (R1CSConstraintSolver.jl:237-395, helpers :205-235) ------
 * The caller of :552, solveWithTrustedFunctions (:502-581), first replaces every window of the main circuit that is
 * isomorphic to a trusted circuit by a special constraint (:538-544), trusted circuits longest first.  With these
 * entry points that pass runs on the GPU: the UNREDUCED system is uploaded once (ecne_abstract_begin), every trusted
 * circuit is matched by kernels and the kept rows are compacted on the device (ecne_abstract_apply, one call per
 * trusted circuit in the order of :527-537; `sub` = its rows, known = its inputs, targets = its outputs; n_matches =
 * windows replaced by this call), and ecne_abstract_upload classifies the reduced system where it lies — it never
 * crosses PCIe and no host core touches a row — giving the same resident handle as ecne_upload.  KeyError of
 * :381-382 (an input / output of the trusted circuit that never appears in it) is ECNE_E_KEYERROR.
 * ecne_abstract_sizes: {rows, stored terms, special constraints, their input wires, their output wires} so far;
 * ecne_abstract_export: D2H of the reduced system and the specials into caller arrays of those sizes (any pointer may
 * be NULL) — parity tests and hosts that want the reduced system back. */
#ifndef ECNE_E_KEYERROR
#define ECNE_E_KEYERROR (-10)
#endif
typedef struct ecne_abstracted ecne_abstracted_t;
int ecne_abstract_begin(const ecne_problem_t* main_circuit, ecne_abstracted_t** out);
int ecne_abstract_apply(ecne_abstracted_t* a, int32_t kind, const ecne_problem_t* sub, uint64_t* n_matches);
/* The part of an apply that depends on the trusted circuit alone (coefficient multisets, wire signatures and their
 * classes: milliseconds of host work) can be done ahead — it touches neither the GPU nor any state of the library, so a
 * host prepares the trusted circuits on another thread while it still reads the main circuit — and handed to the
 * apply together with the same `sub`. */
typedef struct ecne_prepared ecne_prepared_t;
int ecne_abstract_prepare(const ecne_problem_t* sub, ecne_prepared_t** out);
int ecne_abstract_apply_prepared(ecne_abstracted_t* a, int32_t kind, const ecne_problem_t* sub, const ecne_prepared_t* prepared,
                                 uint64_t* n_matches);
void ecne_abstract_prepared_free(ecne_prepared_t* p);
int ecne_abstract_sizes(const ecne_abstracted_t* a, uint64_t sizes[5]);
int ecne_abstract_export(ecne_abstracted_t* a, uint64_t* seg_ptr, uint32_t* col, uint64_t* coef, int32_t* sp_kind,
                         uint64_t* sp_in_ptr, uint32_t* sp_in, uint64_t* sp_out_ptr, uint32_t* sp_out);
int ecne_abstract_upload(ecne_abstracted_t* a, int32_t secp_solve, ecne_resident_t** out);
void ecne_abstract_free(ecne_abstracted_t* a);

/* ---- report path: the "Bad Constraints" listing (R1CSConstraintSolver.jl:1599-1635) ----------
 * The reference walks every constraint, keeps those that mention (getVariables, :36-56: a stored
 * non-zero coefficient) a wire that is not unique (:1612-1620), and prints the state of each of their
 * wires but wire 1 (:1627-1633).  The engine finds those rows and compacts the state of exactly those
 * wires on the device, so the D2H is the bitmap of rows plus one entry per listed wire instead of the
 * whole per-wire state (ecdsa: 140 MB).  Refers to the LAST solve of the handle; all buffers are the
 * caller's.  If cap_wires is too small the call returns ECNE_E_BADARG with bad_row_bits, n_bad_rows and
 * n_wires filled in, so the caller can size the arrays and call again. */
typedef struct ecne_report {
  uint64_t* bad_row_bits; /* [(n_rows+63)/64] required; bit i&63 of word i>>6 <=> 0-based row i is listed */
  uint64_t cap_wires;     /* entries each of the arrays below can hold                                  */
  uint32_t* wire;         /* [cap_wires]   1-based wire ids, ascending; may be NULL (then so are the rest) */
  uint8_t* flags;         /* [cap_wires]   bit 0 .unique, bit 1 .is_known                                */
  uint64_t* lb;           /* [cap_wires*4] .lb.d                                                        */
  uint64_t* ub;           /* [cap_wires*4] .ub.d                                                        */
  uint8_t* nvalues;       /* [cap_wires]   length(.values)                                              */
  uint64_t* values;       /* [cap_wires*8] .values[1..2].d                                              */
  uint64_t n_bad_rows;    /* out: popcount(bad_row_bits)                                                */
  uint64_t n_wires;       /* out: wires (other than wire 1) with a non-zero coefficient in a listed row */
} ecne_report_t;
int ecne_report_resident(ecne_resident_t* r, ecne_report_t* report);

/* ---- row-range sharding across the GPUs of one box ------------------------
 * Either ONE process for all GPUs (ecne_init_multi above) or one process per GPU (ecne_init(local_rank) +
 * ecne_dist_init).  Every rank is given the WHOLE problem by its host and classifies it; wire state and the phases
 * P0 / P2 / P3 / P4 are replicated.  The dense sweeps of the single-row rules — the only rounds that stream every
 * row — are split by row ranges [lo, hi) of equal stored-term weight: their update records are exchanged over
 * NVLink peer mappings inside the solve kernel (one exchange per sharded round: post counts into the peers'
 * mailboxes, pull the peers' record lists, acknowledge).  Frontier-driven rounds cost less than an exchange and
 * are run by every rank on all rows, without any communication.
 * One process per GPU: NCCL only bootstraps — unique_id is the 128-byte ncclUniqueId made by rank 0
 * (ecne_dist_unique_id) and broadcast by the host (torch.distributed / MPI / Julia Distributed); the engine
 * all-gathers its CUDA IPC handles through it.  ecne_upload() and ecne_solve*() are collective calls. */
int ecne_dist_unique_id(uint8_t out[128]);
int ecne_dist_init(int rank, int world, const uint8_t unique_id[128]);
int ecne_dist_rank(void);
int ecne_dist_world(void);
/* Pure host helper (no GPU needed): the contiguous row range [lo, hi) that rank `rank` of `world`
 * sweeps, chosen so that every range holds about the same number of stored terms. */
int ecne_shard_rows(const ecne_problem_t* problem, int rank, int world, uint64_t* lo, uint64_t* hi);

/* ---- engine knobs (testing / benchmarking) ----------------------------------------------
 * "max_rounds" / "max_outer": round guards (ECNE_E_NOCONVERGE when hit); "sparse_max": a Jacobi round
 * whose frontier has at most this many changed wires is frontier-driven instead of a dense sweep (-1: rows/32,
 * 0: always dense); "grid_blocks": launch the solve kernel with fewer blocks than SMs (0: one per SM);
 * "shard_min_rows": on several GPUs, problems with at least this many rows have their dense sweeps split over the
 * GPUs, smaller ones are solved by every GPU in full without any exchange (2 000 000; 0: always shard — must be the
 * same on every rank); "shard_min_rows_per_gpu": ... and only when every GPU gets at least this many rows (4 000 000:
 * every rank applies every record of a sharded round, only the sweep shrinks with the number of GPUs);
 * "chain_open_max": once the linear-system sweep has at most this many rows left to look at (and the frontier is small),
 * ONE block runs whole outer rounds — special constraints, Jacobi rounds, linear systems, IsZero — without any grid
 * barrier (4096; 0: never; results do not depend on it);
 * "shard_upload": one process per GPU (ecne_dist_init): every rank copies 1/world of the rows over its own PCIe link and
 * the ranks gather the slices over NVLink (1; 0: every rank uploads the whole problem; set before ecne_dist_init, same
 * on every rank);
 * "solve_variant": which build of the solve kernel runs (0: by size — a GPU that sweeps at least "wide_min_rows" = 1 500 000 rows
 * takes the 1024-thread x 64-register build, smaller problems the 512 x 128 one; 1 / 2 force them);
 * "p2_hash_bits": bits of the unknown-set hash the linear-system sweep groups by (56; fewer force collisions,
 * which the engine resolves by exact comparison — results do not depend on it). */
int ecne_set_option(const char* key, int64_t value);

/* ---- field-arithmetic known-answer hooks: run the device Montgomery code on n elements ----
 * op: 0 add, 1 sub, 2 mul, 3 inv (b ignored), 4 neg (b ignored), 5 divexact(-a, b); the plain 256-bit
 * integer helpers of Case 5 (:1266-1274): 6 -> 1 if b divides a else 0 (b != 0), 7 -> sign(a*b - p) + 1.
 * a, b, out: [n*4] canonical. */
int ecne_fr_batch(int op, uint64_t n, const uint64_t* a, const uint64_t* b, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* ECNE_ABI_H */
