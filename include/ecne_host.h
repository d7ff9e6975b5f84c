/*
 * ecne_host.h — C ABI of libecne_host.so: the host-side callers either side of the hot path
 * (SURVEY.md §8f "next" rows 1 and 2).  Pure C++ inside, no CUDA.
 *
 *   ecne_read_r1cs      replaces ParseR1CS.readR1CS          (/root/reference/src/ParseR1CS.jl:50-124)
 *   ecne_abstraction    replaces abstraction()               (src/R1CSConstraintSolver.jl:237-395)
 *
 * Both produce the flattened layout ecne_abi.h consumes, so that a host (Julia via ccall, Python
 * via ctypes) can go file -> ecne_problem_t without materialising per-row hash maps.
 */
#ifndef ECNE_HOST_H
#define ECNE_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef ECNE_E_KEYERROR
#define ECNE_E_KEYERROR (-10) /* KeyError at R1CSConstraintSolver.jl:381-382 */
#endif
#define ECNE_E_IO (-11)       /* SystemError opening the file (ParseR1CS.jl:52-53) */
#define ECNE_E_ASSERT (-12)   /* AssertionError at ParseR1CS.jl:58,62,69 */

/* A parsed constraint system; every array is owned by the library (ecne_r1cs_free). */
typedef struct ecne_r1cs {
  uint64_t n_rows;    /* nConstraints                         (ParseR1CS.jl:96)  */
  uint64_t n_vars;    /* num_wires + 1                        (ParseR1CS.jl:123) */
  uint64_t nnz;       /* stored terms, explicit zeros included                   */
  uint64_t* seg_ptr;  /* [3*n_rows+1]                                            */
  uint32_t* col;      /* [nnz] wire+1                         (ParseR1CS.jl:111) */
  uint64_t* coef;     /* [nnz*4] canonical limbs, reduced mod p (F(coeff))       */
  uint32_t* known;    /* [1; 2+nOut .. 1+nOut+nPubIn+nPrvIn]  (ParseR1CS.jl:123) */
  uint64_t n_known;
  uint32_t* targets;  /* [2 .. 1+nOut]                                           */
  uint64_t n_targets;
  uint32_t n_pub_out, n_pub_in, n_prv_in, field_size;
  uint64_t n_labels;
  /* The same rows in the compact form of include/ecne_abi.h (ecne_problem_t.coef_class ...), filled by the reader's
   * fast path while it copies the coefficients anyway; NULL when absent (outputs of ecne_abstraction, repaired files). */
  uint8_t* coef_class;       /* [nnz]                                                   */
  uint64_t* coef_other;      /* [n_coef_other*4]                                        */
  uint32_t* coef_other_term; /* [n_coef_other]                                          */
  uint64_t n_coef_other;
  uint32_t* seg_ptr32;       /* [3*n_rows+1]                                            */
} ecne_r1cs_t;

int ecne_read_r1cs(const char* path, ecne_r1cs_t** out);
/* ... with options.  ECNE_READ_COMPACT_ONLY: the full coefficient array is left out (coef == NULL; the compact arrays are
 * always there then, or the call fails) — for a caller that hands the rows to the device in the compact form. */
#define ECNE_READ_COMPACT_ONLY 1u
int ecne_read_r1cs_opts(const char* path, unsigned int flags, ecne_r1cs_t** out);
int ecne_read_r1cs_mem(const uint8_t* buf, uint64_t len, ecne_r1cs_t** out);
void ecne_r1cs_free(ecne_r1cs_t* r);
/* The large arrays of freed systems are kept for the next read (up to ECNE_HOST_CACHE_MB megabytes, default 1024; 0 =
 * keep nothing): a process that reads circuit after circuit writes into pages it already owns.  This returns them to
 * the system. */
void ecne_host_trim(void);

/* The list of (name, inputs, outputs) "special constraints" (:357-384), CSR over specials. */
typedef struct ecne_specials {
  uint64_t n;
  int32_t* kind;     /* ECNE_SPECIAL_* from the trusted function's name */
  uint64_t* in_ptr;  /* [n+1] */
  uint32_t* in;
  uint64_t* out_ptr; /* [n+1] */
  uint32_t* out;
  uint64_t cap_n, cap_in, cap_out; /* private */
} ecne_specials_t;

ecne_specials_t* ecne_specials_new(void);
void ecne_specials_free(ecne_specials_t* s);

/* One abstraction() call: find every window of `constraints` isomorphic to `sub`, replace it by a
 * special constraint of kind `kind` (appended to `specials`), return the reduced system in
 * *reduced (its known/targets/n_vars are copied from `constraints`).  `n_matches` receives the
 * number of specials added by this call. */
int ecne_abstraction(int32_t kind, const ecne_r1cs_t* constraints, const ecne_r1cs_t* sub,
                     ecne_r1cs_t** reduced, ecne_specials_t* specials, uint64_t* n_matches);

/* The compact form of a coefficient array (include/ecne_abi.h, ecne_problem_t.coef_class / coef_other /
 * coef_other_term): one class byte per stored term — 0: zero, 1: one, 2: p - 1, 3: another value — and the values and
 * term indices of the class-3 terms in term order.  Call with other == other_term == NULL to fill `cls` and learn
 * *n_other, then again with arrays of that size (multi-threaded; `cls` is rewritten identically). */
int ecne_compact_coef(const uint64_t* coef, uint64_t nnz, uint8_t* cls, uint64_t* other, uint32_t* other_term,
                      uint64_t* n_other);

const char* ecne_host_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* ECNE_HOST_H */
