/* solve_r1cs.c — the drop-in boundary from plain C: file -> reduced system (libecne_host.so) -> verdict
 * (libecne_b200.so), nothing but include/ecne_host.h and include/ecne_abi.h.
 *
 *   gcc -std=c99 -Iinclude examples/solve_r1cs.c -Lecneproject_b200 -lecne_host -lecne_b200 \
 *       -Wl,-rpath,$PWD/ecneproject_b200 -o solve_r1cs
 *   ./solve_r1cs [--gpus N] [--secp-solve] main.r1cs [trusted.r1cs TrustedName]...
 *
 * --gpus N: ONE process drives the GPUs 0 .. N-1 of the box (ecne_init_multi): the rows are sharded over them,
 * the result is identical to the one-GPU run.
 *
 * Mirrors solveWithTrustedFunctions (R1CSConstraintSolver.jl:502-581) for trusted circuits given longest first.
 * Exit code: 0 sound, 1 potentially unsound, 2 error (the message of the failing library is printed; without a
 * CUDA device that is ECNE_E_CUDA — there is no CPU fallback). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ecne_abi.h"
#include "ecne_host.h"

int main(int argc, char** argv) {
  int secp_solve = 0; /* the secp_solve keyword of solveWithTrustedFunctions (:511) */
  int gpus = 1;
  if (argc > 2 && strcmp(argv[1], "--gpus") == 0) {
    gpus = atoi(argv[2]);
    argc -= 2;
    argv += 2;
  }
  if (argc > 1 && strcmp(argv[1], "--secp-solve") == 0) {
    secp_solve = 1;
    --argc;
    ++argv;
  }
  if (argc < 2 || (argc % 2) != 0) {
    fprintf(stderr, "usage: %s [--gpus N] [--secp-solve] main.r1cs [trusted.r1cs TrustedName]...\n", argv[0]);
    return 2;
  }
  ecne_r1cs_t* cur = NULL;
  if (ecne_read_r1cs(argv[1], &cur) != 0) {
    fprintf(stderr, "read %s: %s\n", argv[1], ecne_host_last_error());
    return 2;
  }
  const uint64_t n_vars = cur->n_vars;
  ecne_specials_t* sp = ecne_specials_new();
  for (int i = 2; i + 1 < argc; i += 2) {
    ecne_r1cs_t *sub = NULL, *red = NULL;
    uint64_t n = 0;
    int32_t kind = strcmp(argv[i + 1], "BigMultModP") == 0   ? ECNE_SPECIAL_BIGMULTMODP
                   : strcmp(argv[i + 1], "BigLessThan") == 0 ? ECNE_SPECIAL_BIGLESSTHAN
                                                             : ECNE_SPECIAL_GENERIC;
    if (ecne_read_r1cs(argv[i], &sub) != 0 || ecne_abstraction(kind, cur, sub, &red, sp, &n) != 0) {
      fprintf(stderr, "abstraction of %s: %s\n", argv[i], ecne_host_last_error());
      return 2;
    }
    printf("%s: %llu window(s) abstracted, %llu rows left\n", argv[i + 1], (unsigned long long)n,
           (unsigned long long)red->n_rows);
    ecne_r1cs_free(sub);
    ecne_r1cs_free(cur);
    cur = red;
  }
  ecne_problem_t p;
  memset(&p, 0, sizeof p);
  p.n_rows = cur->n_rows;
  p.n_vars = n_vars;
  p.seg_ptr = cur->seg_ptr;
  p.col = cur->col;
  p.coef = cur->coef;
  p.known = cur->known;
  p.n_known = cur->n_known;
  p.targets = cur->targets;
  p.n_targets = cur->n_targets;
  p.n_specials = sp->n;
  p.sp_kind = sp->kind;
  p.sp_in_ptr = sp->in_ptr;
  p.sp_in = sp->in;
  p.sp_out_ptr = sp->out_ptr;
  p.sp_out = sp->out;
  p.secp_solve = secp_solve;
  ecne_result_t r;
  memset(&r, 0, sizeof r);
  const size_t words = (size_t)((n_vars + 63) / 64);
  r.unique_bits = (uint64_t*)calloc(words ? words : 1, 8);
  r.known_bits = (uint64_t*)calloc(words ? words : 1, 8);
  int st = gpus > 1 ? ecne_init_multi(gpus) : 0; /* one GPU: ecne_solve binds device 0 by itself */
  if (st == 0) st = ecne_solve(&p, &r);
  if (st != 0) {
    fprintf(stderr, "ecne_solve: status %d: %s\n", st, ecne_last_error());
    return 2;
  }
  printf("Solved for %llu variables out of %llu total variables\n", (unsigned long long)r.n_unique_nontrivial,
         (unsigned long long)r.n_nontrivial);
  printf("Solved for %llu target variables out of %llu total target variables\n",
         (unsigned long long)r.n_targets_unique, (unsigned long long)cur->n_targets);
  printf("%s (%.3f ms on the device, %llu constraint evaluations, %llu GPU(s))\n",
         r.verdict ? "sound constraints" : "potentially unsound constraints", r.ms_device,
         (unsigned long long)r.constraint_evals, (unsigned long long)r.gpus_used);
  ecne_shutdown();
  return r.verdict ? 0 : 1;
}
