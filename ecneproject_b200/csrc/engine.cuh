// engine.cuh — shared definitions of the sm_100a engine (data layout in HBM, flags, helpers).
//
// Layout.  Everything is SoA and sized once per problem:
//   rows   seg[3N+1] u32 | rflags[N] u32 | aux[N] 32 B | col[nnz] u32 | coef[nnz] 32 B (canonical)
//          only NON-ZERO terms; C terms of linear rows are stored sorted by |fold(coef)| (Case 5)
//   wires  F[2][V+4] u8  (U,K,ABZ bits; OR-monotone)    B[2][V+4] u8 (derived "bounds == [0,1]")
//          LBR/UBR[2][V+1] u32 ranks into the sorted table of every bound value that can occur
//          abz[V+1] i32 | valsrc[V+1] u32
//   double buffering: a Jacobi round reads buffer R and writes buffer W; the update records of a
//   round are replayed into the other buffer during the next round, so one grid barrier per round
//   suffices.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ecne_abi.h"
#include "fr_bn254.cuh"

namespace ecne {

// ---- static per-row classification (Appendix D of SURVEY.md: what depends on coefficients only) --
enum : uint32_t {
  RF_LINEAR = 1u << 0,     // nzk_a, nzk_b empty                         (:944-946)
  RF_CEMPTY = 1u << 1,     // nzk_c empty
  RF_2A = 1u << 2,         // C empty, one non-constant wire v* over A u B (:875-910)
  RF_2A_NOVAR = 1u << 3,   // C empty, no non-constant wire => variable_states[-1] (:916)
  RF_2A_DIVZ = 1u << 4,    // v* missing from A or B => divexact(_, 0)      (:919-920)
  RF_2A_BOOL = 1u << 5,    // roots are {0,1}                               (:923-927)
  RF_2B = 1u << 6,         // linear, one non-constant key in C             (:949-964)
  RF_C3 = 1u << 7,         // C multiset is the (possibly negated) bit-decomposition pattern
  RF_C3_FLIP = 1u << 8,    // ... negated: the reference flips the row in place (:1001-1011)
  RF_C3_L2 = 1u << 9,      // l == 2: {1,-1} matches both patterns, re-flipped on every pop
  RF_C3_TOPBIG = 1u << 10, // 2^(l-1)-1 >= p: the bound update (:1035) can never fire
  RF_4A = 1u << 11,        // stored C values == {1, -1}                    (:1079-1085)
  RF_4B = 1u << 12,        // stored C values == {1@wire1, -1, -1}          (:1150-1162)
  RF_P3 = 1u << 13,        // ABZ shape: C empty, |B| == 1, |A| <= 2        (:1427-1453)
  RF_P3_DIVZ = 1u << 14,   // ... with no non-constant key in A => divexact(_, 0) (:1467)
  RF_P4 = 1u << 15,        // rows (i, i+1) form the IsZero gadget          (:1493-1536)
  RF_LONG = 1u << 16,      // handled by a whole warp
};

struct __align__(16) RowAux {
  uint32_t w1;       // 2a: v*          2b: x
  uint32_t w2;       // C3: new_key (the coefficient-1 key after the flip)
  uint32_t w3;       // P3: b
  uint32_t w4;       // P3: slope_index P4: vk
  uint32_t val_idx;  // 2a: index into roots[]   2b: index into tvals[]
  uint32_t rank_a;   // 2b: rank of t
  uint32_t rank_b;   // C3: rank of 2^(l-1)-1
  uint32_t w5;       // C3 with l == 2: the other key
};

// ---- per-wire flag bits -------------------------------------------------------------------------
enum : uint32_t { WF_U = 1, WF_K = 2, WF_ABZ = 4 };

// update record: OR `bits` into F, max `lbr` into LBR, min `ubr` into UBR
struct __align__(16) Rec {
  uint32_t wire, bits, lbr, ubr;
};
#define ECNE_NO_LB 0u
#define ECNE_NO_UB 0xffffffffu

// valsrc codes: how .values of a wire is materialised at the end (never read by a rule)
#define VS_NONE 0u
#define VS_ONE 1u          // wire 1: [1]                         (:684-685)
#define VS_ONEZERO 2u      // [1, 0]                              (:1205, :1212)
#define VS_2A 0x40000000u  // | roots index: both roots           (:916-922)
#define VS_2B 0x80000000u  // | tvals index: [t]                  (:967)

struct Status {            // device-resident, read back once per outer round
  unsigned long long changed;      // state changes of the current outer round ("successful_steps")
  unsigned long long rounds;       // Jacobi rounds of the single-row sweep
  unsigned long long evals;        // rows visited
  unsigned int err;                // first error (as -status), 0 = none
  unsigned int p2_cand;            // candidates of the linear-system sweep
  unsigned int rec_overflow;
  unsigned int pad;
};

struct Dev {
  // sizes
  uint32_t N, V, nnz, n_long, n_specials, n_known, n_targets, n2a, n2b, table_n;
  uint32_t r0, r1, rpm1;  // ranks of 0, 1, p-1
  int secp_solve;
  // rows
  const uint32_t* seg;
  const uint32_t* col;
  const fr::u256* coef;
  const uint32_t* rflags;
  const RowAux* aux;
  const uint32_t* long_rows;
  uint8_t* solved;
  // static values
  const fr::u256* roots;  // [n2a][2]
  const fr::u256* tvals;  // [n2b]
  const fr::u256* table;  // [table_n] sorted distinct bound values
  // wires (double buffered)
  uint8_t* F[2];
  uint8_t* B[2];
  uint32_t* LBR[2];
  uint32_t* UBR[2];
  int32_t* abz;
  uint32_t* valsrc;
  const uint8_t* nontriv;
  // specials
  const int32_t* sp_kind;
  const uint32_t* sp_in_ptr;
  const uint32_t* sp_in;
  const uint32_t* sp_out_ptr;
  const uint32_t* sp_out;
  uint8_t* sp_solved;
  const uint32_t* known;
  const uint32_t* targets;
  // records: three rotating lists
  Rec* recs[3];
  unsigned int* rec_count;  // [3]
  uint32_t rec_cap;
  // scratch
  unsigned long long* abz_claim;  // [V+1]
  unsigned long long* p2_key;     // [N] candidate keys
  uint32_t* p2_row;               // [N]
  unsigned int* barrier;          // grid barrier counter
  Status* st;
  // sharding: this rank sweeps rows [row_lo, row_hi)
  uint32_t row_lo, row_hi;
};

// ---- state access (L2-coherent loads: the arrays are written by atomics from other SMs) ---------
__device__ __forceinline__ uint32_t ld_flag(const uint8_t* F, uint32_t w) {
  return (uint32_t)__ldcg(F + w);
}
__device__ __forceinline__ uint32_t ld_u32(const uint32_t* p, uint32_t i) { return __ldcg(p + i); }

__device__ __forceinline__ void raise(const Dev& d, int status) {
  atomicCAS(&d.st->err, 0u, (unsigned int)(-status));
}

// OR bits into a byte of F; returns the bits that were newly set
__device__ __forceinline__ uint32_t or_flag(uint8_t* F, uint32_t w, uint32_t bits) {
  unsigned int* word = (unsigned int*)(F + (w & ~3u));
  unsigned int sh = (w & 3u) * 8;
  unsigned int old = atomicOr(word, bits << sh);
  return bits & ~((old >> sh) & 0xffu);
}

// Recompute the derived "bounds == [0,1]" byte after a rank update.  lb only grows and ub only
// shrinks, so re-reading until stable makes the last writer publish the final value.
__device__ __forceinline__ void refresh_b01(const Dev& d, int buf, uint32_t w) {
  volatile uint32_t* lb = d.LBR[buf];
  volatile uint32_t* ub = d.UBR[buf];
  volatile uint8_t* b = d.B[buf];
  uint32_t l = lb[w], u = ub[w];
  while (true) {
    b[w] = (l == d.r0 && u == d.r1) ? 1 : 0;
    __threadfence();
    uint32_t l2 = lb[w], u2 = ub[w];
    if (l2 == l && u2 == u) break;
    l = l2;
    u = u2;
  }
}

// Apply one update to buffer `buf`; returns true when it changed anything there.
__device__ __forceinline__ bool apply_update(const Dev& d, int buf, uint32_t w, uint32_t bits,
                                             uint32_t lbr, uint32_t ubr) {
  bool ch = false;
  if (bits) ch |= or_flag(d.F[buf], w, bits) != 0;
  bool bch = false;
  if (lbr != ECNE_NO_LB) bch |= atomicMax(d.LBR[buf] + w, lbr) < lbr;
  if (ubr != ECNE_NO_UB) bch |= atomicMin(d.UBR[buf] + w, ubr) > ubr;
  if (bch) refresh_b01(d, buf, w);
  return ch | bch;
}

// Apply to the write buffer and, when it changed it, log the record for the other buffer.
__device__ __forceinline__ void emit(const Dev& d, int wbuf, int list, uint32_t w, uint32_t bits,
                                     uint32_t lbr = ECNE_NO_LB, uint32_t ubr = ECNE_NO_UB) {
  if (apply_update(d, wbuf, w, bits, lbr, ubr)) {
    unsigned int i = atomicAdd(d.rec_count + list, 1u);
    if (i < d.rec_cap) {
      Rec r;
      r.wire = w;
      r.bits = bits;
      r.lbr = lbr;
      r.ubr = ubr;
      d.recs[list][i] = r;
    } else {
      d.st->rec_overflow = 1;
    }
  }
}

// grid-wide barrier for a cooperatively launched (co-resident) grid
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int target = (epoch + 1) * gridDim.x;
    atomicAdd(counter, 1u);
    while (*((volatile unsigned int*)counter) < target) {
    }
    __threadfence();
  }
  epoch += 1;
  __syncthreads();
}

}  // namespace ecne
