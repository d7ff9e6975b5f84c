// engine.cuh — shared definitions of the sm_100a engine (data layout in HBM, flags, helpers).
//
// Layout.  Everything is SoA and sized once per problem:
//   rows   seg[3N+1] u32 | rflags[N] u32 | aux[N] 32 B | col[nnz] u32 | coef[nnz] 32 B (canonical)
//          only NON-ZERO terms; C terms of linear rows are stored sorted by |fold(coef)| (Case 5)
//          rec[N] 32 B row records (flags + <= 6 inline wires): what the solve kernel streams
//          inv_head[V+2] 16 B | inv_ptr | inv_row   wire -> rows index (the frontier of a sparse round)
//   wires  F[2][V+4] u8  (U, K, ABZ, BND, UB01/NOT01, HEAVY bits; OR-monotone)
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
  RF_FAST = 1u << 17,      // <= 6 terms, no bound pattern: the inline fast path of the sweep
};

// 32-byte row record streamed by the sweep: flags + the row's wires inline (A u B terms first, then
// C).  p99 row length is 3, so almost every row is evaluated from this one coalesced 32-byte load
// plus its state-byte gathers; longer / pattern rows fall back to the CSR arrays.
#define ROWREC_INLINE 6
struct __align__(32) RowRec {
  uint32_t rf;
  uint8_t nAB, nC, inl, pad;
  uint32_t c[ROWREC_INLINE];
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
enum : uint32_t { WF_U = 1, WF_K = 2, WF_ABZ = 4, WF_BND = 8,  // BND: bounds were ever tightened
                  // the derived test "lb == 0 && ub == 1" (Case 3 asks it of every bit, :1024) as two OR-monotone
                  // bits: UB01 = some update brought ub down to <= 1; NOT01 = some update made lb > 0 or ub < 1.
                  // bounds == [0,1]  <=>  UB01 && !NOT01 (lb only grows from 0, ub only shrinks from p-1)
                  WF_UB01 = 0x10, WF_NOT01 = 0x20,
                  WF_HEAVY = 0x80 };  // static: the wire occurs in more than HEAVY_DEG rows (set at reset, never by a rule)
#define HEAVY_DEG 24u

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

#define ECNE_MAX_WORLD 8
#define P2_BIGQ_CAP 1024u   // big linear-system groups one sweep can queue
#define CH_LIST_CAP 8192u   // open rows a chain stretch can track
#define CH_P4_MAX 1024u     // IsZero pairs it accepts
#define CH_LONG_MAX 2048u   // long rows it accepts

// result of the P2 candidate scan of a long row, with the Jacobi round counter at the time of the scan
struct __align__(16) LongP2 {
  unsigned long long hs, hx;
  uint32_t k, w1, gr, bad;
};

struct Status {            // device-resident, read back once per solve
  unsigned long long changed;      // state changes of the current outer round ("successful_steps")
  unsigned long long rounds;       // Jacobi rounds of the single-row sweep
  unsigned long long evals;        // rows visited
  unsigned long long rule_evals;   // rows whose rule set was actually run (not skipped by the filter)
  unsigned int err;                // first error (as -status), 0 = none
  unsigned int p2_cand;            // candidates of the linear-system sweep
  unsigned int rec_overflow;
  unsigned int bepoch;             // number of rounds so far that tightened a bound
  unsigned int prog;               // progress counter of the outer fixpoint ("successful_steps", mod 2^32)
  unsigned int outer;              // outer rounds executed (:706)
  unsigned int dense_rounds;       // Jacobi rounds that swept every live row (the others are frontier-driven)
  unsigned int p4_fired;           // number of IsZero pairs latched so far (live masks are refreshed when it moves)
  unsigned long long dense_cycles; // SM cycles spent in dense rounds (block 0), for the roofline of the sweep
  unsigned long long dense_evals;  // rows visited by dense rounds
  unsigned long long prof[8];      // block-0 cycles: dense, sparse, p2 scan, p2 resolve, p3 claim, p3/p4, p0+replay, -
  unsigned int n_cand_total, n_cand_max;  // P2 candidates over the solve / largest scan
  unsigned int solo[8];            // where the grid resumes after block 0 ran rounds alone: n, list, rbuf, round, bepoch, gr
  // rows the linear-system sweep left open, by parity of the outer round that counted them (the other slot is cleared
  // for the next round): block 0 takes whole outer rounds over once this is small (kernels.cu "chain stretch")
  unsigned int p2_full;            // some table slot of this sweep has as many members as its sets have unknowns: the
                                   // resolve pass has something to look at (else it, and its barrier, are skipped)
  unsigned int p2_big_n;           // linear-system groups with more than ECNE_P2_KMAX unknowns queued by the resolvers
  unsigned int p2_open_n[2];
  // where the grid resumes after a chain stretch: position (CH_*), outer, prog_prev, n_pl, stop
  unsigned int chain[6];
};
enum : unsigned int { CH_NONE = 0, CH_MIDP1 = 1, CH_DONE = 2 };

struct Dev {
  // sizes
  uint32_t N, V, nnz, n_long, n_specials, n_known, n_targets, n2a, n2b, table_n;
  uint32_t r0, r1, rpm1;  // ranks of 0, 1, p-1
  int secp_solve;
  // secp_solve: some (BigMultModP, BigLessThan) pair lies in the same sets of the equal-wire DSU (:634-678, built at
  // set-up, csrc/setup.cu k_dsu_*) while the BigLessThan has no output: `constraint_j[3][1]` is a BoundsError (:768)
  int p0p_bounds;
  // rows
  const uint32_t* seg;
  const uint32_t* col;
  const fr::u256* coef;
  const uint32_t* rflags;
  const RowAux* aux;
  const uint32_t* long_rows;
  uint8_t* long_done;        // per long row: can never fire again
  const RowRec* rec;
  uint8_t* solved;
  // static values
  const fr::u256* roots;  // [n2a][2]
  const fr::u256* tvals;  // [n2b]
  const fr::u256* table;  // [table_n] sorted distinct bound values
  // wires (double buffered)
  uint8_t* F[2];
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
  uint32_t* sp_tag;  // [V + 2] lowest index of the specials that set the wire in the running P0 pass (0xffffffff: none)
  const uint32_t* known;
  const uint32_t* targets;
  // wire -> rows index (every wire but the constant wire 1): the frontier of a sparse round
  const uint32_t* inv_ptr;  // [V + 2]
  const uint32_t* inv_row;  // [inv_ptr[V + 1]]
  const uint4* inv_head;    // [V + 2] {rows listed, first three rows}: one load for almost every wire
  unsigned int* long_stamp; // [n_long] last round that queued the long row (dedupe inside a sparse round)
  LongP2* long_p2;          // [n_long] the row's last P2 scan (reused while none of its wires has changed)
  // rows with the P3 / P4 shape (static lists)
  const uint32_t* p3_rows;
  const uint32_t* p4_rows;
  uint32_t n_p3, n_p4;
  // P2 grouping: open-addressing table keyed by the 64-bit hash of the unknown set
  unsigned long long* h_key;  // [h_mask + 1], 0 = empty
  uint32_t* h_cnt;            // members of the group
  uint32_t* h_head;           // 1 + index of the last inserted member (0 = none)
  uint32_t* p2_next;          // [N] member list links
  uint32_t* p2_slot;          // [N] table slot of each candidate
  uint32_t* p2_k;             // [N] size of its unknown set
  uint32_t* p2_open;          // [(N + 31) / 32] bit per row: the linear-system sweep still has to look at it
  uint32_t* p2_bigq;          // [P2_BIGQ_CAP] candidate index of the resolver of each queued big group
  uint32_t* p2_list;          // [CH_LIST_CAP] the open short rows as a list, while block 0 runs the outer rounds alone
  long long chain_open_max;   // engine knob "chain_open_max": block 0 takes outer rounds over below this many open rows
  uint32_t h_mask;
  unsigned long long p2_hash_mask;  // testing knob "p2_hash_bits": fewer hash bits force set-hash collisions
  uint32_t sparse_max;        // a round with at most this many frontier records is frontier-driven
  uint32_t max_outer;
  // records: three rotating lists for the Jacobi rounds + lists 3 / 4 for the phase updates of odd /
  // even outer rounds
  Rec* recs[5];
  unsigned int* rec_count;  // [5]
  unsigned int* bnd_flag;   // [5] set when a round tightened any bound
  uint32_t* c5sig;          // [N] state signature of the last failed Case-5 evaluation of a row
  uint32_t rec_cap;
  // scratch
  unsigned long long* abz_claim;  // [V+1]
  uint32_t* p2_row;               // [N]
  unsigned int* barrier;          // grid barrier counter
  unsigned long long* prof;       // per-round / per-block cycle counters (ECNE_PROFILE builds, ECNE_DEBUG_PROF)
  Status* st;
  // sharding: this rank sweeps rows [row_lo, row_hi)
  uint32_t row_lo, row_hi;
  // multi-GPU exchange over NVLink peer mappings (world == 1: unused)
  int world, rank;
  // shard != 0: the dense sweeps are split over the ranks (row ranges, one exchange per sharded round).  A run on
  // several GPUs whose problem is too small for that to pay (ecne_set_option "shard_min_rows") has shard == 0: every
  // rank runs the whole solve on all rows, nothing is exchanged, rank 0 counts the work.
  int shard;
  Rec* xrecs[ECNE_MAX_WORLD][3];             // every rank's three record lists (own entry = recs[])
  unsigned long long* xflag[ECNE_MAX_WORLD]; // every rank's mailbox [ECNE_MAX_WORLD]; we post into slot [rank]
  unsigned int* xcnt;                        // [3][ECNE_MAX_WORLD] record counts per list and rank (local copy)
  unsigned int* xepoch;                      // cross-GPU barrier epoch, persists over the launches of a solve
  // sharded runs: per-list "this wire already has a record in the list" bytes and the count of distinct
  // wires a round changed.  Unlike the record count (two racing rows log a wire once or twice) that count
  // is the same on every rank, so every rank takes the same sharded / replicated decision for a round.
  uint8_t* wflag[5];
  unsigned int* dcnt;                        // [8]
};

// ---- state access -----------------------------------------------------------------------------------
// The buffer a round READS is not written during that round (updates go to the other buffer), and
// rounds are separated by a grid barrier whose acquire makes remote writes visible (and drops the
// SM's L1 lines) — so ordinary L1-cacheable loads are correct here, exactly as after grid.sync(),
// and neighbouring rows that touch neighbouring wires share L1 sectors instead of paying one L2
// sector per state byte.
//
// The exception is the stretch of rounds that ONE warp runs alone (kernels.cu warp_solo): it separates its rounds with
// __syncwarp only — no fence, no acquire, so nothing drops its SM's L1 lines — and therefore reads the state through
// the L2 (ld.global.cg), where its own atomics of the round before have been performed.  The mode travels as a tag in
// the top bit of the state pointer (st_F / st_L / st_U below make the pointers from a buffer index whose bit 1 is the
// tag), so the evaluators need no second copy and the other paths pay one sign test.
#define ST_TAG_CG 2  // bit of a buffer index: read the state through the L2
__device__ __forceinline__ uint32_t ld_flag(const uint8_t* F, uint32_t w) {
  uint32_t v;
  if ((long long)(uintptr_t)F < 0) {
    const uint8_t* p = (const uint8_t*)((uintptr_t)F & 0x7fffffffffffffffULL) + w;
    asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v) : "l"(p));
  } else {
    asm volatile("ld.global.ca.u8 %0, [%1];" : "=r"(v) : "l"(F + w));
  }
  return v;
}
__device__ __forceinline__ uint32_t ld_u32(const uint32_t* p, uint32_t i) {
  uint32_t v;
  if ((long long)(uintptr_t)p < 0) {
    const uint32_t* q = (const uint32_t*)((uintptr_t)p & 0x7fffffffffffffffULL) + i;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(q));
  } else {
    asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p + i));
  }
  return v;
}

__device__ __forceinline__ void raise(const Dev& d, int status) {
  atomicCAS(&d.st->err, 0u, (unsigned int)(-status));
}

// OR bits into a byte of F; returns the byte as it was before
__device__ __forceinline__ uint32_t or_flag_old(uint8_t* F, uint32_t w, uint32_t bits) {
  unsigned int* word = (unsigned int*)(F + (w & ~3u));
  unsigned int sh = (w & 3u) * 8;
  unsigned int old = atomicOr(word, bits << sh);
  return (old >> sh) & 0xffu;
}
// OR bits into a byte of F; returns the bits that were newly set
__device__ __forceinline__ uint32_t or_flag(uint8_t* F, uint32_t w, uint32_t bits) {
  unsigned int* word = (unsigned int*)(F + (w & ~3u));
  unsigned int sh = (w & 3u) * 8;
  unsigned int old = atomicOr(word, bits << sh);
  return bits & ~((old >> sh) & 0xffu);
}

__device__ __forceinline__ bool is01(uint32_t f) { return (f & (WF_UB01 | WF_NOT01)) == WF_UB01; }

// ---- cross-GPU exchange header (one per rank, mapped into every peer) ---------------------------------
// u64 words: [XH_POST + 16 * (epoch & 1) + r]     rank r's post of a sharded round: {epoch, bound bit, record count}
//            [XH_POST + 16 * (epoch & 1) + 8 + r] ... its second word: {epoch, heavy bit, distinct wires changed}
//            [XH_ACK + r]                          the last epoch whose records rank r has finished pulling from us
// Posts alternate between two slots by epoch parity: a rank can only post epoch e + 2 after every peer has posted
// e + 1, which a peer does after it has read e — so a post is never overwritten before it has been read.  Epochs
// grow monotonically over the life of the process (nothing is ever cleared, no start-of-solve rendezvous needed).
#define XH_POST 0
#define XH_ACK 32
#define XH_WORDS 64        // u64 words of mailbox
#define XH_XCNT_OFF 1024   // byte offset of xcnt u32[4][ECNE_MAX_WORLD]
#define XH_EPOCH_OFF 2048  // byte offset of the persistent epoch counter
#define XH_BYTES 4096

__device__ __forceinline__ void spin_guard(const Dev& d, unsigned long long& spins) {
  if (++spins > (1ULL << 24)) {  // (~10 s) a peer died or diverged: fail loudly instead of hanging the box
    atomicCAS(&d.st->err, 0u, (unsigned int)(-ECNE_E_NCCL));
    spins = 0;
  }
}

// Cross-GPU part of the round barrier, executed by the last local arriver only: post
// {epoch, bound bit, own record count} into every peer's mailbox (NVLink peer store), then wait until
// every peer has posted the same epoch into ours.  Returns the total record count (bit 31: a bound
// moved somewhere) and leaves the per-rank counts in d.xcnt[list][*].
__device__ __forceinline__ unsigned int cross_gpu_exchange(const Dev& d, unsigned int list,
                                                           unsigned int own_payload, unsigned int xe) {
  const unsigned long long word = ((unsigned long long)xe << 32) | own_payload;
  // second word: distinct wires this rank's rows changed (bit 31: a heavy wire among them)
  unsigned int own_d = *((volatile const unsigned int*)(d.dcnt + list)) & 0x7fffffffu;
  if (*((volatile const unsigned int*)(d.bnd_flag + list)) & 2u) own_d |= 0x80000000u;
  const unsigned long long word2 = ((unsigned long long)xe << 32) | own_d;
  const unsigned int par = XH_POST + 16u * (xe & 1u);
  asm volatile("fence.acq_rel.sys;" ::: "memory");
  for (int h = 0; h < d.world; ++h) {
    unsigned long long* slot = d.xflag[h] + par + d.rank;
    asm volatile("st.relaxed.sys.u64 [%0], %1;" ::"l"(slot + 8), "l"(word2) : "memory");
    asm volatile("st.release.sys.u64 [%0], %1;" ::"l"(slot), "l"(word) : "memory");
  }
  unsigned int total = 0, bnd = 0, dtot = 0, heavy = 0;
  for (int h = 0; h < d.world; ++h) {
    const unsigned long long* slot = d.xflag[d.rank] + par + h;
    unsigned long long v, v2;
    unsigned long long spins = 0;
    do {
      asm volatile("ld.acquire.sys.u64 %0, [%1];" : "=l"(v) : "l"(slot) : "memory");
      spin_guard(d, spins);
      if (*((volatile unsigned int*)&d.st->err)) break;
    } while ((unsigned int)(v >> 32) != xe);
    asm volatile("ld.relaxed.sys.u64 %0, [%1];" : "=l"(v2) : "l"(slot + 8) : "memory");  // stored before v
    unsigned int nh = (unsigned int)v & 0x7fffffffu;
    bnd |= (unsigned int)v & 0x80000000u;
    d.xcnt[list * ECNE_MAX_WORLD + h] = nh;
    total += nh;
    dtot += (unsigned int)v2 & 0x7fffffffu;
    heavy |= (unsigned int)v2 & 0x80000000u;
  }
  d.xcnt[3 * ECNE_MAX_WORLD + 0] = dtot;
  d.xcnt[3 * ECNE_MAX_WORLD + 1] = heavy ? 1u : 0u;
  if (total > 0x7fffffffu) total = 0x7fffffffu;
  return total | bnd;
}
// "I have pulled your records of epoch xe": posted to every peer by one thread after the pull's grid barrier
__device__ __forceinline__ void cross_gpu_ack(const Dev& d, unsigned int xe) {
  asm volatile("fence.acq_rel.sys;" ::: "memory");
  for (int h = 0; h < d.world; ++h)
    if (h != d.rank) {
      unsigned long long* slot = d.xflag[h] + XH_ACK + d.rank;
      asm volatile("st.release.sys.u64 [%0], %1;" ::"l"(slot), "l"((unsigned long long)xe) : "memory");
    }
}
// Every peer has pulled our records of epoch xe: the lists they live in may be overwritten from now on.
__device__ __forceinline__ void cross_gpu_wait_acks(const Dev& d, unsigned int xe) {
  for (int h = 0; h < d.world; ++h) {
    if (h == d.rank) continue;
    const unsigned long long* slot = d.xflag[d.rank] + XH_ACK + h;
    unsigned long long v, spins = 0;
    do {
      asm volatile("ld.acquire.sys.u64 %0, [%1];" : "=l"(v) : "l"(slot) : "memory");
      spin_guard(d, spins);
      if (*((volatile unsigned int*)&d.st->err)) break;
    } while ((unsigned int)v < xe);  // epochs of one process never wrap in practice (2^32 sharded rounds)
  }
}

#ifdef ECNE_PROFILE
static __device__ long long g_prof_dummy;
#define g_prof_intra prof_intra
#define g_prof_grid prof_grid
#endif
// Grid-wide barrier for a cooperatively launched (co-resident) grid, one arrival per block.
// `bar[0]` counts arrivals; the LAST arriver publishes {epoch, payload} in `bar[32..33]` (another
// 128-byte line), which is what everybody else polls — the contended atomic line is never polled.
// The payload is the value of *payload_src read after every block has arrived (the round's record
// count = the device-wide changed flag; bit 31 = "a bound was tightened this round"), so no second
// round trip is needed to learn it.
__device__ __forceinline__ unsigned int grid_barrier(unsigned int* bar, unsigned int& epoch,
                                                     const unsigned int* payload_src,
                                                     const unsigned int* flag_src = nullptr,
                                                     const Dev* xd = nullptr, unsigned int xlist = 0,
                                                     unsigned int xe = 0
#ifdef ECNE_PROFILE
                                                     , long long* pprof = nullptr
#endif
) {
#ifdef ECNE_PROFILE
  long long prof_intra = 0, prof_grid = 0;
#endif
  __shared__ unsigned int s_payload;
#ifdef ECNE_PROFILE
  long long ta = clock64();
#endif
  __syncthreads();
#ifdef ECNE_PROFILE
  long long tb = clock64();
#endif
  if (threadIdx.x == 0) {
    epoch += 1;
    // arrive: one release-RMW (orders this block's writes — bar.sync above made them cumulative)
    unsigned int old;
    asm volatile("atom.add.release.gpu.u32 %0, [%1], 1;" : "=r"(old) : "l"(bar) : "memory");
    unsigned long long* rel = (unsigned long long*)(bar + 32);
    unsigned int payload;
    if (old == epoch * gridDim.x - 1) {
      // last arriver: everybody's records are visible (acquire through the RMW chain)
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      payload = payload_src ? *((volatile const unsigned int*)payload_src) : 0u;
      if (flag_src && (*((volatile const unsigned int*)flag_src) & 1u)) payload |= 0x80000000u;
      if (xd) payload = cross_gpu_exchange(*xd, xlist, payload, xe);
      unsigned long long v = ((unsigned long long)payload << 32) | epoch;
      asm volatile("st.release.gpu.u64 [%0], %1;" ::"l"(rel), "l"(v) : "memory");
    } else {
      unsigned long long v;
      do {
        asm volatile("ld.acquire.gpu.u64 %0, [%1];" : "=l"(v) : "l"(rel) : "memory");
      } while ((unsigned int)(v & 0xffffffffu) < epoch);
      payload = (unsigned int)(v >> 32);
    }
    s_payload = payload;
#ifdef ECNE_PROFILE
    g_prof_intra += tb - ta;
    g_prof_grid += clock64() - tb;
#endif
  }
  __syncthreads();
#ifdef ECNE_PROFILE
  if (threadIdx.x == 0 && pprof) {
    pprof[0] += prof_intra;
    pprof[1] += prof_grid;
  }
#endif
  return s_payload;
}

// Fast grid-wide barrier for the single-GPU case (same protocol as cooperative_groups' grid.sync():
// one release-RMW per block on a counter whose top bit flips when the last block arrives — block 0
// adds 2^31 - (grid - 1), everybody else 1 — polled with acquire loads).  Measured on B200 with
// 148 x 1024 threads: 1.3 us, against 2.5 us for the publish-style barrier above, which is kept for
// sharded runs (its last arriver performs the cross-GPU exchange before it releases the others).
// Values the next step needs (record counts) are simply loaded from L2 by every thread afterwards.
__device__ __forceinline__ void grid_sync_flip(unsigned int* bar) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int nb = blockIdx.x == 0 ? 0x80000000u - (gridDim.x - 1u) : 1u;
    unsigned int old, cur;
    asm volatile("atom.add.release.gpu.u32 %0, [%1], %2;" : "=r"(old) : "l"(bar), "r"(nb) : "memory");
    do {
      asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(cur) : "l"(bar) : "memory");
    } while (((old ^ cur) & 0x80000000u) == 0u);
  }
  __syncthreads();
}

}  // namespace ecne
