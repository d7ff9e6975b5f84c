// kernels.cu — the propagation fixpoint (/root/reference/src/R1CSConstraintSolver.jl:706-1556) as ONE
// persistent cooperative kernel, plus the small finalisation / reset kernels.
//
//   k_solve       P0 -> Jacobi rounds of the single-row rules (:805-1349) to a fixpoint -> P2 -> P3 -> P4,
//                 repeated while anything changed, all inside one launch: dense rounds sweep every live
//                 row from shared-memory-resident row records, sparse rounds are driven by the previous
//                 round's update records through a wire -> rows index; every step ends in a grid barrier
//                 whose release carries the device-wide count the next step needs
//   k_pack/...    verdict (:1558-1597) + packed bitmaps for the D2H
#include <algorithm>
#include <cstring>

#include "engine_host.h"
#include "sweep.cuh"

namespace ecne {

// One block per SM.  512 threads x 128 registers: the latency-bound paths (warp_solo, sparse_round, the row
// evaluators) compile without register spills — at 1024 x 64 they spill, and every reload is an L2 trip after a
// round boundary's L1 invalidation (-DP1_THREADS=1024 -DP1_MAX_KS=6 -DP1_INFLIGHT=2 builds that variant).
#ifndef P1_THREADS
#define P1_THREADS 512
#endif
#define P1_MIN_BLOCKS 1
#ifndef P1_MAX_KS
#define P1_MAX_KS (6144 / P1_THREADS)  // rows per thread resident in shared memory: 6144 x 32 B = 192 KB of the 227 KB
#endif
#ifndef P1_INFLIGHT
#define P1_INFLIGHT (P1_THREADS >= 1024 ? 2 : 4)  // rows a thread of the dense sweep has in flight
#endif

// A thread's set of rows that can still fire, one bit per row it owns (row = tid + k * nthreads), in W 64-bit
// registers.  148 x 1024 threads x 128 bits cover 19.4 M rows, 148 x 512 threads x 192 bits 14.5 M: the synthetic
// roofline input S16 (11.1 M rows) runs on the masked fast path in either build.
#ifndef P1_LIVE_WORDS
#define P1_LIVE_WORDS (P1_THREADS >= 1024 ? 2 : 3)
#endif
template <int W>
struct LiveMaskT {
  unsigned long long w[W];
  static constexpr uint32_t BITS = 64u * W;
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int i = 0; i < W; ++i) w[i] = 0;
  }
  __device__ __forceinline__ void set(int k) {
#pragma unroll
    for (int i = 0; i < W; ++i)
      if ((k >> 6) == i) w[i] |= 1ULL << (k & 63);
  }
  __device__ __forceinline__ void clear(int k) {
#pragma unroll
    for (int i = 0; i < W; ++i)
      if ((k >> 6) == i) w[i] &= ~(1ULL << (k & 63));
  }
  __device__ __forceinline__ bool test(int k) const {
    unsigned long long v = 0;
#pragma unroll
    for (int i = 0; i < W; ++i)
      if ((k >> 6) == i) v = w[i];
    return ((v >> (k & 63)) & 1ULL) != 0;
  }
  __device__ __forceinline__ bool any() const {
    unsigned long long v = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) v |= w[i];
    return v != 0;
  }
  __device__ __forceinline__ unsigned int count() const {
    unsigned int c = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) c += (unsigned int)__popcll(w[i]);
    return c;
  }
  __device__ __forceinline__ int pop() {  // lowest row of the set, removed; -1 when empty
    int k = -1;
#pragma unroll
    for (int i = 0; i < W; ++i)
      if (k < 0 && w[i]) {
        k = 64 * i + __ffsll((long long)w[i]) - 1;
        w[i] &= w[i] - 1;
      }
    return k;
  }
  __device__ __forceinline__ int pop_nonempty() { return pop(); }
};
typedef LiveMaskT<P1_LIVE_WORDS> LiveMask;

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}
__device__ __forceinline__ void p2_candidate(const Dev& d, uint32_t row, unsigned long long hs,
                                             unsigned long long hx, uint32_t k) {
  // candidate list entry + membership in the group of its unknown set (open-addressing table keyed by
  // the 64-bit set hash; the member list is a stack linked through p2_next)
  // bits 1..7 of the key hold min(k, 127): the members of a slot share k, so "fewer than k members" is an exact
  // reason to skip the slot even when two different unknown sets collide in it (bit 0: 0 = empty slot)
  const unsigned long long key = ((mix64(hs ^ (hx * 0x9e3779b97f4a7c15ULL) ^ k) & d.p2_hash_mask) << 8) |
                                 ((unsigned long long)(k < 127u ? k : 127u) << 1) | 1ULL;
  // warp-aggregated slot allocation (every still-open row is a candidate in every outer round)
  const unsigned int am = __activemask();
  const unsigned int ln = threadIdx.x & 31u;
  const int leader = __ffs((int)am) - 1;
  unsigned int i = 0;
  if ((int)ln == leader) i = atomicAdd(&d.st->p2_cand, (unsigned int)__popc(am));
  i = __shfl_sync(am, i, leader) + (unsigned int)__popc(am & ((1u << ln) - 1u));
  d.p2_row[i] = row;
  d.p2_k[i] = k;
  uint32_t slot = (uint32_t)(key >> 17) & d.h_mask;
  while (true) {
    const unsigned long long prev = atomicCAS(d.h_key + slot, 0ULL, key);
    if (prev == 0ULL || prev == key) break;
    slot = (slot + 1) & d.h_mask;
  }
  d.p2_slot[i] = slot;
  const uint32_t before = atomicAdd(d.h_cnt + slot, 1u);
  d.p2_next[i] = atomicExch(d.h_head + slot, i + 1);
  // a set can only be decided once k rows share it: until some slot holds k members nobody has anything to resolve
  // (two sets colliding in a slot can only make this fire early, never late)
  if (before + 1u >= k && !__ldcg(&d.st->p2_full)) d.st->p2_full = 1u;
}

// P2 qualification of one row through the CSR (generic path): every non-unique wire appears in C
// only (:1364-1385).  k == 1 is decided on the spot; k >= 2 rows become (set-hash, row) candidates.
// `cache` (long rows, single GPU): the scan of a long row walks up to 1025 terms, and its outcome only
// depends on which of the row's wires are unique — so it is reused until a Jacobi round has looked at the
// row again (long_stamp: a sparse round queued it because one of its wires changed; `dense_gr`: a dense
// round swept everything).
template <int G>
__device__ __noinline__ void p2_scan_row(const Dev&, int rbuf, int pl, uint32_t row, LongP2* cache = nullptr,
                                         unsigned int stamp = 0, unsigned int dense_gr = 0, unsigned int gr = 0) {
  const Dev& d = c_dev;
  const uint32_t lane = Grp<G>::lane();
  uint32_t k = 0, w1 = 0, bad = 0;
  unsigned long long hs = 0, hx = 0;
  bool reuse = false;
  if (cache) {
    const LongP2 c = *cache;
    if (c.gr != 0 && stamp < c.gr && dense_gr < c.gr) {
      reuse = true;
      k = c.k;
      w1 = c.w1;
      hs = c.hs;
      hx = c.hx;
      bad = c.bad;
    }
  }
  if (!reuse) {
    const uint8_t* F = d.F[rbuf];
    const uint32_t s0 = d.seg[3 * row], s2 = d.seg[3 * row + 2], s3 = d.seg[3 * row + 3];
    scan_terms<G, SCAN_U(G)>(d, F, s0, s2, lane, [&](uint32_t, uint32_t f) { bad |= (f & WF_U) ? 0u : 1u; });
    bad = Grp<G>::any(bad != 0) ? 1u : 0u;  // a non-unique wire in A or B (:1366, :1378)
    if (!bad) {
      scan_terms<G, SCAN_U(G)>(d, F, s2, s3, lane, [&](uint32_t w, uint32_t f) {
        if (!(f & WF_U)) {
          ++k;
          w1 = w;
          unsigned long long m = mix64(w);
          hs += m;
          hx ^= mix64(m + 0x9e3779b97f4a7c15ULL);
        }
      });
      if (G > 1) {
        k = Grp<G>::sum(k);
        w1 = Grp<G>::max(w1);
        for (int o = 16; o > 0; o >>= 1) {
          hs += __shfl_xor_sync(0xffffffffu, hs, o);
          hx ^= __shfl_xor_sync(0xffffffffu, hx, o);
        }
      }
    }
    if (cache && lane == 0) {
      LongP2 c;
      c.hs = hs;
      c.hx = hx;
      c.k = k;
      c.w1 = w1;
      c.gr = gr + 1;  // scanned after round gr: anything stamped up to gr is included
      c.bad = bad;
      *cache = c;
    }
  }
  if (bad || k == 0 || lane != 0) return;
  if (k == 1) {  // 1x1 "matrix": the stored coefficient is non-zero (:1402)
    emit(d, 1, pl, w1, WF_U | WF_K);
    return;
  }
  p2_candidate(d, row, hs, hx, k);
}

// One inline row (<= 6 wires, all in registers) against the snapshot.  Returns true when the row can
// never fire again.  Rows whose pattern the fast path does not know go through the generic evaluator.
//   plain rows      Case 1 (+ Case 2a when C is empty); Cases 5/6 only when their cheap gates pass
//   2B rows         x = const: applied once (the monotone merge makes it permanent), then retired
//   4A rows         x - y = 0: Case 1 + bound intersection, ranks only loaded when a wire's bounds
//                   were ever tightened (WF_BND); the l == 2 Case-3 pattern these rows also match is
//                   subsumed by exactly these two steps
// a record of another rank's list, read over NVLink: never through a (possibly stale) L1 line
__device__ __forceinline__ Rec ld_peer_rec(const Rec* p) {
  Rec r;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.wire), "=r"(r.bits), "=r"(r.lbr), "=r"(r.ubr)
               : "l"(p));
  return r;
}

struct InlineRow {
  uint32_t rf, meta;
  uint32_t c[ROWREC_INLINE];
};
__device__ __forceinline__ void unpack_row(const uint4& q0, const uint4& q1, InlineRow& r) {
  r.rf = q0.x;
  r.meta = q0.y;
  r.c[0] = q0.z;
  r.c[1] = q0.w;
  r.c[2] = q1.x;
  r.c[3] = q1.y;
  r.c[4] = q1.z;
  r.c[5] = q1.w;
}
__device__ __forceinline__ void load_row(const Dev& d, uint32_t row, InlineRow& r) {
  const uint4* rp = reinterpret_cast<const uint4*>(d.rec + row);
  unpack_row(__ldg(rp), __ldg(rp + 1), r);
}
__device__ __forceinline__ void gather_row(const uint8_t* F, const InlineRow& r, uint32_t* f) {
  const uint32_t nT = (r.meta & 0xffu) + ((r.meta >> 8) & 0xffu);
#pragma unroll
  for (int j = 0; j < ROWREC_INLINE; ++j)
    f[j] = ((r.rf & RF_FAST) && (uint32_t)j < nT) ? ld_flag(F, r.c[j]) : (WF_U | WF_K | WF_ABZ);
}
// Returns EI_DONE when the row can never fire again, EI_GENERIC when the caller still has to run the
// generic evaluator on it (a dense sweep defers those so that one such lane does not stall its warp in
// every iteration), 0 otherwise.
//
// The updates an evaluation wants are not emitted where the cases find them: they are collected — at most two wires per
// inline row, the effects of several cases on one wire merged (flags by OR, bounds by intersection: exactly what the
// state merge would make of them) — and emitted by the CALLER at one place (emit_pending).  The lanes of a warp find
// their updates in different cases; emitted there, every case would run the emit routine on its own, one dependent
// atomic round trip after the other.  Emitted together, the atomics of all lanes are in flight at once.
#define EI_DONE 1u
#define EI_GENERIC 2u
struct Pending {  // (scalar fields, no indexing: the struct must live in registers)
  uint32_t w0, b0, l0, u0, w1, b1, l1, u1;
  __device__ __forceinline__ void reset() {
    w0 = w1 = 0;
    b0 = b1 = 0;
    l0 = l1 = ECNE_NO_LB;
    u0 = u1 = ECNE_NO_UB;
  }
  __device__ __forceinline__ bool has0() const { return (b0 | l0 | ~u0) != 0; }
  __device__ __forceinline__ bool has1() const { return (b1 | l1 | ~u1) != 0; }
  __device__ __forceinline__ void add(uint32_t wire, uint32_t b, uint32_t l = ECNE_NO_LB, uint32_t u = ECNE_NO_UB) {
    if (has0() && w0 != wire) {  // (an inline row updates at most two wires)
      w1 = wire;
      b1 |= b;
      l1 = l > l1 ? l : l1;
      u1 = u < u1 ? u : u1;
    } else {
      w0 = wire;
      b0 |= b;
      l0 = l > l0 ? l : l0;
      u0 = u < u0 ? u : u0;
    }
  }
};
__device__ __forceinline__ void emit_pending(const Dev&, int wbuf, int list, const Pending& p) {
  if (p.has0() || p.has1()) emit2_impl(wbuf, list, p.w0, p.b0, p.l0, p.u0, p.w1, p.b1, p.l1, p.u1);
}
__device__ __forceinline__ uint32_t eval_inline(const Dev&, int rbuf, uint32_t row, const InlineRow& r, const uint32_t* f,
                                            uint32_t bepoch, Pending& out) {
  const Dev& d = c_dev;
  (void)bepoch;
  out.reset();
  const uint32_t rf = r.rf;
  if (!(rf & RF_FAST)) {
    if (rf & RF_LONG) return EI_DONE;  // swept by a whole warp instead
    return EI_GENERIC;
  }
  const uint32_t nAB = r.meta & 0xffu, nT = nAB + ((r.meta >> 8) & 0xffu);
  uint32_t nuAB = 0, nuC = 0, wC = 0, kmiss = 0, abzmiss = 0;
#pragma unroll
  for (int j = 0; j < ROWREC_INLINE; ++j) {
    const bool nu = !(f[j] & WF_U);
    if ((uint32_t)j < nAB) {
      nuAB += nu;
    } else if (nu) {
      nuC += 1;
      wC = r.c[j];
      kmiss += (f[j] & WF_K) ? 0u : 1u;
      abzmiss += (f[j] & WF_ABZ) ? 0u : 1u;
    }
  }
  // Case 1 (:827-873)
  if (nuAB == 0 && nuC == 1) {
    out.add(wC, WF_U | WF_K);
    nuC = 0;
  }
  // Case 2a (:875-942): C is empty and every non-constant wire of A, B is v*
  if (rf & (RF_2A | RF_2A_NOVAR)) {
    if (rf & RF_2A_NOVAR) {
      raise(d, ECNE_E_BOUNDS);
      return true;
    }
    uint32_t fv = WF_K, v = 0;
#pragma unroll
    for (int j = 0; j < ROWREC_INLINE; ++j)
      if ((uint32_t)j < nT && r.c[j] != 1) {
        v = r.c[j];
        fv = f[j];
      }
    if (!(fv & WF_K)) {
      if (rf & RF_2A_DIVZ) {
        raise(d, ECNE_E_DIVZERO);
      } else {
        out.add(v, WF_K, ECNE_NO_LB, (rf & RF_2A_BOOL) ? d.r1 : ECNE_NO_UB);
        d.valsrc[v] = VS_2A | d.aux[row].val_idx;
        d.solved[row] |= 1;
      }
    }
    return true;  // fired (equation_solved) or v* already known: never fires again
  }
  // Case 2b (:949-988) on a row with no other pattern: x = t once and for all
  if (rf & RF_2B) {
    if (d.solved[row] & 2) return true;  // already applied (a sparse round may meet the row again)
    const RowAux a = d.aux[row];
    out.add(a.w1, WF_U | WF_K, a.rank_a, a.rank_a);
    d.valsrc[a.w1] = VS_2B | a.val_idx;
    d.solved[row] |= 2;
    return true;  // x is unique now, so Cases 5/6 have no unknown key left either
  }
  // Case 4a (:1078-1146) — and the l == 2 instance of Case 3 (:991-1076) it subsumes — followed by
  // Cases 5 and 6 specialised to the two-term row x - y = 0 (both magnitudes are 1)
  if (rf & RF_4A) {
    const uint32_t k1 = r.c[0], k2 = r.c[1];
    uint32_t l1 = d.r0, u1 = d.rpm1, l2 = d.r0, u2 = d.rpm1;
    bool K1 = f[0] & WF_K, K2 = f[1] & WF_K;
    if ((f[0] | f[1]) & WF_BND) {
      if (f[0] & WF_BND) {
        l1 = ld_u32(st_L(rbuf), k1);
        u1 = ld_u32(st_U(rbuf), k1);
      }
      if (f[1] & WF_BND) {
        l2 = ld_u32(st_L(rbuf), k2);
        u2 = ld_u32(st_U(rbuf), k2);
      }
      if (u1 != u2 || l1 != l2) {
        const uint32_t mn = u1 < u2 ? u1 : u2, mx = l1 > l2 ? l1 : l2;
        if (u1 > mn || l1 < mx) {
          out.add(k1, WF_K, mx, mn);
          K1 = true;
        }
        if (u2 > mn || l2 < mx) {
          out.add(k2, WF_K, mx, mn);
          K2 = true;
        }
        l1 = l2 = mx;  // what the later cases of this evaluation see
        u1 = u2 = mn;
      }
    }
    if (nuC == 2) {
      // Case 5 (:1235-1298): sorted magnitudes are [1, 1] (the smaller wire first), 1 % 1 == 0, so the
      // chain holds iff 1 > ub - lb of the first key, i.e. its bounds have zero (or negative) width;
      // the top test 1 * (ub + 1) <= p always holds.  Needs both keys is_known.
      bool fire = K1 && K2 && (u1 <= l1);
      // Case 6 (:1304-1348): both keys carry the same ABZ tag
      if (!fire && abzmiss == 0) fire = d.abz[k1] == d.abz[k2];
      if (fire) {
        out.add(k1, WF_U | WF_K);
        out.add(k2, WF_U | WF_K);
      }
    }
    return false;  // bounds may still have to travel through this row later
  }
  if (nuC == 0) return true;  // no non-unique wire left in C: Cases 1/5/6 can never fire again
  // Cases 5 / 6 (:1235-1348) need every non-unique key known resp. ABZ-tagged: rare, generic path
  if ((rf & RF_LINEAR) && (kmiss == 0 || abzmiss == 0)) return EI_GENERIC;
  return 0u;
}

// ---- phase bodies (device functions of the one persistent kernel) --------------------------------
// Phase updates are logged in list 3 (the "phase list"): it is the frontier of the next outer round's
// first Jacobi round, so that round only re-evaluates the rows next to what the phases changed.
#define PL0 3  // phase list written by even outer rounds (and the prologue); odd rounds write PL0 + 1

// log a state change that the caller applied to BOTH buffers itself
__device__ __forceinline__ void log_rec(const Dev& d, int pl, uint32_t w, uint32_t bits) {
  if (d.inv_head[w].x > HEAVY_DEG) atomicOr(d.bnd_flag + pl, 2u);
  if (d.shard) {
    unsigned int* word = (unsigned int*)(d.wflag[pl] + (w & ~3u));
    const unsigned int sh = (w & 3u) * 8;
    if (!((atomicOr(word, 1u << sh) >> sh) & 1u)) atomicAdd(d.dcnt + pl, 1u);
  }
  unsigned int i = atomicAdd(d.rec_count + pl, 1u);
  if (i < d.rec_cap) {
    Rec r;
    r.wire = w;
    r.bits = bits;
    r.lbr = ECNE_NO_LB;
    r.ubr = ECNE_NO_UB;
    d.recs[pl][i] = r;
  } else {
    d.st->rec_overflow = 1;
  }
}
__device__ __forceinline__ uint32_t ld_flag_cg(const uint8_t* F, uint32_t w) { return __ldcg(F + w); }

// P0 / P0' (:718-800): special constraints in list order, one block.  Reads buffer 1 through the L2
// (an earlier special's outputs are visible to a later one exactly as in the reference), writes both.
#define SPC_N 64     // specials whose wire lists block 0 keeps in shared memory
#define SPC_IN 1024
#define SPC_OUT 512
struct SpecialsCache {
  uint32_t in_ptr[SPC_N + 1], out_ptr[SPC_N + 1];
  uint32_t in[SPC_IN], out[SPC_OUT];
  uint8_t solved[SPC_N];
  int ok;         // the lists fit
  int has_pairs;  // the list holds a BigMultModP and a BigLessThan (P0' has something to do)
  int p4_dead;    // chain stretches: a P4 pass found every IsZero pair latched or its vk unique — for good (both are monotone)
};
__device__ __forceinline__ void specials_cache_fill(const Dev& d, SpecialsCache& c) {
  const bool fits = d.n_specials <= SPC_N && (d.n_specials == 0 || (d.sp_in_ptr[d.n_specials] <= SPC_IN &&
                                                                     d.sp_out_ptr[d.n_specials] <= SPC_OUT));
  if (threadIdx.x == 0) {
    c.ok = fits ? 1 : 0;
    bool m = false, l = false;
    for (uint32_t i = 0; i < d.n_specials; ++i) {
      m |= d.sp_kind[i] == ECNE_SPECIAL_BIGMULTMODP;
      l |= d.sp_kind[i] == ECNE_SPECIAL_BIGLESSTHAN;
    }
    c.has_pairs = (m && l) ? 1 : 0;
    c.p4_dead = d.n_p4 == 0 ? 1 : 0;
  }
  if (fits) {
    for (uint32_t i = threadIdx.x; i <= d.n_specials; i += blockDim.x) {
      c.in_ptr[i] = d.sp_in_ptr[i];
      c.out_ptr[i] = d.sp_out_ptr[i];
      if (i < d.n_specials) c.solved[i] = 0;
    }
    const uint32_t ni = d.n_specials ? d.sp_in_ptr[d.n_specials] : 0, no = d.n_specials ? d.sp_out_ptr[d.n_specials] : 0;
    for (uint32_t i = threadIdx.x; i < ni; i += blockDim.x) c.in[i] = d.sp_in[i];
    for (uint32_t i = threadIdx.x; i < no; i += blockDim.x) c.out[i] = d.sp_out[i];
  }
  __syncthreads();
}

#define SP_TAG_NONE 0xffffffffu
// Lists too long for block 0's shared-memory copy (SPC_N specials): the fired set of a pass is computed by ROUNDS instead
// of one firing at a time.  In the reference's walk special k fires iff every input was unique before the pass or was
// set in it by a special with a LOWER index: the warps judge every open special in parallel — an input set during this
// pass carries the lowest index that set it (sp_tag) and counts for later specials only —, all that are ready fire
// together, and the judging repeats until nothing is.  The rounds are as many as the longest chain of specials feeding
// each other inside one pass (one or two), not as many as fire (16 tiles of ecdsa: 400 specials, 16 ready at a time:
// 5.9 M of 24.7 M cycles went into re-judging the list after every single firing).  sp_solved: 1 fired in an earlier
// pass, 2 in this one, 3 ready in this round.
__device__ __noinline__ void phase_p0_rounds(int pl) {
  const Dev& d = c_dev;
  __shared__ unsigned int s_ready;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const uint32_t n = d.n_specials;
  bool fired_any = false;
  for (;;) {
    if (threadIdx.x == 0) s_ready = 0;
    __syncthreads();
    for (uint32_t s = warp; s < n; s += nwarps) {
      if (__ldcg(d.sp_solved + s)) continue;
      const uint32_t k0 = d.sp_in_ptr[s], k1 = d.sp_in_ptr[s + 1];
      bool ok = true;
      for (uint32_t k = k0 + lane; k < k1; k += 32) {
        const uint32_t w = d.sp_in[k];
        if (!(ld_flag_cg(d.F[1], w) & WF_U)) {
          ok = false;
        } else if (fired_any) {
          const uint32_t t = __ldcg(d.sp_tag + w);
          ok &= t == SP_TAG_NONE || t < s;
        }
      }
      if (__all_sync(0xffffffffu, ok) && lane == 0) {
        d.sp_solved[s] = 3;
        atomicAdd(&s_ready, 1u);
      }
    }
    __threadfence();
    __syncthreads();
    if (s_ready == 0) break;
    fired_any = true;
    // the outputs that are not unique yet take the lowest index among the specials that set them now ...
    for (uint32_t s = warp; s < n; s += nwarps) {
      if (__ldcg(d.sp_solved + s) != 3) continue;
      for (uint32_t k = d.sp_out_ptr[s] + lane; k < d.sp_out_ptr[s + 1]; k += 32) {
        const uint32_t w = d.sp_out[k];
        if (!(ld_flag_cg(d.F[1], w) & WF_U)) atomicMin(d.sp_tag + w, s);
      }
    }
    __threadfence();
    __syncthreads();
    // ... and are set
    for (uint32_t s = warp; s < n; s += nwarps) {
      if (__ldcg(d.sp_solved + s) != 3) continue;
      for (uint32_t k = d.sp_out_ptr[s] + lane; k < d.sp_out_ptr[s + 1]; k += 32) {
        const uint32_t w = d.sp_out[k];
        const uint32_t nw = or_flag(d.F[1], w, WF_U | WF_K);
        or_flag(d.F[0], w, WF_U | WF_K);
        if (nw) log_rec(d, pl, w, WF_U | WF_K);
      }
      __syncwarp();
      if (lane == 0) {
        d.sp_solved[s] = 2;
        atomicAdd(&d.st->prog, 1u);  // successful_steps += 1 (:731)
      }
    }
    __threadfence();
    __syncthreads();
  }
  if (fired_any) {  // end of the pass: what fired in it is simply "fired", its marks on the wires are taken back
    for (uint32_t s = warp; s < n; s += nwarps) {
      if (__ldcg(d.sp_solved + s) != 2) continue;
      for (uint32_t k = d.sp_out_ptr[s] + lane; k < d.sp_out_ptr[s + 1]; k += 32) d.sp_tag[d.sp_out[k]] = SP_TAG_NONE;
      __syncwarp();
      if (lane == 0) d.sp_solved[s] = 1;
    }
    __threadfence();
    __syncthreads();
  }
}
__device__ __noinline__ void phase_p0(const Dev&, int pl, SpecialsCache& sc) {
  const Dev& d = c_dev;
  // The reference walks the specials in list order, so a special sees the outputs of an earlier one
  // that fired in the same pass (and not those of a later one).  Same here: the warps judge every open
  // special from `start` on against the current state in parallel, the lowest one that can fire fires,
  // and the search resumes behind it — with the wire lists in shared memory; longer lists: phase_p0_rounds.
  __shared__ unsigned int s_first;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const bool cached = sc.ok != 0;
  if (!cached && d.n_specials) phase_p0_rounds(pl);
  uint32_t start = 0;
  while (cached && start < d.n_specials) {
    if (threadIdx.x == 0) s_first = 0xffffffffu;
    __syncthreads();
    for (uint32_t s = start + warp; s < d.n_specials; s += nwarps) {
      if (sc.solved[s]) continue;
      const uint32_t k0 = sc.in_ptr[s], k1 = sc.in_ptr[s + 1];
      bool ok = true;
      for (uint32_t k = k0 + lane; k < k1; k += 32) ok &= (ld_flag_cg(d.F[1], sc.in[k]) & WF_U) != 0;
      if (__all_sync(0xffffffffu, ok) && lane == 0) atomicMin(&s_first, s);
    }
    __syncthreads();
    const uint32_t f = s_first;
    __syncthreads();
    if (f == 0xffffffffu) break;
    const uint32_t o0 = sc.out_ptr[f], o1 = sc.out_ptr[f + 1];
    for (uint32_t k = o0 + threadIdx.x; k < o1; k += blockDim.x) {
      const uint32_t w = sc.out[k];
      const uint32_t nw = or_flag(d.F[1], w, WF_U | WF_K);
      or_flag(d.F[0], w, WF_U | WF_K);
      if (nw) log_rec(d, pl, w, WF_U | WF_K);
    }
    if (threadIdx.x == 0) {
      d.sp_solved[f] = 1;
      sc.solved[f] = 1;
      atomicAdd(&d.st->prog, 1u);  // successful_steps += 1 (:731)
    }
    __threadfence();
    __syncthreads();
    start = f + 1;
  }
  // P0': every (BigMultModP, BigLessThan) pair marks the BigLessThan's first three inputs (:750-800)
  if (threadIdx.x == 0 && sc.has_pairs) {
    for (uint32_t i = 0; i < d.n_specials; ++i) {
      if (d.sp_kind[i] != ECNE_SPECIAL_BIGMULTMODP) continue;
      for (uint32_t j = 0; j < d.n_specials; ++j) {
        if (d.sp_kind[j] != ECNE_SPECIAL_BIGLESSTHAN) continue;
        if (!d.secp_solve) {  // `dsu` only exists under secp_solve (:634-636, :762)
          raise(d, ECNE_E_NODSU);
          return;
        }
        uint32_t ni = d.sp_in_ptr[i + 1] - d.sp_in_ptr[i], nj = d.sp_in_ptr[j + 1] - d.sp_in_ptr[j];
        if (ni < 9 || nj < 6 || d.p0p_bounds) {  // (both BoundsErrors abort the solve: which pair raises first is not observable)
          raise(d, ECNE_E_BOUNDS);
          return;
        }
        for (uint32_t k = 0; k < 3; ++k) {
          const uint32_t w = d.sp_in[d.sp_in_ptr[j] + k];
          const uint32_t nw = or_flag(d.F[1], w, WF_U | WF_K);
          or_flag(d.F[0], w, WF_U | WF_K);
          if (nw) {  // not a successful_step in the reference, but it re-enqueues rows (:786-795)
            log_rec(d, pl, w, WF_U | WF_K);
            atomicAdd(&d.st->prog, 1u);
          }
        }
      }
    }
  }
}

// ---- P2 (:1357-1417): grouping of the candidates by identical unknown set ------------------------
// sorted unknown set of a candidate row (k <= KMAX) with the matching coefficients
__device__ inline uint32_t p2_unknowns(const Dev& d, const uint8_t* F, uint32_t row, uint32_t* vars,
                                       uint32_t* terms, uint32_t kmax) {
  const uint32_t s2 = d.seg[3 * row + 2], s3 = d.seg[3 * row + 3];
  uint32_t k = 0;
  for (uint32_t t = s2; t < s3; ++t) {
    uint32_t w = d.col[t];
    if (!(ld_flag(F, w) & WF_U)) {
      if (k < kmax) {
        uint32_t j = k;  // insertion sort by wire id (:1386)
        while (j > 0 && vars[j - 1] > w) {
          vars[j] = vars[j - 1];
          terms[j] = terms[j - 1];
          --j;
        }
        vars[j] = w;
        terms[j] = t;
      }
      ++k;
    }
  }
  return k;
}

// slow_det (:1389-1400) = sum over ODD permutations (Combinatorics.parity is 0 for even ones) of the products
// m[j][perm[j]], m[j][x] = coefficient of the x-th unknown in the j-th row.  Montgomery products of canonical
// inputs carry a uniform R^-(k-1) factor, which does not change whether the sum is zero.
__device__ __noinline__ bool p2_odd_permutation_sum_nonzero(const Dev& d, uint32_t k,
                                                            const uint32_t (*terms)[ECNE_P2_KMAX]) {
  uint32_t perm[ECNE_P2_KMAX];
  for (uint32_t j = 0; j < k; ++j) perm[j] = j;
  fr::u256 res = fr::make_u256(0, 0, 0, 0);
  while (true) {
    uint32_t inv = 0;
    for (uint32_t x = 0; x < k; ++x)
      for (uint32_t y = x + 1; y < k; ++y) inv += perm[x] > perm[y];
    if (inv & 1) {
      fr::u256 term = d.coef[terms[0][perm[0]]];
      for (uint32_t j = 1; j < k; ++j) term = fr::mul(term, d.coef[terms[j][perm[j]]]);
      res = fr::add(res, term);
    }
    // next permutation (lexicographic)
    int a = (int)k - 2;
    while (a >= 0 && perm[a] > perm[a + 1]) --a;
    if (a < 0) break;
    int b = (int)k - 1;
    while (perm[b] < perm[a]) --b;
    uint32_t tmp = perm[a];
    perm[a] = perm[b];
    perm[b] = tmp;
    for (int x = a + 1, y = (int)k - 1; x < y; ++x, --y) {
      tmp = perm[x];
      perm[x] = perm[y];
      perm[y] = tmp;
    }
  }
  return !fr::is_zero(res);
}

// The members of one table slot whose unknown set is EXACTLY that of member `leader` (1 + candidate index): the
// first k of them in row order (the reference triggers when the k-th row of a set arrives and uses exactly
// those, :1387-1388) decide the set.  Returns false when some member of the slot has a different unknown set,
// i.e. two sets collided on the 64-bit hash.
__device__ __noinline__ bool p2_resolve_set(const Dev& d, int pl, uint32_t head, uint32_t leader, bool only_if_first) {
  const uint8_t* F = d.F[0];
  uint32_t vars[ECNE_P2_KMAX], terms[ECNE_P2_KMAX][ECNE_P2_KMAX], v2[ECNE_P2_KMAX], t2[ECNE_P2_KMAX];
  const uint32_t k = p2_unknowns(d, F, d.p2_row[leader - 1], vars, terms[0], ECNE_P2_KMAX);
  if (k > ECNE_P2_KMAX) {
    raise(d, ECNE_E_UNSUPPORTED);
    return true;
  }
  uint32_t best[ECNE_P2_KMAX];
  uint32_t nb = 0;
  bool all_same = true, before_leader = true;
  for (uint32_t m = head; m != 0; m = __ldcg(d.p2_next + (m - 1))) {
    const uint32_t r = d.p2_row[m - 1];
    bool same = m == leader;
    if (!same) {
      same = p2_unknowns(d, F, r, v2, t2, ECNE_P2_KMAX) == k;
      for (uint32_t x = 0; same && x < k; ++x) same = v2[x] == vars[x];
    }
    if (m == leader) before_leader = false;
    if (!same) {
      all_same = false;
      continue;
    }
    if (only_if_first && before_leader) return all_same;  // an earlier member of the list leads this set
    if (nb < k) {  // keep the k smallest row ids
      uint32_t j = nb++;
      while (j > 0 && best[j - 1] > r) {
        best[j] = best[j - 1];
        --j;
      }
      best[j] = r;
    } else if (r < best[k - 1]) {
      uint32_t j = k - 1;
      while (j > 0 && best[j - 1] > r) {
        best[j] = best[j - 1];
        --j;
      }
      best[j] = r;
    }
  }
  if (nb < k) return all_same;  // fewer than k rows share this unknown set
  for (uint32_t j = 0; j < k; ++j) p2_unknowns(d, F, best[j], v2, terms[j], ECNE_P2_KMAX);
  if (p2_odd_permutation_sum_nonzero(d, k, terms))
    for (uint32_t j = 0; j < k; ++j) emit(d, 1, pl, vars[j], WF_U | WF_K);
  return all_same;
}

// One group of candidates (all members of table slot `slot`), resolved by the member that was inserted last.
// When two different unknown sets share the slot (a 64-bit hash collision) the resolver decides every distinct
// set of the slot by exact comparison — it never guesses and never gives up.
__device__ __noinline__ void p2_resolve_group(const Dev&, int pl, uint32_t c) {
  const Dev& d = c_dev;
  const uint32_t slot = d.p2_slot[c];
  if (__ldcg(d.h_head + slot) != c + 1) return;  // not the group's resolver
  const uint32_t cnt = __ldcg(d.h_cnt + slot);
  if (cnt < d.p2_k[c]) return;  // fewer than k rows in the slot (almost every group); k is part of the key
  if (d.p2_k[c] > ECNE_P2_KMAX) {
    // more unknowns than the enumeration of k! permutations is good for: queued, and decided by a whole block after the
    // resolve barrier (p2_big_slot).  The members of a slot share k (it is part of the key), so this is a property of the slot.
    if (d.p2_k[c] > ECNE_P2_KBIG) {
      raise(d, ECNE_E_UNSUPPORTED);
      return;
    }
    const unsigned int i = atomicAdd(&d.st->p2_big_n, 1u);  // (every counted entry below the cap is written)
    if (i >= P2_BIGQ_CAP)
      raise(d, ECNE_E_UNSUPPORTED);
    else
      d.p2_bigq[i] = c;
    return;
  }
  if (p2_resolve_set(d, pl, c + 1, c + 1, false)) return;
  for (uint32_t m = __ldcg(d.p2_next + c); m != 0; m = __ldcg(d.p2_next + (m - 1)))
    p2_resolve_set(d, pl, c + 1, m, true);
}

// ---- P2 groups with ECNE_P2_KMAX < k <= ECNE_P2_KBIG unknowns: a whole block per table slot -----------------------------
// slow_det (:1389-1400) sums the products over the ODD permutations.  With perm(A) = even + odd and det(A) = even - odd,
// odd = (perm - det) / 2, and the rule only asks whether it is zero (char != 2): perm by Ryser's formula, the 2^k - 1
// column subsets dealt out over the threads of the block (O(2^k k^2 / threads) additions, 2^k k / threads products — the
// reference's own enumeration costs k! * k products), det by Gaussian elimination in shared memory.  Everything in
// Montgomery form.  The member bookkeeping (which rows share the leader's unknown set, the k smallest row ids, one
// leader per distinct set of the slot) is the logic of p2_resolve_set / p2_resolve_group, run by thread 0.
struct P2Big {
  fr::u256 m[ECNE_P2_KBIG * ECNE_P2_KBIG];
  fr::u256 part[32];
  uint32_t vars[ECNE_P2_KBIG], best[ECNE_P2_KBIG];
  uint32_t k, skip;
  int fire;
};
__device__ inline uint32_t p2_big_unknowns(const Dev& d, const uint8_t* F, uint32_t row, uint32_t* vars) {
  const uint32_t s2 = d.seg[3 * row + 2], s3 = d.seg[3 * row + 3];
  uint32_t k = 0;
  for (uint32_t t = s2; t < s3; ++t) {
    const uint32_t w = d.col[t];
    if (ld_flag(F, w) & WF_U) continue;
    if (k < ECNE_P2_KBIG) {
      uint32_t j = k;  // insertion sort by wire id (:1386)
      while (j > 0 && vars[j - 1] > w) {
        vars[j] = vars[j - 1];
        --j;
      }
      vars[j] = w;
    }
    ++k;
  }
  return k;
}
__device__ inline bool p2_big_same_set(const Dev& d, const uint8_t* F, uint32_t row, const uint32_t* vars, uint32_t k) {
  const uint32_t s2 = d.seg[3 * row + 2], s3 = d.seg[3 * row + 3];
  uint32_t n = 0;
  for (uint32_t t = s2; t < s3; ++t) {
    const uint32_t w = d.col[t];
    if (ld_flag(F, w) & WF_U) continue;
    bool found = false;
    for (uint32_t x = 0; x < k; ++x) found |= vars[x] == w;
    if (!found) return false;
    ++n;
  }
  return n == k;  // (the keys of a form are distinct)
}
__device__ __noinline__ void p2_big_slot(const Dev&, int pl, uint32_t c, P2Big& S) {
  const Dev& d = c_dev;
  const uint8_t* F = d.F[0];
  const uint32_t t = threadIdx.x, nt = blockDim.x, lane = t & 31u, warp = t >> 5;
  const uint32_t head = c + 1;
  for (uint32_t leader = head; leader != 0; leader = __ldcg(d.p2_next + (leader - 1))) {
    if (t == 0) {
      S.skip = 0;
      S.fire = 0;
      const uint32_t k = p2_big_unknowns(d, F, d.p2_row[leader - 1], S.vars);
      S.k = k;
      if (k > ECNE_P2_KBIG) {
        raise(d, ECNE_E_UNSUPPORTED);
        S.skip = 2;
      } else {
        uint32_t nb = 0;
        bool before_leader = true;
        for (uint32_t mm = head; mm != 0; mm = __ldcg(d.p2_next + (mm - 1))) {
          const uint32_t r = d.p2_row[mm - 1];
          const bool same = mm == leader || p2_big_same_set(d, F, r, S.vars, k);
          if (mm == leader) before_leader = false;
          if (!same) continue;
          if (leader != head && before_leader) {  // an earlier member of the list leads this set
            S.skip = 1;
            break;
          }
          if (nb < k) {  // keep the k smallest row ids (:1387-1388)
            uint32_t j = nb++;
            while (j > 0 && S.best[j - 1] > r) {
              S.best[j] = S.best[j - 1];
              --j;
            }
            S.best[j] = r;
          } else if (r < S.best[k - 1]) {
            uint32_t j = k - 1;
            while (j > 0 && S.best[j - 1] > r) {
              S.best[j] = S.best[j - 1];
              --j;
            }
            S.best[j] = r;
          }
        }
        if (!S.skip && nb < k) S.skip = 1;  // fewer than k rows share this unknown set
      }
    }
    __syncthreads();
    const uint32_t skip = S.skip, k = S.k;
    if (skip == 2) return;
    if (skip == 0) {
      // the matrix: m[j][x] = coefficient of the x-th unknown in the j-th row, Montgomery form
      for (uint32_t e = t; e < k * k; e += nt) {
        const uint32_t j = e / k, x = e % k, row = S.best[j], w = S.vars[x];
        const uint32_t s2 = d.seg[3 * row + 2], s3 = d.seg[3 * row + 3];
        fr::u256 v = fr::make_u256(0, 0, 0, 0);
        for (uint32_t q = s2; q < s3; ++q)
          if (d.col[q] == w) v = d.coef[q];
        S.m[j * k + x] = fr::to_mont(v);
      }
      __syncthreads();
      // permanent (Ryser): (-1)^k * sum over the non-empty column subsets s of (-1)^|s| * prod_i sum_{j in s} m[i][j]
      fr::u256 acc = fr::make_u256(0, 0, 0, 0);
      for (uint32_t sub = t + 1; sub < (1u << k); sub += nt) {
        fr::u256 prod = fr::mont_one();
        for (uint32_t i = 0; i < k; ++i) {
          fr::u256 r = fr::make_u256(0, 0, 0, 0);
          for (uint32_t mk = sub; mk; mk &= mk - 1) r = fr::add(r, S.m[i * k + (uint32_t)(__ffs((int)mk) - 1)]);
          prod = fr::mul(prod, r);
        }
        acc = ((__popc(sub) ^ k) & 1u) ? fr::sub(acc, prod) : fr::add(acc, prod);
      }
      // block reduction of acc: shuffles inside a warp, the warps' sums through shared memory
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        fr::u256 other;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          other.v[q] = __shfl_xor_sync(0xffffffffu, acc.v[q], o);
        acc = fr::add(acc, other);
      }
      if (lane == 0) S.part[warp] = acc;
      __syncthreads();
      if (t == 0) {
        fr::u256 perm = S.part[0];
        for (uint32_t w2 = 1; w2 < (nt >> 5); ++w2) perm = fr::add(perm, S.part[w2]);
        // determinant: elimination with row swaps, in place (the permanent is done with the matrix)
        fr::u256 det = fr::mont_one();
        bool neg = false, zero = false;
        for (uint32_t col = 0; col < k && !zero; ++col) {
          uint32_t piv = col;
          while (piv < k && fr::is_zero(S.m[piv * k + col])) ++piv;
          if (piv == k) {
            zero = true;
            break;
          }
          if (piv != col) {
            for (uint32_t x = 0; x < k; ++x) {
              const fr::u256 tmp = S.m[col * k + x];
              S.m[col * k + x] = S.m[piv * k + x];
              S.m[piv * k + x] = tmp;
            }
            neg = !neg;
          }
          const fr::u256 pv = S.m[col * k + col];
          det = fr::mul(det, pv);
          const fr::u256 ipv = fr::inv_mont(pv);
          for (uint32_t r = col + 1; r < k; ++r) {
            if (fr::is_zero(S.m[r * k + col])) continue;
            const fr::u256 f = fr::mul(S.m[r * k + col], ipv);
            for (uint32_t x = col; x < k; ++x) S.m[r * k + x] = fr::sub(S.m[r * k + x], fr::mul(f, S.m[col * k + x]));
          }
        }
        if (zero) det = fr::make_u256(0, 0, 0, 0);
        if (neg) det = fr::neg(det);
        S.fire = !fr::eq(perm, det);  // 2 * (odd-permutation sum) = perm - det != 0
      }
      __syncthreads();
      if (S.fire && t < k) emit(d, 1, pl, S.vars[t], WF_U | WF_K);
    }
    __syncthreads();
  }
}

// the queued big groups, group q by block / caller `first + i * stride`; returns false when none was queued
__device__ __noinline__ bool p2_big_phase(const Dev&, int pl, unsigned int first, unsigned int stride) {
  const Dev& d = c_dev;
  __shared__ P2Big s_big;
  unsigned int n_big = __ldcg(&d.st->p2_big_n);  // (uniform: read behind the resolve barrier)
  if (!n_big) return false;
  if (n_big > P2_BIGQ_CAP) n_big = P2_BIGQ_CAP;
  for (unsigned int q = first; q < n_big; q += stride) p2_big_slot(d, pl, __ldcg(d.p2_bigq + q), s_big);
  return true;
}

// ---- P3 (:1425-1483): the lowest row tags a wire, exactly the reference's order -------------------
__device__ __forceinline__ void p3_claim_row(const Dev& d, uint32_t row) {
  const uint32_t rf = d.rflags[row];
  const RowAux a = d.aux[row];
  if (ld_flag(d.F[1], a.w3) & WF_U) return;  // unique_b (:1448)
  if (rf & RF_P3_DIVZ) {                      // divexact(-intercept, 0) (:1467)
    raise(d, ECNE_E_DIVZERO);
    return;
  }
  atomicMin(d.abz_claim + a.w3, ((unsigned long long)row << 32) | a.w4);
}
__device__ __forceinline__ void p3_commit_row(const Dev& d, int pl, uint32_t row) {
  const uint32_t rf = d.rflags[row];
  if (rf & RF_P3_DIVZ) return;
  const RowAux a = d.aux[row];
  if (ld_flag(d.F[0], a.w3) & WF_U) return;
  unsigned long long cl = __ldcg(d.abz_claim + a.w3);
  if ((uint32_t)(cl >> 32) != row) return;  // a lower row tags this wire first
  d.abz_claim[a.w3] = ~0ULL;
  if (d.abz[a.w3] != -1) return;  // (:1469-1473)
  d.abz[a.w3] = (int32_t)a.w4;
  or_flag(d.F[0], a.w3, WF_K | WF_ABZ);
  or_flag(d.F[1], a.w3, WF_K | WF_ABZ);
  log_rec(d, pl, a.w3, WF_K | WF_ABZ);
}
// ---- P4 (:1492-1550) ------------------------------------------------------------------------------
__device__ __forceinline__ bool p4_row(const Dev& d, int pl, uint32_t row) {
  const uint8_t* F = d.F[0];
  if (d.solved[row] & 1) return false;  // fired before: vk is unique
  const uint32_t vk = d.aux[row].w4;
  if (ld_flag(F, vk) & WF_U) return false;
  const uint32_t s0 = d.seg[3 * row], s1 = d.seg[3 * row + 1];
  for (uint32_t t = s0; t < s1; ++t)
    if (!(ld_flag(F, d.col[t]) & WF_U)) return false;  // a_unique (:1502-1511)
  emit(d, 1, pl, vk, WF_U | WF_K);
  d.solved[row] |= 1;  // equation_solved[i], [i+1] (:1541-1542)
  d.solved[row + 1] |= 1;
  return true;
}

// grid barrier, then the value of *src as every block left it before the barrier
__device__ __forceinline__ unsigned int sync_and_load(const Dev& d, const unsigned int* src) {
  grid_sync_flip(d.barrier + 64);
  return src ? __ldcg(src) : 0u;
}

// One short row next to a changed wire, evaluated by one thread (sparse rounds).  Long rows are
// queued once per round (long_stamp) for a whole warp.
#define SP_LONG_CAP 256
#define SP_HEAVY_CAP 64
#define SOLO_MAX 96u    // frontiers up to this size are swept by block 0 alone (one warp per record), without grid barriers
__device__ __forceinline__ void sparse_row(const Dev&, int rbuf, int wbuf, int list, uint32_t row,
                                           uint32_t bepoch, unsigned int gr, uint32_t* s_long,
                                           unsigned int* s_nlong, unsigned long long& evals, bool all_rows) {
  const Dev& d = c_dev;
  (void)all_rows;  // frontier-driven rounds evaluate every listed row on every rank
  if (row == 0xffffffffu) return;
  const uint4* rp = reinterpret_cast<const uint4*>(d.rec + row);
  const uint4 q0 = __ldg(rp), q1 = __ldg(rp + 1);
  const uint32_t latched = __ldcg(d.solved + row);  // in flight together with the record
  InlineRow ir;
  unpack_row(q0, q1, ir);
  if (ir.rf & RF_LONG) {
    const uint32_t li = ir.c[0];
    if (__ldcg(d.long_done + li)) return;
    if (atomicExch(d.long_stamp + li, gr) == gr) return;  // already queued this round
    const unsigned int slot = atomicAdd(s_nlong, 1u);
    if (slot < SP_LONG_CAP) {
      s_long[slot] = row;
    } else {  // queue full: the (slow) one-thread walk is still exact
      evals += 1;
      if (eval_row<1>(d, rbuf, wbuf, list, row, bepoch)) d.long_done[li] = 1;
    }
    return;
  }
  if (latched & 1) return;  // equation_solved (:820-822)
  uint32_t f[ROWREC_INLINE];
  gather_row(d.F[rbuf], ir, f);
  evals += 1;
  Pending pend;
  const uint32_t ei = eval_inline(d, rbuf, row, ir, f, bepoch, pend);
  emit_pending(d, wbuf, list, pend);
  if (ei & EI_GENERIC) eval_row<1>(d, rbuf, wbuf, list, row, bepoch);
}

// A frontier-driven Jacobi round: every record of the previous round (a state change of one wire) is
// taken by one thread, which evaluates the rows the wire -> rows index lists for that wire against
// buffer `rbuf` and then replays the record into the other buffer.  `solo`: block 0 runs the round
// alone (record i -> thread i); otherwise record i -> block i % grid.  Returns this thread's row visits.
__device__ __noinline__ unsigned long long sparse_round(const Dev&, int rbuf, unsigned int list,
                                                        unsigned int prev_list, unsigned int prev_n,
                                                        uint32_t bepoch, unsigned int gr, bool solo,
                                                        unsigned int prev_own) {
  const Dev& d = c_dev;
  __shared__ uint32_t s_long[SP_LONG_CAP];
  __shared__ uint32_t s_heavy[SP_HEAVY_CAP];
  __shared__ unsigned int s_nlong, s_nheavy;
  const int wbuf = rbuf ^ 1;
  // one WARP per record: its lanes take the rows listed for the record's wire (one row each, so the
  // dependent loads of all of them are in flight together) and one more lane replays the record
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const uint32_t first = solo ? warp : blockIdx.x + warp * gridDim.x;
  const uint32_t stride = solo ? nwarps : gridDim.x * nwarps;
  if (threadIdx.x == 0) {
    s_nlong = 0;
    s_nheavy = 0;
  }
  __syncthreads();
  unsigned long long ev = 0;
  // the list is complete locally (after a sharded round the peers' records were appended behind this rank's own
  // `prev_own` ones when they were pulled and applied to both buffers): only the own records still need their replay
  {
    const Rec* pr = d.recs[prev_list];
    const unsigned int nh = prev_n > d.rec_cap ? d.rec_cap : prev_n;
    for (uint32_t i = first; i < nh; i += stride) {
      const bool own = i < prev_own;
#ifdef ECNE_PROFILE
      long long q0 = clock64();
#endif
      const Rec r = ld_peer_rec(pr + i);  // (volatile: written by other SMs during the previous round)
      // {rows listed for the wire, the first three of them}: one 16-byte load covers 97 % of the wires
      const uint4 hd = __ldcg(reinterpret_cast<const uint4*>(d.inv_head) + r.wire);
#ifdef ECNE_PROFILE
      long long q1 = clock64() + (hd.x & 0);
#endif
      bool queued = false;
      if (hd.x > HEAVY_DEG) {
        unsigned int slot = 0;
        if (lane == 0) slot = atomicAdd(&s_nheavy, 1u);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot < SP_HEAVY_CAP) {
          if (lane == 0) s_heavy[slot] = r.wire;
          queued = true;
        }
      }
      if (!queued) {
        for (uint32_t q = lane; q < hd.x; q += 32) {
          uint32_t row;
          if (q == 0) row = hd.y;
          else if (q == 1) row = hd.z;
          else if (q == 2) row = hd.w;
          else row = d.inv_row[d.inv_ptr[r.wire] + q];
          sparse_row(d, rbuf, wbuf, (int)list, row, bepoch, gr, s_long, &s_nlong, ev, solo);
        }
      }
#ifdef ECNE_PROFILE
      long long q2 = clock64();
#endif
      // the replay, by a lane that has no row to evaluate when there is one
      if (own && lane == (hd.x < 31u ? hd.x : 31u)) {
        apply_update(d, wbuf, r.wire, r.bits, r.lbr, r.ubr);
        consume_rec(d, prev_list, r.wire);
      }
#ifdef ECNE_PROFILE
      if (solo && threadIdx.x == 0 && gr < 2000) {
        d.prof[16000 + 4 * gr + 0] = (unsigned long long)(q1 - q0);   // record + head loads
        d.prof[16000 + 4 * gr + 1] = (unsigned long long)(q2 - q1);   // row evaluations
        d.prof[16000 + 4 * gr + 2] = (unsigned long long)(clock64() - q2);  // replay
        d.prof[16000 + 4 * gr + 3] = hd.x;
      }
#endif
    }
  }
  __syncthreads();
  {  // wires with many rows: the whole block strides their row lists
    const unsigned int nhv = s_nheavy < SP_HEAVY_CAP ? s_nheavy : SP_HEAVY_CAP;
    for (unsigned int x = 0; x < nhv; ++x) {
      const uint32_t w = s_heavy[x];
      const uint32_t lo = d.inv_ptr[w], hi = d.inv_ptr[w + 1];
      for (uint32_t q = lo + threadIdx.x; q < hi; q += blockDim.x)
        sparse_row(d, rbuf, wbuf, (int)list, d.inv_row[q], bepoch, gr, s_long, &s_nlong, ev, solo);
    }
  }
  __syncthreads();
  {  // queued long rows: one warp each
    const unsigned int nlq = s_nlong < SP_LONG_CAP ? s_nlong : SP_LONG_CAP;
    for (unsigned int x = threadIdx.x >> 5; x < nlq; x += blockDim.x >> 5) {
      const uint32_t row = s_long[x];
      const bool done = eval_row<32>(d, rbuf, wbuf, (int)list, row, bepoch);
      if ((threadIdx.x & 31u) == 0) {
        ev += 1;
        if (done) d.long_done[d.rec[row].c[0]] = 1;
      }
    }
  }
  return ev;
}

// The P2 test of one short row from its inline record and the six gathered state bytes (:1364-1385): every
// non-unique wire appears in C only; k == 1 is decided on the spot, k >= 2 rows become candidates.  Returns true
// when the sweep never has to look at the row again.
__device__ __forceinline__ bool p2_scan_short(const Dev& d, int pl, uint32_t row, const InlineRow& r, const uint32_t* ff) {
  if (r.rf & RF_LONG) return true;  // long rows are scanned by a warp each, from their own list
  const uint32_t nAB = r.meta & 0xffu;
  uint32_t kk = 0, w1 = 0;
  bool bad = false;
  unsigned long long hs = 0, hx = 0;
#pragma unroll
  for (int j = 0; j < ROWREC_INLINE; ++j) {
    if (!(ff[j] & WF_U)) {
      if ((uint32_t)j < nAB) {
        bad = true;
      } else {
        ++kk;
        w1 = r.c[j];
        unsigned long long mm = mix64(r.c[j]);
        hs += mm;
        hx ^= mix64(mm + 0x9e3779b97f4a7c15ULL);
      }
    }
  }
  if (!bad && kk == 0) return true;  // every wire is unique: the row can never qualify again
  if (bad) return false;             // a non-unique wire in A or B (:1366, :1378): may qualify later
  if (d.solved[row] & 1) return true;  // equation_solved rows take no part (:1360)
  if (kk == 1)
    emit(d, 1, pl, w1, WF_U | WF_K);
  else
    p2_candidate(d, row, hs, hx, kk);
  return false;
}

// Rounds whose frontier is at most 32 records are run by WARP 0 of block 0 alone (most of ecdsa's rounds
// change one or six wires): the (record, listed row) pairs are dealt out one per lane, a long row among
// them is then evaluated by the whole warp, and the round boundary is __syncwarp + release fence + acquire load — no
// block barrier, no shared-memory queues.  Runs rounds until the frontier is empty or outgrows it and
// leaves the state of the LAST round it ran in `st` (shared memory), exactly as a block-solo round would.
#define WARP_SOLO_MAX 32u
struct SoloState {
  unsigned int n, list, rbuf, round, bepoch, gr, hv, prev_list;
  unsigned int dn;  // distinct wires the last round changed (== n on one GPU)
};
// Unsharded runs take the fast path: between two rounds of the stretch there is only a __syncwarp.  The records stay
// in shared memory (sweep.cuh "records of a stretch"), the state is read through the L2 (ST_TAG_CG) where the warp's
// own atomics have been performed — every atomic of a round has returned before the round ends: the updates of the
// write buffer return the old value that decides whether they are logged, the replays into the other buffer are
// issued at the start of the round as returning atomics whose results are consumed at its end — and nothing touches the global
// record lists or counters until the stretch ends (frontier empty or too large, a heavy wire, a spilled queue): then
// the last round's records are written to the global list, one fence makes everything visible, and the state block is
// left exactly as a round through the global lists would leave it.  Sharded runs (which count distinct wires per
// round through the global arrays) keep the round-by-round global protocol.
template <bool fast>  // fast: unsharded run (the records of a stretch stay in shared memory, see above)
__device__ __noinline__ void warp_solo(const Dev&, SoloState* st, unsigned int prev_n, unsigned int max_rounds,
                                       unsigned long long* evals_io) {
  const Dev& d = c_dev;
  const uint32_t lane = threadIdx.x & 31u;
  unsigned int n = prev_n, list = st->list, prev_list = st->prev_list, round = st->round, bepoch = st->bepoch,
               gr = st->gr, hv = 0, dn = 0;
  int rbuf = (int)st->rbuf;
  unsigned long long ev = 0;
  bool from_smem = false;      // the records of the round before are in s_soloq[qr]
  unsigned int qr = 0;         // queue read this round (fast path); the round writes queue qr ^ 1
  unsigned int prog_acc = 0;
  uint32_t replay_sink = 0;    // consumes the results of the replay atomics (see above)
#ifdef ECNE_PROFILE
  const long long wz0 = clock64();
  unsigned int wrounds = 0;
#endif
  while (true) {
    gr += 1;
#ifdef ECNE_PROFILE
    wrounds += 1;
    const long long rz0 = clock64();
    long long rz1 = rz0;
#endif
    const int wbuf = rbuf ^ 1;
    const int rb = fast ? (rbuf | ST_TAG_CG) : rbuf;
    const unsigned int qw = qr ^ 1u;
    const int elist = fast ? (int)(list | LIST_SOLO | (qw ? LIST_SOLO_Q : 0)) : (int)list;
    // lane i holds record i and the head of its wire's row list; the (record, listed row) pairs are then
    // dealt out one per lane, so every row of the frontier is evaluated in the same step
    const bool have = lane < n;
    Rec r;
    r.wire = 1;
    r.bits = 0;
    r.lbr = ECNE_NO_LB;
    r.ubr = ECNE_NO_UB;
    uint4 hd = make_uint4(0, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    if (have) {
      if (from_smem) {
        r = s_soloq[qr][lane].r;
        hd = s_soloq[qr][lane].head;
      } else {
        r = ld_peer_rec(d.recs[prev_list] + lane);
        hd = __ldcg(reinterpret_cast<const uint4*>(d.inv_head) + r.wire);
      }
    }
#ifdef ECNE_PROFILE
    const long long pz1 = clock64() + (r.wire & 0) + (hd.x & 0);
#endif
    if (fast && lane == 0) {
      s_soloq_n[qw] = 0;
      s_solo_flags = 0;
    }
    // fast path: the replay of the record into the other buffer is issued FIRST, as a returning atomic whose result
    // is consumed when the round ends — it has long returned by then, so it is performed at the L2 before the next
    // round (which reads that buffer) starts, and it costs no trip on the round's critical path
    uint32_t rp0 = 0, rp1 = 0, rp2 = 0;
    if (fast && have) apply_update_issue(wbuf, r.wire, r.bits, r.lbr, r.ubr, rp0, rp1, rp2);
#ifdef ECNE_PROFILE
    const long long pz2 = clock64();
#endif
    __syncwarp();
#ifdef ECNE_PROFILE
    const long long pz3 = clock64();
    if (lane == 0) {
      unsigned long long* q = d.prof + 28000 + 40 * 148 * 4 + 128;
      q[16] += (unsigned long long)(pz1 - rz0);  // records from shared memory
      q[17] += (unsigned long long)(pz2 - pz1);  // replay issued
      q[18] += (unsigned long long)(pz3 - pz2);  // __syncwarp
    }
#endif
    unsigned int incl = hd.x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)lane >= o) incl += t;
    }
    const unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
    const unsigned int excl = incl - hd.x;
    for (unsigned int base = 0; base < total; base += 32) {
      const unsigned int p = base + lane;
      int j = 0;  // the record whose range [excl, incl) holds pair p: the number of records whose range ends at or before p
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {  // (binary search over the non-decreasing prefix sums: 5 shuffles)
        const unsigned int e = __shfl_sync(0xffffffffu, incl, j + step - 1);
        if (p >= e) j += step;
      }
      const uint32_t wj = __shfl_sync(0xffffffffu, r.wire, j);
      const uint32_t hy = __shfl_sync(0xffffffffu, hd.y, j), hz = __shfl_sync(0xffffffffu, hd.z, j),
                     hw = __shfl_sync(0xffffffffu, hd.w, j);
      const unsigned int q = p - __shfl_sync(0xffffffffu, excl, j);
      uint32_t row = 0xffffffffu;
      bool is_long = false;
      Pending pend;  // the updates this lane's row wants: emitted by all lanes together, below
      pend.reset();
#ifdef ECNE_PROFILE
      long long sz0 = clock64(), sz1 = sz0, sz2 = sz0, sz3 = sz0, sz4 = sz0;
      bool was_generic = false;
#endif
      if (p < total) {
        if (q == 0) row = hy;
        else if (q == 1) row = hz;
        else if (q == 2) row = hw;
        else row = d.inv_row[d.inv_ptr[wj] + q];
        if (row != 0xffffffffu) {  // every row, also on a sharded run: solo rounds are replicated
          const uint4* rp = reinterpret_cast<const uint4*>(d.rec + row);
          const uint4 q0v = __ldg(rp), q1v = __ldg(rp + 1);
          const uint32_t latched = __ldcg(d.solved + row);
          InlineRow ir;
          unpack_row(q0v, q1v, ir);
#ifdef ECNE_PROFILE
          sz1 = clock64() + (ir.rf & 0) + (latched & 0);
#endif
          if (ir.rf & RF_LONG) {
            const uint32_t li = ir.c[0];
            is_long = !__ldcg(d.long_done + li) && atomicExch(d.long_stamp + li, gr) != gr;
          } else if (!(latched & 1)) {
            uint32_t f[ROWREC_INLINE];
            gather_row(st_F(rb), ir, f);
            ev += 1;
#ifdef ECNE_PROFILE
            sz2 = clock64() + (f[0] & 0) + (f[1] & 0) + (f[2] & 0);
#endif
            const uint32_t ei = eval_inline(d, rb, row, ir, f, bepoch, pend);
#ifdef ECNE_PROFILE
            sz3 = clock64() + (ei & 0);
            was_generic = (ei & EI_GENERIC) != 0;
#endif
            if (ei & EI_GENERIC)
              eval_row<1>(d, rb, wbuf, elist, row, bepoch);
          }
        }
#ifdef ECNE_PROFILE
        sz4 = clock64();
#endif
      }
#ifdef ECNE_PROFILE
        {
          unsigned long long* q = d.prof + 28000 + 40 * 148 * 4 + 128;
          // slowest lane of the batch, stage by stage
          const unsigned int a = __reduce_max_sync(0xffffffffu, (unsigned int)(sz1 - sz0));
          const unsigned int b = __reduce_max_sync(0xffffffffu, (unsigned int)(sz2 > sz1 ? sz2 - sz1 : 0));
          const unsigned int c = __reduce_max_sync(0xffffffffu, (unsigned int)(sz3 > sz2 ? sz3 - sz2 : 0));
          const unsigned int g = __reduce_max_sync(0xffffffffu, (unsigned int)(was_generic ? sz4 - sz3 : 0));
          const unsigned int ng = __popc(__ballot_sync(0xffffffffu, was_generic));
          if (lane == 0) {
            q[8] += a; q[9] += b; q[10] += c; q[11] += g; q[12] += ng;
          }
        }
#endif
      __syncwarp();
#ifdef ECNE_PROFILE
      const long long ez0 = clock64();
#endif
      emit_pending(d, wbuf, elist, pend);
#ifdef ECNE_PROFILE
      {
        __syncwarp();
        const long long ez1 = clock64();
        if (lane == 0) {
          d.prof[28000 + 40 * 148 * 4 + 128 + 13] += (unsigned long long)(ez1 - ez0);  // the emits of the batch, all lanes together
          d.prof[28000 + 40 * 148 * 4 + 128 + 14] += (unsigned long long)(sz0 - rz0);  // round start -> first batch
          rz1 = ez1;
        }
      }
#endif
      unsigned int m = __ballot_sync(0xffffffffu, is_long);
#ifdef ECNE_PROFILE
      const long long lz0 = clock64();
      const unsigned int lm0 = m;
#endif
      while (m) {
        const int src = __ffs((int)m) - 1;
        m &= m - 1;
        const uint32_t lrow = __shfl_sync(0xffffffffu, row, src);
        const bool done = eval_row<32>(d, rb, wbuf, elist, lrow, bepoch);
        if (lane == 0) {
          ev += 1;
          if (done) d.long_done[d.rec[lrow].c[0]] = 1;
        }
      }
#ifdef ECNE_PROFILE
      if (lane == 0) {
        unsigned long long* q = d.prof + 28000 + 40 * 148 * 4 + 128;
        q[0] += (unsigned long long)(clock64() - lz0);  // cycles in long rows
        q[1] += (unsigned long long)__popc(lm0);        // long rows evaluated
        q[2] += 1;                                       // pair batches
        q[3] += total > base + 32 ? 32 : total - base;   // (record, row) pairs
      }
#endif
    }
    if (have && !fast) {  // the replay
      apply_update(d, wbuf, r.wire, r.bits, r.lbr, r.ubr);
      consume_rec(d, prev_list, r.wire);
    }
#ifdef ECNE_PROFILE
    const long long qz0 = clock64();
#endif
    replay_sink |= rp0 ^ rp1 ^ rp2;  // first use of the replay atomics' results: waits until they have been performed
    if (__all_sync(0xffffffffu, replay_sink == 0xdeadbeefu)) ev += 1;  // (keeps the use alive; practically never true)
#ifdef ECNE_PROFILE
    const long long qz1 = clock64();
#endif
    __syncwarp();
#ifdef ECNE_PROFILE
    if (lane == 0) {
      unsigned long long* q = d.prof + 28000 + 40 * 148 * 4 + 128;
      q[19] += (unsigned long long)(qz0 - rz1);      // last emit -> here (long-row ballot)
      q[20] += (unsigned long long)(qz1 - qz0);      // replay results consumed
      q[21] += (unsigned long long)(clock64() - qz1);  // __syncwarp
    }
#endif
    unsigned int cnt = 0, bf = 0;
    bool leave;
    if (fast) {
      const unsigned int flags = s_solo_flags, nq = s_soloq_n[qw];
      __syncwarp();  // every lane has read the round's flags before lane 0 clears them for the next round (racecheck)
      const bool spilled = (flags & 4u) != 0;
      if (lane == 0 && !from_smem) {  // the list the stretch started from is consumed
        d.rec_count[prev_list] = 0;
        d.bnd_flag[prev_list] = 0;
        d.dcnt[prev_list] = 0;
      }
      bf = flags & 3u;
      cnt = nq < SOLO_Q_CAP ? nq : SOLO_Q_CAP;
      leave = spilled || cnt == 0 || cnt > WARP_SOLO_MAX || (bf & 2u) || round + 1 >= max_rounds;
      if (leave) {
        // hand the last round's records over through the global list (behind what spilled there already)
        unsigned int base = 0;
        if (spilled) {
          if (lane == 0) base = atomicAdd(d.rec_count + list, cnt);
          base = __shfl_sync(0xffffffffu, base, 0);
        }
        for (unsigned int i = lane; i < cnt; i += 32) {
          if (base + i < d.rec_cap)
            d.recs[list][base + i] = s_soloq[qw][i].r;
          else
            d.st->rec_overflow = 1;
        }
        __syncwarp();
        if (lane == 0) {
          if (!spilled) d.rec_count[list] = cnt;
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          unsigned int gbf = 0;
          asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(cnt) : "l"(d.rec_count + list) : "memory");
          asm volatile("ld.relaxed.gpu.u32 %0, [%1];" : "=r"(gbf) : "l"(d.bnd_flag + list) : "memory");
          bf |= gbf & 3u;
          atomicAdd(&d.st->prog, prog_acc + cnt);
        }
        cnt = __shfl_sync(0xffffffffu, cnt, 0);
        bf = __shfl_sync(0xffffffffu, bf, 0);
      } else {
        prog_acc += cnt;
      }
      dn = cnt;
    } else {
      if (lane == 0) {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(cnt) : "l"(d.rec_count + list) : "memory");
        asm volatile("ld.relaxed.gpu.u32 %0, [%1];" : "=r"(bf) : "l"(d.bnd_flag + list) : "memory");
        asm volatile("ld.relaxed.gpu.u32 %0, [%1];" : "=r"(dn) : "l"(d.dcnt + list) : "memory");
        d.rec_count[prev_list] = 0;  // consumed; next written two rounds from now
        d.bnd_flag[prev_list] = 0;
        d.dcnt[prev_list] = 0;
        atomicAdd(&d.st->prog, cnt);
      }
      cnt = __shfl_sync(0xffffffffu, cnt, 0);
      bf = __shfl_sync(0xffffffffu, bf, 0);
      dn = __shfl_sync(0xffffffffu, dn, 0);
      leave = false;
    }
#ifdef ECNE_PROFILE
    if (lane == 0) d.prof[28000 + 40 * 148 * 4 + 128 + 15] += (unsigned long long)(clock64() - rz1);  // last emit -> round end
#endif
    bepoch += bf & 1u;
    hv = bf & 2u;
    round += 1;
    n = cnt;
    if (leave || n == 0 || n > WARP_SOLO_MAX || hv || round >= max_rounds) break;
    prev_list = list;
    list = (list + 1) % 3;
    rbuf ^= 1;
    if (fast) {
      from_smem = true;
      qr = qw;
    }
  }
#ifdef ECNE_PROFILE
  if (lane == 0) {
    unsigned long long* q = d.prof + 28000 + 40 * 148 * 4 + 128;
    q[4] += (unsigned long long)(clock64() - wz0);
    q[5] += wrounds;
    q[6] += 1;
  }
#endif
  *evals_io += ev;
  if (lane == 0) {
    st->n = n;
    st->list = list;
    st->rbuf = (unsigned int)rbuf;
    st->round = round;
    st->bepoch = bepoch;
    st->gr = gr;
    st->hv = hv;
    st->prev_list = prev_list;
    st->dn = dn;
  }
}

// ---- chain stretch: whole outer rounds by block 0 alone ----------------------------------------------------------------
// Most outer rounds of a long dependency chain do almost nothing: one special constraint fires (6 wires), three Jacobi
// rounds of <= 6 records settle it, the linear-system sweep looks at the few hundred rows it still has open and resolves
// nothing, no IsZero pair fires (ecdsa + secp256k1: 24 of 28 outer rounds).  On the whole grid such a round is five grid
// barriers and five phases of a handful of dependent L2 trips each, 147 SMs waiting for one.  Once the linear-system
// sweep has few rows left (engine knob "chain_open_max"), block 0 therefore runs WHOLE outer rounds — Jacobi rounds as
// before (warp_solo / sparse_round), then P2, P4 and the next round's P0 — behind block barriers, and the grid meets
// again when a round needs it (a frontier beyond SOLO_MAX, a heavy wire) or the fixpoint is reached.  Same operations
// on the same buffers in the same order as the grid phases of k_solve; only who executes them differs.

// block barrier that also makes the block's atomics (performed at the L2) visible to the block's cached loads: one
// release fence + one acquire load by the leader (it drops the SM's L1 lines), as between two block-solo rounds.
// Returns *src as of the barrier.
__device__ __forceinline__ unsigned int block_sync_load(const unsigned int* src) {
  __shared__ unsigned int s_bsl;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int v;
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
    s_bsl = v;
  }
  __syncthreads();
  return s_bsl;
}

// the open rows of the linear-system sweep as a list (block 0; the bitmap stays authoritative: closing a row clears
// its bit as well).  Returns the list length, or 0xffffffff when more than `cap` rows are open.
__device__ __noinline__ unsigned int chain_build_list(const Dev&, unsigned int cap) {
  const Dev& d = c_dev;
  __shared__ unsigned int s_n;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const uint32_t n_words = (d.N + 31u) / 32u;
  for (uint32_t base = 0; base < n_words; base += 8u * blockDim.x) {
    uint32_t w[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t i = base + (uint32_t)u * blockDim.x + threadIdx.x;
      w[u] = i < n_words ? __ldcg(d.p2_open + i) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (!w[u]) continue;
      const uint32_t i = base + (uint32_t)u * blockDim.x + threadIdx.x;
      unsigned int at = atomicAdd(&s_n, (unsigned int)__popc(w[u]));
      for (uint32_t m = w[u]; m; m &= m - 1) {
        if (at < cap) d.p2_list[at] = i * 32u + (uint32_t)(__ffs((int)m) - 1);
        ++at;
      }
    }
  }
  __syncthreads();
  const unsigned int n = s_n;
  __syncthreads();
  return n <= cap ? n : 0xffffffffu;
}

// P2 .. P0 of one outer round by block 0 (the state is at the P1 fixpoint in both buffers; `pl`: the phase list this
// round writes).  Returns the number of records in `pl` afterwards (the frontier of the next round's first Jacobi round).
__device__ __noinline__ unsigned int chain_phases(const Dev&, int pl, SpecialsCache& spc, unsigned int list_n,
                                                  unsigned int gr, unsigned int last_dense_gr, unsigned long long* evals_io,
                                                  unsigned int* n_cand_io) {
  const Dev& d = c_dev;
  const uint32_t t = threadIdx.x, nt = blockDim.x, lane = t & 31u, warp = t >> 5, nwarps = nt >> 5;
  __shared__ unsigned int s_nlq;
  uint32_t* lq = d.p2_list + CH_LIST_CAP;  // long rows whose cached scan is stale: one warp each
  const uint8_t* F = d.F[0];
  unsigned long long ev = 0;
  if (t == 0) s_nlq = 0;
  __syncthreads();
  // ---- P2 candidate scan (:1357-1385): the open short rows, a thread each
  for (uint32_t i = t; i < list_n; i += nt) {
    const uint32_t row = __ldcg(d.p2_list + i);
    if (row == 0xffffffffu) continue;
    InlineRow rr;
    load_row(d, row, rr);
    uint32_t ff[ROWREC_INLINE];
#pragma unroll
    for (int j = 0; j < ROWREC_INLINE; ++j) ff[j] = (rr.meta & 0x10000u) ? ld_flag(F, rr.c[j]) : (uint32_t)WF_U;
    ev += 1;  // one visit of the sweep per open row (:1359)
    if (p2_scan_short(d, pl, row, rr, ff)) {
      d.p2_list[i] = 0xffffffffu;
      atomicAnd(d.p2_open + (row >> 5), ~(1u << (row & 31u)));
    }
  }
  // ... the long rows: a thread each decides from the cached scan; stale ones are queued for a warp
  for (uint32_t i = t; i < d.n_long; i += nt) {
    const uint32_t row = d.long_rows[i];
    if (d.solved[row] & 1) continue;
    if (d.long_done[i]) continue;
    const LongP2 c = d.long_p2[i];
    const unsigned int stamp = __ldcg(d.long_stamp + i);
    if (c.gr != 0 && stamp < c.gr && last_dense_gr < c.gr) {
      if (!c.bad && c.k == 1)
        emit(d, 1, pl, c.w1, WF_U | WF_K);
      else if (!c.bad && c.k >= 2)
        p2_candidate(d, row, c.hs, c.hx, c.k);
    } else {
      lq[atomicAdd(&s_nlq, 1u)] = i;  // (n_long <= CH_LONG_MAX: the queue cannot overflow)
    }
  }
  __syncthreads();
  {
    const unsigned int nlq = s_nlq;
    for (unsigned int x = warp; x < nlq; x += nwarps) {
      const uint32_t i = lq[x];
      p2_scan_row<32>(d, 0, pl, d.long_rows[i], d.long_p2 + i, __ldcg(d.long_stamp + i), last_dense_gr, gr);
    }
  }
  unsigned int n_cand = block_sync_load(&d.st->p2_cand);
  if (n_cand > d.N) n_cand = d.N;
  *n_cand_io = n_cand;
  // ---- P2 resolve (:1386-1417): only when some slot holds as many members as its sets have unknowns
  unsigned int n_x;
  if (__ldcg(&d.st->p2_full)) {  // (uniform: read behind the barrier)
    for (uint32_t c = t; c < n_cand; c += nt) p2_resolve_group(d, pl, c);
    n_x = block_sync_load(d.rec_count + pl);
    if (p2_big_phase(d, pl, 0, 1)) n_x = block_sync_load(d.rec_count + pl);
  } else {
    n_x = __ldcg(d.rec_count + pl);  // the k = 1 rows the scan decided on the spot
  }
  {  // replay the P2 updates into buffer 0, clear the table (P3 can only tag in the first outer round: not here)
    const unsigned int nx = n_x > d.rec_cap ? d.rec_cap : n_x;
    for (uint32_t i = t; i < nx; i += nt) {
      const Rec r = d.recs[pl][i];
      apply_update(d, 0, r.wire, r.bits, r.lbr, r.ubr);
    }
    for (uint32_t c = t; c < n_cand; c += nt) {
      const uint32_t slot = d.p2_slot[c];
      d.h_key[slot] = 0ULL;
      d.h_cnt[slot] = 0;
      d.h_head[slot] = 0;
    }
    if (t == 0) {
      d.st->p2_cand = 0;
      d.st->p2_full = 0;
      d.st->p2_big_n = 0;
    }
  }
  if (n_x > 0) block_sync_load(d.rec_count + pl);
  // ---- P4 (:1492-1550): reads buffer 0, U|K to buffer 1.  A pair that is latched, or whose vk is unique, never fires
  // again (both monotone): once a pass has found every pair so, the phase — and its barrier — is skipped for good.
  unsigned int n_y = n_x;
  if (!spc.p4_dead) {
    bool fired = false, open = false;
    for (uint32_t i = t; i < d.n_p4; i += nt) {
      const uint32_t row = d.p4_rows[i];
      if (d.solved[row] & 1) continue;
      if (ld_flag(F, d.aux[row].w4) & WF_U) continue;
      open = true;
      fired |= p4_row(d, pl, row);
    }
    if (fired) atomicAdd(&d.st->p4_fired, 1u);
    const int any_open = __syncthreads_or(open ? 1 : 0);
    if (!any_open) {
      if (t == 0) spc.p4_dead = 1;
    } else {
      n_y = block_sync_load(d.rec_count + pl);
    }
  }
  {  // replay the P4 updates into buffer 0, then the next outer round's P0 (buffer 1 is complete)
    const unsigned int nx = n_x > d.rec_cap ? d.rec_cap : n_x, ny = n_y > d.rec_cap ? d.rec_cap : n_y;
    for (uint32_t i = nx + t; i < ny; i += nt) {
      const Rec r = d.recs[pl][i];
      apply_update(d, 0, r.wire, r.bits, r.lbr, r.ubr);
    }
    if (t == 0) atomicAdd(&d.st->prog, n_y);
    __syncthreads();  // prog += n_y is ordered before P0's atomics on it
    phase_p0(d, pl, spc);
  }
  (void)lane;
  *evals_io += ev;
  return block_sync_load(d.rec_count + pl);
}

// A heavy wire (more than HEAVY_DEG rows) changed: is the next round better off sweeping densely?  Only when the rows
// listed next to the changed heavy wires are a sizeable part of all rows — ecdsa's rounds 6 and 7 follow ~50 changed
// selector wires of 1025 rows each, 51 k row evaluations against a sweep of 400 k live rows.  Every block computes the
// same sum from the same record list (<= HEAVY_SCAN_MAX records; longer lists sweep densely as before).
#define HEAVY_SCAN_MAX 4096u
__device__ __noinline__ bool heavy_wants_dense(const Dev&, unsigned int list, unsigned int n) {
  const Dev& d = c_dev;
  if (n > HEAVY_SCAN_MAX) return true;
  __shared__ unsigned int s_hsum;
  if (threadIdx.x == 0) s_hsum = 0;
  __syncthreads();
  unsigned int sum = 0;
  for (unsigned int i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t w = ld_peer_rec(d.recs[list] + i).wire;
    const uint32_t dg = __ldcg(reinterpret_cast<const uint4*>(d.inv_head) + w).x;
    if (dg > HEAVY_DEG) sum += dg;
  }
  sum = __reduce_add_sync(0xffffffffu, sum);
  if ((threadIdx.x & 31u) == 0 && sum) atomicAdd(&s_hsum, sum);
  __syncthreads();
  const unsigned int total = s_hsum;
  __syncthreads();
  return total > d.N / 8u;
}

// The whole fixpoint (:706-1556) as ONE persistent cooperative launch (148 blocks x 1024 threads):
//
//   P0 -> [Jacobi rounds of the single-row rules until no record] -> P2 -> P3 -> P4 -> repeat while
//   anything changed, every arrow a grid barrier whose release carries the count the next step needs.
//
// Jacobi rounds come in two kinds.  A DENSE round sweeps every row that can still fire: static
// row -> thread mapping, the thread's first <= 6 row records resident in shared memory for the whole
// solve, a 64-bit register mask of its live rows, long rows (> 6 terms) one warp each.  A SPARSE round
// is driven by the previous round's update records: a row none of whose wires changed evaluates to
// exactly what it evaluated to before, so only the rows the wire -> rows index lists next to a changed
// wire are evaluated (one thread per record).  Both read buffer R, write buffer W and log records;
// the records are replayed into the other buffer during the next round (one barrier per round).
__global__ void __launch_bounds__(P1_THREADS, P1_MIN_BLOCKS)
    k_solve(unsigned int max_rounds, int ks) {
  const Dev& d = c_dev;
  extern __shared__ uint4 sm_rec[];  // sm_rec[(2*k + h) * blockDim + thread]: half h of the thread's k-th record
  __shared__ unsigned int s_solo[3];  // records, round flags, distinct wires of a solo round
  __shared__ SoloState s_ws;
  unsigned int epoch = 0;
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nthreads = gridDim.x * blockDim.x;
  const uint32_t rows = d.row_hi - d.row_lo;
  const uint32_t per_thread = (rows + nthreads - 1) / nthreads;
  const uint32_t kmask = per_thread < LiveMask::BITS ? per_thread : LiveMask::BITS;
  const uint32_t warp_in_block = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  unsigned int xe = d.shard ? *d.xepoch : 0u;  // cross-GPU epoch (same on every rank, persists over solves)
  unsigned int ack_pending = 0;                    // epoch of a sharded round whose lists peers may still be reading
  unsigned long long evals = 0, ruleevals = 0, devals = 0, dcycles = 0;
  unsigned int rounds_total = 0, dense_rounds = 0;
  unsigned int gr = 0;  // Jacobi round counter of the solve (stamps the long-row queue)
  unsigned int last_dense_gr = 0;
  uint32_t bepoch = 0;  // rounds so far that tightened a bound (same value in every thread)

  // ---- stage: live mask + the smem-resident row records ---------------------------------------
  LiveMask live;
  live.reset();
  for (uint32_t k = 0; k < kmask; ++k)
    if (tid + k * nthreads < rows) live.set((int)k);
#pragma unroll
  for (int k = 0; k < P1_MAX_KS; ++k) {
    if (k < ks && live.test(k)) {
      const uint32_t r = tid + (uint32_t)k * nthreads;
      const uint4* rp = reinterpret_cast<const uint4*>(d.rec + d.row_lo + r);
      sm_rec[(2 * k) * blockDim.x + threadIdx.x] = __ldg(rp);
      sm_rec[(2 * k + 1) * blockDim.x + threadIdx.x] = __ldg(rp + 1);
    }
  }
  __syncthreads();
  // ---- prologue: P0 of the first outer round ------------------------------------------------------
  __shared__ SpecialsCache s_spc;
  if (blockIdx.x == 0) {
    specials_cache_fill(d, s_spc);
    phase_p0(d, PL0, s_spc);
  }
  unsigned int n_pl = sync_and_load(d, d.rec_count + PL0);
  unsigned int prog_prev = 0, outer = 0, p4_seen = 0;
  bool stop = false;
  unsigned long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned int cand_total = 0, cand_max = 0;
  long long tp = clock64();
#define PROF(i)                               \
  do {                                        \
    long long t_ = clock64();                 \
    pf[i] += (unsigned long long)(t_ - tp);   \
    tp = t_;                                  \
  } while (0)

  // rows latched by P4 leave the live masks (equation_solved, :820-822)
  auto refresh_live = [&]() {
    const unsigned int p4f = __ldcg(&d.st->p4_fired);
    if (p4f != p4_seen) {
      p4_seen = p4f;
      for (LiveMask m = live; m.any();) {
        const int k = m.pop_nonempty();
        if (d.solved[d.row_lo + tid + (uint32_t)k * nthreads] & 1) live.clear(k);
      }
    }
  };
  while (!stop) {
    ++outer;
    int pl_r = PL0 + (int)((outer - 1) & 1u);  // phase list read by this round's first Jacobi round
    int pl = PL0 + (int)(outer & 1u);          // ... written by this round's phases
    bool chain_done = false;  // block 0 ran this round's phases (and possibly whole rounds before it) alone
    // =============================== P1: Jacobi rounds to a fixpoint ===============================
    if (outer == 1 || n_pl > 0) {
      int rbuf = 0;
      unsigned int list = 0, prev_list = (unsigned int)pl_r, prev_n = n_pl > d.rec_cap ? d.rec_cap : n_pl;
      // `dn`: the number of distinct wires the previous round changed.  Mode decisions (dense / sparse /
      // solo) use it instead of the record count on sharded runs, because it is the same on every rank
      unsigned int prev_dn = d.shard ? __ldcg(d.dcnt + pl_r) : n_pl;
      unsigned int prev_own = prev_n;  // leading records of the list that still have to be replayed into the write
                                       // buffer (all of them, except after a sharded round: the phase list is complete)
      bool dense = outer == 1 || prev_dn > d.sparse_max;
      // a changed heavy wire whose rows are not worth a dense sweep: the round is frontier-driven on the WHOLE grid (one
      // block, let alone one warp, would stride tens of thousands of listed rows alone)
      bool heavy_grid = false;
      if (!dense && (__ldcg(d.bnd_flag + pl_r) & 2u) != 0) {
        dense = d.shard || heavy_wants_dense(d, prev_list, prev_n);
        heavy_grid = !dense;
      }
      unsigned int round = 0;
      while (true) {
        if (ack_pending) {  // the peers may still be reading the record lists of our last sharded round
          if (threadIdx.x == 0) cross_gpu_wait_acks(d, ack_pending);
          __syncthreads();
          ack_pending = 0;
        }
        if (!dense && !heavy_grid && prev_dn <= SOLO_MAX) {
          // ---- solo: while the frontier stays small, block 0 runs the Jacobi rounds alone; a round
          // boundary is a block barrier + one release fence + one acquire load (which also drops this
          // SM's L1 lines, as the grid barrier does) instead of a grid barrier
          unsigned int n = 0;
          if (blockIdx.x == 0) {
            unsigned int ch_pos = CH_NONE, ch_list_n = 0xffffffffu, ch_n_pl = 0, ch_stop = 0;
            while (true) {  // chain stretch: rounds, and — while the conditions hold — whole outer rounds
            while (true) {
              if (prev_n <= WARP_SOLO_MAX && prev_own == prev_n) {
                // at most 32 records: warp 0 chases them alone, for as many rounds as that stays so
                if (threadIdx.x == 0) {
                  s_ws.list = list;
                  s_ws.prev_list = prev_list;
                  s_ws.rbuf = (unsigned int)rbuf;
                  s_ws.round = round;
                  s_ws.bepoch = bepoch;
                  s_ws.gr = gr;
                }
                __syncthreads();
                if (warp_in_block == 0) {
                  unsigned long long ev = 0;
                  if (d.shard)
                    warp_solo<false>(d, &s_ws, prev_n, max_rounds, &ev);
                  else
                    warp_solo<true>(d, &s_ws, prev_n, max_rounds, &ev);
                  if (d.rank == 0) {  // replicated work is counted once
                    evals += ev;
                    ruleevals += ev;
                  }
                }
                __syncthreads();
                n = s_ws.n;
                list = s_ws.list;
                rbuf = (int)s_ws.rbuf;
                round = s_ws.round;
                bepoch = s_ws.bepoch;
                gr = s_ws.gr;
                s_solo[1] = s_ws.hv;
                s_solo[2] = s_ws.dn;
#ifdef ECNE_PROFILE
                if (threadIdx.x == 0) {
                  long long t_ = clock64();
                  pf[7] += (unsigned long long)(t_ - tp);
                  tp = t_;
                }
#endif
                __syncthreads();
                if (n == 0 || s_solo[2] > SOLO_MAX || (s_solo[1] & 2u) || round >= max_rounds) break;
                prev_list = list;
                prev_n = n;
                prev_own = n;
                list = (list + 1) % 3;
                rbuf ^= 1;
                continue;
              }
              gr += 1;
#ifdef ECNE_PROFILE
              long long z0 = clock64(), z1 = 0, z2 = 0, z3 = 0, z4 = 0;
#endif
              const unsigned long long ev = sparse_round(d, rbuf, list, prev_list, prev_n, bepoch, gr, true, prev_own);
              if (d.rank == 0) {  // replicated work is counted once
                evals += ev;
                ruleevals += ev;
              }
#ifdef ECNE_PROFILE
              z1 = clock64();
#endif
              __syncthreads();
              if (threadIdx.x == 0) {
                unsigned int cnt, bf;
#ifdef ECNE_PROFILE
                z2 = clock64();
#endif
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
#ifdef ECNE_PROFILE
                z3 = clock64();
#endif
                asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(cnt) : "l"(d.rec_count + list) : "memory");
                asm volatile("ld.relaxed.gpu.u32 %0, [%1];" : "=r"(bf) : "l"(d.bnd_flag + list) : "memory");
                unsigned int dnv = cnt;
                if (d.shard) asm volatile("ld.relaxed.gpu.u32 %0, [%1];" : "=r"(dnv) : "l"(d.dcnt + list) : "memory");
                s_solo[0] = cnt;
                s_solo[1] = bf;
                s_solo[2] = dnv;
#ifdef ECNE_PROFILE
                z4 = clock64();
                if (gr < 1000) {
                  d.prof[24000 + 4 * gr + 0] = (unsigned long long)(z1 - z0);  // sparse_round
                  d.prof[24000 + 4 * gr + 1] = (unsigned long long)(z2 - z1);  // block barrier
                  d.prof[24000 + 4 * gr + 2] = (unsigned long long)(z3 - z2);  // fence
                  d.prof[24000 + 4 * gr + 3] = (unsigned long long)(z4 - z3);  // acquire + flag loads
                }
#endif
                d.rec_count[prev_list] = 0;  // consumed; next written two rounds from now
                d.bnd_flag[prev_list] = 0;
                d.dcnt[prev_list] = 0;
                atomicAdd(&d.st->prog, cnt);
              }
              __syncthreads();
              n = s_solo[0];
              bepoch += s_solo[1] & 1u;
              round += 1;
#ifdef ECNE_PROFILE
              if (threadIdx.x == 0 && gr < 4000) {
                long long t_ = clock64();
                d.prof[4 * gr + 0] = (unsigned long long)(t_ - tp);
                d.prof[4 * gr + 1] = n;
                d.prof[4 * gr + 2] = 2;
                d.prof[4 * gr + 3] = outer;
                pf[7] += (unsigned long long)(t_ - tp);
                tp = t_;
              }
#endif
              if (n == 0 || s_solo[2] > SOLO_MAX || (s_solo[1] & 2u) || round >= max_rounds) break;
              prev_list = list;
              prev_n = n;
              prev_own = n;  // written by this (replicated) round: every record is this rank's own
              list = (list + 1) % 3;
              rbuf ^= 1;
            }
            // ---- the P1 fixpoint of outer round `outer` was reached by this block alone: its phases too? ----
            if (n != 0 || d.shard || outer < 2 || d.chain_open_max <= 0 || d.n_p4 > CH_P4_MAX || d.n_long > CH_LONG_MAX) break;
            if (ch_list_n == 0xffffffffu) {
              // rows the sweep of the round before left open (counted by the grid scan; a stretch keeps its own list)
              if (__ldcg(&d.st->p2_open_n[(outer - 1u) & 1u]) > (unsigned int)d.chain_open_max) break;
              ch_list_n = chain_build_list(d, CH_LIST_CAP);
              if (ch_list_n == 0xffffffffu) break;
            }
            rounds_total += round;
            {
              unsigned long long ev = 0;
              unsigned int nc = 0;
              ch_n_pl = chain_phases(d, pl, s_spc, ch_list_n, gr, last_dense_gr, &ev, &nc);
              if (d.rank == 0) evals += ev;
              cand_total += nc;
              cand_max = nc > cand_max ? nc : cand_max;
            }
            ch_pos = CH_DONE;
            round = 0;
            {
              const unsigned int prog = __ldcg(&d.st->prog);
              const unsigned int err = __ldcg(&d.st->err) | __ldcg(&d.st->rec_overflow);
              if (err || prog == prog_prev) ch_stop = 1;  // successful_steps did not move (:708-711)
              prog_prev = prog;
              if (!ch_stop && outer >= d.max_outer) {
                if (threadIdx.x == 0) raise(d, ECNE_E_NOCONVERGE);
                ch_stop = 1;
              }
            }
#ifdef ECNE_PROFILE
            if (threadIdx.x == 0) {
              long long t_ = clock64();
              pf[6] += (unsigned long long)(t_ - tp);  // (chain phases are booked under "p0+replay")
              tp = t_;
            }
#endif
            // the next outer round starts here when its first Jacobi round can be run by this block as well
            if (ch_stop || ch_n_pl == 0 || ch_n_pl > SOLO_MAX || ch_n_pl > d.rec_cap || (__ldcg(d.bnd_flag + pl) & 2u)) break;
            ++outer;
            pl_r = PL0 + (int)((outer - 1) & 1u);
            pl = PL0 + (int)(outer & 1u);
            ch_pos = CH_MIDP1;
            rbuf = 0;
            list = 0;
            prev_list = (unsigned int)pl_r;
            prev_n = ch_n_pl;
            prev_own = prev_n;
            n = prev_n;
            s_solo[1] = 0;
            s_solo[2] = prev_n;
            __syncthreads();
            }
            if (threadIdx.x == 0) {  // where the other blocks pick the loop up again
              d.st->solo[0] = n;
              d.st->solo[1] = list;
              d.st->solo[2] = (unsigned int)rbuf;
              d.st->solo[3] = round;
              d.st->solo[4] = bepoch;
              d.st->solo[5] = gr;
              d.st->solo[6] = s_solo[1] & 2u;
              d.st->solo[7] = s_solo[2];
              d.st->chain[0] = ch_pos;
              d.st->chain[1] = outer;
              d.st->chain[2] = prog_prev;
              d.st->chain[3] = ch_n_pl;
              d.st->chain[4] = ch_stop;
              if (ch_pos != CH_NONE) {
                // the grid's next entry test reads the count of the last completed round; its next scan adds into the
                // other slot (a stretch's own scans do not count)
                const unsigned int done = ch_pos == CH_DONE ? outer : outer - 1u;
                d.st->p2_open_n[done & 1u] = ch_list_n;
                d.st->p2_open_n[(done + 1u) & 1u] = 0;
              }
            }
          }
          grid_sync_flip(d.barrier + 64);
          n = __ldcg(&d.st->solo[0]);
          list = __ldcg(&d.st->solo[1]);
          rbuf = (int)__ldcg(&d.st->solo[2]);
          round = __ldcg(&d.st->solo[3]);
          bepoch = __ldcg(&d.st->solo[4]);
          gr = __ldcg(&d.st->solo[5]);
          const unsigned int hv = __ldcg(&d.st->solo[6]);
          const unsigned int sdn = __ldcg(&d.st->solo[7]);
          const unsigned int ch = __ldcg(&d.st->chain[0]);
          if (ch != CH_NONE) {  // block 0 went on into later outer rounds
            outer = __ldcg(&d.st->chain[1]);
            prog_prev = __ldcg(&d.st->chain[2]);
            pl_r = PL0 + (int)((outer - 1) & 1u);
            pl = PL0 + (int)(outer & 1u);
            if (ch == CH_DONE) {  // ... and finished the phases of round `outer`
              n_pl = __ldcg(&d.st->chain[3]);
              stop = __ldcg(&d.st->chain[4]) != 0;
              chain_done = true;
              round = 0;
              PROF(7);
              break;
            }
            refresh_live();  // the phases of the rounds in between may have latched IsZero pairs
          }
          PROF(7);
          if (n == 0) break;
          if (round >= max_rounds) {
            if (tid == 0) raise(d, ECNE_E_NOCONVERGE);
            break;
          }
          prev_list = list;
          prev_n = n > d.rec_cap ? d.rec_cap : n;
          list = (list + 1) % 3;
          rbuf ^= 1;
          prev_dn = sdn;
          prev_own = prev_n;
          dense = sdn > d.sparse_max || (hv != 0 && (d.shard || heavy_wants_dense(d, prev_list, prev_n)));
          heavy_grid = hv != 0 && !dense;
          continue;
        }
        const int wbuf = rbuf ^ 1;
        const uint8_t* F = d.F[rbuf];
        const bool sharded = d.shard && dense;  // only dense sweeps are split over the ranks
        const int elist = sharded ? (int)(list | LIST_NOCOUNT) : (int)list;  // (sweep.cuh: no distinct-wire counting)
        gr += 1;
        long long tc0 = 0;
        if (dense && tid == 0) tc0 = clock64();
        if (dense) {
#ifdef ECNE_PROFILE
          long long dz1 = clock64();
#endif
          // (b) replay the previous round's own records into the buffer written this round
          if (prev_own) {
            const Rec* pr = d.recs[prev_list];
            // (consecutive threads take consecutive 16-byte records: a round of S16 replays 5 M of them)
            for (uint32_t j = tid; j < prev_own; j += nthreads) {
              Rec r = pr[j];
              apply_update(d, wbuf, r.wire, r.bits, r.lbr, r.ubr);
              consume_rec(d, prev_list, r.wire);
            }
          }
#ifdef ECNE_PROFILE
          long long dz2 = clock64();
#endif
          // (c) sweep the rows this thread still owns, P1_INFLIGHT of them in flight
          const unsigned int nl = (sharded || d.rank == 0) ? live.count() : 0u;  // replicated work is counted once
          evals += nl;
          devals += nl;
          ruleevals += nl;
          LiveMask slow;
          slow.reset();
          for (LiveMask m = live; m.any();) {
            // P1_INFLIGHT rows of this thread in flight: all their records, then all their state gathers, then
            // the evaluations (an empty slot repeats the first row's loads and is not evaluated)
            int kk[P1_INFLIGHT];
            InlineRow rr[P1_INFLIGHT];
            uint32_t ff[P1_INFLIGHT][ROWREC_INLINE];
#pragma unroll
            for (int h = 0; h < P1_INFLIGHT; ++h) kk[h] = m.pop();
#pragma unroll
            for (int h = 0; h < P1_INFLIGHT; ++h) {
              const int k = kk[h] >= 0 ? kk[h] : kk[0];
              if (k < ks)
                unpack_row(sm_rec[(2 * k) * blockDim.x + threadIdx.x], sm_rec[(2 * k + 1) * blockDim.x + threadIdx.x], rr[h]);
              else
                load_row(d, d.row_lo + tid + (uint32_t)k * nthreads, rr[h]);
            }
#pragma unroll
            for (int h = 0; h < P1_INFLIGHT; ++h)
              if (kk[h] >= 0) gather_row(F, rr[h], ff[h]);
#pragma unroll
            for (int h = 0; h < P1_INFLIGHT; ++h)
              if (kk[h] >= 0) {
                Pending pend;
                const uint32_t e = eval_inline(d, rbuf, d.row_lo + tid + (uint32_t)kk[h] * nthreads, rr[h], ff[h], bepoch, pend);
                // (one emit per update here: measured against the paired emit2_impl — which wins where ONE warp waits for its
                // atomics, the solo and frontier-driven rounds — the plain routine is 1-2 % faster under a full grid)
                if (pend.has0()) emit(d, wbuf, elist, pend.w0, pend.b0, pend.l0, pend.u0);
                if (pend.has1()) emit(d, wbuf, elist, pend.w1, pend.b1, pend.l1, pend.u1);
                if (e & EI_DONE) live.clear(kk[h]);
                if (e & EI_GENERIC) slow.set(kk[h]);
              }
          }
#ifdef ECNE_PROFILE
          long long dzf = clock64();
          if (dense_rounds < 40 && slow.any())
            atomicAdd(d.prof + 28000 + 40 * 148 * 4 + 8 + dense_rounds, (unsigned long long)slow.count());
#endif
          // the rows that need the generic evaluator (bit-decomposition patterns, x + y = 1, Case 5/6
          // candidates), all lanes together: inside the loop above one such lane would stall its warp in
          // almost every iteration
          for (LiveMask m = slow; m.any();) {
            const int k = m.pop_nonempty();
            const uint32_t row = d.row_lo + tid + (uint32_t)k * nthreads;
            const bool done = eval_row<1>(d, rbuf, wbuf, elist, row, bepoch);
            bool fast;
            if (k < ks)
              fast = (sm_rec[(2 * k) * blockDim.x + threadIdx.x].x & RF_FAST) != 0;
            else
              fast = (d.rflags[row] & RF_FAST) != 0;
            if (done && !fast) live.clear(k);  // fast-path rows stay: bounds may still travel through them
          }
          // rows beyond the ones tracked per thread (only for problems far larger than the machine)
          for (uint32_t k = kmask; k < per_thread; ++k) {
            uint32_t r = tid + k * nthreads;
            if (r < rows) {
              uint32_t row = d.row_lo + r;
              if (!(d.rflags[row] & RF_LONG) && !(d.solved[row] & 1)) eval_row<1>(d, rbuf, wbuf, elist, row, bepoch);
              evals += 1;
              devals += 1;
              ruleevals += 1;
            }
          }
#ifdef ECNE_PROFILE
          long long dz4 = clock64();
#endif
          // (a) long rows last, one warp each: block b owns long rows b, b+grid, ...  (not first: in a round that
          // changes 300 k wires their ~30 dependent loads would crawl through the sweep's atomic traffic)
          for (uint32_t i = blockIdx.x + warp_in_block * gridDim.x; i < d.n_long; i += warps_per_block * gridDim.x) {
            if (d.long_done[i]) continue;
            const uint32_t row = d.long_rows[i];
            if (row >= d.row_lo && row < d.row_hi) {
              const bool done = eval_row<32>(d, rbuf, wbuf, elist, row, bepoch);
              if (lane == 0) {
                if (sharded || d.rank == 0) {
                  evals += 1;
                  devals += 1;
                  ruleevals += 1;
                }
                if (done) d.long_done[i] = 1;
              }
            }
          }
#ifdef ECNE_PROFILE
          if (threadIdx.x == 0 && dense_rounds < 40) {
            unsigned long long* q = d.prof + 28000 + ((size_t)dense_rounds * gridDim.x + blockIdx.x) * 4;
            long long dz3 = clock64();
            q[0] = (unsigned long long)(dz3 - dz4);  // long rows
            q[1] = (unsigned long long)(dz2 - dz1);  // replay
            q[2] = (unsigned long long)(dz4 - dz2);  // sweep
            q[3] = gr | ((unsigned long long)(dzf - dz2) << 20);  // ... of which the inline loop
          }
#endif
        } else {
          // frontier-driven rounds are never split over the ranks: every rank evaluates every listed row (a cross-GPU
          // exchange costs more than such a round), so its record list is complete locally afterwards
          const unsigned long long ev = sparse_round(d, rbuf, list, prev_list, prev_n, bepoch, gr, false, prev_own);
          if (d.rank == 0) {  // replicated work is counted once
            evals += ev;
            ruleevals += ev;
          }
        }
        unsigned int n;
        bool heavy = false;  // a wire with very many rows changed: sweep densely instead of chasing its list
        unsigned int dn;
#ifdef ECNE_PROFILE
        long long xz0 = clock64(), xz1 = 0, xz2 = 0;
#endif
        if (sharded) {
          xe += 1;
          n = grid_barrier(d.barrier, epoch, d.rec_count + list, d.bnd_flag + list, &d, list, xe);
          bepoch += n >> 31;
          n &= 0x7fffffffu;
          dn = n;  // the ranks' record counts, summed: the same number on every rank (it was exchanged)
          heavy = __ldcg(d.xcnt + 3 * ECNE_MAX_WORLD + 1) != 0;
        } else {
          grid_sync_flip(d.barrier + 64);
          n = __ldcg(d.rec_count + list);
          const unsigned int bf = __ldcg(d.bnd_flag + list);
          bepoch += bf & 1u;
          heavy = (bf & 2u) != 0;
          dn = d.shard ? __ldcg(d.dcnt + list) : n;
        }
        unsigned int n_own = n;
#ifdef ECNE_PROFILE
        xz1 = clock64();
#endif
        if (sharded) {
          // pull the peers' records of this round over NVLink, apply them to BOTH local buffers (the buffer read
          // next round must already contain them) and append them behind this rank's own records: the local list is
          // complete afterwards and nobody reads a peer's list again once the round's acknowledgement is out.  A
          // peer's record that changes nothing in the buffer this rank wrote during the round repeats one of its own
          // records (two rows on different ranks fixed the same wire) and is not appended: "first writer logs" across
          // the ranks, so the next round's frontier is what one GPU would have.
          n_own = __ldcg(d.xcnt + list * ECNE_MAX_WORLD + d.rank);
          if (n_own > d.rec_cap) n_own = d.rec_cap;
          for (int h = 0; h < d.world; ++h) {
            if (h == d.rank) continue;
            const unsigned int nh = __ldcg(d.xcnt + list * ECNE_MAX_WORLD + h);
            const Rec* pr = d.xrecs[h][list];
            for (uint32_t j = tid; j < nh && j < d.rec_cap; j += nthreads) {
              Rec r = ld_peer_rec(pr + j);
              const bool fresh = (apply_update_t<true>(wbuf, r.wire, r.bits, r.lbr, r.ubr) & 1u) != 0;
              apply_update(d, rbuf, r.wire, r.bits, r.lbr, r.ubr);
              if (fresh) {  // warp-aggregated slot allocation (a big round appends > 10^5 records: one RMW per warp)
                const unsigned int am = __activemask();
                const int leader = __ffs((int)am) - 1;
                unsigned int i = 0;
                if ((int)lane == leader) i = atomicAdd(d.rec_count + list, (unsigned int)__popc(am));
                i = __shfl_sync(am, i, leader) + (unsigned int)__popc(am & ((1u << lane) - 1u));
                if (i < d.rec_cap)
                  d.recs[list][i] = r;
                else
                  d.st->rec_overflow = 1;
              }
            }
          }
#ifdef ECNE_PROFILE
          xz2 = clock64();
#endif
          grid_sync_flip(d.barrier + 64);
          // tell the peers that their lists of this round have been read here; our own lists of this round must
          // not be overwritten before every peer has said the same (checked when the next round starts)
          if (tid == 0) cross_gpu_ack(d, xe);
          ack_pending = xe;
          n = __ldcg(d.rec_count + list);  // own records + the peers' that were new here
        }
#ifdef ECNE_PROFILE
        if (tid == 0 && dense && dense_rounds < 40) {
          unsigned long long* q = d.prof + 27000 + 8 * dense_rounds;
          q[0] = (unsigned long long)(xz1 - xz0);                       // round barrier (+ cross-GPU exchange)
          q[1] = sharded ? (unsigned long long)(xz2 - xz1) : 0;         // pull of the peers' records
          q[2] = sharded ? (unsigned long long)(clock64() - xz2) : 0;   // barrier after the pull + ack
          q[3] = n;
          q[4] = (unsigned long long)(xz0 - tc0);                        // replay + sweep + long rows (block 0)
        }
#endif
        if (tid == 0 && d.prof && gr < 4000) {
          d.prof[4 * gr + 0] = (unsigned long long)(clock64() - tp);
          d.prof[4 * gr + 1] = n;
          d.prof[4 * gr + 2] = dense ? 1 : 0;
          d.prof[4 * gr + 3] = outer;
        }
        if (dense) {
          dense_rounds += 1;
          last_dense_gr = gr;
          if (tid == 0) dcycles += (unsigned long long)(clock64() - tc0);
          PROF(0);
        } else {
          PROF(1);
        }
        round += 1;
        if (tid == 0) {  // the list read this round is consumed; it is next written two rounds from now
          d.rec_count[prev_list] = 0;
          d.bnd_flag[prev_list] = 0;
          d.dcnt[prev_list] = 0;
          atomicAdd(&d.st->prog, n);
        }
        if (n_own > d.rec_cap) n_own = d.rec_cap;
        if (n > d.rec_cap) n = d.rec_cap;
        if (n == 0) break;  // W already holds every earlier record: both buffers are complete
        if (round >= max_rounds) {
          if (tid == 0) raise(d, ECNE_E_NOCONVERGE);
          break;
        }
        prev_list = list;
        prev_n = n;        // the whole round's records (after a sharded round: own ones first, then the peers')
        prev_own = n_own;  // ... of which these still have to be replayed into the other buffer
        list = (list + 1) % 3;
        rbuf = wbuf;
        prev_dn = dn;
        dense = dn > d.sparse_max || (heavy && (d.shard || heavy_wants_dense(d, prev_list, prev_n)));
        heavy_grid = heavy && !dense;
      }
      rounds_total += round;
    }
    if (!chain_done) {
    // =============================== P2: linear systems (:1357-1417) ===============================
    // candidate scan over the rows that can still fire (state is at the P1 fixpoint, in both buffers)
#ifdef ECNE_PROFILE
    long long pz0 = 0, pz1 = 0;
#endif
    {
      const uint8_t* F = d.F[0];
#ifdef ECNE_PROFILE
      pz0 = clock64();
#endif
      // One bit per row says whether the sweep still has to look at it: a row all of whose wires are unique (or
      // that is latched as solved) can never qualify again — uniqueness is monotone — and is closed by the scan
      // that finds it so.  A warp takes 32 consecutive rows (one word of the bitmap, two words in flight), its
      // lanes one row each: one coalesced 1 KB read of the open rows' records and their state-byte gathers.  On a
      // sharded run every rank scans every open row (same state everywhere => same candidates, no exchange);
      // the visits are counted once.
      {
        const uint32_t n_words = (d.N + 31u) / 32u;
        const uint32_t gw = blockIdx.x * warps_per_block + warp_in_block, nw = gridDim.x * warps_per_block;
        unsigned int open_left = 0;  // rows this warp leaves open (lane 0)
        for (uint32_t wi = gw; wi < n_words; wi += 2 * nw) {
          uint32_t open[2];
          InlineRow rr[2];
          uint32_t ff[2][ROWREC_INLINE];
          bool mine[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t w = wi + (uint32_t)h * nw;
            open[h] = w < n_words ? d.p2_open[w] : 0u;
            mine[h] = ((open[h] >> lane) & 1u) != 0;
          }
          if ((open[0] | open[1]) == 0u) continue;
#pragma unroll
          for (int h = 0; h < 2; ++h)
            if (mine[h]) load_row(d, (wi + (uint32_t)h * nw) * 32u + lane, rr[h]);
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < ROWREC_INLINE; ++j)
              ff[h][j] = (mine[h] && (rr[h].meta & 0x10000u)) ? ld_flag(F, rr[h].c[j]) : (uint32_t)WF_U;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            bool close = false;
            if (mine[h]) close = p2_scan_short(d, pl, (wi + (uint32_t)h * nw) * 32u + lane, rr[h], ff[h]);
            const uint32_t cm = __ballot_sync(0xffffffffu, close);
            if (lane == 0 && open[h]) {
              open_left += (unsigned int)__popc(open[h] & ~cm);
              if (cm) d.p2_open[wi + (uint32_t)h * nw] = open[h] & ~cm;
              if (d.rank == 0) evals += (unsigned int)__popc(open[h]);  // one visit of the sweep per open row (:1359)
#ifdef ECNE_PROFILE
              if (outer < 40) atomicAdd(d.prof + 28000 + 40 * 148 * 4 + 48 + outer, (unsigned long long)__popc(open[h]));
#endif
            }
          }
        }
        if (lane == 0 && open_left) atomicAdd(&d.st->p2_open_n[outer & 1u], open_left);
      }
#ifdef ECNE_PROFILE
      pz1 = clock64();
#endif
      for (uint32_t i = blockIdx.x + warp_in_block * gridDim.x; i < d.n_long; i += warps_per_block * gridDim.x) {
        const uint32_t row = d.long_rows[i];
        if (d.solved[row] & 1) continue;
        if (d.long_done[i]) continue;  // no non-unique wire left in C (known to the ranks that evaluated the row)
        p2_scan_row<32>(d, 0, pl, row, d.long_p2 + i, __ldcg(d.long_stamp + i), last_dense_gr, gr);
      }
    }
#ifdef ECNE_PROFILE
    if (threadIdx.x == 0 && outer < 12) {
      unsigned long long* q = d.prof + 28000 + ((size_t)(12 + outer) * gridDim.x + blockIdx.x) * 4;
      long long pz2 = clock64();
      q[0] = (unsigned long long)(pz1 - pz0);  // short rows
      q[1] = (unsigned long long)(pz2 - pz1);  // long rows
      q[2] = (unsigned long long)(pz0 - tp);   // wait before the scan (block 0: since the last PROF point)
      q[3] = outer;
    }
#endif
    unsigned int n_cand = sync_and_load(d, &d.st->p2_cand);
    if (n_cand > d.N) n_cand = d.N;
    cand_total += n_cand;
    cand_max = n_cand > cand_max ? n_cand : cand_max;
    PROF(2);
    unsigned int n_x;
    if (__ldcg(&d.st->p2_full)) {  // (uniform: read behind the scan's barrier)
      for (uint32_t c = tid; c < n_cand; c += nthreads) p2_resolve_group(d, pl, c);
      n_x = sync_and_load(d, d.rec_count + pl);
      // groups with more than ECNE_P2_KMAX unknowns were queued by their resolvers: a block each, one more barrier
      if (p2_big_phase(d, pl, blockIdx.x, gridDim.x)) n_x = sync_and_load(d, d.rec_count + pl);
    } else {
      n_x = __ldcg(d.rec_count + pl);  // the k = 1 rows the scan decided on the spot
    }
    PROF(3);
    // replay the P2 updates into buffer 0, clear the table, P3 claim (reads buffer 1: complete, untouched
    // here).  P3 can only ever tag in the first outer round: a wire it looks at is either unique (for
    // good) or was tagged then, so later rounds skip it — and with nothing to replay, the barrier too.
    const bool p3_round = outer == 1;
    {
      const unsigned int nx = n_x > d.rec_cap ? d.rec_cap : n_x;
      for (uint32_t i = tid; i < nx; i += nthreads) {
        Rec r = d.recs[pl][i];
        apply_update(d, 0, r.wire, r.bits, r.lbr, r.ubr);
      }
      for (uint32_t c = tid; c < n_cand; c += nthreads) {
        const uint32_t slot = d.p2_slot[c];
        d.h_key[slot] = 0ULL;
        d.h_cnt[slot] = 0;
        d.h_head[slot] = 0;
      }
      if (tid == 0) {
        d.st->p2_cand = 0;
        d.st->p2_full = 0;
        d.st->p2_big_n = 0;
        d.st->p2_open_n[(outer + 1u) & 1u] = 0;  // the next round's scan counts the rows it leaves open here
      }
      if (p3_round)
        for (uint32_t i = tid; i < d.n_p3; i += nthreads) p3_claim_row(d, d.p3_rows[i]);
    }
    if (p3_round || n_x > 0) sync_and_load(d, nullptr);
    PROF(4);
    // P3 commit (reads buffer 0, writes K|ABZ to both) and P4 (reads buffer 0, U|K to buffer 1)
    {
      if (p3_round)
        for (uint32_t i = tid; i < d.n_p3; i += nthreads) p3_commit_row(d, pl, d.p3_rows[i]);
      bool fired = false;
      for (uint32_t i = tid; i < d.n_p4; i += nthreads) {
#ifdef ECNE_PROFILE
        const uint32_t r_ = d.p4_rows[i];
        if (outer < 40 && !(d.solved[r_] & 1) && !(ld_flag(d.F[0], d.aux[r_].w4) & WF_U))
          atomicAdd(d.prof + 28000 + 40 * 148 * 4 + 88 + outer, 1ULL);
#endif
        fired |= p4_row(d, pl, d.p4_rows[i]);
      }
      if (fired) atomicAdd(&d.st->p4_fired, 1u);
    }
    const unsigned int n_y = sync_and_load(d, d.rec_count + pl);
    PROF(5);
    // replay the P4 updates into buffer 0; block 0 runs the next outer round's P0 (buffer 1 is complete)
    {
      const unsigned int nx = n_x > d.rec_cap ? d.rec_cap : n_x, ny = n_y > d.rec_cap ? d.rec_cap : n_y;
      for (uint32_t i = nx + tid; i < ny; i += nthreads) {
        Rec r = d.recs[pl][i];
        apply_update(d, 0, r.wire, r.bits, r.lbr, r.ubr);
      }
      if (tid == 0) atomicAdd(&d.st->prog, n_y);
      if (blockIdx.x == 0) {
        __syncthreads();  // prog += n_y is ordered before P0's atomics on it
        phase_p0(d, pl, s_spc);
      }
    }
    n_pl = sync_and_load(d, d.rec_count + pl);
    PROF(6);
    const unsigned int prog = __ldcg(&d.st->prog);
    const unsigned int err = __ldcg(&d.st->err) | __ldcg(&d.st->rec_overflow);
    if (err || prog == prog_prev) stop = true;  // successful_steps did not move (:708-711)
    prog_prev = prog;
    if (!stop && outer >= d.max_outer) {
      if (tid == 0) raise(d, ECNE_E_NOCONVERGE);
      stop = true;
    }
    }  // !chain_done
    if (!stop) refresh_live();
  }
  if (ack_pending && threadIdx.x == 0) cross_gpu_wait_acks(d, ack_pending);  // the next solve rewrites the lists
  // ---- statistics: one atomic per warp ---------------------------------------------------------------
  for (int o = 16; o > 0; o >>= 1) {
    evals += __shfl_xor_sync(0xffffffffu, evals, o);
    ruleevals += __shfl_xor_sync(0xffffffffu, ruleevals, o);
    devals += __shfl_xor_sync(0xffffffffu, devals, o);
  }
  if (lane == 0) {
    atomicAdd(&d.st->evals, evals);
    atomicAdd(&d.st->rule_evals, ruleevals);
    atomicAdd(&d.st->dense_evals, devals);
  }
  if (tid == 0) {
    d.st->rounds = rounds_total;
    d.st->outer = outer;
    d.st->dense_rounds = dense_rounds;
    d.st->dense_cycles = dcycles;
    d.st->bepoch = bepoch;
    for (int i = 0; i < 8; ++i) d.st->prof[i] = pf[i];
    d.st->n_cand_total = cand_total;
    d.st->n_cand_max = cand_max;
    if (d.shard) *d.xepoch = xe;
  }
}

// ---- finalisation ---------------------------------------------------------------------------
__global__ void k_pack(Dev d, int buf, unsigned long long* ubits, unsigned long long* kbits,
                       unsigned long long* counts) {
  // one thread per wire (coalesced byte loads), a ballot packs 32 of them; the two warps of a 64-wire
  // word write its halves; counters are reduced per block
  __shared__ unsigned int s_cnt[3];
  if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // wire i + 1
  const uint32_t nbits = ((d.V + 63) / 64) * 64;
  uint32_t f = 0, nt = 0;
  if (i < d.V) {
    f = d.F[buf][i + 1];
    nt = d.nontriv[i + 1];
  }
  const unsigned int mu = __ballot_sync(0xffffffffu, (f & WF_U) != 0);
  const unsigned int mk = __ballot_sync(0xffffffffu, (f & WF_K) != 0);
  const unsigned int mn = __ballot_sync(0xffffffffu, nt != 0);
  if ((threadIdx.x & 31u) == 0 && i < nbits) {
    reinterpret_cast<unsigned int*>(ubits)[i >> 5] = mu;  // little endian: half (i / 32) & 1 of word i / 64
    reinterpret_cast<unsigned int*>(kbits)[i >> 5] = mk;
    atomicAdd(&s_cnt[0], (unsigned int)__popc(mu));
    atomicAdd(&s_cnt[1], (unsigned int)__popc(mn));
    atomicAdd(&s_cnt[2], (unsigned int)__popc(mu & mn));
  }
  __syncthreads();
  if (threadIdx.x < 3 && s_cnt[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
}
__global__ void k_targets(Dev d, int buf, unsigned long long* counts) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.n_targets) return;
  if (d.F[buf][d.targets[i]] & WF_U) atomicAdd(counts + 3, 1ULL);
}
// materialise lb/ub/values/nvalues from ranks and value sources
__device__ __forceinline__ uint8_t wire_values(const Dev& d, uint32_t w, fr::u256& v0, fr::u256& v1) {
  const uint32_t vsrc = d.valsrc[w];
  uint8_t n = 0;
  v0 = fr::make_u256(0, 0, 0, 0);
  v1 = v0;
  if (vsrc == VS_ONE) {
    n = 1;
    v0 = fr::make_u256(1, 0, 0, 0);
  } else if (vsrc == VS_ONEZERO) {
    n = 2;
    v0 = fr::make_u256(1, 0, 0, 0);
  } else if (vsrc & VS_2B) {
    n = 1;
    v0 = d.tvals[vsrc & 0x3fffffffu];
  } else if (vsrc & VS_2A) {
    n = 2;
    v0 = d.roots[2 * (vsrc & 0x3fffffffu)];
    v1 = d.roots[2 * (vsrc & 0x3fffffffu) + 1];
  }
  return n;
}
__global__ void k_export(Dev d, int buf, fr::u256* lb, fr::u256* ub, uint8_t* nvalues,
                         fr::u256* values) {
  uint32_t w = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (w > d.V) return;
  if (lb) lb[w - 1] = d.table[d.LBR[buf][w]];
  if (ub) ub[w - 1] = d.table[d.UBR[buf][w]];
  if (nvalues || values) {
    fr::u256 v0, v1;
    const uint8_t n = wire_values(d, w, v0, v1);
    if (nvalues) nvalues[w - 1] = n;
    if (values) {
      values[2 * (w - 1)] = v0;
      values[2 * (w - 1) + 1] = v1;
    }
  }
}

// ---- report path (:1599-1635) ---------------------------------------------------------------
// One thread per row: a row is listed when one of its wires (non-zero terms only = getVariables, :36-56)
// is not unique (:1612-1620); its wires but wire 1 are marked for the state listing (:1627-1633).  A
// ballot packs 32 rows into a word of the bitmap.
__global__ void k_bad_rows(Dev d, int buf, unsigned int* row_bits, unsigned int* wire_mark,
                           unsigned long long* counts) {
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  const uint8_t* F = d.F[buf];
  bool bad = false;
  if (row < d.N) {
    const uint32_t b = d.seg[3 * row], e = d.seg[3 * row + 3];
    for (uint32_t j = b; j < e && !bad; ++j) bad = !(F[d.col[j]] & WF_U);
    if (bad)
      for (uint32_t j = b; j < e; ++j) {
        const uint32_t w = d.col[j];
        if (w != 1) atomicOr(wire_mark + ((w - 1) >> 5), 1u << ((w - 1) & 31u));
      }
  }
  const unsigned int m = __ballot_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31u) == 0 && row < ((d.N + 63) / 64) * 64) {  // both halves of the last 64-bit word
    row_bits[row >> 5] = m;
    if (m) atomicAdd(counts, (unsigned long long)__popc(m));
  }
}
// The marked wires in ascending order: ONE block; every thread owns a contiguous run of mask words,
// the run totals are scanned in shared memory, then each thread emits its wires behind its offset.
__global__ void __launch_bounds__(1024) k_report_list(const unsigned int* wire_mark, uint32_t n_words,
                                                      uint32_t* wires, unsigned long long* counts) {
  __shared__ uint32_t s_off[1024];
  const uint32_t per = (n_words + blockDim.x - 1) / blockDim.x;
  const uint32_t lo = min(threadIdx.x * per, n_words), hi = min(lo + per, n_words);
  uint32_t c = 0;
  for (uint32_t i = lo; i < hi; ++i) c += (uint32_t)__popc(wire_mark[i]);
  s_off[threadIdx.x] = c;
  __syncthreads();
  for (uint32_t st = 1; st < blockDim.x; st <<= 1) {  // Hillis-Steele inclusive scan
    const uint32_t v = threadIdx.x >= st ? s_off[threadIdx.x - st] : 0u;
    __syncthreads();
    s_off[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t o = s_off[threadIdx.x] - c;
  for (uint32_t i = lo; i < hi; ++i)
    for (unsigned int m = wire_mark[i]; m; m &= m - 1) wires[o++] = i * 32u + (uint32_t)(__ffs((int)m) - 1) + 1u;
  if (threadIdx.x == blockDim.x - 1) counts[1] = s_off[threadIdx.x];
}
__global__ void k_report_export(Dev d, int buf, const uint32_t* wires, uint32_t n, uint8_t* flags, fr::u256* lb,
                                fr::u256* ub, uint8_t* nvalues, fr::u256* values) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t w = wires[i];
  if (flags) flags[i] = d.F[buf][w] & (WF_U | WF_K);
  if (lb) lb[i] = d.table[d.LBR[buf][w]];
  if (ub) ub[i] = d.table[d.UBR[buf][w]];
  if (nvalues || values) {
    fr::u256 v0, v1;
    const uint8_t nv = wire_values(d, w, v0, v1);
    if (nvalues) nvalues[i] = nv;
    if (values) {
      values[2 * i] = v0;
      values[2 * i + 1] = v1;
    }
  }
}

// ---- state reset ----------------------------------------------------------------------------
// Everything a solve starts from, in ONE launch (a dozen small memsets cost more than the kernel): wire
// state, row latches, the Case-5 cache, long-row bookkeeping, special latches, record counters, the
// barrier words and the status block.
__global__ void k_reset_all(Dev d) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= d.V + 3) {  // wires
    const uint8_t hv = (i <= d.V && d.inv_head[i].x > HEAVY_DEG) ? (uint8_t)WF_HEAVY : (uint8_t)0;
    d.F[0][i] = hv;
    d.F[1][i] = hv;
    if (i <= d.V) {
      d.LBR[0][i] = d.r0;
      d.LBR[1][i] = d.r0;
      d.UBR[0][i] = d.rpm1;
      d.UBR[1][i] = d.rpm1;
      d.abz[i] = -1;
      d.valsrc[i] = VS_NONE;
      d.abz_claim[i] = ~0ULL;
    }
    if (d.shard) {
#pragma unroll
      for (int l = 0; l < 5; ++l) d.wflag[l][i] = 0;
    }
  }
  if (i <= d.N) d.solved[i] = 0;  // rows
  if (i < (d.N + 31u) / 32u) d.p2_open[i] = (i + 1u) * 32u <= d.N ? 0xffffffffu : ((1u << (d.N & 31u)) - 1u);
  if (i < d.N) d.c5sig[i] = 0xffffffffu;
  if (i < d.n_long) {
    d.long_done[i] = 0;
    d.long_stamp[i] = 0;
    LongP2 z;
    z.hs = z.hx = 0;
    z.k = z.w1 = z.gr = z.bad = 0;
    d.long_p2[i] = z;
  }
  if (i < d.n_specials) d.sp_solved[i] = 0;
  if (d.n_specials && i <= d.V + 1) d.sp_tag[i] = SP_TAG_NONE;  // (a solve that stopped on an error inside a pass leaves marks)
  if (i < 8) {
    d.rec_count[i] = 0;
    d.dcnt[i] = 0;
    d.bnd_flag[i] = 0;
  }
  if (i < 128) d.barrier[i] = 0;
  if (i < sizeof(Status) / sizeof(unsigned int)) reinterpret_cast<unsigned int*>(d.st)[i] = 0;
}
__global__ void k_reset_known(Dev d) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.n_known) return;
  uint32_t w = d.known[i];
  or_flag(d.F[0], w, WF_U | WF_K);
  or_flag(d.F[1], w, WF_U | WF_K);
  if (w == 1) d.valsrc[w] = VS_ONE;
}

// ---- host-side launchers ----------------------------------------------------------------------
static inline unsigned int blocks_for(uint64_t n, unsigned int t) { return (unsigned int)((n + t - 1) / t); }

cudaError_t launch_reset(const Dev& d, int grid, cudaStream_t s) {
  (void)grid;
  uint64_t n = (uint64_t)d.V + 4;
  n = std::max<uint64_t>(n, (uint64_t)d.N + 1);
  n = std::max<uint64_t>(n, d.n_long);
  n = std::max<uint64_t>(n, d.n_specials);
  n = std::max<uint64_t>(n, 128);
  k_reset_all<<<blocks_for(n, 256), 256, 0, s>>>(d);
  if (d.n_known) k_reset_known<<<blocks_for(d.n_known, 256), 256, 0, s>>>(d);
  return cudaGetLastError();
}
cudaError_t launch_clear_p2_table(const Dev& d, cudaStream_t s) {
  const size_t cap = (size_t)d.h_mask + 1;
  cudaMemsetAsync(d.h_key, 0, cap * sizeof(unsigned long long), s);
  cudaMemsetAsync(d.h_cnt, 0, cap * sizeof(uint32_t), s);
  cudaMemsetAsync(d.h_head, 0, cap * sizeof(uint32_t), s);
  return cudaGetLastError();
}

int p1_threads() { return P1_THREADS; }
static size_t solve_max_smem() { return (size_t)P1_MAX_KS * 2 * sizeof(uint4) * P1_THREADS; }  // 192 KB of row records
#define ECNE_MAX_DEVICES 64
int p1_grid_size(int device) {
  static int cached[ECNE_MAX_DEVICES] = {0};
  if (device < 0 || device >= ECNE_MAX_DEVICES) return 0;
  if (cached[device]) return cached[device];
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  // (function attributes are per device: the caller has made `device` current)
  cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)solve_max_smem());
  cached[device] = sms;  // one block per SM (persistent, cooperative)
  return cached[device];
}

// the whole fixpoint: one cooperative launch
cudaError_t launch_solve(const Dev& d, unsigned int max_rounds, int grid, cudaStream_t s) {
  // the descriptor goes to constant memory (skipped when it is what the last solve used)
  // (per device: every device has its own copy of the __constant__ symbol; the staging copy must outlive the
  // asynchronous upload, hence one static slot per device)
  static Dev last[ECNE_MAX_DEVICES];
  static bool have_last[ECNE_MAX_DEVICES] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= ECNE_MAX_DEVICES) return cudaErrorInvalidDevice;
  if (!have_last[dev] || memcmp(&last[dev], &d, sizeof(Dev)) != 0) {
    memcpy(&last[dev], &d, sizeof(Dev));
    cudaError_t e = cudaMemcpyToSymbolAsync(c_dev, &last[dev], sizeof(Dev), 0, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    have_last[dev] = true;
  }
  unsigned int mr = max_rounds;
  const uint32_t rows = d.row_hi - d.row_lo;
  const uint32_t nthreads = (uint32_t)grid * P1_THREADS;
  int ks = (int)((rows + nthreads - 1) / nthreads);
  if (ks > P1_MAX_KS) ks = P1_MAX_KS;
  const size_t rec_bytes = (size_t)(ks > 0 ? ks : 1) * 2 * sizeof(uint4) * P1_THREADS;
  size_t smem = rec_bytes;
  void* args[] = {&mr, &ks};
  return cudaLaunchCooperativeKernel((void*)k_solve, dim3(grid), dim3(P1_THREADS), args, smem, s);
}

void launch_finalize(const Dev& d, int buf, unsigned long long* ubits, unsigned long long* kbits,
                     unsigned long long* counts, cudaStream_t s) {
  cudaMemsetAsync(counts, 0, 4 * sizeof(unsigned long long), s);
  k_pack<<<blocks_for((uint64_t)((d.V + 63) / 64) * 64, 256), 256, 0, s>>>(d, buf, ubits, kbits, counts);
  if (d.n_targets) k_targets<<<blocks_for(d.n_targets, 128), 128, 0, s>>>(d, buf, counts);
}
void launch_export(const Dev& d, int buf, fr::u256* lb, fr::u256* ub, uint8_t* nvalues,
                   fr::u256* values, cudaStream_t s) {
  k_export<<<blocks_for(d.V, 256), 256, 0, s>>>(d, buf, lb, ub, nvalues, values);
}

void launch_bad_rows(const Dev& d, int buf, unsigned int* row_bits, unsigned int* wire_mark, uint32_t* wires,
                     unsigned long long* counts, cudaStream_t s) {
  const uint32_t mark_words = (d.V + 31) / 32;
  cudaMemsetAsync(counts, 0, 2 * sizeof(unsigned long long), s);
  cudaMemsetAsync(wire_mark, 0, (size_t)mark_words * 4, s);
  if (d.N) k_bad_rows<<<blocks_for(((uint64_t)d.N + 63) / 64 * 64, 256), 256, 0, s>>>(d, buf, row_bits, wire_mark, counts);
  k_report_list<<<1, 1024, 0, s>>>(wire_mark, mark_words, wires, counts);
}
void launch_report_export(const Dev& d, int buf, const uint32_t* wires, uint32_t n, uint8_t* flags, fr::u256* lb,
                          fr::u256* ub, uint8_t* nvalues, fr::u256* values, cudaStream_t s) {
  if (n) k_report_export<<<blocks_for(n, 256), 256, 0, s>>>(d, buf, wires, n, flags, lb, ub, nvalues, values);
}

}  // namespace ecne

// ---- the second build of this file -----------------------------------------------------------------------------------
// The library carries the solve kernel twice (ecneproject_b200/build.py): 512 threads x 128 registers per block — the
// latency-bound dependency chains of small problems run spill-free — and 1024 x 64, whose 32 warps per SM hide more of a
// bandwidth-bound sweep (S16: dense rounds 1.5x faster).  The second copy is this same file compiled with
// -Decne=ecne_v1024 (its own namespace, its own __constant__ descriptor) and reaches abi.cu through these two C symbols.
#ifdef ECNE_VARIANT_SUFFIX
#define ECNE_CAT2(a, b) a##b
#define ECNE_CAT(a, b) ECNE_CAT2(a, b)
extern "C" cudaError_t ECNE_CAT(ecne_launch_solve_, ECNE_VARIANT_SUFFIX)(const void* dev, unsigned int max_rounds, int grid,
                                                                         cudaStream_t s) {
  return ecne::launch_solve(*reinterpret_cast<const ecne::Dev*>(dev), max_rounds, grid, s);
}
extern "C" int ECNE_CAT(ecne_p1_grid_size_, ECNE_VARIANT_SUFFIX)(int device) { return ecne::p1_grid_size(device); }
extern "C" int ECNE_CAT(ecne_p1_threads_, ECNE_VARIANT_SUFFIX)(void) { return ecne::p1_threads(); }
#endif
