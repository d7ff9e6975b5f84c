// sweep.cuh — the single-row rule set (the body of the reference's queue loop,
// /root/reference/src/R1CSConstraintSolver.jl:805-1349) evaluated for one constraint row against a
// snapshot of the wire state.  One thread per row for ordinary rows (p99 row length is 3), one
// warp per row — lanes striding the terms, ballot/shuffle reductions — for long rows.
//
// Inside one evaluation the cases run in the reference's order and each sees the effects of the
// earlier ones on this row's own wires (a small register overlay), exactly like one pop.
#pragma once
#include "engine.cuh"

namespace ecne {

// The problem descriptor of the running solve lives in constant memory: every device function reads
// its pointers and sizes through the constant cache instead of a by-reference copy of a 900-byte kernel
// parameter in local memory (which costs an L2 round trip per field after every L1 invalidation).
// One solve runs at a time per process (the API is single-threaded).
__constant__ Dev c_dev;

// state pointers of read buffer `rb` (bit 0: the buffer, bit ST_TAG_CG: read through the L2, engine.cuh)
__device__ __forceinline__ uintptr_t st_tag(int rb) { return (uintptr_t)(rb & ST_TAG_CG) << 62; }
__device__ __forceinline__ const uint8_t* st_F(int rb) { return (const uint8_t*)((uintptr_t)c_dev.F[rb & 1] | st_tag(rb)); }
__device__ __forceinline__ const uint32_t* st_L(int rb) { return (const uint32_t*)((uintptr_t)c_dev.LBR[rb & 1] | st_tag(rb)); }
__device__ __forceinline__ const uint32_t* st_U(int rb) { return (const uint32_t*)((uintptr_t)c_dev.UBR[rb & 1] | st_tag(rb)); }

// ---- records of a stretch of rounds that one warp runs alone (kernels.cu warp_solo) -----------------------------------
// The warp is producer and consumer of its update records: they stay in shared memory, with the head of the wire's row
// list fetched by the emitter next to its state atomic — no slot atomic, no record store / load through the L2, no
// fence and no counter read between two rounds.  `list` carries LIST_SOLO and the queue to write (LIST_SOLO_Q).  A
// round that emits more than the queue holds (or fires a whole long row, emit_unique_all) spills to the global list
// and ends the stretch.
#define LIST_SOLO 16
#define LIST_SOLO_Q 32
#define SOLO_Q_CAP 64
struct __align__(16) SoloRec {
  Rec r;
  uint4 head;  // inv_head[r.wire]
};
__shared__ SoloRec s_soloq[2][SOLO_Q_CAP];
__shared__ unsigned int s_soloq_n[2];
__shared__ unsigned int s_solo_flags;  // bit0: a bound moved, bit1: a heavy wire changed, bit2: records spilled to the global list

// Apply one update {OR bits into F, max lbr into LBR, min ubr into UBR} to buffer `buf`.  The bits of
// the derived test "bounds == [0,1]" follow from the update's own values, so every part is a
// commutative, monotone RMW: no ordering between concurrent updates of a wire is needed.
// want_old = false (replays): fire-and-forget REDs.  want_old = true: returns bit0: the update changed
// something in `buf`, bit1: a bound moved, bit2: the wire is a heavy one (WF_HEAVY).
template <bool want_old>
__device__ __noinline__ uint32_t apply_update_t(int buf, uint32_t w, uint32_t bits, uint32_t lbr, uint32_t ubr) {
  const Dev& d = c_dev;
  bool bch = false;
  if (lbr != ECNE_NO_LB || ubr != ECNE_NO_UB) {
    bits |= WF_BND;
    if (ubr != ECNE_NO_UB && ubr <= d.r1) bits |= WF_UB01;
    if ((ubr != ECNE_NO_UB && ubr < d.r1) || (lbr != ECNE_NO_LB && lbr > d.r0)) bits |= WF_NOT01;
    if (want_old) {
      if (lbr != ECNE_NO_LB) bch |= atomicMax(d.LBR[buf] + w, lbr) < lbr;
      if (ubr != ECNE_NO_UB) bch |= atomicMin(d.UBR[buf] + w, ubr) > ubr;
    } else {
      if (lbr != ECNE_NO_LB) atomicMax(d.LBR[buf] + w, lbr);
      if (ubr != ECNE_NO_UB) atomicMin(d.UBR[buf] + w, ubr);
    }
  }
  uint32_t ret = bch ? 3u : 0u;
  if (bits) {
    unsigned int* word = (unsigned int*)(d.F[buf] + (w & ~3u));
    const unsigned int sh = (w & 3u) * 8;
    if (want_old) {
      const uint32_t old = (atomicOr(word, bits << sh) >> sh) & 0xffu;
      if (bits & ~old) ret |= 1u;
      if (old & WF_HEAVY) ret |= 4u;
    } else {
      atomicOr(word, bits << sh);  // result unused: compiles to RED
    }
  }
  return ret;
}
__device__ __forceinline__ void apply_update(const Dev&, int buf, uint32_t w, uint32_t bits, uint32_t lbr,
                                             uint32_t ubr) {
  apply_update_t<false>(buf, w, bits, lbr, ubr);
}
// The same update issued as RETURNING atomics whose raw results are handed back untouched: the caller consumes them
// (any use) at the point where it needs the update to have been performed at the L2 — until then the atomics are
// simply in flight (kernels.cu warp_solo: the replay of a round's records).
__device__ __forceinline__ void apply_update_issue(int buf, uint32_t w, uint32_t bits, uint32_t lbr, uint32_t ubr,
                                                   uint32_t& o0, uint32_t& o1, uint32_t& o2) {
  const Dev& d = c_dev;
  o0 = o1 = o2 = 0;
  if (lbr != ECNE_NO_LB || ubr != ECNE_NO_UB) {
    bits |= WF_BND;
    if (ubr != ECNE_NO_UB && ubr <= d.r1) bits |= WF_UB01;
    if ((ubr != ECNE_NO_UB && ubr < d.r1) || (lbr != ECNE_NO_LB && lbr > d.r0)) bits |= WF_NOT01;
    if (lbr != ECNE_NO_LB) o1 = atomicMax(d.LBR[buf] + w, lbr);
    if (ubr != ECNE_NO_UB) o2 = atomicMin(d.UBR[buf] + w, ubr);
  }
  if (bits) {
    unsigned int* word = (unsigned int*)(d.F[buf] + (w & ~3u));
    o0 = atomicOr(word, bits << ((w & 3u) * 8));
  }
}
// a record of list `list` has been consumed (replayed): its wire may be counted again when the list is
// written next (two rounds from now)
__device__ __forceinline__ void consume_rec(const Dev& d, unsigned int list, uint32_t w) {
  if (d.shard) {
    unsigned int* word = (unsigned int*)(d.wflag[list] + (w & ~3u));
    atomicAnd(word, ~(0xffu << ((w & 3u) * 8)));
  }
}

// Apply an update to the write buffer and, when it changed anything there, log it for the other
// buffer.  A row only calls this when its evaluation against the snapshot wants something the snapshot
// does not have; the write buffer can already hold it only because another row of the SAME round got
// there first, so "first writer logs" makes the round's records an exact, duplicate-free list of state
// changes: "no record in a round" is exactly "fixpoint", and the records are the next round's frontier.
// One copy of this code in the kernel (the fixpoint kernel is instruction-cache bound in its serial
// stretches: every inlined copy costs more than the call).
// `list` may carry LIST_NOCOUNT: the round is a sharded dense sweep, whose size the ranks learn from the record
// counts they exchange anyway — no per-record distinct-wire bookkeeping (an extra dependent atomic per record, which
// made a 2.5 M-record round three times slower).
#define LIST_NOCOUNT 8
__device__ __noinline__ void emit_impl(int wbuf, int list_in, uint32_t w, uint32_t bits, uint32_t lbr, uint32_t ubr) {
  const Dev& d = c_dev;
  const int list = list_in & 7;
  const bool count_distinct = d.shard && !(list_in & LIST_NOCOUNT);
  if (list_in & LIST_SOLO) {
    // one warp alone: the head of the wire's row list is fetched while the state atomic is in flight
    const uint4 hd = __ldcg(reinterpret_cast<const uint4*>(d.inv_head) + w);
    const uint32_t r = apply_update_t<true>(wbuf, w, bits, lbr, ubr);
    const uint32_t want = ((r & 2u) ? 1u : 0u) | (((r & 1u) && (r & 4u)) ? 2u : 0u);
    if (want) atomicOr(&s_solo_flags, want);
    if (!(r & 1u)) return;
    Rec rr;
    rr.wire = w;
    rr.bits = (lbr != ECNE_NO_LB || ubr != ECNE_NO_UB) ? (bits | WF_BND) : bits;
    rr.lbr = lbr;
    rr.ubr = ubr;
    const int q = (list_in & LIST_SOLO_Q) ? 1 : 0;
    const unsigned int i = atomicAdd(&s_soloq_n[q], 1u);
    if (i < SOLO_Q_CAP) {
      s_soloq[q][i].r = rr;
      s_soloq[q][i].head = hd;
      return;
    }
    atomicOr(&s_solo_flags, 4u);  // queue full: this record goes to the global list, the stretch ends with this round
    const unsigned int gi = atomicAdd(d.rec_count + list, 1u);
    if (gi < d.rec_cap)
      d.recs[list][gi] = rr;
    else
      d.st->rec_overflow = 1;
    return;
  }
  const uint32_t r = apply_update_t<true>(wbuf, w, bits, lbr, ubr);
  // round flags (bit0: a bound moved, bit1: a heavy wire changed => the next round is dense): set once —
  // 160 k rows fix a bound in ecdsa's first round, and 160 k REDs on one address serialise in one L2 slice
  const uint32_t want = ((r & 2u) ? 1u : 0u) | (((r & 1u) && (r & 4u)) ? 2u : 0u);
  if (want && (__ldcg(d.bnd_flag + list) & want) != want) atomicOr(d.bnd_flag + list, want);
  if (!(r & 1u)) return;
  if (count_distinct) {  // replicated rounds of a sharded run count the distinct wires they change (same on every rank)
    unsigned int* word = (unsigned int*)(d.wflag[list] + (w & ~3u));
    const unsigned int sh = (w & 3u) * 8;
    if (!((atomicOr(word, 1u << sh) >> sh) & 1u)) atomicAdd(d.dcnt + list, 1u);
  }
  // warp-aggregated slot allocation: the lanes that reach this point together take one atomic (a round
  // that changes 300 k wires would otherwise serialise 300 k RMWs on one L2 address)
  const unsigned int am = __activemask();
  const unsigned int lane = threadIdx.x & 31u;
  const int leader = __ffs((int)am) - 1;
  unsigned int i = 0;
  if ((int)lane == leader) i = atomicAdd(d.rec_count + list, (unsigned int)__popc(am));
  i = __shfl_sync(am, i, leader) + (unsigned int)__popc(am & ((1u << lane) - 1u));
  if (i < d.rec_cap) {
    Rec rr;
    rr.wire = w;
    rr.bits = (lbr != ECNE_NO_LB || ubr != ECNE_NO_UB) ? (bits | WF_BND) : bits;
    rr.lbr = lbr;
    rr.ubr = ubr;
    d.recs[list][i] = rr;
  } else {
    d.st->rec_overflow = 1;
  }
}
__device__ __forceinline__ void emit(const Dev&, int wbuf, int list, uint32_t w, uint32_t bits,
                                     uint32_t lbr = ECNE_NO_LB, uint32_t ubr = ECNE_NO_UB) {
  emit_impl(wbuf, list, w, bits, lbr, ubr);
}

// The (at most two) updates of one inline-row evaluation, emitted TOGETHER: every atomic of both updates — flags, lower
// and upper rank of either wire, and on a solo stretch the heads of their row lists — is issued before the first
// result is looked at, so the emits of a lane cost one atomic round trip instead of up to six dependent ones (measured
// on ecdsa's solo rounds: the emits were 5.2 k of a round's 8 k cycles).  Same updates, same "first writer logs"
// records as two emit() calls.  An update with bits == 0 and no bound is empty.
__device__ __noinline__ void emit2_impl(int wbuf, int list_in, uint32_t w0, uint32_t b0, uint32_t l0, uint32_t u0,
                                        uint32_t w1, uint32_t b1, uint32_t l1, uint32_t u1) {
  const Dev& d = c_dev;
  const int list = list_in & 7;
  const bool has0 = (b0 | l0 | ~u0) != 0, has1 = (b1 | l1 | ~u1) != 0;
  const bool bd0 = l0 != ECNE_NO_LB || u0 != ECNE_NO_UB, bd1 = l1 != ECNE_NO_LB || u1 != ECNE_NO_UB;
  const uint32_t rb0 = bd0 ? (b0 | WF_BND) : b0, rb1 = bd1 ? (b1 | WF_BND) : b1;  // the bits a record carries
  if (bd0) {
    b0 |= WF_BND;
    if (u0 != ECNE_NO_UB && u0 <= d.r1) b0 |= WF_UB01;
    if ((u0 != ECNE_NO_UB && u0 < d.r1) || (l0 != ECNE_NO_LB && l0 > d.r0)) b0 |= WF_NOT01;
  }
  if (bd1) {
    b1 |= WF_BND;
    if (u1 != ECNE_NO_UB && u1 <= d.r1) b1 |= WF_UB01;
    if ((u1 != ECNE_NO_UB && u1 < d.r1) || (l1 != ECNE_NO_LB && l1 > d.r0)) b1 |= WF_NOT01;
  }
  // ---- issue
  const bool solo = (list_in & LIST_SOLO) != 0;
  uint4 hd0 = make_uint4(0, 0, 0, 0), hd1 = hd0;
  uint32_t oL0 = l0, oU0 = u0, oF0 = 0xffu, oL1 = l1, oU1 = u1, oF1 = 0xffu;
  const unsigned int sh0 = (w0 & 3u) * 8, sh1 = (w1 & 3u) * 8;
  if (has0) {
    if (solo) hd0 = __ldcg(reinterpret_cast<const uint4*>(d.inv_head) + w0);
    if (l0 != ECNE_NO_LB) oL0 = atomicMax(d.LBR[wbuf] + w0, l0);
    if (u0 != ECNE_NO_UB) oU0 = atomicMin(d.UBR[wbuf] + w0, u0);
    if (b0) oF0 = atomicOr((unsigned int*)(d.F[wbuf] + (w0 & ~3u)), b0 << sh0) >> sh0;
  }
  if (has1) {
    if (solo) hd1 = __ldcg(reinterpret_cast<const uint4*>(d.inv_head) + w1);
    if (l1 != ECNE_NO_LB) oL1 = atomicMax(d.LBR[wbuf] + w1, l1);
    if (u1 != ECNE_NO_UB) oU1 = atomicMin(d.UBR[wbuf] + w1, u1);
    if (b1) oF1 = atomicOr((unsigned int*)(d.F[wbuf] + (w1 & ~3u)), b1 << sh1) >> sh1;
  }
  // ---- consume: bit0 the update changed something, bit1 a bound moved, bit2 the wire is a heavy one
  oF0 &= 0xffu;
  oF1 &= 0xffu;
  const bool mv0 = has0 && ((l0 != ECNE_NO_LB && oL0 < l0) || (u0 != ECNE_NO_UB && oU0 > u0));
  const bool mv1 = has1 && ((l1 != ECNE_NO_LB && oL1 < l1) || (u1 != ECNE_NO_UB && oU1 > u1));
  const bool ch0 = has0 && (mv0 || (b0 && (b0 & ~oF0))), ch1 = has1 && (mv1 || (b1 && (b1 & ~oF1)));
  const bool hv0 = ch0 && b0 && (oF0 & WF_HEAVY), hv1 = ch1 && b1 && (oF1 & WF_HEAVY);  // a heavy wire changed
  const uint32_t want = ((mv0 || mv1) ? 1u : 0u) | ((hv0 || hv1) ? 2u : 0u);
  Rec r0, r1;
  r0.wire = w0;
  r0.bits = rb0;
  r0.lbr = l0;
  r0.ubr = u0;
  r1.wire = w1;
  r1.bits = rb1;
  r1.lbr = l1;
  r1.ubr = u1;
  if (solo) {
    if (want) atomicOr(&s_solo_flags, want);
    const unsigned int n = (ch0 ? 1u : 0u) + (ch1 ? 1u : 0u);
    if (!n) return;
    const int q = (list_in & LIST_SOLO_Q) ? 1 : 0;
    unsigned int i = atomicAdd(&s_soloq_n[q], n);
    if (ch0) {
      if (i < SOLO_Q_CAP) {
        s_soloq[q][i].r = r0;
        s_soloq[q][i].head = hd0;
      } else {
        atomicOr(&s_solo_flags, 4u);  // queue full: this record goes to the global list, the stretch ends with this round
        const unsigned int gi = atomicAdd(d.rec_count + list, 1u);
        if (gi < d.rec_cap)
          d.recs[list][gi] = r0;
        else
          d.st->rec_overflow = 1;
      }
      ++i;
    }
    if (ch1) {
      if (i < SOLO_Q_CAP) {
        s_soloq[q][i].r = r1;
        s_soloq[q][i].head = hd1;
      } else {
        atomicOr(&s_solo_flags, 4u);
        const unsigned int gi = atomicAdd(d.rec_count + list, 1u);
        if (gi < d.rec_cap)
          d.recs[list][gi] = r1;
        else
          d.st->rec_overflow = 1;
      }
    }
    return;
  }
  if (want && (__ldcg(d.bnd_flag + list) & want) != want) atomicOr(d.bnd_flag + list, want);
  const unsigned int n = (ch0 ? 1u : 0u) + (ch1 ? 1u : 0u);
  if (d.shard && !(list_in & LIST_NOCOUNT)) {  // replicated rounds of a sharded run count the distinct wires they change
    if (ch0) {
      unsigned int* word = (unsigned int*)(d.wflag[list] + (w0 & ~3u));
      if (!((atomicOr(word, 1u << sh0) >> sh0) & 1u)) atomicAdd(d.dcnt + list, 1u);
    }
    if (ch1) {
      unsigned int* word = (unsigned int*)(d.wflag[list] + (w1 & ~3u));
      if (!((atomicOr(word, 1u << sh1) >> sh1) & 1u)) atomicAdd(d.dcnt + list, 1u);
    }
  }
  // warp-aggregated slot allocation over the lanes that are here together: one atomic for all their records
  const unsigned int am = __activemask();
  const unsigned int lane = threadIdx.x & 31u;
  const unsigned int m1 = __ballot_sync(am, n >= 1), m2 = __ballot_sync(am, n >= 2);
  const unsigned int below = (1u << lane) - 1u;
  const unsigned int total = (unsigned int)(__popc(m1) + __popc(m2));
  if (!total) return;
  const int leader = __ffs((int)am) - 1;
  unsigned int base = 0;
  if ((int)lane == leader) base = atomicAdd(d.rec_count + list, total);
  base = __shfl_sync(am, base, leader) + (unsigned int)(__popc(m1 & below) + __popc(m2 & below));
  if (ch0) {
    if (base < d.rec_cap)
      d.recs[list][base] = r0;
    else
      d.st->rec_overflow = 1;
    ++base;
  }
  if (ch1) {
    if (base < d.rec_cap)
      d.recs[list][base] = r1;
    else
      d.st->rec_overflow = 1;
  }
}

template <int G>
struct Grp {
  static __device__ __forceinline__ uint32_t lane() { return G == 1 ? 0u : (threadIdx.x & 31u); }
  static __device__ __forceinline__ uint32_t sum(uint32_t x) {
    if (G == 1) return x;
    return __reduce_add_sync(0xffffffffu, x);
  }
  static __device__ __forceinline__ uint32_t max(uint32_t x) {
    if (G == 1) return x;
    return __reduce_max_sync(0xffffffffu, x);
  }
  static __device__ __forceinline__ uint32_t min(uint32_t x) {
    if (G == 1) return x;
    return __reduce_min_sync(0xffffffffu, x);
  }
  static __device__ __forceinline__ bool all(bool p) {
    if (G == 1) return p;
    return __all_sync(0xffffffffu, p);
  }
  static __device__ __forceinline__ bool any(bool p) {
    if (G == 1) return p;
    return __any_sync(0xffffffffu, p);
  }
};

// Walk terms [s, e) of the CSR with G lanes, U terms per lane in flight: all U wire indices are
// loaded first, then all U state bytes of array `A` (F or B) are gathered, then `fn(wire, byte)`
// runs.  This turns 2*ceil(n/G) dependent L2 round trips into 2*ceil(n/(G*U)).
template <int G, int U, class Fn>
__device__ __forceinline__ void scan_terms(const Dev& d, const uint8_t* A, uint32_t s, uint32_t e,
                                           uint32_t lane, Fn fn) {
  for (uint32_t base = s; base < e; base += G * U) {
    uint32_t w[U], f[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t t = base + (uint32_t)u * G + lane;
      w[u] = t < e ? d.col[t] : 0xffffffffu;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) f[u] = w[u] != 0xffffffffu ? ld_flag(A, w[u]) : 0u;
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (w[u] != 0xffffffffu) fn(w[u], f[u]);
  }
}
#define SCAN_U(G) ((G) == 1 ? 4 : 8)

// What earlier cases of this evaluation did to the row's own wires (group-uniform).
struct Overlay {
  uint32_t w[2], lb[2], ub[2];
  bool k[2];
  int n;
  bool all_unique;  // every C key is unique by now
  __device__ __forceinline__ int find(uint32_t x) const {
    if (n > 0 && w[0] == x) return 0;
    if (n > 1 && w[1] == x) return 1;
    return -1;
  }
  __device__ __forceinline__ void set(uint32_t x, uint32_t l, uint32_t u, bool known) {
    int i = find(x);
    if (i < 0) {
      if (n >= 2) return;  // cannot happen: at most two wires get local bounds
      i = n++;
      w[i] = x;
      k[i] = false;
    }
    lb[i] = l;
    ub[i] = u;
    k[i] = k[i] || known;
  }
};

struct RowCtx {
  int rbuf, wbuf, list;
  uint32_t row, rf, s2, s3;
  Overlay ov;
  __device__ __forceinline__ void bounds(uint32_t w, uint32_t& l, uint32_t& u) const {
    const Dev& d = c_dev;
    int i = ov.find(w);
    if (i >= 0) {
      l = ov.lb[i];
      u = ov.ub[i];
    } else {
      l = ld_u32(st_L(rbuf), w);
      u = ld_u32(st_U(rbuf), w);
    }
  }
  __device__ __forceinline__ bool b01(uint32_t w) const {
    const Dev& d = c_dev;
    int i = ov.find(w);
    if (i >= 0) return ov.lb[i] == d.r0 && ov.ub[i] == d.r1;
    return is01(ld_flag(st_F(rbuf), w));
  }
  __device__ __forceinline__ bool uniq(uint32_t w) const {
    const Dev& d = c_dev;
    return ov.all_unique || (ld_flag(st_F(rbuf), w) & WF_U);
  }
  __device__ __forceinline__ bool known(uint32_t w) const {
    const Dev& d = c_dev;
    int i = ov.find(w);
    if (i >= 0 && ov.k[i]) return true;
    return ld_flag(st_F(rbuf), w) & WF_K;
  }
  __device__ __forceinline__ void out(uint32_t w, uint32_t bits, uint32_t l = ECNE_NO_LB,
                                      uint32_t u = ECNE_NO_UB) const {
    emit(c_dev, wbuf, list, w, bits, l, u);
  }
};

// The firing of Case 5 / Case 6: every non-unique wire among terms [s, e) becomes unique + known — on a long row
// up to 1024 updates from ONE warp (ecdsa: 52 decoder rows fire in one round).  Through c.out() that is 8 dependent
// {atomicOr, flag load, slot atomic} chains per lane and pass (130 k cycles for a 1025-term row); here a lane
// issues its 8 atomics before it looks at any result and the warp takes one record-slot allocation per 256 terms.
// Same updates, same "first writer logs" records (their order inside the list is not observable).
template <int G>
__device__ __noinline__ void emit_unique_all(const RowCtx& c, const uint8_t* F, uint32_t s, uint32_t e, uint32_t lane) {
  const Dev& d = c_dev;
  const int list = c.list & 7;
  if (G == 1 || (d.shard && !(c.list & LIST_NOCOUNT))) {  // thread-per-row callers and rounds with distinct-wire counting keep the plain path
    scan_terms<G, SCAN_U(G)>(d, F, s, e, lane, [&](uint32_t w, uint32_t f) {
      if (!(f & WF_U)) c.out(w, WF_U | WF_K);
    });
    return;
  }
  constexpr int U = 8;
  constexpr uint32_t BITS = WF_U | WF_K;
  for (uint32_t base = s; base < e; base += 32 * U) {
    uint32_t w[U], f[U], old[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t t = base + (uint32_t)u * 32 + lane;
      w[u] = t < e ? d.col[t] : 0xffffffffu;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) f[u] = w[u] != 0xffffffffu ? ld_flag(F, w[u]) : (uint32_t)WF_U;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      old[u] = 0xffu;
      if (!(f[u] & WF_U)) {
        unsigned int* word = (unsigned int*)(d.F[c.wbuf] + (w[u] & ~3u));
        const unsigned int sh = (w[u] & 3u) * 8;
        old[u] = (atomicOr(word, BITS << sh) >> sh) & 0xffu;
      }
    }
    uint32_t chg = 0;
    bool heavy = false;
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (BITS & ~old[u]) {  // this update changed the write buffer: it is logged
        chg |= 1u << u;
        heavy |= (old[u] & WF_HEAVY) != 0;
      }
    if (heavy && !(__ldcg(d.bnd_flag + list) & 2u)) atomicOr(d.bnd_flag + list, 2u);
    const uint32_t n = (uint32_t)__popc(chg);
    uint32_t incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) continue;
    uint32_t i = 0;
    if (lane == 0 && (c.list & LIST_SOLO)) atomicOr(&s_solo_flags, 4u);  // records in the global list: the solo stretch ends
    if (lane == 0) i = atomicAdd(d.rec_count + list, total);
    i = __shfl_sync(0xffffffffu, i, 0) + incl - n;
#pragma unroll
    for (int u = 0; u < U; ++u)
      if ((chg >> u) & 1u) {
        if (i < d.rec_cap) {
          Rec rr;
          rr.wire = w[u];
          rr.bits = BITS;
          rr.lbr = ECNE_NO_LB;
          rr.ubr = ECNE_NO_UB;
          d.recs[list][i] = rr;
        } else {
          d.st->rec_overflow = 1;
        }
        ++i;
      }
  }
}

// |flip_coeffs(c)| of :1245-1257 on the (possibly flipped) stored coefficient
__device__ __forceinline__ fr::u256 case5_mag(const fr::u256& c, bool flipped) {
  fr::u256 x = flipped ? fr::neg(c) : c;
  if (fr::cmp(x, fr::fold_threshold()) > 0) {
    fr::u256 r;
    fr::sub_cc(r, fr::modulus(), x);
    return r;
  }
  return x;
}

// one link of the mixed-radix chain (:1266-1272): lo = term `tp`, hi = term `tc`
__device__ __noinline__ bool case5_pair_ok(const RowCtx& c, uint32_t tp, uint32_t tc) {
  const Dev& d = c_dev;
  bool flipped = (c.rf & RF_C3_FLIP) != 0;
  fr::u256 dlo = case5_mag(d.coef[tp], flipped);
  fr::u256 dhi = case5_mag(d.coef[tc], flipped);
  if (!fr::divides(dlo, dhi)) return false;
  uint32_t l, u;
  c.bounds(d.col[tp], l, u);
  if (u < l) return true;  // ub - lb negative: the quotient (>= 1) is never <= it
  fr::u256 range;
  fr::sub_cc(range, d.table[u], d.table[l]);
  // fail when hi/lo <= range  <=>  hi <= lo*range
  return fr::cmp_mul(dlo, range, dhi) < 0;
}
// the top test (:1274): fail when d * (ub + 1) > p
__device__ __noinline__ bool case5_top_ok(const RowCtx& c, uint32_t t) {
  const Dev& d = c_dev;
  fr::u256 dm = case5_mag(d.coef[t], (c.rf & RF_C3_FLIP) != 0);
  uint32_t l, u;
  c.bounds(d.col[t], l, u);
  fr::u256 ub1;
  fr::add_cc(ub1, d.table[u], fr::make_u256(1, 0, 0, 0));
  return fr::cmp_mul(dm, ub1, fr::modulus()) <= 0;
}

// Case 5 (:1235-1298) on C terms that are stored sorted by magnitude.  Returns true on success.
template <int G>
__device__ __noinline__ bool case5(const RowCtx& c) {
  const Dev& d = c_dev;
  const uint32_t lane = Grp<G>::lane();
  bool ok = true;
  uint32_t prev = 0xffffffffu;  // previous selected term (group-uniform)
  for (uint32_t base = c.s2; base < c.s3; base += G) {
    uint32_t t = base + lane;
    bool sel = false;
    if (t < c.s3) {
      uint32_t w = d.col[t];
      if (!c.uniq(w)) {
        sel = true;
        if (!c.known(w)) ok = false;  // some unknown key is not is_known (:1260-1264)
      }
    }
    if (G == 1) {
      if (sel) {
        if (ok && prev != 0xffffffffu) ok = case5_pair_ok(c, prev, t);
        prev = t;
      }
      if (!ok) return false;
    } else {
      uint32_t m = __ballot_sync(0xffffffffu, sel);
      if (!__all_sync(0xffffffffu, ok)) return false;
      if (sel) {
        uint32_t below = m & ((1u << lane) - 1u);
        uint32_t p = below ? base + (31 - __clz(below)) : prev;
        if (p != 0xffffffffu) ok = case5_pair_ok(c, p, t);
      }
      if (!__all_sync(0xffffffffu, ok)) return false;
      if (m) prev = base + (31 - __clz(m));
    }
  }
  if (prev == 0xffffffffu) return false;  // no unknown key (:1242-1244)
  if (!case5_top_ok(c, prev)) return false;
  return true;
}

// The rule set on one row (generic path: any length, every pattern).
// Returns true when the row can never fire again (so the caller may stop sweeping it): it is
// latched as solved, or it has no bound pattern and no non-unique wire left in C.
template <int G>
__device__ __noinline__ bool eval_row(const Dev&, int rbuf, int wbuf, int list, uint32_t row,
                                      uint32_t bepoch) {
  const Dev& d = c_dev;
  const uint32_t lane = Grp<G>::lane();
  const uint32_t rf = d.rflags[row];
  uint8_t latch = d.solved[row];
  if (latch & 1) return true;  // equation_solved (:820-822)
  const uint8_t* F = st_F(rbuf);
  const uint32_t s0 = d.seg[3 * row], s2 = d.seg[3 * row + 2], s3 = d.seg[3 * row + 3];
#ifdef ECNE_PROFILE
  // slowest long-row evaluation of the solve, stage by stage (maxima; printed with ECNE_DEBUG_PROF=3)
  long long pt[6] = {clock64(), 0, 0, 0, 0, 0};
  auto pstamp = [&](int k) {
    if (G == 32) pt[k] = clock64();
  };
  auto pflush = [&](int last) {
    if (G == 32 && lane == 0) {
      unsigned long long* q = d.prof + 28000 + 40 * 148 * 4;
      for (int k = 1; k <= last; ++k)
        if (pt[k]) atomicMax(q + k, (unsigned long long)(pt[k] - pt[0]));
      atomicMax(q, (unsigned long long)(clock64() - pt[0]));
    }
  };
#define ECNE_PSTAMP(k) pstamp(k)
#define ECNE_PFLUSH(k) pflush(k)
#else
#define ECNE_PSTAMP(k)
#define ECNE_PFLUSH(k)
#endif

  // ---- gather: non-unique counts over A u B and over C --------------------------------------
  uint32_t nuAB = 0, nuC = 0, wC = 0, kmiss = 0, abzmiss = 0;
  scan_terms<G, SCAN_U(G)>(d, F, s0, s2, lane, [&](uint32_t, uint32_t f) { nuAB += (f & WF_U) ? 0u : 1u; });
  scan_terms<G, SCAN_U(G)>(d, F, s2, s3, lane, [&](uint32_t w, uint32_t f) {
    if (!(f & WF_U)) {
      nuC += 1;
      wC = w;
      kmiss += (f & WF_K) ? 0u : 1u;
      abzmiss += (f & WF_ABZ) ? 0u : 1u;
    }
  });
  if (G > 1) {
    nuAB = Grp<G>::sum(nuAB);
    nuC = Grp<G>::sum(nuC);
    wC = Grp<G>::max(wC);
    kmiss = Grp<G>::sum(kmiss);
    abzmiss = Grp<G>::sum(abzmiss);
  }

  ECNE_PSTAMP(1);  // gather
  RowCtx c;
  c.rbuf = rbuf;
  c.wbuf = wbuf;
  c.list = list;
  c.row = row;
  c.rf = rf;
  c.s2 = s2;
  c.s3 = s3;
  c.ov.n = 0;
  c.ov.all_unique = false;

  // ---- Case 1 (:827-873) ------------------------------------------------------------------
  if (nuAB == 0 && nuC == 1) {
    if (lane == 0) c.out(wC, WF_U | WF_K);
    nuC = 0;
    c.ov.all_unique = true;
  }

  // ---- Case 2a (:875-942) -----------------------------------------------------------------
  if (rf & RF_2A_NOVAR) {
    raise(d, ECNE_E_BOUNDS);
    return true;
  }
  if (rf & RF_2A) {
    const RowAux a = d.aux[row];
    if (!(ld_flag(F, a.w1) & WF_K)) {
      if (rf & RF_2A_DIVZ) {
        raise(d, ECNE_E_DIVZERO);
        return true;
      }
      if (lane == 0) {
        c.out(a.w1, WF_K, ECNE_NO_LB, (rf & RF_2A_BOOL) ? d.r1 : ECNE_NO_UB);
        d.valsrc[a.w1] = VS_2A | a.val_idx;
        d.solved[row] = latch | 1;
      }
    }
  }
  const bool plain = !(rf & (RF_2A | RF_2B | RF_C3 | RF_4A | RF_4B));
  if (!(rf & RF_LINEAR)) return plain && nuC == 0;  // (:944-946)
  if (plain && nuC == 0) return true;

  const RowAux a = d.aux[row];
  // ---- Case 2b (:949-988) -----------------------------------------------------------------
  if (rf & RF_2B) {
    if (!(latch & 2)) {
      if (lane == 0) {
        c.out(a.w1, WF_U | WF_K, a.rank_a, a.rank_a);
        d.valsrc[a.w1] = VS_2B | a.val_idx;
        d.solved[row] = latch | 2;  // the monotone merge makes a second application a no-op
      }
    }
    c.ov.set(a.w1, a.rank_a, a.rank_a, true);
  }

  // ---- Case 3 (:991-1076) -----------------------------------------------------------------
  if (rf & RF_C3) {
    const int norient = (rf & RF_C3_L2) ? 2 : 1;
    for (int o = 0; o < norient; ++o) {
      const uint32_t nk = o == 0 ? a.w2 : a.w5;
      bool ok = true;
      scan_terms<G, SCAN_U(G)>(d, F, s2, s3, lane, [&](uint32_t w, uint32_t f) {
        if (w == nk) return;
        const int oi = c.ov.find(w);
        const bool bit = oi >= 0 ? (c.ov.lb[oi] == d.r0 && c.ov.ub[oi] == d.r1) : is01(f);
        if (!bit) ok = false;
      });
      ok = Grp<G>::all(ok);
      if (!ok) continue;
      uint32_t l, u;
      c.bounds(nk, l, u);
      if (!(rf & RF_C3_TOPBIG) && u > a.rank_b) {  // ub.d > 2^(l-1)-1 (:1035)
        if (lane == 0) c.out(nk, WF_K, ECNE_NO_LB, a.rank_b);
        c.ov.set(nk, d.r0, a.rank_b, true);
      }
      if (nuC > 0 && c.uniq(nk)) {  // (:1049-1067)
        scan_terms<G, SCAN_U(G)>(d, F, s2, s3, lane, [&](uint32_t w, uint32_t f) {
          if (w != nk && !(f & WF_U)) c.out(w, WF_U | WF_K);
        });
        nuC = 0;
        c.ov.all_unique = true;
      }
    }
  }

  // ---- Case 4a (:1078-1146) / 4b (:1148-1232) ---------------------------------------------
  if (rf & (RF_4A | RF_4B)) {
    uint32_t k1, k2;
    if (rf & RF_4A) {
      k1 = d.col[s2];
      k2 = d.col[s2 + 1];
    } else {
      uint32_t x0 = d.col[s2], x1 = d.col[s2 + 1], x2 = d.col[s2 + 2];
      k1 = (x0 == 1) ? x1 : x0;
      k2 = (x2 == 1) ? x1 : x2;
    }
    uint32_t l1, u1, l2, u2;
    c.bounds(k1, l1, u1);
    c.bounds(k2, l2, u2);
    // the `unique` fields agree here: Case 1 above already handled "exactly one non-unique"
    if (u1 != u2 || l1 != l2) {
      uint32_t mn = u1 < u2 ? u1 : u2, mx = l1 > l2 ? l1 : l2;
      bool go = (rf & RF_4A) || (mn == d.r1 && mx == d.r0);  // (:1196-1199)
      if (go) {
        if (u1 > mn || l1 < mx) {
          if (lane == 0) {
            c.out(k1, WF_K, mx, mn);
            if (rf & RF_4B) d.valsrc[k1] = VS_ONEZERO;
          }
          c.ov.set(k1, mx, mn, true);
        }
        if (u2 > mn || l2 < mx) {
          if (lane == 0) {
            c.out(k2, WF_K, mx, mn);
            if (rf & RF_4B) d.valsrc[k2] = VS_ONEZERO;
          }
          c.ov.set(k2, mx, mn, true);
        }
      }
    }
  }

  ECNE_PSTAMP(2);  // cases 1-4
  if (nuC == 0) {
    ECNE_PFLUSH(2);
    return plain;
  }
  // ---- Case 5 (:1235-1298) ----------------------------------------------------------------
  bool local_k = (c.ov.n > 0 && c.ov.k[0]) || (c.ov.n > 1 && c.ov.k[1]);
  // Case 5 is a pure function of the row's non-unique set (uniqueness is monotone, so its size
  // identifies it), their is_known bits (all set here) and bounds: skip it when it already failed
  // on exactly this state.
  // (the memo keeps 12 bits of the count and 20 of the epoch: a row with more non-unique keys than that, or a
  // solve with more bound-tightening rounds, is simply evaluated again — never skipped on an aliased signature)
  const bool memo = nuC <= 0xfffu && bepoch <= 0xfffffu;
  const uint32_t sig = (bepoch << 12) | (nuC & 0xfffu);
  if ((kmiss == 0 || local_k) && (!memo || d.c5sig[row] != sig)) {
    if (case5<G>(c)) {
      ECNE_PSTAMP(3);
      emit_unique_all<G>(c, F, s2, s3, lane);
      ECNE_PSTAMP(5);
      ECNE_PFLUSH(5);
      return plain;
    }
    if (lane == 0 && memo && kmiss == 0 && !local_k) d.c5sig[row] = sig;
  }
  ECNE_PSTAMP(3);  // case 5
  // ---- Case 6 (:1304-1348) ----------------------------------------------------------------
  if (abzmiss == 0) {
    uint32_t lo = 0xffffffffu, hi = 0;
    scan_terms<G, SCAN_U(G)>(d, F, s2, s3, lane, [&](uint32_t w, uint32_t f) {
      if (!(f & WF_U)) {
        uint32_t z = (uint32_t)d.abz[w];
        lo = z < lo ? z : lo;
        hi = z > hi ? z : hi;
      }
    });
    lo = Grp<G>::min(lo);
    hi = Grp<G>::max(hi);
    ECNE_PSTAMP(4);  // case 6 scan
    if (lo == hi) {
      emit_unique_all<G>(c, F, s2, s3, lane);
      ECNE_PSTAMP(5);
      ECNE_PFLUSH(5);
      return plain;
    }
  }
  ECNE_PFLUSH(4);
  return false;
}

}  // namespace ecne
