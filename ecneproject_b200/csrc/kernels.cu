// kernels.cu — per-round kernels of the propagation fixpoint.
//
//   k_p1_loop     the queue loop (:805-1349) as Jacobi sweeps inside ONE persistent cooperative
//                 launch: replay last round's records -> sweep all rows -> grid barrier -> repeat
//                 until a round produces no record (the device-wide changed flag)
//   k_p0          special constraints (:718-800), in list order inside one block
//   k_p2_*        linear-system sweep (:1357-1417)
//   k_p3_*        ABZ tagging sweep (:1425-1483), lowest row wins exactly like the reference
//   k_p4          IsZero pair sweep (:1492-1550)
//   k_replay      bring the other state buffer up to date after a phase kernel
//   k_finalize    verdict (:1558-1597) + packed bitmaps for the D2H
#include "engine_host.h"
#include "sweep.cuh"

namespace ecne {

#define P1_THREADS 1024
#define P1_MIN_BLOCKS 1
#define P1_MAX_KS 6     // 6 rows x 32 B x 1024 threads = 192 KB of the 227 KB shared memory
#define CHG_WORDS 128   // 4096-bit changed-wire filter
#define CHG_MAX 512u    // rounds with more records than this are followed by a dense round

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
  unsigned int i = atomicAdd(&d.st->p2_cand, 1u);
  d.p2_key[i] = mix64(hs ^ (hx * 0x9e3779b97f4a7c15ULL) ^ k);
  d.p2_row[i] = row;
}

// P2 qualification of one row through the CSR (generic path): every non-unique wire appears in C
// only (:1364-1385).  k == 1 is decided on the spot; k >= 2 rows become (set-hash, row) candidates.
template <int G>
__device__ __noinline__ void p2_scan_row(const Dev& d, int rbuf, uint32_t row) {
  const uint32_t lane = Grp<G>::lane();
  const uint8_t* F = d.F[rbuf];
  const uint32_t s0 = d.seg[3 * row], s2 = d.seg[3 * row + 2], s3 = d.seg[3 * row + 3];
  uint32_t bad = 0;
  for (uint32_t t = s0 + lane; t < s2; t += G) bad |= (ld_flag(F, d.col[t]) & WF_U) ? 0u : 1u;
  if (Grp<G>::any(bad != 0)) return;  // a non-unique wire in A or B (:1366, :1378)
  uint32_t k = 0, w1 = 0;
  unsigned long long hs = 0, hx = 0;
  for (uint32_t t = s2 + lane; t < s3; t += G) {
    uint32_t w = d.col[t];
    if (!(ld_flag(F, w) & WF_U)) {
      ++k;
      w1 = w;
      unsigned long long m = mix64(w);
      hs += m;
      hx ^= mix64(m + 0x9e3779b97f4a7c15ULL);
    }
  }
  if (G > 1) {
    k = Grp<G>::sum(k);
    w1 = Grp<G>::max(w1);
    for (int o = 16; o > 0; o >>= 1) {
      hs += __shfl_xor_sync(0xffffffffu, hs, o);
      hx ^= __shfl_xor_sync(0xffffffffu, hx, o);
    }
  }
  if (k == 0 || lane != 0) return;
  if (k == 1) {  // 1x1 "matrix": the stored coefficient is non-zero (:1402)
    emit(d, 1, 0, w1, WF_U | WF_K);
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

// one bit of a 32-bit bloom filter per wire (multiplicative hash: neighbouring ids spread out)
__device__ __forceinline__ uint32_t wire_bloom(uint32_t w) { return 1u << ((w * 0x9E3779B1u) >> 27); }

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
__device__ __forceinline__ bool eval_inline(const Dev& d, int rbuf, int wbuf, int list, uint32_t row,
                                            const InlineRow& r, const uint32_t* f, uint32_t bepoch) {
  const uint32_t rf = r.rf;
  if (!(rf & RF_FAST)) {
    if (rf & RF_LONG) return true;  // swept by a whole warp instead
    return eval_row<1>(d, rbuf, wbuf, list, row, bepoch);
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
    emit(d, wbuf, list, wC, WF_U | WF_K);
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
        emit(d, wbuf, list, v, WF_K, ECNE_NO_LB, (rf & RF_2A_BOOL) ? d.r1 : ECNE_NO_UB);
        d.valsrc[v] = VS_2A | d.aux[row].val_idx;
        d.solved[row] |= 1;
      }
    }
    return true;  // fired (equation_solved) or v* already known: never fires again
  }
  // Case 2b (:949-988) on a row with no other pattern: x = t once and for all
  if (rf & RF_2B) {
    const RowAux a = d.aux[row];
    emit(d, wbuf, list, a.w1, WF_U | WF_K, a.rank_a, a.rank_a);
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
        l1 = ld_u32(d.LBR[rbuf], k1);
        u1 = ld_u32(d.UBR[rbuf], k1);
      }
      if (f[1] & WF_BND) {
        l2 = ld_u32(d.LBR[rbuf], k2);
        u2 = ld_u32(d.UBR[rbuf], k2);
      }
      if (u1 != u2 || l1 != l2) {
        const uint32_t mn = u1 < u2 ? u1 : u2, mx = l1 > l2 ? l1 : l2;
        if (u1 > mn || l1 < mx) {
          emit(d, wbuf, list, k1, WF_K, mx, mn);
          K1 = true;
        }
        if (u2 > mn || l2 < mx) {
          emit(d, wbuf, list, k2, WF_K, mx, mn);
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
        emit(d, wbuf, list, k1, WF_U | WF_K);
        emit(d, wbuf, list, k2, WF_U | WF_K);
      }
    }
    return false;  // bounds may still have to travel through this row later
  }
  if (nuC == 0) return true;  // no non-unique wire left in C: Cases 1/5/6 can never fire again
  // Cases 5 / 6 (:1235-1348) need every non-unique key known resp. ABZ-tagged: rare, generic path
  if ((rf & RF_LINEAR) && (kmiss == 0 || abzmiss == 0)) eval_row<1>(d, rbuf, wbuf, list, row, bepoch);
  return false;
}

// The queue loop (:805-1349) as Jacobi sweeps in one persistent cooperative launch.
//   * static row -> thread mapping: thread t owns rows row_lo + t + k*nthreads; a 64-bit register
//     mask tracks which of them can still fire, so resolved rows cost nothing in later rounds;
//   * fast path: one coalesced 32-byte RowRec load + the state-byte gathers of its <= 6 wires, two
//     rows in flight per thread;
//   * long rows get a warp each, spread over all blocks, first in the round;
//   * one grid barrier per round; the round's record count (the device-wide changed flag) comes back
//     with the barrier release;
//   * tail: the P2 candidate scan (:1357-1388) over the rows that are still live.
__global__ void __launch_bounds__(P1_THREADS, P1_MIN_BLOCKS)
    k_p1_loop(Dev d, int rbuf0, unsigned int max_rounds, int ks, int lc_words) {
  // Shared memory, resident for the whole launch:
  //   sm_rec   row records of the first `ks` rows of every thread: sm_rec[(2*k + h) * blockDim + thread]
  //            is half h of the thread's k-th record
  //   sm_chg   4096-bit hash set of the wires changed by the previous round (the frontier filter)
  //   sm_lcol  column lists of this block's long rows (as many as fit in lc_words)
  extern __shared__ uint4 sm_rec[];
  uint32_t* sm_chg = reinterpret_cast<uint32_t*>(sm_rec + (size_t)(ks > 0 ? ks : 1) * 2 * P1_THREADS);
  uint32_t* sm_lcol = sm_chg + CHG_WORDS;
  __shared__ int s_loff[32];       // smem offset of warp j's long row (-1: not cached)
  __shared__ uint32_t s_llen[32];  // its length
  __shared__ uint32_t s_chg32;     // 32-bit bloom of the changed wires (first-level filter)
  unsigned int epoch = 0;
  int rbuf = rbuf0;
  unsigned int list = 0;    // list written this round
  unsigned int prev_n = 0;  // records of the previous round (to replay)
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nthreads = gridDim.x * blockDim.x;
  const uint32_t rows = d.row_hi - d.row_lo;
  const uint32_t per_thread = (rows + nthreads - 1) / nthreads;
  const uint32_t kmask = per_thread < 64 ? per_thread : 64;
  const uint32_t warp_in_block = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
  unsigned int round = 0;
  unsigned int xe = d.world > 1 ? *d.xepoch : 0u;  // cross-GPU epoch (same on every rank)
  unsigned long long evals = 0, changed = 0, ruleevals = 0;
  uint32_t bepoch = d.st->bepoch;  // rounds so far that tightened a bound (same value in every thread)

  // rows latched by P4 / 2a since the last launch leave the mask; the others are staged in smem
  unsigned long long live = d.live[tid];
  uint32_t sig[P1_MAX_KS];  // 32-bit bloom of the wires of the thread's k-th row (registers)
#pragma unroll
  for (int k = 0; k < P1_MAX_KS; ++k) {
    sig[k] = 0xffffffffu;
    if (k < ks && ((live >> k) & 1ULL)) {
      uint32_t r = tid + (uint32_t)k * nthreads;
      if (r >= rows || (d.solved[d.row_lo + r] & 1)) {
        live &= ~(1ULL << k);
      } else {
        const uint4* rp = reinterpret_cast<const uint4*>(d.rec + d.row_lo + r);
        const uint4 q0 = __ldg(rp), q1 = __ldg(rp + 1);
        sm_rec[(2 * k) * blockDim.x + threadIdx.x] = q0;
        sm_rec[(2 * k + 1) * blockDim.x + threadIdx.x] = q1;
        if (q0.y & 0x10000u)  // inline: unused slots hold the constant wire 1, which never changes
          sig[k] = wire_bloom(q0.z) | wire_bloom(q0.w) | wire_bloom(q1.x) | wire_bloom(q1.y) |
                   wire_bloom(q1.z) | wire_bloom(q1.w);
      }
    }
  }
  for (unsigned long long m = live & (~0ULL << ks); m;) {  // ks <= 6
    int k = __ffsll((long long)m) - 1;
    m &= m - 1;
    uint32_t r = tid + (uint32_t)k * nthreads;
    if (r >= rows || (d.solved[d.row_lo + r] & 1)) live &= ~(1ULL << k);
  }

  // stage the column lists of this block's first 32 long rows (warp j owns long row b + j*grid)
  if (threadIdx.x == 0) {
    int off = 0;
    for (uint32_t j = 0; j < 32; ++j) {
      uint32_t i = blockIdx.x + j * gridDim.x;
      s_loff[j] = -1;
      if (i < d.n_long && !d.long_done[i]) {
        uint32_t row = d.long_rows[i];
        int len = (int)(d.seg[3 * row + 3] - d.seg[3 * row]);
        if (off + len <= lc_words) {
          s_loff[j] = off;
          s_llen[j] = (uint32_t)len;
          off += len;
        }
      }
    }
  }
  __syncthreads();
  if (warp_in_block < 32 && s_loff[warp_in_block] >= 0) {
    uint32_t row = d.long_rows[blockIdx.x + warp_in_block * gridDim.x];
    uint32_t s0 = d.seg[3 * row], s3 = d.seg[3 * row + 3];
    for (uint32_t t = s0 + (threadIdx.x & 31); t < s3; t += 32) sm_lcol[s_loff[warp_in_block] + (t - s0)] = d.col[t];
  }
  bool filtered = false;  // this round only evaluates rows that touch a wire changed last round
  __syncthreads();
#ifdef ECNE_PROFILE
  long long pf[6] = {0, 0, 0, 0, 0, 0};  // intra-block wait, grid wait, long rows, replay, sweep, -
#define PROF_T(x) long long x = clock64()
#define PROF_ACC(i, a, b) pf[i] += (b) - (a)
#else
#define PROF_T(x)
#define PROF_ACC(i, a, b)
#endif
  while (true) {
    const int wbuf = rbuf ^ 1;
    const uint8_t* F = d.F[rbuf];
    PROF_T(t0);
    // (a) long rows first (their latency overlaps the rest): block b owns long rows b, b+grid, ...
    for (uint32_t i = blockIdx.x + warp_in_block * gridDim.x, it = 0; i < d.n_long;
         i += warps_per_block * gridDim.x, ++it) {
      if (d.long_done[i]) continue;
      uint32_t row = d.long_rows[i];
      if (row >= d.row_lo && row < d.row_hi) {
        if (filtered) {  // does the row touch a changed wire?
          const bool cached = it == 0 && s_loff[warp_in_block] >= 0;
          const uint32_t s0 = cached ? 0u : d.seg[3 * row];
          const uint32_t len = cached ? s_llen[warp_in_block] : d.seg[3 * row + 3] - s0;
          const uint32_t* cl = cached ? sm_lcol + s_loff[warp_in_block] : d.col + s0;
          bool hit = false;
          for (uint32_t t = threadIdx.x & 31; t < len; t += 32) {
            uint32_t w = cl[t];
            hit |= (sm_chg[(w >> 5) & (CHG_WORDS - 1)] >> (w & 31)) & 1u;
          }
          if (!__any_sync(0xffffffffu, hit)) continue;
        }
        bool done = eval_row<32>(d, rbuf, wbuf, (int)list, row, bepoch);
        if ((threadIdx.x & 31) == 0) {
          evals += 1;
          if (done) d.long_done[i] = 1;
        }
      }
    }
    PROF_T(t1);
    PROF_ACC(2, t0, t1);
    // (b) replay the previous round's records into the buffer written this round
    if (prev_n) {
      const Rec* pr = d.recs[(list + 2) % 3];
      // record i goes to block i % grid so that a short list still spreads over every SM
      for (uint32_t j = threadIdx.x; blockIdx.x + j * gridDim.x < prev_n; j += blockDim.x) {
        Rec r = pr[blockIdx.x + j * gridDim.x];
        apply_update(d, wbuf, r.wire, r.bits, r.lbr, r.ubr);
      }
    }
    PROF_T(t2);
    PROF_ACC(3, t1, t2);
    // (c) sweep the rows this thread still owns, two in flight.  In a filtered round a row is first
    // tested with one AND of its register bloom against the round's changed-set bloom.
    evals += __popcll(live);
    unsigned long long todo = live;
    if (filtered) {
      const uint32_t chg32 = s_chg32;
      unsigned long long cand = ~0ULL << ks;  // rows beyond the smem-resident ones are always taken
#pragma unroll
      for (int k = 0; k < P1_MAX_KS; ++k)
        if (sig[k] & chg32) cand |= 1ULL << k;
      todo &= cand;
    }
    for (unsigned long long m = todo; m;) {
      const int k0 = __ffsll((long long)m) - 1;
      m &= m - 1;
      const int k1 = m ? __ffsll((long long)m) - 1 : -1;
      if (k1 >= 0) m &= m - 1;
      const int kb = k1 >= 0 ? k1 : k0;
      const uint32_t row0 = d.row_lo + tid + (uint32_t)k0 * nthreads;
      const uint32_t row1 = d.row_lo + tid + (uint32_t)kb * nthreads;
      InlineRow r0, r1;
      if (k0 < ks)
        unpack_row(sm_rec[(2 * k0) * blockDim.x + threadIdx.x], sm_rec[(2 * k0 + 1) * blockDim.x + threadIdx.x], r0);
      else
        load_row(d, row0, r0);
      if (kb < ks)
        unpack_row(sm_rec[(2 * kb) * blockDim.x + threadIdx.x], sm_rec[(2 * kb + 1) * blockDim.x + threadIdx.x], r1);
      else
        load_row(d, row1, r1);
      bool go0 = true, go1 = k1 >= 0;
      if (filtered) {
        // Exact shortcut: a row none of whose wires changed state last round evaluates exactly as it
        // did last round, i.e. to nothing new.  (Non-inline rows keep their wires in the CSR: they
        // are few and simply always re-evaluated.)
        if (r0.meta & 0x10000u) {
          uint32_t h = 0;
#pragma unroll
          for (int j = 0; j < ROWREC_INLINE; ++j) h |= (sm_chg[(r0.c[j] >> 5) & (CHG_WORDS - 1)] >> (r0.c[j] & 31)) & 1u;
          go0 = h != 0;
        }
        if (go1 && (r1.meta & 0x10000u)) {
          uint32_t h = 0;
#pragma unroll
          for (int j = 0; j < ROWREC_INLINE; ++j) h |= (sm_chg[(r1.c[j] >> 5) & (CHG_WORDS - 1)] >> (r1.c[j] & 31)) & 1u;
          go1 = h != 0;
        }
        if (!go0 && !go1) continue;
      }
      uint32_t f0[ROWREC_INLINE], f1[ROWREC_INLINE];
      if (go0) gather_row(F, r0, f0);
      if (go1) gather_row(F, r1, f1);
      if (go0) {
        ruleevals += 1;
        if (eval_inline(d, rbuf, wbuf, (int)list, row0, r0, f0, bepoch)) live &= ~(1ULL << k0);
      }
      if (go1) {
        ruleevals += 1;
        if (eval_inline(d, rbuf, wbuf, (int)list, row1, r1, f1, bepoch)) live &= ~(1ULL << k1);
      }
    }
    // rows beyond the 64 tracked per thread (only for problems far larger than the machine)
    for (uint32_t k = kmask; k < per_thread; ++k) {
      uint32_t r = tid + k * nthreads;
      if (r < rows) {
        uint32_t row = d.row_lo + r;
        if (!(d.rflags[row] & RF_LONG)) eval_row<1>(d, rbuf, wbuf, (int)list, row, bepoch);
        evals += 1;
      }
    }
    PROF_T(t3);
    PROF_ACC(4, t2, t3);
#ifdef ECNE_PROFILE
    if (threadIdx.x == 0 && round < 40) {
      // per block: long-row, replay, sweep cycles of this round
      unsigned long long* q = d.prof + 20000 + ((size_t)round * gridDim.x + blockIdx.x) * 4;
      q[0] = (unsigned long long)(t1 - t0);
      q[1] = (unsigned long long)(t2 - t1);
      q[2] = (unsigned long long)(t3 - t2);
    }
    __syncthreads();
    if (threadIdx.x == 0 && round < 40)
      d.prof[20000 + ((size_t)round * gridDim.x + blockIdx.x) * 4 + 3] = (unsigned long long)(clock64() - t0);
    xe += 1;
    unsigned int n = grid_barrier(d.barrier, epoch, d.rec_count + list, d.bnd_flag + list,
                                  d.world > 1 ? &d : nullptr, list, xe, pf);
#else
    xe += 1;
    unsigned int n = grid_barrier(d.barrier, epoch, d.rec_count + list, d.bnd_flag + list,
                                  d.world > 1 ? &d : nullptr, list, xe);
#endif
    bepoch += n >> 31;
    n &= 0x7fffffffu;
    unsigned int n_own = n;
    if (d.world > 1) {
      // pull the peers' records of this round over NVLink and apply them to BOTH local buffers (the
      // buffer read next round must already contain them), then a local barrier
      n_own = __ldcg(d.xcnt + list * ECNE_MAX_WORLD + d.rank);
      for (int h = 0; h < d.world; ++h) {
        if (h == d.rank) continue;
        const unsigned int nh = __ldcg(d.xcnt + list * ECNE_MAX_WORLD + h);
        const Rec* pr = d.xrecs[h][list];
        for (uint32_t j = tid; j < nh && j < d.rec_cap; j += nthreads) {
          Rec r = ld_peer_rec(pr + j);
          apply_update(d, 0, r.wire, r.bits, r.lbr, r.ubr);
          apply_update(d, 1, r.wire, r.bits, r.lbr, r.ubr);
        }
      }
      grid_barrier(d.barrier, epoch, nullptr);
    }
#ifdef ECNE_PROFILE
    if (tid == 0) {
      unsigned long long slot = atomicAdd(&d.prof[7], 1ULL);  // global round index across launches
      if (slot < 4000) {
        d.prof[2048 + 4 * slot + 0] = (unsigned long long)(clock64() - t0);
        d.prof[2048 + 4 * slot + 1] = n;
        d.prof[2048 + 4 * slot + 2] = (unsigned long long)(t3 - t2);
        d.prof[2048 + 4 * slot + 3] = (unsigned long long)(t1 - t0);
      }
    }
#endif
    round += 1;
    changed += n;
    if (tid == 0) {  // replayed this round; next written in two rounds
      d.rec_count[(list + 2) % 3] = 0;
      d.bnd_flag[(list + 2) % 3] = 0;
    }
    if (n_own > d.rec_cap) n_own = d.rec_cap;
    if (n == 0) break;  // W already holds every earlier record: both buffers are complete
    if (round >= max_rounds) {
      if (tid == 0) raise(d, ECNE_E_NOCONVERGE);
      const Rec* pr = d.recs[list];
      for (uint32_t i = tid; i < n_own; i += nthreads) {
        Rec r = pr[i];
        apply_update(d, rbuf, r.wire, r.bits, r.lbr, r.ubr);
      }
      break;
    }
    // frontier filter for the next round: hash set of the wires this round changed (on any rank)
    filtered = n <= CHG_MAX;
    if (filtered) {
      for (uint32_t j = threadIdx.x; j < CHG_WORDS; j += blockDim.x) sm_chg[j] = 0;
      if (threadIdx.x == 0) s_chg32 = 0;
      __syncthreads();
      for (int h = 0; h < (d.world > 1 ? d.world : 1); ++h) {
        const Rec* cr = d.world > 1 ? d.xrecs[h][list] : d.recs[list];
        const unsigned int nh = d.world > 1 ? __ldcg(d.xcnt + list * ECNE_MAX_WORLD + h) : n;
        for (uint32_t j = threadIdx.x; j < nh && j < d.rec_cap; j += blockDim.x) {
          uint32_t w = ld_peer_rec(cr + j).wire;
          atomicOr(&sm_chg[(w >> 5) & (CHG_WORDS - 1)], 1u << (w & 31));
          atomicOr(&s_chg32, wire_bloom(w));
        }
      }
      __syncthreads();
    }
    prev_n = n_own;
    rbuf = wbuf;
    list = (list + 1) % 3;
  }
  d.live[tid] = live;
#ifdef ECNE_PROFILE
  if (threadIdx.x == 0)
    for (int i = 0; i < 6; ++i) d.prof[blockIdx.x * 8 + i] += (unsigned long long)pf[i];
  if (threadIdx.x == 0) d.prof[blockIdx.x * 8 + 6] += round;
#endif
  // ---- tail: P2 candidate scan over the rows that are still live (state is at the P1 fixpoint) ----
  grid_barrier(d.barrier, epoch, nullptr);  // the list counters are all zero and visible from here on
  if (d.world == 1) {  // sharded runs scan every row on every rank instead (k_p2_scan_all)
    const uint8_t* F = d.F[0];
    for (unsigned long long m = live; m;) {
      const int k = __ffsll((long long)m) - 1;
      m &= m - 1;
      const uint32_t row = d.row_lo + tid + (uint32_t)k * nthreads;
      InlineRow r;
      load_row(d, row, r);
      if (r.rf & RF_LONG) continue;
      if (!(r.meta & 0x10000u)) {  // not inline
        p2_scan_row<1>(d, 0, row);
        continue;
      }
      const uint32_t nAB = r.meta & 0xffu, nT = nAB + ((r.meta >> 8) & 0xffu);
      uint32_t kk = 0, w1 = 0;
      bool bad = false;
      unsigned long long hs = 0, hx = 0;
#pragma unroll
      for (int j = 0; j < ROWREC_INLINE; ++j) {
        if ((uint32_t)j < nT && !(ld_flag(F, r.c[j]) & WF_U)) {
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
      if (bad || kk == 0) continue;
      if (kk == 1)
        emit(d, 1, 0, w1, WF_U | WF_K);
      else
        p2_candidate(d, row, hs, hx, kk);
    }
    for (uint32_t k = kmask; k < per_thread; ++k) {
      uint32_t r = tid + k * nthreads;
      if (r < rows && !(d.rflags[d.row_lo + r] & RF_LONG) && !(d.solved[d.row_lo + r] & 1))
        p2_scan_row<1>(d, 0, d.row_lo + r);
    }
    for (uint32_t i = blockIdx.x + warp_in_block * gridDim.x; i < d.n_long; i += warps_per_block * gridDim.x) {
      uint32_t row = d.long_rows[i];
      if (!d.long_done[i] && row >= d.row_lo && row < d.row_hi) p2_scan_row<32>(d, 0, row);
    }
  }
  // statistics: one atomic per warp
  evals = evals + __shfl_xor_sync(0xffffffffu, evals, 16);
  evals = evals + __shfl_xor_sync(0xffffffffu, evals, 8);
  evals = evals + __shfl_xor_sync(0xffffffffu, evals, 4);
  evals = evals + __shfl_xor_sync(0xffffffffu, evals, 2);
  evals = evals + __shfl_xor_sync(0xffffffffu, evals, 1);
  if ((tid & 31) == 0) atomicAdd(&d.st->evals, evals);
  ruleevals += __shfl_xor_sync(0xffffffffu, ruleevals, 16);
  ruleevals += __shfl_xor_sync(0xffffffffu, ruleevals, 8);
  ruleevals += __shfl_xor_sync(0xffffffffu, ruleevals, 4);
  ruleevals += __shfl_xor_sync(0xffffffffu, ruleevals, 2);
  ruleevals += __shfl_xor_sync(0xffffffffu, ruleevals, 1);
  if ((tid & 31) == 0) atomicAdd(&d.st->rule_evals, ruleevals);
  if (tid == 0) {
    d.st->rounds += round;
    d.st->changed += changed;
    d.st->bepoch = bepoch;
    if (d.world > 1) *d.xepoch = xe;
  }
}

// P2 candidate scan over ALL rows (multi-GPU runs: every rank computes the same candidates)
__global__ void k_p2_scan_all(Dev d) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid < d.N) {
    if (!(d.rflags[tid] & RF_LONG) && !(d.solved[tid] & 1)) p2_scan_row<1>(d, 0, tid);
  }
  const uint32_t warp = tid >> 5;
  if (warp < d.n_long) {
    uint32_t row = d.long_rows[warp];
    if (!(d.solved[row] & 1)) p2_scan_row<32>(d, 0, row);
  }
}

// After a phase kernel wrote its updates to buffer `buf ^ 1` and logged them in list 0: apply them
// to `buf` too and clear the list.
__global__ void k_replay(Dev d, int buf) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nthreads = gridDim.x * blockDim.x;
  unsigned int n = d.rec_count[0];
  if (n > d.rec_cap) n = d.rec_cap;
  for (uint32_t i = tid; i < n; i += nthreads) {
    Rec r = d.recs[0][i];
    apply_update(d, buf, r.wire, r.bits, r.lbr, r.ubr);
  }
}
__global__ void k_replay_done(Dev d) {
  d.st->changed += d.rec_count[0];
  d.rec_count[0] = 0;
  d.rec_count[1] = 0;
  d.rec_count[2] = 0;
  d.bnd_flag[0] = 0;
}

// ---- P0 / P0' (:718-800): in-order, one block; writes both buffers in place ------------------
__global__ void k_p0(Dev d) {
  __shared__ int s_ok;
  for (uint32_t s = 0; s < d.n_specials; ++s) {
    if (d.sp_solved[s]) continue;  // uniform: read by all threads from global, written below + barrier
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    for (uint32_t k = d.sp_in_ptr[s] + threadIdx.x; k < d.sp_in_ptr[s + 1]; k += blockDim.x)
      if (!(ld_flag(d.F[0], d.sp_in[k]) & WF_U)) s_ok = 0;
    __syncthreads();
    int ok = s_ok;
    __syncthreads();
    if (!ok) continue;
    for (uint32_t k = d.sp_out_ptr[s] + threadIdx.x; k < d.sp_out_ptr[s + 1]; k += blockDim.x) {
      uint32_t w = d.sp_out[k];
      or_flag(d.F[0], w, WF_U | WF_K);
      or_flag(d.F[1], w, WF_U | WF_K);
    }
    if (threadIdx.x == 0) {
      d.sp_solved[s] = 1;
      d.st->changed += 1;  // successful_steps += 1 (:731)
    }
    __threadfence();
    __syncthreads();
  }
  // P0': every (BigMultModP, BigLessThan) pair marks the BigLessThan's first three inputs (:750-800)
  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < d.n_specials; ++i) {
      if (d.sp_kind[i] != ECNE_SPECIAL_BIGMULTMODP) continue;
      for (uint32_t j = 0; j < d.n_specials; ++j) {
        if (d.sp_kind[j] != ECNE_SPECIAL_BIGLESSTHAN) continue;
        if (!d.secp_solve) {  // `dsu` only exists under secp_solve (:634-636, :762)
          raise(d, ECNE_E_NODSU);
          return;
        }
        uint32_t ni = d.sp_in_ptr[i + 1] - d.sp_in_ptr[i], nj = d.sp_in_ptr[j + 1] - d.sp_in_ptr[j];
        if (ni < 9 || nj < 6) {
          raise(d, ECNE_E_BOUNDS);
          return;
        }
        for (uint32_t k = 0; k < 3; ++k) {
          uint32_t w = d.sp_in[d.sp_in_ptr[j] + k];
          uint32_t nw = or_flag(d.F[0], w, WF_U | WF_K);
          or_flag(d.F[1], w, WF_U | WF_K);
          if (nw & WF_U) d.st->changed += 1;  // not a successful_step in the reference, but it
                                              // re-enqueues rows; a state change keeps us looping
        }
      }
    }
  }
}

// ---- P2 (:1357-1417): grouping of the candidates found by the sweep kernel's tail -------------
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

// candidates sorted by (key, row): thread at a group start takes the first k rows in index order,
// checks the sets really are equal, and evaluates slow_det (:1389-1400) = sum over ODD
// permutations (Combinatorics.parity is 0 for even ones).
__global__ void k_p2_groups(Dev d, int rbuf, uint32_t n_cand, const unsigned long long* keys,
                            const uint32_t* rows) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cand) return;
  if (i > 0 && keys[i - 1] == keys[i]) return;  // not a group start
  const uint8_t* F = d.F[rbuf];
  uint32_t vars[ECNE_P2_KMAX], terms[ECNE_P2_KMAX][ECNE_P2_KMAX], v2[ECNE_P2_KMAX];
  uint32_t k = p2_unknowns(d, F, rows[i], vars, terms[0], ECNE_P2_KMAX);
  // are there k rows in this group?
  if (i + k > n_cand || keys[i + k - 1] != keys[i]) return;
  if (k > ECNE_P2_KMAX) {
    raise(d, ECNE_E_UNSUPPORTED);
    return;
  }
  for (uint32_t j = 1; j < k; ++j) {
    uint32_t kj = p2_unknowns(d, F, rows[i + j], v2, terms[j], ECNE_P2_KMAX);
    bool same = kj == k;
    for (uint32_t x = 0; same && x < k; ++x) same = v2[x] == vars[x];
    if (!same) {  // 64-bit set-hash collision: refuse to guess
      raise(d, ECNE_E_INTERNAL);
      return;
    }
  }
  // odd-permutation sum; Montgomery products of canonical inputs carry a uniform R^-(k-1) factor,
  // which does not change whether the sum is zero
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
  if (!fr::is_zero(res)) {
    for (uint32_t j = 0; j < k; ++j) emit(d, rbuf ^ 1, 0, vars[j], WF_U | WF_K);
  }
}

// ---- P3 (:1425-1483) ------------------------------------------------------------------------
__global__ void k_p3_claim(Dev d, int rbuf) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= d.N) return;
  uint32_t rf = d.rflags[row];
  if (!(rf & RF_P3)) return;
  const RowAux a = d.aux[row];
  if (ld_flag(d.F[rbuf], a.w3) & WF_U) return;  // unique_b (:1448)
  if (rf & RF_P3_DIVZ) {                         // divexact(-intercept, 0) (:1467)
    raise(d, ECNE_E_DIVZERO);
    return;
  }
  atomicMin(d.abz_claim + a.w3, ((unsigned long long)row << 32) | a.w4);
}
__global__ void k_p3_commit(Dev d, int rbuf) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= d.N) return;
  uint32_t rf = d.rflags[row];
  if (!(rf & RF_P3) || (rf & RF_P3_DIVZ)) return;
  const RowAux a = d.aux[row];
  if (ld_flag(d.F[rbuf], a.w3) & WF_U) return;
  unsigned long long cl = d.abz_claim[a.w3];
  if ((uint32_t)(cl >> 32) != row) return;  // a lower row tags this wire first
  d.abz_claim[a.w3] = ~0ULL;
  if (d.abz[a.w3] != -1) return;            // (:1469-1473)
  d.abz[a.w3] = (int32_t)a.w4;
  or_flag(d.F[0], a.w3, WF_K | WF_ABZ);
  or_flag(d.F[1], a.w3, WF_K | WF_ABZ);
  atomicAdd(&d.st->changed, 1ULL);
}

// ---- P4 (:1492-1550) ------------------------------------------------------------------------
__global__ void k_p4(Dev d, int rbuf) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= d.N) return;
  if (!(d.rflags[row] & RF_P4)) return;
  const uint8_t* F = d.F[rbuf];
  const uint32_t s0 = d.seg[3 * row], s1 = d.seg[3 * row + 1];
  for (uint32_t t = s0; t < s1; ++t)
    if (!(ld_flag(F, d.col[t]) & WF_U)) return;  // a_unique (:1502-1511)
  uint32_t vk = d.aux[row].w4;
  if (ld_flag(F, vk) & WF_U) return;
  emit(d, rbuf ^ 1, 0, vk, WF_U | WF_K);
  d.solved[row] |= 1;  // equation_solved[i], [i+1] (:1541-1542)
  d.solved[row + 1] |= 1;
}

// ---- finalisation ---------------------------------------------------------------------------
__global__ void k_pack(Dev d, int buf, unsigned long long* ubits, unsigned long long* kbits,
                       unsigned long long* counts) {
  // one thread per 64 wires
  uint32_t word = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t nwords = (d.V + 63) / 64;
  if (word >= nwords) return;
  unsigned long long u = 0, k = 0;
  uint32_t nu = 0, nnt = 0, nunt = 0;
  for (uint32_t b = 0; b < 64; ++b) {
    uint32_t w = word * 64 + b + 1;
    if (w > d.V) break;
    uint32_t f = d.F[buf][w];
    if (f & WF_U) {
      u |= 1ULL << b;
      ++nu;
    }
    if (f & WF_K) k |= 1ULL << b;
    if (d.nontriv[w]) {
      ++nnt;
      if (f & WF_U) ++nunt;
    }
  }
  ubits[word] = u;
  kbits[word] = k;
  atomicAdd(counts + 0, (unsigned long long)nu);
  atomicAdd(counts + 1, (unsigned long long)nnt);
  atomicAdd(counts + 2, (unsigned long long)nunt);
}
__global__ void k_targets(Dev d, int buf, unsigned long long* counts) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.n_targets) return;
  if (d.F[buf][d.targets[i]] & WF_U) atomicAdd(counts + 3, 1ULL);
}
// materialise lb/ub/values/nvalues from ranks and value sources
__global__ void k_export(Dev d, int buf, fr::u256* lb, fr::u256* ub, uint8_t* nvalues,
                         fr::u256* values) {
  uint32_t w = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (w > d.V) return;
  if (lb) lb[w - 1] = d.table[d.LBR[buf][w]];
  if (ub) ub[w - 1] = d.table[d.UBR[buf][w]];
  if (nvalues || values) {
    uint32_t vsrc = d.valsrc[w];
    uint8_t n = 0;
    fr::u256 v0 = fr::make_u256(0, 0, 0, 0), v1 = v0;
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
    if (nvalues) nvalues[w - 1] = n;
    if (values) {
      values[2 * (w - 1)] = v0;
      values[2 * (w - 1) + 1] = v1;
    }
  }
}

// ---- state reset ----------------------------------------------------------------------------
__global__ void k_reset_wires(Dev d) {
  uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w > d.V + 3) return;
  d.F[0][w] = 0;
  d.F[1][w] = 0;
  d.B[0][w] = 0;
  d.B[1][w] = 0;
  if (w <= d.V) {
    d.LBR[0][w] = d.r0;
    d.LBR[1][w] = d.r0;
    d.UBR[0][w] = d.rpm1;
    d.UBR[1][w] = d.rpm1;
    d.abz[w] = -1;
    d.valsrc[w] = VS_NONE;
    d.abz_claim[w] = ~0ULL;
  }
}
__global__ void k_reset_live(Dev d, uint32_t nthreads) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nthreads) return;
  uint32_t rows = d.row_hi - d.row_lo;
  uint32_t per_thread = (rows + nthreads - 1) / nthreads;
  uint32_t k = per_thread < 64 ? per_thread : 64;
  d.live[t] = k >= 64 ? ~0ULL : ((1ULL << k) - 1ULL);
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
  k_reset_wires<<<blocks_for((uint64_t)d.V + 4, 256), 256, 0, s>>>(d);
  k_reset_live<<<blocks_for((uint64_t)grid * P1_THREADS, 256), 256, 0, s>>>(d, (uint32_t)grid * P1_THREADS);
  if (d.n_known) k_reset_known<<<blocks_for(d.n_known, 256), 256, 0, s>>>(d);
  cudaMemsetAsync(d.solved, 0, (size_t)d.N + 1, s);
  if (d.n_long) cudaMemsetAsync(d.long_done, 0, d.n_long, s);
  if (d.n_specials) cudaMemsetAsync(d.sp_solved, 0, d.n_specials, s);
  cudaMemsetAsync(d.rec_count, 0, 3 * sizeof(unsigned int), s);
  cudaMemsetAsync(d.bnd_flag, 0, 3 * sizeof(unsigned int), s);
  cudaMemsetAsync(d.c5sig, 0xff, (size_t)(d.N ? d.N : 1) * sizeof(uint32_t), s);
  cudaMemsetAsync(d.st, 0, sizeof(Status), s);
  return cudaGetLastError();
}

int p1_threads() { return P1_THREADS; }
int p1_grid_size(int device) {
  static int cached = 0;
  if (cached) return cached;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaFuncSetAttribute(k_p1_loop, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       P1_MAX_KS * 2 * (int)sizeof(uint4) * P1_THREADS + 28 * 1024);
  cached = sms;  // one 1024-thread block per SM (persistent, cooperative)
  return cached;
}

cudaError_t launch_p1(const Dev& d, int rbuf, unsigned int max_rounds, int grid, cudaStream_t s) {
  cudaMemsetAsync(d.barrier, 0, 64 * sizeof(unsigned int), s);
  Dev dd = d;
  int rb = rbuf;
  unsigned int mr = max_rounds;
  const uint32_t rows = d.row_hi - d.row_lo;
  const uint32_t nthreads = (uint32_t)grid * P1_THREADS;
  int ks = (int)((rows + nthreads - 1) / nthreads);
  if (ks > P1_MAX_KS) ks = P1_MAX_KS;
  const size_t rec_bytes = (size_t)(ks > 0 ? ks : 1) * 2 * sizeof(uint4) * P1_THREADS;
  const size_t max_smem = (size_t)P1_MAX_KS * 2 * sizeof(uint4) * P1_THREADS + 28 * 1024;  // 220 KB
  int lc_words = (int)((max_smem - rec_bytes - CHG_WORDS * 4) / 4);
  size_t smem = rec_bytes + CHG_WORDS * 4 + (size_t)lc_words * 4;
  void* args[] = {&dd, &rb, &mr, &ks, &lc_words};
  return cudaLaunchCooperativeKernel((void*)k_p1_loop, dim3(grid), dim3(P1_THREADS), args, smem, s);
}

void launch_replay(const Dev& d, int buf, cudaStream_t s) {
  k_replay<<<148, 256, 0, s>>>(d, buf);
  k_replay_done<<<1, 1, 0, s>>>(d);
}

void launch_p2_scan_all(const Dev& d, cudaStream_t s) {
  uint64_t n = d.N > (uint64_t)d.n_long * 32 ? d.N : (uint64_t)d.n_long * 32;
  if (n) k_p2_scan_all<<<blocks_for(n, 256), 256, 0, s>>>(d);
}
void launch_p0(const Dev& d, cudaStream_t s) {
  if (d.n_specials) k_p0<<<1, 256, 0, s>>>(d);
}
void launch_p2_groups(const Dev& d, int rbuf, uint32_t n_cand, const unsigned long long* keys,
                      const uint32_t* rows, cudaStream_t s) {
  if (n_cand) k_p2_groups<<<blocks_for(n_cand, 128), 128, 0, s>>>(d, rbuf, n_cand, keys, rows);
}
void launch_p3(const Dev& d, int rbuf, cudaStream_t s) {
  if (!d.N) return;
  k_p3_claim<<<blocks_for(d.N, 256), 256, 0, s>>>(d, rbuf);
  k_p3_commit<<<blocks_for(d.N, 256), 256, 0, s>>>(d, rbuf);
}
void launch_p4(const Dev& d, int rbuf, cudaStream_t s) {
  if (d.N) k_p4<<<blocks_for(d.N, 256), 256, 0, s>>>(d, rbuf);
}
void launch_finalize(const Dev& d, int buf, unsigned long long* ubits, unsigned long long* kbits,
                     unsigned long long* counts, cudaStream_t s) {
  cudaMemsetAsync(counts, 0, 4 * sizeof(unsigned long long), s);
  k_pack<<<blocks_for((d.V + 63) / 64, 128), 128, 0, s>>>(d, buf, ubits, kbits, counts);
  if (d.n_targets) k_targets<<<blocks_for(d.n_targets, 128), 128, 0, s>>>(d, buf, counts);
}
void launch_export(const Dev& d, int buf, fr::u256* lb, fr::u256* ub, uint8_t* nvalues,
                   fr::u256* values, cudaStream_t s) {
  k_export<<<blocks_for(d.V, 256), 256, 0, s>>>(d, buf, lb, ub, nvalues, values);
}

}  // namespace ecne
