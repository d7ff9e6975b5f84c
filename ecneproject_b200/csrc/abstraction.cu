// abstraction.cu — abstraction() (/root/reference/src/R1CSConstraintSolver.jl:237-395, helpers :205-235) on the
// device, SURVEY.md §8f-1: the unreduced system is uploaded ONCE, every trusted circuit is matched against it by
// kernels, the kept rows are compacted on the device straight into the arrays the classification reads — the reduced
// system never crosses PCIe and no host core touches a constraint row.
//
//   k_abs_row_hash     a warp per row: order-free hash of each form's non-zero coefficients (:228-235 hashes the three
//                      sorted lists; only equality is ever used, :262)
//   k_abs_candidates   a thread per start row: the first n-1 row hashes line up (:259-270)
//   k_abs_terms        a thread per (candidate, window term): coefficient multisets per form (checkNonZeroValues,
//                      :205-226, as counts against the trusted circuit's sorted lists) and the per-wire appearance
//                      signature (:305-310) accumulated in an open-addressing table per candidate
//   k_abs_counts / k_abs_wires / k_abs_verify / k_abs_pop
//                      every multiset count matches, every wire's signature is EXACTLY the signature of one class of
//                      trusted wires (hash lookup, then term-by-term comparison), every class has as many wires as in
//                      the trusted circuit (:334-347: the two signature-sorted lists are pairwise equal)
//   k_abs_collect      the window's wires in the classes of the trusted circuit's inputs / outputs (:375-384)
//   k_abs_compact_*    kept rows -> the reduced system (:368-388), on the device
//
// The walk over the matches (:368-388, including its stall after an overlapping match, :370) is a loop over a few
// dozen window starts and stays on the host.  Wires with IDENTICAL signatures are paired in wire-id order, the same
// (reference-unpinned, Julia Dict order) tie rule as csrc/host_r1cs.cpp and oracle/abstraction_ref.py.
#include <algorithm>
#include <chrono>
#include <mutex>
#include <deque>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "engine_host.h"

namespace ecne {

#define CKE(x)                                                                     \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      err = std::string(#x) + ": " + cudaGetErrorString(e_);                       \
      return ECNE_E_CUDA;                                                          \
    }                                                                              \
  } while (0)

__host__ __device__ __forceinline__ unsigned long long abs_mix64(unsigned long long x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}
// contribution of one (slot, coefficient) appearance to a wire's signature hash; a signature's slots are distinct,
// so the SUM of the contributions is order-free and needs no per-wire sort
template <class T>
__host__ __device__ __forceinline__ unsigned long long sig_term(unsigned int slot, const T* c, unsigned long long seed) {
  unsigned long long h = abs_mix64(seed ^ (0x9e3779b97f4a7c15ULL * (slot + 1ULL)));
  h = abs_mix64(h ^ c[0]);
  h = abs_mix64(h ^ c[1]);
  h = abs_mix64(h ^ c[2]);
  h = abs_mix64(h ^ c[3]);
  return h;
}
template <class T, class U>
__host__ __device__ __forceinline__ int cmp256(const T* a, const U* b) {
  for (int i = 3; i >= 0; --i)
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  return 0;
}

// ---- row hashes ------------------------------------------------------------------------------------------------
// rows with at most ABS_HASH_SHORT stored terms (all but a few hundred of ecdsa's 1.09 M): a thread each; the sums are
// the same commutative sums the warp version reduces with shuffles
#define ABS_HASH_SHORT 16u
__global__ void k_abs_row_hash_short(uint64_t N, const unsigned long long* seg, const fr::u256* coef, unsigned long long* out) {
  const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= N) return;
  const unsigned long long s0 = seg[3 * row], s3 = seg[3 * row + 3];
  if (s3 - s0 > ABS_HASH_SHORT) return;
  unsigned long long h = 0x1234567ULL;
  unsigned long long b = s0;
  for (int f = 0; f < 3; ++f) {
    const unsigned long long e = seg[3 * row + f + 1];
    unsigned long long s = 0, cnt = 0;
    for (unsigned long long t = b; t < e; ++t) {
      const fr::u256 c = coef[t];
      if (!fr::is_zero(c)) {
        s += sig_term(0u, c.v, 0x51ed270b1ULL);
        cnt += 1;
      }
    }
    h = abs_mix64(h ^ s) + cnt * 0x9e3779b97f4a7c15ULL + (unsigned long long)f;
    b = e;
  }
  out[row] = h;
}
// the few long rows: a warp looks at 32 consecutive rows (one coalesced read of their lengths) and hashes the long ones
// among them with all its lanes, one after the other
__global__ void k_abs_row_hash(uint64_t N, const unsigned long long* seg, const fr::u256* coef,
                               unsigned long long* out) {
  const uint64_t row0 = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u;
  const unsigned int lane = threadIdx.x & 31u;
  if (row0 >= N) return;
  const uint64_t mine = row0 + lane;
  const bool is_long = mine < N && seg[3 * mine + 3] - seg[3 * mine] > ABS_HASH_SHORT;
  for (unsigned int m = __ballot_sync(0xffffffffu, is_long); m; m &= m - 1) {
    const uint64_t row = row0 + (unsigned int)(__ffs((int)m) - 1);
    unsigned long long h = 0x1234567ULL;
    for (int f = 0; f < 3; ++f) {
      unsigned long long s = 0, cnt = 0;
      for (uint64_t t = seg[3 * row + f] + lane; t < seg[3 * row + f + 1]; t += 32) {
        const fr::u256 c = coef[t];
        if (!fr::is_zero(c)) {
          s += sig_term(0u, c.v, 0x51ed270b1ULL);
          cnt += 1;
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      }
      h = abs_mix64(h ^ s) + cnt * 0x9e3779b97f4a7c15ULL + (unsigned long long)f;
    }
    if (lane == 0) out[row] = h;
  }
}
#define ABS_CAND_SERIAL 64u
__global__ void k_abs_candidates(uint64_t N, uint64_t n, const unsigned long long* hc, const unsigned long long* hs,
                                 unsigned long long* cand, unsigned int* n_cand, unsigned int cap) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i + n > N) return;
  // Windows of at most ABS_CAND_SERIAL rows are compared here, a thread each.  Longer ones (secp256k1 has 15 935 rows)
  // are only PROBED at a few positions spread over the window: the 25 threads of ecdsa's true candidates walking 16 k
  // hashes one after the other WERE the kernel (1.9 ms); the survivors are compared in full by a block each
  // (k_abs_candidates_full).  The conjunction does not care about the order of the comparisons.
  if (n > ABS_CAND_SERIAL) {
    for (uint64_t k = 0; k < 16; ++k) {
      const uint64_t j = (k * (n - 2)) / 15;
      if (hc[i + j] != hs[j]) return;
    }
  } else {
    for (uint64_t j = 0; j + 1 < n; ++j)
      if (hc[i + j] != hs[j]) return;
  }
  const unsigned int k = atomicAdd(n_cand, 1u);
  if (k < cap) cand[k] = i;
}
// a block per probed window: the first n - 1 row hashes line up (:259-270); keep[b] = 1 if so
__global__ void k_abs_candidates_full(uint64_t n, const unsigned long long* hc, const unsigned long long* hs,
                                      const unsigned long long* cand, unsigned int* keep) {
  const uint64_t i = cand[blockIdx.x];
  bool ok = true;
  for (uint64_t j = threadIdx.x; j + 1 < n; j += blockDim.x) ok &= hc[i + j] == hs[j];
  const int all = __syncthreads_and(ok ? 1 : 0);
  if (threadIdx.x == 0) keep[blockIdx.x] = all ? 1u : 0u;
}

// ---- the trusted circuit, prepared on the host, as the kernels see it -------------------------------------------
struct SubDev {
  uint64_t n;          // rows
  uint32_t snz;        // non-zero terms
  const uint32_t* snz_ptr;     // [3n + 1] offsets of each form's non-zero coefficients ...
  const fr::u256* snz_coef;    // ... sorted ascending per form
  const uint32_t* snz_expect;  // [snz] run length at the first element of a run of equal values, else 0
  uint32_t n_classes;
  const unsigned long long* class_hash;  // [n_classes] ascending, distinct
  const uint32_t* class_off;             // [n_classes + 1] into class_slot / class_coef (sorted by slot)
  const uint32_t* class_slot;
  const fr::u256* class_coef;
  const uint32_t* class_size;            // wires of the trusted circuit with this signature
  const uint8_t* class_needed;           // an input / output of the trusted circuit is in the class
  unsigned long long seed;
};
struct Batch {
  uint32_t n_cand;       // candidates in this batch
  const unsigned long long* start;  // [n_cand] window start rows
  uint32_t cap;          // slots of a candidate's wire table (power of two)
  uint32_t* t_key;       // [n_cand * cap] wire (0 = empty)
  unsigned long long* t_hash;  // signature hash sums
  uint32_t* t_cnt;       // appearances
  uint32_t* t_cls;       // class of the wire's signature
  uint32_t* cnt;         // [n_cand * snz] multiset counters
  uint32_t* pop;         // [n_cand * n_classes]
  uint32_t* fail;        // [n_cand]
};

// (a window that is not a match can hold more distinct wires than the trusted circuit has terms: probing is bounded,
// a full table fails the candidate)
__device__ __forceinline__ uint32_t table_find_or_insert(const Batch& b, uint32_t c, uint32_t wire) {
  uint32_t* keys = b.t_key + (size_t)c * b.cap;
  uint32_t s = (uint32_t)(abs_mix64(wire) & (b.cap - 1));
  for (uint32_t probes = 0; probes < b.cap; ++probes) {
    const uint32_t prev = atomicCAS(keys + s, 0u, wire);
    if (prev == 0u || prev == wire) return s;
    s = (s + 1) & (b.cap - 1);
  }
  return 0xffffffffu;
}
__device__ __forceinline__ uint32_t table_find(const Batch& b, uint32_t c, uint32_t wire) {
  const uint32_t* keys = b.t_key + (size_t)c * b.cap;
  uint32_t s = (uint32_t)(abs_mix64(wire) & (b.cap - 1));
  for (uint32_t probes = 0; probes < b.cap; ++probes) {
    const uint32_t k = keys[s];
    if (k == wire) return s;
    if (k == 0u) return 0xffffffffu;
    s = (s + 1) & (b.cap - 1);
  }
  return 0xffffffffu;
}

// blockIdx.y = candidate of the batch; the threads of the x dimension stride the window's stored terms
__global__ void k_abs_terms(const unsigned long long* seg, const uint32_t* col, const fr::u256* coef, SubDev S, Batch B) {
  const uint32_t c = blockIdx.y;
  const uint64_t i0 = B.start[c];
  const unsigned long long* wseg = seg + 3 * i0;  // [3n + 1] offsets of the window's forms
  const uint64_t t0 = wseg[0], t1 = wseg[3 * S.n];
  for (uint64_t t = t0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < t1; t += (uint64_t)gridDim.x * blockDim.x) {
    const fr::u256 cf = coef[t];
    if (fr::is_zero(cf)) continue;
    // the form that holds term t: the last segment that starts at or before t
    uint32_t lo = 0, hi = (uint32_t)(3 * S.n);
    while (hi - lo > 1) {
      const uint32_t m = (lo + hi) >> 1;
      if (wseg[m] <= t) lo = m; else hi = m;
    }
    const uint32_t form = lo;  // slot = form + 1 (:296-327 counts from 1)
    // (a) checkNonZeroValues (:205-226): the value must occur in the trusted form; occurrences are counted per run
    {
      uint32_t a = S.snz_ptr[form], e = S.snz_ptr[form + 1];
      const uint32_t e0 = e;
      while (a < e) {  // lower bound
        const uint32_t m = (a + e) >> 1;
        if (cmp256(S.snz_coef[m].v, cf.v) < 0) a = m + 1; else e = m;
      }
      if (a >= e0 || cmp256(S.snz_coef[a].v, cf.v) != 0)
        B.fail[c] = 1;
      else
        atomicAdd(B.cnt + (size_t)c * S.snz + a, 1u);
    }
    // (b) the wire's appearance signature (:305-310)
    const uint32_t s = table_find_or_insert(B, c, col[t]);
    if (s == 0xffffffffu) {
      B.fail[c] = 1;
      continue;
    }
    atomicAdd(B.t_hash + (size_t)c * B.cap + s, sig_term(form + 1, cf.v, S.seed));
    atomicAdd(B.t_cnt + (size_t)c * B.cap + s, 1u);
  }
}
__global__ void k_abs_counts(SubDev S, Batch B) {
  const uint32_t c = blockIdx.y;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < S.snz; q += gridDim.x * blockDim.x)
    if (B.cnt[(size_t)c * S.snz + q] != S.snz_expect[q]) B.fail[c] = 1;
}
__global__ void k_abs_wires(SubDev S, Batch B) {
  const uint32_t c = blockIdx.y;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < B.cap; s += gridDim.x * blockDim.x) {
    const size_t at = (size_t)c * B.cap + s;
    if (B.t_key[at] == 0u) continue;
    const unsigned long long h = B.t_hash[at];
    uint32_t a = 0, e = S.n_classes;
    while (a < e) {
      const uint32_t m = (a + e) >> 1;
      if (S.class_hash[m] < h) a = m + 1; else e = m;
    }
    if (a >= S.n_classes || S.class_hash[a] != h || S.class_off[a + 1] - S.class_off[a] != B.t_cnt[at]) {
      B.t_cls[at] = 0xffffffffu;
      B.fail[c] = 1;
      continue;
    }
    B.t_cls[at] = a;
    atomicAdd(B.pop + (size_t)c * S.n_classes + a, 1u);
  }
}
// every appearance of a window wire is an appearance of its class: same slot, same coefficient.  Together with the
// equal appearance counts (k_abs_wires) and distinct slots that makes the two signatures equal term by term.
__global__ void k_abs_verify(const unsigned long long* seg, const uint32_t* col, const fr::u256* coef, SubDev S, Batch B) {
  const uint32_t c = blockIdx.y;
  if (B.fail[c]) return;
  const uint64_t i0 = B.start[c];
  const unsigned long long* wseg = seg + 3 * i0;
  const uint64_t t0 = wseg[0], t1 = wseg[3 * S.n];
  for (uint64_t t = t0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < t1; t += (uint64_t)gridDim.x * blockDim.x) {
    const fr::u256 cf = coef[t];
    if (fr::is_zero(cf)) continue;
    uint32_t lo = 0, hi = (uint32_t)(3 * S.n);
    while (hi - lo > 1) {
      const uint32_t m = (lo + hi) >> 1;
      if (wseg[m] <= t) lo = m; else hi = m;
    }
    const uint32_t slot = lo + 1;
    const uint32_t s = table_find(B, c, col[t]);
    const uint32_t cls = s == 0xffffffffu ? 0xffffffffu : B.t_cls[(size_t)c * B.cap + s];
    if (cls == 0xffffffffu) {
      B.fail[c] = 1;
      continue;
    }
    uint32_t a = S.class_off[cls], e = S.class_off[cls + 1];
    const uint32_t e0 = e;
    while (a < e) {
      const uint32_t m = (a + e) >> 1;
      if (S.class_slot[m] < slot) a = m + 1; else e = m;
    }
    if (a >= e0 || S.class_slot[a] != slot || cmp256(S.class_coef[a].v, cf.v) != 0) B.fail[c] = 1;
  }
}
__global__ void k_abs_pop(SubDev S, Batch B) {
  const uint32_t c = blockIdx.y;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < S.n_classes; k += gridDim.x * blockDim.x)
    if (B.pop[(size_t)c * S.n_classes + k] != S.class_size[k]) B.fail[c] = 1;
}
// the window's wires that sit in a class of an input / output of the trusted circuit: {candidate, class, wire}
__global__ void k_abs_collect(SubDev S, Batch B, uint32_t* out, unsigned int* n_out, unsigned int cap_out) {
  const uint32_t c = blockIdx.y;
  if (B.fail[c]) return;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < B.cap; s += gridDim.x * blockDim.x) {
    const size_t at = (size_t)c * B.cap + s;
    const uint32_t w = B.t_key[at];
    if (w == 0u) continue;
    const uint32_t cls = B.t_cls[at];
    if (cls == 0xffffffffu || !S.class_needed[cls]) continue;
    const unsigned int k = atomicAdd(n_out, 1u);
    if (k < cap_out) {
      out[3 * k + 0] = c;
      out[3 * k + 1] = cls;
      out[3 * k + 2] = w;
    }
  }
}

__global__ void k_abs_gather_offsets(const unsigned long long* seg, const unsigned long long* rows, uint32_t n,
                                     unsigned long long* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = seg[3 * rows[i]];
}
// ---- compaction of the kept rows (:368-388) --------------------------------------------------------------------
struct Run {
  unsigned long long r0, r1, row_dst, term_src, term_dst;  // rows [r0, r1) of the source land at row_dst / term_dst
};
__global__ void k_abs_compact_rows(const Run* runs, uint32_t n_runs, uint64_t rows_out, const unsigned long long* seg,
                                   unsigned long long* seg_out, unsigned long long terms_out) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > rows_out) return;
  if (r == rows_out) {
    seg_out[3 * rows_out] = terms_out;
    return;
  }
  uint32_t lo = 0, hi = n_runs;  // the run whose destination range holds output row r
  while (hi - lo > 1) {
    const uint32_t m = (lo + hi) >> 1;
    if (runs[m].row_dst <= r) lo = m; else hi = m;
  }
  const Run u = runs[lo];
  const uint64_t src = u.r0 + (r - u.row_dst);
  for (int f = 0; f < 3; ++f) seg_out[3 * r + f] = seg[3 * src + f] - u.term_src + u.term_dst;
}
__global__ void k_abs_compact_terms(const Run* runs, uint32_t n_runs, unsigned long long terms_out, const uint32_t* col,
                                    const fr::u256* coef, uint32_t* col_out, fr::u256* coef_out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= terms_out) return;
  uint32_t lo = 0, hi = n_runs;
  while (hi - lo > 1) {
    const uint32_t m = (lo + hi) >> 1;
    if (runs[m].term_dst <= t) lo = m; else hi = m;
  }
  const uint64_t src = runs[lo].term_src + (t - runs[lo].term_dst);
  col_out[t] = col[src];
  coef_out[t] = coef[src];
}

// ---- host side -----------------------------------------------------------------------------------------------------
namespace {
struct SubHost {  // everything about the trusted circuit that does not depend on the big system
  uint64_t n = 0;
  std::vector<uint32_t> snz_ptr, snz_expect, class_off, class_slot, class_size;
  std::vector<unsigned long long> snz_coef, class_hash, class_coef;  // 4 limbs per value
  std::vector<uint8_t> class_needed;
  unsigned long long seed = 0;
  uint32_t snz = 0;
  // every wire of the trusted circuit that appears: {wire, class of its signature, rank among the class's wires by id}
  struct Where {
    uint32_t wire, cls, rank;
  };
  std::vector<Where> where;  // sorted by wire
  // the trusted circuit's inputs (known, wire 1 left out, :376) and outputs (:382): class and rank inside it
  struct Need {
    uint32_t wire, cls, rank;  // cls == 0xffffffff: the wire never appears with a non-zero coefficient (KeyError)
  };
  std::vector<Need> need_in, need_out;
};

struct SigTerm {
  uint32_t wire, slot;
  const uint64_t* c;
};

bool prepare_sub(const ecne_problem_t* sub, SubHost& H) {
  const uint64_t n = sub->n_rows;
  H.n = n;
  // (a) sorted non-zero coefficients per form
  H.snz_ptr.assign(3 * n + 1, 0);
  std::vector<const uint64_t*> tmp;
  std::vector<SigTerm> terms;
  for (uint64_t s = 0; s < 3 * n; ++s) {
    tmp.clear();
    for (uint64_t t = sub->seg_ptr[s]; t < sub->seg_ptr[s + 1]; ++t) {
      const uint64_t* c = sub->coef + 4 * t;
      if ((c[0] | c[1] | c[2] | c[3]) == 0) continue;
      tmp.push_back(c);
      terms.push_back(SigTerm{sub->col[t], (uint32_t)s + 1, c});
    }
    std::sort(tmp.begin(), tmp.end(), [](const uint64_t* a, const uint64_t* b) {
      return cmp256(a, b) < 0;
    });
    const size_t base = H.snz_expect.size();
    for (size_t k = 0; k < tmp.size(); ++k) {
      for (int l = 0; l < 4; ++l) H.snz_coef.push_back(tmp[k][l]);
      H.snz_expect.push_back(0);
    }
    for (size_t k = 0; k < tmp.size();) {
      size_t e = k + 1;
      while (e < tmp.size() && cmp256(tmp[e], tmp[k]) == 0) ++e;
      H.snz_expect[base + k] = (uint32_t)(e - k);
      k = e;
    }
    H.snz_ptr[s + 1] = (uint32_t)H.snz_expect.size();
  }
  H.snz = (uint32_t)H.snz_expect.size();
  // (b) appearance signatures (:276-292): terms by (wire, slot); a signature is a wire's (slot, coefficient) list
  std::stable_sort(terms.begin(), terms.end(), [](const SigTerm& a, const SigTerm& b) { return a.wire < b.wire; });
  struct W {
    uint32_t wire, b, e;  // range in `terms`
    unsigned long long h;
  };
  std::vector<W> ws;
  for (size_t i = 0; i < terms.size();) {
    size_t e = i + 1;
    while (e < terms.size() && terms[e].wire == terms[i].wire) ++e;
    ws.push_back(W{terms[i].wire, (uint32_t)i, (uint32_t)e, 0});
    i = e;
  }
  auto same_sig = [&](const W& a, const W& b) {
    if (a.e - a.b != b.e - b.b) return false;
    for (uint32_t k = 0; k < a.e - a.b; ++k) {
      const SigTerm &x = terms[a.b + k], &y = terms[b.b + k];
      if (x.slot != y.slot || cmp256(x.c, y.c) != 0) return false;
    }
    return true;
  };
  for (int attempt = 0; attempt < 16; ++attempt) {
    H.seed = abs_mix64(0x6a09e667f3bcc908ULL + 0x9e3779b97f4a7c15ULL * (unsigned long long)attempt);
    for (auto& w : ws) {
      unsigned long long h = 0;
      for (uint32_t k = w.b; k < w.e; ++k) h += sig_term(terms[k].slot, terms[k].c, H.seed);
      w.h = h;
    }
    std::vector<uint32_t> order(ws.size());
    for (size_t i = 0; i < ws.size(); ++i) order[i] = (uint32_t)i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
      if (ws[a].h != ws[b].h) return ws[a].h < ws[b].h;
      return ws[a].wire < ws[b].wire;
    });
    // classes = runs of equal hash; two DIFFERENT signatures under one hash: take another seed
    bool collision = false;
    H.class_hash.clear();
    H.class_off.assign(1, 0);
    H.class_slot.clear();
    H.class_coef.clear();
    H.class_size.clear();
    H.where.clear();
    for (size_t i = 0; i < order.size() && !collision;) {
      size_t e = i + 1;
      while (e < order.size() && ws[order[e]].h == ws[order[i]].h) {
        if (!same_sig(ws[order[e]], ws[order[i]])) collision = true;
        ++e;
      }
      const W& w0 = ws[order[i]];
      H.class_hash.push_back(w0.h);
      for (uint32_t k = w0.b; k < w0.e; ++k) {
        H.class_slot.push_back(terms[k].slot);
        for (int l = 0; l < 4; ++l) H.class_coef.push_back(terms[k].c[l]);
      }
      H.class_off.push_back((uint32_t)H.class_slot.size());
      H.class_size.push_back((uint32_t)(e - i));
      for (size_t k = i; k < e; ++k)  // ascending wire ids (sort key): the rank is the position in the run
        H.where.push_back(SubHost::Where{ws[order[k]].wire, (uint32_t)(H.class_hash.size() - 1), (uint32_t)(k - i)});
      i = e;
    }
    if (!collision) break;
    if (attempt == 15) return false;
  }
  // (c) where the inputs / outputs sit
  H.class_needed.assign(H.class_hash.size(), 0);
  std::sort(H.where.begin(), H.where.end(), [](const SubHost::Where& a, const SubHost::Where& b) { return a.wire < b.wire; });
  auto locate = [&](uint32_t x) {
    SubHost::Need nd{x, 0xffffffffu, 0};
    auto it = std::lower_bound(H.where.begin(), H.where.end(), x, [](const SubHost::Where& a, uint32_t v) { return a.wire < v; });
    if (it != H.where.end() && it->wire == x) {
      nd.cls = it->cls;
      nd.rank = it->rank;
      H.class_needed[it->cls] = 1;
    }
    return nd;
  };
  for (uint64_t k = 0; k < sub->n_known; ++k)
    if (sub->known[k] != 1) H.need_in.push_back(locate(sub->known[k]));
  for (uint64_t k = 0; k < sub->n_targets; ++k) H.need_out.push_back(locate(sub->targets[k]));
  return true;
}

template <class T>
cudaError_t up(Arena& a, const std::vector<T>& v, const T** out, cudaStream_t s, size_t elems_per = 1) {
  T* d = nullptr;
  cudaError_t e = a.alloc(&d, v.size() + 1);
  if (e != cudaSuccess) return e;
  if (!v.empty()) e = cudaMemcpyAsync(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
  (void)elems_per;
  *out = d;
  return e;
}
}  // namespace

// Pageable host arrays reach the device at ~8 GB/s through cudaMemcpyAsync's own bounce buffer.  Large uploads are
// staged instead: a pool of host threads (started once, kept for the life of the process) copies 1 MB slices into
// pinned buffers — one per worker — and sends every slice with its own asynchronous copy, so the memcpy of one slice
// overlaps the DMA of the slices before it.  staged_h2d_async() only queues the slices; staged_flush() returns when
// every queued slice has been handed to its stream (kernels that read the data are launched after it).
namespace {
struct StagePool {
  static constexpr int WORKERS = 12;
  static constexpr size_t SLOT_BYTES = (size_t)1 << 20;
  struct Task {
    char* dst;
    const char* src;
    size_t len;
    cudaStream_t s;
    int dev;
  };
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::deque<Task> q;
  size_t pending = 0;  // queued or being copied
  cudaError_t err = cudaSuccess;
  bool started = false, usable = true;
  void worker() {
    char* buf = nullptr;
    cudaEvent_t done = nullptr;
    int dev = -1;
    for (;;) {
      Task t;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_work.wait(lk, [&] { return !q.empty(); });
        t = q.front();
        q.pop_front();
      }
      cudaError_t e = cudaSuccess;
      if (t.dev != dev) {  // (the buffer is pinned for every context; the event belongs to a device)
        if (done) {
          cudaEventSynchronize(done);
          cudaEventDestroy(done);
          done = nullptr;
        }
        e = cudaSetDevice(t.dev);
        if (e == cudaSuccess && !buf) e = cudaHostAlloc((void**)&buf, SLOT_BYTES, cudaHostAllocPortable);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done, cudaEventDisableTiming);
        if (e == cudaSuccess) dev = t.dev;
      } else {
        e = cudaEventSynchronize(done);  // the copy that last read the buffer
      }
      if (e == cudaSuccess) {
        memcpy(buf, t.src, t.len);
        e = cudaMemcpyAsync(t.dst, buf, t.len, cudaMemcpyHostToDevice, t.s);
      }
      if (e == cudaSuccess) e = cudaEventRecord(done, t.s);
      {
        std::lock_guard<std::mutex> lk(mu);
        if (e != cudaSuccess && err == cudaSuccess) err = e;
        if (--pending == 0) cv_done.notify_all();
      }
    }
  }
  bool start() {
    std::lock_guard<std::mutex> lk(mu);
    if (started) return usable;
    started = true;
    try {
      for (int i = 0; i < WORKERS; ++i) std::thread([this] { worker(); }).detach();
    } catch (...) {
      usable = false;
    }
    return usable;
  }
};
StagePool& stage_pool() {
  static StagePool* p = new StagePool();  // (never destroyed: its threads outlive main)
  return *p;
}
}  // namespace

cudaError_t staged_h2d_async(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  if (!bytes) return cudaSuccess;
  cudaPointerAttributes at;
  const bool pinned = cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost;
  cudaGetLastError();  // (an unregistered host pointer is reported as an error by older runtimes)
  StagePool& P = stage_pool();
  if (pinned || bytes < ((size_t)1 << 20) || !P.start())
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
  int dev = 0;
  cudaGetDevice(&dev);
  const size_t SB = StagePool::SLOT_BYTES;
  {
    std::lock_guard<std::mutex> lk(P.mu);
    for (size_t off = 0; off < bytes; off += SB) {
      P.q.push_back({(char*)dst + off, (const char*)src + off, std::min(SB, bytes - off), s, dev});
      ++P.pending;
    }
  }
  P.cv_work.notify_all();
  return cudaSuccess;
}
cudaError_t staged_flush() {
  StagePool& P = stage_pool();
  std::unique_lock<std::mutex> lk(P.mu);
  P.cv_done.wait(lk, [&] { return P.pending == 0; });
  const cudaError_t e = P.err;
  P.err = cudaSuccess;
  return e;
}
cudaError_t staged_h2d(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  cudaError_t e = staged_h2d_async(dst, src, bytes, s);
  const cudaError_t f = staged_flush();
  return e != cudaSuccess ? e : f;
}

int dev_system_upload(const ecne_problem_t* p, DevSystem* S, cudaStream_t s, std::string& err) {
  int st = problem_rows_ok(p, err);
  if (st != ECNE_OK) return st;
  const uint64_t N = p->n_rows, nnz = problem_nnz(p);
  S->N = N;
  S->V = p->n_vars;
  S->nnz = nnz;
  const bool prof = getenv("ECNE_HOST_PROF") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  CKE(S->arena.alloc(&S->seg, upload_padded<unsigned long long>(3 * N + 2)));
  CKE(S->arena.alloc(&S->col, upload_padded<uint32_t>(nnz + 1)));
  CKE(S->arena.alloc(&S->coef, upload_padded<fr::u256>(nnz + 1)));
  if (prof)
    fprintf(stderr, "[ecne dev] device arrays of the unreduced system allocated in %.3f ms\n",
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  Arena tmp;
  tmp.pool = S->arena.pool;
  st = upload_rows(p, S->seg, S->col, S->coef, tmp, s, err);
  if (tmp.pool && !tmp.slabs.empty()) {
    cudaStreamSynchronize(s);  // the expansion kernels read the scratch
    tmp.release();
  }
  return st;
}

// One abstraction() call on a device-resident system: `S` is replaced by the reduced system, the special
// constraints of the consumed windows are appended to `sp`.
struct PreparedSub {
  SubHost H;
  uint64_t n_rows, nnz;
};
PreparedSub* abstraction_prepare(const ecne_problem_t* sub) {
  if (!sub || !sub->seg_ptr || sub->n_rows == 0 || !sub->col || !sub->coef) return nullptr;
  PreparedSub* p = new (std::nothrow) PreparedSub();
  if (!p) return nullptr;
  p->n_rows = sub->n_rows;
  p->nnz = sub->seg_ptr[3 * sub->n_rows];
  bool ok = false;
  try {
    ok = prepare_sub(sub, p->H);
  } catch (...) {
    ok = false;
  }
  if (!ok) {
    delete p;
    return nullptr;
  }
  return p;
}
void abstraction_prepared_free(PreparedSub* p) { delete p; }

int dev_abstraction(DevSystem* S, int32_t kind, const ecne_problem_t* sub, SpecialsHost* sp, uint64_t* n_matches,
                    cudaStream_t s, std::string& err, AbstractionStats* stats, const std::function<int()>& ready,
                    const PreparedSub* prepared) {
  if (!sub || !sub->seg_ptr || sub->n_rows == 0) {
    err = "trusted circuit without rows";
    return ECNE_E_BADARG;
  }
  if (!sub->col || !sub->coef) {  // (a few thousand rows, prepared on the host: the compact form would buy nothing)
    err = "a trusted circuit is passed with its full 32-byte coefficients";
    return ECNE_E_BADARG;
  }
  auto tp0 = std::chrono::steady_clock::now();
  auto lap = [&](double* acc) {
    cudaStreamSynchronize(s);
    auto t = std::chrono::steady_clock::now();
    if (acc) *acc += std::chrono::duration<double, std::milli>(t - tp0).count();
    tp0 = t;
  };
  const uint64_t n = sub->n_rows;
  SubHost H_local;
  if (prepared) {
    if (prepared->n_rows != sub->n_rows || prepared->nnz != sub->seg_ptr[3 * sub->n_rows]) {
      err = "the prepared trusted circuit is not the one passed with it";
      return ECNE_E_BADARG;
    }
  } else if (!prepare_sub(sub, H_local)) {
    err = "internal: signature hash of the trusted circuit collides under every seed";
    return ECNE_E_INTERNAL;
  }
  const SubHost& H = prepared ? prepared->H : H_local;
  if (stats) stats->ms_prepare += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tp0).count();
  if (ready) {  // the big system is on the device from here on
    const int rst = ready();
    if (rst != ECNE_OK) {
      err = "upload of the unreduced system failed";
      return rst;
    }
  }
  tp0 = std::chrono::steady_clock::now();
  const uint64_t N = S->N;
  Arena tmp;
  struct Guard {
    Arena& a;
    ~Guard() {
      if (a.pool) a.release();
    }
  } guard{tmp};
  // ---- the trusted circuit on the device ------------------------------------------------------------
  DevSystem sd;
  {
    int st = dev_system_upload(sub, &sd, s, err);
    if (st != ECNE_OK) {
      if (sd.arena.pool) sd.arena.release();
      return st;
    }
  }
  struct Guard2 {
    Arena& a;
    ~Guard2() {
      if (a.pool) a.release();
    }
  } guard2{sd.arena};
  SubDev D;
  memset(&D, 0, sizeof(D));
  D.n = n;
  D.snz = H.snz;
  D.n_classes = (uint32_t)H.class_hash.size();
  D.seed = H.seed;
  CKE(up(tmp, H.snz_ptr, &D.snz_ptr, s));
  CKE(up(tmp, H.snz_expect, &D.snz_expect, s));
  CKE(up(tmp, H.snz_coef, (const unsigned long long**)&D.snz_coef, s));
  CKE(up(tmp, H.class_hash, &D.class_hash, s));
  CKE(up(tmp, H.class_off, &D.class_off, s));
  CKE(up(tmp, H.class_slot, &D.class_slot, s));
  CKE(up(tmp, H.class_coef, (const unsigned long long**)&D.class_coef, s));
  CKE(up(tmp, H.class_size, &D.class_size, s));
  CKE(up(tmp, H.class_needed, &D.class_needed, s));
  // ---- row hashes, candidates (:252-270) ----------------------------------------------------------------
  unsigned long long *d_hc = nullptr, *d_hs = nullptr, *d_cand = nullptr;
  unsigned int* d_ncand = nullptr;
  CKE(tmp.alloc(&d_hc, N + 1));
  CKE(tmp.alloc(&d_hs, n + 1));
  CKE(tmp.alloc(&d_ncand, 4));
  CKE(cudaMemsetAsync(d_ncand, 0, 16, s));
  if (N) {
    k_abs_row_hash_short<<<(unsigned int)((N + 255) / 256), 256, 0, s>>>(N, S->seg, S->coef, d_hc);
    k_abs_row_hash<<<(unsigned int)((N + 255) / 256), 256, 0, s>>>(N, S->seg, S->coef, d_hc);
  }
  k_abs_row_hash_short<<<(unsigned int)((n + 255) / 256), 256, 0, s>>>(n, sd.seg, sd.coef, d_hs);
  k_abs_row_hash<<<(unsigned int)((n + 255) / 256), 256, 0, s>>>(n, sd.seg, sd.coef, d_hs);
  lap(stats ? &stats->ms_hash : nullptr);
  std::vector<unsigned long long> cand;
  if (N >= n) {
    unsigned int cap = 1u << 16, nc = 0;
    for (int pass = 0; pass < 2; ++pass) {  // second pass only if the list outgrew its first size
      CKE(tmp.alloc(&d_cand, cap));
      CKE(cudaMemsetAsync(d_ncand, 0, 4, s));
      k_abs_candidates<<<(unsigned int)((N - n + 1 + 255) / 256), 256, 0, s>>>(N, n, d_hc, d_hs, d_cand, d_ncand, cap);
      CKE(cudaMemcpyAsync(&nc, d_ncand, 4, cudaMemcpyDeviceToHost, s));
      CKE(cudaStreamSynchronize(s));
      if (nc <= cap) break;
      cap = nc;
    }
    cand.resize(nc);
    if (nc) CKE(cudaMemcpy(cand.data(), d_cand, (size_t)nc * 8, cudaMemcpyDeviceToHost));
    if (nc && n > ABS_CAND_SERIAL) {  // the probed windows in full, a block each
      unsigned int* d_keep = nullptr;
      CKE(tmp.alloc(&d_keep, nc));
      std::vector<unsigned int> keep(nc);
      for (unsigned int b0 = 0; b0 < nc; b0 += 65535u) {
        const unsigned int nb_ = std::min(65535u, nc - b0);
        k_abs_candidates_full<<<nb_, 256, 0, s>>>(n, d_hc, d_hs, d_cand + b0, d_keep + b0);
      }
      CKE(cudaMemcpyAsync(keep.data(), d_keep, (size_t)nc * 4, cudaMemcpyDeviceToHost, s));
      CKE(cudaStreamSynchronize(s));
      size_t w = 0;
      for (unsigned int k = 0; k < nc; ++k)
        if (keep[k]) cand[w++] = cand[k];
      cand.resize(w);
    }
    std::sort(cand.begin(), cand.end());  // ascending window starts, as the reference's loop finds them (:259)
  }
  lap(stats ? &stats->ms_candidates : nullptr);
  if (stats) stats->n_candidates += cand.size();
  // ---- verification of every candidate (:293-352), in batches that fit a memory budget -------------------
  struct Match {
    unsigned long long start;
    std::vector<std::vector<uint32_t>> cls_wires;  // per needed class: the window's wires in it, ascending
  };
  std::vector<Match> matches;
  std::vector<uint32_t> needed_ids;  // class id -> dense index among the needed classes
  std::vector<int> needed_index(D.n_classes, -1);
  for (uint32_t k = 0; k < D.n_classes; ++k)
    if (H.class_needed[k]) {
      needed_index[k] = (int)needed_ids.size();
      needed_ids.push_back(k);
    }
  if (!cand.empty()) {
    uint32_t cap = 8;
    while (cap < 2u * std::max<uint32_t>(H.snz, 1)) cap <<= 1;
    const size_t per_cand = (size_t)cap * (4 + 8 + 4 + 4) + (size_t)H.snz * 4 + (size_t)D.n_classes * 4 + 64;
    size_t bsz = std::max<size_t>(1, ((size_t)384 << 20) / per_cand);
    bsz = std::min<size_t>(bsz, std::min<size_t>(cand.size(), 65535));
    Batch B;
    memset(&B, 0, sizeof(B));
    B.cap = cap;
    unsigned long long* d_start = nullptr;
    uint32_t* d_out = nullptr;
    unsigned int* d_nout = nullptr;
    const unsigned int cap_out = (unsigned int)std::min<size_t>(bsz * (size_t)std::max<uint32_t>(H.snz, 1), (size_t)1 << 24);
    CKE(tmp.alloc(&d_start, bsz));
    CKE(tmp.alloc(&B.t_key, bsz * cap));
    CKE(tmp.alloc(&B.t_hash, bsz * cap));
    CKE(tmp.alloc(&B.t_cnt, bsz * cap));
    CKE(tmp.alloc(&B.t_cls, bsz * cap));
    CKE(tmp.alloc(&B.cnt, bsz * (size_t)std::max<uint32_t>(H.snz, 1)));
    CKE(tmp.alloc(&B.pop, bsz * (size_t)std::max<uint32_t>(D.n_classes, 1)));
    CKE(tmp.alloc(&B.fail, bsz));
    CKE(tmp.alloc(&d_out, 3 * (size_t)cap_out));
    CKE(tmp.alloc(&d_nout, 4));
    B.start = d_start;
    std::vector<uint32_t> h_fail, h_out;
    for (size_t b0 = 0; b0 < cand.size(); b0 += bsz) {
      const uint32_t nb = (uint32_t)std::min(bsz, cand.size() - b0);
      B.n_cand = nb;
      CKE(cudaMemcpyAsync(d_start, cand.data() + b0, (size_t)nb * 8, cudaMemcpyHostToDevice, s));
      CKE(cudaMemsetAsync(B.t_key, 0, (size_t)nb * cap * 4, s));
      CKE(cudaMemsetAsync(B.t_hash, 0, (size_t)nb * cap * 8, s));
      CKE(cudaMemsetAsync(B.t_cnt, 0, (size_t)nb * cap * 4, s));
      CKE(cudaMemsetAsync(B.cnt, 0, (size_t)nb * std::max<uint32_t>(H.snz, 1) * 4, s));
      CKE(cudaMemsetAsync(B.pop, 0, (size_t)nb * std::max<uint32_t>(D.n_classes, 1) * 4, s));
      CKE(cudaMemsetAsync(B.fail, 0, (size_t)nb * 4, s));
      CKE(cudaMemsetAsync(d_nout, 0, 4, s));
      // window sizes differ little: size the x dimension for the trusted circuit's stored terms
      const uint64_t sub_terms = sub->seg_ptr[3 * n];
      const unsigned int gx = (unsigned int)std::min<uint64_t>(std::max<uint64_t>(1, (sub_terms + 255) / 256), 1024);
      const dim3 gt(gx, nb);
      k_abs_terms<<<gt, 256, 0, s>>>(S->seg, S->col, S->coef, D, B);
      auto gx_for = [](uint64_t items) { return (unsigned int)std::min<uint64_t>(std::max<uint64_t>(1, (items + 255) / 256), 1024); };
      k_abs_counts<<<dim3(gx_for(H.snz), nb), 256, 0, s>>>(D, B);
      k_abs_wires<<<dim3(gx_for(cap), nb), 256, 0, s>>>(D, B);
      k_abs_verify<<<gt, 256, 0, s>>>(S->seg, S->col, S->coef, D, B);
      k_abs_pop<<<dim3(gx_for(D.n_classes), nb), 256, 0, s>>>(D, B);
      k_abs_collect<<<dim3(gx_for(cap), nb), 256, 0, s>>>(D, B, d_out, d_nout, cap_out);
      unsigned int n_out = 0;
      h_fail.resize(nb);
      CKE(cudaMemcpyAsync(h_fail.data(), B.fail, (size_t)nb * 4, cudaMemcpyDeviceToHost, s));
      CKE(cudaMemcpyAsync(&n_out, d_nout, 4, cudaMemcpyDeviceToHost, s));
      CKE(cudaStreamSynchronize(s));
      CKE(cudaGetLastError());
      if (n_out > cap_out) {
        err = "internal: input / output wire list of the matches overflowed";
        return ECNE_E_INTERNAL;
      }
      h_out.resize(3 * (size_t)n_out);
      if (n_out) CKE(cudaMemcpy(h_out.data(), d_out, (size_t)n_out * 12, cudaMemcpyDeviceToHost));
      std::vector<int> slot_of(nb, -1);
      for (uint32_t c = 0; c < nb; ++c)
        if (!h_fail[c]) {
          slot_of[c] = (int)matches.size();
          Match m;
          m.start = cand[b0 + c];
          m.cls_wires.resize(needed_ids.size());
          matches.push_back(std::move(m));
        }
      for (unsigned int k = 0; k < n_out; ++k) {
        const uint32_t c = h_out[3 * k], cls = h_out[3 * k + 1], w = h_out[3 * k + 2];
        if (c < nb && slot_of[c] >= 0 && cls < D.n_classes && needed_index[cls] >= 0)
          matches[slot_of[c]].cls_wires[needed_index[cls]].push_back(w);
      }
    }
    for (auto& m : matches)
      for (auto& v : m.cls_wires) std::sort(v.begin(), v.end());
  }
  lap(stats ? &stats->ms_verify : nullptr);
  if (stats) stats->n_matches += matches.size();
  // ---- the walk (:357-388), including the stall after an overlapping match (:370) ------------------------
  std::vector<size_t> consumed;
  {
    size_t cur = 0;
    uint64_t i = 0;
    while (i < N && cur < matches.size()) {
      if (matches[cur].start < i) break;  // starts inside a consumed window: cur never advances again
      i = matches[cur].start + n;
      consumed.push_back(cur);
      cur += 1;
    }
  }
  uint64_t added = 0;
  for (size_t ci : consumed) {
    const Match& m = matches[ci];
    auto mapped = [&](const SubHost::Need& nd, uint32_t* out) {
      if (nd.cls == 0xffffffffu) return false;  // m[x] of a wire that never appears: KeyError (:381-382)
      const auto& v = m.cls_wires[needed_index[nd.cls]];
      if (nd.rank >= v.size()) return false;
      *out = v[nd.rank];
      return true;
    };
    std::vector<uint32_t> in, outv;
    for (auto& nd : H.need_in) {
      uint32_t w = 0;
      if (!mapped(nd, &w)) {
        err = "KeyError: trusted input wire never appears (:381)";
        return ECNE_E_KEYERROR;
      }
      in.push_back(w);
    }
    for (auto& nd : H.need_out) {
      uint32_t w = 0;
      if (!mapped(nd, &w)) {
        err = "KeyError: trusted output wire never appears (:382)";
        return ECNE_E_KEYERROR;
      }
      outv.push_back(w);
    }
    sp->kind.push_back(kind);
    sp->in.insert(sp->in.end(), in.begin(), in.end());
    sp->out.insert(sp->out.end(), outv.begin(), outv.end());
    sp->in_ptr.push_back(sp->in.size());
    sp->out_ptr.push_back(sp->out.size());
    ++added;
  }
  if (n_matches) *n_matches = added;
  // ---- the kept rows, compacted on the device ----------------------------------------------------------
  if (!consumed.empty()) {
    // the term offsets of the run boundaries: a few dozen values, gathered on the device
    std::vector<unsigned long long> brow, boff;
    {
      uint64_t row_src = 0;
      for (size_t ci : consumed) {
        brow.push_back(row_src);
        brow.push_back(matches[ci].start);
        row_src = matches[ci].start + n;
      }
      brow.push_back(row_src);
      brow.push_back(N);
      unsigned long long *d_rows = nullptr, *d_offs = nullptr;
      CKE(tmp.alloc(&d_rows, brow.size()));
      CKE(tmp.alloc(&d_offs, brow.size()));
      CKE(cudaMemcpyAsync(d_rows, brow.data(), brow.size() * 8, cudaMemcpyHostToDevice, s));
      k_abs_gather_offsets<<<(unsigned int)((brow.size() + 127) / 128), 128, 0, s>>>(S->seg, d_rows, (uint32_t)brow.size(), d_offs);
      boff.resize(brow.size());
      CKE(cudaMemcpyAsync(boff.data(), d_offs, brow.size() * 8, cudaMemcpyDeviceToHost, s));
      CKE(cudaStreamSynchronize(s));
    }
    std::vector<Run> runs;
    uint64_t row_dst = 0, term_dst = 0;
    for (size_t k = 0; k + 1 < brow.size(); k += 2) {
      const uint64_t r0 = brow[k], r1 = brow[k + 1];
      if (r1 <= r0) continue;
      runs.push_back(Run{r0, r1, row_dst, boff[k], term_dst});
      row_dst += r1 - r0;
      term_dst += boff[k + 1] - boff[k];
    }
    const uint64_t rows_out = row_dst, terms_out = term_dst;
    DevSystem R;
    R.N = rows_out;
    R.V = S->V;
    R.nnz = terms_out;
    R.arena.pool = S->arena.pool;
    CKE(R.arena.alloc(&R.seg, 3 * rows_out + 2));
    CKE(R.arena.alloc(&R.col, terms_out + 1));
    CKE(R.arena.alloc(&R.coef, terms_out + 1));
    if (runs.empty()) {
      CKE(cudaMemsetAsync(R.seg, 0, 8, s));
    } else {
      Run* d_runs = nullptr;
      CKE(tmp.alloc(&d_runs, runs.size()));
      CKE(cudaMemcpyAsync(d_runs, runs.data(), runs.size() * sizeof(Run), cudaMemcpyHostToDevice, s));
      k_abs_compact_rows<<<(unsigned int)((rows_out + 1 + 255) / 256), 256, 0, s>>>(d_runs, (uint32_t)runs.size(), rows_out, S->seg,
                                                                                R.seg, terms_out);
      if (terms_out)
        k_abs_compact_terms<<<(unsigned int)((terms_out + 255) / 256), 256, 0, s>>>(d_runs, (uint32_t)runs.size(), terms_out, S->col,
                                                                                    S->coef, R.col, R.coef);
    }
    CKE(cudaStreamSynchronize(s));
    CKE(cudaGetLastError());
    S->arena.release();
    S->arena = std::move(R.arena);
    S->seg = R.seg;
    S->col = R.col;
    S->coef = R.coef;
    S->N = R.N;
    S->nnz = R.nnz;
  }
  lap(stats ? &stats->ms_compact : nullptr);
  return ECNE_OK;
}

}  // namespace ecne
