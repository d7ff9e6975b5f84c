// setup.cu — upload + one-shot classification of a problem (the static half of every rule,
// SURVEY.md Appendix D): streams the full 36-byte CSR terms once, does all field arithmetic
// (Montgomery inverse for -intercept/slope, pattern tests against +-2^i mod p), and lays the rows out
// for the sweep kernels (non-zero terms only; C terms of linear rows sorted by |fold(coef)|).
#include <cub/cub.cuh>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <mutex>

#include "engine_host.h"

namespace ecne {

#define LONG_T 6u           // rows with more non-zero terms than the inline record holds get a whole warp
#define N_CONST 255u        // 2^k-1 (k = 0..253) and p-1
#define LAYOUT_LONG_MIN 64u    // C segments longer than this are ranked by a block in shared memory
#define LAYOUT_LONG_MID 512u   // ... by a small block up to here, by a large one beyond
#define LAYOUT_LONG_MAX 4096u  // ... as long as they fit there (49 B per term)
#define C3_INLINE_MAX 255u  // longer bit-decomposition candidates go through the 2^i mod p table

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      err = std::string(#x) + ": " + cudaGetErrorString(e_);                       \
      return ECNE_E_CUDA;                                                          \
    }                                                                              \
  } while (0)

struct CoefView {  // read-only 32-byte coefficients of the uploaded rows, one 256-bit load each
  const fr::u256* p;
  __device__ __forceinline__ fr::u256 operator[](uint64_t t) const { return fr::ldg256(p + t); }
};
struct Raw {  // the problem as uploaded (explicit zeros included)
  uint32_t N, V;
  uint64_t nnz;
  const unsigned long long* seg;  // [3N+1]
  const uint32_t* col;
  CoefView coef;
};

__global__ void k_keep(Raw r, uint32_t* keep) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > r.nnz) return;
  keep[t] = (t < r.nnz && !fr::is_zero(r.coef[t])) ? 1u : 0u;
}
__global__ void k_seg(Raw r, const uint32_t* pos, uint32_t* seg_nz) {
  uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s > 3ull * r.N) return;
  seg_nz[s] = pos[r.seg[s]];
}

struct ABScan {
  uint32_t nA, nB, v, maxA, bkey;
  bool multi;
  fr::u256 slope_a, icpt_a, slope_b, icpt_b;
};
__device__ inline void scan_ab(const Raw& r, uint32_t row, ABScan& o, bool& bad) {
  o.nA = o.nB = o.v = o.maxA = o.bkey = 0;
  o.multi = false;
  o.slope_a = o.icpt_a = o.slope_b = o.icpt_b = fr::make_u256(0, 0, 0, 0);
  for (int f = 0; f < 2; ++f) {
    for (uint64_t t = r.seg[3ull * row + f]; t < r.seg[3ull * row + f + 1]; ++t) {
      uint32_t w = r.col[t];
      if (w < 1 || w > r.V) {
        bad = true;
        continue;
      }
      fr::u256 c = r.coef[t];
      if (fr::is_zero(c)) continue;
      if (f == 0) {
        o.nA++;
      } else {
        o.nB++;
        o.bkey = w;
      }
      if (w == 1) {
        if (f == 0)
          o.icpt_a = c;
        else
          o.icpt_b = c;
      } else {
        if (f == 0 && w > o.maxA) o.maxA = w;
        if (o.v == 0) o.v = w;
        if (o.v != w) {
          o.multi = true;
        } else if (f == 0) {
          o.slope_a = c;
        } else {
          o.slope_b = c;
        }
      }
    }
  }
}

struct CScan {
  uint32_t nC, stC, n_non1, x, n_one, n_mone, key_one, key_mone;
  bool key1_stored;
  fr::u256 c1, cx;
};
__device__ inline void scan_c(const Raw& r, uint32_t row, CScan& o, bool& bad) {
  o.nC = o.n_non1 = o.x = o.n_one = o.n_mone = o.key_one = o.key_mone = 0;
  o.key1_stored = false;
  o.c1 = o.cx = fr::make_u256(0, 0, 0, 0);
  uint64_t b = r.seg[3ull * row + 2], e = r.seg[3ull * row + 3];
  o.stC = (uint32_t)(e - b);
  for (uint64_t t = b; t < e; ++t) {
    uint32_t w = r.col[t];
    if (w < 1 || w > r.V) {
      bad = true;
      continue;
    }
    fr::u256 c = r.coef[t];
    if (w == 1) {
      o.key1_stored = true;
      o.c1 = c;
    }
    if (fr::is_zero(c)) continue;
    o.nC++;
    if (w != 1) {
      o.n_non1++;
      o.x = w;
      o.cx = c;
    }
    if (fr::is_one(c)) {
      o.n_one++;
      o.key_one = w;
    } else if (fr::is_minus_one(c)) {
      o.n_mone++;
      o.key_mone = w;
    }
  }
}

// exponent e if v == 2^e (as an integer), else -1
__device__ __forceinline__ int log2_exact(const fr::u256& v) {
  if (!fr::is_pow2(v)) return -1;
  return fr::bitlen(v) - 1;
}

// Bit-decomposition pattern test (:999-1013) for l <= C3_INLINE_MAX, where 2^i < p for every i used.
// returns bit0: values == {1} U {-2^i},  bit1: values == {-1} U {2^i}   (i = 0..l-2)
__device__ inline int c3_pattern_small(const Raw& r, uint32_t row, uint32_t l) {
  uint64_t b = r.seg[3ull * row + 2], e = r.seg[3ull * row + 3];
  unsigned long long m1[4] = {0, 0, 0, 0}, m2[4] = {0, 0, 0, 0};
  int ones = 0, mones = 0;
  bool ok1 = true, ok2 = true;
  for (uint64_t t = b; t < e; ++t) {
    fr::u256 c = r.coef[t];
    // T1 membership
    if (fr::is_one(c)) {
      ones++;
    } else {
      fr::u256 n = fr::neg(c);
      int ex = log2_exact(n);
      if (ex < 0 || (uint32_t)ex + 1 >= l || ((m1[ex >> 6] >> (ex & 63)) & 1))
        ok1 = false;
      else
        m1[ex >> 6] |= 1ULL << (ex & 63);
    }
    // T2 membership
    if (fr::is_minus_one(c)) {
      mones++;
    } else {
      int ex = log2_exact(c);
      if (ex < 0 || (uint32_t)ex + 1 >= l || ((m2[ex >> 6] >> (ex & 63)) & 1))
        ok2 = false;
      else
        m2[ex >> 6] |= 1ULL << (ex & 63);
    }
  }
  return ((ok1 && ones == 1) ? 1 : 0) | ((ok2 && mones == 1) ? 2 : 0);
}

struct Counters {
  unsigned int n2a, n2b, n_long, max_c, n_c3_long, bad;
  unsigned int n_longc;  // rows handed to k_classify_long
  unsigned int max_seg_c;  // longest stored C segment of a long row (sizes the shared memory of k_layout_long)
};

// one atomic per converged group of lanes instead of one per row (160 k rows count themselves as 2b)
__device__ __forceinline__ void warp_count(unsigned int* ctr) {
  const unsigned int am = __activemask();
  if ((threadIdx.x & 31u) == (unsigned int)(__ffs((int)am) - 1)) atomicAdd(ctr, (unsigned int)__popc(am));
}
#define CLASSIFY_LONG_C 64u  // rows whose stored C segment is longer get a warp (k_classify_long)
// everything after the two scans of a row: flags, aux, counters
__device__ __forceinline__ void classify_decide(const Raw& r, uint32_t row, const ABScan& ab, const CScan& cs, bool bad,
                                            uint32_t* rflags, RowAux* aux, Counters* cnt, uint32_t* c3_long);

// rows whose stored C segment gets a warp (k_classify_long): listed ahead, so that both classifiers run side by side
__global__ void k_longc(Raw r, Counters* cnt, uint32_t* longc) {
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row < r.N && r.seg[3ull * row + 3] - r.seg[3ull * row + 2] > CLASSIFY_LONG_C) longc[atomicAdd(&cnt->n_longc, 1u)] = row;
}
__global__ void k_classify(Raw r, uint32_t* rflags, RowAux* aux, Counters* cnt, uint32_t* c3_long) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= r.N) return;
  if (r.seg[3ull * row + 3] - r.seg[3ull * row + 2] > CLASSIFY_LONG_C) return;  // k_classify_long (listed by k_longc)
  bool bad = false;
  ABScan ab;
  CScan cs;
  scan_ab(r, row, ab, bad);
  scan_c(r, row, cs, bad);
  classify_decide(r, row, ab, cs, bad, rflags, aux, cnt, c3_long);
}

__device__ __forceinline__ void classify_long_row(const Raw& r, uint32_t row, uint32_t lane, uint32_t* rflags, RowAux* aux,
                                                  Counters* cnt, uint32_t* c3_long);
// One warp per row with a long C segment: the lanes stride the stored terms and their partial scans are
// merged (counts by sum, "last key with ..." by the largest term index), then lane 0 decides as above.
__global__ void k_classify_long(Raw r, uint32_t* rflags, RowAux* aux, Counters* cnt, uint32_t* c3_long,
                                const uint32_t* longc) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const unsigned int n_rows = *((volatile unsigned int*)&cnt->n_longc);
  for (uint32_t li = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; li < n_rows; li += nwarps) classify_long_row(r, longc[li], lane, rflags, aux, cnt, c3_long);
}
__device__ __forceinline__ void classify_long_row(const Raw& r, uint32_t row, uint32_t lane, uint32_t* rflags, RowAux* aux,
                                                  Counters* cnt, uint32_t* c3_long) {
  const uint64_t b = r.seg[3ull * row + 2], e = r.seg[3ull * row + 3];
  uint32_t nC = 0, n_non1 = 0, n_one = 0, n_mone = 0;
  unsigned long long last_x = 0, last_one = 0, last_mone = 0;  // (term index + 1) << 32 | wire
  bool bad = false, key1 = false;
  // (four terms per lane in flight: a 1025-term row is 33 dependent trips to DRAM otherwise)
  for (uint64_t t0 = b + lane; t0 < e; t0 += 128) {
    uint32_t wq[4];
    fr::u256 cq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint64_t t = t0 + 32u * q;
      if (t < e) {
        wq[q] = r.col[t];
        cq[q] = r.coef[t];
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint64_t t = t0 + 32u * q;
      if (t >= e) break;
      const uint32_t w = wq[q];
      if (w < 1 || w > r.V) {
        bad = true;
        continue;
      }
      const fr::u256 c = cq[q];
      if (w == 1) key1 = true;
      if (fr::is_zero(c)) continue;
      const unsigned long long tag = ((unsigned long long)(t - b + 1) << 32) | w;
      nC++;
      if (w != 1) {
        n_non1++;
        last_x = tag;
      }
      if (fr::is_one(c)) {
        n_one++;
        last_one = tag;
      } else if (fr::is_minus_one(c)) {
        n_mone++;
        last_mone = tag;
      }
    }
  }
  nC = __reduce_add_sync(0xffffffffu, nC);
  n_non1 = __reduce_add_sync(0xffffffffu, n_non1);
  n_one = __reduce_add_sync(0xffffffffu, n_one);
  n_mone = __reduce_add_sync(0xffffffffu, n_mone);
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long v;
    v = __shfl_xor_sync(0xffffffffu, last_x, o);
    last_x = v > last_x ? v : last_x;
    v = __shfl_xor_sync(0xffffffffu, last_one, o);
    last_one = v > last_one ? v : last_one;
    v = __shfl_xor_sync(0xffffffffu, last_mone, o);
    last_mone = v > last_mone ? v : last_mone;
  }
  bad = __any_sync(0xffffffffu, bad);
  key1 = __any_sync(0xffffffffu, key1);
  if (lane != 0) return;
  ABScan ab;
  scan_ab(r, row, ab, bad);
  CScan cs;
  cs.nC = nC;
  cs.stC = (uint32_t)(e - b);
  cs.n_non1 = n_non1;
  cs.x = (uint32_t)last_x;
  cs.n_one = n_one;
  cs.n_mone = n_mone;
  cs.key_one = (uint32_t)last_one;
  cs.key_mone = (uint32_t)last_mone;
  cs.key1_stored = key1;
  cs.c1 = cs.cx = fr::make_u256(0, 0, 0, 0);  // only read for rows with a single non-constant key (never long)
  classify_decide(r, row, ab, cs, bad, rflags, aux, cnt, c3_long);
}

__device__ __forceinline__ void classify_decide(const Raw& r, uint32_t row, const ABScan& ab, const CScan& cs, bool bad,
                                            uint32_t* rflags, RowAux* aux, Counters* cnt, uint32_t* c3_long) {
  if (bad) {
    cnt->bad = 1;
    rflags[row] = 0;
    return;
  }
  uint32_t rf = 0;
  RowAux a;
  a.w1 = a.w2 = a.w3 = a.w4 = a.w5 = a.val_idx = 0;
  a.rank_a = a.rank_b = 0;
  const bool linear = ab.nA == 0 && ab.nB == 0;
  if (linear) rf |= RF_LINEAR;
  if (cs.nC == 0) {
    rf |= RF_CEMPTY;
    // Case 2a statics
    if (ab.v == 0) {
      rf |= RF_2A_NOVAR;
    } else if (!ab.multi) {
      rf |= RF_2A;
      a.w1 = ab.v;
      if (fr::is_zero(ab.slope_a) || fr::is_zero(ab.slope_b)) {
        rf |= RF_2A_DIVZ;
      } else {
        // roots {0,1} without dividing: root == 0 <=> intercept == 0, root == 1 <=> intercept + slope == 0
        bool a0 = fr::is_zero(ab.icpt_a), b0 = fr::is_zero(ab.icpt_b);
        bool a1 = fr::is_zero(fr::add(ab.icpt_a, ab.slope_a));
        bool b1 = fr::is_zero(fr::add(ab.icpt_b, ab.slope_b));
        if ((a0 && b1) || (a1 && b0)) rf |= RF_2A_BOOL;
        atomicAdd(&cnt->n2a, 1u);
      }
    }
    // ABZ statics
    if (ab.nB == 1 && ab.nA <= 2) {
      rf |= RF_P3;
      a.w3 = ab.bkey;
      a.w4 = ab.maxA;
      if (ab.maxA == 0) rf |= RF_P3_DIVZ;
    }
  }
  if (linear) {
    if (cs.n_non1 == 1) {
      rf |= RF_2B;
      a.w1 = cs.x;
      warp_count(&cnt->n2b);
    }
    const bool inserted_zero = (rf & RF_2B) && !cs.key1_stored;  // c[1] read inserts a zero (:962)
    const uint32_t l = cs.nC;
    if (l > 0 && cs.stC == l && !inserted_zero) {
      if (l <= C3_INLINE_MAX) {
        int pat = c3_pattern_small(r, row, l);
        if (pat) {
          rf |= RF_C3;
          if (pat & 2) {
            rf |= RF_C3_FLIP;       // (:1001) is tested first
            a.w2 = cs.key_mone;     // coefficient 1 after negation
            a.w5 = cs.key_one;
          } else {
            a.w2 = cs.key_one;
          }
          if (l == 2) rf |= RF_C3_L2;
          if (l - 1 >= 254) rf |= RF_C3_TOPBIG;
        }
      } else {
        c3_long[atomicAdd(&cnt->n_c3_long, 1u)] = row;
      }
    }
    if (cs.nC < 3 && cs.stC == 2 && cs.n_one == 1 && cs.n_mone == 1) rf |= RF_4A;
    if (cs.nC < 4 && cs.stC == 3 && cs.n_one == 1 && cs.n_mone == 2 && cs.key_one == 1) rf |= RF_4B;
  }
  // IsZero pair statics (:1493-1536), row i with row i+1
  if (row + 1 < r.N && cs.nC == 2) {
    bool bad2 = false;
    ABScan nb;
    CScan nc;
    scan_ab(r, row + 1, nb, bad2);
    scan_c(r, row + 1, nc, bad2);
    if (!bad2 && nc.nC == 0 && nb.nB == 1 && nb.bkey != 1) {
      uint32_t vk = nb.bkey;
      bool ok = true;
      // every non-constant key of C_i is vk
      for (uint64_t t = r.seg[3ull * row + 2]; t < r.seg[3ull * row + 3]; ++t) {
        if (fr::is_zero(r.coef[t])) continue;
        uint32_t w = r.col[t];
        if (w != 1 && w != vk) ok = false;
      }
      // a_i == a_{i+1} as stored dictionaries (:1512)
      uint64_t b0 = r.seg[3ull * row], e0 = r.seg[3ull * row + 1];
      uint64_t b1 = r.seg[3ull * row + 3], e1 = r.seg[3ull * row + 4];
      if (e0 - b0 != e1 - b1) ok = false;
      for (uint64_t t = b0; ok && t < e0; ++t) {
        bool found = false;
        for (uint64_t u = b1; u < e1; ++u)
          if (r.col[u] == r.col[t] && fr::eq(r.coef[u], r.coef[t])) {
            found = true;
            break;
          }
        if (!found) ok = false;
      }
      if (ok) {
        rf |= RF_P4;
        a.w4 = vk;
      }
    }
  }
  uint32_t tot = ab.nA + ab.nB + cs.nC;
  if (tot > LONG_T) {
    rf |= RF_LONG;
    atomicAdd(&cnt->n_long, 1u);
    const uint32_t seg_c = (uint32_t)(r.seg[3ull * row + 3] - r.seg[3ull * row + 2]);
    if (seg_c > LAYOUT_LONG_MIN) atomicMax(&cnt->max_seg_c, seg_c);
  }
  if (cs.nC > C3_INLINE_MAX) atomicMax(&cnt->max_c, cs.nC);  // sizes the 2^i mod p table of the long candidates
  rflags[row] = rf;
  aux[row] = a;
}

// Long bit-decomposition candidates: one warp per row, exponents looked up in the sorted table of
// 2^i mod p (i < tab_n).  smem bitmask of seen exponents per warp.
struct Pow2Entry {
  fr::u256 v;
  uint32_t e;
  uint32_t pad[7];
};
__device__ inline int pow2_lookup(const Pow2Entry* tab, uint32_t n, const fr::u256& v) {
  // 2^i < p for i <= 253: a canonical value with one bit set is its own table entry, and below 2^254 the table holds
  // nothing else — the binary search is only for the wrapped powers (rows longer than 254 terms)
  const int bits = __popcll(v.v[0]) + __popcll(v.v[1]) + __popcll(v.v[2]) + __popcll(v.v[3]);
  if (bits == 1) {
    const int ex = v.v[0] ? __ffsll((long long)v.v[0]) - 1
                 : v.v[1] ? 63 + __ffsll((long long)v.v[1])
                 : v.v[2] ? 127 + __ffsll((long long)v.v[2])
                          : 191 + __ffsll((long long)v.v[3]);
    return (uint32_t)ex < n ? ex : -1;
  }
  if (n <= 254) return -1;
  int lo = 0, hi = (int)n - 1;
  while (lo <= hi) {
    int mid = (lo + hi) >> 1;
    int c = fr::cmp(tab[mid].v, v);
    if (c == 0) return (int)tab[mid].e;
    if (c < 0)
      lo = mid + 1;
    else
      hi = mid - 1;
  }
  return -1;
}
__global__ void k_classify_c3_long(Raw r, uint32_t* rflags, RowAux* aux, const uint32_t* list,
                                   uint32_t n_list, const Pow2Entry* tab, uint32_t tab_n,
                                   uint32_t mask_words) {
  // one block per row: every exponent beyond 253 costs a ten-step search in the table, and a 1025-term row is 33
  // dependent rounds of those for a single warp
  extern __shared__ unsigned int smem[];
  __shared__ unsigned int s_ones, s_mones, s_k_one, s_k_mone, s_dead1, s_dead2;
  if (blockIdx.x >= n_list) return;
  unsigned int* m1 = smem;
  unsigned int* m2 = m1 + mask_words;
  for (uint32_t i = threadIdx.x; i < 2 * mask_words; i += blockDim.x) m1[i] = 0;
  if (threadIdx.x == 0) s_ones = s_mones = s_k_one = s_k_mone = s_dead1 = s_dead2 = 0;
  __syncthreads();
  const uint32_t row = list[blockIdx.x];
  const uint64_t b = r.seg[3ull * row + 2], e = r.seg[3ull * row + 3];
  const uint32_t l = (uint32_t)(e - b);
  // (a form that has failed on one term is dead for the row: the others stop looking their exponents up)
  for (uint64_t t = b + threadIdx.x; t < e; t += blockDim.x) {
    // (atomics on both sides: the flags are read while other threads may set them — only to stop early, the verdict is
    // taken behind the barrier — and racecheck has nothing to report)
    const bool dead1 = atomicOr(&s_dead1, 0u) != 0, dead2 = atomicOr(&s_dead2, 0u) != 0;
    if (dead1 && dead2) break;
    const fr::u256 c = r.coef[t];
    if (fr::is_one(c)) {
      atomicAdd(&s_ones, 1u);
      atomicMax(&s_k_one, r.col[t]);
    } else if (!dead1) {
      const int ex = pow2_lookup(tab, tab_n, fr::neg(c));
      if (ex < 0 || (uint32_t)ex + 1 >= l || (atomicOr(m1 + (ex >> 5), 1u << (ex & 31)) & (1u << (ex & 31)))) atomicExch(&s_dead1, 1u);
    }
    if (fr::is_minus_one(c)) {
      atomicAdd(&s_mones, 1u);
      atomicMax(&s_k_mone, r.col[t]);
    } else if (!dead2) {
      const int ex = pow2_lookup(tab, tab_n, c);
      if (ex < 0 || (uint32_t)ex + 1 >= l || (atomicOr(m2 + (ex >> 5), 1u << (ex & 31)) & (1u << (ex & 31)))) atomicExch(&s_dead2, 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const bool ok1 = !s_dead1 && s_ones == 1, ok2 = !s_dead2 && s_mones == 1;
    if (ok1 || ok2) {
      uint32_t rf = rflags[row] | RF_C3;
      RowAux a = aux[row];
      if (ok2) {
        rf |= RF_C3_FLIP;
        a.w2 = s_k_mone;
        a.w5 = s_k_one;
      } else {
        a.w2 = s_k_one;
      }
      if (l - 1 >= 254) rf |= RF_C3_TOPBIG;
      rflags[row] = rf;
      aux[row] = a;
    }
  }
}

// roots (2a) and t (2b): the -intercept/slope divisions (:919-920, :961-964)
__global__ void k_values(Raw r, const uint32_t* rflags, RowAux* aux, fr::u256* roots, fr::u256* tvals,
                         unsigned int* next2a, unsigned int* next2b) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= r.N) return;
  uint32_t rf = rflags[row];
  bool bad = false;
  if ((rf & RF_2A) && !(rf & RF_2A_DIVZ)) {
    ABScan ab;
    scan_ab(r, row, ab, bad);
    unsigned int i = atomicAdd(next2a, 1u);
    roots[2 * i] = fr::neg_div(ab.icpt_a, ab.slope_a);
    roots[2 * i + 1] = fr::neg_div(ab.icpt_b, ab.slope_b);
    aux[row].val_idx = i;
  }
  if (rf & RF_2B) {
    CScan cs;
    scan_c(r, row, cs, bad);
    unsigned int i = atomicAdd(next2b, 1u);
    tvals[N_CONST + i] = fr::neg_div(cs.c1, cs.cx);
    aux[row].val_idx = i;
  }
}

__device__ __forceinline__ fr::u256 mag_of(const fr::u256& c, bool flipped) {
  fr::u256 x = flipped ? fr::neg(c) : c;
  if (fr::cmp(x, fr::fold_threshold()) > 0) {
    fr::u256 o;
    fr::sub_cc(o, fr::modulus(), x);
    return o;
  }
  return x;
}

// Scatter the non-zero terms into the sweep layout.  C terms of linear rows are placed by their
// rank in (|fold(coef)|, wire) order so that Case 5 (:1265) walks them sorted.
// Short rows (everything but RF_LONG): one thread per row walks its three segments — no search for the segment of a
// term, the row's flags read once, and a linear row's C terms ranked among at most LONG_T kept ones.
__device__ __forceinline__ void layout_term(const Raw& r, const uint32_t* keep, const uint32_t* pos, const uint32_t* seg_nz,
                                            uint32_t s, uint64_t seg_b, uint64_t seg_e, uint64_t t, bool ranked, bool fl,
                                            uint32_t* col, fr::u256* coef, uint8_t* nontriv) {
  if (!keep[t]) return;
  const fr::u256 c = r.coef[t];
  const uint32_t w = r.col[t];
  uint32_t dst = pos[t];
  if (ranked) {
    const fr::u256 m = mag_of(c, fl);
    uint32_t rank = 0;
    for (uint64_t u = seg_b; u < seg_e; ++u) {
      if (u == t || !keep[u]) continue;
      const int cm = fr::cmp(mag_of(r.coef[u], fl), m);
      if (cm < 0 || (cm == 0 && r.col[u] < w)) ++rank;
    }
    dst = seg_nz[s] + rank;
  }
  col[dst] = w;
  fr::stg256(coef + dst, c);
  // (a set-once flag: wire 1 — the constant — is in a third of the rows, and a million byte stores to one address
  // queue up in one L2 slice; a stale zero from the L1 only repeats the store)
  if (!nontriv[w]) nontriv[w] = 1;
}
__global__ void k_layout(Raw r, const uint32_t* keep, const uint32_t* pos, const uint32_t* seg_nz,
                         const uint32_t* rflags, uint32_t* col, fr::u256* coef, uint8_t* nontriv) {
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= r.N) return;
  const uint32_t rf = rflags[row];
  if (rf & RF_LONG) return;  // k_layout_long_rows / k_layout_long
  const bool fl = (rf & RF_C3_FLIP) != 0;
  uint64_t seg_b = r.seg[3ull * row];
#pragma unroll 1
  for (uint32_t form = 0; form < 3; ++form) {
    const uint32_t s = 3 * row + form;
    const uint64_t seg_e = r.seg[s + 1];
    const bool ranked = form == 2 && (rf & RF_LINEAR);
    for (uint64_t t = seg_b; t < seg_e; ++t)
      layout_term(r, keep, pos, seg_nz, s, seg_b, seg_e, t, ranked, fl, col, coef, nontriv);
    seg_b = seg_e;
  }
}
// Long rows: one warp per row, lanes striding the terms of a segment.  The C segment of a linear row between
// LAYOUT_LONG_MIN and LAYOUT_LONG_MAX terms is left to k_layout_long (sorted in shared memory).
__global__ void k_layout_long_rows(Raw r, const uint32_t* keep, const uint32_t* pos, const uint32_t* seg_nz,
                                   const uint32_t* rflags, const uint32_t* long_rows, uint32_t n_long, uint32_t* col,
                                   fr::u256* coef, uint8_t* nontriv) {
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
  if (wid >= n_long) return;
  const uint32_t row = long_rows[wid];
  const uint32_t rf = rflags[row];
  const bool fl = (rf & RF_C3_FLIP) != 0;
  for (uint32_t form = 0; form < 3; ++form) {
    const uint32_t s = 3 * row + form;
    const uint64_t seg_b = r.seg[s], seg_e = r.seg[s + 1];
    const bool ranked = form == 2 && (rf & RF_LINEAR);
    if (ranked && seg_e - seg_b > LAYOUT_LONG_MIN && seg_e - seg_b <= LAYOUT_LONG_MAX) continue;
    for (uint64_t t = seg_b + lane; t < seg_e; t += 32)
      layout_term(r, keep, pos, seg_nz, s, seg_b, seg_e, t, ranked, fl, col, coef, nontriv);
  }
}
// Ascending-only bitonic network over a permutation of n items (the first step of every merge mirrors, the others are
// half-cleaners): positions beyond n act as +infinity and are never touched, so n needs no padding.  `less(a, b)`
// compares two ITEMS; all threads of the block call this.
template <class Less>
__device__ __forceinline__ void block_sort_perm(uint32_t* perm, uint32_t n, Less less) {
  uint32_t np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (uint32_t k = 2; k <= np2; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      const bool mirror = j == (k >> 1);
      for (uint32_t t = threadIdx.x; t < (np2 >> 1); t += blockDim.x) {
        const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // t-th pair of the step: bit j of i is clear
        const uint32_t q = mirror ? (i ^ (k - 1)) : (i | j);
        if (q < n) {
          const uint32_t a = perm[i], b = perm[q];
          if (less(b, a)) {
            perm[i] = b;
            perm[q] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}
// 256-bit values of a block in shared memory, one array per limb: a compare reads 8-byte words at random items, which
// spread over the banks (32-byte structs put every item's limb k in one of four bank groups: an 8-way conflict).
struct LimbsSoA {
  unsigned long long* l[4];
  __device__ __forceinline__ void carve(unsigned long long* base, uint32_t cap) {
    for (int k = 0; k < 4; ++k) l[k] = base + (size_t)k * cap;
  }
  __device__ __forceinline__ void put(uint32_t i, const fr::u256& x) const {
    for (int k = 0; k < 4; ++k) l[k][i] = x.v[k];
  }
  __device__ __forceinline__ int cmp(uint32_t a, uint32_t b) const {
    for (int k = 3; k >= 0; --k) {
      const unsigned long long x = l[k][a], y = l[k][b];
      if (x != y) return x < y ? -1 : 1;
    }
    return 0;
  }
};
// 64-bit key that orders 256-bit magnitudes: the value itself below 2^55 (exact: equal keys are equal values), else
// bit length and the 55 bits behind the leading one (equal keys: the limbs decide).
__device__ __forceinline__ unsigned long long order_key(const fr::u256& m) {
  int top = 3;
  while (top > 0 && m.v[top] == 0) --top;
  const unsigned long long hi = top == 3 ? m.v[3] : top == 2 ? m.v[2] : top == 1 ? m.v[1] : m.v[0];
  const unsigned long long below = top == 3 ? m.v[2] : top == 2 ? m.v[1] : top == 1 ? m.v[0] : 0ull;
  if (hi == 0) return 0;
  const int lz = __clzll((long long)hi);
  const int len = 64 * top + 64 - lz;  // bit length
  if (len <= 55) return hi;            // (top == 0 here)
  // 56 bits from the leading one: `hi` shifted to the top of a word, filled from the limb below
  const unsigned long long w = lz ? (hi << lz) | (below >> (64 - lz)) : hi;
  return ((unsigned long long)len << 55) | ((w >> 8) & ((1ull << 55) - 1));
}
// The C segment of one long linear row: magnitudes and wires go to shared memory once and the network sorts an index
// permutation by (dropped?, |fold(coef)|, wire) — the kept terms come out in exactly the order k_layout's rank
// computation gives, in O(n log^2 n) on-chip compares.
__global__ void k_layout_long(Raw r, const uint32_t* keep, const uint32_t* seg_nz, const uint32_t* rflags,
                              const uint32_t* long_rows, uint32_t n_long, uint32_t len_lo, uint32_t len_hi, uint32_t* col,
                              fr::u256* coef, uint8_t* nontriv) {
  extern __shared__ unsigned long long sm_u64[];
  if (blockIdx.x >= n_long) return;
  const uint32_t row = long_rows[blockIdx.x];
  const uint32_t rf = rflags[row];
  if (!(rf & RF_LINEAR)) return;
  const uint32_t s = 3 * row + 2;
  const uint64_t b = r.seg[s], e = r.seg[s + 1];
  const uint32_t n = (uint32_t)(e - b);
  if (n <= len_lo || n > len_hi) return;  // (this launch's length class: the shared memory is sized for len_hi)
  LimbsSoA mag;
  mag.carve(sm_u64, n);
  unsigned long long* key = sm_u64 + 4 * (size_t)n;
  uint32_t* wv = reinterpret_cast<uint32_t*>(key + n);
  uint32_t* idx = wv + n;
  uint8_t* kp = reinterpret_cast<uint8_t*>(idx + n);
  const bool fl = (rf & RF_C3_FLIP) != 0;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    idx[i] = i;
    const fr::u256 m = mag_of(r.coef[b + i], fl);
    mag.put(i, m);
    const uint32_t kept = keep[b + i];
    key[i] = kept ? order_key(m) : ~0ull;  // dropped zeros last (no key reaches 2^63)
    wv[i] = r.col[b + i];
    kp[i] = (uint8_t)kept;
  }
  __syncthreads();
  auto less = [&](uint32_t x, uint32_t y) {  // kept terms first (dropped zeros last), then (magnitude, wire, position)
    const unsigned long long kx = key[x], ky = key[y];
    if (kx != ky) return kx < ky;
    if (kx == ~0ull) return x < y;
    if (kx >> 55) {  // not exact: the limbs
      const int cm = mag.cmp(x, y);
      if (cm != 0) return cm < 0;
    }
    const uint32_t wx = wv[x], wy = wv[y];
    if (wx != wy) return wx < wy;
    return x < y;
  };
  // circom writes a sum in wire order, and the weights of a bit decomposition grow with the wires: many segments are in
  // (|fold(coef)|, wire) order already, with nothing dropped — one pass of neighbour compares instead of the network
  bool in_order = true;
  for (uint32_t i = threadIdx.x; i + 1 < n; i += blockDim.x) in_order = in_order && less(i, i + 1);
  if (!__syncthreads_and(in_order)) block_sort_perm(idx, n, less);
  for (uint32_t p = threadIdx.x; p < n; p += blockDim.x) {
    const uint32_t el = idx[p];
    if (!kp[el]) continue;
    const uint32_t dst = seg_nz[s] + p;  // kept terms occupy the first positions of the sorted order
    col[dst] = wv[el];
    fr::stg256(coef + dst, r.coef[b + el]);
    if (!nontriv[wv[el]]) nontriv[wv[el]] = 1;
  }
}
// 32-byte sweep records: flags + up to 6 inline wires (A u B first, then C)
__global__ void k_rowrec(uint32_t N, const uint32_t* seg, const uint32_t* col, uint32_t* rflags, RowRec* rec) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= N) return;
  uint32_t rf = rflags[row];
  const uint32_t s0 = seg[3 * row], s2 = seg[3 * row + 2], s3 = seg[3 * row + 3];
  RowRec r;
  r.pad = 0;
  for (int j = 0; j < ROWREC_INLINE; ++j) r.c[j] = 1;
  const uint32_t tot = s3 - s0;
  if (tot <= ROWREC_INLINE) {
    r.inl = 1;
    r.nAB = (uint8_t)(s2 - s0);
    r.nC = (uint8_t)(s3 - s2);
    for (uint32_t t = s0; t < s3; ++t) r.c[t - s0] = col[t];
    // fast-path rows: no bound pattern, or exactly "x = const" (2B alone), or exactly "x - y = 0"
    // (4A, which always carries the l == 2 Case-3 flags as well)
    const uint32_t pat = rf & (RF_2B | RF_C3 | RF_4A | RF_4B);
    if (!(rf & RF_LONG) && (pat == 0 || pat == RF_2B || pat == (RF_4A | RF_C3))) rf |= RF_FAST;
  } else {
    r.inl = 0;
    r.nAB = r.nC = 0;
  }
  r.rf = rf;
  rflags[row] = rf;
  rec[row] = r;
}
__global__ void k_mark(const uint32_t* list, uint32_t n, uint8_t* nontriv) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) nontriv[list[i]] = 1;
}
__global__ void k_long_rows(uint32_t N, const uint32_t* rflags, uint32_t* out, unsigned int* n) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row < N && (rflags[row] & RF_LONG)) out[atomicAdd(n, 1u)] = row;
}

// long rows keep their index into long_rows[] in the (otherwise unused) first inline slot
__global__ void k_long_index(uint32_t n, const uint32_t* long_rows, RowRec* rec) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rec[long_rows[i]].c[0] = i;
}
// static lists of the rows with the P3 (ABZ) and P4 (IsZero pair) shapes
__global__ void k_phase_rows(uint32_t N, const uint32_t* rflags, uint32_t* p3, uint32_t* p4, unsigned int* n) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= N) return;
  uint32_t rf = rflags[row];
  if (rf & RF_P3) p3[atomicAdd(n + 0, 1u)] = row;
  if (rf & RF_P4) p4[atomicAdd(n + 1, 1u)] = row;
}
#define INV_LONG_MIN 32u
// wire -> rows index over the non-zero terms (the constant wire 1 never changes state: left out)
__global__ void k_inv_count(uint32_t nnz, const uint32_t* col, uint32_t* cnt) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nnz && col[t] != 1) atomicAdd(cnt + col[t], 1u);
}
__global__ void k_inv_fill(uint32_t N, const uint32_t* seg, const uint32_t* col, const uint32_t* inv_ptr,
                           uint32_t* cursor, uint32_t* inv_row) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= N) return;
  const uint32_t s0 = seg[3 * row], s3 = seg[3 * row + 3];
  if (s3 - s0 > INV_LONG_MIN) return;  // k_inv_fill_long
  for (uint32_t t = s0; t < s3; ++t) {
    const uint32_t w = col[t];
    if (w == 1) continue;
    bool dup = false;  // a wire that occurs in several forms of a short row is listed once
    if (s3 - s0 <= 8)
      for (uint32_t u = s0; u < t; ++u) dup |= col[u] == w;
    if (dup) continue;
    inv_row[inv_ptr[w] + atomicAdd(cursor + w, 1u)] = row;
  }
}
// rows with many terms: one block per row, threads striding its terms
__global__ void k_inv_fill_long(const uint32_t* long_rows, uint32_t n_long, const uint32_t* seg, const uint32_t* col,
                                const uint32_t* inv_ptr, uint32_t* cursor, uint32_t* inv_row) {
  if (blockIdx.x >= n_long) return;
  const uint32_t row = long_rows[blockIdx.x];
  const uint32_t s0 = seg[3 * row], s3 = seg[3 * row + 3];
  if (s3 - s0 <= INV_LONG_MIN) return;
  for (uint32_t t = s0 + threadIdx.x; t < s3; t += blockDim.x) {
    const uint32_t w = col[t];
    if (w == 1) continue;
    inv_row[inv_ptr[w] + atomicAdd(cursor + w, 1u)] = row;
  }
}

__global__ void k_inv_head(uint32_t nw, const uint32_t* inv_ptr, const uint32_t* filled, const uint32_t* inv_row,
                           uint4* head) {
  uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  const uint32_t lo = inv_ptr[w], n = filled[w];
  uint4 h;
  h.x = n;
  h.y = n > 0 ? inv_row[lo] : 0xffffffffu;
  h.z = n > 1 ? inv_row[lo + 1] : 0xffffffffu;
  h.w = n > 2 ? inv_row[lo + 2] : 0xffffffffu;
  head[w] = h;
}

// ---- bound-value table: sort the candidates, rank the distinct values ---------------------------
__global__ void k_consts(fr::u256* cand) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < 254)
    cand[k] = fr::pow2m1((int)k);
  else if (k == 254)
    fr::sub_cc(cand[k], fr::modulus(), fr::make_u256(1, 0, 0, 0));
}
__global__ void k_iota(uint32_t* idx, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = i;
}
__global__ void k_distinct(const fr::u256* cand, const uint32_t* idx, uint32_t n, uint32_t* flag) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == 0 || !fr::eq(cand[idx[i]], cand[idx[i - 1]])) ? 1u : 0u;
}
__global__ void k_ranks(const fr::u256* cand, const uint32_t* idx, const uint32_t* incl, uint32_t n,
                        uint32_t* rank_of, fr::u256* table) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t rk = incl[i] - 1;
  rank_of[idx[i]] = rk;
  if (i == 0 || incl[i] != incl[i - 1]) table[rk] = cand[idx[i]];
}
__global__ void k_fill_ranks(uint32_t N, const uint32_t* rflags, RowAux* aux, const uint32_t* rank_of,
                             const uint32_t* seg_nz) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= N) return;
  uint32_t rf = rflags[row];
  if (rf & RF_2B) aux[row].rank_a = rank_of[N_CONST + aux[row].val_idx];
  if ((rf & RF_C3) && !(rf & RF_C3_TOPBIG)) {
    uint32_t l = seg_nz[3 * row + 3] - seg_nz[3 * row + 2];
    aux[row].rank_b = rank_of[l - 1];  // 2^(l-1) - 1
  }
}

struct U256Less {  // order of the candidate bound values (ties: any order, equal values share a rank)
  __device__ __forceinline__ bool operator()(const fr::u256& a, const fr::u256& b) const { return fr::cmp(a, b) < 0; }
};


// ---- sample sort of the candidate bound values ------------------------------------------------------------------
// ecdsa has 160 k candidates, almost all distinct: a merge sort of that many 32-byte keys is ten latency-bound passes
// (0.32 ms, the longest chain of the set-up).  Here: a hashed sample is sorted by one block and gives B - 1 splitters;
// every value finds its bucket by a search in shared memory and takes a slot in it; one block per bucket sorts it in
// shared memory.  Four launches; the order is (value, index), so equal values do not pile up in one bucket.  A bucket
// that does not fit the shared memory is sorted by the same network in global memory (slow, correct, never seen).
#define SS_MIN 16384u      // below: the library merge sort (a few passes)
#define SS_MAX 524288u     // above: the library merge sort (bandwidth-bound there, and the sample would not fit one block)
#define SS_BUCKET_CAP 2048u
#define SS_OVERSAMPLE 4u
struct SsKey {
  fr::u256 v;
  uint32_t i;
};
__device__ __forceinline__ bool ss_less(const fr::u256& a, uint32_t ia, const fr::u256& b, uint32_t ib) {
  const int c = fr::cmp(a, b);
  return c < 0 || (c == 0 && ia < ib);
}
__device__ __forceinline__ uint32_t ss_hash(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
// one block: sort a hashed sample of n_s values (one per stratum), keep every SS_OVERSAMPLE-th as a splitter
__global__ void k_ss_sample(const fr::u256* cand, uint32_t n, uint32_t n_s, SsKey* split) {
  extern __shared__ unsigned long long ss_smem[];
  LimbsSoA val;
  val.carve(ss_smem, n_s);
  unsigned long long* key = ss_smem + 4 * (size_t)n_s;
  uint32_t* gi = reinterpret_cast<uint32_t*>(key + n_s);
  uint32_t* perm = gi + n_s;
  const uint32_t stride = n / n_s;  // >= 2
  for (uint32_t j = threadIdx.x; j < n_s; j += blockDim.x) {
    const uint32_t at = j * stride + ss_hash(j) % stride;
    const fr::u256 v = fr::ldg256(cand + at);
    val.put(j, v);
    key[j] = order_key(v);
    gi[j] = at;
    perm[j] = j;
  }
  __syncthreads();
  block_sort_perm(perm, n_s, [&](uint32_t x, uint32_t y) {
    const unsigned long long kx = key[x], ky = key[y];
    if (kx != ky) return kx < ky;
    if (kx >> 55) {
      const int c = val.cmp(x, y);
      if (c != 0) return c < 0;
    }
    return gi[x] < gi[y];
  });
  for (uint32_t b = threadIdx.x + 1; b < n_s / SS_OVERSAMPLE; b += blockDim.x) {
    const uint32_t e = perm[b * SS_OVERSAMPLE];
    fr::u256 v;
    for (int k = 0; k < 4; ++k) v.v[k] = val.l[k][e];
    split[b - 1].v = v;
    split[b - 1].i = gi[e];
  }
}
// bucket of every value (number of splitters <= it) and its slot there
__global__ void k_ss_count(const fr::u256* cand, uint32_t n, const SsKey* split, uint32_t n_b, uint32_t* bkt,
                           uint32_t* slot, unsigned int* counts) {
  extern __shared__ unsigned long long ss_smem[];
  fr::u256* sv = reinterpret_cast<fr::u256*>(ss_smem);
  uint32_t* si = reinterpret_cast<uint32_t*>(sv + n_b);
  unsigned int* hist = si + n_b;  // n_b counters, then their global bases
  unsigned int* base = hist + n_b;
  for (uint32_t b = threadIdx.x; b < n_b; b += blockDim.x) {
    if (b + 1 < n_b) {
      sv[b] = split[b].v;
      si[b] = split[b].i;
    }
    hist[b] = 0;
  }
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t b = 0, local = 0;
  if (i < n) {
    const fr::u256 v = fr::ldg256(cand + i);
    uint32_t lo = 0, hi = n_b - 1;  // splitters [0, lo) are <= v, [hi, n_b - 1) are > v
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (ss_less(v, i, sv[mid], si[mid]))
        hi = mid;
      else
        lo = mid + 1;
    }
    b = lo;
    local = atomicAdd(hist + b, 1u);
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < n_b; k += blockDim.x) base[k] = hist[k] ? atomicAdd(counts + k, hist[k]) : 0;
  __syncthreads();
  if (i < n) {
    bkt[i] = b;
    slot[i] = base[b] + local;
  }
}
// exclusive scan of the bucket sizes (one block; n_b <= 1024)
__global__ void k_ss_starts(const unsigned int* counts, uint32_t n_b, uint32_t* starts) {
  __shared__ uint32_t sc[1024];
  const uint32_t t = threadIdx.x;
  sc[t] = t < n_b ? counts[t] : 0;
  __syncthreads();
  for (uint32_t d = 1; d < 1024; d <<= 1) {
    const uint32_t x = t >= d ? sc[t - d] : 0;
    __syncthreads();
    sc[t] += x;
    __syncthreads();
  }
  if (t < n_b) starts[t] = sc[t] - counts[t];
  if (t == 0) starts[n_b] = sc[1023];
}
__global__ void k_ss_scatter(const fr::u256* cand, uint32_t n, const uint32_t* bkt, const uint32_t* slot,
                             const uint32_t* starts, fr::u256* val2, uint32_t* gi2) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t d = starts[bkt[i]] + slot[i];
  fr::stg256(val2 + d, fr::ldg256(cand + i));
  gi2[d] = i;
}
// one block per bucket: sorted indices of the bucket's values into idx[starts[b] ...)
__global__ void k_ss_bucket(const fr::u256* val2, const uint32_t* gi2, const uint32_t* starts, uint32_t* perm_scratch,
                            uint32_t* idx, uint32_t cap) {  // cap <= SS_BUCKET_CAP (smaller only to test the global path)
  extern __shared__ unsigned long long ss_smem[];
  const uint32_t lo = starts[blockIdx.x], n = starts[blockIdx.x + 1] - lo;
  if (n == 0) return;
  if (n <= cap) {
    LimbsSoA val;
    val.carve(ss_smem, SS_BUCKET_CAP);
    unsigned long long* key = ss_smem + 4 * (size_t)SS_BUCKET_CAP;
    uint32_t* gi = reinterpret_cast<uint32_t*>(key + SS_BUCKET_CAP);
    uint32_t* perm = gi + SS_BUCKET_CAP;
    for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) {
      const fr::u256 v = fr::ldg256(val2 + lo + j);
      val.put(j, v);
      key[j] = order_key(v);
      gi[j] = gi2[lo + j];
      perm[j] = j;
    }
    __syncthreads();
    block_sort_perm(perm, n, [&](uint32_t x, uint32_t y) {
      const unsigned long long kx = key[x], ky = key[y];
      if (kx != ky) return kx < ky;
      if (kx >> 55) {
        const int c = val.cmp(x, y);
        if (c != 0) return c < 0;
      }
      return gi[x] < gi[y];
    });
    for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) idx[lo + j] = gi[perm[j]];
  } else {  // (global memory: every step goes through the L2)
    uint32_t* perm = perm_scratch + lo;
    const fr::u256* v = val2 + lo;
    const uint32_t* g = gi2 + lo;
    for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) perm[j] = j;
    __syncthreads();
    block_sort_perm(perm, n, [&](uint32_t x, uint32_t y) { return ss_less(v[x], g[x], v[y], g[y]); });
    for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) idx[lo + j] = g[perm[j]];
  }
}

template <class T>
static cudaError_t h2d(T* dst, const T* src, size_t n, cudaStream_t s) {
  if (!n) return cudaSuccess;
  return cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, s);
}
static inline unsigned int nb(uint64_t n, unsigned int t) { return (unsigned int)((n + t - 1) / t); }

// ---- the disjoint sets of equal wires (:634-678), secp_solve only -----------------------------------------------------
// `DataStructures.IntDisjointSet(num_variables)` + one pushed element per distinct constant: rows `x - y = 0` (stored
// C values exactly {1, p-1}, A and B without non-zero keys) unite their two keys, rows `a x = b` (two stored C keys,
// one of them the constant wire) unite x with the element of the value b/a — here node V + 1 + rank(value), equal
// values share a rank.  Lock-free union-find: the larger root is linked under the smaller with a CAS.  Its only
// readers are the six in_same_set calls per (BigMultModP, BigLessThan) pair (:761-765), whose only observable effect
// is the BoundsError of `constraint_j[3][1]` when the sets agree and the BigLessThan has no output (:768); building
// the sets can raise BoundsErrors of its own on rows with stored zeros (`l[1]` / `l[2]`, :660-662).
__device__ __forceinline__ uint32_t dsu_find(uint32_t* parent, uint32_t x) {
  while (true) {
    const uint32_t p = *((volatile uint32_t*)(parent + x));
    if (p == x) return x;
    x = p;
  }
}
__device__ __forceinline__ void dsu_union(uint32_t* parent, uint32_t a, uint32_t b) {
  while (true) {
    a = dsu_find(parent, a);
    b = dsu_find(parent, b);
    if (a == b) return;
    if (a < b) {
      const uint32_t t = a;
      a = b;
      b = t;
    }
    if (atomicCAS(parent + a, a, b) == a) return;  // a was still a root: now under the smaller root b
  }
}
__global__ void k_dsu_union(Raw r, const uint32_t* rflags, const RowAux* aux, uint32_t* parent, unsigned int* flags) {
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= r.N) return;
  const uint32_t rf = rflags[row];
  if (!(rf & RF_LINEAR)) return;                                        // (:640) nonzeroKeys of a and b are empty
  if (r.seg[3ull * row + 3] - r.seg[3ull * row + 2] != 2) return;       // length(eq.c): STORED keys (:641)
  bool bad = false;
  CScan cs;
  scan_c(r, row, cs, bad);
  if (bad) return;
  if (cs.n_one == 1 && cs.n_mone == 1) {  // sorted values == [1, p - 1] (:642-649)
    dsu_union(parent, cs.key_one, cs.key_mone);
    return;
  }
  if (cs.nC == 0 || (cs.nC == 1 && cs.n_non1 == 0)) {  // `l[1]` of an empty list / `l[2]` of [1] (:660-662)
    flags[0] = 1;
    return;
  }
  if (cs.nC == 2 && cs.n_non1 == 1)  // one key is the constant wire (:656-657): x ~ value (rank of -c1/cx, k_fill_ranks)
    dsu_union(parent, cs.x, r.V + 1u + aux[row].rank_a);
}
__global__ void k_dsu_pairs(uint32_t n_sp, const int32_t* kind, const uint32_t* in_ptr, const uint32_t* in,
                            const uint32_t* out_ptr, uint32_t* parent, unsigned int* flags) {
  const uint64_t pr = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pr >= (uint64_t)n_sp * n_sp) return;
  const uint32_t i = (uint32_t)(pr / n_sp), j = (uint32_t)(pr % n_sp);
  if (kind[i] != ECNE_SPECIAL_BIGMULTMODP || kind[j] != ECNE_SPECIAL_BIGLESSTHAN) return;
  if (in_ptr[i + 1] - in_ptr[i] < 9 || in_ptr[j + 1] - in_ptr[j] < 6) return;  // BoundsError raised by the solve kernel
  if (out_ptr[j + 1] != out_ptr[j]) return;                                     // constraint_j[3][1] exists
  bool same = true;
  for (uint32_t k = 0; k < 6; ++k)
    same &= dsu_find(parent, in[in_ptr[i] + 3 + k]) == dsu_find(parent, in[in_ptr[j] + k]);
  if (same) flags[1] = 1;
}

// ---- the rows of a host problem into device arrays of the on-disk layout ------------------------------------------
// Full form: three copies.  Compact form (include/ecne_abi.h, `coef == NULL`): one class byte per term and the 32-byte
// values of the terms that are not 0, 1 or p - 1 cross PCIe; two kernels write the same `coef` array the full form
// would have been copied into, so nothing downstream knows the difference.
__global__ void k_expand_class(const uint8_t* cls, uint64_t nnz, fr::u256* coef, unsigned int* chk) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int n3 = 0;
  if (t < nnz) {
    const uint8_t c = cls[t];
    if (c == 0) {
      fr::stg256(coef + t, fr::make_u256(0, 0, 0, 0));
    } else if (c == 1) {
      fr::stg256(coef + t, fr::make_u256(1, 0, 0, 0));
    } else if (c == 2) {
      fr::u256 m;
      fr::sub_cc(m, fr::modulus(), fr::make_u256(1, 0, 0, 0));
      fr::stg256(coef + t, m);
    } else if (c == 3) {
      n3 = 1;
    } else {
      chk[1] = 1;  // no such class
    }
  }
  n3 = __reduce_add_sync(0xffffffffu, n3);
  if ((threadIdx.x & 31u) == 0 && n3) atomicAdd(chk, n3);  // number of class-3 terms: must equal n_coef_other
}
__global__ void k_expand_other(const fr::u256* other, const uint32_t* term, uint64_t n_other, uint64_t nnz,
                               const uint8_t* cls, fr::u256* coef, unsigned int* chk) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_other) return;
  const uint32_t t = term[j];
  if (t >= nnz || cls[t] != 3 || (j > 0 && term[j - 1] >= t)) {
    chk[1] = 1;  // not a class-3 term, or the list is not strictly ascending
    return;
  }
  fr::stg256(coef + t, fr::ldg256(other + j));
}
__global__ void k_widen_seg(const uint32_t* seg32, uint64_t n, unsigned long long* seg) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) seg[i] = seg32[i];
}

UploadShard& upload_shard() {
  static UploadShard u;
  return u;
}
// host -> device copy of one array of the problem: whole, or this rank's slice + the all-gather of the ranks' slices
// One array of the problem to the device.  Unsharded: the slices are only QUEUED on the staging pool (the arrays of a
// problem cross back to back, no thread is started or joined per array) — put_done() before anything reads them.
static cudaError_t put(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  UploadShard& U = upload_shard();
  if (U.world <= 1 || !U.allgather || bytes < ((size_t)1 << 20)) return staged_h2d_async(dst, src, bytes, s);
  const size_t slice = ((bytes + (size_t)U.world - 1) / (size_t)U.world + 255) & ~(size_t)255;
  const size_t off = (size_t)U.rank * slice;
  if (off < bytes) {
    cudaError_t e = staged_h2d((char*)dst + off, (const char*)src + off, std::min(slice, bytes - off), s);
    if (e != cudaSuccess) return e;
  }
  return U.allgather(dst, slice, s);
}
static cudaError_t put_done() { return staged_flush(); }

uint64_t problem_nnz(const ecne_problem_t* p) {
  return p->seg_ptr ? p->seg_ptr[3 * p->n_rows] : p->seg_ptr32[3 * p->n_rows];
}
int problem_rows_ok(const ecne_problem_t* p, std::string& err) {
  if (!p || (!p->seg_ptr && !p->seg_ptr32)) {
    err = "null problem arrays";
    return ECNE_E_BADARG;
  }
  if (p->n_rows == 0) return ECNE_OK;
  if (!p->col || (!p->coef && !p->coef_class)) {
    err = "null problem arrays";
    return ECNE_E_BADARG;
  }
  if (!p->coef && p->n_coef_other && (!p->coef_other || !p->coef_other_term)) {
    err = "compact coefficients: coef_other / coef_other_term missing";
    return ECNE_E_BADARG;
  }
  return ECNE_OK;
}

int upload_rows(const ecne_problem_t* p, unsigned long long* d_seg, uint32_t* d_col, fr::u256* d_coef, Arena& tmp,
                cudaStream_t s, std::string& err, cudaStream_t s_col, cudaEvent_t ev_col) {
  const uint64_t N = p->n_rows, nnz = problem_nnz(p);
  static const bool prof = getenv("ECNE_HOST_PROF") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  auto done = [&]() {
    const cudaError_t e = put_done();
    if (prof)
      fprintf(stderr, "[ecne dev] rows queued and handed to the stream in %.3f ms\n",
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    return e;
  };
  if (ev_col) CK(cudaEventRecord(ev_col, s));  // (recorded on every path: a problem without terms returns before the wire ids)
  // every array first (queued on the staging pool), then the kernels that expand them
  struct FlushOnExit {  // (an error return must not leave workers reading the caller's arrays)
    ~FlushOnExit() { staged_flush(); }
  } flush_on_exit;
  uint32_t* d_seg32 = nullptr;
  if (p->seg_ptr) {
    CK(put(d_seg, p->seg_ptr, (3 * N + 1) * 8, s));
  } else {
    CK(tmp.alloc(&d_seg32, upload_padded<uint32_t>(3 * N + 1)));
    CK(put(d_seg32, p->seg_ptr32, (3 * N + 1) * 4, s));
  }
  if (!nnz) {
    CK(done());
    if (d_seg32) k_widen_seg<<<(unsigned int)((3 * N + 1 + 255) / 256), 256, 0, s>>>(d_seg32, 3 * N + 1, d_seg);
    return ECNE_OK;
  }
  if (upload_shard().world > 1 || !ev_col) s_col = nullptr;  // (sharded uploads end in a collective: one stream)
  // the wire ids go last, on their own stream when the caller has one: nothing needs them before the classification
  auto put_col = [&]() -> cudaError_t {
    if (!s_col) return put(d_col, p->col, nnz * 4, s);
    cudaError_t e = cudaEventRecord(ev_col, s);  // (behind the arrays queued so far: an order, not a dependency)
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s_col, ev_col, 0);
    if (e == cudaSuccess) e = put(d_col, p->col, nnz * 4, s_col);
    return e;
  };
  if (p->coef) {
    CK(put(d_coef, p->coef, nnz * 32, s));
    CK(put_col());
    CK(done());
    if (s_col) CK(cudaEventRecord(ev_col, s_col));
    if (d_seg32) k_widen_seg<<<(unsigned int)((3 * N + 1 + 255) / 256), 256, 0, s>>>(d_seg32, 3 * N + 1, d_seg);
    return ECNE_OK;
  }
  const uint64_t n_other = p->n_coef_other;
  uint8_t* d_cls;
  fr::u256* d_other;
  uint32_t* d_term;
  unsigned int* d_chk;
  CK(tmp.alloc(&d_cls, upload_padded<uint8_t>(nnz)));
  CK(tmp.alloc(&d_other, upload_padded<fr::u256>(n_other)));
  CK(tmp.alloc(&d_term, upload_padded<uint32_t>(n_other)));
  CK(tmp.alloc(&d_chk, 2));
  CK(cudaMemsetAsync(d_chk, 0, 2 * sizeof(unsigned int), s));
  CK(put(d_cls, p->coef_class, nnz, s));
  if (n_other) {
    CK(put(d_other, p->coef_other, n_other * 32, s));
    CK(put(d_term, p->coef_other_term, n_other * 4, s));
  }
  CK(put_col());
  CK(done());
  if (s_col) CK(cudaEventRecord(ev_col, s_col));
  if (d_seg32) k_widen_seg<<<(unsigned int)((3 * N + 1 + 255) / 256), 256, 0, s>>>(d_seg32, 3 * N + 1, d_seg);
  k_expand_class<<<(unsigned int)((nnz + 255) / 256), 256, 0, s>>>(d_cls, nnz, d_coef, d_chk);
  if (n_other)
    k_expand_other<<<(unsigned int)((n_other + 255) / 256), 256, 0, s>>>(d_other, d_term, n_other, nnz, d_cls, d_coef, d_chk);
  unsigned int chk[2] = {0, 0};
  CK(cudaMemcpyAsync(chk, d_chk, sizeof(chk), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (chk[1] || chk[0] != n_other) {
    err = "compact coefficients are inconsistent: coef_class has " + std::to_string(chk[0]) + " class-3 terms, n_coef_other is " +
          std::to_string(n_other) + (chk[1] ? " (and a class byte > 3, or coef_other_term is not the ascending list of the class-3 terms)" : "");
    return ECNE_E_BADARG;
  }
  return ECNE_OK;
}

int build_resident(const ecne_problem_t* p, Resident* R, std::string& err, const DevSystem* dev) {
  if (!p) {
    err = "null problem";
    return ECNE_E_BADARG;
  }
  if (!dev) {
    const int st = problem_rows_ok(p, err);
    if (st != ECNE_OK) return st;
  }
  const uint64_t N = dev ? dev->N : p->n_rows, V = p->n_vars, nnz = dev ? dev->nnz : problem_nnz(p);
  if (V < 1 || V >= 0x7fffffffULL || N >= 0x3fffffffULL || nnz >= 0x7fffffffULL) {
    err = "problem too large for 32-bit indices (rows < 2^30, wires, terms < 2^31)";
    return ECNE_E_BADARG;
  }
  for (uint64_t i = 0; i < p->n_known; ++i)
    if (p->known[i] < 1 || p->known[i] > V) {
      err = "BoundsError: known wire outside 1..num_variables (:682)";
      return ECNE_E_BOUNDS;
    }
  for (uint64_t i = 0; i < p->n_targets; ++i)
    if (p->targets[i] < 1 || p->targets[i] > V) {
      err = "BoundsError: target wire outside 1..num_variables (:1580)";
      return ECNE_E_BOUNDS;
    }
  const uint64_t n_sp = p->n_specials;
  const uint64_t n_sp_in = n_sp ? p->sp_in_ptr[n_sp] : 0, n_sp_out = n_sp ? p->sp_out_ptr[n_sp] : 0;
  for (uint64_t i = 0; i < n_sp_in; ++i)
    if (p->sp_in[i] < 1 || p->sp_in[i] > V) {
      err = "BoundsError: special-constraint input outside 1..num_variables (:722)";
      return ECNE_E_BOUNDS;
    }
  for (uint64_t i = 0; i < n_sp_out; ++i)
    if (p->sp_out[i] < 1 || p->sp_out[i] > V) {
      err = "BoundsError: special-constraint output outside 1..num_variables (:733)";
      return ECNE_E_BOUNDS;
    }

  cudaStream_t s = R->stream;
  // ECNE_SETUP_PROF=1: wall-clock time of every stage of the set-up, each closed by a device synchronisation (the
  // stages then no longer overlap: this is a profile, not a benchmark)
  static const bool setup_prof = getenv("ECNE_SETUP_PROF") != nullptr;
  auto sp_t0 = std::chrono::steady_clock::now();
  auto sp_lap = [&](const char* what) {
    if (!setup_prof) return;
    cudaDeviceSynchronize();
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[ecne setup] %-28s %8.1f us\n", what, std::chrono::duration<double, std::micro>(t - sp_t0).count());
    sp_t0 = t;
  };
  Arena tmp;  // freed on return
  struct TmpGuard {
    Arena& a;
    ~TmpGuard() { a.release(); }
  } guard{tmp};
  struct Events {  // destroyed on every return path
    cudaEvent_t e[3];
    Events() {
      for (auto& x : e) cudaEventCreate(&x);
    }
    ~Events() {
      for (auto& x : e) cudaEventDestroy(x);
    }
  } evs;
  cudaEvent_t ev0 = evs.e[0], ev1 = evs.e[1], ev2 = evs.e[2];
  cudaEventRecord(ev0, s);

  struct ColEvent {
    cudaEvent_t e;
    ColEvent() { cudaEventCreate(&e); }
    ~ColEvent() { cudaEventDestroy(e); }
  } col_ev;
  bool col_on_side = false;
  // ---- H2D -------------------------------------------------------------------------------
  unsigned long long* d_seg64;
  uint32_t* d_col_raw;
  fr::u256* d_coef_raw;
  if (dev) {  // the rows are on the device already (abstraction.cu): nothing crosses PCIe
    d_seg64 = dev->seg;
    d_col_raw = dev->col;
    d_coef_raw = dev->coef;
  } else {
    CK(tmp.alloc(&d_seg64, upload_padded<unsigned long long>(3 * N + 2)));
    CK(tmp.alloc(&d_col_raw, upload_padded<uint32_t>(nnz)));
    CK(tmp.alloc(&d_coef_raw, upload_padded<fr::u256>(nnz)));
    // (the wire ids cross last, on the second side stream: the coefficients are expanded and the compaction offsets
    // computed while they are in flight; the classification waits for `col_ev.e`)
    col_on_side = R->side2 != nullptr;
    const int st = upload_rows(p, d_seg64, d_col_raw, d_coef_raw, tmp, s, err, R->side2, col_on_side ? col_ev.e : nullptr);
    if (st != ECNE_OK) return st;
    col_on_side = col_on_side && upload_shard().world <= 1;
  }

  sp_lap("upload of the rows");
  Dev& d = R->d;
  memset(&d, 0, sizeof(d));
  d.N = (uint32_t)N;
  d.V = (uint32_t)V;
  d.n_known = (uint32_t)p->n_known;
  d.n_targets = (uint32_t)p->n_targets;
  d.n_specials = (uint32_t)n_sp;
  d.secp_solve = p->secp_solve;
  d.row_lo = 0;
  d.row_hi = (uint32_t)N;
  d.world = 1;  // single GPU until ecne_upload() shards the rows (setup_exchange)
  d.rank = 0;
  Arena& A = R->arena;
  uint32_t *d_known, *d_targets, *d_sp_in_ptr, *d_sp_in, *d_sp_out_ptr, *d_sp_out;
  int32_t* d_sp_kind;
  CK(A.alloc(&d_known, p->n_known));
  CK(A.alloc(&d_targets, p->n_targets));
  CK(h2d(d_known, p->known, p->n_known, s));
  CK(h2d(d_targets, p->targets, p->n_targets, s));
  CK(A.alloc(&d_sp_kind, n_sp));
  CK(A.alloc(&d_sp_in_ptr, n_sp + 1));
  CK(A.alloc(&d_sp_out_ptr, n_sp + 1));
  CK(A.alloc(&d_sp_in, n_sp_in));
  CK(A.alloc(&d_sp_out, n_sp_out));
  std::vector<uint32_t> ip(n_sp + 1, 0), op(n_sp + 1, 0);
  for (uint64_t i = 0; i <= n_sp && n_sp; ++i) {
    ip[i] = (uint32_t)p->sp_in_ptr[i];
    op[i] = (uint32_t)p->sp_out_ptr[i];
  }
  CK(h2d(d_sp_kind, p->sp_kind, n_sp, s));
  CK(h2d(d_sp_in_ptr, ip.data(), n_sp + 1, s));
  CK(h2d(d_sp_out_ptr, op.data(), n_sp + 1, s));
  CK(h2d(d_sp_in, p->sp_in, n_sp_in, s));
  CK(h2d(d_sp_out, p->sp_out, n_sp_out, s));
  // (ip / op live until the function returns; the counters readback below synchronises the stream)
  cudaEventRecord(ev1, s);
  d.known = d_known;
  d.targets = d_targets;
  d.sp_kind = d_sp_kind;
  d.sp_in_ptr = d_sp_in_ptr;
  d.sp_in = d_sp_in;
  d.sp_out_ptr = d_sp_out_ptr;
  d.sp_out = d_sp_out;

  sp_lap("small arrays + sync");
  Raw raw;
  raw.N = (uint32_t)N;
  raw.V = (uint32_t)V;
  raw.nnz = nnz;
  raw.seg = d_seg64;
  raw.col = d_col_raw;
  raw.coef.p = d_coef_raw;

  // ---- non-zero compaction offsets ------------------------------------------------------------
  uint32_t *d_keep, *d_pos, *d_segnz;
  CK(tmp.alloc(&d_keep, nnz + 1));
  CK(tmp.alloc(&d_pos, nnz + 1));
  CK(A.alloc(&d_segnz, 3 * N + 1));
  // Three independent chains from here to the counters' readback: the compaction offsets (keep / scan / seg) on one side
  // stream, the long rows' classification (a warp per row, latency-bound) on the other, the short rows' on the main one.
  cudaStream_t sa = R->side ? R->side : s, sb = R->side2 ? R->side2 : s;
  struct PreEvents {
    cudaEvent_t up, a, b;
    PreEvents() {
      cudaEventCreateWithFlags(&up, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&a, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&b, cudaEventDisableTiming);
    }
    ~PreEvents() {
      cudaEventDestroy(up);
      cudaEventDestroy(a);
      cudaEventDestroy(b);
    }
  } pre_ev;
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, d_keep, d_pos, (int)(nnz + 1), s);
  void* d_cub;
  {
    size_t b2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b2, (unsigned long long*)nullptr,
                                    (unsigned long long*)nullptr, (uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)std::max<uint64_t>(N + N_CONST, 1024), 0, 64, s);
    cub_bytes = std::max(cub_bytes, b2);
    cub_bytes = std::max(cub_bytes, (size_t)1 << 20);
  }
  CK(A.alloc((uint8_t**)&d_cub, cub_bytes));
  R->d_cub = d_cub;
  R->cub_bytes = cub_bytes;
  uint32_t* d_rflags;
  RowAux* d_aux;
  Counters* d_cnt;
  uint32_t* d_c3_long;
  CK(A.alloc(&d_rflags, N));
  CK(A.alloc(&d_aux, N));
  CK(tmp.alloc(&d_cnt, 1));
  CK(tmp.alloc(&d_c3_long, N));
  uint32_t* d_longc;
  CK(tmp.alloc(&d_longc, N));
  CK(cudaMemsetAsync(d_cnt, 0, sizeof(Counters), s));
  if (N) k_longc<<<nb(N, 256), 256, 0, s>>>(raw, d_cnt, d_longc);
  cudaEventRecord(pre_ev.up, s);  // the rows are on the device, the long ones listed
  if (sa != s) cudaStreamWaitEvent(sa, pre_ev.up, 0);
  if (sb != s) cudaStreamWaitEvent(sb, pre_ev.up, 0);
  k_keep<<<nb(nnz + 1, 256), 256, 0, sa>>>(raw, d_keep);
  {
    size_t b = cub_bytes;
    CK(cub::DeviceScan::ExclusiveSum(d_cub, b, d_keep, d_pos, (int)(nnz + 1), sa));
  }
  k_seg<<<nb(3 * N + 1, 256), 256, 0, sa>>>(raw, d_pos, d_segnz);
  cudaEventRecord(pre_ev.a, sa);
  sp_lap("keep / scan / seg");
  // ---- classify -------------------------------------------------------------------------------
  if (N) k_classify_long<<<592, 256, 0, sb>>>(raw, d_rflags, d_aux, d_cnt, d_c3_long, d_longc);
  cudaEventRecord(pre_ev.b, sb);
  if (col_on_side) cudaStreamWaitEvent(s, col_ev.e, 0);  // the wire ids (k_classify_long runs on the stream that carried them)
  if (N) k_classify<<<nb(N, 128), 128, 0, s>>>(raw, d_rflags, d_aux, d_cnt, d_c3_long);
  if (sa != s) cudaStreamWaitEvent(s, pre_ev.a, 0);
  if (sb != s) cudaStreamWaitEvent(s, pre_ev.b, 0);
  Counters cnt;
  CK(cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, s));
  uint32_t nnz_nz = 0;
  CK(cudaMemcpyAsync(&nnz_nz, d_pos + nnz, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  sp_lap("classify + counters readback");
  if (cnt.bad) {
    err = "wire id outside 1..num_variables in a constraint (BoundsError at :681/:829)";
    return ECNE_E_BOUNDS;
  }
  // ---- static values and the bound table ---------------------------------------------------
  d.n2a = cnt.n2a;
  d.n2b = cnt.n2b;
  fr::u256 *d_roots, *d_tvals;
  CK(A.alloc(&d_roots, 2 * (size_t)cnt.n2a));
  CK(A.alloc(&d_tvals, (size_t)N_CONST + cnt.n2b));  // candidates: constants first, then every t
  unsigned int* d_next;
  CK(tmp.alloc(&d_next, 2));
  CK(cudaMemsetAsync(d_next, 0, 2 * sizeof(unsigned int), s));
  // The bound-table chain (values -> sort -> ranks) only depends on the classification, and the sweep
  // layout below does not depend on it: it runs on a side stream, concurrently with the layout kernels.
  cudaStream_t s2 = R->side;  // side streams of the device context (abi.cu)
  cudaStream_t s3 = R->side2 ? R->side2 : s;  // ... the long rows' chain: Case-3 classification -> list -> layout
  cudaEvent_t ev_fork, ev_join, ev_long;
  cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ev_long, cudaEventDisableTiming);
  // (outputs of the layout, allocated and cleared before the fork: the short rows are laid out on the main stream, the
  // long ones on `s3`, both mark the wires they mention)
  uint32_t* d_col;
  fr::u256* d_coef;
  uint8_t* d_nontriv;
  uint32_t* d_long;
  unsigned int* d_nlong;
  CK(A.alloc(&d_col, (size_t)nnz_nz));
  CK(A.alloc(&d_coef, (size_t)nnz_nz));
  CK(A.alloc(&d_nontriv, V + 4));
  CK(A.alloc(&d_long, (size_t)cnt.n_long));
  CK(tmp.alloc(&d_nlong, 1));
  CK(cudaMemsetAsync(d_nontriv, 0, V + 4, s));
  CK(cudaMemsetAsync(d_nlong, 0, sizeof(unsigned int), s));
  cudaEventRecord(ev_fork, s);
  cudaStreamWaitEvent(s2, ev_fork, 0);
  if (s3 != s) cudaStreamWaitEvent(s3, ev_fork, 0);
  // Order of the launches: the GPU work after the fork is ~0.2 ms in all, about what the host needs to enqueue it — the
  // main stream's layout of the short rows goes first, then the long rows' chain, then the bound table's.
  if (N) k_layout<<<nb(N, 128), 128, 0, s>>>(raw, d_keep, d_pos, d_segnz, d_rflags, d_col, d_coef, d_nontriv);
  // Long bit-decomposition candidates (main stream, while the side stream sorts the bound values)
  std::vector<Pow2Entry> tab;
  cudaEvent_t ev_c3;
  cudaEventCreateWithFlags(&ev_c3, cudaEventDisableTiming);
  if (cnt.n_c3_long) {
    // sorted table of 2^i mod p for i < max_c, built on the host (pure constants: kept between calls)
    uint32_t tn = cnt.max_c;
    static std::mutex tab_mu;
    static std::vector<Pow2Entry> tab_cache;
    static bool tab_repeats = false;
    {
      std::lock_guard<std::mutex> lk(tab_mu);
      if (tab_cache.size() != tn) {
        tab_cache.resize(tn);
        fr::u256 x = fr::make_u256(1, 0, 0, 0);
        for (uint32_t i = 0; i < tn; ++i) {
          tab_cache[i].v = x;
          tab_cache[i].e = i;
          x = fr::add(x, x);
        }
        std::sort(tab_cache.begin(), tab_cache.end(),
                  [](const Pow2Entry& a, const Pow2Entry& b) { return fr::cmp(a.v, b.v) < 0; });
        tab_repeats = false;
        for (uint32_t i = 1; i < tn; ++i) tab_repeats |= fr::eq(tab_cache[i].v, tab_cache[i - 1].v);
      }
      tab = tab_cache;
    }
    if (tab_repeats) {
      err = "2^i mod p repeats below the longest row length";
      cudaStreamSynchronize(s2);  // the side chain works in `tmp`, which is released on return
      cudaStreamSynchronize(s);
      cudaStreamSynchronize(s3);
      return ECNE_E_UNSUPPORTED;
    }
    Pow2Entry* d_tab;
    CK(tmp.alloc(&d_tab, tn));
    CK(cudaMemcpyAsync(d_tab, tab.data(), tn * sizeof(Pow2Entry), cudaMemcpyHostToDevice, s3));
    uint32_t mask_words = (tn + 31) / 32;
    size_t smem = (size_t)2 * mask_words * sizeof(unsigned int);
    if (smem > 200 * 1024) {
      err = "row too long for the bit-decomposition classifier";
      cudaStreamSynchronize(s2);
      cudaStreamSynchronize(s);
      cudaStreamSynchronize(s3);
      return ECNE_E_UNSUPPORTED;
    }
    if (smem > 48 * 1024)
      CK(cudaFuncSetAttribute(k_classify_c3_long, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_classify_c3_long<<<cnt.n_c3_long, 256, smem, s3>>>(raw, d_rflags, d_aux, d_c3_long, cnt.n_c3_long, d_tab, tn, mask_words);
    cudaEventRecord(ev_c3, s3);  // (`tab` lives until the function's final synchronisation)
  }

  // ---- sweep layout ---------------------------------------------------------------------------
  // long rows on `s3` (behind their Case-3 classification, whose flip flag orders their C terms), short rows on the
  // main stream: the 208 sorting blocks of ecdsa's long rows (0.14 ms) run beside the 0.16 ms pass over every term
  if (cnt.n_long) {
    k_long_rows<<<nb(N, 256), 256, 0, s3>>>((uint32_t)N, d_rflags, d_long, d_nlong);
    k_layout_long_rows<<<nb((uint64_t)cnt.n_long * 32, 128), 128, 0, s3>>>(raw, d_keep, d_pos, d_segnz, d_rflags, d_long,
                                                                         cnt.n_long, d_col, d_coef, d_nontriv);
    // two length classes (both kernels apply the same length tests): ecdsa's 4128 rows of ~88 terms sort in 20 KB
    // of shared memory with many blocks per SM, its 208 rows of 1025 terms get the large block
    auto smem_for = [](uint32_t len) {  // per term: four limbs, order key, wire, permutation entry, kept flag
      return (size_t)len * (sizeof(fr::u256) + sizeof(unsigned long long) + 2 * sizeof(uint32_t) + 1) + 64;
    };
    if (cnt.max_seg_c > LAYOUT_LONG_MIN)
      k_layout_long<<<cnt.n_long, 256, smem_for(LAYOUT_LONG_MID), s3>>>(raw, d_keep, d_segnz, d_rflags, d_long, cnt.n_long,
                                                                       LAYOUT_LONG_MIN, LAYOUT_LONG_MID, d_col, d_coef, d_nontriv);
    if (cnt.max_seg_c > LAYOUT_LONG_MID) {
      const size_t smem = smem_for(std::min<uint32_t>(cnt.max_seg_c, LAYOUT_LONG_MAX));
      CK(cudaFuncSetAttribute(k_layout_long, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(LAYOUT_LONG_MAX)));
      k_layout_long<<<cnt.n_long, 1024, smem, s3>>>(raw, d_keep, d_segnz, d_rflags, d_long, cnt.n_long, LAYOUT_LONG_MID,
                                                    LAYOUT_LONG_MAX, d_col, d_coef, d_nontriv);
    }
  }
  cudaEventRecord(ev_long, s3);
  sp_lap("layout of the short rows | c3_long + layout of the long rows");
  k_consts<<<1, 256, 0, s2>>>(d_tvals);
  if (N) k_values<<<nb(N, 128), 128, 0, s2>>>(raw, d_rflags, d_aux, d_roots, d_tvals, d_next, d_next + 1);
  const uint32_t nc = N_CONST + cnt.n2b;
  uint32_t *d_idx, *d_flag, *d_incl, *d_rank_of;
  fr::u256* d_table;
  CK(tmp.alloc(&d_idx, nc));
  CK(tmp.alloc(&d_flag, nc));
  CK(tmp.alloc(&d_incl, nc));
  CK(tmp.alloc(&d_rank_of, nc));
  CK(A.alloc(&d_table, nc));
  void* d_ms = nullptr;
  size_t d_ms_bytes = 0;
  // testing knob: 0 = always the library sort, 1 = sample sort from 1024 values on, 2 = ... and buckets of more than 16
  // values take the global-memory path of the bucket sort
  const char* ss_env = getenv("ECNE_SAMPLE_SORT");
  const bool ss_off = ss_env && atoi(ss_env) == 0;
  const uint32_t ss_min = ss_env && atoi(ss_env) >= 1 ? 1024u : SS_MIN;
  const uint32_t ss_cap = ss_env && atoi(ss_env) == 2 ? 16u : SS_BUCKET_CAP;
  if (!ss_off && nc >= ss_min && nc <= SS_MAX) {
    // sample sort (see k_ss_*): 5 launches, ~40 us for ecdsa's 160 k values
    uint32_t n_b = 64;
    while (n_b < 1024 && n_b * 512u < nc) n_b <<= 1;
    while (n_b > 2 && (nc / (n_b * SS_OVERSAMPLE)) < 2) n_b >>= 1;  // (the knob's small inputs: strata of two values at least)
    const uint32_t n_s = n_b * SS_OVERSAMPLE;
    SsKey* d_split;
    uint32_t *d_bkt, *d_slot, *d_starts, *d_gi2;
    unsigned int* d_counts;
    fr::u256* d_val2;
    CK(tmp.alloc(&d_split, n_b));
    CK(tmp.alloc(&d_bkt, nc));
    CK(tmp.alloc(&d_slot, nc));
    CK(tmp.alloc(&d_starts, n_b + 1));
    CK(tmp.alloc(&d_gi2, nc));
    CK(tmp.alloc(&d_counts, n_b));
    CK(tmp.alloc(&d_val2, nc));
    CK(cudaMemsetAsync(d_counts, 0, n_b * sizeof(unsigned int), s2));
    const size_t sm_sample = (size_t)n_s * (sizeof(fr::u256) + 8 + 2 * sizeof(uint32_t));
    const size_t sm_count = (size_t)n_b * (sizeof(fr::u256) + 3 * sizeof(uint32_t));
    const size_t sm_bucket = (size_t)SS_BUCKET_CAP * (sizeof(fr::u256) + 8 + 2 * sizeof(uint32_t));
    CK(cudaFuncSetAttribute(k_ss_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_sample));
    CK(cudaFuncSetAttribute(k_ss_bucket, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_bucket));
    k_ss_sample<<<1, 1024, sm_sample, s2>>>(d_tvals, nc, n_s, d_split);
    k_ss_count<<<nb(nc, 256), 256, sm_count, s2>>>(d_tvals, nc, d_split, n_b, d_bkt, d_slot, d_counts);
    k_ss_starts<<<1, 1024, 0, s2>>>(d_counts, n_b, d_starts);
    k_ss_scatter<<<nb(nc, 256), 256, 0, s2>>>(d_tvals, nc, d_bkt, d_slot, d_starts, d_val2, d_gi2);
    k_ss_bucket<<<n_b, 512, sm_bucket, s2>>>(d_val2, d_gi2, d_starts, d_slot, d_idx, ss_cap);
    size_t need = 0;  // scratch of the scan below
    cub::DeviceScan::InclusiveSum((void*)nullptr, need, d_flag, d_incl, (int)nc, s2);
    need = std::max<size_t>(need, (size_t)1 << 20);
    CK(tmp.alloc((uint8_t**)&d_ms, need));
    d_ms_bytes = need;
  } else {
    // one merge sort of (value, index) pairs under a 256-bit comparator (a handful of launches; an LSD radix sort
    // over four 64-bit limbs costs 40 launch-bound passes).  The VALUES travel with the indices: sorting the index
    // permutation alone made every comparison two random 32-byte gathers.
    k_iota<<<nb(nc, 256), 256, 0, s2>>>(d_idx, nc);
    fr::u256* d_keys;
    CK(tmp.alloc(&d_keys, nc));
    CK(cudaMemcpyAsync(d_keys, d_tvals, (size_t)nc * sizeof(fr::u256), cudaMemcpyDeviceToDevice, s2));
    size_t need = 0;
    cub::DeviceMergeSort::SortPairs((void*)nullptr, need, d_keys, d_idx, (int)nc, U256Less{}, s2);
    need = std::max<size_t>(need, (size_t)1 << 20);  // own scratch: d_cub is in use on the main stream
    CK(tmp.alloc((uint8_t**)&d_ms, need));
    d_ms_bytes = need;
    CK(cub::DeviceMergeSort::SortPairs(d_ms, need, d_keys, d_idx, (int)nc, U256Less{}, s2));
  }
  k_distinct<<<nb(nc, 256), 256, 0, s2>>>(d_tvals, d_idx, nc, d_flag);
  {
    size_t b = d_ms_bytes;
    CK(cub::DeviceScan::InclusiveSum(d_ms, b, d_flag, d_incl, (int)nc, s2));
  }
  k_ranks<<<nb(nc, 256), 256, 0, s2>>>(d_tvals, d_idx, d_incl, nc, d_rank_of, d_table);
  if (cnt.n_c3_long) cudaStreamWaitEvent(s2, ev_c3, 0);  // k_fill_ranks reads the C3 flags of those rows
  if (N) k_fill_ranks<<<nb(N, 256), 256, 0, s2>>>((uint32_t)N, d_rflags, d_aux, d_rank_of, d_segnz);
  uint32_t h_rank[3], h_tn;
  cudaEventRecord(ev_join, s2);

  sp_lap("values + sort + ranks (side)");
  if (s3 != s) cudaStreamWaitEvent(s, ev_long, 0);  // from here on: final row flags, every row laid out
  if (n_sp_in) k_mark<<<nb(n_sp_in, 256), 256, 0, s>>>(d_sp_in, (uint32_t)n_sp_in, d_nontriv);
  if (n_sp_out) k_mark<<<nb(n_sp_out, 256), 256, 0, s>>>(d_sp_out, (uint32_t)n_sp_out, d_nontriv);
  if (p->n_targets) k_mark<<<nb(p->n_targets, 256), 256, 0, s>>>(d_targets, (uint32_t)p->n_targets, d_nontriv);
  RowRec* d_rec;
  CK(A.alloc(&d_rec, N));
  if (N) k_rowrec<<<nb(N, 256), 256, 0, s>>>((uint32_t)N, d_segnz, d_col, d_rflags, d_rec);
  if (cnt.n_long) k_long_index<<<nb(cnt.n_long, 256), 256, 0, s>>>(cnt.n_long, d_long, d_rec);
  CK(A.alloc(&d.long_done, (size_t)cnt.n_long));
  CK(A.alloc(&d.long_stamp, (size_t)cnt.n_long));
  CK(A.alloc(&d.long_p2, (size_t)cnt.n_long));
  // static P3 / P4 row lists
  uint32_t *d_p3, *d_p4;
  unsigned int* d_nphase;
  CK(A.alloc(&d_p3, N));
  CK(A.alloc(&d_p4, N));
  CK(tmp.alloc(&d_nphase, 2));
  CK(cudaMemsetAsync(d_nphase, 0, 2 * sizeof(unsigned int), s));
  if (N) k_phase_rows<<<nb(N, 256), 256, 0, s>>>((uint32_t)N, d_rflags, d_p3, d_p4, d_nphase);
  unsigned int h_nphase[2] = {0, 0};
  CK(cudaMemcpyAsync(h_nphase, d_nphase, sizeof(h_nphase), cudaMemcpyDeviceToHost, s));
  sp_lap("marks, rowrec, phase rows");
  // wire -> rows index (count, scan, fill)
  uint32_t *d_inv_ptr, *d_inv_row, *d_inv_cur;
  CK(A.alloc(&d_inv_ptr, V + 3));
  CK(A.alloc(&d_inv_row, (size_t)nnz_nz + 1));
  CK(tmp.alloc(&d_inv_cur, V + 3));
  CK(cudaMemsetAsync(d_inv_cur, 0, (V + 3) * sizeof(uint32_t), s));
  if (nnz_nz) k_inv_count<<<nb(nnz_nz, 256), 256, 0, s>>>(nnz_nz, d_col, d_inv_cur);
  {
    size_t b = cub_bytes, need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, d_inv_cur, d_inv_ptr, (int)(V + 2), s);
    if (need > b) {
      err = "internal: scan scratch too small";
      return ECNE_E_INTERNAL;
    }
    CK(cub::DeviceScan::ExclusiveSum(d_cub, b, d_inv_cur, d_inv_ptr, (int)(V + 2), s));
  }
  CK(cudaMemsetAsync(d_inv_cur, 0, (V + 3) * sizeof(uint32_t), s));
  CK(cudaMemsetAsync(d_inv_row, 0xff, ((size_t)nnz_nz + 1) * sizeof(uint32_t), s));  // 0xffffffff = unused slot
  if (N) k_inv_fill<<<nb(N, 256), 256, 0, s>>>((uint32_t)N, d_segnz, d_col, d_inv_ptr, d_inv_cur, d_inv_row);
  if (cnt.n_long)
    k_inv_fill_long<<<cnt.n_long, 256, 0, s>>>(d_long, cnt.n_long, d_segnz, d_col, d_inv_ptr, d_inv_cur, d_inv_row);
  uint4* d_inv_head;
  CK(A.alloc(&d_inv_head, V + 2));
  k_inv_head<<<nb(V + 2, 256), 256, 0, s>>>((uint32_t)(V + 2), d_inv_ptr, d_inv_cur, d_inv_row, d_inv_head);
  d.inv_ptr = d_inv_ptr;
  d.inv_row = d_inv_row;
  d.inv_head = d_inv_head;
  d.p3_rows = d_p3;
  d.p4_rows = d_p4;
  sp_lap("wire -> rows index");
  // P2 grouping table: power of two >= 2 N slots
  {
    uint32_t cap = 1024;
    while (cap < 2 * (uint64_t)N && cap < (1u << 30)) cap <<= 1;
    d.h_mask = cap - 1;
    CK(A.alloc(&d.h_key, (size_t)cap));
    CK(A.alloc(&d.h_cnt, (size_t)cap));
    CK(A.alloc(&d.h_head, (size_t)cap));
    // (cleared on the second side stream, behind the long rows' chain: nothing on the main stream waits for 32 MB of
    // memsets; the final synchronisation of the function covers them)
    CK(cudaMemsetAsync(d.h_key, 0, (size_t)cap * sizeof(unsigned long long), s3));
    CK(cudaMemsetAsync(d.h_cnt, 0, (size_t)cap * sizeof(uint32_t), s3));
    CK(cudaMemsetAsync(d.h_head, 0, (size_t)cap * sizeof(uint32_t), s3));
    CK(A.alloc(&d.p2_next, N));
    CK(A.alloc(&d.p2_slot, N));
    CK(A.alloc(&d.p2_k, N));
    CK(A.alloc(&d.p2_open, (N + 31) / 32 + 1));
    CK(A.alloc(&d.p2_list, (size_t)CH_LIST_CAP + CH_LONG_MAX));
    CK(A.alloc(&d.p2_bigq, (size_t)P2_BIGQ_CAP));
  }

  // ---- wire state, records, scratch -----------------------------------------------------------
  for (int b = 0; b < 2; ++b) {
    CK(A.alloc(&d.F[b], V + 8));
    CK(A.alloc(&d.LBR[b], V + 1));
    CK(A.alloc(&d.UBR[b], V + 1));
  }
  CK(A.alloc(&d.abz, V + 1));
  CK(A.alloc(&d.valsrc, V + 1));
  CK(A.alloc(&d.abz_claim, V + 1));
  CK(A.alloc(&d.solved, N + 2));
  CK(A.alloc(&d.sp_solved, n_sp));
  CK(A.alloc(&d.sp_tag, V + 2));
  CK(cudaMemsetAsync(d.sp_tag, 0xff, (V + 2) * sizeof(uint32_t), s));
  d.rec_cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(2 * V, nnz) + 4096, 0x7ffffff0ULL);
  for (int l = 0; l < 5; ++l) CK(A.alloc(&d.recs[l], d.rec_cap));
  CK(A.alloc(&d.rec_count, 8));
  CK(A.alloc(&d.dcnt, 8));
  for (int l = 0; l < 5; ++l) CK(A.alloc(&d.wflag[l], V + 8));
  CK(A.alloc(&d.bnd_flag, 8));
  CK(A.alloc(&d.c5sig, N));
  CK(A.alloc(&d.barrier, 128));
  CK(A.alloc(&d.prof, (size_t)28000 + 40 * 148 * 4 + 160));
  CK(cudaMemsetAsync(d.prof, 0, ((size_t)28000 + 40 * 148 * 4 + 160) * sizeof(unsigned long long), s));
  CK(A.alloc(&d.st, 1));
  CK(A.alloc(&d.p2_row, N));
  CK(A.alloc(&R->d_ubits, (V + 63) / 64));
  CK(A.alloc(&R->d_kbits, (V + 63) / 64));
  CK(A.alloc(&R->d_counts, 4));

  sp_lap("table memsets + allocations");
  cudaStreamWaitEvent(s, ev_join, 0);  // the timing event below covers the side chain too
  // the disjoint sets of equal wires and their one observable use (:634-678, :760-768)
  unsigned int* d_dsu_flags = nullptr;
  unsigned int h_dsu_flags[2] = {0, 0};
  if (p->secp_solve && N) {
    uint32_t* d_parent;
    CK(tmp.alloc(&d_parent, V + 1 + (size_t)nc));
    CK(tmp.alloc(&d_dsu_flags, 2));
    CK(cudaMemsetAsync(d_dsu_flags, 0, 2 * sizeof(unsigned int), s));
    k_iota<<<nb(V + 1 + (uint64_t)nc, 256), 256, 0, s>>>(d_parent, (uint32_t)(V + 1 + nc));
    k_dsu_union<<<nb(N, 128), 128, 0, s>>>(raw, d_rflags, d_aux, d_parent, d_dsu_flags);
    bool has_m = false, has_l = false;
    for (uint64_t i = 0; i < n_sp; ++i) {
      has_m |= p->sp_kind[i] == ECNE_SPECIAL_BIGMULTMODP;
      has_l |= p->sp_kind[i] == ECNE_SPECIAL_BIGLESSTHAN;
    }
    if (has_m && has_l)
      k_dsu_pairs<<<nb(n_sp * n_sp, 128), 128, 0, s>>>((uint32_t)n_sp, d_sp_kind, d_sp_in_ptr, d_sp_in, d_sp_out_ptr, d_parent,
                                                       d_dsu_flags);
    CK(cudaMemcpyAsync(h_dsu_flags, d_dsu_flags, sizeof(h_dsu_flags), cudaMemcpyDeviceToHost, s));
  }
  CK(cudaMemcpyAsync(&h_rank[0], d_rank_of + 0, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&h_rank[1], d_rank_of + 1, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&h_rank[2], d_rank_of + 254, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&h_tn, d_incl + (nc - 1), 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaStreamSynchronize(s2));
  if (s3 != s) CK(cudaStreamSynchronize(s3));
  cudaEventDestroy(ev_fork);
  cudaEventDestroy(ev_join);
  cudaEventDestroy(ev_c3);
  cudaEventDestroy(ev_long);
  if (h_dsu_flags[0]) {
    err = "BoundsError: a row with two stored C keys and fewer than two usable ones in the disjoint-set construction (:660-662)";
    return ECNE_E_BOUNDS;
  }
  d.p0p_bounds = h_dsu_flags[1] ? 1 : 0;
  d.r0 = h_rank[0];
  d.r1 = h_rank[1];
  d.rpm1 = h_rank[2];
  d.table_n = h_tn;
  d.nnz = nnz_nz;
  d.n_long = cnt.n_long;
  d.n_p3 = h_nphase[0];
  d.n_p4 = h_nphase[1];
  d.seg = d_segnz;
  d.col = d_col;
  d.coef = d_coef;
  d.rflags = d_rflags;
  d.aux = d_aux;
  d.long_rows = d_long;
  d.rec = d_rec;
  d.roots = d_roots;
  d.tvals = d_tvals + N_CONST;
  d.table = d_table;
  d.nontriv = d_nontriv;
  R->n_rows = N;
  R->n_vars = V;
  R->n_targets = p->n_targets;
  sp_lap("dsu + final readbacks");
  cudaEventRecord(ev2, s);
  CK(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, ev0, ev1);
  R->ms_h2d = ms;
  if (col_on_side) {  // (the last array of the upload finished on the side stream; kernels were running by then)
    float ms_col = 0;
    if (cudaEventElapsedTime(&ms_col, ev0, col_ev.e) != cudaSuccess)
      cudaGetLastError();
    else if (ms_col > R->ms_h2d)
      R->ms_h2d = ms_col;
  }
  cudaEventElapsedTime(&ms, ev0, ev2);
  R->ms_classify = ms - R->ms_h2d;
  CK(cudaGetLastError());
  if (d.r0 != 0) {
    err = "internal: rank of 0 is not 0";
    return ECNE_E_INTERNAL;
  }
  return ECNE_OK;
}

}  // namespace ecne
