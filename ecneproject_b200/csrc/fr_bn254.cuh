// fr_bn254.cuh — BN254 scalar field on the device: 4 x 64-bit limbs, Montgomery form (R = 2^256),
// everything in registers; plus the plain 256-bit integer helpers the bound rules need.
//
// Replaces AbstractAlgebra 0.23.0 GF(p) over BigInt (un-vendored; call sites
// /root/reference/src/R1CSConstraintSolver.jl:668, :919-920, :961-964, :999-1000, :1006,
// :1033-1037, :1395-1397, :1467) and the BigInt compare / % / div of :1035, :1113-1114, :1267-1274.
// Carry chains are written as PTX add.cc / addc / mad.lo.cc / madc.hi so that one limb product
// plus accumulation is two instructions (IMAD.WIDE pairs in SASS) instead of a compare-and-add.
#pragma once
#include <stdint.h>

namespace fr {

struct u256 {
  uint64_t v[4];
};

// p, little-endian limbs
#define FR_P0 0x43e1f593f0000001ULL
#define FR_P1 0x2833e84879b97091ULL
#define FR_P2 0xb85045b68181585dULL
#define FR_P3 0x30644e72e131a029ULL
// -p^{-1} mod 2^64
#define FR_NINV 0xc2e1f593efffffffULL

__host__ __device__ __forceinline__ u256 make_u256(uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
  u256 r;
  r.v[0] = a;
  r.v[1] = b;
  r.v[2] = c;
  r.v[3] = d;
  return r;
}
__host__ __device__ __forceinline__ u256 modulus() { return make_u256(FR_P0, FR_P1, FR_P2, FR_P3); }
// R mod p  (Montgomery one)
__host__ __device__ __forceinline__ u256 mont_one() {
  return make_u256(0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL,
                   0x0e0a77c19a07df2fULL);
}
// R^2 mod p
__host__ __device__ __forceinline__ u256 mont_r2() {
  return make_u256(0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL,
                   0x0216d0b17f4e44a5ULL);
}
// the mistyped sign-fold threshold of flip_coeffs (R1CSConstraintSolver.jl:1246-1247)
__host__ __device__ __forceinline__ u256 fold_threshold() {
  return make_u256(0x43e1f593f0000000ULL, 0x9c41be16bb2a8891ULL, 0x045fcd3eea44076aULL,
                   0x2e2e53955f6f1dfeULL);
}

#ifdef __CUDACC__
// One 256-bit load / store per value (LDG.E.256 / STG.E.256 on sm_100): the arrays of u256 in global memory are
// 32-byte aligned (arena allocations of 256 bytes, 32-byte elements); the struct itself only promises 8, so plain
// accesses compile to four 64-bit ones, each a strided warp access.
__device__ __forceinline__ u256 ldg256(const u256* p) {
  u256 r;
  asm("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(r.v[0]), "=l"(r.v[1]), "=l"(r.v[2]), "=l"(r.v[3]) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg256(u256* p, const u256& x) {
  asm volatile("st.global.v4.u64 [%4], {%0,%1,%2,%3};" ::"l"(x.v[0]), "l"(x.v[1]), "l"(x.v[2]), "l"(x.v[3]), "l"(p) : "memory");
}
#endif

__host__ __device__ __forceinline__ bool is_zero(const u256& a) {
  return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0;
}
__host__ __device__ __forceinline__ bool eq(const u256& a, const u256& b) {
  return ((a.v[0] ^ b.v[0]) | (a.v[1] ^ b.v[1]) | (a.v[2] ^ b.v[2]) | (a.v[3] ^ b.v[3])) == 0;
}
__host__ __device__ __forceinline__ bool is_one(const u256& a) {
  return a.v[0] == 1 && (a.v[1] | a.v[2] | a.v[3]) == 0;
}
// -1, i.e. p-1 in canonical form
__host__ __device__ __forceinline__ bool is_minus_one(const u256& a) {
  return a.v[0] == FR_P0 - 1 && a.v[1] == FR_P1 && a.v[2] == FR_P2 && a.v[3] == FR_P3;
}
__host__ __device__ __forceinline__ int cmp(const u256& a, const u256& b) {
#pragma unroll
  for (int i = 3; i >= 0; --i) {
    if (a.v[i] != b.v[i]) return a.v[i] < b.v[i] ? -1 : 1;
  }
  return 0;
}

#ifdef __CUDA_ARCH__
// r = a + b, returns carry-out
__device__ __forceinline__ uint64_t add_cc(u256& r, const u256& a, const u256& b) {
  uint64_t c;
  asm("add.cc.u64 %0, %5, %9;\n\t"
      "addc.cc.u64 %1, %6, %10;\n\t"
      "addc.cc.u64 %2, %7, %11;\n\t"
      "addc.cc.u64 %3, %8, %12;\n\t"
      "addc.u64 %4, 0, 0;"
      : "=l"(r.v[0]), "=l"(r.v[1]), "=l"(r.v[2]), "=l"(r.v[3]), "=l"(c)
      : "l"(a.v[0]), "l"(a.v[1]), "l"(a.v[2]), "l"(a.v[3]), "l"(b.v[0]), "l"(b.v[1]), "l"(b.v[2]),
        "l"(b.v[3]));
  return c;
}
// r = a - b, returns borrow-out (1 if a < b)
__device__ __forceinline__ uint64_t sub_cc(u256& r, const u256& a, const u256& b) {
  uint64_t c;
  asm("sub.cc.u64 %0, %5, %9;\n\t"
      "subc.cc.u64 %1, %6, %10;\n\t"
      "subc.cc.u64 %2, %7, %11;\n\t"
      "subc.cc.u64 %3, %8, %12;\n\t"
      "subc.u64 %4, 0, 0;"
      : "=l"(r.v[0]), "=l"(r.v[1]), "=l"(r.v[2]), "=l"(r.v[3]), "=l"(c)
      : "l"(a.v[0]), "l"(a.v[1]), "l"(a.v[2]), "l"(a.v[3]), "l"(b.v[0]), "l"(b.v[1]), "l"(b.v[2]),
        "l"(b.v[3]));
  return c & 1;
}
#else
inline uint64_t add_cc(u256& r, const u256& a, const u256& b) {
  unsigned __int128 c = 0;
  for (int i = 0; i < 4; ++i) {
    c += (unsigned __int128)a.v[i] + b.v[i];
    r.v[i] = (uint64_t)c;
    c >>= 64;
  }
  return (uint64_t)c;
}
inline uint64_t sub_cc(u256& r, const u256& a, const u256& b) {
  uint64_t borrow = 0;
  for (int i = 0; i < 4; ++i) {
    unsigned __int128 d = (unsigned __int128)a.v[i] - b.v[i] - borrow;
    r.v[i] = (uint64_t)d;
    borrow = (uint64_t)(d >> 64) & 1;
  }
  return borrow;
}
#endif

// ---- modular add / sub / neg on values in [0, p) (canonical or Montgomery alike) ------------
__host__ __device__ __forceinline__ u256 add(const u256& a, const u256& b) {
  u256 s, t;
  add_cc(s, a, b);  // a + b < 2p < 2^255: no carry out
  uint64_t borrow = sub_cc(t, s, modulus());
  return borrow ? s : t;
}
__host__ __device__ __forceinline__ u256 sub(const u256& a, const u256& b) {
  u256 d, t;
  uint64_t borrow = sub_cc(d, a, b);
  add_cc(t, d, modulus());
  return borrow ? t : d;
}
__host__ __device__ __forceinline__ u256 neg(const u256& a) {
  if (is_zero(a)) return a;
  u256 r;
  sub_cc(r, modulus(), a);
  return r;
}

// ---- Montgomery multiplication: returns a*b*R^-1 mod p, inputs/outputs in [0, p) ------------
#ifdef __CUDA_ARCH__
// (hi, lo) = a*b + c + d   — never overflows 128 bits
__device__ __forceinline__ void mac(uint64_t& hi, uint64_t& lo, uint64_t a, uint64_t b, uint64_t c,
                                    uint64_t d) {
  asm("{\n\t"
      ".reg .u64 t;\n\t"
      "mad.lo.cc.u64 %1, %2, %3, %4;\n\t"
      "madc.hi.u64 %0, %2, %3, 0;\n\t"
      "add.cc.u64 %1, %1, %5;\n\t"
      "addc.u64 %0, %0, 0;\n\t"
      "}"
      : "=&l"(hi), "=&l"(lo)
      : "l"(a), "l"(b), "l"(c), "l"(d));
}
#else
inline void mac(uint64_t& hi, uint64_t& lo, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
  unsigned __int128 t = (unsigned __int128)a * b + c + d;
  lo = (uint64_t)t;
  hi = (uint64_t)(t >> 64);
}
#endif

// CIOS, 4 limbs.  p < 2^254 so the running value stays below 2p and one conditional subtract
// suffices.
__host__ __device__ __forceinline__ u256 mul(const u256& a, const u256& b) {
  uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint64_t c, lo;
    mac(c, lo, a.v[0], b.v[i], t0, 0);
    t0 = lo;
    mac(c, lo, a.v[1], b.v[i], t1, c);
    t1 = lo;
    mac(c, lo, a.v[2], b.v[i], t2, c);
    t2 = lo;
    mac(c, lo, a.v[3], b.v[i], t3, c);
    t3 = lo;
    t4 += c;  // t4 <= 1 before, c < 2^64 - 1: no overflow given a, b < p < 2^254
    uint64_t m = t0 * FR_NINV;
    mac(c, lo, m, FR_P0, t0, 0);  // lo == 0
    mac(c, lo, m, FR_P1, t1, c);
    t0 = lo;
    mac(c, lo, m, FR_P2, t2, c);
    t1 = lo;
    mac(c, lo, m, FR_P3, t3, c);
    t2 = lo;
    t3 = t4 + c;
    t4 = (t3 < c) ? 1 : 0;
  }
  u256 r = make_u256(t0, t1, t2, t3), s;
  uint64_t borrow = sub_cc(s, r, modulus());
  return (t4 || !borrow) ? s : r;
}
__host__ __device__ __forceinline__ u256 sqr(const u256& a) { return mul(a, a); }
__host__ __device__ __forceinline__ u256 to_mont(const u256& a) { return mul(a, mont_r2()); }
__host__ __device__ __forceinline__ u256 from_mont(const u256& a) {
  return mul(a, make_u256(1, 0, 0, 0));
}

// a^(p-2) in Montgomery form (Fermat); a != 0.  p-2 has limbs {P0-2, P1, P2, P3}.
__host__ __device__ inline u256 inv_mont(const u256& a) {
  const uint64_t e[4] = {FR_P0 - 2, FR_P1, FR_P2, FR_P3};
  u256 r = mont_one();
  bool started = false;
  for (int i = 3; i >= 0; --i) {
    for (int b = 63; b >= 0; --b) {
      if (started) r = sqr(r);
      if ((e[i] >> b) & 1) {
        r = started ? mul(r, a) : a;
        started = true;
      }
    }
  }
  return r;
}

// divexact(-num, den) on canonical inputs, canonical result; den != 0 (the caller raises the
// DivideError).  Shortcuts for den = +-1 cover almost every circom row.
__host__ __device__ inline u256 neg_div(const u256& num, const u256& den) {
  u256 n = neg(num);
  if (is_one(den)) return n;
  if (is_minus_one(den)) return num;
  if (is_zero(n)) return n;
  u256 dm = to_mont(den);
  u256 di = inv_mont(dm);            // den^-1 * R
  return mul(n, di);                 // n * den^-1 * R * R^-1 = canonical quotient
}

// ---- plain 256-bit integers --------------------------------------------------------------
__host__ __device__ __forceinline__ int bitlen(const u256& a) {
#pragma unroll
  for (int i = 3; i >= 0; --i) {
    if (a.v[i]) {
#ifdef __CUDA_ARCH__
      return 64 * i + 64 - __clzll((long long)a.v[i]);
#else
      return 64 * i + 64 - __builtin_clzll(a.v[i]);
#endif
    }
  }
  return 0;
}
__host__ __device__ __forceinline__ bool is_pow2(const u256& a) {
  int nz = 0;
  uint64_t w = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (a.v[i]) {
      ++nz;
      w = a.v[i];
    }
  return nz == 1 && (w & (w - 1)) == 0;
}
__host__ __device__ __forceinline__ u256 shl1(const u256& a) {
  return make_u256(a.v[0] << 1, (a.v[1] << 1) | (a.v[0] >> 63), (a.v[2] << 1) | (a.v[1] >> 63),
                   (a.v[3] << 1) | (a.v[2] >> 63));
}
// a >> k, k in [0, 256)
__host__ __device__ __forceinline__ u256 shr(const u256& a, int k) {
  const int w = k >> 6, b = k & 63;
  u256 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint64_t lo = (i + w < 4) ? a.v[(i + w) & 3] : 0ULL;
    const uint64_t hi = (i + w + 1 < 4) ? a.v[(i + w + 1) & 3] : 0ULL;
    r.v[i] = b ? ((lo >> b) | (hi << (64 - b))) : lo;
  }
  return r;
}
// 2^k - 1 as an integer, k in [0, 256)
__host__ __device__ __forceinline__ u256 pow2m1(int k) {
  u256 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int lo = 64 * i;
    if (k >= lo + 64)
      r.v[i] = ~0ULL;
    else if (k <= lo)
      r.v[i] = 0;
    else
      r.v[i] = (1ULL << (k - lo)) - 1;
  }
  return r;
}
// a mod b == 0 ?   (b != 0) — shift/subtract, only on the rare generic branch of Case 5
__host__ __device__ inline bool divides(const u256& b, const u256& a) {
  if (is_one(b)) return true;
  int la = bitlen(a), lb = bitlen(b);
  if (la < lb) return is_zero(a);
  if (is_pow2(b)) {  // low (lb-1) bits of a must be zero
    u256 m = pow2m1(lb - 1);
    return ((a.v[0] & m.v[0]) | (a.v[1] & m.v[1]) | (a.v[2] & m.v[2]) | (a.v[3] & m.v[3])) == 0;
  }
  // restoring division, starting from the top lb bits of a: la - lb + 1 compare-subtract steps instead of
  // la (Case 5 divides neighbouring magnitudes of a sorted chain: their lengths are close)
  const int k = la - lb;
  u256 r = shr(a, k);
  {
    u256 t;
    if (!sub_cc(t, r, b)) r = t;
  }
  for (int i = k - 1; i >= 0; --i) {
    uint64_t top = r.v[3] >> 63;
    r = shl1(r);
    r.v[0] |= (a.v[i >> 6] >> (i & 63)) & 1;
    u256 t;
    uint64_t borrow = sub_cc(t, r, b);
    if (top || !borrow) r = t;
  }
  return is_zero(r);
}
// 512-bit product a*b compared with (hi:0, lo:c): returns sign of a*b - c
__host__ __device__ inline int cmp_mul(const u256& a, const u256& b, const u256& c) {
  uint64_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint64_t carry = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint64_t hi, lo;
      mac(hi, lo, a.v[i], b.v[j], t[i + j], carry);
      t[i + j] = lo;
      carry = hi;
    }
    t[i + 4] = carry;
  }
  if (t[4] | t[5] | t[6] | t[7]) return 1;
#pragma unroll
  for (int i = 3; i >= 0; --i) {
    if (t[i] != c.v[i]) return t[i] < c.v[i] ? -1 : 1;
  }
  return 0;
}

}  // namespace fr
