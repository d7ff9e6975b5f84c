// oracle/u256.h — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Deliberately slow, obviously-correct 256-bit integer and BN254-scalar-field arithmetic for the
// CPU oracle.  It restates what AbstractAlgebra 0.23.0 `GF(p)` over BigInt (third-party,
// un-vendored; /root/reference/Manifest.toml:3-7) provides at the reference's call sites
// (R1CSConstraintSolver.jl:668, :919-920, :961-964, :999-1000, :1006, :1033-1037, :1395-1397,
// :1467): + - * ^ divexact == on canonical residues, plus plain BigInt compare / % / div.
// No Montgomery form here on purpose — the device code (csrc/fr_bn254.cuh) uses Montgomery, so the
// two implementations share no algorithm.  Pinned against Python ints in tests/test_field_kat.py.
#pragma once
#include <cstdint>
#include <cstring>

namespace orc {

struct U256 {
  uint64_t l[4];
};
struct U512 {
  uint64_t l[8];
};

inline U256 u256(uint64_t x) { return U256{{x, 0, 0, 0}}; }
inline bool is_zero(const U256& a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
inline bool eq(const U256& a, const U256& b) { return memcmp(a.l, b.l, 32) == 0; }
inline int cmp(const U256& a, const U256& b) {
  for (int i = 3; i >= 0; --i)
    if (a.l[i] != b.l[i]) return a.l[i] < b.l[i] ? -1 : 1;
  return 0;
}
// a + b, returns carry
inline uint64_t add(U256& r, const U256& a, const U256& b) {
  unsigned __int128 c = 0;
  for (int i = 0; i < 4; ++i) {
    c += (unsigned __int128)a.l[i] + b.l[i];
    r.l[i] = (uint64_t)c;
    c >>= 64;
  }
  return (uint64_t)c;
}
// a - b, returns borrow
inline uint64_t sub(U256& r, const U256& a, const U256& b) {
  uint64_t borrow = 0;
  for (int i = 0; i < 4; ++i) {
    unsigned __int128 d = (unsigned __int128)a.l[i] - b.l[i] - borrow;
    r.l[i] = (uint64_t)d;
    borrow = (uint64_t)(d >> 64) & 1;
  }
  return borrow;
}
inline U256 shl1(const U256& a) {
  U256 r;
  r.l[3] = (a.l[3] << 1) | (a.l[2] >> 63);
  r.l[2] = (a.l[2] << 1) | (a.l[1] >> 63);
  r.l[1] = (a.l[1] << 1) | (a.l[0] >> 63);
  r.l[0] = a.l[0] << 1;
  return r;
}
inline U256 shr1(const U256& a) {
  U256 r;
  r.l[0] = (a.l[0] >> 1) | (a.l[1] << 63);
  r.l[1] = (a.l[1] >> 1) | (a.l[2] << 63);
  r.l[2] = (a.l[2] >> 1) | (a.l[3] << 63);
  r.l[3] = a.l[3] >> 1;
  return r;
}
inline int bit(const U256& a, int i) { return (a.l[i >> 6] >> (i & 63)) & 1; }
inline int bitlen(const U256& a) {
  for (int i = 3; i >= 0; --i)
    if (a.l[i]) return 64 * i + 64 - __builtin_clzll(a.l[i]);
  return 0;
}
inline U512 mul_wide(const U256& a, const U256& b) {
  U512 r;
  memset(r.l, 0, sizeof(r.l));
  for (int i = 0; i < 4; ++i) {
    unsigned __int128 carry = 0;
    for (int j = 0; j < 4; ++j) {
      unsigned __int128 t = (unsigned __int128)a.l[i] * b.l[j] + r.l[i + j] + carry;
      r.l[i + j] = (uint64_t)t;
      carry = t >> 64;
    }
    r.l[i + 4] = (uint64_t)carry;
  }
  return r;
}
// compare a 512-bit value with a 256-bit one
inline int cmp512_256(const U512& a, const U256& b) {
  for (int i = 7; i >= 4; --i)
    if (a.l[i]) return 1;
  for (int i = 3; i >= 0; --i)
    if (a.l[i] != b.l[i]) return a.l[i] < b.l[i] ? -1 : 1;
  return 0;
}
// plain integer division by shift-subtract: a = q*b + r (b != 0)
inline void divrem(const U256& a, const U256& b, U256& q, U256& r) {
  q = u256(0);
  r = u256(0);
  for (int i = bitlen(a) - 1; i >= 0; --i) {
    uint64_t top = r.l[3] >> 63;
    r = shl1(r);
    r.l[0] |= (uint64_t)bit(a, i);
    if (top || cmp(r, b) >= 0) {
      sub(r, r, b);
      q.l[i >> 6] |= 1ULL << (i & 63);
    }
  }
}

// ---- the field --------------------------------------------------------------------------
// bjj_p (R1CSConstraintSolver.jl:21-22)
static const U256 P = {{0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL,
                        0x30644e72e131a029ULL}};

inline U256 fadd(const U256& a, const U256& b) {
  U256 r;
  uint64_t c = add(r, a, b);
  if (c || cmp(r, P) >= 0) sub(r, r, P);
  return r;
}
inline U256 fsub(const U256& a, const U256& b) {
  U256 r;
  if (sub(r, a, b)) add(r, r, P);
  return r;
}
inline U256 fneg(const U256& a) { return is_zero(a) ? a : fsub(u256(0), a); }
// 512-bit value mod p by shift-subtract
inline U256 mod512(const U512& x) {
  U256 r = u256(0);
  int top = 511;
  while (top >= 0 && !((x.l[top >> 6] >> (top & 63)) & 1)) --top;
  for (int i = top; i >= 0; --i) {
    uint64_t hi = r.l[3] >> 63;
    r = shl1(r);
    r.l[0] |= (x.l[i >> 6] >> (i & 63)) & 1;
    if (hi || cmp(r, P) >= 0) sub(r, r, P);
  }
  return r;
}
inline U256 fmul(const U256& a, const U256& b) { return mod512(mul_wide(a, b)); }
inline U256 fpow(U256 b, uint64_t e) {
  U256 r = u256(1);
  while (e) {
    if (e & 1) r = fmul(r, b);
    b = fmul(b, b);
    e >>= 1;
  }
  return r;
}
// modular inverse by the binary extended Euclid on (a, p); a in [1, p)
inline U256 finv(const U256& a) {
  U256 u = a, v = P, x1 = u256(1), x2 = u256(0);
  auto half = [](U256& x) {
    if (x.l[0] & 1) {
      uint64_t c = add(x, x, P);
      x = shr1(x);
      if (c) x.l[3] |= 1ULL << 63;
    } else {
      x = shr1(x);
    }
  };
  const U256 one = u256(1);
  while (!eq(u, one) && !eq(v, one)) {
    while (!(u.l[0] & 1)) {
      u = shr1(u);
      half(x1);
    }
    while (!(v.l[0] & 1)) {
      v = shr1(v);
      half(x2);
    }
    if (cmp(u, v) >= 0) {
      sub(u, u, v);
      x1 = fsub(x1, x2);
    } else {
      sub(v, v, u);
      x2 = fsub(x2, x1);
    }
  }
  return eq(u, one) ? x1 : x2;
}

}  // namespace orc
