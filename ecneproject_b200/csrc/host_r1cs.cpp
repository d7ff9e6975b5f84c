// host_r1cs.cpp — .r1cs -> flattened CSR loader and the trusted-function abstraction pass.
//
// Host-side callers of the hot path; behaviour follows
//   /root/reference/src/ParseR1CS.jl:50-124          (readR1CS)
//   /root/reference/src/R1CSConstraintSolver.jl:205-395  (checkNonZeroValues, hash_r1cs_equation,
//                                                        abstraction)
// but the data structure is the engine's CSR (include/ecne_abi.h), not per-row hash maps.
#include "ecne_host.h"
#include "ecne_abi.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <thread>
#include <mutex>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

// BN254 scalar prime, little-endian limbs (R1CSConstraintSolver.jl:21-22).
const uint64_t P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL,
                       0x30644e72e131a029ULL};

struct Fe {
  uint64_t l[4];
  bool operator==(const Fe& o) const { return memcmp(l, o.l, 32) == 0; }
  bool operator!=(const Fe& o) const { return !(*this == o); }
  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
};
inline int cmp(const Fe& a, const Fe& b) {
  for (int i = 3; i >= 0; --i) {
    if (a.l[i] != b.l[i]) return a.l[i] < b.l[i] ? -1 : 1;
  }
  return 0;
}
inline bool geq_p(const uint64_t* a) {
  for (int i = 3; i >= 0; --i) {
    if (a[i] != P[i]) return a[i] > P[i];
  }
  return true;
}
inline void sub_p(uint64_t* a) {
  unsigned __int128 borrow = 0;
  for (int i = 0; i < 4; ++i) {
    unsigned __int128 d = (unsigned __int128)a[i] - P[i] - (uint64_t)borrow;
    a[i] = (uint64_t)d;
    borrow = (d >> 64) & 1;
  }
}

inline uint32_t rd32(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
inline uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }


// Run fn(begin, end) over [0, n) on the host cores (the reference is single-threaded Julia; this is the
// optional fast path of SURVEY.md §8f, and a 142 MB circuit is mostly memory traffic).
template <class Fn>
void parallel_chunks(uint64_t n, uint64_t min_chunk, Fn fn) {
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 1;
  uint64_t nt = std::min<uint64_t>(hw, std::max<uint64_t>(1, n / std::max<uint64_t>(1, min_chunk)));
  if (nt <= 1) {
    fn((uint64_t)0, n);
    return;
  }
  std::vector<std::thread> th;
  const uint64_t per = (n + nt - 1) / nt;
  for (uint64_t t = 0; t < nt; ++t) {
    const uint64_t b = t * per, e = std::min(n, b + per);
    if (b >= e) break;
    th.emplace_back([=]() { fn(b, e); });
  }
  for (auto& x : th) x.join();
}

// The arrays of a parsed circuit are hundreds of megabytes that are written once, right after they were allocated: with
// 4 KB pages the first-touch page faults are a sizeable part of the read (ecdsa: 49 k faults).  Large arrays are
// aligned to 2 MB and offered to the kernel as huge pages (transparent huge pages in `madvise` mode).
// Freed large arrays are kept (up to ECNE_HOST_CACHE_MB, default 1024) and handed out again: a process that reads one
// circuit after another writes into pages it already owns instead of faulting ~10^5 fresh ones per file (a quarter of
// the read of ecdsa.r1cs).  ecne_host_trim() returns them to the system.
namespace {
struct BigCache {
  std::mutex mu;
  std::unordered_map<void*, size_t> live;       // blocks handed out by big_alloc
  std::multimap<size_t, void*> idle;            // freed blocks by size
  size_t idle_bytes = 0;
  size_t cap() {
    static const size_t c = []() {
      const char* e = getenv("ECNE_HOST_CACHE_MB");
      return (size_t)(e ? std::max(0, atoi(e)) : 1024) << 20;
    }();
    return c;
  }
};
BigCache& big_cache() {
  static BigCache* c = new BigCache();  // (never destroyed: arrays may be freed during interpreter shutdown)
  return *c;
}
}  // namespace
static void* big_alloc(size_t bytes) {
  if (bytes < ((size_t)4 << 20)) return malloc(bytes ? bytes : 1);
  const size_t sz = (bytes + (((size_t)2 << 20) - 1)) & ~(((size_t)2 << 20) - 1);
  BigCache& c = big_cache();
  {
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.idle.lower_bound(sz);
    if (it != c.idle.end() && it->first <= sz + sz / 2) {
      void* p = it->second;
      c.live[p] = it->first;
      c.idle_bytes -= it->first;
      c.idle.erase(it);
      return p;
    }
  }
  void* p = nullptr;
  if (posix_memalign(&p, (size_t)2 << 20, sz) != 0) return nullptr;
#ifdef MADV_HUGEPAGE
  madvise(p, sz, MADV_HUGEPAGE);
#endif
  std::lock_guard<std::mutex> lk(c.mu);
  c.live[p] = sz;
  return p;
}
static void big_free(void* p) {
  if (!p) return;
  BigCache& c = big_cache();
  {
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.live.find(p);
    if (it != c.live.end()) {
      const size_t sz = it->second;
      c.live.erase(it);
      if (c.idle_bytes + sz <= c.cap()) {
        c.idle.emplace(sz, p);
        c.idle_bytes += sz;
        return;
      }
    }
  }
  free(p);
}
extern "C" void ecne_host_trim(void) {
  BigCache& c = big_cache();
  std::lock_guard<std::mutex> lk(c.mu);
  for (auto& kv : c.idle) free(kv.second);
  c.idle.clear();
  c.idle_bytes = 0;
}

// ... the same split with the index of the chunk, and the number of chunks it makes, for per-chunk results
inline uint64_t chunk_count(uint64_t n, uint64_t min_chunk) {
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 1;
  return std::max<uint64_t>(1, std::min<uint64_t>(hw, std::max<uint64_t>(1, n / std::max<uint64_t>(1, min_chunk))));
}
template <class Fn>
void parallel_chunks_idx(uint64_t n, uint64_t nt, Fn fn) {
  if (nt <= 1) {
    fn((uint64_t)0, (uint64_t)0, n);
    return;
  }
  std::vector<std::thread> th;
  const uint64_t per = (n + nt - 1) / nt;
  for (uint64_t t = 0; t < nt; ++t) {
    const uint64_t b = std::min(n, t * per), e = std::min(n, b + per);
    th.emplace_back([=]() { fn(t, b, e); });
  }
  for (auto& x : th) x.join();
}

template <class T>
T* dup_vec(const std::vector<T>& v) {
  T* p = (T*)malloc(std::max<size_t>(1, v.size()) * sizeof(T));
  if (!v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
  return p;
}

ecne_r1cs_t* make_r1cs(const std::vector<uint64_t>& seg, const std::vector<uint32_t>& col,
                       const std::vector<uint64_t>& coef, const std::vector<uint32_t>& known,
                       const std::vector<uint32_t>& targets, uint64_t n_vars) {
  ecne_r1cs_t* r = (ecne_r1cs_t*)calloc(1, sizeof(ecne_r1cs_t));
  r->n_rows = (seg.size() - 1) / 3;
  r->n_vars = n_vars;
  r->nnz = col.size();
  r->seg_ptr = dup_vec(seg);
  r->col = dup_vec(col);
  r->coef = dup_vec(coef);
  r->known = dup_vec(known);
  r->n_known = known.size();
  r->targets = dup_vec(targets);
  r->n_targets = targets.size();
  return r;
}

}  // namespace

extern "C" const char* ecne_host_last_error(void) { return g_err.c_str(); }

extern "C" void ecne_r1cs_free(ecne_r1cs_t* r) {
  if (!r) return;
  big_free(r->seg_ptr);
  big_free(r->col);
  big_free(r->coef);
  big_free(r->known);
  big_free(r->targets);
  big_free(r->coef_class);
  big_free(r->coef_other);
  big_free(r->coef_other_term);
  big_free(r->seg_ptr32);
  free(r);
}

// ecne_read_r1cs_opts: ECNE_READ_COMPACT_ONLY leaves out the full coefficient array (coef == NULL) — a caller that hands
// the rows to the device in the compact form never reads it, and it is 77 % of what the reader writes
thread_local unsigned int g_read_flags = 0;

// ---- the compact form (include/ecne_abi.h) beside the full one -----------------------------------------------------------
static const uint64_t FR_PM1[4] = {0x43e1f593f0000000ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static inline uint8_t coef_class_of(const uint64_t* v) {
  if ((v[1] | v[2] | v[3]) == 0 && v[0] <= 1) return (uint8_t)v[0];
  if (v[0] == FR_PM1[0] && v[1] == FR_PM1[1] && v[2] == FR_PM1[2] && v[3] == FR_PM1[3]) return 2;
  return 3;
}
// ParseR1CS.jl:50-124.  Offsets below are 0-based; the Julia is 1-based.
static int read_r1cs_mem_impl(const uint8_t* arr, uint64_t len, ecne_r1cs_t** out) {
  if (!arr || !out) return fail(ECNE_E_BADARG, "null argument");
  *out = nullptr;
  // overflow-safe: `off + n` may wrap for sizes read from a hostile file
  auto need = [&](uint64_t off, uint64_t n) { return n <= len && off <= len - n; };
  if (!need(0, 12)) return fail(ECNE_E_BOUNDS, "file shorter than the 12-byte header");
  if (rd32(arr + 4) != 1) return fail(ECNE_E_ASSERT, "version != 1 (ParseR1CS.jl:58)");
  uint32_t sections = rd32(arr + 8);
  if (sections != 3) return fail(ECNE_E_ASSERT, "nSections != 3 (ParseR1CS.jl:62)");
  uint64_t cur = 12;
  uint64_t starts[4] = {0, 0, 0, 0};
  bool have[4] = {false, false, false, false};
  for (uint32_t i = 0; i < sections; ++i) {
    if (!need(cur, 12)) return fail(ECNE_E_BOUNDS, "truncated section table");
    uint32_t ty = rd32(arr + cur);
    if (ty < 1 || ty > 3) return fail(ECNE_E_ASSERT, "section type outside 1..3 (ParseR1CS.jl:69)");
    starts[ty] = cur;
    have[ty] = true;
    uint64_t sz = rd64(arr + cur + 4);
    if (sz > len - cur - 12) return fail(ECNE_E_BOUNDS, "section size runs past the end of the file");
    cur += 12 + sz;
  }
  if (!have[1] || !have[2]) return fail(ECNE_E_BOUNDS, "header or constraint section missing");
  uint64_t s1 = starts[1] + 12;
  if (!need(s1, 4)) return fail(ECNE_E_BOUNDS, "truncated header section");
  uint32_t field_size = rd32(arr + s1);
  s1 += 4;
  s1 += field_size;  // the prime is read but never checked (ParseR1CS.jl:84)
  if (!need(s1, 28)) return fail(ECNE_E_BOUNDS, "truncated header section");
  uint32_t n_wires = rd32(arr + s1);
  uint32_t pub_out = rd32(arr + s1 + 4);
  uint32_t pub_in = rd32(arr + s1 + 8);
  uint32_t prv_in = rd32(arr + s1 + 12);
  uint64_t n_labels = rd64(arr + s1 + 16);
  uint32_t n_cons = rd32(arr + s1 + 24);

  uint64_t s2 = starts[2] + 12;
  // every form stores at least its 4-byte term count: a header that announces more constraints than the
  // file can hold is rejected before anything is sized from it
  if (!need(s2, 0) || (uint64_t)n_cons * 12 > len - s2)
    return fail(ECNE_E_BOUNDS, "nConstraints does not fit the file (truncated constraint section)");
  {
    // Fast path: one serial walk over the per-form term counts fixes every offset, then the host cores
    // fill the arrays in place.  A form that repeats a wire (Dict assignment overwrites, ParseR1CS.jl:110)
    // changes the lengths: such files take the serial path below.
    const uint64_t nseg = 3 * (uint64_t)n_cons;
    const bool prof = getenv("ECNE_HOST_PROF") != nullptr;  // stage times on stderr
    auto tp0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
      if (!prof) return;
      auto t = std::chrono::steady_clock::now();
      fprintf(stderr, "[ecne host] read_r1cs   %-22s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - tp0).count());
      tp0 = t;
    };
    uint64_t* segp = (uint64_t*)big_alloc((nseg + 1) * sizeof(uint64_t));
    uint64_t* raw = (uint64_t*)big_alloc(std::max<uint64_t>(1, nseg) * sizeof(uint64_t));
    struct FreeRaw {
      uint64_t* p;
      ~FreeRaw() { big_free(p); }
    } free_raw{raw};
    uint64_t total = 0;
    bool ok = segp != nullptr && raw != nullptr;
    // The offsets of the forms are a chain (each header says where the next one is): 3.3 M dependent loads on
    // ecdsa, the largest serial piece of the host path.  Large files are cut into byte ranges walked concurrently:
    // a worker that does not start at the section's first byte looks for the first offset in its range from which
    // the next 64 headers are plausible (term counts and wire ids below nWires, everything inside the section), and
    // the guess is then PROVED: the chain of the range before must end exactly on it, the last chain exactly on the
    // section's end, and the forms must add up to 3 * nConstraints.  Anything else falls back to the serial walk,
    // so the speculation can only cost time, never change a result.
    bool walked = false;
    const uint64_t sec_end = s2 + rd64(arr + starts[2] + 4);
    unsigned hw = std::thread::hardware_concurrency();
    if (ok && nseg >= (1u << 18) && hw > 1 && sec_end <= len && sec_end > s2) {
      unsigned T = std::min<unsigned>(hw, 32);
      if (const char* e = getenv("ECNE_HOST_WALK_RANGES")) T = (unsigned)std::max(2, std::min(256, atoi(e)));  // testing knob
      struct Part {
        uint64_t start = 0, end_pos = 0, terms = 0;
        bool ok = false;
        std::vector<uint64_t> pos;
      };
      std::vector<Part> parts(T);
      // `steps` headers from c stay inside the section and look sane.  One family of false trails survives any number of
      // steps: in a run of one-term forms with coefficient 1 (a * b = c rows), the chain that starts 8 bytes late reads
      // the coefficient's low word as "one term" and its next word as "wire 0", and lands 8 bytes late in the next form
      // again.  Every wire it sees is 0 — no real stretch of 64 forms mentions nothing but the constant, so a trail
      // without a single signal is not taken (the proof below would catch it anyway, at the price of a serial walk of
      // the whole range: 220 k forms of ecdsa.r1cs with 16 ranges).
      // Another: a word that reads as a term count of a few hundred thousand "terms" and lands on a true header
      // megabytes away — one pseudo-form that swallows the range.  A form longer than half a range is not a guess.
      const uint64_t max_form = std::max<uint64_t>((sec_end - s2) / T / 2, 1 << 16);
      auto plausible_from = [&](uint64_t c, int steps) {
        bool signal = false;
        int k = 0;
        for (; k < steps && c < sec_end; ++k) {
          if (c + 4 > sec_end) return false;
          const uint32_t n = rd32(arr + c);
          if (n > n_wires || (uint64_t)n * 36 > max_form || c + 4 + (uint64_t)n * 36 > sec_end) return false;
          if (n) {
            const uint32_t last = rd32(arr + c + 4 + (uint64_t)(n - 1) * 36);
            if (rd32(arr + c + 4) >= n_wires || last >= n_wires) return false;
            signal |= last != 0;
          }
          c += 4 + (uint64_t)n * 36;
        }
        return signal || k < steps;
      };
      {
        Part* pp = parts.data();
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; ++t)
          th.emplace_back([=, &plausible_from]() {
            Part& P = pp[t];
            const uint64_t lo = s2 + (sec_end - s2) / T * t, hi = t + 1 == T ? sec_end : s2 + (sec_end - s2) / T * (t + 1);
            uint64_t c = lo;
            if (t > 0) {
              while (c < hi && !plausible_from(c, 64)) ++c;
              if (c >= hi) return;  // no header found in the range (a form longer than the range): serial walk
            }
            P.start = c;
            P.pos.reserve((hi - lo) / 48);
            uint64_t terms = 0;
            while (c < hi) {
              if (c + 4 > sec_end) return;
              const uint32_t n = rd32(arr + c);
              __builtin_prefetch(arr + c + 512);  // (a dependent chain of loads: the next headers are a few lines ahead)
              __builtin_prefetch(arr + c + 576);
              if (c + 4 + (uint64_t)n * 36 > sec_end) return;
              P.pos.push_back(c);
              terms += n ? n : 1;
              c += 4 + (uint64_t)n * 36;
            }
            P.end_pos = c;
            P.terms = terms;
            P.ok = true;
          });
        for (auto& x : th) x.join();
      }
      lap("  walk: ranges in parallel");
      // proof by construction: part 0 starts on the section's first header, so its chain is the true one and ends on
      // the first true header of part 1.  From there the true chain is followed until it meets part 1's guessed
      // chain (two chains that share a header are identical from it on): a guess that began a few bytes early —
      // inside the zero bytes of a coefficient, which read as empty forms — or on a false trail is repaired by that
      // short serial prefix, its pseudo-forms dropped; and so on for the next part.  Nothing is trusted that was not
      // reached from a proven header.
      bool good = parts[0].ok && parts[0].start == s2;
      std::vector<std::vector<uint64_t>> pre(T);
      std::vector<uint64_t> skip(T, 0);
      uint64_t forms = good ? parts[0].pos.size() : 0;
      for (unsigned t = 1; t < T && good; ++t) {
        Part& P = parts[t];
        const uint64_t hi = t + 1 == T ? sec_end : s2 + (sec_end - s2) / T * (t + 1);
        if (!P.ok) {  // no usable guess: the whole range is walked here
          P.pos.clear();
          P.terms = 0;
        }
        uint64_t cur = parts[t - 1].end_pos;
        size_t i = 0;
        while (true) {
          while (i < P.pos.size() && P.pos[i] < cur) {
            const uint32_t n = rd32(arr + P.pos[i]);
            P.terms -= n ? n : 1;  // a pseudo-form of the guessed chain
            ++i;
          }
          if (i < P.pos.size() && P.pos[i] == cur) break;  // met the guessed chain: true from here on
          if (cur >= hi) {                                 // never met it inside the range
            P.end_pos = cur;
            break;
          }
          if (cur + 4 > sec_end) { good = false; break; }
          const uint32_t n = rd32(arr + cur);
          if (cur + 4 + (uint64_t)n * 36 > sec_end) { good = false; break; }
          pre[t].push_back(cur);
          P.terms += n ? n : 1;
          cur += 4 + (uint64_t)n * 36;
        }
        skip[t] = i;
        forms += pre[t].size() + (P.pos.size() - i);
      }
      good = good && parts[T - 1].end_pos == sec_end;
      lap("  walk: proof");
      if (prof && getenv("ECNE_HOST_PROF_PARTS"))
        for (unsigned t = 0; t < T; ++t)
          fprintf(stderr, "[ecne host] part %u ok %d start +%llu forms %zu repaired %zu dropped %llu end +%llu\n", t, (int)parts[t].ok,
                  (unsigned long long)(parts[t].start - s2), parts[t].pos.size(), pre[t].size(), (unsigned long long)skip[t],
                  (unsigned long long)(parts[t].end_pos - s2));
      if (prof) {
        uint64_t dropped = 0, walked_here = 0;
        for (unsigned t = 0; t < T; ++t) dropped += skip[t], walked_here += pre[t].size();
        fprintf(stderr, "[ecne host] read_r1cs   %u ranges: %llu pseudo-forms dropped, %llu forms walked serially to repair guesses, proof %s\n",
                T, (unsigned long long)dropped, (unsigned long long)walked_here, good && forms == nseg ? "ok" : "FAILED (serial walk)");
      }
      if (good && forms == nseg) {
        std::vector<uint64_t> form_base(T + 1, 0), term_base(T + 1, 0);
        for (unsigned t = 0; t < T; ++t) {
          form_base[t + 1] = form_base[t] + pre[t].size() + parts[t].pos.size() - skip[t];
          term_base[t + 1] = term_base[t] + parts[t].terms;
        }
        total = term_base[T];
        const Part* pp = parts.data();
        const std::vector<uint64_t>* prep = pre.data();
        const uint64_t *fb = form_base.data(), *tb = term_base.data(), *sk = skip.data();
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; ++t)
          th.emplace_back([=]() {
            // (term counts from the distance to the next header — the file is not touched again: the part's true
            // forms are contiguous, the repaired prefix runs into the kept tail, the tail into the part's end)
            uint64_t g = fb[t], run = tb[t];
            auto put = [&](uint64_t c, uint64_t next) {
              const uint64_t n = (next - c - 4) / 36;
              raw[g] = c;
              segp[g] = run;
              run += n ? n : 1;
              ++g;
            };
            const std::vector<uint64_t>& pr = prep[t];
            const std::vector<uint64_t>& ps = pp[t].pos;
            const size_t k0 = (size_t)sk[t];
            const uint64_t tail_first = k0 < ps.size() ? ps[k0] : pp[t].end_pos;
            for (size_t k = 0; k < pr.size(); ++k) put(pr[k], k + 1 < pr.size() ? pr[k + 1] : tail_first);
            for (size_t k = k0; k < ps.size(); ++k) put(ps[k], k + 1 < ps.size() ? ps[k + 1] : pp[t].end_pos);
          });
        for (auto& x : th) x.join();
        walked = true;
      }
    }
    if (!walked) {
      uint64_t pos = s2;
      total = 0;
      for (uint64_t sgi = 0; ok && sgi < nseg; ++sgi) {
        if (!need(pos, 4)) { ok = false; break; }
        const uint32_t n = rd32(arr + pos);
        if (!need(pos + 4, (uint64_t)n * 36)) { ok = false; break; }
        raw[sgi] = pos;
        segp[sgi] = total;
        total += n ? n : 1;
        pos += 4 + (uint64_t)n * 36;
      }
    }
    lap(walked ? "offset walk (parallel)" : "offset walk (serial)");
    if (ok && total < 0x7fffffffULL) {
      segp[nseg] = total;
      const bool compact_only = (g_read_flags & ECNE_READ_COMPACT_ONLY) != 0 && total < 0xffffffffULL;
      uint32_t* colp = (uint32_t*)big_alloc(std::max<uint64_t>(1, total) * sizeof(uint32_t));
      uint64_t* coefp = compact_only ? nullptr : (uint64_t*)big_alloc(std::max<uint64_t>(1, total) * 4 * sizeof(uint64_t));
      uint8_t* clsp = (uint8_t*)big_alloc(std::max<uint64_t>(1, total));  // compact form: class byte per term
      std::vector<uint8_t> dup_flag(1, 0);
      uint8_t* dupf = dup_flag.data();
      const uint64_t* rawp = raw;
      // pass 1 over contiguous ranges of segments: wires, (full coefficients,) class bytes; per range the number of
      // coefficients that are none of 0, 1, p - 1
      const uint64_t n_ranges = chunk_count(nseg, 1 << 14);
      std::vector<uint64_t> others(n_ranges + 1, 0);
      uint64_t* othersp = others.data();
      parallel_chunks_idx(nseg, n_ranges, [=](uint64_t ri, uint64_t b, uint64_t e) {
        std::vector<uint32_t> tmp;
        uint64_t n3 = 0;
        for (uint64_t sgi = b; sgi < e; ++sgi) {
          const uint8_t* q = arr + rawp[sgi];
          const uint32_t n = rd32(q);
          q += 4;
          uint64_t o = segp[sgi];
          if (n == 0) {  // explicit zero on key 1 (ParseR1CS.jl:113-115)
            colp[o] = 1;
            if (coefp) coefp[4 * o] = coefp[4 * o + 1] = coefp[4 * o + 2] = coefp[4 * o + 3] = 0;
            clsp[o] = 0;
            continue;
          }
          for (uint32_t t = 0; t < n; ++t, q += 36, ++o) {
            uint64_t c[4];
            memcpy(c, q + 4, 32);       // 32 bytes hard-coded (ParseR1CS.jl:109)
            while (geq_p(c)) sub_p(c);  // F(coeff) reduces (ParseR1CS.jl:111)
            colp[o] = rd32(q) + 1;
            if (coefp) memcpy(coefp + 4 * o, c, 32);
            const uint8_t cl = coef_class_of(c);
            clsp[o] = cl;
            n3 += cl == 3;
          }
          if (n > 1) {  // a repeated wire?
            const uint32_t* cc = colp + segp[sgi];
            bool dup = false;
            if (n <= 8) {
              for (uint32_t i = 0; i < n && !dup; ++i)
                for (uint32_t j = i + 1; j < n; ++j)
                  if (cc[i] == cc[j]) { dup = true; break; }
            } else {
              tmp.assign(cc, cc + n);
              std::sort(tmp.begin(), tmp.end());
              for (uint32_t i = 1; i < n; ++i)
                if (tmp[i] == tmp[i - 1]) { dup = true; break; }
            }
            if (dup) *dupf = 1;  // benign race: every writer stores 1
          }
        }
        othersp[ri + 1] = n3;
      });
      lap("fill (parallel)");
      // pass 2 over the same ranges: the other values with their term indices (from the file again: a tenth of the
      // terms), the 32-bit copy of the offsets
      uint64_t* otherp = nullptr;
      uint32_t *termp = nullptr, *seg32p = nullptr;
      if (!dup_flag[0] && total < 0xffffffffULL) {
        for (uint64_t ri = 0; ri < n_ranges; ++ri) others[ri + 1] += others[ri];
        const uint64_t n_other = others[n_ranges];
        otherp = (uint64_t*)big_alloc(std::max<uint64_t>(1, n_other) * 32);
        termp = (uint32_t*)big_alloc(std::max<uint64_t>(1, n_other) * sizeof(uint32_t));
        seg32p = (uint32_t*)big_alloc((nseg + 1) * sizeof(uint32_t));
        if (otherp && termp && seg32p) {
          parallel_chunks_idx(nseg, n_ranges, [=](uint64_t ri, uint64_t b, uint64_t e) {
            uint64_t k = othersp[ri];
            for (uint64_t sgi = b; sgi < e; ++sgi) {
              seg32p[sgi] = (uint32_t)segp[sgi];
              const uint8_t* q = arr + rawp[sgi];
              const uint32_t n = rd32(q);
              q += 4;
              uint64_t o = segp[sgi];
              for (uint32_t t = 0; t < n; ++t, q += 36, ++o)
                if (clsp[o] == 3) {
                  uint64_t c[4];
                  memcpy(c, q + 4, 32);
                  while (geq_p(c)) sub_p(c);
                  memcpy(otherp + 4 * k, c, 32);
                  termp[k++] = (uint32_t)o;
                }
            }
          });
          seg32p[nseg] = (uint32_t)total;
        }
        lap("compact form");
      }
      if (!dup_flag[0]) {
        ecne_r1cs_t* r = (ecne_r1cs_t*)calloc(1, sizeof(ecne_r1cs_t));
        r->n_rows = n_cons;
        r->n_vars = (uint64_t)n_wires + 1;
        r->nnz = total;
        r->seg_ptr = segp;
        r->col = colp;
        r->coef = coefp;
        if (otherp && termp && seg32p) {
          r->coef_class = clsp;
          r->coef_other = otherp;
          r->coef_other_term = termp;
          r->n_coef_other = others[n_ranges];
          r->seg_ptr32 = seg32p;
        } else {
          big_free(clsp);
          big_free(otherp);
          big_free(termp);
          big_free(seg32p);
          if (!coefp) {  // (compact only was asked for and cannot be had: terms beyond 32-bit indices never get here)
            big_free(segp);
            big_free(colp);
            free(r);
            return fail(ECNE_E_BOUNDS, "out of memory for the compact form");
          }
        }
        std::vector<uint32_t> known, targets;
        known.push_back(1);
        for (uint64_t i = 2 + (uint64_t)pub_out; i <= 1 + (uint64_t)pub_out + pub_in + prv_in; ++i)
          known.push_back((uint32_t)i);
        for (uint64_t i = 2; i <= 1 + (uint64_t)pub_out; ++i) targets.push_back((uint32_t)i);
        r->known = dup_vec(known);
        r->n_known = known.size();
        r->targets = dup_vec(targets);
        r->n_targets = targets.size();
        r->n_pub_out = pub_out;
        r->n_pub_in = pub_in;
        r->n_prv_in = prv_in;
        r->field_size = field_size;
        r->n_labels = n_labels;
        *out = r;
        return ECNE_OK;
      }
      big_free(colp);
      big_free(coefp);
      big_free(clsp);
      big_free(otherp);
      big_free(termp);
      big_free(seg32p);
    }
    big_free(segp);  // truncated file or repeated wires: the serial walk below reports / handles it
  }
  std::vector<uint64_t> seg;
  seg.reserve(3 * (size_t)n_cons + 1);
  std::vector<uint32_t> col;
  std::vector<uint64_t> coef;
  uint64_t guess = (len - std::min<uint64_t>(len, s2)) / 36 + 3 * (uint64_t)n_cons;
  col.reserve(guess);
  coef.reserve(guess * 4);
  seg.push_back(0);
  std::vector<std::pair<uint32_t, uint32_t>> order;  // (wire, position) for duplicate handling
  for (uint32_t r = 0; r < n_cons; ++r) {
    for (int form = 0; form < 3; ++form) {
      if (!need(s2, 4)) return fail(ECNE_E_BOUNDS, "truncated constraint section");
      uint32_t n = rd32(arr + s2);
      s2 += 4;
      if (!need(s2, (uint64_t)n * 36)) return fail(ECNE_E_BOUNDS, "truncated constraint section");
      size_t base = col.size();
      for (uint32_t t = 0; t < n; ++t) {
        uint32_t idx = rd32(arr + s2);
        uint64_t c[4];
        memcpy(c, arr + s2 + 4, 32);  // 32 bytes hard-coded (ParseR1CS.jl:109)
        s2 += 36;
        while (geq_p(c)) sub_p(c);  // F(coeff) reduces (ParseR1CS.jl:111)
        col.push_back(idx + 1);
        coef.insert(coef.end(), c, c + 4);
      }
      if (n == 0) {  // explicit zero on key 1 (ParseR1CS.jl:113-115)
        col.push_back(1);
        coef.insert(coef.end(), 4, 0ULL);
      } else if (n > 1) {
        // cur_eq[idx+1] = ... : a repeated wire overwrites the earlier value (Dict assignment).
        bool dup = false;
        if (n <= 8) {
          for (uint32_t i = 0; i < n && !dup; ++i)
            for (uint32_t j = i + 1; j < n; ++j)
              if (col[base + i] == col[base + j]) { dup = true; break; }
        } else {
          order.clear();
          for (uint32_t i = 0; i < n; ++i) order.push_back({col[base + i], i});
          std::sort(order.begin(), order.end());
          for (uint32_t i = 1; i < n; ++i)
            if (order[i].first == order[i - 1].first) { dup = true; break; }
        }
        if (dup) {
          std::map<uint32_t, uint32_t> last;  // wire -> last position
          for (uint32_t i = 0; i < n; ++i) last[col[base + i]] = i;
          std::vector<uint32_t> ncol;
          std::vector<uint64_t> ncoef;
          for (auto& kv : last) {
            ncol.push_back(kv.first);
            ncoef.insert(ncoef.end(), coef.begin() + 4 * (base + kv.second),
                         coef.begin() + 4 * (base + kv.second) + 4);
          }
          col.resize(base);
          coef.resize(4 * base);
          col.insert(col.end(), ncol.begin(), ncol.end());
          coef.insert(coef.end(), ncoef.begin(), ncoef.end());
        }
      }
      seg.push_back(col.size());
    }
  }
  std::vector<uint32_t> known, targets;
  known.push_back(1);
  for (uint64_t i = 2 + (uint64_t)pub_out; i <= 1 + (uint64_t)pub_out + pub_in + prv_in; ++i)
    known.push_back((uint32_t)i);
  for (uint64_t i = 2; i <= 1 + (uint64_t)pub_out; ++i) targets.push_back((uint32_t)i);
  ecne_r1cs_t* r = make_r1cs(seg, col, coef, known, targets, (uint64_t)n_wires + 1);
  r->n_pub_out = pub_out;
  r->n_pub_in = pub_in;
  r->n_prv_in = prv_in;
  r->field_size = field_size;
  r->n_labels = n_labels;
  *out = r;
  return ECNE_OK;
}

// No C++ exception crosses the C ABI: an allocation failure is a status code, as the header promises.
extern "C" int ecne_read_r1cs_mem(const uint8_t* arr, uint64_t len, ecne_r1cs_t** out) {
  try {
    return read_r1cs_mem_impl(arr, len, out);
  } catch (const std::bad_alloc&) {
    if (out) *out = nullptr;
    return fail(ECNE_E_BOUNDS, "out of memory while reading the file (sizes in its header exceed what can be allocated)");
  } catch (const std::exception& e) {
    if (out) *out = nullptr;
    return fail(ECNE_E_INTERNAL, std::string("reader: ") + e.what());
  }
}

extern "C" int ecne_read_r1cs_opts(const char* path, unsigned int flags, ecne_r1cs_t** out);
extern "C" int ecne_read_r1cs(const char* path, ecne_r1cs_t** out) { return ecne_read_r1cs_opts(path, 0, out); }
extern "C" int ecne_read_r1cs_opts(const char* path, unsigned int flags, ecne_r1cs_t** out) {
  if (!path || !out) return fail(ECNE_E_BADARG, "null argument");
  struct FlagGuard {
    FlagGuard(unsigned int f) { g_read_flags = f; }
    ~FlagGuard() { g_read_flags = 0; }
  } flag_guard(flags);
  // map the file instead of copying it: the parser touches every byte exactly once
  int fd = open(path, O_RDONLY);
  if (fd < 0) return fail(ECNE_E_IO, std::string("cannot open ") + path);
  struct stat st;
  if (fstat(fd, &st) != 0) {
    close(fd);
    return fail(ECNE_E_IO, std::string("cannot stat ") + path);
  }
  const size_t sz = (size_t)st.st_size;
  if (sz == 0) {
    close(fd);
    uint8_t none = 0;
    return ecne_read_r1cs_mem(&none, 0, out);
  }
  const bool prof = getenv("ECNE_HOST_PROF") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!prof) return;
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[ecne host] read_r1cs   %-22s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t0).count());
    t0 = t;
  };
  void* m = mmap(nullptr, sz, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
  close(fd);
  if (m == MAP_FAILED) return fail(ECNE_E_IO, std::string("cannot map ") + path);
  lap("mmap (populate)");
  const int st_code = ecne_read_r1cs_mem((const uint8_t*)m, sz, out);
  lap("parse (all stages)");
  munmap(m, sz);
  lap("munmap");
  return st_code;
}

// ------------------------------------------------------------------------------------------
// specials container
// ------------------------------------------------------------------------------------------
extern "C" ecne_specials_t* ecne_specials_new(void) {
  ecne_specials_t* s = (ecne_specials_t*)calloc(1, sizeof(ecne_specials_t));
  s->cap_n = 16;
  s->cap_in = s->cap_out = 64;
  s->kind = (int32_t*)malloc(s->cap_n * sizeof(int32_t));
  s->in_ptr = (uint64_t*)malloc((s->cap_n + 1) * sizeof(uint64_t));
  s->out_ptr = (uint64_t*)malloc((s->cap_n + 1) * sizeof(uint64_t));
  s->in = (uint32_t*)malloc(s->cap_in * sizeof(uint32_t));
  s->out = (uint32_t*)malloc(s->cap_out * sizeof(uint32_t));
  s->in_ptr[0] = s->out_ptr[0] = 0;
  return s;
}
extern "C" void ecne_specials_free(ecne_specials_t* s) {
  if (!s) return;
  free(s->kind);
  free(s->in_ptr);
  free(s->out_ptr);
  free(s->in);
  free(s->out);
  free(s);
}
namespace {
void specials_push(ecne_specials_t* s, int32_t kind, const std::vector<uint32_t>& in,
                   const std::vector<uint32_t>& out) {
  if (s->n + 1 > s->cap_n) {
    s->cap_n *= 2;
    s->kind = (int32_t*)realloc(s->kind, s->cap_n * sizeof(int32_t));
    s->in_ptr = (uint64_t*)realloc(s->in_ptr, (s->cap_n + 1) * sizeof(uint64_t));
    s->out_ptr = (uint64_t*)realloc(s->out_ptr, (s->cap_n + 1) * sizeof(uint64_t));
  }
  uint64_t ni = s->in_ptr[s->n], no = s->out_ptr[s->n];
  while (ni + in.size() > s->cap_in) {
    s->cap_in *= 2;
    s->in = (uint32_t*)realloc(s->in, s->cap_in * sizeof(uint32_t));
  }
  while (no + out.size() > s->cap_out) {
    s->cap_out *= 2;
    s->out = (uint32_t*)realloc(s->out, s->cap_out * sizeof(uint32_t));
  }
  if (!in.empty()) memcpy(s->in + ni, in.data(), in.size() * sizeof(uint32_t));
  if (!out.empty()) memcpy(s->out + no, out.data(), out.size() * sizeof(uint32_t));
  s->kind[s->n] = kind;
  s->n += 1;
  s->in_ptr[s->n] = ni + in.size();
  s->out_ptr[s->n] = no + out.size();
}

// hash_r1cs_equation (:228-235): the row's three sorted coefficient multisets, zeros dropped,
// concatenated WITHOUT separators.  Only equality of hashes is ever used (:262), so any hash of
// that list works; rows are verified form-by-form afterwards (:301-329).
struct RowSig {
  std::vector<Fe> list;  // sortedA ++ sortedB ++ sortedC, non-zero
  uint64_t h;
};
uint64_t mix(uint64_t h, uint64_t v) {
  h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
  h *= 0xff51afd7ed558ccdULL;
  return h ^ (h >> 33);
}
void nonzero_sorted(const ecne_r1cs_t* r, uint64_t seg, std::vector<Fe>& out) {
  size_t b = out.size();
  for (uint64_t t = r->seg_ptr[seg]; t < r->seg_ptr[seg + 1]; ++t) {
    Fe f;
    memcpy(f.l, r->coef + 4 * t, 32);
    if (!f.is_zero()) out.push_back(f);
  }
  std::sort(out.begin() + b, out.end(), [](const Fe& x, const Fe& y) { return cmp(x, y) < 0; });
}
uint64_t row_hash(const ecne_r1cs_t* r, uint64_t row, std::vector<Fe>& tmp) {
  tmp.clear();
  for (int f = 0; f < 3; ++f) nonzero_sorted(r, 3 * row + f, tmp);
  uint64_t h = 0x1234567ULL + tmp.size();
  for (auto& x : tmp)
    for (int i = 0; i < 4; ++i) h = mix(h, x.l[i]);
  return h;
}
// checkNonZeroValues (:205-226): same multiset of non-zero values.
bool same_nonzero_multiset(const ecne_r1cs_t* a, uint64_t sa, const ecne_r1cs_t* b, uint64_t sb,
                           std::vector<Fe>& t1, std::vector<Fe>& t2) {
  t1.clear();
  t2.clear();
  nonzero_sorted(a, sa, t1);
  nonzero_sorted(b, sb, t2);
  if (t1.size() != t2.size()) return false;
  for (size_t i = 0; i < t1.size(); ++i)
    if (t1[i] != t2[i]) return false;
  return true;
}

// Appearance signature of one variable: [(slot, coeff)...] in slot order (:282-292, :305-310).  Flat form:
// the window's non-zero terms as (wire, slot, coefficient) sorted by wire — a stable sort, so every wire's
// terms stay in slot order — one contiguous range per wire, and the wires ordered by signature (ties by
// wire id — Julia's Dict order is unpinned, SURVEY.md §7 hard part 7).  No per-wire heap vectors, no hash map:
// a candidate window of secp256k1 (15 935 rows) is verified in a third of the time.
struct Term {
  uint32_t var, slot;
  const uint64_t* c;  // 4 limbs inside the system's coef array
};
struct Appearance {
  std::vector<Term> terms;
  std::vector<uint32_t> start;  // [nv + 1] ranges into terms, one per distinct wire
  std::vector<uint32_t> order;  // wires (as range indices) sorted by signature
  size_t n_vars() const { return start.empty() ? 0 : start.size() - 1; }
  uint32_t var(uint32_t k) const { return terms[start[k]].var; }
};
inline int cmp_limbs(const uint64_t* a, const uint64_t* b) {
  for (int i = 3; i >= 0; --i)
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  return 0;
}
int cmp_sig(const Appearance& A, uint32_t ka, const Appearance& B, uint32_t kb) {
  const Term *a = A.terms.data() + A.start[ka], *b = B.terms.data() + B.start[kb];
  const size_t na = A.start[ka + 1] - A.start[ka], nb = B.start[kb + 1] - B.start[kb];
  const size_t n = std::min(na, nb);
  for (size_t i = 0; i < n; ++i) {
    if (a[i].slot != b[i].slot) return a[i].slot < b[i].slot ? -1 : 1;
    int c = cmp_limbs(a[i].c, b[i].c);
    if (c) return c;
  }
  if (na != nb) return na < nb ? -1 : 1;
  return 0;
}
void appearance(const ecne_r1cs_t* r, uint64_t row0, uint64_t n, Appearance& out) {
  out.terms.clear();
  out.start.clear();
  out.order.clear();
  uint32_t slot = 0;
  for (uint64_t j = 0; j < n; ++j) {
    for (int f = 0; f < 3; ++f) {
      ++slot;
      const uint64_t s = 3 * (row0 + j) + f;
      for (uint64_t t = r->seg_ptr[s]; t < r->seg_ptr[s + 1]; ++t) {
        const uint64_t* c = r->coef + 4 * t;
        if ((c[0] | c[1] | c[2] | c[3]) == 0) continue;
        out.terms.push_back(Term{r->col[t], slot, c});
      }
    }
  }
  std::stable_sort(out.terms.begin(), out.terms.end(), [](const Term& a, const Term& b) { return a.var < b.var; });
  for (size_t i = 0; i < out.terms.size(); ++i)
    if (i == 0 || out.terms[i].var != out.terms[i - 1].var) out.start.push_back((uint32_t)i);
  if (out.terms.empty()) return;
  out.start.push_back((uint32_t)out.terms.size());
  const size_t nv = out.start.size() - 1;
  out.order.resize(nv);
  for (size_t k = 0; k < nv; ++k) out.order[k] = (uint32_t)k;
  std::sort(out.order.begin(), out.order.end(), [&](uint32_t a, uint32_t b) {
    int c = cmp_sig(out, a, out, b);
    if (c) return c < 0;
    return out.var(a) < out.var(b);
  });
}
}  // namespace

// abstraction (:237-395).
static int abstraction_impl(int32_t kind, const ecne_r1cs_t* cons, const ecne_r1cs_t* sub,
                            ecne_r1cs_t** reduced, ecne_specials_t* specials, uint64_t* n_matches);
extern "C" int ecne_abstraction(int32_t kind, const ecne_r1cs_t* cons, const ecne_r1cs_t* sub,
                                ecne_r1cs_t** reduced, ecne_specials_t* specials,
                                uint64_t* n_matches) {
  try {
    return abstraction_impl(kind, cons, sub, reduced, specials, n_matches);
  } catch (const std::bad_alloc&) {
    if (reduced) *reduced = nullptr;
    return fail(ECNE_E_INTERNAL, "out of memory in abstraction()");
  } catch (const std::exception& e) {
    if (reduced) *reduced = nullptr;
    return fail(ECNE_E_INTERNAL, std::string("abstraction: ") + e.what());
  }
}
static int abstraction_impl(int32_t kind, const ecne_r1cs_t* cons, const ecne_r1cs_t* sub,
                            ecne_r1cs_t** reduced, ecne_specials_t* specials, uint64_t* n_matches) {
  if (!cons || !sub || !reduced || !specials) return fail(ECNE_E_BADARG, "null argument");
  *reduced = nullptr;
  const uint64_t N = cons->n_rows, n = sub->n_rows;
  const bool prof = getenv("ECNE_HOST_PROF") != nullptr;  // stage times on stderr
  auto tp0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!prof) return;
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[ecne host] abstraction %-22s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - tp0).count());
    tp0 = t;
  };
  std::vector<Fe> t1, t2;
  std::vector<uint64_t> hc(N), hs(n);
  {  // row hashes on the host cores (each worker with its own scratch)
    uint64_t* hcp = hc.data();
    parallel_chunks(N, 1 << 13, [=](uint64_t b, uint64_t e) {
      std::vector<Fe> tmp;
      for (uint64_t i = b; i < e; ++i) hcp[i] = row_hash(cons, i, tmp);
    });
  }
  for (uint64_t i = 0; i < n; ++i) hs[i] = row_hash(sub, i, t1);

  lap("row hashes");
  // candidates: the first n-1 row hashes line up (:259-270)
  std::vector<uint64_t> candidates;
  if (N + 1 >= n + 1 && N >= n) {
    for (uint64_t i = 0; i + n <= N; ++i) {
      bool ok = true;
      for (uint64_t j = 0; j + 1 < n; ++j) {
        if (hc[i + j] != hs[j]) {
          ok = false;
          break;
        }
      }
      if (ok) candidates.push_back(i);
    }
  }
  lap("candidate scan");
  Appearance orig;
  appearance(sub, 0, n, orig);
  struct Match {
    uint64_t start;
    bool works;
    std::vector<std::pair<uint32_t, uint32_t>> map;  // (sub wire, main wire), sorted by sub wire (:351)
    const uint32_t* find(uint32_t x) const {
      auto it = std::lower_bound(map.begin(), map.end(), std::make_pair(x, (uint32_t)0));
      return it != map.end() && it->first == x ? &it->second : nullptr;
    }
  };
  // every candidate is verified independently (:301-351): coefficient multisets per slot, then the
  // per-variable appearance signatures
  std::vector<Match> verified(candidates.size());
  {
    Match* vp = verified.data();
    const uint64_t* cp = candidates.data();
    const Appearance* origp = &orig;
    parallel_chunks(candidates.size(), 1, [=](uint64_t b, uint64_t e) {
      std::vector<Fe> u1, u2;
      Appearance curv;
      for (uint64_t ci = b; ci < e; ++ci) {
        const uint64_t i = cp[ci];
        Match& m = vp[ci];
        m.start = i;
        m.works = true;
        for (uint64_t j = 0; j < n && m.works; ++j)
          for (int f = 0; f < 3; ++f)
            if (!same_nonzero_multiset(cons, 3 * (i + j) + f, sub, 3 * j + f, u1, u2)) {
              m.works = false;
              break;
            }
        if (!m.works) continue;
        appearance(cons, i, n, curv);
        const size_t nv = curv.n_vars();
        if (nv != origp->n_vars()) {
          m.works = false;
          continue;
        }
        for (size_t k = 0; k < nv; ++k)
          if (cmp_sig(curv, curv.order[k], *origp, origp->order[k]) != 0) {
            m.works = false;
            break;
          }
        if (!m.works) continue;
        m.map.reserve(nv);
        for (size_t k = 0; k < nv; ++k) m.map.emplace_back(origp->var(origp->order[k]), curv.var(curv.order[k]));
        std::sort(m.map.begin(), m.map.end());
      }
    });
  }
  lap("verify candidates");
  std::vector<Match> matches;
  for (auto& m : verified)
    if (m.works) matches.push_back(std::move(m));

  // the walk (:357-388), including the stall after an overlapping match (:370): first which windows it
  // consumes, then the output is sized once and the kept runs of rows are copied
  std::vector<uint64_t> consumed;  // indices into matches
  {
    size_t cur = 0;
    uint64_t i = 0;
    while (i < N && cur < matches.size()) {
      if (matches[cur].start < i) break;  // starts inside a consumed window: cur never advances again
      i = matches[cur].start + n;         // rows up to the window are kept, the window is replaced
      consumed.push_back(cur);
      cur += 1;
    }
  }
  uint64_t added = 0;
  for (uint64_t ci : consumed) {
    std::vector<uint32_t> in, outv;
    for (uint64_t k = 0; k < sub->n_known; ++k) {
      uint32_t x = sub->known[k];
      if (x == 1) continue;
      const uint32_t* it = matches[ci].find(x);
      if (!it) return fail(ECNE_E_KEYERROR, "KeyError: trusted input wire never appears (:381)");
      in.push_back(*it);
    }
    for (uint64_t k = 0; k < sub->n_targets; ++k) {
      const uint32_t* it = matches[ci].find(sub->targets[k]);
      if (!it) return fail(ECNE_E_KEYERROR, "KeyError: trusted output wire never appears (:382)");
      outv.push_back(*it);
    }
    specials_push(specials, kind, in, outv);
    ++added;
  }
  lap("walk + specials");
  const uint64_t rows_out = N - (uint64_t)consumed.size() * n;
  uint64_t terms_out = cons->seg_ptr[3 * N];
  for (uint64_t ci : consumed)
    terms_out -= cons->seg_ptr[3 * (matches[ci].start + n)] - cons->seg_ptr[3 * matches[ci].start];
  // The kept rows go straight into the arrays the caller will own: no zero-filled staging vectors and no
  // second copy (those were 250 of the 430 ms of ecdsa's abstraction), and the host cores share the copy —
  // first touch of ~110 MB of fresh pages is most of its cost.
  ecne_r1cs_t* r = (ecne_r1cs_t*)calloc(1, sizeof(ecne_r1cs_t));
  r->n_rows = rows_out;
  r->n_vars = cons->n_vars;
  r->nnz = terms_out;
  r->seg_ptr = (uint64_t*)big_alloc((3 * rows_out + 1) * sizeof(uint64_t));
  r->col = (uint32_t*)big_alloc(std::max<uint64_t>(1, terms_out) * sizeof(uint32_t));
  r->coef = (uint64_t*)big_alloc(std::max<uint64_t>(1, terms_out) * 4 * sizeof(uint64_t));
  struct Run {
    uint64_t r0, r1, row_dst, term_dst;  // rows [r0, r1) of cons land at row_dst / term_dst
  };
  std::vector<Run> runs;
  {
    uint64_t row_dst = 0, term_dst = 0, row_src = 0;
    auto keep = [&](uint64_t r0, uint64_t r1) {
      if (r1 <= r0) return;
      runs.push_back(Run{r0, r1, row_dst, term_dst});
      row_dst += r1 - r0;
      term_dst += cons->seg_ptr[3 * r1] - cons->seg_ptr[3 * r0];
    };
    for (uint64_t ci : consumed) {
      keep(row_src, matches[ci].start);
      row_src = matches[ci].start + n;
    }
    keep(row_src, N);
    r->seg_ptr[3 * rows_out] = term_dst;
  }
  {
    const Run* rp = runs.data();
    const size_t nr = runs.size();
    uint64_t* seg = r->seg_ptr;
    uint32_t* col = r->col;
    uint64_t* coef = r->coef;
    parallel_chunks(rows_out, 1 << 14, [=](uint64_t b, uint64_t e) {  // output rows [b, e)
      for (size_t k = 0; k < nr; ++k) {
        const Run& u = rp[k];
        const uint64_t lo = std::max(b, u.row_dst), hi = std::min(e, u.row_dst + (u.r1 - u.r0));
        if (lo >= hi) continue;
        const uint64_t a0 = u.r0 + (lo - u.row_dst), a1 = u.r0 + (hi - u.row_dst);  // source rows
        const uint64_t t0 = cons->seg_ptr[3 * a0], t1e = cons->seg_ptr[3 * a1];
        const uint64_t td = u.term_dst + (t0 - cons->seg_ptr[3 * u.r0]);
        memcpy(col + td, cons->col + t0, (t1e - t0) * sizeof(uint32_t));
        memcpy(coef + 4 * td, cons->coef + 4 * t0, (t1e - t0) * 4 * sizeof(uint64_t));
        for (uint64_t sgi = 3 * a0; sgi < 3 * a1; ++sgi) seg[3 * lo + (sgi - 3 * a0)] = cons->seg_ptr[sgi] - t0 + td;
      }
    });
  }
  lap("copy kept rows");
  r->known = (uint32_t*)malloc(std::max<uint64_t>(1, cons->n_known) * sizeof(uint32_t));
  if (cons->n_known) memcpy(r->known, cons->known, cons->n_known * sizeof(uint32_t));
  r->n_known = cons->n_known;
  r->targets = (uint32_t*)malloc(std::max<uint64_t>(1, cons->n_targets) * sizeof(uint32_t));
  if (cons->n_targets) memcpy(r->targets, cons->targets, cons->n_targets * sizeof(uint32_t));
  r->n_targets = cons->n_targets;
  r->n_pub_out = cons->n_pub_out;
  r->n_pub_in = cons->n_pub_in;
  r->n_prv_in = cons->n_prv_in;
  r->field_size = cons->field_size;
  r->n_labels = cons->n_labels;
  *reduced = r;
  if (n_matches) *n_matches = added;
  return ECNE_OK;
}

// ---- compact coefficients (include/ecne_host.h) ---------------------------------------------------------------------
extern "C" int ecne_compact_coef(const uint64_t* coef, uint64_t nnz, uint8_t* cls, uint64_t* other, uint32_t* other_term,
                                 uint64_t* n_other) {
  if ((nnz && (!coef || !cls)) || !n_other || (other == nullptr) != (other_term == nullptr)) {
    g_err = "ecne_compact_coef: null argument";
    return ECNE_E_BADARG;
  }
  if (nnz >= 0xffffffffULL) {
    g_err = "ecne_compact_coef: term indices are 32-bit";
    return ECNE_E_BADARG;
  }
  try {
    // BN254 scalar field modulus minus one, little-endian limbs
    static const uint64_t PM1[4] = {0x43e1f593f0000000ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    const uint64_t CH = 1 << 16;
    const uint64_t n_chunks = (nnz + CH - 1) / CH;
    std::vector<uint64_t> cnt(n_chunks + 1, 0);
    parallel_chunks(n_chunks, 4, [&, coef, cls](uint64_t cb, uint64_t ce) {
      for (uint64_t c = cb; c < ce; ++c) {
        uint64_t k = 0;
        const uint64_t te = std::min(nnz, (c + 1) * CH);
        for (uint64_t t = c * CH; t < te; ++t) {
          const uint64_t* v = coef + 4 * t;
          uint8_t cl = 3;
          if ((v[1] | v[2] | v[3]) == 0 && v[0] <= 1)
            cl = (uint8_t)v[0];
          else if (v[0] == PM1[0] && v[1] == PM1[1] && v[2] == PM1[2] && v[3] == PM1[3])
            cl = 2;
          cls[t] = cl;
          k += cl == 3;
        }
        cnt[c + 1] = k;
      }
    });
    for (uint64_t c = 0; c < n_chunks; ++c) cnt[c + 1] += cnt[c];
    if (other) {
      if (*n_other != cnt[n_chunks]) {
        g_err = "ecne_compact_coef: *n_other does not match the array (call with other == NULL first)";
        return ECNE_E_BADARG;
      }
      parallel_chunks(n_chunks, 4, [&, coef, cls, other, other_term](uint64_t cb, uint64_t ce) {
        for (uint64_t c = cb; c < ce; ++c) {
          uint64_t k = cnt[c];
          const uint64_t te = std::min(nnz, (c + 1) * CH);
          for (uint64_t t = c * CH; t < te; ++t)
            if (cls[t] == 3) {
              memcpy(other + 4 * k, coef + 4 * t, 32);
              other_term[k++] = (uint32_t)t;
            }
        }
      });
    }
    *n_other = cnt[n_chunks];
    return ECNE_OK;
  } catch (const std::bad_alloc&) {
    g_err = "ecne_compact_coef: out of memory";
    return ECNE_E_BOUNDS;
  }
}
