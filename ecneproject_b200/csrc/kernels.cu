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

#define P1_THREADS 256

__global__ void __launch_bounds__(P1_THREADS) k_p1_loop(Dev d, int rbuf0, unsigned int max_rounds) {
  unsigned int epoch = 0;
  int rbuf = rbuf0;
  unsigned int list = 0;          // list written this round
  unsigned int prev_n = 0;        // records of the previous round (to replay)
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nthreads = gridDim.x * blockDim.x;
  const uint32_t rows = d.row_hi - d.row_lo;
  unsigned int round = 0;
  while (true) {
    const int wbuf = rbuf ^ 1;
    // (a) replay the previous round's records into the buffer written this round
    if (prev_n) {
      const Rec* pr = d.recs[(list + 2) % 3];
      for (uint32_t i = tid; i < prev_n; i += nthreads) {
        Rec r = pr[i];
        apply_update(d, wbuf, r.wire, r.bits, r.lbr, r.ubr);
      }
    }
    // (b) sweep: one thread per ordinary row, tiles of blockDim rows round-robin over the grid
    for (uint32_t base = blockIdx.x * blockDim.x; base < rows; base += nthreads) {
      uint32_t r = base + threadIdx.x;
      if (r < rows) {
        uint32_t row = d.row_lo + r;
        if (!(d.rflags[row] & RF_LONG)) eval_row<1>(d, rbuf, wbuf, (int)list, row);
      }
    }
    // long rows: one warp each
    {
      const uint32_t warp = tid >> 5, nwarps = nthreads >> 5;
      for (uint32_t i = warp; i < d.n_long; i += nwarps) {
        uint32_t row = d.long_rows[i];
        if (row >= d.row_lo && row < d.row_hi) eval_row<32>(d, rbuf, wbuf, (int)list, row);
      }
    }
    grid_barrier(d.barrier, epoch);
    unsigned int n = *((volatile unsigned int*)(d.rec_count + list));
    round += 1;
    if (tid == 0) {
      d.st->rounds += 1;
      d.st->evals += rows;
      d.st->changed += n;
      d.rec_count[(list + 2) % 3] = 0;  // replayed during this round; next written in two rounds
    }
    if (n > d.rec_cap) n = d.rec_cap;
    if (n == 0) {
      // W already holds every earlier record (replayed above): both buffers are complete
      break;
    }
    if (round >= max_rounds) {
      if (tid == 0) raise(d, ECNE_E_NOCONVERGE);
      // make the buffers consistent before leaving
      const int nb = rbuf;  // old read buffer lacks this round's records
      const Rec* pr = d.recs[list];
      for (uint32_t i = tid; i < n; i += nthreads) {
        Rec r = pr[i];
        apply_update(d, nb, r.wire, r.bits, r.lbr, r.ubr);
      }
      break;
    }
    prev_n = n;
    rbuf = wbuf;
    list = (list + 1) % 3;
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

// ---- P2 (:1357-1417) ------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

// One thread per row: does the row qualify (every non-unique wire appears in C only)?  k == 1 is
// decided on the spot; k >= 2 rows become (set-hash, row) candidates.
__global__ void k_p2_scan(Dev d, int rbuf) {
  uint32_t row = d.row_lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= d.row_hi) return;
  const uint8_t* F = d.F[rbuf];
  const uint32_t s0 = d.seg[3 * row], s2 = d.seg[3 * row + 2], s3 = d.seg[3 * row + 3];
  for (uint32_t t = s0; t < s2; ++t)
    if (!(ld_flag(F, d.col[t]) & WF_U)) return;  // a non-unique wire in A or B (:1366, :1378)
  uint32_t k = 0, w1 = 0;
  unsigned long long hs = 0, hx = 0;
  for (uint32_t t = s2; t < s3; ++t) {
    uint32_t w = d.col[t];
    if (!(ld_flag(F, w) & WF_U)) {
      ++k;
      w1 = w;
      unsigned long long m = mix64(w);
      hs += m;
      hx ^= mix64(m + 0x9e3779b97f4a7c15ULL);
    }
  }
  if (k == 0) return;
  if (k == 1) {  // 1x1 "matrix": the stored coefficient is non-zero (:1402)
    emit(d, rbuf ^ 1, 0, w1, WF_U | WF_K);
    return;
  }
  unsigned int i = atomicAdd(&d.st->p2_cand, 1u);
  d.p2_key[i] = mix64(hs ^ (hx * 0x9e3779b97f4a7c15ULL) ^ k);
  d.p2_row[i] = row;
}

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

cudaError_t launch_reset(const Dev& d, cudaStream_t s) {
  k_reset_wires<<<blocks_for((uint64_t)d.V + 4, 256), 256, 0, s>>>(d);
  if (d.n_known) k_reset_known<<<blocks_for(d.n_known, 256), 256, 0, s>>>(d);
  cudaMemsetAsync(d.solved, 0, (size_t)d.N + 1, s);
  if (d.n_specials) cudaMemsetAsync(d.sp_solved, 0, d.n_specials, s);
  cudaMemsetAsync(d.rec_count, 0, 3 * sizeof(unsigned int), s);
  cudaMemsetAsync(d.st, 0, sizeof(Status), s);
  return cudaGetLastError();
}

int p1_grid_size(int device) {
  static int cached = 0;
  if (cached) return cached;
  int sms = 0, per_sm = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_p1_loop, P1_THREADS, 0);
  if (per_sm < 1) per_sm = 1;
  cached = sms * per_sm;
  return cached;
}

cudaError_t launch_p1(const Dev& d, int rbuf, unsigned int max_rounds, int grid, cudaStream_t s) {
  cudaMemsetAsync(d.barrier, 0, sizeof(unsigned int), s);
  Dev dd = d;
  int rb = rbuf;
  unsigned int mr = max_rounds;
  void* args[] = {&dd, &rb, &mr};
  return cudaLaunchCooperativeKernel((void*)k_p1_loop, dim3(grid), dim3(P1_THREADS), args, 0, s);
}

void launch_replay(const Dev& d, int buf, cudaStream_t s) {
  k_replay<<<148, 256, 0, s>>>(d, buf);
  k_replay_done<<<1, 1, 0, s>>>(d);
}

void launch_p0(const Dev& d, cudaStream_t s) {
  if (d.n_specials) k_p0<<<1, 256, 0, s>>>(d);
}
void launch_p2_scan(const Dev& d, int rbuf, cudaStream_t s) {
  uint32_t rows = d.row_hi - d.row_lo;
  if (rows) k_p2_scan<<<blocks_for(rows, 256), 256, 0, s>>>(d, rbuf);
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
