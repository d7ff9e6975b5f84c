// engine_host.h — host-side declarations shared by the translation units of libecne_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <functional>
#include <string>
#include <vector>

#include "engine.cuh"

namespace ecne {

// kernels.cu
cudaError_t launch_reset(const Dev& d, int grid, cudaStream_t s);
cudaError_t launch_clear_p2_table(const Dev& d, cudaStream_t s);
int p1_grid_size(int device);
int p1_threads();
// the whole fixpoint (:706-1556): one persistent cooperative launch
cudaError_t launch_solve(const Dev& d, unsigned int max_rounds, int grid, cudaStream_t s);
void launch_finalize(const Dev& d, int buf, unsigned long long* ubits, unsigned long long* kbits,
                     unsigned long long* counts, cudaStream_t s);
void launch_export(const Dev& d, int buf, fr::u256* lb, fr::u256* ub, uint8_t* nvalues,
                   fr::u256* values, cudaStream_t s);

// report path (:1599-1635): rows that mention a non-unique wire (bitmap), their wires in ascending order
// (counts[0] = rows listed, counts[1] = wires listed), then the state of those wires
void launch_bad_rows(const Dev& d, int buf, unsigned int* row_bits, unsigned int* wire_mark, uint32_t* wires,
                     unsigned long long* counts, cudaStream_t s);
void launch_report_export(const Dev& d, int buf, const uint32_t* wires, uint32_t n, uint8_t* flags, fr::u256* lb,
                          fr::u256* ub, uint8_t* nvalues, fr::u256* values, cudaStream_t s);

// Device memory is taken from a process-wide pool of big slabs that survives between calls
// (SURVEY.md §8b "device memory owned by the library, cached between calls, freed in
// ecne_shutdown"): an Arena bump-allocates out of slabs it borrows from the pool and hands them back
// on release(), so a steady-state ecne_solve() performs no cudaMalloc / cudaFree at all.
struct Slab {
  char* base = nullptr;
  size_t size = 0;
};
struct SlabPool {
  std::vector<Slab> free_slabs;
  cudaError_t acquire(size_t min_bytes, Slab* out) {
    // best fit among the cached slabs
    int best = -1;
    for (size_t i = 0; i < free_slabs.size(); ++i)
      if (free_slabs[i].size >= min_bytes && (best < 0 || free_slabs[i].size < free_slabs[best].size))
        best = (int)i;
    if (best >= 0) {
      *out = free_slabs[best];
      free_slabs.erase(free_slabs.begin() + best);
      return cudaSuccess;
    }
    size_t sz = min_bytes < ((size_t)256 << 20) ? ((size_t)256 << 20) : min_bytes;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, sz);
    if (e != cudaSuccess) return e;
    out->base = (char*)p;
    out->size = sz;
    return cudaSuccess;
  }
  void give_back(const Slab& s) { free_slabs.push_back(s); }
  void destroy() {
    for (auto& s : free_slabs) cudaFree(s.base);
    free_slabs.clear();
  }
};
// One pool per device this process drives (ecne_init: one; ecne_init_multi: several); an Arena allocates from the
// pool of the device that is current when it makes its first allocation.
SlabPool& slab_pool();

struct Arena {
  SlabPool* pool = nullptr;
  std::vector<Slab> slabs;
  size_t off = 0;  // bump offset inside slabs.back()
  size_t bytes = 0;
  template <class T>
  cudaError_t alloc(T** out, size_t n) {
    size_t sz = ((n ? n : 1) * sizeof(T) + 255) & ~(size_t)255;
    if (slabs.empty() || off + sz > slabs.back().size) {
      Slab s;
      if (!pool) pool = &slab_pool();
      cudaError_t e = pool->acquire(sz, &s);
      if (e != cudaSuccess) return e;
      slabs.push_back(s);
      off = 0;
    }
    *out = (T*)(slabs.back().base + off);
    off += sz;
    bytes += sz;
    return cudaSuccess;
  }
  void release() {
    for (auto& s : slabs) pool->give_back(s);
    slabs.clear();
    off = 0;
    bytes = 0;
  }
};

// One uploaded + classified problem (ecne_resident in the ABI).
struct Resident {
  Dev d;
  Arena arena;
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;  // side stream of the set-up (bound-table chain)
  cudaStream_t side2 = nullptr; // ... a second one (the long rows' classification and layout)
  int device = 0;
  // host mirrors needed by the solve loop
  uint64_t n_rows = 0, n_vars = 0, n_targets = 0;
  // finalisation buffers
  unsigned long long *d_ubits = nullptr, *d_kbits = nullptr, *d_counts = nullptr;
  // CUB temp storage (set-up only)
  void* d_cub = nullptr;
  size_t cub_bytes = 0;
  bool table_dirty = false;  // a failed solve may leave entries in the P2 grouping table
  bool have_state = false;   // buffer 0 holds the final state of a successful solve (report path)
  // pinned host staging
  Status* h_status = nullptr;
  unsigned long long* h_counts = nullptr;
  double ms_h2d = 0, ms_classify = 0;
  // multi-GPU: exchange buffer header (mailbox, counts, epoch) to clear at the start of a solve
  void* xhdr = nullptr;
};

// A constraint system in its on-disk layout (explicit zeros included) resident on the device: what the
// classification reads, and what abstraction() on the device (abstraction.cu) consumes and produces.
struct DevSystem {
  uint64_t N = 0, V = 0, nnz = 0;
  unsigned long long* seg = nullptr;  // [3N + 1]
  uint32_t* col = nullptr;
  fr::u256* coef = nullptr;
  Arena arena;
};
struct SpecialsHost {  // the special constraints found so far (:357-384), CSR over specials
  std::vector<int32_t> kind;
  std::vector<uint64_t> in_ptr{0}, out_ptr{0};
  std::vector<uint32_t> in, out;
};
struct AbstractionStats {
  double ms_prepare = 0, ms_hash = 0, ms_candidates = 0, ms_verify = 0, ms_compact = 0;
  uint64_t n_candidates = 0, n_matches = 0;
};
#ifndef ECNE_E_KEYERROR
#define ECNE_E_KEYERROR (-10)  // KeyError at R1CSConstraintSolver.jl:381-382 (same code as include/ecne_host.h)
#endif
// abstraction.cu: host -> device copy; pageable sources of 1 MB and more go through a ring of pinned slots filled by
// worker threads, pinned ones (and small ones) straight into cudaMemcpyAsync
cudaError_t staged_h2d(void* dst, const void* src, size_t bytes, cudaStream_t s);
cudaError_t staged_h2d_async(void* dst, const void* src, size_t bytes, cudaStream_t s);  // queue only ...
cudaError_t staged_flush();  // ... every queued slice is on its stream when this returns
// setup.cu: the rows of `p` (seg_ptr / col / coef in either form of include/ecne_abi.h: full 32-byte coefficients, or
// class bytes + the values that are not 0, 1, p-1; 64- or 32-bit offsets) into device arrays of the on-disk layout.
// `tmp` lends the scratch of the compact form.  Returns an ecne_status.
// One process per GPU (ecne_dist_init): rank r copies the r-th slice of every big upload over its OWN PCIe link and the
// ranks all-gather the slices over NVLink — the host -> device traffic of a solve is the problem once, not once per GPU.
// abi.cu installs the gather (ncclAllGather, in place: slice r lies at buf + r * slice_bytes).  Collective: every rank
// uploads the same problem at the same time, which ecne_solve / ecne_upload on several ranks already require.
struct UploadShard {
  int rank = 0, world = 1;
  std::function<cudaError_t(void* buf, size_t slice_bytes, cudaStream_t s)> allgather;
};
UploadShard& upload_shard();
// elements to allocate for an upload target of n elements (room for `world` equal 256-byte-aligned slices)
template <class T>
inline size_t upload_padded(size_t n) {
  return n + (ECNE_MAX_WORLD * 256 + 256) / sizeof(T) + 1;
}
uint64_t problem_nnz(const ecne_problem_t* p);
int problem_rows_ok(const ecne_problem_t* p, std::string& err);
// (s_col / ev_col: when given, the wire ids — the largest array, needed last — cross on that stream behind the others and
// `ev_col` is recorded there, so that the kernels that expand the coefficients run while they are still in flight)
int upload_rows(const ecne_problem_t* p, unsigned long long* d_seg, uint32_t* d_col, fr::u256* d_coef, Arena& tmp,
                cudaStream_t s, std::string& err, cudaStream_t s_col = nullptr, cudaEvent_t ev_col = nullptr);
int dev_system_upload(const ecne_problem_t* p, DevSystem* S, cudaStream_t s, std::string& err);
// `ready`: called after the trusted circuit has been prepared on the host and before the first kernel touches `S`
// (the upload of the big system may still be running until then); returns an ecne_status.
// Everything about a trusted circuit that does not depend on the big system (coefficient multisets per form, wire
// signatures and their classes): host work of a few milliseconds that touches neither CUDA nor any global state, so a
// host may prepare the trusted circuits on another thread while it still reads the main circuit.
struct PreparedSub;
PreparedSub* abstraction_prepare(const ecne_problem_t* sub);  // nullptr: bad argument or no usable signature seed
void abstraction_prepared_free(PreparedSub* p);
// `prepared`: what abstraction_prepare made of the same `sub` (nullptr: prepared here)
int dev_abstraction(DevSystem* S, int32_t kind, const ecne_problem_t* sub, SpecialsHost* sp, uint64_t* n_matches,
                    cudaStream_t s, std::string& err, AbstractionStats* stats, const std::function<int()>& ready,
                    const PreparedSub* prepared = nullptr);

// setup.cu: H2D + classification + layout.  Returns an ecne_status.  With `dev` the rows are taken from a system that
// is already resident on the device (p->seg_ptr / col / coef are not read; sizes come from `dev`).
int build_resident(const ecne_problem_t* p, Resident* r, std::string& err, const DevSystem* dev = nullptr);

}  // namespace ecne
