// engine_host.h — host-side declarations shared by the translation units of libecne_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "engine.cuh"

namespace ecne {

// kernels.cu
cudaError_t launch_reset(const Dev& d, int grid, cudaStream_t s);
int p1_grid_size(int device);
int p1_threads();
cudaError_t launch_p1(const Dev& d, int rbuf, unsigned int max_rounds, int grid, cudaStream_t s);
void launch_replay(const Dev& d, int buf, cudaStream_t s);
void launch_p0(const Dev& d, cudaStream_t s);
void launch_p2_groups(const Dev& d, int rbuf, uint32_t n_cand, const unsigned long long* keys,
                      const uint32_t* rows, cudaStream_t s);
void launch_p3(const Dev& d, int rbuf, cudaStream_t s);
void launch_p4(const Dev& d, int rbuf, cudaStream_t s);
void launch_finalize(const Dev& d, int buf, unsigned long long* ubits, unsigned long long* kbits,
                     unsigned long long* counts, cudaStream_t s);
void launch_export(const Dev& d, int buf, fr::u256* lb, fr::u256* ub, uint8_t* nvalues,
                   fr::u256* values, cudaStream_t s);

// A device allocation arena: every buffer of one resident problem, freed together.
struct Arena {
  std::vector<void*> ptrs;
  size_t bytes = 0;
  template <class T>
  cudaError_t alloc(T** out, size_t n) {
    void* p = nullptr;
    size_t sz = (n ? n : 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&p, sz);
    if (e != cudaSuccess) return e;
    ptrs.push_back(p);
    bytes += sz;
    *out = (T*)p;
    return cudaSuccess;
  }
  void release() {
    for (void* p : ptrs) cudaFree(p);
    ptrs.clear();
    bytes = 0;
  }
};

// One uploaded + classified problem (ecne_resident in the ABI).
struct Resident {
  Dev d;
  Arena arena;
  cudaStream_t stream = nullptr;
  // host mirrors needed by the solve loop
  uint64_t n_rows = 0, n_vars = 0, n_targets = 0;
  // finalisation buffers
  unsigned long long *d_ubits = nullptr, *d_kbits = nullptr, *d_counts = nullptr;
  // CUB temp storage and P2 sort buffers
  void* d_cub = nullptr;
  size_t cub_bytes = 0;
  unsigned long long* d_key2 = nullptr;
  uint32_t* d_row2 = nullptr;
  // pinned host staging
  Status* h_status = nullptr;
  unsigned long long* h_counts = nullptr;
  double ms_h2d = 0, ms_classify = 0;
};

// setup.cu: H2D + classification + layout.  Returns an ecne_status.
int build_resident(const ecne_problem_t* p, Resident* r, std::string& err);
int p2_sort(Resident* r, uint32_t n_cand, std::string& err);

}  // namespace ecne
