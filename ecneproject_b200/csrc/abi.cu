// abi.cu — the C ABI of include/ecne_abi.h and the outer fixpoint loop
// (/root/reference/src/R1CSConstraintSolver.jl:706-1556) that drives the kernels.
#include <dlfcn.h>
#include <nccl.h>  // types only: the library is dlopen()ed on first use (see NcclApi)

#include <algorithm>
#include <chrono>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "engine_host.h"

using namespace ecne;

// the second build of the solve kernel (1024 threads x 64 registers per block; csrc/kernels.cu "the second build")
extern "C" cudaError_t ecne_launch_solve_v1024(const void* dev, unsigned int max_rounds, int grid, cudaStream_t s);
extern "C" int ecne_p1_grid_size_v1024(int device);

namespace {

// One device this process drives: its streams, its slab pool, its pinned staging and its exchange buffer.
struct Ctx {
  int device = -1;
  cudaStream_t stream = nullptr, side = nullptr, side2 = nullptr;
  SlabPool pool;
  void* h_status = nullptr;
  void* h_counts = nullptr;
  char* xbuf = nullptr;  // exchange channel of the rank this context plays (see below)
};

struct Global {
  bool inited = false;
  std::string err;
  // devices of this process: one after ecne_init(device) (one process per GPU), n after ecne_init_multi(n)
  // (ONE process, one host thread, n GPUs: context i plays rank i of world n)
  std::vector<Ctx*> ctx;
  bool multi = false;
  // options
  long long max_rounds = 1000000;     // per P1 launch
  long long max_outer = 100000;
  long long grid_blocks = 0;          // testing knob: launch the solve kernel with fewer blocks than SMs (0: one per SM)
  long long p2_hash_bits = 56;         // testing knob: bits of the P2 set hash that are used (fewer => collisions)
  long long sparse_max = -1;          // records per round up to which a round is frontier-driven (-1: rows / 32)
  // open rows of the linear-system sweep below which block 0 runs whole outer rounds alone (0: never; at most CH_LIST_CAP)
  long long chain_open_max = 4096;
  // one process per GPU: 1 = every rank uploads 1/world of the rows and the ranks gather them over NVLink; 0 = every
  // rank uploads the whole problem itself (set before ecne_dist_init; must be the same on every rank)
  long long shard_upload = 1;
  // On several GPUs the dense sweeps of a problem with at least this many rows are split over the ranks; a smaller
  // problem is solved by every rank on all rows without any exchange (a sharded round costs a cross-GPU barrier,
  // ~15-25 us, which a sweep of a few 10^5 rows does not earn back: ecdsa's 694 k rows sweep in ~20 us)
  long long shard_min_rows = 2000000;
  // ... and at least this many rows PER GPU: every rank applies every record of a sharded round to its replica of the
  // wire state, so only the sweep itself shrinks with the world size — measured on S16 (11.1 M rows): 12.3 ms on one
  // GPU, 12.6 sharded over two, 14.2 sharded over eight.  shard_min_rows == 0 forces sharding regardless.
  long long shard_min_rows_per_gpu = 4000000;
  // which build of the solve kernel: 0 = by size (a GPU that sweeps at least `wide_min_rows` rows takes the 1024-thread
  // build), 1 = 512 threads x 128 registers, 2 = 1024 x 64
  long long solve_variant = 0;
  long long wide_min_rows = 1500000;
  // one process per GPU: rank / world of this process and the communicator that bootstraps the peer mappings
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  // exchange channel: one cudaMalloc'ed buffer per rank, visible to every peer (CUDA IPC mappings between
  // processes, plain peer access inside one process)
  //   [0, 4096)            header (engine.cuh XH_*): mailbox u64[64] @0, xcnt u32[4][8] @1024, epoch u32 @2048
  //   4096 + l*xcap*16     record list l (l = 0..2)
  size_t xcap = 0;  // records per list
  char* xpeer[ECNE_MAX_WORLD] = {nullptr};
  int n_local() const { return (int)ctx.size(); }
  int total_world() const { return multi ? n_local() : world; }
} G;

// NCCL is bound at run time, not link time: a host process that also imports PyTorch must end up
// with ONE libnccl.so.2 (torch bundles 2.28, the system has 2.27), and whichever is already mapped
// is the one dlopen() hands back.
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (h) return true;
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
    GetUniqueId = (decltype(GetUniqueId))dlsym(h, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(h, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
    AllGather = (decltype(AllGather))dlsym(h, "ncclAllGather");
    AllReduce = (decltype(AllReduce))dlsym(h, "ncclAllReduce");
    GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && CommDestroy && AllGather && AllReduce;
  }
} NCCL;

}  // namespace
namespace ecne {
SlabPool& slab_pool() {  // the pool of the device that is current
  static SlabPool fallback;
  int dev = -1;
  cudaGetDevice(&dev);
  for (Ctx* c : G.ctx)
    if (c->device == dev) return c->pool;
  return fallback;
}
}  // namespace ecne
namespace {

int fail(int status, const std::string& msg) {
  G.err = msg;
  return status;
}

#define CKA(x)                                                                              \
  do {                                                                                      \
    cudaError_t e_ = (x);                                                                   \
    if (e_ != cudaSuccess)                                                                  \
      return fail(ECNE_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_));            \
  } while (0)

const char* status_text(int st) {
  switch (st) {
    case ECNE_E_DIVZERO: return "DivideError: divexact by zero (R1CSConstraintSolver.jl:919-920 / :1467)";
    case ECNE_E_BOUNDS: return "BoundsError (R1CSConstraintSolver.jl:916 variable_states[-1] / :762)";
    case ECNE_E_NODSU: return "UndefVarError: dsu not defined (R1CSConstraintSolver.jl:762; secp_solve=false)";
    case ECNE_E_UNSUPPORTED: return "a linear-system group with k > ECNE_P2_KBIG (16) unknowns triggered";
    case ECNE_E_NOCONVERGE: return "round guard hit: the propagation did not reach a fixpoint";
    case ECNE_E_INTERNAL: return "internal error (record overflow or set-hash collision)";
    default: return "error";
  }
}

}  // namespace

// One uploaded problem: a classified copy per device this process drives (rs[0] reports; all copies hold the same
// wire state after a solve).
struct ecne_resident {
  std::vector<Resident> rs;
};

namespace {
int setup_exchange(ecne_resident& H, size_t cap);
int open_ctx(int device, Ctx** out);
void close_all();
}  // namespace

extern "C" int ecne_version(void) { return ECNE_ABI_VERSION; }

extern "C" const char* ecne_last_error(void) { return G.err.c_str(); }

// Layout self-check for foreign-language mirrors of the three structs (ctypes, Julia `struct`): for each of
// ecne_problem_t, ecne_result_t, ecne_report_t the words {sizeof, number of fields, offsetof(field) ...} in
// declaration order.  Returns the number of words the table has; writes at most `cap` of them.
extern "C" int ecne_abi_layout(uint32_t* out, uint32_t cap) {
  std::vector<uint32_t> t;
#define ECNE_F(T, f) t.push_back((uint32_t)offsetof(T, f));
#define ECNE_S(T, n) t.push_back((uint32_t)sizeof(T)); t.push_back(n);
  ECNE_S(ecne_problem_t, 22)
  ECNE_F(ecne_problem_t, n_rows) ECNE_F(ecne_problem_t, n_vars) ECNE_F(ecne_problem_t, seg_ptr)
  ECNE_F(ecne_problem_t, col) ECNE_F(ecne_problem_t, coef) ECNE_F(ecne_problem_t, known)
  ECNE_F(ecne_problem_t, n_known) ECNE_F(ecne_problem_t, targets) ECNE_F(ecne_problem_t, n_targets)
  ECNE_F(ecne_problem_t, n_specials) ECNE_F(ecne_problem_t, sp_kind) ECNE_F(ecne_problem_t, sp_in_ptr)
  ECNE_F(ecne_problem_t, sp_in) ECNE_F(ecne_problem_t, sp_out_ptr) ECNE_F(ecne_problem_t, sp_out)
  ECNE_F(ecne_problem_t, secp_solve) ECNE_F(ecne_problem_t, debug) ECNE_F(ecne_problem_t, coef_class)
  ECNE_F(ecne_problem_t, coef_other) ECNE_F(ecne_problem_t, coef_other_term) ECNE_F(ecne_problem_t, n_coef_other)
  ECNE_F(ecne_problem_t, seg_ptr32)
  ECNE_S(ecne_result_t, 31)
  ECNE_F(ecne_result_t, verdict) ECNE_F(ecne_result_t, status) ECNE_F(ecne_result_t, unique_bits)
  ECNE_F(ecne_result_t, known_bits) ECNE_F(ecne_result_t, lb) ECNE_F(ecne_result_t, ub)
  ECNE_F(ecne_result_t, nvalues) ECNE_F(ecne_result_t, values) ECNE_F(ecne_result_t, abz)
  ECNE_F(ecne_result_t, n_unique_nontrivial) ECNE_F(ecne_result_t, n_nontrivial)
  ECNE_F(ecne_result_t, n_targets_unique) ECNE_F(ecne_result_t, n_unique) ECNE_F(ecne_result_t, outer_rounds)
  ECNE_F(ecne_result_t, inner_rounds) ECNE_F(ecne_result_t, constraint_evals) ECNE_F(ecne_result_t, sweep_launches)
  ECNE_F(ecne_result_t, ms_h2d) ECNE_F(ecne_result_t, ms_classify) ECNE_F(ecne_result_t, ms_solve)
  ECNE_F(ecne_result_t, ms_d2h) ECNE_F(ecne_result_t, ms_exchange) ECNE_F(ecne_result_t, ms_total)
  ECNE_F(ecne_result_t, ms_sweep) ECNE_F(ecne_result_t, rule_evals) ECNE_F(ecne_result_t, dense_rounds)
  ECNE_F(ecne_result_t, dense_evals) ECNE_F(ecne_result_t, dense_cycles) ECNE_F(ecne_result_t, ms_device)
  ECNE_F(ecne_result_t, gpus_used) ECNE_F(ecne_result_t, sharded)
  ECNE_S(ecne_report_t, 10)
  ECNE_F(ecne_report_t, bad_row_bits) ECNE_F(ecne_report_t, cap_wires) ECNE_F(ecne_report_t, wire)
  ECNE_F(ecne_report_t, flags) ECNE_F(ecne_report_t, lb) ECNE_F(ecne_report_t, ub) ECNE_F(ecne_report_t, nvalues)
  ECNE_F(ecne_report_t, values) ECNE_F(ecne_report_t, n_bad_rows) ECNE_F(ecne_report_t, n_wires)
#undef ECNE_F
#undef ECNE_S
  if (out)
    for (size_t i = 0; i < t.size() && i < cap; ++i) out[i] = t[i];
  return (int)t.size();
}

namespace {
int open_ctx(int device, Ctx** out) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(ECNE_E_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") +
                                 cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(ECNE_E_BADARG, "device index out of range");
  CKA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CKA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(ECNE_E_CUDA, std::string("libecne_b200 is built for sm_100a only; device is ") + prop.name);
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
  if (!coop) return fail(ECNE_E_CUDA, "device lacks cooperative launch");
  Ctx* c = new Ctx();
  c->device = device;
  G.ctx.push_back(c);  // (registered first: a failure below is cleaned up by close_all)
  CKA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CKA(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
  CKA(cudaStreamCreateWithFlags(&c->side2, cudaStreamNonBlocking));
  CKA(cudaMallocHost(&c->h_status, sizeof(Status)));
  CKA(cudaMallocHost(&c->h_counts, 4 * sizeof(unsigned long long)));
  *out = c;
  return ECNE_OK;
}
void close_all() {
  if (G.comm) {
    NCCL.CommDestroy(G.comm);
    G.comm = nullptr;
  }
  for (Ctx* c : G.ctx) {
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->side2) cudaStreamDestroy(c->side2);
  }
  if (!G.multi)
    for (int h = 0; h < ECNE_MAX_WORLD; ++h)
      if (G.xpeer[h] && (G.ctx.empty() || G.xpeer[h] != G.ctx[0]->xbuf)) cudaIpcCloseMemHandle(G.xpeer[h]);
  memset(G.xpeer, 0, sizeof(G.xpeer));
  for (Ctx* c : G.ctx) {
    cudaSetDevice(c->device);
    if (c->xbuf) cudaFree(c->xbuf);
    c->pool.destroy();
    if (c->h_status) cudaFreeHost(c->h_status);
    if (c->h_counts) cudaFreeHost(c->h_counts);
    delete c;
  }
  G.ctx.clear();
  G.xcap = 0;
  G.inited = false;
  G.multi = false;
  G.rank = 0;
  G.world = 1;
}
}  // namespace

extern "C" int ecne_init(int device) {
  if (G.inited && !G.multi && G.n_local() == 1 && G.ctx[0]->device == device) return ECNE_OK;
  close_all();
  Ctx* c = nullptr;
  int st = open_ctx(device, &c);
  if (st != ECNE_OK) {
    close_all();
    return st;
  }
  G.inited = true;
  return ECNE_OK;
}

// SURVEY.md §8b "Threading": ONE process, one host thread, the GPUs 0 .. n_gpus-1 of the box.  The devices see each
// other's exchange buffers through plain peer access (no IPC handles, no NCCL, no torch), context i plays rank i.
extern "C" int ecne_init_multi(int n_gpus) {
  if (n_gpus < 1 || n_gpus > ECNE_MAX_WORLD) return fail(ECNE_E_BADARG, "n_gpus outside 1..8");
  if (G.inited && G.multi && G.n_local() == n_gpus) return ECNE_OK;
  close_all();
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(ECNE_E_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
  if (n_gpus > n)
    return fail(ECNE_E_BADARG, "ecne_init_multi(" + std::to_string(n_gpus) + "): the box has " + std::to_string(n) + " GPU(s)");
  for (int i = 0; i < n_gpus; ++i) {
    Ctx* c = nullptr;
    int st = open_ctx(i, &c);
    if (st != ECNE_OK) {
      close_all();
      return st;
    }
  }
  for (int i = 0; i < n_gpus; ++i)
    for (int j = 0; j < n_gpus; ++j) {
      if (i == j) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, i, j);
      if (!can) {
        close_all();
        return fail(ECNE_E_CUDA, "GPUs " + std::to_string(i) + " and " + std::to_string(j) + " have no peer access");
      }
      // (a process may call ecne_init_multi again after ecne_shutdown: the mapping of a pair is enabled once for the
      // life of the process, and asking again would only leave an API error for the tools to report)
      static bool peer_on[ECNE_MAX_WORLD][ECNE_MAX_WORLD];
      if (peer_on[i][j]) continue;
      cudaSetDevice(i);
      cudaError_t pe = cudaDeviceEnablePeerAccess(j, 0);
      if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) {
        close_all();
        return fail(ECNE_E_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe));
      }
      cudaGetLastError();
      peer_on[i][j] = true;
    }
  cudaSetDevice(0);
  G.multi = n_gpus > 1;
  G.inited = true;
  return ECNE_OK;
}

extern "C" void ecne_shutdown(void) { close_all(); }

extern "C" int ecne_set_option(const char* key, int64_t value) {
  if (!key) return fail(ECNE_E_BADARG, "null key");
  std::string k(key);
  if (k == "max_rounds")
    G.max_rounds = value;
  else if (k == "max_outer")
    G.max_outer = value;
  else if (k == "sparse_max")
    G.sparse_max = value;
  else if (k == "grid_blocks")
    G.grid_blocks = value;
  else if (k == "shard_upload")
    G.shard_upload = value ? 1 : 0;
  else if (k == "chain_open_max")
    G.chain_open_max = value < 0 ? 0 : (value > (long long)CH_LIST_CAP ? (long long)CH_LIST_CAP : value);
  else if (k == "solve_variant")
    G.solve_variant = value < 0 || value > 2 ? 0 : value;
  else if (k == "wide_min_rows")
    G.wide_min_rows = value < 0 ? 0 : value;
  else if (k == "shard_min_rows")
    G.shard_min_rows = value < 0 ? 0 : value;
  else if (k == "shard_min_rows_per_gpu")
    G.shard_min_rows_per_gpu = value < 0 ? 0 : value;
  else if (k == "p2_hash_bits")
    G.p2_hash_bits = value < 0 ? 0 : (value > 56 ? 56 : value);
  else
    return fail(ECNE_E_BADARG, "unknown option " + k);
  return ECNE_OK;
}

namespace {
// first row of each rank's range: ranges of equal stored-term weight (what ecne_shard_rows does on host arrays)
__global__ void k_shard_cuts(const unsigned long long* seg, unsigned long long N, int world, unsigned long long* cuts) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > world) return;
  if (r == 0) {
    cuts[0] = 0;
    return;
  }
  if (r == world) {
    cuts[r] = N;
    return;
  }
  const unsigned long long total = seg[3 * N];
  const unsigned long long want = (unsigned long long)(((unsigned __int128)total * (unsigned)r) / (unsigned)world);
  unsigned long long a = 0, b = N;  // smallest row i with seg[3 i] >= want
  while (a < b) {
    const unsigned long long m = (a + b) / 2;
    if (seg[3 * m] >= want) b = m; else a = m + 1;
  }
  cuts[r] = a;
}

// ecne_upload: `dev0` != nullptr: the rows are already resident on the first device (abstraction on the device);
// problem->seg_ptr / col / coef are then not read.
int upload_impl(const ecne_problem_t* problem, const DevSystem* dev0, ecne_resident_t** out) {
  if (!out) return fail(ECNE_E_BADARG, "null out pointer");
  *out = nullptr;
  if (!G.inited) {
    int st = ecne_init(0);
    if (st) return st;
  }
  const int nl = G.n_local(), world = G.total_world();
  if (world > 1 && !G.multi && !G.comm) return fail(ECNE_E_NCCL, "ecne_dist_init has not been called");
  // one process, several GPUs: the rows cross PCIe ONCE (to the first device); the others pull them over NVLink
  DevSystem up0;
  struct Up0Guard {
    DevSystem& s;
    ~Up0Guard() {
      if (s.arena.pool && !s.arena.slabs.empty()) {
        cudaSetDevice(G.ctx[0]->device);
        cudaStreamSynchronize(G.ctx[0]->stream);
        s.arena.release();
      }
    }
  } up0_guard{up0};
  if (nl > 1 && !dev0) {
    cudaSetDevice(G.ctx[0]->device);
    up0.arena.pool = &G.ctx[0]->pool;
    std::string e;
    const int st = dev_system_upload(problem, &up0, G.ctx[0]->stream, e);
    if (st != ECNE_OK) return fail(st, e);
    CKA(cudaStreamSynchronize(G.ctx[0]->stream));
    dev0 = &up0;
  }
  ecne_resident* h = new ecne_resident();
  h->rs.resize(nl);
  std::vector<int> sts(nl, ECNE_OK);
  std::vector<std::string> errs(nl);
  auto build_one = [&](int i) {
    Ctx* c = G.ctx[i];
    cudaSetDevice(c->device);
    Resident& R = h->rs[i];
    R.stream = c->stream;
    R.side = c->side;
    R.side2 = c->side2;
    R.device = c->device;
    R.h_status = (Status*)c->h_status;
    R.h_counts = (unsigned long long*)c->h_counts;
    R.arena.pool = &c->pool;
    if (!dev0) {
      sts[i] = build_resident(problem, &R, errs[i]);
    } else if (i == 0) {
      sts[i] = build_resident(problem, &R, errs[i], dev0);
    } else {
      // the other devices of this process pull the rows from the first one over NVLink (peer copy)
      DevSystem ci;
      ci.N = dev0->N;
      ci.V = dev0->V;
      ci.nnz = dev0->nnz;
      ci.arena.pool = &c->pool;
      cudaError_t e = ci.arena.alloc(&ci.seg, 3 * ci.N + 2);
      if (e == cudaSuccess) e = ci.arena.alloc(&ci.col, ci.nnz + 1);
      if (e == cudaSuccess) e = ci.arena.alloc(&ci.coef, ci.nnz + 1);
      const int d0 = G.ctx[0]->device;
      if (e == cudaSuccess) e = cudaMemcpyPeerAsync(ci.seg, c->device, dev0->seg, d0, (3 * ci.N + 1) * 8, c->stream);
      if (e == cudaSuccess && ci.nnz) e = cudaMemcpyPeerAsync(ci.col, c->device, dev0->col, d0, ci.nnz * 4, c->stream);
      if (e == cudaSuccess && ci.nnz) e = cudaMemcpyPeerAsync(ci.coef, c->device, dev0->coef, d0, ci.nnz * 32, c->stream);
      if (e != cudaSuccess) {
        errs[i] = std::string("peer copy of the reduced system: ") + cudaGetErrorString(e);
        sts[i] = ECNE_E_CUDA;
      } else {
        sts[i] = build_resident(problem, &R, errs[i], &ci);
      }
      cudaStreamSynchronize(c->stream);
      if (ci.arena.pool) ci.arena.release();
    }
  };
  if (nl == 1) {
    build_one(0);
  } else {
    // every device gets and classifies the whole problem (the wire state and the phases are replicated); the
    // copies run concurrently, one host thread per device for the duration of the upload only
    std::vector<std::thread> th;
    for (int i = 0; i < nl; ++i) th.emplace_back(build_one, i);
    for (auto& t : th) t.join();
    cudaSetDevice(G.ctx[0]->device);
  }
  for (int i = 0; i < nl; ++i)
    if (sts[i] != ECNE_OK) {
      const int st = sts[i];
      const std::string e = errs[i];
      ecne_free_resident(h);
      return fail(st, e);
    }
  if (world > 1) {
    const uint64_t n_rows = dev0 ? dev0->N : problem->n_rows;
    const bool shard = G.shard_min_rows == 0 || ((long long)n_rows >= G.shard_min_rows &&
                                                 (long long)(n_rows / (uint64_t)world) >= G.shard_min_rows_per_gpu);
    std::vector<unsigned long long> cuts;
    if (shard && dev0) {  // the row offsets live on the device: the cut points are found there
      cuts.resize(world + 1);
      Arena t;
      unsigned long long* d_cuts = nullptr;
      cudaSetDevice(G.ctx[0]->device);
      CKA(t.alloc(&d_cuts, world + 1));
      k_shard_cuts<<<1, 32, 0, G.ctx[0]->stream>>>(dev0->seg, dev0->N, world, d_cuts);
      CKA(cudaMemcpyAsync(cuts.data(), d_cuts, (world + 1) * 8, cudaMemcpyDeviceToHost, G.ctx[0]->stream));
      CKA(cudaStreamSynchronize(G.ctx[0]->stream));
      t.release();
    }
    for (int i = 0; i < nl; ++i) {
      Dev& di = h->rs[i].d;
      di.world = world;
      di.rank = G.multi ? i : G.rank;
      di.shard = shard ? 1 : 0;
      if (shard) {
        uint64_t lo = 0, hi = 0;
        if (dev0) {
          lo = cuts[di.rank];
          hi = cuts[di.rank + 1];
        } else {
          ecne_shard_rows(problem, di.rank, world, &lo, &hi);
        }
        di.row_lo = (uint32_t)lo;
        di.row_hi = (uint32_t)hi;
      }
    }
    if (shard) {
      int st = setup_exchange(*h, h->rs[0].d.rec_cap);
      if (st != ECNE_OK) {
        ecne_free_resident(h);
        return st;
      }
    }
  }
  *out = h;
  return ECNE_OK;
}
}  // namespace

extern "C" int ecne_upload(const ecne_problem_t* problem, ecne_resident_t** out) { return upload_impl(problem, nullptr, out); }

// ---- abstraction() on the device (SURVEY.md §8f-1, R1CSConstraintSolver.jl:237-395) --------------------------------
struct ecne_abstracted {
  DevSystem sys;
  SpecialsHost sp;
  std::vector<uint32_t> known, targets;
  uint64_t n_vars = 0;
  AbstractionStats stats;
  int device = 0;
  cudaStream_t stream = nullptr;
  double ms_h2d = 0;
  // the upload of the unreduced system runs behind the caller's back until the handle is used for the first time:
  // the host prepares the first trusted circuit (sorted coefficient lists, signature classes) meanwhile
  std::thread uploader;
  int upload_status = ECNE_OK;
  std::string upload_err;
  int wait_upload() {
    if (uploader.joinable()) uploader.join();
    return upload_status;
  }
};

extern "C" int ecne_abstract_begin(const ecne_problem_t* main, ecne_abstracted_t** out) {
  if (!main || !out) return fail(ECNE_E_BADARG, "null argument");
  *out = nullptr;
  if (!G.inited) {
    int st = ecne_init(0);
    if (st) return st;
  }
  Ctx* c = G.ctx[0];
  CKA(cudaSetDevice(c->device));
  ecne_abstracted* a = new ecne_abstracted();
  a->device = c->device;
  a->stream = c->stream;
  a->n_vars = main->n_vars;
  a->known.assign(main->known, main->known + main->n_known);
  a->targets.assign(main->targets, main->targets + main->n_targets);
  a->sys.arena.pool = &c->pool;
  // (the caller's arrays must stay alive until the first ecne_abstract_apply / _sizes / _export / _upload returns)
  a->uploader = std::thread([a, main, c]() {
    cudaSetDevice(c->device);
    auto t0 = std::chrono::steady_clock::now();
    int st = dev_system_upload(main, &a->sys, c->stream, a->upload_err);
    if (st == ECNE_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) {
      st = ECNE_E_CUDA;
      a->upload_err = "upload of the unreduced system failed";
    }
    a->ms_h2d = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    a->upload_status = st;
  });
  *out = a;
  return ECNE_OK;
}
struct ecne_prepared {
  PreparedSub* p = nullptr;
};
// (touches neither CUDA nor the library's global state — not even the last-error string: callable from another thread)
extern "C" int ecne_abstract_prepare(const ecne_problem_t* sub, ecne_prepared_t** out) {
  if (!out) return ECNE_E_BADARG;
  *out = nullptr;
  PreparedSub* p = abstraction_prepare(sub);
  if (!p) return ECNE_E_BADARG;
  ecne_prepared_t* h = new (std::nothrow) ecne_prepared_t();
  if (!h) {
    abstraction_prepared_free(p);
    return ECNE_E_BOUNDS;
  }
  h->p = p;
  *out = h;
  return ECNE_OK;
}
extern "C" void ecne_abstract_prepared_free(ecne_prepared_t* p) {
  if (!p) return;
  abstraction_prepared_free(p->p);
  delete p;
}
static int abstract_apply_impl(ecne_abstracted_t* a, int32_t kind, const ecne_problem_t* sub, const ecne_prepared_t* prepared,
                               uint64_t* n_matches);
extern "C" int ecne_abstract_apply(ecne_abstracted_t* a, int32_t kind, const ecne_problem_t* sub, uint64_t* n_matches) {
  return abstract_apply_impl(a, kind, sub, nullptr, n_matches);
}
extern "C" int ecne_abstract_apply_prepared(ecne_abstracted_t* a, int32_t kind, const ecne_problem_t* sub,
                                            const ecne_prepared_t* prepared, uint64_t* n_matches) {
  if (!prepared || !prepared->p) return fail(ECNE_E_BADARG, "null prepared trusted circuit");
  return abstract_apply_impl(a, kind, sub, prepared, n_matches);
}
static int abstract_apply_impl(ecne_abstracted_t* a, int32_t kind, const ecne_problem_t* sub, const ecne_prepared_t* prepared,
                               uint64_t* n_matches) {
  if (!a || !sub) return fail(ECNE_E_BADARG, "null argument");
  CKA(cudaSetDevice(a->device));
  std::string err;
  int st = dev_abstraction(&a->sys, kind, sub, &a->sp, n_matches, a->stream, err, &a->stats,
                           [a]() { return a->wait_upload(); }, prepared ? prepared->p : nullptr);
  if (st != ECNE_OK && a->upload_status != ECNE_OK) return fail(a->upload_status, a->upload_err);
  if (st != ECNE_OK) return fail(st, err);
  if (getenv("ECNE_HOST_PROF"))
    fprintf(stderr, "[ecne dev] upload of the unreduced system (uploader thread) %.3f ms\n", a->ms_h2d);
  if (getenv("ECNE_HOST_PROF"))
    fprintf(stderr, "[ecne dev] abstraction: trusted circuit prepared on the host %.3f ms | hashes %.3f ms | candidates %.3f ms (%llu) | "
                    "verification %.3f ms (%llu matches) | compaction %.3f ms (cumulative over the calls of this handle)\n",
            a->stats.ms_prepare, a->stats.ms_hash, a->stats.ms_candidates,
            (unsigned long long)a->stats.n_candidates, a->stats.ms_verify, (unsigned long long)a->stats.n_matches, a->stats.ms_compact);
  return ECNE_OK;
}
extern "C" int ecne_abstract_sizes(const ecne_abstracted_t* a, uint64_t sizes[5]) {
  if (!a || !sizes) return fail(ECNE_E_BADARG, "null argument");
  if (const_cast<ecne_abstracted_t*>(a)->wait_upload() != ECNE_OK) return fail(a->upload_status, a->upload_err);
  sizes[0] = a->sys.N;
  sizes[1] = a->sys.nnz;
  sizes[2] = a->sp.kind.size();
  sizes[3] = a->sp.in.size();
  sizes[4] = a->sp.out.size();
  return ECNE_OK;
}
extern "C" int ecne_abstract_export(ecne_abstracted_t* a, uint64_t* seg_ptr, uint32_t* col, uint64_t* coef, int32_t* sp_kind,
                                    uint64_t* sp_in_ptr, uint32_t* sp_in, uint64_t* sp_out_ptr, uint32_t* sp_out) {
  if (!a) return fail(ECNE_E_BADARG, "null argument");
  if (a->wait_upload() != ECNE_OK) return fail(a->upload_status, a->upload_err);
  CKA(cudaSetDevice(a->device));
  if (seg_ptr) CKA(cudaMemcpyAsync(seg_ptr, a->sys.seg, (3 * a->sys.N + 1) * 8, cudaMemcpyDeviceToHost, a->stream));
  if (col && a->sys.nnz) CKA(cudaMemcpyAsync(col, a->sys.col, a->sys.nnz * 4, cudaMemcpyDeviceToHost, a->stream));
  if (coef && a->sys.nnz) CKA(cudaMemcpyAsync(coef, a->sys.coef, a->sys.nnz * 32, cudaMemcpyDeviceToHost, a->stream));
  CKA(cudaStreamSynchronize(a->stream));
  const size_t ns = a->sp.kind.size();
  if (sp_kind && ns) memcpy(sp_kind, a->sp.kind.data(), ns * sizeof(int32_t));
  if (sp_in_ptr) memcpy(sp_in_ptr, a->sp.in_ptr.data(), (ns + 1) * 8);
  if (sp_out_ptr) memcpy(sp_out_ptr, a->sp.out_ptr.data(), (ns + 1) * 8);
  if (sp_in && !a->sp.in.empty()) memcpy(sp_in, a->sp.in.data(), a->sp.in.size() * 4);
  if (sp_out && !a->sp.out.empty()) memcpy(sp_out, a->sp.out.data(), a->sp.out.size() * 4);
  return ECNE_OK;
}
extern "C" int ecne_abstract_upload(ecne_abstracted_t* a, int32_t secp_solve, ecne_resident_t** out) {
  if (!a || !out) return fail(ECNE_E_BADARG, "null argument");
  if (a->wait_upload() != ECNE_OK) return fail(a->upload_status, a->upload_err);
  ecne_problem_t p;
  memset(&p, 0, sizeof p);
  p.n_rows = a->sys.N;
  p.n_vars = a->n_vars;
  p.known = a->known.data();
  p.n_known = a->known.size();
  p.targets = a->targets.data();
  p.n_targets = a->targets.size();
  p.n_specials = a->sp.kind.size();
  p.sp_kind = a->sp.kind.data();
  p.sp_in_ptr = a->sp.in_ptr.data();
  p.sp_in = a->sp.in.data();
  p.sp_out_ptr = a->sp.out_ptr.data();
  p.sp_out = a->sp.out.data();
  p.secp_solve = secp_solve;
  int st = upload_impl(&p, &a->sys, out);
  if (st == ECNE_OK)
    for (auto& R : (*out)->rs) R.ms_h2d = a->ms_h2d;  // what crossed PCIe for this problem: the unreduced system, once
  return st;
}
extern "C" void ecne_abstract_free(ecne_abstracted_t* a) {
  if (!a) return;
  a->wait_upload();
  cudaSetDevice(a->device);
  if (a->stream) cudaStreamSynchronize(a->stream);
  if (a->sys.arena.pool) a->sys.arena.release();
  delete a;
}

extern "C" void ecne_free_resident(ecne_resident_t* h) {
  if (!h) return;
  for (auto& R : h->rs) {
    cudaSetDevice(R.device);
    if (R.stream) cudaStreamSynchronize(R.stream);
    if (R.arena.pool) R.arena.release();
  }
  if (!G.ctx.empty()) cudaSetDevice(G.ctx[0]->device);
  delete h;
}

extern "C" int ecne_solve_resident(ecne_resident_t* h, ecne_result_t* res) {
  if (!h || !res || !res->unique_bits || !res->known_bits) return fail(ECNE_E_BADARG, "null argument");
  const int nl = (int)h->rs.size();
  Resident& R = h->rs[0];  // the copy that reports (every copy holds the same final state)
  Dev& d = R.d;
  R.have_state = false;
  cudaStream_t s = R.stream;
  struct Events {  // destroyed on every return path
    std::vector<cudaEvent_t> e;
    std::vector<int> dev;
    cudaEvent_t make(int device) {
      cudaSetDevice(device);
      cudaEvent_t x;
      cudaEventCreate(&x);
      e.push_back(x);
      dev.push_back(device);
      return x;
    }
    ~Events() {
      for (size_t i = 0; i < e.size(); ++i) {
        cudaSetDevice(dev[i]);
        cudaEventDestroy(e[i]);
      }
      if (!dev.empty()) cudaSetDevice(dev[0]);
    }
  } evs;
  cudaEvent_t e0 = evs.make(R.device), e1 = evs.make(R.device), e2 = evs.make(R.device);
  std::vector<cudaEvent_t> s0(nl), s1(nl);
  for (int i = 0; i < nl; ++i) {
    s0[i] = evs.make(h->rs[i].device);
    s1[i] = evs.make(h->rs[i].device);
  }
  cudaSetDevice(R.device);
  cudaEventRecord(e0, s);
  // (sharded runs need no start-of-solve rendezvous: the exchange epochs grow monotonically over the life of the
  // process and the mailboxes are never cleared, engine.cuh "cross-GPU exchange header")
  // The whole fixpoint (:706-1556) is ONE persistent cooperative launch per device; the host only reads the
  // status blocks back when they have finished.  With several local devices all launches are issued before any
  // is waited for: the kernels meet each other at the sharded rounds.
  for (int i = 0; i < nl; ++i) {
    Resident& Ri = h->rs[i];
    Dev& di = Ri.d;
    CKA(cudaSetDevice(Ri.device));
    const bool wide = G.solve_variant == 2 ||
                      (G.solve_variant == 0 && (long long)(di.row_hi - di.row_lo) >= G.wide_min_rows);
    int grid = wide ? ecne_p1_grid_size_v1024(Ri.device) : p1_grid_size(Ri.device);
    if (G.grid_blocks > 0 && G.grid_blocks < grid) grid = (int)G.grid_blocks;
    CKA(launch_reset(di, grid, Ri.stream));
    if (Ri.table_dirty) {
      CKA(launch_clear_p2_table(di, Ri.stream));
      Ri.table_dirty = false;
    }
    di.max_outer = (uint32_t)std::min<long long>(G.max_outer, 0x7fffffffLL);
    di.p2_hash_mask = G.p2_hash_bits >= 56 ? 0x00ffffffffffffffULL : ((1ULL << G.p2_hash_bits) - 1ULL);
    {
      // from the whole problem, not the shard: every rank must take the same dense / sparse decision
      const long long sm = G.sparse_max >= 0 ? G.sparse_max : std::max<long long>(4096, (long long)di.N / 32);
      di.sparse_max = (uint32_t)std::min<long long>(sm, 0x7fffffffLL);
    }
    di.chain_open_max = G.chain_open_max;
    cudaEventRecord(s0[i], Ri.stream);
    const unsigned int mr = (unsigned int)std::min<long long>(G.max_rounds, 0x7fffffffLL);
    CKA(wide ? ecne_launch_solve_v1024(&di, mr, grid, Ri.stream) : launch_solve(di, mr, grid, Ri.stream));
    cudaEventRecord(s1[i], Ri.stream);
    CKA(cudaMemcpyAsync(Ri.h_status, di.st, sizeof(Status), cudaMemcpyDeviceToHost, Ri.stream));
  }
  for (int i = 0; i < nl; ++i) {
    CKA(cudaSetDevice(h->rs[i].device));
    CKA(cudaStreamSynchronize(h->rs[i].stream));
  }
  CKA(cudaSetDevice(R.device));
  int status = ECNE_OK;
  std::string err;
  float ms_sweep = 0;
  unsigned long long evals_total = 0, rule_evals_total = 0, dense_evals_total = 0, dense_cycles_max = 0;
  for (int i = 0; i < nl; ++i) {
    const Status& S = *h->rs[i].h_status;
    float ms = 0;
    cudaSetDevice(h->rs[i].device);
    cudaEventElapsedTime(&ms, s0[i], s1[i]);
    ms_sweep = std::max(ms_sweep, ms);  // the kernels run concurrently: the slowest one is the solve
    dense_cycles_max = std::max<unsigned long long>(dense_cycles_max, S.dense_cycles);
    if (status == ECNE_OK) {
      if (S.err)
        status = -(int)S.err;
      else if (S.rec_overflow)
        status = ECNE_E_INTERNAL;
    }
    evals_total += S.evals;  // sharded sweeps: every rank its own rows; replicated work is counted by rank 0 only
    rule_evals_total += S.rule_evals;
    dense_evals_total += S.dense_evals;
  }
  cudaSetDevice(R.device);
  const uint64_t outer = R.h_status->outer;
  if (getenv("ECNE_DEBUG_PROF")) {
    const Status& S = *R.h_status;
    fprintf(stderr,
            "[ecne prof] block-0 cycles: dense %llu (%u rounds) sparse %llu (%llu rounds) p2scan %llu p2resolve %llu "
            "p3claim+replay %llu p3commit+p4 %llu p0+replay %llu solo %llu | outer %u | p2 candidates total %u max %u | "
            "dense evals %llu evals %llu\n",
            S.prof[0], S.dense_rounds, S.prof[1], S.rounds - S.dense_rounds, S.prof[2], S.prof[3], S.prof[4], S.prof[5],
            S.prof[6], S.prof[7], S.outer, S.n_cand_total, S.n_cand_max, S.dense_evals, S.evals);
    if (atoi(getenv("ECNE_DEBUG_PROF")) > 2) {
      std::vector<unsigned long long> pr(28000 + 40 * 148 * 4 + 160);
      cudaMemcpy(pr.data(), d.prof, pr.size() * 8, cudaMemcpyDeviceToHost);
      {
        const unsigned long long* q = pr.data() + 28000 + 40 * 148 * 4;
        fprintf(stderr, "[long row] slowest evaluation (maxima, cycles from entry): total %llu | gather %llu | cases 1-4 %llu | case 5 %llu | "
                        "case 6 scan %llu | firing done %llu\n", q[0], q[1], q[2], q[3], q[4], q[5]);
      }
      {
        const unsigned long long* q = pr.data() + 28000 + 40 * 148 * 4 + 128;
        fprintf(stderr, "[warp solo stages] slowest lane per batch, summed: row record + latch %llu | state gather %llu | inline evaluation %llu | "
                        "generic evaluator %llu (%llu rows) | emits of the batch %llu | round start -> first batch %llu | last emit -> round end %llu\n", q[8], q[9], q[10], q[11], q[12], q[13], q[14], q[15]);
        fprintf(stderr, "[warp solo prologue] records from shared memory %llu | replay issued %llu | syncwarp %llu || [epilogue] long-row ballot %llu | "
                        "replay results consumed %llu | syncwarp %llu\n", q[16], q[17], q[18], q[19], q[20], q[21]);
        fprintf(stderr, "[warp solo] %llu stretches, %llu rounds, %llu cycles | %llu pair batches, %llu (record, row) pairs | %llu long rows "
                        "evaluated in %llu cycles (cumulative over the solves of this handle)\n", q[6], q[5], q[4], q[2], q[3], q[1], q[0]);
      }
      fprintf(stderr, "[phases] n_p3 %u n_p4 %u | open rows seen by the P2 scan / P4 rows with a non-unique vk, per outer round:", d.n_p3, d.n_p4);
      for (int o = 1; o < 40; ++o)
        if (pr[28000 + 40 * 148 * 4 + 48 + o] | pr[28000 + 40 * 148 * 4 + 88 + o])
          fprintf(stderr, " %d:%llu/%llu", o, pr[28000 + 40 * 148 * 4 + 48 + o], pr[28000 + 40 * 148 * 4 + 88 + o]);
      fprintf(stderr, "\n");
      for (int r = 12; r < 24; ++r) {
        unsigned long long mx[3] = {0, 0, 0}, sum[3] = {0, 0, 0}, g = 0;
        for (int b = 0; b < 148; ++b)
          for (int k = 0; k < 3; ++k) {
            unsigned long long v = pr[28000 + ((size_t)r * 148 + b) * 4 + k];
            mx[k] = v > mx[k] ? v : mx[k];
            sum[k] += v;
            g = pr[28000 + ((size_t)r * 148 + b) * 4 + 3];
          }
        if (g) fprintf(stderr, "[p2scan] outer %llu: short rows mean %llu max %llu | long rows mean %llu max %llu | pre mean %llu max %llu\n", g,
                       sum[0] / 148, mx[0], sum[1] / 148, mx[1], sum[2] / 148, mx[2]);
      }
      for (int r = 0; r < 12; ++r) {
        const unsigned long long* q = pr.data() + 27000 + 8 * r;
        if (q[0] | q[4])
          fprintf(stderr, "[dense %d] block 0: work %llu | round barrier%s %llu | pull %llu | barrier+ack %llu | records %llu\n", r, q[4],
                  d.world > 1 ? " + exchange" : "", q[0], q[1], q[2], q[3]);
      }
      for (int r = 0; r < 12; ++r) {
        unsigned long long mx[3] = {0, 0, 0}, sum[3] = {0, 0, 0}, g = 0, inl = 0;
        for (int b = 0; b < 148; ++b) {
          for (int k = 0; k < 3; ++k) {
            unsigned long long v = pr[28000 + ((size_t)r * 148 + b) * 4 + k];
            mx[k] = v > mx[k] ? v : mx[k];
            sum[k] += v;
          }
          g = pr[28000 + ((size_t)r * 148 + b) * 4 + 3] & 0xfffffULL;
          inl += pr[28000 + ((size_t)r * 148 + b) * 4 + 3] >> 20;
        }
        if (g) fprintf(stderr, "[dense] round %llu: long rows mean %llu max %llu | replay mean %llu max %llu | sweep mean %llu max %llu "
                       "(inline loop of warp 0: mean %llu; rows sent to the generic evaluator: %llu)\n", g,
                       sum[0] / 148, mx[0], sum[1] / 148, mx[1], sum[2] / 148, mx[2], inl / 148,
                       pr[28000 + 40 * 148 * 4 + 8 + r]);
      }
    }
    if (atoi(getenv("ECNE_DEBUG_PROF")) > 1) {
      std::vector<unsigned long long> pr(28000 + 12 * 148 * 4);
      cudaMemcpy(pr.data(), d.prof, pr.size() * 8, cudaMemcpyDeviceToHost);
      for (unsigned long long g = 1; g <= S.rounds && g < 4000; ++g)
        fprintf(stderr, "[round] %llu cycles %llu records %llu dense %llu outer %llu | thread0: rec+head %llu rows %llu replay %llu deg %llu | solo: round-body %llu bar %llu fence %llu acquire %llu\n", g,
                pr[4 * g], pr[4 * g + 1], pr[4 * g + 2], pr[4 * g + 3], g < 2000 ? pr[16000 + 4 * g] : 0,
                g < 2000 ? pr[16000 + 4 * g + 1] : 0, g < 2000 ? pr[16000 + 4 * g + 2] : 0, g < 2000 ? pr[16000 + 4 * g + 3] : 0, g < 1000 ? pr[24000 + 4 * g] : 0, g < 1000 ? pr[24000 + 4 * g + 1] : 0,
                g < 1000 ? pr[24000 + 4 * g + 2] : 0, g < 1000 ? pr[24000 + 4 * g + 3] : 0);
    }
  }
  // kernels of one call: reset (wires, known), the persistent solve, verdict (pack, targets)
  const uint64_t launches = 1 + 1 + (d.n_known ? 1 : 0) + 1 + (d.n_targets ? 1 : 0);
  // the three whole-set sweeps of every outer round visit: P2 the rows that can still fire (counted by
  // the kernel), P3 / P4 the rows with the ABZ / IsZero shape
  const unsigned long long rounds_total = R.h_status->rounds;
  const uint64_t phase_evals = d.rank == 0 ? outer * ((uint64_t)d.n_p3 + d.n_p4) : 0;  // replicated: counted once
  cudaEventRecord(e1, s);
  res->status = status;
  if (status != ECNE_OK) {
    for (auto& Ri : h->rs) Ri.table_dirty = true;
    cudaStreamSynchronize(s);
    return fail(status, err.empty() ? status_text(status) : err);
  }
  // verdict + D2H
  launch_finalize(d, 0, R.d_ubits, R.d_kbits, R.d_counts, s);
  const size_t words = (R.n_vars + 63) / 64;
  CKA(cudaMemcpyAsync(res->unique_bits, R.d_ubits, words * 8, cudaMemcpyDeviceToHost, s));
  CKA(cudaMemcpyAsync(res->known_bits, R.d_kbits, words * 8, cudaMemcpyDeviceToHost, s));
  CKA(cudaMemcpyAsync(R.h_counts, R.d_counts, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  if (res->lb || res->ub || res->nvalues || res->values) {
    Arena t;
    fr::u256 *dl = nullptr, *du = nullptr, *dv = nullptr;
    uint8_t* dn = nullptr;
    const size_t V = R.n_vars;
    if (res->lb) CKA(t.alloc(&dl, V));
    if (res->ub) CKA(t.alloc(&du, V));
    if (res->values) CKA(t.alloc(&dv, 2 * V));
    if (res->nvalues) CKA(t.alloc(&dn, V));
    launch_export(d, 0, dl, du, dn, dv, s);
    if (dl) CKA(cudaMemcpyAsync(res->lb, dl, V * 32, cudaMemcpyDeviceToHost, s));
    if (du) CKA(cudaMemcpyAsync(res->ub, du, V * 32, cudaMemcpyDeviceToHost, s));
    if (dv) CKA(cudaMemcpyAsync(res->values, dv, V * 64, cudaMemcpyDeviceToHost, s));
    if (dn) CKA(cudaMemcpyAsync(res->nvalues, dn, V, cudaMemcpyDeviceToHost, s));
    CKA(cudaStreamSynchronize(s));
    t.release();
  }
  if (res->abz) CKA(cudaMemcpyAsync(res->abz, d.abz + 1, R.n_vars * 4, cudaMemcpyDeviceToHost, s));
  cudaEventRecord(e2, s);
  CKA(cudaStreamSynchronize(s));
  CKA(cudaGetLastError());
  float ms_solve = 0, ms_d2h = 0, ms_device = 0;
  cudaEventElapsedTime(&ms_solve, e0, e1);
  cudaEventElapsedTime(&ms_d2h, e1, e2);
  cudaEventElapsedTime(&ms_device, e0, e2);
  res->n_unique = R.h_counts[0];
  res->n_nontrivial = R.h_counts[1];
  res->n_unique_nontrivial = R.h_counts[2];
  res->n_targets_unique = R.h_counts[3];
  res->verdict = (R.h_counts[3] == R.n_targets) ? 1 : 0;
  res->outer_rounds = outer;
  res->inner_rounds = rounds_total;
  res->constraint_evals = evals_total + phase_evals;
  res->rule_evals = rule_evals_total + phase_evals;
  res->sweep_launches = launches;
  res->ms_h2d = R.ms_h2d;
  res->ms_classify = R.ms_classify;
  res->ms_solve = ms_solve;
  res->ms_d2h = ms_d2h;
  res->ms_exchange = 0;
  res->ms_sweep = ms_sweep;
  res->dense_rounds = R.h_status->dense_rounds;
  res->dense_evals = dense_evals_total;
  res->dense_cycles = dense_cycles_max;
  res->ms_device = ms_device;
  res->gpus_used = (uint64_t)d.world;
  res->sharded = (uint64_t)d.shard;
  for (auto& Ri : h->rs) Ri.have_state = true;
  res->ms_total = R.ms_h2d + R.ms_classify + ms_solve + ms_d2h;
  R.have_state = true;
  return ECNE_OK;
}

// ---- report path (:1599-1635) ----------------------------------------------------------------------
extern "C" int ecne_report_resident(ecne_resident_t* h, ecne_report_t* rep) {
  if (!h || !rep || !rep->bad_row_bits) return fail(ECNE_E_BADARG, "null argument");
  Resident& R = h->rs[0];
  const Dev& d = R.d;
  cudaSetDevice(R.device);
  if (!R.have_state) return fail(ECNE_E_BADARG, "ecne_report_resident: no successful solve on this handle");
  cudaStream_t s = R.stream;
  Arena t;
  struct Release {
    Arena& a;
    ~Release() { a.release(); }
  } release{t};
  const size_t row_words = ((size_t)R.n_rows + 63) / 64;  // 64-bit words of the row bitmap
  unsigned int *d_rows = nullptr, *d_mark = nullptr;
  uint32_t* d_wires = nullptr;
  unsigned long long* d_counts = nullptr;
  CKA(t.alloc(&d_rows, 2 * row_words));
  CKA(t.alloc(&d_mark, ((size_t)d.V + 31) / 32));
  CKA(t.alloc(&d_wires, (size_t)d.V));
  CKA(t.alloc(&d_counts, 2));
  launch_bad_rows(d, 0, d_rows, d_mark, d_wires, d_counts, s);
  unsigned long long counts[2] = {0, 0};
  CKA(cudaMemcpyAsync(rep->bad_row_bits, d_rows, row_words * 8, cudaMemcpyDeviceToHost, s));
  CKA(cudaMemcpyAsync(counts, d_counts, sizeof(counts), cudaMemcpyDeviceToHost, s));
  CKA(cudaStreamSynchronize(s));
  CKA(cudaGetLastError());
  rep->n_bad_rows = counts[0];
  rep->n_wires = counts[1];
  if (!rep->wire) return ECNE_OK;  // bitmap and counts only
  if (rep->cap_wires < counts[1])
    return fail(ECNE_E_BADARG, "ecne_report_resident: cap_wires " + std::to_string(rep->cap_wires) + " < n_wires " +
                                   std::to_string(counts[1]));
  const size_t n = (size_t)counts[1];
  if (n == 0) return ECNE_OK;
  uint8_t *df = nullptr, *dn = nullptr;
  fr::u256 *dl = nullptr, *du = nullptr, *dv = nullptr;
  if (rep->flags) CKA(t.alloc(&df, n));
  if (rep->lb) CKA(t.alloc(&dl, n));
  if (rep->ub) CKA(t.alloc(&du, n));
  if (rep->nvalues) CKA(t.alloc(&dn, n));
  if (rep->values) CKA(t.alloc(&dv, 2 * n));
  launch_report_export(d, 0, d_wires, (uint32_t)n, df, dl, du, dn, dv, s);
  CKA(cudaMemcpyAsync(rep->wire, d_wires, n * 4, cudaMemcpyDeviceToHost, s));
  if (df) CKA(cudaMemcpyAsync(rep->flags, df, n, cudaMemcpyDeviceToHost, s));
  if (dl) CKA(cudaMemcpyAsync(rep->lb, dl, n * 32, cudaMemcpyDeviceToHost, s));
  if (du) CKA(cudaMemcpyAsync(rep->ub, du, n * 32, cudaMemcpyDeviceToHost, s));
  if (dn) CKA(cudaMemcpyAsync(rep->nvalues, dn, n, cudaMemcpyDeviceToHost, s));
  if (dv) CKA(cudaMemcpyAsync(rep->values, dv, n * 64, cudaMemcpyDeviceToHost, s));
  CKA(cudaStreamSynchronize(s));
  CKA(cudaGetLastError());
  return ECNE_OK;
}

extern "C" int ecne_solve(const ecne_problem_t* problem, ecne_result_t* result) {
  if (!problem || !result) return fail(ECNE_E_BADARG, "null argument");
  auto t0 = std::chrono::steady_clock::now();
  ecne_resident_t* h = nullptr;
  int st = ecne_upload(problem, &h);
  if (st != ECNE_OK) {
    result->status = st;
    return st;
  }
  st = ecne_solve_resident(h, result);
  ecne_free_resident(h);
  auto t1 = std::chrono::steady_clock::now();
  if (st == ECNE_OK) result->ms_total = std::chrono::duration<double, std::milli>(t1 - t0).count();
  return st;
}

// ---- row-range sharding across the GPUs of one box --------------------------------
extern "C" int ecne_shard_rows(const ecne_problem_t* p, int rank, int world, uint64_t* lo, uint64_t* hi) {
  if (!p || !lo || !hi || world < 1 || rank < 0 || rank >= world) return fail(ECNE_E_BADARG, "bad argument");
  const uint64_t N = p->n_rows;
  if (!p->seg_ptr && !p->seg_ptr32) return fail(ECNE_E_BADARG, "null offsets");
  auto seg = [&](uint64_t i) -> uint64_t { return p->seg_ptr ? p->seg_ptr[i] : (uint64_t)p->seg_ptr32[i]; };
  const uint64_t total = N ? seg(3 * N) : 0;
  auto cut = [&](int r) -> uint64_t {  // first row whose prefix term count reaches r/world of the total
    if (r <= 0) return 0;
    if (r >= world) return N;
    const uint64_t want = (uint64_t)(((unsigned __int128)total * (unsigned)r) / (unsigned)world);
    uint64_t a = 0, b = N;  // smallest row i with seg_ptr[3*i] >= want
    while (a < b) {
      uint64_t m = (a + b) / 2;
      if (seg(3 * m) >= want)
        b = m;
      else
        a = m + 1;
    }
    return a;
  };
  *lo = cut(rank);
  *hi = cut(rank + 1);
  return ECNE_OK;
}

extern "C" int ecne_dist_unique_id(uint8_t out[128]) {
  if (!NCCL.load()) return fail(ECNE_E_NCCL, "cannot load libnccl.so.2");
  ncclUniqueId id;
  if (NCCL.GetUniqueId(&id) != ncclSuccess) return fail(ECNE_E_NCCL, "ncclGetUniqueId failed");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(out, &id, 128);
  return ECNE_OK;
}
extern "C" int ecne_dist_init(int rank, int world, const uint8_t unique_id[128]) {
  if (!G.inited) return fail(ECNE_E_CUDA, "call ecne_init first");
  if (G.multi) return fail(ECNE_E_BADARG, "ecne_dist_init: this process drives several GPUs itself (ecne_init_multi)");
  if (world < 1 || world > ECNE_MAX_WORLD || rank < 0 || rank >= world) return fail(ECNE_E_BADARG, "bad rank/world");
  if (G.comm) {
    NCCL.CommDestroy(G.comm);
    G.comm = nullptr;
  }
  G.rank = rank;
  G.world = world;
  upload_shard() = UploadShard();
  if (world == 1) return ECNE_OK;
  if (!unique_id) return fail(ECNE_E_BADARG, "null unique id");
  if (!NCCL.load()) return fail(ECNE_E_NCCL, "cannot load libnccl.so.2");
  ncclUniqueId id;
  memcpy(&id, unique_id, 128);
  cudaSetDevice(G.ctx[0]->device);
  if (NCCL.CommInitRank(&G.comm, world, id, rank) != ncclSuccess)
    return fail(ECNE_E_NCCL, "ncclCommInitRank failed");
  // sharded upload (engine_host.h UploadShard): every rank copies its slice, the slices are gathered over NVLink
  UploadShard& U = upload_shard();
  U.rank = rank;
  U.world = world;
  U.allgather = [](void* buf, size_t slice, cudaStream_t s) -> cudaError_t {
    if (!G.comm || !G.shard_upload) return cudaErrorNotReady;
    return NCCL.AllGather((const char*)buf + (size_t)G.rank * slice, buf, slice, ncclChar, G.comm, s) == ncclSuccess
               ? cudaSuccess
               : cudaErrorUnknown;
  };
  if (!G.shard_upload) U.allgather = nullptr;
  return ECNE_OK;
}
extern "C" int ecne_dist_rank(void) { return G.rank; }
extern "C" int ecne_dist_world(void) { return G.total_world(); }

namespace {
// (Re)build the exchange channel so that every list holds at least `cap` records, and point the
// residents' record lists / peer tables at it.  Collective: every rank calls it with the same cap.
int setup_exchange(ecne_resident& H, size_t cap) {
  const int nl = G.n_local(), world = G.total_world();
  if (G.xcap < cap) {
    if (!G.multi)
      for (int h = 0; h < world; ++h)
        if (h != G.rank && G.xpeer[h]) cudaIpcCloseMemHandle(G.xpeer[h]);
    memset(G.xpeer, 0, sizeof(G.xpeer));
    G.xcap = 0;
    const size_t bytes = XH_BYTES + 3 * cap * sizeof(Rec);
    for (int i = 0; i < nl; ++i) {
      Ctx* c = G.ctx[i];
      CKA(cudaSetDevice(c->device));
      CKA(cudaDeviceSynchronize());
      if (c->xbuf) cudaFree(c->xbuf);
      c->xbuf = nullptr;
      CKA(cudaMalloc((void**)&c->xbuf, bytes));
      CKA(cudaMemset(c->xbuf, 0, XH_BYTES));
      CKA(cudaDeviceSynchronize());  // zeroed before any peer can learn the address
    }
    CKA(cudaSetDevice(G.ctx[0]->device));
    if (G.multi) {
      for (int h = 0; h < nl; ++h) G.xpeer[h] = G.ctx[h]->xbuf;  // plain peer access inside one process
    } else {
      Ctx* c = G.ctx[0];
      cudaStream_t s = c->stream;
      cudaIpcMemHandle_t mine;
      CKA(cudaIpcGetMemHandle(&mine, c->xbuf));
      // all-gather the handles through NCCL
      char* d_all = nullptr;
      CKA(cudaMalloc((void**)&d_all, sizeof(cudaIpcMemHandle_t) * (size_t)(world + 1)));
      CKA(cudaMemcpyAsync(d_all + sizeof(mine) * world, &mine, sizeof(mine), cudaMemcpyHostToDevice, s));
      if (NCCL.AllGather(d_all + sizeof(mine) * world, d_all, sizeof(mine), ncclChar, G.comm, s) != ncclSuccess)
        return fail(ECNE_E_NCCL, "ncclAllGather of the IPC handles failed");
      std::vector<cudaIpcMemHandle_t> all(world);
      CKA(cudaMemcpyAsync(all.data(), d_all, sizeof(mine) * world, cudaMemcpyDeviceToHost, s));
      CKA(cudaStreamSynchronize(s));
      cudaFree(d_all);
      for (int h = 0; h < world; ++h) {
        if (h == G.rank) {
          G.xpeer[h] = c->xbuf;
        } else {
          void* p = nullptr;
          cudaError_t e = cudaIpcOpenMemHandle(&p, all[h], cudaIpcMemLazyEnablePeerAccess);
          if (e != cudaSuccess)
            return fail(ECNE_E_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
          G.xpeer[h] = (char*)p;
        }
      }
    }
    G.xcap = cap;
  }
  for (int i = 0; i < nl; ++i) {
    Resident& R = H.rs[i];
    Dev& d = R.d;
    char* own = G.ctx[i]->xbuf;
    d.world = world;
    d.rank = G.multi ? i : G.rank;
    // d.rec_cap stays this problem's own capacity: the phase lists 3 / 4 live in the problem's arena with exactly
    // that many slots, and the exchange lists (G.xcap >= cap slots each, the stride every rank uses) hold at least
    // as many
    for (int l = 0; l < 3; ++l) d.recs[l] = (Rec*)(own + XH_BYTES + (size_t)l * G.xcap * sizeof(Rec));
    for (int h = 0; h < world; ++h) {
      d.xflag[h] = (unsigned long long*)G.xpeer[h];
      for (int l = 0; l < 3; ++l) d.xrecs[h][l] = (Rec*)(G.xpeer[h] + XH_BYTES + (size_t)l * G.xcap * sizeof(Rec));
    }
    d.xcnt = (unsigned int*)(own + XH_XCNT_OFF);
    d.xepoch = (unsigned int*)(own + XH_EPOCH_OFF);
    R.xhdr = own;
  }
  return ECNE_OK;
}
}  // namespace

// ---- field known-answer hook --------------------------------------------------------------------
__global__ void k_fr_batch(int op, uint64_t n, const fr::u256* a, const fr::u256* b, fr::u256* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fr::u256 x = a[i], y = b ? b[i] : fr::make_u256(0, 0, 0, 0), r;
  switch (op) {
    case 0: r = fr::add(x, y); break;
    case 1: r = fr::sub(x, y); break;
    case 2: r = fr::from_mont(fr::mul(fr::to_mont(x), fr::to_mont(y))); break;
    case 3: r = fr::is_zero(x) ? x : fr::from_mont(fr::inv_mont(fr::to_mont(x))); break;
    case 4: r = fr::neg(x); break;
    case 5: r = fr::neg_div(x, y); break;  // divexact(-x, y)
    case 6: r = fr::make_u256(fr::divides(y, x) ? 1 : 0, 0, 0, 0); break;  // plain integers: y | x  (y != 0)
    case 7: r = fr::make_u256((uint64_t)(fr::cmp_mul(x, y, fr::modulus()) + 1), 0, 0, 0); break;  // sign(x*y - p) + 1
    default: r = x;
  }
  out[i] = r;
}
extern "C" int ecne_fr_batch(int op, uint64_t n, const uint64_t* a, const uint64_t* b, uint64_t* out) {
  if (!G.inited) {
    int st = ecne_init(0);
    if (st) return st;
  }
  if (!a || !out) return fail(ECNE_E_BADARG, "null argument");
  fr::u256 *da = nullptr, *db = nullptr, *dout = nullptr;
  Arena t;
  CKA(t.alloc(&da, n));
  CKA(t.alloc(&dout, n));
  CKA(cudaMemcpy(da, a, n * 32, cudaMemcpyHostToDevice));
  if (b) {
    CKA(t.alloc(&db, n));
    CKA(cudaMemcpy(db, b, n * 32, cudaMemcpyHostToDevice));
  }
  if (n) k_fr_batch<<<(unsigned int)((n + 127) / 128), 128>>>(op, n, da, db, dout);
  CKA(cudaDeviceSynchronize());
  CKA(cudaMemcpy(out, dout, n * 32, cudaMemcpyDeviceToHost));
  t.release();
  return ECNE_OK;
}
