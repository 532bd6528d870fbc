// The path's one exchange step behind the C ABI: an NCCL communicator per process / GPU and an in-place
// fp64 sum all-reduce for (a) the packed normal-equation buffer [G | Gbc | s | N | N_bc] of precompute and (b) the
// projected right-hand sides of solve.  The reference has no collective at all (solver.py:162-199 runs on one device);
// a maintainer binding the C ABI (INTEGRATION.md route B) needs nothing but a way to hand rank 0's 128-byte
// ncclUniqueId to the other ranks.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy the host process already carries - torch bundles one -
// or the system library), so libfps_sm100.so has no link-time dependency on it and single-GPU users never load it.
#include "fps_common.cuh"
#include <dlfcn.h>
#include <string.h>

namespace fps {

struct NcclId { char bytes[FPS_COMM_ID_BYTES]; };          // ncclUniqueId: 128 opaque bytes, passed by value
typedef struct ncclComm* NcclComm_t;
typedef int (*GetUniqueIdFn)(NcclId*);
typedef int (*CommInitRankFn)(NcclComm_t*, int, NcclId, int);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, NcclComm_t, cudaStream_t);
typedef int (*CommDestroyFn)(NcclComm_t);
typedef const char* (*GetErrorStringFn)(int);

struct NcclApi {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllReduceFn all_reduce = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  CommDestroyFn comm_abort = nullptr;
  GetErrorStringFn get_error_string = nullptr;
};

static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.handle) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);   // already in the process (torch)?
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  FPS_REQUIRE(h, "fps_comm: libnccl.so.2 not found (%s)", dlerror());
  g_nccl.get_unique_id = (GetUniqueIdFn)dlsym(h, "ncclGetUniqueId");
  g_nccl.comm_init_rank = (CommInitRankFn)dlsym(h, "ncclCommInitRank");
  g_nccl.all_reduce = (AllReduceFn)dlsym(h, "ncclAllReduce");
  g_nccl.comm_destroy = (CommDestroyFn)dlsym(h, "ncclCommDestroy");
  g_nccl.comm_abort = (CommDestroyFn)dlsym(h, "ncclCommAbort");
  g_nccl.get_error_string = (GetErrorStringFn)dlsym(h, "ncclGetErrorString");
  FPS_REQUIRE(g_nccl.get_unique_id && g_nccl.comm_init_rank && g_nccl.all_reduce && g_nccl.comm_destroy,
              "fps_comm: libnccl.so.2 lacks the expected symbols");
  g_nccl.handle = h;
  return 0;
}

#define FPS_NCCL(call)                                                                                        \
  do {                                                                                                        \
    const int r__ = (call);                                                                                   \
    if (r__ != 0) {                                                                                           \
      fps::set_error("%s:%d %s -> NCCL error %d (%s)", __FILE__, __LINE__, #call, r__,                        \
                     fps::g_nccl.get_error_string ? fps::g_nccl.get_error_string(r__) : "?");                 \
      return 1;                                                                                               \
    }                                                                                                         \
  } while (0)

}  // namespace fps

// A communicator is its own object (not part of a context): one per process / GPU serves every Solver of that process, and
// its lifetime is explicit - ncclCommDestroy is an intra-node collective, so it must never run from a garbage collector.
struct fps_comm {
  fps::NcclComm_t nccl;
  int device, rank, world;
};

using namespace fps;

extern "C" {

int fps_comm_unique_id(void* h_id) {
  FPS_REQUIRE(h_id, "fps_comm_unique_id: null argument");
  if (int rc = nccl_load()) return rc;
  NcclId id;
  memset(&id, 0, sizeof(id));
  FPS_NCCL(g_nccl.get_unique_id(&id));
  memcpy(h_id, &id, sizeof(id));
  return 0;
}

int fps_comm_init(fps_comm** out, int device, const void* h_id, int rank, int world) {
  FPS_REQUIRE(out && h_id && world >= 1 && rank >= 0 && rank < world, "fps_comm_init: bad argument");
  *out = nullptr;
  if (int rc = nccl_load()) return rc;
  int prev = -1;
  cudaGetDevice(&prev);
  FPS_CUDA(cudaSetDevice(device));
  NcclId id;
  memcpy(&id, h_id, sizeof(id));
  NcclComm_t comm = nullptr;
  const int r = g_nccl.comm_init_rank(&comm, world, id, rank);
  if (prev >= 0) cudaSetDevice(prev);
  FPS_REQUIRE(r == 0, "fps_comm_init: ncclCommInitRank failed (%d: %s)", r,
              g_nccl.get_error_string ? g_nccl.get_error_string(r) : "?");
  fps_comm* c = new fps_comm();
  c->nccl = comm;
  c->device = device;
  c->rank = rank;
  c->world = world;
  *out = c;
  return 0;
}

int fps_allreduce_f64(fps_comm* comm, double* buf, int64_t count, void* stream) {
  FPS_REQUIRE(comm && comm->nccl && buf && count >= 0, "fps_allreduce_f64: bad argument");
  if (count == 0) return 0;
  // in place, SUM (ncclSum = 0) of doubles (ncclDouble = 8): reproducible for a fixed rank count (NCCL's ring / tree
  // order is deterministic for a given topology and size)
  FPS_NCCL(g_nccl.all_reduce(buf, buf, (size_t)count, 8, 0, comm->nccl, (cudaStream_t)stream));
  return 0;
}

// COLLECTIVE: every rank of the node must call it (ncclCommDestroy finalises the communicator together with its peers).
int fps_comm_destroy(fps_comm* comm) {
  if (!comm) return 0;
  if (comm->nccl && g_nccl.comm_destroy) g_nccl.comm_destroy(comm->nccl);
  delete comm;
  return 0;
}

// Local teardown without coordination (error paths, interpreter shutdown).
int fps_comm_abort(fps_comm* comm) {
  if (!comm) return 0;
  if (comm->nccl && g_nccl.comm_abort) g_nccl.comm_abort(comm->nccl);
  delete comm;
  return 0;
}

}  // extern "C"
