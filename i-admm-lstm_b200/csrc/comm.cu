// The one collective of the path (models of main.py:336-358 run data parallel): a sum all-reduce of the
// flat LSTM gradient buffer (2,570,601 floats at hidden_dim 800, K = 100) over NCCL / NVLink once per truncated-BPTT window.
// The reference is single process and has no collective at all; this entry point is what the proposed ABI of SURVEY.md section
// 8(b) calls iadmm_allreduce_grads.  NCCL is not a link-time dependency of the library (it must load on a CPU-only box for the
// ABI checks): the already loaded libnccl.so.2 of the host process (torch's) is looked up at the first call.
#include "common.cuh"

#include <dlfcn.h>
#include <string.h>

namespace iadmm {

struct NcclUid { char internal[128]; };             // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128), passed by value
typedef int (*NcclAllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*NcclErrFn)(int);
typedef int (*NcclGetUidFn)(NcclUid*);
typedef int (*NcclInitRankFn)(void**, int, NcclUid, int);
typedef int (*NcclDestroyFn)(void*);

static NcclAllReduceFn g_allreduce = nullptr;
static NcclErrFn g_errstr = nullptr;
static NcclGetUidFn g_getuid = nullptr;
static NcclInitRankFn g_initrank = nullptr;
static NcclDestroyFn g_destroy = nullptr;

static int load_nccl() {
  if (g_allreduce) return IADMM_OK;
  void* h = dlopen("libnccl.so.2", RTLD_LAZY | RTLD_NOLOAD);      // the copy the host process already uses
  if (!h) h = dlopen("libnccl.so.2", RTLD_LAZY);
  if (!h) h = dlopen("libnccl.so", RTLD_LAZY);
  if (!h) IADMM_FAIL(IADMM_ECUDA, "allreduce_grads: libnccl.so.2 is not loaded and cannot be found (%s)", dlerror());
  g_errstr = reinterpret_cast<NcclErrFn>(dlsym(h, "ncclGetErrorString"));
  g_getuid = reinterpret_cast<NcclGetUidFn>(dlsym(h, "ncclGetUniqueId"));
  g_initrank = reinterpret_cast<NcclInitRankFn>(dlsym(h, "ncclCommInitRank"));
  g_destroy = reinterpret_cast<NcclDestroyFn>(dlsym(h, "ncclCommDestroy"));
  NcclAllReduceFn f = reinterpret_cast<NcclAllReduceFn>(dlsym(h, "ncclAllReduce"));
  if (!f) IADMM_FAIL(IADMM_ECUDA, "allreduce_grads: ncclAllReduce not found in libnccl");
  g_allreduce = f;
  return IADMM_OK;
}

__global__ void __launch_bounds__(256) scale_kernel(float* __restrict__ a, size_t count, float s) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) a[i] = __fmul_rn(a[i], s);
}

}  // namespace iadmm

using namespace iadmm;

extern "C" int iadmm_allreduce_grads(float* flat_grads, size_t count, float scale, void* nccl_comm, void* stream) {
  if (!flat_grads || !nccl_comm) IADMM_FAIL(IADMM_EALIGN, "allreduce_grads: NULL pointer");
  if (count == 0) return IADMM_OK;
  int rc = load_nccl();
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ncclFloat32 = 7, ncclSum = 0;                           // nccl.h: ncclDataType_t / ncclRedOp_t
  const int r = g_allreduce(flat_grads, flat_grads, count, ncclFloat32, ncclSum, nccl_comm, st);
  if (r != 0) IADMM_FAIL(IADMM_ECUDA, "allreduce_grads: ncclAllReduce failed: %s", g_errstr ? g_errstr(r) : "unknown NCCL error");
  if (scale != 1.0f) {
    scale_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(flat_grads, count, scale);
    IADMM_LAUNCH_CHECK("scale_kernel");
  }
  return IADMM_OK;
}

// Communicator set-up for iadmm_allreduce_grads without going through the host framework: rank 0 creates the 128-byte unique id,
// the caller hands it to the other ranks by any means (torch.distributed's store in iadmm_b200/dist.py), every rank joins.
extern "C" int iadmm_nccl_unique_id(void* uid128) {
  if (!uid128) IADMM_FAIL(IADMM_EALIGN, "nccl_unique_id: NULL pointer");
  int rc = load_nccl();
  if (rc) return rc;
  if (!g_getuid) IADMM_FAIL(IADMM_ECUDA, "nccl_unique_id: ncclGetUniqueId not found in libnccl");
  NcclUid u;
  const int r = g_getuid(&u);
  if (r != 0) IADMM_FAIL(IADMM_ECUDA, "ncclGetUniqueId failed: %s", g_errstr ? g_errstr(r) : "unknown NCCL error");
  memcpy(uid128, &u, sizeof(u));
  return IADMM_OK;
}

extern "C" int iadmm_nccl_comm_init(void** comm, int world, const void* uid128, int rank) {
  if (!comm || !uid128 || world <= 0 || rank < 0 || rank >= world) IADMM_FAIL(IADMM_ESHAPE, "nccl_comm_init: world=%d rank=%d", world, rank);
  int rc = load_nccl();
  if (rc) return rc;
  if (!g_initrank) IADMM_FAIL(IADMM_ECUDA, "nccl_comm_init: ncclCommInitRank not found in libnccl");
  NcclUid u;
  memcpy(&u, uid128, sizeof(u));
  const int r = g_initrank(comm, world, u, rank);
  if (r != 0) IADMM_FAIL(IADMM_ECUDA, "ncclCommInitRank failed: %s", g_errstr ? g_errstr(r) : "unknown NCCL error");
  return IADMM_OK;
}

extern "C" int iadmm_nccl_comm_destroy(void* comm) {
  if (!comm) return IADMM_OK;
  int rc = load_nccl();
  if (rc) return rc;
  if (g_destroy) g_destroy(comm);
  return IADMM_OK;
}
