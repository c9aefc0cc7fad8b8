// Thin C-ABI wrappers over NCCL for the ONE collective of the path: the in-place all-reduce (SUM) of the flat gradient arena
// (reference: torch DDP's bucketed ncclAllReduce behind wesep/bin/train.py:227-228).  libnccl.so.2 is bound at run time with
// dlopen so that libwesep_b200.so itself carries no link-time dependency on it (single-GPU users never touch it); in a
// PyTorch process the already-loaded NCCL of the torch wheel is the one that resolves.  No torch types cross the boundary:
// the caller exchanges the 128-byte unique id over whatever side channel it has (torch.distributed / MPI / a file).
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace wb {

struct NcclId { char internal[128]; };          // == ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES 128)
using nccl_comm_t = void*;
using fn_get_id = int (*)(NcclId*);
using fn_init_rank = int (*)(nccl_comm_t*, int, NcclId, int);
using fn_allreduce = int (*)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
using fn_destroy = int (*)(nccl_comm_t);
using fn_errstr = const char* (*)(int);

static struct {
  void* lib = nullptr;
  fn_get_id get_id = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_destroy destroy = nullptr;
  fn_errstr errstr = nullptr;
  bool tried = false;
} g_nccl;
static std::mutex g_nccl_mu;

static bool nccl_load() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.tried) return g_nccl.lib != nullptr;
  g_nccl.tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return false;
  g_nccl.get_id = (fn_get_id)dlsym(h, "ncclGetUniqueId");
  g_nccl.init_rank = (fn_init_rank)dlsym(h, "ncclCommInitRank");
  g_nccl.allreduce = (fn_allreduce)dlsym(h, "ncclAllReduce");
  g_nccl.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
  g_nccl.errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
  if (!g_nccl.get_id || !g_nccl.init_rank || !g_nccl.allreduce || !g_nccl.destroy) return false;
  g_nccl.lib = h;
  return true;
}
static int nccl_fail(const char* what, int rc) {
  snprintf(g_err, sizeof(g_err), "%s: NCCL error %d (%s)", what, rc, g_nccl.errstr ? g_nccl.errstr(rc) : "?");
  return -3;
}

}  // namespace wb

using namespace wb;

extern "C" int wesep_b200_nccl_available(void) { return nccl_load() ? 1 : 0; }

extern "C" int wesep_b200_nccl_unique_id(void* id128) {
  if (!id128) return fail(-1, "nccl_unique_id: null");
  if (!nccl_load()) return fail(-2, "nccl: libnccl.so.2 could not be loaded");
  if (int rc = g_nccl.get_id(reinterpret_cast<NcclId*>(id128))) return nccl_fail("ncclGetUniqueId", rc);
  return 0;
}

extern "C" int wesep_b200_nccl_comm_init_rank(void** comm, int nranks, const void* id128, int rank) {
  if (!comm || !id128 || nranks <= 0 || rank < 0 || rank >= nranks) return fail(-1, "nccl_comm_init_rank: bad arguments");
  if (!nccl_load()) return fail(-2, "nccl: libnccl.so.2 could not be loaded");
  NcclId id = *reinterpret_cast<const NcclId*>(id128);
  if (int rc = g_nccl.init_rank(comm, nranks, id, rank)) return nccl_fail("ncclCommInitRank", rc);
  return 0;
}

extern "C" int wesep_b200_nccl_allreduce_flat(float* buf, int64_t count, void* comm, void* stream) {
  if (!buf || count <= 0 || !comm) return fail(-1, "nccl_allreduce_flat: bad arguments");
  if (!nccl_load()) return fail(-2, "nccl: libnccl.so.2 could not be loaded");
  constexpr int kFloat = 7, kSum = 0;            // ncclFloat32, ncclSum
  if (int rc = g_nccl.allreduce(buf, buf, (size_t)count, kFloat, kSum, comm, (cudaStream_t)stream)) return nccl_fail("ncclAllReduce", rc);
  return 0;
}

extern "C" int wesep_b200_nccl_comm_destroy(void* comm) {
  if (!comm) return 0;
  if (!nccl_load()) return fail(-2, "nccl: libnccl.so.2 could not be loaded");
  if (int rc = g_nccl.destroy(comm)) return nccl_fail("ncclCommDestroy", rc);
  return 0;
}
