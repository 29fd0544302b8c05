// the gradient exchange through the C ABI for hosts that are not PyTorch -- a thin layer over NCCL
// (ncclAllReduce over NVLink / NVSwitch), bound at RUN time with dlopen so that the library has no link-time dependency on
// a particular libnccl: inside a PyTorch process the copy torch already loaded is found, elsewhere the system's libnccl.so.2.
//   cgat_comm_unique_id   rank 0 creates the 128-byte id and ships it to the other ranks by its own means (MPI, a socket)
//   cgat_comm_init        one communicator per process / GPU, created once
//   cgat_flat_allreduce   in-place SUM over the flat fp32 gradient buffer (the 1/world is folded into the Adam kernels)
//   cgat_comm_destroy
// The Python package does not use these (torch.distributed owns its communicator there; small models take the
// peer-memory kernel cgat_p2p_allreduce_adam instead).
#include "common.cuh"
#include <dlfcn.h>

namespace cgat {

struct NcclId { char internal[128]; };  // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
using nccl_get_unique_id_t = int (*)(NcclId*);
using nccl_comm_init_rank_t = int (*)(void**, int, NcclId, int);
using nccl_all_reduce_t = int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t);
using nccl_comm_destroy_t = int (*)(void*);
using nccl_get_error_string_t = const char* (*)(int);

struct NcclApi {
  void* handle = nullptr;
  nccl_get_unique_id_t get_unique_id = nullptr;
  nccl_comm_init_rank_t comm_init_rank = nullptr;
  nccl_all_reduce_t all_reduce = nullptr;
  nccl_comm_destroy_t comm_destroy = nullptr;
  nccl_get_error_string_t error_string = nullptr;
};

static const NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.handle = dlopen(name, RTLD_NOW | RTLD_NOLOAD);  // the copy the process already has (torch's)
      if (!api.handle) api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (api.handle) {
      api.get_unique_id = (nccl_get_unique_id_t)dlsym(api.handle, "ncclGetUniqueId");
      api.comm_init_rank = (nccl_comm_init_rank_t)dlsym(api.handle, "ncclCommInitRank");
      api.all_reduce = (nccl_all_reduce_t)dlsym(api.handle, "ncclAllReduce");
      api.comm_destroy = (nccl_comm_destroy_t)dlsym(api.handle, "ncclCommDestroy");
      api.error_string = (nccl_get_error_string_t)dlsym(api.handle, "ncclGetErrorString");
    }
  }
  if (!api.handle || !api.get_unique_id || !api.comm_init_rank || !api.all_reduce || !api.comm_destroy) return nullptr;
  return &api;
}

static int nccl_fail(const NcclApi* api, int rc, const char* what) {
  return fail(CGAT_EINVAL, "%s: NCCL error %d (%s)", what, rc, api->error_string ? api->error_string(rc) : "?");
}

}  // namespace cgat

using namespace cgat;

extern "C" int cgat_comm_available(void) { return nccl_api() != nullptr; }

extern "C" int cgat_comm_unique_id(void* id128) {
  if (!id128) return fail(CGAT_EINVAL, "null argument");
  const NcclApi* api = nccl_api();
  if (!api) return fail(CGAT_EUNSUPPORTED, "libnccl.so.2 not found (dlopen)");
  if (int rc = api->get_unique_id(reinterpret_cast<NcclId*>(id128))) return nccl_fail(api, rc, "ncclGetUniqueId");
  return 0;
}

extern "C" int cgat_comm_init(int32_t rank, int32_t world, const void* id128, void** comm_out) {
  if (!id128 || !comm_out || world < 1 || rank < 0 || rank >= world) return fail(CGAT_EINVAL, "bad argument");
  const NcclApi* api = nccl_api();
  if (!api) return fail(CGAT_EUNSUPPORTED, "libnccl.so.2 not found (dlopen)");
  NcclId id;
  memcpy(&id, id128, sizeof(id));
  if (int rc = api->comm_init_rank(comm_out, world, id, rank)) return nccl_fail(api, rc, "ncclCommInitRank");
  return 0;
}

extern "C" int cgat_flat_allreduce(void* comm, float* buf, int64_t n, void* stream) {
  if (!comm || !buf || n < 1) return fail(CGAT_EINVAL, "bad argument");
  const NcclApi* api = nccl_api();
  if (!api) return fail(CGAT_EUNSUPPORTED, "libnccl.so.2 not found (dlopen)");
  // ncclFloat32 = 7, ncclSum = 0 (nccl.h)
  if (int rc = api->all_reduce(buf, buf, (size_t)n, 7, 0, comm, (cudaStream_t)stream)) return nccl_fail(api, rc, "ncclAllReduce");
  return 0;
}

extern "C" int cgat_comm_destroy(void* comm) {
  if (!comm) return fail(CGAT_EINVAL, "null communicator");
  const NcclApi* api = nccl_api();
  if (!api) return fail(CGAT_EUNSUPPORTED, "libnccl.so.2 not found (dlopen)");
  if (int rc = api->comm_destroy(comm)) return nccl_fail(api, rc, "ncclCommDestroy");
  return 0;
}
