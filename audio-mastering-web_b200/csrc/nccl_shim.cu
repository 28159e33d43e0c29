// NCCL, called straight from C on the context's stream: the exchange step of the time-split path (BASELINE config 5).
//
// The chain of a time slice needs four tiny all-reduces per file (channel sums; channel maxima / negated minima; BS.1770 hop
// sums; output peak).  Going through the Python callback of mm_slice costs a ctypes round trip, a tensor wrap and
// torch.distributed's stream hand-over per scalar; here ncclAllReduce is enqueued on the very stream the kernels run on, so
// nothing but the collective itself sits between two kernels.  libnccl is resolved at run time with dlopen("libnccl.so.2") --
// the copy torch has already loaded when there is one (same SONAME), the system's otherwise -- so the library has no link-time
// dependency on NCCL and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "context.h"

namespace mm {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};

static NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
        api.GetVersion = (decltype(api.GetVersion))dlsym(h, "ncclGetVersion");
        if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString) api.handle = h;
    });
    return api.handle ? &api : nullptr;
}

// dtype 0 = float64, 1 = int64, 2 = float32; op 0 = sum, 1 = min, 2 = max (the codes of mm_allreduce_fn)
int nccl_allreduce(mm_ctx* c, void* comm, void* ptr, int64_t count, int dtype, int op) {
    NcclApi* api = nccl_api();
    if (!api) { set_error("NCCL is not available (libnccl.so.2 could not be loaded)"); return 1; }
    const ncclDataType_t dt = dtype == 0 ? ncclFloat64 : (dtype == 1 ? ncclInt64 : ncclFloat32);
    const ncclRedOp_t ro = op == 0 ? ncclSum : (op == 1 ? ncclMin : ncclMax);
    const ncclResult_t r = api->AllReduce(ptr, ptr, (size_t)count, dt, ro, (ncclComm_t)comm, c->stream);
    if (r != ncclSuccess) { set_error("ncclAllReduce failed: %s", api->GetErrorString(r)); return 1; }
    return 0;
}

}  // namespace mm

using namespace mm;

extern "C" {

int mm_nccl_unique_id(void* id128) {
    NcclApi* api = nccl_api();
    if (!api) { set_error("NCCL is not available (libnccl.so.2 could not be loaded)"); return 1; }
    if (!id128) { set_error("mm_nccl_unique_id: null buffer"); return 1; }
    ncclUniqueId id;
    const ncclResult_t r = api->GetUniqueId(&id);
    if (r != ncclSuccess) { set_error("ncclGetUniqueId failed: %s", api->GetErrorString(r)); return 1; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return 0;
}

int mm_nccl_comm_create(mm_ctx* c, const void* id128, int world, int rank, void** comm_out) {
    if (!c || !id128 || !comm_out || world < 1 || rank < 0 || rank >= world) { set_error("mm_nccl_comm_create: bad arguments"); return 1; }
    NcclApi* api = nccl_api();
    if (!api) { set_error("NCCL is not available (libnccl.so.2 could not be loaded)"); return 1; }
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm = nullptr;
    const ncclResult_t r = api->CommInitRank(&comm, world, id, rank);
    if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
    if (r != ncclSuccess) { set_error("ncclCommInitRank failed: %s", api->GetErrorString(r)); return 1; }
    *comm_out = (void*)comm;
    return 0;
}

int mm_nccl_comm_destroy(void* comm) {
    NcclApi* api = nccl_api();
    if (!api || !comm) return 0;
    api->CommDestroy((ncclComm_t)comm);
    return 0;
}

int mm_nccl_version(void) {
    NcclApi* api = nccl_api();
    int v = 0;
    if (api && api->GetVersion) api->GetVersion(&v);
    return v;
}

}  // extern "C"
