// The one collective of the data-parallel path behind the C ABI: ncclAllReduce(SUM, fp32) over the flat
// gradient buffer.  NCCL is bound at run time (dlopen of libnccl.so.2: the copy a host process has already loaded -- e.g.
// the one PyTorch ships -- or the system one), so the library itself does not link against a particular NCCL build.
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace cvb {

struct NcclUniqueId { char internal[128]; };   // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES 128)
typedef void* NcclComm;
typedef int (*fn_get_unique_id)(NcclUniqueId*);
typedef int (*fn_comm_init_rank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_comm_destroy)(NcclComm);
typedef const char* (*fn_error_string)(int);

struct NcclApi {
    void* handle = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_error_string error_string = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int nccl_api(NcclApi** out) {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (!g_nccl.handle) {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // already in the process?
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        CVB_REQUIRE(h, "libnccl.so.2 not found (%s): the data-parallel collective needs NCCL", dlerror());
        NcclApi a;
        a.handle = h;
        a.get_unique_id = (fn_get_unique_id)dlsym(h, "ncclGetUniqueId");
        a.comm_init_rank = (fn_comm_init_rank)dlsym(h, "ncclCommInitRank");
        a.all_reduce = (fn_all_reduce)dlsym(h, "ncclAllReduce");
        a.comm_destroy = (fn_comm_destroy)dlsym(h, "ncclCommDestroy");
        a.error_string = (fn_error_string)dlsym(h, "ncclGetErrorString");
        CVB_REQUIRE(a.get_unique_id && a.comm_init_rank && a.all_reduce && a.comm_destroy && a.error_string,
                    "libnccl.so.2 lacks an expected symbol");
        g_nccl = a;
    }
    *out = &g_nccl;
    return 0;
}

#define CVB_NCCL(api, call)                                                                    \
    do {                                                                                       \
        int rc_ = (call);                                                                      \
        CVB_REQUIRE(rc_ == 0, "NCCL: %s (%s)", (api)->error_string(rc_), #call);              \
    } while (0)

}  // namespace cvb

using namespace cvb;

extern "C" {

int cvb_comm_unique_id(void* id128) {
    CVB_REQUIRE(id128, "cvb_comm_unique_id: NULL buffer");
    NcclApi* n;
    if (int rc = nccl_api(&n)) return rc;
    NcclUniqueId id;
    CVB_NCCL(n, n->get_unique_id(&id));
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int cvb_comm_init(void** comm, const void* id128, int rank, int world) {
    CVB_REQUIRE(comm && id128 && world >= 1 && rank >= 0 && rank < world, "cvb_comm_init: bad arguments (rank %d of %d)", rank, world);
    NcclApi* n;
    if (int rc = nccl_api(&n)) return rc;
    NcclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NcclComm c = nullptr;
    CVB_NCCL(n, n->comm_init_rank(&c, world, id, rank));   // on the CURRENT CUDA device: one rank per GPU
    *comm = c;
    return 0;
}

int cvb_allreduce_sum(void* comm, float* flat, size_t n_floats, void* stream) {
    CVB_REQUIRE(comm && flat, "cvb_allreduce_sum: NULL argument");
    if (n_floats == 0) return 0;
    NcclApi* n;
    if (int rc = nccl_api(&n)) return rc;
    // in place, SUM (not mean: the reference sums per-utterance losses, train_*.py:1403,1408), ncclFloat32 = 7, ncclSum = 0
    CVB_NCCL(n, n->all_reduce(flat, flat, n_floats, 7, 0, (NcclComm)comm, (cudaStream_t)stream));
    return 0;
}

int cvb_comm_destroy(void* comm) {
    if (!comm) return 0;
    NcclApi* n;
    if (int rc = nccl_api(&n)) return rc;
    CVB_NCCL(n, n->comm_destroy((NcclComm)comm));
    return 0;
}

}  // extern "C"
