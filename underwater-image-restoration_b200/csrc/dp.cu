// Data-parallel helpers of the training step (SURVEY.md §8b "DP helpers ... NCCL all-reduce wrapper taking
// ncclComm_t", §8e): a thin C wrapper over the NCCL the process already carries (PyTorch's bundled libnccl.so.2),
// resolved at run time with dlopen / dlsym so that libuwr_b200.so has no link-time dependency on it and still loads on
// a machine without NCCL (the CPU-side symbol checks).  A raw ncclAllReduce on a caller-chosen stream can be captured
// into a CUDA graph together with the kernels around it, which torch's ProcessGroupNCCL (watchdog thread, work
// objects) could not on this stack — so the whole data-parallel step replays as ONE graph with the bucket all-reduces
// overlapped with the rest of backward (uwr/train.py, uwr/graph.py).
#include <dlfcn.h>

#include "uwr_common.cuh"
#include "../../include/uwr_b200.h"

namespace {

typedef int (*GetUniqueIdFn)(void* id128);
typedef int (*CommInitRankFn)(void** comm, int nranks, uwr_nccl_id id, int rank);
typedef int (*CommDestroyFn)(void* comm);
typedef int (*AllReduceFn)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t stream);
typedef const char* (*GetErrorStringFn)(int);

struct Nccl {
    void* handle = nullptr;
    GetUniqueIdFn get_unique_id = nullptr;
    CommInitRankFn comm_init_rank = nullptr;
    CommDestroyFn comm_destroy = nullptr;
    AllReduceFn all_reduce = nullptr;
    GetErrorStringFn error_string = nullptr;
};

Nccl* nccl() {
    static Nccl n;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);   // PyTorch has it mapped already: same soname, same object
            if (n.handle) break;
        }
        if (n.handle) {
            n.get_unique_id = (GetUniqueIdFn)dlsym(n.handle, "ncclGetUniqueId");
            n.comm_init_rank = (CommInitRankFn)dlsym(n.handle, "ncclCommInitRank");
            n.comm_destroy = (CommDestroyFn)dlsym(n.handle, "ncclCommDestroy");
            n.all_reduce = (AllReduceFn)dlsym(n.handle, "ncclAllReduce");
            n.error_string = (GetErrorStringFn)dlsym(n.handle, "ncclGetErrorString");
        }
    }
    return (n.handle && n.get_unique_id && n.comm_init_rank && n.comm_destroy && n.all_reduce) ? &n : nullptr;
}

int fail(const char* who, int rc) {
    Nccl* n = nccl();
    uwr_set_error("%s: NCCL error %d (%s)", who, rc, (n && n->error_string) ? n->error_string(rc) : "?");
    return -4;
}

}  // namespace

extern "C" int uwr_nccl_available(void) { return nccl() != nullptr; }

extern "C" int uwr_nccl_unique_id(uwr_nccl_id* id) {
    Nccl* n = nccl();
    UWR_REQUIRE(n, "uwr_nccl_unique_id: libnccl.so.2 not found in this process");
    UWR_REQUIRE(id, "uwr_nccl_unique_id: null pointer");
    const int rc = n->get_unique_id(id);
    return rc ? fail("uwr_nccl_unique_id", rc) : 0;
}

extern "C" int uwr_nccl_comm_init(void** comm, int nranks, const uwr_nccl_id* id, int rank) {
    Nccl* n = nccl();
    UWR_REQUIRE(n, "uwr_nccl_comm_init: libnccl.so.2 not found in this process");
    UWR_REQUIRE(comm && id && nranks > 0 && rank >= 0 && rank < nranks, "uwr_nccl_comm_init: bad args");
    const int rc = n->comm_init_rank(comm, nranks, *id, rank);
    return rc ? fail("uwr_nccl_comm_init", rc) : 0;
}

extern "C" int uwr_nccl_comm_destroy(void* comm) {
    Nccl* n = nccl();
    if (!n || !comm) return 0;
    const int rc = n->comm_destroy(comm);
    return rc ? fail("uwr_nccl_comm_destroy", rc) : 0;
}

// in-place sum all-reduce of `count` floats on `stream` (capturable into a CUDA graph)
extern "C" int uwr_nccl_allreduce_sum_f32(void* comm, float* buf, size_t count, uwr_stream_t stream) {
    Nccl* n = nccl();
    UWR_REQUIRE(n, "uwr_nccl_allreduce_sum_f32: libnccl.so.2 not found in this process");
    UWR_REQUIRE(comm && buf, "uwr_nccl_allreduce_sum_f32: null pointer");
    if (count == 0) return 0;
    const int rc = n->all_reduce(buf, buf, count, /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm, (cudaStream_t)stream);
    return rc ? fail("uwr_nccl_allreduce_sum_f32", rc) : 0;
}
