// lw_comm.cpp -- multi-GPU output gather behind the C ABI (the path's one exchange).
//
// One rank per GPU, started by the host's own launcher (MPI, torchrun, ...).  The column axis is block-partitioned over
// the ranks with no collective on the data path; what a host may want is the ranks' output blocks on one rank.  These
// entries give a Fortran / C host that step without Python: NCCL send / receive pairs grouped into one gather, enqueued
// on the caller's stream so that it orders after the rank's kernels.
//   rank 0:    rrtmg_lw_b200_comm_unique_id(id)          -> the host broadcasts the 128 bytes (e.g. MPI_Bcast)
//   all ranks: rrtmg_lw_b200_comm_init(nranks, rank, id)
//   per step:  rrtmg_lw_b200_gather(send, count, recv, root, stream)     recv: nranks x count doubles on the root
//              rrtmg_lw_b200_comm_finalize()
// NCCL is loaded with dlopen("libnccl.so.2") at the first call: the library neither links nor requires NCCL unless
// these entries are used.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstring>

#include "../../include/rrtmg_lw_b200.h"

namespace {

struct NcclId { char internal[128]; };  // ncclUniqueId
typedef void* NcclComm;
constexpr int kNcclFloat64 = 8;  // ncclDataType_t ncclFloat64 / ncclDouble

struct Nccl {
    void* h = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load() {
        if (h) return true;
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { std::fprintf(stderr, "rrtmg_lw_b200: cannot load libnccl.so.2 (%s)\n", dlerror()); return false; }
#define LW_SYM(field, name)                                                        \
    *reinterpret_cast<void**>(&field) = dlsym(h, name);                             \
    if (!field) { std::fprintf(stderr, "rrtmg_lw_b200: %s missing in libnccl\n", name); h = nullptr; return false; }
        LW_SYM(GetUniqueId, "ncclGetUniqueId")
        LW_SYM(CommInitRank, "ncclCommInitRank")
        LW_SYM(CommDestroy, "ncclCommDestroy")
        LW_SYM(Send, "ncclSend")
        LW_SYM(Recv, "ncclRecv")
        LW_SYM(GroupStart, "ncclGroupStart")
        LW_SYM(GroupEnd, "ncclGroupEnd")
        LW_SYM(GetErrorString, "ncclGetErrorString")
#undef LW_SYM
        return true;
    }
} N;

struct Comm {
    NcclComm comm = nullptr;
    int nranks = 0, rank = -1;
};
Comm g_comm[16];  // one per device, like the library's contexts

int cur_dev() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 16) d = 0;
    return d;
}

int fail(int rc, const char* where) {
    std::fprintf(stderr, "rrtmg_lw_b200: NCCL error in %s: %s\n", where, N.GetErrorString ? N.GetErrorString(rc) : "?");
    return RRTMG_LW_ERR_CUDA;
}

}  // namespace

extern "C" int rrtmg_lw_b200_comm_unique_id(void* id128) {
    if (!id128) return RRTMG_LW_ERR_ARG;
    if (!N.load()) return RRTMG_LW_ERR_CUDA;
    NcclId id;
    const int rc = N.GetUniqueId(&id);
    if (rc) return fail(rc, "ncclGetUniqueId");
    std::memcpy(id128, &id, sizeof(id));
    return RRTMG_LW_OK;
}

extern "C" int rrtmg_lw_b200_comm_init(int nranks, int rank, const void* id128) {
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return RRTMG_LW_ERR_ARG;
    if (!N.load()) return RRTMG_LW_ERR_CUDA;
    Comm& c = g_comm[cur_dev()];
    if (c.comm) { N.CommDestroy(c.comm); c = Comm{}; }
    NcclId id;
    std::memcpy(&id, id128, sizeof(id));
    const int rc = N.CommInitRank(&c.comm, nranks, id, rank);
    if (rc) { c = Comm{}; return fail(rc, "ncclCommInitRank"); }
    c.nranks = nranks; c.rank = rank;
    return RRTMG_LW_OK;
}

extern "C" int rrtmg_lw_b200_gather(const double* send, long count, double* recv, int root, void* stream) {
    Comm& c = g_comm[cur_dev()];
    if (!c.comm) return RRTMG_LW_ERR_NOT_INIT;
    if (!send || count < 0 || root < 0 || root >= c.nranks || (c.rank == root && !recv)) return RRTMG_LW_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = N.GroupStart();
    if (rc) return fail(rc, "ncclGroupStart");
    if (c.rank == root) {
        for (int r = 0; r < c.nranks && !rc; ++r) {
            if (r == root) continue;
            rc = N.Recv(recv + (size_t)r * count, (size_t)count, kNcclFloat64, r, c.comm, st);
        }
    } else {
        rc = N.Send(send, (size_t)count, kNcclFloat64, root, c.comm, st);
    }
    const int rc2 = N.GroupEnd();
    if (rc) return fail(rc, "ncclSend/ncclRecv");
    if (rc2) return fail(rc2, "ncclGroupEnd");
    if (c.rank == root && recv + (size_t)root * count != send) {  // the root's own block
        if (cudaMemcpyAsync(recv + (size_t)root * count, send, (size_t)count * sizeof(double), cudaMemcpyDeviceToDevice, st) !=
            cudaSuccess)
            return RRTMG_LW_ERR_CUDA;
    }
    return RRTMG_LW_OK;
}

extern "C" int rrtmg_lw_b200_comm_finalize(void) {
    Comm& c = g_comm[cur_dev()];
    if (c.comm && N.h) N.CommDestroy(c.comm);
    c = Comm{};
    return RRTMG_LW_OK;
}
