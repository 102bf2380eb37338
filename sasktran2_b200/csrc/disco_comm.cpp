// Final gather of a wavelength-sharded solve over NCCL (NVLink 5 / NVSwitch): the only exchange step of the path
// (upstream shards the same way over host threads, cpp/lib/engine/engine.cpp:610-622).
// One process per GPU; every rank has solved its contiguous wavelength block and holds radiances / weighting
// functions on its device.  The non-root ranks ncclSend their blocks, the root ncclRecv's them into staging buffers
// and scatters them into the caller's full-spectrum host arrays with strided device -> host copies.
//
// NCCL is opened at run time (dlopen of libnccl.so.2 - the copy already loaded by the host process when there is one):
// a single-GPU caller needs no NCCL installation.  Only the handful of entry points below are used; their prototypes
// are restated from nccl.h (2.27 / 2.28: ncclUniqueId = 128 bytes, ncclDouble = 8).
#include "disco_comm.h"

#include <dlfcn.h>

#include <cstring>
#include <stdexcept>
#include <string>

namespace disco {

namespace {
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

NcclApi& api() {
    static NcclApi a;
    if (a.handle) return a;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) throw std::runtime_error(std::string("sasktran2_b200: NCCL is not available (dlopen libnccl.so.2): ") + dlerror());
    auto sym = [&](const char* s) {
        void* p = dlsym(a.handle, s);
        if (!p) throw std::runtime_error(std::string("sasktran2_b200: NCCL symbol missing: ") + s);
        return p;
    };
    a.GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
    a.CommInitRank = (int (*)(void**, int, NcclUniqueId, int))sym("ncclCommInitRank");
    a.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
    a.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))sym("ncclSend");
    a.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))sym("ncclRecv");
    a.GroupStart = (int (*)())sym("ncclGroupStart");
    a.GroupEnd = (int (*)())sym("ncclGroupEnd");
    a.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    return a;
}

void check(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + api().GetErrorString(rc));
}
constexpr int kNcclDouble = 8;
}  // namespace

void comm_unique_id(NcclUniqueId* id) { check(api().GetUniqueId(id), "ncclGetUniqueId"); }

Comm::Comm(const NcclUniqueId& id, int rank, int world) : m_rank(rank), m_world(world) {
    if (world < 1 || rank < 0 || rank >= world) throw std::runtime_error("sasktran2_b200: bad rank / world size");
    check(api().CommInitRank(&m_comm, world, id, rank), "ncclCommInitRank");
}
Comm::~Comm() {
    if (m_comm) api().CommDestroy(m_comm);
}
void Comm::group_start() { check(api().GroupStart(), "ncclGroupStart"); }
void Comm::group_end() { check(api().GroupEnd(), "ncclGroupEnd"); }
void Comm::send(const double* buf, size_t n, int peer, cudaStream_t s) {
    if (n) check(api().Send(buf, n, kNcclDouble, peer, m_comm, s), "ncclSend");
}
void Comm::recv(double* buf, size_t n, int peer, cudaStream_t s) {
    if (n) check(api().Recv(buf, n, kNcclDouble, peer, m_comm, s), "ncclRecv");
}

}  // namespace disco
