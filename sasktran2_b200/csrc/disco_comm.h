// NCCL communicator of a wavelength-sharded solve (one process per GPU); see disco_comm.cpp.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace disco {

struct NcclUniqueId {
    char internal[128];  // NCCL_UNIQUE_ID_BYTES
};

// rank 0 creates the id; the host program distributes the 128 bytes to the other ranks (any side channel)
void comm_unique_id(NcclUniqueId* id);

class Comm {
  public:
    Comm(const NcclUniqueId& id, int rank, int world);   // ncclCommInitRank on the current device
    ~Comm();
    Comm(const Comm&) = delete;
    Comm& operator=(const Comm&) = delete;
    int rank() const { return m_rank; }
    int world() const { return m_world; }
    void group_start();
    void group_end();
    void send(const double* buf, size_t n, int peer, cudaStream_t s);
    void recv(double* buf, size_t n, int peer, cudaStream_t s);

  private:
    void* m_comm = nullptr;
    int m_rank = 0, m_world = 1;
};

}  // namespace disco
