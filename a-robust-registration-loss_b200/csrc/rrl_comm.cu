// Peer-memory exchange between the ranks of one box (one process per GPU): the buffers behind CommView
// (rrl_common.cuh).  Each rank cudaMalloc's one buffer, publishes its CUDA IPC handle (the caller moves the 64-byte
// handles between the processes with whatever it has -- torch.distributed all_gather in dist.py), and maps every peer's
// buffer with cudaIpcOpenMemHandle (NVLink peer access).  After that the exchange itself is done entirely by kernels:
// rrl_shard_tail (rrl_sparse.cu: the line shard's counts / entries / partial sums) and rrl_comm_allreduce_f64 below (a
// handful of doubles, e.g. the pose-space gradient).  Ranks living in ONE process (tests; a single-process multi-GPU
// driver) connect with plain pointers instead of IPC handles.
#include <cstring>
#include <new>

#include "rrl_common.cuh"

struct rrl_comm {
    int rank, world, device;
    size_t slot_bytes, bytes;
    char *local;
    char *peer[rrl::kCommMaxWorld];
    bool ipc[rrl::kCommMaxWorld];
    bool connected;
};

namespace rrl {

CommView comm_view(const rrl_comm *c) {
    CommView v;
    v.rank = c->rank; v.world = c->world; v.slot_bytes = c->slot_bytes;
    for (int i = 0; i < kCommMaxWorld; ++i) v.peer[i] = i < c->world ? c->peer[i] : nullptr;
    return v;
}

// all-reduce(sum) of n <= kMaxAllreduce doubles in place: one CTA, one exchange, summed in rank order on every rank
// (same operands, same order: bit-identical results everywhere)
constexpr int kMaxAllreduce = 512;
__global__ void __launch_bounds__(256) comm_allreduce_kernel(CommView comm, double *buf, int n) {
    char *mine = comm.peer[comm.rank];
    const unsigned long long seq = *comm_epoch(mine) + 1ull;
    double *slot = reinterpret_cast<double *>(comm_slot(comm, mine, seq, comm.rank));
    const int npad = (n + 1) & ~1;                                   // 16-byte granules
    for (int i = threadIdx.x; i < npad; i += blockDim.x) slot[i] = i < n ? buf[i] : 0.0;
    __threadfence();
    __syncthreads();
    const bool ok = comm_exchange(comm, seq, (unsigned)npad * 8u);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double acc = 0.0;
        for (int r = 0; r < comm.world; ++r) acc += __ldcg(reinterpret_cast<const double *>(comm_slot(comm, mine, seq, r)) + i);
        buf[i] = ok ? acc : __longlong_as_double(0x7ff8000000000000LL);
    }
    if (threadIdx.x == 0) *comm_epoch(mine) = seq;
}

}  // namespace rrl

using namespace rrl;

extern "C" int rrl_comm_create(int rank, int world, size_t slot_bytes, rrl_comm **out) {
    if (!out || world < 1 || world > kCommMaxWorld || rank < 0 || rank >= world || slot_bytes == 0) return RRL_ERR_ARG;
    rrl_comm *c = new (std::nothrow) rrl_comm();
    if (!c) return RRL_ERR_CUDA;
    std::memset(c, 0, sizeof(*c));
    c->rank = rank; c->world = world;
    if (cudaGetDevice(&c->device) != cudaSuccess) { delete c; return RRL_ERR_CUDA; }
    c->slot_bytes = (slot_bytes + 255) / 256 * 256;
    c->bytes = kCommHeader + 2 * (size_t)world * c->slot_bytes;
    if (cudaMalloc(&c->local, c->bytes) != cudaSuccess) { delete c; return RRL_ERR_CUDA; }
    if (cudaMemset(c->local, 0, kCommHeader) != cudaSuccess) { cudaFree(c->local); delete c; return RRL_ERR_CUDA; }
    c->peer[rank] = c->local;
    c->connected = world == 1;
    *out = c;
    return RRL_OK;
}

extern "C" size_t rrl_comm_slot_bytes(const rrl_comm *c) { return c ? c->slot_bytes : 0; }
extern "C" void *rrl_comm_local_base(const rrl_comm *c) { return c ? c->local : nullptr; }

extern "C" int rrl_comm_ipc_handle(const rrl_comm *c, void *out_handle64) {
    if (!c || !out_handle64) return RRL_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the handle travels as 64 bytes");
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, c->local) != cudaSuccess) { cudaGetLastError(); return RRL_ERR_CUDA; }
    std::memcpy(out_handle64, &h, 64);
    return RRL_OK;
}

// handles64: world x 64 bytes, rank-major (every rank's rrl_comm_ipc_handle)
extern "C" int rrl_comm_connect_ipc(rrl_comm *c, const void *handles64) {
    if (!c || !handles64) return RRL_ERR_ARG;
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char *>(handles64) + (size_t)r * 64, 64);
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return RRL_ERR_CUDA; }
        c->peer[r] = static_cast<char *>(p);
        c->ipc[r] = true;
    }
    c->connected = true;
    return RRL_OK;
}

// same-process peers: bases[r] = rrl_comm_local_base of rank r's communicator (peer access between their devices must be
// enabled by the caller when they differ)
extern "C" int rrl_comm_connect_ptrs(rrl_comm *c, void *const *bases) {
    if (!c || !bases) return RRL_ERR_ARG;
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        if (!bases[r]) return RRL_ERR_ARG;
        c->peer[r] = static_cast<char *>(bases[r]);
    }
    c->connected = true;
    return RRL_OK;
}

extern "C" void rrl_comm_destroy(rrl_comm *c) {
    if (!c) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
        if (c->ipc[r] && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    cudaFree(c->local);
    if (prev >= 0) cudaSetDevice(prev);
    delete c;
}

// 0 = healthy; 1 = an exchange timed out on this rank (a peer never arrived).  Synchronises the device.
extern "C" int rrl_comm_error(const rrl_comm *c) {
    if (!c) return RRL_ERR_ARG;
    unsigned int e = 0;
    if (cudaMemcpy(&e, c->local + (kCommMaxWorld + 1) * 8, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return RRL_ERR_CUDA;
    return (int)e;
}

extern "C" int rrl_comm_allreduce_f64(rrl_comm *c, double *buf, int n, void *stream) {
    if (!c || !buf || n <= 0 || n > kMaxAllreduce || !c->connected) return RRL_ERR_ARG;
    if ((size_t)(n + 1) * 8 > c->slot_bytes) return RRL_ERR_WORKSPACE;
    Range r("rrl_comm_allreduce_f64");
    comm_allreduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(comm_view(c), buf, n);
    count_launch();
    return check_launch();
}

// Fused tail of the line-sharded forward (replaces rrl_shard_counts + the histogram rounds + rrl_shard_stage2/3 and the
// four collectives between them): see shard_tail_kernel in rrl_sparse.cu.  Call after rrl_shard_stage1(_ex) on the same
// workspace and stream, on EVERY rank of the communicator.  The communicator's slots must hold 160 + 64 * nl bytes.
extern "C" int rrl_shard_tail(void *workspace, size_t workspace_bytes, int nf1, int nf2, int nl, rrl_comm *c, float *out_loss,
                              int *out_status, float *out_median, long long *out_stats, void *stream) {
    if (!workspace || !c || !out_loss || nf1 <= 0 || nf2 <= 0 || nl <= 0 || !c->connected) return RRL_ERR_ARG;
    if (rrl_workspace_bytes(1, nf1, nf2, nl) == 0) return RRL_ERR_ARG;
    const Workspace ws = carve(workspace, 1, nf1, nf2, nl);
    if (workspace_bytes < ws.bytes) return RRL_ERR_WORKSPACE;
    if ((size_t)160 + (size_t)64 * nl > c->slot_bytes) return RRL_ERR_WORKSPACE;
    Range r("rrl_shard_tail (peer exchange, global median, Welsch, loss)");
    Geometry g;
    g.B = 1; g.nf1 = nf1; g.nf2 = nf2; g.nl = nl; g.nf1p = pad_points(nf1); g.nf2p = pad_points(nf2);
    return launch_shard_tail(ws, g, comm_view(c), out_loss, out_status, out_median, out_stats, (cudaStream_t)stream);
}
