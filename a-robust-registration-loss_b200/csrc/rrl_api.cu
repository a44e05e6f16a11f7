// extern "C" surface of librrl_b200.so (include/rrl_b200.h): argument checking, workspace carving, stage
// sequencing on the caller's stream, the line-shard stage API and the host-buffer convenience context.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "rrl_common.cuh"

namespace rrl {

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int check_launch() { return cudaGetLastError() == cudaSuccess ? RRL_OK : RRL_ERR_CUDA; }
int sm_count() {
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64) {
        const int c = cached[dev].load(std::memory_order_relaxed);
        if (c > 0) return c;
    }
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (dev >= 0 && dev < 64) cached[dev].store(n, std::memory_order_relaxed);
    return n;
}
void set_dense_variant(int v);

// ---- per-stage timing hook (rrl_measure_stages) ----
constexpr int kStages = 10;
static bool g_timing = false;
static cudaEvent_t g_ev[kStages + 1];
void stage_mark(int stage, cudaStream_t s) {
    if (g_timing && stage >= 0 && stage <= kStages) cudaEventRecord(g_ev[stage], s);
}
void set_param(int id, int v);

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

Workspace carve(void *base, int B, int nf1, int nf2, int nl) {
    Workspace w;
    char *p = reinterpret_cast<char *>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char *r = p + off;
        off = align_up(off + bytes);
        return r;
    };
    const size_t sB = (size_t)B, lines = (size_t)B * nl;
    const int nf1p = pad_points(nf1), nf2p = pad_points(nf2);
    w.hdr = reinterpret_cast<int *>(take(8 * sizeof(int)));
    w.keep = reinterpret_cast<unsigned int *>(take(sB * 8 * sizeof(unsigned int)));
    // ---- per-pair block, contiguous and zeroed by one memset in launch_prep (order matters) ----
    char *pair = take(16 + sB * (6 * 2 * 4 + 4 + 16 * 4 + 4 + 4 * 4 + 32 * 8 + RRL_NSTAT * 8 + 18 * 8));
    w.xcursor = reinterpret_cast<unsigned long long *>(pair);         pair += 16;
    w.pmax = reinterpret_cast<unsigned int *>(pair);                  pair += sB * 2 * 4;
    w.xmax = reinterpret_cast<unsigned int *>(pair);                  pair += sB * 2 * 4;
    w.rmax = reinterpret_cast<unsigned int *>(pair);                  pair += sB * 2 * 4;
    w.smax = reinterpret_cast<unsigned int *>(pair);                  pair += sB * 2 * 4;
    w.tmax = reinterpret_cast<unsigned int *>(pair);                  pair += sB * 2 * 4;
    w.bad = reinterpret_cast<unsigned int *>(pair);                   pair += sB * 2 * 4;
    w.nrec = reinterpret_cast<int *>(pair);                           pair += sB * 4;
    w.n_kj = reinterpret_cast<int *>(pair);                           pair += sB * 16 * 4;
    w.med = reinterpret_cast<float *>(pair);                          pair += sB * 4;
    w.flags = reinterpret_cast<int *>(pair);                          pair += sB * 4 * 4;
    w.sums = reinterpret_cast<unsigned long long *>(pair);            pair += sB * 32 * 8;
    w.stats = reinterpret_cast<long long *>(pair);                    pair += sB * RRL_NSTAT * 8;
    w.gcounts = reinterpret_cast<long long *>(pair);
    // ---- per triplet ----
    w.thr[0] = reinterpret_cast<float *>(take(sB * nf1 * sizeof(float)));
    w.thr[1] = reinterpret_cast<float *>(take(sB * nf2 * sizeof(float)));
    w.perm[0] = reinterpret_cast<int *>(take(sB * nf1p * sizeof(int)));
    w.perm[1] = reinterpret_cast<int *>(take(sB * nf2p * sizeof(int)));
    // triplet records: kNode + 1 float4 per node; node records: 5 float4 per group of 4 nodes (sized for kMinNode)
    w.pt4[0] = reinterpret_cast<float4 *>(take(sB * (nf1p / kMinNode) * (kMinNode + 1) * sizeof(float4)));
    w.pt4[1] = reinterpret_cast<float4 *>(take(sB * (nf2p / kMinNode) * (kMinNode + 1) * sizeof(float4)));
    w.pt12[0] = reinterpret_cast<float4 *>(take(sB * nf1p * 2 * sizeof(float4)));
    w.pt12[1] = reinterpret_cast<float4 *>(take(sB * nf2p * 2 * sizeof(float4)));
    w.node4[0] = reinterpret_cast<float4 *>(take(sB * (nf1p / kMinNode / 4) * 5 * sizeof(float4)));
    w.node4[1] = reinterpret_cast<float4 *>(take(sB * (nf2p / kMinNode / 4) * 5 * sizeof(float4)));
    w.super4[0] = reinterpret_cast<float4 *>(take(sB * (pad_supers(nf1p) / 4) * 5 * sizeof(float4)));
    w.super4[1] = reinterpret_cast<float4 *>(take(sB * (pad_supers(nf2p) / 4) * 5 * sizeof(float4)));
    w.sn8[0] = reinterpret_cast<uint4 *>(take(sB * pad_supers(nf1p) * 9 * sizeof(uint4)));
    w.sn8[1] = reinterpret_cast<uint4 *>(take(sB * pad_supers(nf2p) * 9 * sizeof(uint4)));
    w.sortbuf_bytes = sort_scratch_bytes(nf1p > nf2p ? nf1p : nf2p, B);
    w.sortbuf = reinterpret_cast<unsigned long long *>(take(w.sortbuf_bytes));
    // ---- per line ----
    w.lineC = reinterpret_cast<float4 *>(take(lines * 2 * sizeof(float4)));
    w.cnt[0] = reinterpret_cast<int *>(take(2 * lines * sizeof(int)));
    w.cnt[1] = w.cnt[0] + lines;
    w.hits[0] = reinterpret_cast<int *>(take(lines * kCap * sizeof(int)));
    w.hits[1] = reinterpret_cast<int *>(take(lines * kCap * sizeof(int)));
    w.xcap = (long long)lines * kExactPerLine;
    w.xcand = reinterpret_cast<uint2 *>(take((size_t)w.xcap * sizeof(uint2)));
    // ---- per record ----
    w.recD = reinterpret_cast<float *>(take(lines * 16 * sizeof(float)));
    w.dflat = reinterpret_cast<float *>(take(sB * kMedCache * sizeof(float)));
    w.lpart = reinterpret_cast<unsigned int *>(take(sB * 64 * 2 * sizeof(unsigned int)));
    w.recMeta = reinterpret_cast<int *>(take(lines * 2 * sizeof(int)));
    w.recIdx = reinterpret_cast<int *>(take(lines * 8 * sizeof(int)));
    w.recW = reinterpret_cast<float *>(take(lines * 24 * sizeof(float)));
    w.recQ = reinterpret_cast<float *>(take(lines * 24 * sizeof(float)));
    w.recG = reinterpret_cast<float *>(take(lines * 24 * sizeof(float)));
    w.bytes = off;
    return w;
}

static bool geometry_ok(int B, int nf1, int nf2, int nl) {
    return B > 0 && nf1 > 0 && nf2 > 0 && nl > 0 && B <= 32767 && (long long)B * nl <= (1LL << 30);
}
static bool window_ok(int k_lo, int j_lo, int k_hi, int j_hi) {
    return k_lo >= 1 && j_lo >= 1 && k_hi <= 5 && j_hi <= 5 && k_lo < k_hi && j_lo < j_hi;
}
static Geometry make_geometry(int B, int nf1, int nf2, int nl) {
    Geometry g;
    g.B = B; g.nf1 = nf1; g.nf2 = nf2; g.nl = nl;
    g.nf1p = pad_points(nf1); g.nf2p = pad_points(nf2);
    return g;
}

static int stage_dense_and_build(const float *tri1, const float *tri2, const float *lines, const Workspace &ws,
                                 const Geometry &g, int k_lo, int j_lo, int k_hi, int j_hi, int flags, cudaStream_t s) {
    const int window = k_lo | (j_lo << 8) | (k_hi << 16) | (j_hi << 24);
    int rc;
    {
        Range r("rrl.prep (thresholds, order, bounding spheres)");
        rc = launch_prep(tri1, tri2, lines, ws, g, window, flags, s);
    }
    if (rc) return rc;
    {
        Range r("rrl.dense (filtered point-line predicate + exact test)");
        rc = launch_dense(tri1, tri2, lines, ws, g, s);
    }
    if (rc) return rc;
    Range r("rrl.build (intersection points, D)");
    return launch_build(tri1, tri2, lines, ws, g, k_lo, j_lo, k_hi, j_hi, s);
}

}  // namespace rrl

using namespace rrl;

extern "C" int rrl_version(void) { return RRL_VERSION; }

extern "C" const char *rrl_error_string(int code) {
    switch (code) {
        case RRL_OK: return "ok";
        case RRL_ERR_ARG: return "invalid argument (null pointer, non-positive size, or hit-count window outside 1..4)";
        case RRL_ERR_WORKSPACE: return "workspace smaller than rrl_workspace_bytes()";
        case RRL_ERR_CUDA: return "CUDA runtime call or kernel launch failed";
        case RRL_ERR_STATE: return "workspace does not hold a forward pass for this geometry";
        default: return "unknown error";
    }
}

extern "C" long long rrl_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" size_t rrl_workspace_bytes(int B, int nf1, int nf2, int nl) {
    if (!geometry_ok(B, nf1, nf2, nl)) return 0;
    return carve(nullptr, B, nf1, nf2, nl).bytes;
}

extern "C" int rrl_loss_forward_ex(const float *tri1, const float *tri2, const float *lines, int B, int nf1, int nf2, int nl,
                                   int k_lo, int j_lo, int k_hi, int j_hi, void *workspace, size_t workspace_bytes,
                                   float *out_loss, int *out_status, float *out_median, long long *out_stats, int flags,
                                   void *stream) {
    if (!tri1 || !tri2 || !lines || !workspace || !out_loss) return RRL_ERR_ARG;
    if (!geometry_ok(B, nf1, nf2, nl) || !window_ok(k_lo, j_lo, k_hi, j_hi)) return RRL_ERR_ARG;
    if (flags & ~(RRL_REUSE_ORDER | RRL_REUSE_TARGET)) return RRL_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(workspace) % 256) return RRL_ERR_ARG;
    const Workspace ws = carve(workspace, B, nf1, nf2, nl);
    if (workspace_bytes < ws.bytes) return RRL_ERR_WORKSPACE;
    const Geometry g = make_geometry(B, nf1, nf2, nl);
    cudaStream_t s = (cudaStream_t)stream;
    Range fwd("rrl_loss_forward");
    int rc = stage_dense_and_build(tri1, tri2, lines, ws, g, k_lo, j_lo, k_hi, j_hi, flags, s);
    if (rc) return rc;
    Range r("rrl.tail (median, Welsch, loss)");
    return launch_tail(ws, g, out_loss, out_status, out_median, out_stats, s);
}

extern "C" int rrl_loss_forward(const float *tri1, const float *tri2, const float *lines, int B, int nf1, int nf2, int nl,
                                int k_lo, int j_lo, int k_hi, int j_hi, void *workspace, size_t workspace_bytes,
                                float *out_loss, int *out_status, float *out_median, long long *out_stats, void *stream) {
    return rrl_loss_forward_ex(tri1, tri2, lines, B, nf1, nf2, nl, k_lo, j_lo, k_hi, j_hi, workspace, workspace_bytes, out_loss,
                               out_status, out_median, out_stats, 0, stream);
}

extern "C" int rrl_loss_backward(const void *workspace, size_t workspace_bytes, const float *grad_out, int B, int nf1,
                                 int nf2, int nl, float *grad_tri1, float *grad_tri2, void *stream) {
    if (!workspace || !grad_out || !geometry_ok(B, nf1, nf2, nl)) return RRL_ERR_ARG;
    const Workspace ws = carve(const_cast<void *>(workspace), B, nf1, nf2, nl);
    if (workspace_bytes < ws.bytes) return RRL_ERR_WORKSPACE;
    if (!grad_tri1 && !grad_tri2) return RRL_OK;
    Range r("rrl_loss_backward");
    return launch_backward(ws, make_geometry(B, nf1, nf2, nl), grad_out, grad_tri1, grad_tri2, (cudaStream_t)stream);
}

extern "C" int rrl_se3_chain(const float *twist, const double *acc, int B, float *grad_twist, void *stream);

extern "C" int rrl_loss_backward_twist(const void *workspace, size_t workspace_bytes, const float *grad_out, int B, int nf1,
                                       int nf2, int nl, const float *twist, const float *raw_tri1, double *acc12,
                                       float *grad_twist, void *stream) {
    if (!workspace || !grad_out || !raw_tri1 || !acc12 || !geometry_ok(B, nf1, nf2, nl)) return RRL_ERR_ARG;
    if (grad_twist && !twist) return RRL_ERR_ARG;
    const Workspace ws = carve(const_cast<void *>(workspace), B, nf1, nf2, nl);
    if (workspace_bytes < ws.bytes) return RRL_ERR_WORKSPACE;
    Range r("rrl_loss_backward_twist");
    const int rc = launch_backward_pose(ws, make_geometry(B, nf1, nf2, nl), grad_out, raw_tri1, acc12, (cudaStream_t)stream);
    if (rc || !grad_twist) return rc;
    return rrl_se3_chain(twist, acc12, B, grad_twist, stream);
}

extern "C" int rrl_loss_export_hits(const void *workspace, size_t workspace_bytes, int B, int nf1, int nf2, int nl,
                                    int cloud, int *out_counts, int *out_hits, void *stream) {
    if (!workspace || !out_counts || !out_hits || !geometry_ok(B, nf1, nf2, nl) || (cloud != 1 && cloud != 2)) return RRL_ERR_ARG;
    const Workspace ws = carve(const_cast<void *>(workspace), B, nf1, nf2, nl);
    if (workspace_bytes < ws.bytes) return RRL_ERR_WORKSPACE;
    return launch_export_hits(ws, make_geometry(B, nf1, nf2, nl), cloud - 1, out_counts, out_hits, (cudaStream_t)stream);
}

// ---- line-shard stages (B = 1) -------------------------------------------------------------------------
extern "C" int rrl_shard_stage1_ex(const float *tri1, const float *tri2, const float *lines, int nf1, int nf2, int nl,
                                   int k_lo, int j_lo, int k_hi, int j_hi, void *workspace, size_t workspace_bytes, int flags,
                                   void *stream) {
    if (!tri1 || !tri2 || !lines || !workspace) return RRL_ERR_ARG;
    if (!geometry_ok(1, nf1, nf2, nl) || !window_ok(k_lo, j_lo, k_hi, j_hi)) return RRL_ERR_ARG;
    if (flags & ~(RRL_REUSE_ORDER | RRL_REUSE_TARGET)) return RRL_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(workspace) % 256) return RRL_ERR_ARG;
    const Workspace ws = carve(workspace, 1, nf1, nf2, nl);
    if (workspace_bytes < ws.bytes) return RRL_ERR_WORKSPACE;
    Range r("rrl_shard_stage1");
    return stage_dense_and_build(tri1, tri2, lines, ws, make_geometry(1, nf1, nf2, nl), k_lo, j_lo, k_hi, j_hi, flags,
                                 (cudaStream_t)stream);
}

extern "C" int rrl_shard_stage1(const float *tri1, const float *tri2, const float *lines, int nf1, int nf2, int nl,
                                int k_lo, int j_lo, int k_hi, int j_hi, void *workspace, size_t workspace_bytes, void *stream) {
    return rrl_shard_stage1_ex(tri1, tri2, lines, nf1, nf2, nl, k_lo, j_lo, k_hi, j_hi, workspace, workspace_bytes, 0, stream);
}

extern "C" int rrl_shard_counts(void *workspace, size_t workspace_bytes, int nf1, int nf2, int nl, long long *counts18, void *stream) {
    if (!workspace || !counts18 || !geometry_ok(1, nf1, nf2, nl)) return RRL_ERR_ARG;
    const Workspace ws = carve(workspace, 1, nf1, nf2, nl);
    if (workspace_bytes < ws.bytes) return RRL_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const int rc = launch_local_counts(ws, make_geometry(1, nf1, nf2, nl), s);
    if (rc) return rc;
    return cudaMemcpyAsync(counts18, ws.gcounts, 18 * sizeof(long long), cudaMemcpyDeviceToDevice, s) == cudaSuccess ? RRL_OK : RRL_ERR_CUDA;
}

extern "C" int rrl_shard_pack_entries(void *workspace, size_t workspace_bytes, int nf1, int nf2, int nl, float *out_entries,
                                      long long capacity, void *stream) {
    if (!workspace || !out_entries || capacity < 0 || !geometry_ok(1, nf1, nf2, nl)) return RRL_ERR_ARG;
    const Workspace ws = carve(workspace, 1, nf1, nf2, nl);
    if (workspace_bytes < ws.bytes) return RRL_ERR_WORKSPACE;
    return launch_pack_entries(ws, make_geometry(1, nf1, nf2, nl), out_entries, capacity, (cudaStream_t)stream);
}

extern "C" int rrl_select_lower_median(const float *values, long long n, float *out_median, void *stream) {
    if (!out_median || n < 0 || (n > 0 && !values)) return RRL_ERR_ARG;
    return launch_select_median(values, n, out_median, (cudaStream_t)stream);
}

extern "C" int rrl_shard_select_hist(void *workspace, size_t workspace_bytes, int nf1, int nf2, int nl, int round,
                                     const long long *state2, int *out_hist65536, void *stream) {
    if (!workspace || !state2 || !out_hist65536 || (round != 0 && round != 1) || !geometry_ok(1, nf1, nf2, nl)) return RRL_ERR_ARG;
    const Workspace ws = carve(workspace, 1, nf1, nf2, nl);
    if (workspace_bytes < ws.bytes) return RRL_ERR_WORKSPACE;
    return launch_shard_hist(ws, make_geometry(1, nf1, nf2, nl), round, state2, out_hist65536, (cudaStream_t)stream);
}

extern "C" int rrl_shard_select_pick(int round, const int *global_hist65536, const long long *global_counts18, long long *state2,
                                     float *out_median, void *stream) {
    if (!global_hist65536 || !global_counts18 || !state2 || !out_median || (round != 0 && round != 1)) return RRL_ERR_ARG;
    return launch_shard_pick(round, global_hist65536, global_counts18, state2, out_median, (cudaStream_t)stream);
}

extern "C" int rrl_shard_stage2(void *workspace, size_t workspace_bytes, int nf1, int nf2, int nl,
                                const long long *global_counts18, const float *global_median, long long *sums32, void *stream) {
    if (!workspace || !global_counts18 || !global_median || !sums32 || !geometry_ok(1, nf1, nf2, nl)) return RRL_ERR_ARG;
    const Workspace ws = carve(workspace, 1, nf1, nf2, nl);
    if (workspace_bytes < ws.bytes) return RRL_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemcpyAsync(ws.gcounts, global_counts18, 18 * sizeof(long long), cudaMemcpyDeviceToDevice, s) != cudaSuccess) return RRL_ERR_CUDA;
    if (cudaMemcpyAsync(ws.med, global_median, sizeof(float), cudaMemcpyDeviceToDevice, s) != cudaSuccess) return RRL_ERR_CUDA;
    if (cudaMemsetAsync(ws.sums, 0, 32 * sizeof(unsigned long long), s) != cudaSuccess) return RRL_ERR_CUDA;
    const int rc = launch_welsch(ws, make_geometry(1, nf1, nf2, nl), s);
    if (rc) return rc;
    return cudaMemcpyAsync(sums32, ws.sums, 32 * sizeof(long long), cudaMemcpyDeviceToDevice, s) == cudaSuccess ? RRL_OK : RRL_ERR_CUDA;
}

extern "C" int rrl_shard_stage3(void *workspace, size_t workspace_bytes, int nf1, int nf2, int nl,
                                const long long *global_sums32, float *out_loss, int *out_status, void *stream) {
    if (!workspace || !global_sums32 || !out_loss || !geometry_ok(1, nf1, nf2, nl)) return RRL_ERR_ARG;
    const Workspace ws = carve(workspace, 1, nf1, nf2, nl);
    if (workspace_bytes < ws.bytes) return RRL_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemcpyAsync(ws.sums, global_sums32, 32 * sizeof(long long), cudaMemcpyDeviceToDevice, s) != cudaSuccess) return RRL_ERR_CUDA;
    return launch_finalize(ws, make_geometry(1, nf1, nf2, nl), out_loss, out_status, nullptr, nullptr, s);
}

// ---- host-buffer context ---------------------------------------------------------------------------------
// Two levels of overlap.  (1) The batch is cut into up to kMaxSub sub-batches of whole pairs (pairs are independent),
// each with its own stream and workspace: the H2D copy of sub-batch s+1 runs under the kernels of sub-batch s, and the
// latency-bound sparse stages of one sub-batch run under the dense stage of the next.  (2) The context owns kSlots
// complete sets of device buffers: rrl_host_submit() queues a whole evaluation on the next slot and returns, so the
// copies of evaluation i+1 run under the kernels of evaluation i (a double-buffered input pipeline);
// rrl_host_wait() drains one slot and hands out its results.  rrl_host_loss_fwd_bwd = submit + wait.
constexpr int kMaxSub = 8;
constexpr int kSlots = 2;
// the host-buffer entry points work on the context's device and hand the caller's current device back on return
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = (prev == device) || cudaSetDevice(device) == cudaSuccess;
        if (prev == device) prev = -1;                       // nothing to restore
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
struct HostSlot {
    float *d_tri1, *d_tri2, *d_lines, *d_loss, *d_grad1;
    int *d_status;
    void *d_ws[kMaxSub];
    float *p_loss, *p_grad1;
    int *p_status;
    cudaStream_t stream[kMaxSub];
    bool busy, want_grad;
};
struct rrl_host_ctx {
    int B, nf1, nf2, nl, device, S, next;
    int first[kMaxSub + 1];                      // pairs [first[s], first[s+1]) form sub-batch s
    size_t n_tri1, n_tri2, n_lines;
    size_t ws_bytes[kMaxSub];
    float *d_gout;
    float *p_tri1, *p_tri2, *p_lines;            // optional staging handed out to the caller
    HostSlot slot[kSlots];
};

static int host_subbatches(int B) {
    int S = B / 16;                              // >= 16 pairs per sub-batch keep the grids full (measured: 2 x 16 beats 4 x 8 and 1 x 32 at B = 32)
    if (const char *e = getenv("RRL_HOST_SUBBATCHES")) S = atoi(e);
    if (S > kMaxSub) S = kMaxSub;
    if (S > B) S = B;
    if (S < 1) S = 1;
    return S;
}

extern "C" void rrl_host_destroy(rrl_host_ctx *c) {
    if (!c) return;
    DeviceGuard guard(c->device);
    for (int q = 0; q < kSlots; ++q) {
        HostSlot &t = c->slot[q];
        for (int s = 0; s < c->S; ++s)
            if (t.stream[s]) cudaStreamSynchronize(t.stream[s]);
        cudaFree(t.d_tri1); cudaFree(t.d_tri2); cudaFree(t.d_lines); cudaFree(t.d_loss); cudaFree(t.d_grad1); cudaFree(t.d_status);
        for (int s = 0; s < c->S; ++s) cudaFree(t.d_ws[s]);
        cudaFreeHost(t.p_loss); cudaFreeHost(t.p_grad1); cudaFreeHost(t.p_status);
        for (int s = 0; s < c->S; ++s)
            if (t.stream[s]) cudaStreamDestroy(t.stream[s]);
    }
    cudaFree(c->d_gout);
    cudaFreeHost(c->p_tri1); cudaFreeHost(c->p_tri2); cudaFreeHost(c->p_lines);
    delete c;
}

extern "C" int rrl_host_create(int B, int nf1, int nf2, int nl, int device, rrl_host_ctx **out_ctx) {
    if (!out_ctx || !geometry_ok(B, nf1, nf2, nl)) return RRL_ERR_ARG;
    DeviceGuard guard(device);
    if (!guard.ok) return RRL_ERR_CUDA;
    rrl_host_ctx *c = new (std::nothrow) rrl_host_ctx();
    if (!c) return RRL_ERR_CUDA;
    std::memset(c, 0, sizeof(*c));
    c->B = B; c->nf1 = nf1; c->nf2 = nf2; c->nl = nl; c->device = device;
    c->S = host_subbatches(B);
    for (int s = 0; s <= c->S; ++s) c->first[s] = (int)((long long)B * s / c->S);
    c->n_tri1 = (size_t)B * nf1 * 9; c->n_tri2 = (size_t)B * nf2 * 9; c->n_lines = (size_t)B * nl * 6;
    bool ok = true;
    for (int s = 0; s < c->S; ++s) c->ws_bytes[s] = rrl_workspace_bytes(c->first[s + 1] - c->first[s], nf1, nf2, nl);
    for (int q = 0; q < kSlots && ok; ++q) {
        HostSlot &t = c->slot[q];
        for (int s = 0; s < c->S && ok; ++s)
            ok = cudaStreamCreateWithFlags(&t.stream[s], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaMalloc(&t.d_ws[s], c->ws_bytes[s]) == cudaSuccess;
        ok = ok && cudaMalloc(&t.d_tri1, c->n_tri1 * 4) == cudaSuccess && cudaMalloc(&t.d_tri2, c->n_tri2 * 4) == cudaSuccess;
        ok = ok && cudaMalloc(&t.d_lines, c->n_lines * 4) == cudaSuccess && cudaMalloc(&t.d_loss, (size_t)B * 4) == cudaSuccess;
        ok = ok && cudaMalloc(&t.d_grad1, c->n_tri1 * 4) == cudaSuccess && cudaMalloc(&t.d_status, (size_t)B * 4) == cudaSuccess;
        ok = ok && cudaMallocHost(&t.p_loss, (size_t)B * 4) == cudaSuccess && cudaMallocHost(&t.p_status, (size_t)B * 4) == cudaSuccess;
        ok = ok && cudaMallocHost(&t.p_grad1, c->n_tri1 * 4) == cudaSuccess;
    }
    ok = ok && cudaMalloc(&c->d_gout, (size_t)B * 4) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->p_tri1, c->n_tri1 * 4) == cudaSuccess && cudaMallocHost(&c->p_tri2, c->n_tri2 * 4) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->p_lines, c->n_lines * 4) == cudaSuccess;
    if (ok) {
        float *ones = c->slot[0].p_loss;
        for (int b = 0; b < B; ++b) ones[b] = 1.0f;
        ok = cudaMemcpy(c->d_gout, ones, (size_t)B * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (!ok) {
        rrl_host_destroy(c);
        return RRL_ERR_CUDA;
    }
    *out_ctx = c;
    return RRL_OK;
}

extern "C" float *rrl_host_pinned_tri1(rrl_host_ctx *c) { return c ? c->p_tri1 : nullptr; }
extern "C" float *rrl_host_pinned_tri2(rrl_host_ctx *c) { return c ? c->p_tri2 : nullptr; }
extern "C" float *rrl_host_pinned_lines(rrl_host_ctx *c) { return c ? c->p_lines : nullptr; }
extern "C" int rrl_host_subbatches(rrl_host_ctx *c) { return c ? c->S : 0; }
extern "C" int rrl_host_slots(rrl_host_ctx *c) { return c ? kSlots : 0; }

extern "C" int rrl_host_submit(rrl_host_ctx *c, const float *h_tri1, const float *h_tri2, const float *h_lines,
                               int k_lo, int j_lo, int k_hi, int j_hi, int want_grad_tri1, int *out_ticket) {
    if (!c || !h_tri1 || !h_tri2 || !h_lines || !out_ticket) return RRL_ERR_ARG;
    if (!window_ok(k_lo, j_lo, k_hi, j_hi)) return RRL_ERR_ARG;
    HostSlot &t = c->slot[c->next];
    if (t.busy) return RRL_ERR_STATE;                  // every slot in flight: rrl_host_wait() one first
    DeviceGuard guard(c->device);
    if (!guard.ok) return RRL_ERR_CUDA;
    const size_t t1 = (size_t)c->nf1 * 9, t2 = (size_t)c->nf2 * 9, tl = (size_t)c->nl * 6;
    // straight from the caller's memory: asynchronous when it is pinned (the context's own buffers or any
    // cudaHostAlloc/cudaHostRegister'ed range), staged by the driver when it is pageable.  All copies are queued
    // first, in sub-batch order, so that the copy engine never waits for the host to get through the launches.
    bool ok = true;
    for (int s = 0; s < c->S && ok; ++s) {
        const size_t b0 = (size_t)c->first[s], nb = (size_t)(c->first[s + 1] - c->first[s]);
        cudaStream_t st = t.stream[s];
        ok = cudaMemcpyAsync(t.d_tri1 + b0 * t1, h_tri1 + b0 * t1, nb * t1 * 4, cudaMemcpyHostToDevice, st) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(t.d_tri2 + b0 * t2, h_tri2 + b0 * t2, nb * t2 * 4, cudaMemcpyHostToDevice, st) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(t.d_lines + b0 * tl, h_lines + b0 * tl, nb * tl * 4, cudaMemcpyHostToDevice, st) == cudaSuccess;
    }
    if (!ok) return RRL_ERR_CUDA;
    t.busy = true;                                     // work is queued from here on: the slot must be drained
    t.want_grad = want_grad_tri1 != 0;
    *out_ticket = c->next;
    c->next = (c->next + 1) % kSlots;
    for (int s = 0; s < c->S; ++s) {
        const size_t b0 = (size_t)c->first[s];
        const int nb = c->first[s + 1] - c->first[s];
        cudaStream_t st = t.stream[s];
        int rc = rrl_loss_forward(t.d_tri1 + b0 * t1, t.d_tri2 + b0 * t2, t.d_lines + b0 * tl, nb, c->nf1, c->nf2, c->nl,
                                  k_lo, j_lo, k_hi, j_hi, t.d_ws[s], c->ws_bytes[s], t.d_loss + b0, t.d_status + b0, nullptr,
                                  nullptr, st);
        if (rc) return rc;
        rc = rrl_loss_backward(t.d_ws[s], c->ws_bytes[s], c->d_gout + b0, nb, c->nf1, c->nf2, c->nl, t.d_grad1 + b0 * t1,
                               nullptr, st);
        if (rc) return rc;
        ok = cudaMemcpyAsync(t.p_loss + b0, t.d_loss + b0, (size_t)nb * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(t.p_status + b0, t.d_status + b0, (size_t)nb * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        if (t.want_grad)
            ok = ok && cudaMemcpyAsync(t.p_grad1 + b0 * t1, t.d_grad1 + b0 * t1, (size_t)nb * t1 * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        if (!ok) return RRL_ERR_CUDA;
    }
    return RRL_OK;
}

extern "C" int rrl_host_wait(rrl_host_ctx *c, int ticket, float *h_loss, int *h_status, float *h_grad_tri1) {
    if (!c || ticket < 0 || ticket >= kSlots || !h_loss) return RRL_ERR_ARG;
    HostSlot &t = c->slot[ticket];
    if (!t.busy) return RRL_ERR_STATE;
    if (h_grad_tri1 && !t.want_grad) return RRL_ERR_ARG;
    DeviceGuard guard(c->device);
    if (!guard.ok) return RRL_ERR_CUDA;
    bool ok = true;
    for (int s = 0; s < c->S; ++s) ok = (cudaStreamSynchronize(t.stream[s]) == cudaSuccess) && ok;
    t.busy = false;
    if (!ok) return RRL_ERR_CUDA;
    std::memcpy(h_loss, t.p_loss, (size_t)c->B * 4);
    if (h_status) std::memcpy(h_status, t.p_status, (size_t)c->B * 4);
    if (h_grad_tri1) std::memcpy(h_grad_tri1, t.p_grad1, c->n_tri1 * 4);
    return RRL_OK;
}

extern "C" int rrl_host_loss_fwd_bwd(rrl_host_ctx *c, const float *h_tri1, const float *h_tri2, const float *h_lines,
                                     int k_lo, int j_lo, int k_hi, int j_hi, float *h_loss, int *h_status, float *h_grad_tri1) {
    if (!h_loss) return RRL_ERR_ARG;
    int ticket = -1;
    const int rc = rrl_host_submit(c, h_tri1, h_tri2, h_lines, k_lo, j_lo, k_hi, j_hi, h_grad_tri1 != nullptr, &ticket);
    if (rc) {
        if (ticket >= 0) rrl_host_wait(c, ticket, h_loss, nullptr, nullptr);
        return rc;
    }
    return rrl_host_wait(c, ticket, h_loss, h_status, h_grad_tri1);
}

// ---- measurement ---------------------------------------------------------------------------------------------
extern "C" int rrl_measure_dense(const float *tri1, const float *tri2, const float *lines, int B, int nf1, int nf2, int nl,
                                 void *workspace, size_t workspace_bytes, int iters, float *out_ms_dense, float *out_ms_prep,
                                 void *stream) {
    if (!tri1 || !tri2 || !lines || !workspace || !out_ms_dense || iters <= 0 || !geometry_ok(B, nf1, nf2, nl)) return RRL_ERR_ARG;
    const Workspace ws = carve(workspace, B, nf1, nf2, nl);
    if (workspace_bytes < ws.bytes) return RRL_ERR_WORKSPACE;
    const Geometry g = make_geometry(B, nf1, nf2, nl);
    cudaStream_t s = (cudaStream_t)stream;
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    float td = 0.f, tp = 0.f;
    int rc = RRL_OK;
    for (int it = 0; it < iters + 1 && rc == RRL_OK; ++it) {
        cudaEventRecord(e0, s);
        rc = launch_prep(tri1, tri2, lines, ws, g, 1 | (1 << 8) | (5 << 16) | (5 << 24), 0, s);
        cudaEventRecord(e1, s);
        if (rc == RRL_OK) rc = launch_dense(tri1, tri2, lines, ws, g, s);
        cudaEventRecord(e2, s);
        cudaEventSynchronize(e2);
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, e0, e1);
        cudaEventElapsedTime(&b, e1, e2);
        if (it > 0) { tp += a; td += b; }          // first iteration is warm-up
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    *out_ms_dense = td / iters;
    if (out_ms_prep) *out_ms_prep = tp / iters;
    return rc;
}

// selects the dense-kernel variant (1 = packed FFMA2 [default], 0 = scalar FFMA); measurement only
extern "C" int rrl_debug_set_dense_variant(int v) {
    set_dense_variant(v);
    return RRL_OK;
}

// tuning knobs for A/B measurements: 1 = node size override (0 auto, 8, 16), 2 = target waves of CTAs, 3 = min nodes per chunk
extern "C" int rrl_debug_set_param(int id, int value) {
    set_param(id, value);
    return RRL_OK;
}

// Per-stage device times (ms, CUDA events, averaged over `iters` hot repetitions) of forward + backward:
// out_ms[0..9] = {memset+prep, sort, node, dense, select, build, median, welsch+finalize, backward, total}
extern "C" int rrl_measure_stages(const float *tri1, const float *tri2, const float *lines, int B, int nf1, int nf2, int nl,
                                  void *workspace, size_t workspace_bytes, int iters, float *out_ms, void *stream) {
    if (!tri1 || !tri2 || !lines || !workspace || !out_ms || iters <= 0 || !geometry_ok(B, nf1, nf2, nl)) return RRL_ERR_ARG;
    if (workspace_bytes < rrl_workspace_bytes(B, nf1, nf2, nl)) return RRL_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    float *loss = nullptr, *gout = nullptr, *grad = nullptr;
    int *status = nullptr;
    if (cudaMalloc(&loss, 4 * (size_t)B) != cudaSuccess || cudaMalloc(&gout, 4 * (size_t)B) != cudaSuccess ||
        cudaMalloc(&status, 4 * (size_t)B) != cudaSuccess || cudaMalloc(&grad, 36 * (size_t)B * nf1) != cudaSuccess)
        return RRL_ERR_CUDA;
    cudaMemset(gout, 0, 4 * (size_t)B);
    for (int i = 0; i <= kStages; ++i) cudaEventCreate(&g_ev[i]);
    double acc[kStages] = {0};
    int rc = RRL_OK;
    for (int it = 0; it < iters + 2 && rc == RRL_OK; ++it) {
        g_timing = true;
        stage_mark(0, s);
        rc = rrl_loss_forward(tri1, tri2, lines, B, nf1, nf2, nl, 1, 1, 5, 5, workspace, workspace_bytes, loss, status, nullptr, nullptr, s);
        if (rc == RRL_OK) rc = rrl_loss_backward(workspace, workspace_bytes, gout, B, nf1, nf2, nl, grad, nullptr, s);
        stage_mark(10, s);
        g_timing = false;
        cudaEventSynchronize(g_ev[10]);
        if (it >= 2) {
            for (int i = 0; i < 9; ++i) {
                float ms = 0;
                cudaEventElapsedTime(&ms, g_ev[i], g_ev[i + 1]);
                acc[i] += ms;
            }
            float ms = 0;
            cudaEventElapsedTime(&ms, g_ev[0], g_ev[10]);
            acc[9] += ms;
        }
    }
    for (int i = 0; i < kStages; ++i) out_ms[i] = (float)(acc[i] / iters);
    for (int i = 0; i <= kStages; ++i) cudaEventDestroy(g_ev[i]);
    cudaFree(loss); cudaFree(gout); cudaFree(status); cudaFree(grad);
    return rc;
}
