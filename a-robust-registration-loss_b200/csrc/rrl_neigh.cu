// Pre-processing that feeds the loss: Sample_neighs = farthest-point sampling + k nearest neighbours
// (/root/reference/code/loss.py:473-485, utils.py:275-296,380-385).
//
//   fps_kernel   the reference recurrence, literally: distance[i] = min(distance[i], ((dx^2 + dy^2) + dz^2)) against the
//                latest centroid, next centroid = FIRST index of the maximum (torch.max).  Products and sums are rounded
//                one by one, so the selected indices equal the reference's on identical float32 input bits (the demo
//                casts igl's vertices to float32 first, test_demo_optimized_Lie_Algebra.py:114-115).  A float64 cloud
//                is accepted too: distances in double, the running minimum still the float32 array of utils.py:287
//                (compared in double, rounded when stored) -- the reference itself raises on that input.  One
//                cooperative launch: every block keeps a strided share of the running distances, one grid barrier per
//                sample; the per-block (max, index) candidates are double buffered so that a single barrier suffices.
//   knn_kernel   exact brute-force k nearest neighbours of the sampled points among all points: one warp per query, each
//                lane keeps the k best of its strided share, the warp merges them.  Squared distances in double (what
//                sklearn's KDTree computes on the float64-converted cloud), ties broken by the smaller index.
#include <cooperative_groups.h>

#include "rrl_common.cuh"

namespace cg = cooperative_groups;

namespace rrl {

constexpr int kFpsThreads = 1024;
constexpr int kKnnMax = 8;

template <typename T> __device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }

// candidate = (distance, index); `better` = larger distance, or equal distance and smaller index (first maximum)
__device__ __forceinline__ void take_better(float &d, int &i, float od, int oi) {
    if (od > d || (od == d && oi < i)) { d = od; i = oi; }
}

template <typename T>
__global__ void __launch_bounds__(kFpsThreads) fps_kernel(const T *__restrict__ xyz, int N, int npoint, int start, int *__restrict__ out_idx,
                                                          float *__restrict__ dist, float *cand_d /*[2][grid]*/, int *cand_i /*[2][grid]*/) {
    cg::grid_group grid = cg::this_grid();
    __shared__ float s_d[32];
    __shared__ int s_i[32];
    __shared__ int s_far;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int gtid = blockIdx.x * kFpsThreads + tid, gstride = gridDim.x * kFpsThreads;
    for (int i = gtid; i < N; i += gstride) dist[i] = 1e10f;                    // utils.py:287 (float32)
    int far = start;
    for (int it = 0; it < npoint; ++it) {
        if (gtid == 0) out_idx[it] = far;
        const T cx = xyz[3 * (long long)far], cy = xyz[3 * (long long)far + 1], cz = xyz[3 * (long long)far + 2];
        float best = -1.f;
        int best_i = 0x7fffffff;
        for (int i = gtid; i < N; i += gstride) {
            const T dx = xyz[3 * (long long)i] - cx, dy = xyz[3 * (long long)i + 1] - cy, dz = xyz[3 * (long long)i + 2] - cz;
            const T d = add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz));   // torch.sum over 3: ((a+b)+c)
            float cur = dist[i];
            if (d < (T)cur) { cur = (float)d; dist[i] = cur; }                   // utils.py:293-294
            take_better(best, best_i, cur, i);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            take_better(best, best_i, od, oi);
        }
        if (lane == 0) { s_d[wid] = best; s_i[wid] = best_i; }
        __syncthreads();
        if (wid == 0) {
            best = s_d[lane]; best_i = s_i[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
                take_better(best, best_i, od, oi);
            }
            if (lane == 0) {
                cand_d[(it & 1) * gridDim.x + blockIdx.x] = best;
                cand_i[(it & 1) * gridDim.x + blockIdx.x] = best_i;
            }
        }
        grid.sync();
        // every block reduces the per-block candidates itself (<= a few hundred entries)
        if (wid == 0) {
            best = -1.f; best_i = 0x7fffffff;
            for (int q = lane; q < (int)gridDim.x; q += 32)
                take_better(best, best_i, cand_d[(it & 1) * gridDim.x + q], cand_i[(it & 1) * gridDim.x + q]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
                take_better(best, best_i, od, oi);
            }
            if (lane == 0) s_far = best_i;
        }
        __syncthreads();
        far = s_far;
    }
}

// insertion of (d, i) into an ascending list of k entries (ties: smaller index first)
template <int K>
__device__ __forceinline__ void knn_insert(double (&bd)[K], int (&bi)[K], double d, int i) {
    if (d > bd[K - 1] || (d == bd[K - 1] && i >= bi[K - 1])) return;
    bd[K - 1] = d; bi[K - 1] = i;
#pragma unroll
    for (int q = K - 1; q > 0; --q) {
        if (bd[q] < bd[q - 1] || (bd[q] == bd[q - 1] && bi[q] < bi[q - 1])) {
            const double td = bd[q]; bd[q] = bd[q - 1]; bd[q - 1] = td;
            const int ti = bi[q]; bi[q] = bi[q - 1]; bi[q - 1] = ti;
        }
    }
}

template <typename T, int K>
__global__ void __launch_bounds__(256) knn_kernel(const T *__restrict__ xyz, int N, const int *__restrict__ query_idx, int M, int *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= M) return;
    const long long qi = query_idx[q];
    const double qx = (double)xyz[3 * qi], qy = (double)xyz[3 * qi + 1], qz = (double)xyz[3 * qi + 2];
    double bd[K];
    int bi[K];
#pragma unroll
    for (int s = 0; s < K; ++s) { bd[s] = 1e300; bi[s] = 0x7fffffff; }
    for (int i = lane; i < N; i += 32) {
        const double dx = (double)xyz[3 * (long long)i] - qx, dy = (double)xyz[3 * (long long)i + 1] - qy, dz = (double)xyz[3 * (long long)i + 2] - qz;
        knn_insert<K>(bd, bi, dx * dx + dy * dy + dz * dz, i);
    }
    // K rounds of "warp minimum of the lanes' current heads"; the winning lane pops its head
#pragma unroll
    for (int s = 0; s < K; ++s) {
        double d = bd[0];
        int i = bi[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, d, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (od < d || (od == d && oi < i)) { d = od; i = oi; }
        }
        if (lane == 0) out[(long long)q * K + s] = i;
        if (bi[0] == i && bd[0] == d) {                 // indices are unique across lanes: exactly one lane pops
#pragma unroll
            for (int t = 0; t < K - 1; ++t) { bd[t] = bd[t + 1]; bi[t] = bi[t + 1]; }
            bd[K - 1] = 1e300; bi[K - 1] = 0x7fffffff;
        }
    }
}

template <typename T>
static int launch_fps(const void *xyz, int N, int npoint, int start, int *out_idx, void *scratch, cudaStream_t s) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fps_kernel<T>, kFpsThreads, 0) != cudaSuccess || per_sm < 1) return RRL_ERR_CUDA;
    int blocks = (N + kFpsThreads * 4 - 1) / (kFpsThreads * 4);                  // >= 4 points per thread before another block pays
    if (blocks > sms * per_sm) blocks = sms * per_sm;
    if (blocks < 1) blocks = 1;
    const T *x = reinterpret_cast<const T *>(xyz);
    float *dist = reinterpret_cast<float *>(scratch);
    float *cand_d = dist + (((size_t)N + 63) / 64) * 64;
    int *cand_i = reinterpret_cast<int *>(cand_d + 2 * 1024);
    void *args[] = {(void *)&x, (void *)&N, (void *)&npoint, (void *)&start, (void *)&out_idx, (void *)&dist, (void *)&cand_d, (void *)&cand_i};
    if (cudaLaunchCooperativeKernel((void *)fps_kernel<T>, dim3(blocks), dim3(kFpsThreads), args, 0, s) != cudaSuccess) return RRL_ERR_CUDA;
    count_launch();
    return check_launch();
}

}  // namespace rrl

using namespace rrl;

extern "C" size_t rrl_fps_workspace_bytes(int N) {
    if (N <= 0) return 0;
    return ((((size_t)N + 63) / 64) * 64 + 2 * 1024) * sizeof(float) + 2 * 1024 * sizeof(int);
}

extern "C" int rrl_fps(const void *xyz, int is_double, int N, int npoint, int start, int *out_idx, void *workspace,
                       size_t workspace_bytes, void *stream) {
    if (!xyz || !out_idx || !workspace || N <= 0 || npoint <= 0 || npoint > N || start < 0 || start >= N) return RRL_ERR_ARG;
    if (workspace_bytes < rrl_fps_workspace_bytes(N)) return RRL_ERR_WORKSPACE;
    return is_double ? launch_fps<double>(xyz, N, npoint, start, out_idx, workspace, (cudaStream_t)stream)
                     : launch_fps<float>(xyz, N, npoint, start, out_idx, workspace, (cudaStream_t)stream);
}

extern "C" int rrl_knn(const void *xyz, int is_double, int N, const int *query_idx, int M, int k, int *out_idx, void *stream) {
    if (!xyz || !query_idx || !out_idx || N <= 0 || M <= 0 || k < 1 || k > kKnnMax || k > N) return RRL_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 grid((M + 7) / 8);
#define RRL_KNN_CASE(K)                                                                                                             \
    case K:                                                                                                                         \
        if (is_double) knn_kernel<double, K><<<grid, 256, 0, s>>>(reinterpret_cast<const double *>(xyz), N, query_idx, M, out_idx); \
        else knn_kernel<float, K><<<grid, 256, 0, s>>>(reinterpret_cast<const float *>(xyz), N, query_idx, M, out_idx);             \
        break;
    switch (k) {
        RRL_KNN_CASE(1) RRL_KNN_CASE(2) RRL_KNN_CASE(3) RRL_KNN_CASE(4) RRL_KNN_CASE(5) RRL_KNN_CASE(6) RRL_KNN_CASE(7) RRL_KNN_CASE(8)
    }
#undef RRL_KNN_CASE
    count_launch();
    return check_launch();
}
