// Sparse phase of the intersected-line loss on sm_100a: everything after the per-line intersection sets.
//
//   build     per line with (k, j) hits inside the window: sort the <=4 hit indices ascending (nonzero() order,
//             loss.py:125-131), recompute the three exact distances of every hit triplet, weights
//             w = d / ((d0+d1)+d2) (loss.py:92), intersection points q = ((w0 p0 + w1 p1) + w2 p2) / 3
//             (loss.py:155-163) and the k x j squared distances (loss.py:38-52,165-166); one record per selected line.
//   median    exact lower median (torch.median, loss.py:223-224) of all D entries of a pair by an 8-bit radix select on
//             the float bit patterns.
//   welsch    W = 1 - exp(-(D/med)/2) (loss.py:20-21,226), row/column minima with first-index tie breaking (torch.min),
//             per-(k,j) sums in 2^-40 fixed point (order independent), and the per-record gradient vectors
//             G1[a] = sum_b coef[a,b] dW/dD 2 (q1_a - q2_b), G2[b] = -sum_a ... (closed form).
//   finalize  loss = (1/C) sum_kj exp(-|k-j|/2) (S1/(n k) + S2/(n j))   (loss.py:215,227-230)
//   backward  d loss / d points: scatter (w/3) G grad_out to the hit triplets.
//
// Launches of the single-GPU forward: select (per line) -> build (per record, dense warps) -> median (one CTA per pair)
// -> welsch (per record; the last block of a pair also finalizes).  The line-sharded path (rrl_shard_*) runs the
// same kernels with collectives in between and a separate finalize.
#include "rrl_common.cuh"

namespace rrl {

// weights + intersection point of ONE hit triplet
__device__ __forceinline__ void make_point(const float *__restrict__ tri, int f, const float *ln, float *w /*[3]*/, float *q /*[3]*/) {
    const float *t = tri + (long long)f * 9;
    float v[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) v[c] = __ldg(t + c);
    const float d0 = __fsqrt_rn(point_line_x_exact(v[0], v[1], v[2], ln));
    const float d1 = __fsqrt_rn(point_line_x_exact(v[3], v[4], v[5], ln));
    const float d2 = __fsqrt_rn(point_line_x_exact(v[6], v[7], v[8], ln));
    const float s = __fadd_rn(__fadd_rn(d0, d1), d2);
    w[0] = __fdiv_rn(d0, s); w[1] = __fdiv_rn(d1, s); w[2] = __fdiv_rn(d2, s);
#pragma unroll
    for (int c = 0; c < 3; ++c)
        q[c] = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(w[0], v[c]), __fmul_rn(w[1], v[3 + c])), __fmul_rn(w[2], v[6 + c])), 3.0f);
}

__device__ __forceinline__ void cswap(int &a, int &b) {
    const int lo = min(a, b), hi = max(a, b);
    a = lo; b = hi;
}

// the record of one selected line, written to slot r (global record index).  Everything is fully unrolled with
// static indices and predicated on (a < k), (c < j) so that the per-hit arrays stay in registers.
__device__ __forceinline__ void build_record(const float *__restrict__ tri1, const float *__restrict__ tri2,
                                             const float *__restrict__ lines, const Workspace &ws, const Geometry &g, int b,
                                             int l, int k, int j, long long r) {
    const long long gl = (long long)b * g.nl + l;
    int i1[4], i2[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        i1[a] = a < k ? ws.hits[0][gl * kCap + a] : 0x7fffffff;
        i2[a] = a < j ? ws.hits[1][gl * kCap + a] : 0x7fffffff;
    }
    // ascending (nonzero() order); the padding sorts to the end
    cswap(i1[0], i1[1]); cswap(i1[2], i1[3]); cswap(i1[0], i1[2]); cswap(i1[1], i1[3]); cswap(i1[1], i1[2]);
    cswap(i2[0], i2[1]); cswap(i2[2], i2[3]); cswap(i2[0], i2[2]); cswap(i2[1], i2[3]); cswap(i2[1], i2[2]);
    float ln[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) ln[q] = __ldg(lines + gl * 6 + q);
    float wv[24], qv[24];
#pragma unroll
    for (int a = 0; a < 24; ++a) { wv[a] = 0.f; qv[a] = 0.f; }
    const float *t1 = tri1 + (long long)b * g.nf1 * 9, *t2 = tri2 + (long long)b * g.nf2 * 9;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        if (a < k) make_point(t1, i1[a], ln, wv + a * 3, qv + a * 3);
        if (a < j) make_point(t2, i2[a], ln, wv + 12 + a * 3, qv + 12 + a * 3);
    }
    float4 *D4 = reinterpret_cast<float4 *>(ws.recD + r * 16);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        float d[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
            d[c] = (a < k && c < j) ? sq3_rn(__fsub_rn(qv[a * 3], qv[12 + c * 3]), __fsub_rn(qv[a * 3 + 1], qv[12 + c * 3 + 1]),
                                             __fsub_rn(qv[a * 3 + 2], qv[12 + c * 3 + 2])) : 0.f;
        D4[a] = make_float4(d[0], d[1], d[2], d[3]);
    }
    reinterpret_cast<int2 *>(ws.recMeta)[r] = make_int2(l, k | (j << 8));
    int4 *I4 = reinterpret_cast<int4 *>(ws.recIdx + r * 8);
    I4[0] = make_int4(k > 0 ? i1[0] : -1, k > 1 ? i1[1] : -1, k > 2 ? i1[2] : -1, k > 3 ? i1[3] : -1);
    I4[1] = make_int4(j > 0 ? i2[0] : -1, j > 1 ? i2[1] : -1, j > 2 ? i2[2] : -1, j > 3 ? i2[3] : -1);
    float4 *W4 = reinterpret_cast<float4 *>(ws.recW + r * 24), *Q4 = reinterpret_cast<float4 *>(ws.recQ + r * 24);
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        W4[a] = make_float4(wv[4 * a], wv[4 * a + 1], wv[4 * a + 2], wv[4 * a + 3]);
        Q4[a] = make_float4(qv[4 * a], qv[4 * a + 1], qv[4 * a + 2], qv[4 * a + 3]);
    }
}

// One thread per line selects; the block compacts its selected lines in shared memory (ordered), claims a contiguous
// range of record slots with ONE atomic, and builds the records with densely populated warps.
__global__ void __launch_bounds__(256) build_kernel(const float *__restrict__ tri1, const float *__restrict__ tri2,
                                                    const float *__restrict__ lines, Workspace ws, Geometry g,
                                                    int k_lo, int j_lo, int k_hi, int j_hi) {
    __shared__ int s_line[256], s_kj[256], s_warp[8], s_hist[16], s_base;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int l = blockIdx.x * blockDim.x + tid;
    if (tid < 16) s_hist[tid] = 0;
    int k = 0, j = 0;
    bool sel = false;
    if (l < g.nl) {
        const long long gl = (long long)b * g.nl + l;
        k = ws.cnt[0][gl]; j = ws.cnt[1][gl];
        sel = k >= k_lo && k < k_hi && j >= j_lo && j < j_hi;      // windows are validated to lie inside 1..4
    }
    const unsigned bal = __ballot_sync(0xffffffffu, sel);
    if (lane == 0) s_warp[wid] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        before += w < wid ? s_warp[w] : 0;
        total += s_warp[w];
    }
    if (total == 0) return;
    if (sel) {
        const int pos = before + __popc(bal & ((1u << lane) - 1u));
        s_line[pos] = l;
        s_kj[pos] = k | (j << 8);
        atomicAdd(&s_hist[(k - 1) * 4 + (j - 1)], 1);
    }
    if (tid == 0) s_base = atomicAdd(ws.nrec + b, total);
    __syncthreads();
    if (tid < 16 && s_hist[tid]) atomicAdd(ws.n_kj + b * 16 + tid, s_hist[tid]);
    if (tid < total) {
        const int kj = s_kj[tid];
        build_record(tri1, tri2, lines, ws, g, b, s_line[tid], kj & 255, (kj >> 8) & 255, (long long)b * g.nl + s_base + tid);
    }
}

int launch_build(const float *tri1, const float *tri2, const float *lines, const Workspace &ws, const Geometry &g,
                 int k_lo, int j_lo, int k_hi, int j_hi, cudaStream_t s) {
    stage_mark(5, s);
    build_kernel<<<dim3((g.nl + 255) / 256, g.B), 256, 0, s>>>(tri1, tri2, lines, ws, g, k_lo, j_lo, k_hi, j_hi);
    count_launch();
    stage_mark(6, s);
    return check_launch();
}

// local counts -> gcounts (B,18): n_kj[16], #records, #D entries.  In the single-GPU path these ARE the global counts.
__global__ void local_counts_kernel(Workspace ws, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    long long nD = 0;
    for (int c = 0; c < 16; ++c) {
        const long long n = ws.n_kj[b * 16 + c];
        ws.gcounts[b * 18 + c] = n;
        nD += n * ((c >> 2) + 1) * ((c & 3) + 1);
    }
    ws.gcounts[b * 18 + 16] = ws.nrec[b];
    ws.gcounts[b * 18 + 17] = nD;
}

int launch_local_counts(const Workspace &ws, const Geometry &g, cudaStream_t s) {
    local_counts_kernel<<<(g.B + 127) / 128, 128, 0, s>>>(ws, g.B);
    count_launch();
    return check_launch();
}

// ------------------------------------------------------------------------------------------------------
// exact lower median by radix select (non-negative floats order like their bit patterns)
// ------------------------------------------------------------------------------------------------------
constexpr int kSelThreads = 1024;

// `key(i, valid)` enumerates `slots` slots, `n` of which are valid; returns the key of rank (n-1)/2.  Block-wide.
template <typename KeyFn>
__device__ unsigned radix_select_lower_median(long long slots, long long n, KeyFn key, unsigned *hist /*smem[256]*/,
                                              unsigned *s_prefix, long long *s_rank) {
    unsigned prefix = 0, mask = 0;
    long long rank = (n - 1) / 2;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (long long i0 = 0; i0 < slots; i0 += blockDim.x) {       // block-uniform trip count (warp collectives inside)
            const long long i = i0 + threadIdx.x;
            bool valid = false;
            unsigned kbits = 0;
            if (i < slots) kbits = key(i, valid);
            const bool in = valid && (kbits & mask) == prefix;
            // D values crowd a few bins (same exponent): aggregate equal bins inside the warp before the atomic
            const unsigned bin = in ? ((kbits >> shift) & 255u) : 0xFFFFFFFFu;
            const unsigned peers = __match_any_sync(0xffffffffu, bin);
            if (in && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], (unsigned)__popc(peers));
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // warp 0 finds the bin holding `rank`: 8 bins per lane, exclusive warp scan of the lane totals
            const int lane = threadIdx.x;
            unsigned h[8], tot = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { h[q] = hist[lane * 8 + q]; tot += h[q]; }
            unsigned inc = tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned up = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += up;
            }
            const long long before = (long long)(inc - tot);
            if (rank >= before && rank < before + (long long)tot) {
                long long r = rank - before;
                int q = 0;
                for (; q < 7; ++q) {
                    if (r < (long long)h[q]) break;
                    r -= h[q];
                }
                *s_prefix = prefix | ((unsigned)(lane * 8 + q) << shift);
                *s_rank = r;
            }
        }
        __syncthreads();
        prefix = *s_prefix;
        rank = *s_rank;
        mask |= 255u << shift;
        __syncthreads();
    }
    return prefix;
}

constexpr int kMedCache = 48 * 1024;        // D slots cached in shared memory (192 KB); the rest is re-read through L2

__global__ void __launch_bounds__(kSelThreads) median_kernel(Workspace ws, Geometry g) {
    extern __shared__ unsigned s_keys[];     // [kMedCache]; 0xFFFFFFFF = not a D entry (D >= 0 never has that pattern)
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix;
    __shared__ long long s_rank;
    const int b = blockIdx.x;
    const int nrec = ws.nrec[b];
    long long n = 0;
    for (int c = 0; c < 16; ++c) n += (long long)ws.n_kj[b * 16 + c] * ((c >> 2) + 1) * ((c & 3) + 1);
    if (threadIdx.x < 16) ws.gcounts[b * 18 + threadIdx.x] = ws.n_kj[b * 16 + threadIdx.x];   // single-GPU: local == global
    if (threadIdx.x == 16) ws.gcounts[b * 18 + 16] = nrec;
    if (threadIdx.x == 17) ws.gcounts[b * 18 + 17] = n;
    if (n == 0) {
        if (threadIdx.x == 0) ws.med[b] = 0.f;
        return;
    }
    const float *D = ws.recD + (long long)b * g.nl * 16;
    const int *meta = ws.recMeta + (long long)b * g.nl * 2;
    const long long slots = (long long)nrec * 16;
    auto gkey = [&](long long i) -> unsigned {
        const int kj = meta[(i >> 4) * 2 + 1];
        const int e = (int)(i & 15);
        const bool valid = (e >> 2) < (kj & 255) && (e & 3) < ((kj >> 8) & 255);
        return valid ? __float_as_uint(D[i]) : 0xFFFFFFFFu;
    };
    // one pipelined sweep over the records fills the cache; the four select passes then run out of shared memory
    const int ncache = (int)(slots < kMedCache ? slots : kMedCache);
    // (both loads of a slot are issued unconditionally and eight slots are in flight per thread: with one CTA per pair
    // the sweep is otherwise a chain of dependent L2 round trips)
    for (int i0 = 0; i0 < ncache; i0 += 8 * kSelThreads) {
        int kj[8];
        unsigned kb[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * kSelThreads + threadIdx.x;
            kj[u] = 0; kb[u] = 0;
            if (i < ncache) { kj[u] = meta[(i >> 4) * 2 + 1]; kb[u] = __float_as_uint(D[i]); }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * kSelThreads + threadIdx.x;
            const int e = i & 15;
            const bool valid = (e >> 2) < (kj[u] & 255) && (e & 3) < ((kj[u] >> 8) & 255);
            if (i < ncache) s_keys[i] = valid ? kb[u] : 0xFFFFFFFFu;
        }
    }
    __syncthreads();
    auto key = [&](long long i, bool &valid) -> unsigned {
        const unsigned kb = i < ncache ? s_keys[i] : gkey(i);
        valid = kb != 0xFFFFFFFFu;
        return kb;
    };
    const unsigned bits = radix_select_lower_median(slots, n, key, hist, &s_prefix, &s_rank);
    if (threadIdx.x == 0) ws.med[b] = __uint_as_float(bits);
}

int launch_median(const Workspace &ws, const Geometry &g, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(median_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMedCache * 4) != cudaSuccess) return RRL_ERR_CUDA;
        attr_set = true;
    }
    median_kernel<<<g.B, kSelThreads, kMedCache * 4, s>>>(ws, g);
    count_launch();
    stage_mark(7, s);
    return check_launch();
}

__global__ void __launch_bounds__(kSelThreads) select_flat_kernel(const float *__restrict__ vals, long long n, float *out) {
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix;
    __shared__ long long s_rank;
    if (n <= 0) {
        if (threadIdx.x == 0) *out = 0.f;
        return;
    }
    auto key = [&](long long i, bool &valid) -> unsigned {
        valid = true;
        return __float_as_uint(vals[i]);
    };
    const unsigned bits = radix_select_lower_median(n, n, key, hist, &s_prefix, &s_rank);
    if (threadIdx.x == 0) *out = __uint_as_float(bits);
}

int launch_select_median(const float *vals, long long n, float *out, cudaStream_t s) {
    select_flat_kernel<<<1, kSelThreads, 0, s>>>(vals, n, out);
    count_launch();
    return check_launch();
}

// compact the valid D entries of pair 0 (line-sharded path)
__global__ void pack_entries_kernel(Workspace ws, Geometry g, float *out, long long cap) {
    const int nrec = ws.nrec[0];
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrec) return;
    // entries of a record go to a block claimed with an atomic cursor (order is irrelevant to a median)
    const int kj = ws.recMeta[r * 2 + 1];
    const int k = kj & 255, j = (kj >> 8) & 255;
    const long long pos = (long long)atomicAdd((unsigned long long *)(ws.stats + 7), (unsigned long long)(k * j));   // stats[7]: pack cursor
    for (int a = 0; a < k; ++a)
        for (int c = 0; c < j; ++c) {
            const long long p = pos + a * j + c;
            if (p < cap) out[p] = ws.recD[r * 16 + a * 4 + c];
        }
}

int launch_pack_entries(const Workspace &ws, const Geometry &g, float *out, long long cap, cudaStream_t s) {
    if (cudaMemsetAsync(ws.stats + 7, 0, sizeof(long long), s) != cudaSuccess) return RRL_ERR_CUDA;
    pack_entries_kernel<<<(g.nl + 255) / 256, 256, 0, s>>>(ws, g, out, cap);
    count_launch();
    return check_launch();
}

// ------------------------------------------------------------------------------------------------------
// Welsch + minima + gradient vectors
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float welsch(float D, float med) {
    // 1 - exp(-((x / c)) / 2.0)   loss.py:21
    // exp evaluated in double and rounded once: a well-defined (correctly rounded) float exp, so that the row/column
    // argmin of near-tied entries does not depend on a vendor expf's last bit (the oracle does the same)
    return __fsub_rn(1.0f, (float)exp((double)(-__fdiv_rn(__fdiv_rn(D, med), 2.0f))));
}

// One record: minima -> (s1, s2); the gradient vectors for a unit upstream gradient overwrite the record's
// intersection points (recQ), which nothing reads afterwards.  cw_over_n = exp(-|k-j|/2) / C / n_kj.
__device__ __forceinline__ void welsch_record(const Workspace &ws, long long r, float med, double cw_over_n, int k, int j,
                                              double &s1, double &s2) {
    float W[16], D[16];
    const float4 *D4 = reinterpret_cast<const float4 *>(ws.recD + r * 16);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const float4 d = D4[a];
        D[a * 4] = d.x; D[a * 4 + 1] = d.y; D[a * 4 + 2] = d.z; D[a * 4 + 3] = d.w;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) W[a * 4 + c] = (a < k && c < j) ? welsch(D[a * 4 + c], med) : 0.f;
    int arg_b[4], arg_a[4];
    s1 = 0.0; s2 = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {                   // torch.min(W, 2): first index on ties
        int m = 0;
#pragma unroll
        for (int c = 1; c < 4; ++c) if (c < j && W[a * 4 + c] < W[a * 4 + m]) m = c;
        arg_b[a] = m;
        if (a < k) s1 += (double)W[a * 4 + m];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {                   // torch.min(W, 1)
        int m = 0;
#pragma unroll
        for (int a = 1; a < 4; ++a) if (a < k && W[a * 4 + c] < W[m * 4 + c]) m = a;
        arg_a[c] = m;
        if (c < j) s2 += (double)W[m * 4 + c];
    }
    float4 *Q4 = reinterpret_cast<float4 *>(ws.recQ + r * 24);
    float q[24];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        const float4 v = Q4[a];
        q[4 * a] = v.x; q[4 * a + 1] = v.y; q[4 * a + 2] = v.z; q[4 * a + 3] = v.w;
    }
    double G[24];
#pragma unroll
    for (int a = 0; a < 24; ++a) G[a] = 0.0;
    const double inv2med = 1.0 / (2.0 * (double)med);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (a >= k || c >= j) continue;
            double coef = 0.0;
            if (arg_b[a] == c) coef += cw_over_n / k;
            if (arg_a[c] == a) coef += cw_over_n / j;
            if (coef == 0.0) continue;
            const double f = coef * exp(-(double)D[a * 4 + c] * inv2med) * inv2med * 2.0;
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                const double gv = f * ((double)q[a * 3 + x] - (double)q[12 + c * 3 + x]);
                G[a * 3 + x] += gv;
                G[12 + c * 3 + x] -= gv;
            }
        }
#pragma unroll
    for (int a = 0; a < 6; ++a) Q4[a] = make_float4((float)G[4 * a], (float)G[4 * a + 1], (float)G[4 * a + 2], (float)G[4 * a + 3]);
}

__device__ __forceinline__ int count_combos(const long long *gc) {
    int C = 0;
    for (int c = 0; c < 16; ++c) C += gc[c] > 0;
    return C;
}

__device__ __forceinline__ void finalize_pair(const Workspace &ws, int b, float *out_loss, int *out_status, float *out_median,
                                              long long *out_stats) {
    const long long *gc = ws.gcounts + b * 18;
    int C = 0;
    double loss = 0.0;
    for (int k = 1; k <= 4; ++k)
        for (int j = 1; j <= 4; ++j) {
            const int c = (k - 1) * 4 + (j - 1);
            const long long n = gc[c];
            if (n <= 0) continue;
            ++C;
            const double S1 = (double)__ldcg(ws.sums + b * 32 + c) / kFixScale, S2 = (double)__ldcg(ws.sums + b * 32 + 16 + c) / kFixScale;
            loss += exp(-0.5 * (double)abs(k - j)) * (S1 / ((double)n * k) + S2 / ((double)n * j));
        }
    int status = 0;
    if (C == 0) status |= RRL_STATUS_EMPTY; else loss /= (double)C;
    long long *st = ws.stats + (long long)b * RRL_NSTAT;
    st[0] = gc[16]; st[1] = gc[17]; st[2] = C;
    if (st[6] > 0) status |= RRL_STATUS_NAN;
    if (ws.flags[b * 4]) { status |= RRL_STATUS_NAN; loss = __longlong_as_double(0x7ff8000000000000LL); }
    // |AC|^2 <= (P + X)^2: flag clouds whose own extent already makes the 2e-4 offset smaller than a few ulps of
    // |p|^2 (|AC|^2 >~ 1e3)
    const float pm = fmaxf(__uint_as_float(ws.pmax[b * 2]), __uint_as_float(ws.pmax[b * 2 + 1]));
    if (pm > 250.0f) status |= RRL_STATUS_NAN_RISK;
    out_loss[b] = (float)loss;
    if (out_status) out_status[b] = status;
    if (out_median) out_median[b] = ws.med[b];
    if (out_stats)
        for (int q = 0; q < RRL_NSTAT; ++q) out_stats[(long long)b * RRL_NSTAT + q] = st[q];
}

// gcounts / med hold the GLOBAL values (rrl_shard_stage2) in the line-sharded path.  With out_loss != nullptr
// (single-GPU forward) the last block of a pair to finish also writes the loss (ticket in flags[b*2+1]).
__global__ void __launch_bounds__(256) welsch_kernel(Workspace ws, Geometry g, float *out_loss, int *out_status, float *out_median,
                                                     long long *out_stats) {
    __shared__ unsigned long long s_sum[32];
    __shared__ int s_last;
    const int b = blockIdx.y;
    if (threadIdx.x < 32) s_sum[threadIdx.x] = 0ull;
    __syncthreads();
    const long long nrec = ws.nrec[b];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if ((long long)blockIdx.x * blockDim.x >= nrec && blockIdx.x != 0) return;      // whole block beyond the records
    if (i < nrec) {
        const long long r = (long long)b * g.nl + i;
        const long long *gc = ws.gcounts + b * 18;
        const int C = count_combos(gc);
        const int kj = ws.recMeta[r * 2 + 1];
        const int k = kj & 255, j = (kj >> 8) & 255;
        const int combo = (k - 1) * 4 + (j - 1);
        double s1, s2;
        welsch_record(ws, r, ws.med[b], exp(-0.5 * (double)abs(k - j)) / (double)C / (double)gc[combo], k, j, s1, s2);
        if (!(s1 == s1) || !(s2 == s2)) { ws.flags[b * 4] = 1; s1 = s2 = 0.0; }    // e.g. median 0: the reference's loss is NaN too
        atomicAdd(&s_sum[combo], (unsigned long long)__double2ll_rn(s1 * kFixScale));
        atomicAdd(&s_sum[16 + combo], (unsigned long long)__double2ll_rn(s2 * kFixScale));
    }
    __syncthreads();
    if (threadIdx.x < 32 && s_sum[threadIdx.x]) atomicAdd(ws.sums + b * 32 + threadIdx.x, s_sum[threadIdx.x]);
    if (!out_loss) return;
    // the blocks that hold records (at least block 0) take a ticket; the last one sees every sum
    const int nblocks = nrec > 0 ? (int)((nrec + blockDim.x - 1) / blockDim.x) : 1;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ws.flags + b * 4 + 1, 1) == nblocks - 1);
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        finalize_pair(ws, b, out_loss, out_status, out_median, out_stats);
    }
}

int launch_welsch(const Workspace &ws, const Geometry &g, cudaStream_t s) {
    dim3 grid((g.nl + 255) / 256, g.B);
    welsch_kernel<<<grid, 256, 0, s>>>(ws, g, nullptr, nullptr, nullptr, nullptr);
    count_launch();
    return check_launch();
}

int launch_welsch_finalize(const Workspace &ws, const Geometry &g, float *out_loss, int *out_status, float *out_median,
                           long long *out_stats, cudaStream_t s) {
    dim3 grid((g.nl + 255) / 256, g.B);
    welsch_kernel<<<grid, 256, 0, s>>>(ws, g, out_loss, out_status, out_median, out_stats);
    count_launch();
    stage_mark(8, s);
    return check_launch();
}

__global__ void finalize_kernel(Workspace ws, Geometry g, float *out_loss, int *out_status, float *out_median,
                                long long *out_stats) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= g.B) return;
    finalize_pair(ws, b, out_loss, out_status, out_median, out_stats);
}

int launch_finalize(const Workspace &ws, const Geometry &g, float *out_loss, int *out_status, float *out_median,
                    long long *out_stats, cudaStream_t s) {
    finalize_kernel<<<(g.B + 127) / 128, 128, 0, s>>>(ws, g, out_loss, out_status, out_median, out_stats);
    count_launch();
    return check_launch();
}

// ------------------------------------------------------------------------------------------------------
// backward: scatter (w/3) G grad_out
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) backward_kernel(Workspace ws, Geometry g, const float *__restrict__ grad_out,
                                                       float *__restrict__ g1, float *__restrict__ g2) {
    const int b = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ws.nrec[b]) return;
    const long long r = (long long)b * g.nl + i;
    const int meta = ws.recMeta[r * 2 + 1];
    const int k = meta & 255, j = (meta >> 8) & 255;
    const float go = grad_out[b] * (1.0f / 3.0f);
    const float *G = ws.recQ + r * 24, *Wt = ws.recW + r * 24;
    const int *idx = ws.recIdx + r * 8;
    if (g1) {
        float *O = g1 + (long long)b * g.nf1 * 9;
        for (int a = 0; a < k; ++a) {
            const long long f = idx[a];
            const float gx = G[a * 3] * go, gy = G[a * 3 + 1] * go, gz = G[a * 3 + 2] * go;
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                const float w = Wt[a * 3 + p];
                atomicAdd(O + f * 9 + p * 3, w * gx);
                atomicAdd(O + f * 9 + p * 3 + 1, w * gy);
                atomicAdd(O + f * 9 + p * 3 + 2, w * gz);
            }
        }
    }
    if (g2) {
        float *O = g2 + (long long)b * g.nf2 * 9;
        for (int c = 0; c < j; ++c) {
            const long long f = idx[4 + c];
            const float gx = G[12 + c * 3] * go, gy = G[12 + c * 3 + 1] * go, gz = G[12 + c * 3 + 2] * go;
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                const float w = Wt[12 + c * 3 + p];
                atomicAdd(O + f * 9 + p * 3, w * gx);
                atomicAdd(O + f * 9 + p * 3 + 1, w * gy);
                atomicAdd(O + f * 9 + p * 3 + 2, w * gz);
            }
        }
    }
}

int launch_backward(const Workspace &ws, const Geometry &g, const float *grad_out, float *g1, float *g2, cudaStream_t s) {
    if (g1 && cudaMemsetAsync(g1, 0, sizeof(float) * 9 * (size_t)g.B * g.nf1, s) != cudaSuccess) return RRL_ERR_CUDA;
    if (g2 && cudaMemsetAsync(g2, 0, sizeof(float) * 9 * (size_t)g.B * g.nf2, s) != cudaSuccess) return RRL_ERR_CUDA;
    dim3 grid((g.nl + 127) / 128, g.B);
    backward_kernel<<<grid, 128, 0, s>>>(ws, g, grad_out, g1, g2);
    count_launch();
    stage_mark(9, s);
    return check_launch();
}

// ------------------------------------------------------------------------------------------------------
// export of the per-line intersection sets (sorted ascending, -1 padded)
// ------------------------------------------------------------------------------------------------------
__global__ void export_hits_kernel(Workspace ws, Geometry g, int cloud, int *out_counts, int *out_hits) {
    const long long gl = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gl >= (long long)g.B * g.nl) return;
    const int c = ws.cnt[cloud][gl];
    out_counts[gl] = c;
    int v[kCap];
    const int n = c < kCap ? c : kCap;
    for (int a = 0; a < kCap; ++a) v[a] = a < n ? ws.hits[cloud][gl * kCap + a] : 0x7fffffff;
    for (int a = 1; a < kCap; ++a) {
        int x = v[a], q = a - 1;
        while (q >= 0 && v[q] > x) { v[q + 1] = v[q]; --q; }
        v[q + 1] = x;
    }
    for (int a = 0; a < kCap; ++a) out_hits[gl * kCap + a] = a < n ? v[a] : -1;
}

int launch_export_hits(const Workspace &ws, const Geometry &g, int cloud, int *out_counts, int *out_hits, cudaStream_t s) {
    const long long n = (long long)g.B * g.nl;
    export_hits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ws, g, cloud, out_counts, out_hits);
    count_launch();
    return check_launch();
}

}  // namespace rrl
