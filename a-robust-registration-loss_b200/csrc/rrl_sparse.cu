// Sparse phase of the intersected-line loss on sm_100a: everything after the per-line intersection sets.
//
//   build_kernel     per line with (k, j) hits inside the window: sort the <=4 hit indices ascending
//                    (nonzero() order, loss.py:125-131), recompute the three exact distances of every hit
//                    triplet, weights w = d / ((d0+d1)+d2) (loss.py:92), intersection points
//                    q = ((w0 p0 + w1 p1) + w2 p2) / 3 (loss.py:155-163) and the k x j squared distances
//                    (loss.py:38-52,165-166); appends one record per selected line.
//   median_kernel    exact lower median (torch.median, loss.py:223-224) of all D entries of a pair by an
//                    8-bit radix select on the float bit patterns; one CTA per pair.
//   welsch_kernel    W = 1 - exp(-(D/med)/2) (loss.py:20-21,226), row/column minima with first-index tie
//                    breaking (torch.min), per-(k,j) sums accumulated in 2^-40 fixed point so the result does
//                    not depend on the order records were appended in.
//   finalize_kernel  loss = (1/C) sum_kj exp(-|k-j|/2) (S1/(n k) + S2/(n j))   (loss.py:215,227-230)
//   backward_kernel  closed-form gradient scattered to the hit triplets.
#include "rrl_common.cuh"

namespace rrl {

__device__ __forceinline__ void sort4(int *v, int n) {
    // ascending insertion sort of n <= 4 entries
    for (int i = 1; i < n; ++i) {
        int x = v[i], j = i - 1;
        while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; }
        v[j + 1] = x;
    }
}

// weights + intersection points of the n (<=4) hit triplets of one cloud
__device__ __forceinline__ void make_points(const float *__restrict__ tri, const int *idx, int n, const float *ln,
                                            float *w /*[n][3]*/, float *q /*[n][3]*/) {
    for (int a = 0; a < n; ++a) {
        const float *t = tri + (long long)idx[a] * 9;
        float v[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) v[c] = __ldg(t + c);
        const float d0 = __fsqrt_rn(point_line_x_exact(v[0], v[1], v[2], ln));
        const float d1 = __fsqrt_rn(point_line_x_exact(v[3], v[4], v[5], ln));
        const float d2 = __fsqrt_rn(point_line_x_exact(v[6], v[7], v[8], ln));
        const float s = __fadd_rn(__fadd_rn(d0, d1), d2);
        const float w0 = __fdiv_rn(d0, s), w1 = __fdiv_rn(d1, s), w2 = __fdiv_rn(d2, s);
        w[a * 3 + 0] = w0; w[a * 3 + 1] = w1; w[a * 3 + 2] = w2;
#pragma unroll
        for (int c = 0; c < 3; ++c)
            q[a * 3 + c] = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(w0, v[c]), __fmul_rn(w1, v[3 + c])), __fmul_rn(w2, v[6 + c])), 3.0f);
    }
}

__global__ void __launch_bounds__(128) build_kernel(const float *__restrict__ tri1, const float *__restrict__ tri2,
                                                    const float *__restrict__ lines, Workspace ws, Geometry g,
                                                    int k_lo, int j_lo, int k_hi, int j_hi) {
    const int b = blockIdx.y;
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= g.nl) return;
    const long long gl = (long long)b * g.nl + l;
    const int k = ws.cnt[0][gl], j = ws.cnt[1][gl];
    if (k < k_lo || k >= k_hi || j < j_lo || j >= j_hi) return;     // windows are validated to lie inside 1..4

    int i1[4], i2[4];
    for (int a = 0; a < k; ++a) i1[a] = ws.hits[0][gl * kCap + a];
    for (int a = 0; a < j; ++a) i2[a] = ws.hits[1][gl * kCap + a];
    sort4(i1, k);
    sort4(i2, j);
    float ln[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) ln[q] = __ldg(lines + gl * 6 + q);
    float w1[12], w2[12], q1[12], q2[12];
    make_points(tri1 + (long long)b * g.nf1 * 9, i1, k, ln, w1, q1);
    make_points(tri2 + (long long)b * g.nf2 * 9, i2, j, ln, w2, q2);

    const int slot = atomicAdd(ws.nrec + b, 1);
    atomicAdd(ws.n_kj + b * 16 + (k - 1) * 4 + (j - 1), 1);
    const long long r = (long long)b * g.nl + slot;
    float *D = ws.recD + r * 16;
    for (int a = 0; a < 4; ++a)
        for (int c = 0; c < 4; ++c) {
            float d = 0.f;
            if (a < k && c < j)
                d = sq3_rn(__fsub_rn(q1[a * 3], q2[c * 3]), __fsub_rn(q1[a * 3 + 1], q2[c * 3 + 1]), __fsub_rn(q1[a * 3 + 2], q2[c * 3 + 2]));
            D[a * 4 + c] = d;
        }
    ws.recMeta[r * 2] = l;
    ws.recMeta[r * 2 + 1] = k | (j << 8);
    for (int a = 0; a < 4; ++a) {
        ws.recIdx[r * 8 + a] = a < k ? i1[a] : -1;
        ws.recIdx[r * 8 + 4 + a] = a < j ? i2[a] : -1;
    }
    for (int a = 0; a < 12; ++a) {
        ws.recW[r * 24 + a] = a < 3 * k ? w1[a] : 0.f;
        ws.recW[r * 24 + 12 + a] = a < 3 * j ? w2[a] : 0.f;
        ws.recQ[r * 24 + a] = a < 3 * k ? q1[a] : 0.f;
        ws.recQ[r * 24 + 12 + a] = a < 3 * j ? q2[a] : 0.f;
    }
}

int launch_build(const float *tri1, const float *tri2, const float *lines, const Workspace &ws, const Geometry &g,
                 int k_lo, int j_lo, int k_hi, int j_hi, cudaStream_t s) {
    dim3 grid((g.nl + 127) / 128, g.B);
    build_kernel<<<grid, 128, 0, s>>>(tri1, tri2, lines, ws, g, k_lo, j_lo, k_hi, j_hi);
    count_launch();
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

// Generic body: `key(i, valid)` enumerates `slots` slots, `n` of which are valid; returns the key of rank (n-1)/2.
template <typename KeyFn>
__device__ unsigned radix_select_lower_median(long long slots, long long n, KeyFn key, unsigned *hist /*smem[256]*/,
                                              unsigned *s_prefix, long long *s_rank) {
    unsigned prefix = 0, mask = 0;
    long long rank = (n - 1) / 2;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (long long i = threadIdx.x; i < slots; i += blockDim.x) {
            bool valid;
            const unsigned kbits = key(i, valid);
            if (valid && (kbits & mask) == prefix) atomicAdd(&hist[(kbits >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long r = rank;
            unsigned bin = 0;
            for (; bin < 255; ++bin) {
                if (r < (long long)hist[bin]) break;
                r -= hist[bin];
            }
            *s_prefix = prefix | (bin << shift);
            *s_rank = r;
        }
        __syncthreads();
        prefix = *s_prefix;
        rank = *s_rank;
        mask |= 255u << shift;
        __syncthreads();
    }
    return prefix;
}

__global__ void __launch_bounds__(kSelThreads) median_kernel(Workspace ws, Geometry g) {
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
    auto key = [&](long long i, bool &valid) -> unsigned {
        const int kj = meta[(i >> 4) * 2 + 1];
        const int e = (int)(i & 15);
        valid = (e >> 2) < (kj & 255) && (e & 3) < ((kj >> 8) & 255);
        return __float_as_uint(D[i]);
    };
    const unsigned bits = radix_select_lower_median((long long)nrec * 16, n, key, hist, &s_prefix, &s_rank);
    if (threadIdx.x == 0) ws.med[b] = __uint_as_float(bits);
}

int launch_median(const Workspace &ws, const Geometry &g, cudaStream_t s) {
    median_kernel<<<g.B, kSelThreads, 0, s>>>(ws, g);
    count_launch();
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
    // deterministic position: entries of record r start at the prefix sum of k*j over earlier records -- computed
    // with a per-record atomic cursor instead (order is irrelevant to a median)
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
// Welsch + minima
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float welsch(float D, float med) {
    // 1 - exp(-((x / c)) / 2.0)   loss.py:21
    // exp evaluated in double and rounded once: a well-defined (correctly rounded) float exp, so that the row/column
    // argmin of near-tied entries does not depend on a vendor expf's last bit (the oracle does the same)
    return __fsub_rn(1.0f, (float)exp((double)(-__fdiv_rn(__fdiv_rn(D, med), 2.0f))));
}

__global__ void __launch_bounds__(256) welsch_kernel(Workspace ws, Geometry g) {
    __shared__ unsigned long long s_sum[32];
    const int b = blockIdx.y;
    if (threadIdx.x < 32) s_sum[threadIdx.x] = 0ull;
    __syncthreads();
    const long long nrec = ws.nrec[b];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nrec) {
        const long long r = (long long)b * g.nl + i;
        const float med = ws.med[b];
        const int kj = ws.recMeta[r * 2 + 1];
        const int k = kj & 255, j = (kj >> 8) & 255;
        float W[16];
        const float4 *D4 = reinterpret_cast<const float4 *>(ws.recD + r * 16);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float4 d = D4[a];
            W[a * 4 + 0] = welsch(d.x, med); W[a * 4 + 1] = welsch(d.y, med);
            W[a * 4 + 2] = welsch(d.z, med); W[a * 4 + 3] = welsch(d.w, med);
        }
        unsigned args = 0;
        double s1 = 0.0, s2 = 0.0;
        for (int a = 0; a < k; ++a) {                   // torch.min(W, 2): first index on ties
            int m = 0;
            for (int c = 1; c < j; ++c) if (W[a * 4 + c] < W[a * 4 + m]) m = c;
            args |= (unsigned)m << (2 * a);
            s1 += (double)W[a * 4 + m];
        }
        for (int c = 0; c < j; ++c) {                   // torch.min(W, 1)
            int m = 0;
            for (int a = 1; a < k; ++a) if (W[a * 4 + c] < W[m * 4 + c]) m = a;
            args |= (unsigned)m << (8 + 2 * c);
            s2 += (double)W[m * 4 + c];
        }
        ws.recMeta[r * 2 + 1] = (kj & 0xffff) | (int)(args << 16);
        if (!(s1 == s1) || !(s2 == s2)) { ws.flags[b * 2] = 1; s1 = s2 = 0.0; }    // e.g. median 0: the reference's loss is NaN too
        const int combo = (k - 1) * 4 + (j - 1);
        atomicAdd(&s_sum[combo], (unsigned long long)__double2ll_rn(s1 * kFixScale));
        atomicAdd(&s_sum[16 + combo], (unsigned long long)__double2ll_rn(s2 * kFixScale));
    }
    __syncthreads();
    if (threadIdx.x < 32 && s_sum[threadIdx.x]) atomicAdd(ws.sums + b * 32 + threadIdx.x, s_sum[threadIdx.x]);
}

int launch_welsch(const Workspace &ws, const Geometry &g, cudaStream_t s) {
    dim3 grid((g.nl + 255) / 256, g.B);
    welsch_kernel<<<grid, 256, 0, s>>>(ws, g);
    count_launch();
    return check_launch();
}

__global__ void finalize_kernel(Workspace ws, Geometry g, float *out_loss, int *out_status, float *out_median,
                                long long *out_stats) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= g.B) return;
    const long long *gc = ws.gcounts + b * 18;
    int C = 0;
    double loss = 0.0;
    for (int k = 1; k <= 4; ++k)
        for (int j = 1; j <= 4; ++j) {
            const int c = (k - 1) * 4 + (j - 1);
            const long long n = gc[c];
            if (n <= 0) continue;
            ++C;
            const double S1 = (double)ws.sums[b * 32 + c] / kFixScale, S2 = (double)ws.sums[b * 32 + 16 + c] / kFixScale;
            loss += exp(-0.5 * (double)abs(k - j)) * (S1 / ((double)n * k) + S2 / ((double)n * j));
        }
    int status = 0;
    if (C == 0) status |= RRL_STATUS_EMPTY; else loss /= (double)C;
    long long *st = ws.stats + (long long)b * RRL_NSTAT;
    st[0] = gc[16]; st[1] = gc[17]; st[2] = C;
    if (st[6] > 0) status |= RRL_STATUS_NAN;
    if (ws.flags[b * 2]) { status |= RRL_STATUS_NAN; loss = __longlong_as_double(0x7ff8000000000000LL); }
    // |AC|^2 <= (P1 + X)^2 with X <= ~ the sampling sphere: flag clouds whose own extent already makes the
    // 2e-4 offset smaller than a few ulps of |p|^2 (|AC|^2 >~ 1e3)
    const float pm = fmaxf(__uint_as_float(ws.pmax[b * 2]), __uint_as_float(ws.pmax[b * 2 + 1]));
    if (pm > 250.0f) status |= RRL_STATUS_NAN_RISK;
    out_loss[b] = (float)loss;
    if (out_status) out_status[b] = status;
    if (out_median) out_median[b] = ws.med[b];
    if (out_stats)
        for (int q = 0; q < RRL_NSTAT; ++q) out_stats[(long long)b * RRL_NSTAT + q] = st[q];
}

int launch_finalize(const Workspace &ws, const Geometry &g, float *out_loss, int *out_status, float *out_median,
                    long long *out_stats, cudaStream_t s) {
    finalize_kernel<<<(g.B + 127) / 128, 128, 0, s>>>(ws, g, out_loss, out_status, out_median, out_stats);
    count_launch();
    return check_launch();
}

// ------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) backward_kernel(Workspace ws, Geometry g, const float *__restrict__ grad_out,
                                                       float *__restrict__ g1, float *__restrict__ g2) {
    const int b = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ws.nrec[b]) return;
    const long long *gc = ws.gcounts + b * 18;
    int C = 0;
    for (int c = 0; c < 16; ++c) C += gc[c] > 0;
    const long long r = (long long)b * g.nl + i;
    const int meta = ws.recMeta[r * 2 + 1];
    const int k = meta & 255, j = (meta >> 8) & 255;
    const unsigned args = (unsigned)meta >> 16;
    const double med = (double)ws.med[b];
    const double n = (double)gc[(k - 1) * 4 + (j - 1)];
    const double cw = exp(-0.5 * (double)abs(k - j)) / (double)C * (double)grad_out[b];
    const float *D = ws.recD + r * 16, *Q = ws.recQ + r * 24, *Wt = ws.recW + r * 24;
    double gq1[12], gq2[12];
    for (int a = 0; a < 12; ++a) gq1[a] = gq2[a] = 0.0;
    for (int a = 0; a < k; ++a)
        for (int c = 0; c < j; ++c) {
            double coef = 0.0;
            if ((int)((args >> (2 * a)) & 3u) == c) coef += cw / (n * k);
            if ((int)((args >> (8 + 2 * c)) & 3u) == a) coef += cw / (n * j);
            if (coef == 0.0) continue;
            const double dWdD = exp(-(double)D[a * 4 + c] / (2.0 * med)) / (2.0 * med);
            for (int x = 0; x < 3; ++x) {
                const double gv = coef * dWdD * 2.0 * ((double)Q[a * 3 + x] - (double)Q[12 + c * 3 + x]);
                gq1[a * 3 + x] += gv;
                gq2[c * 3 + x] -= gv;
            }
        }
    if (g1) {
        float *G = g1 + (long long)b * g.nf1 * 9;
        for (int a = 0; a < k; ++a) {
            const int f = ws.recIdx[r * 8 + a];
            for (int p = 0; p < 3; ++p)
                for (int x = 0; x < 3; ++x)
                    atomicAdd(G + (long long)f * 9 + p * 3 + x, (float)((double)Wt[a * 3 + p] / 3.0 * gq1[a * 3 + x]));
        }
    }
    if (g2) {
        float *G = g2 + (long long)b * g.nf2 * 9;
        for (int c = 0; c < j; ++c) {
            const int f = ws.recIdx[r * 8 + 4 + c];
            for (int p = 0; p < 3; ++p)
                for (int x = 0; x < 3; ++x)
                    atomicAdd(G + (long long)f * 9 + p * 3 + x, (float)((double)Wt[12 + c * 3 + p] / 3.0 * gq2[c * 3 + x]));
        }
    }
}

int launch_backward(const Workspace &ws, const Geometry &g, const float *grad_out, float *g1, float *g2, cudaStream_t s) {
    if (g1 && cudaMemsetAsync(g1, 0, sizeof(float) * 9 * (size_t)g.B * g.nf1, s) != cudaSuccess) return RRL_ERR_CUDA;
    if (g2 && cudaMemsetAsync(g2, 0, sizeof(float) * 9 * (size_t)g.B * g.nf2, s) != cudaSuccess) return RRL_ERR_CUDA;
    dim3 grid((g.nl + 127) / 128, g.B);
    backward_kernel<<<grid, 128, 0, s>>>(ws, g, grad_out, g1, g2);
    count_launch();
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
