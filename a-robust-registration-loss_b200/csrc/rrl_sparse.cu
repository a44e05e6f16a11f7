// Sparse phase of the intersected-line loss on sm_100a: everything after the per-line intersection sets.
//
//   build     per line with (k, j) hits inside the window: sort the <=4 hit indices ascending (nonzero() order,
//             loss.py:125-131), recompute the three exact distances of every hit triplet, weights
//             w = d / ((d0+d1)+d2) (loss.py:92), intersection points q = ((w0 p0 + w1 p1) + w2 p2) / 3
//             (loss.py:155-163) and the k x j squared distances (loss.py:38-52,165-166); one record per selected line,
//             its valid D entries also appended to the pair's compact list.
//   median    exact lower median (torch.median, loss.py:223-224) of all D entries of a pair by range-refining selection
//             on the float bit patterns (2^11 buckets over the occupied range per sweep).
//   welsch    W = 1 - exp(-(D/med)/2) (loss.py:20-21,226), row/column minima with first-index tie breaking (torch.min),
//             per-(k,j) sums in 2^-40 fixed point (order independent), and the per-record gradient vectors
//             G1[a] = sum_b coef[a,b] dW/dD 2 (q1_a - q2_b), G2[b] = -sum_a ... (closed form).
//   finalize  loss = (1/C) sum_kj exp(-|k-j|/2) (S1/(n k) + S2/(n j))   (loss.py:215,227-230)
//   backward  d loss / d points: scatter (w/3) G grad_out to the hit triplets (vector reductions), or contracted to pose
//             space on the spot (backward_pose_kernel).
//
// Launches of the single-GPU forward: build (select + records) -> tail (median + Welsch + loss in one launch; the last
// block of a pair finalizes).  The line-sharded path runs build -> shard_tail (exchange over peer memory inside the
// kernel) or, as the fallback, the NCCL protocol's stage kernels with collectives in between and a separate finalize.
#include "rrl_common.cuh"

namespace rrl {

// ---- phase timestamps of block (0, 0) (measurement builds only: -DRRL_MARKS) ----
#ifdef RRL_MARKS
__device__ unsigned long long g_marks[32];
__device__ __forceinline__ void mark(int i) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_marks[i] = t;
    }
}
// start / end of every CTA of one kernel (build_kernel): shows waves, stragglers and the launch's lead-in
__device__ unsigned long long g_cta_t[2][8192];
__device__ __forceinline__ void cta_mark(int which) {
    if (threadIdx.x == 0) {
        const unsigned id = blockIdx.y * gridDim.x + blockIdx.x;
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (id < 8192) g_cta_t[which][id] = t;
    }
}
#else
__device__ __forceinline__ void mark(int) {}
__device__ __forceinline__ void cta_mark(int) {}
#endif

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

// Two lines per thread select; the block compacts its selected lines in shared memory (ordered) and claims a contiguous
// range of record slots with ONE atomic.  The records are then built by one thread per (record, cloud, hit slot) --
// each computes the weights and the intersection point of ONE hit triplet and drops them at the hit's rank among the
// line's hits (ascending triplet index = nonzero() order); the 8 threads of a record then form its 16 D entries.
constexpr int kBuildLines = 512;            // lines per CTA: two per thread (a 15000-line pair is then ONE wave of CTAs on 148 SMs)

__global__ void __launch_bounds__(256) build_kernel(const float *__restrict__ tri1, const float *__restrict__ tri2,
                                                    const float *__restrict__ lines, Workspace ws, Geometry g,
                                                    int k_lo, int j_lo, int k_hi, int j_hi) {
    __shared__ int s_line[kBuildLines], s_kj[kBuildLines], s_eoff[kBuildLines], s_warp[16], s_ewarp[16], s_hist[16], s_base, s_ebase;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    mark(8);
    cta_mark(0);
    if (tid < 16) s_hist[tid] = 0;
    int k[2], j[2], l[2];
    bool sel[2];
    unsigned bal[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {                              // both halves' counters in flight together
        l[u] = blockIdx.x * kBuildLines + u * 256 + tid;
        k[u] = 0; j[u] = 0;
        if (l[u] < g.nl) {
            const long long gl = (long long)b * g.nl + l[u];
            k[u] = ws.cnt[0][gl]; j[u] = ws.cnt[1][gl];
        }
    }
    // ordered compaction of the selected lines (ballots) and, in the same pass, the exclusive prefix of their k j: where a
    // record's valid D entries go in the pair's compact D list (what the tail's median reads: n floats in a row instead of a
    // masked gather over the padded records)
    int ecnt[2], eincl[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        sel[u] = l[u] < g.nl && k[u] >= k_lo && k[u] < k_hi && j[u] >= j_lo && j[u] < j_hi;      // windows are validated to lie inside 1..4
        bal[u] = __ballot_sync(0xffffffffu, sel[u]);
        ecnt[u] = sel[u] ? k[u] * j[u] : 0;
        eincl[u] = ecnt[u];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, eincl[u], d);
            if (lane >= d) eincl[u] += up;
        }
        if (lane == 31) { s_warp[u * 8 + wid] = __popc(bal[u]); s_ewarp[u * 8 + wid] = eincl[u]; }
    }
    __syncthreads();
    int before[2] = {0, 0}, total = 0, ebefore[2] = {0, 0}, etotal = 0;
#pragma unroll
    for (int w = 0; w < 16; ++w) {
        const int c = s_warp[w], e = s_ewarp[w];
        before[0] += w < wid ? c : 0;
        before[1] += w < 8 + wid ? c : 0;
        total += c;
        ebefore[0] += w < wid ? e : 0;
        ebefore[1] += w < 8 + wid ? e : 0;
        etotal += e;
    }
    mark(9);
    if (total == 0) { cta_mark(1); return; }
#pragma unroll
    for (int u = 0; u < 2; ++u)
        if (sel[u]) {                                          // ordered by line index: half 0, then half 1
            const int pos = before[u] + __popc(bal[u] & ((1u << lane) - 1u));
            s_line[pos] = l[u];
            s_kj[pos] = k[u] | (j[u] << 8);
            s_eoff[pos] = ebefore[u] + eincl[u] - ecnt[u];
            atomicAdd(&s_hist[(k[u] - 1) * 4 + (j[u] - 1)], 1);
        }
    if (tid == 0) s_base = atomicAdd(ws.nrec + b, total);
    if (tid == 32) s_ebase = atomicAdd(ws.flags + b * 4 + 2, etotal);
    __syncthreads();
    mark(10);
    if (tid < 16 && s_hist[tid]) atomicAdd(ws.n_kj + b * 16 + tid, s_hist[tid]);
    const long long r0 = (long long)b * g.nl + s_base;
    float *dflat = ws.dflat + (long long)b * kMedCache;
    // one thread per (record, cloud, hit slot): the 8 threads of a record are consecutive lanes and exchange their
    // intersection points with shuffles for the record's 16 D entries (two per lane) -- no shared-memory staging, no barrier
    const int nt = (total * 8 + 31) & ~31;                     // whole warps (full-mask shuffles)
    for (int t = tid; t < nt; t += 256) {
        const bool live = t < total * 8;
        const int rec = live ? t >> 3 : 0, cloud = (t >> 2) & 1, a = t & 3;
        const int kj = s_kj[rec];
        const int kk = kj & 255, jj = (kj >> 8) & 255;
        const int cnt = cloud ? jj : kk;
        const long long gl = (long long)b * g.nl + s_line[rec];
        float w[3] = {0.f, 0.f, 0.f}, q[3] = {0.f, 0.f, 0.f};
        int idx = -1, pos = a;                       // unused slots a >= cnt keep their place behind the hits
        if (live && a < cnt) {
            const int *h = ws.hits[cloud] + gl * kCap;
            idx = h[a];
            pos = 0;
#pragma unroll
            for (int o = 0; o < 4; ++o)
                if (o < cnt && o != a) pos += h[o] < idx;
            float ln[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) ln[c] = __ldg(lines + gl * 6 + c);
            make_point(cloud ? tri2 + (long long)b * g.nf2 * 9 : tri1 + (long long)b * g.nf1 * 9, idx, ln, w, q);
        }
        const long long r = r0 + rec;
        if (live) {
            const int o3 = cloud * 12 + pos * 3;
            ws.recIdx[r * 8 + cloud * 4 + pos] = idx;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                ws.recW[r * 24 + o3 + c] = w[c];
                ws.recQ[r * 24 + o3 + c] = q[c];
            }
        }
        // points into rank order: lane (cloud, a) fetches the point whose rank is a (pos is a permutation of 0..3 per cloud)
        const int gbase = lane & ~7, cbase = gbase + cloud * 4;
        int src = cbase;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int po = __shfl_sync(0xffffffffu, pos, cbase + o);
            if (po == a) src = cbase + o;
        }
        const float sx = __shfl_sync(0xffffffffu, q[0], src), sy = __shfl_sync(0xffffffffu, q[1], src), sz = __shfl_sync(0xffffffffu, q[2], src);
        // D entries e = 2 jl, 2 jl + 1 of the record (jl = lane within the record): a' = e >> 2 from cloud 1, c' = e & 3 from cloud 2
        const int jl = lane & 7, e0 = 2 * jl, ap = e0 >> 2, c0 = e0 & 3;
        const float ax = __shfl_sync(0xffffffffu, sx, gbase + ap), ay = __shfl_sync(0xffffffffu, sy, gbase + ap), az = __shfl_sync(0xffffffffu, sz, gbase + ap);
        const float bx0 = __shfl_sync(0xffffffffu, sx, gbase + 4 + c0), by0 = __shfl_sync(0xffffffffu, sy, gbase + 4 + c0), bz0 = __shfl_sync(0xffffffffu, sz, gbase + 4 + c0);
        const float bx1 = __shfl_sync(0xffffffffu, sx, gbase + 5 + c0), by1 = __shfl_sync(0xffffffffu, sy, gbase + 5 + c0), bz1 = __shfl_sync(0xffffffffu, sz, gbase + 5 + c0);
        if (live) {
            float d0 = 0.f, d1 = 0.f;
            if (ap < kk && c0 < jj) d0 = sq3_rn(__fsub_rn(ax, bx0), __fsub_rn(ay, by0), __fsub_rn(az, bz0));
            if (ap < kk && c0 + 1 < jj) d1 = sq3_rn(__fsub_rn(ax, bx1), __fsub_rn(ay, by1), __fsub_rn(az, bz1));
            *reinterpret_cast<float2 *>(ws.recD + r * 16 + e0) = make_float2(d0, d1);
            const int ei = s_ebase + s_eoff[rec] + ap * jj + c0;            // valid entries of a record, row major
            if (ap < kk && c0 < jj && ei < kMedCache) dflat[ei] = d0;
            if (ap < kk && c0 + 1 < jj && ei + 1 < kMedCache) dflat[ei + 1] = d1;
        }
    }
    mark(11);
    for (int t = tid; t < total; t += 256) reinterpret_cast<int2 *>(ws.recMeta)[r0 + t] = make_int2(s_line[t], s_kj[t]);
    mark(12);
    cta_mark(1);
}

int launch_build(const float *tri1, const float *tri2, const float *lines, const Workspace &ws, const Geometry &g,
                 int k_lo, int j_lo, int k_hi, int j_hi, cudaStream_t s) {
    stage_mark(5, s);
    build_kernel<<<dim3((g.nl + kBuildLines - 1) / kBuildLines, g.B), 256, 0, s>>>(tri1, tri2, lines, ws, g, k_lo, j_lo, k_hi, j_hi);
    count_launch();
    stage_mark(6, s);
    return check_launch();
}

// local counts -> gcounts (B,18): n_kj[16], #records, #D entries.  In the single-GPU path these ARE the global counts.
__global__ void local_counts_kernel(Workspace ws, Geometry g) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= g.B) return;
    if (!hdr_ok(ws, g)) {                                     // no forward of this geometry here: nothing selected
        for (int c = 0; c < 18; ++c) ws.gcounts[b * 18 + c] = 0;
        return;
    }
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
    local_counts_kernel<<<(g.B + 127) / 128, 128, 0, s>>>(ws, g);
    count_launch();
    return check_launch();
}

// ------------------------------------------------------------------------------------------------------
// exact lower median by range-refining selection (non-negative floats order like their bit patterns)
// ------------------------------------------------------------------------------------------------------
// Each round histograms the keys inside the current range [lo, hi] into 2^11 equal-width buckets, finds the bucket
// that holds the wanted rank with a block-wide scan, and narrows the range to it: 31 -> 20 -> 9 -> 0 bits of width,
// i.e. at most three or four sweeps.  Because the buckets subdivide the OCCUPIED range rather than fixed bit fields,
// D values that share an exponent do not pile into a handful of bins, so plain shared-memory atomics suffice.
constexpr int kSelBits = 11;
constexpr int kSelBins = 1 << kSelBits;
constexpr unsigned kNoKey = 0xFFFFFFFFu;     // "not a D entry" (D >= 0 never has that bit pattern)

struct SelectScratch {
    unsigned hist[kSelBins];
    unsigned wtot[32];
    unsigned bin, kmin, kmax;
    long long rank;
};

// `key(i)` enumerates `slots` slots (kNoKey = not valid), n > 0 of which are valid and lie in [kmin, kmax]; returns the
// key of rank (n-1)/2.  Block-wide (kThreads threads), every thread returns the same value.
template <int kThreads, typename KeyFn>
__device__ unsigned select_lower_median(long long slots, long long n, KeyFn key, unsigned kmin, unsigned kmax, SelectScratch &sc) {
    constexpr int kPer = kSelBins / kThreads;
    static_assert(kPer * kThreads == kSelBins && kPer >= 1 && kThreads <= 1024, "one thread owns kPer consecutive bins");
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned lo = kmin, hi = kmax;
    long long rank = (n - 1) / 2;
    for (;;) {
        const unsigned width = hi - lo;
        if (width == 0) return lo;
        const int s = max(0, (32 - __clz(width)) - kSelBits);         // (width >> s) < kSelBins
        for (int i = tid; i < kSelBins; i += kThreads) sc.hist[i] = 0;
        __syncthreads();
        for (long long i0 = 0; i0 < slots; i0 += 4 * kThreads) {      // four independent loads in flight per thread
            unsigned kb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long i = i0 + u * kThreads + tid;
                kb[u] = i < slots ? key(i) : kNoKey;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (kb[u] != kNoKey && kb[u] >= lo && kb[u] <= hi) atomicAdd(&sc.hist[(kb[u] - lo) >> s], 1u);
        }
        __syncthreads();
        unsigned h[kPer], tot = 0;
#pragma unroll
        for (int q = 0; q < kPer; ++q) { h[q] = sc.hist[tid * kPer + q]; tot += h[q]; }
        unsigned inc = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += up;
        }
        if (lane == 31) sc.wtot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const unsigned v = lane < kThreads / 32 ? sc.wtot[lane] : 0u;
            unsigned incw = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned up = __shfl_up_sync(0xffffffffu, incw, d);
                if (lane >= d) incw += up;
            }
            sc.wtot[lane] = incw - v;
        }
        __syncthreads();
        const long long before = (long long)sc.wtot[wid] + (long long)(inc - tot);
        if (tot && rank >= before && rank < before + (long long)tot) {
            long long r = rank - before;
            int q = 0;
            for (; q < kPer - 1; ++q) {
                if (r < (long long)h[q]) break;
                r -= h[q];
            }
            sc.bin = (unsigned)(tid * kPer + q);
            sc.rank = r;
        }
        __syncthreads();
        rank = sc.rank;
        const unsigned long long nlo = (unsigned long long)lo + ((unsigned long long)sc.bin << s);
        const unsigned long long nhi = nlo + ((1ull << s) - 1ull);
        lo = (unsigned)nlo;
        if (nhi < (unsigned long long)hi) hi = (unsigned)nhi;
        if (s == 0) return lo;
        __syncthreads();                               // sc.bin / sc.rank are rewritten in the next round
    }
}

// block-wide min / max of the valid keys a thread has seen -> sc.kmin / sc.kmax (initialised by the caller)
__device__ __forceinline__ void block_minmax(unsigned mn, unsigned mx, SelectScratch &sc) {
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0) {
        if (mn != kNoKey) atomicMin(&sc.kmin, mn);
        atomicMax(&sc.kmax, mx);
    }
}

__global__ void __launch_bounds__(1024) select_flat_kernel(const float *__restrict__ vals, long long n, float *out) {
    __shared__ SelectScratch sc;
    if (n <= 0) {
        if (threadIdx.x == 0) *out = 0.f;
        return;
    }
    if (threadIdx.x == 0) { sc.kmin = kNoKey; sc.kmax = 0u; }
    __syncthreads();
    unsigned mn = kNoKey, mx = 0u;
    for (long long i = threadIdx.x; i < n; i += 1024) {
        const unsigned kb = __float_as_uint(vals[i]);
        mn = min(mn, kb); mx = max(mx, kb);
    }
    block_minmax(mn, mx, sc);
    __syncthreads();
    auto key = [&](long long i) -> unsigned { return __float_as_uint(vals[i]); };
    const unsigned bits = select_lower_median<1024>(n, n, key, sc.kmin, sc.kmax, sc);
    if (threadIdx.x == 0) *out = __uint_as_float(bits);
}

int launch_select_median(const float *vals, long long n, float *out, cudaStream_t s) {
    select_flat_kernel<<<1, 1024, 0, s>>>(vals, n, out);
    count_launch();
    return check_launch();
}

// ---- distributed lower median (line-sharded path): two rounds of 16-bit histograms ------------------------------
// Every rank histograms the keys of ITS records (round 0: the high 16 bits of the float's bit pattern; round 1: the low
// 16 bits of the keys that fall into the bin round 0 chose), the histograms are summed across ranks by the caller
// (all-reduce), and every rank picks the bin that holds the global rank from the summed histogram -- identical inputs,
// identical decision, no host round trip and no variable-length exchange.
// state[0] = key prefix chosen so far, state[1] = rank of the median inside that prefix.
constexpr int kShardBins = 65536;

__global__ void __launch_bounds__(256) shard_hist_kernel(Workspace ws, Geometry g, int round, const long long *__restrict__ state,
                                                          int *__restrict__ hist) {
    if (!hdr_ok(ws, g)) return;
    const long long slots = (long long)ws.nrec[0] * 16;
    const unsigned hi = (unsigned)state[0] >> 16;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += (long long)gridDim.x * blockDim.x) {
        const int kj = ws.recMeta[(i >> 4) * 2 + 1];
        const int e = (int)(i & 15);
        if ((e >> 2) < (kj & 255) && (e & 3) < ((kj >> 8) & 255)) {
            const unsigned key = __float_as_uint(ws.recD[i]);
            if (round == 0) atomicAdd(hist + (key >> 16), 1);
            else if ((key >> 16) == hi) atomicAdd(hist + (key & 0xFFFFu), 1);
        }
    }
}

__global__ void __launch_bounds__(1024) shard_pick_kernel(int round, const int *__restrict__ hist, const long long *__restrict__ gcounts18,
                                                           long long *state, float *out_median) {
    __shared__ long long s_w[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long n = gcounts18[17];
    if (n <= 0) {                                                // no entry anywhere: median 0, like the single-GPU path
        if (tid == 0) { state[0] = 0; state[1] = 0; if (round == 1) *out_median = 0.f; }
        return;
    }
    const long long rank = round == 0 ? (n - 1) / 2 : state[1];
    constexpr int kPer = kShardBins / 1024;
    long long tot = 0;
    for (int q = 0; q < kPer; ++q) tot += hist[tid * kPer + q];
    long long inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long up = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += up;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const long long v = s_w[lane];
        long long incw = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long up = __shfl_up_sync(0xffffffffu, incw, d);
            if (lane >= d) incw += up;
        }
        s_w[lane] = incw - v;
    }
    __syncthreads();
    const long long before = s_w[wid] + inc - tot;
    if (tot > 0 && rank >= before && rank < before + tot) {      // exactly one thread
        long long r = rank - before;
        int q = 0;
        for (; q < kPer - 1; ++q) {
            const long long h = hist[tid * kPer + q];
            if (r < h) break;
            r -= h;
        }
        const unsigned bin = (unsigned)(tid * kPer + q);
        const unsigned prefix = round == 0 ? bin << 16 : (((unsigned)state[0] & 0xFFFF0000u) | bin);
        state[0] = prefix;
        state[1] = r;
        if (round == 1) *out_median = __uint_as_float(prefix);
    }
}

int launch_shard_hist(const Workspace &ws, const Geometry &g, int round, const long long *state, int *hist, cudaStream_t s) {
    if (cudaMemsetAsync(hist, 0, sizeof(int) * kShardBins, s) != cudaSuccess) return RRL_ERR_CUDA;
    shard_hist_kernel<<<sm_count() * 2, 256, 0, s>>>(ws, g, round, state, hist);
    count_launch();
    return check_launch();
}

int launch_shard_pick(int round, const int *hist, const long long *gcounts18, long long *state, float *out_median, cudaStream_t s) {
    shard_pick_kernel<<<1, 1024, 0, s>>>(round, hist, gcounts18, state, out_median);
    count_launch();
    return check_launch();
}

// compact the valid D entries of pair 0 (line-sharded path)
__global__ void pack_entries_kernel(Workspace ws, Geometry g, float *out, long long cap) {
    if (!hdr_ok(ws, g)) return;
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
// e^x for x <= 0 (or NaN) in double, branch free: Cody-Waite reduction x = n ln2 + r, |r| <= ln2/2, degree-13 Taylor polynomial
// (truncation < 2^-57), scaling by exponent bits.  Arguments below -120 are clamped: the result is far below half the
// smallest float subnormal either way.  Error <= 2 ulp of a double.
__device__ __forceinline__ double exp_nonpos(double x) {
    x = x < -120.0 ? -120.0 : x;                                   // NaN stays NaN
    const double t = rint(x * 1.4426950408889634074);
    double r = fma(t, -6.93147180369123816490e-01, x);
    r = fma(t, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821614599e-10;                          // 1/13!
    p = fma(p, r, 2.0876756987868098979e-09);
    p = fma(p, r, 2.5052108385441718775e-08);
    p = fma(p, r, 2.7557319223985890653e-07);
    p = fma(p, r, 2.7557319223985892511e-06);
    p = fma(p, r, 2.4801587301587301566e-05);
    p = fma(p, r, 1.9841269841269841253e-04);
    p = fma(p, r, 1.3888888888888889419e-03);
    p = fma(p, r, 8.3333333333333332177e-03);
    p = fma(p, r, 4.1666666666666664354e-02);
    p = fma(p, r, 1.6666666666666665741e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int n = __double2int_rn(t);                              // in [-174, 0]; 0 for NaN (p is NaN then)
    return p * __longlong_as_double((long long)(1023 + n) << 52);
}

// exp(-((x / c)) / 2.0) of loss.py:21 evaluated in double on the float argument: W = 1 - (float)e is then a well-defined
// (correctly rounded up to the 2^-29 chance of a double-rounding tie) float exp, so that the row/column argmin of
// near-tied entries does not depend on a vendor expf's last bit (the oracle does the same); the double e also serves
// the gradient, dW/dD = e / (2 med).
__device__ __forceinline__ double welsch_exp(float D, float med) {
    return exp_nonpos((double)(-__fdiv_rn(__fdiv_rn(D, med), 2.0f)));
}

// The Welsch stage of 32 records, one per lane of a CONVERGED warp (lanes without a record pass valid = false).
//   * a record has k*j <= 16 entries (3.3 on average) and a double-precision exp is a long dependent chain, so the exps
//     are evaluated on a dense list: the lanes drop their valid D entries into the warp's shared-memory buffer `buf`
//     (>= 512 floats), the warp evaluates the list 32 entries at a time, and every lane reads its (float) exps back.
//     (One thread per record running its own k*j exps measured 4x slower: 16 predicated sites per warp.)
//   * minima with first-index tie breaking (torch.min) -> v1 = sum_a min_b W, v2 = sum_b min_a W as 2^-40 fixed point
//     (W * 2^40 is an exact float product; integer sums are exact and order independent);
//   * gradient vectors for a unit upstream gradient, in float (relative error ~1e-7, bar 1e-5): written to recG.
// cw_over_n = exp(-|k-j|/2) / C / n_kj.
__device__ __forceinline__ void welsch_warp(const Workspace &ws, bool valid, long long r, float med, float cw_over_n, int k, int j,
                                            float *buf, unsigned long long &v1, unsigned long long &v2, bool &nan) {
    const int lane = threadIdx.x & 31;
    float D[16];
#pragma unroll
    for (int a = 0; a < 16; ++a) D[a] = 0.f;
    if (valid) {
        const float4 *D4 = reinterpret_cast<const float4 *>(ws.recD + r * 16);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float4 d = D4[a];
            D[a * 4] = d.x; D[a * 4 + 1] = d.y; D[a * 4 + 2] = d.z; D[a * 4 + 3] = d.w;
        }
    }
    const int cnt = valid ? k * j : 0;
    int inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += up;
    }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    const int off = inc - cnt;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (valid && a < k && c < j) buf[off + a * j + c] = D[a * 4 + c];
    __syncwarp();
    for (int p = lane; p < total; p += 32) buf[p] = (float)welsch_exp(buf[p], med);
    __syncwarp();
    float W[16], E[16];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            E[a * 4 + c] = 0.f;
            W[a * 4 + c] = 0.f;
            if (valid && a < k && c < j) {
                E[a * 4 + c] = buf[off + a * j + c];
                W[a * 4 + c] = __fsub_rn(1.0f, E[a * 4 + c]);
            }
        }
    __syncwarp();                                                  // buf is reused by the caller's next batch
    int arg_b[4], arg_a[4];
    v1 = 0ull; v2 = 0ull;
    nan = false;
#pragma unroll
    for (int a = 0; a < 4; ++a) {                   // torch.min(W, 2): first index on ties
        int m = 0;
#pragma unroll
        for (int c = 1; c < 4; ++c) if (c < j && W[a * 4 + c] < W[a * 4 + m]) m = c;
        arg_b[a] = m;
        if (valid && a < k) {
            const float w = W[a * 4 + m];
            nan |= !(w == w);
            v1 += (unsigned long long)__float2ll_rn(w * (float)kFixScale);
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {                   // torch.min(W, 1)
        int m = 0;
#pragma unroll
        for (int a = 1; a < 4; ++a) if (a < k && W[a * 4 + c] < W[m * 4 + c]) m = a;
        arg_a[c] = m;
        if (valid && c < j) {
            const float w = W[m * 4 + c];
            nan |= !(w == w);
            v2 += (unsigned long long)__float2ll_rn(w * (float)kFixScale);
        }
    }
    if (nan) { v1 = 0ull; v2 = 0ull; }
    if (!valid) return;
    // d loss / d D[a,c] = coef[a,c] e / (2 med); the factor 2 of d D / d q cancels the 1/2
    const float4 *Q4 = reinterpret_cast<const float4 *>(ws.recQ + r * 24);
    float4 *G4 = reinterpret_cast<float4 *>(ws.recG + r * 24);
    float q[24];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        const float4 v = Q4[a];
        q[4 * a] = v.x; q[4 * a + 1] = v.y; q[4 * a + 2] = v.z; q[4 * a + 3] = v.w;
    }
    const float inv_med = 1.0f / med;
    const float ck = cw_over_n / (float)k * inv_med, cj = cw_over_n / (float)j * inv_med;
    float G[24];
#pragma unroll
    for (int x = 0; x < 24; ++x) G[x] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const bool live = a < k && c < j;
            const float coef = (live && arg_b[a] == c ? ck : 0.f) + (live && arg_a[c] == a ? cj : 0.f);
            const float f = coef * E[a * 4 + c];
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                const float gv = f * (q[a * 3 + x] - q[12 + c * 3 + x]);
                G[a * 3 + x] += gv;
                G[12 + c * 3 + x] -= gv;
            }
        }
#pragma unroll
    for (int a = 0; a < 6; ++a) G4[a] = make_float4(G[4 * a], G[4 * a + 1], G[4 * a + 2], G[4 * a + 3]);
}

// exp(-|k-j|/2) of loss.py:229 for |k-j| = 0..3
__device__ __forceinline__ float combo_weight(int k, int j) {
    const int d = abs(k - j);
    return d == 0 ? 1.0f : d == 1 ? 0.60653065971263342f : d == 2 ? 0.36787944117144233f : 0.22313016014842982f;
}

__device__ __forceinline__ int count_combos(const long long *gc) {
    int C = 0;
    for (int c = 0; c < 16; ++c) C += gc[c] > 0;
    return C;
}

// loss = (1/C) sum_kj exp(-|k-j|/2) (S1/(n k) + S2/(n j))   (loss.py:215,227-230).  Called by ONE FULL WARP: lane c < 16
// owns combo c (the per-lane loads would otherwise be a chain of L2 round trips in a single thread); the terms are
// summed in a fixed butterfly order, so the result is deterministic.
__device__ __forceinline__ void finalize_pair(const Workspace &ws, int b, float *out_loss, int *out_status, float *out_median,
                                              long long *out_stats, int extra_status = 0) {
    const int lane = threadIdx.x & 31;
    const int c = lane & 15, k = (c >> 2) + 1, j = (c & 3) + 1;
    const long long n = __ldcg(ws.gcounts + b * 18 + c);
    const long long extra = __ldcg(ws.gcounts + b * 18 + 16 + (lane & 1));           // lane 0: #records, lane 1: #D entries
    const unsigned long long S1b = __ldcg(ws.sums + b * 32 + c), S2b = __ldcg(ws.sums + b * 32 + 16 + c);
    const int nanflag = __ldcg(ws.flags + b * 4);
    long long *st = ws.stats + (long long)b * RRL_NSTAT;
    const long long mystat = lane < RRL_NSTAT ? __ldcg(st + lane) : 0;
    const unsigned pm0 = __ldcg(ws.pmax + b * 2), pm1 = __ldcg(ws.pmax + b * 2 + 1);
    const float med = __ldcg(ws.med + b);
    double term = 0.0;
    const bool live = lane < 16 && n > 0;
    if (live) {
        const double S1 = (double)S1b / kFixScale, S2 = (double)S2b / kFixScale;
        term = exp(-0.5 * (double)abs(k - j)) * (S1 / ((double)n * k) + S2 / ((double)n * j));
    }
    const int C = __popc(__ballot_sync(0xffffffffu, live));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) term += __shfl_xor_sync(0xffffffffu, term, d);
    double loss = term;
    const long long nrec = __shfl_sync(0xffffffffu, extra, 0), nD = __shfl_sync(0xffffffffu, extra, 1);
    const long long nan_count = __shfl_sync(0xffffffffu, mystat, 6);
    int status = extra_status;
    if (C == 0) status |= RRL_STATUS_EMPTY; else loss /= (double)C;
    if (nan_count > 0) status |= RRL_STATUS_NAN;
    if (nanflag) { status |= RRL_STATUS_NAN; loss = __longlong_as_double(0x7ff8000000000000LL); }
    // |AC|^2 <= (P + X)^2: flag clouds whose own extent already makes the 2e-4 offset smaller than a few ulps of
    // |p|^2 (|AC|^2 >~ 1e3)
    if (fmaxf(__uint_as_float(pm0), __uint_as_float(pm1)) > 250.0f) status |= RRL_STATUS_NAN_RISK;
    // a NaN or infinite coordinate in a cloud or a line (prep records it as an infinite extent) gives NaN distances in the
    // reference's dense tensor, which is what makes it exit (loss.py:89-91).  Here such triplets and lines never hit
    // (comparisons with NaN are false) and may never reach the exact test, so the inputs themselves raise the flag.
    if (__ldcg(ws.bad + b * 2) | __ldcg(ws.bad + b * 2 + 1)) status |= RRL_STATUS_NAN;
    // what RRL_REUSE_TARGET needs of cloud 2 in the next forward (the per-pair block is zeroed by its prep stage)
    if (lane == 0) {
        unsigned int *kp = ws.keep + b * 8;
        kp[0] = pm1;
        kp[1] = __ldcg(ws.rmax + b * 2 + 1);
        kp[2] = __ldcg(ws.smax + b * 2 + 1);
        kp[3] = __ldcg(ws.bad + b * 2 + 1);
        kp[5] = __ldcg(ws.tmax + b * 2 + 1);
        // the line extent cloud 2's records are valid for: xmax[1] when the node stage ran under RRL_REUSE_TARGET (it carries the
        // kept or the freshly inflated value), else the extent of this forward's own lines
        const unsigned built = __ldcg(ws.xmax + b * 2 + 1);
        kp[4] = built ? built : __ldcg(ws.xmax + b * 2);
    }
    const long long outstat = lane == 0 ? nrec : lane == 1 ? nD : lane == 2 ? (long long)C : mystat;
    if (lane < 3) st[lane] = outstat;
    if (out_stats && lane < RRL_NSTAT) out_stats[(long long)b * RRL_NSTAT + lane] = outstat;
    if (lane == 0) {
        out_loss[b] = (float)loss;
        if (out_status) out_status[b] = status;
        if (out_median) out_median[b] = med;
    }
}

// Adds the (s1, s2) of every valid lane to the block's per-combo fixed-point sums with ONE shared atomic per (warp,
// combo): a 64-bit shared atomicAdd is a compare-and-swap loop, and hundreds of threads hammering sixteen addresses
// serialise badly.  Lanes of equal combo are summed with warp reductions on 21-bit limbs (32 * 2^21 < 2^32).
// Must be called by converged warps.  Integer sums: exact and order independent.
__device__ __forceinline__ void warp_combo_add(unsigned long long *s_sum, bool valid, int combo, unsigned long long v1,
                                               unsigned long long v2) {
    if (!valid) { v1 = 0ull; v2 = 0ull; }                            // each <= 4 * 2^40
    const int lane = threadIdx.x & 31;
    unsigned remaining = __ballot_sync(0xffffffffu, valid);
    while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const int cb = __shfl_sync(0xffffffffu, combo, leader);
        const bool mine = valid && combo == cb;
        unsigned long long t1 = 0ull, t2 = 0ull;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const unsigned a1 = mine ? (unsigned)((v1 >> (21 * q)) & 0x1FFFFFu) : 0u;
            const unsigned a2 = mine ? (unsigned)((v2 >> (21 * q)) & 0x1FFFFFu) : 0u;
            t1 += (unsigned long long)__reduce_add_sync(0xffffffffu, a1) << (21 * q);
            t2 += (unsigned long long)__reduce_add_sync(0xffffffffu, a2) << (21 * q);
        }
        if (lane == leader) {
            atomicAdd(&s_sum[cb], t1);
            atomicAdd(&s_sum[16 + cb], t2);
        }
        remaining &= ~__ballot_sync(0xffffffffu, mine);
    }
}

// gcounts / med hold the GLOBAL values (rrl_shard_stage2) in the line-sharded path.  With out_loss != nullptr
// (single-GPU forward) the last block of a pair to finish also writes the loss (ticket in flags[b*2+1]).
__global__ void __launch_bounds__(256) welsch_kernel(Workspace ws, Geometry g, float *out_loss, int *out_status, float *out_median,
                                                     long long *out_stats) {
    __shared__ unsigned long long s_sum[32];
    __shared__ float s_buf[8 * 512];
    __shared__ int s_last;
    const int b = blockIdx.y;
    if (!hdr_ok(ws, g)) return;
    if (threadIdx.x < 32) s_sum[threadIdx.x] = 0ull;
    if (blockIdx.x == 0 && b == 0 && threadIdx.x == 0) { ws.hdr[6] = 1; ws.hdr[7] = order_token(g); }   // gradient vectors in place; order complete
    __syncthreads();
    const long long nrec = ws.nrec[b];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if ((long long)blockIdx.x * blockDim.x >= nrec && blockIdx.x != 0) return;      // whole block beyond the records
    {
        const bool valid = i < nrec;
        const long long r = (long long)b * g.nl + i;
        int combo = 0, k = 1, j = 1;
        float cw = 0.f;
        if (valid) {
            const long long *gc = ws.gcounts + b * 18;
            const int kj = ws.recMeta[r * 2 + 1];
            k = kj & 255; j = (kj >> 8) & 255;
            combo = (k - 1) * 4 + (j - 1);
            cw = combo_weight(k, j) / (float)count_combos(gc) / (float)gc[combo];
        }
        unsigned long long v1, v2;
        bool nan;
        welsch_warp(ws, valid, r, ws.med[b], cw, k, j, s_buf + (threadIdx.x >> 5) * 512, v1, v2, nan);
        if (nan) ws.flags[b * 4] = 1;                               // e.g. median 0: the reference's loss is NaN too
        warp_combo_add(s_sum, valid, combo, v1, v2);
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
    if (s_last && threadIdx.x < 32) {
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

// Single-GPU tail of the forward: median -> Welsch -> loss in ONE launch.  grid (S, B): every one of the S blocks of a
// pair recomputes the pair's median from the D entries (a few sweeps over <= 192 KB of shared memory; cheaper than a
// second launch and a grid-wide dependency), then takes every S-th slice of the records through the Welsch stage; the
// last block to finish writes the loss (ticket in flags[b*4+1]).
#ifndef RRL_TAIL_THREADS
#define RRL_TAIL_THREADS 512
#endif
constexpr int kTailThreads = RRL_TAIL_THREADS;

__global__ void __launch_bounds__(kTailThreads) tail_kernel(Workspace ws, Geometry g, float *out_loss, int *out_status,
                                                             float *out_median, long long *out_stats) {
    extern __shared__ unsigned s_keys[];     // [kMedCache]
    __shared__ SelectScratch sc;
    __shared__ unsigned long long s_sum[32];
    __shared__ long long s_gc[16];
    __shared__ int s_last;
    const int b = blockIdx.y, tid = threadIdx.x;
    mark(0);
    if (blockIdx.x == 0 && b == 0 && tid == 0) { ws.hdr[6] = 1; ws.hdr[7] = order_token(g); }   // gradient vectors in place; order complete
    const int nrec = ws.nrec[b];
    if (tid < 16) s_gc[tid] = ws.n_kj[b * 16 + tid];
    if (tid < 32) s_sum[tid] = 0ull;
    if (tid == 0) { sc.kmin = kNoKey; sc.kmax = 0u; }
    __syncthreads();
    long long n = 0;
    int C = 0;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        n += s_gc[c] * ((c >> 2) + 1) * ((c & 3) + 1);
        C += s_gc[c] > 0;
    }
    if (blockIdx.x == 0) {                                     // single GPU: the local counts ARE the global counts
        if (tid < 16) ws.gcounts[b * 18 + tid] = s_gc[tid];
        if (tid == 16) ws.gcounts[b * 18 + 16] = nrec;
        if (tid == 17) ws.gcounts[b * 18 + 17] = n;
    }
    float med = 0.f;
    if (n > 0) {
        const float *D = ws.recD + (long long)b * g.nl * 16;
        const int *meta = ws.recMeta + (long long)b * g.nl * 2;
        const long long slots = (long long)nrec * 16;
        auto gkey = [&](long long i) -> unsigned {
            const int kj = meta[(i >> 4) * 2 + 1];
            const int e = (int)(i & 15);
            const bool valid = (e >> 2) < (kj & 255) && (e & 3) < ((kj >> 8) & 255);
            return valid ? __float_as_uint(D[i]) : kNoKey;
        };
        unsigned mn = kNoKey, mx = 0u;
        if (n <= kMedCache) {
            // the pair's valid D entries, written compactly by the build stage (their order is irrelevant to a median)
            const float *df = ws.dflat + (long long)b * kMedCache;
            for (int i = tid; i < (int)n; i += kTailThreads) {
                const unsigned kb = __float_as_uint(df[i]);
                s_keys[i] = kb;
                mn = min(mn, kb); mx = max(mx, kb);
            }
            block_minmax(mn, mx, sc);
            __syncthreads();
            mark(1);
            auto key = [&](long long i) -> unsigned { return s_keys[i]; };
            med = __uint_as_float(select_lower_median<kTailThreads>(n, n, key, sc.kmin, sc.kmax, sc));
        } else {
            // more entries than the cache holds: every sweep re-reads the records through L2
            for (long long i = tid; i < slots; i += kTailThreads) {
                const unsigned kb = gkey(i);
                if (kb != kNoKey) { mn = min(mn, kb); mx = max(mx, kb); }
            }
            block_minmax(mn, mx, sc);
            __syncthreads();
            mark(1);
            med = __uint_as_float(select_lower_median<kTailThreads>(slots, n, gkey, sc.kmin, sc.kmax, sc));
        }
    }
    if (blockIdx.x == 0 && tid == 0) ws.med[b] = med;
    mark(2);
    float *wbuf = reinterpret_cast<float *>(s_keys) + (tid >> 5) * 512;          // the key cache is free after the select
    for (long long i0 = (long long)blockIdx.x * kTailThreads; i0 < nrec; i0 += (long long)gridDim.x * kTailThreads) {   // block-uniform
        const long long i = i0 + tid;
        const bool valid = i < nrec;
        const long long r = (long long)b * g.nl + i;
        int combo = 0, k = 1, j = 1;
        float cw = 0.f;
        if (valid) {
            const int kj = ws.recMeta[r * 2 + 1];
            k = kj & 255; j = (kj >> 8) & 255;
            combo = (k - 1) * 4 + (j - 1);
            cw = combo_weight(k, j) / (float)C / (float)s_gc[combo];
        }
        unsigned long long v1, v2;
        bool nan;
        welsch_warp(ws, valid, r, med, cw, k, j, wbuf, v1, v2, nan);
        if (nan) ws.flags[b * 4] = 1;                               // e.g. median 0: the reference's loss is NaN too
        warp_combo_add(s_sum, valid, combo, v1, v2);
    }
    __syncthreads();
    mark(3);
    if (gridDim.x == 1) {                                      // the pair's only block: no cross-block hand-off
        if (tid < 32) ws.sums[b * 32 + tid] = s_sum[tid];
        if (tid == 0) s_last = 1;
    } else {
        if (tid < 32 && s_sum[tid]) atomicAdd(ws.sums + b * 32 + tid, s_sum[tid]);
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(ws.flags + b * 4 + 1, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && tid < 32) {
        __threadfence();
        finalize_pair(ws, b, out_loss, out_status, out_median, out_stats);
    }
    mark(4);
}

int launch_tail(const Workspace &ws, const Geometry &g, float *out_loss, int *out_status, float *out_median,
                long long *out_stats, cudaStream_t s) {
    static unsigned long long attr_mask = 0ull;
    if (ensure_dyn_smem(tail_kernel, kMedCache * 4, attr_mask)) return RRL_ERR_CUDA;
    stage_mark(7, s);
    int S = sm_count() / g.B;                       // one block per SM at most; small batches spread a pair's records wider
    if (S < 1) S = 1;
    if (S > 16) S = 16;
    tail_kernel<<<dim3(S, g.B), kTailThreads, kMedCache * 4, s>>>(ws, g, out_loss, out_status, out_median, out_stats);
    count_launch();
    stage_mark(8, s);
    return check_launch();
}

// ------------------------------------------------------------------------------------------------------
// Line shard: exchange + global median + Welsch + exchange + loss in ONE launch per rank (peer memory, no NCCL)
// ------------------------------------------------------------------------------------------------------
// What the NCCL protocol does with ~15 small launches and four collectives (counts, two 256 KB histograms, sums; every
// one of them a latency-bound round trip through the host-side enqueue): one CTA per rank
//   1. compacts the rank's valid D entries and its 18 counts into its slot and PUSHES them to every peer over NVLink
//      (comm_exchange, rrl_common.cuh); when all flags have arrived every rank holds every rank's entries;
//   2. selects the global lower median of ALL entries itself (same data, same deterministic selection on every rank:
//      bit-identical medians without a second round), adds up the global counts;
//   3. runs the Welsch stage on ITS records with the global median and normalisers;
//   4. pushes its 32 fixed-point partial sums (+ NaN / bad-input bits), sums everybody's (integers: exact, order
//      independent) and writes the loss -- identical on every rank.
// The workspace ends up as after a single-GPU forward with GLOBAL med / gcounts, so rrl_loss_backward works unchanged and
// yields the gradient share of this rank's lines.
constexpr int kShardThreads = 1024;
constexpr int kShardEntryOff = 160;          // payload A: 18 int64 counts, padded to a 16-byte boundary, then the float entries

__global__ void __launch_bounds__(kShardThreads) shard_tail_kernel(Workspace ws, Geometry g, CommView comm, float *out_loss,
                                                                    int *out_status, float *out_median, long long *out_stats) {
    extern __shared__ unsigned s_keys[];     // [kMedCache]
    __shared__ SelectScratch sc;
    __shared__ unsigned long long s_sum[32];
    __shared__ long long s_gc[18], s_pref[kCommMaxWorld + 1];
    __shared__ int s_count, s_extra;
    const int tid = threadIdx.x, lane = tid & 31;
    char *mine = comm.peer[comm.rank];
    const unsigned long long seq = *comm_epoch(mine) + 1ull;
    const bool have = hdr_ok(ws, g);          // a rank without a forward still takes part (with nothing), or its peers would wait
    const int nrec = have ? ws.nrec[0] : 0;
    const float *D = ws.recD;
    const int *meta = ws.recMeta;
    if (tid < 32) s_sum[tid] = 0ull;
    if (tid == 0) { sc.kmin = kNoKey; sc.kmax = 0u; s_count = 0; s_extra = 0; ws.hdr[6] = 1; ws.hdr[7] = order_token(g); }
    __syncthreads();
    // ---- 1. payload A: counts + compacted entries -> own slot ------------------------------------------------------
    char *slotA = comm_slot(comm, mine, seq, comm.rank);
    float *ent = reinterpret_cast<float *>(slotA + kShardEntryOff);
    const float4 *D4 = reinterpret_cast<const float4 *>(D);
    for (int i0 = 0; i0 < nrec; i0 += kShardThreads) {                       // block-uniform trip count
        const int i = i0 + tid;
        int kj = 0;
        float4 d[4];
        if (i < nrec) {
            kj = meta[i * 2 + 1];
#pragma unroll
            for (int a = 0; a < 4; ++a) d[a] = D4[(long long)i * 4 + a];
        }
        const int k = kj & 255, j = (kj >> 8) & 255;
        const int cnt = k * j;
        int inc = cnt;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, inc, dd);
            if (lane >= dd) inc += up;
        }
        const int total = __shfl_sync(0xffffffffu, inc, 31);
        int base = 0;
        if (lane == 0 && total) base = atomicAdd(&s_count, total);
        base = __shfl_sync(0xffffffffu, base, 0);
        int pos = base + inc - cnt;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float dv[4] = {d[a].x, d[a].y, d[a].z, d[a].w};
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (a < k && c < j) ent[pos++] = dv[c];
        }
    }
    __syncthreads();
    const int nD_local = s_count;
    if (tid < 18) {
        long long v = 0;
        if (tid < 16) v = have ? ws.n_kj[tid] : 0;
        else if (tid == 16) v = nrec;
        else v = nD_local;
        reinterpret_cast<long long *>(slotA)[tid] = v;
    }
    __threadfence();
    __syncthreads();
    bool ok = comm_exchange(comm, seq, (unsigned)((kShardEntryOff + 4 * nD_local + 15) & ~15));
    // ---- 2. global counts, global lower median ----------------------------------------------------------------------
    if (tid < 18) {
        long long v = 0;
        for (int r = 0; r < comm.world; ++r) v += __ldcg(reinterpret_cast<const long long *>(comm_slot(comm, mine, seq, r)) + tid);
        s_gc[tid] = v;
        ws.gcounts[tid] = v;
    }
    if (tid == 0) {
        long long acc = 0;
        for (int r = 0; r < comm.world; ++r) {
            s_pref[r] = acc;
            acc += __ldcg(reinterpret_cast<const long long *>(comm_slot(comm, mine, seq, r)) + 17);
        }
        s_pref[comm.world] = acc;
    }
    __syncthreads();
    const long long n = s_gc[17];
    int C = 0;
#pragma unroll
    for (int c = 0; c < 16; ++c) C += s_gc[c] > 0;
    float med = 0.f;
    if (n > 0 && ok) {
        auto gkey = [&](long long i) -> unsigned {                           // entry i of the concatenation over the senders
            int r = 0;
            while (r + 1 < comm.world && i >= s_pref[r + 1]) ++r;
            const float *e = reinterpret_cast<const float *>(comm_slot(comm, mine, seq, r) + kShardEntryOff);
            return __float_as_uint(__ldcg(e + (i - s_pref[r])));
        };
        unsigned mn = kNoKey, mx = 0u;
        if (n <= kMedCache) {
            for (long long i = tid; i < n; i += kShardThreads) {
                const unsigned kb = gkey(i);
                s_keys[i] = kb;
                mn = min(mn, kb); mx = max(mx, kb);
            }
            block_minmax(mn, mx, sc);
            __syncthreads();
            auto key = [&](long long i) -> unsigned { return s_keys[i]; };
            med = __uint_as_float(select_lower_median<kShardThreads>(n, n, key, sc.kmin, sc.kmax, sc));
        } else {
            for (long long i = tid; i < n; i += kShardThreads) {
                const unsigned kb = gkey(i);
                mn = min(mn, kb); mx = max(mx, kb);
            }
            block_minmax(mn, mx, sc);
            __syncthreads();
            med = __uint_as_float(select_lower_median<kShardThreads>(n, n, gkey, sc.kmin, sc.kmax, sc));
        }
    }
    if (tid == 0) ws.med[0] = med;
    __syncthreads();                                                         // the key cache becomes the Welsch buffers
    // ---- 3. Welsch stage on this rank's records ------------------------------------------------------------------------
    float *wbuf = reinterpret_cast<float *>(s_keys) + (tid >> 5) * 512;
    for (int i0 = 0; i0 < nrec; i0 += kShardThreads) {                       // block-uniform
        const long long i = i0 + tid;
        const bool valid = i < nrec;
        int combo = 0, k = 1, j = 1;
        float cw = 0.f;
        if (valid) {
            const int kj = meta[i * 2 + 1];
            k = kj & 255; j = (kj >> 8) & 255;
            combo = (k - 1) * 4 + (j - 1);
            cw = combo_weight(k, j) / (float)C / (float)s_gc[combo];
        }
        unsigned long long v1, v2;
        bool nan;
        welsch_warp(ws, valid, i, med, cw, k, j, wbuf, v1, v2, nan);
        if (nan) s_extra = 1;                                                // benign race: every writer stores 1
        warp_combo_add(s_sum, valid, combo, v1, v2);
    }
    __syncthreads();
    // ---- 4. payload B: partial sums + flags; total; loss -----------------------------------------------------------------
    unsigned long long *slotB = reinterpret_cast<unsigned long long *>(comm_slot(comm, mine, seq + 1ull, comm.rank));
    if (tid < 32) slotB[tid] = s_sum[tid];
    if (tid == 32) {
        const long long nan_cand = have ? ws.stats[6] : 0;
        const unsigned bad = have ? (ws.bad[0] | ws.bad[1]) : 0u;
        slotB[32] = (unsigned long long)(s_extra ? 1 : 0) | ((nan_cand > 0 || bad) ? 2ull : 0ull);
        slotB[33] = 0ull;
    }
    __threadfence();
    __syncthreads();
    ok = comm_exchange(comm, seq + 1ull, 34 * 8) && ok;
    if (tid < 33) {
        unsigned long long v = 0ull;
        for (int r = 0; r < comm.world; ++r) {
            const unsigned long long x = __ldcg(reinterpret_cast<const unsigned long long *>(comm_slot(comm, mine, seq + 1ull, r)) + tid);
            v = tid < 32 ? v + x : (v | x);
        }
        if (tid < 32) ws.sums[tid] = v;
        else { ws.flags[0] = (int)(v & 1ull); s_extra = (v & 2ull) ? RRL_STATUS_NAN : 0; }
    }
    if (tid == 0) *comm_epoch(mine) = seq + 1ull;
    __threadfence();
    __syncthreads();
    if (tid < 32) {
        finalize_pair(ws, 0, out_loss, out_status, out_median, out_stats, s_extra | (ok ? 0 : RRL_STATUS_COMM_BIT));
        if (!ok && lane == 0) out_loss[0] = __int_as_float(0x7fc00000);
    }
}

int launch_shard_tail(const Workspace &ws, const Geometry &g, const CommView &comm, float *out_loss, int *out_status,
                      float *out_median, long long *out_stats, cudaStream_t s) {
    static unsigned long long attr_mask = 0ull;
    if (ensure_dyn_smem(shard_tail_kernel, kMedCache * 4, attr_mask)) return RRL_ERR_CUDA;
    shard_tail_kernel<<<1, kShardThreads, kMedCache * 4, s>>>(ws, g, comm, out_loss, out_status, out_median, out_stats);
    count_launch();
    return check_launch();
}

__global__ void __launch_bounds__(128) finalize_kernel(Workspace ws, Geometry g, float *out_loss, int *out_status, float *out_median,
                                                        long long *out_stats) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);           // one warp per pair
    if (b >= g.B || !hdr_ok(ws, g)) return;
    finalize_pair(ws, b, out_loss, out_status, out_median, out_stats);
}

int launch_finalize(const Workspace &ws, const Geometry &g, float *out_loss, int *out_status, float *out_median,
                    long long *out_stats, cudaStream_t s) {
    finalize_kernel<<<(g.B + 3) / 4, 128, 0, s>>>(ws, g, out_loss, out_status, out_median, out_stats);
    count_launch();
    return check_launch();
}

// ------------------------------------------------------------------------------------------------------
// backward: scatter (w/3) G grad_out
// ------------------------------------------------------------------------------------------------------
// The 9 gradient floats of one hit triplet, added with vector reductions (red.global.add.v2/v4.f32, sm_90+): the scatter is
// bound by the number of atomic operations L2 retires, and a triplet's 36 bytes take 3-4 of them (by the alignment of its
// first float) instead of 9.
__device__ __forceinline__ void add9(float *p, const float (&v)[9]) {
    switch ((reinterpret_cast<uintptr_t>(p) >> 2) & 3u) {
    case 0:
        atomicAdd(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
        atomicAdd(reinterpret_cast<float4 *>(p + 4), make_float4(v[4], v[5], v[6], v[7]));
        atomicAdd(p + 8, v[8]);
        break;
    case 1:
        atomicAdd(p, v[0]);
        atomicAdd(reinterpret_cast<float2 *>(p + 1), make_float2(v[1], v[2]));
        atomicAdd(reinterpret_cast<float4 *>(p + 3), make_float4(v[3], v[4], v[5], v[6]));
        atomicAdd(reinterpret_cast<float2 *>(p + 7), make_float2(v[7], v[8]));
        break;
    case 2:
        atomicAdd(reinterpret_cast<float2 *>(p), make_float2(v[0], v[1]));
        atomicAdd(reinterpret_cast<float4 *>(p + 2), make_float4(v[2], v[3], v[4], v[5]));
        atomicAdd(reinterpret_cast<float2 *>(p + 6), make_float2(v[6], v[7]));
        atomicAdd(p + 8, v[8]);
        break;
    default:
        atomicAdd(p, v[0]);
        atomicAdd(reinterpret_cast<float4 *>(p + 1), make_float4(v[1], v[2], v[3], v[4]));
        atomicAdd(reinterpret_cast<float4 *>(p + 5), make_float4(v[5], v[6], v[7], v[8]));
        break;
    }
}
__global__ void __launch_bounds__(128) backward_kernel(Workspace ws, Geometry g, const float *__restrict__ grad_out,
                                                       float *__restrict__ g1, float *__restrict__ g2) {
    const int b = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    mark(16);
    if (!hdr_ok(ws, g) || ws.hdr[6] != 1) return;           // no completed forward of this geometry here: gradients stay zero
    if (i >= ws.nrec[b]) return;
    const long long r = (long long)b * g.nl + i;
    // every load of the record is issued before the first use (the record is whole whatever k and j are: unused slots hold
    // index -1 and zero weights), so the thread waits for ONE round trip to L2 instead of one per hit
    const int4 *I4 = reinterpret_cast<const int4 *>(ws.recIdx + r * 8);
    const float4 *G4 = reinterpret_cast<const float4 *>(ws.recG + r * 24), *W4 = reinterpret_cast<const float4 *>(ws.recW + r * 24);
    const int meta = ws.recMeta[r * 2 + 1];
    const float go = grad_out[b] * (1.0f / 3.0f);
    const int4 ia = I4[0], ib = I4[1];
    float G[24], Wt[24];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const float4 v = G4[q], w = W4[q];
        G[4 * q] = v.x; G[4 * q + 1] = v.y; G[4 * q + 2] = v.z; G[4 * q + 3] = v.w;
        Wt[4 * q] = w.x; Wt[4 * q + 1] = w.y; Wt[4 * q + 2] = w.z; Wt[4 * q + 3] = w.w;
    }
    const int idx[8] = {ia.x, ia.y, ia.z, ia.w, ib.x, ib.y, ib.z, ib.w};
    const int k = meta & 255, j = (meta >> 8) & 255;
    mark(17);
#pragma unroll
    for (int cloud = 0; cloud < 2; ++cloud) {
        float *O0 = cloud ? g2 : g1;
        if (!O0) continue;
        float *O = O0 + (long long)b * (cloud ? g.nf2 : g.nf1) * 9;
        const int n = cloud ? j : k;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            if (a >= n) break;
            const long long f = idx[cloud * 4 + a];
            const float gx = G[cloud * 12 + a * 3] * go, gy = G[cloud * 12 + a * 3 + 1] * go, gz = G[cloud * 12 + a * 3 + 2] * go;
            float v[9];
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                const float w = Wt[cloud * 12 + a * 3 + p];
                v[p * 3] = w * gx; v[p * 3 + 1] = w * gy; v[p * 3 + 2] = w * gz;
            }
            add9(O + f * 9, v);
        }
    }
    mark(18);
}

// Backward straight to pose space (north star (4): "scatters dLoss/dpoint and reduces it to the 6-DoF twist gradient"): when
// cloud 1 is a rigid transform of `raw` (B, nf1, 9), the sparse point gradient of every record is contracted with the RAW
// points on the spot -- acc[b][0..8] += p_raw (x) g, acc[b][9..11] += g, in double -- instead of being scattered into a
// dense (B, nf1, 9) tensor that a second kernel then reads back in full (at 500k triplets: 18 MB zeroed, 18 MB + 18 MB
// re-read, for ~3000 records that touch ~10^4 triplets).  rrl_se3_chain turns acc into the twist gradient.
__global__ void __launch_bounds__(128) backward_pose_kernel(Workspace ws, Geometry g, const float *__restrict__ grad_out,
                                                            const float *__restrict__ raw, double *__restrict__ acc) {
    const int b = blockIdx.y;
    double a[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) a[q] = 0.0;
    const bool ok = hdr_ok(ws, g) && ws.hdr[6] == 1;
    const long long nrec = ok ? ws.nrec[b] : 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nrec; i += (long long)gridDim.x * blockDim.x) {
        const long long r = (long long)b * g.nl + i;
        const int k = ws.recMeta[r * 2 + 1] & 255;
        const float go = grad_out[b] * (1.0f / 3.0f);
        const float *G = ws.recG + r * 24, *Wt = ws.recW + r * 24;
        const int *idx = ws.recIdx + r * 8;
        const float *R0 = raw + (long long)b * g.nf1 * 9;
        for (int h = 0; h < k; ++h) {
            const float *t = R0 + (long long)idx[h] * 9;
            const float gx = G[h * 3] * go, gy = G[h * 3 + 1] * go, gz = G[h * 3 + 2] * go;
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                const float w = Wt[h * 3 + p];
                const double gv[3] = {(double)(w * gx), (double)(w * gy), (double)(w * gz)};      // the float products the scatter adds
                const double pv[3] = {(double)__ldg(t + 3 * p), (double)__ldg(t + 3 * p + 1), (double)__ldg(t + 3 * p + 2)};
#pragma unroll
                for (int m = 0; m < 3; ++m)
#pragma unroll
                    for (int q = 0; q < 3; ++q) a[3 * m + q] += pv[m] * gv[q];
#pragma unroll
                for (int q = 0; q < 3; ++q) a[9 + q] += gv[q];
            }
        }
    }
    __shared__ double sm[4][12];
#pragma unroll
    for (int q = 0; q < 12; ++q) {
        double v = a[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        const double v = (sm[0][threadIdx.x] + sm[1][threadIdx.x]) + (sm[2][threadIdx.x] + sm[3][threadIdx.x]);
        if (v != 0.0) atomicAdd(acc + b * 12 + threadIdx.x, v);
    }
}

int launch_backward_pose(const Workspace &ws, const Geometry &g, const float *grad_out, const float *raw, double *acc, cudaStream_t s) {
    if (cudaMemsetAsync(acc, 0, sizeof(double) * 12 * (size_t)g.B, s) != cudaSuccess) return RRL_ERR_CUDA;
    // records per pair are a small fraction of the lines; the grid covers the capacity and blocks beyond nrec fall through
    int bx = (g.nl + 127) / 128;
    const int cap = (sm_count() * 16 + g.B - 1) / g.B;
    if (bx > cap) bx = cap;
    backward_pose_kernel<<<dim3(bx, g.B), 128, 0, s>>>(ws, g, grad_out, raw, acc);
    count_launch();
    return check_launch();
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
    if (!hdr_ok(ws, g)) {                                     // no forward of this geometry here: report "no hits"
        out_counts[gl] = 0;
        for (int a = 0; a < kCap; ++a) out_hits[gl * kCap + a] = -1;
        return;
    }
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

#ifdef RRL_MARKS
extern "C" int rrl_debug_read_marks(unsigned long long *out32) {
    return cudaMemcpyFromSymbol(out32, rrl::g_marks, sizeof(unsigned long long) * 32) == cudaSuccess ? 0 : -3;
}
extern "C" int rrl_debug_read_cta_times(unsigned long long *out2x8192) {
    return cudaMemcpyFromSymbol(out2x8192, rrl::g_cta_t, sizeof(unsigned long long) * 2 * 8192) == cudaSuccess ? 0 : -3;
}
#endif
