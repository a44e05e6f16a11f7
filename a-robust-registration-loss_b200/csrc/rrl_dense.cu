// Dense phase of the intersected-line loss on sm_100a.
//
// Replaces cal_intersection_batch2_points_with_line (/root/reference/code/loss.py:68-112): for every (line,
// triplet) decide whether all three points of the triplet lie closer to the line than the triplet's local
// threshold, WITHOUT materialising the (nl, nf, 3, 3) tensors the reference builds.
//
//   prep_kernel   per triplet: exact threshold thr_f (reference op order, IEEE sqrt), the filter record
//                 float4(p0, cut_f - |p0|^2) in a point-pair interleaved layout, the cloud's max |p|^2;
//                 also zeroes the per-line hit counters.
//   dense_kernel  streams tiles of filter records through shared memory with 1-D TMA bulk copies
//                 (cp.async.bulk + mbarrier, double buffered) against register-resident lines, evaluates a
//                 conservative FMA-contracted predicate on point 0 of every triplet (7 FP32 ops per
//                 (line, triplet), packed as FFMA2) and queues the rare candidate groups; the queue is drained
//                 with the exact reference-order test of all three points, and confirmed hits are written to
//                 fixed-capacity per-line slots.
//
// The filter is a superset test ("filtered predicate"): the decision itself is always taken by the
// exact arithmetic of loss.py:84-110, so selected indices are bit-exact against the oracle.
#include "rrl_common.cuh"

namespace rrl {

// ------------------------------------------------------------------------------------------------------
// prep
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_kernel(const float *__restrict__ tri1, const float *__restrict__ tri2,
                                                   Workspace ws, Geometry g, int window) {
    const long long n1 = (long long)g.B * g.nf1p, n2 = (long long)g.B * g.nf2p;
    const long long nlines = 2LL * g.B * g.nl;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid0 == 0) {
        ws.hdr[0] = kMagic; ws.hdr[1] = g.B; ws.hdr[2] = g.nf1; ws.hdr[3] = g.nf2; ws.hdr[4] = g.nl; ws.hdr[5] = window;
    }
    // zero the hit counters of both clouds (cnt[0] and cnt[1] are contiguous)
    for (long long i = tid0; i < nlines; i += stride) ws.cnt[0][i] = 0;

    for (long long i = tid0; i < n1 + n2; i += stride) {
        const int cloud = i >= n1;
        const long long r = cloud ? i - n1 : i;
        const int nfp = cloud ? g.nf2p : g.nf1p, nf = cloud ? g.nf2 : g.nf1;
        const int b = (int)(r / nfp), f = (int)(r % nfp);
        float4 rec = make_float4(0.f, 0.f, 0.f, -INFINITY);     // sentinel: never a candidate
        float pm = 0.f;
        if (f < nf) {
            const float *t = (cloud ? tri2 : tri1) + ((long long)b * nf + f) * 9;
            float v[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) v[q] = __ldg(t + q);
            const float thr = triplet_thr_exact(v);
            ws.thr[cloud][(long long)b * nf + f] = thr;
            const double cut = (double)thr * (double)thr - (double)kAddEps;
            const double p2 = (double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2];
            rec = make_float4(v[0], v[1], v[2], (float)(cut - p2));
            pm = fmaxf(fmaxf(sq3_rn(v[0], v[1], v[2]), sq3_rn(v[3], v[4], v[5])), sq3_rn(v[6], v[7], v[8]));
            if (!(pm == pm)) pm = INFINITY;
        }
        // pair-interleaved layout: points (2i, 2i+1) -> {xA,xB,yA,yB}, {zA,zB,wA,wB}
        float *dst = reinterpret_cast<float *>(ws.tri4[cloud] + ((long long)b * nfp + (f & ~1)));
        const int h = f & 1;
        dst[0 + h] = rec.x; dst[2 + h] = rec.y; dst[4 + h] = rec.z; dst[6 + h] = rec.w;
        // cloud max |p|^2 (non-negative floats order like their bit patterns)
        const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(pm));
        // lanes of one warp may straddle a pair boundary: fall back to per-lane atomics in that (rare) case
        const int b0 = __shfl_sync(0xffffffffu, b, 0), c0 = __shfl_sync(0xffffffffu, cloud, 0);
        const bool uniform = __all_sync(0xffffffffu, b == b0 && cloud == c0);
        if (uniform) {
            if ((threadIdx.x & 31) == 0) atomicMax(ws.pmax + b * 2 + cloud, m);
        } else {
            atomicMax(ws.pmax + b * 2 + cloud, __float_as_uint(pm));
        }
    }
}

int launch_prep(const float *tri1, const float *tri2, const Workspace &ws, const Geometry &g, int window, cudaStream_t s) {
    // per-pair block {pmax, nrec, n_kj, med, sums, stats, gcounts} is contiguous, see carve()
    const size_t pair_bytes = (size_t)((char *)(ws.gcounts + (size_t)g.B * 18) - (char *)ws.pmax);
    if (cudaMemsetAsync(ws.pmax, 0, pair_bytes, s) != cudaSuccess) return RRL_ERR_CUDA;
    const long long work = (long long)g.B * (g.nf1p + g.nf2p);
    const long long lines = 2LL * g.B * g.nl;
    long long want = (work > lines / 4 ? work : lines / 4);
    int blocks = (int)((want + 255) / 256);
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 16) blocks = 148 * 16;
    // the warp-uniform __reduce/__shfl calls need whole warps in the loop: the loop bounds are uniform per warp
    // only if every lane runs the same trip count, which grid-stride loops do not guarantee -> pad the triplet
    // loop by running it over a multiple of 32 (nf*p are multiples of 64, so B*(nf1p+nf2p) is).
    prep_kernel<<<blocks, 256, 0, s>>>(tri1, tri2, ws, g, window);
    count_launch();
    return check_launch();
}

// ------------------------------------------------------------------------------------------------------
// dense
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_load(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct DenseArgs {
    const float *tri[2];      // (B, nf, 9) original triplets
    const float *lines;       // (B, nl, 6)
    int chunk_points;         // points per blockIdx.y chunk (multiple of kPointPad)
};

// Exact test of one (line, triplet) -- literal restatement of loss.py:84-110 -- and hit recording.
__device__ __forceinline__ void exact_test_and_record(const float *__restrict__ tri, const float *__restrict__ thr_arr,
                                                      const float *ln, int f, int *cnt_line, int *hits_line,
                                                      int &band, int &nan) {
    const float *t = tri + (long long)f * 9;
    const float thr = __ldg(thr_arr + f);
    const float ulp = ulp_up(thr);
    const float x0 = point_line_x_exact(__ldg(t + 0), __ldg(t + 1), __ldg(t + 2), ln);
    const float d0 = __fsqrt_rn(x0);
    nan += (d0 != d0);
    band += (fabsf(d0 - thr) <= ulp);
    if (!(d0 < thr)) return;
    const float d1 = __fsqrt_rn(point_line_x_exact(__ldg(t + 3), __ldg(t + 4), __ldg(t + 5), ln));
    const float d2 = __fsqrt_rn(point_line_x_exact(__ldg(t + 6), __ldg(t + 7), __ldg(t + 8), ln));
    nan += (d1 != d1) + (d2 != d2);
    band += (fabsf(d1 - thr) <= ulp) + (fabsf(d2 - thr) <= ulp);
    if ((d1 < thr) & (d2 < thr)) {
        const int slot = atomicAdd(cnt_line, 1);
        if (slot < kCap) hits_line[slot] = f;
    }
}

template <bool kPacked>
__global__ void __launch_bounds__(kDenseThreads, 2) dense_kernel(DenseArgs a, Workspace ws, Geometry g) {
    __shared__ __align__(128) float4 stage[2][kTilePoints];
    __shared__ __align__(8) unsigned long long mbar[2];
    __shared__ unsigned queue[kQueueCap];
    __shared__ int q_count;
    __shared__ int s_band, s_nan, s_cand;

    const int tid = threadIdx.x;
    const int b = blockIdx.z >> 1, cloud = blockIdx.z & 1;
    const int nf = cloud ? g.nf2 : g.nf1, nfp = cloud ? g.nf2p : g.nf1p;
    const int p_begin = blockIdx.y * a.chunk_points;
    if (p_begin >= nfp) return;
    const int p_end = min(nfp, p_begin + a.chunk_points);
    const int line_base = blockIdx.x * kLinesPerCta;

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        q_count = 0; s_band = 0; s_nan = 0; s_cand = 0;
    }

    // ---- per-thread lines -> filter constants (double precision, rounded once) --------------------------
    const float *lines_b = a.lines + (long long)b * g.nl * 6;
    const float P = sqrtf(__uint_as_float(ws.pmax[b * 2 + cloud])) * 1.000001f;
    float ux[kLinesPerThread], uy[kLinesPerThread], uz[kLinesPerThread];
    float mx[kLinesPerThread], my[kLinesPerThread], mz[kLinesPerThread], tl[kLinesPerThread];
#pragma unroll
    for (int i = 0; i < kLinesPerThread; ++i) {
        const int l = line_base + tid + i * kDenseThreads;
        ux[i] = uy[i] = uz[i] = mx[i] = my[i] = mz[i] = 0.f;
        tl[i] = INFINITY;                                         // Q > +inf never holds: padding lines are inert
        if (l < g.nl) {
            const float *ln = lines_b + (long long)l * 6;
            const float u0 = __ldg(ln), u1 = __ldg(ln + 1), u2 = __ldg(ln + 2);
            const double x = __ldg(ln + 3), y = __ldg(ln + 4), z = __ldg(ln + 5);
            const double sd = x * u0 + y * u1 + z * u2;
            const double xx = x * x + y * y + z * z;
            ux[i] = u0; uy[i] = u1; uz[i] = u2;
            mx[i] = (float)(2.0 * (x - sd * u0));
            my[i] = (float)(2.0 * (y - sd * u1));
            mz[i] = (float)(2.0 * (z - sd * u2));
            const float X = (float)sqrt(xx) * 1.000001f;
            const float guard = kGuard * 5.9604645e-8f * (P + X) * (P + X) + 1e-12f;
            tl[i] = (float)((xx - sd * sd) - (double)guard);
        }
    }
    __syncthreads();

    const float4 *src = ws.tri4[cloud] + (long long)b * nfp;
    const int ntiles = (p_end - p_begin + kTilePoints - 1) / kTilePoints;
    auto issue = [&](int t) {
        const int s0 = p_begin + t * kTilePoints;
        const int n = min(kTilePoints, p_end - s0);
        const unsigned bytes = (unsigned)n * 16u;
        mbar_expect_tx(&mbar[t & 1], bytes);
        tma_bulk_load(&stage[t & 1][0], src + s0, bytes, &mbar[t & 1]);
    };
    if (tid == 0) issue(0);

    int band = 0, nan = 0, ncand = 0;

    // exact re-test of every queued candidate group; entry = tid<<22 | (group index inside the chunk)
    auto drain = [&]() {
        const int n = min(q_count, kQueueCap);
        for (int e = tid; e < n; e += kDenseThreads) {
            const unsigned ent = queue[e];
            const int owner = ent >> 22;
            const int f0 = p_begin + (int)(ent & 0x3fffffu) * 4;
#pragma unroll 1
            for (int i = 0; i < kLinesPerThread; ++i) {
                const int l = line_base + owner + i * kDenseThreads;
                if (l >= g.nl) break;
                float ln[6];
#pragma unroll
                for (int q = 0; q < 6; ++q) ln[q] = __ldg(lines_b + (long long)l * 6 + q);
                const long long gl = (long long)b * g.nl + l;
#pragma unroll 1
                for (int p = 0; p < 4; ++p) {
                    const int f = f0 + p;
                    if (f < nf)
                        exact_test_and_record(a.tri[cloud] + (long long)b * nf * 9, ws.thr[cloud] + (long long)b * nf, ln, f,
                                              ws.cnt[cloud] + gl, ws.hits[cloud] + gl * kCap, band, nan);
                }
            }
        }
    };

    for (int t = 0; t < ntiles; ++t) {
        if (tid == 0 && t + 1 < ntiles) issue(t + 1);
        mbar_wait(&mbar[t & 1], (t >> 1) & 1);
        const int npts = min(kTilePoints, p_end - (p_begin + t * kTilePoints));     // multiple of kPointPad
        const float4 *sp = &stage[t & 1][0];
        const int group0 = t * (kTilePoints / 4);
#pragma unroll 2
        for (int gi = 0; gi < npts / 4; ++gi) {
            // 4 points = two interleaved pairs: {xA,xB,yA,yB} {zA,zB,wA,wB}
            const float4 a0 = sp[gi * 4 + 0], a1 = sp[gi * 4 + 1], b0 = sp[gi * 4 + 2], b1 = sp[gi * 4 + 3];
            bool any = false;
            if constexpr (kPacked) {
                const float2 xa = make_float2(a0.x, a0.y), ya = make_float2(a0.z, a0.w), za = make_float2(a1.x, a1.y), wa = make_float2(a1.z, a1.w);
                const float2 xb = make_float2(b0.x, b0.y), yb = make_float2(b0.z, b0.w), zb = make_float2(b1.x, b1.y), wb = make_float2(b1.z, b1.w);
#pragma unroll
                for (int i = 0; i < kLinesPerThread; ++i) {
                    const float2 u0 = make_float2(ux[i], ux[i]), u1 = make_float2(uy[i], uy[i]), u2 = make_float2(uz[i], uz[i]);
                    const float2 m0 = make_float2(mx[i], mx[i]), m1 = make_float2(my[i], my[i]), m2 = make_float2(mz[i], mz[i]);
                    float2 ta = __ffma2_rn(za, u2, __ffma2_rn(ya, u1, __fmul2_rn(xa, u0)));
                    float2 sa = __ffma2_rn(za, m2, __ffma2_rn(ya, m1, __ffma2_rn(xa, m0, wa)));
                    float2 qa = __ffma2_rn(ta, ta, sa);
                    float2 tb = __ffma2_rn(zb, u2, __ffma2_rn(yb, u1, __fmul2_rn(xb, u0)));
                    float2 sb = __ffma2_rn(zb, m2, __ffma2_rn(yb, m1, __ffma2_rn(xb, m0, wb)));
                    float2 qb = __ffma2_rn(tb, tb, sb);
                    any |= (qa.x > tl[i]) | (qa.y > tl[i]) | (qb.x > tl[i]) | (qb.y > tl[i]);
                }
            } else {
                const float px[4] = {a0.x, a0.y, b0.x, b0.y}, py[4] = {a0.z, a0.w, b0.z, b0.w};
                const float pz[4] = {a1.x, a1.y, b1.x, b1.y}, pw[4] = {a1.z, a1.w, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < kLinesPerThread; ++i) {
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float tt = fmaf(pz[p], uz[i], fmaf(py[p], uy[i], px[p] * ux[i]));
                        const float ss = fmaf(pz[p], mz[i], fmaf(py[p], my[i], fmaf(px[p], mx[i], pw[p])));
                        any |= fmaf(tt, tt, ss) > tl[i];
                    }
                }
            }
            if (any) {
                const int pos = atomicAdd(&q_count, 1);
                ++ncand;
                if (pos < kQueueCap) {
                    queue[pos] = ((unsigned)tid << 22) | (unsigned)(group0 + gi);
                } else {
                    // queue full: resolve this group right here (rare; keeps the result exact)
                    const int f0 = p_begin + (group0 + gi) * 4;
#pragma unroll 1
                    for (int i = 0; i < kLinesPerThread; ++i) {
                        const int l = line_base + tid + i * kDenseThreads;
                        if (l >= g.nl) break;
                        float ln[6];
#pragma unroll
                        for (int q = 0; q < 6; ++q) ln[q] = __ldg(lines_b + (long long)l * 6 + q);
                        const long long gl = (long long)b * g.nl + l;
#pragma unroll 1
                        for (int p = 0; p < 4; ++p)
                            if (f0 + p < nf)
                                exact_test_and_record(a.tri[cloud] + (long long)b * nf * 9, ws.thr[cloud] + (long long)b * nf, ln,
                                                      f0 + p, ws.cnt[cloud] + gl, ws.hits[cloud] + gl * kCap, band, nan);
                    }
                }
            }
        }
        __syncthreads();                       // everyone is done with stage[t&1] (and with pushing to the queue)
        if (q_count > kQueueCap / 2 || t + 1 == ntiles) {
            drain();
            __syncthreads();
            if (tid == 0) q_count = 0;
            __syncthreads();
        }
    }

    // ---- diagnostics ---------------------------------------------------------------------------------
    if (band) atomicAdd(&s_band, band);
    if (nan) atomicAdd(&s_nan, nan);
    if (ncand) atomicAdd(&s_cand, ncand);
    __syncthreads();
    if (tid == 0) {
        long long *st = ws.stats + (long long)b * RRL_NSTAT;
        if (s_cand) atomicAdd((unsigned long long *)(st + 3 + cloud), (unsigned long long)s_cand);
        if (s_band) atomicAdd((unsigned long long *)(st + 5), (unsigned long long)s_band);
        if (s_nan) atomicAdd((unsigned long long *)(st + 6), (unsigned long long)s_nan);
    }
}

static int g_dense_variant = 1;        // 1 = packed FFMA2, 0 = scalar FFMA (selectable for measurement)
void set_dense_variant(int v) { g_dense_variant = v; }

int launch_dense(const float *tri1, const float *tri2, const float *lines, const Workspace &ws, const Geometry &g, cudaStream_t s) {
    DenseArgs a;
    a.tri[0] = tri1; a.tri[1] = tri2; a.lines = lines;
    const int line_tiles = (g.nl + kLinesPerCta - 1) / kLinesPerCta;
    const int nfp_max = g.nf1p > g.nf2p ? g.nf1p : g.nf2p;
    // split the points so that the grid covers the 148 SMs a few times over when the line tiles alone do not
    const long long base_ctas = (long long)line_tiles * g.B * 2;
    const long long target = 148LL * 2 * 3;
    int chunks = 1;
    if (base_ctas < target) chunks = (int)((target + base_ctas - 1) / base_ctas);
    int chunk_points = (nfp_max + chunks - 1) / chunks;
    chunk_points = ((chunk_points + 255) / 256) * 256;                 // >= 256 points per CTA, multiple of kPointPad
    if (chunk_points > (1 << 22) * 4) return RRL_ERR_ARG;
    chunks = (nfp_max + chunk_points - 1) / chunk_points;
    a.chunk_points = chunk_points;
    dim3 grid(line_tiles, chunks, g.B * 2);
    if (g_dense_variant)
        dense_kernel<true><<<grid, kDenseThreads, 0, s>>>(a, ws, g);
    else
        dense_kernel<false><<<grid, kDenseThreads, 0, s>>>(a, ws, g);
    count_launch();
    return check_launch();
}

}  // namespace rrl
