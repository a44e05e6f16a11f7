// Dense phase of the intersected-line loss on sm_100a.
//
// Replaces cal_intersection_batch2_points_with_line (/root/reference/code/loss.py:68-112): for every (line,
// triplet) decide whether all three points of the triplet lie closer to the line than the triplet's local
// threshold, WITHOUT materialising the (nl, nf, 3, 3) tensors the reference builds.
//
//   prep_kernel    per triplet: exact threshold thr_f (reference op order, IEEE sqrt) and the cloud's max |p|^2;
//                  per line: the filter constants {u, |x0|, M, c} in double precision, the pair's max |x0|^2;
//                  zeroes the per-line hit counters.  NaN / infinite inputs stay out of the extents and raise a flag.
//   sort kernels   Hilbert-curve order of every cloud's points (a register bitonic sort inside small_prep_kernel up
//                  to 4096 triplets; above, ONE CUB radix sort over every cloud of every pair) so that kNode
//                  consecutive triplets are spatial neighbours; kd_refine64 re-partitions windows of 64 sorted
//                  triplets into k-d leaves of kNode (small clouds).
//   node_kernel    per node of kNode sorted triplets: bounding sphere (centre q, radius R covering every
//                  triplet's hit cylinder, upper bounds in float with directed rounding) -> node record
//                  float4(q, R^2 - |q|^2); per triplet the point record float4(p0, cut_f - |p0|^2) in sorted
//                  order; from 16384 triplets on also one SUPER-node record per 256 sorted triplets, the triplet
//                  records in compressed form (16-bit offsets + half cut, error measured and folded into the cut) and
//                  per super node ONE compressed record of its 16 node spheres.
//   small_prep_kernel  all of the above in one launch for clouds up to 4096 triplets.
//   dense_kernel   streams tiles of node (or super-node) records through shared memory with 1-D TMA bulk copies
//                  (cp.async.bulk + mbarrier, double buffered) against register-resident lines.  7 packed FP32
//                  ops (FFMA2) decide per (line, record) whether the line can touch the sphere; results are
//                  accumulated branch-free into per-line bit masks and pushed (ordered, by warp scans) to
//                  warp-private queues; further levels -- node predicate, triplet predicate, refine on points 1
//                  and 2 -- each run one queue entry per lane with converged warps; every level calls the next from
//                  ONE site (code size), masks are built from the sign bits of tl - Q.
//   exact_kernel   the EXACT reference-order test of all three points of every surviving (line, triplet) pair;
//                  confirmed hits go to fixed-capacity per-line slots.
//
// Both filters are superset tests ("filtered predicate"); the decision itself is always taken by the
// literal arithmetic of loss.py:84-110, so the selected indices are bit-exact against the oracle.
#include <cuda_fp16.h>
#include <cub/device/device_radix_sort.cuh>

#include "rrl_common.cuh"

namespace rrl {

// max |p|^2 over the FINITE points of a triplet; a NaN / infinite coordinate (such a point can never pass the test: its
// distance is NaN) stays out of the extent -- it would turn every filter threshold into -inf or NaN -- and raises `bad`
__device__ __forceinline__ float finite_extent(const float *v, int &bad) {
    float m = 0.f;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const float s = sq3_rn(v[3 * q], v[3 * q + 1], v[3 * q + 2]);
        if (s < INFINITY) m = fmaxf(m, s); else bad = 1;
    }
    return m;
}

// ------------------------------------------------------------------------------------------------------
// prep: thresholds, line constants, extents
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_kernel(const float *__restrict__ tri1, const float *__restrict__ tri2,
                                                   const float *__restrict__ lines, Workspace ws, Geometry g, int window,
                                                   int reuse_target) {
    const int b = blockIdx.y;
    // RRL_REUSE_TARGET: cloud 2's thresholds and extent survive from the previous forward of this geometry (hdr[7] is written by
    // that forward's last kernel, by nobody in this launch: every CTA takes the same decision)
    const bool keep2 = reuse_target && ws.hdr[7] == order_token(g);
    const int stride = gridDim.x * blockDim.x;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0 && t0 == 0) {
        ws.hdr[0] = kMagic; ws.hdr[1] = g.B; ws.hdr[2] = g.nf1; ws.hdr[3] = g.nf2; ws.hdr[4] = g.nl; ws.hdr[5] = window;
        ws.hdr[6] = 0;
    }
    float block_max[3], block_thr[2];
    int badv = 0;
#pragma unroll
    for (int cloud = 0; cloud < 2; ++cloud) {
        const int nf = cloud ? g.nf2 : g.nf1;
        const float *tri = (cloud ? tri2 : tri1) + (long long)b * nf * 9;
        float *thr = ws.thr[cloud] + (long long)b * nf;
        float pm = 0.f;                                                           // running maximum: ONE atomic per warp
        float tm = 0.f;
        if (cloud == 1 && keep2) {
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                ws.pmax[b * 2 + 1] = ws.keep[b * 8 + 0];
                ws.bad[b * 2 + 1] = ws.keep[b * 8 + 3];
                ws.tmax[b * 2 + 1] = ws.keep[b * 8 + 5];
            }
            block_max[cloud] = 0.f;
            block_thr[cloud] = 0.f;
            continue;
        }
        for (int base = blockIdx.x * blockDim.x; base < nf; base += stride) {      // warp-uniform trip count
            const int f = base + threadIdx.x;
            if (f < nf) {
                float v[9];
#pragma unroll
                for (int q = 0; q < 9; ++q) v[q] = __ldg(tri + (long long)f * 9 + q);
                const float th = triplet_thr_exact(v);
                thr[f] = th;
                if (th < INFINITY) tm = fmaxf(tm, th);            // a NaN / infinite threshold never hits and stays out
                pm = fmaxf(pm, finite_extent(v, badv));
            }
        }
        block_max[cloud] = pm;
        block_thr[cloud] = tm;
        if (__any_sync(0xffffffffu, badv) && (threadIdx.x & 31) == 0) atomicOr(ws.bad + b * 2 + cloud, 1u);
        badv = 0;
    }
    const float *lb = lines + (long long)b * g.nl * 6;
    float xmw = 0.f;
    for (int base = blockIdx.x * blockDim.x; base < g.nl; base += stride) {
        const int l = base + threadIdx.x;
        float xm = 0.f;
        if (l < g.nl) {
            const float *ln = lb + (long long)l * 6;
            const float u0 = __ldg(ln), u1 = __ldg(ln + 1), u2 = __ldg(ln + 2);
            const double x = __ldg(ln + 3), y = __ldg(ln + 4), z = __ldg(ln + 5);
            const double sd = x * u0 + y * u1 + z * u2;
            const double xx = x * x + y * y + z * z;
            const long long gl = (long long)b * g.nl + l;
            ws.lineC[gl * 2] = make_float4(u0, u1, u2, (float)sqrt(xx) * 1.000001f);
            ws.lineC[gl * 2 + 1] = make_float4((float)(2.0 * (x - sd * u0)), (float)(2.0 * (y - sd * u1)),
                                               (float)(2.0 * (z - sd * u2)), (float)(xx - sd * sd));
            ws.cnt[0][gl] = 0;
            ws.cnt[1][gl] = 0;
            xm = (float)xx * 1.000001f + 0.f * (u0 + u1 + u2);       // NaN for a NaN / infinite position OR direction
            if (!(xm < INFINITY)) { xm = 0.f; badv = 1; }            // such a line never hits: it stays out of the extent
        }
        xmw = fmaxf(xmw, xm);
    }
    block_max[2] = xmw;
    if (__any_sync(0xffffffffu, badv) && (threadIdx.x & 31) == 0) atomicOr(ws.bad + b * 2, 1u);
    // one atomic per CTA and array: thousands of same-address atomics would serialise in L2
    __shared__ unsigned s_max[5];
    if (threadIdx.x < 5) s_max[threadIdx.x] = 0u;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(q < 3 ? block_max[q] : block_thr[q - 3]));
        if ((threadIdx.x & 31) == 0 && m) atomicMax(&s_max[q], m);
    }
    __syncthreads();
    if (threadIdx.x < 5 && s_max[threadIdx.x])
        atomicMax(threadIdx.x < 2 ? ws.pmax + b * 2 + threadIdx.x : (threadIdx.x == 2 ? ws.xmax + b * 2 : ws.tmax + b * 2 + threadIdx.x - 3),
                  s_max[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------------
// Morton order
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned spread10(unsigned v) {
    v &= 1023u;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// Hilbert-curve index (Skilling's transpose algorithm, 10 bits per axis): unlike Morton order, consecutive keys
// are always neighbouring cells, so kNode consecutive triplets form compact nodes
__device__ __forceinline__ unsigned morton_key(const float *p, float P) {
    const float s = P > 0.f ? 512.0f / P : 0.f;              // [-P, P] -> [0, 1024)
    unsigned X[3];
    X[0] = (unsigned)min(1023, max(0, (int)((p[0] + P) * s)));
    X[1] = (unsigned)min(1023, max(0, (int)((p[1] + P) * s)));
    X[2] = (unsigned)min(1023, max(0, (int)((p[2] + P) * s)));
    for (unsigned Q = 512u; Q > 1u; Q >>= 1) {
        const unsigned Pm = Q - 1u;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (X[i] & Q) X[0] ^= Pm;
            else { const unsigned t = (X[0] ^ X[i]) & Pm; X[0] ^= t; X[i] ^= t; }
        }
    }
    X[1] ^= X[0];
    X[2] ^= X[1];
    unsigned t = 0;
    for (unsigned Q = 512u; Q > 1u; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1u;
    X[0] ^= t; X[1] ^= t; X[2] ^= t;
    return (spread10(X[0]) << 2) | (spread10(X[1]) << 1) | spread10(X[2]);
}

// one CTA sorts one cloud (nfp <= kSortSmall) with a bitonic network on (key << 32 | index)
__global__ void __launch_bounds__(1024) sort_small_kernel(const float *__restrict__ tri1, const float *__restrict__ tri2,
                                                          Workspace ws, Geometry g, int sorted) {
    extern __shared__ unsigned long long skeys[];
    const int b = blockIdx.x, cloud = blockIdx.y;
    const int nf = cloud ? g.nf2 : g.nf1, nfp = cloud ? g.nf2p : g.nf1p;
    const float *tri = (cloud ? tri2 : tri1) + (long long)b * nf * 9;
    const float P = sqrtf(__uint_as_float(ws.pmax[b * 2 + cloud]));
    int n2 = 1;
    while (n2 < nfp) n2 <<= 1;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        unsigned long long v = 0xFFFFFFFFFFFFFFFFull;
        if (i < nf) {
            const float p[3] = {__ldg(tri + (long long)i * 9), __ldg(tri + (long long)i * 9 + 1), __ldg(tri + (long long)i * 9 + 2)};
            const unsigned key = sorted ? morton_key(p, P) : 0u;
            v = ((unsigned long long)key << 32) | (unsigned)i;
        }
        skeys[i] = v;
    }
    __syncthreads();
    if (sorted) {
        for (int k = 2; k <= n2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const unsigned long long a = skeys[i], c = skeys[ixj];
                        const bool up = (i & k) == 0;
                        if ((a > c) == up) { skeys[i] = c; skeys[ixj] = a; }
                    }
                }
                __syncthreads();
            }
    }
    int *perm = ws.perm[cloud] + (long long)b * nfp;
    for (int i = threadIdx.x; i < nfp; i += blockDim.x) {
        const unsigned idx = (unsigned)(skeys[i] & 0xFFFFFFFFull);
        perm[i] = idx < (unsigned)nf ? (int)idx : -1;
    }
}

// large clouds: keys for ONE CUB radix sort over both clouds of up to kSortPairs pairs.  Segment = cloud * nb + pair;
// key = segment << (32 - segbits) | padding << (31 - segbits) | the leading 31 - segbits bits of the 30-bit Hilbert index
// (any prefix of a Hilbert index is the index on a coarser grid): the clouds of all pairs sort into place in one go,
// cloud 0 of every pair first (the layout of perm[0], perm[1]), paddings last in each segment.
constexpr int kSortPairs = 512;
__global__ void __launch_bounds__(256) sort_keys_kernel(const float *__restrict__ tri1, const float *__restrict__ tri2, int nb, int nf1, int nf1p,
                                                        int nf2, int nf2p, const unsigned int *pmax_bits, int segbits, unsigned *keys, int *vals,
                                                        int sorted) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n1 = (long long)nb * nf1p;
    if (t >= n1 + (long long)nb * nf2p) return;
    const int cloud = t >= n1;
    const long long r = cloud ? t - n1 : t;
    const int nfp = cloud ? nf2p : nf1p, nf = cloud ? nf2 : nf1;
    const int b = (int)(r / nfp), i = (int)(r % nfp);
    const float *tri = (cloud ? tri2 : tri1) + (long long)b * nf * 9;
    const float Pm = sqrtf(__uint_as_float(pmax_bits[b * 2 + cloud]));
    const int hbits = 31 - segbits < 30 ? 31 - segbits : 30;
    unsigned key = (1u << hbits);                                 // padding flag: after every real key of the segment
    int val = -1;
    if (i < nf) {
        const float p[3] = {__ldg(tri + (long long)i * 9), __ldg(tri + (long long)i * 9 + 1), __ldg(tri + (long long)i * 9 + 2)};
        key = sorted ? morton_key(p, Pm) >> (30 - hbits) : 0u;
        val = i;
    }
    keys[t] = key | ((unsigned)(cloud * nb + b) << (hbits + 1));
    vals[t] = val;
}

// ------------------------------------------------------------------------------------------------------
// local k-d refinement of the Hilbert order
// ------------------------------------------------------------------------------------------------------
// kNode consecutive points of a space-filling curve through a SURFACE are a strip, or two patches when the curve
// leaves the surface in between: bounding spheres several times the area of a compact cluster of kNode points.  One
// warp re-partitions a window of 64 consecutive sorted triplets into k-d leaves of kNode (median splits along the
// longest axis of the segment's bounding box, 64 -> 32 -> 16 (-> 8)): about a third fewer (line, node) candidates.
// Only the ORDER inside the window changes (perm); nothing downstream depends on how the order was made.
//   slot s = lane + 32 e holds element e of the lane.  Per level every element gets the key
//   segment << 29 | 23 leading bits of its order-preserving coordinate | slot  (unique), its rank among the 64 keys IS
//   its new slot; elements move through a per-warp shared-memory scratch.  Paddings and NaNs sort last.
template <int kNode>
__device__ __forceinline__ void kd_refine64(const float *__restrict__ tri, int &f0, int &f1, int lane, float4 *scratch) {
    float x[2], y[2], z[2];
    int f[2] = {f0, f1};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        x[e] = y[e] = z[e] = __int_as_float(0x7fc00000);         // padding: NaN = "not there"
        if (f[e] >= 0) {
            const float *t = tri + (long long)f[e] * 9;
            x[e] = __ldg(t); y[e] = __ldg(t + 1); z[e] = __ldg(t + 2);
        }
    }
#pragma unroll
    for (int m = 64; m > kNode; m >>= 1) {
        // bounding box of the element's segment (fminf / fmaxf drop NaNs); lanes of one segment end up with equal bits
        float lo[2][3], hi[2][3];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            lo[e][0] = hi[e][0] = x[e]; lo[e][1] = hi[e][1] = y[e]; lo[e][2] = hi[e][2] = z[e];
        }
        if (m == 64) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                lo[0][a] = lo[1][a] = fminf(lo[0][a], lo[1][a]);
                hi[0][a] = hi[1][a] = fmaxf(hi[0][a], hi[1][a]);
            }
        }
        const int span = m >= 32 ? 32 : m;                       // lanes per segment
#pragma unroll
        for (int d = 1; d < span; d <<= 1)
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    lo[e][a] = fminf(lo[e][a], __shfl_xor_sync(0xffffffffu, lo[e][a], d));
                    hi[e][a] = fmaxf(hi[e][a], __shfl_xor_sync(0xffffffffu, hi[e][a], d));
                }
        unsigned key[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const float ex = hi[e][0] - lo[e][0], ey = hi[e][1] - lo[e][1], ez = hi[e][2] - lo[e][2];
            const float c = (ex >= ey && ex >= ez) ? x[e] : (ey >= ez ? y[e] : z[e]);
            const unsigned u = __float_as_uint(c);
            unsigned ord = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
            if (!(c == c)) ord = 0xFFFFFFFFu;
            const unsigned slot = (unsigned)(lane + 32 * e);
            key[e] = ((slot / (unsigned)m) << 29) | ((ord >> 9) << 6) | slot;
        }
        int rank[2] = {0, 0};
#pragma unroll 8
        for (int t = 0; t < 32; ++t) {
            const unsigned k0 = __shfl_sync(0xffffffffu, key[0], t), k1 = __shfl_sync(0xffffffffu, key[1], t);
            rank[0] += (k0 < key[0]) + (k1 < key[0]);
            rank[1] += (k0 < key[1]) + (k1 < key[1]);
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 2; ++e) scratch[rank[e]] = make_float4(x[e], y[e], z[e], __int_as_float(f[e]));
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const float4 q = scratch[lane + 32 * e];
            x[e] = q.x; y[e] = q.y; z[e] = q.z; f[e] = __float_as_int(q.w);
        }
    }
    f0 = f[0];
    f1 = f[1];
}

// large clouds: one warp per window of 64 sorted positions, in place on perm
template <int kNode>
__global__ void __launch_bounds__(256) refine_kernel(const float *__restrict__ tri1, const float *__restrict__ tri2, Workspace ws, Geometry g) {
    __shared__ float4 s_kd[8][64];
    const int b = blockIdx.y >> 1, cloud = blockIdx.y & 1;
    const int nf = cloud ? g.nf2 : g.nf1, nfp = cloud ? g.nf2p : g.nf1p;
    const float *tri = (cloud ? tri2 : tri1) + (long long)b * nf * 9;
    int *perm = ws.perm[cloud] + (long long)b * nfp;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int w = blockIdx.x * 8 + wid; w < nfp / 64; w += gridDim.x * 8) {
        int f0 = perm[w * 64 + lane], f1 = perm[w * 64 + 32 + lane];
        kd_refine64<kNode>(tri, f0, f1, lane, s_kd[wid]);
        perm[w * 64 + lane] = f0;
        perm[w * 64 + 32 + lane] = f1;
    }
}

// ------------------------------------------------------------------------------------------------------
// nodes
// ------------------------------------------------------------------------------------------------------
// Upper bound, in float with directed rounding, of  sqrt(cut_f + E) + |p0_f - q|  (cut_f = thr_f^2 - 2e-4): the radius a
// sphere centred at q needs to cover triplet f's hit cylinder.  (The double-precision version of this cost more than
// everything else in the node build: FP64 issue is slow on this part.)
__device__ __forceinline__ float cover_radius_up(float th, float E_up, float px, float py, float pz, float qx, float qy, float qz) {
    const float cut_up = __fadd_ru(__fmul_ru(th, th), -kAddEps);
    const float a = __fsqrt_ru(fmaxf(__fadd_ru(cut_up, E_up), 0.f));
    const float dx = fmaxf(fabsf(__fsub_ru(px, qx)), fabsf(__fsub_rd(px, qx)));
    const float dy = fmaxf(fabsf(__fsub_ru(py, qy)), fabsf(__fsub_rd(py, qy)));
    const float dz = fmaxf(fabsf(__fsub_ru(pz, qz)), fabsf(__fsub_rd(pz, qz)));
    const float d2 = __fmaf_ru(dz, dz, __fmaf_ru(dy, dy, __fmul_ru(dx, dx)));
    return __fadd_ru(a, __fsqrt_ru(d2));
}
// sphere record {q, w = R^2 - |q|^2}, w rounded up (a larger w only admits more candidates); also returns the radius
__device__ __forceinline__ float4 sphere_record_up(float R, float qx, float qy, float qz, float &rad) {
    R = __fmul_ru(R, 1.000002f);
    rad = __fmul_ru(R, 1.000001f);
    const float q2_dn = __fmaf_rd(qz, qz, __fmaf_rd(qy, qy, __fmul_rd(qx, qx)));
    float w = __fsub_ru(__fmul_ru(R, R), q2_dn);
    w = w + fabsf(w) * 2.4e-7f + 1e-30f;
    return make_float4(qx, qy, qz, w);
}

// ---- compressed triplet records of the large-cloud modes (level 2 gathers them from L2, every lane another node) ----------
// A node's point-0 records as 16 + 8 kNode bytes instead of 16 (kNode + 1): header {qb.xyz, scale}, then per PAIR of triplets
// one uint4 {xA | xB << 16, yA | yB << 16, zA | zB << 16, half2(cutA, cutB)} with 16-bit offsets u.  The dense kernel
// reconstructs  p~ = fma(float(2^23 + u), scale, qb)  -- float(2^23 + u) is the bit pattern 0x4B000000 | u, one PRMT -- and
// evaluates the SAME packed-FMA predicate on (p~, cut~).  Superset argument: the node kernel runs the identical
// FMA, MEASURES e >= |p~ - p| (directed rounding), and stores cut~ >= cut + 2 e sqrt(cut + E) + e^2 rounded up to half:
// "p passes the reference test" => d(p)^2 < cut + E => d(p~)^2 <= (d(p) + e)^2 < cut~ + E.  Nothing is assumed about the
// quantiser: a clamped or badly rounded offset only makes e, and with it cut~, larger.  The dense kernel bounds e by
// pc_err_bound() for the |u| > 1 slack; a record that would exceed it gets cut~ = +inf (always a candidate).
constexpr float kPcBias = 8388608.0f + 32768.0f;        // float(2^23 + u) - kPcBias = signed offset in steps
constexpr float kPcSteps = 32000.0f;                    // largest offset in steps (headroom below 32767 for the rounding of qb)
__device__ __forceinline__ float pc_err_bound(float R, float P) {      // >= |p~ - p| for every member of a node of radius <= R
    // sqrt(3) (half a step of R / kPcSteps + one rounding of a value of magnitude <= P + R), generously rounded up
    return 1.7321f * (R * (0.5f / kPcSteps) * 1.01f + 6.0e-8f * 1.01f * (P + R)) + 1e-37f;
}
__device__ __forceinline__ float pc_reconstruct(unsigned u16, float scale, float qb) {
    return fmaf(__uint_as_float(0x4B000000u | u16), scale, qb);
}

// MUFU.SQRT: for the centre heuristics only (WHERE a sphere's centre goes; every lane evaluates the same inputs and gets the
// same bits).  The IEEE sqrtf + division of the refinement steps were a quarter of the node kernel's instructions.
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// One bounding-sphere node, built by kNode CONSECUTIVE LANES: the lane of sorted position i = n * kNode + s loads triplet f
// (or -1 = padding) and writes its point records; centroid, member count and radius are butterfly all-reductions inside
// the lane group (x + y == y + x exactly, so every lane of a group holds the same bits); lane s == 0 writes the node
// record.  Must be called by converged warps (whole groups, full mask).  Returns the node radius in lane s == 0.
template <int kNode>
// pt_stride / grp_stride: float4 per node of the triplet records and per group of 4 node records: kNode + 1 and 5 (one pad: odd
// strides, conflict-free when the records are read from shared memory, and -- measured -- also the better layout for the
// scattered per-lane gathers from L2 of the large-cloud modes).
// pc != nullptr: the point-0 records are written in the compressed form above (uint4 units, stride 1 + kNode / 2) instead of
// pt_base; P_up >= the cloud's extent (for the error bound a record must meet).
__device__ __forceinline__ float make_node_coop(const float *__restrict__ tri, float th, int f, long long i, float E_up,
                                                float4 *pt_base, float4 *pt12, float4 *node4, int ball_iters, int pt_stride,
                                                int grp_stride, uint4 *pc = nullptr, float P_up = 0.f, float4 *out_sphere = nullptr) {
    const long long n = i / kNode;
    const int s = (int)(i % kNode);
    double px = 0, py = 0, pz = 0, cut = 0;
    float4 pr = make_float4(0.f, 0.f, 0.f, -INFINITY);                // padding: never a candidate
    float4 pr1 = pr, pr2 = pr;
    if (f >= 0) {
        const float *t = tri + (long long)f * 9;
        px = __ldg(t); py = __ldg(t + 1); pz = __ldg(t + 2);
        const double ax = __ldg(t + 3), ay = __ldg(t + 4), az = __ldg(t + 5);
        const double bx = __ldg(t + 6), by = __ldg(t + 7), bz = __ldg(t + 8);
        cut = (double)th * (double)th - (double)kAddEps;
        pr = make_float4((float)px, (float)py, (float)pz, (float)(cut - (px * px + py * py + pz * pz)));
        pr1 = make_float4((float)ax, (float)ay, (float)az, (float)(cut - (ax * ax + ay * ay + az * az)));
        pr2 = make_float4((float)bx, (float)by, (float)bz, (float)(cut - (bx * bx + by * by + bz * bz)));
        // a triplet whose point 0 or threshold is NaN / infinite can never hit: it keeps its (inert) records but stays
        // out of the sphere, which would otherwise become NaN and hide the node's other triplets
        if (!(fabs(px) + fabs(py) + fabs(pz) + fabs(cut) < (double)INFINITY)) { f = -1; px = py = pz = cut = 0; }
    }
    // centroid in float: WHERE the centre goes is a heuristic, the radius below is an upper bound for whatever centre
    float cx = (float)px, cy = (float)py, cz = (float)pz;            // 0 for padding lanes
    int cntv = f >= 0;
#pragma unroll
    for (int d = 1; d < kNode; d <<= 1) {
        cx += __shfl_xor_sync(0xffffffffu, cx, d);
        cy += __shfl_xor_sync(0xffffffffu, cy, d);
        cz += __shfl_xor_sync(0xffffffffu, cz, d);
        cntv += __shfl_xor_sync(0xffffffffu, cntv, d);
    }
    float qx = 0.f, qy = 0.f, qz = 0.f;
    float R = 0.f;
    if (cntv > 0) {
        const float inv = 1.0f / (float)cntv;
        qx = cx * inv; qy = cy * inv; qz = cz * inv;
    }
    // Shrink the sphere: the centre that minimises max_f (|p0_f - q| + r_f) instead of the centroid (Badoiu-Clarkson
    // steps q += (p_far - q) / (k + 1) towards the member that currently defines the radius; the best centre seen is
    // kept).  About a quarter fewer (line, node) candidates on surface clouds.  Only a heuristic for WHERE the centre
    // goes: the radius below is computed for whatever centre comes out, so the superset property is untouched.  Every
    // lane of the group computes the same bits (same operations on the same shuffled values).
    if (ball_iters > 0) {
        const float fx = (float)px, fy = (float)py, fz = (float)pz;
        const float rf = f >= 0 ? sqrtf(fmaxf((float)cut + E_up, 0.f)) : 0.f;
        const unsigned gmask = (kNode >= 32 ? 0xffffffffu : ((1u << kNode) - 1u)) << ((threadIdx.x & 31) & ~(kNode - 1));
        float bx = qx, by = qy, bz = qz, bestR = INFINITY;
        float ccx = qx, ccy = qy, ccz = qz;
        for (int k = 1; k <= ball_iters + 1; ++k) {
            const float dx = fx - ccx, dy = fy - ccy, dz = fz - ccz;
            const float d = f >= 0 ? sqrt_approx(dx * dx + dy * dy + dz * dz) + rf : -INFINITY;
            float m = d;
#pragma unroll
            for (int dd = 1; dd < kNode; dd <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, dd));
            if (m < bestR) { bestR = m; bx = ccx; by = ccy; bz = ccz; }
            const unsigned bal = __ballot_sync(0xffffffffu, d == m) & gmask;
            const int far = bal ? __ffs(bal) - 1 : (int)(threadIdx.x & 31);
            const float gx = __shfl_sync(0xffffffffu, fx, far), gy = __shfl_sync(0xffffffffu, fy, far), gz = __shfl_sync(0xffffffffu, fz, far);
            const float step = __fdividef(1.0f, (float)(k + 1));
            ccx += (gx - ccx) * step; ccy += (gy - ccy) * step; ccz += (gz - ccz) * step;
        }
        if (cntv > 0 && bestR < INFINITY) { qx = bx; qy = by; qz = bz; }
    }
    if (cntv > 0 && f >= 0) R = cover_radius_up(th, E_up, (float)px, (float)py, (float)pz, qx, qy, qz);
#pragma unroll
    for (int d = 1; d < kNode; d <<= 1) R = fmaxf(R, __shfl_xor_sync(0xffffffffu, R, d));
    pt12[i * 2] = pr1;
    pt12[i * 2 + 1] = pr2;
    if (pc) {
        // scale: the largest offset component of the node's members / kPcSteps
        float am = f >= 0 ? fmaxf(fmaxf(fabsf((float)px - qx), fabsf((float)py - qy)), fabsf((float)pz - qz)) : 0.f;
#pragma unroll
        for (int d = 1; d < kNode; d <<= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, d));
        float scale = am * (1.0f / kPcSteps);
        if (!(scale > 1e-30f)) scale = 1e-30f;
        const float qbx = fmaf(-kPcBias, scale, qx), qby = fmaf(-kPcBias, scale, qy), qbz = fmaf(-kPcBias, scale, qz);
        unsigned ux = 32768u, uy = 32768u, uz = 32768u;
        __half cuth = __float2half_ru(-INFINITY);                    // padding / non-finite triplet: never a candidate
        if (f >= 0) {
            const double inv = 1.0 / (double)scale;                    // (one double division per lane instead of three)
            auto quant = [&](double pv, float qb) -> unsigned {        // nearest representable offset (double: p - qb is ~2^23 steps)
                double u = rint((pv - (double)qb) * inv) - 8388608.0;
                u = u < 0.0 ? 0.0 : (u > 65535.0 ? 65535.0 : u);
                return (unsigned)u;
            };
            ux = quant(px, qbx); uy = quant(py, qby); uz = quant(pz, qbz);
            const float rx = pc_reconstruct(ux, scale, qbx), ry = pc_reconstruct(uy, scale, qby), rz = pc_reconstruct(uz, scale, qbz);
            const float fx = (float)px, fy = (float)py, fz = (float)pz;
            const float dx = fmaxf(fabsf(__fsub_ru(rx, fx)), fabsf(__fsub_rd(rx, fx)));
            const float dy = fmaxf(fabsf(__fsub_ru(ry, fy)), fabsf(__fsub_rd(ry, fy)));
            const float dz = fmaxf(fabsf(__fsub_ru(rz, fz)), fabsf(__fsub_rd(rz, fz)));
            const float e = __fadd_ru(__fsqrt_ru(__fmaf_ru(dz, dz, __fmaf_ru(dy, dy, __fmul_ru(dx, dx)))), 1e-37f);
            const float cut_up = __fadd_ru(__fmul_ru(th, th), -kAddEps);
            const float a = __fsqrt_ru(fmaxf(__fadd_ru(cut_up, E_up), 0.f));
            float cut_c = __fadd_ru(cut_up, __fmaf_ru(__fmul_ru(2.0f, e), a, __fmul_ru(e, e)));
            if (!(e <= pc_err_bound(R, P_up))) cut_c = INFINITY;     // (cannot happen by the bound's derivation; kept exact anyway)
            cuth = __float2half_ru(cut_c);
        }
        // the pair (s, s ^ 1) shares a uint4: the even lane gathers its neighbour's words and writes all 16 bytes
        const unsigned short ch = __half_as_ushort(cuth);
        const unsigned ox = __shfl_xor_sync(0xffffffffu, ux, 1), oy = __shfl_xor_sync(0xffffffffu, uy, 1), oz = __shfl_xor_sync(0xffffffffu, uz, 1);
        const unsigned oc = __shfl_xor_sync(0xffffffffu, (unsigned)ch, 1);
        uint4 *blk = pc + n * (1 + kNode / 2);
        if ((s & 1) == 0) blk[1 + (s >> 1)] = make_uint4(ux | (ox << 16), uy | (oy << 16), uz | (oz << 16), (unsigned)ch | (oc << 16));
        if (s == 0) blk[0] = make_uint4(__float_as_uint(qbx), __float_as_uint(qby), __float_as_uint(qbz), __float_as_uint(scale));
    } else {   // pair-interleaved like the node records: triplets (2i, 2i+1) -> {xA,xB,yA,yB}{zA,zB,wA,wB};
        // a node occupies kNode + 1 float4 (odd stride: lanes reading different nodes hit different banks)
        float *dp = reinterpret_cast<float *>(pt_base + n * pt_stride + (s & ~1)) + (s & 1);
        dp[0] = pr.x; dp[2] = pr.y; dp[4] = pr.z; dp[6] = pr.w;
    }
    float rad = 0.f;
    if (s == 0) {
        float4 rec = make_float4(0.f, 0.f, 0.f, -INFINITY);           // empty node: never a candidate
        if (cntv > 0) rec = sphere_record_up(R, qx, qy, qz, rad);
        if (!pc && pt_stride > kNode) pt_base[n * pt_stride + kNode] = make_float4(0.f, 0.f, 0.f, 0.f);           // pad slot
        // node records: a group of 4 nodes = two interleaved pairs (+ one pad = 5 float4: odd stride again)
        float4 *grp = node4 + (n >> 2) * grp_stride;
        float *dst = reinterpret_cast<float *>(grp + ((n >> 1) & 1) * 2) + (n & 1);
        dst[0] = rec.x; dst[2] = rec.y; dst[4] = rec.z; dst[6] = rec.w;
        if (grp_stride > 4) reinterpret_cast<float *>(grp + 4)[n & 3] = rad;      // the node's radius (level 1 forms its |u| > 1 slack from it)
        if (out_sphere) *out_sphere = make_float4(qx, qy, qz, cntv > 0 ? rad : -1.f);      // -1: empty node
    }
    return rad;
}

__device__ __forceinline__ int pad_supers_dev(int nfp) { return ((nfp / kSuperPts + kNodePad - 1) / kNodePad) * kNodePad; }

// slack of the node radius for the rounding of the reference-order test (rrl_common.cuh): E = kMargin eps (kRefPX (P + Xmax)^2 +
// kRefT thr_max^2)
__device__ __forceinline__ float node_slack(unsigned pmax_bits, unsigned xmax_bits, unsigned tmax_bits) {   // rounded up throughout
    const float P = __fmul_ru(__fsqrt_ru(__uint_as_float(pmax_bits)), 1.000001f);
    const float Xm = __fmul_ru(__fsqrt_ru(__uint_as_float(xmax_bits)), 1.000001f);
    const float PX = __fadd_ru(P, Xm);
    const float T = __uint_as_float(tmax_bits);
    const float sum = __fadd_ru(__fmul_ru(__fmul_ru(kRefPX, PX), PX), __fmul_ru(__fmul_ru(kRefT, T), T));
    return __fadd_ru(__fmul_ru(kMargin * kEps24, sum), 1e-12f);
}

// Super node = bounding sphere of the kSuperPts = 256 sorted triplets one node_kernel CTA handles per trip (16 nodes of
// 16): the same construction as a node -- centre refined towards the minimum enclosing ball, radius
// R >= sqrt(cut_f + E) + |p0_f - q| for every member, so "member passes => super node passes" holds by the same
// argument -- with block-wide instead of lane-group reductions.  Every thread computes the same centre
// (the partial results are combined in the same order by all).  Thread 0 writes the record.
__device__ __forceinline__ float make_super_block(const float *__restrict__ tri, float th, int f, long long i, float E_up,
                                                  float4 *super4, int nsuper, int nsuperp, int ball_iters, float4 &out_centre) {
    __shared__ float s_sum[8][4];
    __shared__ float s_far[2][8][4];
    __shared__ float s_rad[8];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double px = 0, py = 0, pz = 0, cut = 0;
    if (f >= 0) {
        const float *t = tri + (long long)f * 9;
        px = __ldg(t); py = __ldg(t + 1); pz = __ldg(t + 2);
        cut = (double)th * (double)th - (double)kAddEps;
        if (!(fabs(px) + fabs(py) + fabs(pz) + fabs(cut) < (double)INFINITY)) { f = -1; px = py = pz = cut = 0; }   // as in make_node_coop
    }
    float cx = (float)px, cy = (float)py, cz = (float)pz, cn = f >= 0 ? 1.f : 0.f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        cx += __shfl_xor_sync(0xffffffffu, cx, d);
        cy += __shfl_xor_sync(0xffffffffu, cy, d);
        cz += __shfl_xor_sync(0xffffffffu, cz, d);
        cn += __shfl_xor_sync(0xffffffffu, cn, d);
    }
    if (lane == 0) { s_sum[wid][0] = cx; s_sum[wid][1] = cy; s_sum[wid][2] = cz; s_sum[wid][3] = cn; }
    __syncthreads();
    cx = cy = cz = cn = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { cx += s_sum[w][0]; cy += s_sum[w][1]; cz += s_sum[w][2]; cn += s_sum[w][3]; }
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (cn > 0) { qx = cx / cn; qy = cy / cn; qz = cz / cn; }
    if (ball_iters > 0 && cn > 0) {                              // block-uniform condition
        const float fx = (float)px, fy = (float)py, fz = (float)pz;
        const float rf = f >= 0 ? sqrtf(fmaxf((float)cut + E_up, 0.f)) : 0.f;
        float bx = qx, by = qy, bz = qz, bestR = INFINITY;
        float ccx = qx, ccy = qy, ccz = qz;
        const int steps = ball_iters / 2 + 1;                      // block-wide steps cost a barrier each: half of the nodes' count
        for (int k = 1; k <= steps; ++k) {
            const float dx = fx - ccx, dy = fy - ccy, dz = fz - ccz;
            const float d = f >= 0 ? sqrt_approx(dx * dx + dy * dy + dz * dz) + rf : -INFINITY;
            float m = d;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, dd));
            const unsigned bal = __ballot_sync(0xffffffffu, d == m);
            const int far = bal ? __ffs(bal) - 1 : 0;
            float (*sf)[4] = s_far[k & 1];                        // double buffered: one barrier per step
            if (lane == far) { sf[wid][0] = m; sf[wid][1] = fx; sf[wid][2] = fy; sf[wid][3] = fz; }
            __syncthreads();
            // block maximum of the 8 warp maxima (distances are >= 0 or -inf: as unsigned bits they order like floats
            // once -inf is mapped to 0), first warp on ties: one load, one REDUX, one ballot per thread
            const float mw = sf[lane & 7][0];
            const unsigned mb = mw > 0.f ? __float_as_uint(mw) : 0u;
            const unsigned gb = __reduce_max_sync(0xffffffffu, mb);
            const int gw = __ffs(__ballot_sync(0xffffffffu, mb == gb) & 0xffu) - 1;
            const float gm = __uint_as_float(gb);
            if (gm < bestR) { bestR = gm; bx = ccx; by = ccy; bz = ccz; }
            const float step = __fdividef(1.0f, (float)(k + 1));
            ccx += (sf[gw][1] - ccx) * step; ccy += (sf[gw][2] - ccy) * step; ccz += (sf[gw][3] - ccz) * step;
        }
        if (bestR < INFINITY) { qx = bx; qy = by; qz = bz; }
    }
    float R = 0.f;
    if (f >= 0) R = cover_radius_up(th, E_up, (float)px, (float)py, (float)pz, qx, qy, qz);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) R = fmaxf(R, __shfl_xor_sync(0xffffffffu, R, d));
    if (lane == 0) s_rad[wid] = R;
    __syncthreads();
    float rad = 0.f;
    if (tid == 0) {
#pragma unroll
        for (int w = 0; w < 8; ++w) R = fmaxf(R, s_rad[w]);
        const long long n = i / kSuperPts;
        float4 rec = make_float4(0.f, 0.f, 0.f, -INFINITY);      // empty: never a candidate
        if (cn > 0) rec = sphere_record_up(R, qx, qy, qz, rad);
        auto put = [&](long long m, const float4 &r) {
            float4 *grp = super4 + (m >> 2) * 5;
            float *dst = reinterpret_cast<float *>(grp + ((m >> 1) & 1) * 2) + (m & 1);
            dst[0] = r.x; dst[2] = r.y; dst[4] = r.z; dst[6] = r.w;
            if ((m & 3) == 0) grp[4] = make_float4(0.f, 0.f, 0.f, 0.f);
        };
        put(n, rec);
        if (n == nsuper - 1)                                     // sentinels up to the multiple of 4
            for (long long m = nsuper; m < nsuperp; ++m) put(m, make_float4(0.f, 0.f, 0.f, -INFINITY));
    }
    out_centre = make_float4(qx, qy, qz, cn > 0 ? 1.f : 0.f);     // the same bits in every thread
    __syncthreads();                                             // shared buffers are reused by the next trip
    return rad;
}

// The 16 node spheres of one super node as ONE 144-byte record {qb, scale} + 16 x {3 x 16-bit offsets, half radius} (pairs of
// nodes interleaved like the compressed triplet records): level 1 of the super-node mode reads one such record per fired
// (line, super node) pair instead of four 80-byte group records.  Same construction as the triplet records: the centre is
// reconstructed as q~ = fma(float(2^23 + u), scale, qb), e >= |q~ - q| is MEASURED here and added to the radius, which is then
// rounded up to half -- a sphere of radius R~ around q~ contains the node's sphere of radius R around q, so "node passes =>
// compressed node passes" in exact arithmetic; an empty node gets a NaN radius (no comparison with it holds).
// Called by all 256 threads of a trip; `sph` is valid in the first lane of every node group.
__device__ __forceinline__ void write_super_nodes(uint4 *blk, const float4 &sph, bool node_lane, int jn, const float4 &ctr) {
    __shared__ float4 s_sph[16];
    __shared__ float s_scale;
    const int tid = threadIdx.x;
    if (node_lane) s_sph[jn] = sph;
    __syncthreads();
    if (tid < 32) {
        float am = 0.f;
        if (tid < 16 && s_sph[tid].w >= 0.f)
            am = fmaxf(fmaxf(fabsf(s_sph[tid].x - ctr.x), fabsf(s_sph[tid].y - ctr.y)), fabsf(s_sph[tid].z - ctr.z));
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, d));
        float scale = am * (1.0f / kPcSteps);
        if (!(scale > 1e-30f)) scale = 1e-30f;
        if (tid == 0) s_scale = scale;
    }
    __syncthreads();
    if (tid < 16) {
        const float scale = s_scale;
        const float qbx = fmaf(-kPcBias, scale, ctr.x), qby = fmaf(-kPcBias, scale, ctr.y), qbz = fmaf(-kPcBias, scale, ctr.z);
        const float4 sp = s_sph[tid];
        unsigned ux = 32768u, uy = 32768u, uz = 32768u;
        unsigned short hb = 0x7E00u;                                  // half NaN: an empty node never fires
        if (sp.w >= 0.f) {
            const double inv = 1.0 / (double)scale;
            auto quant = [&](float qv, float qb) -> unsigned {
                double u = rint(((double)qv - (double)qb) * inv) - 8388608.0;
                u = u < 0.0 ? 0.0 : (u > 65535.0 ? 65535.0 : u);
                return (unsigned)u;
            };
            ux = quant(sp.x, qbx); uy = quant(sp.y, qby); uz = quant(sp.z, qbz);
            const float rx = pc_reconstruct(ux, scale, qbx), ry = pc_reconstruct(uy, scale, qby), rz = pc_reconstruct(uz, scale, qbz);
            const float dx = fmaxf(fabsf(__fsub_ru(rx, sp.x)), fabsf(__fsub_rd(rx, sp.x)));
            const float dy = fmaxf(fabsf(__fsub_ru(ry, sp.y)), fabsf(__fsub_rd(ry, sp.y)));
            const float dz = fmaxf(fabsf(__fsub_ru(rz, sp.z)), fabsf(__fsub_rd(rz, sp.z)));
            const float e = __fadd_ru(__fsqrt_ru(__fmaf_ru(dz, dz, __fmaf_ru(dy, dy, __fmul_ru(dx, dx)))), 1e-37f);
            hb = __half_as_ushort(__float2half_ru(__fadd_ru(sp.w, e)));
        }
        unsigned short *rec16 = reinterpret_cast<unsigned short *>(blk + 1 + (tid >> 1));
        const int hs = tid & 1;
        rec16[0 + hs] = (unsigned short)ux; rec16[2 + hs] = (unsigned short)uy; rec16[4 + hs] = (unsigned short)uz; rec16[6 + hs] = hb;
        if (tid == 0) blk[0] = make_uint4(__float_as_uint(qbx), __float_as_uint(qby), __float_as_uint(qbz), __float_as_uint(scale));
    }
    __syncthreads();                                             // s_sph is reused by the next trip
}

// (the prefetch / preload / half-record switches of the level-1 and level-2 gathers -- RRL_PF_LEVEL1/2, RRL_L2_PRELOAD,
// RRL_EXP_HALF_L2 -- were measurement variants of the fp32 records; DESIGN 5.1 has their numbers, the git history their code)
#ifndef RRL_SUPER_MINBLOCKS
#define RRL_SUPER_MINBLOCKS 0
#endif
#ifndef RRL_SUPER_FROM_NODES
#define RRL_SUPER_FROM_NODES 1                // super-node spheres from the 16 node spheres (one warp) instead of the 256 points (block-wide)
#endif
#ifndef RRL_NODE_MINBLOCKS
#define RRL_NODE_MINBLOCKS 4                  // 64 registers: 4 CTAs per SM (139 registers = ONE CTA per SM took 196 us at 500k, 4 take 84)
#endif
// Super node of 16 nodes FROM THE NODE SPHERES (one warp; the 256-point version above costs a quarter of the node kernel in
// block-wide reductions and barriers): centre = refined centroid of the node centres, R >= |q_n - Q| + R_n for every node, so
// a line within R_n of q_n is within R of Q -- "node passes => super node passes" by the triangle inequality, and with it
// "member passes => super node passes" (the |u| > 1 slack has the same form, with the cloud's largest super radius).  A few
// per cent looser than the sphere of the points themselves.  Also writes the super node's compressed node record
// (write_super_nodes' format).  Called by all 256 threads of a trip; returns the super radius in thread 0.
__device__ __forceinline__ float super_from_nodes(float4 *super4, uint4 *blk, long long n, int nsuper, int nsuperp, const float4 &sph,
                                                  bool node_lane, int jn, int ball_iters) {
    __shared__ float4 s_sph[16];
    const int tid = threadIdx.x, lane = tid & 31;
    if (node_lane) s_sph[jn] = sph;
    __syncthreads();
    float rad = 0.f;
    if (tid < 32) {
        const float4 sp = lane < 16 ? s_sph[lane] : make_float4(0.f, 0.f, 0.f, -1.f);
        const bool valid = sp.w >= 0.f;
        const int cnt = __popc(__ballot_sync(0xffffffffu, valid));
        float cx = valid ? sp.x : 0.f, cy = valid ? sp.y : 0.f, cz = valid ? sp.z : 0.f;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            cx += __shfl_xor_sync(0xffffffffu, cx, d);
            cy += __shfl_xor_sync(0xffffffffu, cy, d);
            cz += __shfl_xor_sync(0xffffffffu, cz, d);
        }
        float Qx = 0.f, Qy = 0.f, Qz = 0.f;
        if (cnt > 0) { const float inv = __fdividef(1.0f, (float)cnt); Qx = cx * inv; Qy = cy * inv; Qz = cz * inv; }
        if (cnt > 0 && ball_iters > 0) {                          // centre heuristics (as in make_node_coop)
            float bx = Qx, by = Qy, bz = Qz, bestR = INFINITY, ccx = Qx, ccy = Qy, ccz = Qz;
            for (int k = 1; k <= ball_iters + 1; ++k) {
                const float dx = sp.x - ccx, dy = sp.y - ccy, dz = sp.z - ccz;
                const float d = valid ? sqrt_approx(dx * dx + dy * dy + dz * dz) + sp.w : -INFINITY;
                float m = d;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, dd));
                if (m < bestR) { bestR = m; bx = ccx; by = ccy; bz = ccz; }
                const unsigned bal = __ballot_sync(0xffffffffu, valid && d == m);
                const int far = bal ? __ffs(bal) - 1 : 0;
                const float gx = __shfl_sync(0xffffffffu, sp.x, far), gy = __shfl_sync(0xffffffffu, sp.y, far), gz = __shfl_sync(0xffffffffu, sp.z, far);
                const float step = __fdividef(1.0f, (float)(k + 1));
                ccx += (gx - ccx) * step; ccy += (gy - ccy) * step; ccz += (gz - ccz) * step;
            }
            if (bestR < INFINITY) { Qx = bx; Qy = by; Qz = bz; }
        }
        // radius and the scale of the compressed record, rounded up throughout
        float R = 0.f, am = 0.f;
        if (valid) {
            const float dx = fmaxf(fabsf(__fsub_ru(sp.x, Qx)), fabsf(__fsub_rd(sp.x, Qx)));
            const float dy = fmaxf(fabsf(__fsub_ru(sp.y, Qy)), fabsf(__fsub_rd(sp.y, Qy)));
            const float dz = fmaxf(fabsf(__fsub_ru(sp.z, Qz)), fabsf(__fsub_rd(sp.z, Qz)));
            R = __fadd_ru(__fsqrt_ru(__fmaf_ru(dz, dz, __fmaf_ru(dy, dy, __fmul_ru(dx, dx)))), sp.w);
            am = fmaxf(fmaxf(dx, dy), dz);
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            R = fmaxf(R, __shfl_xor_sync(0xffffffffu, R, d));
            am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, d));
        }
        if (lane == 0) {
            float4 rec = make_float4(0.f, 0.f, 0.f, -INFINITY);      // empty: never a candidate
            if (cnt > 0) rec = sphere_record_up(R, Qx, Qy, Qz, rad);
            auto put = [&](long long m, const float4 &r, float rr) {
                float4 *grp = super4 + (m >> 2) * 5;
                float *dst = reinterpret_cast<float *>(grp + ((m >> 1) & 1) * 2) + (m & 1);
                dst[0] = r.x; dst[2] = r.y; dst[4] = r.z; dst[6] = r.w;
                reinterpret_cast<float *>(grp + 4)[m & 3] = rr;
            };
            put(n, rec, rad);
            if (n == nsuper - 1)                                     // sentinels up to the multiple of 16
                for (long long m = nsuper; m < nsuperp; ++m) put(m, make_float4(0.f, 0.f, 0.f, -INFINITY), 0.f);
        }
        if (blk) {                                                   // compressed node record (see write_super_nodes)
            float scale = am * (1.0f / kPcSteps);
            if (!(scale > 1e-30f)) scale = 1e-30f;
            const float qbx = fmaf(-kPcBias, scale, Qx), qby = fmaf(-kPcBias, scale, Qy), qbz = fmaf(-kPcBias, scale, Qz);
            if (lane < 16) {
                unsigned ux = 32768u, uy = 32768u, uz = 32768u;
                unsigned short hb = 0x7E00u;                          // half NaN: an empty node never fires
                if (valid) {
                    const double inv = 1.0 / (double)scale;
                    auto quant = [&](float qv, float qb) -> unsigned {
                        double u = rint(((double)qv - (double)qb) * inv) - 8388608.0;
                        u = u < 0.0 ? 0.0 : (u > 65535.0 ? 65535.0 : u);
                        return (unsigned)u;
                    };
                    ux = quant(sp.x, qbx); uy = quant(sp.y, qby); uz = quant(sp.z, qbz);
                    const float rx = pc_reconstruct(ux, scale, qbx), ry = pc_reconstruct(uy, scale, qby), rz = pc_reconstruct(uz, scale, qbz);
                    const float dx = fmaxf(fabsf(__fsub_ru(rx, sp.x)), fabsf(__fsub_rd(rx, sp.x)));
                    const float dy = fmaxf(fabsf(__fsub_ru(ry, sp.y)), fabsf(__fsub_rd(ry, sp.y)));
                    const float dz = fmaxf(fabsf(__fsub_ru(rz, sp.z)), fabsf(__fsub_rd(rz, sp.z)));
                    const float e = __fadd_ru(__fsqrt_ru(__fmaf_ru(dz, dz, __fmaf_ru(dy, dy, __fmul_ru(dx, dx)))), 1e-37f);
                    hb = __half_as_ushort(__float2half_ru(__fadd_ru(sp.w, e)));
                }
                unsigned short *rec16 = reinterpret_cast<unsigned short *>(blk + 1 + (lane >> 1));
                const int hs = lane & 1;
                rec16[0 + hs] = (unsigned short)ux; rec16[2 + hs] = (unsigned short)uy; rec16[4 + hs] = (unsigned short)uz; rec16[6 + hs] = hb;
                if (lane == 0) blk[0] = make_uint4(__float_as_uint(qbx), __float_as_uint(qby), __float_as_uint(qbz), __float_as_uint(scale));
            }
        }
    }
    __syncthreads();                                             // s_sph is reused by the next trip
    return rad;
}

template <int kNode>
__global__ void __launch_bounds__(256, RRL_NODE_MINBLOCKS) node_kernel(const float *__restrict__ tri1, const float *__restrict__ tri2, Workspace ws, Geometry g, int ball_iters,
                                                   int supers, int reuse_target, int compressed) {
    const int b = blockIdx.y >> 1, cloud = blockIdx.y & 1;
    const int nf = cloud ? g.nf2 : g.nf1, nfp = cloud ? g.nf2p : g.nf1p;
    const int nnodes = nfp / kNode;
    const float *tri = (cloud ? tri2 : tri1) + (long long)b * nf * 9;
    const float *thr = ws.thr[cloud] + (long long)b * nf;
    const int *perm = ws.perm[cloud] + (long long)b * nfp;
    unsigned xmax_bits = ws.xmax[b * 2];                  // final: prep_kernel has completed
    if (cloud == 1 && reuse_target) {
        // the target's records are valid for every line extent up to the one they were built for (the slack E grows with it):
        // keep[4] is only rewritten by the LAST kernel of a forward, so all CTAs of cloud 2 decide alike
        if (ws.hdr[7] == order_token(g) && xmax_bits <= ws.keep[b * 8 + 4]) {
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                atomicMax(ws.rmax + b * 2 + 1, ws.keep[b * 8 + 1]);
                atomicMax(ws.smax + b * 2 + 1, ws.keep[b * 8 + 2]);
                ws.xmax[b * 2 + 1] = ws.keep[b * 8 + 4];          // the records stay valid for the extent they were built for
            }
            return;
        }
        // (re)build for a 1.1x larger extent, so that the next steps' lines (resampled every step) rarely force another rebuild
        xmax_bits = __float_as_uint(__fmul_ru(__uint_as_float(xmax_bits), 1.21f));
        if (blockIdx.x == 0 && threadIdx.x == 0) ws.xmax[b * 2 + 1] = xmax_bits;      // -> keep[4] at the end of this forward
    }
    const float E = node_slack(ws.pmax[b * 2 + cloud], xmax_bits, ws.tmax[b * 2 + cloud]);
    const float P_up = __fmul_ru(__fsqrt_ru(__uint_as_float(ws.pmax[b * 2 + cloud])), 1.000001f);
    // compressed point-0 records live in the pt4 buffer (16 + 8 kNode <= 16 (kNode + 1) bytes per node)
    uint4 *pc = compressed ? reinterpret_cast<uint4 *>(ws.pt4[cloud]) + (long long)b * nnodes * (1 + kNode / 2) : nullptr;
    // one thread per sorted position; nfp is a multiple of kPointPad = 256 = blockDim.x, so every warp is full
    float rad = 0.f, srad = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nfp; i += (long long)gridDim.x * blockDim.x) {
        int f = perm[i];
        if ((unsigned)f >= (unsigned)nf) f = -1;      // padding; also keeps a violated RRL_REUSE_ORDER contract memory-safe
        const float th = f >= 0 ? thr[f] : 0.f;
        // (sector-aligned strides -- kNode and 4 -- were measured for the super-node mode, whose records are only ever gathered
        // from L2: 25 % SLOWER, 1354 against 1083 us on the large pair.  With every node at the same offset modulo 256 bytes the
        // lanes of a gather collide in the same L1 sets / L2 slices; the odd strides spread them.)
        const int pt_stride = kNode + 1, grp_stride = 5;
        (void)supers;
        float4 sph = make_float4(0.f, 0.f, 0.f, -1.f), ctr;
        rad = fmaxf(rad, make_node_coop<kNode>(tri, th, f, i, E, ws.pt4[cloud] + (long long)b * nnodes * pt_stride,
                                               ws.pt12[cloud] + (long long)b * nfp * 2, ws.node4[cloud] + (long long)b * (nnodes / 4) * grp_stride,
                                               ball_iters, pt_stride, grp_stride, pc, P_up, &sph));
        if (supers) {
            float4 *sup = ws.super4[cloud] + (long long)b * (pad_supers_dev(nfp) / 4) * 5;
            if constexpr (kNode == 16 && RRL_SUPER_FROM_NODES) {   // (a super node = 16 nodes; use_supers() requires kNode == 16)
                srad = fmaxf(srad, super_from_nodes(sup, ws.sn8[cloud] + ((long long)b * pad_supers_dev(nfp) + i / kSuperPts) * 9, i / kSuperPts,
                                                    nfp / kSuperPts, pad_supers_dev(nfp), sph, (i % kNode) == 0,
                                                    (int)((i / kNode) % (kSuperPts / kNode)), ball_iters));
                (void)ctr;
            } else {
                srad = fmaxf(srad, make_super_block(tri, th, f, i, E, sup, nfp / kSuperPts, pad_supers_dev(nfp), ball_iters, ctr));
                if constexpr (kNode == 16)
                    write_super_nodes(ws.sn8[cloud] + ((long long)b * pad_supers_dev(nfp) + i / kSuperPts) * 9, sph, (i % kNode) == 0,
                                      (int)((i / kNode) % (kSuperPts / kNode)), ctr);
            }
        }
    }
    // one block-level maximum and one atomic per CTA (thousands of same-address atomics serialise in L2)
    __shared__ unsigned s_radm;
    if (threadIdx.x == 0) s_radm = 0u;
    __syncthreads();
    const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(rad));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(&s_radm, m);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_radm) atomicMax(ws.rmax + b * 2 + cloud, s_radm);
        if (srad > 0.f) atomicMax(ws.smax + b * 2 + cloud, __float_as_uint(srad));     // thread 0 holds the super radii
    }
}

// ------------------------------------------------------------------------------------------------------
// small clouds (<= kSortSmall padded triplets): memset + prep + sort + nodes in ONE launch
#ifdef RRL_MARKS
// phase timestamps of the cloud-0 CTA of pair 0 (measurement builds only; rrl_debug_read_marks_prep)
__device__ unsigned long long g_marks_prep[32];
__device__ __forceinline__ void pmark(int i) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_marks_prep[i] = t;
    }
}
#else
__device__ __forceinline__ void pmark(int) {}
#endif

// ------------------------------------------------------------------------------------------------------
// grid (B, 2 + line blocks).  blockIdx.y < 2: one CTA per (pair, cloud) runs the whole chain -- thresholds and the
// cloud's extent, the pair's line extent (each cloud CTA scans the pair's lines itself rather than wait for another
// block), Hilbert keys, a bitonic sort in shared memory, the node records -- with block barriers instead of launch
// boundaries.  blockIdx.y >= 2: per-line filter constants and hit counters.  The cloud-0 CTA also zeroes the pair's
// accumulators (what the large-cloud path does with a memset).
__device__ __forceinline__ void line_constants(const float *__restrict__ ln, float4 *lineC2, float &xm) {
    const float u0 = __ldg(ln), u1 = __ldg(ln + 1), u2 = __ldg(ln + 2);
    const double x = __ldg(ln + 3), y = __ldg(ln + 4), z = __ldg(ln + 5);
    const double sd = x * u0 + y * u1 + z * u2;
    const double xx = x * x + y * y + z * z;
    lineC2[0] = make_float4(u0, u1, u2, (float)sqrt(xx) * 1.000001f);
    lineC2[1] = make_float4((float)(2.0 * (x - sd * u0)), (float)(2.0 * (y - sd * u1)), (float)(2.0 * (z - sd * u2)), (float)(xx - sd * sd));
    xm = (float)xx * 1.000001f;
    if (!(xm == xm)) xm = INFINITY;
}

template <int kNode>
__global__ void __launch_bounds__(1024) small_prep_kernel(const float *__restrict__ tri1, const float *__restrict__ tri2,
                                                          const float *__restrict__ lines, Workspace ws, Geometry g, int window,
                                                          int sorted, int line_blocks, int ball_iters, int refine, int reuse, int compressed,
                                                          int allow_handoff) {
    extern __shared__ unsigned long long skeys[];
    __shared__ unsigned s_red[4];                        // bits of max |p|^2, max |x0|^2 (scaled), max node radius, max threshold
    const int b = blockIdx.x, tid = threadIdx.x;
    pmark(0);
    // RRL_REUSE_ORDER is honoured only when an earlier forward of this very geometry completed in this workspace (hdr[7]
    // is written by that forward's LAST kernel and by nobody in this launch, so every CTA reads the same value); on a
    // fresh or differently shaped workspace the cloud is simply sorted
    const bool seen = reuse && ws.hdr[7] == order_token(g);         // (only read when the caller vouches for the workspace: RRL_REUSE_*)
    reuse = seen;
    // Hand-off of the pair's line extent from the line blocks to the cloud CTAs, for SMALL BATCHES inside a SESSION only (the launch
    // allows it for B <= 4 under RRL_REUSE_ORDER / RRL_REUSE_TARGET, i.e. when the caller keeps the workspace untouched between its
    // forwards -- the counter below must survive from one forward to the next, which a workspace handed back to an allocator in between
    // cannot promise): a cloud CTA scanning the pair's 20000 lines itself is 5 us of the demo's 36 us prep kernel, bound by what one SM pulls
    // through L2.  With many pairs the hand-off does not pay (measured: DCP 40.9 against 40.3 us, RPM far worse), so the cloud CTAs scan
    // there.  The cloud CTAs come first in the grid (they are the critical path) and wait for blocks dispatched after them -- all CTAs
    // of such a launch are resident at once -- with a bounded wait and the scan as the fall-back, so the results never depend on the
    // hand-off.  flags[b*4+3] counts the pair's finished line blocks; the second cloud CTA to have read the partials resets it, so the
    // counter is zero again when the kernel ends.  That invariant needs ONE earlier forward of this geometry in this workspace (`seen`,
    // the token RRL_REUSE_ORDER checks: written by a forward's last kernel and by nobody in this launch, so all CTAs decide alike); on a
    // fresh workspace the cloud CTAs scan and the cloud-0 CTA zeroes the counter.
    const bool handoff = seen && allow_handoff;
    const float *lb = lines + (long long)b * g.nl * 6;
    __shared__ unsigned s_lred[2];
    if (blockIdx.y >= 2) {
        const int lblk = blockIdx.y - 2;
        if (tid < 2) s_lred[tid] = 0u;
        __syncthreads();
        float xl = 0.f;
        int badl = 0;
        for (int l = lblk * 1024 + tid; l < g.nl; l += line_blocks * 1024) {
            const long long gl = (long long)b * g.nl + l;
            float xm;
            line_constants(lb + (long long)l * 6, ws.lineC + gl * 2, xm);
            if (handoff) {
                const float *ln = lb + (long long)l * 6;
                float m = xm + 0.f * (__ldg(ln) + __ldg(ln + 1) + __ldg(ln + 2));       // NaN for a NaN / infinite position OR direction
                if (!(m < INFINITY)) { m = 0.f; badl = 1; }                             // such a line never hits: it stays out of the extent
                xl = fmaxf(xl, m);
            }
            ws.cnt[0][gl] = 0;
            ws.cnt[1][gl] = 0;
        }
        if (handoff) {
            const unsigned a = __reduce_max_sync(0xffffffffu, __float_as_uint(xl));
            if ((tid & 31) == 0 && a) atomicMax(&s_lred[0], a);
            const int bad_blk = __syncthreads_or(badl);
            if (tid == 0) {
                unsigned int *lp = ws.lpart + ((long long)b * 64 + lblk) * 2;
                lp[0] = s_lred[0];
                lp[1] = bad_blk ? 1u : 0u;
                __threadfence();
                atomicAdd(ws.flags + b * 4 + 3, 1);
            }
        }
        return;
    }
    const int cloud = blockIdx.y;
    const int nf = cloud ? g.nf2 : g.nf1, nfp = cloud ? g.nf2p : g.nf1p;
    const float *tri = (cloud ? tri2 : tri1) + (long long)b * nf * 9;
    float *thr = ws.thr[cloud] + (long long)b * nf;
    if (tid < 4) s_red[tid] = 0u;
    if (cloud == 0) {
        if (b == 0 && tid == 0) {
            ws.hdr[0] = kMagic; ws.hdr[1] = g.B; ws.hdr[2] = g.nf1; ws.hdr[3] = g.nf2; ws.hdr[4] = g.nl; ws.hdr[5] = window;
            ws.hdr[6] = 0;
            ws.xcursor[0] = 0ull; ws.xcursor[1] = 0ull;
        }
        if (tid < 16) ws.n_kj[b * 16 + tid] = 0;
        if (tid < 32) ws.sums[b * 32 + tid] = 0ull;
        if (tid < RRL_NSTAT) ws.stats[(long long)b * RRL_NSTAT + tid] = 0;
        if (tid < 18) ws.gcounts[b * 18 + tid] = 0;
        if (tid < (handoff ? 3 : 4)) ws.flags[b * 4 + tid] = 0;          // (the hand-off counter belongs to the line blocks then)
        if (tid == 0) { ws.nrec[b] = 0; ws.med[b] = 0.f; ws.xmax[b * 2 + 1] = 0u; }
    }
    __syncthreads();
    pmark(1);
    // thresholds, extent of the cloud
    float pm = 0.f, tm = 0.f;
    int badv = 0;
    for (int f = tid; f < nf; f += 1024) {
        float v[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) v[q] = __ldg(tri + (long long)f * 9 + q);
        const float th = triplet_thr_exact(v);
        thr[f] = th;
        if (th < INFINITY) tm = fmaxf(tm, th);
        pm = fmaxf(pm, finite_extent(v, badv));
    }
    pmark(2);
    // the pair's line extent: here when the launch does not allow the hand-off at all (many pairs: the original order of this kernel),
    // else further down, from the line blocks or by the same scan
    auto scan_lines = [&](float &xm_out, int &bad_out) {
        float xm = 0.f;
#pragma unroll 4
        for (int l = tid; l < g.nl; l += 1024) {
            const float *ln = lb + (long long)l * 6;
            // float suffices: the sum is within 1.8e-7 of |x0|^2 and the factor keeps it an upper bound
            const float x = __ldg(ln + 3), y = __ldg(ln + 4), z = __ldg(ln + 5);
            float m = (x * x + y * y + z * z) * 1.000001f;
            m += 0.f * (__ldg(ln) + __ldg(ln + 1) + __ldg(ln + 2));      // NaN for a NaN / infinite position OR direction
            if (!(m < INFINITY)) { m = 0.f; bad_out = 1; }               // such a line never hits: it stays out of the extent
            xm = fmaxf(xm, m);
        }
        xm_out = xm;
    };
    if (!allow_handoff) {
        float xm = 0.f;
        scan_lines(xm, badv);
        const unsigned c = __reduce_max_sync(0xffffffffu, __float_as_uint(xm));
        if ((tid & 31) == 0 && c) atomicMax(&s_red[1], c);
    }
    {
        const unsigned a = __reduce_max_sync(0xffffffffu, __float_as_uint(pm));
        const unsigned t = __reduce_max_sync(0xffffffffu, __float_as_uint(tm));
        if ((tid & 31) == 0) { atomicMax(&s_red[0], a); atomicMax(&s_red[3], t); }
    }
    const int bad_pts = __syncthreads_or(badv);          // (also the barrier the reductions above need)
    pmark(3);
    if (tid == 0) {
        ws.pmax[b * 2 + cloud] = s_red[0];
        ws.tmax[b * 2 + cloud] = s_red[3];
        ws.smax[b * 2 + cloud] = 0u;                     // (no super nodes on this path; the forward's last kernel copies it to `keep`)
        if (!allow_handoff) {
            if (cloud == 0) ws.xmax[b * 2] = s_red[1];
            ws.bad[b * 2 + cloud] = bad_pts ? 1u : 0u;   // this cloud's points, or any line of the pair (both CTAs scan them)
        }
    }
    // Hilbert order: element i = e * 1024 + tid lives in register v[e]; n2 = E * 1024 >= nfp keys (sentinels sort last).
    // Compare-exchange distances below 32 are warp shuffles, 32..512 go through shared memory (double buffered: one
    // barrier per stage), 1024 and 2048 pair registers of the same thread.
    const float P = sqrtf(__uint_as_float(s_red[0]));
    const int E = nfp <= 1024 ? 1 : (nfp <= 2048 ? 2 : 4);
    // 32-bit keys: (leading bits of the 30-bit curve key) << idx_bits | index -- 22 / 21 / 20 curve bits for up to 1024 / 2048 /
    // 4096 triplets, far below a node's extent; half the shuffles, shared-memory traffic and compare instructions of 64-bit keys.
    // 0xFFFFFFFF = padding: sorts last; its index field is >= nf unless the cloud fills the array, and then there is no padding.
    unsigned v[4];
    const int idx_bits = E == 1 ? 10 : (E == 2 ? 11 : 12);
    unsigned *skeys32 = reinterpret_cast<unsigned *>(skeys);
    int *perm = ws.perm[cloud] + (long long)b * nfp;
    if (reuse) {
        // RRL_REUSE_ORDER: the workspace holds the order of a previous forward of this geometry.  ANY permutation is a
        // valid order (it only decides which triplets share a node); a rigidly moved cloud keeps a good one.
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = e * 1024 + tid;
            v[e] = (e < E && i < nfp) ? (unsigned)perm[i] : 0xFFFFFFFFu;                            // -1 (padding) = 0xFFFFFFFF
        }
    } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int i = e * 1024 + tid;
        v[e] = 0xFFFFFFFFu;
        if (e < E && i < nf) {
            const float p[3] = {__ldg(tri + (long long)i * 9), __ldg(tri + (long long)i * 9 + 1), __ldg(tri + (long long)i * 9 + 2)};
            const unsigned key = sorted ? morton_key(p, P) : 0u;
            v[e] = ((key >> (idx_bits - 2)) << idx_bits) | (unsigned)i;
        }
    }
    }
    if (sorted && !reuse) {
        auto cas = [](unsigned &lo_el, unsigned &hi_el, bool up) {      // lo_el has the lower index
            const unsigned mn = lo_el < hi_el ? lo_el : hi_el, mx = lo_el < hi_el ? hi_el : lo_el;
            lo_el = up ? mn : mx;
            hi_el = up ? mx : mn;
        };
        int buf = 0;
        const int N = E * 1024;
        for (int k = 2; k <= N; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                if (j >= 1024) {
                    // i & k for element (e, tid): k >= 2048 here, so the direction depends on e only
                    if (j == 1024) {
                        cas(v[0], v[1], (0 & k) == 0);
                        if (E > 2) cas(v[2], v[3], (2048 & k) == 0);
                    } else {
                        cas(v[0], v[2], (0 & k) == 0);
                        cas(v[1], v[3], (1024 & k) == 0);
                    }
                } else if (j >= 32) {
                    unsigned *sb = skeys32 + buf * N;
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (e < E) sb[e * 1024 + tid] = v[e];
                    __syncthreads();
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (e < E) {
                            const int i = e * 1024 + tid;
                            const unsigned o = sb[i ^ j];
                            const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
                            v[e] = (v[e] < o) == keep_min ? v[e] : o;
                        }
                    buf ^= 1;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (e < E) {
                            const int i = e * 1024 + tid;
                            const unsigned o = __shfl_xor_sync(0xffffffffu, v[e], j);
                            const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
                            v[e] = (v[e] < o) == keep_min ? v[e] : o;
                        }
                }
            }
    }
    // ---- the pair's line extent: from the line blocks (hand-off), or scanned here ----
    if (allow_handoff) {
        __shared__ int s_have;
        if (tid == 0) {
            int have = 0;
            if (handoff) {
                const volatile int *cnt = ws.flags + b * 4 + 3;
                const long long t0 = clock64();
                while (true) {
                    if ((*cnt & 0xFFFF) == line_blocks) { have = 1; break; }
                    if (clock64() - t0 > 200000000LL) break;         // ~0.1 s: never in a sound run; the scan below keeps the result right
                    __nanosleep(32);
                }
                __threadfence();
            }
            s_have = have;
        }
        __syncthreads();
        int badl = 0;
        if (s_have) {
            if (tid < 32) {
                unsigned x = 0u;
                for (int i = tid; i < line_blocks; i += 32) {
                    x = max(x, __ldcg(ws.lpart + ((long long)b * 64 + i) * 2));
                    badl |= (int)__ldcg(ws.lpart + ((long long)b * 64 + i) * 2 + 1);
                }
                x = __reduce_max_sync(0xffffffffu, x);
                if (tid == 0) s_red[1] = x;
            }
        } else {
            float xm = 0.f;
            scan_lines(xm, badl);
            const unsigned c = __reduce_max_sync(0xffffffffu, __float_as_uint(xm));
            if ((tid & 31) == 0 && c) atomicMax(&s_red[1], c);
        }
        const int bad_lines = __syncthreads_or(badl);    // (also publishes s_red[1])
        if (tid == 0) {
            if (cloud == 0) ws.xmax[b * 2] = s_red[1];
            ws.bad[b * 2 + cloud] = (bad_pts | bad_lines) ? 1u : 0u;     // this cloud's points, or any line of the pair
            if (s_have) {                                // the second reader of the pair leaves the counter at zero for the next forward
                const int old = atomicAdd(ws.flags + b * 4 + 3, 0x10000);
                if ((old >> 16) == 1) atomicExch(ws.flags + b * 4 + 3, 0);
            }
        }
    }
    const int nnodes = nfp / kNode;
    const float Eslack = node_slack(s_red[0], s_red[1], s_red[3]);
    pmark(4);
    __syncthreads();                                     // thr (global) is re-read below by other threads; sort buffers are dead
    // sorted indices -> shared memory, k-d refinement of every window of 64 (one warp each), then the records
    int *sidx = reinterpret_cast<int *>(skeys);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int i = e * 1024 + tid;
        if (e < E && i < nfp) {
            const unsigned idx = v[e] & ((1u << idx_bits) - 1u);
            sidx[i] = idx < (unsigned)nf ? (int)idx : -1;
        }
    }
    __syncthreads();
    pmark(5);
    if (sorted && refine && !reuse) {
        __shared__ float4 s_kd[32][64];
        const int lane = tid & 31, wid = tid >> 5;
        for (int w = wid; w < nfp / 64; w += 32) {
            int f0 = sidx[w * 64 + lane], f1 = sidx[w * 64 + 32 + lane];
            kd_refine64<kNode>(tri, f0, f1, lane, s_kd[wid]);
            sidx[w * 64 + lane] = f0;
            sidx[w * 64 + 32 + lane] = f1;
        }
    }
    __syncthreads();
    pmark(6);
    float rad = 0.f;
    uint4 *pc = compressed ? reinterpret_cast<uint4 *>(ws.pt4[cloud]) + (long long)b * nnodes * (1 + kNode / 2) : nullptr;
    const float P_up = __fmul_ru(__fsqrt_ru(__uint_as_float(s_red[0])), 1.000001f);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int i = e * 1024 + tid;
        if (e < E && i < nfp) {                          // nfp is a multiple of 256: whole warps
            const int f = sidx[i];
            perm[i] = f;
            rad = fmaxf(rad, make_node_coop<kNode>(tri, f >= 0 ? thr[f] : 0.f, f, i, Eslack, ws.pt4[cloud] + (long long)b * nnodes * (kNode + 1),
                                                   ws.pt12[cloud] + (long long)b * nfp * 2, ws.node4[cloud] + (long long)b * (nnodes / 4) * 5, ball_iters,
                                                   kNode + 1, 5, pc, P_up));
        }
    }
    {
        const unsigned a = __reduce_max_sync(0xffffffffu, __float_as_uint(rad));
        if ((tid & 31) == 0 && a) atomicMax(&s_red[2], a);
    }
    __syncthreads();
    if (tid == 0) ws.rmax[b * 2 + cloud] = s_red[2];
    pmark(7);
}

size_t sort_scratch_bytes(int nfp_max, int B) {
    if (nfp_max <= kSortSmall) return 0;
    // keys in/out, values in/out of every cloud of up to kSortPairs pairs + CUB temp (CUB's requirement is checked at
    // launch time)
    const size_t nb = B < kSortPairs ? B : kSortPairs;
    return nb * nfp_max * 2 * 4 * 4 + 1024 + nb * nfp_max * 16 + (4u << 20);
}

// triplets per bounding-sphere node: small clouds are dense in hits per line and want tighter spheres
static int g_param[16] = {0, 0, 16, 32, 0, 0, 0, 0, 8, 0, 1, 0, 0, 0, 4, 0};   // [0] 1 = unfused prep/sort/node launches for small clouds (A/B), [1] node size override, [2] target waves, [3] min nodes per chunk, [4] group-level pushes for small clouds, [5] brute force, [6] lines per thread (2 or 4, 0 = auto), [7] 1 = no super-node level (A/B), [8] enclosing-ball refinement steps of the node centres (0 = centroid), [9] target waves in super-node mode (0 = by the number of line tiles), [10] k-d refinement of the Hilbert order inside windows of 64 (0 = off, 1 = small clouds, 2 = also the large path), [11] entries in flight per thread of the exact kernel (0 = 2), [12] lowest key bit the radix sort of the large path looks at (0 = auto, -1 = every bit), [13] CTAs per SM of the exact kernel's grid (0 = 8), [15] 1 = no line-extent hand-off in the small-cloud prep kernel (A/B), [14] enclosing-ball refinement steps on the large-cloud path (node_kernel: the steps are a third of that kernel, and every rank of a line shard repeats it)
void set_param(int id, int v) { if (id >= 0 && id < 16) g_param[id] = v; }
int node_size(const Geometry &g) {
    if (g_param[1] == 8 || g_param[1] == 16) return g_param[1];
    return (g.nf1 > g.nf2 ? g.nf1 : g.nf2) >= 16384 ? 16 : 8;
}

// the super-node level (one more bounding-sphere level above the nodes) for clouds of kSuperMin triplets and more
int use_supers(const Geometry &g) {
    // measured: pays from about 16k triplets per cloud (node size 16); below, the per-node kernel with its shared-memory
    // point cache is as fast and the node build is cheaper without the extra level
    return g_param[7] == 0 && node_size(g) == 16 && (g.nf1p > g.nf2p ? g.nf1p : g.nf2p) > kSortSmall;
}

// compressed triplet records (make_node_coop) wherever level 2 gathers them from L2 instead of a shared-memory copy
#ifndef RRL_PC8
#define RRL_PC8 1
#endif
#ifndef RRL_SN8
#define RRL_SN8 1
#endif
int use_compressed(const Geometry &g) { return RRL_PC8 && (node_size(g) == 16 || g_param[4] != 0); }

static int g_dense_variant = 1;        // 1 = Morton-sorted nodes (default), 0 = nodes in input order (A/B measurement)
void set_dense_variant(int v) { g_dense_variant = v; }

int launch_prep(const float *tri1, const float *tri2, const float *lines, const Workspace &ws, const Geometry &g,
                int window, int reuse_flags, cudaStream_t s) {
    const int reuse_order = (reuse_flags & (RRL_REUSE_ORDER | RRL_REUSE_TARGET)) ? 1 : 0;
    const int sorted = g_dense_variant ? 1 : 0;
    const int nfp_max = g.nf1p > g.nf2p ? g.nf1p : g.nf2p;
    const int G = node_size(g);
    if (nfp_max <= kSortSmall && g_param[0] == 0) {
        const int n2 = nfp_max <= 1024 ? 1024 : (nfp_max <= 2048 ? 2048 : 4096);
        int line_blocks = (g.nl + 1023) / 1024;
        if (line_blocks > 64) line_blocks = 64;
        const dim3 grid(g.B, 2 + line_blocks);
        const int handoff_ok = g_param[15] == 0 && g.B <= 4 && reuse_order;    // (see small_prep_kernel; [15] = 1 switches the hand-off off)
        static unsigned long long attr_mask8 = 0ull, attr_mask16 = 0ull;
        if (ensure_dyn_smem(small_prep_kernel<8>, 65536, attr_mask8) || ensure_dyn_smem(small_prep_kernel<16>, 65536, attr_mask16))
            return RRL_ERR_CUDA;
        if (G == 8) small_prep_kernel<8><<<grid, 1024, (size_t)n2 * 16, s>>>(tri1, tri2, lines, ws, g, window, sorted, line_blocks, g_param[8], g_param[10], reuse_order, use_compressed(g), handoff_ok);
        else small_prep_kernel<16><<<grid, 1024, (size_t)n2 * 16, s>>>(tri1, tri2, lines, ws, g, window, sorted, line_blocks, g_param[8], g_param[10], reuse_order, use_compressed(g), handoff_ok);
        count_launch();
        stage_mark(1, s);
        stage_mark(2, s);
        stage_mark(3, s);
        return check_launch();
    }
    // per-pair block {pmax ... gcounts} is contiguous, see carve()
    const size_t pair_bytes = (size_t)((char *)(ws.gcounts + (size_t)g.B * 18) - (char *)ws.xcursor);
    if (cudaMemsetAsync(ws.xcursor, 0, pair_bytes, s) != cudaSuccess) return RRL_ERR_CUDA;
    const int most = g.nl > g.nf1 ? (g.nl > g.nf2 ? g.nl : g.nf2) : (g.nf1 > g.nf2 ? g.nf1 : g.nf2);
    int bx = (most + 255) / 256;
    const int cap = (sm_count() * 8 + g.B - 1) / g.B;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    const int reuse_target = (reuse_flags & RRL_REUSE_TARGET) ? 1 : 0;
    prep_kernel<<<dim3(bx, g.B), 256, 0, s>>>(tri1, tri2, lines, ws, g, window, reuse_target);
    count_launch();
    stage_mark(1, s);
    if (reuse_order) {
        // RRL_REUSE_ORDER: perm[] of the previous forward of this geometry stays as it is (see small_prep_kernel)
    } else if (nfp_max <= kSortSmall) {
        int n2 = 1;
        while (n2 < nfp_max) n2 <<= 1;
        sort_small_kernel<<<dim3(g.B, 2), 1024, (size_t)n2 * 8, s>>>(tri1, tri2, ws, g, sorted);
        count_launch();
    } else {
        // every cloud of (up to kSortPairs) pairs in ONE sort: the passes are latency bound at these sizes, so a sort of
        // many segments costs little more than a sort of one.  When a single sort covers the batch, perm[0] and perm[1]
        // are adjacent and the sort writes them directly.
        const int nbmax = g.B < kSortPairs ? g.B : kSortPairs;
        const size_t cap = (size_t)nbmax * ((size_t)g.nf1p + g.nf2p);
        unsigned *keys_in = reinterpret_cast<unsigned *>(ws.sortbuf);
        unsigned *keys_out = keys_in + cap;
        int *vals_in = reinterpret_cast<int *>(keys_out + cap);
        int *vals_tmp = vals_in + cap;
        char *temp = reinterpret_cast<char *>(vals_tmp + cap);
        temp = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(temp) + 255) / 256 * 256);
        if ((size_t)(temp - reinterpret_cast<char *>(ws.sortbuf)) > ws.sortbuf_bytes) return RRL_ERR_WORKSPACE;
        const size_t temp_avail = ws.sortbuf_bytes - (size_t)(temp - reinterpret_cast<char *>(ws.sortbuf));
        for (int b0 = 0; b0 < g.B; b0 += nbmax) {
            const int nb = g.B - b0 < nbmax ? g.B - b0 : nbmax;
            int segbits = 1;
            while ((1 << segbits) < 2 * nb) ++segbits;
            const size_t ntot = (size_t)nb * ((size_t)g.nf1p + g.nf2p);
            if (ntot >= (1ull << 31)) return RRL_ERR_ARG;
            int *perm0 = ws.perm[0] + (long long)b0 * g.nf1p, *perm1 = ws.perm[1] + (long long)b0 * g.nf2p;
            const bool adjacent = perm1 == perm0 + (long long)nb * g.nf1p;
            int *vals_out = adjacent ? perm0 : vals_tmp;
            sort_keys_kernel<<<(unsigned)((ntot + 255) / 256), 256, 0, s>>>(tri1 + (long long)b0 * g.nf1 * 9, tri2 + (long long)b0 * g.nf2 * 9, nb, g.nf1,
                                                                          g.nf1p, g.nf2, g.nf2p, ws.pmax + b0 * 2, segbits, keys_in, vals_in, sorted);
            count_launch();
            // The order only decides which triplets share a node, so the sort skips Morton bits finer than a node: the curve
            // keeps log2(nfp) + 3 bits (an eighth of the mean spacing per axis), at most 8 are dropped -- one 8-bit pass fewer.
            // Measured on 500k / 65k triplets: sort 110 -> 91 / 88 -> 73 us with the dense stage unchanged (12 dropped bits: +8 %).
            const int hbits = 31 - segbits < 30 ? 31 - segbits : 30;
            int keep_bits = 3;
            while (keep_bits < 33 && (1ll << (keep_bits - 3)) < nfp_max) ++keep_bits;
            int begin_bit = hbits - keep_bits;
            begin_bit = begin_bit < 0 ? 0 : (begin_bit > 8 ? 8 : begin_bit);
            if (g_param[12]) begin_bit = g_param[12] < 0 ? 0 : g_param[12];         // A/B override (-1 = every bit)
            size_t need = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, need, keys_in, keys_out, vals_in, vals_out, (int)ntot, begin_bit, 32, s);
            if (need > temp_avail) return RRL_ERR_WORKSPACE;
            if (cub::DeviceRadixSort::SortPairs(temp, need, keys_in, keys_out, vals_in, vals_out, (int)ntot, begin_bit, 32, s) != cudaSuccess)
                return RRL_ERR_CUDA;
            count_launch((32 - begin_bit + 7) / 8 + 2);  // CUB: histogram + exclusive sum + one onesweep kernel per 8-bit pass (profiles/r02_launches_large.csv)
            if (!adjacent) {
                if (cudaMemcpyAsync(perm0, vals_tmp, (size_t)nb * g.nf1p * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
                    cudaMemcpyAsync(perm1, vals_tmp + (size_t)nb * g.nf1p, (size_t)nb * g.nf2p * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
                    return RRL_ERR_CUDA;
            }
        }
    }
    if (sorted && g_param[10] == 2 && !reuse_order) {   // measured on the large path: costs more (33 us) than it saves; A/B only
        int rbx = nfp_max / 64 / 8;
        if (rbx > 2048) rbx = 2048;
        if (rbx < 1) rbx = 1;
        if (G == 8) refine_kernel<8><<<dim3(rbx, g.B * 2), 256, 0, s>>>(tri1, tri2, ws, g);
        else refine_kernel<16><<<dim3(rbx, g.B * 2), 256, 0, s>>>(tri1, tri2, ws, g);
        count_launch();
    }
    stage_mark(2, s);
    int nbx = nfp_max / 256;
    if (nbx > 4096) nbx = 4096;
    if (G == 8) node_kernel<8><<<dim3(nbx, g.B * 2), 256, 0, s>>>(tri1, tri2, ws, g, g_param[14], use_supers(g), reuse_target, use_compressed(g));
    else node_kernel<16><<<dim3(nbx, g.B * 2), 256, 0, s>>>(tri1, tri2, ws, g, g_param[14], use_supers(g), reuse_target, use_compressed(g));
    count_launch();
    stage_mark(3, s);
    return check_launch();
}

// ------------------------------------------------------------------------------------------------------
// dense
// ------------------------------------------------------------------------------------------------------
#ifdef RRL_COUNTERS
// measurement builds: how many entries every queue level consumed (rrl_debug_read_counters)
__device__ unsigned long long g_counters[8];
#define RRL_COUNT(i, n) do { if (lane == 0) atomicAdd(&g_counters[i], (unsigned long long)(n)); } while (0)
#else
#define RRL_COUNT(i, n) do { } while (0)
#endif
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
    int chunk_nodes;          // nodes per blockIdx.y chunk (multiple of kNodePad)
};

// Exact test of one (line, triplet) -- literal restatement of loss.py:84-110 -- and hit recording.
__device__ __forceinline__ void exact_test_and_record(const float *__restrict__ tri, const float *__restrict__ thr_arr,
                                                      const float *ln, int f, int *cnt_line, int *hits_line,
                                                      int &band, int &nan) {
    const float *t = tri + (long long)f * 9;
    const float thr = __ldg(thr_arr + f);
    const float ulp = ulp_up(thr);
    const float d0 = __fsqrt_rn(point_line_x_exact(__ldg(t + 0), __ldg(t + 1), __ldg(t + 2), ln));
    nan += (d0 != d0);
    if (!(d0 < thr) && !(fabsf(d0 - thr) <= ulp)) return;          // point 0 fails outside the band: nothing left to decide
    const float d1 = __fsqrt_rn(point_line_x_exact(__ldg(t + 3), __ldg(t + 4), __ldg(t + 5), ln));
    const float d2 = __fsqrt_rn(point_line_x_exact(__ldg(t + 6), __ldg(t + 7), __ldg(t + 8), ln));
    nan += (d1 != d1) + (d2 != d2);
    band += decisive_band(d0, d1, d2, thr, ulp);
    if (!(d0 < thr)) return;
    if ((d1 < thr) & (d2 < thr)) {
        const int slot = atomicAdd(cnt_line, 1);
        if (slot < kCap) hits_line[slot] = f;
    }
}

constexpr int kNumWarps = kDenseThreads / 32;

// Shared-memory plan of one dense CTA.  LPT = lines per thread: 4 lines amortise every broadcast node read over more
// arithmetic (2 CTAs per SM, 128 registers), 2 lines halve the register and shared-memory footprint so that 4 CTAs
// (32 warps) fit an SM and cover the latency of the queue levels.
template <int kNode, bool kPerNode, int LPT, bool kSuper = false>
struct DenseCfg {
    static_assert(!kSuper || !kPerNode, "super nodes feed the (line, group) queue");
    static constexpr int kLines = kDenseThreads * LPT;                 // lines per CTA
    static constexpr int kTile = LPT >= 4 ? 256 : 128;                 // nodes per TMA stage
    static constexpr int kStage = (kTile / 4) * 5;                     // float4 per stage: 5 per group of 4 nodes
    static constexpr int kWq = LPT >= 4 || (kSuper && LPT > 1) ? 512 : 256;         // per warp: (line, group) entries, or (line, node) when kPerNode
    static constexpr int kNq = 512;                                    // per warp: (line, node) entries of the 3-level pipeline (a level-1 trip appends <= 256)
    static constexpr int kXq = 32 * kNode + 64;                        // per warp: (line, triplet); one level-2 pass appends <= 32 * kNode
    // kPerNode (small clouds): the chunk's point records are staged in shared memory -- the launch sizes the chunks to
    // fit -- and read with LDS; otherwise they are read through L1/L2 with LDG.  Compile-time either way: a pointer that
    // may be shared or global turns every access into a generic load (4 % of the kernel on the DCP batch).
    static constexpr int kPts = !kPerNode ? 0 : (LPT >= 4 ? 576 : 288);    // point-0 records (float4, incl. pads)
    static constexpr int kPts12 = !kPerNode ? 0 : (LPT >= 4 ? 1024 : 512); // point-1/2 records of the chunk
    static constexpr int kOffLine = 2 * kStage * 16;
    static constexpr int kOffWq = kOffLine + kLines * 32;
    static constexpr int kOffNq = kOffWq + kNumWarps * kWq * 4;
    static constexpr int kOffXq = kOffNq + (kPerNode ? 0 : kNumWarps * kNq * 4);
    static constexpr int kOffPts = kOffXq + kNumWarps * kXq * 4;
    static constexpr int kOffPts12 = kOffPts + kPts * 16;
    static constexpr int kSmem = kOffPts12 + kPts12 * 16;
    static constexpr int kFitBlocks = (227 * 1024) / (kSmem + 2048);
    // one line per thread (super-node mode only): the main loop is a small part there, the queue levels want warps
    static constexpr int kMinBlocks = (kSuper && RRL_SUPER_MINBLOCKS > 0) ? (RRL_SUPER_MINBLOCKS < kFitBlocks ? RRL_SUPER_MINBLOCKS : kFitBlocks) : LPT == 1 ? (kFitBlocks >= 5 ? 5 : kFitBlocks) : (kFitBlocks >= 4 && LPT < 4 ? 4 : (kFitBlocks >= 3 && LPT < 4 ? 3 : 2));
    static_assert(kMinBlocks * (kSmem + 2048) <= 227 * 1024, "CTAs per SM must fit (dynamic + static + 1 KB reserved each)");
};

// exclusive prefix sum over the lanes of a (converged) warp of a count c < 2^kBits, by bit planes: kBits independent
// ballots instead of a dependent chain of five shuffles (the queue levels are latency bound, not issue bound)
#ifndef RRL_SCAN_SHFL
#define RRL_SCAN_SHFL 0
#endif

// q[0..] = key + (position of every set bit of `mask`, ascending).  The mask is bit-reversed once, so that an entry costs
// one FLO (bfind: the highest set bit of the reversed mask = the lowest of the mask), the subtraction from key + 31, the
// store, the pointer step and two instructions to clear the bit -- the __ffs / x & (x - 1) form spent a BREV and an address
// computation more per entry, and these loops are 15 % of the instructions of the small-cloud kernel.
__device__ __forceinline__ void push_bits(unsigned *q, unsigned mask, unsigned key) {
    unsigned r = __brev(mask);
    const unsigned key31 = key + 31u;
    --q;
    while (r) {
        unsigned hb;
        asm("bfind.u32 %0, %1;" : "=r"(hb) : "r"(r));
        *++q = key31 - hb;
        r &= ~(1u << hb);
    }
}

template <int kBits>
__device__ __forceinline__ int warp_excl_scan(int c, int lane, int &total) {
#if RRL_SCAN_SHFL
    int inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    total = __shfl_sync(0xffffffffu, inc, 31);
    return inc - c;
#endif
    const unsigned lt = (1u << lane) - 1u;
    int off = 0;
    total = 0;
    // only the bit planes some lane uses (counts are mostly 0..3: two or three ballots instead of kBits)
    const unsigned used = __reduce_or_sync(0xffffffffu, (unsigned)c);
#pragma unroll
    for (int bit = 0; bit < kBits; ++bit) {
        if ((used >> bit) == 0u) break;
        const unsigned bal = __ballot_sync(0xffffffffu, (c >> bit) & 1);
        off += __popc(bal & lt) << bit;
        total += __popc(bal) << bit;
    }
    return off;
}

// Three-level candidate pipeline of one warp (all queues in shared memory, all pushes ordered by warp scans, so
// there are no atomics, no overflow paths, and every level runs with converged, fully populated warps):
//   main loop : (line, node) sphere predicate, packed FFMA2, branch-free bit masks  -> wq: (line, group of 4 nodes)
//   level 1   : one wq entry per lane, the 4 node predicates again                  -> nq: (line, node)
//   level 2   : one nq entry per lane, the triplet predicate on the node's triplets -> xq: (line, triplet)
//   level 3   : one xq entry per lane, the EXACT reference-order test, hit recording
// kPerNode (small, hit-dense clouds): the main loop keeps one mask per node of a group and feeds (line, node) entries
// straight to level 2 -- level 1 would otherwise re-evaluate nearly every group (half of all (line, group) pairs
// fire on a 1024-triplet cloud).
// kSuper (large clouds): the main loop streams SUPER-node records (bounding spheres of 16 nodes = 256 triplets) instead
// of node records and expands every fired (line, super node) pair into its four (line, group) entries; the node
// predicates then run in level 1 only where a line comes near, instead of for all nodes x lines.
template <int kNode, bool kPerNode, int LPT, bool kSuper = false>
__global__ void __launch_bounds__(kDenseThreads, DenseCfg<kNode, kPerNode, LPT, kSuper>::kMinBlocks) dense_kernel(DenseArgs a, Workspace ws, Geometry g) {
    using Cfg = DenseCfg<kNode, kPerNode, LPT, kSuper>;
    constexpr int kLinesPerThread = LPT, kLinesPerCta = Cfg::kLines, kTileNodes = Cfg::kTile, kStageF4 = Cfg::kStage;
    constexpr int kWarpQueue = Cfg::kWq, kNodeQueue = Cfg::kNq, kExactQueue = Cfg::kXq;
    constexpr int kPtStride = kNode + 1;                          // float4 per node of the triplet records (see make_node_coop)
    constexpr int kGrpStride = 5;                                 // float4 per group of 4 node records
    extern __shared__ __align__(128) unsigned char dsm[];
    float4 *stage = reinterpret_cast<float4 *>(dsm);                                       // [2][kStageF4]
    float4 *spts = reinterpret_cast<float4 *>(dsm + Cfg::kOffPts);                         // [kSmemPtsF4]
    float4 *spts12 = reinterpret_cast<float4 *>(dsm + Cfg::kOffPts12);                     // [kSmemPts12F4]
    float4 *slineU = reinterpret_cast<float4 *>(dsm + Cfg::kOffLine);                      // [kLinesPerCta] {u, k_line = 2 sqrt(e)}
    float4 *slineM = slineU + kLinesPerCta;                                                // [kLinesPerCta] {M, tl_point}
    __shared__ __align__(8) unsigned long long mbar[4];
    __shared__ int tile_done[2];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int b = blockIdx.z >> 1, cloud = blockIdx.z & 1;
    const int nf = cloud ? g.nf2 : g.nf1, nfp = cloud ? g.nf2p : g.nf1p;
    const int nnodes = nfp / kNode;
    // records streamed by the main loop: nodes, or super nodes (kSuperNodes nodes each); chunk bounds are in records
    constexpr int kSuperNodes = kSuper ? kSuperPts / kNode : 1;
    const int nrecs = kSuper ? pad_supers_dev(nfp) : nnodes;
    const int n_begin = blockIdx.y * a.chunk_nodes;
    if (n_begin >= nrecs) return;
    const int n_end = min(nrecs, n_begin + a.chunk_nodes);
    const int node_begin = n_begin * kSuperNodes;                  // first NODE of the chunk
    const int line_base = blockIdx.x * kLinesPerCta;
    unsigned *wq = reinterpret_cast<unsigned *>(dsm + Cfg::kOffWq) + wid * kWarpQueue;
    // the (line, node) queue: in kPerNode mode the main loop fills it directly and it takes over the larger region
    constexpr int kNodeCap = kPerNode ? kWarpQueue : kNodeQueue;
    unsigned *nq = kPerNode ? wq : reinterpret_cast<unsigned *>(dsm + Cfg::kOffNq) + wid * kNodeQueue;
    unsigned *xq = reinterpret_cast<unsigned *>(dsm + Cfg::kOffXq) + wid * kExactQueue;

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_init(&mbar[2], 1);
        mbar_init(&mbar[3], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tile_done[0] = 0; tile_done[1] = 0;
    }

    const float4 *src = (kSuper ? ws.super4[cloud] + (long long)b * (nrecs / 4) * 5 : ws.node4[cloud] + (long long)b * (nnodes / 4) * 5) +
                        (long long)(n_begin / 4) * 5;                                                          // chunk start
    const int ntiles = (n_end - n_begin + kTileNodes - 1) / kTileNodes;
    auto issue = [&](int t) {
        const int n = min(kTileNodes, n_end - (n_begin + t * kTileNodes));
        const unsigned bytes = (unsigned)(n / 4) * 5u * 16u;
        mbar_expect_tx(&mbar[t & 1], bytes);
        tma_bulk_load(stage + (t & 1) * kStageF4, src + (long long)t * kStageF4, bytes, &mbar[t & 1]);
    };
    // the chunk's triplet records (level 2) go to shared memory when they fit, else they are read through L2
    const float4 *pt4_c = ws.pt4[cloud] + ((long long)b * nnodes + node_begin) * kPtStride;   // chunk start
    constexpr bool pts_in_smem = kPerNode, pts12_in_smem = kPerNode;
    const float4 *pts = spts;
    if constexpr (!kPerNode) pts = pt4_c;
    // compressed records (make_node_coop) of the modes that gather them from L2
    constexpr bool kCompressed = !kPerNode && RRL_PC8;
    // super-node mode: level 1 reads ONE compressed record of a fired super node's 16 node spheres (write_super_nodes)
    constexpr bool kSn8 = kSuper && kNode == 16 && RRL_SN8;
    [[maybe_unused]] const uint4 *sn_src = ws.sn8[cloud] + ((long long)b * pad_supers_dev(nfp) + (kSuper ? n_begin : 0)) * 9;
    constexpr int kPcStride = 1 + kNode / 2;                      // uint4 per node
    const uint4 *pcs = reinterpret_cast<const uint4 *>(ws.pt4[cloud]) + ((long long)b * nnodes + node_begin) * kPcStride;
    // point-1/2 records of the chunk (refine pass before the hand-off to the exact kernel)
    const float4 *pt12_c = ws.pt12[cloud] + ((long long)b * nfp + (long long)node_begin * kNode) * 2;
    const float4 *pts12 = spts12;
    if constexpr (!kPerNode) pts12 = pt12_c;
    // tid 0 initialised the barriers itself, so it issues the copies BEFORE the line set-up below: the TMA latency runs
    // under the loads and threshold arithmetic of the lines instead of after them
    if (tid == 0) {
        issue(0);
        if (ntiles > 1) issue(1);
        if (pts_in_smem) {
            const unsigned bytes = (unsigned)(n_end - n_begin) * (kNode + 1) * 16u;
            mbar_expect_tx(&mbar[2], bytes);
            tma_bulk_load(spts, pt4_c, bytes, &mbar[2]);
        }
        if (pts12_in_smem) {
            const unsigned bytes = (unsigned)(n_end - n_begin) * kNode * 2u * 16u;
            mbar_expect_tx(&mbar[3], bytes);
            tma_bulk_load(spts12, pt12_c, bytes, &mbar[3]);
        }
    }
    // ---- per-thread lines -> filter thresholds ------------------------------------------------------------
    // every global load of the set-up is issued before the first use (the cloud's extents AND the thread's line records): a
    // CTA lives for ~13 us on the DCP batch, and four dependent L2 round trips at its start were 6 % of the kernel's stall
    // samples.  Padding lines read record 0 (always there) and are masked below.
    const float4 *lineC = ws.lineC + (long long)b * g.nl * 2;
    float4 lc0[kLinesPerThread], lc1[kLinesPerThread];
#pragma unroll
    for (int i = 0; i < kLinesPerThread; ++i) {
        const int l = line_base + tid + i * kDenseThreads;
        const long long lr = l < g.nl ? l : 0;
        lc0[i] = __ldg(lineC + lr * 2);
        lc1[i] = __ldg(lineC + lr * 2 + 1);
    }
    const unsigned pmax_bits = ws.pmax[b * 2 + cloud], rmax_bits = ws.rmax[b * 2 + cloud], tmax_bits = ws.tmax[b * 2 + cloud];
    const unsigned smax_bits = kSuper ? ws.smax[b * 2 + cloud] : 0u;
    const float P = sqrtf(__uint_as_float(pmax_bits)) * 1.000001f;
    const float Rmax = __uint_as_float(rmax_bits);
    const float Tmax = __uint_as_float(tmax_bits);
    const float Smax = __uint_as_float(smax_bits);
    const float2 neg1 = make_float2(-1.0f, -1.0f);
    float ux[kLinesPerThread], uy[kLinesPerThread], uz[kLinesPerThread];
    float mx[kLinesPerThread], my[kLinesPerThread], mz[kLinesPerThread], tl[kLinesPerThread];
    // threshold of the triplet-level predicate and of the node-level predicate for a line
    // k_line = 2 sqrt(e): the level-1 predicate forms a node's slack from the node's OWN radius, tl_point - (k R_n + k^2 / 2)
    auto thresholds = [&](const float4 &c0, const float4 &c1, float &tl_point, float &tl_node, float &tl_super, float &k_line) {
        const float PX = P + c0.w;
        // compressed records: w~ is formed in the kernel with three roundings of magnitude <= P^2 + T instead of one
        constexpr float kPXc = kCompressed ? kFastPX + 2.0f : kFastPX, kTc = kCompressed ? kFastT + 2.0f : kFastT;
        const float guard = kMargin * kEps24 * (kPXc * PX * PX + kTc * Tmax * Tmax) * 1.000001f + 1e-12f;
        tl_point = c1.w - guard - fabsf(c1.w) * 1.2e-7f;                    // rounded down: admits more
        // |u| > 1 makes F slightly indefinite; e = eta (P + |x0|)^2 bounds the deficit, 0 for |u| <= 1 (incl. all-zero
        // lines).  eta = |u|^2 - 1 is formed in double from the float components the exact test uses: a float sum needs a safety
        // factor that makes EVERY normalised line pay the slack (|u|^2 = 1 +- 1e-7), and on large clouds -- sqrt(e) ~ 5e-3 against
        // hit cylinders of radius 4e-3 -- that slack, times the cloud's LARGEST node radius, tripled the (line, node) candidates.
        const double s2d = (double)c0.x * (double)c0.x + (double)c0.y * (double)c0.y + (double)c0.z * (double)c0.z;
        const float eta = s2d > 1.0 ? (float)(s2d - 1.0) * 1.000001f + 1e-30f : 0.f;
        const float e = eta * PX * PX * 1.000001f;
        if (kCompressed && e > 0.f) {
            // F(p~) - F(p) for |p~ - p| <= eb: 2 eb sqrt(F + e) + (1 + eta) eb^2; cut~ covers 2 eb sqrt(cut + E) + eb^2
            const float eb = pc_err_bound(Rmax, P);
            tl_point -= (2.0f * eb * sqrtf(e) + eta * eb * eb) * 1.00001f;
        }
        k_line = 2.0f * sqrtf(e) * 1.00001f;
        const float slack = (2.0f * Rmax * sqrtf(e) + 2.0f * e) * 1.00001f;
        tl_node = tl_point - slack - fabsf(tl_point) * 1.2e-7f;
        const float sslack = (2.0f * Smax * sqrtf(e) + 2.0f * e) * 1.00001f;
        tl_super = tl_point - sslack - fabsf(tl_point) * 1.2e-7f;
    };
#pragma unroll
    for (int i = 0; i < kLinesPerThread; ++i) {
        const int l = line_base + tid + i * kDenseThreads;
        ux[i] = uy[i] = uz[i] = mx[i] = my[i] = mz[i] = 0.f;
        tl[i] = INFINITY;                                         // Q > +inf never holds: padding lines are inert
        if (l < g.nl) {
            const float4 c0 = lc0[i], c1 = lc1[i];
            ux[i] = c0.x; uy[i] = c0.y; uz[i] = c0.z;
            mx[i] = c1.x; my[i] = c1.y; mz[i] = c1.z;
            float tp, tn, ts, kl;
            thresholds(c0, c1, tp, tn, ts, kl);
            tl[i] = kSuper ? ts : tn;                              // threshold of the records the main loop streams
            // only ever re-read by this warp's queue levels, which need the thresholds rather than |x0| and c
            slineU[tid + i * kDenseThreads] = make_float4(c0.x, c0.y, c0.z, kl);
            slineM[tid + i * kDenseThreads] = make_float4(c1.x, c1.y, c1.z, tp);
        }
    }
    __syncthreads();

    // the bulk copies of the chunk's point records are awaited where a warp first needs them (level 2 / the refine pass),
    // not here: the main loop only reads the node records of the stage buffers
    bool pts_ready = !pts_in_smem, pts12_ready = !pts12_in_smem;         // warp-uniform

    int ncand = 0;
    const int *perm_c = ws.perm[cloud] + (long long)b * nfp + (long long)node_begin * kNode;   // from the chunk start
    // node records for level 1: re-read through L1/L2 (a pointer that is sometimes the resident stage would make
    // every access a generic load)
    const float4 *node_src = ws.node4[cloud] + (long long)b * (nnodes / 4) * kGrpStride + (long long)(node_begin / 4) * kGrpStride;   // from the chunk start
    int wq_cnt = 0, nq_cnt = 0, xq_cnt = 0;                      // warp-uniform fill levels

    // level 3 hand-off: (line, triplet) entries that passed both filters go to the launch-wide queue of the exact
    // kernel (one reservation per flush).  When the queue is full the warp runs the exact test itself.
    auto run_exact = [&]() {
        __syncwarp();
        if (!pts12_ready) { mbar_wait(&mbar[3], 0); pts12_ready = true; }
        RRL_COUNT(2, xq_cnt);                                  // (line, triplet) entries that passed the point-0 predicate
        // refine: the same conservative FMA predicate on points 1 and 2 (one entry per lane, compacted in place):
        // about one in eight entries that passed on point 0 survives
        {
            int kept = 0;
            for (int base = 0; base < xq_cnt; base += 32) {
                bool keep = false;
                unsigned xent = 0;
                if (base + lane < xq_cnt) {
                    xent = xq[base + lane];
                    const int lrel = (int)(xent >> 22), pos = (int)(xent & 0x3FFFFFu);
                    const float4 c0 = slineU[lrel], c1 = slineM[lrel];
                    const float4 A = pts12[pos * 2], Bq = pts12[pos * 2 + 1];
                    const float t1 = fmaf(A.z, c0.z, fmaf(A.y, c0.y, A.x * c0.x));
                    const float s1 = fmaf(A.z, c1.z, fmaf(A.y, c1.y, fmaf(A.x, c1.x, A.w)));
                    const float t2 = fmaf(Bq.z, c0.z, fmaf(Bq.y, c0.y, Bq.x * c0.x));
                    const float s2 = fmaf(Bq.z, c1.z, fmaf(Bq.y, c1.y, fmaf(Bq.x, c1.x, Bq.w)));
                    keep = (fmaf(t1, t1, s1) > c1.w) & (fmaf(t2, t2, s2) > c1.w);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                __syncwarp();                                  // every lane has read its entry before any slot is overwritten
                if (keep) xq[kept + __popc(bal & ((1u << lane) - 1u))] = xent;
                kept += __popc(bal);
                __syncwarp();
            }
            xq_cnt = kept;
        }
        RRL_COUNT(3, xq_cnt);                                  // ... and the refine on points 1 and 2: handed to the exact kernel
        if (xq_cnt > 0) {
            // the first 32 entries' triplet indices are fetched BEFORE the reservation: two round trips to L2 in flight
            // together instead of one after the other (most flushes hold fewer than 32 entries)
            const unsigned xent0 = lane < xq_cnt ? xq[lane] : 0u;
            const int f0 = lane < xq_cnt ? __ldg(perm_c + (int)(xent0 & 0x3FFFFFu)) : 0;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(ws.xcursor, (unsigned long long)xq_cnt);
            base = __shfl_sync(0xffffffffu, base, 0);
            const bool fits = (long long)(base + (unsigned long long)xq_cnt) <= ws.xcap;
            for (int i = lane; i < xq_cnt; i += 32) {
                const unsigned xent = i == lane ? xent0 : xq[i];
                const int l = line_base + (int)(xent >> 22);
                const int f = i == lane ? f0 : __ldg(perm_c + (int)(xent & 0x3FFFFFu));
                const long long gl = (long long)b * g.nl + l;
                if (fits) {
                    ws.xcand[base + i] = make_uint2((unsigned)gl, (unsigned)f | ((unsigned)cloud << 31));
                } else {
                    // the reserved slots that still lie inside the queue become sentinels
                    if ((long long)(base + i) < ws.xcap) ws.xcand[base + i] = make_uint2(0xFFFFFFFFu, 0u);
                    float ln[6];
#pragma unroll
                    for (int c = 0; c < 6; ++c) ln[c] = __ldg(a.lines + gl * 6 + c);
                    int band = 0, nan = 0;
                    exact_test_and_record(a.tri[cloud] + (long long)b * nf * 9, ws.thr[cloud] + (long long)b * nf, ln, f,
                                          ws.cnt[cloud] + gl, ws.hits[cloud] + gl * kCap, band, nan);
                    long long *st = ws.stats + (long long)b * RRL_NSTAT;
                    if (band) atomicAdd((unsigned long long *)(st + 5), (unsigned long long)band);
                    if (nan) atomicAdd((unsigned long long *)(st + 6), (unsigned long long)nan);
                }
            }
        }
        __syncwarp();
        xq_cnt = 0;
    };
    // level 2: triplet predicate on the triplets of (line, node) entries, packed two triplets per FFMA2
    // Code size matters here: the levels are lambdas inlined at every call site, and with one site per overflow check the
    // kernel held six copies of level 2 and seven of level 3 (140 KB of SASS; "no instruction" was the top stall reason of the
    // large-cloud kernel).  Every level therefore calls the next one from ONE site -- the overflow check and the final flush
    // (flush = true: drain the levels below as well) share it -- and the main loop pushes from one site per mode.
    auto run_nodes = [&](bool flush) {
        __syncwarp();
        if (!pts_ready) { mbar_wait(&mbar[2], 0); pts_ready = true; }
        RRL_COUNT(1, nq_cnt);                                  // (line, node) entries = triplet-record fetches of (kNode + 1) float4
        for (int base = 0;; base += 32) {
            const bool more = base < nq_cnt;
            if (more ? (xq_cnt + 32 * kNode > kExactQueue) : flush) run_exact();
            if (!more) break;
            unsigned pm = 0, key = 0;
            if (base + lane < nq_cnt) {
                const unsigned ent = nq[base + lane];
                const int lrel = (int)(ent >> 22), nrel = (int)(ent & 0x3FFFFFu);
                const float4 c0 = slineU[lrel], c1 = slineM[lrel];
                const float2 tlp2 = make_float2(c1.w, c1.w);                      // tl_point
                const float2 u0 = make_float2(c0.x, c0.x), u1 = make_float2(c0.y, c0.y), u2 = make_float2(c0.z, c0.z);
                const float2 m0 = make_float2(c1.x, c1.x), m1 = make_float2(c1.y, c1.y), m2 = make_float2(c1.z, c1.z);
                if constexpr (kCompressed) {
                    // every lane another node, straight from L2: 16 + 8 kNode bytes per node.  All loads are issued before the
                    // first use (one round trip per pass); the offsets become floats with one PRMT each (0x4B000000 | u)
                    const uint4 *pp = pcs + nrel * kPcStride;
                    const uint4 hd = __ldg(pp);
                    uint4 rec[kNode / 2];
#pragma unroll
                    for (int j = 0; j < kNode / 2; ++j) rec[j] = __ldg(pp + 1 + j);
                    const float sc = __uint_as_float(hd.w);
                    const float2 sc2 = make_float2(sc, sc);
                    const float2 bx = make_float2(__uint_as_float(hd.x), __uint_as_float(hd.x)), by = make_float2(__uint_as_float(hd.y), __uint_as_float(hd.y));
                    const float2 bz = make_float2(__uint_as_float(hd.z), __uint_as_float(hd.z));
#pragma unroll
                    for (int j = 0; j < kNode / 2; ++j) {
                        const uint4 r = rec[j];
                        const float2 xu = make_float2(__uint_as_float(__byte_perm(r.x, 0x4B000000u, 0x7410)), __uint_as_float(__byte_perm(r.x, 0x4B000000u, 0x7432)));
                        const float2 yu = make_float2(__uint_as_float(__byte_perm(r.y, 0x4B000000u, 0x7410)), __uint_as_float(__byte_perm(r.y, 0x4B000000u, 0x7432)));
                        const float2 zu = make_float2(__uint_as_float(__byte_perm(r.z, 0x4B000000u, 0x7410)), __uint_as_float(__byte_perm(r.z, 0x4B000000u, 0x7432)));
                        const float2 x2 = __ffma2_rn(xu, sc2, bx), y2 = __ffma2_rn(yu, sc2, by), z2 = __ffma2_rn(zu, sc2, bz);
                        const float2 cut2 = __half22float2(*reinterpret_cast<const __half2 *>(&r.w));
                        // w~ = cut~ - |p~|^2 (three more roundings than the stored w of the uncompressed record: kFastPXc / kFastTc)
                        const float2 nx2 = make_float2(-x2.x, -x2.y), ny2 = make_float2(-y2.x, -y2.y), nz2 = make_float2(-z2.x, -z2.y);
                        const float2 w2 = __ffma2_rn(nz2, z2, __ffma2_rn(ny2, y2, __ffma2_rn(nx2, x2, cut2)));
                        const float2 t2 = __ffma2_rn(z2, u2, __ffma2_rn(y2, u1, __fmul2_rn(x2, u0)));
                        const float2 s2 = __ffma2_rn(z2, m2, __ffma2_rn(y2, m1, __ffma2_rn(x2, m0, w2)));
                        const float2 q2 = __ffma2_rn(t2, t2, s2);
                        const float2 d2 = __ffma2_rn(q2, neg1, tlp2);             // sign bits into the mask (see the main loop)
                        pm = __funnelshift_l(__float_as_uint(d2.x), pm, 1);
                        pm = __funnelshift_l(__float_as_uint(d2.y), pm, 1);
                    }
                } else {
                    const float4 *pp = pts + nrel * kPtStride;
#pragma unroll
                    for (int j = 0; j < kNode / 2; ++j) {
                        const float4 A = pp[2 * j], Bq = pp[2 * j + 1];
                        const float2 x2 = make_float2(A.x, A.y), y2 = make_float2(A.z, A.w), z2 = make_float2(Bq.x, Bq.y), w2 = make_float2(Bq.z, Bq.w);
                        const float2 t2 = __ffma2_rn(z2, u2, __ffma2_rn(y2, u1, __fmul2_rn(x2, u0)));
                        const float2 s2 = __ffma2_rn(z2, m2, __ffma2_rn(y2, m1, __ffma2_rn(x2, m0, w2)));
                        const float2 q2 = __ffma2_rn(t2, t2, s2);
                        const float2 d2 = __ffma2_rn(q2, neg1, tlp2);
                        pm = __funnelshift_l(__float_as_uint(d2.x), pm, 1);
                        pm = __funnelshift_l(__float_as_uint(d2.y), pm, 1);
                    }
                }
                pm = __brev(pm) >> (32 - kNode);                                  // bit = triplet index inside the node
                key = ((unsigned)lrel << 22) | (unsigned)(nrel * kNode);
            }
            int total;
            const int pos = xq_cnt + warp_excl_scan<(kNode == 8 ? 4 : 5)>(__popc(pm), lane, total);
            push_bits(xq + pos, pm, key);
            xq_cnt += total;
            __syncwarp();
        }
        nq_cnt = 0;
    };
    // level 1: node predicate on the 4 nodes of (line, group) entries
    auto run_groups = [&](bool flush) {
        __syncwarp();
        RRL_COUNT(0, wq_cnt);                                  // (line, group of 4 nodes) entries = node-record fetches of 5 float4
        if constexpr (kSn8) {
            // one (line, super node) entry per lane: the 16 node predicates from the super node's compressed record (144 bytes, all
            // loads in flight together); the node threshold comes from each node's own radius, tl_point - (k R~ + k^2 / 2)
            for (int base = 0;; base += 32) {
                const bool more = base < wq_cnt;
                unsigned nm = 0, key = 0;
                if (more && base + lane < wq_cnt) {
                    const unsigned ent = wq[base + lane];
                    const int lrel = (int)(ent >> 20), srel = (int)(ent & 0xFFFFFu);      // super node, relative to the chunk
                    const float4 c0 = slineU[lrel], c1 = slineM[lrel];
                    const uint4 *pp = sn_src + srel * 9;
                    const uint4 hd = __ldg(pp);
                    uint4 rec[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) rec[j] = __ldg(pp + 1 + j);
                    const float sc = __uint_as_float(hd.w);
                    const float2 sc2 = make_float2(sc, sc);
                    const float2 bx = make_float2(__uint_as_float(hd.x), __uint_as_float(hd.x)), by = make_float2(__uint_as_float(hd.y), __uint_as_float(hd.y));
                    const float2 bz = make_float2(__uint_as_float(hd.z), __uint_as_float(hd.z));
                    const float2 u0 = make_float2(c0.x, c0.x), u1 = make_float2(c0.y, c0.y), u2 = make_float2(c0.z, c0.z);
                    const float2 m0 = make_float2(c1.x, c1.x), m1 = make_float2(c1.y, c1.y), m2 = make_float2(c1.z, c1.z);
                    const float tl_base = fmaf(-0.5f * c0.w, c0.w * 1.00001f, c1.w) - fabsf(c1.w) * 2.4e-7f;
                    const float2 tlb2 = make_float2(tl_base, tl_base), kn2 = make_float2(-c0.w, -c0.w);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint4 r = rec[j];
                        const float2 xu = make_float2(__uint_as_float(__byte_perm(r.x, 0x4B000000u, 0x7410)), __uint_as_float(__byte_perm(r.x, 0x4B000000u, 0x7432)));
                        const float2 yu = make_float2(__uint_as_float(__byte_perm(r.y, 0x4B000000u, 0x7410)), __uint_as_float(__byte_perm(r.y, 0x4B000000u, 0x7432)));
                        const float2 zu = make_float2(__uint_as_float(__byte_perm(r.z, 0x4B000000u, 0x7410)), __uint_as_float(__byte_perm(r.z, 0x4B000000u, 0x7432)));
                        const float2 x2 = __ffma2_rn(xu, sc2, bx), y2 = __ffma2_rn(yu, sc2, by), z2 = __ffma2_rn(zu, sc2, bz);
                        const float2 r2 = __half22float2(*reinterpret_cast<const __half2 *>(&r.w));        // radii (NaN: empty node)
                        const float2 nx2 = make_float2(-x2.x, -x2.y), ny2 = make_float2(-y2.x, -y2.y), nz2 = make_float2(-z2.x, -z2.y);
                        const float2 w2 = __ffma2_rn(nz2, z2, __ffma2_rn(ny2, y2, __ffma2_rn(nx2, x2, __fmul2_rn(r2, r2))));
                        const float2 t2 = __ffma2_rn(z2, u2, __ffma2_rn(y2, u1, __fmul2_rn(x2, u0)));
                        const float2 s2 = __ffma2_rn(z2, m2, __ffma2_rn(y2, m1, __ffma2_rn(x2, m0, w2)));
                        const float2 q2 = __ffma2_rn(t2, t2, s2);
                        const float2 tn2 = __ffma2_rn(r2, kn2, tlb2);
                        const float2 d2 = __ffma2_rn(q2, neg1, tn2);                  // sign bits into the mask (see the main loop)
                        nm = __funnelshift_l(__float_as_uint(d2.x), nm, 1);
                        nm = __funnelshift_l(__float_as_uint(d2.y), nm, 1);
                    }
                    nm = __brev(nm) >> 16;                                            // bit = node index inside the super node
                    key = ((unsigned)lrel << 22) | (unsigned)(srel * 16);
                }
                int total = 0, off = 0;
                if (more) off = warp_excl_scan<5>(__popc(nm), lane, total);
                if (more ? (nq_cnt + total > kNodeCap) : flush) run_nodes(flush && !more);
                if (!more) break;
                push_bits(nq + nq_cnt + off, nm, key);
                nq_cnt += total;
                __syncwarp();
            }
            wq_cnt = 0;
            return;
        }
        // node predicates of one entry: the records are pair-interleaved ({xA,xB,yA,yB} {zA,zB,wA,wB}, then the 4 radii): two
        // nodes per packed FMA, as in the main loop; the threshold comes from the node's own radius, tl_point - (k R_n + k^2 / 2),
        // k = 2 sqrt(e) (0 for |u| <= 1), rounded down
        auto node_bits = [&](unsigned ent, const float4 &A0, const float4 &A1, const float4 &B0, const float4 &B1, const float4 &R4) -> unsigned {
            const int lrel = (int)(ent >> 20);
            const float4 c0 = slineU[lrel], c1 = slineM[lrel];
            const float tl_base = fmaf(-0.5f * c0.w, c0.w * 1.00001f, c1.w) - fabsf(c1.w) * 2.4e-7f;
            const float tn0 = fmaf(-c0.w, R4.x, tl_base), tn1 = fmaf(-c0.w, R4.y, tl_base);
            const float tn2 = fmaf(-c0.w, R4.z, tl_base), tn3 = fmaf(-c0.w, R4.w, tl_base);
            const float2 u0 = make_float2(c0.x, c0.x), u1 = make_float2(c0.y, c0.y), u2 = make_float2(c0.z, c0.z);
            const float2 m0 = make_float2(c1.x, c1.x), m1 = make_float2(c1.y, c1.y), m2 = make_float2(c1.z, c1.z);
            const float2 xa = make_float2(A0.x, A0.y), ya = make_float2(A0.z, A0.w), za = make_float2(A1.x, A1.y), wa = make_float2(A1.z, A1.w);
            const float2 xb = make_float2(B0.x, B0.y), yb = make_float2(B0.z, B0.w), zb = make_float2(B1.x, B1.y), wb = make_float2(B1.z, B1.w);
            const float2 ta = __ffma2_rn(za, u2, __ffma2_rn(ya, u1, __fmul2_rn(xa, u0)));
            const float2 sa = __ffma2_rn(za, m2, __ffma2_rn(ya, m1, __ffma2_rn(xa, m0, wa)));
            const float2 qa = __ffma2_rn(ta, ta, sa);
            const float2 tb = __ffma2_rn(zb, u2, __ffma2_rn(yb, u1, __fmul2_rn(xb, u0)));
            const float2 sb = __ffma2_rn(zb, m2, __ffma2_rn(yb, m1, __ffma2_rn(xb, m0, wb)));
            const float2 qb = __ffma2_rn(tb, tb, sb);
            return (qa.x > tn0 ? 1u : 0u) | (qa.y > tn1 ? 2u : 0u) | (qb.x > tn2 ? 4u : 0u) | (qb.y > tn3 ? 8u : 0u);
        };
        // TWO entries per lane and trip: a trip is one round trip to L2 for the (scattered) group records and little arithmetic, so
        // both entries' loads are issued before either is used
        for (int base = 0;; base += 64) {
            const bool more = base < wq_cnt;
            if (more ? (nq_cnt + 256 > kNodeCap) : flush) run_nodes(flush && !more);
            if (!more) break;
            const bool hasA = base + lane < wq_cnt, hasB = base + 32 + lane < wq_cnt;
            const unsigned entA = hasA ? wq[base + lane] : 0u, entB = hasB ? wq[base + 32 + lane] : 0u;     // (entry 0 = group 0: a valid address)
            const float4 *pa = node_src + (int)(entA & 0xFFFFFu) * kGrpStride, *pb = node_src + (int)(entB & 0xFFFFFu) * kGrpStride;
            const float4 a0 = pa[0], a1 = pa[1], a2 = pa[2], a3 = pa[3], a4 = pa[4];
            const float4 b0 = pb[0], b1 = pb[1], b2 = pb[2], b3 = pb[3], b4 = pb[4];
            const unsigned nmA = hasA ? node_bits(entA, a0, a1, a2, a3, a4) : 0u;
            const unsigned nmB = hasB ? node_bits(entB, b0, b1, b2, b3, b4) : 0u;
            const int cA = __popc(nmA);
            int total;
            const int pos = nq_cnt + warp_excl_scan<4>(cA + __popc(nmB), lane, total);
            push_bits(nq + pos, nmA, ((entA >> 20) << 22) | ((entA & 0xFFFFFu) * 4u));
            push_bits(nq + pos + cA, nmB, ((entB >> 20) << 22) | ((entB & 0xFFFFFu) * 4u));
            nq_cnt += total;
            __syncwarp();
        }
        wq_cnt = 0;
    };

    for (int t = 0; t < ntiles; ++t) {
        mbar_wait(&mbar[t & 1], (t >> 1) & 1);
        const int nn = min(kTileNodes, n_end - (n_begin + t * kTileNodes));      // multiple of kNodePad
        const float4 *sp = stage + (t & 1) * kStageF4;
        const int ngroups = nn / 4;
        const int group0 = t * (kTileNodes / 4);
        if constexpr (kPerNode || kSuper) {
            // window = 8 groups = 32 records: ONE mask word per line, bit = record index inside the window
            for (int w0 = 0; w0 < ngroups; w0 += 8) {
                const int ng = min(8, ngroups - w0);
                unsigned m[kLinesPerThread];
#pragma unroll
                for (int i = 0; i < kLinesPerThread; ++i) m[i] = 0u;
                // record counts are multiples of kNodePad = 16 = 4 groups: a window is one or two halves of 4 groups,
                // fully unrolled so that every mask bit is an immediate (a bit position computed from a loop counter
                // cost a MOV + SHF per test)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h * 4 >= ng) break;
#pragma unroll
                    for (int gq = 0; gq < 4; ++gq) {
                        const int gi = h * 4 + gq;
                        // 4 records = two interleaved pairs: {xA,xB,yA,yB} {zA,zB,wA,wB}
                        const float4 a0 = sp[(w0 + gi) * 5 + 0], a1 = sp[(w0 + gi) * 5 + 1], b0 = sp[(w0 + gi) * 5 + 2], b1 = sp[(w0 + gi) * 5 + 3];
                        const float2 xa = make_float2(a0.x, a0.y), ya = make_float2(a0.z, a0.w), za = make_float2(a1.x, a1.y), wa = make_float2(a1.z, a1.w);
                        const float2 xb = make_float2(b0.x, b0.y), yb = make_float2(b0.z, b0.w), zb = make_float2(b1.x, b1.y), wb = make_float2(b1.z, b1.w);
#pragma unroll
                        for (int i = 0; i < kLinesPerThread; ++i) {
                            const float2 u0 = make_float2(ux[i], ux[i]), u1 = make_float2(uy[i], uy[i]), u2 = make_float2(uz[i], uz[i]);
                            const float2 m0 = make_float2(mx[i], mx[i]), m1 = make_float2(my[i], my[i]), m2 = make_float2(mz[i], mz[i]);
                            const float2 ta = __ffma2_rn(za, u2, __ffma2_rn(ya, u1, __fmul2_rn(xa, u0)));
                            const float2 sa = __ffma2_rn(za, m2, __ffma2_rn(ya, m1, __ffma2_rn(xa, m0, wa)));
                            const float2 qa = __ffma2_rn(ta, ta, sa);
                            const float2 tb = __ffma2_rn(zb, u2, __ffma2_rn(yb, u1, __fmul2_rn(xb, u0)));
                            const float2 sb = __ffma2_rn(zb, m2, __ffma2_rn(yb, m1, __ffma2_rn(xb, m0, wb)));
                            const float2 qb = __ffma2_rn(tb, tb, sb);
                            // Q > tl  <=>  tl - Q < 0 (a rounded difference keeps the sign of the exact one): the SIGN BITS are
                            // shifted into the mask, one funnel shift per test instead of compare + select + or (records in order,
                            // first record ends up in the highest bit; reversed below)
                            const float2 tl2 = make_float2(tl[i], tl[i]);
                            const float2 da = __ffma2_rn(qa, neg1, tl2), db = __ffma2_rn(qb, neg1, tl2);
                            m[i] = __funnelshift_l(__float_as_uint(da.x), m[i], 1);
                            m[i] = __funnelshift_l(__float_as_uint(da.y), m[i], 1);
                            m[i] = __funnelshift_l(__float_as_uint(db.x), m[i], 1);
                            m[i] = __funnelshift_l(__float_as_uint(db.y), m[i], 1);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < kLinesPerThread; ++i) m[i] = __brev(m[i]) >> (32 - 4 * ng);       // bit = record index inside the window
                // ordered push of the fired (line, node) pairs: one scan and one bit loop per line
                const unsigned node0 = (unsigned)((group0 + w0) * 4);
                if constexpr (kSuper) {
                    // every fired (line, super node) pair -> the (line, group) entries of its kSuperNodes / 4 node groups
                    constexpr int kGroupsPer = kSn8 ? 1 : kSuperNodes / 4;     // (one entry per fired super node with the compressed node records)
                    constexpr int kPart = kWarpQueue / (32 * kGroupsPer);      // bits per part: 32 lanes x kPart x kGroupsPer entries fit
                    static_assert(kPart >= 1 && (kPart & (kPart - 1)) == 0, "part size");
#pragma unroll 1
                    for (int i = 0; i < kLinesPerThread; ++i) {                       // rolled: ONE push site (see run_nodes)
                        const unsigned lrel = (unsigned)(tid + i * kDenseThreads);
                        unsigned mi = m[0];
#pragma unroll
                        for (int k = 1; k < kLinesPerThread; ++k) mi = i == k ? m[k] : mi;
                        if (!__any_sync(0xffffffffu, mi != 0u)) continue;
                        ncand += __popc(mi);
                        // fired bits are sparse (a line comes near a handful of super nodes): one scan over the whole
                        // word; parts of kPart bits only when a window fills more than a queue
                        int total;
                        const int off = warp_excl_scan<6>(__popc(mi), lane, total);
                        const bool whole = total * kGroupsPer <= kWarpQueue;
                        const int step = whole ? 32 : kPart;
#pragma unroll 1
                        for (int part = 0; part < 32; part += step) {
                            unsigned mh = mi;
                            int tot2 = total, off2 = off;
                            if (!whole) {
                                mh = (mi >> part) & ((1u << kPart) - 1u);
                                if (!__any_sync(0xffffffffu, mh != 0u)) continue;
                                off2 = warp_excl_scan<(kPart >= 16 ? 5 : (kPart >= 8 ? 4 : 3))>(__popc(mh), lane, tot2);      // counts <= kPart
                            }
                            if (wq_cnt + tot2 * kGroupsPer > kWarpQueue) run_groups(false);
                            int pos = wq_cnt + off2 * kGroupsPer;
                            const unsigned rec0 = node0 + (unsigned)(whole ? 0 : part);
                            while (mh) {
                                const unsigned rec = rec0 + (unsigned)(__ffs(mh) - 1);              // super node, chunk relative
                                mh &= mh - 1;
#pragma unroll
                                for (int q = 0; q < kGroupsPer; ++q) wq[pos++] = (lrel << 20) | (rec * kGroupsPer + (unsigned)q);
                            }
                            wq_cnt += tot2 * kGroupsPer;
                        }
                    }
                } else {
#pragma unroll 1
                for (int i = 0; i < kLinesPerThread; ++i) {                           // rolled: ONE push site (see run_nodes)
                    const unsigned lrel = (unsigned)(tid + i * kDenseThreads);
                    unsigned mi = m[0];
#pragma unroll
                    for (int k = 1; k < kLinesPerThread; ++k) mi = i == k ? m[k] : mi;
                    const int c = __popc(mi);                                      // <= 32
                    if (!__any_sync(0xffffffffu, c != 0)) continue;
                    ncand += c;
                    int total;
                    const int off = warp_excl_scan<6>(c, lane, total);
                    constexpr int kPart = kNodeCap / 32;           // bits per part: 32 lanes x kPart entries always fit
                    const bool whole = total <= kNodeCap;          // (rarely not: more than a queue-full from one window)
                    const int step = whole ? 32 : kPart;
#pragma unroll 1
                    for (int part = 0; part < 32; part += step) {
                        unsigned mh = mi;
                        int tot2 = total, off2 = off;
                        if (!whole) {
                            mh = (mi >> part) & ((1u << kPart) - 1u);
                            off2 = warp_excl_scan<5>(__popc(mh), lane, tot2);
                        }
                        if (nq_cnt + tot2 > kNodeCap) run_nodes(false);
                        push_bits(nq + nq_cnt + off2, mh, ((lrel << 22) | node0) + (unsigned)(whole ? 0 : part));     // node0 is a multiple of 32
                        nq_cnt += tot2;
                    }
                }
                }   // !kSuper
            }
        } else {
            // window = 32 groups of 4 nodes: one mask word per line, bit = group (the OR of its four node predicates)
            for (int w0 = 0; w0 < ngroups; w0 += 32) {
                const int ng = min(32, ngroups - w0);
                unsigned m[kLinesPerThread];
#pragma unroll
                for (int i = 0; i < kLinesPerThread; ++i) m[i] = 0u;
#pragma unroll 2
                for (int gi = 0; gi < ng; ++gi) {
                    const float4 a0 = sp[(w0 + gi) * 5 + 0], a1 = sp[(w0 + gi) * 5 + 1], b0 = sp[(w0 + gi) * 5 + 2], b1 = sp[(w0 + gi) * 5 + 3];
                    const float2 xa = make_float2(a0.x, a0.y), ya = make_float2(a0.z, a0.w), za = make_float2(a1.x, a1.y), wa = make_float2(a1.z, a1.w);
                    const float2 xb = make_float2(b0.x, b0.y), yb = make_float2(b0.z, b0.w), zb = make_float2(b1.x, b1.y), wb = make_float2(b1.z, b1.w);
                    const unsigned bit = 1u << gi;
#pragma unroll
                    for (int i = 0; i < kLinesPerThread; ++i) {
                        const float2 u0 = make_float2(ux[i], ux[i]), u1 = make_float2(uy[i], uy[i]), u2 = make_float2(uz[i], uz[i]);
                        const float2 m0 = make_float2(mx[i], mx[i]), m1 = make_float2(my[i], my[i]), m2 = make_float2(mz[i], mz[i]);
                        const float2 ta = __ffma2_rn(za, u2, __ffma2_rn(ya, u1, __fmul2_rn(xa, u0)));
                        const float2 sa = __ffma2_rn(za, m2, __ffma2_rn(ya, m1, __ffma2_rn(xa, m0, wa)));
                        const float2 qa = __ffma2_rn(ta, ta, sa);
                        const float2 tb = __ffma2_rn(zb, u2, __ffma2_rn(yb, u1, __fmul2_rn(xb, u0)));
                        const float2 sb = __ffma2_rn(zb, m2, __ffma2_rn(yb, m1, __ffma2_rn(xb, m0, wb)));
                        const float2 qb = __ffma2_rn(tb, tb, sb);
                        const float qmax = fmaxf(fmaxf(qa.x, qa.y), fmaxf(qb.x, qb.y));
                        m[i] |= (qmax > tl[i]) ? bit : 0u;
                    }
                }
                // ordered push of the fired (line, group) pairs, in parts that always fit the queue
#pragma unroll 1
                for (int i = 0; i < kLinesPerThread; ++i) {                           // rolled: ONE push site (see run_nodes)
                    const unsigned lrel = (unsigned)(tid + i * kDenseThreads);
                    unsigned mi = m[0];
#pragma unroll
                    for (int k = 1; k < kLinesPerThread; ++k) mi = i == k ? m[k] : mi;
                    if (!__any_sync(0xffffffffu, mi != 0u)) continue;
                    ncand += __popc(mi);
                    constexpr int kPart = kWarpQueue / 32;         // 16 or 8 groups per part: 32 lanes x kPart entries <= kWarpQueue
#pragma unroll 1
                    for (int part = 0; part < 32; part += kPart) {
                        unsigned mh = (mi >> part) & ((1u << kPart) - 1u);
                        if (!__any_sync(0xffffffffu, mh != 0u)) continue;
                        int total;
                        const int off = warp_excl_scan<5>(__popc(mh), lane, total);
                        if (wq_cnt + total > kWarpQueue) run_groups(false);
                        push_bits(wq + wq_cnt + off, mh, (lrel << 20) + (unsigned)(group0 + w0 + part));
                        wq_cnt += total;
                    }
                }
            }
        }
        // stage t&1 is free once every warp is past it: the last warp to arrive refills it with tile t+2 -- no CTA-wide
        // barrier, so warps with long candidate queues do not hold up the others
        if (t + 2 < ntiles) {
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                if (atomicAdd(&tile_done[t & 1], 1) == kNumWarps - 1) {
                    tile_done[t & 1] = 0;
                    issue(t + 2);
                }
            }
        }
    }
    if constexpr (!kPerNode) run_groups(true);
    else run_nodes(true);

    // ---- diagnostics ---------------------------------------------------------------------------------
    ncand = __reduce_add_sync(0xffffffffu, ncand);
    if (lane == 0 && ncand) atomicAdd((unsigned long long *)(ws.stats + (long long)b * RRL_NSTAT + 3 + cloud), (unsigned long long)ncand);
}

// Level 3: the EXACT reference-order test of every queued (line, triplet) pair.  The chain entry -> line/triplet ->
// counter -> slot is four dependent memory round trips, so every thread keeps kExactIlp entries in flight.
template <int kExactIlp>
__global__ void __launch_bounds__(256) exact_kernel(DenseArgs a, Workspace ws, Geometry g) {
    const unsigned long long reserved = *ws.xcursor;
    const long long n = reserved < (unsigned long long)ws.xcap ? (long long)reserved : ws.xcap;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * kExactIlp) {
        uint2 e[kExactIlp];
#pragma unroll
        for (int u = 0; u < kExactIlp; ++u) {
            const long long i = i0 + u * stride;
            e[u] = i < n ? ws.xcand[i] : make_uint2(0xFFFFFFFFu, 0u);
        }
        float ln[kExactIlp][6], t[kExactIlp][9], thr[kExactIlp];
#pragma unroll
        for (int u = 0; u < kExactIlp; ++u) {
            const bool ok = e[u].x != 0xFFFFFFFFu;
            const long long gl = ok ? e[u].x : 0;
            const int cloud = (int)(e[u].y >> 31), f = (int)(e[u].y & 0x7FFFFFFFu);
            const int b = (int)(gl / g.nl);
            const int nf = cloud ? g.nf2 : g.nf1;
            const float *tp = a.tri[cloud] + ((long long)b * nf + f) * 9;
#pragma unroll
            for (int c = 0; c < 6; ++c) ln[u][c] = __ldg(a.lines + gl * 6 + c);
#pragma unroll
            for (int c = 0; c < 9; ++c) t[u][c] = __ldg(tp + c);
            thr[u] = __ldg(ws.thr[cloud] + (long long)b * nf + f);
        }
#pragma unroll
        for (int u = 0; u < kExactIlp; ++u) {
            if (e[u].x == 0xFFFFFFFFu) continue;
            const long long gl = e[u].x;
            const int cloud = (int)(e[u].y >> 31), f = (int)(e[u].y & 0x7FFFFFFFu);
            const float ulp = ulp_up(thr[u]);
            const float d0 = __fsqrt_rn(point_line_x_exact(t[u][0], t[u][1], t[u][2], ln[u]));
            const float d1 = __fsqrt_rn(point_line_x_exact(t[u][3], t[u][4], t[u][5], ln[u]));
            const float d2 = __fsqrt_rn(point_line_x_exact(t[u][6], t[u][7], t[u][8], ln[u]));
            // NaN accounting as in exact_test_and_record (points 1 and 2 only count when point 0 passes or sits in the band)
            const bool p0 = d0 < thr[u];
            const bool look = p0 || fabsf(d0 - thr[u]) <= ulp;
            const int nan = (d0 != d0) + (look ? (d1 != d1) + (d2 != d2) : 0);
            const int band = decisive_band(d0, d1, d2, thr[u], ulp);
            if (p0 & (d1 < thr[u]) & (d2 < thr[u])) {
                const int slot = atomicAdd(ws.cnt[cloud] + gl, 1);
                if (slot < kCap) ws.hits[cloud][gl * kCap + slot] = f;
            }
            if (band | nan) {
                long long *st = ws.stats + (gl / g.nl) * RRL_NSTAT;
                if (band) atomicAdd((unsigned long long *)(st + 5), (unsigned long long)band);
                if (nan) atomicAdd((unsigned long long *)(st + 6), (unsigned long long)nan);
            }
        }
    }
}

// The reference's own formulation, every (line, triplet) tested exactly: one thread per (line, cloud).  Not on the
// product path: rrl_debug_set_param(5, 1) selects it as an on-device cross-check of the filtered pipeline.
__global__ void __launch_bounds__(128) bruteforce_kernel(DenseArgs a, Workspace ws, Geometry g, int force) {
    const int b = blockIdx.y >> 1, cloud = blockIdx.y & 1;
    (void)force;
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= g.nl) return;
    const int nf = cloud ? g.nf2 : g.nf1;
    const long long gl = (long long)b * g.nl + l;
    const float *tri = a.tri[cloud] + (long long)b * nf * 9;
    const float *thr_arr = ws.thr[cloud] + (long long)b * nf;
    float ln[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) ln[c] = __ldg(a.lines + gl * 6 + c);
    int count = 0, band = 0, nan = 0;
    int *hits_line = ws.hits[cloud] + gl * kCap;
    for (int f = 0; f < nf; ++f) {                  // same early exit (and band/NaN accounting) as exact_test_and_record
        const float *t = tri + (long long)f * 9;
        const float thr = __ldg(thr_arr + f);
        const float ulp = ulp_up(thr);
        const float d0 = __fsqrt_rn(point_line_x_exact(__ldg(t + 0), __ldg(t + 1), __ldg(t + 2), ln));
        nan += (d0 != d0);
        if (!(d0 < thr) && !(fabsf(d0 - thr) <= ulp)) continue;
        const float d1 = __fsqrt_rn(point_line_x_exact(__ldg(t + 3), __ldg(t + 4), __ldg(t + 5), ln));
        const float d2 = __fsqrt_rn(point_line_x_exact(__ldg(t + 6), __ldg(t + 7), __ldg(t + 8), ln));
        nan += (d1 != d1) + (d2 != d2);
        band += decisive_band(d0, d1, d2, thr, ulp);
        if (!(d0 < thr)) continue;
        if ((d1 < thr) & (d2 < thr)) {
            if (count < kCap) hits_line[count] = f;
            ++count;
        }
    }
    ws.cnt[cloud][gl] = count;
    long long *st = ws.stats + (long long)b * RRL_NSTAT;
    if (band) atomicAdd((unsigned long long *)(st + 5), (unsigned long long)band);
    if (nan) atomicAdd((unsigned long long *)(st + 6), (unsigned long long)nan);
}

int launch_bruteforce(const float *tri1, const float *tri2, const float *lines, const Workspace &ws, const Geometry &g,
                      int force, cudaStream_t s) {
    DenseArgs a;
    a.tri[0] = tri1; a.tri[1] = tri2; a.lines = lines; a.chunk_nodes = 0;
    bruteforce_kernel<<<dim3((g.nl + 127) / 128, g.B * 2), 128, 0, s>>>(a, ws, g, force);
    count_launch();
    return check_launch();
}

template <int kNode, bool kPerNode, int LPT, bool kSuper = false>
static int launch_dense_variant(const DenseArgs &a0, const Workspace &ws, const Geometry &g, int G, cudaStream_t s) {
    using Cfg = DenseCfg<kNode, kPerNode, LPT, kSuper>;
    static unsigned long long attr_mask = 0ull;
    if (ensure_dyn_smem(dense_kernel<kNode, kPerNode, LPT, kSuper>, Cfg::kSmem, attr_mask)) return RRL_ERR_CUDA;
    DenseArgs a = a0;
    const int line_tiles = (g.nl + Cfg::kLines - 1) / Cfg::kLines;
    // records the main loop streams: nodes, or super nodes
    const int nfp_max = g.nf1p > g.nf2p ? g.nf1p : g.nf2p;
    const int nn_max = kSuper ? pad_supers(nfp_max) : nfp_max / G;
    const int rec_pad = kNodePad;
    // split the records so that the grid covers the SMs (kMinBlocks CTAs each) g_param[2] times over when the line
    // tiles alone do not; never below g_param[3] records per CTA
    const long long base_ctas = (long long)line_tiles * g.B * 2;
    // super-node mode: every CTA pays a fixed latency chain (set-up loads, the bulk copies, a final flush through three queue
    // levels), so few lines want few, fat CTAs -- measured on the 500k pair: 12 500 lines 74.7 / 79.6 / 87.8 us at 1 / 3 / 6 waves,
    // 50 000 lines 242 / 204 / 213, 100 000 lines 447 / 366 / 364
    const int super_waves = g_param[9] > 0 ? g_param[9] : (base_ctas < 100 ? 1 : (base_ctas < 300 ? 3 : 6));
    const long long target = (long long)sm_count() * Cfg::kMinBlocks * (kSuper ? super_waves : g_param[2]);
    int chunks = 1;
    if (base_ctas < target) chunks = (int)((target + base_ctas - 1) / base_ctas);
    int chunk_nodes = (nn_max + chunks - 1) / chunks;
    if (chunk_nodes < g_param[3]) chunk_nodes = g_param[3];
    chunk_nodes = ((chunk_nodes + rec_pad - 1) / rec_pad) * rec_pad;
    // small clouds: keep the chunk's point records in shared memory (level 2 reads them once per candidate node)
    constexpr int kFit = kPerNode ? (Cfg::kPts / (kNode + 1)) / kNodePad * kNodePad : 0;
    static_assert(!kPerNode || kFit * kNode * 2 <= Cfg::kPts12, "point-1/2 cache must hold what the point-0 cache holds");
    if (kPerNode && chunk_nodes > kFit) chunk_nodes = kFit;
    const long long chunk_real_nodes = (long long)chunk_nodes * (kSuper ? kSuperPts / kNode : 1);
    if (chunk_real_nodes / 4 >= (1 << 20) || chunk_real_nodes * G >= (1 << 22)) return RRL_ERR_ARG;
    chunks = (nn_max + chunk_nodes - 1) / chunk_nodes;
    a.chunk_nodes = chunk_nodes;
    dim3 grid(line_tiles, chunks, g.B * 2);
    dense_kernel<kNode, kPerNode, LPT, kSuper><<<grid, kDenseThreads, Cfg::kSmem, s>>>(a, ws, g);
    count_launch();
    return RRL_OK;
}

int launch_dense(const float *tri1, const float *tri2, const float *lines, const Workspace &ws, const Geometry &g, cudaStream_t s) {
    DenseArgs a;
    a.tri[0] = tri1; a.tri[1] = tri2; a.lines = lines; a.chunk_nodes = 0;
    if (g_param[5]) return launch_bruteforce(tri1, tri2, lines, ws, g, 1, s);      // measurement / cross-check only
    const int G = node_size(g);
    const int lpt = g_param[6] == 2 || g_param[6] == 4 || g_param[6] == 1 ? g_param[6] : 2;
    int rc;
    if (use_supers(g)) {
        if (G == 8) rc = lpt != 4 ? launch_dense_variant<8, false, 2, true>(a, ws, g, G, s) : launch_dense_variant<8, false, 4, true>(a, ws, g, G, s);
        else rc = lpt == 1 ? launch_dense_variant<16, false, 1, true>(a, ws, g, G, s)
                           : (lpt == 2 ? launch_dense_variant<16, false, 2, true>(a, ws, g, G, s) : launch_dense_variant<16, false, 4, true>(a, ws, g, G, s));
    } else if (G == 8 && g_param[4] == 0) rc = lpt != 4 ? launch_dense_variant<8, true, 2>(a, ws, g, G, s) : launch_dense_variant<8, true, 4>(a, ws, g, G, s);
    else if (G == 8) rc = lpt != 4 ? launch_dense_variant<8, false, 2>(a, ws, g, G, s) : launch_dense_variant<8, false, 4>(a, ws, g, G, s);
    else rc = lpt != 4 ? launch_dense_variant<16, false, 2>(a, ws, g, G, s) : launch_dense_variant<16, false, 4>(a, ws, g, G, s);
    if (rc) return rc;
    const int xgrid = sm_count() * (g_param[13] > 0 ? g_param[13] : 8);
    if (g_param[11] == 4) exact_kernel<4><<<xgrid, 256, 0, s>>>(a, ws, g);
    else if (g_param[11] == 1) exact_kernel<1><<<xgrid, 256, 0, s>>>(a, ws, g);
    else exact_kernel<2><<<xgrid, 256, 0, s>>>(a, ws, g);
    count_launch();
    stage_mark(4, s);
    return check_launch();
}

}  // namespace rrl

#ifdef RRL_MARKS
extern "C" int rrl_debug_read_marks_prep(unsigned long long *out32) {
    return cudaMemcpyFromSymbol(out32, rrl::g_marks_prep, sizeof(unsigned long long) * 32) == cudaSuccess ? 0 : -3;
}
#endif
#ifdef RRL_COUNTERS
extern "C" int rrl_debug_read_counters(unsigned long long *out8, int reset) {
    if (cudaMemcpyFromSymbol(out8, rrl::g_counters, sizeof(unsigned long long) * 8) != cudaSuccess) return -3;
    if (reset) {
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(rrl::g_counters, z, sizeof(z));
    }
    return 0;
}
#endif
