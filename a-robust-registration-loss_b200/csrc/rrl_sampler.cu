// On-device random line sampler with bounding-box rejection (sm_100a).  Compiled with -fmad=false: the
// reference's area test is a floating-point knife edge whose acceptance statistics depend on
// every product and sum being rounded separately, like ATen's eager kernels.
//
// Replaces Random_uniform_distribution_lines_batch_efficient_resample (/root/reference/code/loss.py:415-432):
//   bbox_kernel    generate_bbox (loss.py:325-351): per-pair min/max of both clouds;
//   flags_kernel   `rounds` x N candidate chords per pair (loss.py:384-412) from a counter-based Philox4x32-10
//                  stream (or from supplied uniforms), each tested against the 12 triangles of both boxes with
//                  the reference's area test (generate_mesh_by_bbox loss.py:354-362, step1/step2 loss.py:265-316);
//   scan_kernel + scatter_kernel   generate_lines (loss.py:365-381): the first N accepted candidates in (round, index)
//                  order -- an ordered compaction in three parallel steps (accepted count per chunk of 256 candidates
//                  in flags_kernel, exclusive scan of the chunk counts per pair, scatter of every chunk to its
//                  offset); rows that stay unfilled are all-zero, exactly like the reference.  (A single CTA per pair
//                  walking the candidates in order took 0.9 ms for 32 x 15000 lines and 1.1 ms for 1 x 100000.)
#include "rrl_common.cuh"

namespace rrl {

// ---- Philox4x32-10 (Salmon et al., SC'11), counter = (index, round, pair, offset), key = seed -------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

__device__ __forceinline__ float u01(unsigned x) { return (float)(x >> 8) * 5.9604645e-8f; }   // [0,1), 24 bits like torch.rand

struct BoxTris {
    float A[12][3], Bv[12][3], C[12][3], n[12][3], S[12];
};

__device__ const int kBoxFaces[12][3] = {{2, 0, 6}, {0, 4, 6}, {5, 4, 0}, {5, 0, 1}, {6, 4, 5}, {5, 7, 6},
                                         {3, 0, 2}, {1, 0, 3}, {3, 2, 6}, {6, 7, 3}, {5, 1, 3}, {3, 7, 5}};   // loss.py:357-358

__device__ __forceinline__ void cross3(const float *a, const float *b, float *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ float norm3(const float *a) { return __fsqrt_rn((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]); }

__device__ void make_box(const float *lo, const float *hi, BoxTris *bt, int f) {
    // corner order of generate_bbox (loss.py:329-350)
    const int sel[8][3] = {{1, 1, 1}, {1, 1, 0}, {1, 0, 1}, {1, 0, 0}, {0, 1, 1}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}};
    float v[3][3];
    for (int q = 0; q < 3; ++q)
        for (int a = 0; a < 3; ++a) v[q][a] = sel[kBoxFaces[f][q]][a] ? hi[a] : lo[a];
    float e1[3] = {v[1][0] - v[0][0], v[1][1] - v[0][1], v[1][2] - v[0][2]};
    float e2[3] = {v[2][0] - v[0][0], v[2][1] - v[0][1], v[2][2] - v[0][2]};
    float n[3];
    cross3(e1, e2, n);
    const float S = norm3(n);
    const float den = fmaxf(S, 1e-12f);
    for (int a = 0; a < 3; ++a) {
        bt->A[f][a] = v[0][a]; bt->Bv[f][a] = v[1][a]; bt->C[f][a] = v[2][a];
        bt->n[f][a] = n[a] / den;
    }
    bt->S[f] = S;
}

// does the line "hit" at least one of the box's triangles in `mask` by the reference's area test (loss.py:289-316)?  The
// reference counts the hits per box and keeps a line when hits1 * hits2 > 0 (loss.py:430): only whether each count is
// positive matters, so the loop stops at the first hit.  Every lane walks ITS OWN list of triangles (the set bits of its
// mask, see tri_mask): a warp runs as many iterations as its busiest lane has candidates (2-3 faces = 4-6 triangles), not 12.
__device__ bool box_hit(const BoxTris *bt, const float *ln, unsigned mask) {
#pragma unroll 1
    while (mask) {
        const int f = __ffs(mask) - 1;
        mask &= mask - 1u;
        const float *A = bt->A[f], *Bp = bt->Bv[f], *C = bt->C[f], *n = bt->n[f];
        const float num = (n[0] * (A[0] - ln[3]) + n[1] * (A[1] - ln[4])) + n[2] * (A[2] - ln[5]);
        const float den = ((n[0] * ln[0] + n[1] * ln[1]) + n[2] * ln[2]) + 1e-12f;
        const float t = num / den;
        const float X[3] = {t * ln[0] + ln[3], t * ln[1] + ln[4], t * ln[2] + ln[5]};
        const float ca[3] = {X[0] - A[0], X[1] - A[1], X[2] - A[2]};
        const float cb[3] = {X[0] - Bp[0], X[1] - Bp[1], X[2] - Bp[2]};
        const float cc[3] = {X[0] - C[0], X[1] - C[1], X[2] - C[2]};
        float x1[3], x2[3], x3[3];
        cross3(cb, cc, x1);
        cross3(cc, ca, x2);
        cross3(ca, cb, x3);
        const float a = norm3(x1), b = norm3(x2), c = norm3(x3);
        if ((a > 0.f) && (b > 0.f) && (c > 0.f) && ((a + b) + c <= bt->S[f])) return true;
    }
    return false;
}

// candidate chord from four uniforms (loss.py:394-411)
__device__ __forceinline__ void make_line(float r, const float *center, float a1, float u1, float a2, float u2, float *ln) {
    const float PI32 = 3.14159274101257324f;                 // torch.pi rounded to float32 (loss.py:9)
    const float al1 = (a1 * 2.0f) * PI32, z1 = u1 * 2.0f - 1.0f;
    const float al2 = (a2 * 2.0f) * PI32, z2 = u2 * 2.0f - 1.0f;
    const float s1 = __fsqrt_rn(1.0f - z1 * z1), s2 = __fsqrt_rn(1.0f - z2 * z2);
    float sn1, cs1, sn2, cs2;
    sincosf(al1, &sn1, &cs1);
    sincosf(al2, &sn2, &cs2);
    const float q1[3] = {(r * s1) * cs1, (r * sn1) * s1, r * z1};
    const float q2[3] = {(r * s2) * cs2, (r * sn2) * s2, r * z2};
    const float d[3] = {q2[0] - q1[0], q2[1] - q1[1], q2[2] - q1[2]};
    const float nn = fmaxf(norm3(d), 1e-12f);
    ln[0] = d[0] / nn; ln[1] = d[1] / nn; ln[2] = d[2] / nn;
    ln[3] = q1[0] + center[0]; ln[4] = q1[1] + center[1]; ln[5] = q1[2] + center[2];
}

constexpr int kPhases = 6;                    // the rounds are evaluated in phases [0,1) [1,2) [2,3) [3,4) [4,6) [6,rounds): a phase is one
                                              // launch (~3 us); at 42 % acceptance per round the rows fill up during round 2, at 30 % during
                                              // round 3 -- with [2,4) as one phase both cases evaluated four rounds

struct SamplerArgs {
    const float *radius, *centers, *uniforms;
    float *bbox;              // (B, 2, 6): lo, hi
    unsigned char *flags;     // (B, rounds*N)
    int *chunk;               // (B, nchunks): accepted candidates per chunk of kChunk, then (scan_kernel) their exclusive prefix
    int *acc;                 // (B, kPhases + 1): accepted candidates of each phase of rounds (slot 0 stays 0)
    int B, N, rounds, nchunks;
    unsigned long long seed, offset;
    int shard_rank, shard_world;   // candidate shard: this rank evaluates the chunks ch with ch % shard_world == shard_rank
};

__device__ __forceinline__ void candidate(const SamplerArgs &a, int b, int rd, int i, float *ln) {
    float a1, u1, a2, u2;
    if (a.uniforms) {
        const float *U = a.uniforms + (((long long)b * a.rounds + rd) * 4) * a.N;
        a1 = U[i]; u1 = U[a.N + i]; a2 = U[2LL * a.N + i]; u2 = U[3LL * a.N + i];
    } else {
        const uint4 x = philox4x32_10(make_uint4((unsigned)i, (unsigned)rd, (unsigned)b, (unsigned)a.offset),
                                      make_uint2((unsigned)a.seed, (unsigned)(a.seed >> 32) ^ (unsigned)(a.offset >> 32)));
        a1 = u01(x.x); u1 = u01(x.y); a2 = u01(x.z); u2 = u01(x.w);
    }
    make_line(a.radius[b], a.centers + b * 3, a1, u1, a2, u2, ln);
}

__global__ void __launch_bounds__(256) bbox_kernel(const float *__restrict__ v1, const float *__restrict__ v2, int n1, int n2, float *bbox) {
    const int b = blockIdx.x, cloud = blockIdx.y;
    const float *v = (cloud ? v2 : v1) + (long long)b * (cloud ? n2 : n1) * 3;
    const int n = cloud ? n2 : n1;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        for (int a = 0; a < 3; ++a) {
            const float x = v[3 * i + a];
            lo[a] = fminf(lo[a], x);
            hi[a] = fmaxf(hi[a], x);
        }
    __shared__ float s[8][6];
    for (int a = 0; a < 3; ++a) {
        float l = lo[a], h = hi[a];
        for (int o = 16; o > 0; o >>= 1) {
            l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
            h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
        }
        if ((threadIdx.x & 31) == 0) { s[threadIdx.x >> 5][a] = l; s[threadIdx.x >> 5][3 + a] = h; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float r = s[0][threadIdx.x];
        for (int w = 1; w < 8; ++w) r = threadIdx.x < 3 ? fminf(r, s[w][threadIdx.x]) : fmaxf(r, s[w][threadIdx.x]);
        bbox[(b * 2 + cloud) * 6 + threadIdx.x] = r;
    }
}

constexpr int kChunk = 256;                   // candidates per chunk = threads per block of flags_kernel / scatter_kernel

// Which triangles of a box can the line hit at all?  Bit f = triangle f of kBoxFaces.  The two triangles of a face lie in
// the face's plane; the reference intersects the line with that plane (point X) and accepts when the three sub-triangle
// areas add up to the triangle's -- possible only for X inside the triangle.  A face whose plane the line crosses at
// least `margin` (1 % of the box's largest extent, plus a relative allowance for the rounding of X) OUTSIDE the face's
// rectangle fails that test for both of its triangles by a margin orders of magnitude above the rounding of the area sums
// (excess ~ margin x edge length against areas ~ edge^2 with 1e-7 relative rounding), so only the faces the line may
// cross -- two, three near an edge -- are tested at all.  Conservative throughout: a face is dropped only when X is
// provably outside.  Lines (nearly) parallel to a pair of planes keep both faces; degenerate (flat) boxes keep all 12
// triangles (see flags_kernel).  face = axis * 2 + (1 = max side); triangles per face from kBoxFaces + generate_bbox's corners.
__device__ __forceinline__ unsigned tri_mask(const float *lo, const float *hi, float margin, const float *ln) {
    const unsigned kFaceTris[6] = {0x030u, 0x0c0u, 0x300u, 0x00cu, 0xc00u, 0x003u};
    unsigned m = 0u;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int i = (k + 1) % 3, j = (k + 2) % 3;
        const float u = ln[k], x = ln[3 + k];
        if (!(fabsf(u) >= 1e-6f)) {                              // parallel (or NaN): no statement about these two faces
            m |= kFaceTris[2 * k] | kFaceTris[2 * k + 1];
            continue;
        }
        const float inv = 1.0f / u;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const float t = ((side ? hi[k] : lo[k]) - x) * inv;
            const float Xi = ln[3 + i] + t * ln[i], Xj = ln[3 + j] + t * ln[j];
            const float si = margin + 1e-4f * (fabsf(Xi) + fabsf(ln[3 + i])), sj = margin + 1e-4f * (fabsf(Xj) + fabsf(ln[3 + j]));
            const bool outside = Xi < lo[i] - si || Xi > hi[i] + si || Xj < lo[j] - sj || Xj > hi[j] + sj;
            if (!outside) m |= kFaceTris[2 * k + side];
        }
    }
    return m;
}

// One phase of rounds = the chunks [ch_begin, ch_end).  A pair whose rows were already filled by the EARLIER phases
// (their kernels have completed: the counts are final, so the decision is deterministic) skips the phase: later
// candidates cannot be among the first N accepted.  Their chunk counts stay 0 (memset) and scatter_kernel never reads
// their flags.  The reference always evaluates all 10 rounds (loss.py:425-431) and throws the surplus away.
__global__ void __launch_bounds__(kChunk) flags_kernel(SamplerArgs a, int phase, int ch_begin, int ch_end) {
    __shared__ BoxTris bt[2];
    __shared__ float s_box[12];
    const int b = blockIdx.y;
    int before = 0;
    for (int j = 0; j <= phase; ++j) before += a.acc[b * (kPhases + 1) + j];
    if (before >= a.N) return;                                               // block-uniform
    if (threadIdx.x < 24) make_box(a.bbox + (b * 2 + threadIdx.x / 12) * 6, a.bbox + (b * 2 + threadIdx.x / 12) * 6 + 3, &bt[threadIdx.x / 12], threadIdx.x % 12);
    if (threadIdx.x >= 32 && threadIdx.x < 44) s_box[threadIdx.x - 32] = a.bbox[b * 12 + threadIdx.x - 32];
    __syncthreads();
    const long long total = (long long)a.rounds * a.N;
    int mine = 0;
    for (int ch = ch_begin + blockIdx.x; ch < ch_end; ch += gridDim.x) {     // block-uniform trip count
        if (a.shard_world > 1 && ch % a.shard_world != a.shard_rank) continue;   // another rank's chunk (count stays 0)
        const long long c = (long long)ch * kChunk + threadIdx.x;
        int f = 0;
        if (c < total) {
            float ln[6];
            candidate(a, b, (int)(c / a.N), (int)(c % a.N), ln);
            // boxes whose smallest extent is at least 5 % of their largest: only the triangles of the faces the line may cross
            unsigned mask[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float *lo = s_box + q * 6, *hi = lo + 3;
                const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
                const float emax = fmaxf(ex, fmaxf(ey, ez)), emin = fminf(ex, fminf(ey, ez));
                mask[q] = (emin >= 0.05f * emax && emax > 0.f) ? tri_mask(lo, hi, 0.01f * emax, ln) : 0xFFFu;
            }
            f = mask[0] && mask[1] && box_hit(&bt[0], ln, mask[0]) && box_hit(&bt[1], ln, mask[1]);      // loss.py:430
            a.flags[(long long)b * total + c] = (unsigned char)f;
        }
        const int cnt = __syncthreads_count(f);
        if (threadIdx.x == 0) a.chunk[(long long)b * a.nchunks + ch] = cnt;
        mine += cnt;
    }
    if (threadIdx.x == 0 && mine) atomicAdd(a.acc + b * (kPhases + 1) + phase + 1, mine);
}

// exclusive prefix sum of a pair's chunk counts (in place) and the number of filled rows
__global__ void __launch_bounds__(1024) scan_kernel(SamplerArgs a, int *out_filled) {
    __shared__ int warp_tot[32];
    __shared__ int s_carry;
    const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int *cnt = a.chunk + (long long)b * a.nchunks;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < a.nchunks; base += 1024) {                     // block-uniform trip count
        const int i = base + threadIdx.x;
        const int v = i < a.nchunks ? cnt[i] : 0;
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += up;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        int before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < 32; ++w) {
            const int t = warp_tot[w];
            before += w < wid ? t : 0;
            all += t;
        }
        const int carry = s_carry;
        if (i < a.nchunks) cnt[i] = carry + before + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + all;
        __syncthreads();
    }
    if (threadIdx.x == 0) out_filled[b] = min(a.N, s_carry);
}

// every chunk writes its accepted candidates to rows offset[chunk] + rank inside the chunk (rows >= N are dropped, as
// generate_lines stops appending); the rows nobody fills are zeroed
__global__ void __launch_bounds__(kChunk) scatter_kernel(SamplerArgs a, float *out_lines, const int *filled_arr) {
    __shared__ int warp_tot[kChunk / 32];
    const int b = blockIdx.y;
    const long long total = (long long)a.rounds * a.N;
    const unsigned char *fl = a.flags + (long long)b * total;
    const int *off = a.chunk + (long long)b * a.nchunks;
    float *out = out_lines + (long long)b * a.N * 6;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int ch = blockIdx.x; ch < a.nchunks; ch += gridDim.x) {             // block-uniform
        const int base = off[ch];
        if (base >= a.N) break;                                              // offsets only grow: nothing left to place
        const long long c = (long long)ch * kChunk + threadIdx.x;
        const int f = c < total ? fl[c] : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_tot[wid] = __popc(bal);
        __syncthreads();
        int before = 0;
#pragma unroll
        for (int w = 0; w < kChunk / 32; ++w) before += w < wid ? warp_tot[w] : 0;
        const int pos = base + before + __popc(bal & ((1u << lane) - 1u));
        if (f && pos < a.N) {
            float ln[6];
            candidate(a, b, (int)(c / a.N), (int)(c % a.N), ln);
#pragma unroll
            for (int q = 0; q < 6; ++q) out[(long long)pos * 6 + q] = ln[q];
        }
        __syncthreads();
    }
    const int filled = filled_arr[b];
    for (long long i = (long long)filled * 6 + (long long)blockIdx.x * kChunk + threadIdx.x; i < (long long)a.N * 6;
         i += (long long)gridDim.x * kChunk)
        out[i] = 0.f;
}

// ---- candidate shard across ranks -------------------------------------------------------------------
// The candidates are a pure function of (seed, offset, pair, round, index) (counter-based Philox), so any rank can evaluate
// any of them: rank r takes the chunks ch = r (mod world).  The ordered compaction of generate_lines (first N accepted in
// (round, index) order) then needs, per chunk, how many candidates the chunks BEFORE it accepted -- on any rank: the ranks
// sum their per-chunk counts (one all-reduce of nchunks ints, the caller's), every rank scans the summed counts itself and
// keeps, of each of ITS chunks, the accepted candidates whose global row is below N.  The loss does not depend on the
// order of the lines, so nobody needs anybody else's lines: a rank evaluates exactly the rows it produced, plus its share
// of the rows that stay unfilled (all-zero lines, which the reference still evaluates, loss.py:423-432).
//   gcnt: summed counts in, exclusive global prefix out;  lcnt: this rank's counts in, local row offsets of its chunks out;
//   kept (reuses the flags' chunk array): rows each own chunk contributes;  out3 = {rows placed, zero rows appended, filled}
__global__ void __launch_bounds__(1024) shard_scan_kernel(SamplerArgs a, int *gcnt, int *lcnt, int *out3) {
    __shared__ int warp_tot[2][32];
    __shared__ int s_carry[2];
    const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int *g = gcnt + (long long)b * a.nchunks, *l = lcnt + (long long)b * a.nchunks, *kept = a.chunk + (long long)b * a.nchunks;
    if (threadIdx.x < 2) s_carry[threadIdx.x] = 0;
    __syncthreads();
    for (int base = 0; base < a.nchunks; base += 1024) {                     // block-uniform trip count
        const int i = base + threadIdx.x;
        const bool in = i < a.nchunks;
        const int gv = in ? g[i] : 0;
        int ginc = gv;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, ginc, d);
            if (lane >= d) ginc += up;
        }
        if (lane == 31) warp_tot[0][wid] = ginc;
        __syncthreads();
        int gbefore = 0, gall = 0;
#pragma unroll
        for (int w = 0; w < 32; ++w) {
            const int t = warp_tot[0][w];
            gbefore += w < wid ? t : 0;
            gall += t;
        }
        const int goff = s_carry[0] + gbefore + ginc - gv;                  // accepted candidates before chunk i, all ranks
        const bool own = in && (a.shard_world <= 1 || i % a.shard_world == a.shard_rank);
        const int lv = own ? l[i] : 0;
        const int kv = own ? max(0, min(lv, a.N - goff)) : 0;               // rows of chunk i that land below N
        int kinc = kv;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, kinc, d);
            if (lane >= d) kinc += up;
        }
        if (lane == 31) warp_tot[1][wid] = kinc;
        __syncthreads();
        int kbefore = 0, kall = 0;
#pragma unroll
        for (int w = 0; w < 32; ++w) {
            const int t = warp_tot[1][w];
            kbefore += w < wid ? t : 0;
            kall += t;
        }
        if (in) {
            g[i] = goff;
            l[i] = s_carry[1] + kbefore + kinc - kv;                        // local row of the chunk's first kept candidate
            kept[i] = kv;
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_carry[0] += gall; s_carry[1] += kall; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int filled = min(a.N, s_carry[0]);
        // unfilled rows [filled, N): row r belongs to rank r % world
        int zeros = 0;
        if (a.N > filled) {
            const int w = a.shard_world > 1 ? a.shard_world : 1, r = a.shard_world > 1 ? a.shard_rank : 0;
            const int first = filled + ((r - filled % w) % w + w) % w;
            zeros = first < a.N ? (a.N - 1 - first) / w + 1 : 0;
        }
        out3[b * 3] = s_carry[1]; out3[b * 3 + 1] = zeros; out3[b * 3 + 2] = filled;
    }
}

__global__ void __launch_bounds__(kChunk) shard_scatter_kernel(SamplerArgs a, const int *__restrict__ lcnt, float *out_lines,
                                                               const int *__restrict__ out3) {
    __shared__ int warp_tot[kChunk / 32];
    const int b = blockIdx.y;
    const long long total = (long long)a.rounds * a.N;
    const unsigned char *fl = a.flags + (long long)b * total;
    const int *loff = lcnt + (long long)b * a.nchunks, *kept = a.chunk + (long long)b * a.nchunks;
    float *out = out_lines + (long long)b * a.N * 6;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int ch = blockIdx.x; ch < a.nchunks; ch += gridDim.x) {             // block-uniform
        if (a.shard_world > 1 && ch % a.shard_world != a.shard_rank) continue;
        const int keep = kept[ch];
        if (keep <= 0) continue;
        const long long c = (long long)ch * kChunk + threadIdx.x;
        const int f = c < total ? fl[c] : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_tot[wid] = __popc(bal);
        __syncthreads();
        int before = 0;
#pragma unroll
        for (int w = 0; w < kChunk / 32; ++w) before += w < wid ? warp_tot[w] : 0;
        const int rank_in_chunk = before + __popc(bal & ((1u << lane) - 1u));
        if (f && rank_in_chunk < keep) {
            float ln[6];
            candidate(a, b, (int)(c / a.N), (int)(c % a.N), ln);
            const long long pos = loff[ch] + rank_in_chunk;
#pragma unroll
            for (int q = 0; q < 6; ++q) out[pos * 6 + q] = ln[q];
        }
        __syncthreads();
    }
    // the rank's share of the unfilled rows follows its lines; whatever lies beyond is zeroed too (the caller slices)
    const long long placed = out3[b * 3];
    for (long long i = placed * 6 + (long long)blockIdx.x * kChunk + threadIdx.x; i < (long long)a.N * 6; i += (long long)gridDim.x * kChunk)
        out[i] = 0.f;
}

}  // namespace rrl

using namespace rrl;

static size_t sampler_chunks(int N, int rounds) { return ((size_t)rounds * (size_t)N + kChunk - 1) / kChunk; }
static size_t up256(size_t x) { return (x + 255) / 256 * 256; }

extern "C" size_t rrl_sampler_workspace_bytes(int B, int N, int rounds) {
    if (B <= 0 || N <= 0 || rounds <= 0) return 0;
    return up256((size_t)B * 12 * sizeof(float)) + up256((size_t)B * (size_t)rounds * (size_t)N) +
           (size_t)B * (sampler_chunks(N, rounds) + kPhases + 1) * sizeof(int);
}

extern "C" int rrl_sample_lines(const float *radius, const float *centers, const float *verts1, const float *verts2,
                                int B, int n1, int n2, int N, int rounds, unsigned long long seed, unsigned long long offset,
                                const float *uniforms, float *out_lines, int *out_filled,
                                void *workspace, size_t workspace_bytes, void *stream) {
    if (!radius || !centers || !verts1 || !verts2 || !out_lines || !out_filled || !workspace) return RRL_ERR_ARG;
    if (B <= 0 || n1 <= 0 || n2 <= 0 || N <= 0 || rounds <= 0) return RRL_ERR_ARG;
    if ((long long)rounds * N >= (1LL << 31) - kChunk) return RRL_ERR_ARG;
    if (workspace_bytes < rrl_sampler_workspace_bytes(B, N, rounds)) return RRL_ERR_WORKSPACE;
    Range range("rrl_sample_lines");
    cudaStream_t s = (cudaStream_t)stream;
    SamplerArgs a;
    a.radius = radius; a.centers = centers; a.uniforms = uniforms;
    char *w = reinterpret_cast<char *>(workspace);
    a.bbox = reinterpret_cast<float *>(w);
    w += up256((size_t)B * 12 * sizeof(float));
    a.flags = reinterpret_cast<unsigned char *>(w);
    w += up256((size_t)B * (size_t)rounds * (size_t)N);
    a.chunk = reinterpret_cast<int *>(w);
    a.B = B; a.N = N; a.rounds = rounds; a.seed = seed; a.offset = offset;
    a.shard_rank = 0; a.shard_world = 1;
    a.nchunks = (int)sampler_chunks(N, rounds);
    a.acc = a.chunk + (size_t)B * a.nchunks;
    if (cudaMemsetAsync(a.chunk, 0, (size_t)B * (a.nchunks + kPhases + 1) * sizeof(int), s) != cudaSuccess) return RRL_ERR_CUDA;
    bbox_kernel<<<dim3(B, 2), 256, 0, s>>>(verts1, verts2, n1, n2, a.bbox);
    const int cap = (sm_count() * 8 + B - 1) / B;
    // a chunk belongs to the phase of rounds that holds its first candidate
    const int raw_edge[kPhases + 1] = {0, 1, 2, 3, 4, 6, rounds};
    int edge[kPhases + 1];
    for (int q = 0; q <= kPhases; ++q) edge[q] = raw_edge[q] < rounds ? raw_edge[q] : rounds;
    int bx_all = 1;
    for (int ph = 0; ph < kPhases; ++ph) {
        const int ch_begin = (int)(((long long)edge[ph] * N + kChunk - 1) / kChunk);
        const int ch_end = (int)(((long long)edge[ph + 1] * N + kChunk - 1) / kChunk);
        if (ch_end <= ch_begin) continue;
        int bx = ch_end - ch_begin;
        if (bx > cap) bx = cap < 1 ? 1 : cap;
        if (bx > bx_all) bx_all = bx;
        flags_kernel<<<dim3(bx, B), kChunk, 0, s>>>(a, ph, ch_begin, ch_end);
        count_launch();
    }
    const int bx = bx_all;
    scan_kernel<<<B, 1024, 0, s>>>(a, out_filled);
    scatter_kernel<<<dim3(bx, B), kChunk, 0, s>>>(a, out_lines, out_filled);
    count_launch(3);
    return check_launch();
}

// ---- candidate-sharded sampler: two calls with the caller's all-reduce of the chunk counts between them ---------------
static SamplerArgs shard_args(const float *radius, const float *centers, const float *uniforms, int B, int N, int rounds,
                              unsigned long long seed, unsigned long long offset, int rank, int world, void *workspace) {
    SamplerArgs a;
    a.radius = radius; a.centers = centers; a.uniforms = uniforms;
    char *w = reinterpret_cast<char *>(workspace);
    a.bbox = reinterpret_cast<float *>(w);
    w += up256((size_t)B * 12 * sizeof(float));
    a.flags = reinterpret_cast<unsigned char *>(w);
    w += up256((size_t)B * (size_t)rounds * (size_t)N);
    a.chunk = reinterpret_cast<int *>(w);
    a.B = B; a.N = N; a.rounds = rounds; a.seed = seed; a.offset = offset;
    a.shard_rank = rank; a.shard_world = world;
    a.nchunks = (int)sampler_chunks(N, rounds);
    a.acc = a.chunk + (size_t)B * a.nchunks;
    return a;
}

extern "C" int rrl_sampler_num_chunks(int N, int rounds) {
    if (N <= 0 || rounds <= 0 || (long long)rounds * N >= (1LL << 31) - kChunk) return 0;
    return (int)sampler_chunks(N, rounds);
}

extern "C" int rrl_sample_lines_shard_flags(const float *radius, const float *centers, const float *verts1, const float *verts2,
                                            int B, int n1, int n2, int N, int rounds, unsigned long long seed,
                                            unsigned long long offset, const float *uniforms, int shard_rank, int shard_world,
                                            int *out_chunk_counts, void *workspace, size_t workspace_bytes, void *stream) {
    if (!radius || !centers || !verts1 || !verts2 || !out_chunk_counts || !workspace) return RRL_ERR_ARG;
    if (B <= 0 || n1 <= 0 || n2 <= 0 || N <= 0 || rounds <= 0 || shard_world < 1 || shard_rank < 0 || shard_rank >= shard_world) return RRL_ERR_ARG;
    if ((long long)rounds * N >= (1LL << 31) - kChunk) return RRL_ERR_ARG;
    if (workspace_bytes < rrl_sampler_workspace_bytes(B, N, rounds)) return RRL_ERR_WORKSPACE;
    Range range("rrl_sample_lines_shard_flags");
    cudaStream_t s = (cudaStream_t)stream;
    const SamplerArgs a = shard_args(radius, centers, uniforms, B, N, rounds, seed, offset, shard_rank, shard_world, workspace);
    if (cudaMemsetAsync(a.chunk, 0, (size_t)B * (a.nchunks + kPhases + 1) * sizeof(int), s) != cudaSuccess) return RRL_ERR_CUDA;
    bbox_kernel<<<dim3(B, 2), 256, 0, s>>>(verts1, verts2, n1, n2, a.bbox);
    // every round is evaluated (a rank sees only its own counts, so the phases' early exit has nothing to decide on)
    // block x walks the chunks x, x + grid, ...: with a grid that is a multiple of `world`, the blocks x = rank (mod world)
    // own every chunk they visit and the others return at once
    const int cap = (sm_count() * 8 + B - 1) / B;
    int bx = (cap < 1 ? 1 : cap) * shard_world;
    if (bx > a.nchunks) bx = (a.nchunks + shard_world - 1) / shard_world * shard_world;
    flags_kernel<<<dim3(bx, B), kChunk, 0, s>>>(a, 0, 0, a.nchunks);
    count_launch(2);
    if (cudaMemcpyAsync(out_chunk_counts, a.chunk, (size_t)B * a.nchunks * sizeof(int), cudaMemcpyDeviceToDevice, s) != cudaSuccess)
        return RRL_ERR_CUDA;
    return check_launch();
}

extern "C" int rrl_sample_lines_shard_scatter(const float *radius, const float *centers, int B, int N, int rounds,
                                              unsigned long long seed, unsigned long long offset, const float *uniforms,
                                              int shard_rank, int shard_world, int *chunk_counts_global, int *chunk_counts_local,
                                              float *out_lines_local, int *out_counts3, void *workspace, size_t workspace_bytes,
                                              void *stream) {
    if (!radius || !centers || !chunk_counts_global || !chunk_counts_local || !out_lines_local || !out_counts3 || !workspace) return RRL_ERR_ARG;
    if (B <= 0 || N <= 0 || rounds <= 0 || shard_world < 1 || shard_rank < 0 || shard_rank >= shard_world) return RRL_ERR_ARG;
    if ((long long)rounds * N >= (1LL << 31) - kChunk) return RRL_ERR_ARG;
    if (workspace_bytes < rrl_sampler_workspace_bytes(B, N, rounds)) return RRL_ERR_WORKSPACE;
    Range range("rrl_sample_lines_shard_scatter");
    cudaStream_t s = (cudaStream_t)stream;
    const SamplerArgs a = shard_args(radius, centers, uniforms, B, N, rounds, seed, offset, shard_rank, shard_world, workspace);
    shard_scan_kernel<<<B, 1024, 0, s>>>(a, chunk_counts_global, chunk_counts_local, out_counts3);
    int bx = a.nchunks;
    const int cap = (sm_count() * 8 + B - 1) / B;
    if (bx > cap) bx = cap < 1 ? 1 : cap;
    shard_scatter_kernel<<<dim3(bx, B), kChunk, 0, s>>>(a, chunk_counts_local, out_lines_local, out_counts3);
    count_launch(2);
    return check_launch();
}
