// se(3) exponential fused with the per-point rigid transform, and its backward (sm_100a).
//
// Replaces Reconstruction_point.Transform/forward (/root/reference/code/loss.py:455-463) and se3.exp3
// (LieAlgebra/se3.py:83-106) with so3.mat (so3.py:17-27) and sinc1/2/3 (sinc.py:5-17,91-103,120-132):
//   theta = |w|, W = hat(w), S = W W, R = I + sinc1 W + sinc2 S, V = I + sinc2 W + sinc3 S, T = V v,
//   out = p @ R + T  (row vectors).
// The backward reduces d loss / d out to G_R = sum p (x) g and G_T = sum g per pair and chains them through
// the closed-form derivative of exp3 -- the 6-DoF twist gradient never needs a dense Jacobian.
// Also the explicit-(R,t) transform used by the DCP / RPM-Net / FMR hooks (utils.py:32-37).
#include "rrl_common.cuh"

namespace rrl {

struct Sinc {
    double a, b, c, da, db, dc;
};

__device__ __forceinline__ Sinc sinc_all(double t) {
    Sinc s;
    const double t2 = t * t;
    if (fabs(t) < 0.01) {        // Taylor branches (sinc.py:7-11, 95-99, 124-128 and the _dt variants)
        s.a = 1 - t2 / 6 * (1 - t2 / 20 * (1 - t2 / 42));
        s.b = 0.5 * (1 - t2 / 12 * (1 - t2 / 30 * (1 - t2 / 56)));
        s.c = 1.0 / 6 * (1 - t2 / 20 * (1 - t2 / 42 * (1 - t2 / 72)));
        s.da = -t / 3 * (1 - t2 / 10 * (1 - t2 / 28 * (1 - t2 / 54)));
        s.db = -t / 12 * (1 - t2 / 5 * (1.0 / 3 - t2 / 56 * (1.0 / 2 - t2 / 135)));
        s.dc = -t / 60 * (1 - t2 / 21 * (1 - t2 / 24 * (1.0 / 2 - t2 / 165)));
    } else {
        const double sn = sin(t), cs = cos(t);
        s.a = sn / t;
        s.b = (1 - cs) / t2;
        s.c = (t - sn) / (t2 * t);
        s.da = cs / t - sn / t2;
        s.db = sn / t2 - 2 * (1 - cs) / (t2 * t);
        s.dc = (3 * sn - t * (cs + 2)) / (t2 * t2);
    }
    return s;
}

__device__ __forceinline__ void hat(const double *w, double *W) {
    W[0] = 0; W[1] = -w[2]; W[2] = w[1];
    W[3] = w[2]; W[4] = 0; W[5] = -w[0];
    W[6] = -w[1]; W[7] = w[0]; W[8] = 0;
}

__device__ __forceinline__ void mm3(const double *A, const double *B, double *C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

__device__ void exp3(const float *twist, float *R, float *T) {
    const double w[3] = {twist[0], twist[1], twist[2]}, v[3] = {twist[3], twist[4], twist[5]};
    const double t = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const Sinc s = sinc_all(t);
    double W[9], S[9];
    hat(w, W);
    mm3(W, W, S);
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = (float)((i % 4 == 0 ? 1.0 : 0.0) + s.a * W[i] + s.b * S[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double acc = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j) acc += ((i == j ? 1.0 : 0.0) + s.b * W[3 * i + j] + s.c * S[3 * i + j]) * v[j];
        T[i] = (float)acc;
    }
}

__global__ void se3_exp_kernel(const float *__restrict__ twist, int B, float *R, float *T) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float r[9], t[3];
    exp3(twist + b * 6, r, t);
    for (int i = 0; i < 9; ++i) R[b * 9 + i] = r[i];
    for (int i = 0; i < 3; ++i) T[b * 3 + i] = t[i];
}

// FMR's se3.Exp (fmr/se_math/se3.py:60-84): the same R and p = V v as exp3, packed as g = [R p; 0 0 0 1] (B,4,4)
__global__ void se3_exp4_kernel(const float *__restrict__ twist, int B, float *__restrict__ g) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float r[9], t[3];
    exp3(twist + b * 6, r, t);
    float *o = g + (long long)b * 16;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) o[4 * i + j] = r[3 * i + j];
        o[4 * i + 3] = t[i];
    }
    o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
}

// Backward of FMR's ExpMap (fmr/se_math/se3.py:133-165) -- NOT the derivative of exp: the reference propagates
//   grad_x[k] = sum_ij grad_g[i][j] * (gen_k g)[i][j],   gen_k = mat(e_k)  (se3.py:27-55), g = exp(x),
// i.e. the left-trivialised tangent.  gen_k g = [hat(e_k) R, hat(e_k) p; 0] for k < 3 and [0, e_{k-3}; 0] for k >= 3.
__global__ void se3_expmap_backward_kernel(const float *__restrict__ twist, const float *__restrict__ grad_g, int B,
                                           float *__restrict__ grad_twist) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float r[9], t[3];
    exp3(twist + b * 6, r, t);
    const float *go = grad_g + (long long)b * 16;
    double G[3][4], M[3][4];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) { G[i][j] = go[4 * i + j]; M[i][j] = r[3 * i + j]; }
        G[i][3] = go[4 * i + 3];
        M[i][3] = t[i];
    }
    // (hat(e_k) M)[i][j] = sum_m hat(e_k)[i][m] M[m][j];  hat(e_0) = [0 0 0; 0 0 -1; 0 1 0], etc.
    double gx[3] = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        gx[0] += -G[1][j] * M[2][j] + G[2][j] * M[1][j];
        gx[1] += G[0][j] * M[2][j] - G[2][j] * M[0][j];
        gx[2] += -G[0][j] * M[1][j] + G[1][j] * M[0][j];
    }
    float *o = grad_twist + (long long)b * 6;
    o[0] = (float)gx[0]; o[1] = (float)gx[1]; o[2] = (float)gx[2];
    o[3] = go[3]; o[4] = go[7]; o[5] = go[11];
}

// out = p @ R + T with R,T either recomputed from the twist (kFromTwist) or read (column convention: R^T applied)
template <bool kFromTwist>
__global__ void __launch_bounds__(256) apply_kernel(const float *__restrict__ twist_or_R, const float *__restrict__ t_in,
                                                    const float *__restrict__ pts, int n, float *__restrict__ out) {
    __shared__ float sR[9], sT[3];
    const int b = blockIdx.y;
    if (threadIdx.x == 0) {
        if (kFromTwist) {
            exp3(twist_or_R + b * 6, sR, sT);
        } else {
            // out = R p + t  ==  p @ R^T + t
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) sR[3 * i + j] = twist_or_R[b * 9 + 3 * j + i];
            for (int i = 0; i < 3; ++i) sT[i] = t_in[b * 3 + i];
        }
    }
    __syncthreads();
    const long long base = (long long)b * n * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = pts[base + 3 * i], y = pts[base + 3 * i + 1], z = pts[base + 3 * i + 2];
#pragma unroll
        for (int c = 0; c < 3; ++c)
            out[base + 3 * i + c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, sR[c]), __fmul_rn(y, sR[3 + c])), __fmul_rn(z, sR[6 + c])), sT[c]);
    }
}

// per pair: acc[0..8] += sum p[m] g[n],  acc[9..11] += sum g
__global__ void __launch_bounds__(256) reduce_pg_kernel(const float *__restrict__ pts, const float *__restrict__ grad, int n,
                                                        double *__restrict__ acc) {
    const int b = blockIdx.y;
    const long long base = (long long)b * n * 3;
    double a[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) a[i] = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float g0 = grad[base + 3 * i], g1 = grad[base + 3 * i + 1], g2 = grad[base + 3 * i + 2];
        if (g0 == 0.f && g1 == 0.f && g2 == 0.f) continue;               // the point gradient is sparse
        const double p[3] = {pts[base + 3 * i], pts[base + 3 * i + 1], pts[base + 3 * i + 2]};
        const double gg[3] = {g0, g1, g2};
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int q = 0; q < 3; ++q) a[3 * m + q] += p[m] * gg[q];
#pragma unroll
        for (int q = 0; q < 3; ++q) a[9 + q] += gg[q];
    }
    __shared__ double sm[8][12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        double v = a[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        double v = 0;
        for (int w = 0; w < 8; ++w) v += sm[w][threadIdx.x];
        if (v != 0.0) atomicAdd(acc + b * 12 + threadIdx.x, v);
    }
}

// chain G_R, G_T through exp3
__global__ void se3_chain_kernel(const float *__restrict__ twist, const double *__restrict__ acc, int B, float *grad_twist) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double *GR = acc + b * 12, *GT = GR + 9;
    const double w[3] = {twist[b * 6], twist[b * 6 + 1], twist[b * 6 + 2]};
    const double v[3] = {twist[b * 6 + 3], twist[b * 6 + 4], twist[b * 6 + 5]};
    const double t = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const Sinc s = sinc_all(t);
    double W[9], S[9];
    hat(w, W);
    mm3(W, W, S);
    const double dat = t > 0 ? s.da / t : -1.0 / 3, dbt = t > 0 ? s.db / t : -1.0 / 12, dct = t > 0 ? s.dc / t : -1.0 / 60;
    for (int k = 0; k < 3; ++k) {
        double e[3] = {0, 0, 0}, E[9], EW[9], WE[9];
        e[k] = 1;
        hat(e, E);
        mm3(E, W, EW);
        mm3(W, E, WE);
        double acc_k = 0;
        for (int i = 0; i < 9; ++i) acc_k += (s.a * E[i] + s.b * (EW[i] + WE[i]) + dat * w[k] * W[i] + dbt * w[k] * S[i]) * GR[i];
        for (int i = 0; i < 3; ++i) {
            double row = 0;
            for (int j = 0; j < 3; ++j)
                row += (s.b * E[3 * i + j] + s.c * (EW[3 * i + j] + WE[3 * i + j]) + dbt * w[k] * W[3 * i + j] + dct * w[k] * S[3 * i + j]) * v[j];
            acc_k += row * GT[i];
        }
        grad_twist[b * 6 + k] = (float)acc_k;
    }
    for (int j = 0; j < 3; ++j) {
        double a = 0;
        for (int i = 0; i < 3; ++i) a += ((i == j ? 1.0 : 0.0) + s.b * W[3 * i + j] + s.c * S[3 * i + j]) * GT[i];
        grad_twist[b * 6 + 3 + j] = (float)a;
    }
}

// explicit (R,t): out = R p + t.  grad_R[i][j] = sum g[i] p[j] = (sum p (x) g)^T, grad_t = sum g, grad_p = R^T g
__global__ void rigid_chain_kernel(const double *__restrict__ acc, int B, float *grad_R, float *grad_t) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) grad_R[b * 9 + 3 * i + j] = (float)acc[b * 12 + 3 * j + i];
    for (int i = 0; i < 3; ++i) grad_t[b * 3 + i] = (float)acc[b * 12 + 9 + i];
}

__global__ void __launch_bounds__(256) rigid_grad_points_kernel(const float *__restrict__ R, const float *__restrict__ grad, int n,
                                                                float *__restrict__ gp) {
    const int b = blockIdx.y;
    const long long base = (long long)b * n * 3;
    float r[9];
    for (int i = 0; i < 9; ++i) r[i] = R[b * 9 + i];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float g0 = grad[base + 3 * i], g1 = grad[base + 3 * i + 1], g2 = grad[base + 3 * i + 2];
        for (int c = 0; c < 3; ++c) gp[base + 3 * i + c] = r[c] * g0 + r[3 + c] * g1 + r[6 + c] * g2;
    }
}

static dim3 point_grid(int B, int n) {
    int bx = (n + 255) / 256;
    if (bx > 592) bx = 592;
    if (bx < 1) bx = 1;
    return dim3(bx, B);
}

}  // namespace rrl

using namespace rrl;

extern "C" int rrl_se3_exp(const float *twist, int B, float *R, float *T, void *stream) {
    if (!twist || !R || !T || B <= 0) return RRL_ERR_ARG;
    se3_exp_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(twist, B, R, T);
    count_launch();
    return check_launch();
}

extern "C" int rrl_se3_exp4(const float *twist, int B, float *g, void *stream) {
    if (!twist || !g || B <= 0) return RRL_ERR_ARG;
    se3_exp4_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(twist, B, g);
    count_launch();
    return check_launch();
}

extern "C" int rrl_se3_expmap_backward(const float *twist, const float *grad_g, int B, float *grad_twist, void *stream) {
    if (!twist || !grad_g || !grad_twist || B <= 0) return RRL_ERR_ARG;
    se3_expmap_backward_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(twist, grad_g, B, grad_twist);
    count_launch();
    return check_launch();
}

extern "C" int rrl_se3_apply(const float *twist, const float *points, int B, int n, float *out, void *stream) {
    if (!twist || !points || !out || B <= 0 || n <= 0) return RRL_ERR_ARG;
    Range r("rrl_se3_apply");
    apply_kernel<true><<<point_grid(B, n), 256, 0, (cudaStream_t)stream>>>(twist, nullptr, points, n, out);
    count_launch();
    return check_launch();
}

extern "C" int rrl_se3_apply_backward(const float *twist, const float *points, const float *grad_out, int B, int n,
                                      float *grad_twist, double *scratch, void *stream) {
    if (!twist || !points || !grad_out || !grad_twist || !scratch || B <= 0 || n <= 0) return RRL_ERR_ARG;
    Range r("rrl_se3_apply_backward");
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(scratch, 0, sizeof(double) * 12 * (size_t)B, s) != cudaSuccess) return RRL_ERR_CUDA;
    reduce_pg_kernel<<<point_grid(B, n), 256, 0, s>>>(points, grad_out, n, scratch);
    se3_chain_kernel<<<(B + 127) / 128, 128, 0, s>>>(twist, scratch, B, grad_twist);
    count_launch(2);
    return check_launch();
}

extern "C" int rrl_se3_chain(const float *twist, const double *acc, int B, float *grad_twist, void *stream) {
    if (!twist || !acc || !grad_twist || B <= 0) return RRL_ERR_ARG;
    se3_chain_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(twist, acc, B, grad_twist);
    count_launch();
    return check_launch();
}

extern "C" int rrl_rigid_apply(const float *R, const float *t, const float *points, int B, int n, float *out, void *stream) {
    if (!R || !t || !points || !out || B <= 0 || n <= 0) return RRL_ERR_ARG;
    apply_kernel<false><<<point_grid(B, n), 256, 0, (cudaStream_t)stream>>>(R, t, points, n, out);
    count_launch();
    return check_launch();
}

extern "C" int rrl_rigid_apply_backward(const float *R, const float *points, const float *grad_out, int B, int n,
                                        float *grad_R, float *grad_t, float *grad_points, double *scratch, void *stream) {
    if (!R || !points || !grad_out || !grad_R || !grad_t || !scratch || B <= 0 || n <= 0) return RRL_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(scratch, 0, sizeof(double) * 12 * (size_t)B, s) != cudaSuccess) return RRL_ERR_CUDA;
    reduce_pg_kernel<<<point_grid(B, n), 256, 0, s>>>(points, grad_out, n, scratch);
    rigid_chain_kernel<<<(B + 127) / 128, 128, 0, s>>>(scratch, B, grad_R, grad_t);
    count_launch(2);
    if (grad_points) {
        rigid_grad_points_kernel<<<point_grid(B, n), 256, 0, s>>>(R, grad_out, n, grad_points);
        count_launch();
    }
    return check_launch();
}
