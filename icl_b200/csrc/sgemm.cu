// fp32 CUDA-core GEMM family for the ICL heads: every nn.Linear /
// 1x1x1 conv / Conv1d(k=1) on the path is C = act(A*B + bias) with small M or small K, where
// fp32 FFMA keeps exact reference numerics.  The two 13 824 x 13 824 mlp2 Linears at the 24^3
// scale are pure weight streaming (rows <= 16 at K=2): skinny_* kernels read each weight once
// with 128-bit coalesced loads and keep the 16 row-vectors in shared memory / registers.
#include "common.cuh"

#define GM 64
#define GN 64
#define GK 16

// C[m*scm + n*scn] (+)= act( sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] + bias )
// bias_mode: 0 none, 1 bias[n], 2 bias[m].  act: 0 none, 1 GELU(erf).
// pre (optional): the pre-activation value is also stored there (same strides as C) for backward.
__global__ void __launch_bounds__(256) sgemm_k(
    int M, int N, int K, const float* __restrict__ A, long long sam, long long sak, long long sA,
    const float* __restrict__ Bm, long long sbk, long long sbn, long long sB, float* __restrict__ C, long long scm, long long scn,
    long long sC, const float* __restrict__ bias, int bias_mode, int act, int accumulate, float* __restrict__ pre, int ksplit, int k_per) {
  __shared__ __align__(16) float As[GK][GM + 4];
  __shared__ __align__(16) float Bs[GK][GN + 4];
  // split-K (ksplit > 1): blockIdx.z = batch * ksplit + slice; slices combine with atomics into a C the host zeroed
  const int bz = blockIdx.z / ksplit, ks = blockIdx.z % ksplit;
  A += (long long)bz * sA;
  Bm += (long long)bz * sB;
  C += (long long)bz * sC;
  if (pre) pre += (long long)bz * sC;
  const int kbeg = ks * k_per, kend = min(K, kbeg + k_per);
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  const int t = threadIdx.x, tx = t % 16, ty = t / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_kfast = (sak == 1), b_nfast = (sbn == 1);
  for (int k0 = kbeg; k0 < kend; k0 += GK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int m, k;
      if (a_kfast) { k = t % GK; m = t / GK + 16 * j; } else { m = t % GM; k = t / GM + 4 * j; }
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < kend) ? A[gm * sam + gk * sak] : 0.f;
      int n, kb;
      if (b_nfast) { n = t % GN; kb = t / GN + 4 * j; } else { kb = t % GK; n = t / GK + 16 * j; }
      const int gn = n0 + n, gkb = k0 + kb;
      Bs[kb][n] = (gn < N && gkb < kend) ? Bm[gkb * sbk + gn * sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      const long long o = gm * scm + gn * scn;
      if (ksplit > 1) {
        if (ks == 0) { if (bias_mode == 1) v += bias[gn]; else if (bias_mode == 2) v += bias[gm]; }
        atomicAdd(C + o, v);
        continue;
      }
      if (bias_mode == 1) v += bias[gn];
      else if (bias_mode == 2) v += bias[gm];
      if (pre) pre[o] = v;
      if (act == 1) v = gelu_erf(v);
      C[o] = accumulate ? C[o] + v : v;
    }
  }
}

__global__ void zero_strided_k(float* __restrict__ C, int M, int N, long long scm, long long scn, long long sC) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)M * N) C[(long long)blockIdx.y * sC + (i / N) * scm + (i % N) * scn] = 0.f;
}

ICL_API int icl_sgemm(int M, int N, int K, const float* A, long long sam, long long sak, long long sA, const float* Bm, long long sbk,
                      long long sbn, long long sB, float* C, long long scm, long long scn, long long sC, int batch, const float* bias,
                      int bias_mode, int act, int accumulate, float* pre, void* stream) {
  ICL_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0 && batch <= 65535, "sgemm: bad sizes M=%d N=%d K=%d batch=%d", M, N, K, batch);
  dim3 grid(cdiv(N, GN), cdiv(M, GM), batch);
  ICL_REQUIRE(grid.y <= 65535, "sgemm: M too large for grid.y");
  // few output tiles + long reduction (weight gradients of 1x1x1 convs / Linears over all voxels): split K over the SMs
  const long long tiles = (long long)grid.x * grid.y * batch;
  int ksplit = 1, k_per = K;
  if (tiles < 148 && K >= 4096 && act == 0 && pre == nullptr) {
    ksplit = (int)((148 * 4 + tiles - 1) / tiles);
    if (ksplit > K / 1024) ksplit = K / 1024;
    if ((long long)batch * ksplit > 65535) ksplit = 65535 / batch;
    k_per = cdiv(cdiv(K, ksplit), GK) * GK;
    ksplit = cdiv(K, k_per);
  }
  if (ksplit > 1) {
    if (!accumulate) {
      zero_strided_k<<<dim3(cdiv((long long)M * N, 256), batch), 256, 0, as_stream(stream)>>>(C, M, N, scm, scn, sC);
      icl_count_launch(1);
    }
    grid.z = batch * ksplit;
  }
  sgemm_k<<<grid, 256, 0, as_stream(stream)>>>(M, N, K, A, sam, sak, sA, Bm, sbk, sbn, sB, C, scm, scn, sC, bias, bias_mode, act,
                                               accumulate, pre, ksplit, k_per);
  ICL_LAUNCHED("sgemm");
}

// ------------------------------------------------------------------------------------------
// skinny NT:  y[m, n] = act( sum_k x[m, k] * W[n, k] + bias[n] ),  m < M <= 16 per pass.
// x row-major [M, K], W row-major [N, K] (nn.Linear weight).  Block = 4 warps x 4 n each.
// ------------------------------------------------------------------------------------------
#define SK_M 16
#define SK_KC 256
__global__ void __launch_bounds__(128) skinny_nt_k(int M, int N, int K, const float* __restrict__ x, const float* __restrict__ Wt,
                                                   const float* __restrict__ bias, float* __restrict__ y, float* __restrict__ pre, int act) {
  __shared__ float xs[SK_M][SK_KC];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nb = blockIdx.x * 16 + wid * 4;
  for (int mg = 0; mg < M; mg += SK_M) {
    float acc[4][SK_M];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int m = 0; m < SK_M; ++m) acc[j][m] = 0.f;
    for (int k0 = 0; k0 < K; k0 += SK_KC) {
      __syncthreads();
      for (int i = threadIdx.x; i < SK_M * SK_KC; i += 128) {
        const int m = i / SK_KC, k = i % SK_KC;
        xs[m][k] = (mg + m < M && k0 + k < K) ? x[(long long)(mg + m) * K + k0 + k] : 0.f;
      }
      __syncthreads();
#pragma unroll 2
      for (int kk = lane; kk < SK_KC; kk += 32) {
        const int k = k0 + kk;
        float wv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) wv[j] = (k < K && nb + j < N) ? __ldg(&Wt[(long long)(nb + j) * K + k]) : 0.f;
#pragma unroll
        for (int m = 0; m < SK_M; ++m) {
          const float xv = xs[m][kk];
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j][m] = fmaf(wv[j], xv, acc[j][m]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int m = 0; m < SK_M; ++m) {
        const float s = warp_sum(acc[j][m]);
        if (lane == 0 && nb + j < N && mg + m < M) {
          float v = s + (bias ? bias[nb + j] : 0.f);
          const long long o = (long long)(mg + m) * N + nb + j;
          if (pre) pre[o] = v;
          y[o] = act == 1 ? gelu_erf(v) : v;
        }
      }
  }
}
ICL_API int icl_skinny_linear_fwd(int M, int N, int K, const float* x, const float* Wt, const float* bias, float* y, float* pre, int act,
                                  void* stream) {
  skinny_nt_k<<<cdiv(N, 16), 128, 0, as_stream(stream)>>>(M, N, K, x, Wt, bias, y, pre, act);
  ICL_LAUNCHED("skinny_linear_fwd");
}

// ------------------------------------------------------------------------------------------
// skinny NN (data gradient):  dx[m, k] += sum_n dy[m, n] * W[n, k].   dx must be zeroed by the
// caller; the n range is split over blockIdx.y and combined with atomics.
// ------------------------------------------------------------------------------------------
#define SN_NC 64
__global__ void __launch_bounds__(128) skinny_nn_k(int M, int N, int K, const float* __restrict__ dy, const float* __restrict__ Wt,
                                                   float* __restrict__ dx, int n_per) {
  __shared__ float ds[SK_M][SN_NC];
  const int k = (blockIdx.x * 128 + threadIdx.x) * 4;
  const int nbeg = blockIdx.y * n_per, nend = min(N, nbeg + n_per);
  const bool kvalid = k < K;  // K % 4 == 0 required
  for (int mg = 0; mg < M; mg += SK_M) {
    float acc[SK_M][4];
#pragma unroll
    for (int m = 0; m < SK_M; ++m) { acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f; }
    for (int n0 = nbeg; n0 < nend; n0 += SN_NC) {
      __syncthreads();
      for (int i = threadIdx.x; i < SK_M * SN_NC; i += 128) {
        const int m = i / SN_NC, n = i % SN_NC;
        ds[m][n] = (mg + m < M && n0 + n < nend) ? dy[(long long)(mg + m) * N + n0 + n] : 0.f;
      }
      __syncthreads();
      if (kvalid) {
        const int nn = min(SN_NC, nend - n0);
#pragma unroll 4
        for (int n = 0; n < nn; ++n) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(&Wt[(long long)(n0 + n) * K + k]));
#pragma unroll
          for (int m = 0; m < SK_M; ++m) {
            const float g = ds[m][n];
            acc[m][0] = fmaf(g, w4.x, acc[m][0]); acc[m][1] = fmaf(g, w4.y, acc[m][1]);
            acc[m][2] = fmaf(g, w4.z, acc[m][2]); acc[m][3] = fmaf(g, w4.w, acc[m][3]);
          }
        }
      }
    }
    if (kvalid) {
#pragma unroll
      for (int m = 0; m < SK_M; ++m)
        if (mg + m < M) {
          float* dst = dx + (long long)(mg + m) * K + k;
          atomicAdd(dst + 0, acc[m][0]); atomicAdd(dst + 1, acc[m][1]); atomicAdd(dst + 2, acc[m][2]); atomicAdd(dst + 3, acc[m][3]);
        }
    }
  }
}
ICL_API int icl_skinny_linear_dgrad(int M, int N, int K, const float* dy, const float* Wt, float* dx, void* stream) {
  ICL_REQUIRE(K % 4 == 0, "skinny_linear_dgrad: K=%d must be a multiple of 4", K);
  const int gx = cdiv(K, 512);
  int splits = max(1, min(cdiv(N, SN_NC), (148 * 4) / gx));
  const int n_per = cdiv(cdiv(N, splits), SN_NC) * SN_NC;
  splits = cdiv(N, n_per);
  skinny_nn_k<<<dim3(gx, splits), 128, 0, as_stream(stream)>>>(M, N, K, dy, Wt, dx, n_per);
  ICL_LAUNCHED("skinny_linear_dgrad");
}

// ------------------------------------------------------------------------------------------
// rank-M weight gradient:  dW[n, k] = sum_{m<M} dy[m, n] * x[m, k]   (M <= 64), write-bound.
// Also db[n] = sum_m dy[m, n] (done by blockIdx.x == 0).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) outer_wgrad_k(int M, int N, int K, const float* __restrict__ dy, const float* __restrict__ x,
                                                     float* __restrict__ dW, float* __restrict__ db, int accumulate) {
  // block: 16 n rows x 256 k columns (thread = 4 consecutive k of 4 n rows)
  extern __shared__ float sm[];  // dys[M][16] , xs[M][256]
  float* dys = sm;
  float* xs = sm + M * 16;
  const int n0 = blockIdx.y * 16, k0 = blockIdx.x * 256;
  for (int i = threadIdx.x; i < M * 16; i += 256) {
    const int m = i / 16, n = i % 16;
    dys[i] = (n0 + n < N) ? dy[(long long)m * N + n0 + n] : 0.f;
  }
  for (int i = threadIdx.x; i < M * 256; i += 256) {
    const int m = i / 256, k = i % 256;
    xs[i] = (k0 + k < K) ? x[(long long)m * K + k0 + k] : 0.f;
  }
  __syncthreads();
  const int kq = (threadIdx.x % 64) * 4, ng = (threadIdx.x / 64) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  for (int m = 0; m < M; ++m) {
    const float4 xv = *reinterpret_cast<const float4*>(&xs[m * 256 + kq]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float g = dys[m * 16 + ng + i];
      acc[i][0] = fmaf(g, xv.x, acc[i][0]); acc[i][1] = fmaf(g, xv.y, acc[i][1]);
      acc[i][2] = fmaf(g, xv.z, acc[i][2]); acc[i][3] = fmaf(g, xv.w, acc[i][3]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ng + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + kq + j;
      if (k < K) {
        float* d = dW + (long long)n * K + k;
        *d = accumulate ? *d + acc[i][j] : acc[i][j];
      }
    }
  }
  if (db && blockIdx.x == 0 && threadIdx.x < 16 && n0 + threadIdx.x < N) {
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += dys[m * 16 + threadIdx.x];
    db[n0 + threadIdx.x] = accumulate ? db[n0 + threadIdx.x] + s : s;
  }
}
ICL_API int icl_outer_wgrad(int M, int N, int K, const float* dy, const float* x, float* dW, float* db, int accumulate, void* stream) {
  ICL_REQUIRE(M >= 1 && M <= 64, "outer_wgrad: M=%d out of range [1,64]", M);
  dim3 grid(cdiv(K, 256), cdiv(N, 16));
  ICL_REQUIRE(grid.y <= 65535, "outer_wgrad: N too large");
  outer_wgrad_k<<<grid, 256, (size_t)M * (16 + 256) * sizeof(float), as_stream(stream)>>>(M, N, K, dy, x, dW, db, accumulate);
  ICL_LAUNCHED("outer_wgrad");
}

// column sums: out[n] (+)= sum_m a[m*N + n]   (bias gradients of Linear layers)
__global__ void colsum_k(const float* __restrict__ a, float* __restrict__ out, long long M, int N, int accumulate, long long m_per) {
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int r = threadIdx.x >> 5;  // 8 row lanes
  __shared__ float red[8][33];
  const long long mbeg = (long long)blockIdx.y * m_per, mend = min(M, mbeg + m_per);
  float s = 0.f;
  if (n < N)
    for (long long m = mbeg + r; m < mend; m += 8) s += a[m * N + n];
  red[r][threadIdx.x & 31] = s;
  __syncthreads();
  if (r == 0 && n < N) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x & 31];
    if (gridDim.y > 1) atomicAdd(out + n, t);
    else out[n] = accumulate ? out[n] + t : t;
  }
}
// narrow matrices (N in {1,2,4,8,16}, e.g. the bias gradient of the K-class 1x1x1 head over all voxels): element i
// belongs to column i % N; the grid-stride is a multiple of 32, so a thread stays on column (lane % N) and a warp
// reads contiguous memory.  Lanes of equal column combine by xor-shuffles, warps through shared memory, blocks by atomics.
__global__ void __launch_bounds__(256) colsum_narrow_k(const float* __restrict__ a, float* __restrict__ out, long long total, int N) {
  __shared__ float red[8][16];
  float s = 0.f;
  const long long stride = (long long)gridDim.x * 256;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += stride) s += a[i];
  for (int o = 16; o >= N; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane < N) red[wid][lane] = s;
  __syncthreads();
  if (threadIdx.x < N) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(out + threadIdx.x, t);
  }
}
ICL_API int icl_colsum(const float* a, float* out, long long M, int N, int accumulate, void* stream) {
  if (N <= 16 && (32 % N) == 0 && M >= 65536) {
    if (!accumulate) cudaMemsetAsync(out, 0, sizeof(float) * N, as_stream(stream));
    colsum_narrow_k<<<148 * 4, 256, 0, as_stream(stream)>>>(a, out, M * N, N);
    ICL_LAUNCHED("colsum_narrow");
  }
  int ysplit = 1;
  long long m_per = M;
  if (M >= 16384 && cdiv(N, 32) < 148) {
    ysplit = (148 * 2) / cdiv(N, 32);
    if (ysplit > M / 2048) ysplit = (int)(M / 2048);
    if (ysplit < 1) ysplit = 1;
    m_per = (M + ysplit - 1) / ysplit;
    ysplit = (int)((M + m_per - 1) / m_per);
  }
  if (ysplit > 1 && !accumulate) cudaMemsetAsync(out, 0, sizeof(float) * N, as_stream(stream));
  colsum_k<<<dim3(cdiv(N, 32), ysplit), 256, 0, as_stream(stream)>>>(a, out, M, N, accumulate, m_per);
  ICL_LAUNCHED("colsum");
}

// elementwise GELU backward: dx = dy * gelu'(pre)
__global__ void gelu_bwd_k(const float* __restrict__ dy, const float* __restrict__ pre, float* __restrict__ dx, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dx[i] = dy[i] * gelu_erf_grad(pre[i]);
}
ICL_API int icl_gelu_bwd(const float* dy, const float* pre, float* dx, long long n, void* stream) {
  gelu_bwd_k<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(dy, pre, dx, n);
  ICL_LAUNCHED("gelu_bwd");
}
