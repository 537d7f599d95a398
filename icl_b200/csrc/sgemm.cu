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
  // register double buffering: the global loads of K-step i+1 are in flight while step i is multiplied
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int m, k;
      if (a_kfast) { k = t % GK; m = t / GK + 16 * j; } else { m = t % GM; k = t / GM + 4 * j; }
      const int gm = m0 + m, gk = k0 + k;
      ra[j] = (gm < M && gk < kend) ? A[gm * sam + gk * sak] : 0.f;
      int n, kb;
      if (b_nfast) { n = t % GN; kb = t / GN + 4 * j; } else { kb = t % GK; n = t / GK + 16 * j; }
      const int gn = n0 + n, gkb = k0 + kb;
      rb[j] = (gn < N && gkb < kend) ? Bm[gkb * sbk + gn * sbn] : 0.f;
    }
  };
  if (kbeg < kend) fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += GK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int m, k;
      if (a_kfast) { k = t % GK; m = t / GK + 16 * j; } else { m = t % GM; k = t / GM + 4 * j; }
      As[k][m] = ra[j];
      int n, kb;
      if (b_nfast) { n = t % GN; kb = t / GN + 4 * j; } else { kb = t % GK; n = t / GK + 16 * j; }
      Bs[kb][n] = rb[j];
    }
    __syncthreads();
    if (k0 + GK < kend) fetch(k0 + GK);
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

// bias / activation / pre-activation copy of a split-K result (contiguous [batch][M][N])
__global__ void sgemm_post_k(float* __restrict__ C, float* __restrict__ pre, const float* __restrict__ bias, int bias_mode, int act,
                             long long total, int M, int N) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float v = C[i];
    if (bias_mode == 1) v += bias[i % N];
    else if (bias_mode == 2) v += bias[(i / N) % M];
    if (pre) pre[i] = v;
    C[i] = act == 1 ? gelu_erf(v) : v;
  }
}
__global__ void zero_strided_k(float* __restrict__ C, int M, int N, long long scm, long long scn, long long sC) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)M * N) C[(long long)blockIdx.y * sC + (i / N) * scm + (i % N) * scn] = 0.f;
}

// ------------------------------------------------------------------------------------------
// M <= 8 rows (the class-proxy queries: B*K tokens through fc_q / proj / mlp, networks/unet_3D_icl.py:277-280,304-306).
// The tiled kernel above is a chain of K/16 dependent load->sync->FMA steps (~1.5 us each) for these; here every thread
// issues all of its loads up front, so a launch costs about one memory latency.
//   NT: C[m][n] = act(sum_k A[m][k] * W[n][k] + bias[n])   one warp per n, lanes stride k   (A, W rows K-contiguous)
//   NN: C[m][k] = sum_n A[m][n] * W[n][k]                  one thread per k, n split over blockIdx.y (atomics)
// ------------------------------------------------------------------------------------------
#define SM_MAXM 8
__global__ void __launch_bounds__(256) smallm_nt_k(int M, int N, int K, const float* __restrict__ A, const float* __restrict__ W,
                                                   float* __restrict__ C, const float* __restrict__ bias, int act, float* __restrict__ pre) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  float acc[SM_MAXM];
#pragma unroll
  for (int m = 0; m < SM_MAXM; ++m) acc[m] = 0.f;
  const float* w = W + (long long)n * K;
  if ((K & 3) == 0) {
    for (int k = lane * 4; k < K; k += 128) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + k));
#pragma unroll
      for (int m = 0; m < SM_MAXM; ++m)
        if (m < M) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(A + (long long)m * K + k));
          acc[m] += a.x * wv.x + a.y * wv.y + a.z * wv.z + a.w * wv.w;
        }
    }
  } else {
    for (int k = lane; k < K; k += 32) {
      const float wv = __ldg(w + k);
#pragma unroll
      for (int m = 0; m < SM_MAXM; ++m)
        if (m < M) acc[m] += __ldg(A + (long long)m * K + k) * wv;
    }
  }
#pragma unroll
  for (int m = 0; m < SM_MAXM; ++m) {
    if (m >= M) break;
    float v = warp_sum(acc[m]);
    if (lane == 0) {
      v += bias ? bias[n] : 0.f;
      const long long o = (long long)m * N + n;
      if (pre) pre[o] = v;
      C[o] = act == 1 ? gelu_erf(v) : v;
    }
  }
}
__global__ void __launch_bounds__(128) smallm_nn_k(int M, int N, int K, const float* __restrict__ A, const float* __restrict__ W,
                                                   float* __restrict__ C, int n_per) {
  const int k = blockIdx.x * 128 + threadIdx.x;
  const int nbeg = blockIdx.y * n_per, nend = min(N, nbeg + n_per);
  if (k >= K) return;
  float acc[SM_MAXM];
#pragma unroll
  for (int m = 0; m < SM_MAXM; ++m) acc[m] = 0.f;
#pragma unroll 8
  for (int n = nbeg; n < nend; ++n) {
    const float wv = __ldg(W + (long long)n * K + k);
#pragma unroll
    for (int m = 0; m < SM_MAXM; ++m)
      if (m < M) acc[m] = fmaf(__ldg(A + (long long)m * N + n), wv, acc[m]);
  }
#pragma unroll
  for (int m = 0; m < SM_MAXM; ++m)
    if (m < M) {
      if (gridDim.y > 1) atomicAdd(C + (long long)m * K + k, acc[m]);
      else C[(long long)m * K + k] = acc[m];
    }
}

ICL_API int icl_sgemm(int M, int N, int K, const float* A, long long sam, long long sak, long long sA, const float* Bm, long long sbk,
                      long long sbn, long long sB, float* C, long long scm, long long scn, long long sC, int batch, const float* bias,
                      int bias_mode, int act, int accumulate, float* pre, void* stream) {
  ICL_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0 && batch <= 65535, "sgemm: bad sizes M=%d N=%d K=%d batch=%d", M, N, K, batch);
  if (M <= SM_MAXM && batch == 1 && !accumulate && K >= 32) {
    // A [M][K] row-major; W given as B[k][n]:  NT when W is [N][K] K-contiguous, NN when W is [K][N] N-contiguous
    if (sak == 1 && sam == K && sbk == 1 && sbn == K && scn == 1 && scm == N && bias_mode != 2) {
      smallm_nt_k<<<cdiv(N, 8), 256, 0, as_stream(stream)>>>(M, N, K, A, Bm, C, bias_mode == 1 ? bias : nullptr, act, pre);
      ICL_LAUNCHED("sgemm_smallm_nt");
    }
    if (sak == 1 && sam == K && sbn == 1 && sbk == N && scn == 1 && scm == N && bias_mode == 0 && act == 0 && pre == nullptr) {
      // reduction axis K of this GEMM runs over the rows of W [K][N]
      const int gx = cdiv(N, 128);
      int splits = max(1, min(K / 64, (148 * 4) / gx));
      const int n_per = cdiv(K, splits);
      splits = cdiv(K, n_per);
      if (splits > 1) cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, as_stream(stream));
      smallm_nn_k<<<dim3(gx, splits), 128, 0, as_stream(stream)>>>(M, K, N, A, Bm, C, n_per);
      ICL_LAUNCHED("sgemm_smallm_nn");
    }
  }
  dim3 grid(cdiv(N, GN), cdiv(M, GM), batch);
  ICL_REQUIRE(grid.y <= 65535, "sgemm: M too large for grid.y");
  // few output tiles: split K over the SMs (slices combine with atomics into a zeroed C); bias / activation / the
  // pre-activation copy then run in a second tiny pass.  A K loop is a chain of dependent load -> sync -> FMA steps
  // (~1 us each), so 4 tiles x 14 steps on 4 SMs costs ~30 us where 24 CTAs x 3 steps cost ~5.
  const long long tiles = (long long)grid.x * grid.y * batch;
  int ksplit = 1, k_per = K;
  const bool contiguous_c = (scn == 1 && scm == N && (batch == 1 || sC == (long long)M * N));
  const bool needs_post = (act != 0 || pre != nullptr);
  if (tiles < 148 && K >= 64 && !accumulate && (!needs_post || contiguous_c)) {
    ksplit = (int)((148 * 2 + tiles - 1) / tiles);
    if (ksplit > K / 32) ksplit = K / 32;
    if ((long long)batch * ksplit > 65535) ksplit = 65535 / batch;
    if (ksplit < 1) ksplit = 1;
    k_per = cdiv(cdiv(K, ksplit), GK) * GK;
    ksplit = cdiv(K, k_per);
  }
  const bool post = ksplit > 1 && needs_post;
  if (ksplit > 1) {
    zero_strided_k<<<dim3(cdiv((long long)M * N, 256), batch), 256, 0, as_stream(stream)>>>(C, M, N, scm, scn, sC);
    icl_count_launch(1);
    grid.z = batch * ksplit;
  }
  sgemm_k<<<grid, 256, 0, as_stream(stream)>>>(M, N, K, A, sam, sak, sA, Bm, sbk, sbn, sB, C, scm, scn, sC, post ? nullptr : bias,
                                               post ? 0 : bias_mode, post ? 0 : act, accumulate, post ? nullptr : pre, ksplit, k_per);
  if (post) {
    icl_count_launch(1);
    const long long total = (long long)batch * M * N;
    sgemm_post_k<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(C, pre, bias, bias_mode, act, total, M, N);
  }
  ICL_LAUNCHED("sgemm");
}

// ------------------------------------------------------------------------------------------
// Skinny GEMMs against a huge fp32 weight (mlp2 = MLP(N, N, N) over the spatial axis, 13 824 x 13 824 at 24^3,
// networks/unet_3D_icl.py:258-259,267; rows = B*K*heads <= 64).  They stream every weight exactly once, so the bound is
// HBM — but 16+ FMAs per weight on the CUDA cores would cap them far below it.  The products therefore run on the
// tensor cores as error-compensated 3xTF32 (hi*hi + lo*hi + hi*lo with hi = tf32(x), lo = tf32(x - hi): ~21 mantissa
// bits, fp32-level accuracy) via warp-level mma.sync m16n8k8; W fragments come straight from global memory as
// 128-bit loads (the K index inside a 16-wide block is permuted identically on both operands so that no shuffle is
// needed), the small activation operand is split once per chunk into shared memory.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = f2tf32(x);
  lo = f2tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// skinny NT:  y[m, n] = act( sum_k x[m, k] * W[n, k] + bias[n] ),  M <= 16*MT.  Block = 4 warps x 8 weight rows.
// A lane reads 32 contiguous bytes of its weight row per 32-wide K block (4 lanes = one 128-byte line per row).
#define SK_KC 256
#define SK_LD (SK_KC + 2)  // row stride = 2 (mod 32) words: 2*g + 8*t + {0,1} -> conflict-free 64-bit fragment loads
template <int MT>
__global__ void __launch_bounds__(128) skinny_nt_k(int M, int N, int K, const float* __restrict__ x, const float* __restrict__ Wt,
                                                   const float* __restrict__ bias, float* __restrict__ y, float* __restrict__ pre, int act,
                                                   int k_per) {
  extern __shared__ __align__(16) uint32_t sk_smem[];
  uint32_t* xh = sk_smem;                         // [MT*16][SK_LD] tf32 hi
  uint32_t* xl = sk_smem + MT * 16 * SK_LD;       // [MT*16][SK_LD] tf32 lo
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int n0 = (blockIdx.x * 4 + wid) * 8;
  const int nrow = min(n0 + g, N - 1);  // clamp: rows past N are computed but never stored
  const float* wrow = Wt + (long long)nrow * K;
  float acc[MT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) acc[mt][0] = acc[mt][1] = acc[mt][2] = acc[mt][3] = 0.f;
  // gridDim.y > 1: the K range is split over blockIdx.y (more CTAs in flight to cover HBM latency); partial sums are
  // combined with atomics into y (zeroed by the host) and bias / activation run in skinny_bias_act_k afterwards.
  const int kbeg = blockIdx.y * k_per, kend = min(K, kbeg + k_per);
  for (int k0 = kbeg; k0 < kend; k0 += SK_KC) {
    // the weight stream does not depend on the staged activations: issue the whole chunk's loads (16 x 128-bit per lane,
    // 16 KB per warp in flight) BEFORE the activation staging and the barrier, so HBM latency overlaps both
    float4 wv[SK_KC / 32][2];
#pragma unroll
    for (int u = 0; u < SK_KC / 32; ++u) {
      const int k = k0 + u * 32 + 8 * t;
      wv[u][0] = (k < kend) ? __ldg(reinterpret_cast<const float4*>(wrow + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      wv[u][1] = (k + 4 < kend) ? __ldg(reinterpret_cast<const float4*>(wrow + k + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < MT * 16 * (SK_KC / 4); i += 128) {
      const int m = i / (SK_KC / 4), kq = (i % (SK_KC / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M && k0 + kq < kend) v = *reinterpret_cast<const float4*>(x + (long long)m * K + k0 + kq);
      uint4 h, l;
      split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
      *reinterpret_cast<uint2*>(&xh[m * SK_LD + kq]) = make_uint2(h.x, h.y);
      *reinterpret_cast<uint2*>(&xh[m * SK_LD + kq + 2]) = make_uint2(h.z, h.w);
      *reinterpret_cast<uint2*>(&xl[m * SK_LD + kq]) = make_uint2(l.x, l.y);
      *reinterpret_cast<uint2*>(&xl[m * SK_LD + kq + 2]) = make_uint2(l.z, l.w);
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < SK_KC / 32; ++u) {
      const float wf[8] = {wv[u][0].x, wv[u][0].y, wv[u][0].z, wv[u][0].w, wv[u][1].x, wv[u][1].y, wv[u][1].z, wv[u][1].w};
#pragma unroll
      for (int sidx = 0; sidx < 4; ++sidx) {
        // MMA k-slot t   <- weight/activation column 32*u + 8t + 2*sidx
        // MMA k-slot t+4 <- column 32*u + 8t + 2*sidx + 1
        uint32_t b0h, b0l, b1h, b1l;
        split_tf32(wf[2 * sidx], b0h, b0l);
        split_tf32(wf[2 * sidx + 1], b1h, b1l);
        const int col = u * 32 + 8 * t + 2 * sidx;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const uint2 ah0 = *reinterpret_cast<const uint2*>(&xh[(mt * 16 + g) * SK_LD + col]);
          const uint2 ah1 = *reinterpret_cast<const uint2*>(&xh[(mt * 16 + g + 8) * SK_LD + col]);
          const uint2 al0 = *reinterpret_cast<const uint2*>(&xl[(mt * 16 + g) * SK_LD + col]);
          const uint2 al1 = *reinterpret_cast<const uint2*>(&xl[(mt * 16 + g + 8) * SK_LD + col]);
          mma_tf32(acc[mt], ah0.x, ah1.x, ah0.y, ah1.y, b0h, b1h);
          mma_tf32(acc[mt], al0.x, al1.x, al0.y, al1.y, b0h, b1h);
          mma_tf32(acc[mt], ah0.x, ah1.x, ah0.y, ah1.y, b0l, b1l);
        }
      }
    }
  }
  // C fragment: c0 (row g, col 2t), c1 (g, 2t+1), c2 (g+8, 2t), c3 (g+8, 2t+1)
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = mt * 16 + g + (i >> 1) * 8, n = n0 + 2 * t + (i & 1);
      if (m < M && n < N) {
        const long long o = (long long)m * N + n;
        if (gridDim.y > 1) { atomicAdd(y + o, acc[mt][i]); continue; }
        float v = acc[mt][i] + (bias ? bias[n] : 0.f);
        if (pre) pre[o] = v;
        y[o] = act == 1 ? gelu_erf(v) : v;
      }
    }
}
// skinny NT, second form: the WEIGHTS are the MMA A operand (16 weight rows per m16 tile, RW tiles per warp) and the activations
// the B operand (8 activation rows per n8 tile), so one set of shared-memory B fragments serves RW*16 weight rows instead of 8:
// 4x fewer shared-memory wavefronts and 4x less activation re-staging per weight byte than skinny_nt_k (ncu r01p: that kernel
// sat at 51 % of the shared-memory pipe and 57 % issue for 2.6 TB/s of weight stream).
// k-slot mapping as above: MMA k-slot t <- column 32u + 8t + 2s, k-slot t+4 <- column 32u + 8t + 2s + 1, so a lane's A
// fragment comes from its own two 32-byte weight segments and nothing is shuffled.
#define S2_KC 64                 // K per register chunk of weights: RW*16 rows x 64 floats in flight per warp
#define S2_XC 256                // K per staged activation chunk
#define S2_LD (S2_XC + 2)
template <int RW, int NT>
__global__ void __launch_bounds__(128) skinny_nt2_k(int M, int N, int K, const float* __restrict__ x, const float* __restrict__ Wt,
                                                    const float* __restrict__ bias, float* __restrict__ y, float* __restrict__ pre, int act,
                                                    int k_per) {
  extern __shared__ __align__(16) uint32_t sk_smem[];
  uint32_t* xh = sk_smem;                       // [NT*8][S2_LD] tf32 hi
  uint32_t* xl = sk_smem + NT * 8 * S2_LD;      // [NT*8][S2_LD] tf32 lo
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int n0 = (blockIdx.x * 4 + wid) * (RW * 16);
  const float* wrow[2 * RW];
#pragma unroll
  for (int r = 0; r < 2 * RW; ++r) wrow[r] = Wt + (long long)min(n0 + g + 8 * r, N - 1) * K;  // clamp: rows past N are never stored
  float acc[RW][NT][4];
#pragma unroll
  for (int mt = 0; mt < RW; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
  const int kbeg = blockIdx.y * k_per, kend = min(K, kbeg + k_per);
  for (int k0 = kbeg; k0 < kend; k0 += S2_XC) {
    __syncthreads();
    for (int i = threadIdx.x; i < NT * 8 * (S2_XC / 4); i += 128) {
      const int m = i / (S2_XC / 4), kq = (i % (S2_XC / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M && k0 + kq < kend) v = *reinterpret_cast<const float4*>(x + (long long)m * K + k0 + kq);
      uint4 h, l;
      split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
      *reinterpret_cast<uint2*>(&xh[m * S2_LD + kq]) = make_uint2(h.x, h.y);
      *reinterpret_cast<uint2*>(&xh[m * S2_LD + kq + 2]) = make_uint2(h.z, h.w);
      *reinterpret_cast<uint2*>(&xl[m * S2_LD + kq]) = make_uint2(l.x, l.y);
      *reinterpret_cast<uint2*>(&xl[m * S2_LD + kq + 2]) = make_uint2(l.z, l.w);
    }
    __syncthreads();
#pragma unroll 1
    for (int kc = 0; kc < S2_XC && k0 + kc < kend; kc += S2_KC) {
      float4 wv[2 * RW][S2_KC / 32][2];
#pragma unroll
      for (int r = 0; r < 2 * RW; ++r)
#pragma unroll
        for (int u = 0; u < S2_KC / 32; ++u) {
          const int k = k0 + kc + u * 32 + 8 * t;
          wv[r][u][0] = (k < kend) ? __ldg(reinterpret_cast<const float4*>(wrow[r] + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
          wv[r][u][1] = (k + 4 < kend) ? __ldg(reinterpret_cast<const float4*>(wrow[r] + k + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int u = 0; u < S2_KC / 32; ++u) {
#pragma unroll
        for (int sidx = 0; sidx < 4; ++sidx) {
          const int col = kc + u * 32 + 8 * t + 2 * sidx;
          uint2 bh[NT], bl[NT];
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            bh[nt] = *reinterpret_cast<const uint2*>(&xh[(nt * 8 + g) * S2_LD + col]);
            bl[nt] = *reinterpret_cast<const uint2*>(&xl[(nt * 8 + g) * S2_LD + col]);
          }
#pragma unroll
          for (int mt = 0; mt < RW; ++mt) {
            const float4 q0 = wv[2 * mt][u][sidx >> 1], q1 = wv[2 * mt + 1][u][sidx >> 1];
            const float w00 = (sidx & 1) ? q0.z : q0.x, w01 = (sidx & 1) ? q0.w : q0.y;   // row g      : columns col, col + 1
            const float w10 = (sidx & 1) ? q1.z : q1.x, w11 = (sidx & 1) ? q1.w : q1.y;   // row g + 8
            uint32_t a0h, a0l, a1h, a1l, a2h, a2l, a3h, a3l;
            split_tf32(w00, a0h, a0l); split_tf32(w10, a1h, a1l); split_tf32(w01, a2h, a2l); split_tf32(w11, a3h, a3l);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              mma_tf32(acc[mt][nt], a0h, a1h, a2h, a3h, bh[nt].x, bh[nt].y);
              mma_tf32(acc[mt][nt], a0l, a1l, a2l, a3l, bh[nt].x, bh[nt].y);
              mma_tf32(acc[mt][nt], a0h, a1h, a2h, a3h, bl[nt].x, bl[nt].y);
            }
          }
        }
      }
    }
  }
  // C fragment of tile (mt, nt): c0 (weight row g, activation row 2t), c1 (g, 2t+1), c2 (g+8, 2t), c3 (g+8, 2t+1)
#pragma unroll
  for (int mt = 0; mt < RW; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = n0 + mt * 16 + g + (i >> 1) * 8, m = nt * 8 + 2 * t + (i & 1);
        if (m < M && n < N) {
          const long long o = (long long)m * N + n;
          if (gridDim.y > 1) { atomicAdd(y + o, acc[mt][nt][i]); continue; }
          float v = acc[mt][nt][i] + (bias ? bias[n] : 0.f);
          if (pre) pre[o] = v;
          y[o] = act == 1 ? gelu_erf(v) : v;
        }
      }
}
__global__ void skinny_bias_act_k(float* __restrict__ y, float* __restrict__ pre, const float* __restrict__ bias, long long total, int N, int act) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float v = y[i] + (bias ? bias[i % N] : 0.f);
    if (pre) pre[i] = v;
    y[i] = act == 1 ? gelu_erf(v) : v;
  }
}
static int skinny_fwd_rows(int M, int N, int K, const float* x, const float* Wt, const float* bias, float* y, float* pre, int act,
                           void* stream) {
  if ((long long)N * K < (1LL << 24)) {
    // small weight matrices: 8 rows per warp (more, smaller CTAs) keeps more of the machine busy
    const int gx = cdiv(N, 32);
    int ksplit = 1, k_per = K;
    if (K >= 4 * SK_KC && gx < 148 * 8) {
      ksplit = (148 * 8 + gx - 1) / gx;
      if (ksplit > K / (2 * SK_KC)) ksplit = K / (2 * SK_KC);
      if (ksplit > 16) ksplit = 16;
      if (ksplit < 1) ksplit = 1;
      k_per = cdiv(cdiv(K, ksplit), SK_KC) * SK_KC;
      ksplit = cdiv(K, k_per);
    }
    if (ksplit > 1) cudaMemsetAsync(y, 0, sizeof(float) * (size_t)M * N, as_stream(stream));
    const dim3 grid(gx, ksplit);
#define SK_LAUNCH(MT)                                                                                                    \
  {                                                                                                                      \
    const size_t smem = (size_t)2 * MT * 16 * SK_LD * 4;                                                                 \
    static bool cfg = false;                                                                                             \
    if (!cfg) { cudaFuncSetAttribute(skinny_nt_k<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cfg = true; } \
    skinny_nt_k<MT><<<grid, 128, smem, as_stream(stream)>>>(M, N, K, x, Wt, bias, y, pre, act, k_per);                          \
  }
    if (M <= 16) SK_LAUNCH(1) else if (M <= 32) SK_LAUNCH(2) else SK_LAUNCH(4)
#undef SK_LAUNCH
    if (ksplit > 1) {
      icl_count_launch(1);
      skinny_bias_act_k<<<grid_for((long long)M * N, 256), 256, 0, as_stream(stream)>>>(y, pre, bias, (long long)M * N, N, act);
    }
    ICL_LAUNCHED("skinny_linear_fwd");
  }
  // large weight matrices (mlp2 at 24^3: 13824 x 13824): weights as the A operand, RW m16 tiles per warp
  const int RW = M <= 32 ? 2 : 1;
  const int gx = cdiv(N, 64 * RW);
  // K split: as many CTAs as stay resident at once (5 per SM at 90 registers) — measured on 16 x 13824 x 13824: 648 CTAs
  // 0.198 ms, 432 CTAs (better balanced, fewer bytes in flight) 0.218 ms.  kend multiples of S2_XC keep the float4 loads aligned.
  int ksplit = 1, k_per = K;
  for (int ks = 2; ks <= 16 && ks * 2 * S2_XC <= K; ++ks) {
    const int kp = cdiv(cdiv(K, ks), S2_XC) * S2_XC;
    if (gx * cdiv(K, kp) > 148 * 5) break;
    ksplit = cdiv(K, kp); k_per = kp;
  }
  if (ksplit > 1) cudaMemsetAsync(y, 0, sizeof(float) * (size_t)M * N, as_stream(stream));
  const dim3 grid(gx, ksplit);
#define S2_LAUNCH(RW_, NT_)                                                                                                      \
  {                                                                                                                              \
    const size_t smem = (size_t)2 * NT_ * 8 * S2_LD * 4;                                                                         \
    static bool cfg = false;                                                                                                     \
    if (!cfg) { cudaFuncSetAttribute(skinny_nt2_k<RW_, NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cfg = true; } \
    skinny_nt2_k<RW_, NT_><<<grid, 128, smem, as_stream(stream)>>>(M, N, K, x, Wt, bias, y, pre, act, k_per);                      \
  }
  if (M <= 16) { if (RW == 2) S2_LAUNCH(2, 2) else S2_LAUNCH(1, 2) }
  else if (M <= 32) { if (RW == 2) S2_LAUNCH(2, 4) else S2_LAUNCH(1, 4) }
  else S2_LAUNCH(1, 8)
#undef S2_LAUNCH
  if (ksplit > 1) {
    icl_count_launch(1);
    skinny_bias_act_k<<<grid_for((long long)M * N, 256), 256, 0, as_stream(stream)>>>(y, pre, bias, (long long)M * N, N, act);
  }
  ICL_LAUNCHED("skinny_linear_fwd");
}

ICL_API int icl_skinny_linear_fwd(int M, int N, int K, const float* x, const float* Wt, const float* bias, float* y, float* pre, int act,
                                  void* stream) {
  ICL_REQUIRE(M >= 1 && M <= 256 && K % 4 == 0, "skinny_linear_fwd: need 1 <= M <= 256 and K %% 4 == 0 (M=%d K=%d)", M, K);
  // more than 64 rows (config 3: K = 16 classes x 4 heads x 2 samples = 128): one weight pass per block of 64 rows
  for (int m0 = 0; m0 < M; m0 += 64) {
    const int rc = skinny_fwd_rows(min(64, M - m0), N, K, x + (size_t)m0 * K, Wt, bias, y + (size_t)m0 * N, pre ? pre + (size_t)m0 * N : nullptr,
                                   act, stream);
    if (rc) return rc;
  }
  return 0;
}

// skinny NN (data gradient):  dx[m, k] += sum_n dy[m, n] * W[n, k],  M <= 16*MT.  dx must be zeroed by the caller; the n
// range is split over blockIdx.y and combined with atomics.  Warp = 32 output columns (4 MMA column tiles sharing the
// A fragment: a weight float4 at [n][kb + 4g .. 4g+3] feeds column g of tiles 0..3), block = 4 warps = 128 columns.
#define SN_NC 256
#define SN_UB 4
#define SN_LD (SN_NC + 4)  // 4*g + t -> conflict-free 32-bit fragment loads
template <int MT>
__global__ void __launch_bounds__(128) skinny_nn_k(int M, int N, int K, const float* __restrict__ dy, const float* __restrict__ Wt,
                                                   float* __restrict__ dx, int n_per) {
  extern __shared__ __align__(16) uint32_t sk_smem[];
  uint32_t* dh = sk_smem;
  uint32_t* dl = sk_smem + MT * 16 * SN_LD;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int kb = (blockIdx.x * 4 + wid) * 32;
  const int kcol = min(kb + 4 * g, K - 4);  // clamp (K % 4 == 0): columns past K are computed but never stored
  const int nbeg = blockIdx.y * n_per, nend = min(N, nbeg + n_per);
  float acc[MT][4][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[mt][j][0] = acc[mt][j][1] = acc[mt][j][2] = acc[mt][j][3] = 0.f;
  for (int n0 = nbeg; n0 < nend; n0 += SN_NC) {
    __syncthreads();
    for (int i = threadIdx.x; i < MT * 16 * SN_NC; i += 128) {
      const int m = i / SN_NC, n = i % SN_NC;
      const float v = (m < M && n0 + n < nend) ? dy[(long long)m * N + n0 + n] : 0.f;
      split_tf32(v, dh[m * SN_LD + n], dl[m * SN_LD + n]);
    }
    __syncthreads();
    const int steps = min(SN_NC, nend - n0 + 7) / 8;  // rows past nend multiply zeroed dy (weights clamped in range)
#pragma unroll 1
    for (int s0 = 0; s0 < steps; s0 += SN_UB) {
      float4 w0[SN_UB], w1[SN_UB];   // SN_UB steps of 8 weight rows in flight per warp (8 KB)
#pragma unroll
      for (int u = 0; u < SN_UB; ++u) {
        const int r0 = min(n0 + (s0 + u) * 8 + t, N - 1), r1 = min(n0 + (s0 + u) * 8 + t + 4, N - 1);
        w0[u] = __ldg(reinterpret_cast<const float4*>(Wt + (long long)r0 * K + kcol));
        w1[u] = __ldg(reinterpret_cast<const float4*>(Wt + (long long)r1 * K + kcol));
      }
#pragma unroll
      for (int u = 0; u < SN_UB; ++u) {
        if (s0 + u >= steps) break;
        const int nl = (s0 + u) * 8;
        const float f0[4] = {w0[u].x, w0[u].y, w0[u].z, w0[u].w}, f1[4] = {w1[u].x, w1[u].y, w1[u].z, w1[u].w};
        uint32_t ah[MT][4], al[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          ah[mt][0] = dh[(mt * 16 + g) * SN_LD + nl + t];     ah[mt][1] = dh[(mt * 16 + g + 8) * SN_LD + nl + t];
          ah[mt][2] = dh[(mt * 16 + g) * SN_LD + nl + t + 4]; ah[mt][3] = dh[(mt * 16 + g + 8) * SN_LD + nl + t + 4];
          al[mt][0] = dl[(mt * 16 + g) * SN_LD + nl + t];     al[mt][1] = dl[(mt * 16 + g + 8) * SN_LD + nl + t];
          al[mt][2] = dl[(mt * 16 + g) * SN_LD + nl + t + 4]; al[mt][3] = dl[(mt * 16 + g + 8) * SN_LD + nl + t + 4];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t b0h, b0l, b1h, b1l;
          split_tf32(f0[j], b0h, b0l);
          split_tf32(f1[j], b1h, b1l);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma_tf32(acc[mt][j], ah[mt][0], ah[mt][1], ah[mt][2], ah[mt][3], b0h, b1h);
            mma_tf32(acc[mt][j], al[mt][0], al[mt][1], al[mt][2], al[mt][3], b0h, b1h);
            mma_tf32(acc[mt][j], ah[mt][0], ah[mt][1], ah[mt][2], ah[mt][3], b0l, b1l);
          }
        }
      }
    }
  }
  // tile j, C fragment column c (= 2t, 2t+1) is the output column kb + 4*c + j
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = mt * 16 + g + (i >> 1) * 8, k = kb + 4 * (2 * t + (i & 1)) + j;
        if (m < M && k < K) atomicAdd(dx + (long long)m * K + k, acc[mt][j][i]);
      }
}
static int skinny_dgrad_rows(int M, int N, int K, const float* dy, const float* Wt, float* dx, void* stream) {
  const int gx = cdiv(K, 128);
  // all CTAs resident at once (6 per SM at 78 registers / 33 KB): more would run as a partial second wave
  int splits = max(1, min(cdiv(N, SN_NC), (148 * 6) / gx));
  const int n_per = cdiv(cdiv(N, splits), SN_NC) * SN_NC;
  splits = cdiv(N, n_per);
#define SN_LAUNCH(MT)                                                                                                    \
  {                                                                                                                      \
    const size_t smem = (size_t)2 * MT * 16 * SN_LD * 4;                                                                 \
    static bool cfg = false;                                                                                             \
    if (!cfg) { cudaFuncSetAttribute(skinny_nn_k<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cfg = true; } \
    skinny_nn_k<MT><<<dim3(gx, splits), 128, smem, as_stream(stream)>>>(M, N, K, dy, Wt, dx, n_per);                      \
  }
  if (M <= 16) SN_LAUNCH(1) else if (M <= 32) SN_LAUNCH(2) else SN_LAUNCH(4)
#undef SN_LAUNCH
  ICL_LAUNCHED("skinny_linear_dgrad");
}

ICL_API int icl_skinny_linear_dgrad(int M, int N, int K, const float* dy, const float* Wt, float* dx, void* stream) {
  ICL_REQUIRE(M >= 1 && M <= 256 && K % 4 == 0 && K >= 4, "skinny_linear_dgrad: need 1 <= M <= 256 and K %% 4 == 0 (M=%d K=%d)", M, K);
  for (int m0 = 0; m0 < M; m0 += 64) {
    const int rc = skinny_dgrad_rows(min(64, M - m0), N, K, dy + (size_t)m0 * N, Wt, dx + (size_t)m0 * K, stream);
    if (rc) return rc;
  }
  return 0;
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
  static bool cfg = false;
  if (!cfg) { cudaFuncSetAttribute(outer_wgrad_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * (16 + 256) * (int)sizeof(float)); cfg = true; }
  outer_wgrad_k<<<grid, 256, (size_t)M * (16 + 256) * sizeof(float), as_stream(stream)>>>(M, N, K, dy, x, dW, db, accumulate);
  ICL_LAUNCHED("outer_wgrad");
}

// ------------------------------------------------------------------------------------------
// Fused rank-R weight gradient + momentum-SGD update for the mlp2 weights:
//   g[n,k] = sum_r dy[r,n] * x[r,k] + wd * p[n,k];   m = mu * m + g;   p -= lr * m
// The 764 MB gradient of a 13 824 x 13 824 Linear is never written to or re-read from HBM: per element the kernel
// reads p and m and writes them back (16 B instead of 4 (dW write) + 8 (grad accumulation) + 20 (SGD)).  Rows are staged
// through shared memory 32 at a time, so any number of (rank-local or all-gathered) factor rows is accepted.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sgd_factored_k(int R, int N, int K, const float* __restrict__ dy, const float* __restrict__ x,
                                                      float* __restrict__ p, float* __restrict__ m, const float* __restrict__ lr_ptr,
                                                      float mu, float wd) {
  __shared__ float dys[32 * 16];
  __shared__ __align__(16) float xs[32 * 256];
  const int n0 = blockIdx.y * 16, k0 = blockIdx.x * 256;
  const int kq = (threadIdx.x % 64) * 4, ng = (threadIdx.x / 64) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  for (int r0 = 0; r0 < R; r0 += 32) {
    const int rows = min(32, R - r0);
    __syncthreads();
    for (int i = threadIdx.x; i < rows * 16; i += 256) {
      const int r = i / 16, n = i % 16;
      dys[i] = (n0 + n < N) ? dy[(long long)(r0 + r) * N + n0 + n] : 0.f;
    }
    for (int i = threadIdx.x; i < rows * 256; i += 256) {
      const int r = i / 256, k = i % 256;
      xs[i] = (k0 + k < K) ? x[(long long)(r0 + r) * K + k0 + k] : 0.f;
    }
    __syncthreads();
    for (int r = 0; r < rows; ++r) {
      const float4 xv = *reinterpret_cast<const float4*>(&xs[r * 256 + kq]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float g = dys[r * 16 + ng + i];
        acc[i][0] = fmaf(g, xv.x, acc[i][0]); acc[i][1] = fmaf(g, xv.y, acc[i][1]);
        acc[i][2] = fmaf(g, xv.z, acc[i][2]); acc[i][3] = fmaf(g, xv.w, acc[i][3]);
      }
    }
  }
  const float lr = lr_ptr[0];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ng + i;
    if (n >= N) continue;
    const long long o = (long long)n * K + k0 + kq;
    if (k0 + kq + 3 < K && (K & 3) == 0) {
      float4 pv = *reinterpret_cast<float4*>(p + o), mv = *reinterpret_cast<float4*>(m + o);
      mv.x = mu * mv.x + acc[i][0] + wd * pv.x; mv.y = mu * mv.y + acc[i][1] + wd * pv.y;
      mv.z = mu * mv.z + acc[i][2] + wd * pv.z; mv.w = mu * mv.w + acc[i][3] + wd * pv.w;
      pv.x -= lr * mv.x; pv.y -= lr * mv.y; pv.z -= lr * mv.z; pv.w -= lr * mv.w;
      *reinterpret_cast<float4*>(m + o) = mv;
      *reinterpret_cast<float4*>(p + o) = pv;
    } else {
      for (int j = 0; j < 4; ++j)
        if (k0 + kq + j < K) {
          const float b = mu * m[o + j] + acc[i][j] + wd * p[o + j];
          m[o + j] = b; p[o + j] -= lr * b;
        }
    }
  }
}
ICL_API int icl_sgd_factored(int R, int N, int K, const float* dy, const float* x, float* p, float* m, const float* lr_ptr, float mu, float wd,
                             void* stream) {
  ICL_REQUIRE(R >= 1 && N >= 1 && K >= 1, "sgd_factored: bad sizes R=%d N=%d K=%d", R, N, K);
  dim3 grid(cdiv(K, 256), cdiv(N, 16));
  ICL_REQUIRE(grid.y <= 65535, "sgd_factored: N too large");
  sgd_factored_k<<<grid, 256, 0, as_stream(stream)>>>(R, N, K, dy, x, p, m, lr_ptr, mu, wd);
  ICL_LAUNCHED("sgd_factored");
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
