// Kernels of the Inherent-Consistent-Learning heads:
// LayerNorm over the last axis (C or the spatial axis N), the voxel -> class-proxy cross
// attention (Query_Attention, networks/unet_3D_icl.py:283-297) and the depthwise-separable
// conv + batch-stat BatchNorm3d stack (SeparableConv3d, :317-345) on planar [B*K, H, d,h,w] maps.
// All are bandwidth/latency bound: coalesced streaming + warp-shuffle reductions, fp32 math.
#include "common.cuh"

// scratch of the chunked two-stage reductions in this file (see the section before dwconv3d_wgrad_partial_k)
#define RED_CHUNKS 296
#define RED_WS_BYTES (RED_CHUNKS * 16 * 32 * 8)

// ------------------------------------------------------------------------------------------
// LayerNorm over the last axis, one block per row.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_fwd_k(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                       float* __restrict__ y, float* __restrict__ mean_rstd, int C, float eps) {
  __shared__ float red[33];
  const long long row = blockIdx.x;
  const float* xr = x + row * C;
  float s = 0.f;
  for (int i = threadIdx.x; i < C; i += blockDim.x) s += xr[i];
  const float mean = block_sum(s, red) / C;
  float q = 0.f;
  for (int i = threadIdx.x; i < C; i += blockDim.x) { const float d = xr[i] - mean; q += d * d; }
  const float var = block_sum(q, red) / C;
  const float rstd = rsqrtf(var + eps);
  for (int i = threadIdx.x; i < C; i += blockDim.x) y[row * C + i] = (xr[i] - mean) * rstd * w[i] + b[i];
  if (threadIdx.x == 0) { mean_rstd[row * 2] = mean; mean_rstd[row * 2 + 1] = rstd; }
}
ICL_API int icl_layernorm_fwd(const float* x, const float* w, const float* b, float* y, float* mean_rstd, long long rows, int C, float eps,
                              void* stream) {
  ICL_REQUIRE(rows > 0 && rows < 2147483647LL, "layernorm_fwd: bad rows");
  layernorm_fwd_k<<<(unsigned)rows, C >= 1024 ? 256 : (C >= 128 ? 128 : 64), 0, as_stream(stream)>>>(x, w, b, y, mean_rstd, C, eps);
  ICL_LAUNCHED("layernorm_fwd");
}

// dx = rstd * (dy*w - mean(dy*w) - xh * mean(dy*w*xh))
__global__ void __launch_bounds__(256) layernorm_bwd_dx_k(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ mean_rstd, float* __restrict__ dx, int C) {
  __shared__ float red[33];
  const long long row = blockIdx.x;
  const float mean = mean_rstd[row * 2], rstd = mean_rstd[row * 2 + 1];
  const float* xr = x + row * C; const float* gr = dy + row * C;
  float a = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    const float g = gr[i] * w[i], xh = (xr[i] - mean) * rstd;
    a += g; c += g * xh;
  }
  a = block_sum(a, red) / C;
  c = block_sum(c, red) / C;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    const float g = gr[i] * w[i], xh = (xr[i] - mean) * rstd;
    dx[row * C + i] = rstd * (g - a - xh * c);
  }
}
// dw[c] += sum_rows dy*xh ; db[c] += sum_rows dy.  grid (col tiles of 32, row chunks); atomics at the end.
__global__ void __launch_bounds__(256) layernorm_bwd_wb_k(const float* __restrict__ dy, const float* __restrict__ x,
                                                          const float* __restrict__ mean_rstd, float* __restrict__ dw, float* __restrict__ db,
                                                          long long rows, int C, long long rows_per) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  const long long r0 = (long long)blockIdx.y * rows_per, r1 = min(rows, r0 + rows_per);
  float a = 0.f, bsum = 0.f;
  if (c < C)
    for (long long r = r0 + rl; r < r1; r += 8) {
      const float g = dy[r * C + c];
      a += g * (x[r * C + c] - mean_rstd[r * 2]) * mean_rstd[r * 2 + 1];
      bsum += g;
    }
  __shared__ float red[2][8][33];
  red[0][rl][threadIdx.x & 31] = a; red[1][rl][threadIdx.x & 31] = bsum;
  __syncthreads();
  if (rl == 0 && c < C) {
    float ta = 0.f, tb = 0.f;
    for (int k = 0; k < 8; ++k) { ta += red[0][k][threadIdx.x & 31]; tb += red[1][k][threadIdx.x & 31]; }
    atomicAdd(&dw[c], ta); atomicAdd(&db[c], tb);
  }
}
ICL_API int icl_layernorm_bwd(const float* dy, const float* x, const float* w, const float* mean_rstd, float* dx, float* dw, float* db,
                              long long rows, int C, void* stream) {
  if (dx) {
    layernorm_bwd_dx_k<<<(unsigned)rows, C >= 1024 ? 256 : (C >= 128 ? 128 : 64), 0, as_stream(stream)>>>(dy, x, w, mean_rstd, dx, C);
    icl_count_launch(1);
  }
  if (dw) {  // dw / db must be zero-initialised (or hold the value to accumulate into)
    const int gx = cdiv(C, 32);
    int gy = (int)max((long long)1, min((rows + 63) / 64, (long long)max(1, 148 * 4 / gx)));
    const long long rows_per = (rows + gy - 1) / gy;
    layernorm_bwd_wb_k<<<dim3(gx, gy), 256, 0, as_stream(stream)>>>(dy, x, mean_rstd, dw, db, rows, C, rows_per);
    icl_count_launch(1);
  }
  return icl_check_launch("layernorm_bwd");
}

// ------------------------------------------------------------------------------------------
// Proxy cross-attention.  ql: fc_q output, flat [B, K*C] read as [B,H,K,hd];
// kv: fc_kv output [B,N,2C].  map[b,k,h,n] = scale * <q[b,h,k,:], k[b,n,h,:]> (PRE-softmax, the
// tensor the reference returns); xv[b,h,k,:] = sum_n softmax_n(map[b,k,h,:])[n] * v[b,n,h,:].
// ------------------------------------------------------------------------------------------
#define PA_MAXHD 64
__global__ void __launch_bounds__(256) proxy_logits_k(const float* __restrict__ ql, const float* __restrict__ kv, float* __restrict__ map,
                                                      int B, int N, int C, int H, int K, float scale) {
  extern __shared__ float qs[];  // [K][hd] for this (b, h)
  const int hd = C / H;
  const int b = blockIdx.z, h = blockIdx.y;
  for (int i = threadIdx.x; i < K * hd; i += blockDim.x) qs[i] = ql[(long long)b * K * C + (long long)h * K * hd + i];
  __syncthreads();
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    float kr[PA_MAXHD];
    const float* kp = kv + ((long long)b * N + n) * 2 * C + h * hd;
#pragma unroll 4
    for (int d = 0; d < hd; ++d) kr[d] = kp[d];
    for (int k = 0; k < K; ++k) {
      float s = 0.f;
      for (int d = 0; d < hd; ++d) s = fmaf(qs[k * hd + d], kr[d], s);
      map[(((long long)b * K + k) * H + h) * N + n] = s * scale;
    }
  }
}
// one block per (b,h,k): softmax stats over N and xv
__global__ void __launch_bounds__(256) proxy_av_k(const float* __restrict__ map, const float* __restrict__ kv, float* __restrict__ xv,
                                                  float* __restrict__ mstat, int B, int N, int C, int H, int K) {
  __shared__ float red[33];
  __shared__ float accs[8][PA_MAXHD];
  const int hd = C / H;
  int id = blockIdx.x;
  const int k = id % K; id /= K;
  const int h = id % H; const int b = id / H;
  const float* row = map + (((long long)b * K + k) * H + h) * N;
  float mx = -INFINITY;
  for (int n = threadIdx.x; n < N; n += blockDim.x) mx = fmaxf(mx, row[n]);
  mx = warp_max(mx);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = -INFINITY;
  for (int i = 0; i < (blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float se = 0.f;
  float acc[PA_MAXHD];
  for (int d = 0; d < hd; ++d) acc[d] = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float e = __expf(row[n] - mx);
    se += e;
    const float* vp = kv + ((long long)b * N + n) * 2 * C + C + h * hd;
    for (int d = 0; d < hd; ++d) acc[d] = fmaf(e, vp[d], acc[d]);
  }
  se = block_sum(se, red);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int d = 0; d < hd; ++d) {
    const float s = warp_sum(acc[d]);
    if (lane == 0) accs[wid][d] = s;
  }
  __syncthreads();
  if (threadIdx.x < hd) {
    float s = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) s += accs[i][threadIdx.x];
    xv[(long long)b * K * C + ((long long)h * K + k) * hd + threadIdx.x] = s / se;
  }
  if (threadIdx.x == 0) { mstat[blockIdx.x * 2] = mx; mstat[blockIdx.x * 2 + 1] = se; }
}
// ------------------------------------------------------------------------------------------
// head_dim 16 form (every ICL head of the reference: C / heads = 16) with the voxel axis split over blocks.  The reductions over
// N — softmax statistics, softmax(map) @ v, and dq = dl^T k in backward — are "weighted column sums" sum_n w[n] * X[n, 0:16]
// computed per chunk of PA_CH voxels (grid = chunks x (b,h,k): 224 blocks at 24^3, K = 2, where one block per (b,h,k) gives 16)
// and added in a fixed order by a small combine kernel.  The softmax max comes from per-chunk maxima written by the logits
// kernel, so chunk partials add without rescaling; backward needs no pass for sum_n p[n] dP[n]: it equals <dxv, xv>.
// ------------------------------------------------------------------------------------------
#define PA_HD 16
#define PA_CH 1024
__device__ __forceinline__ void load16(const float* __restrict__ p, float* r) {
#pragma unroll
  for (int q = 0; q < 4; ++q) { const float4 t = reinterpret_cast<const float4*>(p)[q]; r[4 * q] = t.x; r[4 * q + 1] = t.y; r[4 * q + 2] = t.z; r[4 * q + 3] = t.w; }
}
__global__ void __launch_bounds__(256) proxy_logits16_k(const float* __restrict__ ql, const float* __restrict__ kv, float* __restrict__ map,
                                                        float* __restrict__ pmax, int B, int N, int C, int H, int K, float scale, int chunks) {
  __shared__ float qs[16 * PA_HD];
  __shared__ float red[8][16];
  const int b = blockIdx.z, h = blockIdx.y, chunk = blockIdx.x;
  for (int i = threadIdx.x; i < K * PA_HD; i += 256) qs[i] = ql[(long long)b * K * C + (long long)h * K * PA_HD + i];
  __syncthreads();
  float mx[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) mx[k] = -INFINITY;
  const int n1 = min(N, (chunk + 1) * PA_CH);
  for (int n = chunk * PA_CH + threadIdx.x; n < n1; n += 256) {
    float kr[PA_HD];
    load16(kv + ((long long)b * N + n) * 2 * C + h * PA_HD, kr);
#pragma unroll
    for (int k = 0; k < 16; ++k) if (k < K) {
      float sacc = 0.f;
#pragma unroll
      for (int d = 0; d < PA_HD; ++d) sacc = fmaf(qs[k * PA_HD + d], kr[d], sacc);
      sacc *= scale;
      map[(((long long)b * K + k) * H + h) * N + n] = sacc;
      mx[k] = fmaxf(mx[k], sacc);
    }
  }
  if (pmax) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 16; ++k) if (k < K) { const float m = warp_max(mx[k]); if (lane == 0) red[wid][k] = m; }
    __syncthreads();
    if (threadIdx.x < K) {
      float m = -INFINITY;
      for (int w = 0; w < 8; ++w) m = fmaxf(m, red[w][threadIdx.x]);
      pmax[((long long)(b * H + h) * K + threadIdx.x) * chunks + chunk] = m;
    }
  }
}
// part[bhk][chunk][0..15] = sum_n w[n] * X[n][0:16], part[..][16] = sum_n w[n]
//   MODE 0: w = exp(map[n] - max over the chunk maxima), X = v (x_off = C)       MODE 1: w = wsrc[n] (dl), X = k (x_off = 0)
template <int MODE>
__global__ void __launch_bounds__(256) proxy_wsum16_k(const float* __restrict__ wsrc, const float* __restrict__ kv, int x_off,
                                                      const float* __restrict__ pmax, float* __restrict__ part, int B, int N, int C, int H, int K,
                                                      int chunks) {
  __shared__ float red[8][17];
  const int bhk = blockIdx.y, chunk = blockIdx.x;
  const int k = bhk % K, h = (bhk / K) % H, b = bhk / (K * H);
  float gmax = 0.f;
  if (MODE == 0) {
    gmax = -INFINITY;
    for (int c = 0; c < chunks; ++c) gmax = fmaxf(gmax, pmax[(long long)bhk * chunks + c]);
  }
  const float* row = wsrc + (((long long)b * K + k) * H + h) * N;
  float acc[PA_HD], se = 0.f;
#pragma unroll
  for (int d = 0; d < PA_HD; ++d) acc[d] = 0.f;
  const int n1 = min(N, (chunk + 1) * PA_CH);
  for (int n = chunk * PA_CH + threadIdx.x; n < n1; n += 256) {
    const float wv = MODE == 0 ? __expf(row[n] - gmax) : row[n];
    float xr[PA_HD];
    load16(kv + ((long long)b * N + n) * 2 * C + x_off + h * PA_HD, xr);
    se += wv;
#pragma unroll
    for (int d = 0; d < PA_HD; ++d) acc[d] = fmaf(wv, xr[d], acc[d]);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int d = 0; d < PA_HD; ++d) { const float t = warp_sum(acc[d]); if (lane == 0) red[wid][d] = t; }
  { const float t = warp_sum(se); if (lane == 0) red[wid][16] = t; }
  __syncthreads();
  if (threadIdx.x < 17) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    part[((long long)bhk * chunks + chunk) * 17 + threadIdx.x] = t;
  }
}
// MODE 0: xv[b, (h*K + k)*16 + d] = sum_c part[d] / sum_c part[16]; mstat[bhk] = (max, sum exp).   MODE 1: dql[...] = scale * sum_c part[d]
template <int MODE>
__global__ void proxy_combine16_k(const float* __restrict__ part, const float* __restrict__ pmax, float* __restrict__ out, float* __restrict__ mstat,
                                  int BHK, int H, int K, int C, int chunks, float scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BHK * PA_HD) return;
  const int bhk = i / PA_HD, d = i % PA_HD;
  const int k = bhk % K, h = (bhk / K) % H, b = bhk / (K * H);
  float a = 0.f, se = 0.f;
  for (int c = 0; c < chunks; ++c) { a += part[((long long)bhk * chunks + c) * 17 + d]; se += part[((long long)bhk * chunks + c) * 17 + 16]; }
  const long long o = (long long)b * K * C + ((long long)h * K + k) * PA_HD + d;
  if (MODE == 0) {
    out[o] = a / se;
    if (d == 0) {
      float gmax = -INFINITY;
      for (int c = 0; c < chunks; ++c) gmax = fmaxf(gmax, pmax[(long long)bhk * chunks + c]);
      mstat[2 * bhk] = gmax; mstat[2 * bhk + 1] = se;
    }
  } else {
    out[o] = a * scale;
  }
}
// backward per voxel (thread per (b, n, h)): dl[k, n] = dmap + p (dP - <dxv, xv>), dk = scale * sum_k dl q_k, dv = sum_k p dxv_k
__global__ void __launch_bounds__(256) proxy_bwd_vox16_k(const float* __restrict__ dmap, const float* __restrict__ dxv, const float* __restrict__ xv,
                                                         const float* __restrict__ map, const float* __restrict__ ql, const float* __restrict__ kv,
                                                         const float* __restrict__ mstat, float* __restrict__ dl, float* __restrict__ dkv, int B,
                                                         int N, int C, int H, int K, float scale) {
  __shared__ float qs[16 * PA_HD], gs[16 * PA_HD], ms[32], Ds[16];
  const int b = blockIdx.z, h = blockIdx.y;
  for (int i = threadIdx.x; i < K * PA_HD; i += 256) {
    qs[i] = ql[(long long)b * K * C + (long long)h * K * PA_HD + i];
    gs[i] = dxv ? dxv[(long long)b * K * C + (long long)h * K * PA_HD + i] : 0.f;
  }
  if (threadIdx.x < K) {
    const int id = (b * H + h) * K + threadIdx.x;
    ms[2 * threadIdx.x] = mstat ? mstat[id * 2] : 0.f;
    ms[2 * threadIdx.x + 1] = mstat ? 1.f / mstat[id * 2 + 1] : 0.f;
    float dsum = 0.f;
    if (dxv) {
      const long long o = (long long)b * K * C + ((long long)h * K + threadIdx.x) * PA_HD;
      for (int d = 0; d < PA_HD; ++d) dsum = fmaf(dxv[o + d], xv[o + d], dsum);
    }
    Ds[threadIdx.x] = dsum;
  }
  __syncthreads();
  for (int n = blockIdx.x * 256 + threadIdx.x; n < N; n += gridDim.x * 256) {
    float kr[PA_HD], vr[PA_HD], dk[PA_HD], dv[PA_HD];
    const float* base = kv + ((long long)b * N + n) * 2 * C + h * PA_HD;
    load16(base, kr);
    if (dxv) load16(base + C, vr);
#pragma unroll
    for (int d = 0; d < PA_HD; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
    for (int k = 0; k < K; ++k) {
      const long long o = (((long long)b * K + k) * H + h) * N + n;
      float g = dmap ? dmap[o] : 0.f;
      if (dxv) {
        const float p = __expf(map[o] - ms[2 * k]) * ms[2 * k + 1];
        float dP = 0.f;
#pragma unroll
        for (int d = 0; d < PA_HD; ++d) dP = fmaf(gs[k * PA_HD + d], vr[d], dP);
        g += p * (dP - Ds[k]);
#pragma unroll
        for (int d = 0; d < PA_HD; ++d) dv[d] = fmaf(p, gs[k * PA_HD + d], dv[d]);
      }
      dl[o] = g;
      const float gsc = g * scale;
#pragma unroll
      for (int d = 0; d < PA_HD; ++d) dk[d] = fmaf(gsc, qs[k * PA_HD + d], dk[d]);
    }
    float* o = dkv + ((long long)b * N + n) * 2 * C + h * PA_HD;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      reinterpret_cast<float4*>(o)[q] = make_float4(dk[4 * q], dk[4 * q + 1], dk[4 * q + 2], dk[4 * q + 3]);
      reinterpret_cast<float4*>(o + C)[q] = make_float4(dv[4 * q], dv[4 * q + 1], dv[4 * q + 2], dv[4 * q + 3]);
    }
  }
}
static inline bool proxy16_ok(int C, int H, int K, int N, float* ws) {
  const int chunks = cdiv(N, PA_CH);
  return ws != nullptr && C % H == 0 && C / H == PA_HD && K <= 16 && C % 4 == 0 && (long long)chunks * 18 * 4 <= RED_WS_BYTES;
}

ICL_API int icl_proxy_attn_fwd(const float* ql, const float* kv, float* map, float* xv, float* mstat, int B, int N, int C, int H, int K,
                               float scale, int want_xv, float* ws, void* stream) {
  const int hd = C / H;
  ICL_REQUIRE(C % H == 0 && hd <= PA_MAXHD, "proxy_attn: head_dim %d unsupported (max %d)", hd, PA_MAXHD);
  const int chunks = cdiv(N, PA_CH), BHK = B * H * K;
  if (proxy16_ok(C, H, K, N, ws) && (long long)BHK * chunks * 18 * 4 <= RED_WS_BYTES) {
    float* pmax = ws;
    float* part = ws + (long long)BHK * chunks;
    proxy_logits16_k<<<dim3(chunks, H, B), 256, 0, as_stream(stream)>>>(ql, kv, map, want_xv ? pmax : nullptr, B, N, C, H, K, scale, chunks);
    icl_count_launch(1);
    if (want_xv) {
      proxy_wsum16_k<0><<<dim3(chunks, BHK), 256, 0, as_stream(stream)>>>(map, kv, C, pmax, part, B, N, C, H, K, chunks);
      icl_count_launch(1);
      proxy_combine16_k<0><<<cdiv(BHK * PA_HD, 128), 128, 0, as_stream(stream)>>>(part, pmax, xv, mstat, BHK, H, K, C, chunks, scale);
      icl_count_launch(1);
    }
    return icl_check_launch("proxy_attn_fwd");
  }
  proxy_logits_k<<<dim3(cdiv(N, 256), H, B), 256, K * hd * sizeof(float), as_stream(stream)>>>(ql, kv, map, B, N, C, H, K, scale);
  icl_count_launch(1);
  if (want_xv) {
    proxy_av_k<<<B * H * K, 256, 0, as_stream(stream)>>>(map, kv, xv, mstat, B, N, C, H, K);
    icl_count_launch(1);
  }
  return icl_check_launch("proxy_attn_fwd");
}

// backward, stage 1 (block per (b,h,k)): total gradient of the logits
//   dl[n] = dmap[n] + p[n] * (dP[n] - sum_n' p[n'] dP[n']),  dP[n] = <dxv, v[n]>   (second term only if dxv)
// written to dl (may alias a scratch buffer), and dq[b,h,k,:] = scale * sum_n dl[n] * k[n,:].
__global__ void __launch_bounds__(256) proxy_bwd1_k(const float* __restrict__ dmap, const float* __restrict__ dxv, const float* __restrict__ map,
                                                    const float* __restrict__ kv, const float* __restrict__ mstat, float* __restrict__ dl,
                                                    float* __restrict__ dql, int B, int N, int C, int H, int K, float scale) {
  __shared__ float red[33];
  __shared__ float accs[8][PA_MAXHD];
  __shared__ float gx[PA_MAXHD];
  const int hd = C / H;
  int id = blockIdx.x;
  const int k = id % K; id /= K;
  const int h = id % H; const int b = id / H;
  const long long ro = (((long long)b * K + k) * H + h) * N;
  float Dsum = 0.f;
  const float mx = mstat ? mstat[blockIdx.x * 2] : 0.f, inv = mstat ? 1.f / mstat[blockIdx.x * 2 + 1] : 0.f;
  if (dxv) {
    if (threadIdx.x < hd) gx[threadIdx.x] = dxv[(long long)b * K * C + ((long long)h * K + k) * hd + threadIdx.x];
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      const float p = __expf(map[ro + n] - mx) * inv;
      const float* vp = kv + ((long long)b * N + n) * 2 * C + C + h * hd;
      float dP = 0.f;
      for (int d = 0; d < hd; ++d) dP = fmaf(gx[d], vp[d], dP);
      Dsum += p * dP;
    }
    Dsum = block_sum(Dsum, red);
  }
  float acc[PA_MAXHD];
  for (int d = 0; d < hd; ++d) acc[d] = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float g = dmap ? dmap[ro + n] : 0.f;
    if (dxv) {
      const float p = __expf(map[ro + n] - mx) * inv;
      const float* vp = kv + ((long long)b * N + n) * 2 * C + C + h * hd;
      float dP = 0.f;
      for (int d = 0; d < hd; ++d) dP = fmaf(gx[d], vp[d], dP);
      g += p * (dP - Dsum);
    }
    dl[ro + n] = g;
    const float* kp = kv + ((long long)b * N + n) * 2 * C + h * hd;
    for (int d = 0; d < hd; ++d) acc[d] = fmaf(g, kp[d], acc[d]);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int d = 0; d < hd; ++d) {
    const float s = warp_sum(acc[d]);
    if (lane == 0) accs[wid][d] = s;
  }
  __syncthreads();
  if (threadIdx.x < hd) {
    float s = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) s += accs[i][threadIdx.x];
    dql[(long long)b * K * C + ((long long)h * K + k) * hd + threadIdx.x] = s * scale;
  }
}
// stage 2 (thread per (b,n,h)): dk = scale * sum_k dl[k,n] q[k,:],  dv = sum_k p[k,n] dxv[k,:]
__global__ void __launch_bounds__(256) proxy_bwd2_k(const float* __restrict__ dl, const float* __restrict__ dxv, const float* __restrict__ map,
                                                    const float* __restrict__ ql, const float* __restrict__ mstat, float* __restrict__ dkv,
                                                    int B, int N, int C, int H, int K, float scale) {
  extern __shared__ float sm[];  // qs[K*hd], gs[K*hd], ms[2K]
  const int hd = C / H;
  float* qs = sm; float* gs = sm + K * hd; float* ms = gs + K * hd;
  const int b = blockIdx.z, h = blockIdx.y;
  for (int i = threadIdx.x; i < K * hd; i += blockDim.x) {
    qs[i] = ql[(long long)b * K * C + (long long)h * K * hd + i];
    gs[i] = dxv ? dxv[(long long)b * K * C + (long long)h * K * hd + i] : 0.f;
  }
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const int id = (b * H + h) * K + i;
    ms[2 * i] = mstat ? mstat[id * 2] : 0.f; ms[2 * i + 1] = mstat ? 1.f / mstat[id * 2 + 1] : 0.f;
  }
  __syncthreads();
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    float dk[PA_MAXHD], dv[PA_MAXHD];
    for (int d = 0; d < hd; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
    for (int k = 0; k < K; ++k) {
      const long long o = (((long long)b * K + k) * H + h) * N + n;
      const float g = dl[o] * scale;
      for (int d = 0; d < hd; ++d) dk[d] = fmaf(g, qs[k * hd + d], dk[d]);
      if (dxv) {
        const float p = __expf(map[o] - ms[2 * k]) * ms[2 * k + 1];
        for (int d = 0; d < hd; ++d) dv[d] = fmaf(p, gs[k * hd + d], dv[d]);
      }
    }
    float* o = dkv + ((long long)b * N + n) * 2 * C + h * hd;
    for (int d = 0; d < hd; ++d) { o[d] = dk[d]; o[C + d] = dv[d]; }
  }
}
ICL_API int icl_proxy_attn_bwd(const float* dmap, const float* dxv, const float* xv, const float* map, const float* ql, const float* kv,
                               const float* mstat, float* dl_scratch, float* dql, float* dkv, int B, int N, int C, int H, int K, float scale,
                               float* ws, void* stream) {
  const int hd = C / H;
  ICL_REQUIRE(C % H == 0 && hd <= PA_MAXHD, "proxy_attn_bwd: head_dim %d unsupported", hd);
  ICL_REQUIRE(dxv == nullptr || mstat != nullptr, "proxy_attn_bwd: softmax stats required when dxv is given");
  const int chunks = cdiv(N, PA_CH), BHK = B * H * K;
  if (proxy16_ok(C, H, K, N, ws) && (long long)BHK * chunks * 18 * 4 <= RED_WS_BYTES && (dxv == nullptr || xv != nullptr)) {
    float* part = ws + (long long)BHK * chunks;
    proxy_bwd_vox16_k<<<dim3(cdiv(N, 256), H, B), 256, 0, as_stream(stream)>>>(dmap, dxv, xv, map, ql, kv, mstat, dl_scratch, dkv, B, N, C, H, K, scale);
    icl_count_launch(1);
    proxy_wsum16_k<1><<<dim3(chunks, BHK), 256, 0, as_stream(stream)>>>(dl_scratch, kv, 0, nullptr, part, B, N, C, H, K, chunks);
    icl_count_launch(1);
    proxy_combine16_k<1><<<cdiv(BHK * PA_HD, 128), 128, 0, as_stream(stream)>>>(part, nullptr, dql, nullptr, BHK, H, K, C, chunks, scale);
    icl_count_launch(1);
    return icl_check_launch("proxy_attn_bwd");
  }
  proxy_bwd1_k<<<B * H * K, 256, 0, as_stream(stream)>>>(dmap, dxv, map, kv, mstat, dl_scratch, dql, B, N, C, H, K, scale);
  icl_count_launch(1);
  proxy_bwd2_k<<<dim3(cdiv(N, 256), H, B), 256, (2 * K * hd + 2 * K) * sizeof(float), as_stream(stream)>>>(
      dl_scratch, dxv, map, ql, mstat, dkv, B, N, C, H, K, scale);
  icl_count_launch(1);
  return icl_check_launch("proxy_attn_bwd");
}

// ------------------------------------------------------------------------------------------
// Planar [NB, CH, d, h, w] depthwise 3x3x3 conv (groups = CH, no bias, zero pad 1).
// flip = 1 gives the data gradient (correlation with the flipped kernel).
// ------------------------------------------------------------------------------------------
__global__ void dwconv3d_k(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, int NB, int CH, int d, int h,
                           int wd, int flip) {
  const long long S = (long long)d * h * wd, total = (long long)NB * CH * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long v = i;
    const int xx = (int)(v % wd); v /= wd;
    const int yy = (int)(v % h); v /= h;
    const int zz = (int)(v % d); v /= d;
    const int c = (int)(v % CH);
    const float* xp = x + (i - (((long long)zz * h + yy) * wd + xx));
    const float* wp = w + c * 27;
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const int z = zz + t / 9 - 1, yq = yy + (t / 3) % 3 - 1, xq = xx + t % 3 - 1;
      if (z >= 0 && z < d && yq >= 0 && yq < h && xq >= 0 && xq < wd)
        s = fmaf(xp[((long long)z * h + yq) * wd + xq], wp[flip ? 26 - t : t], s);
    }
    y[i] = s;
  }
}
ICL_API int icl_dwconv3d(const float* x, const float* w, float* y, int NB, int CH, int d, int h, int wd, int flip, void* stream) {
  dwconv3d_k<<<grid_for((long long)NB * CH * d * h * wd, 256), 256, 0, as_stream(stream)>>>(x, w, y, NB, CH, d, h, wd, flip);
  ICL_LAUNCHED("dwconv3d");
}
// ------------------------------------------------------------------------------------------
// Reductions over all (sample, voxel) positions of a planar [NB, CH, S] map.  At K = 16 classes the ICL heads run on
// NB = B*K = 32 maps, so "one block per channel" (CH = 4 ... 16 blocks on 148 SMs) is latency-bound by two orders of
// magnitude: the position axis is cut into up to RED_CHUNKS chunks (grid = chunks x channels), every block writes its
// partial sums to a workspace and a small second kernel adds the partials in a fixed order (deterministic).
// `ws` arguments: device scratch of at least ICL_RED_WS_BYTES bytes (icl_reduce_workspace_bytes()).
// ------------------------------------------------------------------------------------------
ICL_API int icl_reduce_workspace_bytes(void) { return RED_WS_BYTES; }

static inline int red_chunks(long long n, int per_chunk_min) {
  long long c = (n + per_chunk_min - 1) / per_chunk_min;
  if (c < 1) c = 1;
  if (c > RED_CHUNKS) c = RED_CHUNKS;
  return (int)c;
}

// dw[c][t] = sum_{nb, v} x[nb,c,v+t-1] * dy[nb,c,v].  Block = (chunk of positions, channel); a thread keeps the 27 tap sums of its
// positions in registers (dy read once, the 27 x neighbours from L1), then warp / block reduction -> partial[c][chunk][27].
__global__ void __launch_bounds__(256) dwconv3d_wgrad_partial_k(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ part,
                                                                int NB, int CH, int d, int h, int wd, int chunks) {
  __shared__ float red[8][27];
  const int c = blockIdx.y, chunk = blockIdx.x;
  const long long S = (long long)d * h * wd, n = (long long)NB * S;
  const long long i0 = n * chunk / chunks, i1 = n * (chunk + 1) / chunks;
  float acc[27];
#pragma unroll
  for (int t = 0; t < 27; ++t) acc[t] = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    const long long nb = i / S; long long v = i - nb * S;
    const int xx = (int)(v % wd); v /= wd;
    const int yy = (int)(v % h); const int zz = (int)(v / h);
    const long long base = (nb * CH + c) * S;
    const float g = dy[base + ((long long)zz * h + yy) * wd + xx];
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const int z = zz + t / 9 - 1, yq = yy + (t / 3) % 3 - 1, xq = xx + t % 3 - 1;
      if (z >= 0 && z < d && yq >= 0 && yq < h && xq >= 0 && xq < wd) acc[t] = fmaf(x[base + ((long long)z * h + yq) * wd + xq], g, acc[t]);
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < 27; ++t) {
    const float v = warp_sum(acc[t]);
    if (lane == 0) red[wid][t] = v;
  }
  __syncthreads();
  if (threadIdx.x < 27) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    part[((long long)c * chunks + chunk) * 27 + threadIdx.x] = v;
  }
}
// out[j] = sum_chunk part[(j / width) * chunks * width + chunk * width + j % width]   (one thread per output, fixed order)
__global__ void reduce_chunks_k(const float* __restrict__ part, float* __restrict__ out, int groups, int width, int chunks) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= groups * width) return;
  const int gidx = j / width, e = j % width;
  const float* p = part + (long long)gidx * chunks * width + e;
  float a = 0.f;
  for (int cidx = 0; cidx < chunks; ++cidx) a += p[(long long)cidx * width];
  out[j] = a;
}
ICL_API int icl_dwconv3d_wgrad(const float* x, const float* dy, float* dw, int NB, int CH, int d, int h, int wd, float* ws, void* stream) {
  ICL_REQUIRE(ws != nullptr, "dwconv3d_wgrad: workspace of icl_reduce_workspace_bytes() bytes required");
  const int chunks = red_chunks((long long)NB * d * h * wd, 2048);
  ICL_REQUIRE((long long)CH * chunks * 27 * 4 <= RED_WS_BYTES, "dwconv3d_wgrad: CH=%d too large for the reduction workspace", CH);
  dwconv3d_wgrad_partial_k<<<dim3(chunks, CH), 256, 0, as_stream(stream)>>>(x, dy, ws, NB, CH, d, h, wd, chunks);
  icl_count_launch(1);
  reduce_chunks_k<<<cdiv(CH * 27, 128), 128, 0, as_stream(stream)>>>(ws, dw, CH, 27, chunks);
  ICL_LAUNCHED("dwconv3d_wgrad");
}

// ------------------------------------------------------------------------------------------
// BatchNorm3d in training mode + ReLU on planar [NB, CH, S]: batch statistics per channel over
// NB*S (biased var for normalisation, unbiased for the running update, momentum 0.1).
// ------------------------------------------------------------------------------------------
// partial sums of (a, b) over a chunk of positions of channel c: MODE 0: (x, x^2); MODE 1: (g, g * xhat) with g = dy * [y > 0]
template <int MODE>
__global__ void __launch_bounds__(256) bn_partial_k(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ y,
                                                    const float* __restrict__ mean_rstd, double* __restrict__ part, int NB, int CH, long long S,
                                                    int chunks) {
  __shared__ double red[66];
  const int c = blockIdx.y, chunk = blockIdx.x;
  const long long n = (long long)NB * S;
  const long long i0 = n * chunk / chunks, i1 = n * (chunk + 1) / chunks;
  float m = 0.f, r = 1.f;
  if (MODE == 1) { m = mean_rstd[2 * c]; r = mean_rstd[2 * c + 1]; }
  double a = 0.0, q = 0.0;
  for (long long nb = i0 / S; nb * S < i1; ++nb) {
    const long long lo = i0 > nb * S ? i0 : nb * S, hi = i1 < (nb + 1) * S ? i1 : (nb + 1) * S;
    const long long base = (nb * CH + c) * S - nb * S;
    float fa = 0.f, fq = 0.f;
    for (long long i = lo + threadIdx.x; i < hi; i += 256) {
      if (MODE == 0) {
        const float v = x[base + i];
        fa += v; fq = fmaf(v, v, fq);
      } else {
        const float g = y[base + i] > 0.f ? dy[base + i] : 0.f;
        fa += g; fq = fmaf(g, (x[base + i] - m) * r, fq);
      }
    }
    a += (double)fa; q += (double)fq;
  }
  a = block_sum_d(a, red); q = block_sum_d(q, red + 33);
  if (threadIdx.x == 0) { part[((long long)c * chunks + chunk) * 2] = a; part[((long long)c * chunks + chunk) * 2 + 1] = q; }
}
// one warp per channel adds the chunk partials in a fixed order
__global__ void bn_stats_finalize_k(const double* __restrict__ part, int chunks, float* __restrict__ mean_rstd, float* __restrict__ run_mean,
                                    float* __restrict__ run_var, double n, float eps, float momentum) {
  const int c = blockIdx.x, lane = threadIdx.x;
  double a = 0.0, q = 0.0;
  for (int i = lane; i < chunks; i += 32) { a += part[((long long)c * chunks + i) * 2]; q += part[((long long)c * chunks + i) * 2 + 1]; }
  a = warp_sum_d(a); q = warp_sum_d(q);
  if (lane == 0) {
    const double m = a / n;
    double var = q / n - m * m;
    if (var < 0) var = 0;
    mean_rstd[2 * c] = (float)m;
    mean_rstd[2 * c + 1] = (float)(1.0 / sqrt(var + (double)eps));
    if (run_mean) {
      const double unb = n > 1 ? var * n / (n - 1) : var;
      run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * (float)m;
      run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)unb;
    }
  }
}
__global__ void bn_bwd_finalize_k(const double* __restrict__ part, int chunks, float* __restrict__ sums) {
  const int c = blockIdx.x, lane = threadIdx.x;
  double a = 0.0, q = 0.0;
  for (int i = lane; i < chunks; i += 32) { a += part[((long long)c * chunks + i) * 2]; q += part[((long long)c * chunks + i) * 2 + 1]; }
  a = warp_sum_d(a); q = warp_sum_d(q);
  if (lane == 0) { sums[c] = (float)a; sums[gridDim.x + c] = (float)q; }   // [2][CH]: row 0 = dbeta, row 1 = dgamma
}
__global__ void bn_relu_apply_k(const float* __restrict__ x, const float* __restrict__ mean_rstd, const float* __restrict__ g,
                                const float* __restrict__ b, float* __restrict__ y, int CH, long long S, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i / S) % CH);
    y[i] = fmaxf((x[i] - mean_rstd[2 * c]) * mean_rstd[2 * c + 1] * g[c] + b[c], 0.f);
  }
}
ICL_API int icl_bn_relu_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean_rstd, float* run_mean, float* run_var,
                            int NB, int CH, long long S, float eps, float momentum, double* ws, void* stream) {
  ICL_REQUIRE(ws != nullptr, "bn_relu_fwd: workspace of icl_reduce_workspace_bytes() bytes required");
  const int chunks = red_chunks((long long)NB * S, 4096);
  ICL_REQUIRE((long long)CH * chunks * 16 <= RED_WS_BYTES, "bn_relu_fwd: CH=%d too large for the reduction workspace", CH);
  bn_partial_k<0><<<dim3(chunks, CH), 256, 0, as_stream(stream)>>>(x, nullptr, nullptr, nullptr, ws, NB, CH, S, chunks);
  icl_count_launch(1);
  bn_stats_finalize_k<<<CH, 32, 0, as_stream(stream)>>>(ws, chunks, mean_rstd, run_mean, run_var, (double)NB * (double)S, eps, momentum);
  icl_count_launch(1);
  const long long total = (long long)NB * CH * S;
  bn_relu_apply_k<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(x, mean_rstd, gamma, beta, y, CH, S, total);
  ICL_LAUNCHED("bn_relu_fwd");
}
// backward: g = dy * [y > 0];  dgamma = sum g*xh; dbeta = sum g; dx = gamma*rstd*(g - mean(g) - xh*mean(g*xh))
__global__ void bn_relu_bwd_apply_k(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                                    const float* __restrict__ mean_rstd, const float* __restrict__ g, const float* __restrict__ sums,
                                    float* __restrict__ dx, int CH, long long S, long long total, float inv_n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i / S) % CH);
    const float m = mean_rstd[2 * c], r = mean_rstd[2 * c + 1];
    const float gg = y[i] > 0.f ? dy[i] : 0.f;
    const float xh = (x[i] - m) * r;
    dx[i] = g[c] * r * (gg - sums[c] * inv_n - xh * sums[CH + c] * inv_n);
  }
}
ICL_API int icl_bn_relu_bwd(const float* dy, const float* x, const float* y, const float* mean_rstd, const float* gamma, float* sums /*[2][CH]: dbeta row, dgamma row*/,
                            float* dx, int NB, int CH, long long S, double* ws, void* stream) {
  ICL_REQUIRE(ws != nullptr, "bn_relu_bwd: workspace of icl_reduce_workspace_bytes() bytes required");
  const int chunks = red_chunks((long long)NB * S, 4096);
  ICL_REQUIRE((long long)CH * chunks * 16 <= RED_WS_BYTES, "bn_relu_bwd: CH=%d too large for the reduction workspace", CH);
  bn_partial_k<1><<<dim3(chunks, CH), 256, 0, as_stream(stream)>>>(x, dy, y, mean_rstd, ws, NB, CH, S, chunks);
  icl_count_launch(1);
  bn_bwd_finalize_k<<<CH, 32, 0, as_stream(stream)>>>(ws, chunks, sums);
  icl_count_launch(1);
  const long long total = (long long)NB * CH * S;
  bn_relu_bwd_apply_k<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(dy, x, y, mean_rstd, gamma, sums, dx, CH, S, total,
                                                                         1.f / (float)((long long)NB * S));
  ICL_LAUNCHED("bn_relu_bwd");
}

// pointwise (1x1x1) weight gradient on planar maps: dw[o][i] = sum_{nb, s} dy[nb,o,s] * x[nb,i,s]; db[o] = sum dy.
// Block = chunk of positions; the dy / x values of 256 positions are staged in shared memory, then thread (o, i) adds its products
// (CO * (CI + 1) <= 272 pairs, threads stride over them) -> partial[chunk][CO * (CI + 1)]; reduce_chunks_k finishes.
#define PW_MAXPAIRS 1024
__global__ void __launch_bounds__(256) planar_pw_wgrad_partial_k(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ part,
                                                                 int NB, int CO, int CI, long long S, int chunks, int T) {
  extern __shared__ float sm[];  // [CO][T + 1] dy, then [CI][T + 1] x (odd row stride: the pair threads of a warp read different rows)
  const int TS = T + 1;
  float* sdy = sm;
  float* sx = sm + CO * TS;
  const int chunk = blockIdx.x;
  const long long n = (long long)NB * S;
  const long long i0 = n * chunk / chunks, i1 = n * (chunk + 1) / chunks;
  const int pairs = CO * (CI + 1);
  float acc[PW_MAXPAIRS / 256];   // pairs threadIdx.x + 256 * k
#pragma unroll
  for (int k = 0; k < PW_MAXPAIRS / 256; ++k) acc[k] = 0.f;
  for (long long t0 = i0; t0 < i1; t0 += T) {
    for (int j = threadIdx.x; j < T; j += 256) {
      const long long p = t0 + j;
      const bool ok = p < i1;
      const long long nb = ok ? p / S : 0, sp = ok ? p - nb * S : 0;
      for (int o = 0; o < CO; ++o) sdy[o * TS + j] = ok ? dy[(nb * CO + o) * S + sp] : 0.f;
      for (int i = 0; i < CI; ++i) sx[i * TS + j] = ok ? x[(nb * CI + i) * S + sp] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PW_MAXPAIRS / 256; ++k) {
      const int pr = threadIdx.x + 256 * k;
      if (pr < pairs) {
        const int o = pr / (CI + 1), i = pr % (CI + 1);
        const float* a = sdy + o * TS;
        float s = 0.f;
        if (i < CI) {
          const float* b = sx + i * TS;
#pragma unroll 8
          for (int t = 0; t < T; ++t) s = fmaf(a[t], b[t], s);
        } else {
#pragma unroll 8
          for (int t = 0; t < T; ++t) s += a[t];
        }
        acc[k] += s;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < PW_MAXPAIRS / 256; ++k) {
    const int pr = threadIdx.x + 256 * k;
    if (pr < pairs) part[(long long)chunk * pairs + pr] = acc[k];
  }
}
// One warp per (output, input) pair: lane l sums chunks l, l + 32, ... with four loads in flight, then a shuffle tree — a fixed
// order, so the result is reproducible.  (One thread per pair walking all chunks serially took 14-31 us: a chain of dependent L2
// round trips over up to 1 728 chunks.)
__global__ void __launch_bounds__(256) planar_pw_finalize_k(const float* __restrict__ part, int chunks, int CO, int CI, float* __restrict__ dw,
                                                            float* __restrict__ db) {
  const int pr = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, pairs = CO * (CI + 1);
  if (pr >= pairs) return;
  const float* src = part + pr;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int c = lane;
  for (; c + 96 < chunks; c += 128) {
    const float v0 = src[(long long)c * pairs], v1 = src[(long long)(c + 32) * pairs], v2 = src[(long long)(c + 64) * pairs],
                v3 = src[(long long)(c + 96) * pairs];
    a0 += v0; a1 += v1; a2 += v2; a3 += v3;
  }
  for (; c < chunks; c += 32) a0 += src[(long long)c * pairs];
  const float a = warp_sum((a0 + a1) + (a2 + a3));
  if (lane == 0) {
    const int o = pr / (CI + 1), i = pr % (CI + 1);
    if (i < CI) dw[o * CI + i] = a;
    else if (db) db[o] = a;
  }
}
ICL_API int icl_planar_pw_wgrad(const float* dy, const float* x, float* dw, float* db, int NB, int CO, int CI, long long S, float* ws, void* stream) {
  ICL_REQUIRE(ws != nullptr, "planar_pw_wgrad: workspace of icl_reduce_workspace_bytes() bytes required");
  ICL_REQUIRE(CO >= 1 && CI >= 1 && CO * (CI + 1) <= PW_MAXPAIRS, "planar_pw_wgrad: CO=%d CI=%d not supported (CO * (CI + 1) <= %d)", CO, CI, PW_MAXPAIRS);
  int T = 256;
  while (T > 32 && (size_t)(CO + CI) * (T + 1) * 4 > 40 * 1024) T >>= 1;
  ICL_REQUIRE((size_t)(CO + CI) * (T + 1) * 4 <= 48 * 1024, "planar_pw_wgrad: CO=%d CI=%d not supported (shared-memory staging)", CO, CI);
  int chunks = red_chunks((long long)NB * S, T);   // one staged tile per block while blocks are scarce
  while (chunks > 1 && (long long)chunks * CO * (CI + 1) * 4 > RED_WS_BYTES) chunks >>= 1;
  planar_pw_wgrad_partial_k<<<chunks, 256, (size_t)(CO + CI) * (T + 1) * 4, as_stream(stream)>>>(dy, x, ws, NB, CO, CI, S, chunks, T);
  icl_count_launch(1);
  planar_pw_finalize_k<<<cdiv(CO * (CI + 1), 8), 256, 0, as_stream(stream)>>>(ws, chunks, CO, CI, dw, db);
  ICL_LAUNCHED("planar_pw_wgrad");
}

// pointwise (1x1x1) conv on planar maps, forward and data gradient:  y[nb][o][s] = sum_i W[o*w_so + i*w_si] * x[nb][i][s] (+ bias[o]).
// (data gradient = the same kernel with the weight strides swapped.)  One thread per (nb, s) position: CI coalesced loads, CO
// coalesced stores, the CO x CI weights in shared memory — the maps stream through once (CI, CO <= 16).
__global__ void __launch_bounds__(256) planar_pw_k(const float* __restrict__ x, const float* __restrict__ w, int w_so, int w_si,
                                                   const float* __restrict__ bias, float* __restrict__ y, int NB, int CI, int CO, long long S) {
  __shared__ float ws[16 * 16 + 16];
  for (int i = threadIdx.x; i < CO * CI; i += 256) ws[i] = w[(i / CI) * w_so + (i % CI) * w_si];
  for (int i = threadIdx.x; i < CO; i += 256) ws[256 + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const long long n = (long long)NB * S;
  for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < n; p += (long long)gridDim.x * 256) {
    const long long nb = p / S, sp = p - nb * S;
    float xv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) xv[i] = i < CI ? x[(nb * CI + i) * S + sp] : 0.f;
    for (int o = 0; o < CO; ++o) {
      float a = ws[256 + o];
#pragma unroll
      for (int i = 0; i < 16; ++i) if (i < CI) a = fmaf(ws[o * CI + i], xv[i], a);
      y[(nb * CO + o) * S + sp] = a;
    }
  }
}
ICL_API int icl_planar_pw(const float* x, const float* w, int w_so, int w_si, const float* bias, float* y, int NB, int CI, int CO, long long S,
                          void* stream) {
  ICL_REQUIRE(CI >= 1 && CI <= 16 && CO >= 1 && CO <= 16, "planar_pw: CI=%d CO=%d unsupported (<= 16)", CI, CO);
  planar_pw_k<<<grid_for((long long)NB * S, 256, 148 * 8), 256, 0, as_stream(stream)>>>(x, w, w_so, w_si, bias, y, NB, CI, CO, S);
  ICL_LAUNCHED("planar_pw");
}
