// Bandwidth-bound activation-side kernels of the 3D U-Net backbone:
// InstanceNorm3d+ReLU apply / backward, MaxPool3d(2), trilinear x2 upsample, Dropout, and the
// fp32 -> split-bf16 operand packing that feeds the tcgen05 convolution.
//
// Layouts:
//   F32CL : float  [B][D][H][W][C]                      (torch channels_last_3d of [B,C,D,H,W])
//   PK    : bf16   [P][B][C/8][D][H][W][8], P=2 planes  (hi, lo) with x ~= hi + lo
// All kernels are pure streaming: 128/256-bit accesses, voxel-fastest thread mapping so that both
// the F32CL side (32 B sectors) and the PK side (16 B per voxel, contiguous along W) coalesce.
#include "common.cuh"

struct F8 { float4 a, b; };
__device__ __forceinline__ F8 ld8(const float* p) {
  F8 r; r.a = *reinterpret_cast<const float4*>(p); r.b = *reinterpret_cast<const float4*>(p + 4); return r;
}
__device__ __forceinline__ void st8(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void unpack8(const F8& f, float* v) {
  v[0] = f.a.x; v[1] = f.a.y; v[2] = f.a.z; v[3] = f.a.w; v[4] = f.b.x; v[5] = f.b.y; v[6] = f.b.z; v[7] = f.b.w;
}
// writes 8 values as split bf16 into the hi plane (and lo plane when lo != nullptr)
__device__ __forceinline__ uint2 pack4_bf16(__nv_bfloat16 a, __nv_bfloat16 b, __nv_bfloat16 c, __nv_bfloat16 d) {
  const __nv_bfloat162 p0 = __halves2bfloat162(a, b), p1 = __halves2bfloat162(c, d);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
}
__device__ __forceinline__ void st_pk8(__nv_bfloat16* hi, __nv_bfloat16* lo, const float* v) {
  __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split_bf16(v[i], h[i], l[i]);
  *reinterpret_cast<uint4*>(hi) = *reinterpret_cast<const uint4*>(h);
  if (lo) *reinterpret_cast<uint4*>(lo) = *reinterpret_cast<const uint4*>(l);
}

// ------------------------------------------------------------------------------------------
// InstanceNorm statistics finalize: (sum, sumsq) in double -> (mean, rstd) in float.
// ------------------------------------------------------------------------------------------
__global__ void instnorm_finalize_k(const double* __restrict__ stats, float* __restrict__ mr, int n, double inv_count, float eps) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double m = stats[2 * i] * inv_count;
  double var = stats[2 * i + 1] * inv_count - m * m;
  if (var < 0) var = 0;
  mr[2 * i] = (float)m;
  mr[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

ICL_API int icl_instnorm_finalize(const double* stats, float* mr, int B, int C, long long S, float eps, void* stream) {
  int n = B * C;
  instnorm_finalize_k<<<cdiv(n, 128), 128, 0, as_stream(stream)>>>(stats, mr, n, 1.0 / (double)S, eps);
  ICL_LAUNCHED("instnorm_finalize");
}

// Plain per-(b,c) sum / sumsq over S for tensors whose producer did not emit statistics.
__global__ void instnorm_stats_k(const float* __restrict__ y, double* __restrict__ stats, int C, long long S, int chunks) {
  // grid: (chunks, B); each block reduces a slab of voxels for all C channels (C <= 1024).
  __shared__ double red[66];
  const int b = blockIdx.y;
  const long long per = (S + chunks - 1) / chunks;
  const long long s0 = (long long)blockIdx.x * per, s1 = min(S, s0 + per);
  for (int c = 0; c < C; ++c) {
    double a = 0, q = 0;
    for (long long s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
      float v = y[((long long)b * S + s) * C + c];
      a += v; q += (double)v * v;
    }
    a = block_sum_d(a, red);
    q = block_sum_d(q, red + 33);
    if (threadIdx.x == 0) {
      atomicAdd(&stats[((long long)b * C + c) * 2], a);
      atomicAdd(&stats[((long long)b * C + c) * 2 + 1], q);
    }
  }
}
ICL_API int icl_instnorm_stats(const float* y, double* stats, int B, int C, long long S, void* stream) {
  int chunks = (int)min((long long)64, (S + 4095) / 4096);
  if (chunks < 1) chunks = 1;
  instnorm_stats_k<<<dim3(chunks, B), 256, 0, as_stream(stream)>>>(y, stats, C, S, chunks);
  ICL_LAUNCHED("instnorm_stats");
}

// ------------------------------------------------------------------------------------------
// InstanceNorm + ReLU apply.  a = relu((y - mean) * rstd); optional PK output.
// grid: (chunks over S, B*C8); block 256.
// ------------------------------------------------------------------------------------------
// gamma / beta (optional, per channel) and slope generalise InstanceNorm+ReLU to BatchNorm2d(affine)+LeakyReLU: with the
// images of a 2D batch stacked along D of ONE sample, per-(sample, channel) statistics are exactly BatchNorm's batch
// statistics (networks/unet_icl.py:46-54).  a = act(gamma * yh + beta), act(v) = v > 0 ? v : slope * v.
__global__ void instnorm_relu_fwd_k(const float* __restrict__ y, const float* __restrict__ mr, float* __restrict__ a,
                                    __nv_bfloat16* __restrict__ pk, int write_lo, int B, int C, long long S,
                                    const float* __restrict__ gamma, const float* __restrict__ beta, float slope) {
  const int C8 = C >> 3;
  const int b = blockIdx.y / C8, c8 = blockIdx.y % C8;
  float mean[8], rstd[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mean[i] = mr[((long long)b * C + c8 * 8 + i) * 2];
    rstd[i] = mr[((long long)b * C + c8 * 8 + i) * 2 + 1] * (gamma ? gamma[c8 * 8 + i] : 1.f);
    sh[i] = beta ? beta[c8 * 8 + i] : 0.f;
  }
  const long long plane = (long long)B * C8 * S * 8;
  __nv_bfloat16* hi = pk ? pk + ((long long)b * C8 + c8) * S * 8 : nullptr;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (long long)gridDim.x * blockDim.x) {
    const long long off = ((long long)b * S + s) * C + c8 * 8;
    float v[8];
    unpack8(ld8(y + off), v);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float t = (v[i] - mean[i]) * rstd[i] + sh[i]; v[i] = t > 0.f ? t : slope * t; }
    if (a) st8(a + off, v);
    if (hi) st_pk8(hi + s * 8, write_lo ? hi + plane + s * 8 : nullptr, v);
  }
}
__global__ void instnorm_relu_fwd_generic_k(const float* __restrict__ y, const float* __restrict__ mr, float* __restrict__ a,
                                            int C, long long S, long long total, const float* __restrict__ gamma,
                                            const float* __restrict__ beta, float slope) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long b = i / (S * C);
    const float m = mr[(b * C + c) * 2], r = mr[(b * C + c) * 2 + 1];
    const float t = (y[i] - m) * r * (gamma ? gamma[c] : 1.f) + (beta ? beta[c] : 0.f);
    a[i] = t > 0.f ? t : slope * t;
  }
}
ICL_API int icl_normact_fwd(const float* y, const float* mr, const float* gamma, const float* beta, float slope, float* a, void* pk, int write_lo,
                            int B, int C, long long S, void* stream) {
  if (C % 8 == 0) {
    int gx = (int)min((long long)cdiv(S, 256), (long long)max(1, 148 * 8 / (B * (C / 8)) + 1));
    instnorm_relu_fwd_k<<<dim3(gx, B * (C / 8)), 256, 0, as_stream(stream)>>>(y, mr, a, (__nv_bfloat16*)pk, write_lo, B, C, S, gamma, beta, slope);
  } else {
    ICL_REQUIRE(pk == nullptr && a != nullptr, "normact_fwd: PK output needs C %% 8 == 0 (C=%d)", C);
    long long total = (long long)B * S * C;
    instnorm_relu_fwd_generic_k<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(y, mr, a, C, S, total, gamma, beta, slope);
  }
  ICL_LAUNCHED("normact_fwd");
}
ICL_API int icl_instnorm_relu_fwd(const float* y, const float* mr, float* a, void* pk, int write_lo, int B, int C, long long S,
                                  void* stream) {
  return icl_normact_fwd(y, mr, nullptr, nullptr, 0.f, a, pk, write_lo, B, C, S, stream);
}

// ------------------------------------------------------------------------------------------
// InstanceNorm + ReLU backward:  yh = (y-mean)*rstd, g = dA * [yh > 0],
//   dY = rstd * (g - mean_S(g) - yh * mean_S(g*yh)).
// Pass 1 reduces (sum g, sum g*yh) per (b,c); pass 2 applies and optionally emits dY as PK and
// accumulates sum_S dY (the conv-bias gradient; mathematically ~0 for InstanceNorm).
// ------------------------------------------------------------------------------------------
// With affine / leaky activation:  pre = gamma * yh + beta,  g' = dA * act'(pre);  the two sums reduced here are
// sum g' (= dbeta) and sum g' * yh (= dgamma);  dY = rstd * gamma * (g' - mean(g') - yh * mean(g' * yh)).
__device__ __forceinline__ float act_grad(float yh, float gm, float bt, float slope) { return (gm * yh + bt) > 0.f ? 1.f : slope; }
__global__ void instnorm_relu_bwd_reduce_k(const float* __restrict__ dA, const float* __restrict__ y, const float* __restrict__ mr,
                                           double* __restrict__ red, int C, long long S, int chunks, const float* __restrict__ gamma,
                                           const float* __restrict__ beta, float slope) {
  // grid (chunks, B); thread t owns channel group: channels handled as c = t % C lanes when C <= blockDim
  // Generic mapping: each thread walks elements i = s*C + c with stride blockDim (C divides blockDim or not).
  extern __shared__ double sm[];  // [2][C]
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.0;
  __syncthreads();
  const long long per = (S + chunks - 1) / chunks;
  const long long s0 = (long long)blockIdx.x * per, s1 = min(S, s0 + per);
  const long long n = (s1 - s0) * C;
  const float* dAb = dA + ((long long)b * S + s0) * C;
  const float* yb = y + ((long long)b * S + s0) * C;
  // C % 4 == 0 and (4 * blockDim) % C == 0: every thread keeps a fixed group of 4 channels and streams float4s.
  if (C % 4 == 0 && (4 * blockDim.x) % C == 0) {
    const int c = (threadIdx.x * 4) % C;
    float m[4], r[4], gm[4], bt[4], sg[4] = {0.f, 0.f, 0.f, 0.f}, sgy[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      m[j] = mr[((long long)b * C + c + j) * 2]; r[j] = mr[((long long)b * C + c + j) * 2 + 1];
      gm[j] = gamma ? gamma[c + j] : 1.f; bt[j] = beta ? beta[c + j] : 0.f;
    }
    const long long n4 = n >> 2;
    const float4* y4 = reinterpret_cast<const float4*>(yb);
    const float4* d4 = reinterpret_cast<const float4*>(dAb);
#pragma unroll 4
    for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 yv = y4[i], dv = d4[i];
      const float ya[4] = {yv.x, yv.y, yv.z, yv.w}, da[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float yh = (ya[j] - m[j]) * r[j];
        const float g = da[j] * act_grad(yh, gm[j], bt[j], slope);
        sg[j] += g; sgy[j] += g * yh;
      }
    }
    // lanes l and l + C/4 own the same channels: combine them by shuffles first (the shared double atomics are CAS loops)
    const int P4 = C >> 2;
    for (int o = 16; o >= P4; o >>= 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { sg[j] += __shfl_xor_sync(0xffffffffu, sg[j], o); sgy[j] += __shfl_xor_sync(0xffffffffu, sgy[j], o); }
    }
    if ((int)(threadIdx.x & 31) < P4 || P4 >= 32) {
      // P4 <= 8 (C <= 32): the few surviving lanes of a warp hit distinct addresses, and the 8 warps collide at most 8-way
#pragma unroll
      for (int j = 0; j < 4; ++j) { atomicAdd(&sm[c + j], (double)sg[j]); atomicAdd(&sm[C + c + j], (double)sgy[j]); }
    }
  } else if (blockDim.x % C == 0) {
    const int c = threadIdx.x % C;
    const float m = mr[((long long)b * C + c) * 2], r = mr[((long long)b * C + c) * 2 + 1];
    const float gm = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
    float sg = 0.f, sgy = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const float yh = (yb[i] - m) * r;
      const float g = dAb[i] * act_grad(yh, gm, bt, slope);
      sg += g; sgy += g * yh;
    }
    atomicAdd(&sm[c], (double)sg);
    atomicAdd(&sm[C + c], (double)sgy);
  } else {
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const int c = (int)(i % C);
      const float m = mr[((long long)b * C + c) * 2], r = mr[((long long)b * C + c) * 2 + 1];
      const float yh = (yb[i] - m) * r;
      const float g = dAb[i] * act_grad(yh, gamma ? gamma[c] : 1.f, beta ? beta[c] : 0.f, slope);
      atomicAdd(&sm[c], (double)g);
      atomicAdd(&sm[C + c], (double)(g * yh));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(&red[((long long)b * C + i) * 2], sm[i]);
    atomicAdd(&red[((long long)b * C + i) * 2 + 1], sm[C + i]);
  }
}

__global__ void instnorm_relu_bwd_apply_k(const float* __restrict__ dA, const float* __restrict__ y, const float* __restrict__ mr,
                                          const double* __restrict__ red, float* __restrict__ dY, __nv_bfloat16* __restrict__ pk,
                                          int write_lo, int B, int C, long long S, float* __restrict__ dbias,
                                          const float* __restrict__ gamma, const float* __restrict__ beta, float slope) {
  __shared__ float bsum[8][8];
  float bacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int C8 = C >> 3;
  const int b = blockIdx.y / C8, c8 = blockIdx.y % C8;
  float mean[8], rstd[8], mg[8], mgy[8], gm[8], bt[8];
  const float invS = 1.f / (float)S;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long k = (long long)b * C + c8 * 8 + i;
    mean[i] = mr[k * 2]; rstd[i] = mr[k * 2 + 1];
    mg[i] = (float)(red[k * 2] / (double)S); mgy[i] = (float)(red[k * 2 + 1] / (double)S);
    gm[i] = gamma ? gamma[c8 * 8 + i] : 1.f; bt[i] = beta ? beta[c8 * 8 + i] : 0.f;
  }
  (void)invS;
  const long long plane = (long long)B * C8 * S * 8;
  __nv_bfloat16* hi = pk ? pk + ((long long)b * C8 + c8) * S * 8 : nullptr;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (long long)gridDim.x * blockDim.x) {
    const long long off = ((long long)b * S + s) * C + c8 * 8;
    float yv[8], gv[8], o[8];
    unpack8(ld8(y + off), yv);
    unpack8(ld8(dA + off), gv);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float yh = (yv[i] - mean[i]) * rstd[i];
      const float g = gv[i] * act_grad(yh, gm[i], bt[i], slope);
      o[i] = rstd[i] * gm[i] * (g - mg[i] - yh * mgy[i]);
      bacc[i] += o[i];
    }
    if (dY) st8(dY + off, o);
    if (hi) st_pk8(hi + s * 8, write_lo ? hi + plane + s * 8 : nullptr, o);
  }
  if (dbias) {  // conv-bias gradient = sum over voxels (and samples) of dY: warp -> block -> one atomic per channel
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float t = warp_sum(bacc[i]);
      if (lane == 0) bsum[wid][i] = t;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += bsum[w][threadIdx.x];
      atomicAdd(dbias + c8 * 8 + threadIdx.x, t);
    }
  }
}
// Same arithmetic with the 8-channel group as the FASTEST thread index (C/8 a power of two <= 32): consecutive lanes read
// consecutive 32-byte segments of a voxel, so every DRAM line is fetched once.  With one CTA per channel group (kernel above)
// the two groups of a C = 16 layer run in different waves and each 64-byte line of dA / y was read from DRAM twice
// (ncu r01p: 448 MB read for 226 MB of operands).
__global__ void __launch_bounds__(256, 3) instnorm_relu_bwd_apply_cl_k(
    const float* __restrict__ dA, const float* __restrict__ y, const float* __restrict__ mr, const double* __restrict__ red,
    float* __restrict__ dY, __nv_bfloat16* __restrict__ pk, int write_lo, int B, int C, long long S, float* __restrict__ dbias,
    const float* __restrict__ gamma, const float* __restrict__ beta, float slope) {
  __shared__ float bsum[8][32][8];
  const int C8 = C >> 3;
  const int b = blockIdx.y, c8 = threadIdx.x & (C8 - 1);
  float mean[8], rstd[8], mg[8], mgy[8], gm[8], bt[8], bacc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long k = (long long)b * C + c8 * 8 + i;
    mean[i] = mr[k * 2]; rstd[i] = mr[k * 2 + 1];
    mg[i] = (float)(red[k * 2] / (double)S); mgy[i] = (float)(red[k * 2 + 1] / (double)S);
    gm[i] = gamma ? gamma[c8 * 8 + i] : 1.f; bt[i] = beta ? beta[c8 * 8 + i] : 0.f;
    bacc[i] = 0.f;
  }
  const long long plane = (long long)B * C8 * S * 8;
  __nv_bfloat16* hi = pk ? pk + ((long long)b * C8 + c8) * S * 8 : nullptr;
  const long long total = S * C8;  // (voxel, group) pairs of this sample; blockDim and the grid stride are multiples of C8
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / C8;
    const long long off = ((long long)b * S + s) * C + c8 * 8;
    float yv[8], gv[8], o[8];
    unpack8(ld8(y + off), yv);
    unpack8(ld8(dA + off), gv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float yh = (yv[j] - mean[j]) * rstd[j];
      const float g = gv[j] * act_grad(yh, gm[j], bt[j], slope);
      o[j] = rstd[j] * gm[j] * (g - mg[j] - yh * mgy[j]);
      bacc[j] += o[j];
    }
    if (dY) st8(dY + off, o);
    if (hi) st_pk8(hi + s * 8, write_lo ? hi + plane + s * 8 : nullptr, o);
  }
  if (dbias) {  // lanes l and l + C8 own the same channels: fold them, then warps -> block -> one atomic per channel
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = bacc[j];
      for (int o2 = 16; o2 >= C8; o2 >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o2);
      if (lane < C8) bsum[wid][lane][j] = t;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C8 * 8; i += blockDim.x) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += bsum[w][i >> 3][i & 7];
      atomicAdd(dbias + i, t);
    }
  }
}
__global__ void instnorm_relu_bwd_apply_generic_k(const float* __restrict__ dA, const float* __restrict__ y, const float* __restrict__ mr,
                                                  const double* __restrict__ red, float* __restrict__ dY, int C, long long S, long long total,
                                                  const float* __restrict__ gamma, const float* __restrict__ beta, float slope) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long k = (i / (S * C)) * C + c;
    const float m = mr[k * 2], r = mr[k * 2 + 1];
    const float mg = (float)(red[k * 2] / (double)S), mgy = (float)(red[k * 2 + 1] / (double)S);
    const float yh = (y[i] - m) * r;
    const float gmc = gamma ? gamma[c] : 1.f;
    const float g = dA[i] * act_grad(yh, gmc, beta ? beta[c] : 0.f, slope);
    dY[i] = r * gmc * (g - mg - yh * mgy);
  }
}
ICL_API int icl_normact_bwd(const float* dA, const float* y, const float* mr, const float* gamma, const float* beta, float slope,
                            double* red /* [B,C,2] zeroed; out: sum g' (= dbeta), sum g' yh (= dgamma) */, float* dY, void* pk, int write_lo,
                            float* dbias /* [C] zeroed, or null */, int B, int C, long long S, void* stream) {
  ICL_REQUIRE(C <= 1024, "normact_bwd: C=%d > 1024", C);
  // one wave: the reduce kernel holds 3 CTAs per SM (66 registers); 592 CTAs ran as 1.33 waves (ncu r01p)
  int chunks = (int)min((long long)max(1, 148 * 3 / B), (S * C + 16383) / 16384);
  if (chunks < 1) chunks = 1;
  instnorm_relu_bwd_reduce_k<<<dim3(chunks, B), 256, 2 * C * sizeof(double), as_stream(stream)>>>(dA, y, mr, red, C, S, chunks, gamma, beta, slope);
  icl_count_launch(1);
  const int C8 = C / 8;
  if (C % 8 == 0 && C8 <= 32 && (C8 & (C8 - 1)) == 0) {
    int gx = (int)min((long long)cdiv(S * C8, 256), (long long)max(1, 148 * 3 / B));
    instnorm_relu_bwd_apply_cl_k<<<dim3(gx, B), 256, 0, as_stream(stream)>>>(dA, y, mr, red, dY, (__nv_bfloat16*)pk, write_lo, B, C, S, dbias, gamma, beta, slope);
  } else if (C % 8 == 0) {
    int gx = (int)min((long long)cdiv(S, 256), (long long)max(1, 148 * 8 / (B * (C / 8)) + 1));
    instnorm_relu_bwd_apply_k<<<dim3(gx, B * (C / 8)), 256, 0, as_stream(stream)>>>(dA, y, mr, red, dY, (__nv_bfloat16*)pk, write_lo, B, C, S, dbias, gamma, beta, slope);
  } else {
    ICL_REQUIRE(pk == nullptr && dY != nullptr && dbias == nullptr, "instnorm_relu_bwd: PK / bias-gradient outputs need C %% 8 == 0 (C=%d)", C);
    long long total = (long long)B * S * C;
    instnorm_relu_bwd_apply_generic_k<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(dA, y, mr, red, dY, C, S, total, gamma, beta, slope);
  }
  ICL_LAUNCHED("normact_bwd");
}
ICL_API int icl_instnorm_relu_bwd(const float* dA, const float* y, const float* mr, double* red, float* dY, void* pk, int write_lo, float* dbias,
                                  int B, int C, long long S, void* stream) {
  return icl_normact_bwd(dA, y, mr, nullptr, nullptr, 0.f, red, dY, pk, write_lo, dbias, B, C, S, stream);
}

// ------------------------------------------------------------------------------------------
// fp32 F32CL -> PK split-bf16 operand (C % 8 == 0).
// ------------------------------------------------------------------------------------------
__global__ void pack_pk_k(const float* __restrict__ x, __nv_bfloat16* __restrict__ pk, int write_lo, int B, int C, long long S) {
  const int C8 = C >> 3;
  const int b = blockIdx.y / C8, c8 = blockIdx.y % C8;
  const long long plane = (long long)B * C8 * S * 8;
  __nv_bfloat16* hi = pk + ((long long)b * C8 + c8) * S * 8;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (long long)gridDim.x * blockDim.x) {
    float v[8];
    unpack8(ld8(x + ((long long)b * S + s) * C + c8 * 8), v);
    st_pk8(hi + s * 8, write_lo ? hi + plane + s * 8 : nullptr, v);
  }
}
ICL_API int icl_pack_pk(const float* x, void* pk, int write_lo, int B, int C, long long S, void* stream) {
  ICL_REQUIRE(C % 8 == 0, "pack_pk: C=%d not a multiple of 8", C);
  int gx = (int)min((long long)cdiv(S, 256), (long long)max(1, 148 * 8 / (B * (C / 8)) + 1));
  pack_pk_k<<<dim3(gx, B * (C / 8)), 256, 0, as_stream(stream)>>>(x, (__nv_bfloat16*)pk, write_lo, B, C, S);
  ICL_LAUNCHED("pack_pk");
}

// ------------------------------------------------------------------------------------------
// MaxPool3d(2) forward/backward.  First maximum in (d,h,w) scan order wins ties:
// strict '>' while scanning; NaN propagates like torch (val != val takes over).
// idx stores the winning position 0..7 per output element.
// ------------------------------------------------------------------------------------------
__global__ void maxpool_fwd_k(const float* __restrict__ x, float* __restrict__ out, unsigned char* __restrict__ idx,
                              __nv_bfloat16* __restrict__ pk, int write_lo, int B, int C, int D, int H, int W) {
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const long long So = (long long)Do * Ho * Wo;
  const long long total = (long long)B * So * C;
  const int C8 = (C % 8 == 0) ? C / 8 : 0;
  const long long plane = (long long)B * C8 * So * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long v = i / C;
    const int wo = (int)(v % Wo); v /= Wo;
    const int ho = (int)(v % Ho); v /= Ho;
    const int dd = (int)(v % Do);
    const int b = (int)(v / Do);
    float best = 0.f; int bi = 0;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int d = 2 * dd + (p >> 2), h = 2 * ho + ((p >> 1) & 1), w = 2 * wo + (p & 1);
      const float val = x[((((long long)b * D + d) * H + h) * W + w) * C + c];
      if (p == 0 || val > best || val != val) { best = val; bi = p; }
    }
    if (out) out[i] = best;
    idx[i] = (unsigned char)bi;
    if (pk) {
      const long long s = ((long long)dd * Ho + ho) * Wo + wo;
      __nv_bfloat16 hi, lo; split_bf16(best, hi, lo);
      const long long o = (((long long)b * C8 + (c >> 3)) * So + s) * 8 + (c & 7);
      pk[o] = hi;
      if (write_lo) pk[plane + o] = lo;
    }
  }
}
// C % 8 == 0: thread = one output voxel x 4 channels (channel quad fastest): eight 128-bit loads, 32-bit index math, one 4-byte index
// store, 8-byte PK stores (a lane pair fills one 16-byte brick row).  Same first-max / NaN rule as above.
__global__ void __launch_bounds__(256) maxpool_fwd_v4_k(const float* __restrict__ x, float* __restrict__ out, unsigned char* __restrict__ idx,
                                                        __nv_bfloat16* __restrict__ pk, int write_lo, int B, int C, int D, int H, int W) {
  const int Do = D / 2, Ho = H / 2, Wo = W / 2, C4 = C >> 2, C8 = C >> 3;
  const size_t So = (size_t)Do * Ho * Wo;
  const size_t plane = (size_t)B * C8 * So * 8;
  const unsigned total = (unsigned)B * Do * Ho * Wo * C4;   // host guarantees < 2^31
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int q = (int)(i % (unsigned)C4);
    unsigned v = i / (unsigned)C4;
    const int wo = (int)(v % (unsigned)Wo); v /= (unsigned)Wo;
    const int ho = (int)(v % (unsigned)Ho); v /= (unsigned)Ho;
    const int dd = (int)(v % (unsigned)Do);
    const int b = (int)(v / (unsigned)Do);
    const float* xb = x + ((((size_t)b * D + 2 * dd) * H + 2 * ho) * W + 2 * wo) * C + q * 4;
    float best[4]; int bi[4];
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const float4 t = *reinterpret_cast<const float4*>(xb + ((size_t)((p >> 2) * H + ((p >> 1) & 1)) * W + (p & 1)) * C);
      const float val[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (p == 0 || val[j] > best[j] || val[j] != val[j]) { best[j] = val[j]; bi[j] = p; }
    }
    const size_t s = ((size_t)dd * Ho + ho) * Wo + wo;
    const size_t o = ((size_t)b * So + s) * C + q * 4;
    if (out) *reinterpret_cast<float4*>(out + o) = make_float4(best[0], best[1], best[2], best[3]);
    *reinterpret_cast<uchar4*>(idx + o) = make_uchar4((unsigned char)bi[0], (unsigned char)bi[1], (unsigned char)bi[2], (unsigned char)bi[3]);
    if (pk) {
      __nv_bfloat16 hh[4], ll[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split_bf16(best[j], hh[j], ll[j]);
      __nv_bfloat16* dst = pk + (((size_t)b * C8 + (q >> 1)) * So + s) * 8 + (q & 1) * 4;
      *reinterpret_cast<uint2*>(dst) = pack4_bf16(hh[0], hh[1], hh[2], hh[3]);
      if (write_lo) *reinterpret_cast<uint2*>(dst + plane) = pack4_bf16(ll[0], ll[1], ll[2], ll[3]);
    }
  }
}
ICL_API int icl_maxpool3d_fwd(const float* x, float* out, unsigned char* idx, void* pk, int write_lo, int B, int C, int D, int H, int W,
                              void* stream) {
  ICL_REQUIRE(D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool3d: odd spatial size %dx%dx%d", D, H, W);
  ICL_REQUIRE(pk == nullptr || C % 8 == 0, "maxpool3d: PK output needs C %% 8 == 0");
  long long total = (long long)B * C * (D / 2) * (H / 2) * (W / 2);
  if (C % 8 == 0 && total / 4 < (1LL << 31)) {
    maxpool_fwd_v4_k<<<grid_for(total / 4, 256, 148 * 32), 256, 0, as_stream(stream)>>>(x, out, idx, (__nv_bfloat16*)pk, write_lo, B, C, D, H, W);
    ICL_LAUNCHED("maxpool3d_fwd");
  }
  maxpool_fwd_k<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(x, out, idx, (__nv_bfloat16*)pk, write_lo, B, C, D, H, W);
  ICL_LAUNCHED("maxpool3d_fwd");
}

// dx[in voxel] (+)= (idx[out voxel] == my position) ? dout : 0     (gather form, no atomics)
__global__ void maxpool_bwd_k(const float* __restrict__ dout, const unsigned char* __restrict__ idx, float* __restrict__ dx, int accumulate,
                              int B, int C, int D, int H, int W) {
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const long long total = (long long)B * D * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long v = i / C;
    const int w = (int)(v % W); v /= W;
    const int h = (int)(v % H); v /= H;
    const int d = (int)(v % D);
    const int b = (int)(v / D);
    const long long o = ((((long long)b * Do + d / 2) * Ho + h / 2) * Wo + w / 2) * C + c;
    const int p = ((d & 1) << 2) | ((h & 1) << 1) | (w & 1);
    const float g = (idx[o] == p) ? dout[o] : 0.f;
    dx[i] = accumulate ? dx[i] + g : g;
  }
}
// C % 4 == 0: thread = one input voxel x 4 channels (index math once per voxel, 128-bit accesses)
__global__ void __launch_bounds__(256) maxpool_bwd_v4_k(const float* __restrict__ dout, const unsigned char* __restrict__ idx,
                                                        float* __restrict__ dx, int accumulate, int B, int C, int D, int H, int W) {
  const int Do = D / 2, Ho = H / 2, Wo = W / 2, C4 = C >> 2;
  const long long total = (long long)B * D * H * W * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    long long v = i / C4;
    const int w = (int)(v % W); v /= W;
    const int h = (int)(v % H); v /= H;
    const int d = (int)(v % D);
    const int b = (int)(v / D);
    const long long o = ((((long long)b * Do + d / 2) * Ho + h / 2) * Wo + w / 2) * C + c;
    const int p = ((d & 1) << 2) | ((h & 1) << 1) | (w & 1);
    const uchar4 id = *reinterpret_cast<const uchar4*>(idx + o);
    const float4 g = *reinterpret_cast<const float4*>(dout + o);
    float4 r = make_float4(id.x == p ? g.x : 0.f, id.y == p ? g.y : 0.f, id.z == p ? g.z : 0.f, id.w == p ? g.w : 0.f);
    float4* dst = reinterpret_cast<float4*>(dx + ((((long long)b * D + d) * H + h) * W + w) * C + c);
    if (accumulate) { const float4 q = *dst; r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w; }
    *dst = r;
  }
}
ICL_API int icl_maxpool3d_bwd(const float* dout, const unsigned char* idx, float* dx, int accumulate, int B, int C, int D, int H, int W,
                              void* stream) {
  if (C % 4 == 0) {
    maxpool_bwd_v4_k<<<grid_for((long long)B * (C / 4) * D * H * W, 256), 256, 0, as_stream(stream)>>>(dout, idx, dx, accumulate, B, C, D, H, W);
    ICL_LAUNCHED("maxpool3d_bwd");
  }
  long long total = (long long)B * C * D * H * W;
  maxpool_bwd_k<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(dout, idx, dx, accumulate, B, C, D, H, W);
  ICL_LAUNCHED("maxpool3d_bwd");
}

// ------------------------------------------------------------------------------------------
// 2D path (networks/unet_icl.py): a batch of images is ONE sample with the images stacked along D, F32CL [1][Bimg][H][W][C].
// MaxPool2d(2) (unet_icl.py:66) pools H and W only; first maximum in (h, w) scan order wins ties; idx in 0..3.
// ------------------------------------------------------------------------------------------
__global__ void maxpool2d_fwd_k(const float* __restrict__ x, float* __restrict__ out, unsigned char* __restrict__ idx,
                                __nv_bfloat16* __restrict__ pk, int write_lo, int N, int C, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  const long long So = (long long)N * Ho * Wo;
  const long long total = So * C;
  const int C8 = (C % 8 == 0) ? C / 8 : 0;
  const long long plane = (long long)C8 * So * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long v = i / C;
    const int wo = (int)(v % Wo); v /= Wo;
    const int ho = (int)(v % Ho);
    const int n = (int)(v / Ho);
    float best = 0.f; int bi = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int h = 2 * ho + (p >> 1), w = 2 * wo + (p & 1);
      const float val = x[(((long long)n * H + h) * W + w) * C + c];
      if (p == 0 || val > best || val != val) { best = val; bi = p; }
    }
    if (out) out[i] = best;
    idx[i] = (unsigned char)bi;
    if (pk) {
      const long long sidx = ((long long)n * Ho + ho) * Wo + wo;
      __nv_bfloat16 hi, lo; split_bf16(best, hi, lo);
      const long long o = ((long long)(c >> 3) * So + sidx) * 8 + (c & 7);
      pk[o] = hi;
      if (write_lo) pk[plane + o] = lo;
    }
  }
}
__global__ void maxpool2d_bwd_k(const float* __restrict__ dout, const unsigned char* __restrict__ idx, float* __restrict__ dx, int accumulate,
                                int N, int C, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long v = i / C;
    const int w = (int)(v % W); v /= W;
    const int h = (int)(v % H);
    const int n = (int)(v / H);
    const long long o = (((long long)n * Ho + h / 2) * Wo + w / 2) * C + c;
    const int p = ((h & 1) << 1) | (w & 1);
    const float g = (idx[o] == p) ? dout[o] : 0.f;
    dx[i] = accumulate ? dx[i] + g : g;
  }
}
ICL_API int icl_maxpool2d_fwd(const float* x, float* out, unsigned char* idx, void* pk, int write_lo, int N, int C, int H, int W, void* stream) {
  ICL_REQUIRE(H % 2 == 0 && W % 2 == 0, "maxpool2d: odd spatial size %dx%d", H, W);
  ICL_REQUIRE(pk == nullptr || C % 8 == 0, "maxpool2d: PK output needs C %% 8 == 0");
  maxpool2d_fwd_k<<<grid_for((long long)N * C * (H / 2) * (W / 2), 256), 256, 0, as_stream(stream)>>>(x, out, idx, (__nv_bfloat16*)pk, write_lo, N, C,
                                                                                                   H, W);
  ICL_LAUNCHED("maxpool2d_fwd");
}
ICL_API int icl_maxpool2d_bwd(const float* dout, const unsigned char* idx, float* dx, int accumulate, int N, int C, int H, int W, void* stream) {
  maxpool2d_bwd_k<<<grid_for((long long)N * C * H * W, 256), 256, 0, as_stream(stream)>>>(dout, idx, dx, accumulate, N, C, H, W);
  ICL_LAUNCHED("maxpool2d_bwd");
}

// nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) (unet_icl.py:84-85): src = dst * (n - 1) / (2n - 1).
__device__ __forceinline__ void up2ac_src(int j, int n, int& i0, int& i1, float& l1) {
  const float s = (n > 1) ? (float)j * ((float)(n - 1) / (float)(2 * n - 1)) : 0.f;
  i0 = (int)s;
  i1 = min(i0 + 1, n - 1);
  l1 = s - (float)i0;
}
__global__ void upsample2x_ac2d_fwd_k(const float* __restrict__ x, float* __restrict__ out, __nv_bfloat16* __restrict__ pk, int write_lo,
                                      int N, int C, int h, int w) {
  const int H = 2 * h, W = 2 * w;
  const long long S = (long long)N * H * W;
  const long long total = S * C;
  const int C8 = (C % 8 == 0) ? C / 8 : 0;
  const long long plane = (long long)C8 * S * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long v = i / C;
    const int X = (int)(v % W); v /= W;
    const int Y = (int)(v % H);
    const int n = (int)(v / H);
    int y0, y1, x0, x1; float ly, lx;
    up2ac_src(Y, h, y0, y1, ly); up2ac_src(X, w, x0, x1, lx);
    const float* xb = x + (long long)n * h * w * C + c;
#define AT(yy, xx) xb[((long long)(yy) * w + (xx)) * C]
    // same association order as ATen's upsample_bilinear2d: w-lerp inside h-lerp
    const float r = (1.f - ly) * ((1.f - lx) * AT(y0, x0) + lx * AT(y0, x1)) + ly * ((1.f - lx) * AT(y1, x0) + lx * AT(y1, x1));
#undef AT
    if (out) out[i] = r;
    if (pk) {
      const long long sidx = ((long long)n * H + Y) * W + X;
      __nv_bfloat16 hi, lo; split_bf16(r, hi, lo);
      const long long o = ((long long)(c >> 3) * S + sidx) * 8 + (c & 7);
      pk[o] = hi;
      if (write_lo) pk[plane + o] = lo;
    }
  }
}
__device__ __forceinline__ float up2ac_wt(int j, int n, int i) {
  if (j < 0 || j >= 2 * n) return 0.f;
  int i0, i1; float l1;
  up2ac_src(j, n, i0, i1, l1);
  float wgt = 0.f;
  if (i0 == i) wgt += 1.f - l1;
  if (i1 == i) wgt += l1;
  return wgt;
}
// adjoint in gather form: coarse pixel i collects from the fine pixels j with src(j) in (i - 1, i + 1), i.e. j in [2i - 2, 2i + 3]
__global__ void upsample2x_ac2d_bwd_k(const float* __restrict__ dout, int Cd, int c_off, float* __restrict__ dx, int accumulate, int N, int C, int h,
                                      int w) {
  const int H = 2 * h, W = 2 * w;
  const long long total = (long long)N * h * w * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long v = i / C;
    const int x = (int)(v % w); v /= w;
    const int y = (int)(v % h);
    const int n = (int)(v / h);
    float acc = 0.f;
    for (int dy = -2; dy <= 3; ++dy) {
      const int Y = 2 * y + dy; const float wy = up2ac_wt(Y, h, y);
      if (wy == 0.f) continue;
      for (int dxx = -2; dxx <= 3; ++dxx) {
        const int X = 2 * x + dxx; const float wx = up2ac_wt(X, w, x);
        if (wx == 0.f) continue;
        acc += wy * wx * dout[(((long long)n * H + Y) * W + X) * Cd + c_off + c];
      }
    }
    dx[i] = accumulate ? dx[i] + acc : acc;
  }
}
ICL_API int icl_upsample2x_ac2d_fwd(const float* x, float* out, void* pk, int write_lo, int N, int C, int h, int w, void* stream) {
  ICL_REQUIRE(pk == nullptr || C % 8 == 0, "upsample2x_ac2d: PK output needs C %% 8 == 0");
  ICL_REQUIRE(pk != nullptr || out != nullptr, "upsample2x_ac2d: no output requested");
  upsample2x_ac2d_fwd_k<<<grid_for((long long)N * C * 4 * h * w, 256), 256, 0, as_stream(stream)>>>(x, out, (__nv_bfloat16*)pk, write_lo, N, C, h, w);
  ICL_LAUNCHED("upsample2x_ac2d_fwd");
}
ICL_API int icl_upsample2x_ac2d_bwd(const float* dout, int Cd, int c_off, float* dx, int accumulate, int N, int C, int h, int w, void* stream) {
  upsample2x_ac2d_bwd_k<<<grid_for((long long)N * C * h * w, 256), 256, 0, as_stream(stream)>>>(dout, Cd, c_off, dx, accumulate, N, C, h, w);
  ICL_LAUNCHED("upsample2x_ac2d_bwd");
}

// ------------------------------------------------------------------------------------------
// Trilinear x2 upsample, align_corners=False: src = (dst+0.5)/2 - 0.5 clamped at 0.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void up2_src(int j, int n, int& i0, int& i1, float& l1) {
  float s = (j + 0.5f) * 0.5f - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  i1 = min(i0 + 1, n - 1);
  l1 = s - (float)i0;
}
__global__ void upsample2x_fwd_k(const float* __restrict__ x, float* __restrict__ out, __nv_bfloat16* __restrict__ pk, int write_lo,
                                 int B, int C, int d, int h, int w) {
  const int D = 2 * d, H = 2 * h, W = 2 * w;
  const long long S = (long long)D * H * W;
  const long long total = (long long)B * S * C;
  const int C8 = (C % 8 == 0) ? C / 8 : 0;
  const long long plane = (long long)B * C8 * S * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long v = i / C;
    const int X = (int)(v % W); v /= W;
    const int Y = (int)(v % H); v /= H;
    const int Z = (int)(v % D);
    const int b = (int)(v / D);
    int z0, z1, y0, y1, x0, x1; float lz, ly, lx;
    up2_src(Z, d, z0, z1, lz); up2_src(Y, h, y0, y1, ly); up2_src(X, w, x0, x1, lx);
    const float* xb = x + (long long)b * d * h * w * C + c;
#define AT(zz, yy, xx) xb[(((long long)(zz) * h + (yy)) * w + (xx)) * C]
    // same association order as ATen's upsample_trilinear3d: w-lerp inside h-lerp inside d-lerp
    const float v00 = (1.f - lx) * AT(z0, y0, x0) + lx * AT(z0, y0, x1);
    const float v01 = (1.f - lx) * AT(z0, y1, x0) + lx * AT(z0, y1, x1);
    const float v10 = (1.f - lx) * AT(z1, y0, x0) + lx * AT(z1, y0, x1);
    const float v11 = (1.f - lx) * AT(z1, y1, x0) + lx * AT(z1, y1, x1);
#undef AT
    const float r = (1.f - lz) * ((1.f - ly) * v00 + ly * v01) + lz * ((1.f - ly) * v10 + ly * v11);
    if (out) out[i] = r;
    if (pk) {
      const long long s = ((long long)Z * H + Y) * W + X;
      __nv_bfloat16 hi, lo; split_bf16(r, hi, lo);
      const long long o = (((long long)b * C8 + (c >> 3)) * S + s) * 8 + (c & 7);
      pk[o] = hi;
      if (write_lo) pk[plane + o] = lo;
    }
  }
}
// C % 8 == 0: thread = one fine voxel x 8 channels (grid.y = sample x channel group); the 8 coarse corners are 32-byte
// loads, the PK store is 16 B per plane contiguous along w, the optional fp32 store 32 B.
__global__ void __launch_bounds__(256) upsample2x_fwd_v8_k(const float* __restrict__ x, float* __restrict__ out, __nv_bfloat16* __restrict__ pk,
                                                           int write_lo, int B, int C, int d, int h, int w) {
  const int D = 2 * d, H = 2 * h, W = 2 * w, C8 = C >> 3;
  const long long S = (long long)D * H * W;
  const int b = blockIdx.y / C8, c8 = blockIdx.y % C8;
  const long long plane = (long long)B * C8 * S * 8;
  const float* xb = x + (long long)b * d * h * w * C + c8 * 8;
  __nv_bfloat16* hi = pk ? pk + ((long long)b * C8 + c8) * S * 8 : nullptr;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(s % W), Y = (int)((s / W) % H), Z = (int)(s / ((long long)W * H));
    int z0, z1, y0, y1, x0, x1; float lz, ly, lx;
    up2_src(Z, d, z0, z1, lz); up2_src(Y, h, y0, y1, ly); up2_src(X, w, x0, x1, lx);
    float c[8][8];
#define LD(i, zz, yy, xx) unpack8(ld8(xb + (((long long)(zz) * h + (yy)) * w + (xx)) * C), c[i])
    LD(0, z0, y0, x0); LD(1, z0, y0, x1); LD(2, z0, y1, x0); LD(3, z0, y1, x1);
    LD(4, z1, y0, x0); LD(5, z1, y0, x1); LD(6, z1, y1, x0); LD(7, z1, y1, x1);
#undef LD
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // same association order as ATen's upsample_trilinear3d: w-lerp inside h-lerp inside d-lerp
      const float v00 = (1.f - lx) * c[0][i] + lx * c[1][i], v01 = (1.f - lx) * c[2][i] + lx * c[3][i];
      const float v10 = (1.f - lx) * c[4][i] + lx * c[5][i], v11 = (1.f - lx) * c[6][i] + lx * c[7][i];
      r[i] = (1.f - lz) * ((1.f - ly) * v00 + ly * v01) + lz * ((1.f - ly) * v10 + ly * v11);
    }
    if (out) st8(out + ((long long)b * S + s) * C + c8 * 8, r);
    if (hi) st_pk8(hi + s * 8, write_lo ? hi + plane + s * 8 : nullptr, r);
  }
}
// Same with the 8-channel group as the fastest thread index (C/8 a power of two <= 32, S * C/8 < 2^31): neighbouring lanes read
// neighbouring 32-byte segments of the same coarse voxel, so a corner load touches C/32 lines per voxel instead of one line per
// lane (the kernel above needs 32 L1 wavefronts per 128-bit request at C = 32), and the index arithmetic is 32-bit.
__global__ void __launch_bounds__(256) upsample2x_fwd_cl_k(const float* __restrict__ x, float* __restrict__ out, __nv_bfloat16* __restrict__ pk,
                                                           int write_lo, int B, int C, int d, int h, int w, int c8_log2) {
  const int D = 2 * d, H = 2 * h, W = 2 * w, C8 = C >> 3;
  const long long S = (long long)D * H * W;
  const int b = blockIdx.y, c8 = threadIdx.x & (C8 - 1);
  const long long plane = (long long)B * C8 * S * 8;
  const float* xb = x + (long long)b * d * h * w * C + c8 * 8;
  __nv_bfloat16* hi = pk ? pk + ((long long)b * C8 + c8) * S * 8 : nullptr;
  float* ob = out ? out + (long long)b * S * C + c8 * 8 : nullptr;
  const unsigned total = (unsigned)(S << c8_log2);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned s = i >> c8_log2;
    const unsigned q = s / (unsigned)W;
    const int X = (int)(s - q * (unsigned)W), Z = (int)(q / (unsigned)H), Y = (int)(q - (unsigned)Z * (unsigned)H);
    int z0, z1, y0, y1, x0, x1; float lz, ly, lx;
    up2_src(Z, d, z0, z1, lz); up2_src(Y, h, y0, y1, ly); up2_src(X, w, x0, x1, lx);
    float c[8][8];
#define LD(i_, zz, yy, xx) unpack8(ld8(xb + (size_t)(((zz) * h + (yy)) * w + (xx)) * C), c[i_])
    LD(0, z0, y0, x0); LD(1, z0, y0, x1); LD(2, z0, y1, x0); LD(3, z0, y1, x1);
    LD(4, z1, y0, x0); LD(5, z1, y0, x1); LD(6, z1, y1, x0); LD(7, z1, y1, x1);
#undef LD
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // same association order as ATen's upsample_trilinear3d: w-lerp inside h-lerp inside d-lerp
      const float v00 = (1.f - lx) * c[0][j] + lx * c[1][j], v01 = (1.f - lx) * c[2][j] + lx * c[3][j];
      const float v10 = (1.f - lx) * c[4][j] + lx * c[5][j], v11 = (1.f - lx) * c[6][j] + lx * c[7][j];
      r[j] = (1.f - lz) * ((1.f - ly) * v00 + ly * v01) + lz * ((1.f - ly) * v10 + ly * v11);
    }
    if (ob) st8(ob + (size_t)s * C, r);
    if (hi) st_pk8(hi + (size_t)s * 8, write_lo ? hi + plane + (size_t)s * 8 : nullptr, r);
  }
}
// Thread -> (coarse voxel, channel quad) with a COMPACT 3-D tile of coarse voxels per 256-thread block (up to 4 x 4 x 2: x, y, z), the
// quad fastest: the neighbourhoods of a block's voxels overlap, so most of their loads hit L1 (a linear sweep along x re-reads every
// y / z neighbour from L2: 4x the compulsory traffic, which made the adjoint kernel L2-bound).  C/4 must be a power of two <= 256.
struct Up2Tile { int b, z, y, x, q; bool valid; };
__device__ __forceinline__ Up2Tile up2_tile(int C4, int d, int h, int w) {
  const int vpb = 256 / C4;
  const int tx = min(4, vpb), ty = min(4, vpb / tx), tz = vpb / (tx * ty);
  const int ntx = cdiv(w, tx), nty = cdiv(h, ty), ntz = cdiv(d, tz);
  int t = blockIdx.x;
  const int bx = t % ntx; t /= ntx;
  const int by = t % nty; t /= nty;
  const int bz = t % ntz;
  Up2Tile o;
  o.b = t / ntz;
  o.q = threadIdx.x % C4;
  const int lv = threadIdx.x / C4;
  o.x = bx * tx + lv % tx;
  o.y = by * ty + (lv / tx) % ty;
  o.z = bz * tz + lv / (tx * ty);
  o.valid = o.x < w && o.y < h && o.z < d;
  return o;
}
static inline long long up2_tile_blocks(int B, int C4, int d, int h, int w) {
  const int vpb = 256 / C4;
  const int tx = vpb < 4 ? vpb : 4, ty = (vpb / tx) < 4 ? (vpb / tx) : 4, tz = vpb / (tx * ty);
  return (long long)B * cdiv(d, tz) * cdiv(h, ty) * cdiv(w, tx);
}
static inline bool up2_tile_ok(int C) { const int c4 = C / 4; return C % 8 == 0 && c4 <= 256 && (c4 & (c4 - 1)) == 0; }
// Thread = one COARSE voxel x 4 channels (channel quad fastest across lanes): the 27 coarse neighbours are loaded once (float4 each,
// fully coalesced: a warp reads whole 128-byte lines) and all 8 fine voxels of the cell are formed from them, 3.4 loads per fine voxel
// instead of 8 (the per-fine-voxel kernels above are bound by L1 load wavefronts).  Arithmetic per output is ATen's, in ATen's order:
// w-lerp inside h-lerp inside d-lerp with the (index, weight) pairs of up2_src, so values are unchanged.  For the PK store a lane pair
// (channel quads 2g, 2g+1 = one 8-channel brick) swaps halves so that each lane writes one whole 16-byte brick row (X = 2x or 2x+1).
__device__ __forceinline__ float4 lerp4(float l, const float4& a, const float4& b) {
  const float m = 1.f - l;
  return make_float4(m * a.x + l * b.x, m * a.y + l * b.y, m * a.z + l * b.z, m * a.w + l * b.w);
}
__device__ __forceinline__ float4 sel4(bool c, const float4& a, const float4& b) { return c ? a : b; }
__global__ void __launch_bounds__(256, 2) upsample2x_fwd_c27_k(const float* __restrict__ x, float* __restrict__ out, __nv_bfloat16* __restrict__ pk,
                                                               int write_lo, int B, int C, int d, int h, int w) {
  const int D = 2 * d, H = 2 * h, W = 2 * w, C4 = C >> 2, C8 = C >> 3;
  const size_t S = (size_t)D * H * W;
  const size_t plane = (size_t)B * C8 * S * 8;
  const unsigned total = (unsigned)B * d * h * w * C4;            // host guarantees < 2^31, C4 even: lane pairs never straddle the end
  const unsigned nthreads = gridDim.x * blockDim.x;
  const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x);
  // (a linear sweep, x fastest: the compact tiles of the adjoint kernel below made THIS kernel slower — it is bound by instruction
  // issue, not by L2, and one short-lived block per tile costs more than the grid-stride loop)
  for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 - (threadIdx.x & 31) < total; i0 += nthreads) {
    const bool valid = i0 < total;
    const unsigned i = valid ? i0 : total - 1;
    const int q = (int)(i % (unsigned)C4);
    unsigned v = i / (unsigned)C4;
    const int cx = (int)(v % (unsigned)w); v /= (unsigned)w;
    const int cy = (int)(v % (unsigned)h); v /= (unsigned)h;
    const int cz = (int)(v % (unsigned)d);
    const int b = (int)(v / (unsigned)d);
    // Neighbours clamped to the volume.  Fine index 2c blends (c-1, c) with weight .75 on c; at c = 0 up2_src clamps to (0, 1) with
    // weight 0, i.e. the value of voxel 0 itself: with the clamped neighbour (voxel 0 again) and weight 0 that is (1-0)*v0 + 0*v0,
    // the same number.  Fine index 2c+1 blends (c, c+1) with weight .25, c+1 clamped at the end (up2_src does the same).
    const float lx0 = cx > 0 ? 0.75f : 0.f, ly0 = cy > 0 ? 0.75f : 0.f, lz0 = cz > 0 ? 0.75f : 0.f;
    // offsets in float4 units (host guarantees the coarse tensor has < 2^31 floats)
    const unsigned xo[3] = {(unsigned)max(cx - 1, 0) * C4, (unsigned)cx * C4, (unsigned)min(cx + 1, w - 1) * C4};
    const unsigned yo[3] = {(unsigned)max(cy - 1, 0) * w * C4, (unsigned)cy * w * C4, (unsigned)min(cy + 1, h - 1) * w * C4};
    const unsigned hw = (unsigned)h * w * C4;
    const unsigned zo[3] = {(unsigned)max(cz - 1, 0) * hw, (unsigned)cz * hw, (unsigned)min(cz + 1, d - 1) * hw};
    const float4* xb = x4 + (size_t)b * d * hw + q;
    float4 P[3][2][2];   // [coarse plane][fine dy][fine dx]: h-lerp of w-lerps
#pragma unroll
    for (int kz = 0; kz < 3; ++kz) {
      float4 vx[3][2];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const float4* row = xb + (zo[kz] + yo[ky]);
        const float4 n0 = row[xo[0]], n1 = row[xo[1]], n2 = row[xo[2]];
        vx[ky][0] = lerp4(lx0, n0, n1);
        vx[ky][1] = lerp4(0.25f, n1, n2);
      }
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        P[kz][0][dx] = lerp4(ly0, vx[0][dx], vx[1][dx]);
        P[kz][1][dx] = lerp4(0.25f, vx[1][dx], vx[2][dx]);
      }
    }
    const int odd = q & 1;
    // 32-bit element offsets (host guarantees B*C*S < 2^31)
    const unsigned S32 = (unsigned)D * H * W;
    const unsigned s00 = ((unsigned)(2 * cz) * H + 2 * cy) * W + 2 * cx;
    __nv_bfloat16* const ph = pk + (size_t)((((unsigned)b * C8 + (q >> 1)) * S32 + s00 + odd) * 8u);
    __nv_bfloat16* const pl = ph + plane;
    float* const po = out + (size_t)(((unsigned)b * S32 + s00) * (unsigned)C + q * 4);
#pragma unroll
    for (int dz = 0; dz < 2; ++dz) {
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        float4 r[2];
#pragma unroll
        for (int dx = 0; dx < 2; ++dx)
          r[dx] = dz == 0 ? lerp4(lz0, P[0][dy][dx], P[1][dy][dx]) : lerp4(0.25f, P[1][dy][dx], P[2][dy][dx]);
        const unsigned ds = ((unsigned)dz * H + dy) * W;   // fine-voxel offset of this (dz, dy) row
        if (out && valid) {
          float* o = po + ds * (unsigned)C;
          *reinterpret_cast<float4*>(o) = r[0];
          *reinterpret_cast<float4*>(o + C) = r[1];
        }
        if (pk) {
          uint2 hi[2], lo[2];
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            // split x = hi + lo two channels at a time (same roundings as split_bf16)
            const __nv_bfloat162 h01 = __floats2bfloat162_rn(r[dx].x, r[dx].y), h23 = __floats2bfloat162_rn(r[dx].z, r[dx].w);
            const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
            const __nv_bfloat162 l01 = __floats2bfloat162_rn(r[dx].x - f01.x, r[dx].y - f01.y);
            const __nv_bfloat162 l23 = __floats2bfloat162_rn(r[dx].z - f23.x, r[dx].w - f23.y);
            hi[dx] = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
            lo[dx] = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
          }
          // even quad keeps X = 2x (dx 0) and receives the partner's channels 4..7 of it; odd quad keeps X = 2x+1
          const uint2 sh = odd ? hi[0] : hi[1], sl = odd ? lo[0] : lo[1];
          uint2 gh, gl;
          gh.x = __shfl_xor_sync(0xffffffffu, sh.x, 1); gh.y = __shfl_xor_sync(0xffffffffu, sh.y, 1);
          gl.x = __shfl_xor_sync(0xffffffffu, sl.x, 1); gl.y = __shfl_xor_sync(0xffffffffu, sl.y, 1);
          if (valid) {
            const uint2 mh = odd ? hi[1] : hi[0], ml = odd ? lo[1] : lo[0];
            const uint4 vh = odd ? make_uint4(gh.x, gh.y, mh.x, mh.y) : make_uint4(mh.x, mh.y, gh.x, gh.y);
            *reinterpret_cast<uint4*>(ph + ds * 8u) = vh;
            if (write_lo) {
              const uint4 vl = odd ? make_uint4(gl.x, gl.y, ml.x, ml.y) : make_uint4(ml.x, ml.y, gl.x, gl.y);
              *reinterpret_cast<uint4*>(pl + ds * 8u) = vl;
            }
          }
        }
      }
    }
  }
}
ICL_API int icl_upsample2x_fwd(const float* x, float* out, void* pk, int write_lo, int B, int C, int d, int h, int w, void* stream) {
  ICL_REQUIRE(pk == nullptr || C % 8 == 0, "upsample2x: PK output needs C %% 8 == 0");
  ICL_REQUIRE(pk != nullptr || out != nullptr, "upsample2x: no output requested");
  const int C8u = C / 8;
  if (C % 8 == 0 && 8LL * B * d * h * w * C < (1LL << 31)) {
    const long long threads = (long long)B * d * h * w * (C / 4);
    upsample2x_fwd_c27_k<<<grid_for(threads, 256, 148 * 32), 256, 0, as_stream(stream)>>>(x, out, (__nv_bfloat16*)pk, write_lo, B, C, d, h, w);
    ICL_LAUNCHED("upsample2x_fwd");
  }
  if (C % 8 == 0 && C8u <= 32 && (C8u & (C8u - 1)) == 0 && 8LL * d * h * w * C8u < (1LL << 31)) {
    const long long S = 8LL * d * h * w;
    int lg = 0;
    while ((1 << lg) < C8u) ++lg;
    int gx = (int)min((long long)cdiv(S * C8u, 256), (long long)max(1, 148 * 8 / B));
    upsample2x_fwd_cl_k<<<dim3(gx, B), 256, 0, as_stream(stream)>>>(x, out, (__nv_bfloat16*)pk, write_lo, B, C, d, h, w, lg);
    ICL_LAUNCHED("upsample2x_fwd");
  }
  if (C % 8 == 0) {
    const long long S = 8LL * d * h * w;
    int gx = (int)min((long long)cdiv(S, 256), (long long)max(1, 148 * 8 / (B * (C / 8)) + 1));
    upsample2x_fwd_v8_k<<<dim3(gx, B * (C / 8)), 256, 0, as_stream(stream)>>>(x, out, (__nv_bfloat16*)pk, write_lo, B, C, d, h, w);
    ICL_LAUNCHED("upsample2x_fwd");
  }
  long long total = (long long)B * C * 8 * d * h * w;
  upsample2x_fwd_k<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(x, out, (__nv_bfloat16*)pk, write_lo, B, C, d, h, w);
  ICL_LAUNCHED("upsample2x_fwd");
}

// adjoint in gather form: coarse voxel i collects from fine j in [2i-1, 2i+2]
__device__ __forceinline__ float up2_wt(int j, int n, int i) {
  if (j < 0 || j >= 2 * n) return 0.f;
  int i0, i1; float l1;
  up2_src(j, n, i0, i1, l1);
  float wgt = 0.f;
  if (i0 == i) wgt += 1.f - l1;
  if (i1 == i) wgt += l1;
  return wgt;
}
__global__ void upsample2x_bwd_k(const float* __restrict__ dout, int Cd, int c_off, float* __restrict__ dx, int accumulate,
                                 int B, int C, int d, int h, int w) {
  // dout is F32CL with Cd channels; this op's channels start at c_off (lets the caller pass the
  // [skip|up] gradient of a concatenated conv input without slicing).
  const int D = 2 * d, H = 2 * h, W = 2 * w;
  const long long total = (long long)B * d * h * w * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long v = i / C;
    const int x = (int)(v % w); v /= w;
    const int y = (int)(v % h); v /= h;
    const int z = (int)(v % d);
    const int b = (int)(v / d);
    float acc = 0.f;
    for (int dz = -1; dz <= 2; ++dz) {
      const int Z = 2 * z + dz; const float wz = up2_wt(Z, d, z);
      if (wz == 0.f) continue;
      for (int dy = -1; dy <= 2; ++dy) {
        const int Y = 2 * y + dy; const float wy = up2_wt(Y, h, y);
        if (wy == 0.f) continue;
        for (int dxx = -1; dxx <= 2; ++dxx) {
          const int X = 2 * x + dxx; const float wx = up2_wt(X, w, x);
          if (wx == 0.f) continue;
          acc += wz * wy * wx * dout[((((long long)b * D + Z) * H + Y) * W + X) * Cd + c_off + c];
        }
      }
    }
    dx[i] = accumulate ? dx[i] + acc : acc;
  }
}
// 4 channels per thread (float4): used when C, Cd and c_off are multiples of 4
__global__ void __launch_bounds__(256) upsample2x_bwd_v4_k(const float* __restrict__ dout, int Cd, int c_off, float* __restrict__ dx,
                                                           int accumulate, int B, int C, int d, int h, int w) {
  const int D = 2 * d, H = 2 * h, W = 2 * w, C4 = C >> 2;
  const long long total = (long long)B * d * h * w * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    long long v = i / C4;
    const int x = (int)(v % w); v /= w;
    const int y = (int)(v % h); v /= h;
    const int z = (int)(v % d);
    const int b = (int)(v / d);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int dz = -1; dz <= 2; ++dz) {
      const int Z = 2 * z + dz; const float wz = up2_wt(Z, d, z);
      if (wz == 0.f) continue;
#pragma unroll
      for (int dy = -1; dy <= 2; ++dy) {
        const int Y = 2 * y + dy; const float wy = up2_wt(Y, h, y);
        if (wy == 0.f) continue;
#pragma unroll
        for (int dxx = -1; dxx <= 2; ++dxx) {
          const int X = 2 * x + dxx; const float wx = up2_wt(X, w, x);
          if (wx == 0.f) continue;
          const float wgt = wz * wy * wx;
          const float4 g = *reinterpret_cast<const float4*>(dout + ((((long long)b * D + Z) * H + Y) * W + X) * Cd + c_off + c);
          acc.x += wgt * g.x; acc.y += wgt * g.y; acc.z += wgt * g.z; acc.w += wgt * g.w;
        }
      }
    }
    float4* o = reinterpret_cast<float4*>(dx + (v = 0, ((((long long)b * d + z) * h + y) * w + x) * C + c));
    if (accumulate) { const float4 p = *o; acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w; }
    *o = acc;
  }
}
// Separable gather: the 4 weights per axis are formed once per thread (up2_wt, so the boundary rules are the forward's), the
// x-sum of a fine row is taken first and scaled by wz*wy once (64 loads, 80 float4 FMAs, no per-tap weight logic).
__global__ void __launch_bounds__(256) upsample2x_bwd_sep_k(const float* __restrict__ dout, int Cd, int c_off, float* __restrict__ dx,
                                                            int accumulate, int B, int C, int d, int h, int w) {
  const int D = 2 * d, H = 2 * h, W = 2 * w, C4 = C >> 2;
  {
    const Up2Tile t = up2_tile(C4, d, h, w);
    if (!t.valid) return;
    const int c = t.q * 4, x = t.x, y = t.y, z = t.z, b = t.b;
    const float* base = dout + (size_t)b * D * H * W * Cd + c_off + c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (x > 0 && x < w - 1 && y > 0 && y < h - 1 && z > 0 && z < d - 1) {
      // interior cell: the weights of fine offsets -1..2 are (.25, .75, .75, .25) on every axis (what up2_wt returns there)
      const float* p0 = base + (((size_t)(2 * z - 1) * H + (2 * y - 1)) * W + (2 * x - 1)) * Cd;
      const size_t rs = (size_t)W * Cd, ps = (size_t)H * W * Cd;
#pragma unroll
      for (int kz = 0; kz < 4; ++kz) {
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
          const float* row = p0 + kz * ps + ky * rs;
          const float4 g0 = *reinterpret_cast<const float4*>(row), g1 = *reinterpret_cast<const float4*>(row + Cd);
          const float4 g2 = *reinterpret_cast<const float4*>(row + 2 * Cd), g3 = *reinterpret_cast<const float4*>(row + 3 * Cd);
          float4 t;   // same order of the adds as the general path below
          t.x = 0.25f * g0.x; t.x += 0.75f * g1.x; t.x += 0.75f * g2.x; t.x += 0.25f * g3.x;
          t.y = 0.25f * g0.y; t.y += 0.75f * g1.y; t.y += 0.75f * g2.y; t.y += 0.25f * g3.y;
          t.z = 0.25f * g0.z; t.z += 0.75f * g1.z; t.z += 0.75f * g2.z; t.z += 0.25f * g3.z;
          t.w = 0.25f * g0.w; t.w += 0.75f * g1.w; t.w += 0.75f * g2.w; t.w += 0.25f * g3.w;
          const float wzy = ((kz == 0 || kz == 3) ? 0.25f : 0.75f) * ((ky == 0 || ky == 3) ? 0.25f : 0.75f);
          acc.x += wzy * t.x; acc.y += wzy * t.y; acc.z += wzy * t.z; acc.w += wzy * t.w;
        }
      }
    } else {
      float wz[4], wy[4], wx[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        wz[k] = up2_wt(2 * z - 1 + k, d, z);
        wy[k] = up2_wt(2 * y - 1 + k, h, y);
        wx[k] = up2_wt(2 * x - 1 + k, w, x);
      }
#pragma unroll
      for (int kz = 0; kz < 4; ++kz) {
        if (wz[kz] == 0.f) continue;
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
          if (wy[ky] == 0.f) continue;
          const float* row = base + ((size_t)(2 * z - 1 + kz) * H + (2 * y - 1 + ky)) * W * Cd;
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int kx = 0; kx < 4; ++kx) {
            if (wx[kx] == 0.f) continue;
            const float4 g = *reinterpret_cast<const float4*>(row + (size_t)(2 * x - 1 + kx) * Cd);
            t.x += wx[kx] * g.x; t.y += wx[kx] * g.y; t.z += wx[kx] * g.z; t.w += wx[kx] * g.w;
          }
          const float wzy = wz[kz] * wy[ky];
          acc.x += wzy * t.x; acc.y += wzy * t.y; acc.z += wzy * t.z; acc.w += wzy * t.w;
        }
      }
    }
    float4* o = reinterpret_cast<float4*>(dx + ((((size_t)b * d + z) * h + y) * w + x) * C + c);
    if (accumulate) { const float4 p = *o; acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w; }
    *o = acc;
  }
}
ICL_API int icl_upsample2x_bwd(const float* dout, int Cd, int c_off, float* dx, int accumulate, int B, int C, int d, int h, int w,
                               void* stream) {
  if (up2_tile_ok(C) && Cd % 4 == 0 && c_off % 4 == 0 && up2_tile_blocks(B, C / 4, d, h, w) < (1LL << 31)) {
    upsample2x_bwd_sep_k<<<(unsigned)up2_tile_blocks(B, C / 4, d, h, w), 256, 0, as_stream(stream)>>>(dout, Cd, c_off, dx, accumulate,
                                                                                                                 B, C, d, h, w);
    ICL_LAUNCHED("upsample2x_bwd");
  }
  if (C % 4 == 0 && Cd % 4 == 0 && c_off % 4 == 0) {
    upsample2x_bwd_v4_k<<<grid_for((long long)B * (C / 4) * d * h * w, 256), 256, 0, as_stream(stream)>>>(dout, Cd, c_off, dx, accumulate, B, C,
                                                                                                        d, h, w);
    ICL_LAUNCHED("upsample2x_bwd");
  }
  long long total = (long long)B * C * d * h * w;
  upsample2x_bwd_k<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(dout, Cd, c_off, dx, accumulate, B, C, d, h, w);
  ICL_LAUNCHED("upsample2x_bwd");
}

// ------------------------------------------------------------------------------------------
// Dropout (inverted, p): out = x * keep / (1-p).  keep comes either from an explicit byte mask
// (parity tests replay torch-drawn masks) or from Philox4x32 keyed by (seed, element index / 4),
// regenerated identically in backward so no mask is stored.
// ------------------------------------------------------------------------------------------
__global__ void dropout_k(const float* __restrict__ x, float* __restrict__ out, const unsigned char* __restrict__ mask,
                          unsigned long long seed, const unsigned long long* __restrict__ seed_ptr, float p, long long total) {
  if (seed_ptr) seed = seed_ptr[0];  // seed in device memory: the launch can be captured in a CUDA graph and re-seeded per replay
  const float scale = 1.f / (1.f - p);
  const long long n4 = (total + 3) / 4;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
    uint4 r = make_uint4(0, 0, 0, 0);
    if (!mask) r = philox4x32(make_uint4((uint32_t)q, (uint32_t)(q >> 32), 0x1c1u, 0xb200u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = q * 4 + j;
      if (i >= total) break;
      bool keep;
      if (mask) keep = mask[i] != 0;
      else keep = (rr[j] >> 8) * (1.0f / 16777216.0f) >= p;
      out[i] = keep ? x[i] * scale : 0.f;
    }
  }
}
// total % 4 == 0, 16-byte aligned tensors: one 128-bit load and store per Philox draw (the scalar loop above is L1-bound, 2.8 TB/s);
// same counter -> element mapping, so both kernels draw the same mask.
__global__ void __launch_bounds__(256) dropout_v4_k(const float4* __restrict__ x, float4* __restrict__ out, const uchar4* __restrict__ mask,
                                                    unsigned long long seed, const unsigned long long* __restrict__ seed_ptr, float p, long long n4) {
  if (seed_ptr) seed = seed_ptr[0];
  const float scale = 1.f / (1.f - p);
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
    bool k0, k1, k2, k3;
    if (mask) {
      const uchar4 m = mask[q];
      k0 = m.x != 0; k1 = m.y != 0; k2 = m.z != 0; k3 = m.w != 0;
    } else {
      const uint4 r = philox4x32(make_uint4((uint32_t)q, (uint32_t)(q >> 32), 0x1c1u, 0xb200u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      k0 = (r.x >> 8) * (1.0f / 16777216.0f) >= p; k1 = (r.y >> 8) * (1.0f / 16777216.0f) >= p;
      k2 = (r.z >> 8) * (1.0f / 16777216.0f) >= p; k3 = (r.w >> 8) * (1.0f / 16777216.0f) >= p;
    }
    const float4 v = x[q];
    out[q] = make_float4(k0 ? v.x * scale : 0.f, k1 ? v.y * scale : 0.f, k2 ? v.z * scale : 0.f, k3 ? v.w * scale : 0.f);
  }
}
ICL_API int icl_dropout(const float* x, float* out, const unsigned char* mask, unsigned long long seed, const unsigned long long* seed_ptr,
                        float p, long long total, void* stream) {
  ICL_REQUIRE(p >= 0.f && p < 1.f, "dropout: p=%f out of range", p);
  if (total % 4 == 0 && ((uintptr_t)x | (uintptr_t)out) % 16 == 0 && (uintptr_t)mask % 4 == 0) {
    dropout_v4_k<<<grid_for(total / 4, 256, 148 * 32), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(out),
                                                                                     reinterpret_cast<const uchar4*>(mask), seed, seed_ptr, p, total / 4);
    ICL_LAUNCHED("dropout");
  }
  dropout_k<<<grid_for((total + 3) / 4, 256), 256, 0, as_stream(stream)>>>(x, out, mask, seed, seed_ptr, p, total);
  ICL_LAUNCHED("dropout");
}

// ------------------------------------------------------------------------------------------
// `final` 1x1x1 Conv3d(16 -> K) at full resolution (networks/unet_3D_icl.py:65,117) and its backward.  M = all voxels,
// N = K <= 16, reduction = 16 channels: pure streaming (64 B in, 4K B out per voxel) — a tiled GEMM wastes > 90 % of
// its tile on N = 2.  Backward reads g and x once and emits dx, dW and db together (K <= 4: dW / db live in registers).
// ------------------------------------------------------------------------------------------
#define HD_C 16
__global__ void __launch_bounds__(256) head1x1_fwd_k(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                     float* __restrict__ out, long long rows, int K) {
  __shared__ __align__(16) float ws[16 * HD_C + 16];
  for (int i = threadIdx.x; i < K * HD_C; i += 256) ws[i] = w[i];
  for (int i = threadIdx.x; i < K; i += 256) ws[16 * HD_C + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  for (long long v = (long long)blockIdx.x * 256 + threadIdx.x; v < rows; v += (long long)gridDim.x * 256) {
    float xv[HD_C];
    const float4* xp = reinterpret_cast<const float4*>(x + v * HD_C);
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float4 t = xp[q]; xv[4 * q] = t.x; xv[4 * q + 1] = t.y; xv[4 * q + 2] = t.z; xv[4 * q + 3] = t.w; }
    if ((K & 3) == 0) {
      // four classes at a time: 128-bit weight reads from shared memory (a scalar read per FMA made K = 16 LDS-bound: 212 us for
      // 450 MB of traffic) and one 128-bit store per four outputs
      for (int k = 0; k < K; k += 4) {
        float a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          a[j] = ws[16 * HD_C + k + j];
          const float4* wr = reinterpret_cast<const float4*>(ws + (k + j) * HD_C);
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const float4 w4 = wr[c4];
            a[j] = fmaf(xv[4 * c4], w4.x, a[j]); a[j] = fmaf(xv[4 * c4 + 1], w4.y, a[j]);
            a[j] = fmaf(xv[4 * c4 + 2], w4.z, a[j]); a[j] = fmaf(xv[4 * c4 + 3], w4.w, a[j]);
          }
        }
        *reinterpret_cast<float4*>(out + v * K + k) = make_float4(a[0], a[1], a[2], a[3]);
      }
    } else {
      for (int k = 0; k < K; ++k) {
        float a = ws[16 * HD_C + k];
#pragma unroll
        for (int c = 0; c < HD_C; ++c) a = fmaf(xv[c], ws[k * HD_C + c], a);
        out[v * K + k] = a;
      }
    }
  }
}
ICL_API int icl_head1x1_fwd(const float* x, const float* w, const float* bias, float* out, long long rows, int C, int K, void* stream) {
  ICL_REQUIRE(C == HD_C && K >= 1 && K <= 16, "head1x1_fwd: C=%d K=%d unsupported (C must be 16, K <= 16)", C, K);
  head1x1_fwd_k<<<grid_for(rows, 256, 148 * 8), 256, 0, as_stream(stream)>>>(x, w, bias, out, rows, K);
  ICL_LAUNCHED("head1x1_fwd");
}
template <int KT>
__global__ void __launch_bounds__(256) head1x1_bwd_k(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ w,
                                                     float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, long long rows) {
  __shared__ float ws[KT * HD_C];
  __shared__ float red[8][KT * HD_C + KT];
  for (int i = threadIdx.x; i < KT * HD_C; i += 256) ws[i] = w[i];
  __syncthreads();
  float aw[KT][HD_C], ab[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    ab[k] = 0.f;
#pragma unroll
    for (int c = 0; c < HD_C; ++c) aw[k][c] = 0.f;
  }
  for (long long v = (long long)blockIdx.x * 256 + threadIdx.x; v < rows; v += (long long)gridDim.x * 256) {
    float xv[HD_C], gv[KT], o[HD_C];
    const float4* xp = reinterpret_cast<const float4*>(x + v * HD_C);
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float4 t = xp[q]; xv[4 * q] = t.x; xv[4 * q + 1] = t.y; xv[4 * q + 2] = t.z; xv[4 * q + 3] = t.w; }
#pragma unroll
    for (int k = 0; k < KT; ++k) gv[k] = g[v * KT + k];
#pragma unroll
    for (int c = 0; c < HD_C; ++c) o[c] = 0.f;
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      ab[k] += gv[k];
#pragma unroll
      for (int c = 0; c < HD_C; ++c) { o[c] = fmaf(gv[k], ws[k * HD_C + c], o[c]); aw[k][c] = fmaf(gv[k], xv[c], aw[k][c]); }
    }
    float4* dp = reinterpret_cast<float4*>(dx + v * HD_C);
#pragma unroll
    for (int q = 0; q < 4; ++q) dp[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < KT; ++k) {
#pragma unroll
    for (int c = 0; c < HD_C; ++c) { const float t = warp_sum(aw[k][c]); if (lane == 0) red[wid][k * HD_C + c] = t; }
    const float t = warp_sum(ab[k]);
    if (lane == 0) red[wid][KT * HD_C + k] = t;
  }
  __syncthreads();
  if (threadIdx.x < KT * HD_C + KT) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    if (threadIdx.x < KT * HD_C) atomicAdd(dw + threadIdx.x, t);
    else atomicAdd(db + (threadIdx.x - KT * HD_C), t);
  }
}
// K = 16 classes (config 3): 256 weight-gradient accumulators do not fit one thread, so FOUR threads share a voxel, each owning a
// quarter of the classes: 64 accumulators per thread, the partial dx of the four quarters is combined with two xor-shuffles and
// every thread stores one float4 of it.  Same traffic as the small-K kernel: g, x read once (x from L1 for three of the four),
// dx written once.
__global__ void __launch_bounds__(256) head1x1_bwd16_k(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ w,
                                                       float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, long long rows) {
  __shared__ __align__(16) float ws[16 * HD_C];
  __shared__ float red[8][4 * 68];
  for (int i = threadIdx.x; i < 16 * HD_C; i += 256) ws[i] = w[i];
  __syncthreads();
  const int q = threadIdx.x & 3;
  float aw[4][HD_C], ab[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    ab[j] = 0.f;
#pragma unroll
    for (int c = 0; c < HD_C; ++c) aw[j][c] = 0.f;
  }
  const long long iters = (rows + (long long)gridDim.x * 64 - 1) / ((long long)gridDim.x * 64);
  for (long long it = 0; it < iters; ++it) {   // uniform trip count: the shuffles below need the whole warp
    const long long v = (it * gridDim.x + blockIdx.x) * 64 + (threadIdx.x >> 2);
    const bool ok = v < rows;
    float xv[HD_C], o[HD_C];
    float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) {
      const float4* xp = reinterpret_cast<const float4*>(x + v * HD_C);
#pragma unroll
      for (int t = 0; t < 4; ++t) { const float4 u = xp[t]; xv[4 * t] = u.x; xv[4 * t + 1] = u.y; xv[4 * t + 2] = u.z; xv[4 * t + 3] = u.w; }
      gq = *reinterpret_cast<const float4*>(g + v * 16 + 4 * q);
    } else {
#pragma unroll
      for (int c = 0; c < HD_C; ++c) xv[c] = 0.f;
    }
    const float gv[4] = {gq.x, gq.y, gq.z, gq.w};
#pragma unroll
    for (int c = 0; c < HD_C; ++c) o[c] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ab[j] += gv[j];
      const float4* wr = reinterpret_cast<const float4*>(ws + (4 * q + j) * HD_C);   // 128-bit shared-memory reads: a scalar read per FMA pair was the bound
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const float4 w4 = wr[c4];
        o[4 * c4] = fmaf(gv[j], w4.x, o[4 * c4]); o[4 * c4 + 1] = fmaf(gv[j], w4.y, o[4 * c4 + 1]);
        o[4 * c4 + 2] = fmaf(gv[j], w4.z, o[4 * c4 + 2]); o[4 * c4 + 3] = fmaf(gv[j], w4.w, o[4 * c4 + 3]);
      }
#pragma unroll
      for (int c = 0; c < HD_C; ++c) aw[j][c] = fmaf(gv[j], xv[c], aw[j][c]);
    }
#pragma unroll
    for (int c = 0; c < HD_C; ++c) {
      o[c] += __shfl_xor_sync(0xffffffffu, o[c], 1);
      o[c] += __shfl_xor_sync(0xffffffffu, o[c], 2);
    }
    if (ok) {
      float4 r;
      r.x = q == 0 ? o[0] : (q == 1 ? o[4] : (q == 2 ? o[8] : o[12]));
      r.y = q == 0 ? o[1] : (q == 1 ? o[5] : (q == 2 ? o[9] : o[13]));
      r.z = q == 0 ? o[2] : (q == 1 ? o[6] : (q == 2 ? o[10] : o[14]));
      r.w = q == 0 ? o[3] : (q == 1 ? o[7] : (q == 2 ? o[11] : o[15]));
      *reinterpret_cast<float4*>(dx + v * HD_C + 4 * q) = r;
    }
  }
  // lanes with equal q: xor-reduce over lane bits 2..4, then lanes 0..3 hold the warp totals of quarters 0..3
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int c = 0; c < HD_C; ++c) {
      float t = aw[j][c];
      t += __shfl_xor_sync(0xffffffffu, t, 4); t += __shfl_xor_sync(0xffffffffu, t, 8); t += __shfl_xor_sync(0xffffffffu, t, 16);
      if (lane < 4) red[wid][lane * 68 + j * HD_C + c] = t;
    }
    float t = ab[j];
    t += __shfl_xor_sync(0xffffffffu, t, 4); t += __shfl_xor_sync(0xffffffffu, t, 8); t += __shfl_xor_sync(0xffffffffu, t, 16);
    if (lane < 4) red[wid][lane * 68 + 64 + j] = t;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * 68; i += 256) {
    float t = 0.f;
    for (int wv = 0; wv < 8; ++wv) t += red[wv][i];
    const int qq = i / 68, e = i % 68;
    if (e < 64) atomicAdd(dw + (4 * qq + e / HD_C) * HD_C + (e % HD_C), t);
    else atomicAdd(db + 4 * qq + (e - 64), t);
  }
}
ICL_API int icl_head1x1_bwd(const float* g, const float* x, const float* w, float* dx, float* dw /* zeroed [K][16] */, float* db /* zeroed [K] */,
                            long long rows, int C, int K, void* stream) {
  ICL_REQUIRE(C == HD_C && (K == 1 || K == 2 || K == 4 || K == 16), "head1x1_bwd: C=%d K=%d unsupported (C = 16, K in {1,2,4,16})", C, K);
  const int grid = grid_for(rows, 256, 148 * 4);
  if (K == 16) { head1x1_bwd16_k<<<grid_for(rows, 64, 148 * 4), 256, 0, as_stream(stream)>>>(g, x, w, dx, dw, db, rows); ICL_LAUNCHED("head1x1_bwd"); }
  if (K == 1) head1x1_bwd_k<1><<<grid, 256, 0, as_stream(stream)>>>(g, x, w, dx, dw, db, rows);
  else if (K == 2) head1x1_bwd_k<2><<<grid, 256, 0, as_stream(stream)>>>(g, x, w, dx, dw, db, rows);
  else head1x1_bwd_k<4><<<grid, 256, 0, as_stream(stream)>>>(g, x, w, dx, dw, db, rows);
  ICL_LAUNCHED("head1x1_bwd");
}

// ------------------------------------------------------------------------------------------
// small utilities:  y = alpha*x + beta*y ;  rows scaled by a per-row factor (DropPath)
// ------------------------------------------------------------------------------------------
__global__ void axpby_k(const float* __restrict__ x, float* __restrict__ y, float alpha, float beta, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = alpha * x[i] + (beta == 0.f ? 0.f : beta * y[i]);
}
ICL_API int icl_axpby(const float* x, float* y, float alpha, float beta, long long n, void* stream) {
  axpby_k<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(x, y, alpha, beta, n);
  ICL_LAUNCHED("axpby");
}
// out[r, :] = a[r, :] * sa[r] + b[r, :] * sb[r]   (sa/sb may be null => 1)
__global__ void row_combine_k(const float* __restrict__ a, const float* __restrict__ sa, const float* __restrict__ b,
                              const float* __restrict__ sb, float* __restrict__ out, long long rows, long long cols) {
  const long long n = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    float v = a[i] * (sa ? sa[r] : 1.f);
    if (b) v += b[i] * (sb ? sb[r] : 1.f);
    out[i] = v;
  }
}
ICL_API int icl_row_combine(const float* a, const float* sa, const float* b, const float* sb, float* out, long long rows, long long cols,
                            void* stream) {
  row_combine_k<<<grid_for(rows * cols, 256), 256, 0, as_stream(stream)>>>(a, sa, b, sb, out, rows, cols);
  ICL_LAUNCHED("row_combine");
}
