// (Shifted-)window multi-head self-attention of the Swin blocks, forward and backward — SURVEY §8 row a20:
// WindowAttention.forward (networks/swinunet_icl.py:120-155) together with the cyclic shift, window partition / reverse and
// the shift mask that SwinTransformerBlock.forward wraps around it (:249-293, mask construction :217-245).
//
// One CTA per (window, head).  Everything the reference does with roll / view / permute copies is address arithmetic here:
// the CTA gathers its ws*ws tokens from the token-major qkv tensor [B, H*W, 3*C] at their ORIGINAL (un-rolled) positions
// and scatters the result back to the same positions of [B, H*W, C], so no rolled or windowed copy of the activations is
// ever written.  The shift mask (-100 between tokens of different wrap-around regions) and the relative-position bias are
// evaluated from coordinates; neither the [nW, N, N] mask buffer nor the [nH, N, N] gathered bias exists on the device.
//
// Shapes on this path: ws = 7 (N = 49 tokens), head_dim = 32, C = 96..768.  49 x 49 x 32 tiles are far below a tcgen05 MMA
// atom (M >= 64) — the work per CTA is 0.6 MFLOP against 19 KB of operands — so the kernel is a shared-memory fp32 kernel
// bound by the qkv / output streams (algorithmic bytes: forward 16 B per token-channel, backward 32 B).
#include "common.cuh"

namespace {

constexpr int WA_HD = 32;        // head_dim (C / num_heads) on every Swin stage of the reference configs
constexpr int WA_MAXN = 64;      // tokens per window (ws <= 8)
constexpr int WA_LD = WA_HD + 1; // padded row: conflict-free column walks
constexpr int WA_THREADS = 128;

struct WaGeom {
  int B, H, W, C, nH, ws, shift, nWx, nWy;
  float scale;
};

// token t of window (wy, wx): row offset (in tokens) of its original position inside the sample, and its mask region.
__device__ __forceinline__ void wa_token(const WaGeom& g, int wy, int wx, int t, int& pos, int& region) {
  const int i = t / g.ws, j = t - i * g.ws;
  const int ys = wy * g.ws + i, xs = wx * g.ws + j;  // coordinates in the rolled frame
  int y = ys + g.shift, x = xs + g.shift;            // torch.roll(x, -shift): rolled[ys] = x[(ys + shift) mod H]
  if (y >= g.H) y -= g.H;
  if (x >= g.W) x -= g.W;
  pos = y * g.W + x;
  const int ry = ys < g.H - g.ws ? 0 : (ys < g.H - g.shift ? 1 : 2);
  const int rx = xs < g.W - g.ws ? 0 : (xs < g.W - g.shift ? 1 : 2);
  region = g.shift > 0 ? ry * 3 + rx : 0;
}

__device__ __forceinline__ int wa_rel(const WaGeom& g, int ti, int tj) {
  const int yi = ti / g.ws, xi = ti - yi * g.ws, yj = tj / g.ws, xj = tj - yj * g.ws;
  return (yi - yj + g.ws - 1) * (2 * g.ws - 1) + (xi - xj + g.ws - 1);
}

// loads q (pre-scaled), k, v rows of this (window, head) into shared memory; spos / sreg get the token positions / regions
__device__ __forceinline__ void wa_load(const WaGeom& g, const float* __restrict__ qkv, int b, int wy, int wx, int h, int N,
                                        float (*sq)[WA_LD], float (*sk)[WA_LD], float (*sv)[WA_LD], int* spos, int* sreg) {
  for (int t = threadIdx.x; t < N; t += blockDim.x) wa_token(g, wy, wx, t, spos[t], sreg[t]);
  __syncthreads();
  const long long base = (long long)b * g.H * g.W;
  for (int idx = threadIdx.x; idx < N * WA_HD; idx += blockDim.x) {
    const int t = idx >> 5, d = idx & 31;
    const float* row = qkv + (base + spos[t]) * (3LL * g.C) + h * WA_HD + d;
    sq[t][d] = row[0] * g.scale;
    sk[t][d] = row[g.C];
    sv[t][d] = row[2 * g.C];
  }
  __syncthreads();
}

// S = q k^T + bias + mask, then row softmax in place (one warp per row)
__device__ __forceinline__ void wa_probs(const WaGeom& g, const float* __restrict__ table, int h, int N, float (*sq)[WA_LD],
                                         float (*sk)[WA_LD], const int* sreg, float (*sp)[WA_MAXN + 1]) {
  for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
    const int i = idx / N, j = idx - i * N;
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < WA_HD; ++d) acc = fmaf(sq[i][d], sk[j][d], acc);
    acc += table[wa_rel(g, i, j) * g.nH + h];
    if (sreg[i] != sreg[j]) acc += -100.0f;
    sp[i][j] = acc;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = wid; i < N; i += nw) {
    const float a0 = lane < N ? sp[i][lane] : -INFINITY, a1 = lane + 32 < N ? sp[i][lane + 32] : -INFINITY;
    const float m = warp_max(fmaxf(a0, a1));
    const float e0 = lane < N ? expf(a0 - m) : 0.f, e1 = lane + 32 < N ? expf(a1 - m) : 0.f;
    const float inv = 1.f / warp_sum(e0 + e1);
    if (lane < N) sp[i][lane] = e0 * inv;
    if (lane + 32 < N) sp[i][lane + 32] = e1 * inv;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(WA_THREADS) window_attn_fwd_k(const float* __restrict__ qkv, const float* __restrict__ table,
                                                                float* __restrict__ out, WaGeom g) {
  __shared__ float sq[WA_MAXN][WA_LD], sk[WA_MAXN][WA_LD], sv[WA_MAXN][WA_LD];
  __shared__ float sp[WA_MAXN][WA_MAXN + 1];
  __shared__ int spos[WA_MAXN], sreg[WA_MAXN];
  const int N = g.ws * g.ws, h = blockIdx.y;
  const int win = blockIdx.x % (g.nWy * g.nWx), b = blockIdx.x / (g.nWy * g.nWx);
  const int wy = win / g.nWx, wx = win - wy * g.nWx;
  wa_load(g, qkv, b, wy, wx, h, N, sq, sk, sv, spos, sreg);
  wa_probs(g, table, h, N, sq, sk, sreg, sp);
  const long long base = (long long)b * g.H * g.W;
  for (int idx = threadIdx.x; idx < N * WA_HD; idx += blockDim.x) {
    const int i = idx >> 5, d = idx & 31;
    float acc = 0.f;
    for (int j = 0; j < N; ++j) acc = fmaf(sp[i][j], sv[j][d], acc);
    out[(base + spos[i]) * g.C + h * WA_HD + d] = acc;
  }
}

__global__ void __launch_bounds__(WA_THREADS) window_attn_bwd_k(const float* __restrict__ qkv, const float* __restrict__ table,
                                                                const float* __restrict__ dout, float* __restrict__ dqkv,
                                                                float* __restrict__ dtable, WaGeom g) {
  extern __shared__ float smem[];
  float (*sq)[WA_LD] = reinterpret_cast<float (*)[WA_LD]>(smem);
  float (*sk)[WA_LD] = sq + WA_MAXN;
  float (*sv)[WA_LD] = sk + WA_MAXN;
  float (*sdo)[WA_LD] = sv + WA_MAXN;
  float (*sp)[WA_MAXN + 1] = reinterpret_cast<float (*)[WA_MAXN + 1]>(sdo + WA_MAXN);
  float (*sds)[WA_MAXN + 1] = sp + WA_MAXN;
  float* sbin = reinterpret_cast<float*>(sds + WA_MAXN);  // (2 ws - 1)^2 <= 225 bins
  int* spos = reinterpret_cast<int*>(sbin + 225);
  int* sreg = spos + WA_MAXN;
  const int N = g.ws * g.ws, h = blockIdx.y, nbins = (2 * g.ws - 1) * (2 * g.ws - 1);
  const int win = blockIdx.x % (g.nWy * g.nWx), b = blockIdx.x / (g.nWy * g.nWx);
  const int wy = win / g.nWx, wx = win - wy * g.nWx;
  for (int i = threadIdx.x; i < nbins; i += blockDim.x) sbin[i] = 0.f;
  wa_load(g, qkv, b, wy, wx, h, N, sq, sk, sv, spos, sreg);
  const long long base = (long long)b * g.H * g.W;
  for (int idx = threadIdx.x; idx < N * WA_HD; idx += blockDim.x) {
    const int t = idx >> 5, d = idx & 31;
    sdo[t][d] = dout[(base + spos[t]) * g.C + h * WA_HD + d];
  }
  wa_probs(g, table, h, N, sq, sk, sreg, sp);  // its barriers also publish sdo
  // dP = dO v^T
  for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
    const int i = idx / N, j = idx - i * N;
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < WA_HD; ++d) acc = fmaf(sdo[i][d], sv[j][d], acc);
    sds[i][j] = acc;
  }
  __syncthreads();
  // dS = P * (dP - sum_j P dP); the bias gradient is dS binned by relative position
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = wid; i < N; i += nw) {
    const float p0 = lane < N ? sp[i][lane] : 0.f, p1 = lane + 32 < N ? sp[i][lane + 32] : 0.f;
    const float d0 = lane < N ? sds[i][lane] : 0.f, d1 = lane + 32 < N ? sds[i][lane + 32] : 0.f;
    const float r = warp_sum(p0 * d0 + p1 * d1);
    if (lane < N) {
      const float v = p0 * (d0 - r);
      sds[i][lane] = v;
      atomicAdd(&sbin[wa_rel(g, i, lane)], v);
    }
    if (lane + 32 < N) {
      const float v = p1 * (d1 - r);
      sds[i][lane + 32] = v;
      atomicAdd(&sbin[wa_rel(g, i, lane + 32)], v);
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < N * WA_HD; idx += blockDim.x) {
    const int t = idx >> 5, d = idx & 31;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int j = 0; j < N; ++j) {
      dq = fmaf(sds[t][j], sk[j][d], dq);   // dq_scaled = dS k
      dk = fmaf(sds[j][t], sq[j][d], dk);   // dk = dS^T (q * scale)
      dv = fmaf(sp[j][t], sdo[j][d], dv);   // dv = P^T dO
    }
    float* row = dqkv + (base + spos[t]) * (3LL * g.C) + h * WA_HD + d;
    row[0] = dq * g.scale;
    row[g.C] = dk;
    row[2 * g.C] = dv;
  }
  for (int i = threadIdx.x; i < nbins; i += blockDim.x) atomicAdd(&dtable[i * g.nH + h], sbin[i]);
}

constexpr size_t WA_BWD_SMEM = (4 * WA_MAXN * WA_LD + 2 * WA_MAXN * (WA_MAXN + 1) + 225) * sizeof(float) + 2 * WA_MAXN * sizeof(int);

int wa_geom(WaGeom& g, int B, int H, int W, int C, int nH, int ws, int shift) {
  ICL_REQUIRE(B > 0 && H > 0 && W > 0 && nH > 0 && C == nH * WA_HD, "window_attn: head_dim must be %d (C=%d, heads=%d)", WA_HD, C, nH);
  ICL_REQUIRE(ws >= 1 && ws * ws <= WA_MAXN && H % ws == 0 && W % ws == 0, "window_attn: window %d does not tile %dx%d (<= %d tokens)", ws, H, W,
              WA_MAXN);
  ICL_REQUIRE(shift >= 0 && shift < ws, "window_attn: shift %d must be in [0, %d)", shift, ws);
  g.B = B; g.H = H; g.W = W; g.C = C; g.nH = nH; g.ws = ws; g.shift = shift; g.nWy = H / ws; g.nWx = W / ws;
  g.scale = 1.0f / sqrtf((float)WA_HD);
  return 0;
}

}  // namespace

ICL_API int icl_window_attn_fwd(const float* qkv, const float* table, float* out, int B, int H, int W, int C, int nH, int ws, int shift,
                                void* stream) {
  WaGeom g;
  if (wa_geom(g, B, H, W, C, nH, ws, shift)) return -1;
  window_attn_fwd_k<<<dim3(B * g.nWy * g.nWx, nH), WA_THREADS, 0, as_stream(stream)>>>(qkv, table, out, g);
  ICL_LAUNCHED("window_attn_fwd");
}

ICL_API int icl_window_attn_bwd(const float* qkv, const float* table, const float* dout, float* dqkv, float* dtable /* zeroed */, int B, int H,
                                int W, int C, int nH, int ws, int shift, void* stream) {
  WaGeom g;
  if (wa_geom(g, B, H, W, C, nH, ws, shift)) return -1;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(window_attn_bwd_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WA_BWD_SMEM);
    attr = true;
  }
  window_attn_bwd_k<<<dim3(B * g.nWy * g.nWx, nH), WA_THREADS, WA_BWD_SMEM, as_stream(stream)>>>(qkv, table, dout, dqkv, dtable, g);
  ICL_LAUNCHED("window_attn_bwd");
}
