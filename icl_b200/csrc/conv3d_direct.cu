// fp32 CUDA-core 3x3x3 convolution kernels (stride 1, zero padding 1) on F32CL tensors:
//   * conv3d_direct_fwd : generic forward / data-gradient kernel (any Cin, Cout; two channel-
//     concatenated sources so UnetUp3_CT's cat([skip, up]) is never materialised).  It serves
//     the layers the tcgen05 implicit-GEMM kernel does not take (Cin = 1 stem, channel counts
//     that are not multiples of 16) and is the on-device cross-check for that kernel.
//   * conv3d_wgrad      : weight (+bias) gradient, fp32 FFMA with register-tiled sliding window.
// Reference semantics: nn.Conv3d(k=3, s=1, p=1, bias=True), networks/utils.py:104,107.
#include "common.cuh"
#include <stdlib.h>

// ------------------------------------------------------------------------------------------
// weight repacking:  torch [Cout][Cin][27]  ->  fwd  : wp[tap][ci][co]
//                                              dgrad : wp[26-tap][co][ci]  (flipped, transposed)
// so that the same kernel computes  out[v][n] = sum_{tap,k} in[v + tap - 1][k] * wp[tap][k][n].
// ------------------------------------------------------------------------------------------
__global__ void repack_w_k(const float* __restrict__ w, float* __restrict__ wp, int Cout, int Cin, int dgrad) {
  const int total = Cout * Cin * 27;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % 27, ci = (i / 27) % Cin, co = i / (27 * Cin);
    if (!dgrad) wp[((long long)tap * Cin + ci) * Cout + co] = w[i];
    else wp[((long long)(26 - tap) * Cout + co) * Cin + ci] = w[i];
  }
}
ICL_API int icl_repack_w_f32(const float* w, float* wp, int Cout, int Cin, int dgrad, void* stream) {
  repack_w_k<<<grid_for((long long)Cout * Cin * 27, 256), 256, 0, as_stream(stream)>>>(w, wp, Cout, Cin, dgrad);
  ICL_LAUNCHED("repack_w_f32");
}

// ------------------------------------------------------------------------------------------
// forward.  Block = 256 threads = 2x8x16 output voxels; each thread accumulates NB=16 output
// channels; K loop over input channels in chunks of 8 staged through shared memory
// (x halo tile [8][4][10][18] + weights [27][8][16]).
// ------------------------------------------------------------------------------------------
#define DT_D 2
#define DT_H 8
#define DT_W 16
#define DT_CI 8
#define DT_NB 16
#define DT_HALO ((DT_D + 2) * (DT_H + 2) * (DT_W + 2))

__global__ void __launch_bounds__(256) conv3d_direct_fwd_k(
    const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1, const float* __restrict__ wp,
    const float* __restrict__ bias, float* __restrict__ y, int ldy, int y_coff, double* __restrict__ stats,
    int B, int D, int H, int W, int Cout) {
  __shared__ float xs[DT_CI][DT_HALO + 1];
  __shared__ __align__(16) float ws[27][DT_CI][DT_NB];
  __shared__ float red[8][2 * DT_NB];
  const int Cin = C0 + C1;
  const int tw = (W + DT_W - 1) / DT_W, th = (H + DT_H - 1) / DT_H, td = (D + DT_D - 1) / DT_D;
  int t = blockIdx.x;
  const int bw = t % tw; t /= tw;
  const int bh = t % th; t /= th;
  const int bd = t % td;
  const int b = t / td;
  const int n0 = blockIdx.y * DT_NB;
  const int tid = threadIdx.x;
  const int lw = tid % DT_W, lh = (tid / DT_W) % DT_H, ld = tid / (DT_W * DT_H);
  const int d = bd * DT_D + ld, h = bh * DT_H + lh, w = bw * DT_W + lw;
  const bool valid = d < D && h < H && w < W;

  float acc[DT_NB];
#pragma unroll
  for (int i = 0; i < DT_NB; ++i) acc[i] = 0.f;

  for (int c0 = 0; c0 < Cin; c0 += DT_CI) {
    __syncthreads();
    // stage input halo tile (zero outside the volume / beyond Cin)
    for (int i = tid; i < DT_HALO * DT_CI; i += 256) {
      const int ci = i % DT_CI, pos = i / DT_CI;
      const int pw = pos % (DT_W + 2), ph = (pos / (DT_W + 2)) % (DT_H + 2), pd = pos / ((DT_W + 2) * (DT_H + 2));
      const int gd = bd * DT_D + pd - 1, gh = bh * DT_H + ph - 1, gw = bw * DT_W + pw - 1;
      const int c = c0 + ci;
      float v = 0.f;
      if (c < Cin && gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W) {
        const long long vox = (((long long)b * D + gd) * H + gh) * W + gw;
        v = c < C0 ? x0[vox * C0 + c] : x1[vox * C1 + (c - C0)];
      }
      xs[ci][pos] = v;
    }
    for (int i = tid; i < 27 * DT_CI * DT_NB; i += 256) {
      const int n = i % DT_NB, ci = (i / DT_NB) % DT_CI, tap = i / (DT_NB * DT_CI);
      const int c = c0 + ci, co = n0 + n;
      ws[tap][ci][n] = (c < Cin && co < Cout) ? wp[((long long)tap * Cin + c) * Cout + co] : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int tap = 0; tap < 27; ++tap) {
      const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
      const int pos = ((ld + kd) * (DT_H + 2) + (lh + kh)) * (DT_W + 2) + lw + kw;
#pragma unroll
      for (int ci = 0; ci < DT_CI; ++ci) {
        const float xv = xs[ci][pos];
        const float4* wv = reinterpret_cast<const float4*>(&ws[tap][ci][0]);
#pragma unroll
        for (int q = 0; q < DT_NB / 4; ++q) {
          const float4 w4 = wv[q];
          acc[q * 4 + 0] = fmaf(xv, w4.x, acc[q * 4 + 0]);
          acc[q * 4 + 1] = fmaf(xv, w4.y, acc[q * 4 + 1]);
          acc[q * 4 + 2] = fmaf(xv, w4.z, acc[q * 4 + 2]);
          acc[q * 4 + 3] = fmaf(xv, w4.w, acc[q * 4 + 3]);
        }
      }
    }
  }
  // epilogue: bias, store, per-(b, co) sum / sumsq for InstanceNorm
  const long long vox = (((long long)b * D + d) * H + h) * W + w;
#pragma unroll
  for (int n = 0; n < DT_NB; ++n) {
    const int co = n0 + n;
    float v = acc[n] + ((bias && co < Cout) ? bias[co] : 0.f);
    if (!valid || co >= Cout) v = 0.f;
    else y[vox * ldy + y_coff + co] = v;
    acc[n] = v;
  }
  if (stats) {
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int n = 0; n < DT_NB; ++n) {
      const float s = warp_sum(acc[n]), q = warp_sum(acc[n] * acc[n]);
      if (lane == 0) { red[wid][n] = s; red[wid][DT_NB + n] = q; }
    }
    __syncthreads();
    if (tid < 2 * DT_NB) {
      double tsum = 0.0;
      for (int k = 0; k < 8; ++k) tsum += (double)red[k][tid];
      const int n = tid % DT_NB, co = n0 + n;
      if (co < Cout) atomicAdd(&stats[((long long)b * Cout + co) * 2 + (tid >= DT_NB ? 1 : 0)], tsum);
    }
  }
}

ICL_API int icl_conv3d_direct_fwd(const float* x0, int C0, const float* x1, int C1, const float* wp, const float* bias, float* y,
                                  int ldy, int y_coff, double* stats, int B, int D, int H, int W, int Cout, void* stream) {
  ICL_REQUIRE(C0 > 0 && Cout > 0 && (x1 != nullptr || C1 == 0), "conv3d_direct_fwd: bad channel arguments");
  const long long tiles = (long long)B * cdiv(D, DT_D) * cdiv(H, DT_H) * cdiv(W, DT_W);
  ICL_REQUIRE(tiles < 2147483647LL, "conv3d_direct_fwd: too many tiles");
  dim3 grid((unsigned)tiles, cdiv(Cout, DT_NB));
  conv3d_direct_fwd_k<<<grid, 256, 0, as_stream(stream)>>>(x0, C0, x1, C1, wp, bias, y, ldy, y_coff, stats, B, D, H, W, Cout);
  ICL_LAUNCHED("conv3d_direct_fwd");
}

// ------------------------------------------------------------------------------------------
// weight gradient:  dw[co][ci_off+ci][tap] += sum_{b,v} x[b, v+tap-1, ci] * dy[b, v, co]
// Block = 128 threads = 8 ci x 16 co pairs, each owning all 27 taps in registers.  The block walks
// spatial tiles of 4x8x8 voxels (persistent, strided), staging x (halo) and dy through shared
// memory; along W a 3-wide register window is slid so each new voxel costs 9 x-loads + 1 dy-load
// for 27 FMAs.  One atomicAdd per (pair, tap) per block at the end.
// ------------------------------------------------------------------------------------------
#define WG_D 4
#define WG_H 8
#define WG_W 8
#define WG_CI 8
#define WG_CO 16
#define WG_HALO ((WG_D + 2) * (WG_H + 2) * (WG_W + 2))
#define WG_VOX (WG_D * WG_H * WG_W)

__global__ void __launch_bounds__(128) conv3d_wgrad_k(
    const float* __restrict__ x, int Cx, const float* __restrict__ dy, int Cout, float* __restrict__ dw, int Cin_total, int ci_off,
    float* __restrict__ dbias, int B, int D, int H, int W, int tiles_total) {
  __shared__ float xs[WG_CI][WG_HALO + 1];
  __shared__ float ds[WG_CO][WG_VOX + 1];
  const int tid = threadIdx.x;
  const int ci = tid % WG_CI, co = tid / WG_CI;  // a warp = 8 ci x 4 co
  const int cib = blockIdx.y % cdiv(Cx, WG_CI), cob = blockIdx.y / cdiv(Cx, WG_CI);
  const int c_base = cib * WG_CI, n_base = cob * WG_CO;
  const int tw = cdiv(W, WG_W), th = cdiv(H, WG_H), td = cdiv(D, WG_D);
  float acc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) acc[i] = 0.f;
  float bsum = 0.f;

  for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
    int t = tile;
    const int bw = t % tw; t /= tw;
    const int bh = t % th; t /= th;
    const int bd = t % td;
    const int b = t / td;
    __syncthreads();
    for (int i = tid; i < WG_HALO * WG_CI; i += 128) {
      const int cc = i % WG_CI, pos = i / WG_CI;
      const int pw = pos % (WG_W + 2), ph = (pos / (WG_W + 2)) % (WG_H + 2), pd = pos / ((WG_W + 2) * (WG_H + 2));
      const int gd = bd * WG_D + pd - 1, gh = bh * WG_H + ph - 1, gw = bw * WG_W + pw - 1;
      const int c = c_base + cc;
      float v = 0.f;
      if (c < Cx && gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W)
        v = x[((((long long)b * D + gd) * H + gh) * W + gw) * Cx + c];
      xs[cc][pos] = v;
    }
    for (int i = tid; i < WG_VOX * WG_CO; i += 128) {
      const int n = i % WG_CO, pos = i / WG_CO;
      const int pw = pos % WG_W, ph = (pos / WG_W) % WG_H, pd = pos / (WG_W * WG_H);
      const int gd = bd * WG_D + pd, gh = bh * WG_H + ph, gw = bw * WG_W + pw;
      const int o = n_base + n;
      float v = 0.f;
      if (o < Cout && gd < D && gh < H && gw < W) v = dy[((((long long)b * D + gd) * H + gh) * W + gw) * Cout + o];
      ds[n][pos] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int row = 0; row < WG_D * WG_H; ++row) {
      const int ld = row / WG_H, lh = row % WG_H;
      float win[9][3];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const int base = ((ld + k / 3) * (WG_H + 2) + (lh + k % 3)) * (WG_W + 2);
        win[k][0] = 0.f; win[k][1] = xs[ci][base]; win[k][2] = xs[ci][base + 1];
      }
#pragma unroll
      for (int lw = 0; lw < WG_W; ++lw) {
        const float g = ds[co][row * WG_W + lw];
        bsum += g;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const int base = ((ld + k / 3) * (WG_H + 2) + (lh + k % 3)) * (WG_W + 2);
          win[k][0] = win[k][1]; win[k][1] = win[k][2]; win[k][2] = xs[ci][base + lw + 2];
          acc[k * 3 + 0] = fmaf(win[k][0], g, acc[k * 3 + 0]);
          acc[k * 3 + 1] = fmaf(win[k][1], g, acc[k * 3 + 1]);
          acc[k * 3 + 2] = fmaf(win[k][2], g, acc[k * 3 + 2]);
        }
      }
    }
  }
  const int c = c_base + ci, o = n_base + co;
  if (c < Cx && o < Cout) {
    float* dst = dw + ((long long)o * Cin_total + ci_off + c) * 27;
#pragma unroll
    for (int k = 0; k < 27; ++k) atomicAdd(dst + k, acc[k]);
  }
  if (dbias && cib == 0 && ci == 0 && o < Cout) atomicAdd(dbias + o, bsum);
}

ICL_API int icl_conv3d_wgrad(const float* x, int Cx, const float* dy, int Cout, float* dw, int Cin_total, int ci_off, float* dbias,
                             int B, int D, int H, int W, void* stream) {
  const long long tiles = (long long)B * cdiv(D, WG_D) * cdiv(H, WG_H) * cdiv(W, WG_W);
  ICL_REQUIRE(tiles < 2147483647LL, "conv3d_wgrad: too many tiles");
  const int gy = cdiv(Cx, WG_CI) * cdiv(Cout, WG_CO);
  int gx = (int)min(tiles, (long long)max(1, (148 * 8) / gy));
  conv3d_wgrad_k<<<dim3(gx, gy), 128, 0, as_stream(stream)>>>(x, Cx, dy, Cout, dw, Cin_total, ci_off, dbias, B, D, H, W, (int)tiles);
  ICL_LAUNCHED("conv3d_wgrad");
}

// ------------------------------------------------------------------------------------------
// Stem: Cin = 1 -> Cout = 16 (conv1.conv1 of the backbone at full resolution, networks/utils.py:104 with
// in_size = in_channels = 1).  Bandwidth-bound (4 B in, 64 B out per voxel): the generic kernels above would
// spend 8x the FMAs on zero-padded channels.  Thread = 4 consecutive w voxels x 16 output channels.
// ------------------------------------------------------------------------------------------
#define ST_D 4
#define ST_H 8
#define ST_W 32
#define ST_CO 16
#define ST_HW (ST_W + 2)
#define ST_HALO ((ST_D + 2) * (ST_H + 2) * ST_HW)

__global__ void __launch_bounds__(256) conv3d_stem_fwd_k(const float* __restrict__ x, const float* __restrict__ w /* [16][1][27] */,
                                                         const float* __restrict__ bias, float* __restrict__ y, double* __restrict__ stats,
                                                         int B, int D, int H, int W) {
  __shared__ float xs[ST_HALO];
  __shared__ __align__(16) float ws[27][ST_CO];
  __shared__ float red[8][2 * ST_CO];
  const int tw = cdiv(W, ST_W), th = cdiv(H, ST_H), td = cdiv(D, ST_D);
  int t = blockIdx.x;
  const int bw = t % tw; t /= tw;
  const int bh = t % th; t /= th;
  const int bd = t % td;
  const int b = t / td;
  const int tid = threadIdx.x;
  for (int i = tid; i < 27 * ST_CO; i += 256) ws[i % 27][i / 27] = w[i];  // w[co*27 + tap]
  for (int i = tid; i < ST_HALO; i += 256) {
    const int pw = i % ST_HW, ph = (i / ST_HW) % (ST_H + 2), pd = i / (ST_HW * (ST_H + 2));
    const int gd = bd * ST_D + pd - 1, gh = bh * ST_H + ph - 1, gw = bw * ST_W + pw - 1;
    xs[i] = (gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W) ? x[(((long long)b * D + gd) * H + gh) * W + gw] : 0.f;
  }
  __syncthreads();
  const int lw = (tid % 8) * 4, lh = (tid / 8) % ST_H, ld = tid / (8 * ST_H);
  float acc[4][ST_CO];
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int c = 0; c < ST_CO; ++c) acc[v][c] = 0.f;
#pragma unroll
  for (int k9 = 0; k9 < 9; ++k9) {
    const float* row = &xs[((ld + k9 / 3) * (ST_H + 2) + lh + k9 % 3) * ST_HW + lw];
    float xv[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) xv[i] = row[i];
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const float4* wv = reinterpret_cast<const float4*>(&ws[k9 * 3 + kw][0]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 w4 = wv[q];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          acc[v][q * 4 + 0] = fmaf(xv[v + kw], w4.x, acc[v][q * 4 + 0]);
          acc[v][q * 4 + 1] = fmaf(xv[v + kw], w4.y, acc[v][q * 4 + 1]);
          acc[v][q * 4 + 2] = fmaf(xv[v + kw], w4.z, acc[v][q * 4 + 2]);
          acc[v][q * 4 + 3] = fmaf(xv[v + kw], w4.w, acc[v][q * 4 + 3]);
        }
      }
    }
  }
  const int d = bd * ST_D + ld, h = bh * ST_H + lh, w0 = bw * ST_W + lw;
  float s[ST_CO], q2[ST_CO];
#pragma unroll
  for (int c = 0; c < ST_CO; ++c) { s[c] = 0.f; q2[c] = 0.f; }
  const bool rowok = d < D && h < H;
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    if (rowok && w0 + v < W) {
      float* dst = y + ((((long long)b * D + d) * H + h) * W + w0 + v) * ST_CO;
#pragma unroll
      for (int c = 0; c < ST_CO; ++c) {
        acc[v][c] += bias ? bias[c] : 0.f;
        s[c] += acc[v][c]; q2[c] += acc[v][c] * acc[v][c];
      }
#pragma unroll
      for (int c = 0; c < ST_CO; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(acc[v][c], acc[v][c + 1], acc[v][c + 2], acc[v][c + 3]);
    }
  }
  if (stats) {
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int c = 0; c < ST_CO; ++c) {
      const float a = warp_sum(s[c]), bq = warp_sum(q2[c]);
      if (lane == 0) { red[wid][c] = a; red[wid][ST_CO + c] = bq; }
    }
    __syncthreads();
    if (tid < 2 * ST_CO) {
      double tsum = 0.0;
      for (int k = 0; k < 8; ++k) tsum += (double)red[k][tid];
      atomicAdd(&stats[((long long)b * ST_CO + (tid % ST_CO)) * 2 + (tid >= ST_CO ? 1 : 0)], tsum);
    }
  }
}
// Persistent form of the kernel above (same tile, same per-thread arithmetic in the same order): two CTAs per SM walk the tiles, the
// halo of the NEXT tile arrives by cp.async (zero-fill outside the volume) while the current one is computed, and the InstanceNorm
// statistics are reduced across the warp with a 31-shuffle transposing reduction (lane l ends up with the warp total of value l:
// 16 sums, 16 sums of squares) and accumulated in double precision in shared memory until the sample changes — the one-tile-per-CTA
// kernel spent its time in the load / barrier / 160-shuffle phases around the FMA loop (ncu: 40 % issue utilisation, 29 % FMA pipe).
__device__ __forceinline__ void cp_async4_zfill(float* dst, const float* src, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int n = ok ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__global__ void __launch_bounds__(256, 2) conv3d_stem_fwd2_k(const float* __restrict__ x, const float* __restrict__ w /* [16][1][27] */,
                                                             const float* __restrict__ bias, float* __restrict__ y, double* __restrict__ stats,
                                                             int B, int D, int H, int W, int tiles_total) {
  __shared__ float xs[2][ST_HALO];
  __shared__ __align__(16) float ws[27][ST_CO];
  __shared__ double red[8][32];
  __shared__ float bs[ST_CO];
  const int tw = cdiv(W, ST_W), th = cdiv(H, ST_H), td = cdiv(D, ST_D);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < 27 * ST_CO; i += 256) ws[i % 27][i / 27] = w[i];  // w[co*27 + tap]
  if (tid < ST_CO) bs[tid] = bias ? bias[tid] : 0.f;
  red[wid][lane] = 0.0;
  // this thread's halo slots (the same for every tile): slot k is element tid + 256 k of the [ST_D+2][ST_H+2][ST_HW] halo
  int hpd[8], hph[8], hpw[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = tid + k * 256;
    hpw[k] = i % ST_HW; hph[k] = (i / ST_HW) % (ST_H + 2); hpd[k] = i / (ST_HW * (ST_H + 2));
  }
  auto prefetch = [&](int tile, int buf) {
    int t = tile;
    const int bw = t % tw; t /= tw;
    const int bh = t % th; t /= th;
    const int bd = t % td;
    const int b = t / td;
    const float* xb = x + (long long)b * D * H * W;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = tid + k * 256;
      if (i < ST_HALO) {
        const int gd = bd * ST_D + hpd[k] - 1, gh = bh * ST_H + hph[k] - 1, gw = bw * ST_W + hpw[k] - 1;
        const bool ok = gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W;
        cp_async4_zfill(&xs[buf][i], ok ? xb + ((long long)gd * H + gh) * W + gw : x, ok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto flush = [&](int b) {   // block-uniform call
    __syncthreads();
    if (tid < 32) {
      double tsum = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) tsum += red[k][tid];
      atomicAdd(&stats[((long long)b * ST_CO + (tid & 15)) * 2 + (tid >> 4)], tsum);
    }
    __syncthreads();
    red[wid][lane] = 0.0;
  };
  const int lw = (tid % 8) * 4, lh = (tid / 8) % ST_H, ld = tid / (8 * ST_H);
  int buf = 0, cur_b = -1;
  if ((int)blockIdx.x < tiles_total) prefetch(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x, buf ^= 1) {
    const int next = tile + gridDim.x;
    if (next < tiles_total) prefetch(next, buf ^ 1);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    int t = tile;
    const int bw = t % tw; t /= tw;
    const int bh = t % th; t /= th;
    const int bd = t % td;
    const int b = t / td;
    if (b != cur_b) {
      if (cur_b >= 0 && stats) flush(cur_b);
      cur_b = b;
    }
    float acc[4][ST_CO];
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int c = 0; c < ST_CO; ++c) acc[v][c] = 0.f;
#pragma unroll
    for (int k9 = 0; k9 < 9; ++k9) {
      const float* row = &xs[buf][((ld + k9 / 3) * (ST_H + 2) + lh + k9 % 3) * ST_HW + lw];
      float xv[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) xv[i] = row[i];
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float4* wv = reinterpret_cast<const float4*>(&ws[k9 * 3 + kw][0]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w4 = wv[q];
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            acc[v][q * 4 + 0] = fmaf(xv[v + kw], w4.x, acc[v][q * 4 + 0]);
            acc[v][q * 4 + 1] = fmaf(xv[v + kw], w4.y, acc[v][q * 4 + 1]);
            acc[v][q * 4 + 2] = fmaf(xv[v + kw], w4.z, acc[v][q * 4 + 2]);
            acc[v][q * 4 + 3] = fmaf(xv[v + kw], w4.w, acc[v][q * 4 + 3]);
          }
        }
      }
    }
    const int d = bd * ST_D + ld, h = bh * ST_H + lh, w0 = bw * ST_W + lw;
    float sv[2 * ST_CO];
#pragma unroll
    for (int c = 0; c < 2 * ST_CO; ++c) sv[c] = 0.f;
    const bool rowok = d < D && h < H;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      if (rowok && w0 + v < W) {
        float* dst = y + ((((long long)b * D + d) * H + h) * W + w0 + v) * ST_CO;
#pragma unroll
        for (int c = 0; c < ST_CO; ++c) {
          acc[v][c] += bs[c];
          sv[c] += acc[v][c]; sv[ST_CO + c] += acc[v][c] * acc[v][c];
        }
#pragma unroll
        for (int c = 0; c < ST_CO; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(acc[v][c], acc[v][c + 1], acc[v][c + 2], acc[v][c + 3]);
      }
    }
    if (stats) {
      // transposing reduction: after the step with offset o, slot j of a lane holds value j + (lane & (32 - o)) summed over the
      // lanes that differ from it in bits >= o; 16 + 8 + 4 + 2 + 1 shuffles leave the warp total of value `lane` in slot 0
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < off; ++j) {
          const float send = up ? sv[j] : sv[j + off];
          const float keep = up ? sv[j + off] : sv[j];
          sv[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      red[wid][lane] += (double)sv[0];
    }
    __syncthreads();   // every warp is done with xs[buf] before the next iteration's prefetch overwrites it
  }
  if (cur_b >= 0 && stats) flush(cur_b);
}
ICL_API int icl_conv3d_stem_fwd(const float* x, const float* w, const float* bias, float* y, double* stats, int B, int D, int H, int W, int Cout,
                                void* stream) {
  ICL_REQUIRE(Cout == ST_CO, "conv3d_stem_fwd: Cout=%d (only 16 is built)", Cout);
  const long long tiles = (long long)B * cdiv(D, ST_D) * cdiv(H, ST_H) * cdiv(W, ST_W);
  ICL_REQUIRE(tiles < 2147483647LL, "conv3d_stem_fwd: too many tiles");
  if (getenv("ICL_STEM_V1") == nullptr) {
    const int grid = (int)min(tiles, (long long)148 * 2);
    conv3d_stem_fwd2_k<<<grid, 256, 0, as_stream(stream)>>>(x, w, bias, y, stats, B, D, H, W, (int)tiles);
    ICL_LAUNCHED("conv3d_stem_fwd");
  }
  conv3d_stem_fwd_k<<<(unsigned)tiles, 256, 0, as_stream(stream)>>>(x, w, bias, y, stats, B, D, H, W);
  ICL_LAUNCHED("conv3d_stem_fwd");
}

#define SW_D 2
#define SW_HALO ((SW_D + 2) * (ST_H + 2) * ST_HW)
// stem weight gradient: dw[co][0][tap] = sum_v dy[v][co] * x[v + tap - 1].  Thread = (4 co) x (kd,kh) x (3 kw) = 12 accumulators,
// 36 threads cover the 16 x 27 outputs, 3 such groups split a tile's rows; a 3-wide register window slides along w.
__global__ void __launch_bounds__(128) conv3d_stem_wgrad_k(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                                                           int B, int D, int H, int W, int tiles_total) {
  __shared__ float xs[SW_HALO];
  __shared__ __align__(16) float ds[SW_D * ST_H * ST_W][ST_CO];
  const int tid = threadIdx.x;
  const int co4 = tid % 4, k9 = (tid / 4) % 9, part = tid / 36;  // part 3 (threads 108..127) only helps staging
  const int tw = cdiv(W, ST_W), th = cdiv(H, ST_H), td = cdiv(D, SW_D);
  float acc[3][4];
#pragma unroll
  for (int i = 0; i < 3; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
    int t = tile;
    const int bw = t % tw; t /= tw;
    const int bh = t % th; t /= th;
    const int bd = t % td;
    const int b = t / td;
    __syncthreads();
    for (int i = tid; i < SW_HALO; i += 128) {
      const int pw = i % ST_HW, ph = (i / ST_HW) % (ST_H + 2), pd = i / (ST_HW * (ST_H + 2));
      const int gd = bd * SW_D + pd - 1, gh = bh * ST_H + ph - 1, gw = bw * ST_W + pw - 1;
      xs[i] = (gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W) ? x[(((long long)b * D + gd) * H + gh) * W + gw] : 0.f;
    }
    for (int i = tid; i < SW_D * ST_H * ST_W * 4; i += 128) {
      const int q = i % 4, pos = i / 4;
      const int pw = pos % ST_W, ph = (pos / ST_W) % ST_H, pd = pos / (ST_W * ST_H);
      const int gd = bd * SW_D + pd, gh = bh * ST_H + ph, gw = bw * ST_W + pw;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gd < D && gh < H && gw < W) v = *reinterpret_cast<const float4*>(dy + ((((long long)b * D + gd) * H + gh) * W + gw) * ST_CO + q * 4);
      *reinterpret_cast<float4*>(&ds[pos][q * 4]) = v;
    }
    __syncthreads();
    if (part < 3) {
      for (int row = part; row < SW_D * ST_H; row += 3) {
        const int ld = row / ST_H, lh = row % ST_H;
        const float* xr = &xs[((ld + k9 / 3) * (ST_H + 2) + lh + k9 % 3) * ST_HW];
        float x0 = xr[0], x1 = xr[1];
#pragma unroll 4
        for (int lw = 0; lw < ST_W; ++lw) {
          const float x2 = xr[lw + 2];
          const float4 g = *reinterpret_cast<const float4*>(&ds[row * ST_W + lw][co4 * 4]);
          acc[0][0] = fmaf(x0, g.x, acc[0][0]); acc[0][1] = fmaf(x0, g.y, acc[0][1]); acc[0][2] = fmaf(x0, g.z, acc[0][2]); acc[0][3] = fmaf(x0, g.w, acc[0][3]);
          acc[1][0] = fmaf(x1, g.x, acc[1][0]); acc[1][1] = fmaf(x1, g.y, acc[1][1]); acc[1][2] = fmaf(x1, g.z, acc[1][2]); acc[1][3] = fmaf(x1, g.w, acc[1][3]);
          acc[2][0] = fmaf(x2, g.x, acc[2][0]); acc[2][1] = fmaf(x2, g.y, acc[2][1]); acc[2][2] = fmaf(x2, g.z, acc[2][2]); acc[2][3] = fmaf(x2, g.w, acc[2][3]);
          x0 = x1; x1 = x2;
        }
      }
    }
  }
  if (part < 3) {
#pragma unroll
    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
      for (int c = 0; c < 4; ++c) atomicAdd(dw + (co4 * 4 + c) * 27 + k9 * 3 + kw, acc[kw][c]);
  }
}
// The same kernel with the NEXT tile's operands (x halo, dy tile) arriving by cp.async into a second buffer while the current tile is
// accumulated (the synchronous loads left the FMA pipe idle two thirds of the time: ncu 33 % FMA, 21 % warps active).
#define SW2_BUF_FLOATS (SW_HALO + SW_D * ST_H * ST_W * ST_CO)
__device__ __forceinline__ void cp_async16_zfill(float* dst, const float* src, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int n = ok ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__global__ void __launch_bounds__(128) conv3d_stem_wgrad2_k(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                                                            int B, int D, int H, int W, int tiles_total) {
  extern __shared__ __align__(16) float sw2[];   // [2][ ds: SW_D*ST_H*ST_W x 16 | xs: SW_HALO ]
  const int tid = threadIdx.x;
  const int co4 = tid % 4, k9 = (tid / 4) % 9, part = tid / 36;  // part 3 (threads 108..127) only helps staging
  const int tw = cdiv(W, ST_W), th = cdiv(H, ST_H), td = cdiv(D, SW_D);
  float acc[3][4];
#pragma unroll
  for (int i = 0; i < 3; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  auto prefetch = [&](int tile, int buf) {
    float* ds = sw2 + (size_t)buf * SW2_BUF_FLOATS;
    float* xs = ds + SW_D * ST_H * ST_W * ST_CO;
    int t = tile;
    const int bw = t % tw; t /= tw;
    const int bh = t % th; t /= th;
    const int bd = t % td;
    const int b = t / td;
    for (int i = tid; i < SW_HALO; i += 128) {
      const int pw = i % ST_HW, ph = (i / ST_HW) % (ST_H + 2), pd = i / (ST_HW * (ST_H + 2));
      const int gd = bd * SW_D + pd - 1, gh = bh * ST_H + ph - 1, gw = bw * ST_W + pw - 1;
      const bool ok = gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W;
      cp_async4_zfill(&xs[i], ok ? x + (((long long)b * D + gd) * H + gh) * W + gw : x, ok);
    }
    for (int i = tid; i < SW_D * ST_H * ST_W * 4; i += 128) {
      const int q = i % 4, pos = i / 4;
      const int pw = pos % ST_W, ph = (pos / ST_W) % ST_H, pd = pos / (ST_W * ST_H);
      const int gd = bd * SW_D + pd, gh = bh * ST_H + ph, gw = bw * ST_W + pw;
      const bool ok = gd < D && gh < H && gw < W;
      cp_async16_zfill(&ds[pos * ST_CO + q * 4], ok ? dy + ((((long long)b * D + gd) * H + gh) * W + gw) * ST_CO + q * 4 : dy, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int buf = 0;
  if ((int)blockIdx.x < tiles_total) prefetch(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x, buf ^= 1) {
    const int next = tile + gridDim.x;
    if (next < tiles_total) prefetch(next, buf ^ 1);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const float* ds = sw2 + (size_t)buf * SW2_BUF_FLOATS;
    const float* xs = ds + SW_D * ST_H * ST_W * ST_CO;
    if (part < 3) {
      for (int row = part; row < SW_D * ST_H; row += 3) {
        const int ld = row / ST_H, lh = row % ST_H;
        const float* xr = &xs[((ld + k9 / 3) * (ST_H + 2) + lh + k9 % 3) * ST_HW];
        float x0 = xr[0], x1 = xr[1];
#pragma unroll 4
        for (int lw = 0; lw < ST_W; ++lw) {
          const float x2 = xr[lw + 2];
          const float4 g = *reinterpret_cast<const float4*>(&ds[(row * ST_W + lw) * ST_CO + co4 * 4]);
          acc[0][0] = fmaf(x0, g.x, acc[0][0]); acc[0][1] = fmaf(x0, g.y, acc[0][1]); acc[0][2] = fmaf(x0, g.z, acc[0][2]); acc[0][3] = fmaf(x0, g.w, acc[0][3]);
          acc[1][0] = fmaf(x1, g.x, acc[1][0]); acc[1][1] = fmaf(x1, g.y, acc[1][1]); acc[1][2] = fmaf(x1, g.z, acc[1][2]); acc[1][3] = fmaf(x1, g.w, acc[1][3]);
          acc[2][0] = fmaf(x2, g.x, acc[2][0]); acc[2][1] = fmaf(x2, g.y, acc[2][1]); acc[2][2] = fmaf(x2, g.z, acc[2][2]); acc[2][3] = fmaf(x2, g.w, acc[2][3]);
          x0 = x1; x1 = x2;
        }
      }
    }
    __syncthreads();   // done with this buffer before the next iteration's prefetch refills it
  }
  if (part < 3) {
#pragma unroll
    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
      for (int c = 0; c < 4; ++c) atomicAdd(dw + (co4 * 4 + c) * 27 + k9 * 3 + kw, acc[kw][c]);
  }
}
ICL_API int icl_conv3d_stem_wgrad(const float* x, const float* dy, float* dw /* zeroed [16][1][27] */, int B, int D, int H, int W, int Cout,
                                  void* stream) {
  ICL_REQUIRE(Cout == ST_CO, "conv3d_stem_wgrad: Cout=%d (only 16 is built)", Cout);
  const long long tiles = (long long)B * cdiv(D, SW_D) * cdiv(H, ST_H) * cdiv(W, ST_W);
  ICL_REQUIRE(tiles < 2147483647LL, "conv3d_stem_wgrad: too many tiles");
  if (getenv("ICL_STEM_V1") == nullptr) {
    const size_t smem = 2 * (size_t)SW2_BUF_FLOATS * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
      ICL_REQUIRE(cudaFuncSetAttribute(conv3d_stem_wgrad2_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess,
                  "conv3d_stem_wgrad: cannot raise the dynamic shared-memory limit");
      attr_done = true;
    }
    const int grid = (int)min(tiles, (long long)148 * 2);
    conv3d_stem_wgrad2_k<<<grid, 128, smem, as_stream(stream)>>>(x, dy, dw, B, D, H, W, (int)tiles);
    ICL_LAUNCHED("conv3d_stem_wgrad");
  }
  const int grid = (int)min(tiles, (long long)148 * 4);
  conv3d_stem_wgrad_k<<<grid, 128, 0, as_stream(stream)>>>(x, dy, dw, B, D, H, W, (int)tiles);
  ICL_LAUNCHED("conv3d_stem_wgrad");
}
