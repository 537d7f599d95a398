// Loss-side reductions of the ICL training step plus fused SGD (a16)
// and the on-device sliding-window accumulation / argmax / Dice counts (a17-a18).
//
// One streaming kernel family ("class statistics") covers CrossEntropy + DiceLoss on the main
// logits, AuxLoss3D (trilinear interpolation of a coarse class map to the label grid fused into
// the load, so the [B,K,96^3] upsampled tensor is never materialised) and PseudoSoftLoss3D
// (softmax-Dice against detached soft targets).  Per voxel: softmax over K in registers, then
// per-class partial sums reduced warp -> block -> one double atomic per block.
// Reference: utils/losses.py:22-30,42-59,68-90,195-231,254-299.
#include "common.cuh"
#include <stdlib.h>

struct SrcGeom {
  int planar;          // 1: [B,K,rz,ry,rx]   0: channels-last [B,rz,ry,rx,K]
  int rz, ry, rx;      // source grid
  int Z, Y, X;         // loss grid (labels / targets)
  float sz, sy, sx;    // rz/Z ...
};

__device__ __forceinline__ void lin_src(int j, float scale, int n, int& i0, int& i1, float& l1) {
  float s = scale * (j + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > n - 1) i0 = n - 1;
  i1 = min(i0 + 1, n - 1);
  l1 = s - (float)i0;
}

template <int KT>
__device__ __forceinline__ void load_logits(const float* __restrict__ src, const SrcGeom& g, int K, int b, int z, int y, int x, float* L) {
  if (g.rz == g.Z && g.ry == g.Y && g.rx == g.X) {
    const long long v = ((long long)z * g.Y + y) * g.X + x, S = (long long)g.Z * g.Y * g.X;
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if (k < K) L[k] = g.planar ? src[((long long)b * K + k) * S + v] : src[((long long)b * S + v) * K + k];
    return;
  }
  int z0, z1, y0, y1, x0, x1; float lz, ly, lx;
  lin_src(z, g.sz, g.rz, z0, z1, lz); lin_src(y, g.sy, g.ry, y0, y1, ly); lin_src(x, g.sx, g.rx, x0, x1, lx);
  const long long S = (long long)g.rz * g.ry * g.rx;
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    if (k >= K) break;
    const float* p = g.planar ? src + ((long long)b * K + k) * S : src + (long long)b * S * K + k;
    const long long st = g.planar ? 1 : K;
#define AT(zz, yy, xx) p[(((long long)(zz) * g.ry + (yy)) * g.rx + (xx)) * st]
    const float v00 = (1.f - lx) * AT(z0, y0, x0) + lx * AT(z0, y0, x1);
    const float v01 = (1.f - lx) * AT(z0, y1, x0) + lx * AT(z0, y1, x1);
    const float v10 = (1.f - lx) * AT(z1, y0, x0) + lx * AT(z1, y0, x1);
    const float v11 = (1.f - lx) * AT(z1, y1, x0) + lx * AT(z1, y1, x1);
#undef AT
    L[k] = (1.f - lz) * ((1.f - ly) * v00 + ly * v01) + lz * ((1.f - ly) * v10 + ly * v11);
  }
}

template <int KT>
__device__ __forceinline__ float softmax_inplace(float* L, int K, float& mx, float& lse) {
  mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < KT; ++k) if (k < K) mx = fmaxf(mx, L[k]);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < KT; ++k) if (k < K) { L[k] = expf(L[k] - mx); s += L[k]; }
  const float inv = 1.f / s;
#pragma unroll
  for (int k = 0; k < KT; ++k) if (k < K) L[k] *= inv;
  lse = logf(s);
  return inv;
}

// sums layout (double): [0]=ce_sum, then per class k: [1+3k]=A (inter), [2+3k]=Bp (sum p^2 | sum p), [3+3k]=Cq (count | sum q)
template <int KT>
__global__ void __launch_bounds__(256) class_stats_fwd_k(const float* __restrict__ src, SrcGeom g, int B, int K,
                                                         const long long* __restrict__ labels, const float* __restrict__ tgt,
                                                         int is_prob, double* __restrict__ sums) {
  __shared__ float red[8][3 * KT + 1];
  float acc[3 * KT + 1];
#pragma unroll
  for (int i = 0; i < 3 * KT + 1; ++i) acc[i] = 0.f;
  const long long S = (long long)g.Z * g.Y * g.X, total = (long long)B * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long v = i % S; const int b = (int)(i / S);
    const int x = (int)(v % g.X); v /= g.X;
    const int y = (int)(v % g.Y); const int z = (int)(v / g.Y);
    float L[KT];
    load_logits<KT>(src, g, K, b, z, y, x, L);
    float mx, lse;
    if (labels) {
      const int lab = (int)labels[i];
      float raw = 0.f;
#pragma unroll
      for (int k = 0; k < KT; ++k) if (k == lab) raw = L[k];
      if (!is_prob) {
        softmax_inplace<KT>(L, K, mx, lse);
        acc[0] += (mx + lse) - raw;
      }
#pragma unroll
      for (int k = 0; k < KT; ++k) if (k < K) {
        const float t = (k == lab) ? 1.f : 0.f;
        acc[1 + 3 * k] += L[k] * t; acc[2 + 3 * k] += L[k] * L[k]; acc[3 + 3 * k] += t;
      }
    } else {
      float Q[KT];
#pragma unroll
      for (int k = 0; k < KT; ++k) if (k < K) Q[k] = tgt[i * K + k];
      softmax_inplace<KT>(L, K, mx, lse);
      softmax_inplace<KT>(Q, K, mx, lse);
#pragma unroll
      for (int k = 0; k < KT; ++k) if (k < K) { acc[1 + 3 * k] += L[k] * Q[k]; acc[2 + 3 * k] += L[k]; acc[3 + 3 * k] += Q[k]; }
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 3 * KT + 1; ++i) {
    const float s = warp_sum(acc[i]);
    if (lane == 0) red[wid][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < 3 * K + 1) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += (double)red[w][threadIdx.x];
    atomicAdd(&sums[threadIdx.x], t);
  }
}

// Finalize: out[0] = CE mean (label mode) ; out[1] = mean_k (1 - (2A+s)/(Bp+Cq+s)).
__global__ void class_stats_finalize_k(const double* __restrict__ sums, int K, double n_vox, const float* __restrict__ class_w,
                                       float* __restrict__ out) {
  if (threadIdx.x == 0) {
    double d = 0.0;
    for (int k = 0; k < K; ++k)
      d += (class_w ? (double)class_w[k] : 1.0) * (1.0 - (2.0 * sums[1 + 3 * k] + 1e-5) / (sums[2 + 3 * k] + sums[3 + 3 * k] + 1e-5));
    out[0] = (float)(sums[0] / n_vox);
    out[1] = (float)(d / K);
  }
}

// Gradient of the class-statistics losses w.r.t. the (interpolated) logits of ONE loss-grid voxel.
template <int KT>
__device__ __forceinline__ void voxel_dlogits(const float* __restrict__ src, const SrcGeom& g, int K, int b, int z, int y, int x, long long i,
                                              const long long* __restrict__ labels, const float* __restrict__ tgt, int is_prob,
                                              const float* cA, const float* cB, float cCE, float* dl) {
  float L[KT];
  load_logits<KT>(src, g, K, b, z, y, x, L);
  float mx, lse;
  if (!is_prob) softmax_inplace<KT>(L, K, mx, lse);
  float dot = 0.f;
  if (labels) {
    const int lab = (int)labels[i];
#pragma unroll
    for (int k = 0; k < KT; ++k) if (k < K) {
      dl[k] = cA[k] * ((k == lab) ? 1.f : 0.f) + cB[k] * 2.f * L[k];
      dot += dl[k] * L[k];
    }
    if (!is_prob) {
#pragma unroll
      for (int k = 0; k < KT; ++k) if (k < K) dl[k] = L[k] * (dl[k] - dot) + cCE * (L[k] - ((k == lab) ? 1.f : 0.f));
    }
  } else {
    float Q[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) if (k < K) Q[k] = tgt[i * K + k];
    softmax_inplace<KT>(Q, K, mx, lse);
#pragma unroll
    for (int k = 0; k < KT; ++k) if (k < K) { dl[k] = cA[k] * Q[k] + cB[k]; dot += dl[k] * L[k]; }
#pragma unroll
    for (int k = 0; k < KT; ++k) if (k < K) dl[k] = L[k] * (dl[k] - dot);
  }
}

// Adjoint of the trilinear interpolation (align_corners=False), gather form: one thread group per COARSE cell sums
// weight * dfine over the (2*scale)^3 loss-grid voxels whose footprint touches the cell.  No atomics, deterministic.
// dfine is planar [B][K][Z][Y][X]; dsrc has the layout of the coarse source (planar or channels-last).
__global__ void __launch_bounds__(256) trilinear_adjoint_k(const float* __restrict__ dfine, SrcGeom g, int K, float* __restrict__ dsrc, int G,
                                                           long long cells) {
  __shared__ float red[8];
  __shared__ float wzs[8][72];  // per thread group: the 1-D z weights of its cell (footprint <= 2 * scale + 2 <= 66 planes)
  const int per_block = 256 / G;
  const long long cell = (long long)blockIdx.x * per_block + threadIdx.x / G;  // cell index includes the class: ((b*K + k)*rz + cz)...
  const int gt = threadIdx.x % G, grp = threadIdx.x / G;
  const bool live = cell < cells;
  float acc = 0.f;
  int bk = 0, cz = 0, cy = 0, cx = 0;
  auto w1d = [](int j, float sc, int n, int c) {
    int i0, i1; float l1;
    lin_src(j, sc, n, i0, i1, l1);
    return (i0 == c ? 1.f - l1 : 0.f) + (i1 == c ? l1 : 0.f);
  };
  long long c = live ? cell : 0;
  cx = (int)(c % g.rx); c /= g.rx;
  cy = (int)(c % g.ry); c /= g.ry;
  cz = (int)(c % g.rz); bk = (int)(c / g.rz);
  const int fz = g.Z / g.rz, fy = g.Y / g.ry, fx = g.X / g.rx;  // integer scale factors (host-checked)
  const int z0 = max(0, cz * fz - fz / 2 - 1), z1 = min(g.Z - 1, cz * fz + (3 * fz) / 2);
  const int y0 = max(0, cy * fy - fy / 2 - 1), y1 = min(g.Y - 1, cy * fy + (3 * fy) / 2);
  const int x0 = max(0, cx * fx - fx / 2 - 1), x1 = min(g.X - 1, cx * fx + (3 * fx) / 2);
  const int nz = min(72, z1 - z0 + 1), ny = y1 - y0 + 1, nx = x1 - x0 + 1;
  for (int i = gt; i < nz; i += G) wzs[grp][i] = w1d(z0 + i, g.sz, g.rz, cz);
  __syncthreads();
  if (live) {
    // separable weights: a thread owns (y, x) positions of the footprint (one index division per position, not per voxel),
    // keeps wy * wx in a register and walks z with the tabulated z weights; consecutive threads read consecutive x.
    const float* src = dfine + (long long)bk * g.Z * g.Y * g.X;
    const long long zs = (long long)g.Y * g.X;
    for (int pr = gt; pr < ny * nx; pr += G) {
      const int yy = pr / nx, xx = pr - yy * nx;
      const float wyx = w1d(y0 + yy, g.sy, g.ry, cy) * w1d(x0 + xx, g.sx, g.rx, cx);
      if (wyx == 0.f) continue;
      const float* col = src + (long long)z0 * zs + (long long)(y0 + yy) * g.X + (x0 + xx);
      float a = 0.f;
#pragma unroll 4
      for (int i = 0; i < nz; ++i) a = fmaf(wzs[grp][i], col[i * zs], a);
      acc += wyx * a;
    }
  }
  acc = warp_sum(acc);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (G > 32) {
    if (lane == 0) red[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { acc = 0.f; for (int w = 0; w < 8; ++w) acc += red[w]; }
  }
  if (cell < cells && ((G > 32) ? threadIdx.x == 0 : lane == 0)) {
    const long long Sr = (long long)g.rz * g.ry * g.rx, sp = ((long long)cz * g.ry + cy) * g.rx + cx;
    const int b = bk / K, k = bk % K;
    if (g.planar) dsrc[((long long)b * K + k) * Sr + sp] = acc; else dsrc[((long long)b * Sr + sp) * K + k] = acc;
  }
}

// Backward.  Coefficients per class from the saved sums and the upstream gradients (device scalars):
//   dL/dA_k = -2 gd / (K den_k),  dL/dBp_k = gd num_k / (K den_k^2),  dL/dce_sum = gc / n_vox.
// dp_k = cA_k t_k + cB_k (2 p_k | 1);  dlogit_j = p_j (dp_j - sum_k dp_k p_k) + cCE (p_j - [lab==j]).
// Direct mode writes dsrc (channels-last or planar, same layout as src); interpolated mode
// scatters through a shared-memory tile of the coarse footprint, then one atomic per touched cell.
#define BT_Z 4
#define BT_Y 8
#define BT_X 8
#define FP_Z 4
#define FP_Y 6
#define FP_X 6
template <int KT>
__global__ void __launch_bounds__(256) class_stats_bwd_k(const float* __restrict__ src, SrcGeom g, int B, int K,
                                                         const long long* __restrict__ labels, const float* __restrict__ tgt,
                                                         int is_prob, const float* __restrict__ class_w,
                                                         const double* __restrict__ sums, const float* __restrict__ g_ce,
                                                         const float* __restrict__ g_dice, float w_ce, float w_dice, float* __restrict__ dsrc,
                                                         float* __restrict__ fine_out) {
  __shared__ float cA[KT], cB[KT];
  __shared__ float cCE;
  __shared__ float tile[KT][FP_Z * FP_Y * FP_X];
  // fine_out: write the per-voxel gradient at loss-grid resolution (planar) and leave the adjoint interpolation to
  // trilinear_adjoint_k (two-stage backward of interpolated sources)
  const bool direct = (g.rz == g.Z && g.ry == g.Y && g.rx == g.X) || fine_out != nullptr;
  if (threadIdx.x < K) {
    const int k = threadIdx.x;
    const double num = 2.0 * sums[1 + 3 * k] + 1e-5, den = sums[2 + 3 * k] + sums[3 + 3 * k] + 1e-5;
    const double gd = (g_dice ? (double)g_dice[0] : 0.0) * w_dice / K * (class_w ? (double)class_w[k] : 1.0);
    cA[k] = (float)(-2.0 * gd / den);
    cB[k] = (float)(gd * num / (den * den));
  }
  if (threadIdx.x == 0) cCE = (g_ce && labels && !is_prob) ? g_ce[0] * w_ce / (float)((double)B * g.Z * g.Y * g.X) : 0.f;
  const int tz = cdiv(g.Z, BT_Z), ty = cdiv(g.Y, BT_Y), tx = cdiv(g.X, BT_X);
  int t = blockIdx.x;
  const int bx = t % tx; t /= tx;
  const int by = t % ty; t /= ty;
  const int bz = t % tz; const int b = t / tz;
  const int lx_ = threadIdx.x % BT_X, ly_ = (threadIdx.x / BT_X) % BT_Y, lz_ = threadIdx.x / (BT_X * BT_Y);
  const int z = bz * BT_Z + lz_, y = by * BT_Y + ly_, x = bx * BT_X + lx_;
  const bool valid = z < g.Z && y < g.Y && x < g.X;
  int fz0 = 0, fy0 = 0, fx0 = 0;
  if (!direct) {
    for (int i = threadIdx.x; i < KT * FP_Z * FP_Y * FP_X; i += 256) (&tile[0][0])[i] = 0.f;
    int a, c; float l;
    lin_src(bz * BT_Z, g.sz, g.rz, a, c, l); fz0 = a;
    lin_src(by * BT_Y, g.sy, g.ry, a, c, l); fy0 = a;
    lin_src(bx * BT_X, g.sx, g.rx, a, c, l); fx0 = a;
  }
  __syncthreads();
  if (valid) {
    const long long S = (long long)g.Z * g.Y * g.X;
    const long long i = (long long)b * S + ((long long)z * g.Y + y) * g.X + x;
    float dl[KT];
    voxel_dlogits<KT>(src, g, K, b, z, y, x, i, labels, tgt, is_prob, cA, cB, cCE, dl);
    if (direct) {
      const long long v = ((long long)z * g.Y + y) * g.X + x;
#pragma unroll
      for (int k = 0; k < KT; ++k) if (k < K) {
        if (fine_out) fine_out[((long long)b * K + k) * S + v] = dl[k];
        else if (g.planar) dsrc[((long long)b * K + k) * S + v] = dl[k];
        else dsrc[((long long)b * S + v) * K + k] = dl[k];
      }
    } else {
      int z0, z1, y0, y1, x0, x1; float lz, ly, lx;
      lin_src(z, g.sz, g.rz, z0, z1, lz); lin_src(y, g.sy, g.ry, y0, y1, ly); lin_src(x, g.sx, g.rx, x0, x1, lx);
      const int zi[2] = {z0 - fz0, z1 - fz0}, yi[2] = {y0 - fy0, y1 - fy0}, xi[2] = {x0 - fx0, x1 - fx0};
      const float wz[2] = {1.f - lz, lz}, wy[2] = {1.f - ly, ly}, wx[2] = {1.f - lx, lx};
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float wgt = wz[a] * wy[c] * wx[e];
            if (wgt == 0.f) continue;
            const int cell = (zi[a] * FP_Y + yi[c]) * FP_X + xi[e];
#pragma unroll
            for (int k = 0; k < KT; ++k) if (k < K) atomicAdd(&tile[k][cell], wgt * dl[k]);
          }
    }
  }
  if (!direct) {
    __syncthreads();
    const long long Sr = (long long)g.rz * g.ry * g.rx;
    for (int i = threadIdx.x; i < K * FP_Z * FP_Y * FP_X; i += 256) {
      const int k = i / (FP_Z * FP_Y * FP_X), cell = i % (FP_Z * FP_Y * FP_X);
      const float v = tile[k][cell];
      if (v == 0.f) continue;
      const int cz = fz0 + cell / (FP_Y * FP_X), cy = fy0 + (cell / FP_X) % FP_Y, cx = fx0 + cell % FP_X;
      if (cz >= g.rz || cy >= g.ry || cx >= g.rx) continue;
      const long long sp = ((long long)cz * g.ry + cy) * g.rx + cx;
      atomicAdd(g.planar ? &dsrc[((long long)b * K + k) * Sr + sp] : &dsrc[((long long)b * Sr + sp) * K + k], v);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Row form of the class-statistics kernels (the path every BASELINE shape takes: X <= 256).
// A block walks over whole x-rows (b, z, y) of the loss grid, threads along x.  For an interpolated source the z / y
// interpolation of the row is done ONCE per row into a shared-memory table tab[k][cx] (K * rx values from the four coarse
// rows), so a voxel costs 2 shared-memory reads per class instead of 8 global ones; labels / soft targets / direct logits
// are read with full-width vector loads (a voxel's K contiguous channels-last values = one or more float4).
// Backward: the per-voxel gradient of the row goes to shared memory and the x-axis adjoint of the interpolation
// (X -> rx) is applied inside the block; the kernel writes [B][K][Z][Y][rx] (X / rx times smaller than the fine gradient
// the old two-stage path wrote), and trilinear_adjoint_k finishes the y / z adjoint on that reduced tensor.
// ------------------------------------------------------------------------------------------
#define ROW_MAXX 256
#define ROW_MAXRX 128   // rx <= X / 2

template <int KT>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, int K, float* L) {
  if (K == KT && KT % 4 == 0) {
#pragma unroll
    for (int k = 0; k < KT; k += 4) {
      const float4 v = *reinterpret_cast<const float4*>(p + k);
      L[k] = v.x; L[k + 1] = v.y; L[k + 2] = v.z; L[k + 3] = v.w;
    }
  } else if (K == KT && KT == 2) {
    const float2 v = *reinterpret_cast<const float2*>(p);
    L[0] = v.x; L[1] = v.y;
  } else {
#pragma unroll
    for (int k = 0; k < KT; ++k) if (k < K) L[k] = p[k];
  }
}
template <int KT>
__device__ __forceinline__ void store_vec(float* __restrict__ p, int K, const float* L) {
  if (K == KT && KT % 4 == 0) {
#pragma unroll
    for (int k = 0; k < KT; k += 4) *reinterpret_cast<float4*>(p + k) = make_float4(L[k], L[k + 1], L[k + 2], L[k + 3]);
  } else if (K == KT && KT == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(L[0], L[1]);
  } else {
#pragma unroll
    for (int k = 0; k < KT; ++k) if (k < K) p[k] = L[k];
  }
}

// One WARP owns a row (no block-wide barriers inside the row loop; 8 rows in flight per block).
// tab[k][cx] (row stride tw = rx + 1) = z/y-interpolated coarse row for fine row (z, y) of sample b.  Caller __syncwarp()s.
template <int KT>
__device__ __forceinline__ void row_table(const float* __restrict__ src, const SrcGeom& g, int K, int b, int z, int y, float* tab, int tw, int lane) {
  int z0, z1, y0, y1; float lz, ly;
  lin_src(z, g.sz, g.rz, z0, z1, lz); lin_src(y, g.sy, g.ry, y0, y1, ly);
  const float w00 = (1.f - lz) * (1.f - ly), w01 = (1.f - lz) * ly, w10 = lz * (1.f - ly), w11 = lz * ly;
  const long long Sr = (long long)g.rz * g.ry * g.rx;
  const long long r00 = ((long long)z0 * g.ry + y0) * g.rx, r01 = ((long long)z0 * g.ry + y1) * g.rx;
  const long long r10 = ((long long)z1 * g.ry + y0) * g.rx, r11 = ((long long)z1 * g.ry + y1) * g.rx;
  const int n = K * g.rx;
  for (int i = lane; i < n; i += 32) {
    int k, cx; long long st; const float* p;
    if (g.planar) { k = i / g.rx; cx = i - k * g.rx; p = src + ((long long)b * K + k) * Sr + cx; st = 1; }
    else { cx = i / K; k = i - cx * K; p = src + (long long)b * Sr * K + (long long)cx * K + k; st = K; }
    tab[k * tw + cx] = w00 * p[r00 * st] + w01 * p[r01 * st] + w10 * p[r10 * st] + w11 * p[r11 * st];
  }
}

template <int KT>
__device__ __forceinline__ void row_logits(const float* __restrict__ src, const SrcGeom& g, bool direct, int K, long long vox_in_batch, int b,
                                           int x, const float* tab, int tw, float* L) {
  if (direct) {
    const long long S = (long long)g.Z * g.Y * g.X;
    if (g.planar) {
#pragma unroll
      for (int k = 0; k < KT; ++k) if (k < K) L[k] = src[((long long)b * K + k) * S + vox_in_batch];
    } else {
      load_vec<KT>(src + ((long long)b * S + vox_in_batch) * K, K, L);
    }
  } else {
    int x0, x1; float lx;
    lin_src(x, g.sx, g.rx, x0, x1, lx);
#pragma unroll
    for (int k = 0; k < KT; ++k) if (k < K) L[k] = (1.f - lx) * tab[k * tw + x0] + lx * tab[k * tw + x1];
  }
}

template <int KT>
__global__ void __launch_bounds__(256, 2) class_stats_row_fwd_k(const float* __restrict__ src, SrcGeom g, int B, int K,
                                                                const long long* __restrict__ labels, const float* __restrict__ tgt,
                                                                int is_prob, double* __restrict__ sums) {
  extern __shared__ float dsm[];   // per warp: tab[KT][rx + 1]
  __shared__ float red[8][3 * KT + 1];
  float acc[3 * KT + 1];
#pragma unroll
  for (int i = 0; i < 3 * KT + 1; ++i) acc[i] = 0.f;
  const bool direct = g.rz == g.Z && g.ry == g.Y && g.rx == g.X;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int tw = g.rx + 1;
  float* tab = dsm + (size_t)wid * KT * tw;
  const long long rows = (long long)B * g.Z * g.Y;
  for (long long row = (long long)blockIdx.x * nw + wid; row < rows; row += (long long)gridDim.x * nw) {
    const int b = (int)(row / ((long long)g.Z * g.Y));
    const int zy = (int)(row - (long long)b * g.Z * g.Y);
    const int z = zy / g.Y, y = zy - z * g.Y;
    if (!direct) {
      __syncwarp();   // the previous row's readers are done with the table
      row_table<KT>(src, g, K, b, z, y, tab, tw, lane);
      __syncwarp();
    }
    for (int x = lane; x < g.X; x += 32) {
      const long long v = (long long)zy * g.X + x, i = row * g.X + x;
      float L[KT];
      row_logits<KT>(src, g, direct, K, v, b, x, tab, tw, L);
      float mx, lse;
      if (labels) {
        const int lab = (int)labels[i];
        float raw = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k) if (k == lab) raw = L[k];
        if (!is_prob) {
          softmax_inplace<KT>(L, K, mx, lse);
          acc[0] += (mx + lse) - raw;
        }
#pragma unroll
        for (int k = 0; k < KT; ++k) if (k < K) {
          const float t = (k == lab) ? 1.f : 0.f;
          acc[1 + 3 * k] += L[k] * t; acc[2 + 3 * k] += L[k] * L[k]; acc[3 + 3 * k] += t;
        }
      } else {
        float Q[KT];
        load_vec<KT>(tgt + i * K, K, Q);
        softmax_inplace<KT>(L, K, mx, lse);
        softmax_inplace<KT>(Q, K, mx, lse);
#pragma unroll
        for (int k = 0; k < KT; ++k) if (k < K) { acc[1 + 3 * k] += L[k] * Q[k]; acc[2 + 3 * k] += L[k]; acc[3 + 3 * k] += Q[k]; }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 3 * KT + 1; ++i) {
    const float sv = warp_sum(acc[i]);
    if (lane == 0) red[wid][i] = sv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * K + 1; i += blockDim.x) {
    double t = 0.0;
    for (int w = 0; w < nw; ++w) t += (double)red[w][i];
    atomicAdd(&sums[i], t);
  }
}

// gradient w.r.t. the logits of one voxel from its (already loaded) logits L (overwritten with probabilities)
template <int KT>
__device__ __forceinline__ void dlogits_from(float* L, int K, long long i, const long long* __restrict__ labels, const float* __restrict__ tgt,
                                             int is_prob, const float* cA, const float* cB, float cCE, float* dl) {
  float mx, lse;
  if (!is_prob) softmax_inplace<KT>(L, K, mx, lse);
  float dot = 0.f;
  if (labels) {
    const int lab = (int)labels[i];
#pragma unroll
    for (int k = 0; k < KT; ++k) if (k < K) {
      dl[k] = cA[k] * ((k == lab) ? 1.f : 0.f) + cB[k] * 2.f * L[k];
      dot += dl[k] * L[k];
    }
    if (!is_prob) {
#pragma unroll
      for (int k = 0; k < KT; ++k) if (k < K) dl[k] = L[k] * (dl[k] - dot) + cCE * (L[k] - ((k == lab) ? 1.f : 0.f));
    }
  } else {
    float Q[KT];
    load_vec<KT>(tgt + i * K, K, Q);
    softmax_inplace<KT>(Q, K, mx, lse);
#pragma unroll
    for (int k = 0; k < KT; ++k) if (k < K) { dl[k] = cA[k] * Q[k] + cB[k]; dot += dl[k] * L[k]; }
#pragma unroll
    for (int k = 0; k < KT; ++k) if (k < K) dl[k] = L[k] * (dl[k] - dot);
  }
}

// direct sources: dsrc (layout of src) = per-voxel gradient.  Interpolated sources: xred[b][k][z][y][cx] = x-adjoint of the row.
template <int KT>
__global__ void __launch_bounds__(256, 2) class_stats_row_bwd_k(const float* __restrict__ src, SrcGeom g, int B, int K,
                                                                const long long* __restrict__ labels, const float* __restrict__ tgt,
                                                                int is_prob, const float* __restrict__ class_w, const double* __restrict__ sums,
                                                                const float* __restrict__ g_ce, const float* __restrict__ g_dice, float w_ce,
                                                                float w_dice, float* __restrict__ dsrc, float* __restrict__ xred) {
  __shared__ float cA[KT], cB[KT];
  __shared__ float cCE;
  extern __shared__ float dsm[];   // per warp (interpolated sources only): tab[KT][rx + 1], then sdl[KT][X + 1]
  const bool direct = g.rz == g.Z && g.ry == g.Y && g.rx == g.X;
  if (threadIdx.x < K) {
    const int k = threadIdx.x;
    const double num = 2.0 * sums[1 + 3 * k] + 1e-5, den = sums[2 + 3 * k] + sums[3 + 3 * k] + 1e-5;
    const double gd = (g_dice ? (double)g_dice[0] : 0.0) * w_dice / K * (class_w ? (double)class_w[k] : 1.0);
    cA[k] = (float)(-2.0 * gd / den);
    cB[k] = (float)(gd * num / (den * den));
  }
  if (threadIdx.x == 0) cCE = (g_ce && labels && !is_prob) ? g_ce[0] * w_ce / (float)((double)B * g.Z * g.Y * g.X) : 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long rows = (long long)B * g.Z * g.Y, S = (long long)g.Z * g.Y * g.X;
  const int tw = g.rx + 1, XS = g.X + 1, fx = direct ? 1 : g.X / g.rx;
  float* tab = dsm + (size_t)wid * KT * (tw + XS);
  float* sdl = tab + KT * tw;
  for (long long row = (long long)blockIdx.x * nw + wid; row < rows; row += (long long)gridDim.x * nw) {
    const int b = (int)(row / ((long long)g.Z * g.Y));
    const int zy = (int)(row - (long long)b * g.Z * g.Y);
    const int z = zy / g.Y, y = zy - z * g.Y;
    if (!direct) {
      __syncwarp();   // previous row: table readers and sdl readers are done
      row_table<KT>(src, g, K, b, z, y, tab, tw, lane);
      __syncwarp();
    }
    for (int x = lane; x < g.X; x += 32) {
      const long long v = (long long)zy * g.X + x, i = row * g.X + x;
      float L[KT], dl[KT];
      row_logits<KT>(src, g, direct, K, v, b, x, tab, tw, L);
      dlogits_from<KT>(L, K, i, labels, tgt, is_prob, cA, cB, cCE, dl);
      if (direct) {
        if (g.planar) {
#pragma unroll
          for (int k = 0; k < KT; ++k) if (k < K) dsrc[((long long)b * K + k) * S + v] = dl[k];
        } else {
          store_vec<KT>(dsrc + ((long long)b * S + v) * K, K, dl);
        }
      } else {
#pragma unroll
        for (int k = 0; k < KT; ++k) if (k < K) sdl[k * XS + x] = dl[k];
      }
    }
    if (!direct) {
      __syncwarp();
      // x-adjoint: coarse column cx collects the fine columns whose interpolation footprint touches it
      const int n = K * g.rx;
      for (int i = lane; i < n; i += 32) {
        const int k = i / g.rx, cx = i - k * g.rx;
        const int x0 = max(0, cx * fx - fx / 2 - 1), x1 = min(g.X - 1, cx * fx + (3 * fx) / 2);
        float a = 0.f;
        for (int x = x0; x <= x1; ++x) {
          int i0, i1; float l1;
          lin_src(x, g.sx, g.rx, i0, i1, l1);
          const float w = (i0 == cx ? 1.f - l1 : 0.f) + (i1 == cx ? l1 : 0.f);
          a = fmaf(w, sdl[k * XS + x], a);
        }
        xred[(((long long)b * K + k) * g.Z * g.Y + zy) * g.rx + cx] = a;
      }
    }
  }
}

// y / z adjoint of the trilinear interpolation on the x-reduced gradient xr[B*K][Z][Y][rx], small footprints: one THREAD per coarse
// cell (consecutive threads = consecutive cx: coalesced rows of rx floats), (2 fy + 2) x (2 fz + 2) taps each.
__global__ void __launch_bounds__(256) yz_adjoint_cell_k(const float* __restrict__ xr, SrcGeom g, int K, float* __restrict__ dsrc, long long cells) {
  const long long cell = (long long)blockIdx.x * 256 + threadIdx.x;
  if (cell >= cells) return;
  long long c = cell;
  const int cx = (int)(c % g.rx); c /= g.rx;
  const int cy = (int)(c % g.ry); c /= g.ry;
  const int cz = (int)(c % g.rz); const int bk = (int)(c / g.rz);
  const int fz = g.Z / g.rz, fy = g.Y / g.ry;
  const int z0 = max(0, cz * fz - fz / 2 - 1), z1 = min(g.Z - 1, cz * fz + (3 * fz) / 2);
  const int y0 = max(0, cy * fy - fy / 2 - 1), y1 = min(g.Y - 1, cy * fy + (3 * fy) / 2);
  const float* base = xr + (long long)bk * g.Z * g.Y * g.rx + cx;
  float acc = 0.f;
  for (int z = z0; z <= z1; ++z) {
    int i0, i1; float l1;
    lin_src(z, g.sz, g.rz, i0, i1, l1);
    const float wz = (i0 == cz ? 1.f - l1 : 0.f) + (i1 == cz ? l1 : 0.f);
    if (wz == 0.f) continue;
    float a = 0.f;
    for (int y = y0; y <= y1; ++y) {
      lin_src(y, g.sy, g.ry, i0, i1, l1);
      const float wy = (i0 == cy ? 1.f - l1 : 0.f) + (i1 == cy ? l1 : 0.f);
      a = fmaf(wy, base[((long long)z * g.Y + y) * g.rx], a);
    }
    acc = fmaf(wz, a, acc);
  }
  const long long Sr = (long long)g.rz * g.ry * g.rx, sp = ((long long)cz * g.ry + cy) * g.rx + cx;
  const int b = bk / K, k = bk % K;
  if (g.planar) dsrc[((long long)b * K + k) * Sr + sp] = acc; else dsrc[((long long)b * Sr + sp) * K + k] = acc;
}

static int make_geom(SrcGeom& g, int planar, int rz, int ry, int rx, int Z, int Y, int X) {
  g.planar = planar; g.rz = rz; g.ry = ry; g.rx = rx; g.Z = Z; g.Y = Y; g.X = X;
  g.sz = (float)rz / (float)Z; g.sy = (float)ry / (float)Y; g.sx = (float)rx / (float)X;
  const bool direct = rz == Z && ry == Y && rx == X;
  // an axis is either untouched (r == full, e.g. the depth-1 axis of the 2D path) or at most half the loss grid
  if (!direct && ((rz != Z && 2 * rz > Z) || (ry != Y && 2 * ry > Y) || (rx != X && 2 * rx > X))) {
    icl_set_error("class_stats: interpolated source must be at most half the loss grid (%d,%d,%d -> %d,%d,%d)", rz, ry, rx, Z, Y, X);
    return -1;
  }
  return 0;
}

// the row kernels take every loss grid whose rows fit their shared-memory staging (all BASELINE shapes: X = 96 or 256)
// (ICL_DISABLE_ROW_LOSS=1 routes everything to the generic per-voxel kernels — a test knob that keeps them covered)
static inline bool row_path_ok(bool direct, int rx, int X) {
  const char* e = getenv("ICL_DISABLE_ROW_LOSS");
  if (e && e[0] == '1') return false;
  return X <= ROW_MAXX && (direct || (rx <= ROW_MAXRX && 2 * rx <= X));
}

// warps (= rows in flight) per block of the row kernels: up to 8, limited by their per-warp shared-memory staging
static inline int row_warps(size_t per_warp_bytes) {
  int w = 8;
  while (w > 1 && per_warp_bytes * w > 90 * 1024) w >>= 1;
  return w;
}

#define DISPATCH_K(K, CALL)                                      \
  if (K <= 2) { CALL(2); } else if (K <= 4) { CALL(4); } else if (K <= 8) { CALL(8); } else { CALL(16); }

ICL_API int icl_class_stats_fwd(const float* src, int planar, int rz, int ry, int rx, int B, int K, int Z, int Y, int X,
                                const long long* labels, const float* tgt, int is_prob, const float* class_w,
                                double* sums /* zeroed [3K+1] */, float* out2, void* stream) {
  ICL_REQUIRE(K >= 1 && K <= 16, "class_stats: K=%d unsupported (max 16)", K);
  ICL_REQUIRE((labels != nullptr) != (tgt != nullptr), "class_stats: exactly one of labels / soft targets");
  SrcGeom g;
  if (make_geom(g, planar, rz, ry, rx, Z, Y, X)) return -1;
  const long long total = (long long)B * Z * Y * X;
  const bool direct = rz == Z && ry == Y && rx == X;
  if (row_path_ok(direct, rx, X)) {
    const long long rows = (long long)B * Z * Y;
#define CALL(KT) do { \
      const size_t per_warp = direct ? 0 : (size_t)KT * (rx + 1) * sizeof(float); \
      const int warps = row_warps(per_warp); \
      static bool cfg = false; \
      if (!cfg) { cudaFuncSetAttribute(class_stats_row_fwd_k<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); cfg = true; } \
      const int grid = (int)(cdiv(rows, warps) < 148 * 2 ? cdiv(rows, warps) : 148 * 2); \
      class_stats_row_fwd_k<KT><<<grid, warps * 32, per_warp * warps, as_stream(stream)>>>(src, g, B, K, labels, tgt, is_prob, sums); } while (0)
    DISPATCH_K(K, CALL)
#undef CALL
  } else {
#define CALL(KT) class_stats_fwd_k<KT><<<grid_for(total, 256, 148 * 8), 256, 0, as_stream(stream)>>>(src, g, B, K, labels, tgt, is_prob, sums)
    DISPATCH_K(K, CALL)
#undef CALL
  }
  icl_count_launch(1);
  class_stats_finalize_k<<<1, 32, 0, as_stream(stream)>>>(sums, K, (double)total, class_w, out2);
  ICL_LAUNCHED("class_stats_fwd");
}

ICL_API int icl_class_stats_bwd(const float* src, int planar, int rz, int ry, int rx, int B, int K, int Z, int Y, int X,
                                const long long* labels, const float* tgt, int is_prob, const float* class_w, const double* sums,
                                const float* g_ce, const float* g_dice, float w_ce, float w_dice, float* dsrc /* zeroed when interpolated */,
                                float* workspace /* B*K*Z*Y*X floats for interpolated sources, or null (then: atomic scatter) */, void* stream) {
  ICL_REQUIRE(K >= 1 && K <= 16, "class_stats: K=%d unsupported (max 16)", K);
  SrcGeom g;
  if (make_geom(g, planar, rz, ry, rx, Z, Y, X)) return -1;
  const long long blocks = (long long)B * cdiv(Z, BT_Z) * cdiv(Y, BT_Y) * cdiv(X, BT_X);
  const bool direct = rz == Z && ry == Y && rx == X;
  if (row_path_ok(direct, rx, X) && (direct || (workspace && Z % rz == 0 && Y % ry == 0 && X % rx == 0))) {
    const long long rows = (long long)B * Z * Y;
#define CALL(KT) do { \
      const size_t per_warp = direct ? 0 : (size_t)KT * (rx + 1 + X + 1) * sizeof(float); \
      const int warps = row_warps(per_warp); \
      static bool cfg = false; \
      if (!cfg) { cudaFuncSetAttribute(class_stats_row_bwd_k<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); cfg = true; } \
      const int grid = (int)(cdiv(rows, warps) < 148 * 2 ? cdiv(rows, warps) : 148 * 2); \
      class_stats_row_bwd_k<KT><<<grid, warps * 32, per_warp * warps, as_stream(stream)>>>( \
          src, g, B, K, labels, tgt, is_prob, class_w, sums, g_ce, g_dice, w_ce, w_dice, dsrc, workspace); } while (0)
    DISPATCH_K(K, CALL)
#undef CALL
    if (!direct) {
      icl_count_launch(1);
      // y / z adjoint on the x-reduced gradient [B][K][Z][Y][rx]: the same gather kernel with an untouched x axis
      SrcGeom g2;
      if (make_geom(g2, planar, rz, ry, rx, Z, Y, rx)) return -1;
      const long long cells = (long long)B * K * rz * ry * rx;
      const long long foot = (2LL * (Z / rz) + 2) * (2LL * (Y / ry) + 2);
      if (foot <= 400 && cells >= 148 * 256) {
        yz_adjoint_cell_k<<<(unsigned)cdiv(cells, 256), 256, 0, as_stream(stream)>>>(workspace, g2, K, dsrc, cells);
      } else {
        const int G = 4 * foot >= 4096 ? 256 : 32;
        const long long nb = (cells + (256 / G) - 1) / (256 / G);
        trilinear_adjoint_k<<<(unsigned)nb, 256, 0, as_stream(stream)>>>(workspace, g2, K, dsrc, G, cells);
      }
    }
    ICL_LAUNCHED("class_stats_bwd");
  }
  float* fine = (!direct && workspace && Z % rz == 0 && Y % ry == 0 && X % rx == 0) ? workspace : nullptr;
#define CALL(KT) class_stats_bwd_k<KT><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(src, g, B, K, labels, tgt, is_prob, class_w, sums, g_ce, g_dice, w_ce, w_dice, dsrc, fine)
  DISPATCH_K(K, CALL)
#undef CALL
  if (fine) {
    icl_count_launch(1);
    const long long foot = 8LL * (Z / rz) * (Y / ry) * (X / rx);
    const int G = foot >= 4096 ? 256 : 32;
    const long long cells = (long long)B * K * rz * ry * rx;
    const long long nb = (cells + (256 / G) - 1) / (256 / G);
    trilinear_adjoint_k<<<(unsigned)nb, 256, 0, as_stream(stream)>>>(fine, g, K, dsrc, G, cells);
  }
  ICL_LAUNCHED("class_stats_bwd");
}

// ------------------------------------------------------------------------------------------
// softmax-MSE consistency on planar [B,K,S] pairs: loss_sum += sum (softmax(a)-softmax(b))^2;
// backward: da_j = coef * 2 p_j (e_j - sum_k e_k p_k), e = p - q, coef = upstream * w / numel.
// ------------------------------------------------------------------------------------------
template <int KT>
__global__ void softmax_mse_k(const float* __restrict__ a, const float* __restrict__ bq, int B, int K, long long S, double* __restrict__ sum,
                              const float* __restrict__ gup, float w, float* __restrict__ da) {
  __shared__ float red[33];
  float acc = 0.f;
  const float coef = da ? gup[0] * w / (float)((double)B * K * S) : 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)B * S; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / S, s = i % S;
    float P[KT], Q[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) if (k < K) { P[k] = a[(b * K + k) * S + s]; Q[k] = bq[(b * K + k) * S + s]; }
    float mx, lse;
    softmax_inplace<KT>(P, K, mx, lse);
    softmax_inplace<KT>(Q, K, mx, lse);
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < KT; ++k) if (k < K) { const float e = P[k] - Q[k]; acc += e * e; dot += e * P[k]; }
    if (da) {
#pragma unroll
      for (int k = 0; k < KT; ++k) if (k < K) da[(b * K + k) * S + s] = coef * 2.f * P[k] * ((P[k] - Q[k]) - dot);
    }
  }
  if (sum) {
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(sum, (double)acc);
  }
}
ICL_API int icl_softmax_mse(const float* a, const float* b, int B, int K, long long S, double* sum, const float* gup, float w, float* da,
                            void* stream) {
  ICL_REQUIRE(K >= 1 && K <= 16, "softmax_mse: K=%d unsupported (max 16)", K);
#define CALL(KT) softmax_mse_k<KT><<<grid_for((long long)B * S, 256, 148 * 4), 256, 0, as_stream(stream)>>>(a, b, B, K, S, sum, gup, w, da)
  DISPATCH_K(K, CALL)
#undef CALL
  ICL_LAUNCHED("softmax_mse");
}

// out[0] = (float)(sum[0] * scale)
__global__ void scale_to_float_k(const double* s, double scale, float* out) { out[0] = (float)(s[0] * scale); }
ICL_API int icl_scale_to_float(const double* s, double scale, float* out, void* stream) {
  scale_to_float_k<<<1, 1, 0, as_stream(stream)>>>(s, scale, out);
  ICL_LAUNCHED("scale_to_float");
}

// ------------------------------------------------------------------------------------------
// Fused multi-tensor SGD (momentum, weight decay; torch.optim.SGD semantics, dampening 0):
//   d = g + wd*p ; buf = first ? d : mu*buf + d ; p -= lr*buf.   lr is read from device memory so a
// captured CUDA graph can be replayed with a new poly-LR value.
// ------------------------------------------------------------------------------------------
struct SgdTensor { float* p; const float* g; float* m; long long n; };
#define SGD_CHUNK 16384
__global__ void __launch_bounds__(256) sgd_multi_k(const SgdTensor* __restrict__ tab, const int* __restrict__ chunk_tensor,
                                                   const long long* __restrict__ chunk_off, const float* __restrict__ lr_ptr, float mu,
                                                   float wd, int first) {
  const SgdTensor t = tab[chunk_tensor[blockIdx.x]];
  const long long off = chunk_off[blockIdx.x];
  const long long n = min((long long)SGD_CHUNK, t.n - off);
  const float lr = lr_ptr[0];
  float* p = t.p + off; const float* g = t.g + off; float* m = t.m + off;
  const bool vec = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m)) & 15) == 0;
  if (vec) {
    const long long n4 = n / 4;
    for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 pv = reinterpret_cast<float4*>(p)[i];
      const float4 gv = reinterpret_cast<const float4*>(g)[i];
      float4 mv = first ? make_float4(0, 0, 0, 0) : reinterpret_cast<float4*>(m)[i];
      float4 d = make_float4(gv.x + wd * pv.x, gv.y + wd * pv.y, gv.z + wd * pv.z, gv.w + wd * pv.w);
      mv = first ? d : make_float4(mu * mv.x + d.x, mu * mv.y + d.y, mu * mv.z + d.z, mu * mv.w + d.w);
      pv.x -= lr * mv.x; pv.y -= lr * mv.y; pv.z -= lr * mv.z; pv.w -= lr * mv.w;
      reinterpret_cast<float4*>(m)[i] = mv;
      reinterpret_cast<float4*>(p)[i] = pv;
    }
    for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      const float d = g[i] + wd * p[i];
      const float b = first ? d : mu * m[i] + d;
      m[i] = b; p[i] -= lr * b;
    }
  } else {
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const float d = g[i] + wd * p[i];
      const float b = first ? d : mu * m[i] + d;
      m[i] = b; p[i] -= lr * b;
    }
  }
}
ICL_API int icl_sgd_multi(const void* tab, const int* chunk_tensor, const long long* chunk_off, int n_chunks, const float* lr_ptr, float mu,
                          float wd, int first, void* stream) {
  if (n_chunks <= 0) return 0;
  sgd_multi_k<<<n_chunks, 256, 0, as_stream(stream)>>>((const SgdTensor*)tab, chunk_tensor, chunk_off, lr_ptr, mu, wd, first);
  ICL_LAUNCHED("sgd_multi");
}
ICL_API int icl_sgd_chunk(void) { return SGD_CHUNK; }

// ------------------------------------------------------------------------------------------
// Sliding-window inference (test_3D_BraTS.py:110-135): scores and visit counts stay on the
// device; one launch per window adds softmax(logits) into score[K][W][H][D] at the window
// origin; finalize divides by the count and takes the first-max argmax.
// ------------------------------------------------------------------------------------------
template <int KT>
__global__ void sw_accumulate_k(const float* __restrict__ logits /* [pw,ph,pd,K] channels-last */, int K, int pw, int ph, int pd,
                                float* __restrict__ score, float* __restrict__ cnt, int W, int H, int D, int xs, int ys, int zs) {
  const long long total = (long long)pw * ph * pd, S = (long long)W * H * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long v = i;
    const int z = (int)(v % pd); v /= pd;
    const int y = (int)(v % ph); const int x = (int)(v / ph);
    float L[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) if (k < K) L[k] = logits[i * K + k];
    float mx, lse;
    softmax_inplace<KT>(L, K, mx, lse);
    const long long o = ((long long)(xs + x) * H + (ys + y)) * D + (zs + z);
#pragma unroll
    for (int k = 0; k < KT; ++k) if (k < K) score[k * S + o] += L[k];
    cnt[o] += 1.f;
  }
}
ICL_API int icl_sw_accumulate(const float* logits, int K, int pw, int ph, int pd, float* score, float* cnt, int W, int H, int D, int xs,
                              int ys, int zs, void* stream) {
  ICL_REQUIRE(K >= 1 && K <= 16, "sw_accumulate: K=%d unsupported", K);
  ICL_REQUIRE(xs >= 0 && ys >= 0 && zs >= 0 && xs + pw <= W && ys + ph <= H && zs + pd <= D, "sw_accumulate: window out of bounds");
#define CALL(KT) sw_accumulate_k<KT><<<grid_for((long long)pw * ph * pd, 256), 256, 0, as_stream(stream)>>>(logits, K, pw, ph, pd, score, cnt, W, H, D, xs, ys, zs)
  DISPATCH_K(K, CALL)
#undef CALL
  ICL_LAUNCHED("sw_accumulate");
}
__global__ void sw_finalize_k(const float* __restrict__ score, const float* __restrict__ cnt, int K, long long S, long long* __restrict__ label) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (long long)gridDim.x * blockDim.x) {
    const float c = cnt[i];
    float best = score[i] / c; int bi = 0;
    for (int k = 1; k < K; ++k) {
      const float v = score[k * S + i] / c;
      if (v > best) { best = v; bi = k; }
    }
    label[i] = bi;
  }
}
ICL_API int icl_sw_finalize(const float* score, const float* cnt, int K, long long S, long long* label, void* stream) {
  sw_finalize_k<<<grid_for(S, 256), 256, 0, as_stream(stream)>>>(score, cnt, K, S, label);
  ICL_LAUNCHED("sw_finalize");
}
// counts[0] = |pred>0 & gt>0|, counts[1] = |pred>0|, counts[2] = |gt>0|  (exact integers)
__global__ void dice_counts_k(const long long* __restrict__ pred, const long long* __restrict__ gt, long long n, unsigned long long* __restrict__ counts) {
  unsigned long long a = 0, b = 0, c = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const bool p = pred[i] > 0, g = gt[i] > 0;
    a += (p && g); b += p; c += g;
  }
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&counts[0], a); atomicAdd(&counts[1], b); atomicAdd(&counts[2], c); }
}
ICL_API int icl_dice_counts(const long long* pred, const long long* gt, long long n, unsigned long long* counts, void* stream) {
  dice_counts_k<<<grid_for(n, 256, 148 * 4), 256, 0, as_stream(stream)>>>(pred, gt, n, counts);
  ICL_LAUNCHED("dice_counts");
}
