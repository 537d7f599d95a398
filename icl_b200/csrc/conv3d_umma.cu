// 3x3x3 convolution (stride 1, zero pad 1) as an implicit GEMM on the 5th-gen tensor cores:
// TMA halo-tile loads -> shared memory -> tcgen05.mma (accumulators in TMEM) -> tcgen05.ld
// epilogue (+bias, fp32 NDHWC store, per-(sample, channel) sum / sum-of-squares for the
// InstanceNorm that always follows, networks/utils.py:104-109).  The same kernel computes the
// data gradient (weights packed flipped + transposed).
//
// GEMM view per output tile:  D[M=128 voxels][N=NT out-channels] += A[M][K] * B[K][N],
//   M tile = 16 (h) x 8 (w) voxels of one depth plane; K runs over 3 depth taps x Cin/16 chunks
//   (one pipeline stage each) x 9 in-plane taps x 16 channels (one MMA each).
// Operands are bf16.  "parity" mode (P=2) keeps fp32-level accuracy by splitting every operand
// into hi+lo bf16 planes and issuing hi*hi + hi*lo + lo*hi into the same fp32 accumulator
// (SURVEY.md §7.3(2)); "fast" mode (P=1) uses the hi plane only.
//
// Shared-memory operand layout = UMMA canonical K-major, no swizzle ("interleave"):
//   A stage  [plane p][k8 chunk j(2)][halo line (18)][halo w (10)][8 ch]   (written by ONE 4-D TMA box per
//            plane from the PK activation tensor [P][B][C/8][D][H][W][8]); core matrix = 8 consecutive w of
//            one line (8 x 16 B contiguous), SBO = one halo line (160 B), LBO = one k8 chunk (2880 B).
//            An in-plane tap (kh,kw) is just a start-address offset of (kh*10 + kw)*16 B, so the 9 taps
//            reuse one staged tile.
//   B stage  [plane p][tap 9][k8 chunk 2][n NT][8 ch]  (one 1-D bulk copy; weights pre-packed per stage),
//            core matrix = 8 n x 16 B, SBO = 128 B, LBO = NT*16 B.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer, warps 2-5 = epilogue.
#include "umma.cuh"
#include <stdlib.h>

// Optional cycle accounting of the three roles (CTA 0 only; thread-local clock64 deltas, env ICL_UMMA_PROF=1).
#define PROF_DECL() long long prof_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long prof_t0_ = 0; const bool prof_on = p.prof != nullptr && blockIdx.x == 0
#define PROF_T0() do { if (prof_on) prof_t0_ = clock64(); } while (0)
#define PROF_ADD(i) do { if (prof_on) { const long long n_ = clock64(); prof_acc[i] += n_ - prof_t0_; prof_t0_ = n_; } } while (0)
#define PROF_FLUSH(lo, hi) do { if (prof_on) for (int i_ = lo; i_ <= hi; ++i_) p.prof[i_] = prof_acc[i_]; } while (0)

#define UM_TH 16
#define UM_TW 8
#define UM_HL (UM_TH + 2)
#define UM_HW (UM_TW + 2)
#define UM_KC 16
#define UM_A_PLANE_BYTES (2 * UM_HL * UM_HW * 16)  // 5760: two k8 chunks of one halo plane
#define UM_A_LBO (UM_HL * UM_HW * 16)              // 2880
#define UM_A_SBO (UM_HW * 16)                      // 160
#define UM_MAX_STAGES 8
#define UM_MAX_COUT 256
#define UM_MAX_ACC 8

struct UmmaConvParams {
  const __nv_bfloat16* wp;  // packed weights [nt][kd][chunk][p][tap9][2][NT][8]
  const float* bias;        // [Cout] or null
  float* y0; int ld0;       // output columns [0, split)
  float* y1; int ld1;       // output columns [split, Cout)
  int split;
  double* stats;            // [B][Cout][2] or null
  int B, D, H, W;
  int C0, C1;               // channels of source 0 / source 1 (virtual concat)
  int Cout, NT, n_tiles;    // N tiling
  int tiles_h, tiles_w;
  long long num_tiles;
  int stages, P;
  int tmem_cols;
  int n_acc;  // TMEM accumulator ring depth (MMA of tile t+n_acc waits for the epilogue of tile t)
  long long* prof;  // optional [16] cycle counters written by CTA 0 (env ICL_UMMA_PROF=1), else null
};

struct TileCoord { int nt, b, d, h0, w0; };
// Tile t = ((((b * D + d) * tiles_h + th) * tiles_w + tw) * n_tiles + nt), walked with stride gridDim.x (neighbouring CTAs
// work on neighbouring tiles, so halo planes are re-read from L2).  Integer division is ~200 dependent cycles per call
// on the single-thread roles, so the mixed-radix digits are advanced with carries instead of being re-derived per tile.
struct TileWalker {
  int nt, tw, th, d, b;
  int snt, stw, sth, sd, sb;
  __device__ __forceinline__ static void digits(int t, const UmmaConvParams& p, int& nt, int& tw, int& th, int& d, int& b) {
    nt = t % p.n_tiles; t /= p.n_tiles;
    tw = t % p.tiles_w; t /= p.tiles_w;
    th = t % p.tiles_h; t /= p.tiles_h;
    d = t % p.D;
    b = t / p.D;
  }
  __device__ __forceinline__ TileWalker(int t0, int stride, const UmmaConvParams& p) {
    digits(t0, p, nt, tw, th, d, b);
    digits(stride, p, snt, stw, sth, sd, sb);
  }
  __device__ __forceinline__ void advance(const UmmaConvParams& p) {
    nt += snt; int c = nt >= p.n_tiles; nt -= c ? p.n_tiles : 0;
    tw += stw + c; c = tw >= p.tiles_w; tw -= c ? p.tiles_w : 0;
    th += sth + c; c = th >= p.tiles_h; th -= c ? p.tiles_h : 0;
    d += sd + c; c = d >= p.D; d -= c ? p.D : 0;
    b += sb + c;
  }
  __device__ __forceinline__ TileCoord coord() const {
    TileCoord c; c.nt = nt; c.b = b; c.d = d; c.h0 = th * UM_TH; c.w0 = tw * UM_TW; return c;
  }
};

__global__ void __launch_bounds__(192, 1)
conv3d_umma_k(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1, const UmmaConvParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * UM_MAX_STAGES + 2 * UM_MAX_ACC];
  __shared__ uint32_t tmem_base_s;
  __shared__ float sstat[UM_MAX_COUT][2];
  __shared__ float sbias[UM_MAX_COUT + 128];  // Cout <= 256 with statistics, <= 384 for the data gradient (no bias there)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = p.P, NT = p.NT, stages = p.stages;
  const uint32_t a_bytes = UM_A_PLANE_BYTES * P;        // per stage
  const uint32_t b_plane = 9u * NT * 32u;               // per plane per stage
  const uint32_t b_bytes = b_plane * P;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t smem0 = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[UM_MAX_STAGES]);
  const uint32_t tfull0 = smem_u32(&bars[2 * UM_MAX_STAGES]), tempty0 = smem_u32(&bars[2 * UM_MAX_STAGES + UM_MAX_ACC]);
  const int nchunks = (p.C0 + p.C1) / UM_KC;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < p.n_acc; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA1) : "memory");
  }
  for (int i = threadIdx.x; i < UM_MAX_COUT * 2; i += blockDim.x) (&sstat[0][0])[i] = 0.f;
  for (int i = threadIdx.x; i < UM_MAX_COUT + 128; i += blockDim.x) sbias[i] = (p.bias && i < p.Cout) ? p.bias[i] : 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ================================ TMA producer (warp-uniform, one elected lane issues) ================================
    {
      PROF_DECL();
      int stage = 0; uint32_t phase = 0;
      TileWalker tw_(blockIdx.x, gridDim.x, p);
      for (int t = blockIdx.x; t < (int)p.num_tiles; t += gridDim.x, tw_.advance(p)) {
        const TileCoord tc = tw_.coord();
        for (int kd = 0; kd < 3; ++kd) {
          const int dz = tc.d + kd - 1;
          if (dz < 0 || dz >= p.D) continue;
          for (int c = 0; c < nchunks; ++c) {
            PROF_T0(); mbar_wait(empty0 + 8 * stage, phase ^ 1, 100 + stage); PROF_ADD(0);
            const uint32_t sa = smem0 + stage * stage_bytes;
            const uint32_t fb = full0 + 8 * stage;
            const int k0 = c * UM_KC;
            const bool src0 = k0 < p.C0;
            const CUtensorMap* map = src0 ? &mapA0 : &mapA1;
            const int C8 = (src0 ? p.C0 : p.C1) / 8;
            const int ch8 = (src0 ? k0 : k0 - p.C0) / 8;
            const __nv_bfloat16* wsrc = p.wp + ((((long long)tc.nt * 3 + kd) * nchunks + c) * (long long)(b_bytes / 2));
            if (elect_one()) {
              mbar_expect_tx(fb, stage_bytes);
              for (int pl = 0; pl < P; ++pl)
                tma_load_4d(sa + pl * UM_A_PLANE_BYTES, map, fb, (tc.w0 - 1) * 8, tc.h0 - 1, dz, (pl * p.B + tc.b) * C8 + ch8);
              bulk_load(sa + a_bytes, wsrc, b_bytes, fb);
            }
            __syncwarp();
            PROF_ADD(1);
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (lane == 0) PROF_FLUSH(0, 1);
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (warp-uniform, one elected lane issues) ================================
    {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      PROF_DECL();
      TileWalker tw_(blockIdx.x, gridDim.x, p);
      for (int t = blockIdx.x; t < (int)p.num_tiles; t += gridDim.x, tw_.advance(p)) {
        const TileCoord tc = tw_.coord();
        PROF_T0(); mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1, 200 + acc); PROF_ADD(2);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * NT);
        uint32_t accumulate = 0;
        for (int kd = 0; kd < 3; ++kd) {
          const int dz = tc.d + kd - 1;
          if (dz < 0 || dz >= p.D) continue;
          for (int c = 0; c < nchunks; ++c) {
            mbar_wait(full0 + 8 * stage, phase, 300 + stage); PROF_ADD(3);
            tc_fence_after(); PROF_ADD(8);
            const uint32_t sa = smem0 + stage * stage_bytes;
            // Descriptors differ between MMAs only in the 14-bit start-address field (bits 0..13, units of 16 B):
            // build the four bases once per stage, then each MMA costs one 32-bit add per operand.
            const uint64_t a_hi0 = umma_desc(sa, UM_A_LBO, UM_A_SBO), a_lo0 = umma_desc(sa + UM_A_PLANE_BYTES, UM_A_LBO, UM_A_SBO);
            const uint64_t b_hi0 = umma_desc(sa + a_bytes, NT * 16, 128), b_lo0 = umma_desc(sa + a_bytes + b_plane, NT * 16, 128);
            const uint32_t b_step = (uint32_t)(NT * 32) >> 4;
            if (elect_one()) {
#pragma unroll
              for (int t9 = 0; t9 < 9; ++t9) {
                const uint32_t aoff = (uint32_t)((t9 / 3) * UM_HW + (t9 % 3));  // (kh * halo_w + kw) * 16 B >> 4
                const uint64_t a_hi = a_hi0 + aoff, b_hi = b_hi0 + t9 * b_step;
                umma_bf16(tmem_d, a_hi, b_hi, idesc, accumulate);
                accumulate = 1;
                if (P == 2) {
                  umma_bf16(tmem_d, a_hi, b_lo0 + t9 * b_step, idesc, 1);
                  umma_bf16(tmem_d, a_lo0 + aoff, b_hi, idesc, 1);
                }
              }
              PROF_ADD(9);
              umma_commit(empty0 + 8 * stage);
            }
            accumulate = 1;
            __syncwarp();
            PROF_ADD(4);
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
        if (elect_one()) umma_commit(tfull0 + 8 * acc);
        __syncwarp();
        if (++acc == p.n_acc) { acc = 0; acc_phase ^= 1; }
        if (prof_on) prof_acc[5] += 1;
      }
      if (lane == 0) { PROF_FLUSH(2, 5); PROF_FLUSH(8, 9); }
    }
  } else {
    // ================================ epilogue (warps 2..5) ================================
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;        // accumulator row == voxel within the tile
    const int hl = row >> 3, wl = row & 7;
    const int et = threadIdx.x - 64;      // 0..127
    int acc = 0; uint32_t acc_phase = 0;
    int cur_b = -1;
    PROF_DECL();
    PROF_T0();
    TileWalker tw_(blockIdx.x, gridDim.x, p);
    for (int t = blockIdx.x; t < (int)p.num_tiles; t += gridDim.x, tw_.advance(p)) {
      const TileCoord tc = tw_.coord();
      if (p.stats && tc.b != cur_b) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (cur_b >= 0) {
          for (int i = et; i < p.Cout * 2; i += 128) {
            const float v = (&sstat[0][0])[i];
            if (v != 0.f) atomicAdd(&p.stats[(long long)cur_b * p.Cout * 2 + i], (double)v);
            (&sstat[0][0])[i] = 0.f;
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        cur_b = tc.b;
      }
      mbar_wait(tfull0 + 8 * acc, acc_phase, 400 + acc); PROF_ADD(6);
      tc_fence_after();
      const int h = tc.h0 + hl, w = tc.w0 + wl;
      const bool valid = h < p.H && w < p.W;
      const long long vox = (((long long)tc.b * p.D + tc.d) * p.H + h) * p.W + w;
      for (int c0 = 0; c0 < NT; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT + c0), r);
        const int n = tc.nt * NT + c0;  // first global output column of this chunk
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + sbias[n + i];
        if (valid) {
          float* dst = (n < p.split) ? p.y0 + vox * p.ld0 + n : p.y1 + vox * p.ld1 + (n - p.split);
#pragma unroll
          for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
        if (p.stats) {
          // column sums over the warp's 32 rows by a transposing butterfly: every exchange halves the number of
          // columns a lane still carries (16 -> 8 -> 4 -> 2 -> 1), 16 shuffles per statistic instead of 80; lane l
          // ends with column (l >> 1): even lanes hold its sum, odd lanes its sum of squares.
          float a[16], q[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) { a[i] = valid ? v[i] : 0.f; q[i] = a[i] * a[i]; }
#pragma unroll
          for (int o = 16; o >= 2; o >>= 1) {
            const int half = o >> 1;
            const bool upper = (lane & o) != 0;
#pragma unroll
            for (int j = 0; j < half; ++j) {
              const float sa = upper ? a[j] : a[j + half], sq = upper ? q[j] : q[j + half];
              const float ka = upper ? a[j + half] : a[j], kq = upper ? q[j + half] : q[j];
              a[j] = ka + __shfl_xor_sync(0xffffffffu, sa, o);
              q[j] = kq + __shfl_xor_sync(0xffffffffu, sq, o);
            }
          }
          const float ta = a[0] + __shfl_xor_sync(0xffffffffu, a[0], 1);
          const float tq = q[0] + __shfl_xor_sync(0xffffffffu, q[0], 1);
          atomicAdd(&sstat[n + (lane >> 1)][lane & 1], (lane & 1) ? tq : ta);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      PROF_ADD(7);
      if (++acc == p.n_acc) { acc = 0; acc_phase ^= 1; }
    }
    if (et == 0) PROF_FLUSH(6, 7);
    if (p.stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (cur_b >= 0)
        for (int i = et; i < p.Cout * 2; i += 128) {
          const float v = (&sstat[0][0])[i];
          if (v != 0.f) atomicAdd(&p.stats[(long long)cur_b * p.Cout * 2 + i], (double)v);
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// weight packing: torch fp32 [Cout][Cin][27] -> bf16 [nt][kd][chunk][p][tap9][half][NT][8]
//   fwd   : B[n = co][k = ci] for tap (kd,kh,kw)
//   dgrad : B[n = ci][k = co] for the flipped tap (2-kd, 2-kh, 2-kw)
// ------------------------------------------------------------------------------------------
__global__ void pack_w_umma_k(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int Cout, int Cin, int dgrad, int NT, int P) {
  const int Nn = dgrad ? Cin : Cout, Kk = dgrad ? Cout : Cin;
  const int nchunks = Kk / 16, n_tiles = Nn / NT;
  const long long total = (long long)Nn * Kk * 27;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int e = (int)(r % 8); r /= 8;
    const int nl = (int)(r % NT); r /= NT;
    const int half = (int)(r % 2); r /= 2;
    const int t9 = (int)(r % 9); r /= 9;
    const int c = (int)(r % nchunks); r /= nchunks;
    const int kd = (int)(r % 3); r /= 3;
    const int nt = (int)r;
    const int n = nt * NT + nl, k = c * 16 + half * 8 + e;
    const int tap = kd * 9 + t9;
    const float v = dgrad ? w[((long long)k * Cin + n) * 27 + (26 - tap)] : w[((long long)n * Cin + k) * 27 + tap];
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    // destination index with the plane dimension inserted after `chunk`
    const long long stage = ((long long)nt * 3 + kd) * nchunks + c;
    const long long in_plane = (((long long)t9 * 2 + half) * NT + nl) * 8 + e;
    const long long plane_sz = 9LL * 2 * NT * 8;
    wp[(stage * P + 0) * plane_sz + in_plane] = hi;
    if (P == 2) wp[(stage * P + 1) * plane_sz + in_plane] = lo;
  }
  (void)n_tiles;
}
ICL_API int icl_pack_w_umma(const float* w, void* wp, int Cout, int Cin, int dgrad, int NT, int P, void* stream) {
  const int Nn = dgrad ? Cin : Cout, Kk = dgrad ? Cout : Cin;
  ICL_REQUIRE(Kk % 16 == 0 && Nn % NT == 0 && (P == 1 || P == 2), "pack_w_umma: unsupported shape N=%d K=%d NT=%d P=%d", Nn, Kk, NT, P);
  pack_w_umma_k<<<grid_for((long long)Nn * Kk * 27, 256), 256, 0, as_stream(stream)>>>((const float*)w, (__nv_bfloat16*)wp, Cout, Cin, dgrad, NT, P);
  ICL_LAUNCHED("pack_w_umma");
}

// N tile the kernel will use for a given number of output columns (must divide it).
ICL_API int icl_umma_ntile(int N) {
  if (N <= 0 || N % 16 != 0) return 0;
  if (N <= 128) return N;  // every multiple of 16 up to 256 is a legal UMMA N at M=128
  if (N % 128 == 0) return 128;
  if (N % 96 == 0) return 96;
  if (N % 64 == 0) return 64;
  if (N % 48 == 0) return 48;
  if (N % 32 == 0) return 32;
  return 16;
}

// ------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------
static int make_pk_map(CUtensorMap* map, const void* pk, int P, int B, int C, int D, int H, int W) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { icl_set_error("cuTensorMapEncodeTiled entry point unavailable"); return -1; }
  // (w, 8 channels) is contiguous in PK, so it is ONE tensor-map dimension of 8*W elements: a halo line is a single
  // 160-byte box row instead of ten 16-byte ones (TMA cost is per box row); out-of-range w still zero-fills.
  const cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)P * B * (C / 8)};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
  const cuuint32_t box[4] = {8 * UM_HW, UM_HL, 1, 2};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(pk), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { icl_set_error("cuTensorMapEncodeTiled failed (%d) for PK [%d,%d,%d,%d,%d,%d]", (int)r, P, B, C, D, H, W); return -1; }
  return 0;
}

ICL_API int icl_conv3d_umma_fwd(const void* pk0, int C0, const void* pk1, int C1, const void* wp, const float* bias, float* y0, int ld0,
                                float* y1, int ld1, int split, double* stats, int B, int D, int H, int W, int Cout, int P, int max_ctas,
                                void* stream) {
  ICL_REQUIRE(C0 > 0 && C0 % 16 == 0 && C1 % 16 == 0 && Cout % 16 == 0, "conv3d_umma: channels must be multiples of 16 (C0=%d C1=%d Cout=%d)", C0, C1, Cout);
  ICL_REQUIRE(Cout <= UM_MAX_COUT || stats == nullptr, "conv3d_umma: Cout=%d > %d with stats", Cout, UM_MAX_COUT);
  ICL_REQUIRE(Cout <= UM_MAX_COUT + 128, "conv3d_umma: Cout=%d > %d", Cout, UM_MAX_COUT + 128);
  ICL_REQUIRE(P == 1 || P == 2, "conv3d_umma: P must be 1 or 2");
  ICL_REQUIRE(split % 16 == 0, "conv3d_umma: split must be a multiple of 16");
  UmmaConvParams p;
  p.wp = (const __nv_bfloat16*)wp; p.bias = bias; p.y0 = y0; p.ld0 = ld0; p.y1 = y1 ? y1 : y0; p.ld1 = y1 ? ld1 : ld0;
  p.split = y1 ? split : Cout; p.stats = stats;
  if (!y1) { p.split = Cout; }
  p.B = B; p.D = D; p.H = H; p.W = W; p.C0 = C0; p.C1 = C1; p.Cout = Cout; p.P = P;
  p.NT = icl_umma_ntile(Cout);
  ICL_REQUIRE(p.NT >= 16 && Cout % p.NT == 0, "conv3d_umma: no N tile for Cout=%d", Cout);
  ICL_REQUIRE(y1 == nullptr || split % 16 == 0, "conv3d_umma: bad split");
  p.n_tiles = Cout / p.NT;
  p.prof = nullptr;
  static long long* prof_buf = nullptr;
  const bool prof = getenv("ICL_UMMA_PROF") != nullptr;
  if (prof) {
    if (!prof_buf) cudaMalloc(&prof_buf, 16 * sizeof(long long));
    cudaMemsetAsync(prof_buf, 0, 16 * sizeof(long long), as_stream(stream));
    p.prof = prof_buf;
  }
  p.tiles_h = cdiv(H, UM_TH); p.tiles_w = cdiv(W, UM_TW);
  p.num_tiles = (long long)B * D * p.tiles_h * p.tiles_w * p.n_tiles;
  ICL_REQUIRE(p.num_tiles < (1LL << 31), "conv3d_umma: too many tiles");
  const size_t stage_bytes = (size_t)P * (UM_A_PLANE_BYTES + 9 * p.NT * 32);
  int stages = (int)((200 * 1024) / stage_bytes);
  if (stages > UM_MAX_STAGES) stages = UM_MAX_STAGES;
  ICL_REQUIRE(stages >= 2, "conv3d_umma: stage of %zu bytes does not fit twice in shared memory", stage_bytes);
  p.stages = stages;
  p.n_acc = 512 / p.NT < UM_MAX_ACC ? 512 / p.NT : UM_MAX_ACC;
  { const char* e = getenv("ICL_UMMA_NACC"); if (e && atoi(e) >= 1 && atoi(e) <= p.n_acc) p.n_acc = atoi(e); }
  int cols = 32;
  while (cols < p.n_acc * p.NT) cols *= 2;
  p.tmem_cols = cols;
  CUtensorMap m0, m1;
  if (make_pk_map(&m0, pk0, P, B, C0, D, H, W)) return -1;
  if (C1 > 0) { if (make_pk_map(&m1, pk1, P, B, C1, D, H, W)) return -1; } else m1 = m0;
  const size_t smem = stage_bytes * stages + 128;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(conv3d_umma_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024));
    if (e != cudaSuccess) { icl_set_error("conv3d_umma: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
    configured = 220 * 1024;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long grid = p.num_tiles < sms ? p.num_tiles : sms;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  conv3d_umma_k<<<(unsigned)grid, 192, smem, as_stream(stream)>>>(m0, m1, p);
  if (prof) {
    long long h[16];
    cudaStreamSynchronize(as_stream(stream));
    cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost);
    const double tiles = h[5] > 0 ? (double)h[5] : 1.0;
    fprintf(stderr, "[umma prof] CTA0 tiles %lld  per tile clk: producer wait_empty %.0f issue %.0f | mma wait_tempty %.0f wait_full %.0f issue+commit %.0f | "
                    "epilogue wait_tfull %.0f work %.0f | mma: fence %.0f descs+mma issue %.0f commit %.0f\n", h[5], h[0] / tiles, h[1] / tiles, h[2] / tiles, h[3] / tiles, (h[4] + h[8] + h[9]) / tiles, h[6] / tiles, h[7] / tiles, h[8] / tiles, h[9] / tiles, h[4] / tiles);
  }
  ICL_LAUNCHED("conv3d_umma_fwd");
}
