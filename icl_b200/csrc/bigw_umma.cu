// Skinny GEMMs against a huge fp32 matrix on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), for the
// mlp2 = MLP(N, N, N) Linears that act over the SPATIAL axis of the proxy-attention map
// (reference: networks/unet_3D_icl.py:258-259,267; N = 13 824 at the 24^3 level, so each weight is 764 MB of fp32
// and rows = B*K*heads is 16 (K = 2) ... 128 (K = 16)).  Every weight is streamed from HBM exactly once per use:
//
//   bigw_gemm_k<0>  y[r, m]  = sum_k S[r, k] * W[m, k]      (forward of nn.Linear, S = activations)
//   bigw_gemm_k<1>  dx[r, m] = sum_k S[r, k] * W[k, m]      (data gradient,        S = dY)
//   sgd_factored_umma_k      g = dY^T X (rank R, formed in TMEM), m = mu*m + g + wd*p, p -= lr*m   (weight gradient fused
//                            into the momentum-SGD update: the 764 MB gradient never exists in HBM)
//
// Numerics: fp32-parity through split bf16 (x = hi + lo, ~16 mantissa bits) and three MMAs hi*hi + hi*lo + lo*hi into
// one fp32 TMEM accumulator — the same scheme as the convolution kernels.
//
// bigw_gemm_k.  The fp32 W tile (128 x 32, 16 KB) lands in shared memory by TMA; four converter warps read it (one thread
// per W row / column), split it into bf16 hi/lo and write it with tcgen05.st into TMEM, where it is the A operand of
// tcgen05.mma (.ts form: A from TMEM, K-major, lane = output column m).  Reading A from TMEM keeps the shared-memory
// pipe for the TMA fill, one conversion read and the small B operand (an SS-form kernel would write the converted tile
// back to shared memory and read it three more times — shared memory, not HBM, would bound it).  The skinny operand S is
// pre-split once per call into T[plane][k/8][rows][8] bf16 (bigw_pack_k): a 32-k block of it is one contiguous piece per
// plane (bulk copy), and each [rows][8] group is a run of K-major UMMA core matrices (SWIZZLE_NONE, SBO 128 B,
// LBO rows*16 B).  D[m][r] accumulates over a K slice in TMEM; the reduction axis is cut into `splits` slices so that
// tiles x splits fills 2 CTAs on every SM; partials go to [split][r][m] and bigw_finish_k adds them in a fixed order
// (deterministic) with bias / GELU.
// Warp roles (192 threads): warps 0-3 converters, then epilogue (TMEM lane quarter = warp); warp 4 TMA producer;
// warp 5 TMEM alloc + MMA issuer.
#include "umma.cuh"

#define BW_KB 32
#define BW_MT 128
#define BW_W_BYTES (BW_MT * BW_KB * 4)
#define BW_MAX_STAGES 8
#define BW_SLOT_COLS 32  // TMEM columns per A slot: 16 (hi, 32 bf16 of k) + 16 (lo)

struct BigwParams {
  float* part;             // [splits][rows][M]; with `direct`: y [M][ldy] row-major (+ r0), written by the epilogue itself
  float* pre;              // direct: optional pre-activation copy, same layout as y
  const float* bias;       // direct: optional bias[rows]
  int direct, act, ldy, valid_rows;
  const __nv_bfloat16* T;  // [2][kblocks*4][rows][8]
  long long t_plane;       // elements per precision plane of T
  int M, rows, kblocks, splits, stages, tmem_cols, acc_cols;
};

template <int TRANS>
__global__ void __launch_bounds__(192, 2) bigw_gemm_k(const __grid_constant__ CUtensorMap mapW, const BigwParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[3 * BW_MAX_STAGES + 1];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = p.rows, stages = p.stages;
  const uint32_t s_bytes = (uint32_t)rows * 64u;  // one precision plane of one 32-k block of S
  const uint32_t stage_bytes = BW_W_BYTES + 2u * s_bytes;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[BW_MAX_STAGES]), afull0 = smem_u32(&bars[2 * BW_MAX_STAGES]),
                 accfull = smem_u32(&bars[3 * BW_MAX_STAGES]);
  const int mt = blockIdx.x / p.splits, sp = blockIdx.x % p.splits;
  const int kb0 = (int)((long long)p.kblocks * sp / p.splits), kb1 = (int)((long long)p.kblocks * (sp + 1) / p.splits);

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); mbar_init(afull0 + 8 * s, 128); }
    mbar_init(accfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tmem_a0 = tmem_base + (uint32_t)p.acc_cols;

  if (warp == 4) {
    // ================================ TMA producer ================================
    int stage = 0; uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(empty0 + 8 * stage, phase ^ 1, 100 + stage);
      const uint32_t sW = smem0 + stage * stage_bytes, fb = full0 + 8 * stage;
      if (elect_one()) {
        mbar_expect_tx(fb, stage_bytes);
        if (TRANS) tma_load_2d(sW, &mapW, fb, mt * BW_MT, kb * BW_KB);   // box {128 m (contiguous), 32 k}
        else       tma_load_2d(sW, &mapW, fb, kb * BW_KB, mt * BW_MT);   // box {32 k (contiguous, 128 B, swizzled), 128 m}
        const __nv_bfloat16* src = p.T + (long long)kb * 4 * rows * 8;
        bulk_load(sW + BW_W_BYTES, src, s_bytes, fb);
        bulk_load(sW + BW_W_BYTES + s_bytes, src + p.t_plane, s_bytes, fb);
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 5) {
    // ================================ MMA issuer ================================
    // kind::f16: D fp32 (bit 4), A = B = bf16 (bits 7, 10), both K-major, N = rows, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(rows >> 3) << 17) | ((uint32_t)(BW_MT >> 4) << 24);
    int stage = 0; uint32_t phase = 0;
    uint32_t accumulate = 0;
    const uint32_t kstep = ((uint32_t)rows * 32u) >> 4;  // 16 k = two [rows][8] groups, in 16-byte descriptor units
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(afull0 + 8 * stage, phase, 200 + stage);
      mbar_wait(full0 + 8 * stage, phase, 300 + stage);
      tc_fence_after();
      const uint32_t sS = smem0 + stage * stage_bytes + BW_W_BYTES;
      const uint64_t b_hi0 = umma_desc(sS, (uint32_t)rows * 16u, 128), b_lo0 = umma_desc(sS + s_bytes, (uint32_t)rows * 16u, 128);
      const uint32_t ta = tmem_a0 + (uint32_t)(stage * BW_SLOT_COLS);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          umma_bf16_ts(tmem_base, ta + ks * 8, b_hi0 + ks * kstep, idesc, ks == 0 ? accumulate : 1u);
          umma_bf16_ts(tmem_base, ta + ks * 8, b_lo0 + ks * kstep, idesc, 1u);
          umma_bf16_ts(tmem_base, ta + 16 + ks * 8, b_hi0 + ks * kstep, idesc, 1u);
        }
        umma_commit(empty0 + 8 * stage);
      }
      __syncwarp();
      accumulate = 1;
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(accfull);
    __syncwarp();
  } else {
    // ================================ converters (fp32 W tile -> bf16 hi/lo in TMEM), then epilogue ================================
    const int row = threadIdx.x;                                  // 0..127 = output column within the tile = TMEM lane
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    int stage = 0; uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(full0 + 8 * stage, phase, 400 + stage);
      const uint32_t sW = smem0 + stage * stage_bytes;
      float v[32];
      if (TRANS) {
        // tile [32 k][128 m] (no swizzle): this thread's column; a warp reads 128 contiguous bytes per k
#pragma unroll
        for (int k = 0; k < 32; ++k)
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[k]) : "r"(sW + (uint32_t)(k * 512 + row * 4)));
      } else {
        // tile [128 m][32 k], SWIZZLE_128B: 16-byte chunk c of row m sits at chunk c ^ (m & 7)
#pragma unroll
        for (int c = 0; c < 8; ++c)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(v[4 * c]), "=f"(v[4 * c + 1]), "=f"(v[4 * c + 2]), "=f"(v[4 * c + 3])
                       : "r"(sW + (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4))));
      }
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) split2_bf16(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
      const uint32_t ta = tmem_a0 + lane_base + (uint32_t)(stage * BW_SLOT_COLS);
      tmem_st16(ta, hi);
      tmem_st16(ta + 16, lo);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(afull0 + 8 * stage);
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    mbar_wait(accfull, 0, 500);
    tc_fence_after();
    const int m = mt * BW_MT + row;
    if (p.direct) {
      // token-major GEMM (the streamed matrix is the activation, one split): y[m][c] = act(D[m][c] + bias[c]), 64 contiguous bytes per store group
      float* dst = p.part + (long long)m * p.ldy;
      float* dpre = p.pre ? p.pre + (long long)m * p.ldy : nullptr;
      for (int c0 = 0; c0 < p.valid_rows; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + lane_base + (uint32_t)c0, r);
        if (m < p.M) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + ((p.bias && c0 + i < p.valid_rows) ? p.bias[c0 + i] : 0.f);
          if (c0 + 16 <= p.valid_rows && (p.ldy & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              if (dpre) *reinterpret_cast<float4*>(dpre + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
              const float4 o = p.act ? make_float4(gelu_erf(v[i]), gelu_erf(v[i + 1]), gelu_erf(v[i + 2]), gelu_erf(v[i + 3]))
                                     : make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
              *reinterpret_cast<float4*>(dst + c0 + i) = o;
            }
          } else {
            for (int i = 0; i < 16 && c0 + i < p.valid_rows; ++i) {
              if (dpre) dpre[c0 + i] = v[i];
              dst[c0 + i] = p.act ? gelu_erf(v[i]) : v[i];
            }
          }
        }
      }
    } else {
    float* dst = p.part + (long long)sp * rows * p.M + m;
    for (int c0 = 0; c0 < rows; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_base + lane_base + (uint32_t)c0, r);
      if (m < p.M) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dst[(long long)(c0 + i) * p.M] = __uint_as_float(r[i]);   // a warp stores 128 contiguous bytes per row
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// S[rows][K] fp32 (row stride lds) * scale -> T[plane][K32/8][rows_pad][8] split bf16, written at row offset r0; rows in
// [r0 + rows, rows_pad) of a chunk and columns >= K are zero-filled when `zero_pad` (the caller packs the last piece last).
__global__ void __launch_bounds__(256) bigw_pack_k(const float* __restrict__ S, long long lds, int rows, int K, float scale, __nv_bfloat16* __restrict__ T,
                                                   int rows_pad, int r0, int k8_total, int fill_rows, long long ldk = 1) {
  // one thread per (k8 chunk, row): rows fastest, so a warp writes 512 contiguous bytes per plane
  const long long total = (long long)k8_total * fill_rows;
  const long long plane = (long long)k8_total * rows_pad * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i % fill_rows);
    const long long k8 = i / fill_rows;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const long long k = k8 * 8 + e;
      v[e] = (r < rows && k < K) ? S[(long long)r * lds + k * ldk] * scale : 0.f;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split2_bf16(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
    __nv_bfloat16* d = T + (k8 * rows_pad + r0 + r) * 8;
    *reinterpret_cast<uint4*>(d) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(d + plane) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// y[r][m] = act(sum_s part[s][r][m] + bias[m]); pre (optional) = the pre-activation value
// t_ld > 0: transposed output y[m * t_ld + r] (the token-major weight gradient: partials are [feature r][output row m])
__global__ void __launch_bounds__(256) bigw_finish_k(const float* __restrict__ part, int splits, int rows_pad, int rows, int M, const float* __restrict__ bias,
                                                     int act, float* __restrict__ y, float* __restrict__ pre, int t_ld = 0) {
  const long long total = (long long)rows * M;
  const long long sstride = (long long)rows_pad * M;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float a = part[i];
    for (int s = 1; s < splits; ++s) a += part[i + s * sstride];
    if (t_ld > 0) { y[(i % M) * t_ld + i / M] = a; continue; }
    if (bias) a += bias[i % M];
    if (pre) pre[i] = a;
    y[i] = act ? gelu_erf(a) : a;
  }
}

static inline uint32_t pow2_cols(uint32_t c) {
  uint32_t v = 32;
  while (v < c) v <<= 1;
  return v;
}

static int make_w_map(CUtensorMap* map, const float* W, int n_rows, int n_cols, int trans) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { icl_set_error("cuTensorMapEncodeTiled entry point unavailable"); return -1; }
  const cuuint64_t dims[2] = {(cuuint64_t)n_cols, (cuuint64_t)n_rows};
  const cuuint64_t strides[1] = {(cuuint64_t)n_cols * 4};
  const cuuint32_t box[2] = {trans ? (cuuint32_t)BW_MT : (cuuint32_t)BW_KB, trans ? (cuuint32_t)BW_KB : (cuuint32_t)BW_MT};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(W), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   trans ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { icl_set_error("cuTensorMapEncodeTiled failed (%d) for a %d x %d fp32 weight", (int)r, n_rows, n_cols); return -1; }
  return 0;
}

// Plan of one pass (<= 128 rows): padded rows, pipeline depth, split count, workspace.
struct BigwPlan { int rows_pad, stages, splits, m_tiles, kblocks, tmem_cols, acc_cols; size_t smem; };
static BigwPlan bigw_plan(int rows, int M, int K) {
  BigwPlan pl;
  pl.rows_pad = ((rows + 15) / 16) * 16;
  pl.m_tiles = cdiv(M, BW_MT);
  pl.kblocks = cdiv(K, BW_KB);
  const size_t stage_bytes = BW_W_BYTES + (size_t)pl.rows_pad * 128;
  pl.acc_cols = pl.rows_pad < 32 ? 32 : pl.rows_pad;
  int stages = (int)((110 * 1024) / stage_bytes);       // two CTAs per SM
  const int tm_stages = (256 - pl.acc_cols) / BW_SLOT_COLS;
  if (stages > tm_stages) stages = tm_stages;
  if (stages > BW_MAX_STAGES) stages = BW_MAX_STAGES;
  if (stages < 2) stages = 2;
  pl.stages = stages;
  pl.tmem_cols = (int)pow2_cols((uint32_t)(pl.acc_cols + stages * BW_SLOT_COLS));
  pl.smem = stage_bytes * stages + 1024;
  // splits: fill 2 CTAs x 148 SMs with whole waves
  const int slots = 296;
  int best = 1; double best_eff = 0.0;
  const int max_splits = pl.kblocks < 32 ? pl.kblocks : 32;
  for (int s = 1; s <= max_splits; ++s) {
    const long long ctas = (long long)pl.m_tiles * s;
    const double waves = (double)ctas / slots;
    const double eff = waves / (double)((ctas + slots - 1) / slots);
    // prefer fewer splits (less partial traffic) unless efficiency improves by > 3 %
    if (eff > best_eff + 0.03) { best_eff = eff; best = s; }
  }
  pl.splits = best;
  return pl;
}

// workspace bytes for icl_bigw_linear_{fwd,dgrad} with `rows` rows against a weight whose output axis has M entries and
// reduction axis K entries: packed operand T + split partials
ICL_API long long icl_bigw_workspace(int rows, int M, int K) {
  long long total = 0;
  for (int r0 = 0; r0 < rows; r0 += 128) {
    const int rr = rows - r0 < 128 ? rows - r0 : 128;
    const BigwPlan pl = bigw_plan(rr, M, K);
    const long long t = 2LL * pl.kblocks * 4 * pl.rows_pad * 8 * 2;
    const long long part = (long long)pl.splits * pl.rows_pad * M * 4;
    const long long need = ((t + 255) / 256) * 256 + part;
    if (need > total) total = need;
  }
  return total;
}

static int bigw_run(int trans, int rows, int M, int K, const float* S, const float* W, int w_rows, int w_cols, const float* bias, int act, float* y,
                    float* pre, void* workspace, void* stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(bigw_gemm_k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(112 * 1024));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bigw_gemm_k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(112 * 1024));
    if (e != cudaSuccess) { icl_set_error("bigw: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
    configured = true;
  }
  CUtensorMap map;
  if (make_w_map(&map, W, w_rows, w_cols, trans)) return -1;
  for (int r0 = 0; r0 < rows; r0 += 128) {
    const int rr = rows - r0 < 128 ? rows - r0 : 128;
    const BigwPlan pl = bigw_plan(rr, M, K);
    __nv_bfloat16* T = reinterpret_cast<__nv_bfloat16*>(workspace);
    const long long t_bytes = 2LL * pl.kblocks * 4 * pl.rows_pad * 8 * 2;
    float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ((t_bytes + 255) / 256) * 256);
    const int k8_total = pl.kblocks * 4;
    bigw_pack_k<<<grid_for((long long)k8_total * pl.rows_pad, 256), 256, 0, as_stream(stream)>>>(S + (long long)r0 * K, K, rr, K, 1.f, T, pl.rows_pad, 0,
                                                                                                   k8_total, pl.rows_pad);
    icl_count_launch(1);
    BigwParams p;
    p.part = part; p.T = T; p.t_plane = (long long)k8_total * pl.rows_pad * 8;
    p.M = M; p.rows = pl.rows_pad; p.kblocks = pl.kblocks; p.splits = pl.splits; p.stages = pl.stages; p.tmem_cols = pl.tmem_cols;
    p.acc_cols = pl.acc_cols;
    p.direct = 0; p.pre = nullptr; p.bias = nullptr; p.act = 0; p.ldy = 0; p.valid_rows = pl.rows_pad;
    const unsigned grid = (unsigned)(pl.m_tiles * pl.splits);
    if (trans) bigw_gemm_k<1><<<grid, 192, pl.smem, as_stream(stream)>>>(map, p);
    else       bigw_gemm_k<0><<<grid, 192, pl.smem, as_stream(stream)>>>(map, p);
    icl_count_launch(1);
    bigw_finish_k<<<grid_for((long long)rr * M, 256), 256, 0, as_stream(stream)>>>(part, pl.splits, pl.rows_pad, rr, M, bias, act,
                                                                                    y + (long long)r0 * M, pre ? pre + (long long)r0 * M : nullptr);
    icl_count_launch(1);
  }
  return icl_check_launch(trans ? "bigw_linear_dgrad" : "bigw_linear_fwd");
}

// y[rows, N] = act(x[rows, K] @ W[N, K]^T + bias)   (nn.Linear forward; W row-major, K % 4 == 0 for the TMA row stride)
ICL_API int icl_bigw_linear_fwd(int rows, int N, int K, const float* x, const float* W, const float* bias, float* y, float* pre, int act,
                                void* workspace, void* stream) {
  ICL_REQUIRE(rows > 0 && N > 0 && K > 0 && K % 4 == 0, "bigw_linear_fwd: bad shape rows=%d N=%d K=%d (K %% 4 == 0 required)", rows, N, K);
  ICL_REQUIRE(workspace != nullptr, "bigw_linear_fwd: workspace of icl_bigw_workspace(rows, N, K) bytes required");
  return bigw_run(0, rows, N, K, x, W, N, K, bias, act, y, pre, workspace, stream);
}

// dx[rows, K] = dy[rows, N] @ W[N, K]
ICL_API int icl_bigw_linear_dgrad(int rows, int N, int K, const float* dy, const float* W, float* dx, void* workspace, void* stream) {
  ICL_REQUIRE(rows > 0 && N > 0 && K > 0 && K % 4 == 0, "bigw_linear_dgrad: bad shape rows=%d N=%d K=%d (K %% 4 == 0 required)", rows, N, K);
  ICL_REQUIRE(workspace != nullptr, "bigw_linear_dgrad: workspace of icl_bigw_workspace(rows, K, N) bytes required");
  return bigw_run(1, rows, K, N, dy, W, N, K, nullptr, 0, dx, nullptr, workspace, stream);
}

// ---------------------------------------------------------------------------------------------------------------------
// Token-major Linears on the same kernel (nn.Linear over many tokens: Swin-UNet qkv / proj / mlp, the ICL-head token
// projections; reference networks/swinunet_icl.py:120-155, unet_3D_icl.py:277-306).  The roles swap: the STREAMED fp32
// matrix is the activation x [tokens][K] (TMA tile 128 tokens x 32 k, split to bf16 hi/lo into TMEM as the A operand), the
// small packed operand is the weight (<= 128 output features per pass), the reduction axis is short (one split), and the
// epilogue writes y[token][feature] row-major with bias / GELU itself.
//   fwd  : y[M][N]  = act(x[M][K] @ W[N][K]^T + b)      dgrad : dx[M][K] = dy[M][N] @ W[N][K]   (packs W transposed)
ICL_API long long icl_tok_linear_workspace(int N, int K) {
  const int np = N < 128 ? ((N + 15) / 16) * 16 : 128;
  return 2LL * cdiv(K, BW_KB) * 4 * np * 8 * 2 + 256;
}

static int tok_linear_run(int M, int N, int K, const float* x, const float* W, long long w_lds, long long w_ldk, const float* bias, int act, float* y,
                          float* pre, void* workspace, void* stream, const char* what) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(bigw_gemm_k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(112 * 1024));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bigw_gemm_k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(112 * 1024));
    if (e != cudaSuccess) { icl_set_error("tok_linear: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
    configured = true;
  }
  CUtensorMap map;
  if (make_w_map(&map, x, M, K, 0)) return -1;
  __nv_bfloat16* T = reinterpret_cast<__nv_bfloat16*>(workspace);
  for (int r0 = 0; r0 < N; r0 += 128) {
    const int rr = N - r0 < 128 ? N - r0 : 128;
    BigwPlan pl = bigw_plan(rr, M, K);
    pl.splits = 1;
    const int k8_total = pl.kblocks * 4;
    bigw_pack_k<<<grid_for((long long)k8_total * pl.rows_pad, 256), 256, 0, as_stream(stream)>>>(W + (long long)r0 * w_lds, w_lds, rr, K, 1.f, T, pl.rows_pad, 0,
                                                                                                   k8_total, pl.rows_pad, w_ldk);
    icl_count_launch(1);
    BigwParams p;
    p.part = y + r0; p.T = T; p.t_plane = (long long)k8_total * pl.rows_pad * 8;
    p.M = M; p.rows = pl.rows_pad; p.kblocks = pl.kblocks; p.splits = 1; p.stages = pl.stages; p.tmem_cols = pl.tmem_cols;
    p.acc_cols = pl.acc_cols;
    p.direct = 1; p.pre = pre ? pre + r0 : nullptr; p.bias = bias ? bias + r0 : nullptr; p.act = act; p.ldy = N; p.valid_rows = rr;
    bigw_gemm_k<0><<<(unsigned)pl.m_tiles, 192, pl.smem, as_stream(stream)>>>(map, p);
    icl_count_launch(1);
  }
  return icl_check_launch(what);
}

ICL_API int icl_tok_linear_fwd(int M, int N, int K, const float* x, const float* W, const float* bias, float* y, float* pre, int act, void* workspace,
                               void* stream) {
  ICL_REQUIRE(M > 0 && N > 0 && K > 0 && K % 4 == 0, "tok_linear_fwd: bad shape M=%d N=%d K=%d (K %% 4 == 0 required)", M, N, K);
  ICL_REQUIRE(workspace != nullptr, "tok_linear_fwd: workspace of icl_tok_linear_workspace(N, K) bytes required");
  return tok_linear_run(M, N, K, x, W, K, 1, bias, act, y, pre, workspace, stream, "tok_linear_fwd");
}

ICL_API int icl_tok_linear_dgrad(int M, int N, int K, const float* dy, const float* W, float* dx, void* workspace, void* stream) {
  ICL_REQUIRE(M > 0 && N > 0 && K > 0 && N % 4 == 0, "tok_linear_dgrad: bad shape M=%d N=%d K=%d (N %% 4 == 0 required)", M, N, K);
  ICL_REQUIRE(workspace != nullptr, "tok_linear_dgrad: workspace of icl_tok_linear_workspace(K, N) bytes required");
  // output features = K (rows of W^T), reduction axis = N: element (r = k, kk = n) of the packed operand is W[n][k]
  return tok_linear_run(M, K, N, dy, W, 1, K, nullptr, 0, dx, nullptr, workspace, stream, "tok_linear_dgrad");
}

//   wgrad: dW[N][K] = dy[M][N]^T @ x[M][K]: the streamed matrix is dy in [reduction = token][output row n] order (the TRANS tile path),
//          the packed operand x^T (<= 128 weight columns per pass), the token axis is cut into splits whose partials are summed in a
//          fixed order and written transposed into dW.
ICL_API long long icl_tok_linear_wgrad_workspace(int M, int N, int K) {
  const int kp = K < 128 ? ((K + 15) / 16) * 16 : 128;
  const long long t = 2LL * cdiv(M, BW_KB) * 4 * kp * 8 * 2;
  return ((t + 255) / 256) * 256 + 128LL * kp * N * 4;
}

ICL_API int icl_tok_linear_wgrad(int M, int N, int K, const float* dy, const float* x, float* dW, void* workspace, void* stream) {
  ICL_REQUIRE(M > 0 && N > 0 && K > 0 && N % 4 == 0, "tok_linear_wgrad: bad shape M=%d N=%d K=%d (N %% 4 == 0 required)", M, N, K);
  ICL_REQUIRE(workspace != nullptr, "tok_linear_wgrad: workspace of icl_tok_linear_wgrad_workspace(M, N, K) bytes required");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(bigw_gemm_k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(112 * 1024));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bigw_gemm_k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(112 * 1024));
    if (e != cudaSuccess) { icl_set_error("tok_linear_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
    configured = true;
  }
  CUtensorMap map;
  if (make_w_map(&map, dy, M, N, 1)) return -1;
  for (int r0 = 0; r0 < K; r0 += 128) {
    const int rr = K - r0 < 128 ? K - r0 : 128;
    BigwPlan pl = bigw_plan(rr, N, M);
    // few output tiles, a long reduction: cut the token axis until the CTAs fill the chip (at most 128 partial slabs)
    int splits = 296 / pl.m_tiles;
    if (splits > 128) splits = 128;
    if (splits > pl.kblocks) splits = pl.kblocks;
    if (splits < 1) splits = 1;
    pl.splits = splits;
    __nv_bfloat16* T = reinterpret_cast<__nv_bfloat16*>(workspace);
    const long long t_bytes = 2LL * pl.kblocks * 4 * pl.rows_pad * 8 * 2;
    float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ((t_bytes + 255) / 256) * 256);
    const int k8_total = pl.kblocks * 4;
    // packed operand row r = weight column r0 + r, reduction index = token: x[token][r0 + r]
    bigw_pack_k<<<grid_for((long long)k8_total * pl.rows_pad, 256), 256, 0, as_stream(stream)>>>(x + r0, 1, rr, M, 1.f, T, pl.rows_pad, 0, k8_total,
                                                                                                   pl.rows_pad, (long long)K);
    icl_count_launch(1);
    BigwParams p;
    p.part = part; p.T = T; p.t_plane = (long long)k8_total * pl.rows_pad * 8;
    p.M = N; p.rows = pl.rows_pad; p.kblocks = pl.kblocks; p.splits = pl.splits; p.stages = pl.stages; p.tmem_cols = pl.tmem_cols;
    p.acc_cols = pl.acc_cols;
    p.direct = 0; p.pre = nullptr; p.bias = nullptr; p.act = 0; p.ldy = 0; p.valid_rows = pl.rows_pad;
    bigw_gemm_k<1><<<(unsigned)(pl.m_tiles * pl.splits), 192, pl.smem, as_stream(stream)>>>(map, p);
    icl_count_launch(1);
    bigw_finish_k<<<grid_for((long long)rr * N, 256), 256, 0, as_stream(stream)>>>(part, pl.splits, pl.rows_pad, rr, N, nullptr, 0, dW + r0, nullptr, K);
    icl_count_launch(1);
  }
  return icl_check_launch("tok_linear_wgrad");
}

// =====================================================================================================================
// Fused rank-R weight gradient + momentum-SGD update of a huge fp32 weight p[N][K] (optim.SGD step for the mlp2 weights,
// train_inherent_consistent_unet_3D_BraTS.py:85-86,115, with the autograd weight gradient of nn.Linear folded in):
//     g[n][k] = sum_r dY[r][n] * X[r][k] + wd * p[n][k];   m = mu * m + g;   p -= lr * m
// The factors arrive packed by bigw_pack_k as T[plane][cols/8][Rpad][8]: a [8 cols][8 r] brick is one MN-major UMMA core
// matrix, so TMA box loads of T are tcgen05.mma operands as they are (A = dY^T: M = 128 weight rows, B = X^T: N = 256 weight
// columns, K = R).  A persistent CTA walks over 128 x 256 output tiles: the MMA warp accumulates dY^T X over R in one of
// two TMEM accumulators while the epilogue warps update the previous tile: p and m tiles (128 x 32 fp32, SWIZZLE_128B)
// stream in by TMA, each thread updates its row in shared memory with the gradient read from TMEM, and the tiles stream
// back by TMA store — 16 B of HBM traffic per parameter, all of it full-line bulk transfers, for any R.
// Warp roles (224 threads): warps 0-3 epilogue, 4 factor producer, 5 MMA issuer, 6 p/m tile producer.
// =====================================================================================================================
#define SF_MT 128
#define SF_NT 256
#define SF_RC 16                                   // factor rows per pipeline stage (one K = 16 MMA step)
#define SF_A_PLANE (SF_MT / 8 * SF_RC * 16)        // 4 KB
#define SF_B_PLANE (SF_NT / 8 * SF_RC * 16)        // 8 KB
#define SF_STAGE (2 * SF_A_PLANE + 2 * SF_B_PLANE) // 24 KB
#define SF_MAX_STAGES 6
#define SF_SL 32                                   // weight columns per p/m slice
#define SF_PM_TILE (SF_MT * SF_SL * 4)             // 16 KB
#define SF_MAX_PM_STAGES 4
#define SF_BAND 8                                  // o-tiles per rasterisation band (tile order: see sf_tile)

struct SgdFacParams {
  int tiles_n, tiles_k, num_tiles, rchunks;
  int op_stages, pm_stages;   // pipeline depths: factor stages (24 KB each) and p/m slice stages (32 KB each)
  int band;                   // weight-row tiles per rasterisation band (1 = row-major tile order)
  int a_groups, b_groups;  // N/8 and K/8 (outer extent of one precision plane in the factor maps)
  float mu, wd;
  const float* lr;
};

// Tile order.  A wave of 148 consecutive tiles covers SF_BAND weight-row tiles x ~18 weight-column tiles, so the factor slabs it
// reads (8 x A + 18 x B) are shared through L2 by the whole wave instead of every CTA of a wave reading a different B slab
// (at R = 1024 the packed factors are 113 MB: row-major order re-reads them from HBM, 8.9 GB per weight).
__device__ __forceinline__ void sf_tile(int t, const SgdFacParams& p, int& tn, int& tk) {
  const int per_band = p.band * p.tiles_k;
  const int band = t / per_band, local = t - band * per_band;
  const int rows = min(p.band, p.tiles_n - band * p.band);
  tk = local / rows;
  tn = band * p.band + (local - tk * rows);
}

__global__ void __launch_bounds__(224, 1)
sgd_factored_umma_k(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapP,
                    const __grid_constant__ CUtensorMap mapM, const SgdFacParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * SF_MAX_STAGES + 2 * SF_MAX_PM_STAGES + 4];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int SF_STAGES = p.op_stages, SF_PM_STAGES = p.pm_stages;
  const uint32_t pm0 = smem0;                                       // p/m stages first (1024-byte aligned swizzled tiles)
  const uint32_t op0 = smem0 + SF_PM_STAGES * 2 * SF_PM_TILE;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[SF_MAX_STAGES]);
  const uint32_t pmfull0 = smem_u32(&bars[2 * SF_MAX_STAGES]), pmempty0 = smem_u32(&bars[2 * SF_MAX_STAGES + SF_MAX_PM_STAGES]);
  const uint32_t accfull0 = smem_u32(&bars[2 * SF_MAX_STAGES + 2 * SF_MAX_PM_STAGES]), accempty0 = accfull0 + 16;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SF_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int s = 0; s < SF_PM_STAGES; ++s) { mbar_init(pmfull0 + 8 * s, 1); mbar_init(pmempty0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(accfull0 + 8 * a, 1); mbar_init(accempty0 + 8 * a, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapP) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapM) : "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 4) {
    // ================================ factor producer ================================
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      int tn, tk;
      sf_tile(t, p, tn, tk);
      for (int rc = 0; rc < p.rchunks; ++rc) {
        mbar_wait(empty0 + 8 * stage, phase ^ 1, 100 + stage);
        const uint32_t sa = op0 + stage * SF_STAGE, fb = full0 + 8 * stage;
        if (elect_one()) {
          mbar_expect_tx(fb, SF_STAGE);
          for (int pl = 0; pl < 2; ++pl) {
            tma_load_2d(sa + pl * SF_A_PLANE, &mapA, fb, rc * SF_RC * 8, pl * p.a_groups + tn * (SF_MT / 8));
            tma_load_2d(sa + 2 * SF_A_PLANE + pl * SF_B_PLANE, &mapB, fb, rc * SF_RC * 8, pl * p.b_groups + tk * (SF_NT / 8));
          }
        }
        __syncwarp();
        if (++stage == SF_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    // ================================ MMA issuer ================================
    // A and B MN-major (bits 15, 16), N = 256, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(SF_NT >> 3) << 17) | ((uint32_t)(SF_MT >> 4) << 24);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      mbar_wait(accempty0 + 8 * acc, acc_phase ^ 1, 200 + acc);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * SF_NT);
      uint32_t accumulate = 0;
      for (int rc = 0; rc < p.rchunks; ++rc) {
        mbar_wait(full0 + 8 * stage, phase, 300 + stage);
        tc_fence_after();
        const uint32_t sa = op0 + stage * SF_STAGE, sb = sa + 2 * SF_A_PLANE;
        // MN-major, no swizzle: core matrix = [8 r][8 cols] (128 B); LBO = next group of 8 r (128 B), SBO = next group of 8 columns
        const uint64_t a_hi0 = umma_desc(sa, 128, SF_RC * 16), a_lo0 = umma_desc(sa + SF_A_PLANE, 128, SF_RC * 16);
        const uint64_t b_hi0 = umma_desc(sb, 128, SF_RC * 16), b_lo0 = umma_desc(sb + SF_B_PLANE, 128, SF_RC * 16);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < SF_RC / 16; ++ks) {
            const uint32_t o = (uint32_t)(ks * 256) >> 4;  // 16 r = 256 B
            umma_bf16(tmem_d, a_hi0 + o, b_hi0 + o, idesc, ks == 0 ? accumulate : 1u);
            umma_bf16(tmem_d, a_hi0 + o, b_lo0 + o, idesc, 1u);
            umma_bf16(tmem_d, a_lo0 + o, b_hi0 + o, idesc, 1u);
          }
          umma_commit(empty0 + 8 * stage);
        }
        __syncwarp();
        accumulate = 1;
        if (++stage == SF_STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(accfull0 + 8 * acc);
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp == 6) {
    // ================================ p / m tile producer ================================
    int ps = 0; uint32_t pphase = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      int tn, tk;
      sf_tile(t, p, tn, tk);
      for (int sl = 0; sl < SF_NT / SF_SL; ++sl) {
        mbar_wait(pmempty0 + 8 * ps, pphase ^ 1, 600 + ps);
        const uint32_t sp_ = pm0 + ps * 2 * SF_PM_TILE, fb = pmfull0 + 8 * ps;
        if (elect_one()) {
          mbar_expect_tx(fb, 2 * SF_PM_TILE);
          tma_load_2d(sp_, &mapP, fb, tk * SF_NT + sl * SF_SL, tn * SF_MT);
          tma_load_2d(sp_ + SF_PM_TILE, &mapM, fb, tk * SF_NT + sl * SF_SL, tn * SF_MT);
        }
        __syncwarp();
        if (++ps == SF_PM_STAGES) { ps = 0; pphase ^= 1; }
      }
    }
  } else {
    // ================================ epilogue: TMEM gradient + streamed p, m -> updated p, m ================================
    const int row = threadIdx.x;  // 0..127
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const float lr = *p.lr, mu = p.mu, wd = p.wd;
    int ps = 0; uint32_t pphase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    int prev_ps = -1;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      int tn, tk;
      sf_tile(t, p, tn, tk);
      mbar_wait(accfull0 + 8 * acc, acc_phase, 400 + acc);
      tc_fence_after();
      for (int sl = 0; sl < SF_NT / SF_SL; ++sl) {
        uint32_t g[32];
        tmem_ld32_nowait(tmem_base + lane_base + (uint32_t)(acc * SF_NT + sl * SF_SL), g);
        mbar_wait(pmfull0 + 8 * ps, pphase, 500 + ps);
        tmem_ld_wait();
        const uint32_t sp_ = pm0 + ps * 2 * SF_PM_TILE + (uint32_t)(row * 128), sm_ = sp_ + SF_PM_TILE;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t off = (uint32_t)((c ^ (row & 7)) << 4);
          float4 pv, mv;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pv.x), "=f"(pv.y), "=f"(pv.z), "=f"(pv.w) : "r"(sp_ + off));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(mv.x), "=f"(mv.y), "=f"(mv.z), "=f"(mv.w) : "r"(sm_ + off));
          mv.x = mu * mv.x + (__uint_as_float(g[4 * c + 0]) + wd * pv.x); pv.x -= lr * mv.x;
          mv.y = mu * mv.y + (__uint_as_float(g[4 * c + 1]) + wd * pv.y); pv.y -= lr * mv.y;
          mv.z = mu * mv.z + (__uint_as_float(g[4 * c + 2]) + wd * pv.z); pv.z -= lr * mv.z;
          mv.w = mu * mv.w + (__uint_as_float(g[4 * c + 3]) + wd * pv.w); pv.w -= lr * mv.w;
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sp_ + off), "f"(pv.x), "f"(pv.y), "f"(pv.z), "f"(pv.w) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sm_ + off), "f"(mv.x), "f"(mv.y), "f"(mv.z), "f"(mv.w) : "memory");
        }
        fence_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 0) {
          const uint32_t tp = pm0 + ps * 2 * SF_PM_TILE;
          tma_store_2d(&mapP, tp, tk * SF_NT + sl * SF_SL, tn * SF_MT);
          tma_store_2d(&mapM, tp + SF_PM_TILE, tk * SF_NT + sl * SF_SL, tn * SF_MT);
          bulk_commit();
          if (SF_PM_STAGES <= 2) {
            // shallow p/m pipeline (large R): refill the stage as soon as the stores have finished READING it (a few hundred
            // cycles; the other 127 threads are already on the next slice)
            bulk_wait_read<0>();
            mbar_arrive(pmempty0 + 8 * ps);
          } else if (prev_ps >= 0) {
            // the previous slice's stores have finished reading shared memory once at most one group is pending
            bulk_wait_read<1>();
            mbar_arrive(pmempty0 + 8 * prev_ps);
          }
        }
        prev_ps = ps;
        if (++ps == SF_PM_STAGES) { ps = 0; pphase ^= 1; }
      }
      tc_fence_before();
      mbar_arrive(accempty0 + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (threadIdx.x == 0) bulk_wait_all<0>();  // all stores complete (global writes visible at kernel end)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

static int make_factor_map(CUtensorMap* map, const __nv_bfloat16* T, int groups, int Rpad, int box_groups) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { icl_set_error("cuTensorMapEncodeTiled entry point unavailable"); return -1; }
  const cuuint64_t dims[2] = {(cuuint64_t)Rpad * 8, (cuuint64_t)groups * 2};
  const cuuint64_t strides[1] = {(cuuint64_t)Rpad * 16};
  const cuuint32_t box[2] = {(cuuint32_t)(SF_RC * 8), (cuuint32_t)box_groups};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(T), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { icl_set_error("cuTensorMapEncodeTiled failed (%d) for a factor map (groups %d, Rpad %d)", (int)r, groups, Rpad); return -1; }
  return 0;
}

static int make_pm_map(CUtensorMap* map, float* P, int N, int K) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { icl_set_error("cuTensorMapEncodeTiled entry point unavailable"); return -1; }
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  const cuuint32_t box[2] = {SF_SL, SF_MT};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, P, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { icl_set_error("cuTensorMapEncodeTiled failed (%d) for a %d x %d parameter", (int)r, N, K); return -1; }
  return 0;
}

// bytes of the packed-factor workspace for R total factor rows of a weight [N][K]
ICL_API long long icl_sgd_factored_workspace(int R, int N, int K) {
  const long long Rpad = ((R + SF_RC - 1) / SF_RC) * SF_RC;
  const long long n8 = (N + 7) / 8, k8 = (K + 7) / 8;
  return ((2 * n8 * Rpad * 16 + 255) / 256) * 256 + 2 * k8 * Rpad * 16;
}

// Packs one factor pair (dy [rows][N], x [rows][K], both scaled: dy by `scale`) at row offset r0 of the workspace laid out for
// R_total rows.  Call once per factor pair, then icl_sgd_factored_apply.
ICL_API int icl_sgd_factored_pack(const float* dy, const float* x, int rows, int r0, int R_total, int N, int K, float scale, void* workspace,
                                  void* stream) {
  ICL_REQUIRE(rows > 0 && r0 >= 0 && r0 + rows <= R_total && N % 8 == 0 && K % 8 == 0, "sgd_factored_pack: bad shape rows=%d r0=%d R=%d N=%d K=%d", rows,
              r0, R_total, N, K);
  const int Rpad = ((R_total + SF_RC - 1) / SF_RC) * SF_RC;
  __nv_bfloat16* TA = reinterpret_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* TB = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(workspace) + ((2LL * (N / 8) * Rpad * 16 + 255) / 256) * 256);
  // the last piece also zero-fills the padding rows [R_total, Rpad)
  const int fill = (r0 + rows == R_total) ? Rpad - r0 : rows;
  bigw_pack_k<<<grid_for((long long)(N / 8) * fill, 256), 256, 0, as_stream(stream)>>>(dy, N, rows, N, scale, TA, Rpad, r0, N / 8, fill);
  icl_count_launch(1);
  bigw_pack_k<<<grid_for((long long)(K / 8) * fill, 256), 256, 0, as_stream(stream)>>>(x, K, rows, K, 1.f, TB, Rpad, r0, K / 8, fill);
  ICL_LAUNCHED("sgd_factored_pack");
}

ICL_API int icl_sgd_factored_apply(int R_total, int N, int K, const void* workspace, float* p, float* m, const float* lr_ptr, float mu, float wd,
                                   int max_ctas, void* stream) {
  ICL_REQUIRE(R_total > 0 && N % 8 == 0 && K % 8 == 0 && K % 4 == 0, "sgd_factored_apply: bad shape R=%d N=%d K=%d", R_total, N, K);
  const int Rpad = ((R_total + SF_RC - 1) / SF_RC) * SF_RC;
  const __nv_bfloat16* TA = reinterpret_cast<const __nv_bfloat16*>(workspace);
  const __nv_bfloat16* TB = reinterpret_cast<const __nv_bfloat16*>(reinterpret_cast<const char*>(workspace) + ((2LL * (N / 8) * Rpad * 16 + 255) / 256) * 256);
  CUtensorMap ma, mb, mp, mm;
  if (make_factor_map(&ma, TA, N / 8, Rpad, SF_MT / 8)) return -1;
  if (make_factor_map(&mb, TB, K / 8, Rpad, SF_NT / 8)) return -1;
  if (make_pm_map(&mp, p, N, K)) return -1;
  if (make_pm_map(&mm, m, N, K)) return -1;
  SgdFacParams q;
  q.tiles_n = cdiv(N, SF_MT); q.tiles_k = cdiv(K, SF_NT); q.num_tiles = q.tiles_n * q.tiles_k; q.rchunks = Rpad / SF_RC;
  q.a_groups = N / 8; q.b_groups = K / 8; q.mu = mu; q.wd = wd; q.lr = lr_ptr;
  // few factor rows: the p / m stream is everything (deep p/m pipeline); many rows: the MMAs and their operand loads dominate and the
  // epilogue hides behind them through the double-buffered accumulator (deep factor pipeline)
  if (q.rchunks <= 8) { q.pm_stages = 4; q.op_stages = 3; q.band = 1; }
  else if (q.rchunks <= 24) { q.pm_stages = 3; q.op_stages = 4; q.band = 1; }
  else { q.pm_stages = 2; q.op_stages = SF_MAX_STAGES; q.band = SF_BAND; }
  const size_t smem = (size_t)q.pm_stages * 2 * SF_PM_TILE + (size_t)q.op_stages * SF_STAGE + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(sgd_factored_umma_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(226 * 1024));
    if (e != cudaSuccess) { icl_set_error("sgd_factored_apply: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
    configured = true;
  }
  int grid = max_ctas > 0 ? max_ctas : 148;
  if (grid > q.num_tiles) grid = q.num_tiles;
  sgd_factored_umma_k<<<(unsigned)grid, 224, smem, as_stream(stream)>>>(ma, mb, mp, mm, q);
  ICL_LAUNCHED("sgd_factored_apply");
}
