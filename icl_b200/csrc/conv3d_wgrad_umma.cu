// Weight gradient of the 3x3x3 convolution on the 5th-gen tensor cores (reference: autograd of nn.Conv3d in
// UnetConv3, networks/utils.py:104,107):
//
//     dW[co][ci][kd][kh][kw] = sum_{b, voxel v} dY[b, v][co] * X[b, v + (kd-1, kh-1, kw-1)][ci]
//
// i.e. a GEMM whose reduction axis is the VOXEL axis and whose output is tiny (Cout x 27*Cin).  Both operands
// are already in HBM as PK split-bf16 tensors [P][B][C/8][D][H][W][8] (X from the forward pass, dY from the
// InstanceNorm backward), and an (8 voxels along w) x (8 channels) brick of PK is exactly one UMMA "MN-major,
// no-swizzle" core matrix (8 K-steps 16 B apart, 8 MN elements contiguous).  So plain TMA box loads of PK
// tiles are directly tcgen05.mma operands with K = voxels; no im2col, no transposition:
//
//   A (M x K) = X halo tile, M = 64 rows = 4 consecutive depth planes x 16 input channels
//   B (N x K) = dY tile,     N = 32 cols = 2 consecutive depth planes x 16 output channels
//   D[(p,ci)][(q,co)] for in-plane tap (kh,kw) accumulates X[plane z+p] * dY[plane z+1+q]  ->  kd = p - q.
//
// With the X window 4 planes deep and the dY window 2 planes deep (advancing 2 planes per step) every
// (dY plane, kd) pair is produced exactly once and 6 of the 8 (p,q) blocks are useful — this is how the 16-channel
// layers (half of the network's FLOPs) fill the M >= 64 the tensor core needs.  The 9 in-plane taps are start-address
// offsets into the same staged X tile and own one 32-column TMEM accumulator each (288 of 512 columns).
//
// Work split: an output block = (16 ci) x (16 co); the voxel axis of a block is cut into `splits` ranges, one
// CTA each; a CTA keeps its accumulators in TMEM over its whole range and writes them once to a partial buffer,
// which wgrad_reduce_k sums in a fixed order (deterministic) into the torch-layout dW.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer, warps 2-5 = epilogue.
#include "umma.cuh"

#define WG_TH 16
#define WG_TW 8
#define WG_HL (WG_TH + 2)
#define WG_HW (WG_TW + 2)
#define WG_PX 4
#define WG_PQ 2
#define WG_A_GRP (WG_HL * WG_HW * 16)       // 2880 B: one (chunk, plane) group of the X tile = stride between M groups
#define WG_A_BYTES (2 * WG_PX * WG_A_GRP)   // 23040 B per precision plane: [chunk 2][plane 4][line 18][w 10][8 ch]
#define WG_B_GRP (WG_TH * WG_TW * 16)       // 2048 B: one (chunk, plane) group of the dY tile = stride between N groups
#define WG_B_BYTES (2 * WG_PQ * WG_B_GRP)   // 8192 B per precision plane:  [chunk 2][plane 2][line 16][w 8][8 ch]
#define WG_M 64
#define WG_N 32
#define WG_TAPS 9
#define WG_MAX_STAGES 6
#define WG_PARTIAL (WG_TAPS * WG_M * WG_N)  // floats per (block, split)

struct WgradParams {
  float* partial;
  int B, D, H, W;
  int Bx;  // samples the X tensor was ALLOCATED with (its precision-plane stride); B <= Bx samples from its front are used
  int n_co_tiles;
  int tiles_h, tiles_w, zsteps;
  int items, splits;
  int stages, P;
  int Ci8, Co8;
};

struct WgItem { int b, zs, h0, w0; };
__device__ __forceinline__ WgItem wg_decode(int it, const WgradParams& p) {
  WgItem r;
  r.zs = it % p.zsteps; it /= p.zsteps;
  r.w0 = (it % p.tiles_w) * WG_TW; it /= p.tiles_w;
  r.h0 = (it % p.tiles_h) * WG_TH;
  r.b = it / p.tiles_h;
  return r;
}

__global__ void __launch_bounds__(192, 1)
conv3d_wgrad_umma_k(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapY, const WgradParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * WG_MAX_STAGES + 1];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = p.P, stages = p.stages;
  const uint32_t a_bytes = WG_A_BYTES * P, b_bytes = WG_B_BYTES * P, stage_bytes = a_bytes + b_bytes;
  const uint32_t smem0 = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[WG_MAX_STAGES]), tfull = smem_u32(&bars[2 * WG_MAX_STAGES]);
  const int ob = blockIdx.x / p.splits, sp = blockIdx.x % p.splits;
  const int ci_tile = ob / p.n_co_tiles, co_tile = ob % p.n_co_tiles;
  const int it0 = (int)((long long)p.items * sp / p.splits), it1 = (int)((long long)p.items * (sp + 1) / p.splits);

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapY) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ================================ TMA producer (warp-uniform, one elected lane issues) ================================
    int stage = 0; uint32_t phase = 0;
    for (int it = it0; it < it1; ++it) {
      const WgItem w = wg_decode(it, p);
      mbar_wait(empty0 + 8 * stage, phase ^ 1, 100 + stage);
      const uint32_t sa = smem0 + stage * stage_bytes, fb = full0 + 8 * stage;
      if (elect_one()) {
        mbar_expect_tx(fb, stage_bytes);
        for (int pl = 0; pl < P; ++pl) {
          // X planes 2*zs-1 .. 2*zs+2 (out-of-range planes / halo rows / halo columns zero-fill)
          tma_load_4d(sa + pl * WG_A_BYTES, &mapX, fb, (w.w0 - 1) * 8, w.h0 - 1, 2 * w.zs - 1, (pl * p.Bx + w.b) * p.Ci8 + ci_tile * 2);
          // dY planes 2*zs, 2*zs+1
          tma_load_4d(sa + a_bytes + pl * WG_B_BYTES, &mapY, fb, w.w0 * 8, w.h0, 2 * w.zs, (pl * p.B + w.b) * p.Co8 + co_tile * 2);
        }
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (warp-uniform, one elected lane issues) ================================
    // kind::f16, D=f32 (bit 4), A=B=bf16 (bits 7, 10), A and B MN-major (bits 15, 16), N>>3 at bit 17, M>>4 at bit 24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(WG_N >> 3) << 17) |
                           ((uint32_t)(WG_M >> 4) << 24);
    int stage = 0; uint32_t phase = 0;
    uint32_t accumulate = 0;
    for (int it = it0; it < it1; ++it) {
      mbar_wait(full0 + 8 * stage, phase, 300 + stage);
      tc_fence_after();
      const uint32_t sa = smem0 + stage * stage_bytes, sb = sa + a_bytes;
      // descriptor fields for MN-major / no swizzle: LBO = stride between the two K groups of one MMA (next line of 8 voxels),
      // SBO = stride between 8-channel MN groups.  Only the start-address field (units of 16 B) changes between MMAs.
      const uint64_t a_hi0 = umma_desc(sa, WG_HW * 16, WG_A_GRP), a_lo0 = umma_desc(sa + WG_A_BYTES, WG_HW * 16, WG_A_GRP);
      const uint64_t b_hi0 = umma_desc(sb, WG_TW * 16, WG_B_GRP), b_lo0 = umma_desc(sb + WG_B_BYTES, WG_TW * 16, WG_B_GRP);
      if (elect_one()) {
#pragma unroll
        for (int t9 = 0; t9 < WG_TAPS; ++t9) {
          const uint32_t aoff = (uint32_t)((t9 / 3) * WG_HW + (t9 % 3));  // in 16-byte units
          const uint32_t tmem_d = tmem_base + (uint32_t)(t9 * WG_N);
#pragma unroll
          for (int j = 0; j < WG_TH / 2; ++j) {  // one MMA = K 16 voxels = 2 lines of 8
            const uint32_t ao = aoff + j * 2 * WG_HW, bo = j * 2 * WG_TW;
            umma_bf16(tmem_d, a_hi0 + ao, b_hi0 + bo, idesc, j == 0 ? accumulate : 1u);
            if (P == 2) {
              umma_bf16(tmem_d, a_hi0 + ao, b_lo0 + bo, idesc, 1);
              umma_bf16(tmem_d, a_lo0 + ao, b_hi0 + bo, idesc, 1);
            }
          }
        }
        umma_commit(empty0 + 8 * stage);
      }
      __syncwarp();
      accumulate = 1;
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(tfull);
    __syncwarp();
  } else {
    // ================================ epilogue (warps 2..5): TMEM -> partial buffer ================================
    const int q = warp & 3;  // TMEM lane quarter of this warp; an M=64 accumulator keeps rows 16q..16q+15 in its lanes 0..15
    mbar_wait(tfull, 0, 400);
    tc_fence_after();
    float* dst = p.partial + (long long)blockIdx.x * WG_PARTIAL;
    for (int t9 = 0; t9 < WG_TAPS; ++t9) {
#pragma unroll
      for (int c0 = 0; c0 < WG_N; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t9 * WG_N + c0), r);
        if (lane < 16) {
          float* o = dst + ((long long)t9 * WG_M + q * 16 + lane) * WG_N + c0;
          const bool any = it1 > it0;  // a CTA without work still defines its slot
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(o + i) = any ? make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]),
                                                                  __uint_as_float(r[i + 3]))
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// dW[co][ci_off + ci][kd][kh][kw] (+)= sum over splits and over the two (p, q) pairs with p - q = kd.
// One CTA per (output block, in-plane tap): the 64 x 32 partial tiles of all splits are summed with coalesced float4 reads
// in a fixed order (deterministic), staged in shared memory, then the 16 x 16 x 3 weights of this tap are gathered from it.
__global__ void __launch_bounds__(256) wgrad_reduce_k(const float* __restrict__ partial, float* __restrict__ dw, int Cout, int Cin, int Cin_total,
                                                      int ci_off, int n_co_tiles, int splits, int accumulate) {
  __shared__ __align__(16) float tile[WG_M * WG_N];
  const int ob = blockIdx.x / WG_TAPS, t9 = blockIdx.x % WG_TAPS;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
  const float4* base = reinterpret_cast<const float4*>(partial + ((long long)ob * splits * WG_TAPS + t9) * (WG_M * WG_N)) + threadIdx.x;
  const long long sp_stride = (long long)WG_TAPS * (WG_M * WG_N) / 4;
  for (int sp = 0; sp < splits; ++sp) {
    const float4 v0 = base[sp * sp_stride], v1 = base[sp * sp_stride + 256];
    a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
    a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
  }
  reinterpret_cast<float4*>(tile)[threadIdx.x] = a0;
  reinterpret_cast<float4*>(tile)[threadIdx.x + 256] = a1;
  __syncthreads();
  const int ci_tile = ob / n_co_tiles, co_tile = ob % n_co_tiles;
  for (int i = threadIdx.x; i < 16 * 16 * 3; i += 256) {
    const int co = i % 16, ci = (i / 16) % 16, kd = i / 256;
    const int c = ci / 8, e = ci % 8, c2 = co / 8, e2 = co % 8;
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < WG_PQ; ++q) s += tile[((c * WG_PX + q + kd) * 8 + e) * WG_N + (c2 * WG_PQ + q) * 8 + e2];
    float* d = dw + ((long long)(co_tile * 16 + co) * Cin_total + ci_off + ci_tile * 16 + ci) * 27 + kd * 9 + t9;
    *d = accumulate ? *d + s : s;
  }
  (void)Cout; (void)Cin;
}

static int make_wg_map(CUtensorMap* map, const void* pk, int P, int B, int C, int D, int H, int W, int box_w, int box_h, int box_d) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { icl_set_error("cuTensorMapEncodeTiled entry point unavailable"); return -1; }
  const cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)P * B * (C / 8)};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
  const cuuint32_t box[4] = {(cuuint32_t)(8 * box_w), (cuuint32_t)box_h, (cuuint32_t)box_d, 2};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(pk), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { icl_set_error("cuTensorMapEncodeTiled failed (%d) for wgrad PK [%d,%d,%d,%d,%d,%d]", (int)r, P, B, C, D, H, W); return -1; }
  return 0;
}

// Workspace of icl_conv3d_wgrad_umma for this shape, in slots of 9*64*32 floats (0 = shape not supported).
ICL_API int icl_conv3d_wgrad_umma_slots(int Cin, int Cout, int B, int D, int H, int W) {
  if (Cin <= 0 || Cout <= 0 || Cin % 16 || Cout % 16 || D % 2) return 0;
  const int blocks = (Cin / 16) * (Cout / 16);
  const int items = B * (D / 2) * cdiv(H, WG_TH) * cdiv(W, WG_TW);
  int splits = blocks >= 148 ? 1 : 148 / blocks;
  if (splits > items) splits = items;
  return blocks * splits;
}

ICL_API int icl_conv3d_wgrad_umma(const void* x_pk, int Cin, const void* dy_pk, int Cout, float* dw, int Cin_total, int ci_off, float* workspace,
                                  int B, int D, int H, int W, int P, int accumulate, int Bx, void* stream) {
  ICL_REQUIRE(Cin > 0 && Cin % 16 == 0 && Cout % 16 == 0, "conv3d_wgrad_umma: channels must be multiples of 16 (Cin=%d Cout=%d)", Cin, Cout);
  ICL_REQUIRE(D % 2 == 0, "conv3d_wgrad_umma: depth %d must be even", D);
  ICL_REQUIRE(P == 1 || P == 2, "conv3d_wgrad_umma: P must be 1 or 2");
  WgradParams p;
  if (Bx <= 0) Bx = B;
  ICL_REQUIRE(Bx >= B, "conv3d_wgrad_umma: Bx=%d < B=%d", Bx, B);
  p.partial = workspace; p.B = B; p.Bx = Bx; p.D = D; p.H = H; p.W = W;
  p.n_co_tiles = Cout / 16;
  p.tiles_h = cdiv(H, WG_TH); p.tiles_w = cdiv(W, WG_TW); p.zsteps = D / 2;
  p.items = B * p.zsteps * p.tiles_h * p.tiles_w;
  const int blocks = (Cin / 16) * (Cout / 16);
  p.splits = blocks >= 148 ? 1 : 148 / blocks;
  if (p.splits > p.items) p.splits = p.items;
  p.P = P; p.Ci8 = Cin / 8; p.Co8 = Cout / 8;
  const size_t stage_bytes = (size_t)P * (WG_A_BYTES + WG_B_BYTES);
  int stages = (int)((200 * 1024) / stage_bytes);
  if (stages > WG_MAX_STAGES) stages = WG_MAX_STAGES;
  p.stages = stages;
  CUtensorMap mx, my;
  if (make_wg_map(&mx, x_pk, P, Bx, Cin, D, H, W, WG_HW, WG_HL, WG_PX)) return -1;
  if (make_wg_map(&my, dy_pk, P, B, Cout, D, H, W, WG_TW, WG_TH, WG_PQ)) return -1;
  const size_t smem = stage_bytes * stages + 128;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv3d_wgrad_umma_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024));
    if (e != cudaSuccess) { icl_set_error("conv3d_wgrad_umma: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
    configured = true;
  }
  conv3d_wgrad_umma_k<<<(unsigned)(blocks * p.splits), 192, smem, as_stream(stream)>>>(mx, my, p);
  icl_count_launch(1);
  wgrad_reduce_k<<<(unsigned)(blocks * WG_TAPS), 256, 0, as_stream(stream)>>>(workspace, dw, Cout, Cin, Cin_total, ci_off, p.n_co_tiles, p.splits, accumulate);
  ICL_LAUNCHED("conv3d_wgrad_umma");
}
