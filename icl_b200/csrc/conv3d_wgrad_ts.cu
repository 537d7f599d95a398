// Weight gradient of the 3x3x3 convolution with the dY operand in TENSOR MEMORY (reference: autograd of nn.Conv3d in
// UnetConv3, networks/utils.py:104,107):
//
//     dW[co][ci][kd][kh][kw] = sum_{b, voxel u} dY[b, u][co] * X[b, u + (kd-1, kh-1, kw-1)][ci]
//
// a GEMM whose reduction axis K is the voxel axis and whose output (16 co x 27 taps x Cin) is tiny.  conv3d_wgrad_umma.cu
// feeds both operands from shared memory, which (a) forces M rows to be an arithmetic progression in shared memory, so taps
// can only enter M as whole-plane shifts (6 of 8 blocks useful, M = 64 at half rate), and (b) makes every MMA re-read its A
// tile from shared memory: (M + N) * 32 B per MMA is more than the 128 B/clk the SM delivers for N < 128.
//
// Here the M side lives in TMEM (tcgen05.mma ".ts" form: A from tensor memory, K-major, lane = output row):
//
//   A1 (128 lanes x 16 k) : lane = 16 * j + co, j = 3 * kd + kh for the first 8 of the 9 (kd, kh) pairs;
//                           A1[(j, co)][v] = dY[v + (1 - kd, 1 - kh, 0)][co]          (8 shifted copies of the dY tile)
//   A2 (3 x 16 lanes)     : the ninth pair (kd, kh) = (2, 2) with kw in the lanes:  A2[(kw, co)][v] = dY[v + (-1, -1, 1 - kw)][co]
//   B  (N' x 16 k)        : PK tiles of X in shared memory exactly as TMA delivers them (an 8 voxel x 8 channel brick is one
//                           MN-major UMMA core matrix).  TMA loads the X tile three times, shifted by kw - 1 voxels along w, and
//                           each copy as hi and lo plane: all 8-channel groups [kw 3][hi, lo][N / 8] sit at one constant stride,
//                           so ONE descriptor spans N' = 6N columns = every (kw, precision) combination of the N input channels.
//
//   D1[(j, co)][(kw, prec, ci)] += A1_hi * B' + A1_lo * B'   (N' = 6N)      D2[(kw, co)][(prec, ci)] += A2_hi * B'' + A2_lo * B''
//                                                                           (B'' = the unshifted copy, N'' = 2N)
//
// Why N' is made this wide: every tcgen05.mma of M = 128 costs at least ~32 clk whatever its N (measured here: N = 16, 32, 64
// all issue at 33 clk per MMA), so a 16-voxel K step is paid in MMA INSTRUCTIONS.  4 MMAs (2 x 6N + 2 x 2N columns) cover all
// 27 taps and all split-bf16 products hi*hi + hi*lo + lo*hi (+ lo*lo, free) of a K step; the precision halves of D are added in
// the epilogue.  432 of the 512 issued rows are useful and the tensor pipe reads only the B tile from shared memory.  The
// shifted dY copies are not materialised in memory: converter warps read the dY PK tile with ldmatrix.trans (8 voxels x
// 8 channels -> channel-major fragments, any voxel shift is just a different row address) and write them with
// tcgen05.st.16x128b, whose register layout is exactly ldmatrix's (probe: tools/probes/tmem_st16x128.cu).
//
// A CTA owns (16 co) x (N = 16 or 32 ci) and a contiguous range of items = (sample, 16x8 in-plane tile, plane d), walking
// along d: the three dY planes an item needs stay in a ring of plane slots (one new plane per item), the X copies are staged
// per item.  Accumulators (8N columns) stay in TMEM over the whole range and go once to a partial buffer, which
// wgrad_ts_reduce_k sums over the CTAs of a block in a fixed order (deterministic) into the torch-layout dW.
// Warps (576 threads): 0-15 converters, then epilogue: group g = warp / 4 converts K steps 2g, 2g+1 of every item into A slot g
// (one hand-off per quarter item); TMEM lane quarter = warp & 3; warp 16 = TMA producer; warp 17 = TMEM alloc + MMA issuer.
#include "umma.cuh"

#define WT_TH 16
#define WT_TW 8
#define WT_YW (WT_TW + 2)
#define WT_YH (WT_TH + 2)
#define WT_XLINE (WT_TW * 16)           // 128 B: one line of an X copy (8 voxels x 8 channels)
#define WT_XCHUNK (WT_TH * WT_XLINE)    // 2048 B: one 8-channel group of an X copy (16 lines) = stride between N groups
#define WT_YLINE (WT_YW * 16)           // 160 B: one line of the dY tile (10 voxels)
#define WT_YCHUNK (WT_YH * WT_YLINE)    // 2880 B: one 8-channel chunk of a dY plane tile (18 lines)
#define WT_YPREC (2 * WT_YCHUNK)        // 5760 B: 16 channels, one precision plane
#define WT_KSTEPS (WT_TH / 2)           // 8 K steps of 2 lines x 8 voxels per item
#define WT_GROUPS 4                     // converter groups = A slots
#define WT_KPG (WT_KSTEPS / WT_GROUPS)  // K steps per group and item
#define WT_KCOLS 32                     // TMEM columns of one K step: A1 hi 8 | A1 lo 8 | A2 hi 8 | A2 lo 8
#define WT_SLOT_COLS (WT_KPG * WT_KCOLS)
#define WT_MAX_XS 4
#define WT_MAX_RY 6
#define WT_THREADS (32 * (4 * WT_GROUPS + 2))

struct WtParams {
  float* partial;   // [cta][4][128][N]
  int B, Bx, D, H, W;
  int n_ci_tiles;   // Cin / N
  int N;            // ci per CTA (16 or 32)
  int tiles_h, tiles_w;
  long long items;  // B * tiles_h * tiles_w * D
  int splits;
  int xs, ry;       // X stages, dY plane slots
  int P;
  int Ci8, Co8;
  int tmem_cols;
};

struct WtItem { int b, h0, w0, d; };
__device__ __forceinline__ WtItem wt_decode(long long it, const WtParams& p) {
  WtItem r;
  r.d = (int)(it % p.D); it /= p.D;
  r.w0 = (int)(it % p.tiles_w) * WT_TW; it /= p.tiles_w;
  r.h0 = (int)(it % p.tiles_h) * WT_TH;
  r.b = (int)(it / p.tiles_h);
  return r;
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// registers -> TMEM, 16 lanes x 8 columns: r0 = (row t/4, col t%4), r1 = (row t/4 + 8, col t%4), r2, r3 = the same rows, col + 4
__device__ __forceinline__ void tmem_st_16x128b_x2(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}

__global__ void __launch_bounds__(WT_THREADS, 1)
conv3d_wgrad_ts_k(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapY, const WtParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * WT_MAX_XS + 2 * WT_MAX_RY + 2 * WT_GROUPS + 1];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = p.P, N = p.N, xs = p.xs, ry = p.ry;
  const uint32_t xcopy = (uint32_t)(N / 8) * WT_XCHUNK;        // one precision plane of one kw copy of the X tile
  const uint32_t xstage = 3u * P * xcopy, yslot = (uint32_t)WT_YPREC * P;
  const uint32_t smem0 = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t smem_y = smem0 + xs * xstage;
  const uint32_t xfull0 = smem_u32(&bars[0]), xempty0 = xfull0 + 8 * WT_MAX_XS, yfull0 = xempty0 + 8 * WT_MAX_XS,
                 yempty0 = yfull0 + 8 * WT_MAX_RY, afull0 = yempty0 + 8 * WT_MAX_RY, aempty0 = afull0 + 8 * WT_GROUPS,
                 dfull = aempty0 + 8 * WT_GROUPS;
  const int ob = blockIdx.x / p.splits, sp = blockIdx.x % p.splits;
  const int co_tile = ob / p.n_ci_tiles, ci_tile = ob % p.n_ci_tiles;
  const long long it0 = p.items * sp / p.splits, it1 = p.items * (sp + 1) / p.splits;
  const int n_items = (int)(it1 - it0);
  const uint32_t acs = (uint32_t)(P * N);                      // accumulator columns per (kw or A2) group: [prec][ci]

  if (threadIdx.x == 0) {
    for (int s = 0; s < WT_MAX_XS; ++s) { mbar_init(xfull0 + 8 * s, 1); mbar_init(xempty0 + 8 * s, 1); }
    for (int s = 0; s < WT_MAX_RY; ++s) { mbar_init(yfull0 + 8 * s, 1); mbar_init(yempty0 + 8 * s, 4 * WT_GROUPS); }
    for (int s = 0; s < WT_GROUPS; ++s) { mbar_init(afull0 + 8 * s, 4); mbar_init(aempty0 + 8 * s, 1); }
    mbar_init(dfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapY) : "memory");
  }
  if (warp == 4 * WT_GROUPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tmem_a0 = tmem_base + 4u * acs;

  if (warp == 4 * WT_GROUPS) {
    // ================================ TMA producer ================================
    int xst = 0; uint32_t xph = 0;
    int yst = 0; uint32_t yph = 0;
    for (int i = 0; i < n_items; ++i) {
      const WtItem w = wt_decode(it0 + i, p);
      const bool new_walk = (i == 0) || (w.d == 0);
      // dY planes d-1, d, d+1 (the first two only at the start of a walk); out-of-range planes / lines / columns zero-fill
      for (int pz = new_walk ? w.d - 1 : w.d + 1; pz <= w.d + 1; ++pz) {
        mbar_wait(yempty0 + 8 * yst, yph ^ 1, 100 + yst);
        if (elect_one()) {
          const uint32_t dst = smem_y + yst * yslot, fb = yfull0 + 8 * yst;
          mbar_expect_tx(fb, yslot);
          for (int pl = 0; pl < P; ++pl)
            tma_load_4d(dst + pl * WT_YPREC, &mapY, fb, (w.w0 - 1) * 8, w.h0 - 1, pz, (pl * p.B + w.b) * p.Co8 + co_tile * 2);
        }
        __syncwarp();
        if (++yst == ry) { yst = 0; yph ^= 1; }
      }
      mbar_wait(xempty0 + 8 * xst, xph ^ 1, 120 + xst);
      if (elect_one()) {
        const uint32_t dst = smem0 + xst * xstage, fb = xfull0 + 8 * xst;
        mbar_expect_tx(fb, xstage);
        // [kw copy][precision][N / 8 groups][16 lines][8 voxels][8 ch]
        for (int kw = 0; kw < 3; ++kw)
          for (int pl = 0; pl < P; ++pl)
            tma_load_4d(dst + (kw * P + pl) * xcopy, &mapX, fb, (w.w0 - 1 + kw) * 8, w.h0, w.d, (pl * p.Bx + w.b) * p.Ci8 + ci_tile * (N / 8));
      }
      __syncwarp();
      if (++xst == xs) { xst = 0; xph ^= 1; }
    }
  } else if (warp == 4 * WT_GROUPS + 1) {
    // ================================ MMA issuer ================================
    // kind::f16: D fp32 (bit 4), A = B = bf16 (bits 7, 10), A K-major from TMEM, B MN-major (bit 16), N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc1 = idesc0 | ((uint32_t)((3 * P * N) >> 3) << 17);   // A1: all kw copies and precisions
    const uint32_t idesc2 = idesc0 | ((uint32_t)((P * N) >> 3) << 17);       // A2: the unshifted copy, all precisions
    int xst = 0; uint32_t xph = 0;
    for (int i = 0; i < n_items; ++i) {
      mbar_wait(xfull0 + 8 * xst, xph, 300 + xst);
      const uint32_t sx = smem0 + xst * xstage;
      // B descriptors, MN-major / no swizzle: LBO = stride between the two K groups (next line), SBO = stride between 8-channel groups
      const uint64_t b1 = umma_desc(sx, WT_XLINE, WT_XCHUNK), b2 = umma_desc(sx + P * xcopy, WT_XLINE, WT_XCHUNK);
      for (int g = 0; g < WT_GROUPS; ++g) {
        mbar_wait(afull0 + 8 * g, (uint32_t)(i & 1), 320 + g);
        tc_fence_after();
        if (elect_one()) {
          // the short A2 MMAs first, the wide A1 MMAs last: the pipe still has ~100 clk of queued work while this warp waits for the next slot
#pragma unroll
          for (int k = 0; k < WT_KPG; ++k) {
            const int ks = g * WT_KPG + k;
            const uint32_t ta = tmem_a0 + (uint32_t)(g * WT_SLOT_COLS + k * WT_KCOLS);
            const uint32_t acc = (i == 0 && ks == 0) ? 0u : 1u;
            const uint32_t bo = (uint32_t)(ks * 2 * WT_TW);  // 16-byte units: two lines of 8 voxels per K step
            umma_bf16_ts(tmem_base + 3u * acs, ta + 16, b2 + bo, idesc2, acc);
            if (P == 2) umma_bf16_ts(tmem_base + 3u * acs, ta + 24, b2 + bo, idesc2, 1u);
          }
#pragma unroll
          for (int k = 0; k < WT_KPG; ++k) {
            const int ks = g * WT_KPG + k;
            const uint32_t ta = tmem_a0 + (uint32_t)(g * WT_SLOT_COLS + k * WT_KCOLS);
            const uint32_t acc = (i == 0 && ks == 0) ? 0u : 1u;
            const uint32_t bo = (uint32_t)(ks * 2 * WT_TW);
            umma_bf16_ts(tmem_base, ta, b1 + bo, idesc1, acc);
            if (P == 2) umma_bf16_ts(tmem_base, ta + 8, b1 + bo, idesc1, 1u);
          }
          umma_commit(aempty0 + 8 * g);
          if (g == WT_GROUPS - 1) umma_commit(xempty0 + 8 * xst);
        }
        __syncwarp();
      }
      if (++xst == xs) { xst = 0; xph ^= 1; }
    }
    if (elect_one()) umma_commit(dfull);
    __syncwarp();
  } else {
    // ================================ converters: dY PK tile -> shifted channel-major copies in TMEM ================================
    const int q = warp & 3, grp = warp >> 2;
    const uint32_t lane_q = (uint32_t)(q * 32) << 16;
    // ldmatrix row address of this lane: matrix m = lane / 8 -> chunk m & 1, line + (m >> 1); row = voxel lane % 8
    const uint32_t lm_off = (uint32_t)((lane >> 3) & 1) * WT_YCHUNK + (uint32_t)(lane >> 4) * WT_YLINE + (uint32_t)(lane & 7) * 16;
    // this warp's A1 shifts j = 2q, 2q+1 (kd = j / 3, kh = j % 3) and, for q >= 1, the A2 shift kw = q - 1
    const int j0 = 2 * q, j1 = 2 * q + 1;
    const int kd0 = j0 / 3, kh0 = j0 % 3, kd1 = j1 / 3, kh1 = j1 % 3;
    // tile line of dY needed by X line l under shift kh: l + (1 - kh) + 1 (the tile starts at h0 - 1); column: +1 (+ (1 - kw) for A2)
    const uint32_t koff = (uint32_t)(grp * WT_KPG * 2) * WT_YLINE;  // this group's first K step
    const uint32_t off0 = lm_off + koff + (uint32_t)(2 - kh0) * WT_YLINE + 16, off1 = lm_off + koff + (uint32_t)(2 - kh1) * WT_YLINE + 16;
    const uint32_t off2 = lm_off + koff + (uint32_t)(2 - (q - 1)) * 16;  // (kd, kh) = (2, 2): plane d - 1, line l - 1
    const uint32_t ta0 = tmem_a0 + lane_q + (uint32_t)(grp * WT_SLOT_COLS);
    int yseq = 0;  // ring position of plane d - 1 of the current item
    uint32_t yph = 0;
    for (int i = 0; i < n_items; ++i) {
      const int d = (int)((it0 + i) % p.D);
      const bool new_walk = (i == 0) || (d == 0);
      const bool last_of_walk = (i == n_items - 1) || (d == p.D - 1);
      // slots of planes d-1, d, d+1
      int s0 = yseq, s1 = yseq + 1, s2 = yseq + 2;
      uint32_t ph1 = yph, ph2 = yph;
      if (s1 >= ry) { s1 -= ry; ph1 ^= 1; }
      if (s2 >= ry) { s2 -= ry; ph2 ^= 1; }
      if (new_walk) { mbar_wait(yfull0 + 8 * s0, yph, 400); mbar_wait(yfull0 + 8 * s1, ph1, 401); }
      mbar_wait(yfull0 + 8 * s2, ph2, 402);
      // kd -> plane d + 1 - kd: kd 0 = slot s2, 1 = s1, 2 = s0
      const uint32_t a0 = smem_y + (uint32_t)(kd0 == 0 ? s2 : (kd0 == 1 ? s1 : s0)) * yslot + off0;
      const uint32_t a1 = smem_y + (uint32_t)(kd1 == 0 ? s2 : (kd1 == 1 ? s1 : s0)) * yslot + off1;
      const uint32_t a2 = smem_y + (uint32_t)s0 * yslot + off2;
      uint32_t r[WT_KPG][6][4];
#pragma unroll
      for (int k = 0; k < WT_KPG; ++k) {
        const uint32_t lo = (uint32_t)(k * 2) * WT_YLINE;
        ldmatrix_x4_trans(a0 + lo, r[k][0]);
        ldmatrix_x4_trans(a1 + lo, r[k][1]);
        if (q) ldmatrix_x4_trans(a2 + lo, r[k][2]);
        if (P == 2) {
          ldmatrix_x4_trans(a0 + lo + WT_YPREC, r[k][3]);
          ldmatrix_x4_trans(a1 + lo + WT_YPREC, r[k][4]);
          if (q) ldmatrix_x4_trans(a2 + lo + WT_YPREC, r[k][5]);
        }
      }
      mbar_wait(aempty0 + 8 * grp, (uint32_t)((i & 1) ^ 1), 420 + grp);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < WT_KPG; ++k) {
        const uint32_t ta = ta0 + (uint32_t)(k * WT_KCOLS);
        tmem_st_16x128b_x2(ta, r[k][0]);
        tmem_st_16x128b_x2(ta + (16u << 16), r[k][1]);
        if (q) tmem_st_16x128b_x2(ta + 16, r[k][2]);
        if (P == 2) {
          tmem_st_16x128b_x2(ta + 8, r[k][3]);
          tmem_st_16x128b_x2(ta + (16u << 16) + 8, r[k][4]);
          if (q) tmem_st_16x128b_x2(ta + 24, r[k][5]);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      // the shared-memory reads of this item are complete (ldmatrix results are in registers): hand the A slot to the issuer and
      // release the oldest plane (all three at the end of a walk)
      if (lane == 0) {
        mbar_arrive(afull0 + 8 * grp);
        mbar_arrive(yempty0 + 8 * s0);
        if (last_of_walk) { mbar_arrive(yempty0 + 8 * s1); mbar_arrive(yempty0 + 8 * s2); }
      }
      yseq += last_of_walk ? 3 : 1;
      if (yseq >= ry) { yseq -= ry; yph ^= 1; }
    }
    // ================================ epilogue: TMEM accumulators -> partial buffer ================================
    // accumulator group a (kw = 0, 1, 2 of A1; 3 = A2) keeps [prec][ci] in columns a * P * N ...: the precision halves are added here
    mbar_wait(dfull, 0, 500);
    tc_fence_after();
    const int row = q * 32 + lane;
    const int nblk = 4 * (N / 16);
    for (int bi = grp; bi < nblk; bi += WT_GROUPS) {
      const int a = bi / (N / 16), c0 = (bi % (N / 16)) * 16;
      float* dst = p.partial + (((long long)blockIdx.x * 4 + a) * 128 + row) * N + c0;
      uint32_t r0[16], r1[16];
      tmem_ld16(tmem_base + lane_q + (uint32_t)a * acs + (uint32_t)c0, r0);
      if (P == 2) tmem_ld16(tmem_base + lane_q + (uint32_t)a * acs + (uint32_t)(N + c0), r1);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float v = __uint_as_float(r0[k]);
        if (P == 2) v += __uint_as_float(r1[k]);
        r0[k] = n_items > 0 ? __float_as_uint(v) : 0u;
      }
#pragma unroll
      for (int k = 0; k < 16; k += 4) *reinterpret_cast<uint4*>(dst + k) = make_uint4(r0[k], r0[k + 1], r0[k + 2], r0[k + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4 * WT_GROUPS + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// dW[co][ci_off + ci][kd][kh][kw] (+)= sum over the CTAs of a block, in a fixed order.  A CTA of 256 threads = (256 / L) outputs x
// L split lanes: lane l sums splits l, l + L, ... (four independent loads in flight), the L lanes are then added in order through
// shared memory.  L = 8 for the thin layers (up to 148 splits), L = 1 for the deep ones (many blocks, few splits).
template <int L>
__global__ void __launch_bounds__(256) wgrad_ts_reduce_k(const float* __restrict__ partial, float* __restrict__ dw, int N, int n_ci_tiles,
                                                         int Cin_total, int ci_off, int splits, int accumulate) {
  constexpr int OUTS = 256 / L;
  __shared__ float red[L][OUTS];
  const int ob = blockIdx.y;
  const int co_tile = ob / n_ci_tiles, ci_tile = ob % n_ci_tiles;
  const int o = threadIdx.x % OUTS, l = threadIdx.x / OUTS;
  const int i = blockIdx.x * OUTS + o;  // 27 * 16 * N outputs per block: a multiple of 256
  const int ci = i % N, co = (i / N) % 16, tap = i / (16 * N);
  const int kw = tap % 3, j = tap / 3;  // j = 3 * kd + kh
  const int a = j < 8 ? kw : 3;
  const int row = j < 8 ? j * 16 + co : 32 * (kw + 1) + co;
  const long long sstride = (long long)4 * 128 * N;
  const float* src = partial + (((long long)ob * splits * 4 + a) * 128 + row) * N + ci;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int sp = l;
  for (; sp + 3 * L < splits; sp += 4 * L) {
    const float v0 = src[sp * sstride], v1 = src[(sp + L) * sstride], v2 = src[(sp + 2 * L) * sstride], v3 = src[(sp + 3 * L) * sstride];
    s0 += v0; s1 += v1; s2 += v2; s3 += v3;
  }
  for (; sp < splits; sp += L) s0 += src[sp * sstride];
  float s = (s0 + s1) + (s2 + s3);
  if (L > 1) {
    red[l][o] = s;
    __syncthreads();
    if (l == 0) {
#pragma unroll
      for (int k = 1; k < L; ++k) s += red[k][o];
    }
  }
  if (l == 0) {
    float* d = dw + ((long long)(co_tile * 16 + co) * Cin_total + ci_off + ci_tile * N + ci) * 27 + tap;
    *d = accumulate ? *d + s : s;
  }
}

static int make_wt_map(CUtensorMap* map, const void* pk, int P, int B, int C, int D, int H, int W, int box_w, int box_h, int box_c8) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { icl_set_error("cuTensorMapEncodeTiled entry point unavailable"); return -1; }
  const cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)P * B * (C / 8)};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
  const cuuint32_t box[4] = {(cuuint32_t)(8 * box_w), (cuuint32_t)box_h, 1, (cuuint32_t)box_c8};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(pk), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { icl_set_error("cuTensorMapEncodeTiled failed (%d) for wgrad PK [%d,%d,%d,%d,%d,%d]", (int)r, P, B, C, D, H, W); return -1; }
  return 0;
}

struct WtPlan { int N, blocks, splits; long long items; };
static WtPlan wt_plan(int Cin, int Cout, int B, int D, int H, int W) {
  WtPlan pl;
  pl.N = Cin % 32 == 0 ? 32 : 16;
  pl.blocks = (Cout / 16) * (Cin / pl.N);
  pl.items = (long long)B * cdiv(H, WT_TH) * cdiv(W, WT_TW) * D;
  int splits = pl.blocks >= 148 ? 1 : 148 / pl.blocks;
  if (splits > pl.items) splits = (int)pl.items;
  pl.splits = splits;
  return pl;
}

// Workspace of icl_conv3d_wgrad_ts in floats (0 = shape not supported).
ICL_API long long icl_conv3d_wgrad_ts_workspace(int Cin, int Cout, int B, int D, int H, int W) {
  if (Cin <= 0 || Cout <= 0 || Cin % 16 || Cout % 16 || B <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
  const WtPlan pl = wt_plan(Cin, Cout, B, D, H, W);
  return (long long)pl.blocks * pl.splits * 4 * 128 * pl.N;
}

ICL_API int icl_conv3d_wgrad_ts(const void* x_pk, int Cin, const void* dy_pk, int Cout, float* dw, int Cin_total, int ci_off, float* workspace,
                                int B, int D, int H, int W, int P, int accumulate, int Bx, void* stream) {
  ICL_REQUIRE(Cin > 0 && Cin % 16 == 0 && Cout > 0 && Cout % 16 == 0, "conv3d_wgrad_ts: channels must be multiples of 16 (Cin=%d Cout=%d)", Cin, Cout);
  ICL_REQUIRE(P == 1 || P == 2, "conv3d_wgrad_ts: P must be 1 or 2");
  if (Bx <= 0) Bx = B;
  ICL_REQUIRE(Bx >= B, "conv3d_wgrad_ts: Bx=%d < B=%d", Bx, B);
  const WtPlan pl = wt_plan(Cin, Cout, B, D, H, W);
  WtParams p;
  p.partial = workspace; p.B = B; p.Bx = Bx; p.D = D; p.H = H; p.W = W;
  p.N = pl.N; p.n_ci_tiles = Cin / pl.N;
  p.tiles_h = cdiv(H, WT_TH); p.tiles_w = cdiv(W, WT_TW);
  p.items = pl.items; p.splits = pl.splits;
  p.P = P; p.Ci8 = Cin / 8; p.Co8 = Cout / 8;
  const int cols = 4 * P * pl.N + WT_GROUPS * WT_SLOT_COLS;
  p.tmem_cols = cols <= 256 ? 256 : 512;
  const size_t xstage = (size_t)3 * P * (pl.N / 8) * WT_XCHUNK, yslot = (size_t)WT_YPREC * P;
  p.ry = WT_MAX_RY;
  int xs = (int)((212 * 1024 - p.ry * yslot) / xstage);
  if (xs > WT_MAX_XS) xs = WT_MAX_XS;
  p.xs = xs;
  CUtensorMap mx, my;
  if (make_wt_map(&mx, x_pk, P, Bx, Cin, D, H, W, WT_TW, WT_TH, pl.N / 8)) return -1;
  if (make_wt_map(&my, dy_pk, P, B, Cout, D, H, W, WT_YW, WT_YH, 2)) return -1;
  const size_t smem = xstage * xs + yslot * p.ry + 128;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv3d_wgrad_ts_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024));
    if (e != cudaSuccess) { icl_set_error("conv3d_wgrad_ts: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
    configured = true;
  }
  conv3d_wgrad_ts_k<<<(unsigned)(pl.blocks * pl.splits), WT_THREADS, smem, as_stream(stream)>>>(mx, my, p);
  icl_count_launch(1);
  const int per_block = 27 * 16 * pl.N;
  if (pl.splits >= 8)
    wgrad_ts_reduce_k<8><<<dim3(per_block / 32, pl.blocks), 256, 0, as_stream(stream)>>>(workspace, dw, pl.N, p.n_ci_tiles, Cin_total, ci_off,
                                                                                     pl.splits, accumulate);
  else
    wgrad_ts_reduce_k<1><<<dim3(per_block / 256, pl.blocks), 256, 0, as_stream(stream)>>>(workspace, dw, pl.N, p.n_ci_tiles, Cin_total, ci_off,
                                                                                      pl.splits, accumulate);
  ICL_LAUNCHED("conv3d_wgrad_ts");
}
