// "Plane-walk" variant of the tcgen05 implicit-GEMM 3x3x3 convolution for the thin layers (Cout 16..80, which hold
// ~70 % of the network's FLOPs; reference: nn.Conv3d in UnetConv3, networks/utils.py:104,107).
//
// Why: an SS-mode MMA of M = 128 voxels, K = 16 channels reads a 4 KB A slice from shared memory (~30 clk, measured)
// whatever N is, so a layer with N = Cout = 16 is bound by A reads at 1/4 of the N = 64 rate.  Here the three DEPTH taps
// are folded into N: a CTA walks along d for a fixed (h, w) tile, and input plane z is multiplied ONCE against the
// weights of all three depth taps,  D[voxel][(j, co)] += X_z[voxel + (kh, kw)][ci] * W[kd = 2 - j][kh][kw][ci][co],
// whose three column blocks j = 0, 1, 2 belong to output planes z - 1, z, z + 1.  Output planes own consecutive slots of
// a TMEM accumulator ring, so the 3 * Cout columns of one MMA land directly in the accumulators of three different output
// planes — A is read 9 times per plane instead of 27 and N is 3x wider.  An output plane is complete once input plane z + 1
// has been issued; the epilogue drains it (bias, fp32 NDHWC store, InstanceNorm statistics) and zero-fills the slot for
// its next owner (all MMAs accumulate).  The packed weights of the whole layer stay resident in shared memory.
//
// Shared-memory operand layouts are those of conv3d_umma.cu (K-major, no swizzle; an in-plane tap is a start-address
// offset into the staged halo tile).  Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 2-5 = epilogue.
#include "umma.cuh"
#include <stdlib.h>

// Optional cycle accounting of the roles (CTA 0 only; env ICL_UMMA_PROF=1), see conv3d_umma.cu
#define PROF_DECL() long long prof_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long prof_t0_ = 0; const bool prof_on = p.prof != nullptr && blockIdx.x == 0
#define PROF_T0() do { if (prof_on) prof_t0_ = clock64(); } while (0)
#define PROF_ADD(i) do { if (prof_on) { const long long n_ = clock64(); prof_acc[i] += n_ - prof_t0_; prof_t0_ = n_; } } while (0)
#define PROF_FLUSH(lo, hi) do { if (prof_on) for (int i_ = lo; i_ <= hi; ++i_) p.prof[i_] = prof_acc[i_]; } while (0)

#define WK_TH 16
#define WK_TW 8
#define WK_HL (WK_TH + 2)
#define WK_HW (WK_TW + 2)
#define WK_A_PLANE_BYTES (2 * WK_HL * WK_HW * 16)  // 5760
#define WK_A_LBO (WK_HL * WK_HW * 16)
#define WK_A_SBO (WK_HW * 16)
#define WK_MAX_STAGES 16
#define WK_MAX_SLOTS 32
#define WK_MAX_N 48  // columns per output plane (Cout, or Cin for the data gradient)

struct WalkParams {
  const __nv_bfloat16* wp;  // packed weights [chunk][plane P][tap 9][k8 2][n = (j, co)][8], kd = 2 - j
  const float* bias;
  float* y0; int ld0;
  float* y1; int ld1;
  int split;
  double* stats;            // [B][Cout][2] or null
  int B, D, H, W;
  int C0, C1, Cout;
  int tiles_h, tiles_w, nseg, seg_len;
  int num_items;
  int stages, P, slots_log2;
  uint32_t w_bytes;         // resident weights
  long long* prof;
};

struct WalkItem { int b, d0, d1, h0, w0; };
__device__ __forceinline__ WalkItem walk_item(int it, const WalkParams& p) {
  WalkItem r;
  r.w0 = (it % p.tiles_w) * WK_TW; it /= p.tiles_w;
  r.h0 = (it % p.tiles_h) * WK_TH; it /= p.tiles_h;
  const int seg = it % p.nseg;
  r.b = it / p.nseg;
  r.d0 = seg * p.seg_len;
  r.d1 = min(p.D, r.d0 + p.seg_len);
  return r;
}

__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

#define WK_THREADS 320  // warp 0 producer, warp 1 MMA issuer, warps 2-5 and 6-9: two epilogue groups draining alternate planes
__global__ void __launch_bounds__(WK_THREADS, 1)
conv3d_umma_walk_k(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1, const WalkParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[WK_MAX_STAGES + 2 * WK_MAX_SLOTS + 2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float sbias[WK_MAX_N];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = p.P, NT = p.Cout, stages = p.stages;
  const int RL = p.slots_log2, R = 1 << RL, RM = R - 1;  // ring size is a power of two: slot = k & RM, phase = (k >> RL) & 1 (no integer division in the roles)
  const uint32_t a_bytes = WK_A_PLANE_BYTES * P;
  const uint32_t smem0 = (smem_u32(smem_raw) + 127u) & ~127u;  // resident weights first, then the A stage ring
  const uint32_t sa0 = smem0 + p.w_bytes;
  // full[s]: TMA landed stage s.  pdone[k % R]: every MMA of the k-th input plane this CTA processed has completed — ONE
  // tcgen05.commit per plane releases both the plane's smem stages (producer) and the output plane it completes (epilogue).
  // tempty[k % R]: the epilogue drained and zero-filled the accumulator slot of the k-th output plane.
  const uint32_t full0 = smem_u32(&bars[0]);
  const uint32_t pdone0 = smem_u32(&bars[WK_MAX_STAGES]), tempty0 = smem_u32(&bars[WK_MAX_STAGES + WK_MAX_SLOTS]);
  const uint32_t wfull = smem_u32(&bars[WK_MAX_STAGES + 2 * WK_MAX_SLOTS]), tready = wfull + 8;
  const int nchunks = (p.C0 + p.C1) / 16;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(full0 + 8 * s, 1);
    for (int a = 0; a < R; ++a) { mbar_init(pdone0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 4); }
    mbar_init(wfull, 1);
    mbar_init(tready, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA1) : "memory");
  }
  for (int i = threadIdx.x; i < WK_MAX_N; i += blockDim.x) sbias[i] = (p.bias && i < p.Cout) ? p.bias[i] : 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      mbar_expect_tx(wfull, p.w_bytes);
      for (uint32_t off = 0; off < p.w_bytes; off += 32768u) {
        const uint32_t n = min(32768u, p.w_bytes - off);
        bulk_load(smem0 + off, reinterpret_cast<const unsigned char*>(p.wp) + off, n, wfull);
      }
    }
    __syncwarp();
    int stage = 0;
    int kin = 0;                            // running input-plane counter of this CTA
    const int ahead = stages / nchunks;     // planes that fit in the stage ring (host guarantees >= 1)
    PROF_DECL();
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      const WalkItem w = walk_item(it, p);
      const int z0 = max(0, w.d0 - 1), z1 = min(p.D - 1, w.d1);
      for (int z = z0; z <= z1; ++z, ++kin) {
        PROF_T0();
        if (kin >= ahead) {                 // the stages about to be refilled were read by input plane kin - ahead
          const int kp = kin - ahead;
          mbar_wait(pdone0 + 8 * (kp & RM), (kp >> RL) & 1, 100 + (kp & RM));
        }
        PROF_ADD(0);
        for (int c = 0; c < nchunks; ++c) {
          const uint32_t sa = sa0 + stage * a_bytes, fb = full0 + 8 * stage;
          const int k0 = c * 16;
          const bool src0 = k0 < p.C0;
          const CUtensorMap* map = src0 ? &mapA0 : &mapA1;
          const int C8 = (src0 ? p.C0 : p.C1) / 8;
          const int ch8 = (src0 ? k0 : k0 - p.C0) / 8;
          if (elect_one()) {
            mbar_expect_tx(fb, a_bytes);
            for (int pl = 0; pl < P; ++pl)
              tma_load_4d(sa + pl * WK_A_PLANE_BYTES, map, fb, (w.w0 - 1) * 8, w.h0 - 1, z, (pl * p.B + w.b) * C8 + ch8);
          }
          __syncwarp();
          if (++stage == ahead * nchunks) stage = 0;
        }
        PROF_ADD(1);
      }
    }
    if (lane == 0) PROF_FLUSH(0, 1);
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t tap_bytes = 96u * (uint32_t)NT;  // [k8 2][n 3*NT][8] bf16
    const uint32_t b_lbo = 48u * (uint32_t)NT;      // stride between the two k8 chunks = 3*NT rows of 16 B
    mbar_wait(wfull, 0, 500);
    mbar_wait(tready, 0, 501);   // accumulator ring zero-filled by the epilogue warps
    const int ring = (stages / nchunks) * nchunks;  // stages actually used (whole planes)
    int kin = 0;
    tc_fence_after();
    int stage = 0; uint32_t phase = 0;
    int kslot = 0;               // running slot counter of the NEXT output plane whose slot has not been acquired yet
    PROF_DECL();
    PROF_T0();
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      const WalkItem w = walk_item(it, p);
      const int z0 = max(0, w.d0 - 1), z1 = min(p.D - 1, w.d1);
      const int kbase = kslot;   // slot counter of output plane d0
      int acquired = w.d0;       // output planes < acquired have their slot
      for (int z = z0; z <= z1; ++z) {
        const int dlo = max(w.d0, z - 1), dhi = min(w.d1 - 1, z + 1);  // output planes this input plane feeds
        while (acquired <= dhi) {
          const int k = kbase + (acquired - w.d0);
          mbar_wait(tempty0 + 8 * (k & RM), ((k >> RL) & 1) ^ 1, 200 + (k & RM));
          ++acquired;
        }
        PROF_ADD(2);
        tc_fence_after();
        // pieces: consecutive ring slots; split where the ring wraps
        const int k_lo = kbase + (dlo - w.d0), n_planes = dhi - dlo + 1;
        const int s_lo = k_lo & RM;
        const int n1 = min(n_planes, R - s_lo), n2 = n_planes - n1;
        const int j_lo = dlo - (z - 1);  // weight column block of the first fed plane
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(full0 + 8 * stage, phase, 300 + stage);
          PROF_ADD(3);
          tc_fence_after();
          const uint32_t sa = sa0 + stage * a_bytes;
          const uint64_t a_hi0 = umma_desc(sa, WK_A_LBO, WK_A_SBO), a_lo0 = umma_desc(sa + WK_A_PLANE_BYTES, WK_A_LBO, WK_A_SBO);
          const uint32_t wb = smem0 + (uint32_t)(c * P) * 9u * tap_bytes + (uint32_t)(j_lo * NT) * 16u;
          const uint64_t b_hi0 = umma_desc(wb, b_lbo, 128), b_lo0 = umma_desc(wb + 9u * tap_bytes, b_lbo, 128);
          const uint32_t idesc1 = idesc_base | ((uint32_t)((n1 * NT) >> 3) << 17);
          const uint32_t idesc2 = idesc_base | ((uint32_t)((n2 * NT) >> 3) << 17);
          const uint32_t d1 = tmem_base + (uint32_t)(s_lo * NT), d2 = tmem_base;  // a wrapped piece restarts at slot 0
          const uint32_t b2off = (uint32_t)(n1 * NT);                               // its weight columns, in 16-byte rows
          if (elect_one()) {
#pragma unroll
            for (int t9 = 0; t9 < 9; ++t9) {
              const uint32_t aoff = (uint32_t)((t9 / 3) * WK_HW + (t9 % 3));
              const uint32_t boff = (uint32_t)t9 * (tap_bytes >> 4);
              umma_bf16(d1, a_hi0 + aoff, b_hi0 + boff, idesc1, 1);
              if (P == 2) {
                umma_bf16(d1, a_hi0 + aoff, b_lo0 + boff, idesc1, 1);
                umma_bf16(d1, a_lo0 + aoff, b_hi0 + boff, idesc1, 1);
              }
              if (n2 > 0) {
                umma_bf16(d2, a_hi0 + aoff, b_hi0 + boff + b2off, idesc2, 1);
                if (P == 2) {
                  umma_bf16(d2, a_hi0 + aoff, b_lo0 + boff + b2off, idesc2, 1);
                  umma_bf16(d2, a_lo0 + aoff, b_hi0 + boff + b2off, idesc2, 1);
                }
              }
            }
          }
          __syncwarp();
          PROF_ADD(4);
          if (++stage == ring) { stage = 0; phase ^= 1; }
        }
        // one commit per input plane: frees its stages and completes output plane z - 1 (and z, for the last plane of the volume)
        if (elect_one()) umma_commit(pdone0 + 8 * (kin & RM));
        __syncwarp();
        ++kin;
        PROF_ADD(8);
        if (prof_on) prof_acc[5] += 1;
      }
      kslot = kbase + (w.d1 - w.d0);
    }
    if (lane == 0) { PROF_FLUSH(2, 5); PROF_FLUSH(8, 8); }
  } else {
    // ================================ epilogue (warps 2..5 = group 0, warps 6..9 = group 1) ================================
    // One output plane is drained by ONE group (4 warps = the 4 TMEM lane quarters); the groups take alternate planes so
    // that two planes are in flight — the per-plane chain (barrier wake-up, tcgen05.ld, stores, tcgen05.st, arrive) is
    // latency-bound and was the limiter with a single group.
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int hl = row >> 3, wl = row & 7;
    if (grp == 0) {  // zero-fill the accumulator ring once
      for (int c0 = 0; c0 < R * NT; c0 += 16) tmem_st16_zero(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tready);
    }
    int kslot = 0;
    int kin_base = 0;  // input-plane counter of this item's first input plane
    PROF_DECL();
    PROF_T0();
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      const WalkItem w = walk_item(it, p);
      const int z0 = max(0, w.d0 - 1), z1 = min(p.D - 1, w.d1);
      const int h = w.h0 + hl, x = w.w0 + wl;
      const bool valid = h < p.H && x < p.W;
      float ssum[WK_MAX_N], ssq[WK_MAX_N];
#pragma unroll
      for (int i = 0; i < WK_MAX_N; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
      for (int d = w.d0; d < w.d1; ++d, ++kslot) {
        if ((kslot & 1) != grp) continue;
        const int slot = kslot & RM;
        const int kdone = kin_base + (min(d + 1, z1) - z0);  // output plane d is complete after input plane d + 1 (or the last one)
        mbar_wait(pdone0 + 8 * (kdone & RM), (kdone >> RL) & 1, 400 + slot);
        PROF_ADD(6);
        tc_fence_after();
        const long long vox = (((long long)w.b * p.D + d) * p.H + h) * p.W + x;
#pragma unroll
        for (int cc = 0; cc < WK_MAX_N / 16; ++cc) {
          if (cc * 16 < NT) {
            const int c0 = cc * 16;
            uint32_t r[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * NT + c0);
            tmem_ld16(taddr, r);
            tmem_st16_zero(taddr);  // the slot's next owner accumulates onto zeros
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + sbias[c0 + i];
            if (valid) {
              float* dst = (c0 < p.split) ? p.y0 + vox * p.ld0 + c0 : p.y1 + vox * p.ld1 + (c0 - p.split);
#pragma unroll
              for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
#pragma unroll
              for (int i = 0; i < 16; ++i) { ssum[cc * 16 + i] += v[i]; ssq[cc * 16 + i] += v[i] * v[i]; }
            }
          }
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * slot);
        PROF_ADD(7);
      }
      if (p.stats) {
        // per-thread column sums of the whole segment -> warp (transposing butterfly, see conv3d_umma.cu) -> global
#pragma unroll
        for (int cc = 0; cc < WK_MAX_N / 16; ++cc) {
          if (cc * 16 < NT) {
            float a[16], qv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { a[i] = ssum[cc * 16 + i]; qv[i] = ssq[cc * 16 + i]; }
#pragma unroll
            for (int o = 16; o >= 2; o >>= 1) {
              const int half = o >> 1;
              const bool upper = (lane & o) != 0;
#pragma unroll
              for (int j = 0; j < half; ++j) {
                const float sa_ = upper ? a[j] : a[j + half], sq_ = upper ? qv[j] : qv[j + half];
                const float ka = upper ? a[j + half] : a[j], kq = upper ? qv[j + half] : qv[j];
                a[j] = ka + __shfl_xor_sync(0xffffffffu, sa_, o);
                qv[j] = kq + __shfl_xor_sync(0xffffffffu, sq_, o);
              }
            }
            const float ta = a[0] + __shfl_xor_sync(0xffffffffu, a[0], 1);
            const float tq = qv[0] + __shfl_xor_sync(0xffffffffu, qv[0], 1);
            atomicAdd(&p.stats[((long long)w.b * p.Cout + cc * 16 + (lane >> 1)) * 2 + (lane & 1)], (double)((lane & 1) ? tq : ta));
          }
        }
      }
      PROF_ADD(9);
      kin_base += z1 - z0 + 1;
    }
    if (warp == 2 && lane == 0) { PROF_FLUSH(6, 7); PROF_FLUSH(9, 9); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// weight packing: torch fp32 [Cout][Cin][27] -> bf16 [chunk][plane P][tap9][k8 2][n = (j, nl)][8],  kd = 2 - j
//   fwd   : B[n = co][k = ci] for tap (kd, kh, kw);   dgrad : B[n = ci][k = co] for the flipped tap 26 - tap
__global__ void pack_w_walk_k(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int Cout, int Cin, int dgrad, int P) {
  const int Nn = dgrad ? Cin : Cout, Kk = dgrad ? Cout : Cin;
  const int nchunks = Kk / 16;
  const long long total = (long long)Nn * Kk * 27;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int e = (int)(r % 8); r /= 8;
    const int nl = (int)(r % Nn); r /= Nn;
    const int j = (int)(r % 3); r /= 3;
    const int half = (int)(r % 2); r /= 2;
    const int t9 = (int)(r % 9); r /= 9;
    const int c = (int)r;
    const int k = c * 16 + half * 8 + e;
    const int tap = (2 - j) * 9 + t9;
    const float v = dgrad ? w[((long long)k * Cin + nl) * 27 + (26 - tap)] : w[((long long)nl * Cin + k) * 27 + tap];
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    const long long plane_sz = 9LL * 2 * 3 * Nn * 8;
    const long long in_plane = ((((long long)t9 * 2 + half) * 3 + j) * Nn + nl) * 8 + e;
    wp[((long long)c * P + 0) * plane_sz + in_plane] = hi;
    if (P == 2) wp[((long long)c * P + 1) * plane_sz + in_plane] = lo;
  }
  (void)nchunks;
}
ICL_API int icl_pack_w_walk(const float* w, void* wp, int Cout, int Cin, int dgrad, int P, void* stream) {
  const int Nn = dgrad ? Cin : Cout, Kk = dgrad ? Cout : Cin;
  ICL_REQUIRE(Kk % 16 == 0 && Nn % 16 == 0 && Nn <= WK_MAX_N && (P == 1 || P == 2), "pack_w_walk: unsupported shape N=%d K=%d P=%d", Nn, Kk, P);
  pack_w_walk_k<<<grid_for((long long)Nn * Kk * 27, 256), 256, 0, as_stream(stream)>>>(w, (__nv_bfloat16*)wp, Cout, Cin, dgrad, P);
  ICL_LAUNCHED("pack_w_walk");
}

// 1 if the plane-walk kernel takes this layer: thin output, the whole packed layer resident in shared memory next to
// at least 3 operand stages, and enough depth to amortise the two halo planes of a segment.
ICL_API int icl_conv3d_umma_walk_ok(int Cin_total, int Cout, int D, int P) {
  if (Cin_total <= 0 || Cin_total % 16 || Cout % 16 || Cout < 16 || Cout > WK_MAX_N || D < 16) return 0;
  const long long w_bytes = (long long)(Cin_total / 16) * P * 9 * 96 * Cout;
  const long long min_stages = (Cin_total / 16) > 3 ? (Cin_total / 16) : 3;  // at least one whole input plane in the stage ring
  return (min_stages <= WK_MAX_STAGES && w_bytes + min_stages * P * WK_A_PLANE_BYTES <= 216 * 1024) ? 1 : 0;
}

static int make_walk_map(CUtensorMap* map, const void* pk, int P, int B, int C, int D, int H, int W) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { icl_set_error("cuTensorMapEncodeTiled entry point unavailable"); return -1; }
  const cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)P * B * (C / 8)};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
  const cuuint32_t box[4] = {8 * WK_HW, WK_HL, 1, 2};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(pk), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { icl_set_error("cuTensorMapEncodeTiled failed (%d) for PK [%d,%d,%d,%d,%d,%d]", (int)r, P, B, C, D, H, W); return -1; }
  return 0;
}

ICL_API int icl_conv3d_umma_walk_fwd(const void* pk0, int C0, const void* pk1, int C1, const void* wp, const float* bias, float* y0, int ld0,
                                     float* y1, int ld1, int split, double* stats, int B, int D, int H, int W, int Cout, int P, int max_ctas,
                                     void* stream) {
  ICL_REQUIRE(icl_conv3d_umma_walk_ok(C0 + C1, Cout, D, P) && C0 % 16 == 0 && C1 % 16 == 0, "conv3d_umma_walk: unsupported shape C0=%d C1=%d Cout=%d D=%d", C0, C1,
              Cout, D);
  ICL_REQUIRE(y1 == nullptr || split % 16 == 0, "conv3d_umma_walk: split must be a multiple of 16");
  WalkParams p;
  p.wp = (const __nv_bfloat16*)wp; p.bias = bias; p.y0 = y0; p.ld0 = ld0; p.y1 = y1 ? y1 : y0; p.ld1 = y1 ? ld1 : ld0;
  p.split = y1 ? split : Cout; p.stats = stats;
  p.B = B; p.D = D; p.H = H; p.W = W; p.C0 = C0; p.C1 = C1; p.Cout = Cout; p.P = P;
  p.tiles_h = cdiv(H, WK_TH); p.tiles_w = cdiv(W, WK_TW);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // depth segments: trade SM balance (items per wave) against the two halo planes every segment re-reads
  const int columns = B * p.tiles_h * p.tiles_w;
  int best_len = D;
  double best = -1.0;
  for (int ns = 1; ns <= D / 4; ++ns) {
    const int len = cdiv(D, ns), n = cdiv(D, len);
    const long long items = (long long)columns * n;
    const double balance = (double)items / (double)(cdiv(items, sms) * (long long)sms);
    const double eff = balance * (double)len / (double)(len + 2);
    if (eff > best + 1e-9) { best = eff; best_len = len; }
  }
  p.seg_len = best_len;
  p.nseg = cdiv(D, p.seg_len);
  p.num_items = columns * p.nseg;
  p.w_bytes = (uint32_t)((C0 + C1) / 16) * P * 9u * 96u * Cout;
  int stages = (int)((216 * 1024 - (long long)p.w_bytes) / ((long long)P * WK_A_PLANE_BYTES));
  if (stages > WK_MAX_STAGES) stages = WK_MAX_STAGES;
  p.stages = stages;
  p.slots_log2 = 0;
  while ((2 << p.slots_log2) * Cout <= 512 && (2 << p.slots_log2) <= WK_MAX_SLOTS) ++p.slots_log2;
  CUtensorMap m0, m1;
  if (make_walk_map(&m0, pk0, P, B, C0, D, H, W)) return -1;
  if (C1 > 0) { if (make_walk_map(&m1, pk1, P, B, C1, D, H, W)) return -1; } else m1 = m0;
  const size_t smem = (size_t)p.w_bytes + (size_t)stages * P * WK_A_PLANE_BYTES + 128;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv3d_umma_walk_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(224 * 1024));
    if (e != cudaSuccess) { icl_set_error("conv3d_umma_walk: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
    configured = true;
  }
  int grid = p.num_items < sms ? p.num_items : sms;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  p.prof = nullptr;
  static long long* prof_buf = nullptr;
  const bool prof = getenv("ICL_UMMA_PROF") != nullptr;
  if (prof) {
    if (!prof_buf) cudaMalloc(&prof_buf, 16 * sizeof(long long));
    cudaMemsetAsync(prof_buf, 0, 16 * sizeof(long long), as_stream(stream));
    p.prof = prof_buf;
  }
  conv3d_umma_walk_k<<<(unsigned)grid, WK_THREADS, smem, as_stream(stream)>>>(m0, m1, p);
  if (prof) {
    long long h[16];
    cudaStreamSynchronize(as_stream(stream));
    cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost);
    const double n = h[5] > 0 ? (double)h[5] : 1.0;
    fprintf(stderr, "[walk prof] CTA0 input planes %lld (items %d, seg_len %d, stages %d, slots %d)  per plane clk: producer wait_empty %.0f issue %.0f | "
                    "mma acquire %.0f wait_full %.0f issue+commit %.0f tfull-commit %.0f | epilogue(grp0) wait_tfull %.0f work %.0f stats %.0f\n",
            h[5], p.num_items, p.seg_len, p.stages, 1 << p.slots_log2, h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[8] / n, h[6] / n, h[7] / n, h[9] / n);
  }
  ICL_LAUNCHED("conv3d_umma_walk_fwd");
}
