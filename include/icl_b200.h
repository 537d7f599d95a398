/*
 * icl_b200 — C-ABI of the B200-native (sm_100a) ICL hot path.
 *
 * The reference (zhuye98/ICL) has no native layer and no FFI: its hot path is the PyTorch
 * nn.Module / loss-callable API.  Every entry point below therefore replaces the
 * torch op(s) that a reference line dispatches; the citation after each group is that line
 * (paths relative to the reference's code/ directory).  The Python mirror of the reference API
 * (icl_b200/networks, icl_b200/utils/losses.py) reaches these through ctypes; INTEGRATION.md shows
 * the binding a reference maintainer would add.
 *
 * Conventions: plain pointers + sizes, no torch types; all pointers are DEVICE pointers unless
 * noted; `stream` is a cudaStream_t passed as void*; kernels are asynchronous on that stream and
 * hold no global mutable state; return 0 on success, negative on error (text: icl_last_error()).
 *
 * Layouts:  F32CL = float [B][D][H][W][C]  (torch channels_last_3d of a [B,C,D,H,W] tensor)
 *           PK    = bf16  [P][B][C/8][D][H][W][8], P planes (hi[, lo]) with x ~= hi + lo
 *           planar= float [B][K][d][h][w]
 */
#ifndef ICL_B200_H
#define ICL_B200_H
#ifdef __cplusplus
extern "C" {
#endif

const char* icl_last_error(void);
int icl_version(void);
unsigned long long icl_launch_count(void);

/* ---- Conv3d 3x3x3 s1 p1 (+bias) and its gradients: nn.Conv3d in UnetConv3, networks/utils.py:104,107 ---- */
/* tcgen05/TMEM/TMA implicit GEMM; also the data gradient when wp was packed with dgrad=1.
   Emits per-(b,co) sum / sum-of-squares for the InstanceNorm3d that follows (networks/utils.py:105,108). */
int icl_conv3d_umma_fwd(const void* pk0, int C0, const void* pk1, int C1, const void* wp, const float* bias, float* y0, int ld0,
                        float* y1, int ld1, int split, double* stats, int B, int D, int H, int W, int Cout, int P, int max_ctas,
                        void* stream);
/* plane-walk variant for thin layers (Cout <= 48, whole packed layer resident in shared memory): the three depth taps are
   folded into N and accumulate into a ring of TMEM slots owned by consecutive output planes.  Same arguments as above;
   weights packed by icl_pack_w_walk; icl_conv3d_umma_walk_ok() says whether a layer qualifies. */
int icl_conv3d_umma_walk_ok(int Cin_total, int Cout, int D, int P);
int icl_pack_w_walk(const float* w, void* wp, int Cout, int Cin, int dgrad, int P, void* stream);
int icl_conv3d_umma_walk_fwd(const void* pk0, int C0, const void* pk1, int C1, const void* wp, const float* bias, float* y0, int ld0,
                             float* y1, int ld1, int split, double* stats, int B, int D, int H, int W, int Cout, int P, int max_ctas,
                             void* stream);
int icl_pack_w_umma(const float* w, void* wp, int Cout, int Cin, int dgrad, int NT, int P, void* stream);
int icl_umma_ntile(int N);
/* fp32 CUDA-core path for the Cin=1 stem and channel counts that are not multiples of 16. */
int icl_conv3d_direct_fwd(const float* x0, int C0, const float* x1, int C1, const float* wp, const float* bias, float* y, int ldy,
                          int y_coff, double* stats, int B, int D, int H, int W, int Cout, void* stream);
/* Cin = 1 stem (conv1.conv1 at full resolution): bandwidth-bound specialisations, torch weight layout [16][1][27] */
int icl_conv3d_stem_fwd(const float* x, const float* w, const float* bias, float* y, double* stats, int B, int D, int H, int W, int Cout,
                        void* stream);
int icl_conv3d_stem_wgrad(const float* x, const float* dy, float* dw, int B, int D, int H, int W, int Cout, void* stream);
int icl_repack_w_f32(const float* w, float* wp, int Cout, int Cin, int dgrad, void* stream);
/* tcgen05 weight gradient: reduction over voxels, PK operands used in place as MN-major UMMA tiles; `workspace` holds
   icl_conv3d_wgrad_umma_slots() * 9*64*32 floats of per-CTA partial sums, reduced in a fixed order into dw. */
int icl_conv3d_wgrad_umma_slots(int Cin, int Cout, int B, int D, int H, int W);
int icl_conv3d_wgrad_umma(const void* x_pk, int Cin, const void* dy_pk, int Cout, float* dw, int Cin_total, int ci_off, float* workspace, int B,
                          int D, int H, int W, int P, int accumulate, int Bx, void* stream);
/* tcgen05 weight gradient, dY operand in tensor memory (conv3d_wgrad_ts.cu): converter warps turn the dY PK tile into 8 (kd,kh)-shifted
   channel-major copies in TMEM (ldmatrix.trans -> tcgen05.st), the X PK tile is the shared-memory B operand; 4 MMAs of M = 128 cover
   the 27 taps of a 16-voxel K step.  Any depth.  `workspace`: icl_conv3d_wgrad_ts_workspace() floats of per-CTA partial sums. */
long long icl_conv3d_wgrad_ts_workspace(int Cin, int Cout, int B, int D, int H, int W);
int icl_conv3d_wgrad_ts(const void* x_pk, int Cin, const void* dy_pk, int Cout, float* dw, int Cin_total, int ci_off, float* workspace, int B,
                        int D, int H, int W, int P, int accumulate, int Bx, void* stream);
int icl_conv3d_wgrad(const float* x, int Cx, const float* dy, int Cout, float* dw, int Cin_total, int ci_off, float* dbias, int B, int D,
                     int H, int W, void* stream);

/* ---- InstanceNorm3d(affine=False, eps 1e-5) + ReLU: networks/utils.py:105-106,108-109 ---- */
int icl_instnorm_stats(const float* y, double* stats, int B, int C, long long S, void* stream);
int icl_instnorm_finalize(const double* stats, float* mr, int B, int C, long long S, float eps, void* stream);
int icl_instnorm_relu_fwd(const float* y, const float* mr, float* a, void* pk, int write_lo, int B, int C, long long S, void* stream);
int icl_instnorm_relu_bwd(const float* dA, const float* y, const float* mr, double* red, float* dY, void* pk, int write_lo, float* dbias,
                          int B, int C, long long S, void* stream);
int icl_pack_pk(const float* x, void* pk, int write_lo, int B, int C, long long S, void* stream);
/* generalisation used by the 2D path: optional per-channel affine (gamma, beta) and LeakyReLU slope.  With the images of a 2D
   batch stacked along D of one sample, the per-(sample, channel) statistics are nn.BatchNorm2d's batch statistics, so this is
   Conv2d -> BatchNorm2d -> LeakyReLU of ConvBlock (networks/unet_icl.py:46-54).  red returns (sum g' = dbeta, sum g' yh = dgamma). */
int icl_normact_fwd(const float* y, const float* mr, const float* gamma, const float* beta, float slope, float* a, void* pk, int write_lo, int B,
                    int C, long long S, void* stream);
int icl_normact_bwd(const float* dA, const float* y, const float* mr, const float* gamma, const float* beta, float slope, double* red, float* dY,
                    void* pk, int write_lo, float* dbias, int B, int C, long long S, void* stream);

/* ---- 2D path: nn.MaxPool2d(2) (unet_icl.py:66) and nn.Upsample(scale_factor=2, bilinear, align_corners=True) (unet_icl.py:84-85) on
        F32CL [1][N images][H][W][C] ---- */
int icl_maxpool2d_fwd(const float* x, float* out, unsigned char* idx, void* pk, int write_lo, int N, int C, int H, int W, void* stream);
int icl_maxpool2d_bwd(const float* dout, const unsigned char* idx, float* dx, int accumulate, int N, int C, int H, int W, void* stream);
int icl_upsample2x_ac2d_fwd(const float* x, float* out, void* pk, int write_lo, int N, int C, int h, int w, void* stream);
int icl_upsample2x_ac2d_bwd(const float* dout, int Cd, int c_off, float* dx, int accumulate, int N, int C, int h, int w, void* stream);

/* ---- nn.MaxPool3d(2): networks/unet_3D_icl.py:41,45,49,53 (first max in scan order wins ties) ---- */
int icl_maxpool3d_fwd(const float* x, float* out, unsigned char* idx, void* pk, int write_lo, int B, int C, int D, int H, int W, void* stream);
int icl_maxpool3d_bwd(const float* dout, const unsigned char* idx, float* dx, int accumulate, int B, int C, int D, int H, int W, void* stream);

/* ---- nn.Upsample(scale_factor=2, mode='trilinear') in UnetUp3_CT: networks/utils.py:264,272 ---- */
int icl_upsample2x_fwd(const float* x, float* out, void* pk, int write_lo, int B, int C, int d, int h, int w, void* stream);
int icl_upsample2x_bwd(const float* dout, int Cd, int c_off, float* dx, int accumulate, int B, int C, int d, int h, int w, void* stream);

/* ---- nn.Dropout(p=0.3): networks/unet_3D_icl.py:67-68,110,116 (mask bytes, or Philox keyed by seed; seed_ptr = seed in device memory,
        so that a captured CUDA graph can be re-seeded per replay) ---- */
int icl_dropout(const float* x, float* out, const unsigned char* mask, unsigned long long seed, const unsigned long long* seed_ptr, float p,
                long long total, void* stream);

/* ---- `final` 1x1x1 Conv3d(16 -> K) over all voxels and its fused backward (dx, dW, db in one pass): networks/unet_3D_icl.py:65,117 ---- */
int icl_head1x1_fwd(const float* x, const float* w, const float* bias, float* out, long long rows, int C, int K, void* stream);
int icl_head1x1_bwd(const float* g, const float* x, const float* w, float* dx, float* dw, float* db, long long rows, int C, int K, void* stream);

/* ---- elementwise helpers for residual adds / DropPath row scaling: networks/unet_3D_icl.py:264-267 ---- */
int icl_axpby(const float* x, float* y, float alpha, float beta, long long n, void* stream);
int icl_row_combine(const float* a, const float* sa, const float* b, const float* sb, float* out, long long rows, long long cols, void* stream);

/* ---- nn.Linear / 1x1x1 Conv3d / Conv1d(k=1) as strided batched fp32 GEMM:
        networks/unet_3D_icl.py:65 (final), :186 (proj_layers), :196 (attn_convs1), :197 (query_convs),
        :277-280 (fc_q, fc_kv, proj), :304-306 (MLP fc1/fc2), :325 (pointwise) ---- */
int icl_sgemm(int M, int N, int K, const float* A, long long sam, long long sak, long long sA, const float* Bm, long long sbk, long long sbn,
              long long sB, float* C, long long scm, long long scn, long long sC, int batch, const float* bias, int bias_mode, int act,
              int accumulate, float* pre, void* stream);
/* mlp2 = MLP(N, N, N) over the spatial axis, networks/unet_3D_icl.py:258-259,267: weight-streaming kernels; M <= 256 rows
 * (one weight pass per block of 64 rows), K % 4 == 0 */
int icl_skinny_linear_fwd(int M, int N, int K, const float* x, const float* Wt, const float* bias, float* y, float* pre, int act, void* stream);
int icl_skinny_linear_dgrad(int M, int N, int K, const float* dy, const float* Wt, float* dx, void* stream);
int icl_outer_wgrad(int M, int N, int K, const float* dy, const float* x, float* dW, float* db, int accumulate, void* stream);
/* the same Linears on the 5th-gen tensor cores (tcgen05 / TMEM / TMA, split-bf16 x3): the fp32 weight is streamed once by TMA,
 * converted to bf16 hi/lo into TMEM (A operand), any number of rows (128 per weight pass); `workspace` = device buffer of
 * icl_bigw_workspace(rows, out_features, in_features) bytes (dgrad: (rows, in_features, out_features)) */
long long icl_bigw_workspace(int rows, int M, int K);
int icl_bigw_linear_fwd(int rows, int N, int K, const float* x, const float* W, const float* bias, float* y, float* pre, int act, void* workspace,
                        void* stream);
int icl_bigw_linear_dgrad(int rows, int N, int K, const float* dy, const float* W, float* dx, void* workspace, void* stream);
/* token-major nn.Linear on the same tcgen05 kernel (roles swapped: the streamed fp32 matrix is the activation, the packed operand the
   weight; epilogue writes y row-major with bias / GELU): Swin-UNet qkv / proj / mlp (networks/swinunet_icl.py:120-155) and the ICL-head
   token projections (networks/unet_3D_icl.py:277-306).  workspace: icl_tok_linear_workspace(output features, reduction length) bytes. */
long long icl_tok_linear_workspace(int N, int K);
int icl_tok_linear_fwd(int M, int N, int K, const float* x, const float* W, const float* bias, float* y, float* pre, int act, void* workspace,
                       void* stream);
int icl_tok_linear_dgrad(int M, int N, int K, const float* dy, const float* W, float* dx, void* workspace, void* stream);
long long icl_tok_linear_wgrad_workspace(int M, int N, int K);
int icl_tok_linear_wgrad(int M, int N, int K, const float* dy, const float* x, float* dW, void* workspace, void* stream);
int icl_colsum(const float* a, float* out, long long M, int N, int accumulate, void* stream);
int icl_gelu_bwd(const float* dy, const float* pre, float* dx, long long n, void* stream);

/* ---- nn.LayerNorm over C and over N: networks/unet_3D_icl.py:187,248-249,254,258 ---- */
int icl_layernorm_fwd(const float* x, const float* w, const float* b, float* y, float* mean_rstd, long long rows, int C, float eps, void* stream);
int icl_layernorm_bwd(const float* dy, const float* x, const float* w, const float* mean_rstd, float* dx, float* dw, float* db, long long rows,
                      int C, void* stream);

/* ---- Query_Attention (voxel -> class-proxy cross attention): networks/unet_3D_icl.py:283-297 ---- */
/* `ws`: device scratch of icl_reduce_workspace_bytes() bytes (per-chunk partial sums of the split-N reductions); `xv` in backward is
 * the forward output (sum_n p dP equals <dxv, xv>, so backward needs no reduction pass for it) */
int icl_proxy_attn_fwd(const float* ql, const float* kv, float* map, float* xv, float* mstat, int B, int N, int C, int H, int K, float scale,
                       int want_xv, float* ws, void* stream);
int icl_proxy_attn_bwd(const float* dmap, const float* dxv, const float* xv, const float* map, const float* ql, const float* kv,
                       const float* mstat, float* dl_scratch, float* dql, float* dkv, int B, int N, int C, int H, int K, float scale,
                       float* ws, void* stream);

/* ---- SeparableConv3d (depthwise 3^3 + BatchNorm3d(train) + ReLU + pointwise + BN + ReLU): networks/unet_3D_icl.py:317-345 ---- */
int icl_dwconv3d(const float* x, const float* w, float* y, int NB, int CH, int d, int h, int wd, int flip, void* stream);
/* `ws`: device scratch of icl_reduce_workspace_bytes() bytes (chunked two-stage reductions over all (sample, voxel) positions) */
int icl_reduce_workspace_bytes(void);
int icl_dwconv3d_wgrad(const float* x, const float* dy, float* dw, int NB, int CH, int d, int h, int wd, float* ws, void* stream);
int icl_bn_relu_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean_rstd, float* run_mean, float* run_var, int NB,
                    int CH, long long S, float eps, float momentum, double* ws, void* stream);
int icl_bn_relu_bwd(const float* dy, const float* x, const float* y, const float* mean_rstd, const float* gamma, float* sums, float* dx, int NB,
                    int CH, long long S, double* ws, void* stream);
int icl_planar_pw(const float* x, const float* w, int w_so, int w_si, const float* bias, float* y, int NB, int CI, int CO, long long S,
                  void* stream);
int icl_planar_pw_wgrad(const float* dy, const float* x, float* dw, float* db, int NB, int CO, int CI, long long S, float* ws, void* stream);

/* ---- losses: CrossEntropyLoss + DiceLoss (utils/losses.py:195-231), AuxLoss3D (:254-271, trilinear
        interpolation fused), PseudoSoftLoss3D / softmax_dice_loss (:287-299, :42-59), softmax_mse_loss (:68-90) ---- */
int icl_class_stats_fwd(const float* src, int planar, int rz, int ry, int rx, int B, int K, int Z, int Y, int X, const long long* labels,
                        const float* tgt, int is_prob, const float* class_w, double* sums, float* out2, void* stream);
int icl_class_stats_bwd(const float* src, int planar, int rz, int ry, int rx, int B, int K, int Z, int Y, int X, const long long* labels,
                        const float* tgt, int is_prob, const float* class_w, const double* sums, const float* g_ce, const float* g_dice,
                        float w_ce, float w_dice, float* dsrc, float* workspace, void* stream);
int icl_softmax_mse(const float* a, const float* b, int B, int K, long long S, double* sum, const float* gup, float w, float* da, void* stream);
int icl_scale_to_float(const double* s, double scale, float* out, void* stream);

/* ---- Swin (shifted-)window attention: WindowAttention.forward networks/swinunet_icl.py:120-155 with the cyclic shift, window
 * partition / reverse and shift mask of SwinTransformerBlock.forward :249-293 (mask :217-245) folded into the addressing.
 * qkv [B, H*W, 3*C] token-major (channel = which*C + head*32 + d), table [(2*ws-1)^2, nH], out / dout [B, H*W, C];
 * head_dim 32, ws*ws <= 64.  dtable must be zeroed by the caller (accumulated with atomics). ---- */
int icl_window_attn_fwd(const float* qkv, const float* table, float* out, int B, int H, int W, int C, int nH, int ws, int shift, void* stream);
int icl_window_attn_bwd(const float* qkv, const float* table, const float* dout, float* dqkv, float* dtable, int B, int H, int W, int C, int nH,
                        int ws, int shift, void* stream);

/* ---- optim.SGD(momentum=0.9, weight_decay=1e-4).step(): train_inherent_consistent_unet_3D_BraTS.py:85-86,115 ---- */
int icl_sgd_multi(const void* tab, const int* chunk_tensor, const long long* chunk_off, int n_chunks, const float* lr_ptr, float mu, float wd,
                  int first, void* stream);
int icl_sgd_chunk(void);
/* fused rank-R weight gradient + SGD update for the mlp2 weights: the gradient dy^T x is formed on the fly, never stored */
int icl_sgd_factored(int R, int N, int K, const float* dy, const float* x, float* p, float* m, const float* lr_ptr, float mu, float wd,
                     void* stream);

/* the same update on the tensor cores: factor pairs are packed (split bf16) into a workspace of icl_sgd_factored_workspace bytes with
 * one icl_sgd_factored_pack per (dy [rows][N], x [rows][K]) pair (dy scaled by `scale`, e.g. 1/world), then one
 * icl_sgd_factored_apply forms dy^T x in TMEM and streams p, m through TMA tiles (16 B of HBM traffic per parameter for any R) */
long long icl_sgd_factored_workspace(int R, int N, int K);
int icl_sgd_factored_pack(const float* dy, const float* x, int rows, int r0, int R_total, int N, int K, float scale, void* workspace, void* stream);
int icl_sgd_factored_apply(int R_total, int N, int K, const void* workspace, float* p, float* m, const float* lr_ptr, float mu, float wd,
                           int max_ctas, void* stream);

/* ---- sliding-window inference + Dice counts: test_3D_BraTS.py:110-135,175-187 (val_3D.py:43-97) ---- */
int icl_sw_accumulate(const float* logits, int K, int pw, int ph, int pd, float* score, float* cnt, int W, int H, int D, int xs, int ys, int zs,
                      void* stream);
int icl_sw_finalize(const float* score, const float* cnt, int K, long long S, long long* label, void* stream);
int icl_dice_counts(const long long* pred, const long long* gt, long long n, unsigned long long* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif
