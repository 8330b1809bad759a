/* satmvs_b200.h — C ABI of the B200-native SatMVS plane-sweep path.
 *
 * The reference (WHU-GPCV/SatMVS) is 100 % Python and has no FFI; its "operator API" for this
 * path is the set of Python call sites listed in SURVEY.md §8b.  Each entry point below names
 * the reference function (file:line under /root/reference) it replaces.  The Python host side
 * (satmvs_b200/*.py) binds these with ctypes and keeps the reference's signatures.
 *
 * Conventions
 *  - every pointer documented "device" is a CUDA device pointer valid on the *current* device;
 *    "host" pointers are plain CPU memory read before the call returns;
 *  - tensors are dense, row-major (NCHW / NCDHW) fp32 unless stated otherwise;
 *  - cameras are tiny (170 doubles / 16 doubles): they are passed as HOST pointers and travel to
 *    the kernel by value in the launch parameters (constant bank), so calls are stateless and
 *    re-entrant; one call handles one batch element's camera set — the `_batched` host loop lives
 *    in the Python wrapper;
 *  - the caller owns every buffer; nothing here allocates, frees or synchronises;
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *  - return value 0 = ok, non-zero = error, text via satmvs_last_error() (thread-local).
 */
#ifndef SATMVS_B200_H
#define SATMVS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SATMVS_ABI_VERSION 1
#define SATMVS_MAX_SRC_VIEWS 8
#define SATMVS_MAX_PEERS 8
#define SATMVS_RPC_LEN 170

enum { SATMVS_OK = 0, SATMVS_EINVAL = 1, SATMVS_ECUDA = 2 };

/* Device-side failures detected inside a kernel (tensor-core completion timeout = 1, producer timeout = 2, cluster flag
 * timeout = 3) are parked in pinned host memory instead of trapping the CUDA context: the next library call of the same host
 * thread fails with SATMVS_ECUDA, and this function returns and clears the code (call it after synchronising). */
int satmvs_async_error(void);
int satmvs_abi_version(void);
const char* satmvs_last_error(void);

/* Measurement aid (bench.py): between begin and end every kernel launch of the calling thread is
 * bracketed by CUDA events on its stream; end synchronises the device and returns, per kernel class,
 * the summed device time [ms] and the launch count.  Classes: 0 sweep, 1 batched convs (conv engine),
 * 2 GRU gate conv, 3 GRU output conv, 4 GRU pointwise, 5 RED decoder, 6 CostRegNet, 7 heads. */
#define SATMVS_PROFILE_CLASSES 8
int satmvs_profile_begin(void);
int satmvs_profile_end(float* ms_by_class, int* launches_by_class);
/* the same, plus per class the time during which at least one of its launches was running (launches of concurrent streams overlap) */
int satmvs_profile_end_ex(float* ms_by_class, int* launches_by_class, float* busy_ms_by_class);

/* ---- fused plane sweep: per-hypothesis geometry + bilinear gather + variance over views ----
 * Replaces networks/casred.py:26-53 (== networks/casmvs.py:30-59; per-plane form casred.py:191-212)
 * with rpc_warping (modules/warping.py:310-365) inlined for every source view.
 *   ref_fea   device [C,H,W]           reference-view features of ONE batch element
 *   src_feas  host array of n_src device pointers, each [C,H,W]
 *   ref_rpc   host double[170]; src_rpcs host double[n_src*170]     (layout data_io.py:78-92)
 *   depth     device; depth_per_pixel=0: [D] planes, =1: [D,H,W] per-pixel hypotheses
 *   out_var   device [C,D,H,W]:  var = Q/V - (S/V)^2,  V = n_src+1
 */
int satmvs_cost_volume_rpc_fwd(const float* ref_fea, const float* const* src_feas, int n_src,
                               const double* ref_rpc, const double* src_rpcs,
                               const float* depth, int depth_per_pixel,
                               int C, int D, int H, int W, float* out_var, void* stream);

/* Same with the pin-hole homography of homo_warping (modules/warping.py:6-44).
 *   ref_proj host double[16], src_projs host double[n_src*16]  (K·E, row-major 4x4) */
int satmvs_cost_volume_homo_fwd(const float* ref_fea, const float* const* src_feas, int n_src,
                                const double* ref_proj, const double* src_projs,
                                const float* depth, int depth_per_pixel,
                                int C, int D, int H, int W, float* out_var, void* stream);

/* ---- depth-sharded sweep (multi-GPU) ----
 * Same kernels, sweeping only D planes (hypotheses `depth` = the shard's [D] / [D,H,W]) and writing them
 * at plane offset d0 of output volumes that hold D_total planes ([C,D_total,H,W]).  With n_outs == 1
 * this fills one rank's slab of a shared layout; with n_outs == world size and `outs` = the peer-mapped
 * pointers of every rank's volume (CUDA IPC / symmetric memory), the store loop IS the all-gather:
 * each value is written once per peer over NVLink while the sweep is still computing, and the slab
 * never makes a second trip through HBM.  The caller provides the cross-rank barrier afterwards.
 * n_outs == -1: outs[0] is an NVLink-switch MULTICAST address of the symmetric volume (NVLS): every value leaves the SM
 * once as a multimem.st and the switch replicates it to all GPUs (egress 1x instead of (G-1)x).
 * workspace: optional caller-owned scratch of n_src*C*H*W*4 bytes (16-byte aligned).  When given and C is
 * a multiple of 4, the source features are re-packed to 4-channel words and the vectorised kernel runs
 * (same results); with NULL the scalar kernel runs.  satmvs_cost_volume_*_fwd == this with d0 = 0,
 * D_total = D, one output and no workspace. */
int satmvs_cost_volume_rpc_fwd_sharded(const float* ref_fea, const float* const* src_feas, int n_src,
                                       const double* ref_rpc, const double* src_rpcs,
                                       const float* depth, int depth_per_pixel,
                                       int C, int D, int H, int W, int d0, int D_total,
                                       float* const* outs, int n_outs,
                                       void* workspace, size_t workspace_bytes, void* stream);
int satmvs_cost_volume_homo_fwd_sharded(const float* ref_fea, const float* const* src_feas, int n_src,
                                        const double* ref_proj, const double* src_projs,
                                        const float* depth, int depth_per_pixel,
                                        int C, int D, int H, int W, int d0, int D_total,
                                        float* const* outs, int n_outs,
                                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- single-view warps with the reference operator's meaning ----
 * rpc_warping (modules/warping.py:310-365) / homo_warping (:6-44): out [C,D,H,W] warped volume. */
int satmvs_rpc_warp_fwd(const float* src_fea, const double* src_rpc, const double* ref_rpc,
                        const float* depth, int depth_per_pixel,
                        int C, int D, int H, int W, float* out, void* stream);
int satmvs_homo_warp_fwd(const float* src_fea, const double* src_proj, const double* ref_proj,
                         const float* depth, int depth_per_pixel,
                         int C, int D, int H, int W, float* out, void* stream);

/* Backward of the two warps w.r.t. the source features only (the reference builds the grid
 * under no_grad, warping.py:322-356): grad_src [C,H,W] += scatter of grad_out [C,D,H,W].
 * grad_src must be zero-initialised by the caller. */
int satmvs_rpc_warp_bwd(const float* grad_out, const double* src_rpc, const double* ref_rpc,
                        const float* depth, int depth_per_pixel,
                        int C, int D, int H, int W, float* grad_src, void* stream);
int satmvs_homo_warp_bwd(const float* grad_out, const double* src_proj, const double* ref_proj,
                         const float* depth, int depth_per_pixel,
                         int C, int D, int H, int W, float* grad_src, void* stream);

/* Backward of the fused variance volume: grad_var [C,D,H,W] -> grad_ref [C,H,W] and
 * grad_srcs[v] [C,H,W] (zero-initialised by the caller; host array of n_src device pointers). */
int satmvs_cost_volume_rpc_bwd(const float* grad_var, const float* ref_fea, const float* const* src_feas,
                               int n_src, const double* ref_rpc, const double* src_rpcs,
                               const float* depth, int depth_per_pixel, int C, int D, int H, int W,
                               float* grad_ref, float* const* grad_srcs, void* stream);
int satmvs_cost_volume_homo_bwd(const float* grad_var, const float* ref_fea, const float* const* src_feas,
                                int n_src, const double* ref_proj, const double* src_projs,
                                const float* depth, int depth_per_pixel, int C, int D, int H, int W,
                                float* grad_ref, float* const* grad_srcs, void* stream);

/* ---- RPC localisation / projection on flat fp64 point lists ----
 * RPC_Photo2Obj (modules/warping.py:255-307) == RPCModelParameter.RPC_PHOTO2OBJ
 * (tools/rpc_tensor.py:138-165); RPC_Obj2Photo (warping.py:218-252) == RPC_OBJ2PHOTO
 * (rpc_tensor.py:109-136).  rpc host double[170]; all point arrays device double[n]. */
int satmvs_rpc_localise(const double* rpc, const double* samp, const double* line, const double* hei,
                        int64_t n, double* lat, double* lon, void* stream);
int satmvs_rpc_project(const double* rpc, const double* lat, const double* lon, const double* hei,
                       int64_t n, double* samp, double* line, void* stream);

/* ---- bilinear gather of a height map at projected positions (geometric-consistency filter) ----
 * tools/rpc_filter.py:29-30: cv2.remap(depth_src, x, y, INTER_LINEAR, BORDER_CONSTANT, borderValue) with float32 maps:
 * coordinates rounded to 1/32 pixel, out-of-range taps replaced by `border` one by one.
 * src device float[Hs*Ws]; mapx / mapy / out device float[n]. */
int satmvs_remap_bilinear(const float* src, int Hs, int Ws, const float* mapx, const float* mapy, int64_t n,
                          float border, float* out, void* stream);

/* ---- soft-argmin heads ----
 * mode 0: RED train head (networks/casred.py:58-62): softmax over D, depth = sum p*d, conf = max p.
 * mode 1: CasMVS head (networks/casmvs.py:66-74): conf = sum of the 4 probabilities around the
 *         regressed plane index.  logits device [D,H,W]; depth as above; out_* device [H,W].
 * mode 2: depth_regression (modules/module.py:433-439): `logits` holds probabilities, depth = sum p*d;
 *         out_conf may be NULL. */
int satmvs_softargmin_fwd(const float* logits, const float* depth, int depth_per_pixel, int mode,
                          int D, int H, int W, float* out_depth, float* out_conf, void* stream);

/* Streaming fp64 head of the inference net (networks/casred.py:182-184, :218-236).
 * state device double[3*H*W] = (exp_sum, depth_acc, max_e), zero-initialised by the caller.
 * update: one plane, reg device [H,W], depth_plane device [H,W] (or [1] when depth_per_pixel=0).
 * finish: depth = depth_acc/(exp_sum+1e-10), conf = max_e/(exp_sum+1e-10) -> fp32 [H,W]. */
int satmvs_softargmin_stream_update(const float* reg, const float* depth_plane, int depth_per_pixel,
                                    int H, int W, double* state, void* stream);
/* The same update for K consecutive planes in one launch (reg device [K,H,W], depth device [K,H,W] or [K]): planes are
 * accumulated in order with the per-plane arithmetic of satmvs_softargmin_stream_update. */
int satmvs_softargmin_stream_update_planes(const float* reg, const float* depth, int depth_per_pixel, int K,
                                           int H, int W, double* state, void* stream);
/* Plane sweep without a regulariser (the exchange-light consumer of a depth-sharded sweep): folds the K planes of a
 * variance slab [C,K,H,W] into the same fp64 state with reg_k = scale * mean_c var[c,k] (scale < 0). */
int satmvs_softargmin_stream_update_volume(const float* var, const float* depth, int depth_per_pixel, float scale, int C, int K,
                                           int H, int W, double* state, void* stream);
int satmvs_softargmin_stream_finish(const double* state, int H, int W,
                                    float* out_depth, float* out_conf, void* stream);

/* ---- depth hypotheses + cascade resampling ----
 * get_depth_range_samples (modules/depth_range.py:4-42) fused with the bilinear up-sampling of the
 * previous stage's depth and the trilinear resize to the stage grid (networks/casred.py:132-145).
 *   prev_depth device [hp,wp] previous-stage depth, or NULL for the first stage, in which case
 *   depth_range device [n_range] gives the range (planes from [0] to [n_range-1] inclusive);
 *   out device [D,h,w] hypotheses on the stage grid (h = Himg/scale, w = Wimg/scale). */
int satmvs_depth_hypotheses(const float* prev_depth, int hp, int wp, const float* depth_range, int n_range,
                            int D, float interval, int Himg, int Wimg, int h, int w, float* out, void* stream);

/* F.interpolate(x, [ho, wo], mode="bilinear", align_corners=False) over N planes: the resize of the
 * hypotheses inside depth_regression (modules/module.py:437).  in device [N,hi,wi], out device [N,ho,wo]. */
int satmvs_resize_bilinear(const float* in, int N, int hi, int wi, int ho, int wo, float* out, void* stream);

/* ---- RED regulariser: 2-D conv-GRU UNet recurring over depth planes ----
 * RED_Regularization.forward (modules/module.py:614-649) with D planes, and, with D = 1 and explicit
 * states, slice_RED_Regularization.forward (modules/module.py:672-693).  base_channels is 8 (hidden
 * channels 8/16/32/64 are hard-coded in the reference, module.py:617-620).
 * All pointers device fp32.  Weight fields carry the reference's state_dict names:
 *   gate_w[l]  conv_gru{l+1}.gate_conv.weight   [2*ch, cx+ch, 3, 3]    gate_b[l]  .gate_conv.bias
 *   out_w[l]   conv_gru{l+1}.output_conv.weight [ch, cx+ch, 3, 3]      out_b[l]   .output_conv.bias
 *   rn_*, un_*, on_*  reset_gate_norm / update_gate_norm / output_norm  .weight / .bias   [ch]
 *   conv_w[i]  conv{i+1}.conv.weight;  upconv_w[i]  upconv{i+1}.conv.weight  ([Cin, Cout, 3, 3])
 *   upconv2d_w [8,1,3,3], upconv2d_b [1]
 */
typedef struct satmvs_red_weights {
  const float* gate_w[4]; const float* gate_b[4]; const float* out_w[4]; const float* out_b[4];
  const float* rn_w[4]; const float* rn_b[4]; const float* un_w[4]; const float* un_b[4];
  const float* on_w[4]; const float* on_b[4];
  const float* conv_w[3];
  const float* upconv_w[3];
  const float* upconv2d_w; const float* upconv2d_b;
} satmvs_red_weights;

/* bytes of caller-owned scratch satmvs_red_forward needs (0 if the shape is unsupported:
 * H and W must be multiples of 8). */
size_t satmvs_red_workspace_bytes(int C, int D, int H, int W);

/*   volume   [C,D,H,W] variance cost volume of one batch element (the regulariser negates it itself)
 *   state_in  host array of 4 device pointers [8,H,W] [16,H/2,W/2] [32,H/4,W/4] [64,H/8,W/8], or NULL /
 *             NULL entries for zero initial states;  state_out likewise, receives the states after plane D-1
 *   logits   [D,H,W] regularised cost (the reference's prob_volume before softmax) */
int satmvs_red_forward(const satmvs_red_weights* w, const float* volume, int C, int D, int H, int W,
                       const float* const* state_in, float* const* state_out, float* logits,
                       void* workspace, size_t workspace_bytes, void* stream);

/* The same with a caller-owned region for the packed tensor-core weights (satmvs_red_pack_bytes(C) bytes, 256-byte aligned):
 * the packs are written on the first call and reused while *pack_tag (host memory) still holds the signature the library
 * stored there; zero the tag whenever the weights change.  Saves ~11 small launches per forward. */
size_t satmvs_red_pack_bytes(int C);
int satmvs_red_forward_packed(const satmvs_red_weights* w, const float* volume, int C, int D, int H, int W,
                              const float* const* state_in, float* const* state_out, float* logits,
                              void* workspace, size_t workspace_bytes, void* pack, size_t pack_bytes,
                              unsigned long long* pack_tag, void* stream);

/* Which implementation of the depth recurrence the last satmvs_red_forward of this thread used: 3 = tensor-core cluster
 * kernel (csrc/red_tc.cuh) running concurrently with the batched convolutions that feed it (side stream, per-plane ready
 * counters; SATMVS_RED_NO_OVERLAP=1 disables), 2 = the same kernel after them, 1 = FFMA cluster kernel (csrc/red_cluster.cuh), 0 = per-plane kernel chain, -1 = none yet.
 * Diagnostic for tests and the bench line; SATMVS_RED_NO_TC=1 / SATMVS_RED_NO_CLUSTER=1 force the later ones. */
int satmvs_red_last_path(void);

/* ---- FeatureNet: the 2-D UNet in front of the plane sweep (modules/module.py:442-543, arch_mode "unet", num_stage 3) ----
 * Replaces `self.feature(img)` of networks/casred.py:116-119 / :288-290 for ALL views of a stack in one call.
 *   block[i]: conv weight + inference-mode BatchNorm folded to scale / shift, in the reference's order
 *     0 conv0.0 [b,3,3,3]   1 conv0.1 [b,b,3,3]      2 conv1.0 [2b,b,5,5] (stride 2)   3 conv1.1   4 conv1.2 [2b,2b,3,3]
 *     5 conv2.0 [4b,2b,5,5] (stride 2)   6 conv2.1   7 conv2.2 [4b,4b,3,3]
 *     8 deconv1.deconv [4b,2b,3,3] (ConvTranspose2d, stride 2)   9 deconv1.conv [2b,4b,3,3]
 *    10 deconv2.deconv [2b,b,3,3]                               11 deconv2.conv [b,2b,3,3]
 *   out_w[k]: the bare 1x1 heads out1 [4b,4b], out2 [2b,2b], out3 [b,b] (no bias).
 *   images [3,V,H,W] (channel-major, the V views as planes); out1 [V,4b,H/4,W/4], out2 [V,2b,H/2,W/2], out3 [V,b,H,W]
 *   (view-major: every view's outputs["stage1".."stage3"] is a contiguous [C,h,w] tensor).  H and W multiples of 4. */
typedef struct satmvs_conv_bn { const float* w; const float* scale; const float* shift; } satmvs_conv_bn;
typedef struct satmvs_featurenet_weights { satmvs_conv_bn block[12]; const float* out_w[3]; } satmvs_featurenet_weights;
size_t satmvs_featurenet_workspace_bytes(int base_channels, int V, int H, int W);
int satmvs_featurenet_forward(const satmvs_featurenet_weights* w, const float* images, int base_channels, int V, int H, int W,
                              float* out1, float* out2, float* out3, void* workspace, size_t workspace_bytes, void* stream);

/* ---- one convolution block of the regularisers ----
 * ConvReLU (modules/module.py:178-186): 3x3 per depth plane (NZ = 1); Conv3d (modules/module.py:324-366): 3x3x3 (NZ = 3);
 * padding 1, out = relu?( conv(in, w) * acc_scale * scale[co] + shift[co] ), scale / shift may be NULL.
 *   in [Cin,D,H,W], w [Cout,Cin,(3,)3,3], out [Cout,Do,Ho,Wo]; stride 1 or 2 (NZ = 1 halves H, W; NZ = 3 halves D, H, W)
 *   engine 0 automatic, 1 tcgen05 tensor cores (3xTF32 split; SATMVS_EINVAL when the shape does not fit), 2 fp32 FFMA kernels,
 *          3 tcgen05 with the nine in-plane taps in the N dimension (Cout <= 8, stride 1)
 *   workspace: satmvs_conv_workspace_bytes(Cin, Cout, NZ) bytes of device scratch (packed weights), may be NULL for engine 2 */
size_t satmvs_conv_workspace_bytes(int Cin, int Cout, int NZ);
int satmvs_conv_forward(const float* in, int Cin, int D, int H, int W, const float* w, const float* scale, const float* shift,
                        int Cout, int NZ, int stride, int relu, float acc_scale, float* out, int engine,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- CostRegNet: 3-D conv UNet regulariser (CasMVSNet / UCS-Net) ----
 * CostRegNet.forward (modules/module.py:546-577), inference-mode BatchNorm.
 *   conv_w[0..6]  conv0..conv6 .conv.weight [Cout,Cin,3,3,3];  conv_w[7..9]  conv7, conv9, conv11
 *                 .conv.weight (ConvTranspose3d layout [Cin,Cout,3,3,3])
 *   bn_scale[i] = bn.weight / sqrt(bn.running_var + eps),  bn_shift[i] = bn.bias - bn.running_mean * bn_scale[i]
 *   prob_w        prob.weight [1,base,3,3,3]
 *   x [Cin,D,H,W] -> out [D,H,W] (the reference's [1,D,H,W]); D, H, W multiples of 8. */
typedef struct satmvs_costreg_weights {
  const float* conv_w[10];
  const float* bn_scale[10];
  const float* bn_shift[10];
  const float* prob_w;
} satmvs_costreg_weights;

size_t satmvs_costreg_workspace_bytes(int base_channels, int D, int H, int W);
int satmvs_costreg_forward(const satmvs_costreg_weights* w, const float* x, int Cin, int base_channels,
                           int D, int H, int W, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- training form of the Conv3d / Deconv3d blocks (modules/module.py:324-410; train.py:267-287 runs model.train() and
 * loss.backward()) ----
 * satmvs_conv3d_raw: out = conv(in) without BatchNorm / activation, weight read as w[co * w_co + ci * w_ci + tap];
 *   NZ 3: 3x3x3 on [C,D,H,W];  NZ 1: 3x3 per plane.  mode 0: stride 1, padding 1;  1: stride 2, padding 1 (even sizes);
 *   2: stride 1 with mirrored taps (with w_co / w_ci swapped this is the data gradient of mode 0);
 *   3: transposed, stride 2, padding 1, output_padding 1 (nn.ConvTranspose3d; reading a Conv3d weight [Cout,Cin,..] with
 *      w_ci = Cin*taps, w_co = taps it is the data gradient of mode 1, and mode 1 reading a ConvTranspose3d weight is the data
 *      gradient of mode 3).
 * satmvs_conv3d_wgrad: dw[co * dw_co + ci * dw_ci + tap] (+)= sum_o dy[co, o] x[ci, stride * o + k - 1]  (padding 1);
 *   for a transposed block call it with x = the output gradient (large tensor) and dy = the block input (small tensor), stride 2.
 * satmvs_bn_train_fwd / _bwd: BatchNorm3d on batch statistics (+ ReLU, + skip tensor added after the ReLU, module.py:573-575)
 *   on [B,C,n] tensors; mean / var (biased) are outputs of _fwd and inputs of _bwd; acc = 2 C doubles of scratch;
 *   dz2 (may be null) is a second gradient added to dz (a block output that also feeds a skip connection).
 * satmvs_softargmin_bwd: gradient of depth = sum_d softmax(logits)_d * depth_d to the logits [D,H,W] (casmvs.py:66-68). */
int satmvs_conv3d_raw(const float* in, int Cin, int Di, int Hi, int Wi, const float* w, long long w_co, long long w_ci,
                      int NZ, int mode, float* out, int Cout, void* stream);
size_t satmvs_conv3d_wgrad_workspace_bytes(int Cin, int Cout, int Di, int Hi, int Wi, int NZ, int stride);
int satmvs_conv3d_wgrad(const float* x, int Cin, int Di, int Hi, int Wi, const float* dy, int Cout, int NZ, int stride,
                        float* dw, long long dw_co, long long dw_ci, int accumulate, void* workspace, size_t workspace_bytes,
                        void* stream);
/* Per-plane K x K (K = 1, 3, 5; padding K/2) forms of the two primitives above on [C][N][H][W] tensors (FeatureNet's layers,
 * modules/module.py:442-543): same modes; satmvs_conv2d_wgrad: dw[co, ci, ky * K + kx]. */
int satmvs_conv2d_raw(const float* in, int Cin, int N, int Hi, int Wi, const float* w, long long w_co, long long w_ci, int K, int mode,
                      float* out, int Cout, void* stream);
size_t satmvs_conv2d_wgrad_workspace_bytes(int Cin, int Cout, int N, int Hi, int Wi, int K, int stride);
int satmvs_conv2d_wgrad(const float* x, int Cin, int N, int Hi, int Wi, const float* dy, int Cout, int K, int stride,
                        float* dw, long long dw_co, long long dw_ci, int accumulate, void* workspace, size_t workspace_bytes,
                        void* stream);
int satmvs_bn_train_fwd(const float* y, int B, int C, long long n, const float* gamma, const float* beta, float eps, int relu,
                        const float* post_add, float* z, float* mean, float* var, double* acc, void* stream);
int satmvs_bn_train_bwd(const float* dz, const float* dz2, const float* y, int B, int C, long long n, const float* gamma, const float* beta,
                        const float* mean, const float* var, float eps, int relu, float* dy, float* dgamma, float* dbeta,
                        double* acc, void* stream);
int satmvs_softargmin_bwd(const float* logits, const float* depth_values, int per_pixel, int D, int H, int W,
                          const float* grad_depth, float* grad_logits, void* stream);

/* ---- backward of the RED regulariser (RED_Regularization.forward, modules/module.py:614-649; train.py:284 loss.backward()) ----
 * The forward (satmvs_red_forward) keeps, in its workspace, the hidden-state history S_l [ch_l][D+1][h_l][w_l] (slot 0 = initial
 * state), the encoder outputs E_i [ch][D][h][w] and the decoder tensors U_l [ch_l][D+1][h_l][w_l]; satmvs_red_workspace_layout
 * returns their byte offsets: offsets[0..3] = S_0..S_3, [4..6] = E_1..E_3, [7..9] = U_0..U_2.
 * The host side (satmvs_b200/training.py) recomputes the gate / output pre-activations batched over planes, walks the planes
 * backwards with satmvs_red_recurrence_bwd and finishes
 * with batched data / weight gradients.  Tensors are [C][D][px] (channel stride D*px unless a stride argument says otherwise). */
int satmvs_red_workspace_layout(int C, int D, int H, int W, size_t* offsets10);
/* out = act(GroupNorm(1, ch)(pre + bias)) per plane and group (groups 2: reset | update halves; act 0 sigmoid, 1 tanh);
 * stats [D][groups][2] = (mean, rstd) out; acc = D*groups*2 doubles of scratch. */
int satmvs_gn_act_fwd(const float* pre, const float* bias, const float* gamma, const float* beta, int ch, int groups, int D, int px,
                      float eps, int act, float* out, float* stats, double* acc, void* stream);
/* out = scale * (a [+ b]) [* mul], zeroed where (m1 [- m2]) <= 0 when m1 is given; every operand with its own channel stride. */
int satmvs_elementwise(const float* a, long long a_cs, const float* b, long long b_cs, const float* mul, long long mul_cs,
                       const float* m1, long long m1_cs, const float* m2, long long m2_cs, float scale, float* out, long long out_cs,
                       int C, long long n_per_c, void* stream);
int satmvs_channel_sum(const float* t, int C, long long n_per_c, float* out, double* acc, void* stream);
int satmvs_gn_param_grad(const float* dout, const float* pre, const float* bias, const float* stats, int ch, int groups, int D, int px,
                         float* dgamma, float* dbeta, double* acc, void* stream);
/* The sequential part: for each level (independent recurrences, one stream each between a fork from / join into `stream`) walk
 * planes D-1 .. 0.  S: state history [ch][D+1][px]; ru [2ch][D][px] (reset | update gates), y, opre [ch][D][px], gpre [2ch][D][px]
 * (conv results WITHOUT bias); ostat [D][2], gstat [D][2][2] (mean, rstd); dec [ch][D][px] = gradient reaching h'(d) from the
 * decoder; wo_h / wg_h: the output / gate filters offset to their hidden input channels, w_ci = (cx + ch) * 9 their stride
 * between output channels.  Outputs [.][D][px]: dyn, dgn (gradients at the GroupNorm outputs, for the affine gradients), dO, dG
 * (gradients at the conv outputs).  scratch: 12 * D + 4 + 14 * ch * px floats. */
typedef struct satmvs_gru_bwd_level {
  const float* S; const float* ru; const float* y; const float* opre; const float* gpre;
  const float* ob; const float* gb; const float* on_w; const float* rn_w; const float* un_w;
  const float* ostat; const float* gstat; const float* dec;
  const float* wo_h; const float* wg_h;
  long long w_ci;
  float* dyn; float* dgn; float* dO; float* dG; float* scratch;
  int ch, h, w, D;
} satmvs_gru_bwd_level;
int satmvs_red_recurrence_bwd(const satmvs_gru_bwd_level* levels, int nlevels, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SATMVS_B200_H */
