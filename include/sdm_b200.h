/* sdm_b200 -- C ABI of the B200 (sm_100a) diffusion hot path.
 *
 * The reference (Vinmwaura/Simple-Diffusion-Model) has no FFI layer: its hot path is PyTorch eager
 * (models/custom_layers.py, models/U_Net.py, degraders.py, diffusion_sampling_algorithms.py).  Each entry
 * point below names the reference call site whose arithmetic it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, nonzero on error; b2_last_error() returns a thread-local message;
 *   - all pointers are DEVICE pointers borrowed for the duration of the call (the caller owns all memory);
 *   - every launch is asynchronous on `stream` (a cudaStream_t passed as void*), allocation-free,
 *     host-sync-free and CUDA-graph capturable;
 *   - activations are NHWC; `ld*` is the per-pixel channel stride in ELEMENTS, so a tensor may be a channel
 *     slice of a wider buffer (zero-copy concat, reference models/U_Net.py:168);
 *   - dtype: 0 = bf16 storage / kind::f16 tensor cores, 1 = fp32 storage / kind::tf32 tensor cores
 *     ("parity mode"); accumulation is always fp32 (TMEM).
 */
#ifndef SDM_B200_H
#define SDM_B200_H
#ifdef __cplusplus
extern "C" {
#endif

const char* b2_last_error(void);
int b2_version(void);

/* Kernel-selection switches for A/B measurements and tests (defaults from SDM_B200_HALO / SDM_B200_SWAP_AB): "halo" = halo-tile
 * 3x3 convolutions, "swap_ab" = swapped-operand convolutions for <= 128 output channels (results are the same up to the
 * order of fp32 accumulation); "sm_limit" = number of SMs the persistent grids occupy (0 = all; SDM_B200_SM_LIMIT). */
int b2_set_option(const char* name, int value);
/* Zero-fills `bytes` of device memory on `stream` (the flat gradient buffer before a backward pass; the reference's
 * optimizer.zero_grad(), train_diffusion.py:317). */
int b2_zero(void* ptr, long long bytes, void* stream);
/* Deterministic mode (also SDM_B200_DETERMINISTIC=1): GroupNorm statistics from a fixed-order pass instead of the conv
 * epilogue's fp32 atomics and no split-K on the forward kernel, so an image's result is bitwise independent of batch size,
 * sharding and timing (sharded sampling == unsharded sampling); the weight-gradient kernel's split-K becomes
 * ordered (per-split partial tiles summed in split order: bitwise repeatable) instead of fp32 atomics.  Returns 0. */
int b2_set_deterministic(int on);
/* Optional split-K workspace for the tensor-core kernels: a device buffer (256-byte aligned, >= 2 MiB, ZEROED: each half
 * starts with 4 KiB of per-tile arrival counters) owned by the caller and registered for the CURRENT device (one per device).
 * The first half serves the NT kernel (forward / data gradients), the second the TN kernel (ordered weight-gradient split-K),
 * so the two may run on different streams; kernels sharing a half must be stream-ordered.  Layers whose output tiles cannot
 * fill the 148 SMs then split their K loop; counters are handed back zeroed.  Passing NULL disables split-K on that device. */
int b2_set_workspace(void* ws, long long bytes);

/* ---- dense contractions (tcgen05 implicit GEMM) ------------------------------------------------------- */

/* mode 0: Conv2d 3x3 stride 1 pad 1            (custom_layers.py:224-228), x = [N][H][W][Cin]
 * mode 1: Conv2d 3x3 stride 2 pad 1            (custom_layers.py:196-201), x = parity planes
 *         [2][2][N][H][W][Cin] made by b2_space_to_depth2, (H, W) = OUTPUT size
 * mode 2: ConvTranspose2d 4x4 stride 2 pad 1   (custom_layers.py:174-179), (H, W) = INPUT size, y is 2H x 2W
 * mode 3: data gradient of mode 1: x = dz [N][H][W][Cin=fwd Cout], y = dx [N][2H][2W][Cout=fwd Cin], weights kind 5
 * mode 4: data gradient of mode 2: x = parity planes of dz [2][2][N][H][W][fwd Cout], y = dx [N][H][W], weights kind 6
 *         (data gradient of mode 0 is mode 0 itself with weights of kind 1)
 * mode 5: data gradient of mode 0 straight from the FORWARD weights (bf16 only): x = dz [N][H][W][Cin = fwd Cout],
 *         wpacked = the forward kernel layout [fwd Cout][9][fwd Cin] (kind 0 / the optimiser's bf16 copy of a channels-last
 *         stored weight), Cout = fwd Cin; both channel counts multiples of 64.  The weights are consumed MN-major with the
 *         taps mirrored, so no transposed copy (kind 1) is made (autograd of custom_layers.py:224).
 * wpacked: weights in kernel layout from b2_pack_conv_weight. act: 0 none, 1 Swish (custom_layers.py:18-20), 2 tanh,
 * 3 = store the pre-activation but accumulate the GroupNorm statistics of Swish(value) (training forward).
 * residual (optional, mode 0/1): added after the activation. gn_stats (optional): [N][gn_groups][2] fp32,
 * must be zeroed by the caller; receives per-(image, group) sum and sum of squares of the written values.
 * out_mode 0: y is NHWC in `dtype`; out_mode 1 (mode 0 only): y is fp32 NCHW [N][Cout][H][W] -- the final
 * layer writes the network output directly (models/U_Net.py:172); act 2 = tanh (image_recon, U_Net.py:126). */
int b2_conv2d_nhwc(int mode, const void* x, int N, int H, int W, int Cin, long long ldx, const void* wpacked,
                   const float* bias, int Cout, void* y, long long ldy, int act, const void* residual,
                   long long ldr, float* gn_stats, int gn_groups, int out_mode, int dtype, void* stream);

/* Network-edge 3x3 convolutions on CUDA cores (too narrow for a 128 x 64 tensor-core tile: models/U_Net.py:55-66, :113-130).
 * first: x fp32 NCHW [N][Cin][H][W], Cin 3 or 6 (fuses the NCHW->NHWC edge); w_kc = weight as [Cin*9][Cout] fp32;
 *        y NHWC `dtype`, act 0 none / 1 Swish.
 * last : x NHWC `dtype`; w_tc4 = weight as [9][Cin][4] fp32 (Cout <= 4, unused slots zero); y fp32 NCHW [N][Cout][H][W],
 *        act 0 none / 2 tanh (image_recon, U_Net.py:126). */
int b2_conv3x3_first(const float* x, const float* w_kc, const float* bias, void* y, long long ldy, int N, int Cin, int H, int W,
                     int Cout, int act, int dtype, void* stream);
int b2_conv3x3_last(const void* x, long long ldx, const float* w_tc4, const float* bias, float* y, int N, int H, int W, int Cin,
                    int Cout, int act, int dtype, void* stream);

/* C = alpha * A . B^T (+bias) (act) (+residual); A [M][K], B [Ncols][K] (nn.Linear weight layout,
 * custom_layers.py:116,119; q.k^T custom_layers.py:144).  batch1/batch2 > 1: batched with element strides
 * *_s1 / *_s2 for A, B and C. */
int b2_gemm_nt(const void* A, long long lda, long long a_s1, long long a_s2, const void* B, long long ldb,
               long long b_s1, long long b_s2, void* C, long long ldc, long long c_s1, long long c_s2, int M,
               int Ncols, int K, int batch1, int batch2, const float* bias, float alpha, int act,
               const void* residual, long long ldr, int out_fp32, int dtype, void* stream);
/* b2_gemm_nt with B given UN-transposed, [batch2][batch1][K][Ncols] (row stride ldb): C = alpha * A . B, bf16, Ncols % 64 == 0.
 * The attention backward products dV = P^T dO and dK = dS^T Q (autograd of custom_layers.py:144-150) read dO / Q in place --
 * the tensor-core kernel consumes B MN-major -- instead of through transposed copies; likewise the Linear data gradients
 * dX = dY W (+ residual, unbatched) read the forward weight [out][in] in place. */
int b2_gemm_nt_bmn(const void* A, long long lda, long long a_s1, long long a_s2, const void* B, long long ldb, long long b_s1,
                   long long b_s2, void* C, long long ldc, long long c_s1, long long c_s2, int M, int Ncols, int K, int batch1,
                   int batch2, float alpha, const void* residual, long long ldr, int dtype, void* stream);

/* Fused attention scores (custom_layers.py:144-147): P^T[n][h][j][i] = softmax over the QUERY index i of
 * scale * q_i . k_j, computed as S^T = K Q^T on the tensor cores with the softmax in the epilogue (one thread owns one
 * key row in TMEM), so the fp32 score matrix is never written.  k, q: [P][d] slices of the packed qkv tensor (row stride
 * ld, head stride sh, image stride sn, in elements); pt: [N][heads][P][ldp], ldp a multiple of 8.  work: 2*N*heads*P*
 * ceil(P/256) floats, used only when P > 256 (per-tile statistics + a normalising fix-up pass).  P.V then is
 * b2_gemm_tn(pt, v): no transpose of V is needed. */
int b2_attn_scores_softmax(const void* k, const void* q, long long ld, long long sh, long long sn, void* pt, long long ldp,
                           int P, int d, int heads, int N, float scale, float* work, int dtype, void* stream);
/* Backward of the above in one tensor-core pass: dS^T = scale * P^T .* (V dO^T - dot[key]), dot from b2_rowdot(v, dv). */
int b2_attn_scores_bwd(const void* v, long long ld, long long sh, long long sn, const void* d_o, long long ld_do,
                       long long do_sh, long long do_sn, const void* pt, const float* dot, void* dst, long long ldp, int P,
                       int d, int heads, int N, float scale, int dtype, void* stream);
/* out[r][h] = sum_{c<d} a[r][h*head_stride + c] * b[r][h*head_stride + c]  (fp32). */
int b2_rowdot(const void* a, long long lda, const void* b, long long ldb, long long head_stride, float* out, long long rows,
              int heads, int d, int dtype, void* stream);

/* ---- gradients of the dense contractions (tcgen05 "TN" GEMM: contraction over pixel / sequence rows) -------- */

/* Weight gradient of b2_conv2d_nhwc's three modes (autograd's convolution_backward in the reference), written (plain
 * stores when one work item owns a tile, fp32 atomics under split-K) into a ZEROED buffer in kernel layout: mode 0/1 [Cout][9][Cin], mode 2 [4][Cout][4][Cin].
 * x: forward input (mode 1: its parity planes); dz: gradient w.r.t. the conv pre-activation output
 * (mode 2: [N][2H][2W][Cout]); (H, W) as in b2_conv2d_nhwc. */
int b2_conv2d_wgrad(int mode, const void* x, int N, int H, int W, int Cin, long long ldx, const void* dz, int Cout,
                    long long lddz, float* grad_packed, int dtype, void* stream);
/* n_jobs weight gradients in as few launches as possible: desc = n_jobs rows of 11 values {mode, x, N, H, W, Cin, ldx, dz, Cout,
 * lddz, grad_packed} (pointers as integers; meaning as in b2_conv2d_wgrad; host memory, consumed before the call returns).  Weight
 * gradients feed nothing but the optimiser (the reference's autograd computes them wherever it likes, train_diffusion.py:358), so
 * the host may defer a module's worth and run them as ONE persistent kernel whose work items carry a job index -- at small
 * batches a step is bound by the number of ~16 us dependent launches, not by their arithmetic.  Equals n_jobs calls of
 * b2_conv2d_wgrad up to the order of fp32 atomic adds; layers the grouped kernel cannot take are launched one by one. */
int b2_conv2d_wgrad_batch(int n_jobs, const long long* desc, int dtype, void* stream);
/* C (+)= alpha * A^T . B with A [K][M], B [K][Ncols] (rows = contraction index), optionally batched.
 * out_mode 0: fp32 result added into a ZEROED C (Linear weight grads; atomics only under split-K);
 * out_mode 1: store in `dtype` (attention products). */
int b2_gemm_tn(const void* A, long long lda, long long a_s1, long long a_s2, const void* B, long long ldb,
               long long b_s1, long long b_s2, void* C, long long ldc, long long c_s1, long long c_s2, int M,
               int Ncols, int K, int batch1, int batch2, float alpha, int out_mode, int dtype, void* stream);

/* ---- memory-bound forward kernels ----------------------------------------------------------------------- */

/* fp32 NCHW image -> NHWC `dtype` with channels zero-padded to Cpad (network input edge, U_Net.py:155). */
int b2_nchw_to_nhwc_pad(const float* x, void* y, int N, int C, int H, int W, int Cpad, int dtype, void* stream);
/* NHWC `dtype` (row stride ldx) -> fp32 NCHW (module-boundary edge of the standalone blocks). */
int b2_nhwc_to_nchw(const void* x, long long ldx, float* y, int N, int C, int H, int W, int dtype, void* stream);
/* planes[pr][pc][n][i][j][:] = x[n][2i+pr][2j+pc][:]; feeds b2_conv2d_nhwc mode 1 (custom_layers.py:196). */
int b2_space_to_depth2(const void* x, long long ldx, void* planes, int N, int H, int W, int C, int dtype, void* stream);
/* fp32 master weights -> kernel layout; k_pad zero-pads the contraction side to the 128-byte K block.
 * kind 0 Conv2d [Cout][Cin][3][3] -> [Cout][9][k_pad>=Cin]; 1 same -> [Cin][9 flipped][k_pad>=Cout] (dgrad, s1);
 * 2 ConvTranspose2d [Cin][Cout][4][4] -> [4 parities][Cout][4 taps][Cin]; 3 Linear [Cout][Cin] -> [Cout][k_pad>=Cin];
 * 4 Linear -> transposed [Cin][k_pad>=Cout] (dgrad); 5 Conv2d -> [4][Cin][4 zero-padded taps][Cout] (dgrad, s2);
 * 6 ConvTranspose2d -> [Cin][16][Cout] (dgrad).  dtype 1 rounds to TF32 (nearest) so the MMA's truncation is exact. */
int b2_pack_weight(int kind, const float* w, void* out, int Cout, int Cin, int k_pad, int dtype, void* stream);
/* Batched forms: every stale kernel-layout weight of a training step in ONE launch each.  jobs_dev: device int64 records,
 * pack: {w, out, kind, Cout, Cin, k_pad, first element, element count} (first elements ascending, total = sum of counts);
 * transpose: {in, out, Cout, Cin, first tile, tile count} with tiles = (Cin/64)*(Cout/64)*9 per weight. */
int b2_pack_weight_multi(const long long* jobs_dev, int njobs, long long total, int dtype, void* stream);
int b2_transpose_weight_cl_multi(const long long* jobs_dev, int njobs, long long total_tiles, void* stream);
/* bf16 [Cout][9][Cin] (a 3x3 weight stored channels-last, the layout of kind 0) -> [Cin][9 flipped][Cout] (kind 1):
 * data-gradient weights derived from the optimiser's bf16 copy; Cout, Cin multiples of 64. */
int b2_transpose_weight_cl(const void* w_cl, void* out, int Cout, int Cin, void* stream);
/* bf16 Linear weight [rows][cols] -> [cols][rows] (kind 4, the data-gradient layout of custom_layers.py:116,119) from the
 * optimiser's bf16 copy; rows, cols multiples of 64. */
int b2_transpose_linear_weight(const void* w, void* out, int rows, int cols, void* stream);
/* bf16 ConvTranspose2d weight [Cin][Cout][4][4] (custom_layers.py:174) -> fwd (kind 2 layout) and/or dgrad (kind 6 layout),
 * either may be NULL; Cin, Cout multiples of 32. */
int b2_pack_convt_bf16(const void* w, void* fwd, void* dgrad, int Cin, int Cout, void* stream);
/* GroupNorm statistics as a separate pass: stats[n][g] += (sum, sum of squares) of y (pre_swish: of Swish(y)); used for
 * group widths the conv epilogue does not fuse and by the standalone AdaGN module (custom_layers.py:35-45). */
int b2_gn_stats(const void* y, long long ldy, float* stats, int N, int HW, int C, int groups, int pre_swish, int dtype, void* stream);
/* out = s*(gamma*(y-mean)*rstd+beta) + s (+residual): GroupNorm x AdaGN (custom_layers.py:35-45) fused with the
 * ResidualBlock add (custom_layers.py:282-287).  stats from b2_conv2d_nhwc; s = y_scale(emb) [B][C] with row
 * stride s_bstride (0 broadcasts one embedding over the batch, as the samplers do).  pre_swish: y holds the conv
 * pre-activation (training forward keeps it for the backward pass) and Swish is applied on load. */
int b2_adagn_apply(const void* y, long long ldy, const float* stats, const float* gamma, const float* beta,
                   const float* s, long long s_bstride, const void* residual, long long ldr, void* out, long long ldo,
                   int N, int HW, int C, int groups, float eps, int pre_swish, int dtype, void* stream);
/* P[b][i][j] = softmax over the QUERY index i of S[b][i][j] (custom_layers.py:147); S fp32, P `dtype`, row stride ldp. */
int b2_softmax_query_axis(const float* S, void* P, int B, int Pq, int Pk, long long ldp, int dtype, void* stream);
/* out[b1][b2][c][r] = in[b1][b2][r][c] (V^T for P.V, custom_layers.py:150). */
int b2_transpose_batched(const void* in, long long ld_in, long long in_s1, long long in_s2, void* out, long long ld_out,
                         long long out_s1, long long out_s2, int R, int Ccols, int B1, int B2, int dtype, void* stream);
/* [sin(t f_k), cos(t f_k)] (custom_layers.py:84-90); t int64 [B]. */
int b2_sinusoid_embedding(const long long* t, float* out, int B, int dim, void* stream);
/* fp32 CUDA-core GEMM for the tiny embedding / AdaGN-scale linears (custom_layers.py:30,60-77):
 * C (+)= op(A).op(B) (+bias)(Swish). ta: A stored [K][M]; tb 0: B stored [N][K], tb 1: B stored [K][N]. */
int b2_small_gemm(const float* A, long long lda, int ta, const float* B, long long ldb, int tb, float* C, long long ldc,
                  int M, int N, int K, const float* bias, int act, int accumulate, void* stream);

/* ---- memory-bound backward kernels (the reference gets these from autograd) ------------------------------- */

/* Backward of Conv -> Swish -> AdaGN (custom_layers.py:240-245, :35-45) from dout to the conv pre-activation:
 * dz = rstd*(s*gamma*dout - m1 - xh*m2) * swish'(z).  Also accumulates ds (gradient of the AdaGN scale vector, row
 * stride ds_bstride, 0 = embedding broadcast over the batch), dgamma, dbeta, dbias (+=, fp32).  z: pre-activation
 * saved by the forward; stats: the forward's (sum, sumsq); work: 2*N*C floats of scratch, ZEROED by the caller. */
int b2_adagn_bwd(const void* dout, long long ldd, const void* z, long long ldz, const float* stats, const float* gamma,
                 const float* beta, const float* s, long long s_bstride, float* work, float* ds, long long ds_bstride,
                 float* dgamma, float* dbeta, void* dz, long long lddz, float* dbias, int N, int HW, int C, int groups,
                 float eps, int dtype, void* stream);
/* mode 0: out = swish(z); mode 1: out = a * swish'(z), dbias += column sums; mode 2: dbias += column sums of a. */
int b2_act(int mode, const void* a, long long lda, const void* z, long long ldz, void* out, long long ldo, float* dbias,
           long long rows, int C, int dtype, void* stream);
/* fp32 helpers for the embedding MLPs: mode 0 swish(z), 1 a*swish'(z), 2 a*(1 - z^2) (tanh backward, z = tanh output). */
int b2_f32_act(int mode, const float* a, const float* z, float* out, long long n, void* stream);
/* dS = scale * P * (dP - sum_i P dP): backward of the query-axis softmax (custom_layers.py:147). */
int b2_softmax_query_axis_bwd(const void* P, const float* dP, void* dS, int B, int Pq, int Pk, long long ldp, float scale,
                              int dtype, void* stream);
/* out = a + b on NHWC views (merges the two consumers of a skip tensor, models/U_Net.py:160,168, in the backward pass). */
int b2_add(const void* a, long long lda, const void* b, long long ldb, void* out, long long ldo, long long rows, int C,
           int dtype, void* stream);
/* Kernel-layout fp32 weight gradient -> parameter layout (kind 0: Conv2d, kind 2: ConvTranspose2d). */
int b2_unpack_weight_grad(int kind, const float* packed, float* grad, int Cout, int Cin, int Cin_pad, int accumulate,
                          void* stream);

/* ---- diffusion process (fp32 NCHW tensors) -------------------------------------------------------------- */

/* Standard normals from Philox4x32-10 keyed on (seed, offset, global element index): sharded draws == unsharded. */
int b2_philox_normal(float* out, long long n, unsigned long long seed, unsigned long long offset, long long first_elem,
                     void* stream);
/* q(x_t|x_0) = sqrt(abar_t) img + sqrt(1-abar_t) eps (degraders.py:51-59 table gather when abar_table != NULL,
 * degraders.py:70-82,96-104 cosine closed form otherwise); steps int64, 1 or N entries. */
int b2_qsample(const float* img, const float* eps, float* out, const long long* steps, int steps_count,
               const float* abar_table, int max_step, int N, long long per_image, void* stream);
/* Same, with eps ~ N(0, I) drawn IN the kernel (the reference's `torch.randn_like`, train_diffusion.py:310 /
 * degraders.py:53-54) from Philox keyed on (seed, offset, first_elem + element index); offset_dev != NULL overrides
 * `offset` with (u64)*offset_dev (device-resident step counter: CUDA-graph replays draw fresh noise); eps_out optional.
 * A timestep outside [0, max_step] yields NaN for that image in both variants (the reference's gather raises). */
int b2_qsample_philox(const float* img, float* out, float* eps_out, const long long* steps, int steps_count,
                      const float* abar_table, int max_step, int N, long long per_image, unsigned long long seed,
                      unsigned long long offset, const float* offset_dev, long long first_elem, void* stream);
/* diffusion_sampling_algorithms.py:107-136: x0 = c_scale*(x - c_s*e); x' = c_an*x0 + c_dir*e + sigma*noise. */
int b2_ddim_step(const float* x_t, const float* eps_hat, const float* noise, float* x_out, float* x0_out, long long n,
                 float c_scale, float c_s, float c_an, float c_dir, float sigma, int last, void* stream);
/* diffusion_sampling_algorithms.py:42-55: x' = scale1*(x - scale2*e) + sigma*z (z given | Philox | none). */
int b2_ddpm_step(const float* x_t, const float* eps_hat, const float* z, float* out, long long n, float scale1,
                 float scale2, float sigma, int use_philox, unsigned long long seed, unsigned long long offset,
                 long long first_elem, void* stream);
/* diffusion_sampling_algorithms.py:193-208: x' = x - (a_t x0 + b_t noise) + (a_n x0 + b_n noise). */
int b2_cold_step(const float* x_t, const float* x0_hat, const float* noise, float* out, long long n, float a_t,
                 float b_t, float a_n, float b_n, void* stream);
/* F.interpolate(mode="area") on fp32 NCHW planes (train_SR_diffusion.py:321-328, generate_sr_images_diffusion.py:170-173):
 * adaptive average pooling, out[o] = mean in[floor(o*I/O) .. ceil((o+1)*I/O)); planes = N*C. */
int b2_area_resample(const float* x, float* y, long long planes, int H, int W, int OH, int OW, void* stream);
/* loss = mean((pred-target)^2) (train_diffusion.py:350), grad (optional) = 2 (pred-target)/n * grad_scale. */
int b2_mse_loss_grad(const float* pred, const float* target, float* grad, float* loss, long long n, float grad_scale,
                     void* stream);

/* Same with target = the eps b2_qsample_philox drew for (seed, offset | *offset_dev, first_elem): re-generated in the kernel,
 * never stored (eps-prediction target, train_diffusion.py:336-350). */
int b2_mse_loss_grad_philox(const float* pred, float* grad, float* loss, long long n, float grad_scale,
                            unsigned long long seed, unsigned long long offset, const float* offset_dev, long long first_elem,
                            void* stream);

/* ---- image input / output edges (uint8 <-> fp32, on the device) ---------------------------------------------- */

/* uint8 [N][H][W][C] (cv2 BGR bytes) -> fp32 [N][C][H][W] = (x - 127.5) / 127.5, computed in double and rounded once like
 * custom_dataset/img_dataset.py:26-35 and generate_sr_images_diffusion.py:117-126; flip_flags (optional, N bytes): nonzero =
 * horizontal flip of that image (torchvision RandomHorizontalFlip per image, train_diffusion.py:312-314). */
int b2_u8_to_image(const void* src_u8_nhwc, float* dst_nchw, const void* flip_flags, int N, int H, int W, int C, void* stream);
/* Per-image horizontal flip of an fp32 [N][C][H][W] batch (train_diffusion.py:312-314). */
int b2_flip_images(const float* x_nchw, float* out_nchw, const void* flip_flags, int N, int C, int H, int W, void* stream);
/* fp32 [N][C][H][W] -> uint8 [N][H][W][C]: clamp to [lo, hi], (v - lo) / (hi - lo), *255 + 0.5, truncate -- the uint8-range
 * image generate_sr_images_diffusion.py:106-126 takes as `lr_img` (cascade hand-off), produced without leaving the device. */
int b2_image_to_u8(const float* x_nchw, void* out_u8_nhwc, int N, int C, int H, int W, float lo, float hi, void* stream);
/* utils/utils.py:39-65 (plot_sampled_images): channel swap (swap_rb: BGR -> RGB), torchvision make_grid(nrow, padding,
 * normalize=True, value_range=(lo, hi), pad_value=0) and save_image's *255 + 0.5 quantisation, as ONE pass writing the uint8
 * [GH][GW][C] picture.  N == 1: GH = H, GW = W (make_grid returns a single image unframed); otherwise xmaps = min(nrow, N),
 * GH = ceil(N / xmaps) * (H + padding) + padding, GW = xmaps * (W + padding) + padding. */
int b2_image_grid_u8(const float* x_nchw, void* grid_u8_hwc, int N, int C, int H, int W, int nrow, int padding, int swap_rb,
                     float lo, float hi, void* stream);

/* ---- optimiser ------------------------------------------------------------------------------------------- */

/* torch.optim.Adam (train_diffusion.py:214-218: betas (0.5, 0.999), eps 1e-8, no weight decay) over flat fp32
 * buffers: m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= step_size * m / (sqrt(v) * inv_bc2_sqrt + eps), with
 * g = grad * grad_scale (1/world_size after a sum all-reduce), step_size = lr/(1-b1^t), inv_bc2_sqrt = 1/sqrt(1-b2^t).
 * shadow_bf16 (optional): receives bf16(p) in the same pass -- the copy the tensor-core kernels read. */
int b2_adam_flat(float* p, const float* g, float* m, float* v, long long n, double beta1, double beta2, float eps,
                 float step_size, float inv_bc2_sqrt, float grad_scale, void* shadow_bf16, void* stream);

/* CUDA-graph friendly form: `state` is a DEVICE float[8] = {steps taken, lr, grad_scale, (out) step_size, (out)
 * inv_bc2_sqrt, ...}; with advance != 0 the call first advances the step counter on the device, so a captured graph
 * replays correctly.
 * The host changes the learning rate (train_diffusion.py:368-371) by writing state[1]. */
int b2_adam_flat_graph(float* p, const float* g, float* m, float* v, long long n, double beta1, double beta2, float eps,
                       float* state, void* shadow_bf16, int advance, void* stream);
/* Advances the device-side step state alone; b2_adam_flat_graph(..., advance = 0) then updates any number of sub-ranges of
 * the flat buffers with the same bias corrections (bucket-wise updates overlapped with the backward pass). */
int b2_adam_advance(float* state, double beta1, double beta2, void* stream);

/* b2_adam_flat / b2_adam_flat_graph with bf16 gradients (data-parallel transport buffer, reference: none -- the reference is
 * single-process); state == NULL: host scalars step_size / inv_bc2_sqrt / grad_scale, else the device float[8] of the graph form. */
int b2_adam_flat_g16(float* p, const void* g_bf16, float* m, float* v, long long n, double beta1, double beta2, float eps,
                     float step_size, float inv_bc2_sqrt, float grad_scale, float* state, void* shadow_bf16, void* stream);
/* to_f32 == 0: dst_bf16[i] = bf16(src[i]); to_f32 != 0: src[i] = float(dst_bf16[i])  (gradient buckets around the bf16 all-reduce). */
int b2_cast_f32_bf16(const float* src, void* dst_bf16, long long n, int to_f32, void* stream);

/* b2_conv2d_nhwc (mode 0, bf16) whose epilogue also accumulates the pass-1 sums of the GroupNorm x AdaGN backward that will
 * consume y as its `dout` (autograd of custom_layers.py:35-45 after :224-245): per (image, output channel)
 * cs_s1 += sum_p y[p], cs_s2 += sum_p y[p] * swish(cs_z[p]); cs_z = that layer's pre-activation [N][H][W][Cout] (per-pixel
 * stride cs_ldz), cs_s1 / cs_s2 = [N][Cout] fp32 zeroed by the caller.  Pair with b2_adagn_bwd_fused(sums_ready = 1). */
int b2_conv2d_nhwc_colsum(int mode, const void* x, int N, int H, int W, int Cin, long long ldx, const void* wpacked,
                          const float* bias, int Cout, void* y, long long ldy, int act, const void* residual, long long ldr,
                          float* gn_stats, int gn_groups, int out_mode, int dtype, void* stream, const void* cs_z,
                          long long cs_ldz, float* cs_s1, float* cs_s2);
/* b2_conv2d_nhwc (modes 0..2, act 0, no statistics / residual) with a SECOND output y_act = Swish(y) (NHWC, per-pixel stride
 * ldy_act, same dtype): the training forward of the un-normalised convs (custom_layers.py:224-245 with use_norm False, :174-201)
 * keeps the pre-activation y for autograd and hands Swish(y) on -- no separate activation pass.  y_act equals b2_act(mode 0) of
 * the stored y bit for bit. */
int b2_conv2d_nhwc_dual(int mode, const void* x, int N, int H, int W, int Cin, long long ldx, const void* wpacked,
                        const float* bias, int Cout, void* y, long long ldy, void* y_act, long long ldy_act, int dtype,
                        void* stream);
/* b2_adagn_bwd with sums_ready != 0: `work` = [2][N][C] already holds (sum_p dout, sum_p dout * swish(z)) from
 * b2_conv2d_nhwc_colsum; the reduce pass is skipped (one pass of 6 bytes per element instead of 10). */
int b2_adagn_bwd_fused(const void* dout, long long ldd, const void* z, long long ldz, const float* stats, const float* gamma,
                       const float* beta, const float* s, long long s_bstride, float* work, float* ds, long long ds_bstride,
                       float* dgamma, float* dbeta, void* dz, long long lddz, float* dbias, int N, int HW, int C, int groups,
                       float eps, int sums_ready, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDM_B200_H */
