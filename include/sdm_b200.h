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

/* ---- dense contractions (tcgen05 implicit GEMM) ------------------------------------------------------- */

/* mode 0: Conv2d 3x3 stride 1 pad 1            (custom_layers.py:224-228), x = [N][H][W][Cin]
 * mode 1: Conv2d 3x3 stride 2 pad 1            (custom_layers.py:196-201), x = parity planes
 *         [2][2][N][H][W][Cin] made by b2_space_to_depth2, (H, W) = OUTPUT size
 * mode 2: ConvTranspose2d 4x4 stride 2 pad 1   (custom_layers.py:174-179), (H, W) = INPUT size, y is 2H x 2W
 * wpacked: weights in kernel layout from b2_pack_conv_weight. act: 0 none, 1 Swish (custom_layers.py:18-20).
 * residual (optional, mode 0/1): added after the activation. gn_stats (optional): [N][gn_groups][2] fp32,
 * must be zeroed by the caller; receives per-(image, group) sum and sum of squares of the written values. */
int b2_conv2d_nhwc(int mode, const void* x, int N, int H, int W, int Cin, long long ldx, const void* wpacked,
                   const float* bias, int Cout, void* y, long long ldy, int act, const void* residual,
                   long long ldr, float* gn_stats, int gn_groups, int dtype, void* stream);

/* C = alpha * A . B^T (+bias) (act) (+residual); A [M][K], B [Ncols][K] (nn.Linear weight layout,
 * custom_layers.py:116,119; q.k^T custom_layers.py:144).  batch1/batch2 > 1: batched with element strides
 * *_s1 / *_s2 for A, B and C. */
int b2_gemm_nt(const void* A, long long lda, long long a_s1, long long a_s2, const void* B, long long ldb,
               long long b_s1, long long b_s2, void* C, long long ldc, long long c_s1, long long c_s2, int M,
               int Ncols, int K, int batch1, int batch2, const float* bias, float alpha, int act,
               const void* residual, long long ldr, int out_fp32, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDM_B200_H */
