// Device pieces of the denoising-score-matching losses (losses.py:55-150 in the reference): the forward perturbation of a
// batch and the per-image reduction of the squared residual.  The network evaluation in between is the engine's
// (b200_ncsnpp_forward); the per-image scalars (marginal mean coefficient, std, sigma, g^2) come from the SDE's own torch
// ops on [B] tensors, exactly as the reference computes them.
//
// Arithmetic follows the reference's op for op in fp32 (separate multiply / divide / add roundings, no FMA contraction), so
// the perturbed batch is bit-equal to the reference's and the residual differs only by the reduction order (fp64 partial
// sums in a fixed order here: deterministic, and closer to the exact sum than a float32 tree).
#include "kernels.h"
#include "../../include/scoresde_b200.h"

namespace b200 {
namespace {

constexpr int DSM_THREADS = 256;
constexpr int DSM_MAX_PARTS = 64;     // partial sums per image

// out = a[img] * x + s[img] * z      (losses.py:86-87: mean + std[:, None, None, None] * z with mean = a * x;
//                                     :111-112 SMLD: a = 1, s = sigma; :133-134 DDPM: a = sqrt_alphas_cumprod, s = sqrt_1m_alphas_cumprod)
__global__ void __launch_bounds__(DSM_THREADS) dsm_perturb_kernel(const float* __restrict__ x, const float* __restrict__ z,
                                                                  const float* __restrict__ a, const float* __restrict__ s,
                                                                  float* __restrict__ out, long long n_per_img) {
  pdl_wait(); pdl_trigger();
  const int img = blockIdx.y;
  const float ai = a ? __ldg(a + img) : 1.f, si = __ldg(s + img);
  const long long base = (long long)img * n_per_img;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_per_img; i += (long long)gridDim.x * blockDim.x) {
    const float m = a ? __fmul_rn(ai, x[base + i]) : x[base + i];
    out[base + i] = __fadd_rn(m, __fmul_rn(si, z[base + i]));
  }
}

// residual of one element, per loss (all fp32, the reference's operation order):
//   mode 0  score * std + z                       get_sde_loss_fn, likelihood_weighting=False   (losses.py:90-91)
//   mode 1  score + z / std                       get_sde_loss_fn, likelihood_weighting=True    (:94-95)
//   mode 2  score - (-(z * sigma) / sigma^2)      get_smld_loss_fn (:110-115; w = sigma, w2 = sigma ** 2)
//   mode 3  score - z                             get_ddpm_loss_fn (:136)
__device__ __forceinline__ float dsm_residual(int mode, float sc, float zz, float w, float w2) {
  if (mode == 0) return __fadd_rn(__fmul_rn(sc, w), zz);
  if (mode == 1) return __fadd_rn(sc, __fdiv_rn(zz, w));
  if (mode == 2) return __fsub_rn(sc, __fdiv_rn(-__fmul_rn(zz, w), w2));
  return __fsub_rn(sc, zz);
}

__global__ void __launch_bounds__(DSM_THREADS) dsm_partial_kernel(const float* __restrict__ score, const float* __restrict__ z,
                                                                  const float* __restrict__ w, const float* __restrict__ w2,
                                                                  long long n_per_img, int mode, double* __restrict__ part) {
  pdl_wait(); pdl_trigger();
  const int img = blockIdx.y;
  const float wi = w ? __ldg(w + img) : 1.f, w2i = w2 ? __ldg(w2 + img) : 1.f;
  const long long base = (long long)img * n_per_img;
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_per_img; i += (long long)gridDim.x * blockDim.x) {
    const float r = dsm_residual(mode, score[base + i], z[base + i], wi, w2i);
    acc += (double)__fmul_rn(r, r);                    // torch.square in fp32, then the sum
  }
  __shared__ double sh[DSM_THREADS / 32];
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = threadIdx.x < DSM_THREADS / 32 ? sh[threadIdx.x] : 0.0;
    t = warp_sum_d(t);
    if (threadIdx.x == 0) part[(long long)img * gridDim.x + blockIdx.x] = t;
  }
}

// losses[img] = reduce_mean ? sum / n : 0.5 * sum        (losses.py:71: torch.mean, or 0.5 * torch.sum)
__global__ void dsm_final_kernel(const double* __restrict__ part, int parts, long long n_per_img, int reduce_mean, float* __restrict__ out, int nimg) {
  pdl_wait(); pdl_trigger();
  const int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= nimg) return;
  double s = 0.0;
  for (int j = 0; j < parts; ++j) s += part[(long long)img * parts + j];
  out[img] = (float)(reduce_mean ? s / (double)n_per_img : 0.5 * s);
}

int dsm_parts(long long n_per_img) {
  return (int)std::max<long long>(1, std::min<long long>((n_per_img + DSM_THREADS * 8 - 1) / (DSM_THREADS * 8), DSM_MAX_PARTS));
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int b200_dsm_perturb_f32(const float* x, const float* z, const float* mean_coef, const float* noise_coef, float* out, int nimg,
                         long long n_per_img, void* stream) {
  B200_REQUIRE(x && z && noise_coef && out && nimg > 0 && nimg <= 65535 && n_per_img > 0, "dsm_perturb: bad argument");
  const int gx = (int)std::max<long long>(1, std::min<long long>((n_per_img + DSM_THREADS - 1) / DSM_THREADS, 1024));
  launch_kernel(dsm_perturb_kernel, dim3(gx, nimg), dim3(DSM_THREADS), 0, static_cast<cudaStream_t>(stream), x, z, mean_coef, noise_coef, out, n_per_img);
  B200_CHECK_LAUNCH();
  return 0;
}

long long b200_dsm_workspace_doubles(int nimg, long long n_per_img) { return (long long)nimg * dsm_parts(n_per_img); }

int b200_dsm_loss_f32(const float* score, const float* z, const float* w, const float* w2, float* losses, int nimg,
                      long long n_per_img, int mode, int reduce_mean, double* ws, void* stream) {
  B200_REQUIRE(score && z && losses && ws && nimg > 0 && nimg <= 65535 && n_per_img > 0, "dsm_loss: bad argument");
  B200_REQUIRE(mode >= 0 && mode <= 3 && (mode == 3 || w) && (mode != 2 || w2), "dsm_loss: mode %d with missing per-image scalars", mode);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int parts = dsm_parts(n_per_img);
  launch_kernel(dsm_partial_kernel, dim3(parts, nimg), dim3(DSM_THREADS), 0, st, score, z, w, w2, n_per_img, mode, ws);
  launch_kernel(dsm_final_kernel, dim3((nimg + 127) / 128), dim3(128), 0, st, (const double*)ws, parts, n_per_img, reduce_mean, losses, nimg);
  B200_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
