// HBM-bound kernels of the NCSN++ forward: GroupNorm statistics / apply(+SiLU),
// FIR resampling (upfirdn2d), fused bias+activation, row softmax, the
// time-embedding path, layout/packing utilities.
//
// All activations are NHWC fp32 ([image][pixel][channel], channel contiguous) so a
// 128-bit access covers 4 channels of one pixel and a GroupNorm group (4/8/12/16
// channels here) never straddles a float4.
#include "common.cuh"
#include "kernels.h"

namespace b200 {

// ============================================================================
// GroupNorm statistics.  Reference semantics: nn.GroupNorm(min(C/4,32), C, eps=1e-6)
// (layerspp.py:67,219,231; ncsnpp.py:226) -> biased variance over (C/G)*H*W.
// One CTA per image streams the whole [HW, C] slab once (coalesced float4),
// accumulating per-thread sums in fp64 (robust to mean^2 >> var cancellation),
// then folds threads that share a group through shared-memory fp64 atomics.
// The input may be a virtual channel-concat of two tensors (U-Net skip joins,
// ncsnpp.py:318) so torch.cat never materialises.
// ============================================================================
__global__ void __launch_bounds__(384) gn_stats_kernel(
    const float* __restrict__ x1, int C1, const float* __restrict__ x2, int C2,
    int HW, int G, float eps, float2* __restrict__ stats) {
  extern __shared__ double sacc[];   // [2*G]
  const int C = C1 + C2, Q = C >> 2, cpg = C / G;
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sacc[i] = 0.0;
  __syncthreads();
  const long long units = (long long)HW * Q;
  const float* p1 = x1 + (long long)b * HW * C1;
  const float* p2 = x2 ? x2 + (long long)b * HW * C2 : nullptr;
  double s = 0.0, ss = 0.0;
  int cur_g = -1;
  for (long long u = threadIdx.x; u < units; u += blockDim.x) {
    const int pix = (int)(u / Q), c0 = (int)(u % Q) << 2;
    const int g = c0 / cpg;
    if (g != cur_g) {
      if (cur_g >= 0) { atomicAdd(&sacc[2 * cur_g], s); atomicAdd(&sacc[2 * cur_g + 1], ss); }
      s = 0.0; ss = 0.0; cur_g = g;
    }
    float4 v = (c0 < C1) ? __ldg(reinterpret_cast<const float4*>(p1 + (long long)pix * C1 + c0))
                         : __ldg(reinterpret_cast<const float4*>(p2 + (long long)pix * C2 + (c0 - C1)));
    s += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
    ss += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  if (cur_g >= 0) { atomicAdd(&sacc[2 * cur_g], s); atomicAdd(&sacc[2 * cur_g + 1], ss); }
  __syncthreads();
  const double n = (double)HW * cpg;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const double mean = sacc[2 * g] / n;
    double var = sacc[2 * g + 1] / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[(long long)b * G + g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
  }
}

int launch_gn_stats(const float* x1, int C1, const float* x2, int C2, int B, int HW, int G,
                    float eps, float* stats, cudaStream_t st) {
  const int C = C1 + C2;
  B200_REQUIRE(C % 4 == 0 && C1 % 4 == 0 && C % G == 0 && (C / G) % 4 == 0,
               "gn_stats: C=%d (C1=%d) G=%d must give 4-aligned groups", C, C1, G);
  gn_stats_kernel<<<B, 384, 2 * G * sizeof(double), st>>>(x1, C1, x2, C2, HW, G, eps,
                                                          reinterpret_cast<float2*>(stats));
  B200_CHECK_LAUNCH();
  return 0;
}

// ============================================================================
// GroupNorm apply (+ optional SiLU, + optional TF32 rounding of the stored value
// because the consumer is a tcgen05 kind::tf32 contraction).  Optionally also
// emits `raw`: the TF32-rounded *un-normalised* (concatenated) input, which is the
// A operand of the ResBlock's 1x1 skip convolution (layerspp.py:268-269).
// ============================================================================
__global__ void __launch_bounds__(256) gn_apply_kernel(
    const float* __restrict__ x1, int C1, const float* __restrict__ x2, int C2,
    const float2* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
    long long total_units, int HW, int G, int act, int round_out,
    float* __restrict__ y, float* __restrict__ raw) {
  const int C = C1 + C2, Q = C >> 2, cpg = C / G;
  for (long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x; u < total_units;
       u += (long long)gridDim.x * blockDim.x) {
    const long long pixg = u / Q;               // global pixel index (b*HW + pix)
    const int c0 = (int)(u % Q) << 2;
    const int b = (int)(pixg / HW);
    float4 v = (c0 < C1) ? __ldg(reinterpret_cast<const float4*>(x1 + pixg * C1 + c0))
                         : __ldg(reinterpret_cast<const float4*>(x2 + pixg * C2 + (c0 - C1)));
    const float2 mr = __ldg(&stats[(long long)b * G + c0 / cpg]);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c0));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c0));
    float4 o;
    o.x = (v.x - mr.x) * mr.y * ga.x + be.x;
    o.y = (v.y - mr.x) * mr.y * ga.y + be.y;
    o.z = (v.z - mr.x) * mr.y * ga.z + be.z;
    o.w = (v.w - mr.x) * mr.y * ga.w + be.w;
    if (act) { o.x = silu_f(o.x); o.y = silu_f(o.y); o.z = silu_f(o.z); o.w = silu_f(o.w); }
    if (round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
    *reinterpret_cast<float4*>(y + pixg * C + c0) = o;
    if (raw) {
      float4 r = v;
      if (round_out) { r.x = round_tf32(r.x); r.y = round_tf32(r.y); r.z = round_tf32(r.z); r.w = round_tf32(r.w); }
      *reinterpret_cast<float4*>(raw + pixg * C + c0) = r;
    }
  }
}

int launch_gn_apply(const float* x1, int C1, const float* x2, int C2, const float* stats,
                    const float* gamma, const float* beta, int B, int HW, int G, int act,
                    int round_out, float* y, float* raw, cudaStream_t st) {
  const int C = C1 + C2;
  const long long units = (long long)B * HW * (C / 4);
  const int grid = (int)std::min<long long>((units + 255) / 256, 148LL * 32);
  gn_apply_kernel<<<grid, 256, 0, st>>>(x1, C1, x2, C2, reinterpret_cast<const float2*>(stats), gamma,
                                        beta, units, HW, G, act, round_out, y, raw);
  B200_CHECK_LAUNCH();
  return 0;
}

// ============================================================================
// upfirdn2d: zero-insert upsample (up), zero pad, correlate with the flipped FIR,
// decimate (down).  Successor of the reference's native op
// (op/upfirdn2d_kernel.cu:49-207, host dispatch :209-369); same tensor convention
// [major, in_h, in_w, minor] -> [major, out_h, out_w, minor].  The reference always
// calls it with major=N*C, minor=1 (op/upfirdn2d.py:99); the engine calls it with
// major=N, minor=C, i.e. directly on NHWC, so a thread produces VEC consecutive
// channels of one output pixel with 128-bit accesses and the FIR taps in registers.
//   out[oy,ox] = sum_{a,b} k[kh-1-a, kw-1-b] * u[oy*down_y + a, ox*down_x + b]
//   u[Y,X] = x[(Y-pad_y0)/up_y, (X-pad_x0)/up_x] when divisible and in range, else 0.
// ============================================================================
struct FirParams {
  int major, in_h, in_w, minor, out_h, out_w;
  int kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_y0;
  int round_out;
  float k[64];   // row-major [kh][kw], kh*kw <= 64
};

template <int VEC>
__global__ void __launch_bounds__(256) upfirdn2d_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                       const FirParams p) {
  const int mv = p.minor / VEC;
  const long long total = (long long)p.major * p.out_h * p.out_w * mv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cm = (int)(i % mv) * VEC;
    long long t = i / mv;
    const int ox = (int)(t % p.out_w); t /= p.out_w;
    const int oy = (int)(t % p.out_h);
    const int n = (int)(t / p.out_h);
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    for (int a = 0; a < p.kh; ++a) {
      const int Y = oy * p.down_y + a - p.pad_y0;
      if (Y < 0 || (Y % p.up_y) != 0) continue;
      const int iy = Y / p.up_y;
      if (iy >= p.in_h) continue;
      for (int b = 0; b < p.kw; ++b) {
        const int X = ox * p.down_x + b - p.pad_x0;
        if (X < 0 || (X % p.up_x) != 0) continue;
        const int ix = X / p.up_x;
        if (ix >= p.in_w) continue;
        const float w = p.k[(p.kh - 1 - a) * p.kw + (p.kw - 1 - b)];
        const float* src = x + (((long long)n * p.in_h + iy) * p.in_w + ix) * p.minor + cm;
        if (VEC == 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src));
          acc[0] += w * v.x; acc[1] += w * v.y; acc[2] += w * v.z; acc[3] += w * v.w;
        } else {
          acc[0] += w * __ldg(src);
        }
      }
    }
    float* dst = y + (((long long)n * p.out_h + oy) * p.out_w + ox) * p.minor + cm;
    if (p.round_out) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] = round_tf32(acc[v]);
    }
    if (VEC == 4) *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else dst[0] = acc[0];
  }
}

int launch_upfirdn2d(const float* x, const float* kernel_host, float* y, int major, int in_h, int in_w,
                     int minor, int kh, int kw, int up_x, int up_y, int down_x, int down_y,
                     int pad_x0, int pad_x1, int pad_y0, int pad_y1, int round_out, cudaStream_t st) {
  B200_REQUIRE(kh * kw <= 64 && kh > 0 && kw > 0, "upfirdn2d: FIR %dx%d exceeds 64 taps", kh, kw);
  B200_REQUIRE(up_x > 0 && up_y > 0 && down_x > 0 && down_y > 0, "upfirdn2d: up/down must be positive");
  FirParams p;
  p.major = major; p.in_h = in_h; p.in_w = in_w; p.minor = minor;
  p.out_h = (in_h * up_y + pad_y0 + pad_y1 - kh) / down_y + 1;
  p.out_w = (in_w * up_x + pad_x0 + pad_x1 - kw) / down_x + 1;
  p.kh = kh; p.kw = kw; p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y;
  p.pad_x0 = pad_x0; p.pad_y0 = pad_y0; p.round_out = round_out;
  for (int i = 0; i < kh * kw; ++i) p.k[i] = kernel_host[i];
  B200_REQUIRE(p.out_h > 0 && p.out_w > 0, "upfirdn2d: empty output %dx%d", p.out_h, p.out_w);
  const bool vec = (minor % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) % 16 == 0);
  const long long total = (long long)major * p.out_h * p.out_w * (vec ? minor / 4 : minor);
  if (total == 0) return 0;
  const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 64);
  if (vec) upfirdn2d_kernel<4><<<grid, 256, 0, st>>>(x, y, p);
  else upfirdn2d_kernel<1><<<grid, 256, 0, st>>>(x, y, p);
  B200_CHECK_LAUNCH();
  return 0;
}

// ============================================================================
// fused bias + activation (successor of op/fused_bias_act_kernel.cu:19-98):
// y = act(x + b[(i / step_b) % size_b]) * scale, act 1 = linear, 3 = leaky-relu(alpha);
// grad 1 gates on `ref` instead of x (first derivative), grad 2 yields zeros.
// ============================================================================
__global__ void __launch_bounds__(256) fused_bias_act_kernel(
    const float* __restrict__ x, const float* __restrict__ b, const float* __restrict__ ref,
    float* __restrict__ y, long long n, int step_b, int size_b, int act, int grad, float alpha, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float v = x[i];
    if (b) v += __ldg(&b[(i / step_b) % size_b]);
    const float r = ref ? ref[i] : 0.f;
    float o;
    if (grad == 2) o = 0.f;
    else if (act == 3) o = ((grad == 1 ? r : v) > 0.f) ? v : v * alpha;
    else o = v;
    y[i] = o * scale;
  }
}

int launch_fused_bias_act(const float* x, const float* b, const float* ref, float* y, long long n,
                          int step_b, int size_b, int act, int grad, float alpha, float scale,
                          cudaStream_t st) {
  if (n == 0) return 0;
  B200_REQUIRE(act == 1 || act == 3, "fused_bias_act: act=%d unsupported (1 linear, 3 lrelu)", act);
  B200_REQUIRE(!b || (step_b > 0 && size_b > 0), "fused_bias_act: bad bias geometry");
  const int grid = (int)std::min<long long>((n + 255) / 256, 148LL * 64);
  fused_bias_act_kernel<<<grid, 256, 0, st>>>(x, b, ref, y, n, step_b, size_b, act, grad, alpha, scale);
  B200_CHECK_LAUNCH();
  return 0;
}

// ============================================================================
// Row softmax for the attention logits (layerspp.py:82-85): rows of length T,
// logits pre-multiplied by `scale` = C^-1/2.  One warp per row, values held in
// registers (T <= 1024), warp-shuffle max/sum.
// ============================================================================
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, float* __restrict__ p,
                                                          long long rows, int T, float scale, int round_out) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* src = s + row * T;
  float v[32];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int c = lane + j * 32;
    v[j] = (c < T) ? src[c] * scale : -INFINITY;
    mx = fmaxf(mx, v[j]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int c = lane + j * 32;
    v[j] = (c < T) ? expf(v[j] - mx) : 0.f;
    sum += v[j];
  }
  sum = warp_sum(sum);
  float* dst = p + row * T;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int c = lane + j * 32;
    if (c < T) {
      float o = v[j] / sum;
      dst[c] = round_out ? round_tf32(o) : o;
    }
  }
}

int launch_softmax_rows(const float* s, float* p, long long rows, int T, float scale, int round_out,
                        cudaStream_t st) {
  B200_REQUIRE(T > 0 && T <= 1024, "softmax_rows: T=%d out of range (1..1024)", T);
  const int wpb = 8;
  softmax_rows_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(s, p, rows, T, scale, round_out);
  B200_CHECK_LAUNCH();
  return 0;
}

// ============================================================================
// Time-embedding path (ncsnpp.py:236-255, layerspp.py:39-41).
// fourier: emb[r] = [sin(p), cos(p)], p = ((log(sigma_r) * W_j) * 2) * fp32(pi) with the
// reference's operation order; accurate sinf/cosf/logf (phases reach ~1e3 rad).
// ============================================================================
__global__ void fourier_embed_kernel(const float* __restrict__ sigma, long long sigma_stride,
                                     const float* __restrict__ W, int nf, float* __restrict__ emb) {
  const int r = blockIdx.x;
  const float lv = logf(sigma[r * sigma_stride]);
  for (int j = threadIdx.x; j < nf; j += blockDim.x) {
    const float ph = ((lv * W[j]) * 2.0f) * 3.14159265358979323846f;
    emb[(long long)r * 2 * nf + j] = sinf(ph);
    emb[(long long)r * 2 * nf + nf + j] = cosf(ph);
  }
}

int launch_fourier_embed(const float* sigma, long long sigma_stride, const float* W, int nf, int rows,
                         float* emb, cudaStream_t st) {
  fourier_embed_kernel<<<rows, 128, 0, st>>>(sigma, sigma_stride, W, nf, emb);
  B200_CHECK_LAUNCH();
  return 0;
}

// y[r][n] = sum_k act(x[r][k]) * W[n][k] + b[n]   (torch Linear layout, fp32 exact).
// A CTA stages up to RB rows of act(x) in shared memory; each warp owns output
// columns and reuses every W row it loads for all staged rows.
constexpr int LIN_RB = 8;
__global__ void __launch_bounds__(256) linear_rows_kernel(
    const float* __restrict__ x, long long ldx, const float* __restrict__ W, const float* __restrict__ bias,
    int rows, int N, int K, int act_in, float* __restrict__ y, long long ldy) {
  extern __shared__ float sx[];   // [LIN_RB][K]
  const int r0 = blockIdx.y * LIN_RB;
  const int nr = min(LIN_RB, rows - r0);
  for (int i = threadIdx.x; i < nr * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    float v = x[(long long)(r0 + r) * ldx + k];
    sx[r * K + k] = act_in ? silu_f(v) : v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int n = blockIdx.x * nwarps + warp; n < N; n += gridDim.x * nwarps) {
    float acc[LIN_RB];
#pragma unroll
    for (int r = 0; r < LIN_RB; ++r) acc[r] = 0.f;
    const float* w = W + (long long)n * K;
    for (int k = lane; k < K; k += 32) {
      const float wv = __ldg(&w[k]);
#pragma unroll
      for (int r = 0; r < LIN_RB; ++r)
        if (r < nr) acc[r] += sx[r * K + k] * wv;
    }
#pragma unroll
    for (int r = 0; r < LIN_RB; ++r) {
      const float t = warp_sum(acc[r]);
      if (lane == 0 && r < nr) y[(long long)(r0 + r) * ldy + n] = t + (bias ? bias[n] : 0.f);
    }
  }
}

int launch_linear_rows(const float* x, long long ldx, const float* W, const float* bias, int rows, int N,
                       int K, int act_in, float* y, long long ldy, cudaStream_t st) {
  B200_REQUIRE(K * LIN_RB * 4 <= 96 * 1024, "linear_rows: K=%d too large for the row stage", K);
  const size_t smem = (size_t)LIN_RB * K * sizeof(float);
  if (smem > 48 * 1024)
    B200_CHECK_CUDA(cudaFuncSetAttribute(linear_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(std::min(ceil_div(N, 8), 148 * 4), ceil_div(rows, LIN_RB));
  linear_rows_kernel<<<grid, 256, smem, st>>>(x, ldx, W, bias, rows, N, K, act_in, y, ldy);
  B200_CHECK_LAUNCH();
  return 0;
}

// ============================================================================
// Utilities
// ============================================================================
__global__ void fill_from_table_kernel(const float* __restrict__ table, const int* __restrict__ step,
                                       float* __restrict__ dst, int n) {
  const float v = table[*step];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = v;
}
int launch_fill_from_table(const float* table, const int* step, float* dst, int n, cudaStream_t st) {
  fill_from_table_kernel<<<ceil_div(n, 256), 256, 0, st>>>(table, step, dst, n);
  B200_CHECK_LAUNCH();
  return 0;
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int HW, int C) {
  const long long total = (long long)B * HW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const long long t = i / HW;
    const int c = (int)(t % C);
    const long long b = t / C;
    dst[i] = src[(b * HW + p) * C + c];
  }
}
int launch_nhwc_to_nchw(const float* src, float* dst, int B, int HW, int C, cudaStream_t st) {
  const long long total = (long long)B * HW * C;
  nhwc_to_nchw_kernel<<<(int)std::min<long long>((total + 255) / 256, 148LL * 64), 256, 0, st>>>(src, dst, B, HW, C);
  B200_CHECK_LAUNCH();
  return 0;
}

// dst[tap][o][i] = src[o*so + i*si + tap*st]  (OIHW conv weights: so=I*R*S, si=R*S, st=1;
// NIN W[in][out]: taps=1, so=1, si=out).  Optional TF32 rounding for tensor-core layers.
__global__ void pack_weight_kernel(const float* __restrict__ src, float* __restrict__ dst, int taps, int O, int I,
                                   long long so, long long si, long long stp, int round_out) {
  const long long total = (long long)taps * O * I;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % I);
    const long long t = idx / I;
    const int o = (int)(t % O);
    const int tap = (int)(t / O);
    const float v = src[o * so + i * si + tap * stp];
    dst[idx] = round_out ? round_tf32(v) : v;
  }
}
int launch_pack_weight(const float* src, float* dst, int taps, int O, int I, long long so, long long si,
                       long long stp, int round_out, cudaStream_t st) {
  const long long total = (long long)taps * O * I;
  pack_weight_kernel<<<(int)std::min<long long>((total + 255) / 256, 148LL * 16), 256, 0, st>>>(
      src, dst, taps, O, I, so, si, stp, round_out);
  B200_CHECK_LAUNCH();
  return 0;
}

}  // namespace b200
