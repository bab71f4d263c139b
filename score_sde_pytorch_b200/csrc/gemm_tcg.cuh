// 3x3 convolution that applies GroupNorm (+ SiLU) to its input ON LOAD: the normalised activation never exists in
// HBM.  Included by gemm_tc.cu inside its anonymous namespace (PTX wrappers, descriptors, epilogue block routines and
// the cluster helpers of gemm_tc2.cuh are reused).  fp16 operands (tcgen05 kind::f16), stride 1, 'same' padding,
// image width 16 or 32, C_out a multiple of 256, C_in a multiple of 64 per source.
//
// Replaces, per resblock (models/layerspp.py:242-266 in the reference), the pair
//     a = SiLU(GroupNorm(x))  [gn_apply_stream_kernel: read x, write a (+ a rounded copy of x for the skip projection)]
//     h = Conv3x3(a)          [gemm_tc2_kernel: nine TMA tap loads of a per 64-channel K chunk]
// by one kernel that reads x once per tile (plus a one-row halo above and below).  Round 1 measured the separate
// GroupNorm pass at 21 % of a PC step and 71 GB of DRAM traffic per step (profiles/r01_c34_dram_traffic.md).
//
// How the operand is built.  A CTA owns 128 consecutive pixels = R = 128/W whole image rows.  For one 64-channel
// chunk, four TRANSFORM warps read the (R+2) x W patch (rows above/below the image are zero) straight from global
// memory with 128-bit loads, apply  y = SiLU(x * scale[img][c] + shift[img][c])  in fp32 (the coefficients come from
// gn_coeff_kernel: same arithmetic as the stand-alone GroupNorm kernel, so the fp16 operand is bit-identical), round
// to fp16 and store the patch K-major with the 128-byte swizzle the UMMA descriptor expects: "copy 0".  Two more
// copies of the patch are derived in shared memory, shifted by one pixel along W with the column that would wrap
// around zeroed (dw = -1 and dw = +1).  With those three copies every filter tap (dh, dw) is a plain descriptor:
// copy[dw] starting (dh+1)*W pixel rows in (a multiple of 1024 B for W = 16 or 32, so no base-offset games), 128
// rows long.  The transform does 1.25-1.5x the elements of a stand-alone pass (halo rows) instead of the 9x a
// per-tap transform would need, the A side writes 72 KB into shared memory per chunk where nine TMA tap loads wrote
// 144 KB, and the weights stream through their own TMA ring exactly as in gemm_tc2_kernel.
//
// The optional extra 1x1 phase (a resblock's skip projection, layerspp.py:268-274) goes through the same path with an
// identity transform: fp32 block input -> fp16 operand, one un-shifted 128-row copy per chunk, so the rounded copy of
// the block input that the GroupNorm pass used to write is gone as well.
//
// Roles per CTA (640 threads; setmaxnreg moves registers from the producer / issuer warpgroup to the epilogue warpgroups): warp 0 lane 0 weight TMA producer (both CTAs), warp 1 lane 0 of the leader issues
// tcgen05.mma.cta_group::2 for the pair, warp 2 TMEM allocation, warps 4..11 epilogue (identical to gemm_tc2_kernel),
// warps 12..19 transform.  Barriers: wfull/wempty (weight ring), tfull/tempty (ring of operand copies; tfull lives in
// the leader and counts one arrival per transform warp of BOTH CTAs), tmem_full/tmem_empty.

constexpr int TG_TS = 5;                    // operand-copy slots (three per chunk in flight + two being built)
constexpr int TG_WS = 4;                    // weight ring stages
constexpr int TG_SLOT_BYTES = 192 * 128;    // (R+2)*W <= 192 pixel rows of 64 fp16 channels
constexpr int TG_W_BYTES = 128 * 128;       // this CTA's 128 of the 256 output channels x 64 k
constexpr int TG_NTW = 8;                   // transform warps per CTA (two warpgroups)
constexpr int TG_THREADS = 384 + 32 * TG_NTW;

struct SmemG {
  static constexpr int W_OFFSET = TG_TS * TG_SLOT_BYTES;
  static constexpr int TRN_OFFSET = W_OFFSET + TG_WS * TG_W_BYTES;    // 8 epilogue warps x 4 KB transposition scratch
  static constexpr int BAR_OFFSET = TRN_OFFSET + 8 * 4096;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
  static_assert(TOTAL <= 232448, "exceeds the 227 KB shared-memory limit of sm_100");
};

struct TcgParams {
  CUtensorMap tmW, tmW2;
  const void* src[4];        // 0,1: 3x3 phase (channel concat), 2,3: extra 1x1 phase
  int srcC[4], srcF16[4], kch[4];
  const float* scale; const float* shift; int Cgn;   // [img][Cgn] GroupNorm affine of the 3x3 phase (null: identity)
  int act;
  int H, W, R;               // image size; R = 128 / W rows per tile
  int N_total, tiles_n;
  long long tiles_m;         // 128-pixel tiles
  double* qstats;
  Epilogue epi;
};

__device__ __forceinline__ uint4 ldg128(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
// SiLU on the MUFU pipe, five instructions: x * rcp(1 + 2^(-x log2 e)).  (__fdividef / __expf wrap the same two MUFU ops in
// range-scaling code - 9-10 instructions per element, measured in the first version of the transform.)  x -> +inf: e flushes
// to 0, y = x; x -> -inf: e = inf, rcp = 0, y = -0.  elementwise.cu's silu_fast is the same function.
__device__ __forceinline__ float silu_approx(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}

// eight consecutive channels of one pixel: raw bits as loaded (fp32: 32 B in a|b, fp16: 16 B in a)
struct Raw8 { uint4 a, b; };
template <bool F16IN>
__device__ __forceinline__ Raw8 load_raw8(const void* gp) {
  Raw8 r;
  r.a = ldg128(gp);
  r.b = F16IN ? make_uint4(0u, 0u, 0u, 0u) : ldg128(reinterpret_cast<const uint8_t*>(gp) + 16);
  return r;
}
// -> eight fp16 (four packed words) after the affine (GroupNorm; scale 1 / shift 0 is the exact identity) and the
// optional SiLU, computed in fp32
struct Coef8 { float4 s0, s1, h0, h1; };      // scale[0..7], shift[0..7] of this thread's channel octet
template <bool F16IN>
__device__ __forceinline__ uint4 apply8(const Raw8 r, const Coef8 cf, bool act) {
  const float sc[8] = {cf.s0.x, cf.s0.y, cf.s0.z, cf.s0.w, cf.s1.x, cf.s1.y, cf.s1.z, cf.s1.w};
  const float sh[8] = {cf.h0.x, cf.h0.y, cf.h0.z, cf.h0.w, cf.h1.x, cf.h1.y, cf.h1.z, cf.h1.w};
  float v[8];
  if constexpr (F16IN) {
    const uint32_t w[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  } else {
    v[0] = __uint_as_float(r.a.x); v[1] = __uint_as_float(r.a.y); v[2] = __uint_as_float(r.a.z); v[3] = __uint_as_float(r.a.w);
    v[4] = __uint_as_float(r.b.x); v[5] = __uint_as_float(r.b.y); v[6] = __uint_as_float(r.b.z); v[7] = __uint_as_float(r.b.w);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], sc[i], sh[i]);
  if (act) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = silu_approx(v[i]);
  }
  return make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
}

// One chunk of the operand for one transform thread, as straight-line code.  The patch is a CONTIGUOUS range of 32 * NI
// NHWC pixels starting one image row above the tile, so with 256 transform threads thread t owns channel octet t & 7 of
// patch pixels pl + 32 j (pl = t >> 3, j < NI), and everything about an item except j is a per-thread constant: its column
// w = pl & (W-1) (W divides 32), its swizzle phase pl & 7, hence its shared-memory address in each of the three copies
// (a0 / a1 / a2 + 4096 j, immediate offsets) and its global address (gp + gstep j).  Copy -1 holds T[hh][w-1] at (hh, w):
// this thread's value goes one pixel row further (pl + 1), except that the thread owning column W-1 writes the zero of
// column 0 instead (W-1 rows back: same swizzle phase): m1 = 0 for that thread, ~0 otherwise; copy +1 mirrored (m2).
// Only the first and the last item can lie in a halo row outside the image: their results are ANDed with mfirst / mlast
// (0 or ~0; fp16 zero is all-zero bits) and their loads are redirected to the neighbouring item's address by the caller,
// so there is not a single branch per item.  Loads run two items ahead of the arithmetic (64 B per thread in flight).
// The first version of this routine (runtime trip count, per-item validity branches, __fdividef/__expf) executed 265
// instructions per item, 80 of them arithmetic; the transform, not the tensor pipe, set the pace (profiles/r02_g4_*).
template <int NI, bool F16IN, bool THREE>
__device__ __forceinline__ void tg_build(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t m1, uint32_t m2, const uint8_t* gp,
                                         uint32_t gstep, uint32_t mfirst, uint32_t mlast, const Coef8 cf, bool act, Raw8 ra0, Raw8 ra1) {
  Raw8 rb0, rb1;
  const uint8_t* glast = gp + (uint32_t)(NI - 1) * gstep - (mlast ? 0u : gstep);     // a halo item reads its (valid) neighbour
#define B200_TG_PTR(J) ((J) == NI - 1 ? glast : gp + (uint32_t)(J) * gstep)
#define B200_TG_EMIT(J, R)                                                                             \
  do {                                                                                                 \
    uint4 v_ = apply8<F16IN>(R, cf, act);                                                              \
    if ((J) == 0) { v_.x &= mfirst; v_.y &= mfirst; v_.z &= mfirst; v_.w &= mfirst; }                  \
    if ((J) == NI - 1) { v_.x &= mlast; v_.y &= mlast; v_.z &= mlast; v_.w &= mlast; }                 \
    sts128(a0 + 4096u * (J), v_.x, v_.y, v_.z, v_.w);                                                  \
    if (THREE) {                                                                                       \
      sts128(a1 + 4096u * (J), v_.x & m1, v_.y & m1, v_.z & m1, v_.w & m1);                            \
      sts128(a2 + 4096u * (J), v_.x & m2, v_.y & m2, v_.z & m2, v_.w & m2);                            \
    }                                                                                                  \
  } while (0)
#pragma unroll
  for (int j = 0; j < NI; j += 4) {
    if (j + 2 < NI) rb0 = load_raw8<F16IN>(B200_TG_PTR(j + 2));
    if (j + 3 < NI) rb1 = load_raw8<F16IN>(B200_TG_PTR(j + 3));
    B200_TG_EMIT(j, ra0);
    if (j + 1 < NI) B200_TG_EMIT(j + 1, ra1);
    if (j + 4 < NI) ra0 = load_raw8<F16IN>(B200_TG_PTR(j + 4));
    if (j + 5 < NI) ra1 = load_raw8<F16IN>(B200_TG_PTR(j + 5));
    if (j + 2 < NI) B200_TG_EMIT(j + 2, rb0);
    if (j + 3 < NI) B200_TG_EMIT(j + 3, rb1);
  }
#undef B200_TG_PTR
#undef B200_TG_EMIT
}
// Pull the (R+2) x W patch of ALL channels of one source towards L2 (fire-and-forget, no registers): issued for the
// NEXT tile while the current one is transformed, so build_copy0's loads pay L2 latency instead of HBM latency.
__device__ __forceinline__ void prefetch_patch(const uint8_t* src_img, int C, int esz, int ih0, int H, int W, int rows_img, int tid) {
  const int row_bytes = W * C * esz;                         // one image row, contiguous in NHWC
  for (int hh = 0; hh < rows_img; ++hh) {
    const int ih = ih0 + hh;
    if (ih < 0 || ih >= H) continue;
    const uint8_t* base = src_img + (long long)ih * row_bytes;
    for (int off = tid * 128; off < row_bytes; off += 32 * TG_NTW * 128) prefetch_l2(base + off);
  }
}
template <bool W32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TG_THREADS, 1) conv_gn2_kernel(const __grid_constant__ TcgParams p) {
  using L = SmemG;
  constexpr int BN = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* wfull = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);   // leader only
  uint64_t* wempty = wfull + TG_WS;
  uint64_t* tfull = wempty + TG_WS;                                      // leader only
  uint64_t* tempty = tfull + TG_TS;
  uint64_t* tmem_full = tempty + TG_TS;
  uint64_t* tmem_empty = tmem_full + 2;                                  // leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TG_WS; ++s) { mbar_init(&wfull[s], 2); mbar_init(&wempty[s], 1); }
    for (int s = 0; s < TG_TS; ++s) { mbar_init(&tfull[s], 2 * TG_NTW); mbar_init(&tempty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait(); pdl_trigger();   // prologue done; nothing above touched global memory (common.cuh)

  const int HW = p.H * p.W;
  const long long pairs_m = (p.tiles_m + 1) / 2;
  const long long total_pairs = pairs_m * p.tiles_n;
  const long long cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int chunks33 = p.kch[0] + p.kch[1], chunks11 = p.kch[2] + p.kch[3];

  // Register budget per warpgroup.  The CTA is launched with 640 threads x 96 registers; setmaxnreg.inc can only take what
  // setmaxnreg.dec of the same CTA released, so the two must balance: the producer / issuer / allocator warpgroup drops to
  // 32 (frees 128 x 64 = 8192), the two epilogue warpgroups rise to 120 (their block routine holds two 32-register tiles;
  // 2 x 128 x 24 = 6144) and the two transform warpgroups to 104 (2 x 128 x 8 = 2048).
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
  if (warp == 0 && lane == 0) {
    // ======================= weight TMA producer (both CTAs) =======================
    uint32_t stage = 0, phase = 0;
    for (long long pair = cid; pair < total_pairs; pair += nclusters) {
      const int nt = (int)(pair % p.tiles_n);
      const int wrow0 = nt * BN + (int)rank * (BN / 2);
      for (int src = 0; src < 4; ++src) {
        const int nch = p.kch[src];
        const CUtensorMap* tmW = src < 2 ? &p.tmW : &p.tmW2;
        const int wcol0 = src == 1 ? p.srcC[0] : src == 3 ? p.srcC[2] : 0;
        const int ntaps = src < 2 ? 9 : 1;
        for (int kc = 0; kc < nch; ++kc) {
          for (int t = 0; t < ntaps; ++t) {
            // consumption order of the operand copies: dw = 0, -1, +1; inside a copy dh = -1, 0, +1
            const int d = t / 3, dh = t % 3;
            const int dwi = d == 0 ? 1 : d == 1 ? 0 : 2;               // column of the 3x3 filter
            const int tap = src < 2 ? dh * 3 + dwi : 0;
            mbar_wait(&wempty[stage], phase ^ 1);
            const uint32_t lead_full = map_to_cta(smem_u32(&wfull[stage]), 0);
            if (leader) mbar_expect_tx(&wfull[stage], 2 * TG_W_BYTES);
            tma2_load_2d(tmW, smem + L::W_OFFSET + stage * TG_W_BYTES, lead_full, wcol0 + kc * 64, wrow0 + tap * p.N_total);
            if (!leader) mbar_arrive_cluster(lead_full);
            if (++stage == TG_WS) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ======================= MMA issuer (leader CTA, for the pair) =======================
    const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // kind::f16, M = 256
    uint32_t ws = 0, wphase = 0, ts = 0, tphase = 0, acc = 0, acc_phase = 0;
    const uint32_t t_base = smem_u32(smem), w_base = smem_u32(smem + L::W_OFFSET);
    for (long long pair = cid; pair < total_pairs; pair += nclusters) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      uint32_t first = 1;
      for (int c = 0; c < chunks33; ++c) {
        for (int d = 0; d < 3; ++d) {
          mbar_wait(&tfull[ts], tphase);
          tc_fence_after();
          for (int dh = 0; dh < 3; ++dh) {
            mbar_wait(&wfull[ws], wphase);
            tc_fence_after();
            const uint64_t adesc = make_smem_desc(t_base + ts * TG_SLOT_BYTES + dh * p.W * 128);
            const uint64_t bdesc = make_smem_desc(w_base + ws * TG_W_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma2_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
            first = 0;
            umma2_commit_mc(&wempty[ws]);
            if (++ws == TG_WS) { ws = 0; wphase ^= 1; }
          }
          umma2_commit_mc(&tempty[ts]);
          if (++ts == TG_TS) { ts = 0; tphase ^= 1; }
        }
      }
      for (int c = 0; c < chunks11; ++c) {
        mbar_wait(&tfull[ts], tphase);
        tc_fence_after();
        mbar_wait(&wfull[ws], wphase);
        tc_fence_after();
        const uint64_t adesc = make_smem_desc(t_base + ts * TG_SLOT_BYTES);
        const uint64_t bdesc = make_smem_desc(w_base + ws * TG_W_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma2_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
        first = 0;
        umma2_commit_mc(&wempty[ws]);
        if (++ws == TG_WS) { ws = 0; wphase ^= 1; }
        umma2_commit_mc(&tempty[ts]);
        if (++ts == TG_TS) { ts = 0; tphase ^= 1; }
      }
      umma2_commit_mc(&tmem_full[acc]);
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
  }
  } else if (warp < 12) {
    // ======================= epilogue (both CTAs, own 128 rows) =======================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
    const int q = (warp - 4) & 3, half = (warp - 4) >> 2;
    const Epilogue& e = p.epi;
    uint32_t acc = 0, acc_phase = 0;
    for (long long pair = cid; pair < total_pairs; pair += nclusters) {
      const int nt = (int)(pair % p.tiles_n);
      const long long mg = (pair / p.tiles_n) * 2 + rank;
      const long long row0 = mg * BM + q * 32;
      const int rows_valid = (mg < p.tiles_m) ? 32 : 0;                    // tiles are whole (HW % 128 == 0)
      const long long gm0 = row0;
      const int img0 = rows_valid > 0 ? (int)(gm0 / e.rows_per_img) : 0;
      if (e.residual && lane < rows_valid) {
#pragma unroll
        for (int i = 0; i < BN / 2; i += 32) prefetch_l2(e.residual + (gm0 + lane) * e.ld_res + nt * BN + half * (BN / 2) + i);
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int j = half * (BN / 64); j < (half + 1) * (BN / 64); ++j) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + j * 32, v);
        row_chunk_t_dispatch(v, smem + L::TRN_OFFSET + (warp - 4) * 4096, e, p.qstats, p.N_total, gm0, rows_valid,
                             nt * BN + j * 32, img0, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&tmem_empty[acc]), 0));
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ======================= transform (both CTAs, own 128 pixels; 8 warps = 256 threads) =======================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int tid = threadIdx.x - 384, o = tid & 7, pl = tid >> 3;      // channel octet, patch pixel lane (0..31)
    const uint32_t t_base = smem_u32(smem);
    uint32_t ts = 0, tphase = 0;
    constexpr int W = W32 ? 32 : 16, NI3 = W32 ? 6 : 5;                  // (R + 2) * W = 32 * NI3 patch pixels
    const int w = pl & (W - 1);
    // per-thread constants of the three copies (see tg_build): row of item 0 and swizzle phase in each copy
    const uint32_t off0 = (uint32_t)pl * 128 + (uint32_t)((o ^ (pl & 7)) << 4);
    const bool z1 = w == W - 1, z2 = w == 0;
    const uint32_t m1 = z1 ? 0u : 0xffffffffu, m2 = z2 ? 0u : 0xffffffffu;
    const int r1 = z1 ? pl - (W - 1) : pl + 1, r2 = z2 ? pl + (W - 1) : pl - 1;
    const uint32_t off1 = (uint32_t)r1 * 128 + (uint32_t)((o ^ ((pl + 1) & 7)) << 4);
    const uint32_t off2 = (uint32_t)r2 * 128 + (uint32_t)((o ^ ((pl - 1) & 7)) << 4);
    auto arrive_full = [&](uint32_t slot_idx) {
      if (leader) mbar_arrive(&tfull[slot_idx]);
      else mbar_arrive_cluster(map_to_cta(smem_u32(&tfull[slot_idx]), 0));
    };
    const bool act = p.act != 0 && p.scale != nullptr;
    for (long long pair = cid; pair < total_pairs; pair += nclusters) {
      const long long mg = (pair / p.tiles_n) * 2 + rank;
      const bool valid = mg < p.tiles_m;
      const long long p0 = mg * BM;
      const int img = valid ? (int)(p0 / HW) : 0;
      const int h0 = valid ? (int)(p0 % HW) / W : 0;
      {   // the next tile's input patches -> L2 while this tile is being transformed
        const long long pn = pair + nclusters;
        const long long mgn = (pn / p.tiles_n) * 2 + rank;
        if (pn < total_pairs && mgn < p.tiles_m) {
          const long long pn0 = mgn * BM;
          const int imgn = (int)(pn0 / HW), hn = (int)(pn0 % HW) / W;
          for (int src = 0; src < 4; ++src) {
            if (p.kch[src] == 0) continue;
            const int esz = p.srcF16[src] ? 2 : 4;
            prefetch_patch(reinterpret_cast<const uint8_t*>(p.src[src]) + (long long)imgn * HW * p.srcC[src] * esz, p.srcC[src], esz,
                           src < 2 ? hn - 1 : hn, p.H, W, src < 2 ? p.R + 2 : p.R, tid);
          }
        }
      }
      // ---- 3x3 phase: GroupNorm (+SiLU) on load, three shifted copies per chunk, written straight from registers ----
      // only item 0 (top halo row) and item NI3-1 (bottom halo row) can lie outside the image; a tile past the end is all zero
      const uint32_t mfirst = (valid && !(h0 == 0 && pl < W)) ? 0xffffffffu : 0u;
      const uint32_t mlast = (valid && !(h0 + p.R == p.H && pl >= 32 - W)) ? 0xffffffffu : 0u;
      const uint32_t mmid = valid ? 0xffffffffu : 0u;
      const int pix33 = (h0 - 1) * W + pl;                               // image pixel index of this thread's item 0 (may be < 0)
      for (int src = 0; src < 2; ++src) {
        const int nch = p.kch[src];
        if (nch == 0) continue;
        const int C = p.srcC[src];
        const bool f16in = p.srcF16[src] != 0;
        const uint32_t esz = f16in ? 2 : 4, pixb = (uint32_t)C * esz, gstep = 32 * pixb;
        const uint8_t* gbase = reinterpret_cast<const uint8_t*>(p.src[src]) + ((long long)img * HW + pix33) * (long long)pixb + (uint32_t)(o * 8) * esz;
        const int cg0 = (src == 1 ? p.srcC[0] : 0) + o * 8;
        for (int kc = 0; kc < nch; ++kc) {
          const uint8_t* gp = gbase + (uint32_t)(kc * 64) * esz;
          Coef8 cf;
          cf.s0 = cf.s1 = make_float4(1.f, 1.f, 1.f, 1.f); cf.h0 = cf.h1 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.scale) {
            const float4* sp = reinterpret_cast<const float4*>(p.scale + (long long)img * p.Cgn + cg0 + kc * 64);
            const float4* hp = reinterpret_cast<const float4*>(p.shift + (long long)img * p.Cgn + cg0 + kc * 64);
            cf.s0 = __ldg(sp); cf.s1 = __ldg(sp + 1); cf.h0 = __ldg(hp); cf.h1 = __ldg(hp + 1);
          }
          if (!valid) { cf.s0 = cf.s1 = cf.h0 = cf.h1 = make_float4(0.f, 0.f, 0.f, 0.f); }   // tile past the end: SiLU(0) = 0 everywhere
          // first two items' loads go out before the slot waits (item 0 reads item 1's address when it is a halo row)
          Raw8 ra0, ra1;
          const uint8_t* g0 = mfirst ? gp : gp + gstep;
          if (f16in) { ra0 = load_raw8<true>(g0); ra1 = load_raw8<true>(gp + gstep); }
          else { ra0 = load_raw8<false>(g0); ra1 = load_raw8<false>(gp + gstep); }
          const uint32_t s0 = ts, ph0 = tphase;
          const uint32_t s1 = s0 + 1 == TG_TS ? 0 : s0 + 1, ph1 = s0 + 1 == TG_TS ? ph0 ^ 1 : ph0;
          const uint32_t s2 = s1 + 1 == TG_TS ? 0 : s1 + 1, ph2 = s1 + 1 == TG_TS ? ph1 ^ 1 : ph1;
          mbar_wait(&tempty[s0], ph0 ^ 1); mbar_wait(&tempty[s1], ph1 ^ 1); mbar_wait(&tempty[s2], ph2 ^ 1);
          const uint32_t a0 = t_base + s0 * TG_SLOT_BYTES + off0, a1 = t_base + s1 * TG_SLOT_BYTES + off1, a2 = t_base + s2 * TG_SLOT_BYTES + off2;
          if (f16in) tg_build<NI3, true, true>(a0, a1, a2, m1, m2, gp, gstep, mfirst, mlast, cf, act, ra0, ra1);
          else tg_build<NI3, false, true>(a0, a1, a2, m1, m2, gp, gstep, mfirst, mlast, cf, act, ra0, ra1);
          // this warp's generic-proxy stores -> visible to the async proxy (UMMA), then one arrival per copy
          fence_async_smem();
          __syncwarp();
          if (lane == 0) { arrive_full(s0); arrive_full(s1); arrive_full(s2); }
          ts = s2 + 1 == TG_TS ? 0 : s2 + 1; tphase = s2 + 1 == TG_TS ? ph2 ^ 1 : ph2;
        }
      }
      // ---- extra 1x1 phase: identity transform (fp32 -> fp16), one un-shifted 128-row copy per chunk ----
      // A chunk is only four MMAs (512 tensor cycles) but 32 KB of fp32 input per CTA, so this phase runs at load latency
      // unless the loads are far ahead: a whole chunk (4 items per thread) is requested while the previous one is converted
      // and stored (measured with a two-item prefetch: the skip-projection convolutions at 0.69 PFLOP/s, profiles/r02_g5_*).
      const int pix11 = h0 * W + pl;
      for (int src = 2; src < 4; ++src) {
        const int nch = p.kch[src];
        if (nch == 0) continue;
        const int C = p.srcC[src];
        const bool f16in = p.srcF16[src] != 0;
        const uint32_t esz = f16in ? 2 : 4, pixb = (uint32_t)C * esz, gstep = 32 * pixb;
        const uint8_t* gbase = reinterpret_cast<const uint8_t*>(p.src[src]) + ((long long)img * HW + pix11) * (long long)pixb + (uint32_t)(o * 8) * esz;
        Coef8 ident;
        ident.s0 = ident.s1 = valid ? make_float4(1.f, 1.f, 1.f, 1.f) : make_float4(0.f, 0.f, 0.f, 0.f);
        ident.h0 = ident.h1 = make_float4(0.f, 0.f, 0.f, 0.f);
        Raw8 q0, q1, q2, q3, n0, n1, n2, n3;
#define B200_TG_LOAD4(G, A, B, C_, D)                                                                                   \
        do { if (f16in) { A = load_raw8<true>(G); B = load_raw8<true>((G) + gstep); C_ = load_raw8<true>((G) + 2 * gstep); D = load_raw8<true>((G) + 3 * gstep); } \
             else { A = load_raw8<false>(G); B = load_raw8<false>((G) + gstep); C_ = load_raw8<false>((G) + 2 * gstep); D = load_raw8<false>((G) + 3 * gstep); } } while (0)
#define B200_TG_STORE1(J, R)                                                                                            \
        do { const uint4 v_ = f16in ? apply8<true>(R, ident, false) : apply8<false>(R, ident, false);                  \
             sts128(a0 + 4096u * (J), v_.x, v_.y, v_.z, v_.w); } while (0)
        B200_TG_LOAD4(gbase, q0, q1, q2, q3);
        for (int kc = 0; kc < nch; ++kc) {
          if (kc + 1 < nch) { const uint8_t* gn = gbase + (uint32_t)((kc + 1) * 64) * esz; B200_TG_LOAD4(gn, n0, n1, n2, n3); }
          mbar_wait(&tempty[ts], tphase ^ 1);
          const uint32_t a0 = t_base + ts * TG_SLOT_BYTES + off0;
          B200_TG_STORE1(0, q0); B200_TG_STORE1(1, q1); B200_TG_STORE1(2, q2); B200_TG_STORE1(3, q3);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) arrive_full(ts);
          if (++ts == TG_TS) { ts = 0; tphase ^= 1; }
          q0 = n0; q1 = n1; q2 = n2; q3 = n3;
        }
#undef B200_TG_LOAD4
#undef B200_TG_STORE1
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
}
