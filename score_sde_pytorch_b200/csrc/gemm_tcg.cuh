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
// Roles per CTA (512 threads): warp 0 lane 0 weight TMA producer (both CTAs), warp 1 lane 0 of the leader issues
// tcgen05.mma.cta_group::2 for the pair, warp 2 TMEM allocation, warps 4..11 epilogue (identical to gemm_tc2_kernel),
// warps 12..15 transform.  Barriers: wfull/wempty (weight ring), tfull/tempty (ring of operand copies; tfull lives in
// the leader and counts one arrival per transform warp of BOTH CTAs), tmem_full/tmem_empty.

constexpr int TG_TS = 5;                    // operand-copy slots (three per chunk in flight + two being built)
constexpr int TG_WS = 4;                    // weight ring stages
constexpr int TG_SLOT_BYTES = 192 * 128;    // (R+2)*W <= 192 pixel rows of 64 fp16 channels
constexpr int TG_W_BYTES = 128 * 128;       // this CTA's 128 of the 256 output channels x 64 k
constexpr int TG_NTW = 4;                   // transform warps per CTA
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
__device__ __forceinline__ void transform_bar() { asm volatile("bar.sync 1, %0;" ::"n"(32 * TG_NTW) : "memory"); }
__device__ __forceinline__ float silu_approx(float x) { return __fdividef(x, 1.0f + __expf(-x)); }   // == silu_fast (elementwise.cu)

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
template <bool F16IN>
__device__ __forceinline__ uint4 apply8(const Raw8& r, const float (&sc)[8], const float (&sh)[8], bool act) {
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

// Build "copy 0" of one chunk: `rows` pixel rows starting at patch row 0 (image row ih0 = first image row of the
// patch, may be -1; rows outside [0, H) are zero).  Thread `tid` (0..127) owns channel octet tid & 7 of pixel rows
// (tid >> 3) + 16 j.  Software-pipelined in batches of four pixels per thread: the loads of batch b+1 are in flight
// while batch b is transformed and stored (the first version loaded, transformed and stored batch by batch and was
// bound by global-load latency: 16 KB in flight per SM for a third of the time, fused convolutions 2.3x slower than the
// unfused pair, profiles/r02_g1_*).
template <bool F16IN>
__device__ __forceinline__ void build_copy0(uint32_t slot, const uint8_t* src_img, int C, int c0, int ih0, int H, int W, int rows,
                                            const float (&sc)[8], const float (&sh)[8], bool act, int tid, bool tile_valid) {
  const int o = tid & 7;
  const int esz = F16IN ? 2 : 4;
  const int wshift = W == 32 ? 5 : 4;                     // W is 16 or 32
  const uint32_t cb = (uint32_t)c0 * esz, pixb = (uint32_t)C * esz;
  // item j of this thread: pixel row pp = (tid >> 3) + 16 j; two items per batch, two batches in flight
  const int nitems = (rows - (tid >> 3) + 15) >> 4;
  Raw8 a0, a1, b0, b1;
  bool va0 = false, va1 = false, vb0 = false, vb1 = false;
#define B200_TG_FETCH(J, R, V)                                                                   \
  do {                                                                                           \
    const int pp_ = (tid >> 3) + 16 * (J);                                                       \
    const int ih_ = ih0 + (pp_ >> wshift);                                                       \
    V = (J) < nitems && tile_valid && ih_ >= 0 && ih_ < H;                                       \
    if (V) R = load_raw8<F16IN>(src_img + ((uint32_t)(ih_ * W + (pp_ & (W - 1))) * pixb + cb));   \
  } while (0)
#define B200_TG_EMIT(J, R, V)                                                                    \
  do {                                                                                           \
    if ((J) < nitems) {                                                                          \
      const int pp_ = (tid >> 3) + 16 * (J);                                                     \
      uint4 v_ = make_uint4(0u, 0u, 0u, 0u);                                                     \
      if (V) v_ = apply8<F16IN>(R, sc, sh, act);                                                \
      sts128(slot + pp_ * 128 + ((o ^ (pp_ & 7)) << 4), v_.x, v_.y, v_.z, v_.w);                 \
    }                                                                                            \
  } while (0)
  B200_TG_FETCH(0, a0, va0); B200_TG_FETCH(1, a1, va1);
#pragma unroll 1
  for (int j = 0; j < nitems; j += 4) {
    B200_TG_FETCH(j + 2, b0, vb0); B200_TG_FETCH(j + 3, b1, vb1);
    B200_TG_EMIT(j, a0, va0); B200_TG_EMIT(j + 1, a1, va1);
    B200_TG_FETCH(j + 4, a0, va0); B200_TG_FETCH(j + 5, a1, va1);
    B200_TG_EMIT(j + 2, b0, vb0); B200_TG_EMIT(j + 3, b1, vb1);
  }
#undef B200_TG_FETCH
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
// dst[hh][w] = copy0[hh][w + d] (zero where w + d leaves the row), d = -1 or +1
__device__ __forceinline__ void shift_copy(uint32_t dst, uint32_t src, int d, int W, int rows, int tid) {
  const int o = tid & 7;
  for (int pp = tid >> 3; pp < rows; pp += 16) {
    const int w = pp & (W - 1);                           // W is 16 or 32
    const int sp = pp + d;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if ((unsigned)(w + d) < (unsigned)W) v = lds128u(src + sp * 128 + ((o ^ (sp & 7)) << 4));
    sts128(dst + pp * 128 + ((o ^ (pp & 7)) << 4), v.x, v.y, v.z, v.w);
  }
}

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

  const int HW = p.H * p.W;
  const long long pairs_m = (p.tiles_m + 1) / 2;
  const long long total_pairs = pairs_m * p.tiles_n;
  const long long cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int chunks33 = p.kch[0] + p.kch[1], chunks11 = p.kch[2] + p.kch[3];

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
  } else if (warp >= 4 && warp < 12) {
    // ======================= epilogue (both CTAs, own 128 rows) =======================
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
  } else if (warp >= 12) {
    // ======================= transform (both CTAs, own 128 pixels) =======================
    const int tid = threadIdx.x - 384, o = tid & 7;
    const uint32_t t_base = smem_u32(smem);
    uint32_t ts = 0, tphase = 0;
    // one arrival per transform warp on the LEADER's tfull barrier, after this warp's generic-proxy stores have
    // been made visible to the async proxy (the tensor core reads the copy through it)
    auto publish = [&](uint32_t slot_idx) {
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tfull[slot_idx]);
        else mbar_arrive_cluster(map_to_cta(smem_u32(&tfull[slot_idx]), 0));
      }
    };
    auto next_slot = [&]() { if (++ts == TG_TS) { ts = 0; tphase ^= 1; } };
    const int rows33 = (p.R + 2) * p.W;
    for (long long pair = cid; pair < total_pairs; pair += nclusters) {
      const long long mg = (pair / p.tiles_n) * 2 + rank;
      const bool valid = mg < p.tiles_m;
      const long long p0 = mg * BM;
      const int img = valid ? (int)(p0 / HW) : 0;
      const int h0 = valid ? (int)(p0 % HW) / p.W : 0;
      {   // the next tile's input patches -> L2 while this tile is being transformed
        const long long pn = pair + nclusters;
        const long long mgn = (pn / p.tiles_n) * 2 + rank;
        if (pn < total_pairs && mgn < p.tiles_m) {
          const long long pn0 = mgn * BM;
          const int imgn = (int)(pn0 / HW), hn = (int)(pn0 % HW) / p.W;
          for (int src = 0; src < 4; ++src) {
            if (p.kch[src] == 0) continue;
            const int esz = p.srcF16[src] ? 2 : 4;
            prefetch_patch(reinterpret_cast<const uint8_t*>(p.src[src]) + (long long)imgn * HW * p.srcC[src] * esz, p.srcC[src], esz,
                           src < 2 ? hn - 1 : hn, p.H, p.W, src < 2 ? p.R + 2 : p.R, tid);
          }
        }
      }
      // ---- 3x3 phase: GroupNorm (+SiLU) on load, three shifted copies per chunk ----
      for (int src = 0; src < 2; ++src) {
        const int nch = p.kch[src];
        if (nch == 0) continue;
        const int C = p.srcC[src];
        const bool f16in = p.srcF16[src] != 0;
        const uint8_t* img_base = reinterpret_cast<const uint8_t*>(p.src[src]) + (long long)img * HW * C * (f16in ? 2 : 4);
        const int cg0 = src == 1 ? p.srcC[0] : 0;
        for (int kc = 0; kc < nch; ++kc) {
          const int c0 = kc * 64 + o * 8;
          float sc[8], sh[8];
          if (p.scale) {
            const float4* sp = reinterpret_cast<const float4*>(p.scale + (long long)img * p.Cgn + cg0 + c0);
            const float4* hp = reinterpret_cast<const float4*>(p.shift + (long long)img * p.Cgn + cg0 + c0);
            const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1), h0v = __ldg(hp), h1v = __ldg(hp + 1);
            sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
            sh[0] = h0v.x; sh[1] = h0v.y; sh[2] = h0v.z; sh[3] = h0v.w; sh[4] = h1v.x; sh[5] = h1v.y; sh[6] = h1v.z; sh[7] = h1v.w;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) { sc[i] = 1.f; sh[i] = 0.f; }
          }
          const uint32_t s0 = ts, s0addr = t_base + ts * TG_SLOT_BYTES;
          mbar_wait(&tempty[ts], tphase ^ 1);
          if (f16in) build_copy0<true>(s0addr, img_base, C, c0, h0 - 1, p.H, p.W, rows33, sc, sh, p.act != 0 && p.scale, tid, valid);
          else build_copy0<false>(s0addr, img_base, C, c0, h0 - 1, p.H, p.W, rows33, sc, sh, p.act != 0 && p.scale, tid, valid);
          next_slot();
          fence_async_smem();
          transform_bar();                       // copy 0 complete: neighbours' pixels are readable
          if (lane == 0) {
            if (leader) mbar_arrive(&tfull[s0]);
            else mbar_arrive_cluster(map_to_cta(smem_u32(&tfull[s0]), 0));
          }
#pragma unroll 1
          for (int d = -1; d <= 1; d += 2) {
            mbar_wait(&tempty[ts], tphase ^ 1);
            shift_copy(t_base + ts * TG_SLOT_BYTES, s0addr, d, p.W, rows33, tid);
            publish(ts);
            next_slot();
          }
          // copy 0 of this chunk is overwritten only TG_TS slots later, and every thread passes transform_bar()
          // of the next chunk first, so no thread can still be reading it then
        }
      }
      // ---- extra 1x1 phase: identity transform (fp32 -> fp16), one un-shifted 128-row copy per chunk ----
      // (the 1x1 chunks recycle slots without the per-chunk barrier of the 3x3 phase: make sure no transform thread is
      // still deriving a shifted copy from a copy 0 that is about to be overwritten)
      if (chunks11) transform_bar();
      for (int src = 2; src < 4; ++src) {
        const int nch = p.kch[src];
        if (nch == 0) continue;
        const int C = p.srcC[src];
        const bool f16in = p.srcF16[src] != 0;
        const uint8_t* img_base = reinterpret_cast<const uint8_t*>(p.src[src]) + (long long)img * HW * C * (f16in ? 2 : 4);
        const float one[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f}, zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int kc = 0; kc < nch; ++kc) {
          const int c0 = kc * 64 + o * 8;
          mbar_wait(&tempty[ts], tphase ^ 1);
          if (f16in) build_copy0<true>(t_base + ts * TG_SLOT_BYTES, img_base, C, c0, h0, p.H, p.W, BM, one, zero, false, tid, valid);
          else build_copy0<false>(t_base + ts * TG_SLOT_BYTES, img_base, C, c0, h0, p.H, p.W, BM, one, zero, false, tid, valid);
          publish(ts);
          next_slot();
        }
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
