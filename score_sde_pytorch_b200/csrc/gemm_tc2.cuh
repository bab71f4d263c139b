// Two-CTA (cta_group::2) variant of the tcgen05 implicit GEMM: a cluster of two CTAs on one TPC computes a
// 256-row x 256-column output tile.  Included by gemm_tc.cu inside its anonymous namespace (it reuses the PTX
// wrappers, TcParams, descriptors and the GroupNorm-sum helper defined there).
//
// Why: with one CTA per tile every K step writes 48 KB into shared memory (TMA) and reads 48 KB back (UMMA)
// per 512 tensor-pipe cycles = 192 B/clk against the SM's ~128 B/clk shared-memory port, which caps the
// tensor pipe near 67 % (measured 70 %, profiles/r01_*).  In a CTA pair each CTA stages its own 128 rows of A
// and only HALF of the W tile (128 of the 256 output channels); the pair's tensor cores share the two W halves,
// so a CTA writes 32 KB and reads 32 KB per K step = 128 B/clk.
//
// Roles per CTA (384 threads): warp 0 TMA producer (both CTAs, one elect.sync lane issuing; transaction bytes are
// accounted on the leader's `full` barrier), warp 1 of the LEADER (one elected lane) issues tcgen05.mma.cta_group::2 for the pair and
// multicasts its commits to both CTAs' `empty` / `tmem_full` barriers, warp 2 allocates TMEM (cta_group::2),
// warps 4..11 of each CTA drain their own 128 TMEM lanes (direct 128-bit stores, fused bias / time-embedding /
// residual / scale / TF32 rounding / GroupNorm quad sums) and release the accumulator stage on the leader's
// `tmem_empty` barrier (16 arrivals).

template <int BN, int STAGES>
struct Smem2 {
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + (BN / 2) * BKE * 4;   // A rows (16 KB) + this CTA's half of the W tile
  static constexpr int TRN_OFFSET = STAGES * STAGE_BYTES;                  // 8 epilogue warps x 4 KB transposition scratch
  static constexpr int BAR_OFFSET = TRN_OFFSET + 8 * 4096;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
  static_assert(TOTAL <= 232448, "exceeds the 227 KB shared-memory limit of sm_100");
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank)); return r;
}
// Remote arrive.  NOT `.release.cluster`: that form compiles to MEMBAR.ALL.GPU + ERRBAR in front of the arrive
// (measured: ~1600 cycles per K step in the peer's producer, halving the kernel's throughput — profiles/r01_c7).
// The arrive only has to count; the data it guards is delivered by TMA complete_tx on the same barrier.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(const CUtensorMap* tm, void* dst, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* tm, void* dst, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// commit: arrive (once) on the barrier at this shared::cta offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1) gemm_tc2_kernel(const __grid_constant__ TcParams p) {
  using L = Smem2<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);   // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;                                     // used in the leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* wfull_bar = tmem_empty + 3;                                     // halo form: barriers of the weight-half ring
  uint64_t* wempty_bar = wfull_bar + HALO2_WS;
  static_assert((2 * STAGES + 4 + 1 + 2 * HALO2_WS) * 8 <= 256 && STAGES >= HALO_XS, "barrier block");

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp index provably warp-uniform
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }   // full: one arrive per CTA
    for (int s = 0; s < HALO2_WS; ++s) { mbar_init(&wfull_bar[s], 2); mbar_init(&wempty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 16); }     // empty: 8 warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // both CTAs' barriers and TMEM exist before any cross-CTA traffic
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait(); pdl_trigger();   // prologue done; nothing above touched global memory (common.cuh)

  const int kiters = (p.kchunks1 + p.kchunks2) * p.taps + p.kchunks3 + p.kchunks4;
  const int HW = p.H * p.W;
  const long long tiles_m_total = (long long)p.nbatch * p.tiles_m_per_batch;
  const long long pairs_m = (tiles_m_total + 1) / 2;
  const long long total_pairs = pairs_m * p.tiles_n;
  const long long cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp == 0) {
    // ======================= TMA producer (both CTAs; whole warp, one elected lane issues - see elect_one) =======================
    const bool issue = elect_one();
    uint32_t stage = 0, phase = 0;
    if (p.halo) {
      // halo form (see gemm_tc_kernel): per channel chunk three halo copies of this CTA's 128 pixels (whole rows of one
      // image), each followed by this CTA's half of the three weight slices of that filter column
      uint32_t ws = 0, wphase = 0;
      uint8_t* const wring = smem + HALO_XS * HALO2_X_BYTES;
      for (long long pair = cid; pair < total_pairs; pair += nclusters) {
        const int nt = (int)(pair % p.tiles_n);
        const long long mg = (pair / p.tiles_n) * 2 + rank;
        const long long p0 = mg * BM;
        const int img0 = mg >= tiles_m_total ? (1 << 28) : (int)(p0 / HW);      // past the end -> TMA zero fill
        const int h0 = (int)(p0 % HW) / p.W;
        const int wrow0 = nt * BN + (int)rank * (BN / 2);
        if (issue && p.halo_prefetch && pair + nclusters < total_pairs && (pair + nclusters) / p.tiles_n != pair / p.tiles_n) {
          const long long qg = ((pair + nclusters) / p.tiles_n) * 2 + rank;               // this CTA's next 128 pixels, every chunk
          if (qg < tiles_m_total) {
            const int qi = (int)(qg * BM / HW), qh = (int)(qg * BM % HW) / p.W;
            for (int kc = 0; kc < p.kchunks1; ++kc) tma_prefetch_4d(&p.tmH1, kc * p.bke, 0, qh - 1, qi);
            for (int kc = 0; kc < p.kchunks2; ++kc) tma_prefetch_4d(&p.tmH2, kc * p.bke, 0, qh - 1, qi);
          }
        }
        for (int src = 0; src < 4; ++src) {
          const int nch = src == 0 ? p.kchunks1 : src == 1 ? p.kchunks2 : src == 2 ? p.kchunks3 : p.kchunks4;
          if (nch == 0) continue;
          const int wcol0 = src == 1 ? p.C1 : src == 3 ? p.C3 : 0;
          for (int kc = 0; kc < nch; ++kc) {
            const int ncopy = src < 2 ? 3 : 1;
            for (int dwi = 0; dwi < ncopy; ++dwi) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint32_t lead = map_to_cta(smem_u32(&full_bar[stage]), 0);
              if (issue) {
                if (src < 2) {
                  if (leader) mbar_expect_tx(&full_bar[stage], 2 * (uint32_t)p.halo_copy_bytes);
                  tma2_load_4d(src == 0 ? &p.tmH1 : &p.tmH2, smem + stage * HALO2_X_BYTES, lead, kc * p.bke, dwi - 1, h0 - 1, img0);
                } else {
                  if (leader) mbar_expect_tx(&full_bar[stage], 2 * A_STAGE_BYTES);
                  tma2_load_4d(src == 2 ? &p.tmA3 : &p.tmA4, smem + stage * HALO2_X_BYTES, lead, kc * p.bke, 0, h0, img0);
                }
                if (!leader) mbar_arrive_cluster(lead);
              }
              if (++stage == HALO_XS) { stage = 0; phase ^= 1; }
              const int ntap = src < 2 ? 3 : 1;
              for (int dhi = 0; dhi < ntap; ++dhi) {
                mbar_wait(&wempty_bar[ws], wphase ^ 1);
                lead = map_to_cta(smem_u32(&wfull_bar[ws]), 0);
                if (issue) {
                  if (leader) mbar_expect_tx(&wfull_bar[ws], 2 * A_STAGE_BYTES);
                  if (src < 2) tma2_load_2d(&p.tmW, wring + ws * A_STAGE_BYTES, lead, wcol0 + kc * p.bke, wrow0 + (dhi * 3 + dwi) * p.N_total);
                  else tma2_load_2d(&p.tmW2, wring + ws * A_STAGE_BYTES, lead, wcol0 + kc * p.bke, wrow0);
                  if (!leader) mbar_arrive_cluster(lead);
                }
                if (++ws == HALO2_WS) { ws = 0; wphase ^= 1; }
              }
            }
          }
        }
      }
    } else
    for (long long pair = cid; pair < total_pairs; pair += nclusters) {
      const int nt = (int)(pair % p.tiles_n);
      const long long mg = (pair / p.tiles_n) * 2 + rank;      // this CTA's 128-row tile (may be one past the end)
      const int b = (int)(mg / p.tiles_m_per_batch);
      const int mt = (int)(mg % p.tiles_m_per_batch);
      int img0 = 0, h0 = 0, w0 = 0;
      if (p.conv) {
        const long long p0 = (long long)mt * BM;
        img0 = (int)(p0 / HW);
        const int rem = (int)(p0 % HW);
        h0 = rem / p.W; w0 = rem % p.W;
        if (mg >= tiles_m_total) img0 = 1 << 28;               // fully out of bounds -> TMA zero fill
      }
      const int arow0 = (mg >= tiles_m_total) ? (1 << 30) : b * p.a_batch_rows + mt * BM;
      const int bsh = (int)((pair / p.tiles_n) * 2 / p.tiles_m_per_batch);   // batch of the pair (both CTAs share it)
      const int wrow0 = bsh * p.w_batch_rows + nt * BN + (int)rank * (BN / 2);   // this CTA's half of the W tile
      for (int src = 0; src < 4; ++src) {                        // 2,3: the extra 1x1 phase (see gemm_tc_kernel)
        const int nch = src == 0 ? p.kchunks1 : src == 1 ? p.kchunks2 : src == 2 ? p.kchunks3 : p.kchunks4;
        if (nch == 0) continue;
        const CUtensorMap* tmA = src == 0 ? &p.tmA1 : src == 1 ? &p.tmA2 : src == 2 ? &p.tmA3 : &p.tmA4;
        const CUtensorMap* tmW = src < 2 ? &p.tmW : &p.tmW2;
        const int wcol0 = src == 1 ? p.C1 : src == 3 ? p.C3 : 0;
        const int ntaps = src < 2 ? p.taps : 1;
        for (int tap = 0; tap < ntaps; ++tap) {                 // K order: filter tap, then channel chunk (see gemm_tc_kernel)
          const int dh = src < 2 ? tap / p.S - p.pad : 0, dw = src < 2 ? tap % p.S - p.pad : 0;
          for (int kc = 0; kc < nch; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);            // own smem slot released by the pair's MMA commit
            uint8_t* sa = smem + stage * L::STAGE_BYTES;
            uint8_t* sb = sa + A_STAGE_BYTES;
            const uint32_t lead_full = map_to_cta(smem_u32(&full_bar[stage]), 0);
            if (issue) {
              if (leader) mbar_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);     // bytes of BOTH CTAs land on this barrier
              if (p.conv) tma2_load_4d(tmA, sa, lead_full, kc * p.bke, w0 * p.stride + dw, h0 * p.stride + dh, img0);
              else tma2_load_4d(tmA, sa, lead_full, kc * p.bke, arow0, 0, 0);
              tma2_load_2d(tmW, sb, lead_full, wcol0 + kc * p.bke, wrow0 + tap * p.N_total);
              if (!leader) mbar_arrive_cluster(lead_full);                          // second of the barrier's two arrivals
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ======================= MMA issuer (leader CTA, for the pair; whole warp, one elected lane issues) =======================
    const bool issue = elect_one();
    // instruction: M = 256 (128 rows per CTA), N = BN, K = 8 tf32; D fp32 in each CTA's own TMEM
    const bool f16 = p.f16 != 0;
    const uint32_t idesc = (1u << 4) | (f16 ? 0u : ((2u << 7) | (2u << 10))) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    if (p.halo) {
      // halo form: A = 128 pixel rows of the halo copy starting one filter row further in per dh (same offset in both
      // CTAs), B = the pair's two weight halves; slices are released per tap, a copy after its three taps
      uint32_t ws = 0, wphase = 0;
      const uint32_t wring = smem_u32(smem + HALO_XS * HALO2_X_BYTES);
      for (long long pair = cid; pair < total_pairs; pair += nclusters) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        int it = 0;
        for (int src = 0; src < 4; ++src) {
          const int nch = src == 0 ? p.kchunks1 : src == 1 ? p.kchunks2 : src == 2 ? p.kchunks3 : p.kchunks4;
          const int ncopies = src < 2 ? 3 * nch : nch, ntap = src < 2 ? 3 : 1;
          for (int c = 0; c < ncopies; ++c) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sx = smem_u32(smem + stage * HALO2_X_BYTES);
            for (int dhi = 0; dhi < ntap; ++dhi, ++it) {
              mbar_wait(&wfull_bar[ws], wphase);
              tc_fence_after();
              const uint64_t adesc = make_smem_desc(sx + (src < 2 ? dhi * p.halo_dh_bytes : 0));
              const uint64_t bdesc = make_smem_desc(wring + ws * A_STAGE_BYTES);
              if (issue) {
                if (f16) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) umma2_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
                } else {
#pragma unroll
                  for (int k = 0; k < 4; ++k) umma2_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
                }
                umma2_commit_mc(&wempty_bar[ws]);
              }
              if (++ws == HALO2_WS) { ws = 0; wphase ^= 1; }
            }
            if (issue) umma2_commit_mc(&empty_bar[stage]);
            if (++stage == HALO_XS) { stage = 0; phase ^= 1; }
          }
        }
        if (issue) umma2_commit_mc(&tmem_full[acc]);
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
      }
    } else
    for (long long pair = cid; pair < total_pairs; pair += nclusters) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int it = 0; it < kiters; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
        const uint64_t adesc = make_smem_desc(sa);
        const uint64_t bdesc = make_smem_desc(sa + A_STAGE_BYTES);
        if (issue) {
          if (f16) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma2_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma2_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
          }
          umma2_commit_mc(&empty_bar[stage]);                   // frees the slot in both CTAs
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (issue) umma2_commit_mc(&tmem_full[acc]);              // accumulator ready in both CTAs
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ======================= epilogue (both CTAs, own 128 rows) =======================
    const int q = (warp - 4) & 3, half = (warp - 4) >> 2;   // two warps per TMEM lane quarter, half the columns each
    const Epilogue& e = p.epi;
    uint32_t acc = 0, acc_phase = 0;
    for (long long pair = cid; pair < total_pairs; pair += nclusters) {
      const int nt = (int)(pair % p.tiles_n);
      const long long mg = (pair / p.tiles_n) * 2 + rank;
      const int b = (int)(mg / p.tiles_m_per_batch);
      const int mt = (int)(mg % p.tiles_m_per_batch);
      const int row0 = mt * BM + q * 32;                                   // first row of this warp's 32-row blocks
      const int rows_valid = (mg < tiles_m_total) ? min(32, max(0, p.M_per_batch - row0)) : 0;
      const long long gm0 = (long long)b * p.M_per_batch + row0;
      const int img0 = rows_valid > 0 ? (int)(gm0 / e.rows_per_img) : 0;
      if (e.residual && lane < rows_valid) {   // residual rows -> L2 while the accumulator is still being produced
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
      if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&tmem_empty[acc]), 0));   // leader's barrier, from either CTA
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // the peer's smem / barriers / TMEM stay valid until both are done
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
}
