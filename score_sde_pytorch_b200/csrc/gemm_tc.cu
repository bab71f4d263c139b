// tcgen05 / TMEM / TMA implicit GEMM for the NCSN++ contractions (3x3 and 1x1
// convolutions, NIN projections, the attention products) on sm_100a.
//
//   out[b][m, n] = epi( sum_{src, tap, c} A_src(b, m, tap, c) * W(b)[tap][n][koff_src + c]
//                       (+ an optional extra 1x1 phase: the fused skip projection) )
//
// Operands are written by their producers in the contraction's format: fp32 bit patterns
// rounded to the TF32 grid (kind::tf32, 32 channels per 128-byte K step) or IEEE fp16
// (kind::f16, 64 channels per K step); either way 11 significand bits, fp32 accumulation in TMEM.
//
// Kernels (persistent, warp-specialised, 384 threads unless noted):
//   gemm_tc_kernel<BN, STAGES, false>  one CTA per tile: 128 rows x BN columns, or - `swap` - 128 output
//                                      channels x 256 pixels (D^T = W X^T) for 128-channel convolutions,
//                                      1x1 convolutions and the network head;
//   gemm_tc2_kernel<BN, STAGES>        gemm_tc2.cuh: a CTA pair (cta_group::2) per 256-row tile, each CTA
//                                      stages its own A rows and half of the W tile;
//   attn_tc_kernel<F16>                attn_tc.cuh: QK^T, softmax, PV, NIN_3 + residual in one kernel;
//   gemm_tc_kernel<BN, STAGES, true>   256 threads: the smem-staged TMA-store epilogue kept for A/B runs.
// Roles: warp 0 TMA producer (4-D box of the NHWC tensor shifted by the filter tap - or, in the halo form, three
// W-shifted copies of the tile with its halo per channel chunk - zero halo and tail rows from out-of-bounds fill,
// SWIZZLE_128B, ring with full/empty mbarriers); warp 1 MMA issuer (four MMAs per K step into one of two TMEM
// accumulator stages, tcgen05.commit frees the slot / publishes the accumulator); both run as WHOLE warps whose
// loop state is warp-uniform, with the TMA / tcgen05 instructions predicated on one elect.sync lane (see elect_one:
// under a `lane == 0` branch every such instruction was wrapped in a lane-serialising loop and the issue loop became
// the bound of the kernel); warp 2 TMEM allocation; warps 4..11 epilogue, two per TMEM lane
// quarter: tcgen05.ld, then a block routine compiled per (residual, store format, stats) combination -
// `row_chunk_t` (row-major outputs: 32x32 blocks transposed through warp-private shared memory so global
// accesses are coalesced) or `swap_chunk` (lane = channel, already coalesced) - with bias, time-embedding
// row, residual, scale, fp32 / TF32 / fp16 store and the GroupNorm quad sums fused.
#include "kernels.h"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <mutex>
#include <cstdlib>
#include <cstring>

namespace b200 {

namespace {

constexpr int BM = 128;          // rows (pixels) per tile == UMMA M
constexpr int BKE = 32;          // fp32 elements per K step == one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BM * BKE * 4;   // 16 KiB

// Halo form of the 3x3 mainloop.  The nine filter taps of one channel chunk read three W-shifted
// copies of the tile *with its one-row halo* (TMA box [chunk, W, rows + 2], origin w = -1 / 0 / +1, zero fill outside the
// image); the operand of tap (dh, dw) is copy[dw] starting (dh + 1) * W pixels in - a multiple of 1024 bytes, so a plain
// swizzled descriptor.  L2 -> shared-memory traffic per chunk: 3 x (rows + 2) / rows tiles instead of 9.
constexpr int HALO_XS = 4;                  // halo copies in flight (both kernels)
constexpr int HALO_X_BYTES = 40 * 1024;     // single-CTA swapped form: (8 + 2) rows x 32 pixels x 128 B
constexpr int HALO_WS = 4;                  //   its weight-slice ring: 4 x 16 KB   (4 x 40 + 4 x 16 = 224 KB = the plain form's footprint)
constexpr int HALO2_X_BYTES = 24 * 1024;    // CTA pair: (4 + 2) rows x 32 pixels x 128 B per CTA
constexpr int HALO2_WS = 6;                 //   its weight-half ring: 6 x 16 KB   (4 x 24 + 6 x 16 = 192 KB = the plain form's stages)

struct TcParams {
  CUtensorMap tmA1, tmA2, tmW;
  CUtensorMap tmA3, tmA4, tmW2;            // optional extra 1x1 K phase (fused skip projection): out += [A3|A4] W2^T
  CUtensorMap tmH1, tmH2;                  // halo form: box = [bke channels, W, tile rows + 2, 1] of the two filter sources
  int halo;                                // 1: 3x3 taps read W-shifted halo copies of the tile (3 loads per channel chunk instead of 9)
  int halo_dh_bytes;                       // W * 128: bytes between the operand windows of consecutive filter rows inside a copy
  int halo_copy_bytes;                     // (tile rows + 2) * W * 128: bytes of one halo copy
  int halo_prefetch;                       // 1: the halo producers prefetch the next tile's halo boxes (all channel chunks) into L2 (opt-in: measured -2 %)
  int chunk_major;                         // nine-load loop of the single-CTA kernel walks K as (chunk, filter column, filter row): shapes with a halo form
  int conv, H, W, taps, pad, S, stride;   // H, W: OUTPUT spatial size; S = filter width (3 or 1)
  int kchunks1, kchunks2, C1;
  int kchunks3, kchunks4, C3;              // extra phase: channel chunks of its (two-source) input
  int N_total, tiles_n;
  int nbatch, tiles_m_per_batch, M_per_batch;
  int a_batch_rows, w_batch_rows;
  int f16;                                 // 1: A and W are fp16 (tcgen05 kind::f16, 64-channel K steps); 0: TF32-grid fp32 (32-channel K steps)
  int bke;                                 // channels per K step: one 128-byte swizzle row = 32 fp32 or 64 fp16
  int swap;                                // 1: operands swapped (D^T = W X^T): 128 output channels x 256 pixels per tile
  double* qstats;                          // optional [img][N_total/4][2] GroupNorm quad sums (sum, sum of squares)
  long long total_tiles;
  Epilogue epi;
};

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (launch error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) {   // ~4 s at 2 GHz
      printf("gemm_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* tm, void* dst, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, void* dst, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// fire-and-forget: pull a 4-D box towards L2 (UTMAPF); the halo producers use it for the NEXT tile's whole channel vector so
// that DRAM sees each pixel row once, contiguously, instead of one 128-byte chunk per K phase
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Pull one 128-byte line towards L2 without occupying a register: the epilogues use it for their residual rows
// while they still wait for the accumulator, so the later loads hit L2 instead of paying an HBM round trip
// inside the per-chunk critical path (the 112-register cap that lets GroupNorm CTAs co-reside leaves no room
// for a register-held prefetch).
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Explicit shared-space 128-bit accesses.  Through a generic pointer the compiler emitted LD.E / ST.E for the
// epilogue's transposition scratch (it cannot prove the address space of a pointer derived from the dynamic smem
// base), i.e. long-scoreboard loads that it then serialised one row at a time: the attention kernel's `final`
// phase spent most of its samples waiting on them (profiles/r01_c31_ncu_attn_tc_f16.md).
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// the four MMAs of one 128-byte K step: 4 x K=8 (tf32) or 4 x K=16 (f16); either way +32 B per slice
template <bool F16>
__device__ __forceinline__ void umma_kstep(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, int it) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    // advance 32 B along K inside the 128-B swizzle row: +2 in the (addr>>4) field
    if (F16) umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
    else umma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
  }
}
// One lane of a converged warp (elect.sync).  The MMA issuer runs as a whole warp with its loop counters, barrier
// addresses and descriptors warp-uniform, and only the tcgen05 instructions predicated on this: under a `lane == 0`
// branch the compiler cannot prove a single active lane and wraps every UTCHMMA / UTCBAR in an ELECT / R2UR.BROADCAST /
// BRA.U.ANY loop (~8 extra instructions per MMA), which made the issue loop - not data, not the tensor pipe - the
// bound of every convolution (profiles/r02_h1_ncu_swap_halo.md: the issuer's samples sit on ALU latencies).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled operand tile: rows are 128 B, 8-row atoms are 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  const uint64_t lo = (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16);            // start address, LBO (unused for swizzled K-major)
  const uint64_t hi = (uint64_t)(1024 >> 4) | (1ull << 14) | (2ull << 29);           // SBO = 1024 B, version 1 (sm_100), SWIZZLE_128B
  return lo | (hi << 32);
}

template <int BN>
__host__ __device__ constexpr uint32_t make_idesc() {
  // D fp32 (1<<4), A tf32 (2<<7), B tf32 (2<<10), both K-major, N>>3 at bit 17, M>>4 at bit 24.
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
template <int BN>
__host__ __device__ constexpr uint32_t make_idesc_f16() {
  // kind::f16: D fp32 (1<<4), A fp16 (0<<7), B fp16 (0<<10)
  return (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int B_STAGE_BYTES = BN * BKE * 4;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int TRN_OFFSET = STAGES * STAGE_BYTES;   // 8 epilogue warps x 4 KB transposition scratch
  static constexpr int BAR_OFFSET = TRN_OFFSET + 8 * 4096;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;   // barriers + slack for 1024-B alignment
  static_assert(TOTAL <= 232448, "exceeds the 227 KB shared-memory limit of sm_100");
};

// GroupNorm quad sums of one 32-column chunk held one row per lane.  st[0..7] = per-quad sums of the lane's
// row, st[8..15] = per-quad sums of squares (zeros for rows past the end).  Halving butterfly: each exchange
// sends half of the live entries to the partner lane and adds the partner's other half, so after four
// exchanges a lane holds ONE entry summed over 16 lanes (16 shuffles instead of 80); a final exchange
// completes the 32-row sum.  rows_per_img == 16 (4x4 images): the two half-warps are different images and are
// reduced separately.  The entry totals are then accumulated in fp64 by 16 (or 32) lanes in parallel.
__device__ __forceinline__ void quad_stats_commit_raw(double* qstats, int n_total, int rows_per_img, float (&st)[16],
                                                      int img, bool valid, int col0, int lane) {
  const bool halves = rows_per_img < 32;
  int idx = 0;
  if (!halves) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool up = lane & 16;
      const float recv = __shfl_xor_sync(0xffffffffu, up ? st[i] : st[i + 8], 16);
      st[i] = (up ? st[i + 8] : st[i]) + recv;
    }
    idx = (lane & 16) ? 8 : 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool up = lane & 8;
      const float recv = __shfl_xor_sync(0xffffffffu, up ? st[i] : st[i + 4], 8);
      st[i] = (up ? st[i + 4] : st[i]) + recv;
    }
    idx += (lane & 8) ? 4 : 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const bool up = lane & 4;
      const float recv = __shfl_xor_sync(0xffffffffu, up ? st[i] : st[i + 2], 4);
      st[i] = (up ? st[i + 2] : st[i]) + recv;
    }
    idx += (lane & 4) ? 2 : 0;
    {
      const bool up = lane & 2;
      const float recv = __shfl_xor_sync(0xffffffffu, up ? st[0] : st[1], 2);
      st[0] = (up ? st[1] : st[0]) + recv;
    }
    idx += (lane & 2) ? 1 : 0;
    st[0] += __shfl_xor_sync(0xffffffffu, st[0], 1);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool up = lane & 8;
      const float recv = __shfl_xor_sync(0xffffffffu, up ? st[i] : st[i + 8], 8);
      st[i] = (up ? st[i + 8] : st[i]) + recv;
    }
    idx = (lane & 8) ? 8 : 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool up = lane & 4;
      const float recv = __shfl_xor_sync(0xffffffffu, up ? st[i] : st[i + 4], 4);
      st[i] = (up ? st[i + 4] : st[i]) + recv;
    }
    idx += (lane & 4) ? 4 : 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const bool up = lane & 2;
      const float recv = __shfl_xor_sync(0xffffffffu, up ? st[i] : st[i + 2], 2);
      st[i] = (up ? st[i + 2] : st[i]) + recv;
    }
    idx += (lane & 2) ? 2 : 0;
    {
      const bool up = lane & 1;
      const float recv = __shfl_xor_sync(0xffffffffu, up ? st[0] : st[1], 1);
      st[0] = (up ? st[1] : st[0]) + recv;
    }
    idx += (lane & 1) ? 1 : 0;
  }
  // image / validity of the (half-)warp's rows: taken from its first row
  const int img_w = __shfl_sync(0xffffffffu, img, halves ? (lane & 16) : 0);
  const bool val_w = __shfl_sync(0xffffffffu, (int)valid, halves ? (lane & 16) : 0) != 0;
  const bool writer = halves ? true : ((lane & 1) == 0);
  if (writer && val_w && st[0] != 0.f) {
    const int quad = idx & 7, which = idx >> 3;          // which: 0 = sum, 1 = sum of squares
    atomicAdd(qstats + ((long long)img_w * (n_total >> 2) + (col0 >> 2) + quad) * 2 + which, (double)st[0]);
  }
}
__device__ __forceinline__ void quad_stats_commit(const TcParams& p, const Epilogue& e, float (&st)[16], int img,
                                                  bool valid, int col0, int lane) {
  quad_stats_commit_raw(p.qstats, p.N_total, e.rows_per_img, st, img, valid, col0, lane);
}

// One 32-pixel chunk of the swapped-operand epilogue (lane = output channel, v[i] = pixel i of the chunk), specialised
// on what the launch needs so the per-element instruction stream is minimal: the profile of the first version
// (profiles/r01_c19_ncu_swap_f16.md) showed the eight epilogue warps ISSUE-bound at ~40 instructions per element
// (runtime-uniform branches, an IEEE divide path for an unused divisor, 64-bit index multiplies) - longer than the
// fp16 main loop of a 128-channel convolution.  RES: add the residual; MODE 0 fp32 store, 1 TF32-rounded, 2 fp16.
template <bool RES, int MODE>
__device__ __forceinline__ void swap_chunk(const uint32_t (&v)[32], float add, float scale, float* out, const float* res,
                                           long long ld_out, long long ld_res, float& ssum, float& ssq) {
  float rr[32];
  if (RES) {
    const float* rp = res;
#pragma unroll
    for (int i = 0; i < 32; ++i) { rr[i] = __ldg(rp); rp += ld_res; }
  }
  float* op = out;                                        // MODE 2: `out` addresses fp16 elements, ld_out counts them
  uint16_t* oh = reinterpret_cast<uint16_t*>(out);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float o = __uint_as_float(v[i]) + add;
    if (RES) o += rr[i];
    o *= scale;
    if (MODE == 1) o = round_tf32(o);
    if (MODE == 2) {
      uint16_t h;
      asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(o));
      *oh = h; oh += ld_out;
    } else { *op = o; op += ld_out; }
    ssum += o; ssq = fmaf(o, o, ssq);
  }
}

// One 32-row x 32-column block of the direct epilogue (row-major outputs), specialised like swap_chunk: RES residual
// add, MODE store format (0 fp32, 1 TF32-rounded fp32, 2 fp16), STATS GroupNorm quad sums of the stored fp32 values.
// tcgen05.ld hands each lane one accumulator ROW; storing from that layout touches 32 different 128-byte lines with
// every 128-bit load/store (the in-step ncu capture of the +residual 256-channel convolution showed the eight
// epilogue warps bound by L1 wavefronts, tensor pipe 48 % vs 70 % without the residual; measured A/B of the two
// forms: 173 -> 153 us for that launch, -2.2 % per PC step, profiles/r01_c23_*).  So the warp first
// transposes its 32x32 block through 4 KB of warp-private shared memory (XOR-swizzled 16-byte slots: conflict-free
// both ways), after which lane l owns column quad l%8 of rows l/8, l/8+4, ...: eight consecutive lanes cover one
// 128-byte line, a warp instruction covers 4 lines instead of 32, bias / time-embedding quads are loaded once per
// chunk, and the GroupNorm quad sums need 2 shuffle steps instead of the 16-shuffle butterfly.
//   gm0: global row of the block's first row; rows_valid: rows of the block inside the matrix (0..32);
//   img0: image of the first row; rows of one block span two images only when rows_per_img == 16.
template <bool RES, int MODE, bool STATS>
__device__ __forceinline__ void row_chunk_t(const uint32_t (&v)[32], uint8_t* tbuf, const Epilogue& e, double* qstats, int n_total,
                                            long long gm0, int rows_valid, int n0, int img0, int lane) {
  const uint32_t tb = smem_u32(tbuf);
#pragma unroll
  for (int qd = 0; qd < 8; ++qd)
    sts128(tb + lane * 128 + ((qd ^ (lane & 7)) << 4), v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
  __syncwarp();
  const int cq = lane & 7, r0 = lane >> 3, col = n0 + cq * 4;
  const bool two_img = e.rows_per_img < 32;
  float4 addA = make_float4(0.f, 0.f, 0.f, 0.f);
  if (e.bias) addA = __ldg(reinterpret_cast<const float4*>(e.bias + col));
  float4 addB = addA;
  if (e.rowvec) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(e.rowvec + img0 * e.rowvec_ld + col));
    addA.x += t.x; addA.y += t.y; addA.z += t.z; addA.w += t.w;
    if (two_img) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(e.rowvec + (img0 + 1) * e.rowvec_ld + col));
      addB.x += u.x; addB.y += u.y; addB.z += u.z; addB.w += u.w;
    } else addB = addA;
  }
  const float scale = e.scale;
  float sA = 0.f, qA = 0.f, sB = 0.f, qB = 0.f;
  // all eight residual quads are requested before the first store: the output may alias the residual as far as
  // the compiler knows, so loads left inside the store loop are serialised behind the stores (seen as one
  // long-scoreboard stall per row in the attention kernel's profile)
  float4 rr[8];
  if (RES) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = r0 + 4 * i;
      rr[i] = row < rows_valid ? __ldg(reinterpret_cast<const float4*>(e.residual + (gm0 + row) * e.ld_res + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = r0 + 4 * i;
    float4 o = lds128(tb + row * 128 + ((cq ^ (row & 7)) << 4));
    if (row < rows_valid) {
      const long long g = gm0 + row;
      const bool second = two_img && i >= 4;                 // rows 16..31 of the block: the next image
      const float4 ad = second ? addB : addA;
      o.x += ad.x; o.y += ad.y; o.z += ad.z; o.w += ad.w;
      if (RES) { o.x += rr[i].x; o.y += rr[i].y; o.z += rr[i].z; o.w += rr[i].w; }
      o.x *= scale; o.y *= scale; o.z *= scale; o.w *= scale;
      if (MODE == 1) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
      if (MODE == 2) *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(e.out) + g * e.ld_out + col) = make_uint2(pack_half2(o.x, o.y), pack_half2(o.z, o.w));
      else *reinterpret_cast<float4*>(e.out + g * e.ld_out + col) = o;
      if (STATS) {
        const float s = (o.x + o.y) + (o.z + o.w), q2 = fmaf(o.x, o.x, o.y * o.y) + fmaf(o.z, o.z, o.w * o.w);
        if (second) { sB += s; qB += q2; } else { sA += s; qA += q2; }
      }
    }
  }
  __syncwarp();                                            // the scratch block is free for the next chunk
  if (STATS) {
    sA += __shfl_xor_sync(0xffffffffu, sA, 8); qA += __shfl_xor_sync(0xffffffffu, qA, 8);
    sA += __shfl_xor_sync(0xffffffffu, sA, 16); qA += __shfl_xor_sync(0xffffffffu, qA, 16);
    if (two_img) {
      sB += __shfl_xor_sync(0xffffffffu, sB, 8); qB += __shfl_xor_sync(0xffffffffu, qB, 8);
      sB += __shfl_xor_sync(0xffffffffu, sB, 16); qB += __shfl_xor_sync(0xffffffffu, qB, 16);
    }
    if (r0 == 0 && rows_valid > 0) {
      double* qd = qstats + ((long long)img0 * (n_total >> 2) + (col >> 2)) * 2;
      atomicAdd(qd, (double)sA); atomicAdd(qd + 1, (double)qA);
      if (two_img && rows_valid > 16) { qd += (long long)(n_total >> 2) * 2; atomicAdd(qd, (double)sB); atomicAdd(qd + 1, (double)qB); }
    }
  }
}
__device__ __forceinline__ void row_chunk_t_dispatch(const uint32_t (&v)[32], uint8_t* tbuf, const Epilogue& e, double* qstats, int n_total,
                                                     long long gm0, int rows_valid, int n0, int img0, int lane) {
  const int mode = e.round_tf32;
  const bool has_res = e.residual != nullptr, stats = qstats != nullptr;
#define B200_ROWT(R, M) do { if (stats) row_chunk_t<R, M, true>(v, tbuf, e, qstats, n_total, gm0, rows_valid, n0, img0, lane); \
                             else row_chunk_t<R, M, false>(v, tbuf, e, qstats, n_total, gm0, rows_valid, n0, img0, lane); } while (0)
  if (has_res) { if (mode == 0) B200_ROWT(true, 0); else if (mode == 1) B200_ROWT(true, 1); else B200_ROWT(true, 2); }
  else { if (mode == 0) B200_ROWT(false, 0); else if (mode == 1) B200_ROWT(false, 1); else B200_ROWT(false, 2); }
#undef B200_ROWT
}

// ---------------------------------------------------------------------------
// Kernel
// ---------------------------------------------------------------------------
template <int BN, int STAGES>
__global__ void __maxnreg__(160) gemm_tc_kernel(const __grid_constant__ TcParams p) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* wfull_bar = tmem_empty + 3;          // halo form: the weight-slice ring has its own barriers
  uint64_t* wempty_bar = wfull_bar + HALO_WS;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp index provably warp-uniform

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < HALO_WS; ++s) { mbar_init(&wfull_bar[s], 1); mbar_init(&wempty_bar[s], 1); }
    // one arrival per epilogue warp (384 threads: two warps per TMEM lane quarter)
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], (blockDim.x >> 5) - 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait(); pdl_trigger();   // barriers and TMEM are set up; nothing above touched global memory (common.cuh)

  const int kiters = (p.kchunks1 + p.kchunks2) * p.taps + p.kchunks3 + p.kchunks4;
  const int HW = p.H * p.W;
  const int bke = p.bke;

  if (warp == 0) {
    // ======================= TMA producer (whole warp; one elected lane issues, see elect_one) =======================
    const bool issue = elect_one();
    uint32_t stage = 0, phase = 0;
    if (BN == 256 && p.halo) {
      // halo form (swapped operands, one image per 256-pixel tile, W <= 32): per channel chunk three halo copies, each
      // followed by the three 16 KB weight slices of its filter column
      uint32_t ws = 0, wphase = 0;
      uint8_t* const wring = smem + HALO_XS * HALO_X_BYTES;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = (int)(tile % p.tiles_n);
        const long long p0 = (tile / p.tiles_n) * 256;
        const int img0 = (int)(p0 / HW), h0 = (int)(p0 % HW) / p.W;
        const int wrow0 = nt * 128;
        if (issue && p.halo_prefetch && tile + gridDim.x < p.total_tiles && (tile + gridDim.x) / p.tiles_n != tile / p.tiles_n) {
          const long long q0 = ((tile + gridDim.x) / p.tiles_n) * 256;                 // next tile of this CTA: its pixels, every chunk
          const int qi = (int)(q0 / HW), qh = (int)(q0 % HW) / p.W;
          for (int kc = 0; kc < p.kchunks1; ++kc) tma_prefetch_4d(&p.tmH1, kc * bke, 0, qh - 1, qi);
          for (int kc = 0; kc < p.kchunks2; ++kc) tma_prefetch_4d(&p.tmH2, kc * bke, 0, qh - 1, qi);
        }
        for (int src = 0; src < 4; ++src) {
          const int nch = src == 0 ? p.kchunks1 : src == 1 ? p.kchunks2 : src == 2 ? p.kchunks3 : p.kchunks4;
          if (nch == 0) continue;
          const int wcol0 = src == 1 ? p.C1 : src == 3 ? p.C3 : 0;
          for (int kc = 0; kc < nch; ++kc) {
            if (src < 2) {
              const CUtensorMap* tmH = src == 0 ? &p.tmH1 : &p.tmH2;
              for (int dwi = 0; dwi < 3; ++dwi) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (issue) {
                  mbar_expect_tx(&full_bar[stage], (uint32_t)p.halo_copy_bytes);
                  tma_load_4d(tmH, smem + stage * HALO_X_BYTES, &full_bar[stage], kc * bke, dwi - 1, h0 - 1, img0);
                }
                if (++stage == HALO_XS) { stage = 0; phase ^= 1; }
                for (int dhi = 0; dhi < 3; ++dhi) {
                  mbar_wait(&wempty_bar[ws], wphase ^ 1);
                  if (issue) {
                    mbar_expect_tx(&wfull_bar[ws], A_STAGE_BYTES);
                    tma_load_2d(&p.tmW, wring + ws * A_STAGE_BYTES, &wfull_bar[ws], wcol0 + kc * bke, wrow0 + (dhi * 3 + dwi) * p.N_total);
                  }
                  if (++ws == HALO_WS) { ws = 0; wphase ^= 1; }
                }
              }
            } else {
              // extra 1x1 phase (fused skip projection): the plain 256-pixel tile as two 128-pixel boxes, its own weights
              const CUtensorMap* tmA = src == 2 ? &p.tmA3 : &p.tmA4;
              const int h1 = h0 + 128 / p.W;
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if (issue) {
                mbar_expect_tx(&full_bar[stage], 2 * A_STAGE_BYTES);
                tma_load_4d(tmA, smem + stage * HALO_X_BYTES, &full_bar[stage], kc * bke, 0, h0, img0);
                tma_load_4d(tmA, smem + stage * HALO_X_BYTES + A_STAGE_BYTES, &full_bar[stage], kc * bke, 0, h1, img0);
              }
              if (++stage == HALO_XS) { stage = 0; phase ^= 1; }
              mbar_wait(&wempty_bar[ws], wphase ^ 1);
              if (issue) {
                mbar_expect_tx(&wfull_bar[ws], A_STAGE_BYTES);
                tma_load_2d(&p.tmW2, wring + ws * A_STAGE_BYTES, &wfull_bar[ws], wcol0 + kc * bke, wrow0);
              }
              if (++ws == HALO_WS) { ws = 0; wphase ^= 1; }
            }
          }
        }
      }
    } else
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = (int)(tile % p.tiles_n);
      const long long mg = tile / p.tiles_n;
      const int b = (int)(mg / p.tiles_m_per_batch);
      const int mt = (int)(mg % p.tiles_m_per_batch);
      int img0 = 0, h0 = 0, w0 = 0;
      if (p.conv) {
        const long long p0 = (long long)mt * BM;
        img0 = (int)(p0 / HW);
        const int rem = (int)(p0 % HW);
        h0 = rem / p.W; w0 = rem % p.W;
      }
      const int arow0 = b * p.a_batch_rows + mt * BM;
      const int wrow0 = b * p.w_batch_rows + nt * (p.swap ? 128 : BN);
      int img1 = 0, h1 = 0, w1 = 0;                     // swap mode: second 128-pixel box of the 256-pixel tile
      if (p.swap) {
        const long long p0 = (long long)mt * 256;
        img0 = (int)(p0 / HW); h0 = (int)(p0 % HW) / p.W; w0 = (int)(p0 % HW) % p.W;
        const long long p1 = p0 + 128;
        img1 = (int)(p1 / HW); h1 = (int)(p1 % HW) / p.W; w1 = (int)(p1 % HW) % p.W;
      }
      // sources 0,1: the (two-source) filter input, all taps; sources 2,3: the extra 1x1 phase (centre tap, own weights)
      for (int src = 0; src < 4; ++src) {
        const int nch = src == 0 ? p.kchunks1 : src == 1 ? p.kchunks2 : src == 2 ? p.kchunks3 : p.kchunks4;
        if (nch == 0) continue;
        const CUtensorMap* tmA = src == 0 ? &p.tmA1 : src == 1 ? &p.tmA2 : src == 2 ? &p.tmA3 : &p.tmA4;
        const CUtensorMap* tmW = src < 2 ? &p.tmW : &p.tmW2;
        const int wcol0 = src == 1 ? p.C1 : src == 3 ? p.C3 : 0;
        const int ntaps = src < 2 ? p.taps : 1;
        // K order.  Default: filter tap, then channel chunk - consecutive loads sweep the channel vector of the same shifted
        // pixels (contiguous 128-byte segments of every pixel row; measured ~30 % faster than the other order in the
        // CTA-pair kernel, profiles/r02_f3_*).  Shapes that have a halo form (p.chunk_major) walk K the way that form must -
        // channel chunk, filter column, filter row - so that halo on / off add the same products in the same order and
        // are bit-identical (the nine-load loop of such a shape only runs in A/B checks).
        const int nouter = p.chunk_major ? nch : ntaps, ninner = p.chunk_major ? ntaps : nch;
        for (int o = 0; o < nouter; ++o) {
          for (int i = 0; i < ninner; ++i) {
            const int kc = p.chunk_major ? o : i, t = p.chunk_major ? i : o;
            const int tap = (ntaps == 1 || !p.chunk_major) ? t : (t % p.S) * p.S + t / p.S;
            const int dh = src < 2 ? tap / p.S - p.pad : 0, dw = src < 2 ? tap % p.S - p.pad : 0;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * L::STAGE_BYTES;
            uint8_t* sb = sa + A_STAGE_BYTES;
            if (issue) {
              mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
              if (p.swap) {
                // first 16 KiB: 128 output channels x 32 k of W (UMMA A); next 32 KiB: 256 pixels x 32 k (UMMA B)
                tma_load_2d(tmW, sa, &full_bar[stage], wcol0 + kc * bke, wrow0 + tap * p.N_total);
                tma_load_4d(tmA, sb, &full_bar[stage], kc * bke, w0 + dw, h0 + dh, img0);
                tma_load_4d(tmA, sb + A_STAGE_BYTES, &full_bar[stage], kc * bke, w1 + dw, h1 + dh, img1);
              } else {
                if (p.conv) tma_load_4d(tmA, sa, &full_bar[stage], kc * bke, w0 * p.stride + dw, h0 * p.stride + dh, img0);
                else tma_load_4d(tmA, sa, &full_bar[stage], kc * bke, arow0, 0, 0);
                tma_load_2d(tmW, sb, &full_bar[stage], wcol0 + kc * bke, wrow0 + tap * p.N_total);
              }
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (whole warp; one elected lane issues) =======================
    const bool issue = elect_one();
    const bool f16 = p.f16 != 0;
    const uint32_t idesc = f16 ? make_idesc_f16<BN>() : make_idesc<BN>();
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    if (BN == 256 && p.halo) {
      // halo form: A = the 16 KB weight slice (128 output channels), B = 256 pixel rows of a halo copy starting one filter
      // row (W pixels) further in per dh; the slice's slot is released per tap, the copy's after its three taps
      uint32_t ws = 0, wphase = 0;
      const uint32_t wring = smem_u32(smem + HALO_XS * HALO_X_BYTES);
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
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
            const uint32_t sx = smem_u32(smem + stage * HALO_X_BYTES);
            for (int dhi = 0; dhi < ntap; ++dhi, ++it) {
              mbar_wait(&wfull_bar[ws], wphase);
              tc_fence_after();
              const uint64_t adesc = make_smem_desc(wring + ws * A_STAGE_BYTES);
              const uint64_t bdesc = make_smem_desc(sx + (src < 2 ? dhi * p.halo_dh_bytes : 0));
              if (issue) {
                if (f16) umma_kstep<true>(d_tmem, adesc, bdesc, idesc, it);
                else umma_kstep<false>(d_tmem, adesc, bdesc, idesc, it);
                umma_commit(&wempty_bar[ws]);
              }
              if (++ws == HALO_WS) { ws = 0; wphase ^= 1; }
            }
            if (issue) umma_commit(&empty_bar[stage]);
            if (++stage == HALO_XS) { stage = 0; phase ^= 1; }
          }
        }
        if (issue) umma_commit(&tmem_full[acc]);
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
      }
    } else
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
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
          if (f16) umma_kstep<true>(d_tmem, adesc, bdesc, idesc, it);
          else umma_kstep<false>(d_tmem, adesc, bdesc, idesc, it);
          umma_commit(&empty_bar[stage]);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (issue) umma_commit(&tmem_full[acc]);
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ======================= epilogue (direct stores) =======================
    // Eight warps: warp w may touch TMEM lanes 32*(w%4).., so warps 4..7 and 8..11 pair up on each lane quarter
    // and split the tile's columns — twice the loads/stores in flight for the output-bound (small-K) launches.
    const int q = (warp - 4) & 3;              // TMEM lane quarter owned by this warp
    const int half = (warp - 4) >> 2;          // which half of the tile's columns
    const int r = q * 32 + lane;               // row of the tile held by this thread
    const Epilogue& e = p.epi;
    uint32_t acc = 0, acc_phase = 0;
    if (p.swap) {
      // Swapped operands (D^T = W X^T, 128-channel layers): TMEM lane = output channel, column = pixel.  A warp's
      // 32 lanes are 32 consecutive channels of one pixel, so every scalar load/store instruction below touches
      // exactly one 128-byte line of the NHWC tensor (fully coalesced), and the M=128 x N=256 instruction keeps the
      // single issuing thread ahead of the tensor pipe (N=128 instructions retire in 64 cycles, faster than one
      // thread can issue them).
      if constexpr (BN == 256) {
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
          const int co = (int)(tile % p.tiles_n) * 128 + r;
          const long long row_base = (tile / p.tiles_n) * 256;
          const int img = (int)(row_base / e.rows_per_img);          // rows_per_img % 256 == 0: one image per tile
          const float add = (e.bias ? __ldg(e.bias + co) : 0.f) + (e.rowvec ? __ldg(e.rowvec + img * e.rowvec_ld + co) : 0.f);
          float ssum = 0.f, ssq = 0.f;
          if (e.out_nchw) {
            // Network head (ncsnpp.py:374-380): the 128-row weight tile is zero-padded above the n_valid image
            // channels, so only lanes < n_valid of lane quarter 0 hold results: (acc + bias) / sigma[img], written
            // as NCHW with 128-bit stores along the pixel axis.  The other warps just release the accumulator.
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            if (q == 0) {
              const float dv = e.per_img_div ? __ldg(e.per_img_div + img * e.div_stride) : 1.f;
              const int hw = e.rows_per_img, pix_in_img = (int)(row_base - (long long)img * hw);
#pragma unroll 1
              for (int j = half * 4; j < half * 4 + 4; ++j) {
                uint32_t v[32];
                tmem_ld32(tmem_base + acc * BN + j * 32, v);
                if (lane < e.n_valid) {
                  float* dst = e.out + ((long long)img * e.n_valid + lane) * hw + pix_in_img + j * 32;
#pragma unroll
                  for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(dst + i) = make_float4((__uint_as_float(v[i]) + add) / dv, (__uint_as_float(v[i + 1]) + add) / dv,
                                                                      (__uint_as_float(v[i + 2]) + add) / dv, (__uint_as_float(v[i + 3]) + add) / dv);
                }
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            acc ^= 1; if (acc == 0) acc_phase ^= 1;
            continue;
          }
          if (e.residual) {
#pragma unroll
            for (int j = 0; j < 4; ++j)   // this lane's share of the warp's 128 residual lines (one per pixel)
              prefetch_l2(e.residual + (row_base + half * 128 + j * 32 + lane) * e.ld_res + (int)(tile % p.tiles_n) * 128 + q * 32);
          }
          mbar_wait(&tmem_full[acc], acc_phase);
          tc_fence_after();
          const int mode = e.round_tf32;
          const bool has_res = e.residual != nullptr;
#pragma unroll 1
          for (int j = half * 4; j < half * 4 + 4; ++j) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + j * 32, v);
            const long long pix0 = row_base + j * 32;
            const float* res = has_res ? e.residual + pix0 * e.ld_res + co : nullptr;
            // element offset of (pix0, co); fp16 outputs count halves from the same base pointer
            float* dst = mode == 2 ? reinterpret_cast<float*>(reinterpret_cast<uint16_t*>(e.out) + pix0 * e.ld_out + co)
                                   : e.out + pix0 * e.ld_out + co;
            if (has_res) {
              if (mode == 0) swap_chunk<true, 0>(v, add, e.scale, dst, res, e.ld_out, e.ld_res, ssum, ssq);
              else if (mode == 1) swap_chunk<true, 1>(v, add, e.scale, dst, res, e.ld_out, e.ld_res, ssum, ssq);
              else swap_chunk<true, 2>(v, add, e.scale, dst, res, e.ld_out, e.ld_res, ssum, ssq);
            } else {
              if (mode == 0) swap_chunk<false, 0>(v, add, e.scale, dst, res, e.ld_out, e.ld_res, ssum, ssq);
              else if (mode == 1) swap_chunk<false, 1>(v, add, e.scale, dst, res, e.ld_out, e.ld_res, ssum, ssq);
              else swap_chunk<false, 2>(v, add, e.scale, dst, res, e.ld_out, e.ld_res, ssum, ssq);
            }
          }
          if (p.qstats) {
            ssum += __shfl_xor_sync(0xffffffffu, ssum, 1); ssq += __shfl_xor_sync(0xffffffffu, ssq, 1);
            ssum += __shfl_xor_sync(0xffffffffu, ssum, 2); ssq += __shfl_xor_sync(0xffffffffu, ssq, 2);
            if ((lane & 3) == 0) {
              double* qd = p.qstats + ((long long)img * (p.N_total >> 2) + (co >> 2)) * 2;
              atomicAdd(qd, (double)ssum); atomicAdd(qd + 1, (double)ssq);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          acc ^= 1; if (acc == 0) acc_phase ^= 1;
        }
      }
    } else
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = (int)(tile % p.tiles_n);
      const long long mg = tile / p.tiles_n;
      const int b = (int)(mg / p.tiles_m_per_batch);
      const int mt = (int)(mg % p.tiles_m_per_batch);
      const int row0 = mt * BM + q * 32;                                   // first row of this warp's 32-row blocks
      const int rows_valid = min(32, max(0, p.M_per_batch - row0));
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
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
}

#include "gemm_tc2.cuh"
#include "attn_tc.cuh"
#include "gemm_tcg.cuh"

// ---------------------------------------------------------------------------
// Host side: tensor maps, plan, launch
// ---------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
  });
  return fn;
}

int encode_map(CUtensorMap* tm, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, const uint32_t* elem_strides = nullptr, bool f16 = false) {
  auto fn = get_encode_fn();
  B200_REQUIRE(fn != nullptr, "gemm_tc: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "gemm_tc: operand base %p not 16-byte aligned", (const void*)base);
  uint32_t estr[5] = {1, 1, 1, 1, 1};
  if (elem_strides) for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
  CUresult r = fn(tm, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), dims,
                  strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, "gemm_tc: cuTensorMapEncodeTiled failed with CUresult %d "
               "(rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)", (int)r, rank,
               (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
               (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  return 0;
}

int num_sms() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
  return n;
}

}  // namespace

struct TcGemmPlan {
  TcParams prm;
  int bn;
  bool two_cta;
};

static int tc_configure();

bool tc_gemm_supported(const TcGemmDesc& d, const char** why) {
  static const char* w;
  auto fail = [&](const char* m) { w = m; if (why) *why = w; return false; };
  const int bke = d.f16 ? 64 : BKE;
  if (d.C1 % bke || d.C2 % bke || d.C1 <= 0) return fail("channel counts must be multiples of 32 (tf32) / 64 (f16)");
  if (d.N_total % 128) return fail("N must be a multiple of 128");
  if (d.taps != 1 && d.taps != 9) return fail("only 1x1 and 3x3 filters");
  if (d.conv) {
    if (d.nbatch != 1) return fail("conv mode is unbatched");
    if (d.stride > 2) return fail("stride must be 1 or 2");
    if (d.stride == 2 && d.a2) return fail("strided conv is single-source");
    const int HW = d.H * d.W;
    if (HW >= BM) {
      if (d.W >= BM) { if (d.W % BM) return fail("image width does not tile 128 pixels"); }
      else if ((BM % d.W) || (d.H % (BM / d.W))) return fail("image rows do not tile 128 pixels");
    } else if (BM % HW) return fail("image size does not divide 128 pixels");
  } else {
    if (d.nbatch > 1 && (d.M_per_batch % BM)) return fail("batched gemm needs M_per_batch % 128 == 0");
    if (d.a_ld % (d.f16 ? 8 : 4)) return fail("A row pitch must be a multiple of 16 bytes");
  }
  if (d.a3) {
    if (!d.conv || d.stride == 2 || !d.w2) return fail("extra 1x1 phase needs a stride-1 convolution and its weights");
    if (d.C3 % bke || d.C3 <= 0 || (d.a4 && d.C4 % bke)) return fail("extra-phase channel counts must be multiples of 32 (tf32) / 64 (f16)");
  }
  if (d.epi.out_nchw) {   // network head: zero-padded 128-row weight tile, swapped form, NCHW store of the real channels
    if (!(d.conv && d.N_total == 128 && d.epi.n_valid > 0 && d.epi.n_valid <= 32 && (d.H * d.W) % 256 == 0 && (d.W <= BM || d.W % BM == 0) && d.stride != 2 &&
          !d.qstats && !d.epi.residual && !d.epi.rowvec && d.epi.round_tf32 == 0))
      return fail("NCHW head needs a 128-row padded weight tile, HW % 256 == 0 and a plain epilogue");
  } else if (d.epi.per_img_div) return fail("per-image divisor only with the NCHW head epilogue");
  // a 32-row epilogue block spans two images only for 4x4 images (rows_per_img == 16)
  if ((d.qstats || d.epi.rowvec) && !(d.epi.rows_per_img % 32 == 0 || d.epi.rows_per_img == 16)) return fail("per-image epilogue terms need rows_per_img % 32 == 0 or == 16");
  if (d.epi.ld_out % 4 || (d.epi.residual && d.epi.ld_res % 4)) return fail("output pitch must be a multiple of 4 elements");
  if (d.epi.round_tf32 == 2 && d.epi.ld_out % 8) return fail("fp16 output pitch must be a multiple of 8 elements");
  return true;
}

int tc_gemm_plan_create(const TcGemmDesc& d, TcGemmPlan** out) {
  const char* why = nullptr;
  B200_REQUIRE(tc_gemm_supported(d, &why), "gemm_tc: unsupported shape: %s", why ? why : "?");
  if (int r = tc_configure()) return r;
  TcGemmPlan* pl = new TcGemmPlan();
  TcParams& p = pl->prm;
  memset(&p, 0, sizeof(p));
  pl->bn = (d.N_total % 256 == 0) ? 256 : 128;
  p.conv = d.conv; p.H = d.conv ? d.H : 1; p.W = d.conv ? d.W : 1; p.taps = d.taps;
  p.S = d.taps == 9 ? 3 : 1; p.pad = (d.taps == 9 && !d.valid_pad) ? 1 : 0;
  p.stride = d.stride == 2 ? 2 : 1;
  p.f16 = d.f16 ? 1 : 0;

  const bool f16 = d.f16 != 0;
  const int bke = f16 ? 64 : BKE;            // elements per 128-byte K step
  const uint64_t es = f16 ? 2 : 4;           // operand element size
  p.bke = bke;
  {
    // Form selection (each choice measured against the alternatives in round 1):
    //  * swapped operands (D^T = W X^T: 128 output channels x 256 pixels per tile) for 128-channel outputs, whose
    //    M=128,N=128 MMAs are issue-bound, and for 1x1 convolutions (output-bound launches: the swapped form's
    //    epilogue is perfectly coalesced, 12-20 % faster, profiles/r01_c11_exp.log); the NCHW head exists only in this form;
    //  * CTA pairs (cta_group::2) for 256-column tiles when the launch has at least one 256-row pair per cluster
    //    slot (small launches fill the SMs better with single-CTA tiles); `no_pair` in the descriptor opts out;
    //  * launches too small to give every SM a 256-column tile are cut into 128-column tiles: twice the CTAs at work.
    const long long Mtot = (long long)d.nimg * d.H * d.W;
    // (image rows wider than 128 pixels - the 256..1024-pixel families - are cut into 128-pixel boxes like any other: the
    // two boxes of a 256-pixel tile are then two halves of one row or of consecutive rows)
    const bool can_swap = d.conv && p.stride == 1 && (d.N_total % 256 != 0 || d.taps == 1) && (d.H * d.W) % 256 == 0 &&
                          (d.W <= BM || d.W % BM == 0) && Mtot % 256 == 0 && d.epi.rows_per_img % 256 == 0;
    p.swap = can_swap ? 1 : 0;
    if (d.epi.out_nchw && !p.swap) { delete pl; B200_REQUIRE(false, "gemm_tc: the NCHW head needs the swapped-operand form"); }
    if (p.swap) pl->bn = 256;
    p.qstats = d.qstats;
    const int tmb = d.conv ? 1 : (d.M_per_batch + BM - 1) / BM;
    const long long m_tiles = d.conv ? (Mtot + BM - 1) / BM : (long long)d.nbatch * tmb;
    const long long n_tiles = d.N_total % 256 == 0 ? d.N_total / 256 : d.N_total / 128;
    pl->two_cta = !d.no_pair && !p.swap && d.N_total % 256 == 0 &&
                  (d.conv || d.nbatch == 1 || tmb % 2 == 0) && (m_tiles / 2) * n_tiles >= num_sms() / 2;
    if (!pl->two_cta && !p.swap && pl->bn == 256 && m_tiles * n_tiles <= num_sms() / 2) pl->bn = 128;
  }
  p.kchunks1 = d.C1 / bke; p.kchunks2 = d.a2 ? d.C2 / bke : 0; p.C1 = d.C1;
  p.N_total = d.N_total; p.tiles_n = p.swap ? d.N_total / 128 : d.N_total / pl->bn;
  p.a_batch_rows = d.a_batch_rows; p.w_batch_rows = d.w_batch_rows;
  p.epi = d.epi;

  int rc = 0;
  if (d.conv) {
    const int HW = d.H * d.W;
    const long long M = (long long)d.nimg * HW;
    p.nbatch = 1; p.M_per_batch = (int)M; p.tiles_m_per_batch = p.swap ? (int)(M / 256) : (int)((M + BM - 1) / BM);
    uint32_t box[4];
    if (HW >= BM) { box[1] = std::min(d.W, BM); box[2] = BM / box[1]; box[3] = 1; }
    else { box[1] = d.W; box[2] = d.H; box[3] = BM / HW; }
    box[0] = (uint32_t)bke;
    // stride 2: the box *traverses* 2x as many pixels and TMA keeps every other one
    const uint32_t estr[4] = {1, (uint32_t)p.stride, (uint32_t)p.stride, 1};
    box[1] *= p.stride; box[2] *= p.stride;
    const int Hin = d.Hin ? d.Hin : d.H, Win = d.Win ? d.Win : d.W;
    for (int s = 0; s < 2; ++s) {
      const float* base = s ? d.a2 : d.a1;
      const int C = s ? d.C2 : d.C1;
      if (!base) continue;
      uint64_t dims[4] = {(uint64_t)C, (uint64_t)Win, (uint64_t)Hin, (uint64_t)d.nimg};
      uint64_t str[3] = {(uint64_t)C * es, (uint64_t)Win * C * es, (uint64_t)Hin * Win * C * es};
      rc = encode_map(s ? &p.tmA2 : &p.tmA1, base, 4, dims, str, box, estr, f16);
      if (rc) { delete pl; return rc; }
    }
  } else {
    p.nbatch = d.nbatch; p.M_per_batch = d.M_per_batch; p.tiles_m_per_batch = (d.M_per_batch + BM - 1) / BM;
    uint32_t box[4] = {(uint32_t)bke, BM, 1, 1};
    for (int s = 0; s < 2; ++s) {
      const float* base = s ? d.a2 : d.a1;
      if (!base) continue;
      uint64_t dims[4] = {(uint64_t)(s ? d.C2 : d.C1), (uint64_t)d.a_rows, 1, 1};
      uint64_t str[3] = {(uint64_t)d.a_ld * es, (uint64_t)d.a_ld * es * d.a_rows, (uint64_t)d.a_ld * es * d.a_rows};
      rc = encode_map(s ? &p.tmA2 : &p.tmA1, base, 4, dims, str, box, nullptr, f16);
      if (rc) { delete pl; return rc; }
    }
  }
  if (!d.a2) p.tmA2 = p.tmA1;
  p.tmA3 = p.tmA1; p.tmA4 = p.tmA1;
  p.tmH1 = p.tmA1; p.tmH2 = p.tmA1;
  {
    // Halo form: 'same'-padded stride-1 3x3 filters on 16- or 32-pixel-wide images, where the CTA's tile
    // (256 pixels swapped, 128 pixels per CTA of a pair) is whole rows of ONE image: 3 loads of (rows + 2) x W pixels per
    // channel chunk instead of 9 loads of rows x W - the plain form is bound by the L2 -> SM fill rate, not the tensor pipe.
    const int tile_px = p.swap ? 256 : BM;
    const bool halo = (d.no_halo & 3) != 1 && d.conv && d.taps == 9 && p.pad == 1 && p.stride == 1 && (p.swap || (pl->two_cta && (d.no_halo & 3) == 2)) &&
                      (d.W == 16 || d.W == 32) && (d.H * d.W) % tile_px == 0 && (d.Hin == 0 || d.Hin == d.H) && (d.Win == 0 || d.Win == d.W);
    // swapped-form shapes that have a halo form keep its K order when it is switched off (bit-identical A/B).  `no_halo & 8`
    // asks the same of the single-CTA row-major kernel for shapes the CTA-pair kernel would run in halo form: a small-batch
    // plan then adds the same products in the same order as the pair plan of a large batch (plan-agreement tests; the
    // chunk-major nine-load loop is ~30 % slower, so nothing else sets it)
    const bool halo_shape = d.conv && d.taps == 9 && p.pad == 1 && p.stride == 1 && (d.W == 16 || d.W == 32) && (d.H * d.W) % tile_px == 0 &&
                            (d.Hin == 0 || d.Hin == d.H) && (d.Win == 0 || d.Win == d.W);
    p.chunk_major = (halo_shape && (p.swap || ((d.no_halo & 8) && !pl->two_cta))) ? 1 : 0;
    if (halo) {
      const int rows = tile_px / d.W;
      p.halo = 1; p.halo_dh_bytes = d.W * 128; p.halo_copy_bytes = (rows + 2) * d.W * 128;
      p.halo_prefetch = (d.no_halo & 4) ? 1 : 0;
      B200_REQUIRE(p.halo_copy_bytes <= (p.swap ? HALO_X_BYTES : HALO2_X_BYTES), "gemm_tc: halo copy of %d bytes does not fit its slot", p.halo_copy_bytes);
      uint32_t box[4] = {(uint32_t)bke, (uint32_t)d.W, (uint32_t)(rows + 2), 1};
      for (int s = 0; s < 2; ++s) {
        const float* base = s ? d.a2 : d.a1;
        const int C = s ? d.C2 : d.C1;
        if (!base) continue;
        uint64_t dims[4] = {(uint64_t)C, (uint64_t)d.W, (uint64_t)d.H, (uint64_t)d.nimg};
        uint64_t str[3] = {(uint64_t)C * es, (uint64_t)d.W * C * es, (uint64_t)d.H * d.W * C * es};
        rc = encode_map(s ? &p.tmH2 : &p.tmH1, base, 4, dims, str, box, nullptr, f16);
        if (rc) { delete pl; return rc; }
      }
    }
  }
  if (d.a3) {
    // extra 1x1 phase: same pixel box as the main input (stride 1, same spatial size), its own channel counts
    const int HW = d.H * d.W;
    uint32_t box[4];
    if (HW >= BM) { box[1] = std::min(d.W, BM); box[2] = BM / box[1]; box[3] = 1; }
    else { box[1] = d.W; box[2] = d.H; box[3] = BM / HW; }
    box[0] = (uint32_t)bke;
    for (int s = 0; s < 2; ++s) {
      const float* base = s ? d.a4 : d.a3;
      const int C = s ? d.C4 : d.C3;
      if (!base) continue;
      uint64_t dims[4] = {(uint64_t)C, (uint64_t)d.W, (uint64_t)d.H, (uint64_t)d.nimg};
      uint64_t str[3] = {(uint64_t)C * es, (uint64_t)d.W * C * es, (uint64_t)d.H * d.W * C * es};
      rc = encode_map(s ? &p.tmA4 : &p.tmA3, base, 4, dims, str, box, nullptr, f16);
      if (rc) { delete pl; return rc; }
    }
    p.kchunks3 = d.C3 / bke; p.kchunks4 = d.a4 ? d.C4 / bke : 0; p.C3 = d.C3;
  }
  {
    uint64_t dims[2] = {(uint64_t)d.K_total, (uint64_t)d.w_rows};
    uint64_t str[1] = {(uint64_t)(d.w_ld ? d.w_ld : d.K_total) * es};
    uint32_t box[2] = {(uint32_t)bke, (uint32_t)(pl->two_cta ? pl->bn / 2 : (p.swap ? 128 : pl->bn))};
    rc = encode_map(&p.tmW, d.w, 2, dims, str, box, nullptr, f16);
    if (rc) { delete pl; return rc; }
    p.tmW2 = p.tmW;
    if (d.a3) {
      const uint64_t K3 = (uint64_t)d.C3 + (d.a4 ? d.C4 : 0);
      uint64_t dims2[2] = {K3, (uint64_t)d.N_total};
      uint64_t str2[1] = {K3 * es};
      rc = encode_map(&p.tmW2, d.w2, 2, dims2, str2, box, nullptr, f16);
      if (rc) { delete pl; return rc; }
    }
  }
  p.total_tiles = (long long)p.nbatch * p.tiles_m_per_batch * p.tiles_n;
  *out = pl;
  return 0;
}

void tc_gemm_plan_destroy(TcGemmPlan* p) { delete p; }
const char* tc_gemm_form(const TcGemmPlan* p) {
  if (p->two_cta) return p->prm.halo ? "pair256-halo" : "pair256";
  if (p->prm.swap) return p->prm.halo ? "swap-halo" : "swap";
  return p->bn == 256 ? "single256" : "single128";
}
void tc_gemm_set_rowvec_ld(TcGemmPlan* p, long long ld) { p->prm.epi.rowvec_ld = ld; }
void tc_gemm_set_head(TcGemmPlan* p, float* out_nchw, const float* per_img_div, long long div_stride) {
  p->prm.epi.out = out_nchw; p->prm.epi.per_img_div = per_img_div; p->prm.epi.div_stride = div_stride;
}
// ---- fused attention core (attn_tc.cuh) ----
struct TcAttnPlan { AttnParams prm; bool f16; };

bool tc_attn_supported(int T, int C) { return T == AT_T && C == AT_C; }

int tc_attn_plan_create(const TcAttnDesc& d, TcAttnPlan** out) {
  B200_REQUIRE(tc_attn_supported(d.T, d.C), "attn_tc: only T=256, C=256 (got T=%d C=%d)", d.T, d.C);
  B200_REQUIRE(d.qk && d.vT && d.w3 && d.bv && d.b3 && d.x && d.out && d.nimg > 0, "attn_tc: null argument");
  if (int r = tc_configure()) return r;
  TcAttnPlan* pl = new TcAttnPlan();
  AttnParams& p = pl->prm;
  memset(&p, 0, sizeof(p));
  pl->f16 = d.f16 != 0;
  const bool f16 = pl->f16;
  const uint64_t es = f16 ? 2 : 4;
  const uint32_t bka = f16 ? 64 : 32;
  int rc = 0;
  {
    uint64_t dims[2] = {(uint64_t)2 * AT_C, (uint64_t)d.nimg * AT_T};
    uint64_t str[1] = {(uint64_t)2 * AT_C * es};
    uint32_t boxq[2] = {bka, (uint32_t)BM}, boxk[2] = {bka, 256};
    rc = encode_map(&p.tmQ, d.qk, 2, dims, str, boxq, nullptr, f16);
    if (!rc) rc = encode_map(&p.tmK, d.qk, 2, dims, str, boxk, nullptr, f16);
  }
  if (!rc) {
    uint64_t dims[2] = {(uint64_t)AT_T, (uint64_t)d.nimg * AT_C};
    uint64_t str[1] = {(uint64_t)AT_T * es};
    uint32_t box[2] = {bka, 256};
    rc = encode_map(&p.tmVT, d.vT, 2, dims, str, box, nullptr, f16);
  }
  if (!rc) {
    uint64_t dims[2] = {(uint64_t)AT_C, (uint64_t)AT_C};
    uint64_t str[1] = {(uint64_t)AT_C * es};
    uint32_t box[2] = {bka, 256};
    rc = encode_map(&p.tmW3, d.w3, 2, dims, str, box, nullptr, f16);
  }
  if (rc) { delete pl; return rc; }
  p.bv = d.bv; p.b3 = d.b3; p.x = d.x; p.out = d.out; p.qstats = d.qstats; p.nimg = d.nimg;
  p.logit_scale = (float)((1.0 / std::sqrt((double)AT_C)) * 1.4426950408889634);
  p.out_scale = d.out_scale;
  *out = pl;
  return 0;
}
void tc_attn_plan_destroy(TcAttnPlan* p) { delete p; }
int tc_attn_launch(const TcAttnPlan* pl, cudaStream_t st) {
  const long long tiles = 2LL * pl->prm.nimg;
  const int grid = (int)std::min<long long>(tiles, num_sms());
  if (pl->f16) launch_kernel(attn_tc_kernel<true>, dim3(grid), dim3(384), AttnSmem<true>::TOTAL, st, pl->prm);
  else launch_kernel(attn_tc_kernel<false>, dim3(grid), dim3(384), AttnSmem<false>::TOTAL, st, pl->prm);
  B200_CHECK_LAUNCH();
  return 0;
}

// ---- convolution with GroupNorm (+SiLU) applied on load (gemm_tcg.cuh) ----
struct TcgPlan { TcgParams prm; };

bool tcg_supported(const TcgDesc& d, const char** why) {
  static const char* w;
  auto fail = [&](const char* m) { w = m; if (why) *why = w; return false; };
  if (d.W != 16 && d.W != 32) return fail("image width must be 16 or 32");
  if (d.H <= 0 || (d.H * d.W) % 128 || d.H % (128 / d.W)) return fail("image rows do not tile 128 pixels");
  if (d.N_total % 256) return fail("C_out must be a multiple of 256");
  if (d.C1 <= 0 || d.C1 % 64 || d.C2 % 64 || d.C3 % 64 || d.C4 % 64) return fail("channel counts must be multiples of 64");
  if ((d.a2 == nullptr) != (d.C2 == 0) || (d.a3 == nullptr) != (d.C3 == 0) || (d.a4 == nullptr) != (d.C4 == 0)) return fail("source / channel-count mismatch");
  if (d.a4 && !d.a3) return fail("second extra source without the first");
  if (d.a3 && !d.w2) return fail("extra 1x1 phase without weights");
  if ((d.gn_scale == nullptr) != (d.gn_shift == nullptr)) return fail("GroupNorm scale and shift come together");
  if (d.epi.out_nchw || d.epi.per_img_div) return fail("plain NHWC epilogue only");
  if (d.epi.ld_out % 4 || (d.epi.residual && d.epi.ld_res % 4)) return fail("output pitch must be a multiple of 4 elements");
  if (d.epi.round_tf32 == 2 && d.epi.ld_out % 8) return fail("fp16 output pitch must be a multiple of 8 elements");
  if (d.epi.rows_per_img != d.H * d.W) return fail("rows_per_img must be H*W");
  return true;
}

int tcg_plan_create(const TcgDesc& d, TcgPlan** out) {
  const char* why = nullptr;
  B200_REQUIRE(tcg_supported(d, &why), "gemm_tcg: unsupported shape: %s", why ? why : "?");
  B200_REQUIRE(d.a1 && d.w && d.nimg > 0, "gemm_tcg: null argument");
  if (int r = tc_configure()) return r;
  TcgPlan* pl = new TcgPlan();
  TcgParams& p = pl->prm;
  memset(&p, 0, sizeof(p));
  p.src[0] = d.a1; p.srcC[0] = d.C1; p.srcF16[0] = d.a1_f16; p.kch[0] = d.C1 / 64;
  p.src[1] = d.a2; p.srcC[1] = d.C2; p.srcF16[1] = d.a2_f16; p.kch[1] = d.a2 ? d.C2 / 64 : 0;
  p.src[2] = d.a3; p.srcC[2] = d.C3; p.srcF16[2] = d.a3_f16; p.kch[2] = d.a3 ? d.C3 / 64 : 0;
  p.src[3] = d.a4; p.srcC[3] = d.C4; p.srcF16[3] = d.a4_f16; p.kch[3] = d.a4 ? d.C4 / 64 : 0;
  for (int s = 0; s < 4; ++s)
    if (p.src[s] && (reinterpret_cast<uintptr_t>(p.src[s]) & 15)) { delete pl; B200_REQUIRE(false, "gemm_tcg: source %d not 16-byte aligned", s); }
  p.scale = d.gn_scale; p.shift = d.gn_shift; p.Cgn = d.C1 + d.C2; p.act = d.act;
  p.H = d.H; p.W = d.W; p.R = 128 / d.W;
  p.N_total = d.N_total; p.tiles_n = d.N_total / 256;
  p.tiles_m = (long long)d.nimg * d.H * d.W / 128;
  p.qstats = d.qstats; p.epi = d.epi;
  {
    const uint64_t K = (uint64_t)d.C1 + d.C2;
    uint64_t dims[2] = {K, (uint64_t)9 * d.N_total};
    uint64_t str[1] = {K * 2};
    uint32_t box[2] = {64, 128};
    int rc = encode_map(&p.tmW, d.w, 2, dims, str, box, nullptr, true);
    p.tmW2 = p.tmW;
    if (!rc && d.a3) {
      const uint64_t K3 = (uint64_t)d.C3 + d.C4;
      uint64_t dims2[2] = {K3, (uint64_t)d.N_total};
      uint64_t str2[1] = {K3 * 2};
      rc = encode_map(&p.tmW2, d.w2, 2, dims2, str2, box, nullptr, true);
    }
    if (rc) { delete pl; return rc; }
  }
  *out = pl;
  return 0;
}
void tcg_plan_destroy(TcgPlan* p) { delete p; }
void tcg_set_rowvec_ld(TcgPlan* p, long long ld) { p->prm.epi.rowvec_ld = ld; }
int tcg_launch(const TcgPlan* pl, cudaStream_t st) {
  const TcgParams& q = pl->prm;
  const long long pairs = ((q.tiles_m + 1) / 2) * q.tiles_n;
  if (pairs == 0) return 0;
  const int grid = (int)std::min<long long>(2 * pairs, (long long)(num_sms() & ~1));
  if (q.W == 32) launch_kernel(conv_gn2_kernel<true>, dim3(grid), dim3(TG_THREADS), SmemG::TOTAL, st, q);
  else launch_kernel(conv_gn2_kernel<false>, dim3(grid), dim3(TG_THREADS), SmemG::TOTAL, st, q);
  B200_CHECK_LAUNCH();
  return 0;
}

// Opt in to the large dynamic shared-memory carve-out once, outside any stream capture.
static int tc_configure() {
  static bool configured = false;
  if (configured) return 0;
  B200_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout<256, 4>::TOTAL));
  B200_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<128, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout<128, 6>::TOTAL));
  B200_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<256, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem2<256, 6>::TOTAL));
  B200_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<false>::TOTAL));
  B200_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<true>::TOTAL));
  B200_CHECK_CUDA(cudaFuncSetAttribute(conv_gn2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemG::TOTAL));
  B200_CHECK_CUDA(cudaFuncSetAttribute(conv_gn2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemG::TOTAL));
  configured = true;
  return 0;
}

template <int BN, int STAGES>
static int launch_impl(const TcGemmPlan* pl, cudaStream_t st) {
  using L = SmemLayout<BN, STAGES>;
  const int grid = (int)std::min<long long>(pl->prm.total_tiles, num_sms());
  launch_kernel(gemm_tc_kernel<BN, STAGES>, dim3(grid), dim3(384), L::TOTAL, st, pl->prm);
  B200_CHECK_LAUNCH();
  return 0;
}

int tc_gemm_launch(const TcGemmPlan* pl, cudaStream_t st) {
  if (pl->prm.total_tiles == 0) return 0;
  if (pl->two_cta) {
    const TcParams& q = pl->prm;
    const long long pairs = (((long long)q.nbatch * q.tiles_m_per_batch + 1) / 2) * q.tiles_n;
    const int grid = (int)std::min<long long>(2 * pairs, (long long)(num_sms() & ~1));
    launch_kernel(gemm_tc2_kernel<256, 6>, dim3(grid), dim3(384), Smem2<256, 6>::TOTAL, st, pl->prm);
    B200_CHECK_LAUNCH();
    return 0;
  }
  return pl->bn == 256 ? launch_impl<256, 4>(pl, st) : launch_impl<128, 6>(pl, st);
}

}  // namespace b200
