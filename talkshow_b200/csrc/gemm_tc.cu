// talkshow_b200 — tensor-core implicit-GEMM for the dense contractions (face wav2vec2 convs /
// transformer, VQ decoder convs): tcgen05.mma kind::tf32 with fp32-grade accuracy by the 3xTF32
// split  x = hi + lo,  A*B ~= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  (fp32 accumulate in TMEM).
//
//  * operands arrive pre-split (hi = fp32 with the 13 low mantissa bits cleared, lo = x - hi, exact)
//    as K-major tiles: TMA (cp.async.bulk.tensor, 128B swizzle) stages [128 x 32] fp32 boxes of
//    A_hi, A_lo, W_hi, W_lo per k-block into a 3-stage shared-memory ring;
//  * one elected thread issues 12 tcgen05.mma (4 k-steps x 3 products) per stage into a 128x128 fp32
//    accumulator in TMEM; tcgen05.commit releases the stage / publishes the accumulator;
//  * 4 epilogue warps read TMEM (tcgen05.ld 32x32b), add bias / residual, apply the activation and
//    write either plain fp32 or the (hi, lo) pair the next tensor-core GEMM consumes.
//  A Conv1d(k, stride s) is the same kernel: the A tensor map is 3-D {C, s, rows/s} over the padded
//  channel-last buffer, tap t reads box (c0, t % s, j + t / s); GEMM rows are indexed by the padded row
//  index j, rows that fall on padding are computed and discarded by the epilogue.
//
// fp16-split variant (ts_set_tensor_cores(e, 6), the default): the same two-term split carried by fp16 instead of tf32 —
// x = h + l with h = fp16(x), l = fp16(x - h): 11 + 11 significant bits, the same as two tf32 terms — and the three
// products issued as kind::f16, which runs at twice the tf32 rate: a k-block of 128 bytes per row holds 64 K values
// instead of 32, so the same 12 MMAs per stage cover twice the K.  Half the tensor time, half the operand bytes
// (HBM, L2 -> shared memory and shared memory -> tensor core).  Range: |x| < 65504 (fp16); weights are pre-scaled by a
// power of two per layer so their low plane stays normal (undone exactly in the epilogue); activations whose low plane
// underflows lose absolute accuracy below 2^-25 only.
#include <cuda.h>
#include <cuda_fp16.h>

#include "kernels.h"

namespace ts {

// 128 x 256 CTA tile: with N = 256 one UMMA reads 4 KB of A + 8 KB of B from shared memory per 128
// tensor clocks (96 B/clk) instead of 8 KB per 64 clocks at N = 128 (128 B/clk = the whole smem port),
// which measured as the limiter of the 128 x 128 version (multicast and L2 prefetch changed nothing).
constexpr int TC_BM = 128, TC_BN = 256, TC_BK = 32, TC_STAGES = 2;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;               // 16 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 4;               // 32 KB
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;   // A_hi, A_lo, B_hi, B_lo = 96 KB
constexpr int TC_EPI_WARPS = 8;                             // 4 TMEM lane quadrants x 2 column halves
constexpr int TC_SMEM = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
// The tensor core's fp32 accumulator truncates (measured: error grows linearly with K, biased toward
// zero), so K is accumulated in TMEM only over chunks of TC_CHUNK k-blocks (K = 256); the epilogue
// warps drain each chunk into fp32 registers (round-to-nearest adds) while the MMA warp fills the
// other of the two TMEM accumulator buffers.
constexpr int TC_CHUNK = 8;
constexpr int TC_PREFETCH = 8;   // k-blocks of L2 prefetch distance for the A operand
constexpr int TC_THREADS = 128 + 32 * TC_EPI_WARPS;          // warpgroup 0: TMA, MMA (+2 idle); warps 4-19 epilogue

struct TcArgs {
  int taps, cblocks, stride;        // K loop = taps x (C / 32)
  int C;                            // input channels (K per tap)
  int rows_in, off, T_out, nbatch;  // epilogue row mapping (see header)
  int Rs;                           // GEMM rows (padded row index / stride)
  int N;
  float* c_hi;                      // output (plain fp32 when c_lo == nullptr)
  float* c_lo;
  long c_bs, c_rs;
  const float* bias;
  const float* r_hi;                // residual: plain (r_lo null) or split
  const float* r_lo;
  long r_bs, r_rs;
  int act;
  const float* a_hi; const float* a_lo;   // raw A arrays (row r = a + r * C), for whole-row L2 prefetches
  long a_rows;                      // rows in the A arrays
  int nprod;                        // 3 (normal) or 1 (hi*hi only: throughput experiment)
  int prefetch_rows;                // pair kernel: whole-row L2 prefetch of the A operand one tap ahead
  int cn, cm;                       // cluster shape: cn N-tiles x cm M-tiles share operands by TMA multicast
  int chunk;                        // k-blocks accumulated in TMEM before the epilogue drains them (K = 256)
  float oscale;                     // accumulator scale (undoes the power-of-two weight scale of the fp16 split), 1 for tf32
  unsigned short* c_h16;            // fp16-split copy of the output (c_hi then holds the full value, c_lo is null)
  unsigned short* c_l16;
};

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(b)), "r"(c)); }
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nTCW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra TCD;\nbra TCW;\nTCD:\n}\n" ::"r"(s_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_3d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s_u32(dst)),
               "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(s_u32(bar))
               : "memory");
}
// TMA prefetch of a box into L2 (no shared-memory destination): hides HBM latency of k-blocks that
// do not fit in the 3-stage ring yet
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(m), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s_u32(dst)),
               "l"(m), "r"(c0), "r"(c1), "r"(s_u32(bar))
               : "memory");
}
// multicast variants: the box lands at the same smem offset in every CTA of `mask` and completes tx
// bytes on the mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_3d_mc(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(
          s_u32(dst)),
      "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(s_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_2d_mc(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
          s_u32(dst)),
      "l"(m), "r"(c0), "r"(c1), "r"(s_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(s_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// K-major, 128B-swizzled operand tile: 8-row groups are 1024 B apart (SBO), LBO unused (=1), version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(a), "l"(b),
      "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(taddr));
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

__device__ __forceinline__ float tc_act(float v, int act) {
  if (act == ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == ACT_LRELU) return v > 0.f ? v : 0.2f * v;
  if (act == ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
  return v;
}

// Epilogue warps (4..11): TMEM lane quadrant = warp % 4, column half = (warp - 4) / 4.
//  1. drain: after every K chunk the warp adds its 32 lanes x 128 columns of the TMEM accumulator into
//     fp32 registers (round-to-nearest) and frees the buffer (`tempty_addr[buf]`: shared::cluster address of
//     the MMA issuer's barrier — own CTA, or the leader CTA of a pair);
//  2. the 128 x 256 tile is parked in shared memory (the operand ring is idle by then) and written out by
//     a compact loop with threads along N: coalesced bias / residual loads and float4 stores, and ~150
//     instructions of code instead of a 10k-instruction unrolled epilogue (the first version spent 30 % of
//     its warp samples in instruction-fetch stalls).
constexpr int TC_SLD = TC_BN + 4;   // padded row of the parked tile (floats)

// Write-out of the parked tile.  Thread -> 4 fixed columns (bias loaded once), rows strided by 4; the first version did the
// row / column arithmetic, a runtime activation switch and scalar tails per element: ~420 instructions per float4, 28 us
// per tile -- more than a K = 768 tile's whole main loop (ncu, profiles/r02_gemm_f16_ncu_summary.md).
// CG = float4 column groups per parked row (the 256 epilogue threads cover 256 / CG rows per pass), SLD = parked row stride.
template <int ACT, int CG, int SLD>
__device__ __forceinline__ void tc_writeout(const TcArgs& P, const float* S, const long* rowc, const long* rowr, int n0) {
  const int et = threadIdx.x - 128;
  const int c4 = (et % CG) * 4, n = n0 + c4;
  if (n >= P.N) return;
  const int nv = min(4, P.N - n);                   // valid columns of this thread (4 except at a ragged N edge)
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (P.bias)
    for (int q = 0; q < nv; ++q) bv[q] = P.bias[n + q];
  const float osc = P.oscale;
  const bool rvec = nv == 4 && P.r_hi && (((P.r_bs | P.r_rs) & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.r_hi) & 15) == 0) &&
                    (!P.r_lo || (reinterpret_cast<uintptr_t>(P.r_lo) & 15) == 0);
#pragma unroll 4
  for (int r = et / CG; r < TC_BM; r += 256 / CG) {
    const long co = rowc[r];
    if (co < 0) continue;
    const float4 sv = *reinterpret_cast<const float4*>(&S[r * SLD + c4]);
    float o[4] = {sv.x * osc + bv[0], sv.y * osc + bv[1], sv.z * osc + bv[2], sv.w * osc + bv[3]};
    if (P.r_hi) {
      const long roff = rowr[r] + n;
      if (rvec) {
        const float4 rh = *reinterpret_cast<const float4*>(P.r_hi + roff);
        if (P.r_lo) {
          const float4 rl = *reinterpret_cast<const float4*>(P.r_lo + roff);
          o[0] += rh.x + rl.x; o[1] += rh.y + rl.y; o[2] += rh.z + rl.z; o[3] += rh.w + rl.w;
        } else {
          o[0] += rh.x; o[1] += rh.y; o[2] += rh.z; o[3] += rh.w;
        }
      } else {
        for (int q = 0; q < nv; ++q) o[q] += P.r_lo ? (P.r_hi[roff + q] + P.r_lo[roff + q]) : P.r_hi[roff + q];
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (ACT == ACT_GELU) o[q] = 0.5f * o[q] * (1.0f + erff(o[q] * 0.70710678118654752440f));
      else if (ACT != ACT_NONE) o[q] = tc_act(o[q], P.act);
    }
    const long coff = co + n;
    if (nv == 4) {
      if (P.c_h16) {
        if (P.c_hi) *reinterpret_cast<float4*>(P.c_hi + coff) = make_float4(o[0], o[1], o[2], o[3]);
        ushort4 h, l;
        split16(o[0], h.x, l.x); split16(o[1], h.y, l.y); split16(o[2], h.z, l.z); split16(o[3], h.w, l.w);
        *reinterpret_cast<ushort4*>(P.c_h16 + coff) = h;
        *reinterpret_cast<ushort4*>(P.c_l16 + coff) = l;
      } else if (P.c_lo) {
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(o[0]) & 0xffffe000u); l.x = o[0] - h.x;
        h.y = __uint_as_float(__float_as_uint(o[1]) & 0xffffe000u); l.y = o[1] - h.y;
        h.z = __uint_as_float(__float_as_uint(o[2]) & 0xffffe000u); l.z = o[2] - h.z;
        h.w = __uint_as_float(__float_as_uint(o[3]) & 0xffffe000u); l.w = o[3] - h.w;
        *reinterpret_cast<float4*>(P.c_hi + coff) = h;
        *reinterpret_cast<float4*>(P.c_lo + coff) = l;
      } else {
        *reinterpret_cast<float4*>(P.c_hi + coff) = make_float4(o[0], o[1], o[2], o[3]);
      }
    } else {
      for (int q = 0; q < nv; ++q) {
        if (P.c_h16) {
          if (P.c_hi) P.c_hi[coff + q] = o[q];
          split16(o[q], P.c_h16[coff + q], P.c_l16[coff + q]);
        } else if (P.c_lo) {
          const float h = __uint_as_float(__float_as_uint(o[q]) & 0xffffe000u);
          P.c_hi[coff + q] = h;
          P.c_lo[coff + q] = o[q] - h;
        } else {
          P.c_hi[coff + q] = o[q];
        }
      }
    }
  }
}

__device__ __forceinline__ void tc_epilogue(const TcArgs& P, unsigned char* smem, int warp, int lane, int j0, int n0, int nk, uint32_t tmem_base,
                                            uint64_t* tfull, const uint32_t* tempty_addr) {
  const int quad = warp & 3;
  const int half = (warp - 4) >> 2;
  constexpr int EC = TC_BN / 2;
  const int row = quad * 32 + lane;
  float acc[EC];
#pragma unroll
  for (int i = 0; i < EC; ++i) acc[i] = 0.f;
  const int nchunks = (nk + P.chunk - 1) / P.chunk;
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    mb_wait(&tfull[buf], (c >> 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int cc = 0; cc < EC / 16; ++cc) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + buf * TC_BN + half * EC + cc * 16, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[cc * 16 + i] += __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(tempty_addr[buf]) : "memory");
  }
  // ---- park the tile: S[row][col], plus the output / residual offsets of every GEMM row ----------------------
  float* S = reinterpret_cast<float*>(smem);
  long* rowc = reinterpret_cast<long*>(S + TC_BM * TC_SLD);
  long* rowr = rowc + TC_BM;
#pragma unroll
  for (int c = 0; c < EC; c += 4)
    *reinterpret_cast<float4*>(&S[row * TC_SLD + half * EC + c]) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
  if (half == 0) {  // GEMM row j <-> padded input row j*stride -> (batch, output step)
    const int j = j0 + row;
    long co = -1, ro = 0;
    if (j < P.Rs) {
      const long in_row = (long)j * P.stride;
      const int bb = (int)(in_row / P.rows_in);
      const int tin = (int)(in_row - (long)bb * P.rows_in) - P.off;
      if (bb < P.nbatch && tin >= 0 && (tin % P.stride) == 0 && tin / P.stride < P.T_out) {
        const int t = tin / P.stride;
        co = (long)bb * P.c_bs + (long)t * P.c_rs;
        ro = (long)bb * P.r_bs + (long)t * P.r_rs;
      }
    }
    rowc[row] = co;
    rowr[row] = ro;
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 epilogue warps only
  if (P.act == ACT_NONE) tc_writeout<ACT_NONE, TC_BN / 4, TC_SLD>(P, S, rowc, rowr, n0);
  else if (P.act == ACT_GELU) tc_writeout<ACT_GELU, TC_BN / 4, TC_SLD>(P, S, rowc, rowr, n0);
  else tc_writeout<-1, TC_BN / 4, TC_SLD>(P, S, rowc, rowr, n0);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mA_hi, const __grid_constant__ CUtensorMap mA_lo,
               const __grid_constant__ CUtensorMap mB_hi, const __grid_constant__ CUtensorMap mB_lo, TcArgs P) {
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
  uint64_t* empty = full + TC_STAGES;
  uint64_t* tfull = empty + TC_STAGES;   // [2] accumulator buffer ready
  uint64_t* tempty = tfull + 2;          // [2] accumulator buffer drained (4 epilogue warps arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // N tiles vary fastest: the CTAs that share an A row block run together and hit it in L2
  const int j0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_BN;
  const int nk = P.taps * P.cblocks;
  // cluster = cn x cm tiles; CTAs in my row (same M tile) share the A box, CTAs in my column share the B box
  const uint32_t crank = cluster_ctarank();
  const int cx = crank % P.cn, cy = crank / P.cn;
  const uint16_t mask_row = (uint16_t)(((1u << P.cn) - 1u) << (cy * P.cn));
  uint16_t mask_col = 0;
  for (int j = 0; j < P.cm; ++j) mask_col |= (uint16_t)(1u << (j * P.cn + cx));
  const bool clustered = P.cn * P.cm > 1;

  if (threadIdx.x == 0) {
    // a stage is free once every CTA that received one of my slices has consumed it: cn + cm - 1 arrivals
    for (int i = 0; i < TC_STAGES; ++i) { mb_init(&full[i], 1); mb_init(&empty[i], P.cn + P.cm - 1); }
    mb_init(&tfull[0], 1); mb_init(&tfull[1], 1);
    mb_init(&tempty[0], TC_EPI_WARPS); mb_init(&tempty[1], TC_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mA_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mA_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mB_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mB_lo) : "memory");
  }
  if (warp == 1) {  // TMEM allocation: two 256-column fp32 accumulator buffers (all 512 columns)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "n"(2 * TC_BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (clustered) cluster_sync_all();   // peers' barriers are initialised before anyone multicasts / arrives remotely
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;


  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      for (int kb = 0; kb < nk; ++kb) {
        const int st = kb % TC_STAGES, ph = (kb / TC_STAGES) & 1;
        mb_wait(&empty[st], ph ^ 1);
        const int tap = kb / P.cblocks, cb = kb - tap * P.cblocks;
        unsigned char* base = smem + st * TC_STAGE_BYTES;
        mb_expect(&full[st], TC_STAGE_BYTES);
        // my slices: rows [cx*128/cn, +128/cn) of the A box, rows [cy*128/cm, +128/cm) of the B box
        const int ar = TC_BM / P.cn, br = TC_BN / P.cm;
        const int c0 = cb * TC_BK, c1 = tap % P.stride, c2 = j0 + tap / P.stride + cx * ar;
        tma_3d_mc(base + cx * ar * 128, &mA_hi, c0, c1, c2, &full[st], mask_row);
        tma_3d_mc(base + TC_A_BYTES + cx * ar * 128, &mA_lo, c0, c1, c2, &full[st], mask_row);
        const int kcol = tap * P.C + cb * TC_BK;
        tma_2d_mc(base + 2 * TC_A_BYTES + cy * br * 128, &mB_hi, kcol, n0 + cy * br, &full[st], mask_col);
        tma_2d_mc(base + 2 * TC_A_BYTES + TC_B_BYTES + cy * br * 128, &mB_lo, kcol, n0 + cy * br, &full[st], mask_col);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      // instruction descriptor: D=F32, A=B=TF32, K-major both, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int kb = 0; kb < nk; ++kb) {
        const int st = kb % TC_STAGES, ph = (kb / TC_STAGES) & 1;
        const int chunk = kb / P.chunk, kin = kb - chunk * P.chunk, buf = chunk & 1;
        if (kin == 0) {                       // this accumulator buffer must have been drained
          mb_wait(&tempty[buf], ((chunk >> 1) & 1) ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        mb_wait(&full[st], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = s_u32(smem + st * TC_STAGE_BYTES), a_lo = a_hi + TC_A_BYTES, b_hi = a_hi + 2 * TC_A_BYTES,
                       b_lo = b_hi + TC_B_BYTES;
        const uint32_t d = tmem_base + buf * TC_BN;
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {
          const uint32_t o = k * 32;  // 8 tf32 = 32 bytes along K inside the 128B swizzle atom
          if (P.nprod == 3) {
            umma_tf32(d, umma_desc(a_lo + o), umma_desc(b_hi + o), idesc, (kin | k) != 0);
            umma_tf32(d, umma_desc(a_hi + o), umma_desc(b_lo + o), idesc, 1);
            umma_tf32(d, umma_desc(a_hi + o), umma_desc(b_hi + o), idesc, 1);
          } else {
            umma_tf32(d, umma_desc(a_hi + o), umma_desc(b_hi + o), idesc, (kin | k) != 0);
          }
        }
        umma_commit_mc(&empty[st], mask_row | mask_col);   // release the stage to every CTA that filled it
        if (kin == P.chunk - 1 || kb == nk - 1) umma_commit(&tfull[buf]);   // chunk accumulated
      }
    }
  } else if (warp >= 4) {
    const uint32_t te[2] = {s_u32(&tempty[0]), s_u32(&tempty[1])};
    tc_epilogue(P, smem, warp, lane, j0, n0, nk, tmem_base, tfull, te);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (clustered) cluster_sync_all();   // nobody exits while a peer may still multicast into it / arrive on its barriers
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * TC_BN) : "memory");
  }
}

// ---- CTA-pair variant (cta_group::2) -------------------------------------------------------------
// Two CTAs of a (1,2) cluster form one 256 x 256 tile: each CTA stages its own 128 A rows and HALF of the
// 256-row B tile (64 KB per k-block instead of 96 KB: the per-SM L2->smem delivery that limits the
// single-CTA kernel), the leader issues tcgen05.mma.cta_group::2 (M = 256) which reads both halves of B,
// and each CTA's epilogue drains its own 128 TMEM lanes.  TMA completions of both CTAs land on the leader's
// `full` barrier; tcgen05.commit multicasts the `empty` / `tfull` arrivals to both CTAs.
constexpr int T2_STAGES = 3;
constexpr int T2_BHALF = TC_BN / 2 * TC_BK * 4;                      // 16 KB: 128 rows of the B tile
constexpr int T2_STAGE_BYTES = 2 * TC_A_BYTES + 2 * T2_BHALF;      // 64 KB
constexpr int T2_SMEM = T2_STAGES * T2_STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ void tma_3d_2sm(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint32_t leader_bar) {
  asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   s_u32(dst)),
               "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(leader_bar)
               : "memory");
}
__device__ __forceinline__ void tma_2d_2sm(void* dst, const CUtensorMap* m, int c0, int c1, uint32_t leader_bar) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   s_u32(dst)),
               "l"(m), "r"(c0), "r"(c1), "r"(leader_bar)
               : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(a), "l"(b),
      "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(a), "l"(b),
      "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {   // arrives on the barrier at this offset in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(s_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

// ONCHIP = true (round 2 experiment, ts_set_tensor_cores(e, 5)): the operands arrive as PLAIN fp32 — one TMA box per operand
// and k-block instead of two, i.e. half the L2 -> shared-memory bytes — and the two otherwise idle warps
// of warpgroup 0 split them in shared memory: lo = x - (x & 0xffffe000) is written to the `lo` tile at the SAME offset
// (elementwise, so the 128 B swizzle is irrelevant); x itself serves as the `hi` operand — measured: kind::tf32 ignores the
// 13 low mantissa bits, results are bit-identical to the pre-split kernel.  The converters of both CTAs then arrive on the
// leader's `full` barrier (generic -> async proxy fence first), which the MMA issuer waits on.  Measured (face, 64 clips x
// 10 s): 66.5 ms vs 54.1 ms with operands pre-split in HBM — the extra hop (TMA -> converter -> MMA) on a 3-stage ring costs
// more than the halved L2 -> SM traffic gains, so pre-split stays the default (profiles/r02_summary.md).
// F16 = true: operands are fp16 planes (h16, l16), k-blocks of 64 K values, kind::f16 (see the file header).
template <bool ONCHIP, bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc2_gemm_kernel(const __grid_constant__ CUtensorMap mA_hi, const __grid_constant__ CUtensorMap mA_lo,
                const __grid_constant__ CUtensorMap mB_hi, const __grid_constant__ CUtensorMap mB_lo, TcArgs P) {
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + T2_STAGES * T2_STAGE_BYTES);
  uint64_t* empty = full + T2_STAGES;
  uint64_t* tfull = empty + T2_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint64_t* raw = tempty + 4;                       // ONCHIP: this CTA's plain operand boxes have landed
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader of the pair
  // 1-D grid, (2,1,1) clusters: blockIdx.x = 2 * (n_tile + tiles_n * m_pair) + member  (P.cn carries tiles_n);
  // the pair covers M tiles 2*m_pair and 2*m_pair+1 of one N tile, N tiles vary fastest (A row blocks shared in L2)
  const int lin = blockIdx.x >> 1, nt = lin % P.cn, mp = lin / P.cn;
  const int j0 = (2 * mp + (int)rank) * TC_BM, n0 = nt * TC_BN;
  const int nk = P.taps * P.cblocks;
  constexpr int BK = F16 ? 2 * TC_BK : TC_BK;      // K values per 128-byte k-block row
  static_assert(!(ONCHIP && F16), "the on-chip split experiment exists for tf32 operands only");

  if (threadIdx.x == 0) {
    // full: TMA bytes of both CTAs (pre-split operands) or the 2 x 2 converter warps of the pair (ONCHIP)
    for (int i = 0; i < T2_STAGES; ++i) { mb_init(&full[i], ONCHIP ? 4 : 1); mb_init(&empty[i], 1); mb_init(&raw[i], 1); }
    mb_init(&tfull[0], 1); mb_init(&tfull[1], 1);
    mb_init(&tempty[0], 2 * TC_EPI_WARPS); mb_init(&tempty[1], 2 * TC_EPI_WARPS);   // epilogue warps of both CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mA_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mA_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mB_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mB_lo) : "memory");
  }
  if (warp == 1) {  // pair-wide TMEM allocation (same warp id, same slot offset in both CTAs)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "n"(2 * TC_BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // Whole-row L2 prefetch: a tap's A rows (C*4 bytes contiguous each) are pulled into L2 as full rows one
    // tap ahead, so DRAM sees 2-4 KB bursts instead of the 128-byte pieces the k-block boxes would request
    auto prefetch_tap = [&](int tap) {
      if (tap >= P.taps) return;
      const uint32_t bytes = (uint32_t)P.C * 4u;   // tf32 operands only (P.prefetch_rows is 0 for fp16 planes)
      for (int r = lane; r < TC_BM; r += 32) {
        const long row = (long)(j0 + r + tap / P.stride) * P.stride + tap % P.stride;
        if (row < P.a_rows) {
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.a_hi + row * P.C), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.a_lo + row * P.C), "r"(bytes) : "memory");
        }
      }
    };
    if (P.prefetch_rows) { prefetch_tap(0); prefetch_tap(1); }
    __syncwarp();
    for (int kbp = 0; kbp < nk; kbp += P.cblocks) {   // one iteration per tap
      if (P.prefetch_rows && kbp > 0) prefetch_tap(kbp / P.cblocks + 1);
      __syncwarp();
      if (lane == 0)
      for (int kb = kbp; kb < kbp + P.cblocks; ++kb) {
        const int st = kb % T2_STAGES, ph = (kb / T2_STAGES) & 1;
        mb_wait(&empty[st], ph ^ 1);
        const int tap = kb / P.cblocks, cb = kb - tap * P.cblocks;
        unsigned char* base = smem + st * T2_STAGE_BYTES;
        const int c0 = cb * BK, c1 = tap % P.stride, c2 = j0 + tap / P.stride;
        const int kcol = tap * P.C + cb * BK, nrow = n0 + (int)rank * (TC_BN / 2);
        if constexpr (ONCHIP) {
          mb_expect(&raw[st], TC_A_BYTES + T2_BHALF);                       // this CTA's two plain boxes, on its own barrier
          tma_3d(base, &mA_hi, c0, c1, c2, &raw[st]);
          tma_2d(base + 2 * TC_A_BYTES, &mB_hi, kcol, nrow, &raw[st]);
        } else {
          if (rank == 0) mb_expect(&full[st], 2 * T2_STAGE_BYTES);          // bytes of both CTAs land on the leader's barrier
          const uint32_t lbar = s_u32(&full[st]) & 0xFEFFFFFFu;             // peer bit cleared -> CTA 0 of the pair
          tma_3d_2sm(base, &mA_hi, c0, c1, c2, lbar);
          tma_3d_2sm(base + TC_A_BYTES, &mA_lo, c0, c1, c2, lbar);
          tma_2d_2sm(base + 2 * TC_A_BYTES, &mB_hi, kcol, nrow, lbar);
          tma_2d_2sm(base + 2 * TC_A_BYTES + T2_BHALF, &mB_lo, kcol, nrow, lbar);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {  // ===== MMA issuer: leader CTA only =====
      // instruction descriptor: D = F32 (bit 4), A / B format (bits 7.., 10..) TF32 = 2 or F16 = 0, N >> 3 at 17, M >> 4 at 24
      const uint32_t fmt = F16 ? 0u : 2u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)((2 * TC_BM) >> 4) << 24);
      for (int kb = 0; kb < nk; ++kb) {
        const int st = kb % T2_STAGES, ph = (kb / T2_STAGES) & 1;
        const int chunk = kb / P.chunk, kin = kb - chunk * P.chunk, buf = chunk & 1;
        if (kin == 0) {
          mb_wait(&tempty[buf], ((chunk >> 1) & 1) ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        mb_wait(&full[st], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = s_u32(smem + st * T2_STAGE_BYTES), a_lo = a_hi + TC_A_BYTES, b_hi = a_hi + 2 * TC_A_BYTES, b_lo = b_hi + T2_BHALF;
        const uint32_t d = tmem_base + buf * TC_BN;
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {
          const uint32_t o = k * 32;    // 32 bytes along K inside the 128 B swizzle atom: 8 tf32 or 16 fp16 = one MMA's K
          if constexpr (F16) {
            umma2_f16(d, umma_desc(a_lo + o), umma_desc(b_hi + o), idesc, (kin | k) != 0);
            umma2_f16(d, umma_desc(a_hi + o), umma_desc(b_lo + o), idesc, 1);
            umma2_f16(d, umma_desc(a_hi + o), umma_desc(b_hi + o), idesc, 1);
          } else {
            umma2_tf32(d, umma_desc(a_lo + o), umma_desc(b_hi + o), idesc, (kin | k) != 0);
            umma2_tf32(d, umma_desc(a_hi + o), umma_desc(b_lo + o), idesc, 1);
            umma2_tf32(d, umma_desc(a_hi + o), umma_desc(b_hi + o), idesc, 1);
          }
        }
        umma2_commit(&empty[st]);
        if (kin == P.chunk - 1 || kb == nk - 1) umma2_commit(&tfull[buf]);
      }
    }
  } else if (warp < 4) {
    if constexpr (ONCHIP) {
      // ===== converter warps 2, 3 (64 threads) of BOTH CTAs: hi / lo split of the two plain tiles in shared memory =====
      const int ct = (warp - 2) * 32 + lane;
      for (int kb = 0; kb < nk; ++kb) {
        const int st = kb % T2_STAGES, ph = (kb / T2_STAGES) & 1;
        mb_wait(&raw[st], ph);
        unsigned char* base = smem + st * T2_STAGE_BYTES;
#pragma unroll
        for (int op = 0; op < 2; ++op) {
          uint4* x4 = reinterpret_cast<uint4*>(base + (op ? 2 * TC_A_BYTES : 0));
          float4* l4 = reinterpret_cast<float4*>(base + (op ? 2 * TC_A_BYTES + T2_BHALF : TC_A_BYTES));
#pragma unroll 4
          for (int i = ct; i < TC_A_BYTES / 16; i += 64) {       // A tile and B half tile are both 16 KB
            uint4 v = x4[i];
            uint4 h = make_uint4(v.x & 0xffffe000u, v.y & 0xffffe000u, v.z & 0xffffe000u, v.w & 0xffffe000u);
            l4[i] = make_float4(__uint_as_float(v.x) - __uint_as_float(h.x), __uint_as_float(v.y) - __uint_as_float(h.y),
                                __uint_as_float(v.z) - __uint_as_float(h.z), __uint_as_float(v.w) - __uint_as_float(h.w));
            if (P.nprod == 4) x4[i] = h;   // TS_TC_NPROD=4: also truncate x in place (the tensor core ignores the 13 low bits itself)
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core's reads
        __syncwarp();
        if (lane == 0) {
          uint32_t fa;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(fa) : "r"(s_u32(&full[st])), "r"(0));
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(fa) : "memory");
        }
      }
    }
  } else {
    // the "buffer drained" arrivals of both CTAs go to the leader's tempty barriers
    uint32_t te[2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(te[i]) : "r"(s_u32(&tempty[i])), "r"(0));
    tc_epilogue(P, smem, warp, lane, j0, n0, nk, tmem_base, tfull, te);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * TC_BN) : "memory");
  }
}

// ---- persistent CTA-pair variant ------------------------------------------------------------------------
// One CTA pair per TPC loops over the 256 x 256 tiles (tile = pair + i * pairs, N tiles fastest).  The roles keep running
// counters (k-blocks for the operand ring, chunks for the two TMEM buffers), so the TMA producer and the MMA issuer run
// straight into the next tile while the epilogue warps of both CTAs are still writing the previous one out: the write-out
// (8 - 10 us per tile, as long as a K = 768 main loop) is hidden instead of serialised, and barrier init / TMEM allocation /
// pipeline fill are paid once per CTA instead of once per tile.  The parked tile can no longer borrow the operand ring, so
// every epilogue warp transposes its 32 rows x 128 columns through a private 32 x 16 park, 16 columns at a time.
constexpr int TP_COLS = 16;                          // columns per write-out step of one warp
constexpr int TP_SLD = TP_COLS + 4;                  // parked row stride (floats): the row-per-lane STS.128 is conflict-free
constexpr int TP_PARK = TC_EPI_WARPS * 32 * TP_SLD * 4;   // 20 KB: a private 32 x 16 park per epilogue warp
constexpr int TP_ROWS = TC_EPI_WARPS * 32 * 2 * 8;   // per-warp output / residual row offsets
constexpr int TP_SMEM = T2_STAGES * T2_STAGE_BYTES + TP_PARK + TP_ROWS + 256 + 1024;

// One write-out step of one epilogue warp: its 32 rows x 16 columns, parked row-per-lane, leave as 8 rows x 64 B per store
// instruction (lane -> row lane / 4 + 8 i, 4 columns lane % 4).  Warp-private: no block barrier, the 8 warps of a CTA
// interleave freely and hide each other's load latency.
template <int ACT>
__device__ __forceinline__ void tp_rows(const TcArgs& P, const float* Sw, const long* rc, const long* rr, int n, int lane) {
  const int cg = lane & 3;
  n += cg * 4;
  if (n >= P.N) return;
  const int nv = min(4, P.N - n);
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (P.bias)
    for (int q = 0; q < nv; ++q) bv[q] = P.bias[n + q];
  const float osc = P.oscale;
  const bool rvec = nv == 4 && P.r_hi && (((P.r_bs | P.r_rs) & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.r_hi) & 15) == 0) &&
                    (!P.r_lo || (reinterpret_cast<uintptr_t>(P.r_lo) & 15) == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (lane >> 2) + 8 * i;
    const long co = rc[r];
    if (co < 0) continue;
    const float4 sv = *reinterpret_cast<const float4*>(&Sw[r * TP_SLD + cg * 4]);
    float o[4] = {sv.x * osc + bv[0], sv.y * osc + bv[1], sv.z * osc + bv[2], sv.w * osc + bv[3]};
    if (P.r_hi) {
      const long roff = rr[r] + n;
      if (rvec) {
        const float4 rh = *reinterpret_cast<const float4*>(P.r_hi + roff);
        if (P.r_lo) {
          const float4 rl = *reinterpret_cast<const float4*>(P.r_lo + roff);
          o[0] += rh.x + rl.x; o[1] += rh.y + rl.y; o[2] += rh.z + rl.z; o[3] += rh.w + rl.w;
        } else {
          o[0] += rh.x; o[1] += rh.y; o[2] += rh.z; o[3] += rh.w;
        }
      } else {
        for (int q = 0; q < nv; ++q) o[q] += P.r_lo ? (P.r_hi[roff + q] + P.r_lo[roff + q]) : P.r_hi[roff + q];
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (ACT == ACT_GELU) o[q] = 0.5f * o[q] * (1.0f + erff(o[q] * 0.70710678118654752440f));
      else if (ACT != ACT_NONE) o[q] = tc_act(o[q], P.act);
    }
    const long coff = co + n;
    if (nv == 4) {
      if (P.c_h16) {
        if (P.c_hi) *reinterpret_cast<float4*>(P.c_hi + coff) = make_float4(o[0], o[1], o[2], o[3]);
        ushort4 h, l;
        split16(o[0], h.x, l.x); split16(o[1], h.y, l.y); split16(o[2], h.z, l.z); split16(o[3], h.w, l.w);
        *reinterpret_cast<ushort4*>(P.c_h16 + coff) = h;
        *reinterpret_cast<ushort4*>(P.c_l16 + coff) = l;
      } else if (P.c_lo) {
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(o[0]) & 0xffffe000u); l.x = o[0] - h.x;
        h.y = __uint_as_float(__float_as_uint(o[1]) & 0xffffe000u); l.y = o[1] - h.y;
        h.z = __uint_as_float(__float_as_uint(o[2]) & 0xffffe000u); l.z = o[2] - h.z;
        h.w = __uint_as_float(__float_as_uint(o[3]) & 0xffffe000u); l.w = o[3] - h.w;
        *reinterpret_cast<float4*>(P.c_hi + coff) = h;
        *reinterpret_cast<float4*>(P.c_lo + coff) = l;
      } else {
        *reinterpret_cast<float4*>(P.c_hi + coff) = make_float4(o[0], o[1], o[2], o[3]);
      }
    } else {
      for (int q = 0; q < nv; ++q) {
        if (P.c_h16) {
          if (P.c_hi) P.c_hi[coff + q] = o[q];
          split16(o[q], P.c_h16[coff + q], P.c_l16[coff + q]);
        } else if (P.c_lo) {
          const float h = __uint_as_float(__float_as_uint(o[q]) & 0xffffe000u);
          P.c_hi[coff + q] = h;
          P.c_lo[coff + q] = o[q] - h;
        } else {
          P.c_hi[coff + q] = o[q];
        }
      }
    }
  }
}
// All 8 steps of one tile for one warp.  The step loop is NOT unrolled (the body with erff must exist once per activation);
// the accumulator registers need constant indices, so the 16 values of step j are picked by a switch.  No function call and
// no address of P taken: a first version passed P to a __noinline__ helper, which moved the whole argument block to local
// memory -- with 222 KB of the SM's 228 KB given to shared memory those loads missed L1 in every role's inner loop and
// the kernel ran 40 % slower than the one-tile-per-CTA kernel.
template <int ACT>
__device__ __forceinline__ void tp_tile(const TcArgs& P, const float (&acc)[TC_BN / 2], float* Sw, const long* rc, const long* rr, int nbase,
                                        int lane) {
#pragma unroll 1
  for (int j = 0; j < (TC_BN / 2) / TP_COLS; ++j) {
    float q[TP_COLS];
    switch (j) {
#define TP_PICK(J)                                                  \
  case J:                                                           \
    _Pragma("unroll") for (int c = 0; c < TP_COLS; ++c) q[c] = acc[J * TP_COLS + c]; \
    break;
      TP_PICK(0) TP_PICK(1) TP_PICK(2) TP_PICK(3) TP_PICK(4) TP_PICK(5) TP_PICK(6)
      default:
#pragma unroll
        for (int c = 0; c < TP_COLS; ++c) q[c] = acc[7 * TP_COLS + c];
        break;
#undef TP_PICK
    }
#pragma unroll
    for (int c = 0; c < TP_COLS; c += 4) *reinterpret_cast<float4*>(&Sw[lane * TP_SLD + c]) = make_float4(q[c], q[c + 1], q[c + 2], q[c + 3]);
    __syncwarp();
    tp_rows<ACT>(P, Sw, rc, rr, nbase + j * TP_COLS, lane);
    __syncwarp();
  }
}

template <bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc2p_gemm_kernel(const __grid_constant__ CUtensorMap mA_hi, const __grid_constant__ CUtensorMap mA_lo,
                 const __grid_constant__ CUtensorMap mB_hi, const __grid_constant__ CUtensorMap mB_lo, TcArgs P, int ntiles) {
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  float* S = reinterpret_cast<float*>(smem + T2_STAGES * T2_STAGE_BYTES);
  long* rows = reinterpret_cast<long*>(smem + T2_STAGES * T2_STAGE_BYTES + TP_PARK);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + T2_STAGES * T2_STAGE_BYTES + TP_PARK + TP_ROWS);
  uint64_t* empty = full + T2_STAGES;
  uint64_t* tfull = empty + T2_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int nk = P.taps * P.cblocks;
  constexpr int BK = F16 ? 2 * TC_BK : TC_BK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < T2_STAGES; ++i) { mb_init(&full[i], 1); mb_init(&empty[i], 1); }
    mb_init(&tfull[0], 1); mb_init(&tfull[1], 1);
    mb_init(&tempty[0], 2 * TC_EPI_WARPS); mb_init(&tempty[1], 2 * TC_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mA_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mA_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mB_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mB_lo) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "n"(2 * TC_BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      uint32_t it = 0;
      for (int tile = pair; tile < ntiles; tile += npairs) {
        const int nt = tile % P.cn, mp = tile / P.cn;
        const int j0 = (2 * mp + (int)rank) * TC_BM, n0 = nt * TC_BN;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int st = it % T2_STAGES, ph = (it / T2_STAGES) & 1;
          mb_wait(&empty[st], ph ^ 1);
          const int tap = kb / P.cblocks, cb = kb - tap * P.cblocks;
          unsigned char* base = smem + st * T2_STAGE_BYTES;
          const int c0 = cb * BK, c1 = tap % P.stride, c2 = j0 + tap / P.stride;
          const int kcol = tap * P.C + cb * BK, nrow = n0 + (int)rank * (TC_BN / 2);
          if (rank == 0) mb_expect(&full[st], 2 * T2_STAGE_BYTES);
          const uint32_t lbar = s_u32(&full[st]) & 0xFEFFFFFFu;
          tma_3d_2sm(base, &mA_hi, c0, c1, c2, lbar);
          tma_3d_2sm(base + TC_A_BYTES, &mA_lo, c0, c1, c2, lbar);
          tma_2d_2sm(base + 2 * TC_A_BYTES, &mB_hi, kcol, nrow, lbar);
          tma_2d_2sm(base + 2 * TC_A_BYTES + T2_BHALF, &mB_lo, kcol, nrow, lbar);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {  // ===== MMA issuer: leader CTA only =====
      const uint32_t fmt = F16 ? 0u : 2u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)((2 * TC_BM) >> 4) << 24);
      uint32_t it = 0, gch = 0;
      for (int tile = pair; tile < ntiles; tile += npairs) {
        int kin = 0;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int st = it % T2_STAGES, ph = (it / T2_STAGES) & 1;
          const uint32_t buf = gch & 1;
          if (kin == 0) {
            mb_wait(&tempty[buf], ((gch >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          mb_wait(&full[st], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = s_u32(smem + st * T2_STAGE_BYTES), a_lo = a_hi + TC_A_BYTES, b_hi = a_hi + 2 * TC_A_BYTES, b_lo = b_hi + T2_BHALF;
          const uint32_t d = tmem_base + buf * TC_BN;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint32_t o = k * 32;
            if constexpr (F16) {
              umma2_f16(d, umma_desc(a_lo + o), umma_desc(b_hi + o), idesc, (kin | k) != 0);
              umma2_f16(d, umma_desc(a_hi + o), umma_desc(b_lo + o), idesc, 1);
              umma2_f16(d, umma_desc(a_hi + o), umma_desc(b_hi + o), idesc, 1);
            } else {
              umma2_tf32(d, umma_desc(a_lo + o), umma_desc(b_hi + o), idesc, (kin | k) != 0);
              umma2_tf32(d, umma_desc(a_hi + o), umma_desc(b_lo + o), idesc, 1);
              umma2_tf32(d, umma_desc(a_hi + o), umma_desc(b_hi + o), idesc, 1);
            }
          }
          umma2_commit(&empty[st]);
          if (++kin == P.chunk || kb == nk - 1) {
            umma2_commit(&tfull[buf]);
            ++gch;
            kin = 0;
          }
        }
      }
    }
  } else if (warp >= 4) {  // ===== epilogue warps of both CTAs =====
    uint32_t te[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(te[i]) : "r"(s_u32(&tempty[i])), "r"(0));
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int row = quad * 32 + lane;
    const int nchunks = (nk + P.chunk - 1) / P.chunk;
    float* Sw = S + (warp - 4) * (32 * TP_SLD);          // this warp's park
    long* rc = rows + (warp - 4) * 64;                   // its rows' output offsets (-1: discarded) and residual offsets
    long* rr = rc + 32;
    constexpr int EC = TC_BN / 2;                        // 128 columns per warp: [half * 128, +128)
    uint32_t gch = 0;
    for (int tile = pair; tile < ntiles; tile += npairs) {
      const int nt = tile % P.cn, mp = tile / P.cn;
      const int j0 = (2 * mp + (int)rank) * TC_BM, n0 = nt * TC_BN;
      float acc[EC];
#pragma unroll
      for (int i = 0; i < EC; ++i) acc[i] = 0.f;
      for (int c = 0; c < nchunks; ++c, ++gch) {
        const uint32_t buf = gch & 1;
        mb_wait(&tfull[buf], (gch >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int cc = 0; cc < EC / 16; ++cc) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + buf * TC_BN + half * EC + cc * 16, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[cc * 16 + i] += __uint_as_float(v[i]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(te[buf]) : "memory");
      }
      {  // GEMM row j <-> padded input row j*stride -> (batch, output step)
        const int j = j0 + row;
        long co = -1, ro = 0;
        if (j < P.Rs) {
          const long in_row = (long)j * P.stride;
          const int bb = (int)(in_row / P.rows_in);
          const int tin = (int)(in_row - (long)bb * P.rows_in) - P.off;
          if (bb < P.nbatch && tin >= 0 && (tin % P.stride) == 0 && tin / P.stride < P.T_out) {
            const int t = tin / P.stride;
            co = (long)bb * P.c_bs + (long)t * P.c_rs;
            ro = (long)bb * P.r_bs + (long)t * P.r_rs;
          }
        }
        rc[lane] = co;
        rr[lane] = ro;
      }
      __syncwarp();
      if (P.act == ACT_NONE) tp_tile<ACT_NONE>(P, acc, Sw, rc, rr, n0 + half * EC, lane);
      else if (P.act == ACT_GELU) tp_tile<ACT_GELU>(P, acc, Sw, rc, rr, n0 + half * EC, lane);
      else tp_tile<-1>(P, acc, Sw, rc, rr, n0 + half * EC, lane);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * TC_BN) : "memory");
  }
}

// ---- split kernels ------------------------------------------------------------------------------
__global__ void split_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float v = x[i];
    float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    hi[i] = h;
    lo[i] = v - h;
  }
}
void split_hi_lo(ts_engine* e, const float* x, float* hi, float* lo, long n, cudaStream_t s) {
  if (e->ws.sizing || n <= 0) return;
  split_kernel<<<(int)std::min<long>((n + 255) / 256, 148 * 16), 256, 0, s>>>(x, hi, lo, n);
  e->launches++;
  TS_CUDA(cudaGetLastError());
}
__global__ void split16_kernel(const float* __restrict__ x, unsigned short* __restrict__ h, unsigned short* __restrict__ l, long n, float scale) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) split16(x[i] * scale, h[i], l[i]);
}
// fp16 planes of W * 2^shift; the shift puts max|W| just below 2^14 (clamped to [0, 14]) and comes back as 2^-shift
static float split16_host(const std::vector<float>& w, std::vector<unsigned short>* h, std::vector<unsigned short>* l) {
  float mx = 0.f;
  for (float v : w) mx = std::max(mx, std::fabs(v));
  int shift = 0;
  if (mx > 0.f && std::isfinite(mx)) {
    int ex;
    std::frexp(mx, &ex);            // mx = m * 2^ex, m in [0.5, 1)
    shift = std::min(14, std::max(0, 14 - ex));
  }
  const float sc = std::ldexp(1.0f, shift);
  h->resize(w.size());
  l->resize(w.size());
  for (size_t i = 0; i < w.size(); ++i) {
    const float v = w[i] * sc;
    const __half hh = __float2half_rn(v);
    const __half ll = __float2half_rn(v - __half2float(hh));
    (*h)[i] = __half_as_ushort(hh);
    (*l)[i] = __half_as_ushort(ll);
  }
  return std::ldexp(1.0f, -shift);
}
void split_host(const std::vector<float>& w, std::vector<float>* hi, std::vector<float>* lo) {
  hi->resize(w.size());
  lo->resize(w.size());
  for (size_t i = 0; i < w.size(); ++i) {
    uint32_t u;
    memcpy(&u, &w[i], 4);
    u &= 0xffffe000u;
    float h;
    memcpy(&h, &u, 4);
    (*hi)[i] = h;
    (*lo)[i] = w[i] - h;
  }
}

// ---- host launcher -------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    TS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) fail(TS_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    fn = (EncodeTiledFn)p;
  }
  return fn;
}
static CUtensorMap make_map(const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
                            bool f16 = false) {
  CUtensorMap m;
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode()(&m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, (void*)base, dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(TS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rank %d dims %llu,%llu", (int)r, rank, (unsigned long long)dims[0],
                              (unsigned long long)dims[1]);
  return m;
}

bool tc_conv_supported(ts_engine* e, const Layer& L, const Act3& x, int stride, int pd) {
  const bool onchip = e->tc_pair && e->tc_onchip;     // plain operands, split on chip (CTA-pair kernel only)
  const bool f16 = e->tc_pair && e->tc_f16;           // fp16 planes
  if (f16) {
    if (!L.W_h16 || !x.h16 || x.C % (2 * TC_BK)) return false;
  } else {
    if (onchip ? (!L.W || x.split) : (!L.W_hi || !x.lo)) return false;
    if (x.C % TC_BK) return false;
  }
  const int rows_in = x.T + 2 * x.pad + x.tail;
  if (rows_in % stride || (x.pad - pd) % stride || x.pad < pd) return false;
  return true;
}

// y(b, t*y_tmul + y_toff, :) = act( conv(x)(b,t,:) + bias + res(b,t,:) );  x and W must be split (hi/lo)
void tc_conv1d(ts_engine* e, const Layer& L, const Act3& x, int k, int stride, int pd, const Act3& y, int T_out, int act, const Act3* res,
               cudaStream_t s, int y_tmul, int y_toff, int coff) {
  if (e->ws.sizing) return;
  if (!tc_conv_supported(e, L, x, stride, pd)) fail(TS_ERR_INVALID, "tc_conv1d: unsupported geometry");
  if (!y.p && !y.h16) fail(TS_ERR_INVALID, "tc_conv1d: output without storage");
  const bool onchip = e->tc_pair && e->tc_onchip;
  const bool f16 = e->tc_pair && e->tc_f16;
  const int esz = f16 ? 2 : 4, bk = f16 ? 2 * TC_BK : TC_BK;
  if (L.taps != k || L.cin != x.C) fail(TS_ERR_INVALID, "tc_conv1d: layer/input mismatch");
  const int rows_in = x.T + 2 * x.pad + x.tail;
  const long R = (long)x.B * rows_in;
  const long Rs = R / stride;
  const void* base_hi = f16 ? (const void*)x.h16 : (const void*)x.p;   // first padded row of batch 0 (allocation start)
  const void* base_lo = f16 ? (const void*)x.l16 : onchip ? (const void*)x.p : (const void*)x.lo;
  cuuint64_t adims[3] = {(cuuint64_t)x.C, (cuuint64_t)stride, (cuuint64_t)Rs};
  cuuint64_t astr[2] = {(cuuint64_t)x.C * esz, (cuuint64_t)x.C * esz * stride};
  // cluster shape: up to 4 N-tiles x 2 M-tiles share their operand boxes by TMA multicast
  const int tiles_n = (L.N + TC_BN - 1) / TC_BN, tiles_m = (int)((Rs + TC_BM - 1) / TC_BM);
  int cn = 1, cm = 1;
  if (e->tc_multicast) {
    cn = (tiles_n % 4 == 0) ? 4 : (tiles_n % 2 == 0) ? 2 : 1;
    cm = tiles_m >= 2 ? 2 : 1;
  }
  const bool pair = e->tc_pair;   // CTA-pair kernel: (1,2) cluster, each CTA loads half of the B tile
  if (pair) { cn = 1; cm = 2; }
  cuuint32_t abox[3] = {(cuuint32_t)bk, 1, (cuuint32_t)(pair ? TC_BM : TC_BM / cn)};
  CUtensorMap mAh = make_map(base_hi, 3, adims, astr, abox, f16), mAl = make_map(base_lo, 3, adims, astr, abox, f16);
  cuuint64_t bdims[2] = {(cuuint64_t)L.K, (cuuint64_t)L.N};
  cuuint64_t bstr[1] = {(cuuint64_t)L.K * esz};
  cuuint32_t bbox[2] = {(cuuint32_t)bk, (cuuint32_t)(TC_BN / cm)};   // pair: cm == 2 -> 128-row halves
  const void* wh = f16 ? (const void*)L.W_h16 : onchip ? (const void*)L.W : (const void*)L.W_hi;
  const void* wl = f16 ? (const void*)L.W_l16 : onchip ? (const void*)L.W : (const void*)L.W_lo;
  CUtensorMap mBh = make_map(wh, 2, bdims, bstr, bbox, f16), mBl = make_map(wl, 2, bdims, bstr, bbox, f16);
  TcArgs P;
  P.chunk = 256 / bk;                  // K = 256 per TMEM accumulation chunk
  P.oscale = f16 ? L.w_unscale : 1.f;
  P.c_h16 = y.h16 ? y.row_h16(0, y_toff) + coff : nullptr;
  P.c_l16 = y.h16 ? y.row_l16(0, y_toff) + coff : nullptr;
  P.taps = k; P.cblocks = x.C / bk; P.stride = stride; P.C = x.C;
  P.rows_in = rows_in; P.off = x.pad - pd; P.T_out = T_out; P.nbatch = x.B; P.Rs = (int)Rs; P.N = L.N;
  P.c_hi = y.p ? y.row(0, y_toff) + coff : nullptr; P.c_lo = y.lo ? y.row_lo(0, y_toff) + coff : nullptr;
  P.c_bs = y.bstride(); P.c_rs = (long)y_tmul * y.C;
  P.bias = L.bias;
  P.r_hi = res ? res->row(0, 0) : nullptr;
  P.r_lo = (res && res->lo) ? res->row_lo(0, 0) : nullptr;
  P.r_bs = res ? res->bstride() : 0; P.r_rs = res ? res->C : 0;
  P.act = act;
  P.cn = pair ? tiles_n : cn; P.cm = cm;
  { const char* np = getenv("TS_TC_NPROD"); P.nprod = (np && np[0] == '1') ? 1 : (np && np[0] == '4') ? 4 : 3; }
  P.a_hi = (const float*)base_hi; P.a_lo = (const float*)base_lo; P.a_rows = R;
  // whole-row L2 prefetch measured slower (96 vs 89 ms per face pass): off unless TS_TC_ROWPF=1
  { const char* pf = getenv("TS_TC_ROWPF"); P.prefetch_rows = (pf && pf[0] == '1' && !f16) ? 1 : 0; }
  if (!e->tc_attr_set) {   // the max-dynamic-smem attribute is per device: cached per engine, not per process
    TS_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    TS_CUDA(cudaFuncSetAttribute(tc2_gemm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM));
    TS_CUDA(cudaFuncSetAttribute(tc2_gemm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM));
    TS_CUDA(cudaFuncSetAttribute(tc2_gemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM));
    TS_CUDA(cudaFuncSetAttribute(tc2p_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM));
    TS_CUDA(cudaFuncSetAttribute(tc2p_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM));
    e->tc_attr_set = true;
  }
  // grid padded to whole clusters; surplus tiles fall outside Rs / N and are masked (TMA zero-fills OOB)
  dim3 grid((unsigned)(((tiles_n + cn - 1) / cn) * cn), (unsigned)(((tiles_m + cm - 1) / cm) * cm));
  if (pair) grid = dim3((unsigned)(2 * tiles_n * ((tiles_m + 1) / 2)), 1, 1);
  // TS_TC_PERSIST=0 switches it off (A/B): face alone at 64 clips 31.0 vs 31.3 ms, whole step next to the sampler 51.0 vs 52.1 ms
  static const bool persist_env = !(getenv("TS_TC_PERSIST") && getenv("TS_TC_PERSIST")[0] == '0');
  const int ntiles = tiles_n * ((tiles_m + 1) / 2);
  // persistent CTA pairs (one per TPC) looping over the tiles -- when every pair gets at least two tiles; below that there is
  // nothing to overlap and the block-wide write-out of the one-tile kernel is the faster one (face, 8 clips: 5.7 vs 5.9 ms)
  const bool persist = pair && !onchip && persist_env && ntiles >= 2 * std::max(1, e->sm_count / 2);
  if (persist) grid = dim3((unsigned)(2 * std::min(ntiles, std::max(1, e->sm_count / 2))), 1, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = persist ? TP_SMEM : pair ? T2_SMEM : TC_SMEM;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = pair ? 2 : cn; at[0].val.clusterDim.y = pair ? 1 : cm; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (persist && f16) TS_CUDA(cudaLaunchKernelEx(&cfg, tc2p_gemm_kernel<true>, mAh, mAl, mBh, mBl, P, ntiles));
  else if (persist) TS_CUDA(cudaLaunchKernelEx(&cfg, tc2p_gemm_kernel<false>, mAh, mAl, mBh, mBl, P, ntiles));
  else if (f16) TS_CUDA(cudaLaunchKernelEx(&cfg, tc2_gemm_kernel<false, true>, mAh, mAl, mBh, mBl, P));
  else if (pair && onchip) TS_CUDA(cudaLaunchKernelEx(&cfg, tc2_gemm_kernel<true, false>, mAh, mAl, mBh, mBl, P));
  else if (pair) TS_CUDA(cudaLaunchKernelEx(&cfg, tc2_gemm_kernel<false, false>, mAh, mAl, mBh, mBl, P));
  else TS_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel, mAh, mAl, mBh, mBl, P));
  e->launches++;
  TS_CUDA(cudaGetLastError());
}

void upload_weights(ts_engine* e, const std::vector<float>& W, Layer* L) {
  L->W = e->upload(W);
  std::vector<float> hi, lo;
  split_host(W, &hi, &lo);
  L->W_hi = e->upload(hi);
  L->W_lo = e->upload(lo);
  std::vector<unsigned short> h16, l16;
  L->w_unscale = split16_host(W, &h16, &l16);
  L->W_h16 = e->upload(h16);
  L->W_l16 = e->upload(l16);
}

void conv_auto(ts_engine* e, const Layer& L, const Act3& x, int k, int stride, int pd, const Act3& y, int T_out, int act, const Act3* res,
               cudaStream_t s, int y_tmul, int y_toff, int coff) {
  if (e->use_tc && (y.C % 4) == 0 && (coff % 4) == 0 && tc_conv_supported(e, L, x, stride, pd)) tc_conv1d(e, L, x, k, stride, pd, y, T_out, act, res, s, y_tmul, y_toff, coff);
  else conv1d(e, L, x, k, stride, pd, y, T_out, act, res, s, y_tmul, y_toff, 0, coff);
}

}  // namespace ts

using namespace ts;

extern "C" int ts_set_tensor_cores(ts_engine* e, int enable) {
  if (!e) return TS_ERR_INVALID;
  e->use_tc = enable != 0;
  e->tc_pair = enable == 1 || enable == 3 || enable == 5 || enable == 6;   // 1 / 3 = CTA-pair (cta_group::2) 256x256 kernel, 3xTF32
  e->tc_f16 = enable == 6;                   // 6 (default) = the CTA-pair kernel on fp16-split operands (kind::f16, see the file header)
  e->tc_multicast = enable == 2;             // 2 = single-CTA 128x256 kernel in (n x 2) clusters with TMA multicast
                                             // 4 = single-CTA 128x256 kernel, no cluster
  e->tc_onchip = enable == 5;                // 5 = CTA-pair kernel on PLAIN operands, hi / lo split in shared memory (experiment)
  return TS_OK;
}

// debug entry: dense C[M,N] = act(A[M,K] W[N,K]^T + bias), mode 0 = FFMA kernel, 1 = tcgen05 3xTF32
extern "C" int ts_debug_gemm(ts_engine* e, int mode, const float* A, const float* W, const float* bias, float* C, int M, int N, int K, int act,
                             void* stream) {
  TS_API_BEGIN(e)
  cudaStream_t s = (cudaStream_t)stream;
  if (mode == 0) {
    GemmP p;
    p.A = A; p.W = W; p.bias = bias; p.C = C; p.M = M; p.N = N; p.K = K; p.mper = M; p.a_rs = K; p.kc = K; p.a_ts = K; p.c_rs = N; p.act = act; p.ldw = K;
    e->ws.sizing = false;
    launch_gemm(e, p, s);
  } else if (mode == 2) {
    if (K % TC_BK) fail(TS_ERR_INVALID, "ts_debug_gemm: K must be a multiple of 32 for the tensor-core path");
    if (!(e->tc_pair && e->tc_onchip)) fail(TS_ERR_INVALID, "ts_debug_gemm mode 2 needs ts_set_tensor_cores(e, 5)");
    e->ws.sizing = false;
    Layer L;
    L.N = N; L.K = K; L.taps = 1; L.cin = K; L.W = const_cast<float*>(W); L.bias = const_cast<float*>(bias);
    Act3 x; x.p = const_cast<float*>(A); x.split = false; x.B = 1; x.T = M; x.C = K; x.pad = 0;
    Act3 y; y.p = C; y.B = 1; y.T = M; y.C = N; y.pad = 0;
    tc_conv1d(e, L, x, 1, 1, 0, y, M, act, nullptr, s, 1, 0, 0);
  } else if (mode == 3) {
    if (K % (2 * TC_BK)) fail(TS_ERR_INVALID, "ts_debug_gemm: K must be a multiple of 64 for the fp16-split path");
    const bool f0 = e->tc_f16, p0 = e->tc_pair, oc = e->tc_onchip;
    e->tc_f16 = true; e->tc_pair = true; e->tc_onchip = false;
    e->ws.sizing = false;
    unsigned short *ah, *al, *wh, *wl;
    float* full;
    TS_CUDA(cudaMalloc(&ah, (size_t)M * K * 2)); TS_CUDA(cudaMalloc(&al, (size_t)M * K * 2));
    TS_CUDA(cudaMalloc(&wh, (size_t)N * K * 2)); TS_CUDA(cudaMalloc(&wl, (size_t)N * K * 2));
    TS_CUDA(cudaMalloc(&full, 16));
    split16_kernel<<<148 * 8, 256, 0, s>>>(A, ah, al, (long)M * K, 1.f);
    split16_kernel<<<148 * 8, 256, 0, s>>>(W, wh, wl, (long)N * K, 256.f);
    Layer L;
    L.N = N; L.K = K; L.taps = 1; L.cin = K; L.W_h16 = wh; L.W_l16 = wl; L.w_unscale = 1.f / 256.f; L.bias = const_cast<float*>(bias);
    Act3 x; x.p = const_cast<float*>(A); x.h16 = ah; x.l16 = al; x.split = true; x.B = 1; x.T = M; x.C = K; x.pad = 0;
    Act3 y; y.p = C; y.B = 1; y.T = M; y.C = N; y.pad = 0;
    tc_conv1d(e, L, x, 1, 1, 0, y, M, act, nullptr, s, 1, 0, 0);
    e->tc_f16 = f0; e->tc_pair = p0; e->tc_onchip = oc;
    TS_CUDA(cudaStreamSynchronize(s));
    cudaFree(ah); cudaFree(al); cudaFree(wh); cudaFree(wl); cudaFree(full);
  } else {
    if (K % TC_BK) fail(TS_ERR_INVALID, "ts_debug_gemm: K must be a multiple of 32 for the tensor-core path");
    const bool oc = e->tc_onchip, f0 = e->tc_f16;
    e->tc_f16 = false;
    e->tc_onchip = false;
    e->ws.sizing = false;
    float *ah, *al, *wh, *wl;
    TS_CUDA(cudaMalloc(&ah, (size_t)M * K * 4)); TS_CUDA(cudaMalloc(&al, (size_t)M * K * 4));
    TS_CUDA(cudaMalloc(&wh, (size_t)N * K * 4)); TS_CUDA(cudaMalloc(&wl, (size_t)N * K * 4));
    split_hi_lo(e, A, ah, al, (long)M * K, s);
    split_hi_lo(e, W, wh, wl, (long)N * K, s);
    Layer L;
    L.N = N; L.K = K; L.taps = 1; L.cin = K; L.W_hi = wh; L.W_lo = wl; L.bias = const_cast<float*>(bias);
    Act3 x; x.p = ah; x.lo = al; x.split = true; x.B = 1; x.T = M; x.C = K; x.pad = 0;
    Act3 y; y.p = C; y.B = 1; y.T = M; y.C = N; y.pad = 0;
    tc_conv1d(e, L, x, 1, 1, 0, y, M, act, nullptr, s, 1, 0, 0);
    e->tc_onchip = oc; e->tc_f16 = f0;
    TS_CUDA(cudaStreamSynchronize(s));
    cudaFree(ah); cudaFree(al); cudaFree(wh); cudaFree(wl);
  }
  TS_API_END(e)
}
