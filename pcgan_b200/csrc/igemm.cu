// Implicit-GEMM convolution engine for sm_100a: TMA -> 128B-swizzled shared memory ->
// tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> tcgen05.ld epilogue.
//
// One persistent, warp-specialised kernel, two operand arrangements:
//   KMAJOR (forward / data-gradient): A = activation box [<=128 pixels][64 ch] per
//     (tap, channel chunk), B = weights [block_n][64]; both K-major.
//   WGRAD (weight-gradient): A = dY box [64 pixels][128 cout], B = X box
//     [64 pixels][block_n cin]; both MN-major, K = pixels, split across CTAs.
// Everything that distinguishes one convolution from another (padding mode, stride,
// transposed phases, reflect halos, packed stems) is data in pcgan_igemm_desc.
//
// 384 threads, 1 CTA / SM.  Warps 4-7 and 8-11 are two epilogue groups (TMEM lane quarter = warp % 4); warp 2
// allocates TMEM.  Two instantiations of the kernel:
//   igemm_kernel<true>  (desc.pair, full-width N tiles): clusters of two CTAs on adjacent M tiles of the same N tile.
//     Warp 0 of each CTA loads its own A tile and half of the B tile, warp 1 of the leader issues one
//     tcgen05.mma.cta_group::2 of M = 256 per K step for both (operand traffic through shared memory is halved per
//     SM), tiles alternate between two 256-column accumulators, each CTA's epilogue groups drain their own 128 rows.
//   igemm_kernel<false> (everything narrower): the single-thread issue loops, not the tensor pipe, bound small tiles,
//     so the CTA runs two independent pipelines: warps 0 -> 1 -> group 0 on its even tiles, warps 3 -> 2 -> group 1 on
//     its odd tiles, each with half of the ring, its own barriers and (N <= 128) a double-buffered accumulator.
// Barriers: ring full / empty (TMA <-> MMA), accumulator full / empty (MMA <-> epilogue); persistent tile loop over a
// static contiguous schedule.
#include <cuda.h>
#include <mutex>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>

#include "common.cuh"

namespace pcgan {

static constexpr int kMaxStages = 16;
static constexpr int kDataBytes = 192 * 1024;      // operand ring: num_stages x (A + B) chosen per plan
static constexpr int kBoxBytesMN = 64 * 128;       // one MN-major box: 64 K-rows x 128 B
static constexpr int kTmemCols = 512;              // 2 accumulators x 256 fp32 columns (or 4 x 128: DevParams::acc_sub)
static constexpr int kAccCols = 256;
static constexpr int kNumThreads = 384;            // warps 0-3: TMA, MMA, TMEM alloc (+ MMA 2), TMA 2; 4-7 and 8-11: two epilogue groups
static constexpr int kEpiGroups = 2;               // group g drains accumulator g (every other tile of the CTA)
// tail after the ring: barriers (512 B) | bias [2 groups][256] f32 | per-warp scratch [8 warps][32][17] f32: the
// statistics partial sums [2][256] of each epilogue warp, or the row-exchange tile of the shift-sum epilogue
static constexpr int kTailBias = 512;
static constexpr int kTailTr = kTailBias + kEpiGroups * 256 * 4;
static constexpr int kWarpScratch = 32 * 17 * 4;   // bytes per epilogue warp (>= 2 * 256 * 4)
static constexpr int kTailBytes = kTailTr + 8 * kWarpScratch;
static_assert(kDataBytes + 1024 + kTailBytes <= 227 * 1024, "shared memory budget");
static constexpr int kSmemBytes = kDataBytes + 1024 /*align*/ + kTailBytes;

// Division by a divisor that is fixed per plan, for 0 <= x < 2^31: q = (umulhi(x, mul) + x) >> shift with
// shift = ceil(log2 d), mul = floor(2^32 (2^shift - d) / d) + 1 (Granlund-Montgomery).  The per-tile coordinate arithmetic
// of every role (digits of the tile index, components of the output coordinate) is on the critical path of tiles with
// little work: a hardware integer division is ~20 dependent instructions, this is three.
struct FastDiv {
  uint32_t mul, shift;
};
static FastDiv make_fastdiv(int32_t d) {
  FastDiv f{1u, 0u};
  if (d <= 1) return f;
  uint32_t sh = 0;
  while ((1ull << sh) < static_cast<uint64_t>(d)) ++sh;
  f.shift = sh;
  f.mul = static_cast<uint32_t>((((1ull << sh) - static_cast<uint64_t>(d)) << 32) / static_cast<uint64_t>(d)) + 1u;
  return f;
}
__device__ __forceinline__ int32_t fdiv(int32_t x, const FastDiv& f) {
  const uint32_t u = static_cast<uint32_t>(x);
  return static_cast<int32_t>((__umulhi(u, f.mul) + u) >> f.shift);
}
// the same arithmetic on the host (pcgan_selftest_fastdiv)
static int32_t fdiv_host(int32_t x, const FastDiv& f) {
  const uint32_t u = static_cast<uint32_t>(x);
  const uint32_t hi = static_cast<uint32_t>((static_cast<uint64_t>(u) * f.mul) >> 32);
  return static_cast<int32_t>((hi + u) >> f.shift);
}

struct DevParams {
  int32_t kind, block_n, a_rows, a_ch;
  int32_t num_stages, stage_bytes, a_alloc;
  int32_t a_window;    // KMAJOR: A is the raw 8-channel pixel row, read through an overlapping no-swizzle descriptor
  int32_t a_bytes;     // bytes one A box brings into a stage
  int32_t b_res;       // KMAJOR, > 0: the whole B operand stays resident in front of the ring, b_res bytes per K chunk
  int32_t ring_off;    // byte offset of the ring behind the resident B operand
  int32_t wg_box_dim;  // WGRAD: the 64-column boxes of an N tile step along this B tensor dim (filter rows in N), 0 = channels
  int32_t tf32;        // fp32 operands, tcgen05.mma.kind::tf32
  int32_t kc;          // elements of one 128-byte K chunk: 64 (bf16) or 32 (tf32)
  int32_t split_prod;  // paired WGRAD: warp 3 issues the X boxes, warp 0 the dY boxes
  int32_t kps;         // KMAJOR, unpaired: K chunks per ring stage (one barrier round trip and one commit for all of them)
  int32_t sub_bytes;   //   bytes of one chunk's slot inside a stage (stage_bytes = kps * sub_bytes)
  int32_t acc_sub;     // 1 (dual, block_n <= 128): each pipeline double-buffers its accumulator in two 128-column halves
  int32_t dual;        // 1: two independent producer -> MMA -> epilogue pipelines (even / odd tiles of the CTA), each with
                       //    num_stages stages of the ring and one TMEM accumulator
  int32_t t_count[4];
  FastDiv fd_t[4], fd_nt, fd_p1[4], fd_p2[4], fd_sdiv;   // t_count, n_tiles, e_p1, e_p2, stats_div
  int32_t a_base[4], a_step[4][4];
  int32_t b_base[4], b_step[4][4];
  int32_t n_tiles, m_tiles, ksplit, num_m_tiles, total_tiles;
  int32_t shift_taps, shift_cpad;   // shift-sum epilogue (include/pcgan_kernels.h)
  int32_t pair;        // 1: clusters of 2 CTAs on two M tiles of the same N tile, one tcgen05.mma.cta_group::2 for both
  int32_t sched_items; // work items per CTA slot schedule: tiles, or pair-tiles when pair
  int32_t num_taps, cchunks;
  int32_t tap_off[PCGAN_MAX_TAPS][4];
  int32_t tap_c0[PCGAN_MAX_TAPS];
  int32_t tap_bk[PCGAN_MAX_TAPS];
  int32_t box[4];
  int32_t e_base[4], e_step[4][4], e_p1[4], e_p2[4];
  pcgan_comp e_comp[4][3];
  int32_t out_dtype, act;
  float act_slope;
  int32_t n_valid;
  int64_t out_cstride;
  int32_t stats_mode, stats_dim, stats_comp, stats_div;
  int32_t m_valid, wg_ncols;
  int64_t ldo;
  void* out;
  const float* bias;
  float* stats;
};

struct Digits {
  int32_t t[4];
};
__device__ __forceinline__ Digits decompose(int32_t idx, const int32_t (&count)[4], const FastDiv (&fd)[4]) {
  Digits d;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int32_t q = fdiv(idx, fd[j]);
    d.t[j] = idx - q * count[j];
    idx = q;
  }
  return d;
}
__device__ __forceinline__ int32_t coord(const Digits& d, const int32_t (&base)[4], const int32_t (&step)[4][4],
                                         int dim) {
  return base[dim] + d.t[0] * step[0][dim] + d.t[1] * step[1][dim] + d.t[2] * step[2][dim] + d.t[3] * step[3][dim];
}

// Statistics group of a tile: the chosen component of the first row's coordinate along stats_dim.
__device__ __forceinline__ int32_t tile_group(const DevParams& P, const Digits& d) {
  if (P.stats_dim < 0) return 0;
  const int dim = P.stats_dim;
  const int32_t g0 = coord(d, P.e_base, P.e_step, dim);
  int32_t s0 = 0, srem = g0, s1;
  if (g0 < 0) {   // never for a plan with statistics; keeps the exact semantics if it ever happens
    if (P.e_p1[dim] > 0) { s0 = g0 / P.e_p1[dim]; srem = g0 % P.e_p1[dim]; }
    if (P.e_p2[dim] > 0) { s1 = srem / P.e_p2[dim]; } else { s1 = srem; }
    const int32_t smp_ = P.stats_comp == 0 ? s0 : s1;
    return P.stats_div > 1 ? smp_ / P.stats_div : smp_;
  }
  if (P.e_p1[dim] > 0) { s0 = fdiv(g0, P.fd_p1[dim]); srem = g0 - s0 * P.e_p1[dim]; }
  if (P.e_p2[dim] > 0) { s1 = fdiv(srem, P.fd_p2[dim]); } else { s1 = srem; }
  const int32_t smp = P.stats_comp == 0 ? s0 : s1;
  return P.stats_div > 1 ? fdiv(smp, P.fd_sdiv) : smp;
}

// ------------------------------------------------------------------ schedule
// A CTA (or CTA pair) owns a contiguous range of work items.  Unpaired: item = tile.  Paired: item = two M tiles
// (2*m2 + rank) of the same N tile [, tap, K split]; the odd one out of an odd M-tile count is computed twice and
// dropped (`valid` false).
struct Sched {
  int32_t begin, end;
  uint32_t rank;   // CTA rank in the pair (0 when unpaired)
};
__device__ __forceinline__ Sched make_sched(const DevParams& P) {
  Sched s;
  const int32_t slots = P.pair ? static_cast<int32_t>(gridDim.x >> 1) : static_cast<int32_t>(gridDim.x);
  const int32_t slot = P.pair ? static_cast<int32_t>(blockIdx.x >> 1) : static_cast<int32_t>(blockIdx.x);
  s.begin = static_cast<int32_t>(static_cast<int64_t>(P.sched_items) * slot / slots);
  s.end = static_cast<int32_t>(static_cast<int64_t>(P.sched_items) * (slot + 1) / slots);
  s.rank = P.pair ? cluster_ctarank() : 0u;
  return s;
}
__device__ __forceinline__ void kmajor_item(const DevParams& P, const Sched& sc, int32_t item, int32_t& mt, int32_t& nt, bool& valid) {
  const int32_t q = fdiv(item, P.fd_nt);
  nt = item - q * P.n_tiles;
  mt = P.pair ? 2 * q + static_cast<int32_t>(sc.rank) : q;
  valid = mt < P.num_m_tiles;
  if (!valid) mt = P.num_m_tiles - 1;
}
__device__ __forceinline__ void wgrad_item(const DevParams& P, const Sched& sc, int32_t item, int32_t& ks, int32_t& nt, int32_t& mt,
                                           int32_t& tap) {
  int32_t t = item;
  ks = t % P.ksplit; t /= P.ksplit;
  nt = t % P.n_tiles; t /= P.n_tiles;
  const int32_t mdiv = P.pair ? (P.m_tiles >> 1) : P.m_tiles;
  const int32_t m2 = t % mdiv; t /= mdiv;
  mt = P.pair ? 2 * m2 + static_cast<int32_t>(sc.rank) : m2;
  tap = t;
}

struct EpiShared {
  uint32_t s_bias;     // shared address of this group's bias [256]
  uint32_t s_tr;       // shared address of this warp's scratch: statistics partial sums [2][256] or exchange tile [32][17]
  uint64_t* tmem_full; // barriers of this group's accumulator(s): [2], the second one used when P.acc_sub
  uint64_t* tmem_empty;
  uint32_t group;      // 0 / 1: also the TMEM stage and the parity of the CTA-local tile index it handles
};

// The accumulator stage is drained: tell the MMA warp (the pair's leader's when paired).
__device__ __forceinline__ void epi_release(const DevParams& P, const EpiShared& es, uint32_t sub) {
  if (P.pair) mbar_arrive_leader(es.tmem_empty + sub);
  else mbar_arrive(es.tmem_empty + sub);
}
// Accumulator of a group's j-th tile: which half (0 unless P.acc_sub) and the parity of its barriers' phase.
__device__ __forceinline__ void acc_slot(const DevParams& P, uint32_t j, uint32_t& sub, uint32_t& phase) {
  sub = P.acc_sub ? (j & 1u) : 0u;
  phase = P.acc_sub ? ((j >> 1) & 1u) : (j & 1u);
}

template <int ACT>
__device__ __forceinline__ float act_ct(float x, float slope) {
  if (ACT == PCGAN_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == PCGAN_ACT_LRELU) return x > 0.f ? x : x * slope;
  if (ACT == PCGAN_ACT_TANH) return tanhf(x);
  if (ACT == PCGAN_ACT_SIGMOID) return 1.f / (1.f + expf(-x));
  return x;
}

// Forward / data-gradient epilogue of one epilogue group (4 warps = 128 accumulator rows): TMEM -> registers ->
// (+bias, statistics, activation) -> global.  Thread = one accumulator row (output pixel), 32 fp32 columns per
// tcgen05.ld.  The two groups of a CTA alternate tiles, each on its own TMEM stage.
// x[i] = value of column i in this lane's row (zero for rows that are not stored; squared if SQ).  Returns, in lane l,
// the sum over the warp's 32 rows of column l: at each step a lane keeps the half of its columns that matches its lane
// bit and receives the partner's partial sums of that half.
template <bool SQ>
__device__ __forceinline__ float warp_column_sums(const float (&v)[32], bool valid, uint32_t lane) {
  float a[16];
  {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float lo = valid ? v[i] : 0.f, hi = valid ? v[i + 16] : 0.f;
      if (SQ) { lo *= lo; hi *= hi; }
      a[i] = (up ? hi : lo) + __shfl_xor_sync(0xffffffffu, up ? lo : hi, 16);
    }
  }
  float b[8];
  {
    const bool up = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = (up ? a[i + 8] : a[i]) + __shfl_xor_sync(0xffffffffu, up ? a[i] : a[i + 8], 8);
  }
  float c[4];
  {
    const bool up = (lane & 4) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = (up ? b[i + 4] : b[i]) + __shfl_xor_sync(0xffffffffu, up ? b[i] : b[i + 4], 4);
  }
  float d[2];
  {
    const bool up = (lane & 2) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) d[i] = (up ? c[i + 2] : c[i]) + __shfl_xor_sync(0xffffffffu, up ? c[i] : c[i + 2], 2);
  }
  const bool up = (lane & 1) != 0;
  return (up ? d[1] : d[0]) + __shfl_xor_sync(0xffffffffu, up ? d[0] : d[1], 1);
}

template <int ACT, bool BF16, bool STATS>
__device__ __forceinline__ void epilogue_kmajor(const DevParams& P, const EpiShared es, uint32_t tmem_base, const Sched sch) {
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q = warp & 3;
  const uint32_t row = q * 32 + lane;
  const uint32_t et = (threadIdx.x - 128) & 127;
  const uint32_t bar_stats = 1 + es.group, bar_bias = 3 + es.group;
  const uint32_t my_part = es.s_tr;                      // this warp's column sums [256] | sums of squares [256]
  const uint32_t group_part = es.s_tr - q * kWarpScratch;   // the group's four warps are contiguous
  const int32_t block_n = P.block_n, n_valid = P.n_valid;
  const int64_t cs = P.out_cstride;
  const float* bias = P.bias;
  const float slope = P.act_slope;
  const uint32_t s_bias = es.s_bias;
  uint32_t jt = 0;   // tiles this group has drained
  int32_t cur_nt = -1, cur_group = -1;

  // box-local index of this row along the four outer box dims (tile-invariant)
  int32_t il[4];
  {
    uint32_t r = row;
#pragma unroll
    for (int dim = 0; dim < 4; ++dim) {
      const int32_t bx = P.box[dim];
      il[dim] = 0;
      if (bx > 1) { il[dim] = r % bx; r /= bx; }
    }
  }
  const bool row_in_box = row < static_cast<uint32_t>(P.a_rows);

  auto flush_stats = [&](int32_t group, int32_t nt) {
    named_bar_sync(bar_stats, 128);
    float* gs = P.stats + static_cast<int64_t>(group) * n_valid * 2;
    for (int32_t i = et; i < 512; i += 128) {
      const int32_t col = i & 255, which = i >> 8;
      if (col >= block_n) continue;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint32_t a = group_part + w * kWarpScratch + i * 4;
        v += lds_f32(a);
        sts_f32(a, 0.f);
      }
      const int32_t ch = nt * block_n + col;
      if (ch < n_valid && v != 0.f) atomicAdd(gs + ch * 2 + which, v);
    }
    named_bar_sync(bar_stats, 128);
  };

  for (int32_t tile = sch.begin + static_cast<int32_t>(es.group); tile < sch.end; tile += kEpiGroups) {
    int32_t mt, nt;
    bool tile_valid;
    kmajor_item(P, sch, tile, mt, nt, tile_valid);
    const Digits d = decompose(mt, P.t_count, P.fd_t);
    bool valid = row_in_box && tile_valid;
    int64_t off = 0;
#pragma unroll
    for (int dim = 0; dim < 4; ++dim) {
      const int32_t g = coord(d, P.e_base, P.e_step, dim) + il[dim];
      int32_t k0 = 0, rem = g, k1, k2 = 0;
      // g < 0 (a row outside the output) gives garbage components here; `valid` below is false for it
      if (P.e_p1[dim] > 0) { k0 = fdiv(g, P.fd_p1[dim]); rem = g - k0 * P.e_p1[dim]; }
      if (P.e_p2[dim] > 0) { k1 = fdiv(rem, P.fd_p2[dim]); k2 = rem - k1 * P.e_p2[dim]; } else { k1 = rem; }
      const pcgan_comp& m0 = P.e_comp[dim][0];
      const pcgan_comp& m1 = P.e_comp[dim][1];
      const pcgan_comp& m2 = P.e_comp[dim][2];
      valid = valid && g >= 0 && k0 >= m0.lo && k0 < m0.hi && k1 >= m1.lo && k1 < m1.hi && k2 >= m2.lo && k2 < m2.hi;
      off += (k0 - m0.lo) * m0.stride + (k1 - m1.lo) * m1.stride + (k2 - m2.lo) * m2.stride;
    }
    if (STATS) {
      const int32_t group = tile_group(P, d);
      if (cur_group >= 0 && (group != cur_group || nt != cur_nt)) flush_stats(cur_group, cur_nt);
      cur_group = group;
    }
    if (nt != cur_nt) {
      if (bias != nullptr) {
        named_bar_sync(bar_bias, 128);   // everybody in the group is done with the previous tile's bias
        for (int32_t i = et; i < block_n; i += 128) {
          const int32_t ch = nt * block_n + i;
          sts_f32(s_bias + i * 4, ch < n_valid ? __ldg(bias + ch) : 0.f);
        }
        named_bar_sync(bar_bias, 128);
      }
      cur_nt = nt;
    }
    const int32_t ncol_limit = min(n_valid - nt * block_n, block_n);
    using OutT = typename std::conditional<BF16, __nv_bfloat16, float>::type;
    OutT* orow = reinterpret_cast<OutT*>(P.out) + off + static_cast<int64_t>(nt) * block_n * cs;
    const bool fast_rows = cs == 1 && ((reinterpret_cast<uintptr_t>(orow) & 15) == 0);

    uint32_t sub, acc_phase;
    acc_slot(P, jt, sub, acc_phase);
    mbar_wait(es.tmem_full + sub, acc_phase);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_base + ((q * 32u) << 16) + es.group * kAccCols + sub * 128u;

    for (int32_t c0 = 0; c0 < ncol_limit; c0 += 32) {
      uint32_t raw[32];
      tmem_ld_32x32(taddr + c0, raw);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
      if (bias != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = lds_f32x4(s_bias + (c0 + 4 * i) * 4);
          v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
        }
      }
      if (STATS) {
        // column sums over the warp's 32 rows: a butterfly of 31 shuffles leaves the total of column c0 + l in lane l
        // (once for the values, once for their squares), which the lane adds to the warp's own partial sums: no
        // shared-memory atomics, no transposes
        const float s1 = warp_column_sums<false>(v, valid, lane);
        const float s2 = warp_column_sums<true>(v, valid, lane);
        const uint32_t a = my_part + (c0 + lane) * 4;
        sts_f32(a, lds_f32(a) + s1);
        sts_f32(a + 1024, lds_f32(a + 1024) + s2);
      }
      if (valid) {
        const int32_t ncols = ncol_limit - c0;
        if (fast_rows && ncols >= 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = act_ct<ACT>(v[i], slope);
          if (BF16) {
            uint4* o4 = reinterpret_cast<uint4*>(orow + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 w;
              w.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
              w.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
              w.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
              w.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
              o4[i] = w;
            }
          } else {
            float4* o4 = reinterpret_cast<float4*>(orow + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        } else {
          OutT* o = orow + static_cast<int64_t>(c0) * cs;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i < ncols) {
              const float y = act_ct<ACT>(v[i], slope);
              if (BF16) o[i * cs] = __float2bfloat16(y);
              else o[i * cs] = y;
            }
          }
        }
      }
    }
    // accumulator drained: hand the TMEM stage back to the MMA warp
    tcgen05_fence_before();
    epi_release(P, es, sub);
    ++jt;
  }
  if (STATS && cur_group >= 0) flush_stats(cur_group, cur_nt);
}

// Shift-sum epilogue (few output channels, horizontal taps in N): out[i][c] = sum_j acc[i + j][j*cpad + c].  The 128
// accumulator rows of the group meet in a [128][17] shared tile, 16 columns at a time.
template <int ACT, bool BF16>
__device__ __forceinline__ void epilogue_shift(const DevParams& P, const EpiShared es, uint32_t tmem_base, const Sched sch) {
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q = warp & 3;
  const uint32_t row = q * 32 + lane;
  const uint32_t bar = 5 + es.group;
  const uint32_t tr = es.s_tr - q * kWarpScratch;     // the group's four warp tiles are contiguous: [128][17]
  const int32_t kw = P.shift_taps, cp = P.shift_cpad, nv = P.n_valid;
  const int64_t cs = P.out_cstride;
  const float slope = P.act_slope;
  float bias[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) bias[c] = (P.bias != nullptr && c < nv) ? __ldg(P.bias + c) : 0.f;
  int32_t il[4];
  {
    uint32_t r = row;
#pragma unroll
    for (int dim = 0; dim < 4; ++dim) {
      const int32_t bx = P.box[dim];
      il[dim] = 0;
      if (bx > 1) { il[dim] = r % bx; r /= bx; }
    }
  }
  const bool row_out = static_cast<int32_t>(row) < P.a_rows - (kw - 1);
  uint32_t jt = 0;   // tiles this group has drained
  for (int32_t tile = sch.begin + static_cast<int32_t>(es.group); tile < sch.end; tile += kEpiGroups) {
    int32_t mt, nt;
    bool tile_valid;
    kmajor_item(P, sch, tile, mt, nt, tile_valid);
    const Digits d = decompose(mt, P.t_count, P.fd_t);
    bool valid = row_out && tile_valid;
    int64_t off = 0;
#pragma unroll
    for (int dim = 0; dim < 4; ++dim) {
      const int32_t g = coord(d, P.e_base, P.e_step, dim) + il[dim];
      int32_t k0 = 0, rem = g, k1, k2 = 0;
      // g < 0 (a row outside the output) gives garbage components here; `valid` below is false for it
      if (P.e_p1[dim] > 0) { k0 = fdiv(g, P.fd_p1[dim]); rem = g - k0 * P.e_p1[dim]; }
      if (P.e_p2[dim] > 0) { k1 = fdiv(rem, P.fd_p2[dim]); k2 = rem - k1 * P.e_p2[dim]; } else { k1 = rem; }
      const pcgan_comp& m0 = P.e_comp[dim][0];
      const pcgan_comp& m1 = P.e_comp[dim][1];
      const pcgan_comp& m2 = P.e_comp[dim][2];
      valid = valid && g >= 0 && k0 >= m0.lo && k0 < m0.hi && k1 >= m1.lo && k1 < m1.hi && k2 >= m2.lo && k2 < m2.hi;
      off += (k0 - m0.lo) * m0.stride + (k1 - m1.lo) * m1.stride + (k2 - m2.lo) * m2.stride;
    }
    uint32_t sub, acc_phase;
    acc_slot(P, jt, sub, acc_phase);
    mbar_wait(es.tmem_full + sub, acc_phase);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_base + ((q * 32u) << 16) + es.group * kAccCols + sub * 128u;
    uint32_t raw[32];
    tmem_ld_32x32(taddr, raw);
    tmem_ld_wait();
    // the accumulator is in registers: hand the TMEM stage back before the exchange
    tcgen05_fence_before();
    epi_release(P, es, sub);
    ++jt;
    float o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      named_bar_sync(bar, 128);       // everybody has read the previous half
#pragma unroll
      for (int i = 0; i < 16; ++i) sts_f32(tr + (row * 17 + i) * 4, __uint_as_float(raw[16 * h + i]));
      named_bar_sync(bar, 128);
      if (row_out) {
        if (cp == 4) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int32_t j = (16 * h + i) >> 2;             // static after unrolling: o[] stays in registers
            if (j < kw) o[i & 3] += lds_f32(tr + ((row + j) * 17 + i) * 4);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int32_t j = (16 * h + i) >> 3;
            if (j < kw) o[i & 7] += lds_f32(tr + ((row + j) * 17 + i) * 4);
          }
        }
      }
    }
    if (valid) {
#pragma unroll
      for (int c = 0; c < 8; ++c) o[c] = act_ct<ACT>(o[c] + bias[c], slope);
      if (BF16) {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(P.out) + off;
        if (cs == 1 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
          // the destination is an 8-channel NHWC pixel: channels >= n_valid are written as zeros
          uint4 w;
          w.x = pack_bf16x2(nv > 0 ? o[0] : 0.f, nv > 1 ? o[1] : 0.f);
          w.y = pack_bf16x2(nv > 2 ? o[2] : 0.f, nv > 3 ? o[3] : 0.f);
          w.z = pack_bf16x2(nv > 4 ? o[4] : 0.f, nv > 5 ? o[5] : 0.f);
          w.w = pack_bf16x2(nv > 6 ? o[6] : 0.f, nv > 7 ? o[7] : 0.f);
          *reinterpret_cast<uint4*>(dst) = w;
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c < nv) dst[c * cs] = __float2bfloat16(o[c]);
        }
      } else {
        float* dst = reinterpret_cast<float*>(P.out) + off;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < nv) dst[c * cs] = o[c];
      }
    }
  }
}

template <int ACT>
__device__ __forceinline__ void epi_dispatch(bool bf16, bool stats, const DevParams& P, const EpiShared& es, uint32_t tmem_base,
                                             const Sched& sch) {
  if (P.shift_taps > 0) {
    if (bf16) epilogue_shift<ACT, true>(P, es, tmem_base, sch);
    else epilogue_shift<ACT, false>(P, es, tmem_base, sch);
    return;
  }
  if (bf16) {
    if (stats) epilogue_kmajor<ACT, true, true>(P, es, tmem_base, sch);
    else epilogue_kmajor<ACT, true, false>(P, es, tmem_base, sch);
  } else {
    if (stats) epilogue_kmajor<ACT, false, true>(P, es, tmem_base, sch);
    else epilogue_kmajor<ACT, false, false>(P, es, tmem_base, sch);
  }
}

// Weight-gradient epilogue: fp32 partial tiles added into the packed gradient with vector reductions.
__device__ __forceinline__ void epilogue_wgrad(const DevParams& P, const EpiShared es, uint32_t tmem_base, const Sched sch,
                                               int32_t total_kb, int32_t kb_per_split) {
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q = warp & 3;
  const uint32_t row = q * 32 + lane;
  const int32_t block_n = P.block_n;
  uint32_t jt = 0;   // tiles this group has drained
  for (int32_t tile = sch.begin + static_cast<int32_t>(es.group); tile < sch.end; tile += kEpiGroups) {
    int32_t ks, nt, mt, tap;
    wgrad_item(P, sch, tile, ks, nt, mt, tap);
    const int32_t grow = mt * 128 + row;
    const bool valid = grow < P.m_valid && ks * kb_per_split < total_kb;
    const int32_t ncol_limit = min(P.wg_ncols - nt * block_n, block_n);
    float* orow = reinterpret_cast<float*>(P.out) + static_cast<int64_t>(grow) * P.ldo + P.tap_bk[tap] + nt * block_n;
    const bool aligned = (reinterpret_cast<uintptr_t>(orow) & 15) == 0;
    uint32_t sub, acc_phase;
    acc_slot(P, jt, sub, acc_phase);
    mbar_wait(es.tmem_full + sub, acc_phase);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_base + ((q * 32u) << 16) + es.group * kAccCols + sub * 128u;
    for (int32_t c0 = 0; c0 < ncol_limit; c0 += 32) {
      uint32_t raw[32];
      tmem_ld_32x32(taddr + c0, raw);
      tmem_ld_wait();
      if (!valid) continue;
      const int32_t ncols = ncol_limit - c0;
      float* o = orow + c0;
      if (ncols >= 32 && aligned) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + i), "r"(raw[i]), "r"(raw[i + 1]),
                       "r"(raw[i + 2]), "r"(raw[i + 3])
                       : "memory");
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < ncols) atomicAdd(o + i, __uint_as_float(raw[i]));
      }
    }
    tcgen05_fence_before();
    epi_release(P, es, sub);
    ++jt;
  }
}

// kPair: the instantiation launched as clusters of two CTAs (desc.pair).  The cta_group::2 instructions live only in it: a
// kernel that contains them cannot be launched without a cluster.
template <bool kPair>
__global__ void __launch_bounds__(kNumThreads, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
             const __grid_constant__ DevParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tail = smem + kDataBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;
  uint64_t* tmem_empty = tmem_full + 4;    // [group][half]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 4);
  uint64_t* bres_bar = tmem_empty + 5;
  const uint32_t s_bias = smem_u32(tail + kTailBias);        // [2 groups][256] f32
  const uint32_t s_tr = smem_u32(tail + kTailTr);            // [8 warps] x kWarpScratch

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const bool wgrad = P.kind == PCGAN_IGEMM_WGRAD;
  const int32_t nstages = P.num_stages;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kPair ? 256 : 128);   // paired: the leader's MMA warp waits for both CTAs' epilogues
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 2) {
    if (kPair) { tmem_alloc_2cta(tmem_slot, kTmemCols); tmem_relinquish_2cta(); }
    else { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  }
  if (threadIdx.x >= 128) {
    for (int i = threadIdx.x - 128; i < 8 * kWarpScratch / 4; i += kNumThreads - 128) sts_f32(s_tr + i * 4, 0.f);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above touched no global data: under programmatic dependent launch it overlapped the previous kernel's tail
  griddep_wait();
  griddep_launch();

  const int32_t k_chunks_fwd = P.num_taps * P.cchunks;
  const int32_t total_kb = P.t_count[0] * P.t_count[1] * P.t_count[2] * P.t_count[3];  // WGRAD: pixel blocks
  const int32_t kb_per_split = wgrad ? (total_kb + P.ksplit - 1) / P.ksplit : 0;
  // static schedule: every CTA (pair) owns one contiguous range of tiles (neighbouring tiles share halos in L2 and, for
  // per-sample statistics, usually the same sample, so the statistics are flushed once per sample, not per tile)
  const Sched sch = make_sched(P);
  const uint16_t pair_mask = 0x3;
  // Single pipeline: warp 0 feeds warp 1, which alternates between the two accumulators.  Dual (P.dual): the single-thread
  // issue loops, not the tensor pipe, bound tiles with little work per K chunk, so the CTA's even tiles run through
  // warps 0 -> 1 -> epilogue group 0 and its odd tiles through warps 3 -> 2 -> epilogue group 1, each pipeline with its
  // own half of the ring, its own barriers and its own accumulator.
  // paired weight gradients: the MMA warp waits for TMA while the ring never fills (one warp issues four boxes per K chunk
  // and steps the pixel-block digits), so the otherwise idle warp 3 issues the X boxes and warp 0 the dY boxes
  const bool split_producer = kPair && wgrad && P.split_prod;
  const bool producer_warp = warp == 0 || (warp == 3 && (P.dual || split_producer));
  const bool mma_warp = warp == 1 || (warp == 2 && P.dual);
  const int32_t pipe = (P.dual && (warp == 2 || warp == 3)) ? 1 : 0;
  const int32_t tile_step = P.dual ? 2 : 1;
  uint8_t* ring = smem + P.ring_off + pipe * nstages * P.stage_bytes;
  full_bar += pipe * (kMaxStages / 2);
  empty_bar += pipe * (kMaxStages / 2);

  if (producer_warp) {
    // ------------------------------------------------------------ TMA producer
    // The whole warp walks the (uniform) schedule; one elected lane issues the copies.
    int32_t stage = 0;
    uint32_t phase = 0;
    // shared addresses of ring stage 0 and of its barriers; the loops step them
    const uint32_t ring0 = smem_u32(ring), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    const uint32_t stage_bytes = static_cast<uint32_t>(P.stage_bytes), a_alloc = static_cast<uint32_t>(P.a_alloc);
    uint32_t sa = ring0, full_a = full0, empty_a = empty0;
    if (P.b_res && pipe == 0 && sch.begin < sch.end) {
      // small weight matrices (one N tile) are fetched once per CTA, not once per tile
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(bres_bar, static_cast<uint32_t>(k_chunks_fwd * P.block_n * 128));
        for (int32_t tap = 0; tap < P.num_taps; ++tap)
          for (int32_t cc = 0; cc < P.cchunks; ++cc)
            tma_load_5d(smem_u32(smem) + (tap * P.cchunks + cc) * P.b_res, &tma_b, smem_u32(bres_bar), P.tap_bk[tap] + cc * P.kc, 0, 0, 0, 0);
      }
      __syncwarp();
    }
    for (int32_t tile = sch.begin + pipe; tile < sch.end; tile += tile_step) {
      if (!wgrad) {
        int32_t mt, nt;
        bool tile_valid;
        kmajor_item(P, sch, tile, mt, nt, tile_valid);
        const Digits d = decompose(mt, P.t_count, P.fd_t);
        int32_t c[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) c[q] = coord(d, P.a_base, P.a_step, q);
        const int32_t half_rows = P.block_n >> 1;   // paired: this CTA holds rows [rank*half, +half) of the B tile
        // paired: both CTAs' copies (own A tile + own half of B each) are counted on the leader's barrier
        const uint32_t bytes = kPair ? 2 * (P.a_bytes + half_rows * 128) : P.a_bytes + (P.b_res ? 0 : P.block_n * 128);
        const bool b_resident = P.b_res != 0;
        const int32_t kps = P.kps;                      // 1 when paired
        const uint32_t sub_bytes = static_cast<uint32_t>(P.sub_bytes);
        int32_t kc = 0, slot = 0;                       // K chunk of the tile, its slot in the current stage
        for (int32_t tap = 0; tap < P.num_taps; ++tap) {
          const int32_t o0 = P.tap_off[tap][0], o1 = P.tap_off[tap][1], o2 = P.tap_off[tap][2],
                        o3 = P.tap_off[tap][3];
          const int32_t ac0 = P.tap_c0[tap], bk0 = P.tap_bk[tap];
          for (int32_t cc = 0; cc < P.cchunks; ++cc) {
            // a stage holds kps consecutive K chunks of the tile (the last stage of a tile may hold fewer): one wait and
            // one expectation for all of them
            if (slot == 0) mbar_wait_addr(empty_a, phase ^ 1);
            if (elect_one_sync()) {
              if (kPair) {
                if (sch.rank == 0) mbar_arrive_expect_tx_addr(full_a, bytes);
                tma_load_5d_2cta(sa, &tma_a, full_a, ac0 + cc * 64, c[0] + o0, c[1] + o1, c[2] + o2, c[3] + o3);      // pairs: bf16 only
                tma_load_5d_2cta(sa + a_alloc, &tma_b, full_a, bk0 + cc * 64, nt * P.block_n + sch.rank * half_rows, 0, 0, 0);
              } else {
                if (slot == 0) mbar_arrive_expect_tx_addr(full_a, bytes * static_cast<uint32_t>(min(kps, k_chunks_fwd - kc)));
                const uint32_t dst = sa + static_cast<uint32_t>(slot) * sub_bytes;
                tma_load_5d(dst, &tma_a, full_a, ac0 + cc * P.kc, c[0] + o0, c[1] + o1, c[2] + o2, c[3] + o3);
                if (!b_resident) tma_load_5d(dst + a_alloc, &tma_b, full_a, bk0 + cc * P.kc, nt * P.block_n, 0, 0, 0);
              }
            }
            __syncwarp();
            ++kc;
            if (++slot == kps || kc == k_chunks_fwd) {
              slot = 0;
              if (++stage == nstages) { stage = 0; phase ^= 1; sa = ring0; full_a = full0; empty_a = empty0; }
              else { sa += stage_bytes; full_a += 8; empty_a += 8; }
            }
          }
        }
      } else {
        int32_t ks, nt, mt, tap;
        wgrad_item(P, sch, tile, ks, nt, mt, tap);
        const int32_t kb0 = ks * kb_per_split;
        const int32_t kb1 = min(kb0 + kb_per_split, total_kb);
        const int32_t kc = P.kc;                    // channels per box: 64 (bf16) or 32 (tf32)
        const int32_t nb = P.block_n / kc;
        const int32_t amax = 128 / kc;              // boxes of one 128-channel M tile
        const int32_t o0 = P.tap_off[tap][0], o1 = P.tap_off[tap][1], o2 = P.tap_off[tap][2], o3 = P.tap_off[tap][3];
        // the nb boxes of this N tile: 64-channel slices of one pixel box, or (filter rows in N) the same channels at
        // box coordinates nt*nb + j along dim wg_box_dim
        const int32_t bdim = P.wg_box_dim;
        const int32_t bc0 = P.tap_c0[tap] + (bdim ? 0 : nt * P.block_n), bcs = bdim ? 0 : kc;
        const int32_t box0 = bdim ? nt * nb : 0;
        const int32_t e0 = bdim == 1, e1 = bdim == 2, e2 = bdim == 3, e3 = bdim == 4;
        // pixel blocks kb0 .. kb1-1 are consecutive: the mixed-radix digits are stepped, not re-divided, per block
        Digits d = decompose(kb0 < total_kb ? kb0 : 0, P.t_count, P.fd_t);
        // kc-channel boxes of this M tile (and of the peer's) that exist
        const int32_t a_boxes = min((P.a_ch - mt * 128 + kc - 1) / kc, amax);
        const int32_t peer_boxes = min((P.a_ch - (mt ^ 1) * 128 + kc - 1) / kc, amax);
        for (int32_t kb = kb0; kb < kb1; ++kb) {
          const int32_t a0 = coord(d, P.a_base, P.a_step, 0), a1 = coord(d, P.a_base, P.a_step, 1),
                        a2 = coord(d, P.a_base, P.a_step, 2), a3 = coord(d, P.a_base, P.a_step, 3);
          const int32_t b0 = coord(d, P.b_base, P.b_step, 0) + o0, b1 = coord(d, P.b_base, P.b_step, 1) + o1,
                        b2 = coord(d, P.b_base, P.b_step, 2) + o2, b3 = coord(d, P.b_base, P.b_step, 3) + o3;
          if (++d.t[0] == P.t_count[0]) {
            d.t[0] = 0;
            if (++d.t[1] == P.t_count[1]) {
              d.t[1] = 0;
              if (++d.t[2] == P.t_count[2]) { d.t[2] = 0; ++d.t[3]; }
            }
          }
          mbar_wait_addr(empty_a, phase ^ 1);
          if (elect_one_sync()) {
            // a box that lies entirely beyond the tensor's channels is not fetched: its rows of the tile are >= m_valid
            // and never stored, whatever the shared memory holds
            if (kPair) {
              // both CTAs' copies are counted on the leader's barrier; this CTA holds its own M tile and half of the X boxes.
              // warp 0: the expectation (all bytes of both CTAs and both warps) and the dY boxes; warp 3: the X boxes, whose
              // bytes may land before the expectation is posted (the barrier's pending arrival keeps the phase open)
              const bool w_a = warp == 0, w_b = split_producer ? warp == 3 : true;
              const int32_t hb = w_b ? (nb >> 1) : 0;
              if (sch.rank == 0 && warp == 0) mbar_arrive_expect_tx_addr(full_a, (a_boxes + peer_boxes + nb) * kBoxBytesMN);
#pragma unroll
              for (int j = 0; j < 2; ++j)
                if (w_a && j < a_boxes) tma_load_5d_2cta(sa + j * kBoxBytesMN, &tma_a, full_a, mt * 128 + j * 64, a0, a1, a2, a3);
              for (int j = 0; j < hb; ++j) {
                const int32_t jj = static_cast<int>(sch.rank) * hb + j, bx = bdim ? box0 + jj : 0;
                tma_load_5d_2cta(sa + a_alloc + j * kBoxBytesMN, &tma_b, full_a, bc0 + jj * bcs, b0 + e0 * bx, b1 + e1 * bx,
                                 b2 + e2 * bx, b3 + e3 * bx);
              }
            } else {
              mbar_arrive_expect_tx_addr(full_a, (a_boxes + nb) * kBoxBytesMN);
              for (int j = 0; j < a_boxes; ++j)
                tma_load_5d(sa + j * kBoxBytesMN, &tma_a, full_a, mt * 128 + j * kc, a0, a1, a2, a3);
              for (int j = 0; j < nb; ++j) {
                const int32_t bx = bdim ? box0 + j : 0;
                tma_load_5d(sa + a_alloc + j * kBoxBytesMN, &tma_b, full_a, bc0 + j * bcs, b0 + e0 * bx, b1 + e1 * bx, b2 + e2 * bx,
                            b3 + e3 * bx);
              }
            }
          }
          __syncwarp();
          if (++stage == nstages) { stage = 0; phase ^= 1; sa = ring0; full_a = full0; empty_a = empty0; }
          else { sa += stage_bytes; full_a += 8; empty_a += 8; }
        }
      }
    }
  } else if (mma_warp) {
    // -------------------------------------------------------------- MMA issuer
    // Converged warp, one elected lane issues tcgen05.mma / commit (always the same lane, so the commits track its MMAs).
    // paired: one tcgen05.mma.cta_group::2 of M = 256 per K step, issued by the leader; the peer's MMA warp has nothing to do
    const uint32_t idesc = P.tf32 ? make_idesc_tf32(128, P.block_n, wgrad ? 1u : 0u, wgrad ? 1u : 0u)
                                  : make_idesc_bf16(kPair ? 256 : 128, P.block_n, wgrad ? 1u : 0u, wgrad ? 1u : 0u);
    // K-major: 8-row groups 1024 B apart; one UMMA_K (16 bf16) = 32 B along the swizzled row.
    // MN-major: 64-element column groups one box (8192 B) apart, 8 K-rows = 1024 B; UMMA_K = 16 rows = 2048 B.
    // TF32, MN-major: the 32-bit-granular layout (128-byte rows swizzled in 32-byte units over groups of 4 K rows:
    // CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B on the TMA side, descriptor layout type 1), K step 8 pixel rows = 1024 B.
    const bool tf32 = P.tf32 != 0;
    const uint32_t lbo = wgrad ? kBoxBytesMN : 16;
    const uint32_t sbo = (wgrad && tf32) ? 512 : 1024;
    const uint32_t kstep = wgrad ? ((tf32 ? 1024 : 2048) >> 4) : (32 >> 4);
    const uint32_t ksteps = (wgrad && tf32) ? 8u : 4u;       // MMAs per ring stage
    const uint64_t ltype = (wgrad && tf32) ? (1ull << 61) ^ (2ull << 61) : 0ull;   // swaps layout type 2 (SWIZZLE_128B) for 1
    int32_t stage = 0;
    uint32_t phase = 0, jt = 0;   // jt: tiles this warp has issued
    // descriptors and barrier addresses of ring stage 0; the loop steps them (the 14-bit address field cannot overflow:
    // shared addresses are below 256 KB)
    const uint32_t ring_a = smem_u32(ring);
    const uint64_t da0 = (P.a_window ? make_smem_desc_noswizzle(ring_a, 16, 128) : make_smem_desc(ring_a, lbo, sbo)) ^ ltype;
    const uint64_t db0 = make_smem_desc(P.b_res ? smem_u32(smem) : ring_a + P.a_alloc, lbo, sbo) ^ ltype;
    const uint32_t dstage = static_cast<uint32_t>(P.stage_bytes) >> 4, dres = static_cast<uint32_t>(P.b_res) >> 4;
    const uint32_t dsub = static_cast<uint32_t>(P.sub_bytes) >> 4;
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    const bool b_res = P.b_res != 0;
    uint64_t da = da0, db_ring = db0;
    uint32_t full_a = full0, empty_a = empty0;
    if (P.b_res && sch.begin + pipe < sch.end) mbar_wait(bres_bar, 0);
    const int32_t mma_end = (kPair && sch.rank != 0) ? sch.begin : sch.end;
    for (int32_t tile = sch.begin + pipe; tile < mma_end; tile += tile_step) {
      int32_t nk;
      if (!wgrad) {
        nk = k_chunks_fwd;
      } else {
        const int32_t ks = tile % P.ksplit;
        const int32_t kb0 = ks * kb_per_split;
        nk = max(min(kb0 + kb_per_split, total_kb) - kb0, 0);
      }
      // single pipeline: tiles alternate between the two groups' accumulators; dual: this pipeline's own group
      const uint32_t group = P.dual ? static_cast<uint32_t>(pipe) : (jt & 1u);
      uint32_t sub, acc_phase;
      acc_slot(P, P.dual ? jt : (jt >> 1), sub, acc_phase);
      const uint32_t acc = group * 2 + sub;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + group * kAccCols + sub * 128u;
      // windowed A (da0 without swizzle): pixel rows of 16 B; row m of K chunk j (8 channels of window pixel j) sits at
      // (m + j) * 16, so the core matrices overlap: 16 B between the two K chunks of one MMA, 128 B between 8-row groups
      uint64_t db_res = db0;   // resident B: K chunk kc of the tile
      const int32_t kps = (wgrad || kPair) ? 1 : P.kps;   // K chunks per stage
      for (int32_t kc = 0; kc < nk; kc += kps) {
        mbar_wait_addr(full_a, phase);
        tcgen05_fence_after();
        const int32_t nchunk = min(kps, nk - kc);
        if (elect_one_sync()) {
          if (kPair) {
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k)
              umma_bf16_2cta(tmem_d, da + k * kstep, db_ring + k * kstep, idesc, (kc | k) != 0 ? 1u : 0u);
            tcgen05_commit_2cta_addr(empty_a, pair_mask);   // frees the stage in both CTAs
          } else {
            for (int32_t j = 0; j < nchunk; ++j) {
              const uint64_t daj = da + static_cast<uint32_t>(j) * dsub;
              const uint64_t dbj = b_res ? db_res + static_cast<uint32_t>(j) * dres : db_ring + static_cast<uint32_t>(j) * dsub;
              if (tf32) {
#pragma unroll 4
                for (uint32_t k = 0; k < ksteps; ++k)
                  umma_tf32(tmem_d, daj + k * kstep, dbj + k * kstep, idesc, (kc | j | k) != 0 ? 1u : 0u);
              } else {
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k)
                  umma_bf16(tmem_d, daj + k * kstep, dbj + k * kstep, idesc, (kc | j | k) != 0 ? 1u : 0u);
              }
            }
            tcgen05_commit_addr(empty_a);
          }
        }
        __syncwarp();
        db_res += static_cast<uint32_t>(nchunk) * dres;
        if (++stage == nstages) {
          stage = 0; phase ^= 1;
          da = da0; db_ring = db0; full_a = full0; empty_a = empty0;
        } else {
          da += dstage; db_ring += dstage; full_a += 8; empty_a += 8;
        }
      }
      if (elect_one_sync()) {
        if (kPair) tcgen05_commit_2cta(&tmem_full[acc], pair_mask);   // both CTAs' epilogues drain their half of the rows
        else tcgen05_commit(&tmem_full[acc]);
      }
      __syncwarp();
      ++jt;
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue
    const uint32_t eg = (warp - 4) >> 2;
    EpiShared es{s_bias + eg * 1024, s_tr + (warp - 4) * kWarpScratch, &tmem_full[eg * 2], &tmem_empty[eg * 2], eg};
    if (wgrad) {
      epilogue_wgrad(P, es, tmem_base, sch, total_kb, kb_per_split);
    } else {
      const bool bf = P.out_dtype == PCGAN_DT_BF16;
      const bool st = P.stats_mode != PCGAN_STATS_NONE;
      // one specialised copy of the tile loop per (activation, output type, statistics): the per-element code is
      // straight-line, nothing about the layer is decided inside the column loops
      switch (P.act) {
        case PCGAN_ACT_RELU: epi_dispatch<PCGAN_ACT_RELU>(bf, st, P, es, tmem_base, sch); break;
        case PCGAN_ACT_LRELU: epi_dispatch<PCGAN_ACT_LRELU>(bf, st, P, es, tmem_base, sch); break;
        case PCGAN_ACT_TANH: epi_dispatch<PCGAN_ACT_TANH>(bf, st, P, es, tmem_base, sch); break;
        case PCGAN_ACT_SIGMOID: epi_dispatch<PCGAN_ACT_SIGMOID>(bf, st, P, es, tmem_base, sch); break;
        default: epi_dispatch<PCGAN_ACT_NONE>(bf, st, P, es, tmem_base, sch); break;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();   // no CTA leaves while the pair's MMAs may still read its operands or arrive on its barriers
  if (warp == 2) {
    tcgen05_fence_after();
    if (kPair) tmem_dealloc_2cta(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------- host ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// mode: 0 no swizzle, 1 SWIZZLE_128B, 2 SWIZZLE_128B_ATOM_32B (MN-major fp32 operands)
static int encode_tmap(CUtensorMap* out, const pcgan_tmap& t, const void* base, int mode = 1, bool f32 = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(PCGAN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5], estr[5];
  for (int i = 0; i < 5; ++i) {
    dims[i] = t.dims[i];
    box[i] = t.box[i];
    estr[i] = 1;
  }
  for (int i = 0; i < 4; ++i) strides[i] = t.strides[i + 1];
  const CUtensorMapSwizzle sw = mode == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : (mode == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
  CUresult r = fn(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(PCGAN_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d): dims=[%llu,%llu,%llu,%llu,%llu] strides=[%llu,%llu,%llu,%llu] "
                "box=[%u,%u,%u,%u,%u] base=%p",
                (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                (unsigned long long)dims[3], (unsigned long long)dims[4], (unsigned long long)strides[0],
                (unsigned long long)strides[1], (unsigned long long)strides[2], (unsigned long long)strides[3],
                box[0], box[1], box[2], box[3], box[4], base);
  return PCGAN_OK;
}

}  // namespace pcgan

struct pcgan_igemm_plan {
  pcgan_igemm_desc desc;
  pcgan::DevParams dev;
  std::mutex mu;
  const void* cached_a = nullptr;
  const void* cached_b = nullptr;
  CUtensorMap map_a, map_b;
  int grid = 0;
};

using namespace pcgan;

// PCGAN_BRES=0 turns the resident weight operand off (measurement only).
static bool resident_b_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("PCGAN_BRES");
    cached = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return cached == 1;
}

// PCGAN_KPS=0 keeps one K chunk per ring stage (measurement only).
static bool kps_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("PCGAN_KPS");
    cached = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return cached == 1;
}

// PCGAN_SPLITP=0 keeps one producer warp for paired weight gradients (measurement only).
static bool splitp_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("PCGAN_SPLITP");
    cached = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return cached == 1;
}

// PCGAN_DUAL=0 turns the second pipeline off (measurement only).
static bool dual_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("PCGAN_DUAL");
    cached = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return cached == 1;
}

static int validate_tmap(const pcgan_tmap& t, const char* name, uint32_t inner = 64, uint64_t esz = 2) {
  (void)esz;
  if (t.box[0] != inner) return fail(PCGAN_ERR_INVALID, "%s.box[0] must be %u", name, inner);
  for (int i = 0; i < 5; ++i) {
    if (t.dims[i] == 0 || t.dims[i] > (1ull << 32)) return fail(PCGAN_ERR_INVALID, "%s.dims[%d] out of range", name, i);
    if (t.box[i] == 0 || t.box[i] > 256) return fail(PCGAN_ERR_INVALID, "%s.box[%d] out of range", name, i);
    if (i > 0 && (t.strides[i] % 16 != 0 || t.strides[i] == 0 || t.strides[i] >= (1ull << 40)))
      return fail(PCGAN_ERR_INVALID, "%s.strides[%d]=%llu must be a non-zero multiple of 16 bytes", name, i,
                  (unsigned long long)t.strides[i]);
  }
  return PCGAN_OK;
}

extern "C" int pcgan_igemm_plan_create(const pcgan_igemm_desc* d, pcgan_igemm_plan** out) {
  if (!d || !out) return fail(PCGAN_ERR_INVALID, "null argument");
  *out = nullptr;
  if (d->kind != PCGAN_IGEMM_KMAJOR && d->kind != PCGAN_IGEMM_WGRAD) return fail(PCGAN_ERR_INVALID, "bad kind");
  const bool wg = d->kind == PCGAN_IGEMM_WGRAD;
  if (d->block_n < 16 || d->block_n > 256 || d->block_n % 16 != 0)
    return fail(PCGAN_ERR_UNSUPPORTED, "block_n=%d must be a multiple of 16 in [16,256]", d->block_n);
  if (wg && d->block_n % 64 != 0) return fail(PCGAN_ERR_UNSUPPORTED, "WGRAD block_n=%d must be a multiple of 64", d->block_n);
  int rc;
  if (d->a_window != 0 && d->a_window != 8) return fail(PCGAN_ERR_INVALID, "a_window must be 0 or 8");
  if (d->tf32 != 0 && d->tf32 != 1) return fail(PCGAN_ERR_INVALID, "tf32 must be 0 or 1");
  if (d->tf32 && (d->pair || d->a_window)) return fail(PCGAN_ERR_UNSUPPORTED, "TF32 plans: no pairing, no windowed A");
  const uint32_t kc = d->tf32 ? 32 : 64;       // elements of a 128-byte K chunk
  const uint64_t esz = d->tf32 ? 4 : 2;
  if ((rc = validate_tmap(d->a, "a", d->a_window ? 8 : kc, esz)) != PCGAN_OK) return rc;
  if ((rc = validate_tmap(d->b, "b", kc, esz)) != PCGAN_OK) return rc;
  if (d->a_window) {
    if (wg || d->a.dims[0] != 8 || d->a.strides[1] != 16 || d->a.box[1] < 8 || d->a.box[2] != 1 || d->a.box[3] != 1 ||
        d->a.box[4] != 1 || d->cchunks != 1 || d->pair)
      return fail(PCGAN_ERR_INVALID, "windowed A: KMAJOR, 8-channel tensor with 16-byte pixels, a single-row box of rows + 7 pixels, cchunks == 1, no pairing");
  }
  // windowed A: the box holds the tile's rows plus the 7 pixels the last row's window reaches into
  const int64_t a_rows = (int64_t)d->a.box[1] * d->a.box[2] * d->a.box[3] * d->a.box[4] - (d->a_window ? 7 : 0);
  const int64_t b_rows = (int64_t)d->b.box[1] * d->b.box[2] * d->b.box[3] * d->b.box[4];
  if (!wg) {
    if (a_rows < 1 || a_rows > 128) return fail(PCGAN_ERR_INVALID, "A box has %lld rows (1..128)", (long long)a_rows);
    if (b_rows != d->block_n) return fail(PCGAN_ERR_INVALID, "B box rows %lld != block_n %d", (long long)b_rows, d->block_n);
    if (d->ksplit != 1) return fail(PCGAN_ERR_INVALID, "KMAJOR needs ksplit == 1");
    if (d->n_valid < 1 || d->n_valid > d->n_tiles * d->block_n) return fail(PCGAN_ERR_INVALID, "n_valid out of range");
    if (d->out_dtype != PCGAN_DT_BF16 && d->out_dtype != PCGAN_DT_F32) return fail(PCGAN_ERR_INVALID, "bad out_dtype");
    if (d->stats_mode != PCGAN_STATS_NONE && d->stats_dim >= 4) return fail(PCGAN_ERR_INVALID, "bad stats_dim");
  } else {
    if (a_rows != 64 || b_rows != 64) return fail(PCGAN_ERR_INVALID, "WGRAD boxes must hold exactly 64 pixels (got %lld, %lld)", (long long)a_rows, (long long)b_rows);
    if (d->ksplit < 1 || d->m_tiles < 1) return fail(PCGAN_ERR_INVALID, "WGRAD needs ksplit >= 1 and m_tiles >= 1");
    if (d->m_valid < 1 || d->m_valid > d->m_tiles * 128) return fail(PCGAN_ERR_INVALID, "m_valid out of range");
    if (d->wg_ncols < 1 || d->wg_ncols > d->n_tiles * d->block_n) return fail(PCGAN_ERR_INVALID, "wg_ncols out of range");
    if (d->ldo < 1) return fail(PCGAN_ERR_INVALID, "ldo must be positive");
  }
  if (d->num_taps < 1 || d->num_taps > PCGAN_MAX_TAPS) return fail(PCGAN_ERR_UNSUPPORTED, "num_taps=%d (1..%d)", d->num_taps, PCGAN_MAX_TAPS);
  if (d->cchunks < 1 || d->n_tiles < 1) return fail(PCGAN_ERR_INVALID, "cchunks / n_tiles must be >= 1");
  int64_t tiles = 1;
  for (int j = 0; j < 4; ++j) {
    if (d->t_count[j] < 1) return fail(PCGAN_ERR_INVALID, "t_count[%d] must be >= 1", j);
    tiles *= d->t_count[j];
  }
  int64_t total = wg ? (int64_t)d->num_taps * d->m_tiles * d->n_tiles * d->ksplit : tiles * d->n_tiles;
  if (total < 1 || total > 0x7fffffff) return fail(PCGAN_ERR_INVALID, "tile count out of range");
  if (d->pair != 0 && d->pair != 1) return fail(PCGAN_ERR_INVALID, "pair must be 0 or 1");
  if (d->shift_taps != 0) {
    if (wg || d->shift_taps < 1 || d->shift_taps > 8 || (d->shift_cpad != 4 && d->shift_cpad != 8) || d->shift_taps * d->shift_cpad > 32 ||
        d->block_n != 32 || d->n_tiles != 1 || d->n_valid > d->shift_cpad || d->stats_mode != PCGAN_STATS_NONE ||
        a_rows <= d->shift_taps - 1)
      return fail(PCGAN_ERR_INVALID, "shift-sum epilogue: KMAJOR, block_n == 32, one N tile, shift_taps*shift_cpad <= 32, n_valid <= shift_cpad in {4, 8}, no statistics");
  }
  if (d->wg_box_dim != 0 && (!wg || d->wg_box_dim < 1 || d->wg_box_dim > 4 || d->num_taps != 1))
    return fail(PCGAN_ERR_INVALID, "wg_box_dim: WGRAD with num_taps == 1, dim 1..4");
  if (d->pair && wg && (d->m_tiles % 2 != 0 || (d->block_n / 64) % 2 != 0))
    return fail(PCGAN_ERR_INVALID, "paired WGRAD needs an even m_tiles and an even number of 64-column B boxes");

  pcgan_igemm_plan* p = new pcgan_igemm_plan();
  p->desc = *d;
  DevParams& v = p->dev;
  memset(&v, 0, sizeof(v));
  v.kind = d->kind; v.block_n = d->block_n; v.a_rows = (int32_t)a_rows; v.a_ch = (int32_t)d->a.dims[0];
  {
    // operand ring: as many stages as fit (small tiles get a deep ring, which is what hides the TMA latency
    // between the many short tiles of the tiny-K layers)
    v.a_window = d->a_window;
    v.a_bytes = d->a_window ? (int32_t)(a_rows + 7) * 16 : (int32_t)a_rows * 128;
    const int32_t a_alloc = wg ? (128 / (int32_t)kc) * kBoxBytesMN : (v.a_bytes + 1023) / 1024 * 1024;
    // paired: each CTA holds half of the B tile (tcgen05.mma.cta_group::2 reads the other half from the peer)
    const int32_t b_rows_cta = d->pair ? d->block_n / 2 : d->block_n;
    const int32_t b_alloc = wg ? (b_rows_cta / (int32_t)kc) * kBoxBytesMN : (b_rows_cta * 128 + 1023) / 1024 * 1024;
    // a weight matrix of one N tile that fits in half of the operand memory is fetched once per CTA and stays
    const int64_t b_total = (int64_t)d->num_taps * d->cchunks * b_alloc;
    const bool resident = resident_b_enabled() && !wg && !d->pair && d->n_tiles == 1 && b_total <= kDataBytes / 2;
    v.a_alloc = a_alloc;
    v.b_res = resident ? b_alloc : 0;
    v.ring_off = resident ? (int32_t)b_total : 0;
    v.sub_bytes = resident ? a_alloc : a_alloc + b_alloc;
    // K chunks per stage (unpaired forward / data gradient): as many as leave each pipeline three stages, at most 8; the
    // single-thread issue loops pay their barrier round trip once per stage
    int32_t kps = 1;
    if (!wg && !d->pair && kps_enabled()) {
      const int64_t per_pipe = (kDataBytes - v.ring_off) / 2;
      const int64_t k_total = (int64_t)d->num_taps * d->cchunks;
      while (kps < 8 && kps < k_total && (int64_t)(kps + 1) * v.sub_bytes * 3 <= per_pipe) ++kps;
    }
    v.kps = kps;
    v.split_prod = (wg && d->pair && splitp_enabled()) ? 1 : 0;
    v.tf32 = d->tf32;
    v.kc = (int32_t)kc;
    v.stage_bytes = kps * v.sub_bytes;
    int32_t ns = (kDataBytes - v.ring_off) / v.stage_bytes;
    // two pipelines when each still gets at least two stages and the plan is not paired (a pair already shares one MMA
    // stream between two SMs)
    v.dual = dual_enabled() && !d->pair && ns >= 4 ? 1 : 0;
    if (v.dual) ns /= 2;
    v.acc_sub = v.dual && d->block_n <= 128 ? 1 : 0;
    const int32_t cap = v.dual ? kMaxStages / 2 : kMaxStages;
    v.num_stages = ns > cap ? cap : (ns < 2 ? 2 : ns);
  }
  memcpy(v.t_count, d->t_count, sizeof(v.t_count));
  memcpy(v.a_base, d->a_base, sizeof(v.a_base)); memcpy(v.a_step, d->a_step, sizeof(v.a_step));
  memcpy(v.b_base, d->b_base, sizeof(v.b_base)); memcpy(v.b_step, d->b_step, sizeof(v.b_step));
  v.n_tiles = d->n_tiles; v.m_tiles = wg ? d->m_tiles : 1; v.ksplit = wg ? d->ksplit : 1;
  v.num_m_tiles = (int32_t)tiles; v.total_tiles = (int32_t)total;
  v.pair = d->pair;
  v.shift_taps = d->shift_taps; v.shift_cpad = d->shift_cpad;
  v.wg_box_dim = d->wg_box_dim;
  v.sched_items = !d->pair ? (int32_t)total : (wg ? (int32_t)(total / 2) : (int32_t)(((tiles + 1) / 2) * d->n_tiles));
  v.num_taps = d->num_taps; v.cchunks = d->cchunks;
  memcpy(v.tap_off, d->tap_off, sizeof(v.tap_off)); memcpy(v.tap_c0, d->tap_c0, sizeof(v.tap_c0));
  memcpy(v.tap_bk, d->tap_bk, sizeof(v.tap_bk));
  for (int i = 0; i < 4; ++i) v.box[i] = d->a.box[i + 1];
  memcpy(v.e_base, d->e_base, sizeof(v.e_base)); memcpy(v.e_step, d->e_step, sizeof(v.e_step));
  memcpy(v.e_p1, d->e_p1, sizeof(v.e_p1)); memcpy(v.e_p2, d->e_p2, sizeof(v.e_p2));
  for (int i = 0; i < 4; ++i) {
    v.fd_t[i] = make_fastdiv(v.t_count[i]);
    v.fd_p1[i] = make_fastdiv(v.e_p1[i]);
    v.fd_p2[i] = make_fastdiv(v.e_p2[i]);
  }
  v.fd_nt = make_fastdiv(v.n_tiles);
  v.fd_sdiv = make_fastdiv(d->stats_div);
  memcpy(v.e_comp, d->e_comp, sizeof(v.e_comp));
  v.out_dtype = d->out_dtype; v.act = d->act; v.act_slope = d->act_slope; v.n_valid = d->n_valid;
  v.out_cstride = d->out_cstride; v.stats_mode = d->stats_mode; v.stats_dim = d->stats_dim; v.stats_comp = d->stats_comp; v.stats_div = d->stats_div;
  v.m_valid = d->m_valid; v.wg_ncols = d->wg_ncols; v.ldo = d->ldo;
  *out = p;
  return PCGAN_OK;
}

extern "C" void pcgan_igemm_plan_destroy(pcgan_igemm_plan* p) { delete p; }

extern "C" int64_t pcgan_selftest_fastdiv(uint32_t seed, int32_t iters) {
  // divisors: small, plan-like (image sizes, tile counts), powers of two and their neighbours, up to 2^31 - 1;
  // dividends: 0, multiples of the divisor and their neighbours, random values below 2^31
  uint64_t st = seed ? seed : 1u;
  auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return static_cast<uint32_t>(st >> 33); };
  int64_t bad = 0;
  for (int32_t it = 0; it < iters; ++it) {
    int32_t d;
    switch (it & 3) {
      case 0: d = 1 + static_cast<int32_t>(rnd() % 70000u); break;
      case 1: d = static_cast<int32_t>(1u << (rnd() % 31u)) + static_cast<int32_t>(rnd() % 3u) - 1; break;
      case 2: d = 1 + static_cast<int32_t>(rnd() % 0x7fffffffu); break;
      default: { static const int32_t k[] = {3, 7, 49, 98, 112, 134, 224, 230, 12544, 17956, 16384, 65536}; d = k[rnd() % 12u]; }
    }
    if (d < 1) d = 1;
    const FastDiv f = make_fastdiv(d);
    const uint32_t m = rnd() % 100000u;
    const int64_t xs[6] = {0, static_cast<int64_t>(rnd()) & 0x7fffffff, static_cast<int64_t>(d) * m, static_cast<int64_t>(d) * m - 1,
                           static_cast<int64_t>(d) * m + 1, 0x7fffffff};
    for (int64_t x : xs) {
      if (x < 0 || x > 0x7fffffff) continue;
      if (fdiv_host(static_cast<int32_t>(x), f) != static_cast<int32_t>(x / d)) ++bad;
    }
  }
  return bad;
}

extern "C" int pcgan_igemm_run(pcgan_igemm_plan* p, const void* a, const void* b, void* out, const float* bias,
                               float* stats, pcgan_stream_t stream) {
  if (!p || !a || !b || !out) return fail(PCGAN_ERR_INVALID, "null argument");
  if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15))
    return fail(PCGAN_ERR_INVALID, "operand pointers must be 16-byte aligned");
  if (p->desc.kind == PCGAN_IGEMM_KMAJOR && p->desc.stats_mode != PCGAN_STATS_NONE && !stats)
    return fail(PCGAN_ERR_INVALID, "stats_mode set but stats pointer is null");
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, []() {
    attr_err = cudaFuncSetAttribute(igemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(igemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  });
  if (attr_err != cudaSuccess) return fail(PCGAN_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  CUtensorMap ma, mb;
  DevParams dev;
  {
    std::lock_guard<std::mutex> lock(p->mu);
    if (p->cached_a != a) {
      const bool f32 = p->desc.tf32 != 0;
      const int mode = p->desc.a_window ? 0 : ((f32 && p->desc.kind == PCGAN_IGEMM_WGRAD) ? 2 : 1);
      int rc = encode_tmap(&p->map_a, p->desc.a, a, mode, f32);
      if (rc != PCGAN_OK) return rc;
      p->cached_a = a;
    }
    if (p->cached_b != b) {
      pcgan_tmap tb = p->desc.b;
      if (p->desc.pair && p->desc.kind == PCGAN_IGEMM_KMAJOR) tb.box[1] = p->desc.block_n / 2;   // each CTA of a pair fetches half of B
      const bool f32 = p->desc.tf32 != 0;
      int rc = encode_tmap(&p->map_b, tb, b, (f32 && p->desc.kind == PCGAN_IGEMM_WGRAD) ? 2 : 1, f32);
      if (rc != PCGAN_OK) return rc;
      p->cached_b = b;
    }
    if (p->grid == 0) {
      int sms = sm_count();
      if (sms <= 0) return fail(PCGAN_ERR_CUDA, "cannot query SM count");
      if (!p->desc.pair) {
        p->grid = p->dev.total_tiles < sms ? p->dev.total_tiles : sms;
      } else {
        const int pairs = p->dev.sched_items < sms / 2 ? p->dev.sched_items : sms / 2;
        p->grid = 2 * pairs;
      }
    }
    ma = p->map_a; mb = p->map_b; dev = p->dev;
  }
  dev.out = out; dev.bias = bias; dev.stats = stats;
  {
    cudaError_t e = p->desc.pair
                        ? launch_pdl(igemm_kernel<true>, dim3(p->grid), dim3(kNumThreads), kSmemBytes, static_cast<cudaStream_t>(stream), 2, ma, mb, dev)
                        : launch_pdl(igemm_kernel<false>, dim3(p->grid), dim3(kNumThreads), kSmemBytes, static_cast<cudaStream_t>(stream), 1, ma, mb, dev);
    if (e != cudaSuccess) return fail(PCGAN_ERR_CUDA, "cudaLaunchKernelEx(igemm_kernel): %s", cudaGetErrorString(e));
  }
  PCGAN_LAUNCH_OK("igemm_kernel");
  return PCGAN_OK;
}
