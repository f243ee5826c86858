// Normalisation kernels of the wsgan_emb step (HBM-bound): statistics finalize, normalise + activation
// (+ residual) with halo write, halo fold, and the two-pass backward.
//
// Layout: activations are NHWC bf16 in physically padded buffers.  Every kernel runs on a 2-D grid
// (chunks of one sample, sample): a block owns a contiguous run of 16-byte vectors (8 channels) of ONE sample,
// so all index arithmetic is 32-bit (one division per vector), the per-channel constants of a thread are loaded
// once (256 % (C/8) == 0: a thread always sees the same 8 channels), and each thread keeps kU independent
// 16-byte loads in flight.
#include <mutex>

#include "common.cuh"

namespace pcgan {

static constexpr int kT = 256;   // threads per block

// Statistics group of sample n: one per sample (InstanceNorm), one for the batch (BatchNorm), or `groups` equal runs of
// consecutive samples (several independent BatchNorm batches in one launch: the passes of a network batched together)
__device__ __forceinline__ int group_of(int n, int groups, int n_total) { return groups > 1 ? n / (n_total / groups) : 0; }

__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void unpack8(const uint4 u, float (&v)[8]) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void load_f8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// -------------------------------------------------------------- norm finalize
// block = 32 channels x 8 group lanes; one launch covers every (group, channel)
__global__ void __launch_bounds__(kT) norm_finalize_kernel(pcgan_norm_finalize_args a) {
  __shared__ float sm[8][33], sv[8][33];
  griddep_wait();
  griddep_launch();
  const int cl = threadIdx.x & 31, gl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const bool combine = a.in_groups > a.groups;   // per-sample statistics folded into one batch statistic (groups == 1)
  float macc = 0.f, vacc = 0.f;
  if (c < a.c && !combine) {
    const float gam = a.gamma ? a.gamma[c] : 1.f;
    const float bet = a.beta ? a.beta[c] : 0.f;
    for (int g = gl; g < a.groups; g += 8) {
      const int64_t o = static_cast<int64_t>(g) * a.c + c;
      const float2 s = reinterpret_cast<const float2*>(a.stats)[o];
      const float mean = s.x / a.count;
      float var = s.y / a.count - mean * mean;
      var = var > 0.f ? var : 0.f;
      const float rstd = rsqrtf(var + a.eps);
      if (a.mean) a.mean[o] = mean;
      if (a.rstd) a.rstd[o] = rstd;
      if (a.scale) a.scale[o] = gam * rstd;
      if (a.shift) a.shift[o] = bet - mean * gam * rstd;
      macc += mean;
      vacc += var;
    }
  }
  if (c < a.c && combine) {
    for (int g = gl; g < a.in_groups; g += 8) {
      const int64_t o = static_cast<int64_t>(g) * a.c + c;
      const float2 s = reinterpret_cast<const float2*>(a.stats)[o];
      const float m = a.drop_mask ? a.drop_mask[o] : 1.f;
      macc += m * s.x;        // here: partial sums, not means
      vacc += m * m * s.y;
    }
  }
  sm[gl][cl] = macc;
  sv[gl][cl] = vacc;
  __syncthreads();
  if (gl != 0 || c >= a.c) return;
  float m = 0.f, v = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { m += sm[i][cl]; v += sv[i][cl]; }
  if (combine) {
    const float gam = a.gamma ? a.gamma[c] : 1.f;
    const float bet = a.beta ? a.beta[c] : 0.f;
    const float mean = m / a.count;
    float var = v / a.count - mean * mean;
    var = var > 0.f ? var : 0.f;
    const float rstd = rsqrtf(var + a.eps);
    if (a.mean) a.mean[c] = mean;
    if (a.rstd) a.rstd[c] = rstd;
    if (a.scale) a.scale[c] = gam * rstd;
    if (a.shift) a.shift[c] = bet - mean * gam * rstd;
    m = mean;
    v = var;
  } else {
    m /= a.groups;
    v /= a.groups;
  }
  if (a.running_mean) a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * m;
  if (a.running_var) {
    const float unbias = a.count > 1.f ? a.count / (a.count - 1.f) : 1.f;
    a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (v * unbias);
  }
}

// blockIdx.y = layer, one thread per channel; groups are walked in order (coalesced across channels)
__global__ void __launch_bounds__(kT) norm_running_batched_kernel(const pcgan_running_item* __restrict__ items) {
  griddep_wait();
  griddep_launch();
  const pcgan_running_item it = items[blockIdx.y];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && it.num_batches_tracked) *it.num_batches_tracked += it.sequential ? it.groups : 1;
  if (c >= it.c) return;
  const float unbias_ = it.count > 1.f ? it.count / (it.count - 1.f) : 1.f;
  if (it.sequential) {
    // the groups are successive BatchNorm batches (passes of the network batched into one launch): one EMA step each, in order
    float rm = it.running_mean ? it.running_mean[c] : 0.f, rv = it.running_var ? it.running_var[c] : 0.f;
    for (int g = 0; g < it.groups; ++g) {
      const float2 s = reinterpret_cast<const float2*>(it.stats)[static_cast<int64_t>(g) * it.c + c];
      const float mean = s.x / it.count;
      float var = s.y / it.count - mean * mean;
      var = var > 0.f ? var : 0.f;
      rm = (1.f - it.momentum) * rm + it.momentum * mean;
      rv = (1.f - it.momentum) * rv + it.momentum * (var * unbias_);
    }
    if (it.running_mean) it.running_mean[c] = rm;
    if (it.running_var) it.running_var[c] = rv;
    return;
  }
  float macc = 0.f, vacc = 0.f;
  for (int g = 0; g < it.groups; ++g) {
    const float2 s = reinterpret_cast<const float2*>(it.stats)[static_cast<int64_t>(g) * it.c + c];
    const float mean = s.x / it.count;
    float var = s.y / it.count - mean * mean;
    var = var > 0.f ? var : 0.f;
    macc += mean;
    vacc += var;
  }
  const float m = macc / it.groups, v = vacc / it.groups;
  if (it.running_mean) it.running_mean[c] = (1.f - it.momentum) * it.running_mean[c] + it.momentum * m;
  if (it.running_var) {
    const float unbias = it.count > 1.f ? it.count / (it.count - 1.f) : 1.f;
    it.running_var[c] = (1.f - it.momentum) * it.running_var[c] + it.momentum * (v * unbias);
  }
}

// ------------------------------------------------------------- stream pipeline
// The three hot kernels (norm_apply, norm_bwd_reduce, norm_bwd_apply) stream their inputs through shared memory
// with 1-D bulk async copies (cp.async.bulk + mbarrier complete_tx): a producer warp keeps kStages segments of every
// input stream in flight per block, independent of registers, which is what an HBM-bound kernel on B200 needs
// (~45 KB in flight per SM at full bandwidth).  A segment = seg_px consecutive pixels of one image row
// (<= 8 KB per stream); 8 consumer warps process it (thread = 16-byte vector, fixed 8 channels) and hand the stage
// back through an "empty" mbarrier.
static constexpr int kConsumers = 256;                    // consumer threads of the backward kernels
static constexpr int kStreamThreads = kConsumers + 32;    // + producer warp
// norm_apply keeps half as many consumer threads: its per-thread state is small, so four vectors per thread and segment
// (instead of two) halve the per-segment bookkeeping and the per-block prologue, which is what bounds it at the
// ResnetBlock shape (33 MB, L2 resident: 16.1 -> 13.2 us; profiles/README.md, round 2)
static constexpr int kApplyConsumers = 128;
static constexpr int kReduceConsumers = 128;              // norm_bwd_reduce: no stores, small state: as norm_apply
static constexpr int kApplyThreads = kApplyConsumers + 32;
static constexpr int kStages = 4;
static constexpr int kSegBytes = 8192;

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct SegGeom {
  int32_t seg_px, segs_per_row, seg_vec;   // pixels and 16-byte vectors per segment
};

// One input stream: element offset of interior pixel (y, x) of sample n is base + ((y + pad) * wp + x + pad) * c
struct StreamSrc {
  const __nv_bfloat16* base;   // sample base (already offset by n)
  int32_t pad, wp;
};

template <int NT>
struct Pipe {
  uint64_t* full;    // [kStages]
  uint64_t* empty;   // [kStages]
  uint8_t* data;     // [kStages][NT][kSegBytes]
  __device__ __forceinline__ const uint4* stage(int s, int t) const {
    return reinterpret_cast<const uint4*>(data + (static_cast<size_t>(s) * NT + t) * kSegBytes);
  }
};

template <int NT>
__device__ __forceinline__ Pipe<NT> pipe_init(uint8_t* smem, int consumers = kConsumers) {
  Pipe<NT> p;
  p.data = smem;
  p.full = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(kStages) * NT * kSegBytes);
  p.empty = p.full + kStages;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&p.full[i], 1);
      mbar_init(&p.empty[i], consumers / 32);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();
  griddep_wait();      // the barriers above were set up while the previous kernel drained (programmatic dependent launch)
  griddep_launch();
  return p;
}

// producer: lane 0 of the last warp; segments [g0, g1) of one sample
template <int NT>
__device__ __forceinline__ void pipe_produce(const Pipe<NT>& p, const SegGeom& sg, const StreamSrc (&src)[NT], int32_t c, int32_t g0,
                                             int32_t g1) {
  const uint32_t bytes = static_cast<uint32_t>(sg.seg_vec) * 16u;
  int s = 0;
  uint32_t ph = 0;
  for (int32_t g = g0; g < g1; ++g) {
    const int32_t y = g / sg.segs_per_row;
    const int32_t x0 = (g - y * sg.segs_per_row) * sg.seg_px;
    mbar_wait(&p.empty[s], ph ^ 1);
    mbar_arrive_expect_tx(&p.full[s], bytes * NT);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const __nv_bfloat16* gsrc = src[t].base + (static_cast<int64_t>(y + src[t].pad) * src[t].wp + x0 + src[t].pad) * c;
      bulk_load(const_cast<uint4*>(p.stage(s, t)), gsrc, bytes, &p.full[s]);
    }
    if (++s == kStages) { s = 0; ph ^= 1; }
  }
}

// consumer side of one segment: wait, run body(stage index), release
#define PIPE_CONSUME_BEGIN(p, s, ph) mbar_wait(&(p).full[s], ph)
#define PIPE_CONSUME_END(p, s, ph)                           \
  do {                                                       \
    __syncwarp();                                            \
    if ((threadIdx.x & 31) == 0) mbar_arrive(&(p).empty[s]); \
    if (++s == kStages) { s = 0; ph ^= 1; }                  \
  } while (0)

// ----------------------------------------------------------------- norm apply
// y = act(sc*x + sh [+ rsc*res + rsh]) into the interior of y; under reflect halo every interior pixel within `pad`
// of a border is also stored at its mirror positions (a zero halo is never written: the buffer is allocated zeroed
// and only interiors are ever stored).
// x within `p` of a border but not on it: the pixels ReflectionPad2d copies into the halo (x in [1, p] or [w-1-p, w-2])
__device__ __forceinline__ bool mirrored(int x, int w, int p) {
  return static_cast<unsigned>(x - 1) < static_cast<unsigned>(p) || static_cast<unsigned>(x - (w - 1 - p)) < static_cast<unsigned>(p);
}
// ACT >= 0: the activation is fixed at compile time (the hot instantiations); ACT < 0: taken from the arguments
template <int ACT>
__device__ __forceinline__ float act_t(float v, int act, float slope) {
  if constexpr (ACT == PCGAN_ACT_NONE) return v;
  else if constexpr (ACT == PCGAN_ACT_RELU) return fmaxf(v, 0.f);
  else if constexpr (ACT == PCGAN_ACT_LRELU) return v > 0.f ? v : v * slope;
  else return apply_act(v, act, slope);
}

template <bool RES, int ACT>
__global__ void __launch_bounds__(kApplyThreads) norm_apply_kernel(pcgan_norm_apply_args a, SegGeom sg, int segs_per_block, int lcv) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int NT = RES ? 2 : 1;
  Pipe<NT> pipe = pipe_init<NT>(smem_raw, kApplyConsumers);
  const int n = blockIdx.y;
  const int32_t total_segs = a.h * sg.segs_per_row;
  const int32_t g0 = blockIdx.x * segs_per_block, g1 = min(g0 + segs_per_block, total_segs);
  const int cv = a.c >> 3;
  StreamSrc src[NT];
  src[0].pad = a.x_pad; src[0].wp = a.w + 2 * a.x_pad;
  src[0].base = reinterpret_cast<const __nv_bfloat16*>(a.x) + static_cast<int64_t>(n) * (a.h + 2 * a.x_pad) * src[0].wp * a.c;
  if constexpr (RES) {
    src[1].pad = a.res_pad; src[1].wp = a.w + 2 * a.res_pad;
    src[1].base = reinterpret_cast<const __nv_bfloat16*>(a.res) + static_cast<int64_t>(n) * (a.h + 2 * a.res_pad) * src[1].wp * a.c;
  }
  if (threadIdx.x >= kApplyConsumers) {
    if (threadIdx.x == kApplyConsumers) pipe_produce<NT>(pipe, sg, src, a.c, g0, g1);
    return;
  }
  const int c0 = (threadIdx.x & (cv - 1)) << 3;
  float sc[8], sh[8], rsc[RES ? 8 : 1], rsh[RES ? 8 : 1];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = 1.f; sh[j] = 0.f; }
#pragma unroll
  for (int j = 0; j < (RES ? 8 : 1); ++j) { rsc[j] = 1.f; rsh[j] = 0.f; }
  if (a.stats) {
    // fused finalize: the same arithmetic as norm_finalize_kernel, for this thread's 8 channels of its group
    const int64_t so = static_cast<int64_t>(group_of(n, a.groups, a.n)) * a.c + c0;
    float st[16];
    load_f8(a.stats + so * 2, *reinterpret_cast<float(*)[8]>(st));
    load_f8(a.stats + so * 2 + 8, *reinterpret_cast<float(*)[8]>(st + 8));
    float gam[8], bet[8], mean[8], rstd[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { gam[j] = 1.f; bet[j] = 0.f; }
    if (a.gamma) load_f8(a.gamma + c0, gam);
    if (a.beta) load_f8(a.beta + c0, bet);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mean[j] = st[2 * j] / a.count;
      float var = st[2 * j + 1] / a.count - mean[j] * mean[j];
      var = var > 0.f ? var : 0.f;
      rstd[j] = rsqrtf(var + a.eps);
      sc[j] = gam[j] * rstd[j];
      sh[j] = bet[j] - mean[j] * gam[j] * rstd[j];
    }
    if (blockIdx.x == 0 && threadIdx.x < cv && (a.groups > 1 || n == 0)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (a.mean_out) a.mean_out[so + j] = mean[j];
        if (a.rstd_out) a.rstd_out[so + j] = rstd[j];
        if (a.scale_out) a.scale_out[so + j] = sc[j];
        if (a.shift_out) a.shift_out[so + j] = sh[j];
      }
    }
  } else if (a.scale) {
    const int64_t so = static_cast<int64_t>(group_of(n, a.groups, a.n)) * a.c + c0;
    load_f8(a.scale + so, sc);
    load_f8(a.shift + so, sh);
  }
  if (a.drop_mask) {
    float m[8];
    load_f8(a.drop_mask + static_cast<int64_t>(n) * a.c + c0, m);
#pragma unroll
    for (int j = 0; j < 8; ++j) sc[j] *= m[j];
  }
  if constexpr (RES) {
    if (a.res_scale) {
      const int64_t ro = static_cast<int64_t>(group_of(n, a.res_groups, a.n)) * a.c + c0;
      load_f8(a.res_scale + ro, rsc);
      load_f8(a.res_shift + ro, rsh);
    }
  }
  float pm[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) pm[j] = 1.f;
  const bool has_pm = a.post_mask != nullptr;
  if (has_pm) load_f8(a.post_mask + static_cast<int64_t>(n) * a.c + c0, pm);
  // Everything the per-vector loop needs, in registers: the loop itself is what bounds this kernel (issue slots, not HBM:
  // profiles/README.md), so no divisions, no 64-bit index arithmetic and no halo tests for interior pixels in it.
  const int p = a.y_pad, wp = a.w + 2 * p, w = a.w, h = a.h;
  const int ych = a.y_c > 0 ? a.y_c : a.c;       // channels per pixel of the output buffer (a channel slice of a wider one)
  __nv_bfloat16* ys = reinterpret_cast<__nv_bfloat16*>(a.y) + static_cast<int64_t>(n) * (h + 2 * p) * wp * ych + a.y_c0 + c0;
  const bool reflect = a.y_halo == PCGAN_HALO_REFLECT && p > 0;
  const int act = a.act;
  const float slope = a.act_slope;
  const int px0 = threadIdx.x >> lcv, dpx = kApplyConsumers >> lcv;    // this thread's pixels within a segment: px0, px0 + dpx, ...
  const int seg_vec = sg.seg_vec, seg_px = sg.seg_px;
  int32_t y = g0 / sg.segs_per_row;
  int32_t x0 = (g0 - y * sg.segs_per_row) * seg_px;
  int s = 0;
  uint32_t ph = 0;
  for (int32_t g = g0; g < g1; ++g) {
    __nv_bfloat16* yrow = ys + ((y + p) * wp + x0 + p) * ych;      // element offsets fit 31 bits (checked on the host)
    const bool rowm = reflect && mirrored(y, h, p);
    PIPE_CONSUME_BEGIN(pipe, s, ph);
    const uint4* sx = pipe.stage(s, 0);
    const uint4* sr = RES ? pipe.stage(s, 1) : nullptr;
    int px = px0;
#pragma unroll 2
    for (int v = threadIdx.x; v < seg_vec; v += kApplyConsumers, px += dpx) {
      float x8[8], o[8];
      unpack8(sx[v], x8);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(sc[j], x8[j], sh[j]);
      if constexpr (RES) {
        float r8[8];
        unpack8(sr[v], r8);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += fmaf(rsc[j], r8[j], rsh[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = act_t<ACT>(o[j], act, slope);
      if (has_pm) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] *= pm[j];
      }
      uint4 wv;
      wv.x = pack_bf16x2(o[0], o[1]); wv.y = pack_bf16x2(o[2], o[3]); wv.z = pack_bf16x2(o[4], o[5]); wv.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(yrow + px * ych) = wv;
      if (reflect) {
        const int x = x0 + px;
        if (rowm || mirrored(x, w, p)) {
          // a pixel within `pad` of a border is also stored at its mirror positions in the halo
          const int ya = y + p, xa = x + p;
          const int yb = (y >= 1 && y <= p) ? p - y : -1;
          const int yc = (y <= h - 2 && y >= h - 1 - p) ? p + 2 * (h - 1) - y : -1;
          const int xb = (x >= 1 && x <= p) ? p - x : -1;
          const int xc = (x <= w - 2 && x >= w - 1 - p) ? p + 2 * (w - 1) - x : -1;
#pragma unroll
          for (int iy = 0; iy < 3; ++iy) {
            const int yy = iy == 0 ? ya : (iy == 1 ? yb : yc);
            if (yy < 0) continue;
#pragma unroll
            for (int ix = 0; ix < 3; ++ix) {
              const int xx = ix == 0 ? xa : (ix == 1 ? xb : xc);
              if (xx < 0 || (iy == 0 && ix == 0)) continue;
              *reinterpret_cast<uint4*>(ys + (yy * wp + xx) * ych) = wv;
            }
          }
        }
      }
    }
    PIPE_CONSUME_END(pipe, s, ph);
    x0 += seg_px;
    if (x0 >= w) { x0 = 0; ++y; }
  }
}

// ------------------------------------------------------------------ halo fold
// 16-byte vector of the gradient at interior pixel (y, x) of a padded-grid gradient, halo folded onto its mirror
// (a pixel within p of a border also receives the halo rows / columns that ReflectionPad2d copied from it).
// `have_center`: acc already holds the pixel's own value (streamed), only the mirrors are added.
__device__ __forceinline__ void folded_load(const __nv_bfloat16* g, int y, int x, int h, int w, int p, int c, bool reflect,
                                            float (&acc)[8], bool have_center = false) {
  const int wp = w + 2 * p;
  const int ya = y + p, xa = x + p;
  if (!have_center) unpack8(ldg16(g + (ya * wp + xa) * c), acc);
  if (!reflect || !(mirrored(y, h, p) || mirrored(x, w, p))) return;   // interior pixel: nothing mirrors onto it (the common case)
  const int yb = (y >= 1 && y <= p) ? p - y : -1;
  const int yc = (y <= h - 2 && y >= h - 1 - p) ? p + 2 * (h - 1) - y : -1;
  const int xb = (x >= 1 && x <= p) ? p - x : -1;
  const int xc = (x <= w - 2 && x >= w - 1 - p) ? p + 2 * (w - 1) - x : -1;
  if ((yb & yc & xb & xc) == -1) return;   // interior pixel: nothing mirrors onto it
#pragma unroll
  for (int iy = 0; iy < 3; ++iy) {
    const int yy = iy == 0 ? ya : (iy == 1 ? yb : yc);
    if (yy < 0) continue;
#pragma unroll
    for (int ix = 0; ix < 3; ++ix) {
      const int xx = ix == 0 ? xa : (ix == 1 ? xb : xc);
      if (xx < 0 || (iy == 0 && ix == 0)) continue;
      float v[8];
      unpack8(ldg16(g + (yy * wp + xx) * c), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
}

__global__ void __launch_bounds__(kT) halo_fold_kernel(pcgan_fold_args a, int vec_per_block, int lcv) {
  griddep_wait();
  griddep_launch();
  const int n = blockIdx.y;
  const int cv = a.c >> 3;
  const int total = a.h * a.w * cv;
  const int vb = blockIdx.x * vec_per_block, ve = min(vb + vec_per_block, total);
  const int c0 = (threadIdx.x & (cv - 1)) << 3;
  const int p = a.g_pad;
  const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(a.gpad) +
                           static_cast<int64_t>(n) * (a.h + 2 * p) * (a.w + 2 * p) * a.c + c0;
  const int wap = a.w + 2 * a.add_pad, wop = a.w + 2 * a.out_pad;
  const __nv_bfloat16* add = a.add ? reinterpret_cast<const __nv_bfloat16*>(a.add) +
                                         static_cast<int64_t>(n) * (a.h + 2 * a.add_pad) * wap * a.c + c0
                                   : nullptr;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out) + static_cast<int64_t>(n) * (a.h + 2 * a.out_pad) * wop * a.c + c0;
  const bool reflect = a.halo == PCGAN_HALO_REFLECT;
  for (int v0 = vb + threadIdx.x; v0 < ve; v0 += kT * 2) {
    float acc[2][8];
    uint4 av[2];
    int yy[2], xx[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int v = v0 + u * kT;
      if (v < ve) {
        const int pix = v >> lcv;
        yy[u] = pix / a.w;
        xx[u] = pix - yy[u] * a.w;
        folded_load(g, yy[u], xx[u], a.h, a.w, p, a.c, reflect, acc[u]);
        if (add) av[u] = ldg16(add + ((yy[u] + a.add_pad) * wap + xx[u] + a.add_pad) * a.c);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int v = v0 + u * kT;
      if (v < ve) {
        if (add) {
          float t[8];
          unpack8(av[u], t);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[u][j] += t[j];
        }
        store8(out + ((yy[u] + a.out_pad) * wop + xx[u] + a.out_pad) * a.c, acc[u]);
      }
    }
  }
}

// In-place form of the reflect fold: only the interior pixels within `pad` of a border receive anything (3 % of a 128 x 128
// map, 12 % of a 32 x 32 one), so a small launch adds the mirrored halo values onto them and the consumers (norm backward
// with dy_fold = 1, halo_fold with a zero halo) then stream the interior at full speed.  Streaming kernels that fold on the
// fly stall one warp per segment on its mirror loads (+15 / +27 us per 134 MB pass: tools/norm_bench.py).
// Pixel list of one sample: the 2p mirrored rows in full, then the 2p mirrored columns of the remaining rows.
__global__ void __launch_bounds__(kT) halo_accumulate_kernel(pcgan_fold_args a, int lcv, int band_px, int total_px) {
  griddep_wait();
  griddep_launch();
  const int n = blockIdx.y;
  const int cv = a.c >> 3, p = a.g_pad, h = a.h, w = a.w;
  const int v = blockIdx.x * kT + threadIdx.x;
  const int pix = v >> lcv;
  if (pix >= total_px) return;
  const int c0 = (v & (cv - 1)) << 3;
  int y, x;
  if (pix < band_px) {                       // rows 1..p and h-1-p..h-2, every column
    const int r = pix / w;
    x = pix - r * w;
    y = r < p ? 1 + r : h - 1 - p + (r - p);
  } else {                                   // the other rows: columns 1..p and w-1-p..w-2
    const int q = pix - band_px;
    const int r = q / (2 * p), k = q - r * (2 * p);
    x = k < p ? 1 + k : w - 1 - p + (k - p);
    // rows that are not mirrored rows, in order: 0, p+1 .. h-2-p, h-1
    y = r == 0 ? 0 : (r <= h - 2 * p - 2 ? p + r : h - 1);
  }
  __nv_bfloat16* g = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(a.gpad)) +
                     static_cast<int64_t>(n) * (h + 2 * p) * (w + 2 * p) * a.c + c0;
  float acc[8];
  folded_load(g, y, x, h, w, p, a.c, true, acc);
  store8(g + ((y + p) * (w + 2 * p) + x + p) * a.c, acc);
}

// -------------------------------------------------------------- norm backward
// Variants (compile time): GEN = false is the lean path of the generator (InstanceNorm without affine, no dropout mask,
// residual not needed for the activation mask): scale == rstd and shift == -mean*rstd, so the normalised value xhat IS
// the pre-activation sc*x + sh.  GEN = true carries mean / rstd / mask separately (BatchNorm with gamma, beta);
// RES adds the residual branch to the pre-activation (ResNet BasicBlock: relu(bn(x) + shortcut)).
// Streams: 0 = dy, 1 = x, 2 = residual (RES).
template <bool GEN, bool RES>
struct BwdCtx {
  float sc[8], sh[8];                               // pre = sc*(x*mask) + sh (+ residual)
  float mean[GEN ? 8 : 1], rstd[GEN ? 8 : 1], mk[GEN ? 8 : 1], pm[GEN ? 8 : 1];
  float rsc[RES ? 8 : 1], rsh[RES ? 8 : 1];
  const __nv_bfloat16* dyg;                         // sample base of dy (+c0) for the mirror loads of a folded gradient
  int fold;                                         // 0 plain, 2 reflect fold

  __device__ __forceinline__ void init(const pcgan_norm_bwd_args& a, int n, int c0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = 1.f; sh[j] = 0.f; }
#pragma unroll
    for (int j = 0; j < (GEN ? 8 : 1); ++j) { mean[j] = 0.f; rstd[j] = 0.f; mk[j] = 1.f; pm[j] = 1.f; }
#pragma unroll
    for (int j = 0; j < (RES ? 8 : 1); ++j) { rsc[j] = 1.f; rsh[j] = 0.f; }
    const int64_t so = static_cast<int64_t>(group_of(n, a.groups, a.n)) * a.c + c0;
    if (a.scale) { load_f8(a.scale + so, sc); load_f8(a.shift + so, sh); }
    if constexpr (GEN) {
      if (a.mean) { load_f8(a.mean + so, mean); load_f8(a.rstd + so, rstd); }
      if (a.drop_mask) load_f8(a.drop_mask + static_cast<int64_t>(n) * a.c + c0, mk);
      if (a.post_mask) load_f8(a.post_mask + static_cast<int64_t>(n) * a.c + c0, pm);
    }
    if constexpr (RES) {
      if (a.res_scale) {
        const int64_t ro = static_cast<int64_t>(group_of(n, a.res_groups, a.n)) * a.c + c0;
        load_f8(a.res_scale + ro, rsc);
        load_f8(a.res_shift + ro, rsh);
      }
    }
    dyg = reinterpret_cast<const __nv_bfloat16*>(a.dy) + static_cast<int64_t>(n) * (a.h + 2 * a.dy_pad) * (a.w + 2 * a.dy_pad) * a.c + c0;
    fold = a.dy_fold;
  }

  // g = dy * act'(pre) (returned in dyv) and xhat, from the streamed vectors
  __device__ __forceinline__ void grad(const pcgan_norm_bwd_args& a, int y, int x, const uint4 dyraw, const uint4 xraw,
                                       const uint4 rraw, float (&dyv)[8], float (&xh)[8]) const {
    unpack8(dyraw, dyv);
    if (fold == 2) folded_load(dyg, y, x, a.h, a.w, a.dy_pad, a.c, true, dyv, true);
    if constexpr (GEN) {
#pragma unroll
      for (int j = 0; j < 8; ++j) dyv[j] *= pm[j];
    }
    float x8[8], pre[8];
    unpack8(xraw, x8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if constexpr (GEN) {
        const float xm = x8[j] * mk[j];
        pre[j] = fmaf(sc[j], xm, sh[j]);
        xh[j] = (xm - mean[j]) * rstd[j];
      } else {
        pre[j] = fmaf(sc[j], x8[j], sh[j]);
        xh[j] = pre[j];
      }
    }
    if constexpr (RES) {
      float r8[8];
      unpack8(rraw, r8);
#pragma unroll
      for (int j = 0; j < 8; ++j) pre[j] += fmaf(rsc[j], r8[j], rsh[j]);
    }
    if (a.act == PCGAN_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) dyv[j] = pre[j] > 0.f ? dyv[j] : 0.f;
    } else if (a.act == PCGAN_ACT_LRELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) dyv[j] = pre[j] > 0.f ? dyv[j] : dyv[j] * a.act_slope;
    }
  }
};

template <bool RES>
__device__ __forceinline__ void bwd_sources(const pcgan_norm_bwd_args& a, int n, StreamSrc (&src)[RES ? 3 : 2]) {
  src[0].pad = a.dy_pad; src[0].wp = a.w + 2 * a.dy_pad;
  src[0].base = reinterpret_cast<const __nv_bfloat16*>(a.dy) + static_cast<int64_t>(n) * (a.h + 2 * a.dy_pad) * src[0].wp * a.c;
  src[1].pad = a.x_pad; src[1].wp = a.w + 2 * a.x_pad;
  src[1].base = reinterpret_cast<const __nv_bfloat16*>(a.x) + static_cast<int64_t>(n) * (a.h + 2 * a.x_pad) * src[1].wp * a.c;
  if constexpr (RES) {
    src[2].pad = a.res_pad; src[2].wp = a.w + 2 * a.res_pad;
    src[2].base = reinterpret_cast<const __nv_bfloat16*>(a.res) + static_cast<int64_t>(n) * (a.h + 2 * a.res_pad) * src[2].wp * a.c;
  }
}

template <bool GEN, bool RES, int NC>
__global__ void __launch_bounds__(NC + 32) norm_bwd_reduce_kernel(pcgan_norm_bwd_args a, SegGeom sg, int segs_per_block, int lcv) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int NT = RES ? 3 : 2;
  Pipe<NT> pipe = pipe_init<NT>(smem_raw, NC);
  const int n = blockIdx.y;
  const int32_t total_segs = a.h * sg.segs_per_row;
  const int32_t g0 = blockIdx.x * segs_per_block, g1 = min(g0 + segs_per_block, total_segs);
  const int cv = a.c >> 3;
  StreamSrc src[NT];
  bwd_sources<RES>(a, n, src);
  if (threadIdx.x >= NC) {
    if (threadIdx.x == NC) pipe_produce<NT>(pipe, sg, src, a.c, g0, g1);
  } else {
    const int c0 = (threadIdx.x & (cv - 1)) << 3;
    BwdCtx<GEN, RES> k;
    k.init(a, n, c0);
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    const int px0 = threadIdx.x >> lcv, dpx = NC >> lcv;
    const int seg_vec = sg.seg_vec, seg_px = sg.seg_px, w = a.w;
    int32_t y = g0 / sg.segs_per_row;
    int32_t x0 = (g0 - y * sg.segs_per_row) * seg_px;
    int s = 0;
    uint32_t ph = 0;
    for (int32_t g = g0; g < g1; ++g) {
      PIPE_CONSUME_BEGIN(pipe, s, ph);
      const uint4* sd = pipe.stage(s, 0);
      const uint4* sx = pipe.stage(s, 1);
      const uint4* sr = RES ? pipe.stage(s, 2) : sx;
      int px = px0;
#pragma unroll 2
      for (int v = threadIdx.x; v < seg_vec; v += NC, px += dpx) {
        float gv[8], xh[8];
        k.grad(a, y, x0 + px, sd[v], sx[v], sr[v], gv, xh);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s1[j] += gv[j]; s2[j] = fmaf(gv[j], xh[j], s2[j]); }
      }
      PIPE_CONSUME_END(pipe, s, ph);
      x0 += seg_px;
      if (x0 >= w) { x0 = 0; ++y; }
    }
    // all segments consumed: the ring is free, reuse its first NC * 64 bytes for the block reduction
    named_bar_sync(1, NC);
    float* red = reinterpret_cast<float*>(smem_raw);
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[threadIdx.x * 16 + j] = s1[j]; red[threadIdx.x * 16 + 8 + j] = s2[j]; }
    named_bar_sync(1, NC);
    const int lanes = NC / cv;
    for (int t = threadIdx.x; t < cv * 16; t += NC) {
      const int c = t >> 4, slot = t & 15;
      float sum = 0.f;
      for (int l = 0; l < lanes; ++l) sum += red[(l * cv + c) * 16 + slot];
      const int ch = (c << 3) + (slot & 7);
      const int64_t o = (static_cast<int64_t>(group_of(n, a.groups, a.n)) * a.c + ch) * 2 + (slot >> 3);
      atomicAdd(a.sums + o, sum);
    }
  }
}

template <bool GEN, bool RES, int NC>
__global__ void __launch_bounds__(NC + 32) norm_bwd_apply_kernel(pcgan_norm_bwd_args a, SegGeom sg, int segs_per_block, int lcv) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int NT = RES ? 3 : 2;
  Pipe<NT> pipe = pipe_init<NT>(smem_raw, NC);
  // blocks walk the samples (and chunks) in the opposite order to the reduce pass: what that pass read last is
  // still in L2 when this one starts
  const int n = gridDim.y - 1 - blockIdx.y;
  const int bx = gridDim.x - 1 - blockIdx.x;
  const int32_t total_segs = a.h * sg.segs_per_row;
  const int32_t g0 = bx * segs_per_block, g1 = min(g0 + segs_per_block, total_segs);
  const int cv = a.c >> 3;
  StreamSrc src[NT];
  bwd_sources<RES>(a, n, src);
  if (threadIdx.x >= NC) {
    if (threadIdx.x == NC) pipe_produce<NT>(pipe, sg, src, a.c, g0, g1);
    return;
  }
  const int c0 = (threadIdx.x & (cv - 1)) << 3;
  BwdCtx<GEN, RES> k;
  k.init(a, n, c0);
  float A[8], B[8];
  const float inv = a.count > 0.f ? 1.f / a.count : 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { A[j] = 0.f; B[j] = 0.f; }
  if (a.count > 0.f) {
    const int64_t so = (static_cast<int64_t>(group_of(n, a.groups, a.n)) * a.c + c0) * 2;
    float t0[8], t1[8];
    load_f8(a.sums + so, t0);
    load_f8(a.sums + so + 8, t1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      A[j] = t0[2 * j] * inv; B[j] = t0[2 * j + 1] * inv;
      A[4 + j] = t1[2 * j] * inv; B[4 + j] = t1[2 * j + 1] * inv;
    }
  }
  const bool scaled = a.scale != nullptr;
  const int wop = a.w + 2 * a.dx_pad, wsp = a.w + 2 * a.dres_pad, c = a.c, w = a.w;
  __nv_bfloat16* dx = a.dx ? reinterpret_cast<__nv_bfloat16*>(a.dx) + static_cast<int64_t>(n) * (a.h + 2 * a.dx_pad) * wop * c + c0 : nullptr;
  __nv_bfloat16* dres = a.dres ? reinterpret_cast<__nv_bfloat16*>(a.dres) + static_cast<int64_t>(n) * (a.h + 2 * a.dres_pad) * wsp * c + c0 : nullptr;
  const int px0 = threadIdx.x >> lcv, dpx = NC >> lcv;
  const int seg_vec = sg.seg_vec, seg_px = sg.seg_px;
  int32_t y = g0 / sg.segs_per_row;
  int32_t x0 = (g0 - y * sg.segs_per_row) * seg_px;
  int s = 0;
  uint32_t ph = 0;
  for (int32_t g = g0; g < g1; ++g) {
    // element offsets fit 31 bits (checked on the host)
    __nv_bfloat16* dxrow = dx ? dx + ((y + a.dx_pad) * wop + x0 + a.dx_pad) * c : nullptr;
    __nv_bfloat16* drrow = dres ? dres + ((y + a.dres_pad) * wsp + x0 + a.dres_pad) * c : nullptr;
    PIPE_CONSUME_BEGIN(pipe, s, ph);
    const uint4* sd = pipe.stage(s, 0);
    const uint4* sx = pipe.stage(s, 1);
    const uint4* sr = RES ? pipe.stage(s, 2) : sx;
    int px = px0;
#pragma unroll 2
    for (int v = threadIdx.x; v < seg_vec; v += NC, px += dpx) {
      float gv[8], xh[8];
      k.grad(a, y, x0 + px, sd[v], sx[v], sr[v], gv, xh);
      if (drrow) store8(drrow + px * c, gv);
      if (dxrow) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = gv[j] - A[j] - xh[j] * B[j];
          if (scaled) t *= k.sc[j];
          if constexpr (GEN) t *= k.mk[j];
          o[j] = t;
        }
        store8(dxrow + px * c, o);
      }
    }
    PIPE_CONSUME_END(pipe, s, ph);
    x0 += seg_px;
    if (x0 >= w) { x0 = 0; ++y; }
  }
}

// ------------------------------------------------------ one-pass norm backward
// Lean InstanceNorm path (affine == 0, no masks; the generator's ResnetBlocks, networks.py:621-652).  The two-pass
// backward reads dy and x twice (reduce, then apply: 5 tensor passes over HBM); here a CLUSTER of kFusedCL CTAs owns one
// sample, each CTA keeps its h / kFusedCL image rows of dy and x in shared memory (bulk async copies, read from HBM
// once), the per-channel sums are reduced across the cluster through distributed shared memory, and dx is produced from
// the resident copies: 3 tensor passes, one launch.  Under dy_fold == 2 the streamed dy is the interior of a padded-grid
// gradient and the mirrored halo values are added from global memory (L2) for the few border pixels, exactly as in the
// two-pass kernels.
static constexpr int kFusedCL = 8;
static constexpr int kFusedT = 1024;  // 32 warps: one CTA per SM holds 128 KB of rows, so the warps of that one CTA must hide the latencies
static constexpr int kFusedMaxTensorBytes = 64 * 1024;   // per CTA and tensor

__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 dsmem_ld_f32x4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

__global__ void __launch_bounds__(kFusedT, 1) norm_bwd_fused_kernel(pcgan_norm_bwd_args a, int rows, int lcv) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int n = blockIdx.y;
  const uint32_t rank = cluster_ctarank();
  const int y0 = static_cast<int>(rank) * rows;
  const int cv = a.c >> 3;
  const int row_vec = a.w * cv;            // 16-byte vectors per image row
  const int nvec = rows * row_vec;
  uint4* sX = reinterpret_cast<uint4*>(smem_raw);
  uint4* sG = sX + nvec;
  float* red = reinterpret_cast<float*>(sG + nvec);            // [kFusedT][16]
  float* part = red + kFusedT * 16;                            // [c][2]: this CTA's (sum g, sum g*xhat)
  uint64_t* bar = reinterpret_cast<uint64_t*>(part + a.c * 2);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();
  griddep_wait();
  griddep_launch();
  const int wdp = a.w + 2 * a.dy_pad, wxp = a.w + 2 * a.x_pad;
  const __nv_bfloat16* dyn = reinterpret_cast<const __nv_bfloat16*>(a.dy) + static_cast<int64_t>(n) * (a.h + 2 * a.dy_pad) * wdp * a.c;
  const __nv_bfloat16* xn = reinterpret_cast<const __nv_bfloat16*>(a.x) + static_cast<int64_t>(n) * (a.h + 2 * a.x_pad) * wxp * a.c;
  if (threadIdx.x == 0) {
    const uint32_t row_bytes = static_cast<uint32_t>(row_vec) * 16u;
    mbar_arrive_expect_tx(bar, 2u * row_bytes * rows);
    for (int r = 0; r < rows; ++r) {
      const int y = y0 + r;
      bulk_load(sX + r * row_vec, xn + (static_cast<int64_t>(y + a.x_pad) * wxp + a.x_pad) * a.c, row_bytes, bar);
      bulk_load(sG + r * row_vec, dyn + (static_cast<int64_t>(y + a.dy_pad) * wdp + a.dy_pad) * a.c, row_bytes, bar);
    }
  }
  const int c0 = (threadIdx.x & (cv - 1)) << 3;
  float sc[8], sh[8];
  {
    const int64_t so = static_cast<int64_t>(group_of(n, a.groups, a.n)) * a.c + c0;
    load_f8(a.scale + so, sc);
    load_f8(a.shift + so, sh);
  }
  const bool fold = a.dy_fold == 2;
  const bool relu = a.act == PCGAN_ACT_RELU, lrelu = a.act == PCGAN_ACT_LRELU;
  const __nv_bfloat16* dyg = dyn + c0;
  mbar_wait(bar, 0);

  // g = dy * act'(pre) and xhat = pre for vector v of this CTA's rows
  auto grad = [&](int v, float (&g)[8], float (&xh)[8]) {
    unpack8(sG[v], g);
    if (fold) {
      const int pix = v >> lcv;
      const int r = pix / a.w;
      folded_load(dyg, y0 + r, pix - r * a.w, a.h, a.w, a.dy_pad, a.c, true, g, true);
    }
    float x8[8];
    unpack8(sX[v], x8);
#pragma unroll
    for (int j = 0; j < 8; ++j) xh[j] = fmaf(sc[j], x8[j], sh[j]);
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = xh[j] > 0.f ? g[j] : 0.f;
    } else if (lrelu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = xh[j] > 0.f ? g[j] : g[j] * a.act_slope;
    }
  };

  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  for (int v = threadIdx.x; v < nvec; v += kFusedT) {
    float g[8], xh[8];
    grad(v, g, xh);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] += g[j]; s2[j] = fmaf(g[j], xh[j], s2[j]); }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[threadIdx.x * 16 + j] = s1[j]; red[threadIdx.x * 16 + 8 + j] = s2[j]; }
  __syncthreads();
  const int lanes = kFusedT / cv;
  for (int t = threadIdx.x; t < cv * 16; t += kFusedT) {
    const int c = t >> 4, slot = t & 15;
    float sum = 0.f;
    for (int l = 0; l < lanes; ++l) sum += red[(l * cv + c) * 16 + slot];
    part[((c << 3) + (slot & 7)) * 2 + (slot >> 3)] = sum;
  }
  cluster_sync_all();      // every CTA's partial sums are in place (release / acquire at cluster scope)
  // a few threads fetch the peers' partial sums (remote shared-memory requests are expensive: one float4 per thread and
  // peer, not one set per consumer thread) and leave the cluster totals / count in local shared memory
  float* tot = red;        // the block-reduction scratch is free again
  if (static_cast<int>(threadIdx.x) * 4 < a.c * 2) {
    const uint32_t mine = smem_u32(part + threadIdx.x * 4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (uint32_t r = 0; r < kFusedCL; ++r) {
      const float4 t = dsmem_ld_f32x4(dsmem_addr(mine, r));
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    const float inv = 1.f / a.count;
    reinterpret_cast<float4*>(tot)[threadIdx.x] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  }
  cluster_sync_all();      // nobody overwrites or leaves while a peer may still read its partial sums; `tot` is visible
  float A[8], B[8];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 t = reinterpret_cast<const float4*>(tot + c0 * 2)[q];
    A[2 * q] = t.x; B[2 * q] = t.y; A[2 * q + 1] = t.z; B[2 * q + 1] = t.w;
  }
  const int wop = a.w + 2 * a.dx_pad;
  __nv_bfloat16* dx = reinterpret_cast<__nv_bfloat16*>(a.dx) + static_cast<int64_t>(n) * (a.h + 2 * a.dx_pad) * wop * a.c + c0;
  for (int v = threadIdx.x; v < nvec; v += kFusedT) {
    float g[8], xh[8], o[8];
    grad(v, g, xh);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (g[j] - A[j] - xh[j] * B[j]) * sc[j];
    const int pix = v >> lcv;
    const int r = pix / a.w;
    store8(dx + (static_cast<int64_t>(y0 + r + a.dx_pad) * wop + (pix - r * a.w) + a.dx_pad) * a.c, o);
  }
}

static size_t fused_smem(int rows, int w, int c) {
  return 2 * static_cast<size_t>(rows) * w * c * 2 + kFusedT * 16 * sizeof(float) + static_cast<size_t>(c) * 2 * sizeof(float) + 16;
}

// ----------------------------------------------------------------------- host
static int log2_pow2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return (1 << l) == v ? l : -1;
}
static int check_c(int c, const char* who, int* lcv) {
  if (c < 8 || c % 8 != 0) return fail(PCGAN_ERR_UNSUPPORTED, "%s: channels=%d must be a multiple of 8", who, c);
  const int l = log2_pow2(c / 8);
  if (l < 0 || c / 8 > kT) return fail(PCGAN_ERR_UNSUPPORTED, "%s: channels/8=%d must be a power of two <= %d", who, c / 8, kT);
  *lcv = l;
  return PCGAN_OK;
}
// chunks of a sample so that the whole grid is a few waves of blocks, each with a useful amount of work
static int chunking(int64_t vec_per_sample, int n, int unit, int* per_block) {
  const int sms = sm_count() > 0 ? sm_count() : 148;
  int64_t want_blocks = static_cast<int64_t>(sms) * 16;        // 2 waves at 8 blocks / SM
  int64_t chunks = (want_blocks + n - 1) / n;
  int64_t per = (vec_per_sample + chunks - 1) / chunks;
  const int64_t min_per = static_cast<int64_t>(kT) * 8;        // at least 8 vectors per thread
  if (per < min_per) per = min_per;
  per = (per + unit - 1) / unit * unit;
  *per_block = static_cast<int>(per);
  return static_cast<int>((vec_per_sample + per - 1) / per);
}

// Segments of the stream pipeline: the largest divisor of the row width whose bytes fit one stage.
static SegGeom seg_geom(int w, int c) {
  SegGeom sg;
  int best = 1;
  for (int d = 1; d <= w; ++d)
    if (w % d == 0 && static_cast<int64_t>(d) * c * 2 <= kSegBytes) best = d;
  sg.seg_px = best;
  sg.segs_per_row = w / best;
  sg.seg_vec = best * (c / 8);
  return sg;
}
// segments per block: ONE wave of blocks (`blocks_per_sm` of them are resident per SM) whenever a block then still gets
// a handful of segments — a second, partly filled wave costs a whole block time on these short kernels — and never
// fewer than two rings of segments per block
static int seg_chunking(int total_segs, int n, int blocks_per_sm, int* segs_per_block) {
  const int sms = sm_count() > 0 ? sm_count() : 148;
  const int64_t slots = static_cast<int64_t>(sms) * blocks_per_sm;
  int64_t chunks = slots / n;                 // chunks per sample that fit one wave
  if (chunks < 1) chunks = 1;
  int64_t per = (total_segs + chunks - 1) / chunks;
  if (per < 2 * kStages) per = 2 * kStages;
  if (per > total_segs) per = total_segs;
  *segs_per_block = static_cast<int>(per);
  return static_cast<int>((total_segs + per - 1) / per);
}
template <int NT>
static constexpr size_t pipe_smem() { return static_cast<size_t>(kStages) * NT * kSegBytes + 2 * kStages * sizeof(uint64_t); }

template <typename K>
static int set_smem(K kernel, size_t bytes, const char* name) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) return fail(PCGAN_ERR_CUDA, "cudaFuncSetAttribute(%s): %s", name, cudaGetErrorString(e));
  return PCGAN_OK;
}
// one-time opt-in to > 48 KB of dynamic shared memory for every instantiation
static int norm_kernels_ready() {
  static std::once_flag once;
  static int rc = PCGAN_OK;
  std::call_once(once, []() {
    int r;
    if ((r = set_smem(norm_apply_kernel<false, -1>, pipe_smem<1>(), "norm_apply<0>"))) { rc = r; return; }
    if ((r = set_smem(norm_apply_kernel<false, PCGAN_ACT_NONE>, pipe_smem<1>(), "norm_apply<0,none>"))) { rc = r; return; }
    if ((r = set_smem(norm_apply_kernel<false, PCGAN_ACT_RELU>, pipe_smem<1>(), "norm_apply<0,relu>"))) { rc = r; return; }
    if ((r = set_smem(norm_apply_kernel<false, PCGAN_ACT_LRELU>, pipe_smem<1>(), "norm_apply<0,lrelu>"))) { rc = r; return; }
    if ((r = set_smem(norm_apply_kernel<true, -1>, pipe_smem<2>(), "norm_apply<1>"))) { rc = r; return; }
    if ((r = set_smem(norm_apply_kernel<true, PCGAN_ACT_NONE>, pipe_smem<2>(), "norm_apply<1,none>"))) { rc = r; return; }
    if ((r = set_smem(norm_apply_kernel<true, PCGAN_ACT_RELU>, pipe_smem<2>(), "norm_apply<1,relu>"))) { rc = r; return; }
    if ((r = set_smem(norm_bwd_reduce_kernel<false, false, kReduceConsumers>, pipe_smem<2>(), "norm_bwd_reduce<0,0>"))) { rc = r; return; }
    if ((r = set_smem(norm_bwd_reduce_kernel<true, false, kReduceConsumers>, pipe_smem<2>(), "norm_bwd_reduce<1,0>"))) { rc = r; return; }
    if ((r = set_smem(norm_bwd_reduce_kernel<true, true, kReduceConsumers>, pipe_smem<3>(), "norm_bwd_reduce<1,1>"))) { rc = r; return; }
    if ((r = set_smem(norm_bwd_apply_kernel<false, false, kConsumers>, pipe_smem<2>(), "norm_bwd_apply<0,0>"))) { rc = r; return; }
    if ((r = set_smem(norm_bwd_apply_kernel<true, false, kConsumers>, pipe_smem<2>(), "norm_bwd_apply<1,0>"))) { rc = r; return; }
    if ((r = set_smem(norm_bwd_apply_kernel<true, true, kConsumers>, pipe_smem<3>(), "norm_bwd_apply<1,1>"))) { rc = r; return; }
  });
  return rc;
}

}  // namespace pcgan

using namespace pcgan;
#define STREAM(s) static_cast<cudaStream_t>(s)

extern "C" int pcgan_norm_finalize(const pcgan_norm_finalize_args* a, pcgan_stream_t s) {
  if (!a || !a->stats || a->groups < 1 || a->c < 1 || a->count <= 0.f) return fail(PCGAN_ERR_INVALID, "norm_finalize: bad argument");
  if (a->in_groups > a->groups && a->groups != 1) return fail(PCGAN_ERR_INVALID, "norm_finalize: per-sample statistics combine into exactly one group");
  if (a->drop_mask && a->in_groups <= a->groups) return fail(PCGAN_ERR_INVALID, "norm_finalize: drop_mask needs per-sample statistics (in_groups = N)");
  PCGAN_CUDA_OK(launch_pdl(norm_finalize_kernel, dim3((a->c + 31) / 32), dim3(kT), 0, STREAM(s), 1, *a));
  PCGAN_LAUNCH_OK("norm_finalize_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_norm_running_batched(const pcgan_running_item* items, int32_t count, int32_t max_c, pcgan_stream_t s) {
  if (!items || count < 1 || count > 65535 || max_c < 1) return fail(PCGAN_ERR_INVALID, "norm_running_batched: bad argument");
  PCGAN_CUDA_OK(launch_pdl(norm_running_batched_kernel, dim3((max_c + kT - 1) / kT, count), dim3(kT), 0, STREAM(s), 1, items));
  PCGAN_LAUNCH_OK("norm_running_batched_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_norm_apply(const pcgan_norm_apply_args* a, pcgan_stream_t s) {
  if (!a || !a->x || !a->y) return fail(PCGAN_ERR_INVALID, "norm_apply: null argument");
  if (a->stats && a->count <= 0.f) return fail(PCGAN_ERR_INVALID, "norm_apply: fused finalize needs count > 0");
  int lcv, rc = check_c(a->c, "norm_apply", &lcv);
  if (rc) return rc;
  if (a->c / 8 > kApplyConsumers) return fail(PCGAN_ERR_UNSUPPORTED, "norm_apply: channels=%d (<= %d)", a->c, 8 * kApplyConsumers);
  if ((a->scale == nullptr) != (a->shift == nullptr)) return fail(PCGAN_ERR_INVALID, "norm_apply: scale and shift go together");
  if (a->y_halo == PCGAN_HALO_REFLECT && (2 * a->y_pad + 1 > a->h || 2 * a->y_pad + 1 > a->w)) return fail(PCGAN_ERR_INVALID, "norm_apply: image smaller than 2*pad+1 under reflect halo");
  if (a->n < 1 || a->n > 65535) return fail(PCGAN_ERR_UNSUPPORTED, "norm_apply: n=%d (1..65535)", a->n);
  if (a->y_c != 0 && (a->y_c < a->c + a->y_c0 || a->y_c % 8 != 0 || a->y_c0 % 8 != 0 || a->y_c0 < 0))
    return fail(PCGAN_ERR_INVALID, "norm_apply: channel slice [%d, %d) of %d channels", a->y_c0, a->y_c0 + a->c, a->y_c);
  const int mp = a->y_pad > a->x_pad ? (a->y_pad > a->res_pad ? a->y_pad : a->res_pad) : (a->x_pad > a->res_pad ? a->x_pad : a->res_pad);
  if (static_cast<int64_t>(a->h + 2 * mp) * (a->w + 2 * mp) * (a->y_c > a->c ? a->y_c : a->c) >= (1ll << 31)) return fail(PCGAN_ERR_UNSUPPORTED, "norm_apply: sample too large");
  if ((rc = norm_kernels_ready())) return rc;
  const SegGeom sg = seg_geom(a->w, a->c);
  int per;
  const int chunks = seg_chunking(a->h * sg.segs_per_row, a->n, a->res ? 3 : 6, &per);
  const dim3 grid(chunks, a->n);
  const dim3 blk(kApplyThreads);
#define PCGAN_APPLY(RES, ACT, NT) PCGAN_CUDA_OK(launch_pdl(norm_apply_kernel<RES, ACT>, grid, blk, pipe_smem<NT>(), STREAM(s), 1, *a, sg, per, lcv))
  if (a->res) {
    if (a->act == PCGAN_ACT_NONE) PCGAN_APPLY(true, PCGAN_ACT_NONE, 2);
    else if (a->act == PCGAN_ACT_RELU) PCGAN_APPLY(true, PCGAN_ACT_RELU, 2);
    else PCGAN_APPLY(true, -1, 2);
  } else {
    if (a->act == PCGAN_ACT_NONE) PCGAN_APPLY(false, PCGAN_ACT_NONE, 1);
    else if (a->act == PCGAN_ACT_RELU) PCGAN_APPLY(false, PCGAN_ACT_RELU, 1);
    else if (a->act == PCGAN_ACT_LRELU) PCGAN_APPLY(false, PCGAN_ACT_LRELU, 1);
    else PCGAN_APPLY(false, -1, 1);
  }
#undef PCGAN_APPLY
  PCGAN_LAUNCH_OK("norm_apply_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_halo_fold(const pcgan_fold_args* a, pcgan_stream_t s) {
  if (!a || !a->gpad || !a->out) return fail(PCGAN_ERR_INVALID, "halo_fold: null argument");
  int lcv, rc = check_c(a->c, "halo_fold", &lcv);
  if (rc) return rc;
  if (a->halo == PCGAN_HALO_REFLECT && (2 * a->g_pad + 1 > a->h || 2 * a->g_pad + 1 > a->w)) return fail(PCGAN_ERR_UNSUPPORTED, "halo_fold: image smaller than 2*pad+1");
  if (a->n < 1 || a->n > 65535) return fail(PCGAN_ERR_UNSUPPORTED, "halo_fold: n=%d (1..65535)", a->n);
  const int64_t vps = static_cast<int64_t>(a->h) * a->w * (a->c / 8);
  if (static_cast<int64_t>(a->h + 2 * a->g_pad) * (a->w + 2 * a->g_pad) * a->c >= (1ll << 31)) return fail(PCGAN_ERR_UNSUPPORTED, "halo_fold: sample too large");
  int per;
  const int chunks = chunking(vps, a->n, kT, &per);
  PCGAN_CUDA_OK(launch_pdl(halo_fold_kernel, dim3(chunks, a->n), dim3(kT), 0, STREAM(s), 1, *a, per, lcv));
  PCGAN_LAUNCH_OK("halo_fold_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_halo_accumulate(const pcgan_fold_args* a, pcgan_stream_t s) {
  if (!a || !a->gpad) return fail(PCGAN_ERR_INVALID, "halo_accumulate: null argument");
  int lcv, rc = check_c(a->c, "halo_accumulate", &lcv);
  if (rc) return rc;
  const int p = a->g_pad;
  if (p < 1 || 2 * p + 2 > a->h || 2 * p + 2 > a->w) return fail(PCGAN_ERR_UNSUPPORTED, "halo_accumulate: needs pad >= 1 and an image of at least 2*pad+2");
  if (a->n < 1 || a->n > 65535) return fail(PCGAN_ERR_UNSUPPORTED, "halo_accumulate: n=%d (1..65535)", a->n);
  if (static_cast<int64_t>(a->h + 2 * p) * (a->w + 2 * p) * a->c >= (1ll << 31)) return fail(PCGAN_ERR_UNSUPPORTED, "halo_accumulate: sample too large");
  const int band_px = 2 * p * a->w, total_px = band_px + (a->h - 2 * p) * 2 * p;
  const int64_t vec = static_cast<int64_t>(total_px) * (a->c / 8);
  PCGAN_CUDA_OK(launch_pdl(halo_accumulate_kernel, dim3(static_cast<unsigned>((vec + kT - 1) / kT), a->n), dim3(kT), 0, STREAM(s), 1, *a, lcv, band_px, total_px));
  PCGAN_LAUNCH_OK("halo_accumulate_kernel");
  return PCGAN_OK;
}

static int check_bwd(const pcgan_norm_bwd_args* a, int* lcv) {
  if (!a || !a->dy || !a->x) return fail(PCGAN_ERR_INVALID, "norm_bwd: null argument");
  int rc = check_c(a->c, "norm_bwd", lcv);
  if (rc) return rc;
  if (a->count > 0.f && (!a->mean || !a->rstd || !a->sums)) return fail(PCGAN_ERR_INVALID, "norm_bwd: statistics missing");
  if ((a->scale == nullptr) != (a->shift == nullptr)) return fail(PCGAN_ERR_INVALID, "norm_bwd: scale and shift go together");
  if (a->n < 1 || a->n > 65535) return fail(PCGAN_ERR_UNSUPPORTED, "norm_bwd: n=%d (1..65535)", a->n);
  if (a->dy_fold < 0 || a->dy_fold > 2) return fail(PCGAN_ERR_INVALID, "norm_bwd: dy_fold must be 0 (plain), 1 (zero halo dropped) or 2 (reflect fold)");
  if (a->dy_fold == 2 && (2 * a->dy_pad + 1 > a->h || 2 * a->dy_pad + 1 > a->w)) return fail(PCGAN_ERR_UNSUPPORTED, "norm_bwd: image smaller than 2*pad+1");
  int mp = a->dy_pad > a->x_pad ? a->dy_pad : a->x_pad;
  if (a->res_pad > mp) mp = a->res_pad;
  if (a->dx_pad > mp) mp = a->dx_pad;
  if (static_cast<int64_t>(a->h + 2 * mp) * (a->w + 2 * mp) * a->c >= (1ll << 31)) return fail(PCGAN_ERR_UNSUPPORTED, "norm_bwd: sample too large");
  return norm_kernels_ready();
}

extern "C" int pcgan_norm_bwd_reduce(const pcgan_norm_bwd_args* a, pcgan_stream_t s) {
  int lcv, rc = check_bwd(a, &lcv);
  if (rc) return rc;
  if (!a->sums) return fail(PCGAN_ERR_INVALID, "norm_bwd_reduce: sums is null");
  const bool res = a->res != nullptr && a->act != PCGAN_ACT_NONE;   // the residual only matters through the activation mask
  const bool gen = a->affine != 0 || a->drop_mask != nullptr || a->post_mask != nullptr || res;
  const SegGeom sg = seg_geom(a->w, a->c);
  int per;
  const int chunks = seg_chunking(a->h * sg.segs_per_row, a->n, gen ? 2 : 3, &per);
  const dim3 grid(chunks, a->n);
  const dim3 blk(kReduceConsumers + 32);
  if (a->c / 8 > kReduceConsumers) return fail(PCGAN_ERR_UNSUPPORTED, "norm_bwd_reduce: channels=%d (<= %d)", a->c, 8 * kReduceConsumers);
  if (!gen) PCGAN_CUDA_OK(launch_pdl(norm_bwd_reduce_kernel<false, false, kReduceConsumers>, grid, blk, pipe_smem<2>(), STREAM(s), 1, *a, sg, per, lcv));
  else if (!res) PCGAN_CUDA_OK(launch_pdl(norm_bwd_reduce_kernel<true, false, kReduceConsumers>, grid, blk, pipe_smem<2>(), STREAM(s), 1, *a, sg, per, lcv));
  else PCGAN_CUDA_OK(launch_pdl(norm_bwd_reduce_kernel<true, true, kReduceConsumers>, grid, blk, pipe_smem<3>(), STREAM(s), 1, *a, sg, per, lcv));
  PCGAN_LAUNCH_OK("norm_bwd_reduce_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_norm_bwd_apply(const pcgan_norm_bwd_args* a, pcgan_stream_t s) {
  int lcv, rc = check_bwd(a, &lcv);
  if (rc) return rc;
  if (!a->dx && !a->dres) return fail(PCGAN_ERR_INVALID, "norm_bwd_apply: no output");
  const bool res = a->res != nullptr && a->act != PCGAN_ACT_NONE;
  const bool gen = a->affine != 0 || a->drop_mask != nullptr || a->post_mask != nullptr || res;
  const SegGeom sg = seg_geom(a->w, a->c);
  int per;
  const int chunks = seg_chunking(a->h * sg.segs_per_row, a->n, 2, &per);
  const dim3 grid(chunks, a->n);
  const dim3 blk(kStreamThreads);
  if (!gen) PCGAN_CUDA_OK(launch_pdl(norm_bwd_apply_kernel<false, false, kConsumers>, grid, blk, pipe_smem<2>(), STREAM(s), 1, *a, sg, per, lcv));
  else if (!res) PCGAN_CUDA_OK(launch_pdl(norm_bwd_apply_kernel<true, false, kConsumers>, grid, blk, pipe_smem<2>(), STREAM(s), 1, *a, sg, per, lcv));
  else PCGAN_CUDA_OK(launch_pdl(norm_bwd_apply_kernel<true, true, kConsumers>, grid, blk, pipe_smem<3>(), STREAM(s), 1, *a, sg, per, lcv));
  PCGAN_LAUNCH_OK("norm_bwd_apply_kernel");
  return PCGAN_OK;
}

static int fused_active_clusters = -1;
/* clusters of the one-pass kernel the device holds at once at the ResnetBlock shape (diagnostic; -1 before the first launch) */
extern "C" int pcgan_norm_bwd_fused_active_clusters(void) { return fused_active_clusters; }

extern "C" int pcgan_norm_bwd_fused_supported(const pcgan_norm_bwd_args* a) {
  if (!a || !a->dy || !a->x || !a->dx || a->dres) return 0;
  if (a->affine != 0 || a->drop_mask || a->post_mask || !a->scale || !a->shift || a->count <= 0.f) return 0;
  if (a->res && a->act != PCGAN_ACT_NONE) return 0;
  if (a->act != PCGAN_ACT_NONE && a->act != PCGAN_ACT_RELU && a->act != PCGAN_ACT_LRELU) return 0;
  if (a->dy_fold != 0 && a->dy_fold != 2) return 0;
  if (a->c < 8 || a->c % 8 != 0 || log2_pow2(a->c / 8) < 0 || a->c / 8 > kFusedT || a->c > 2048) return 0;
  if (a->h % kFusedCL != 0 || a->n < 1 || a->n > 65535) return 0;
  const int rows = a->h / kFusedCL;
  if (static_cast<int64_t>(rows) * a->w * a->c * 2 > kFusedMaxTensorBytes) return 0;
  if (a->dy_fold == 2 && (2 * a->dy_pad + 1 > a->h || 2 * a->dy_pad + 1 > a->w)) return 0;
  return 1;
}

extern "C" int pcgan_norm_bwd_fused(const pcgan_norm_bwd_args* a, pcgan_stream_t s) {
  int lcv, rc = check_bwd(a, &lcv);
  if (rc) return rc;
  if (!pcgan_norm_bwd_fused_supported(a)) return fail(PCGAN_ERR_UNSUPPORTED, "norm_bwd_fused: lean InstanceNorm path, h %% %d == 0, %d KB of rows per CTA", kFusedCL, kFusedMaxTensorBytes / 1024);
  const int rows = a->h / kFusedCL;
  const size_t smem = fused_smem(rows, a->w, a->c);
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, []() {
    // two resident tensors + the block-reduction scratch + per-channel partial sums of up to 2048 channels
    attr_err = cudaFuncSetAttribute(norm_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    2 * kFusedMaxTensorBytes + kFusedT * 16 * 4 + 2048 * 2 * 4 + 16);
    if (attr_err == cudaSuccess && kFusedCL > 8)
      attr_err = cudaFuncSetAttribute(norm_bwd_fused_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (attr_err == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(kFusedCL, 64);
      cfg.blockDim = dim3(kFusedT);
      cfg.dynamicSmemBytes = 2 * kFusedMaxTensorBytes + kFusedT * 16 * 4 + 256 * 2 * 4 + 16;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = kFusedCL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, norm_bwd_fused_kernel, &cfg) == cudaSuccess) fused_active_clusters = nc;
      else (void)cudaGetLastError();
    }
  });
  if (attr_err != cudaSuccess) return fail(PCGAN_ERR_CUDA, "cudaFuncSetAttribute(norm_bwd_fused): %s", cudaGetErrorString(attr_err));
  PCGAN_CUDA_OK(launch_pdl(norm_bwd_fused_kernel, dim3(kFusedCL, a->n), dim3(kFusedT), smem, STREAM(s), kFusedCL, *a, rows, lcv));
  PCGAN_LAUNCH_OK("norm_bwd_fused_kernel");
  return PCGAN_OK;
}
