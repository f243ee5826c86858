// HBM-bound kernels of the wsgan_emb step: layout packing, normalisation forward and
// backward, halo folding, pooling, losses, Adam.  All activations are NHWC bf16 in
// physically padded buffers; every kernel moves 16-byte vectors (8 channels) per
// thread with consecutive threads on consecutive channel vectors (coalesced), and
// grids are sized in whole waves of the SM count.
#include "common.cuh"

namespace pcgan {

static constexpr int kThreads = 256;

static inline int grid_for(int64_t work_items, int per_block = kThreads, int waves = 8) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  int64_t cap = static_cast<int64_t>(sm_count() > 0 ? sm_count() : 148) * waves;
  if (blocks > cap) blocks = cap;  // grid-stride loops cover the rest
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// ------------------------------------------------------------ gather / scatter
__global__ void gather_cast_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx,
                                   __nv_bfloat16* __restrict__ dst, int64_t n) {
  griddep_wait();
  griddep_launch();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t j = idx[i];
    dst[i] = __float2bfloat16(j >= 0 ? src[j] : 0.f);
  }
}
// the fp32 packed operand of a TF32 plan, rounded to nearest TF32 here so that the tensor core's truncation of the low
// mantissa bits loses nothing more
__global__ void gather_tf32_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, float* __restrict__ dst, int64_t n) {
  griddep_wait();
  griddep_launch();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t j = idx[i];
    uint32_t r;
    const float v = j >= 0 ? src[j] : 0.f;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    dst[i] = __uint_as_float(r);
  }
}
__global__ void scatter_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, float* __restrict__ dst,
                               int64_t n, int accumulate) {
  griddep_wait();
  griddep_launch();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t j = idx[i];
    if (j >= 0) dst[j] = accumulate ? dst[j] + src[i] : src[i];
  }
}

// one launch for a whole network: blockIdx.y = tensor
__global__ void gather_cast_batched_kernel(const pcgan_batch_item* __restrict__ items) {
  griddep_wait();
  griddep_launch();
  const pcgan_batch_item it = items[blockIdx.y];
  const float* __restrict__ src = it.src;
  const int32_t* __restrict__ idx = it.idx;
  __nv_bfloat16* __restrict__ dst = reinterpret_cast<__nv_bfloat16*>(it.dst);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < it.n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t j = idx[i];
    dst[i] = __float2bfloat16(j >= 0 ? src[j] : 0.f);
  }
}
__global__ void scatter_batched_kernel(const pcgan_batch_item* __restrict__ items, int accumulate) {
  griddep_wait();
  griddep_launch();
  const pcgan_batch_item it = items[blockIdx.y];
  const float* __restrict__ src = it.src;
  const int32_t* __restrict__ idx = it.idx;
  float* __restrict__ dst = reinterpret_cast<float*>(it.dst);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < it.n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t j = idx[i];
    if (j >= 0) dst[j] = accumulate ? dst[j] + src[i] : src[i];
  }
}

// ------------------------------------------------------------------ pack_nchw
__device__ __forceinline__ void bilinear_setup(int o, int in, int out, int& i0, int& i1, float& l0, float& l1) {
  // torch.nn.functional.interpolate(mode='bilinear', align_corners=True): util/util.py:117
  const float scale = out > 1 ? static_cast<float>(in - 1) / static_cast<float>(out - 1) : 0.f;
  const float s = scale * static_cast<float>(o);
  i0 = static_cast<int>(s);
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 < in - 1 ? i0 + 1 : i0;
  l1 = s - static_cast<float>(i0);
  l0 = 1.f - l1;
}

// grid (x blocks, padded rows, samples): no per-thread division; one 16-byte store per 8 destination channels
__global__ void pack_nchw_kernel(pcgan_pack_args a) {
  griddep_wait();
  griddep_launch();
  const int hp = a.ho + 2 * a.pad, wp = a.wo + 2 * a.pad;
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  const int py = blockIdx.y;
  const int n = blockIdx.z;
  if (px >= wp) return;
  const int64_t nstride = a.dst_n_stride ? a.dst_n_stride : static_cast<int64_t>(hp) * wp * a.cd;
  const bool resize = (a.ho != a.h) || (a.wo != a.w);
  int y = py - a.pad, x = px - a.pad;
  const bool halo = y < 0 || y >= a.ho || x < 0 || x >= a.wo;
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(a.dst) + n * nstride + (static_cast<int64_t>(py) * wp + px) * a.cd;
  if (halo && a.halo == PCGAN_HALO_ZERO) {
    for (int c = 0; c < a.cd; c += 8) *reinterpret_cast<uint4*>(d + c) = make_uint4(0, 0, 0, 0);
    return;
  }
  y = reflect_idx(y, a.ho);
  x = reflect_idx(x, a.wo);
  int y0 = y, y1 = y, x0 = x, x1 = x;
  float ly0 = 1.f, ly1 = 0.f, lx0 = 1.f, lx1 = 0.f;
  if (resize) {
    bilinear_setup(y, a.h, a.ho, y0, y1, ly0, ly1);
    bilinear_setup(x, a.w, a.wo, x0, x1, lx0, lx1);
  }
  const int64_t plane = static_cast<int64_t>(a.h) * a.w;
  for (int cb = 0; cb < a.cd; cb += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cb + j;
      float t = 0.f;
      if (c < a.cs) {
        const float* s = a.src + (static_cast<int64_t>(n) * a.cs + c) * plane;
        if (resize) {
          t = ly0 * (lx0 * __ldg(s + y0 * a.w + x0) + lx1 * __ldg(s + y0 * a.w + x1)) +
              ly1 * (lx0 * __ldg(s + y1 * a.w + x0) + lx1 * __ldg(s + y1 * a.w + x1));
        } else {
          t = __ldg(s + y * a.w + x);
          if (a.mul_out) {
            const float o = __ldg(a.mul_out + (static_cast<int64_t>(n) * a.cs + c) * plane + y * a.w + x);
            t *= a.mul_kind == PCGAN_ACT_SIGMOID ? o * (1.f - o) : 1.f - o * o;
          }
        }
      } else if (c == a.cs && a.z != nullptr) {
        t = a.z[n];
      }
      v[j] = t;
    }
    store8(d + cb, v);
  }
}

__global__ void unpack_resize_bwd_kernel(pcgan_unpack_args a) {
  griddep_wait();
  griddep_launch();
  const int64_t total = static_cast<int64_t>(a.n) * a.h * a.w;
  const int hsp = a.hs + 2 * a.pad, wsp = a.ws + 2 * a.pad;
  const bool resize = (a.hs != a.h) || (a.ws != a.w);
  const float sy = a.hs > 1 ? static_cast<float>(a.h - 1) / static_cast<float>(a.hs - 1) : 0.f;
  const float sx = a.ws > 1 ? static_cast<float>(a.w - 1) / static_cast<float>(a.ws - 1) : 0.f;
  const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(a.g);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = static_cast<int>(i % a.w);
    const int y = static_cast<int>((i / a.w) % a.h);
    const int n = static_cast<int>(i / (static_cast<int64_t>(a.w) * a.h));
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    if (!resize) {
      const __nv_bfloat16* p = g + ((static_cast<int64_t>(n) * hsp + y + a.pad) * wsp + x + a.pad) * a.c;
      for (int c = 0; c < a.cd; ++c) acc[c] = __bfloat162float(p[c]);
    } else {
      // adjoint of the bilinear gather: visit every resized pixel whose footprint touches (y, x)
      int ylo = sy > 0.f ? static_cast<int>(floorf((y - 1) / sy)) : 0;
      int yhi = sy > 0.f ? static_cast<int>(ceilf((y + 1) / sy)) : a.hs - 1;
      int xlo = sx > 0.f ? static_cast<int>(floorf((x - 1) / sx)) : 0;
      int xhi = sx > 0.f ? static_cast<int>(ceilf((x + 1) / sx)) : a.ws - 1;
      ylo = max(ylo, 0); xlo = max(xlo, 0); yhi = min(yhi, a.hs - 1); xhi = min(xhi, a.ws - 1);
      for (int yy = ylo; yy <= yhi; ++yy) {
        int y0, y1; float l0, l1;
        bilinear_setup(yy, a.h, a.hs, y0, y1, l0, l1);
        const float wy = (y0 == y ? l0 : 0.f) + (y1 == y ? l1 : 0.f);
        if (wy == 0.f) continue;
        for (int xx = xlo; xx <= xhi; ++xx) {
          int x0, x1; float m0, m1;
          bilinear_setup(xx, a.w, a.ws, x0, x1, m0, m1);
          const float wx = (x0 == x ? m0 : 0.f) + (x1 == x ? m1 : 0.f);
          if (wx == 0.f) continue;
          const __nv_bfloat16* p = g + ((static_cast<int64_t>(n) * hsp + yy + a.pad) * wsp + xx + a.pad) * a.c;
          const float wgt = wy * wx;
          for (int c = 0; c < a.cd; ++c) acc[c] += wgt * __bfloat162float(p[c]);
        }
      }
    }
    for (int c = 0; c < a.cd; ++c) {
      float* d = a.dst + ((static_cast<int64_t>(n) * a.cd + c) * a.h + y) * a.w + x;
      const float v = acc[c] * a.scale;
      *d = a.accumulate ? *d + v : v;
    }
  }
}

// grid (x blocks, output rows, planes): the row coefficients are per block, no per-thread division
__global__ void resize_nchw_fwd_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t planes, int h, int w,
                                       int ho, int wo) {
  griddep_wait();
  griddep_launch();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= wo) return;
  int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
  bilinear_setup(y, h, ho, y0, y1, ly0, ly1);
  bilinear_setup(x, w, wo, x0, x1, lx0, lx1);
  for (int64_t pl = blockIdx.z; pl < planes; pl += gridDim.z) {
    const float* s = src + pl * h * w;
    dst[(pl * ho + y) * wo + x] = ly0 * (lx0 * __ldg(s + y0 * w + x0) + lx1 * __ldg(s + y0 * w + x1)) +
                                  ly1 * (lx0 * __ldg(s + y1 * w + x0) + lx1 * __ldg(s + y1 * w + x1));
  }
}

// adjoint: grid (x blocks, source rows, planes); every source pixel gathers the resized pixels whose footprint touches it
__global__ void resize_nchw_bwd_kernel(const float* __restrict__ gdst, float* __restrict__ gsrc, int64_t planes, int h, int w,
                                       int ho, int wo) {
  griddep_wait();
  griddep_launch();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= w) return;
  const float sy = ho > 1 ? static_cast<float>(h - 1) / static_cast<float>(ho - 1) : 0.f;
  const float sx = wo > 1 ? static_cast<float>(w - 1) / static_cast<float>(wo - 1) : 0.f;
  int ylo = sy > 0.f ? static_cast<int>(floorf((y - 1) / sy)) : 0;
  int yhi = sy > 0.f ? static_cast<int>(ceilf((y + 1) / sy)) : ho - 1;
  int xlo = sx > 0.f ? static_cast<int>(floorf((x - 1) / sx)) : 0;
  int xhi = sx > 0.f ? static_cast<int>(ceilf((x + 1) / sx)) : wo - 1;
  ylo = max(ylo, 0); xlo = max(xlo, 0); yhi = min(yhi, ho - 1); xhi = min(xhi, wo - 1);
  for (int64_t pl = blockIdx.z; pl < planes; pl += gridDim.z) {
    const float* g = gdst + pl * ho * wo;
    float acc = 0.f;
    for (int yy = ylo; yy <= yhi; ++yy) {
      int y0, y1; float l0, l1;
      bilinear_setup(yy, h, ho, y0, y1, l0, l1);
      const float wy = (y0 == y ? l0 : 0.f) + (y1 == y ? l1 : 0.f);
      if (wy == 0.f) continue;
      for (int xx = xlo; xx <= xhi; ++xx) {
        int x0, x1; float m0, m1;
        bilinear_setup(xx, w, wo, x0, x1, m0, m1);
        const float wx = (x0 == x ? m0 : 0.f) + (x1 == x ? m1 : 0.f);
        if (wx != 0.f) acc += wy * wx * __ldg(g + yy * wo + xx);
      }
    }
    gsrc[(pl * h + y) * w + x] = acc;
  }
}

// -------------------------------------------------------------------- maxpool
__global__ void maxpool_fwd_kernel(pcgan_maxpool_args a) {
  griddep_wait();
  griddep_launch();
  const int pp = a.pool_pad;
  const int ho = (a.h + 2 * pp - 3) / 2 + 1, wo = (a.w + 2 * pp - 3) / 2 + 1, cv = a.c >> 3;
  const int64_t total = static_cast<int64_t>(a.n) * ho * wo * cv;
  const __nv_bfloat16* xin = reinterpret_cast<const __nv_bfloat16*>(a.x);
  __nv_bfloat16* yout = reinterpret_cast<__nv_bfloat16*>(a.y);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = static_cast<int>(i % cv) << 3;
    int64_t q = i / cv;
    const int ox = static_cast<int>(q % wo); q /= wo;
    const int oy = static_cast<int>(q % ho);
    const int n = static_cast<int>(q / ho);
    float best[8];
    int bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 4; }
    for (int r = 0; r < 3; ++r) {
      const int y = 2 * oy - pp + r;
      if (y < 0 || y >= a.h) continue;
      for (int s = 0; s < 3; ++s) {
        const int x = 2 * ox - pp + s;
        if (x < 0 || x >= a.w) continue;
        float v[8];
        load8(xin + pix_off(n, y, x, a.h, a.w, a.c, a.x_pad) + c0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (v[j] > best[j]) { best[j] = v[j]; bi[j] = r * 3 + s; }
      }
    }
    store8(yout + pix_off(n, oy, ox, ho, wo, a.c, a.y_pad) + c0, best);
    uint8_t* ip = a.idx + ((static_cast<int64_t>(n) * ho + oy) * wo + ox) * a.c + c0;
    uint2 packed;
    packed.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
    packed.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
    *reinterpret_cast<uint2*>(ip) = packed;
  }
}

__global__ void maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int dy_pad, const uint8_t* __restrict__ idx,
                                   __nv_bfloat16* __restrict__ dx, int dx_pad, int n_, int h, int w, int c, int pp) {
  griddep_wait();
  griddep_launch();
  const int ho = (h + 2 * pp - 3) / 2 + 1, wo = (w + 2 * pp - 3) / 2 + 1, cv = c >> 3;
  const int64_t total = static_cast<int64_t>(n_) * h * w * cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = static_cast<int>(i % cv) << 3;
    int64_t q = i / cv;
    const int x = static_cast<int>(q % w); q /= w;
    const int y = static_cast<int>(q % h);
    const int n = static_cast<int>(q / h);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    // windows that contain (y, x): 2*o - pp <= coordinate <= 2*o - pp + 2
    for (int oy = max(y + pp - 2, 0) / 2; oy <= (y + pp) / 2; ++oy) {
      if (oy >= ho) continue;
      const int r = y - (2 * oy - pp);
      if (r < 0 || r > 2) continue;
      for (int ox = max(x + pp - 2, 0) / 2; ox <= (x + pp) / 2; ++ox) {
        if (ox >= wo) continue;
        const int s = x - (2 * ox - pp);
        if (s < 0 || s > 2) continue;
        const int pos = r * 3 + s;
        const uint2 packed = *reinterpret_cast<const uint2*>(idx + ((static_cast<int64_t>(n) * ho + oy) * wo + ox) * c + c0);
        float g[8];
        load8(dy + pix_off(n, oy, ox, ho, wo, c, dy_pad) + c0, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t word = j < 4 ? packed.x : packed.y;
          if (static_cast<int>((word >> (8 * (j & 3))) & 0xff) == pos) acc[j] += g[j];
        }
      }
    }
    store8(dx + pix_off(n, y, x, h, w, c, dx_pad) + c0, acc);
  }
}

// ---------------------------------------------------- activation backward / cast
// dx = dy where the stored post-activation output y is positive, else slope * dy (nn.ReLU / nn.LeakyReLU backward from
// the sign of the output), any channel count that is a multiple of 8; the three buffers carry their own halo widths.
__global__ void act_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int dy_pad, const __nv_bfloat16* __restrict__ y, int y_pad,
                               __nv_bfloat16* __restrict__ dx, int dx_pad, int n_, int h, int w, int c, float slope) {
  griddep_wait();
  griddep_launch();
  const int cv = c >> 3;
  const int64_t total = static_cast<int64_t>(n_) * h * w * cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = static_cast<int>(i % cv) << 3;
    int64_t q = i / cv;
    const int x = static_cast<int>(q % w); q /= w;
    const int yy = static_cast<int>(q % h);
    const int n = static_cast<int>(q / h);
    float g[8], o[8];
    load8(dy + pix_off(n, yy, x, h, w, c, dy_pad) + c0, g);
    load8(y + pix_off(n, yy, x, h, w, c, y_pad) + c0, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = o[j] > 0.f ? g[j] : g[j] * slope;
    store8(dx + pix_off(n, yy, x, h, w, c, dx_pad) + c0, g);
  }
}

// NHWC bf16 (padded) <-> contiguous NHWC fp32: the feature map a frozen network hands to a torch loss, and its gradient
__global__ void nhwc_cast_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int pad, int n_, int h, int w, int c) {
  griddep_wait();
  griddep_launch();
  const int cv = c >> 3;
  const int64_t total = static_cast<int64_t>(n_) * h * w * cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = static_cast<int>(i % cv) << 3;
    int64_t q = i / cv;
    const int x = static_cast<int>(q % w); q /= w;
    const int yy = static_cast<int>(q % h);
    const int n = static_cast<int>(q / h);
    float v[8];
    load8(src + pix_off(n, yy, x, h, w, c, pad) + c0, v);
    float4* d = reinterpret_cast<float4*>(dst + i * 8);
    d[0] = make_float4(v[0], v[1], v[2], v[3]);
    d[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}
__global__ void nhwc_uncast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int pad, int n_, int h, int w, int c) {
  griddep_wait();
  griddep_launch();
  const int cv = c >> 3;
  const int64_t total = static_cast<int64_t>(n_) * h * w * cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = static_cast<int>(i % cv) << 3;
    int64_t q = i / cv;
    const int x = static_cast<int>(q % w); q /= w;
    const int yy = static_cast<int>(q % h);
    const int n = static_cast<int>(q / h);
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + i * 8)), b = __ldg(reinterpret_cast<const float4*>(src + i * 8) + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    store8(dst + pix_off(n, yy, x, h, w, c, pad) + c0, v);
  }
}

// ------------------------------------------------------------- input pipeline
// Resize (PIL's antialiased bicubic: two separable passes with 8-bit intermediate rounding) + crop + horizontal flip +
// ToTensor + Normalize(0.5, 0.5) of one decoded RGB image per blockIdx.y, straight into the fp32 NCHW batch tensor
// (data/base_dataset.py:24-64).  Thread = one output pixel, three channels.
__device__ __forceinline__ float bicubic_w(float x) {
  const float a = -0.5f;
  x = fabsf(x);
  if (x < 1.f) return ((a + 2.f) * x - (a + 3.f)) * x * x + 1.f;
  if (x < 2.f) return (((x - 5.f) * x + 8.f) * x - 4.f) * a;
  return 0.f;
}
// the source range [lo, lo + cnt) and filter scale of output coordinate r when `in` samples are resized to `out`
__device__ __forceinline__ void resample_window(int r, int in, int out, int& lo, int& cnt, float& center, float& ss) {
  const float scale = static_cast<float>(in) / static_cast<float>(out);
  const float fs = scale > 1.f ? scale : 1.f;
  const float support = 2.f * fs;
  center = (r + 0.5f) * scale;
  ss = 1.f / fs;
  lo = static_cast<int>(center - support + 0.5f);
  if (lo < 0) lo = 0;
  int hi = static_cast<int>(center + support + 0.5f);
  if (hi > in) hi = in;
  cnt = hi - lo;
}
__device__ __forceinline__ float clip8(float v) {
  v = floorf(v + 0.5f);
  return v < 0.f ? 0.f : (v > 255.f ? 255.f : v);
}

__global__ void augment_kernel(pcgan_augment_args a) {
  griddep_wait();
  griddep_launch();
  const pcgan_image_item it = a.items[blockIdx.y];
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= a.fine * a.fine) return;
  const int oy = pix / a.fine, ox = pix - oy * a.fine;
  const int ry = oy + it.crop_y;
  const int rx = (it.flip ? a.fine - 1 - ox : ox) + it.crop_x;
  int x0, xn, y0, yn;
  float cx, sx, cy, sy;
  resample_window(rx, it.w, a.load, x0, xn, cx, sx);
  resample_window(ry, it.h, a.load, y0, yn, cy, sy);
  float wxs = 0.f, wys = 0.f;
  for (int i = 0; i < xn; ++i) wxs += bicubic_w((i + x0 - cx + 0.5f) * sx);
  for (int j = 0; j < yn; ++j) wys += bicubic_w((j + y0 - cy + 0.5f) * sy);
  float acc[3] = {0.f, 0.f, 0.f};
  for (int j = 0; j < yn; ++j) {
    const uint8_t* row = it.src + (static_cast<int64_t>(y0 + j) * it.w + x0) * 3;
    float h[3] = {0.f, 0.f, 0.f};
    for (int i = 0; i < xn; ++i) {
      const float wgt = bicubic_w((i + x0 - cx + 0.5f) * sx);
      h[0] = fmaf(wgt, static_cast<float>(row[3 * i]), h[0]);
      h[1] = fmaf(wgt, static_cast<float>(row[3 * i + 1]), h[1]);
      h[2] = fmaf(wgt, static_cast<float>(row[3 * i + 2]), h[2]);
    }
    const float wy = bicubic_w((j + y0 - cy + 0.5f) * sy) / wys;
    // the horizontal pass is stored as 8-bit pixels before the vertical pass runs (PIL's ImagingResample)
    acc[0] = fmaf(wy, clip8(h[0] / wxs), acc[0]);
    acc[1] = fmaf(wy, clip8(h[1] / wxs), acc[1]);
    acc[2] = fmaf(wy, clip8(h[2] / wxs), acc[2]);
  }
  float* dst = a.dst + static_cast<int64_t>(blockIdx.y) * 3 * a.fine * a.fine + pix;
#pragma unroll
  for (int c = 0; c < 3; ++c) dst[static_cast<int64_t>(c) * a.fine * a.fine] = (clip8(acc[c]) / 255.f - 0.5f) / 0.5f;
}

// ----------------------------------------------------------------------- loss
__global__ void loss_kernel(pcgan_loss_args a) {
  griddep_wait();
  griddep_launch();
  __shared__ float red[kThreads / 32];
  float acc = 0.f;
  const float invn = 1.f / static_cast<float>(a.n);
  const float wgt = a.weight * (a.weight_dev ? *a.weight_dev : 1.f);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    const float p = a.p[i];
    const float t = a.per_sample > 0 ? a.target[i / a.per_sample] : a.target[i];
    float l, g;
    switch (a.kind) {
      case PCGAN_LOSS_BCE: {
        // nn.BCELoss: log terms clamped at -100; grad (p - t) / max(p (1 - p), 1e-12)
        const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(logf(1.f - p), -100.f);
        l = -(t * lp + (1.f - t) * lq);
        g = (p - t) / fmaxf(p * (1.f - p), 1e-12f);
        break;
      }
      case PCGAN_LOSS_MSE: l = (p - t) * (p - t); g = 2.f * (p - t); break;
      case PCGAN_LOSS_L1: l = fabsf(p - t); g = p > t ? 1.f : (p < t ? -1.f : 0.f); break;
      case PCGAN_LOSS_ELO_NLL_SCORE: {  // p is the rating difference: prob = sigmoid(p) (siamese.py:675), then as below
        const float e = 1e-20f;
        const float pr = 1.f / (1.f + expf(-p));
        l = -(t * logf(pr + e) + (1.f - t) * logf(1.f - pr + e));
        g = -(t / (pr + e) - (1.f - t) / (1.f - pr + e)) * pr * (1.f - pr);
        break;
      }
      default: {  // PCGAN_LOSS_ELO_NLL, networks.py:479-481 with MAGIC_EPS = 1e-20
        const float e = 1e-20f;
        l = -(t * logf(p + e) + (1.f - t) * logf(1.f - p + e));
        g = -(t / (p + e) - (1.f - t) / (1.f - p + e));
        break;
      }
    }
    acc += l;
    if (a.grad) a.grad[i] = wgt * g * invn;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kThreads / 32 ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0 && a.loss) atomicAdd(a.loss, wgt * v * invn);
  }
}

// ----------------------------------------------------------------------- adam
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, const float* lr_p, float b1, float b2, float eps,
                            const float* step_p) {
  griddep_wait();
  griddep_launch();
  const float lr = *lr_p, step = *step_p;
  const float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
  const float step_size = lr / bc1, rsq_bc2 = rsqrtf(bc2);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * rsq_bc2 + eps);
  }
}

// Multi-tensor form: blockIdx.y = tensor.  `step` holds the number of updates already applied; this update is number
// step + 1 and step_inc_kernel, launched behind it, stores that.
__global__ void adam_batched_kernel(const pcgan_adam_item* __restrict__ items, const float* lr_p, float b1, float b2, float c1,
                                    float c2, float eps, const float* step_p) {
  griddep_wait();
  griddep_launch();
  const pcgan_adam_item it = items[blockIdx.y];
  const float lr = *lr_p, step = *step_p + 1.f;
  const float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
  const float step_size = lr / bc1, rsq_bc2 = rsqrtf(bc2);
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(it.p) | reinterpret_cast<uintptr_t>(it.g) | reinterpret_cast<uintptr_t>(it.m) |
                     reinterpret_cast<uintptr_t>(it.v)) & 15) == 0;
  int64_t done = 0;
  if (vec) {
    const int64_t n4 = it.n >> 2;
    float4* p4 = reinterpret_cast<float4*>(it.p);
    const float4* g4 = reinterpret_cast<const float4*>(it.g);
    float4* m4 = reinterpret_cast<float4*>(it.m);
    float4* v4 = reinterpret_cast<float4*>(it.v);
    for (int64_t i = tid; i < n4; i += nthr) {
      const float4 g = g4[i];
      float4 m = m4[i], v = v4[i], p = p4[i];
      m.x = b1 * m.x + c1 * g.x; m.y = b1 * m.y + c1 * g.y; m.z = b1 * m.z + c1 * g.z; m.w = b1 * m.w + c1 * g.w;
      v.x = b2 * v.x + c2 * g.x * g.x; v.y = b2 * v.y + c2 * g.y * g.y; v.z = b2 * v.z + c2 * g.z * g.z; v.w = b2 * v.w + c2 * g.w * g.w;
      p.x -= step_size * m.x / (sqrtf(v.x) * rsq_bc2 + eps);
      p.y -= step_size * m.y / (sqrtf(v.y) * rsq_bc2 + eps);
      p.z -= step_size * m.z / (sqrtf(v.z) * rsq_bc2 + eps);
      p.w -= step_size * m.w / (sqrtf(v.w) * rsq_bc2 + eps);
      m4[i] = m; v4[i] = v; p4[i] = p;
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < it.n; i += nthr) {
    const float g = it.g[i];
    const float m = b1 * it.m[i] + c1 * g;
    const float v = b2 * it.v[i] + c2 * g * g;
    it.m[i] = m;
    it.v[i] = v;
    it.p[i] -= step_size * m / (sqrtf(v) * rsq_bc2 + eps);
  }
}

__global__ void step_inc_kernel(float* step) {
  griddep_wait();   // the update that reads the old value has completed
  griddep_launch();
  *step += 1.f;
}

}  // namespace pcgan

using namespace pcgan;
#define STREAM(s) static_cast<cudaStream_t>(s)

extern "C" int pcgan_gather_cast_bf16(const float* src, const int32_t* idx, void* dst, int64_t n, pcgan_stream_t s) {
  if (!src || !idx || !dst || n < 0) return fail(PCGAN_ERR_INVALID, "gather_cast: bad argument");
  if (n == 0) return PCGAN_OK;
  PCGAN_CUDA_OK(launch_pdl(gather_cast_kernel, dim3(grid_for(n)), dim3(kThreads), 0, STREAM(s), 1, src, idx, reinterpret_cast<__nv_bfloat16*>(dst), n));
  PCGAN_LAUNCH_OK("gather_cast_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_gather_tf32(const float* src, const int32_t* idx, float* dst, int64_t n, pcgan_stream_t s) {
  if (!src || !idx || !dst || n < 0) return fail(PCGAN_ERR_INVALID, "gather_tf32: bad argument");
  if (n == 0) return PCGAN_OK;
  PCGAN_CUDA_OK(launch_pdl(gather_tf32_kernel, dim3(grid_for(n)), dim3(kThreads), 0, STREAM(s), 1, src, idx, dst, n));
  PCGAN_LAUNCH_OK("gather_tf32_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_scatter_f32(const float* src, const int32_t* idx, float* dst, int64_t n, int32_t accumulate,
                                 pcgan_stream_t s) {
  if (!src || !idx || !dst || n < 0) return fail(PCGAN_ERR_INVALID, "scatter: bad argument");
  if (n == 0) return PCGAN_OK;
  PCGAN_CUDA_OK(launch_pdl(scatter_kernel, dim3(grid_for(n)), dim3(kThreads), 0, STREAM(s), 1, src, idx, dst, n, accumulate));
  PCGAN_LAUNCH_OK("scatter_kernel");
  return PCGAN_OK;
}

static int batched_grid(int32_t count, int64_t max_n, dim3* grid) {
  if (count < 1 || count > 65535 || max_n < 1) return fail(PCGAN_ERR_INVALID, "batched gather/scatter: count=%d max_n=%lld", count, (long long)max_n);
  const int sms = sm_count() > 0 ? sm_count() : 148;
  int64_t bx = (max_n + kThreads * 4 - 1) / (kThreads * 4);      // about four elements per thread for the largest tensor
  const int64_t cap = (static_cast<int64_t>(sms) * 16 + count - 1) / count;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  *grid = dim3(static_cast<unsigned>(bx), static_cast<unsigned>(count));
  return PCGAN_OK;
}

extern "C" int pcgan_gather_cast_bf16_batched(const pcgan_batch_item* items, int32_t count, int64_t max_n, pcgan_stream_t s) {
  if (!items) return fail(PCGAN_ERR_INVALID, "gather_cast_batched: null table");
  dim3 grid;
  int rc = batched_grid(count, max_n, &grid);
  if (rc) return rc;
  PCGAN_CUDA_OK(launch_pdl(gather_cast_batched_kernel, grid, dim3(kThreads), 0, STREAM(s), 1, items));
  PCGAN_LAUNCH_OK("gather_cast_batched_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_scatter_f32_batched(const pcgan_batch_item* items, int32_t count, int64_t max_n, int32_t accumulate,
                                         pcgan_stream_t s) {
  if (!items) return fail(PCGAN_ERR_INVALID, "scatter_batched: null table");
  dim3 grid;
  int rc = batched_grid(count, max_n, &grid);
  if (rc) return rc;
  PCGAN_CUDA_OK(launch_pdl(scatter_batched_kernel, grid, dim3(kThreads), 0, STREAM(s), 1, items, static_cast<int>(accumulate)));
  PCGAN_LAUNCH_OK("scatter_batched_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_pack_nchw(const pcgan_pack_args* a, pcgan_stream_t s) {
  if (!a || !a->src || !a->dst) return fail(PCGAN_ERR_INVALID, "pack_nchw: null argument");
  if (a->cd % 8 != 0 || a->cd < a->cs + (a->z ? 1 : 0)) return fail(PCGAN_ERR_INVALID, "pack_nchw: cd=%d must be a multiple of 8 holding %d channels", a->cd, a->cs + (a->z ? 1 : 0));
  if (a->n < 1 || a->h < 1 || a->w < 1 || a->ho < 1 || a->wo < 1 || a->pad < 0) return fail(PCGAN_ERR_INVALID, "pack_nchw: bad geometry");
  if (a->halo == PCGAN_HALO_REFLECT && (a->pad >= a->ho || a->pad >= a->wo)) return fail(PCGAN_ERR_INVALID, "pack_nchw: reflect pad too large");
  if (a->mul_out && (a->ho != a->h || a->wo != a->w)) return fail(PCGAN_ERR_UNSUPPORTED, "pack_nchw: mul_out with resize");
  if (a->n > 65535 || a->ho + 2 * a->pad > 65535) return fail(PCGAN_ERR_UNSUPPORTED, "pack_nchw: n / rows out of range");
  const int wp = a->wo + 2 * a->pad;
  const int bx = wp >= 128 ? 128 : 64;
  PCGAN_CUDA_OK(launch_pdl(pack_nchw_kernel, dim3((wp + bx - 1) / bx, a->ho + 2 * a->pad, a->n), dim3(bx), 0, STREAM(s), 1, *a));
  PCGAN_LAUNCH_OK("pack_nchw_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_resize_nchw_fwd(const float* src, float* dst, int64_t planes, int32_t h, int32_t w, int32_t ho, int32_t wo,
                                     pcgan_stream_t s) {
  if (!src || !dst || planes < 1 || h < 1 || w < 1 || ho < 1 || wo < 1) return fail(PCGAN_ERR_INVALID, "resize_fwd: bad argument");
  if (ho > 65535) return fail(PCGAN_ERR_UNSUPPORTED, "resize_fwd: ho out of range");
  PCGAN_CUDA_OK(launch_pdl(resize_nchw_fwd_kernel, dim3((wo + 127) / 128, ho, static_cast<unsigned>(planes < 64 ? planes : 64)), dim3(128), 0, STREAM(s), 1,
                           src, dst, planes, h, w, ho, wo));
  PCGAN_LAUNCH_OK("resize_nchw_fwd_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_resize_nchw_bwd(const float* gdst, float* gsrc, int64_t planes, int32_t h, int32_t w, int32_t ho, int32_t wo,
                                     pcgan_stream_t s) {
  if (!gdst || !gsrc || planes < 1 || h < 1 || w < 1 || ho < 1 || wo < 1) return fail(PCGAN_ERR_INVALID, "resize_bwd: bad argument");
  if (h > 65535) return fail(PCGAN_ERR_UNSUPPORTED, "resize_bwd: h out of range");
  PCGAN_CUDA_OK(launch_pdl(resize_nchw_bwd_kernel, dim3((w + 127) / 128, h, static_cast<unsigned>(planes < 64 ? planes : 64)), dim3(128), 0, STREAM(s), 1,
                           gdst, gsrc, planes, h, w, ho, wo));
  PCGAN_LAUNCH_OK("resize_nchw_bwd_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_unpack_resize_bwd(const pcgan_unpack_args* a, pcgan_stream_t s) {
  if (!a || !a->g || !a->dst) return fail(PCGAN_ERR_INVALID, "unpack: null argument");
  if (a->cd < 1 || a->cd > 8 || a->cd > a->c) return fail(PCGAN_ERR_UNSUPPORTED, "unpack: cd=%d (1..8, <= c)", a->cd);
  const int64_t total = static_cast<int64_t>(a->n) * a->h * a->w;
  PCGAN_CUDA_OK(launch_pdl(unpack_resize_bwd_kernel, dim3(grid_for(total)), dim3(kThreads), 0, STREAM(s), 1, *a));
  PCGAN_LAUNCH_OK("unpack_resize_bwd_kernel");
  return PCGAN_OK;
}

static int check_cvec(int c, const char* who) {
  if (c < 8 || c % 8 != 0) return fail(PCGAN_ERR_UNSUPPORTED, "%s: channels=%d must be a multiple of 8", who, c);
  return PCGAN_OK;
}

extern "C" int pcgan_maxpool3x3s2_fwd(const pcgan_maxpool_args* a, pcgan_stream_t s) {
  if (!a || !a->x || !a->y || !a->idx) return fail(PCGAN_ERR_INVALID, "maxpool: null argument");
  int rc = check_cvec(a->c, "maxpool");
  if (rc) return rc;
  if (a->pool_pad != 0 && a->pool_pad != 1) return fail(PCGAN_ERR_INVALID, "maxpool: pool_pad must be 0 or 1");
  if (a->h + 2 * a->pool_pad < 3 || a->w + 2 * a->pool_pad < 3) return fail(PCGAN_ERR_INVALID, "maxpool: image smaller than the window");
  const int64_t total = static_cast<int64_t>(a->n) * ((a->h + 2 * a->pool_pad - 3) / 2 + 1) * ((a->w + 2 * a->pool_pad - 3) / 2 + 1) * (a->c / 8);
  PCGAN_CUDA_OK(launch_pdl(maxpool_fwd_kernel, dim3(grid_for(total)), dim3(kThreads), 0, STREAM(s), 1, *a));
  PCGAN_LAUNCH_OK("maxpool_fwd_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_maxpool3x3s2_bwd(const void* dy, int32_t dy_pad, const uint8_t* idx, void* dx, int32_t dx_pad,
                                      int32_t n, int32_t h, int32_t w, int32_t c, int32_t pool_pad, pcgan_stream_t s) {
  if (!dy || !idx || !dx) return fail(PCGAN_ERR_INVALID, "maxpool_bwd: null argument");
  if (pool_pad != 0 && pool_pad != 1) return fail(PCGAN_ERR_INVALID, "maxpool_bwd: pool_pad must be 0 or 1");
  int rc = check_cvec(c, "maxpool_bwd");
  if (rc) return rc;
  const int64_t total = static_cast<int64_t>(n) * h * w * (c / 8);
  PCGAN_CUDA_OK(launch_pdl(maxpool_bwd_kernel, dim3(grid_for(total)), dim3(kThreads), 0, STREAM(s), 1, reinterpret_cast<const __nv_bfloat16*>(dy), dy_pad, idx,
                                                                 reinterpret_cast<__nv_bfloat16*>(dx), dx_pad, n, h, w, c, pool_pad));
  PCGAN_LAUNCH_OK("maxpool_bwd_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_act_bwd(const void* dy, int32_t dy_pad, const void* y, int32_t y_pad, void* dx, int32_t dx_pad, int32_t n, int32_t h,
                             int32_t w, int32_t c, float slope, pcgan_stream_t s) {
  if (!dy || !y || !dx || n < 1 || h < 1 || w < 1) return fail(PCGAN_ERR_INVALID, "act_bwd: bad argument");
  int rc = check_cvec(c, "act_bwd");
  if (rc) return rc;
  const int64_t total = static_cast<int64_t>(n) * h * w * (c / 8);
  PCGAN_CUDA_OK(launch_pdl(act_bwd_kernel, dim3(grid_for(total)), dim3(kThreads), 0, STREAM(s), 1, reinterpret_cast<const __nv_bfloat16*>(dy), dy_pad,
                           reinterpret_cast<const __nv_bfloat16*>(y), y_pad, reinterpret_cast<__nv_bfloat16*>(dx), dx_pad, n, h, w, c, slope));
  PCGAN_LAUNCH_OK("act_bwd_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_nhwc_cast(const void* src, void* dst, int32_t pad, int32_t n, int32_t h, int32_t w, int32_t c, int32_t to_f32, pcgan_stream_t s) {
  if (!src || !dst || n < 1 || h < 1 || w < 1) return fail(PCGAN_ERR_INVALID, "nhwc_cast: bad argument");
  int rc = check_cvec(c, "nhwc_cast");
  if (rc) return rc;
  const int64_t total = static_cast<int64_t>(n) * h * w * (c / 8);
  if (to_f32) PCGAN_CUDA_OK(launch_pdl(nhwc_cast_kernel, dim3(grid_for(total)), dim3(kThreads), 0, STREAM(s), 1, reinterpret_cast<const __nv_bfloat16*>(src),
                                       reinterpret_cast<float*>(dst), pad, n, h, w, c));
  else PCGAN_CUDA_OK(launch_pdl(nhwc_uncast_kernel, dim3(grid_for(total)), dim3(kThreads), 0, STREAM(s), 1, reinterpret_cast<const float*>(src),
                                reinterpret_cast<__nv_bfloat16*>(dst), pad, n, h, w, c));
  PCGAN_LAUNCH_OK("nhwc_cast_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_augment(const pcgan_augment_args* a, pcgan_stream_t s) {
  if (!a || !a->items || !a->dst || a->n < 1 || a->n > 65535 || a->load < 1 || a->fine < 1 || a->fine > a->load)
    return fail(PCGAN_ERR_INVALID, "augment: bad argument");
  const int px = a->fine * a->fine;
  PCGAN_CUDA_OK(launch_pdl(augment_kernel, dim3((px + kThreads - 1) / kThreads, a->n), dim3(kThreads), 0, STREAM(s), 1, *a));
  PCGAN_LAUNCH_OK("augment_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_loss(const pcgan_loss_args* a, pcgan_stream_t s) {
  if (!a || !a->p || !a->target || a->n < 1) return fail(PCGAN_ERR_INVALID, "loss: bad argument");
  if (a->kind < PCGAN_LOSS_BCE || a->kind > PCGAN_LOSS_ELO_NLL_SCORE) return fail(PCGAN_ERR_INVALID, "loss: bad kind");
  PCGAN_CUDA_OK(launch_pdl(loss_kernel, dim3(grid_for(a->n, kThreads, 2)), dim3(kThreads), 0, STREAM(s), 1, *a));
  PCGAN_LAUNCH_OK("loss_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_adam(float* p, const float* g, float* m, float* v, int64_t n, const float* lr, float beta1,
                          float beta2, float eps, const float* step, pcgan_stream_t s) {
  if (!p || !g || !m || !v || !lr || !step || n < 0) return fail(PCGAN_ERR_INVALID, "adam: bad argument");
  if (n == 0) return PCGAN_OK;
  PCGAN_CUDA_OK(launch_pdl(adam_kernel, dim3(grid_for(n)), dim3(kThreads), 0, STREAM(s), 1, p, g, m, v, n, lr, beta1, beta2, eps, step));
  PCGAN_LAUNCH_OK("adam_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_adam_batched(const pcgan_adam_item* items, int32_t count, int64_t max_n, const float* lr, double beta1,
                                  double beta2, double eps, float* step, pcgan_stream_t s) {
  if (!items || !lr || !step) return fail(PCGAN_ERR_INVALID, "adam_batched: null argument");
  dim3 grid;
  int rc = batched_grid(count, max_n, &grid);
  if (rc) return rc;
  // 1 - beta is rounded from double, as torch.optim.Adam's `value=1 - beta2` is
  PCGAN_CUDA_OK(launch_pdl(adam_batched_kernel, grid, dim3(kThreads), 0, STREAM(s), 1, items, lr, static_cast<float>(beta1),
                           static_cast<float>(beta2), static_cast<float>(1.0 - beta1), static_cast<float>(1.0 - beta2),
                           static_cast<float>(eps), static_cast<const float*>(step)));
  PCGAN_LAUNCH_OK("adam_batched_kernel");
  PCGAN_CUDA_OK(launch_pdl(step_inc_kernel, dim3(1), dim3(1), 0, STREAM(s), 1, step));
  PCGAN_LAUNCH_OK("step_inc_kernel");
  return PCGAN_OK;
}
