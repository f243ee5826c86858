// Normalisation kernels of the wsgan_emb step (HBM-bound): statistics finalize, normalise + activation
// (+ residual) with halo write, halo fold, and the two-pass backward.
//
// Layout: activations are NHWC bf16 in physically padded buffers.  Every kernel runs on a 2-D grid
// (chunks of one sample, sample): a block owns a contiguous run of 16-byte vectors (8 channels) of ONE sample,
// so all index arithmetic is 32-bit (one division per vector), the per-channel constants of a thread are loaded
// once (256 % (C/8) == 0: a thread always sees the same 8 channels), and each thread keeps kU independent
// 16-byte loads in flight.
#include "common.cuh"

namespace pcgan {

static constexpr int kT = 256;   // threads per block

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
  const int cl = threadIdx.x & 31, gl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float macc = 0.f, vacc = 0.f;
  if (c < a.c) {
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
  sm[gl][cl] = macc;
  sv[gl][cl] = vacc;
  __syncthreads();
  if (gl == 0 && c < a.c && (a.running_mean || a.running_var)) {
    float m = 0.f, v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { m += sm[i][cl]; v += sv[i][cl]; }
    if (a.running_mean) a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * (m / a.groups);
    if (a.running_var) {
      const float unbias = a.count > 1.f ? a.count / (a.count - 1.f) : 1.f;
      a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (v / a.groups * unbias);
    }
  }
}

// ----------------------------------------------------------------- norm apply
// y = act(sc*x + sh [+ rsc*res + rsh]) over the whole padded grid of y (interior + halo).
template <bool RES>
__global__ void __launch_bounds__(kT, 3) norm_apply_kernel(pcgan_norm_apply_args a, int rows_per_block, int lcv) {
  constexpr int kU = RES ? 2 : 4;
  const int n = blockIdx.y;
  const int cv = a.c >> 3;
  const int hp = a.h + 2 * a.y_pad, wp = a.w + 2 * a.y_pad;
  const int rowvec = wp * cv;
  const int row0 = blockIdx.x * rows_per_block;
  const int nrows = min(rows_per_block, hp - row0);
  const int total = nrows * rowvec;
  const int c0 = (threadIdx.x & (cv - 1)) << 3;
  float sc[8], sh[8], rsc[RES ? 8 : 1], rsh[RES ? 8 : 1];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = 1.f; sh[j] = 0.f; }
#pragma unroll
  for (int j = 0; j < (RES ? 8 : 1); ++j) { rsc[j] = 1.f; rsh[j] = 0.f; }
  if (a.scale) {
    const int64_t so = static_cast<int64_t>(a.groups > 1 ? n : 0) * a.c + c0;
    load_f8(a.scale + so, sc);
    load_f8(a.shift + so, sh);
  }
  if (a.drop_mask) {
    float m[8];
    load_f8(a.drop_mask + static_cast<int64_t>(n) * a.c + c0, m);
#pragma unroll
    for (int j = 0; j < 8; ++j) sc[j] *= m[j];
  }
  constexpr bool has_res = RES;
  if constexpr (RES) {
    if (a.res_scale) {
      const int64_t ro = static_cast<int64_t>(a.res_groups > 1 ? n : 0) * a.c + c0;
      load_f8(a.res_scale + ro, rsc);
      load_f8(a.res_shift + ro, rsh);
    }
  }
  const int wxp = a.w + 2 * a.x_pad, wrp = a.w + 2 * a.res_pad;
  const __nv_bfloat16* xs = reinterpret_cast<const __nv_bfloat16*>(a.x) +
                            static_cast<int64_t>(n) * (a.h + 2 * a.x_pad) * wxp * a.c + c0;
  const __nv_bfloat16* rs = has_res ? reinterpret_cast<const __nv_bfloat16*>(a.res) +
                                          static_cast<int64_t>(n) * (a.h + 2 * a.res_pad) * wrp * a.c + c0
                                    : nullptr;
  __nv_bfloat16* ys = reinterpret_cast<__nv_bfloat16*>(a.y) + (static_cast<int64_t>(n) * hp + row0) * wp * a.c;
  const bool zero_halo = a.y_halo == PCGAN_HALO_ZERO;

  for (int v0 = threadIdx.x; v0 < total; v0 += kT * kU) {
    uint4 xv[kU], rv[RES ? kU : 1];
    int st[kU];   // 0: out of range, 1: zero halo, 2: value
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int v = v0 + u * kT;
      st[u] = 0;
      if (v < total) {
        const int r = v / rowvec;
        const int px = (v - r * rowvec) >> lcv;
        int y = row0 + r - a.y_pad, x = px - a.y_pad;
        const bool halo = y < 0 || y >= a.h || x < 0 || x >= a.w;
        if (halo && zero_halo) {
          st[u] = 1;
        } else {
          st[u] = 2;
          y = reflect_idx(y, a.h);
          x = reflect_idx(x, a.w);
          xv[u] = ldg16(xs + ((y + a.x_pad) * wxp + x + a.x_pad) * a.c);
          if constexpr (RES) rv[u] = ldg16(rs + ((y + a.res_pad) * wrp + x + a.res_pad) * a.c);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (st[u] == 0) continue;
      __nv_bfloat16* d = ys + static_cast<int64_t>(v0 + u * kT) * 8;
      if (st[u] == 1) {
        *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
        continue;
      }
      float x8[8], o[8];
      unpack8(xv[u], x8);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(sc[j], x8[j], sh[j]);
      if constexpr (RES) {
        float r8[8];
        unpack8(rv[u], r8);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += fmaf(rsc[j], r8[j], rsh[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = apply_act(o[j], a.act, a.act_slope);
      store8(d, o);
    }
  }
}

// ------------------------------------------------------------------ halo fold
// 16-byte vector of the gradient at interior pixel (y, x) of a padded-grid gradient, halo folded onto its mirror
// (a pixel within p of a border also receives the halo rows / columns that ReflectionPad2d copied from it)
__device__ __forceinline__ void folded_load(const __nv_bfloat16* g, int y, int x, int h, int w, int p, int c, bool reflect,
                                            float (&acc)[8]) {
  const int wp = w + 2 * p;
  const int ya = y + p, xa = x + p;
  unpack8(ldg16(g + (ya * wp + xa) * c), acc);
  if (!reflect) return;
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

// -------------------------------------------------------------- norm backward
// Variants (compile time): GEN = false is the lean path of the generator (InstanceNorm without affine, no dropout mask,
// residual not needed for the activation mask): scale == rstd and shift == -mean*rstd, so the normalised value xhat IS
// the pre-activation sc*x + sh.  GEN = true carries mean / rstd / mask separately (BatchNorm with gamma, beta);
// RES adds the residual branch to the pre-activation (ResNet BasicBlock: relu(bn(x) + shortcut)).
template <bool GEN, bool RES>
struct BwdCtx {
  float sc[8], sh[8];                               // pre = sc*(x*mask) + sh (+ residual)
  float mean[GEN ? 8 : 1], rstd[GEN ? 8 : 1], mk[GEN ? 8 : 1];
  float rsc[RES ? 8 : 1], rsh[RES ? 8 : 1];
  const __nv_bfloat16* dy; const __nv_bfloat16* x; const __nv_bfloat16* res;
  int wdp, wxp, wrp;
  int fold;                                         // 0 plain, 2 reflect fold

  __device__ __forceinline__ void init(const pcgan_norm_bwd_args& a, int n, int c0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = 1.f; sh[j] = 0.f; }
#pragma unroll
    for (int j = 0; j < (GEN ? 8 : 1); ++j) { mean[j] = 0.f; rstd[j] = 0.f; mk[j] = 1.f; }
#pragma unroll
    for (int j = 0; j < (RES ? 8 : 1); ++j) { rsc[j] = 1.f; rsh[j] = 0.f; }
    const int64_t so = static_cast<int64_t>(a.groups > 1 ? n : 0) * a.c + c0;
    if (a.scale) { load_f8(a.scale + so, sc); load_f8(a.shift + so, sh); }
    if constexpr (GEN) {
      if (a.mean) { load_f8(a.mean + so, mean); load_f8(a.rstd + so, rstd); }
      if (a.drop_mask) load_f8(a.drop_mask + static_cast<int64_t>(n) * a.c + c0, mk);
    }
    wdp = a.w + 2 * a.dy_pad; wxp = a.w + 2 * a.x_pad; wrp = a.w + 2 * a.res_pad;
    dy = reinterpret_cast<const __nv_bfloat16*>(a.dy) + static_cast<int64_t>(n) * (a.h + 2 * a.dy_pad) * wdp * a.c + c0;
    x = reinterpret_cast<const __nv_bfloat16*>(a.x) + static_cast<int64_t>(n) * (a.h + 2 * a.x_pad) * wxp * a.c + c0;
    res = nullptr;
    if constexpr (RES) {
      res = reinterpret_cast<const __nv_bfloat16*>(a.res) + static_cast<int64_t>(n) * (a.h + 2 * a.res_pad) * wrp * a.c + c0;
      if (a.res_scale) {
        const int64_t ro = static_cast<int64_t>(a.res_groups > 1 ? n : 0) * a.c + c0;
        load_f8(a.res_scale + ro, rsc);
        load_f8(a.res_shift + ro, rsh);
      }
    }
    fold = a.dy_fold;
  }

  __device__ __forceinline__ void load(const pcgan_norm_bwd_args& a, int y, int xx, float (&dyv)[8], uint4& xraw, uint4& rraw) const {
    if (fold) folded_load(dy, y, xx, a.h, a.w, a.dy_pad, a.c, fold == 2, dyv);
    else unpack8(ldg16(dy + ((y + a.dy_pad) * wdp + xx + a.dy_pad) * a.c), dyv);
    xraw = ldg16(x + ((y + a.x_pad) * wxp + xx + a.x_pad) * a.c);
    if constexpr (RES) rraw = ldg16(res + ((y + a.res_pad) * wrp + xx + a.res_pad) * a.c);
  }

  // g = dy * act'(pre) (in place in dyv) and xhat
  __device__ __forceinline__ void grad(const pcgan_norm_bwd_args& a, float (&dyv)[8], const uint4 xraw, const uint4 rraw,
                                       float (&xh)[8]) const {
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

template <bool GEN, bool RES>
__global__ void __launch_bounds__(kT, GEN ? 2 : 3) norm_bwd_reduce_kernel(pcgan_norm_bwd_args a, int vec_per_block, int lcv) {
  __shared__ float red[kT * 16];
  const int n = blockIdx.y;
  const int cv = a.c >> 3;
  const int total = a.h * a.w * cv;
  const int vb = blockIdx.x * vec_per_block, ve = min(vb + vec_per_block, total);
  const int c0 = (threadIdx.x & (cv - 1)) << 3;
  BwdCtx<GEN, RES> k;
  k.init(a, n, c0);
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  for (int v0 = vb + threadIdx.x; v0 < ve; v0 += kT * 2) {
    float g[2][8];
    uint4 xr[2], rr[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int v = v0 + u * kT;
      if (v < ve) {
        const int pix = v >> lcv;
        const int y = pix / a.w;
        k.load(a, y, pix - y * a.w, g[u], xr[u], rr[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (v0 + u * kT < ve) {
        float xh[8];
        k.grad(a, g[u], xr[u], rr[u], xh);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s1[j] += g[u][j]; s2[j] = fmaf(g[u][j], xh[j], s2[j]); }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[threadIdx.x * 16 + j] = s1[j]; red[threadIdx.x * 16 + 8 + j] = s2[j]; }
  __syncthreads();
  const int lanes = kT / cv;
  for (int t = threadIdx.x; t < cv * 16; t += kT) {
    const int c = t >> 4, slot = t & 15;
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[(l * cv + c) * 16 + slot];
    const int ch = (c << 3) + (slot & 7);
    const int64_t o = (static_cast<int64_t>(a.groups > 1 ? n : 0) * a.c + ch) * 2 + (slot >> 3);
    atomicAdd(a.sums + o, s);
  }
}

template <bool GEN, bool RES>
__global__ void __launch_bounds__(kT, GEN ? 2 : 3) norm_bwd_apply_kernel(pcgan_norm_bwd_args a, int vec_per_block, int lcv) {
  // blocks walk the samples (and chunks) in the opposite order to the reduce pass: what that pass read last is
  // still in L2 when this one starts
  const int n = gridDim.y - 1 - blockIdx.y;
  const int bx = gridDim.x - 1 - blockIdx.x;
  const int cv = a.c >> 3;
  const int total = a.h * a.w * cv;
  const int vb = bx * vec_per_block, ve = min(vb + vec_per_block, total);
  const int c0 = (threadIdx.x & (cv - 1)) << 3;
  BwdCtx<GEN, RES> k;
  k.init(a, n, c0);
  float A[8], B[8];
  const float inv = a.count > 0.f ? 1.f / a.count : 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { A[j] = 0.f; B[j] = 0.f; }
  if (a.count > 0.f) {
    const int64_t so = (static_cast<int64_t>(a.groups > 1 ? n : 0) * a.c + c0) * 2;
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
  const int wop = a.w + 2 * a.dx_pad, wsp = a.w + 2 * a.dres_pad;
  __nv_bfloat16* dx = a.dx ? reinterpret_cast<__nv_bfloat16*>(a.dx) + static_cast<int64_t>(n) * (a.h + 2 * a.dx_pad) * wop * a.c + c0 : nullptr;
  __nv_bfloat16* dres = a.dres ? reinterpret_cast<__nv_bfloat16*>(a.dres) + static_cast<int64_t>(n) * (a.h + 2 * a.dres_pad) * wsp * a.c + c0 : nullptr;
  for (int v0 = vb + threadIdx.x; v0 < ve; v0 += kT * 2) {
    float g[2][8];
    uint4 xr[2], rr[2];
    int yy[2], xx[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int v = v0 + u * kT;
      if (v < ve) {
        const int pix = v >> lcv;
        yy[u] = pix / a.w;
        xx[u] = pix - yy[u] * a.w;
        k.load(a, yy[u], xx[u], g[u], xr[u], rr[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (v0 + u * kT < ve) {
        float xh[8];
        k.grad(a, g[u], xr[u], rr[u], xh);
        if (dres) store8(dres + ((yy[u] + a.dres_pad) * wsp + xx[u] + a.dres_pad) * a.c, g[u]);
        if (dx) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float t = g[u][j] - A[j] - xh[j] * B[j];
            if (scaled) t *= k.sc[j];
            if constexpr (GEN) t *= k.mk[j];
            o[j] = t;
          }
          store8(dx + ((yy[u] + a.dx_pad) * wop + xx[u] + a.dx_pad) * a.c, o);
        }
      }
    }
  }
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

}  // namespace pcgan

using namespace pcgan;
#define STREAM(s) static_cast<cudaStream_t>(s)

extern "C" int pcgan_norm_finalize(const pcgan_norm_finalize_args* a, pcgan_stream_t s) {
  if (!a || !a->stats || a->groups < 1 || a->c < 1 || a->count <= 0.f) return fail(PCGAN_ERR_INVALID, "norm_finalize: bad argument");
  norm_finalize_kernel<<<(a->c + 31) / 32, kT, 0, STREAM(s)>>>(*a);
  PCGAN_LAUNCH_OK("norm_finalize_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_norm_apply(const pcgan_norm_apply_args* a, pcgan_stream_t s) {
  if (!a || !a->x || !a->y) return fail(PCGAN_ERR_INVALID, "norm_apply: null argument");
  int lcv, rc = check_c(a->c, "norm_apply", &lcv);
  if (rc) return rc;
  if ((a->scale == nullptr) != (a->shift == nullptr)) return fail(PCGAN_ERR_INVALID, "norm_apply: scale and shift go together");
  if (a->y_halo == PCGAN_HALO_REFLECT && (a->y_pad >= a->h || a->y_pad >= a->w)) return fail(PCGAN_ERR_INVALID, "norm_apply: reflect pad too large");
  if (a->n < 1 || a->n > 65535) return fail(PCGAN_ERR_UNSUPPORTED, "norm_apply: n=%d (1..65535)", a->n);
  const int hp = a->h + 2 * a->y_pad, wp = a->w + 2 * a->y_pad;
  const int64_t rowvec = static_cast<int64_t>(wp) * (a->c / 8);
  if (rowvec * hp >= (1ll << 28)) return fail(PCGAN_ERR_UNSUPPORTED, "norm_apply: sample too large");
  int per;
  chunking(rowvec * hp, a->n, static_cast<int>(rowvec), &per);
  const int rows = per / static_cast<int>(rowvec);
  const dim3 grid((hp + rows - 1) / rows, a->n);
  if (a->res) norm_apply_kernel<true><<<grid, kT, 0, STREAM(s)>>>(*a, rows, lcv);
  else norm_apply_kernel<false><<<grid, kT, 0, STREAM(s)>>>(*a, rows, lcv);
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
  halo_fold_kernel<<<dim3(chunks, a->n), kT, 0, STREAM(s)>>>(*a, per, lcv);
  PCGAN_LAUNCH_OK("halo_fold_kernel");
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
  const int mp = a->dy_pad > a->x_pad ? a->dy_pad : a->x_pad;
  if (static_cast<int64_t>(a->h + 2 * mp) * (a->w + 2 * mp) * a->c >= (1ll << 31)) return fail(PCGAN_ERR_UNSUPPORTED, "norm_bwd: sample too large");
  return PCGAN_OK;
}

extern "C" int pcgan_norm_bwd_reduce(const pcgan_norm_bwd_args* a, pcgan_stream_t s) {
  int lcv, rc = check_bwd(a, &lcv);
  if (rc) return rc;
  if (!a->sums) return fail(PCGAN_ERR_INVALID, "norm_bwd_reduce: sums is null");
  int per;
  const int chunks = chunking(static_cast<int64_t>(a->h) * a->w * (a->c / 8), a->n, kT, &per);
  const dim3 grid(chunks, a->n);
  const bool res = a->res != nullptr && a->act != PCGAN_ACT_NONE;   // the residual only matters through the activation mask
  const bool gen = a->affine != 0 || a->drop_mask != nullptr || res;
  if (!gen) norm_bwd_reduce_kernel<false, false><<<grid, kT, 0, STREAM(s)>>>(*a, per, lcv);
  else if (!res) norm_bwd_reduce_kernel<true, false><<<grid, kT, 0, STREAM(s)>>>(*a, per, lcv);
  else norm_bwd_reduce_kernel<true, true><<<grid, kT, 0, STREAM(s)>>>(*a, per, lcv);
  PCGAN_LAUNCH_OK("norm_bwd_reduce_kernel");
  return PCGAN_OK;
}

extern "C" int pcgan_norm_bwd_apply(const pcgan_norm_bwd_args* a, pcgan_stream_t s) {
  int lcv, rc = check_bwd(a, &lcv);
  if (rc) return rc;
  if (!a->dx && !a->dres) return fail(PCGAN_ERR_INVALID, "norm_bwd_apply: no output");
  int per;
  const int chunks = chunking(static_cast<int64_t>(a->h) * a->w * (a->c / 8), a->n, kT, &per);
  const dim3 grid(chunks, a->n);
  const bool res = a->res != nullptr && a->act != PCGAN_ACT_NONE;
  const bool gen = a->affine != 0 || a->drop_mask != nullptr || res;
  if (!gen) norm_bwd_apply_kernel<false, false><<<grid, kT, 0, STREAM(s)>>>(*a, per, lcv);
  else if (!res) norm_bwd_apply_kernel<true, false><<<grid, kT, 0, STREAM(s)>>>(*a, per, lcv);
  else norm_bwd_apply_kernel<true, true><<<grid, kT, 0, STREAM(s)>>>(*a, per, lcv);
  PCGAN_LAUNCH_OK("norm_bwd_apply_kernel");
  return PCGAN_OK;
}
