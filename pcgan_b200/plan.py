"""Host-side planner: turns a convolution (forward, data-gradient, weight-gradient,
transposed) over padded NHWC bf16 buffers into pcgan_igemm_desc data for the
tcgen05 implicit-GEMM kernel (include/pcgan_kernels.h).

Nothing here touches the GPU: a plan is pure data (TMA views, tap tables, output
maps, weight index maps), so it is validated on the CPU by the emulator in
oracle/igemm_emulator.py against torch.nn.functional convolutions.

Geometry vocabulary
  Geom(n, h, w, c, pad): a padded NHWC buffer [n][h+2*pad][w+2*pad][c]; "interior"
  is the h x w image, the halo is written by the producer kernel (zeros or
  reflection).  All element offsets below are in bf16 elements.

Reference semantics implemented (phymhan/pc-gan):
  nn.Conv2d / nn.ConvTranspose2d as used in models/networks.py:578-605, :621-648,
  :747-775, :1014-1027 and models/resnet.py:20-28,134.
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch

from . import _lib as L

MAX_TAPS = L.MAX_TAPS


@dataclass(frozen=True)
class Geom:
    n: int
    h: int
    w: int
    c: int
    pad: int

    @property
    def hp(self):
        return self.h + 2 * self.pad

    @property
    def wp(self):
        return self.w + 2 * self.pad

    @property
    def numel(self):
        return self.n * self.hp * self.wp * self.c

    def off(self, n, y, x):
        """element offset of interior pixel (n, y, x), channel 0"""
        return ((n * self.hp + y + self.pad) * self.wp + x + self.pad) * self.c


SLACK = 512  # elements appended to every activation buffer (packed-row windows read past the end)


@dataclass
class OutMap:
    """Where output pixel (n, y, x), channel ch lands: base + n*sn + y*sy + x*sx + ch*sc (elements)."""
    base: int
    sn: int
    sy: int
    sx: int
    sc: int = 1
    dtype: int = L.DT_BF16

    @staticmethod
    def nhwc(g: Geom, dtype=L.DT_BF16, ystep=1, xstep=1, y0=0, x0=0):
        """interior of a padded NHWC buffer; (ystep, y0) write every ystep-th row starting at y0 (sub-pixel phases)"""
        return OutMap(base=g.off(0, y0, x0), sn=g.hp * g.wp * g.c, sy=g.wp * g.c * ystep, sx=g.c * xstep, sc=1, dtype=dtype)

    @staticmethod
    def nchw(n, c, h, w, dtype=L.DT_F32):
        return OutMap(base=0, sn=c * h * w, sy=w, sx=1, sc=h * w, dtype=dtype)


@dataclass
class IgemmSpec:
    kind: int = L.IGEMM_KMAJOR
    block_n: int = 64
    a_dims: List[int] = field(default_factory=lambda: [1] * 5)
    a_strides: List[int] = field(default_factory=lambda: [0] * 5)  # bytes
    a_box: List[int] = field(default_factory=lambda: [64, 1, 1, 1, 1])
    b_dims: List[int] = field(default_factory=lambda: [1] * 5)
    b_strides: List[int] = field(default_factory=lambda: [0] * 5)
    b_box: List[int] = field(default_factory=lambda: [64, 1, 1, 1, 1])
    t_count: List[int] = field(default_factory=lambda: [1] * 4)
    a_base: List[int] = field(default_factory=lambda: [0] * 4)
    a_step: List[List[int]] = field(default_factory=lambda: [[0] * 4 for _ in range(4)])
    b_base: List[int] = field(default_factory=lambda: [0] * 4)
    b_step: List[List[int]] = field(default_factory=lambda: [[0] * 4 for _ in range(4)])
    n_tiles: int = 1
    m_tiles: int = 1
    ksplit: int = 1
    cchunks: int = 1
    tap_off: List[List[int]] = field(default_factory=list)
    tap_c0: List[int] = field(default_factory=list)
    tap_bk: List[int] = field(default_factory=list)
    e_base: List[int] = field(default_factory=lambda: [0] * 4)
    e_step: List[List[int]] = field(default_factory=lambda: [[0] * 4 for _ in range(4)])
    e_p1: List[int] = field(default_factory=lambda: [0] * 4)
    e_p2: List[int] = field(default_factory=lambda: [0] * 4)
    # e_comp[d][k] = (lo, hi, stride)
    e_comp: List[List[Tuple[int, int, int]]] = field(
        default_factory=lambda: [[(0, 1, 0), (0, 1 << 30, 0), (0, 1, 0)] for _ in range(4)])
    out_dtype: int = L.DT_BF16
    act: int = L.ACT_NONE
    act_slope: float = 0.0
    n_valid: int = 1
    out_cstride: int = 1
    stats_mode: int = L.STATS_NONE
    stats_dim: int = -1
    stats_comp: int = 1
    stats_div: int = 0         # > 1: statistics group = sample // stats_div (several BatchNorm batches in one launch)
    m_valid: int = 1
    wg_ncols: int = 1
    ldo: int = 1
    # host-side extras (not part of the C struct)
    a_elem_offset: int = 0     # added to the A base pointer (elements)
    b_elem_offset: int = 0
    out_elem_offset: int = 0   # added to the output base pointer (elements)
    b_rows: int = 0            # packed weight matrix shape [b_rows][b_k] (KMAJOR) / wgrad output [m][ldo]
    b_k: int = 0
    flops: int = 0             # 2*MACs actually issued (incl. padding waste), for bookkeeping
    note: str = ""
    shift_taps: int = 0        # shift-sum epilogue: horizontal taps carried in N (include/pcgan_kernels.h)
    shift_cpad: int = 0
    pair: int = 0              # CTA pairs issuing one tcgen05.mma.cta_group::2 of M = 256 (include/pcgan_kernels.h)
    a_window: int = 0          # 8: A is the plain 8-channel tensor read through an overlapping descriptor (pcgan_kernels.h)
    wg_box_dim: int = 0        # WGRAD: the 64-column boxes of an N tile step along this B tensor dim (filter rows in N)
    tf32: bool = False         # operands are fp32 tensors multiplied as TF32 (tcgen05.mma.kind::tf32): K chunks of 32 elements
    swap_operands: bool = False  # WGRAD: M side = input activations, N side = dY (ConvRT.backward_weight passes them so)
    box_taps: Optional[List[int]] = None   # wg_box_dim: kidx of the filter row held by 64-column box i of the output

    @property
    def num_taps(self):
        return len(self.tap_off)

    @property
    def kc(self):
        """elements of one 128-byte K chunk"""
        return 32 if self.tf32 else 64

    @property
    def esz(self):
        return 4 if self.tf32 else 2

    def to_desc(self) -> L.IgemmDesc:
        d = L.IgemmDesc()
        d.kind, d.block_n = self.kind, self.block_n
        for i in range(5):
            d.a.dims[i], d.a.strides[i], d.a.box[i] = self.a_dims[i], self.a_strides[i], self.a_box[i]
            d.b.dims[i], d.b.strides[i], d.b.box[i] = self.b_dims[i], self.b_strides[i], self.b_box[i]
        for j in range(4):
            d.t_count[j] = self.t_count[j]
            d.a_base[j], d.b_base[j], d.e_base[j] = self.a_base[j], self.b_base[j], self.e_base[j]
            d.e_p1[j], d.e_p2[j] = self.e_p1[j], self.e_p2[j]
            for k in range(4):
                d.a_step[j][k], d.b_step[j][k], d.e_step[j][k] = self.a_step[j][k], self.b_step[j][k], self.e_step[j][k]
            for k in range(3):
                lo, hi, st = self.e_comp[j][k]
                d.e_comp[j][k].lo, d.e_comp[j][k].hi, d.e_comp[j][k].stride = lo, hi, st
        d.n_tiles, d.m_tiles, d.ksplit = self.n_tiles, self.m_tiles, self.ksplit
        if self.num_taps > MAX_TAPS:
            raise ValueError("too many taps: %d" % self.num_taps)
        d.num_taps, d.cchunks = self.num_taps, self.cchunks
        for t in range(self.num_taps):
            for k in range(4):
                d.tap_off[t][k] = self.tap_off[t][k]
            d.tap_c0[t], d.tap_bk[t] = self.tap_c0[t], self.tap_bk[t]
        d.out_dtype, d.act, d.act_slope, d.n_valid = self.out_dtype, self.act, self.act_slope, self.n_valid
        d.out_cstride = self.out_cstride
        d.stats_mode, d.stats_dim, d.stats_comp = self.stats_mode, self.stats_dim, self.stats_comp
        d.m_valid, d.wg_ncols, d.ldo = self.m_valid, self.wg_ncols, self.ldo
        d.pair = self.pair
        d.shift_taps, d.shift_cpad = self.shift_taps, self.shift_cpad
        d.a_window = self.a_window
        d.wg_box_dim = self.wg_box_dim
        d.tf32 = int(self.tf32)
        d.stats_div = int(self.stats_div)
        return d


def _ceil(a, b):
    return -(-a // b)


def _block_n(cout):
    """UMMA N for `cout` output columns: multiple of 16, <= 256, minimising padded columns.  (128-wide tiles were
    measured 25 % slower per FLOP than 256-wide ones on the 256 -> 256 3x3 convolution: the A operand is re-read from
    shared memory for half as many columns.)"""
    if cout <= 256:
        return max(16, _ceil(cout, 16) * 16)
    best = None
    for bn in (256, 192, 128):
        waste = _ceil(cout, bn) * bn - cout
        if best is None or waste < best[0]:
            best = (waste, bn)
    return best[1]


PAIRING = True   # CTA pairs (one cta_group::2 MMA for two M tiles) where it pays: full-width N tiles, enough M tiles
TAPBOX = True    # packed-window weight gradients: the filter rows ride in N, the 64-channel operand is read once per N tile
WINDOW = True    # 8-channel inputs: windowed A operand (pcgan_igemm_desc.a_window) instead of overlapping-stride TMA boxes


def _pair_kmajor(s: "IgemmSpec", m_tiles: int) -> int:
    """Pair two M tiles (one tcgen05.mma.cta_group::2 of M = 256) for full-width N tiles; narrower tiles run faster
    unpaired through two pipelines with double-buffered accumulators (measured: tools/bench_conv.py, [pair] / [dual]).
    TF32 plans run unpaired."""
    return int(PAIRING and not s.tf32 and s.block_n >= 256 and m_tiles * s.n_tiles >= 2)


ANY = (0, 1 << 30, 0)
ONE = (0, 1, 0)


def _set_weights_tmap(s: IgemmSpec, rows: int, k: int):
    E = s.esz
    s.b_dims = [k, rows, 1, 1, 1]
    s.b_strides = [0, k * E, k * E * max(rows, 1), k * E * max(rows, 1), k * E * max(rows, 1)]
    s.b_box = [s.kc, s.block_n, 1, 1, 1]
    s.b_rows, s.b_k = rows, k


# ------------------------------------------------------------------------------------------
# KMAJOR plans: forward / data-gradient
# ------------------------------------------------------------------------------------------
def _choose_box(wo, ho, n, single_image):
    bw = min(wo, 128)
    bh = max(1, min(ho, 128 // bw))
    bn = 1
    if not single_image and bh == ho and bw == wo:
        bn = max(1, min(n, 128 // (bw * bh)))
    return bw, bh, bn


def plan_box(xg: Geom, taps: List[Tuple[int, int, int]], cin: int, cout: int, ho: int, wo: int, stride: int,
             out: OutMap, *, act=L.ACT_NONE, act_slope=0.0, stats=False, per_sample_stats=False,
             note="", tf32=False) -> IgemmSpec:
    """Generic "box" plan.  Output pixel (n, y, x) = sum over taps (dy, dx, kidx) and channels of
    Xpadded[n][stride*y + dy][stride*x + dx][:] . Wpacked[:, kidx*cin : (kidx+1)*cin]   (dy, dx >= 0 are
    coordinates in the PADDED buffer).  Requires cin % 64 == 0 and xg.c == cin.
    stride 1 uses a 4-D view (c, x, y, n); stride 2 a 5-D phase view (c, px, X, py, Y*n).
    """
    s = IgemmSpec(kind=L.IGEMM_KMAJOR, note=note, tf32=tf32)
    kc, E = s.kc, s.esz
    assert xg.c == cin and cin % kc == 0, (xg, cin)
    assert stride in (1, 2)
    s.block_n = _block_n(cout)
    s.n_tiles = _ceil(cout, s.block_n)
    s.n_valid = cout
    s.cchunks = cin // kc
    C, Hp, Wp, N = xg.c, xg.hp, xg.wp, xg.n
    bw, bh, bn = _choose_box(wo, ho, N, per_sample_stats or stride == 2)
    tx, ty, tn = _ceil(wo, bw), _ceil(ho, bh), _ceil(N, bn)
    s.t_count = [tx, ty, tn, 1]
    if stride == 1:
        s.a_dims = [C, Wp, Hp, N, 1]
        s.a_strides = [0, C * E, Wp * C * E, Hp * Wp * C * E, N * Hp * Wp * C * E]
        s.a_box = [kc, bw, bh, bn, 1]
        s.a_step[0][0], s.a_step[1][1], s.a_step[2][2] = bw, bh, bn
        for (dy, dx, kidx) in taps:
            s.tap_off.append([dx, dy, 0, 0])
            s.tap_c0.append(0)
            s.tap_bk.append(kidx * cin)
        # epilogue: box-local dims are (x, y, n, -)
        s.e_step[0][0], s.e_step[1][1], s.e_step[2][2] = bw, bh, bn
        s.e_comp = [[ONE, (0, wo, out.sx), ONE], [ONE, (0, ho, out.sy), ONE], [ONE, (0, N, out.sn), ONE], [ONE, ANY, ONE]]
        if stats:
            s.stats_mode = L.STATS_ON
            s.stats_dim, s.stats_comp = (2, 1) if per_sample_stats else (-1, 1)
    else:
        assert Hp % 2 == 0 and Wp % 2 == 0, "stride-2 phase view needs even padded extents"
        s.a_dims = [C, 2, Wp // 2, 2, (Hp // 2) * N]
        s.a_strides = [0, C * E, 2 * C * E, Wp * C * E, 2 * Wp * C * E]
        s.a_box = [kc, 1, bw, 1, bh]
        # outer dims: d0 = px, d1 = X, d2 = py, d3 = Y (+ n * Hp/2)
        s.a_step[0][1] = bw
        s.a_step[1][3] = bh
        s.a_step[2][3] = Hp // 2
        for (dy, dx, kidx) in taps:
            s.tap_off.append([dx % 2, dx // 2, dy % 2, dy // 2])
            s.tap_c0.append(0)
            s.tap_bk.append(kidx * cin)
        period = ty * bh
        s.e_step[0][1] = bw
        s.e_step[1][3] = bh
        s.e_step[2][3] = period
        s.e_p1[3] = period
        s.e_comp = [[ONE, ANY, ONE], [ONE, (0, wo, out.sx), ONE], [ONE, ANY, ONE], [(0, N, out.sn), (0, ho, out.sy), ONE]]
        if stats:
            s.stats_mode = L.STATS_ON
            s.stats_dim, s.stats_comp = (3, 0) if per_sample_stats else (-1, 1)
    ktot = (max(k for _, _, k in taps) + 1) * cin
    _set_weights_tmap(s, cout, ktot)
    s.out_dtype, s.out_cstride, s.out_elem_offset = out.dtype, out.sc, out.base
    s.act, s.act_slope = act, act_slope
    s.flops = 2 * tx * ty * tn * 128 * s.n_tiles * s.block_n * len(taps) * cin
    s.pair = _pair_kmajor(s, tx * ty * tn)
    return s


def plan_flat(xg: Geom, taps: List[Tuple[int, int, int]], cin: int, cout: int, out: OutMap,
              yr: Tuple[int, int], xr: Tuple[int, int], *, act=L.ACT_NONE, act_slope=0.0, stats=False,
              note="", tf32=False) -> IgemmSpec:
    """"Flat" plan over the flattened padded grid of xg: for every padded position q = (n, Y, X)
    out(q) = sum over taps (dy, dx, kidx) of Xpadded[q + dy*Wp + dx] . W[:, kidx*cin:...]   (dy, dx may be
    negative; rows wrap into neighbouring rows/images, which is harmless when the wrapped reads hit a zero
    halo or the wrapped outputs are masked).  Only positions with yr[0] <= Y < yr[1], xr[0] <= X < xr[1] are
    stored, at out(n, Y - yr[0], X - xr[0]).  Statistics are per channel over all stored rows (batch norm)."""
    s = IgemmSpec(kind=L.IGEMM_KMAJOR, note=note, tf32=tf32)
    kc, E = s.kc, s.esz
    assert xg.c == cin and cin % kc == 0
    s.block_n = _block_n(cout)
    s.n_tiles = _ceil(cout, s.block_n)
    s.n_valid = cout
    s.cchunks = cin // kc
    C, Hp, Wp, N = xg.c, xg.hp, xg.wp, xg.n
    P = N * Hp * Wp
    s.a_dims = [C, P, 1, 1, 1]
    s.a_strides = [0, C * E, P * C * E, P * C * E, P * C * E]
    s.a_box = [kc, 128, 1, 1, 1]
    # only tiles that intersect stored rows: positions from first valid to last valid
    q_lo = yr[0] * Wp + xr[0]
    q_hi = (N - 1) * Hp * Wp + (yr[1] - 1) * Wp + xr[1]
    first = q_lo // 128
    tiles = _ceil(q_hi, 128) - first
    s.t_count = [tiles, 1, 1, 1]
    s.a_base[0] = first * 128
    s.a_step[0][0] = 128
    for (dy, dx, kidx) in taps:
        s.tap_off.append([dy * Wp + dx, 0, 0, 0])
        s.tap_c0.append(0)
        s.tap_bk.append(kidx * cin)
    s.e_base[0] = first * 128
    s.e_step[0][0] = 128
    s.e_p1[0], s.e_p2[0] = Hp * Wp, Wp
    s.e_comp = [[(0, N, out.sn), (yr[0], yr[1], out.sy), (xr[0], xr[1], out.sx)], [ONE, ANY, ONE], [ONE, ANY, ONE], [ONE, ANY, ONE]]
    if stats:
        s.stats_mode, s.stats_dim = L.STATS_ON, -1
    ktot = (max(k for _, _, k in taps) + 1) * cin
    _set_weights_tmap(s, cout, ktot)
    s.out_dtype, s.out_cstride, s.out_elem_offset = out.dtype, out.sc, out.base
    s.act, s.act_slope = act, act_slope
    s.flops = 2 * tiles * 128 * s.n_tiles * s.block_n * len(taps) * cin
    s.pair = _pair_kmajor(s, tiles)
    return s


def plan_shift_flat(xg: Geom, kh: int, kw: int, cin: int, cout: int, row_taps: List[Tuple[int, int, int]], out: OutMap,
                    yr: Tuple[int, int], xr: Tuple[int, int], *, out_shift=0, act=L.ACT_NONE, act_slope=0.0, note="", tf32=False) -> IgemmSpec:
    """Few-output-channel convolution over the flattened padded grid of xg with the horizontal taps moved into N
    ("shift-sum" epilogue): for GEMM row position t
        partial[t][j*4 + c] = sum over filter rows (dy, dx, r) in row_taps and channels of
                              Xpadded[t + dy*Wp + dx][:] . W[j*4 + c][r*cin : (r+1)*cin]
    and the value stored at flat position g = t + out_shift is sum_j partial[t + j][j*4 + c], j < kw.  Tiles are 128 rows
    stepping by 128 - (kw - 1).  Positions with yr[0] <= Y < yr[1], xr[0] <= X < xr[1] are stored at
    out(n, Y - yr[0], X - xr[0]).  Every activation row is fetched kh times instead of kh*kw times."""
    s = IgemmSpec(kind=L.IGEMM_KMAJOR, note=note, tf32=tf32)
    kc, E = s.kc, s.esz
    assert xg.c == cin and cin % kc == 0 and cout <= 4 and kw * 4 <= 32
    s.block_n, s.n_tiles, s.n_valid = 32, 1, cout
    s.shift_taps, s.shift_cpad = kw, 4
    s.cchunks = cin // kc
    C, Hp, Wp, N = xg.c, xg.hp, xg.wp, xg.n
    P = N * Hp * Wp
    S = 128 - (kw - 1)
    s.a_dims = [C, P, 1, 1, 1]
    s.a_strides = [0, C * E, P * C * E, P * C * E, P * C * E]
    s.a_box = [kc, 128, 1, 1, 1]
    q_lo = yr[0] * Wp + xr[0]
    q_hi = (N - 1) * Hp * Wp + (yr[1] - 1) * Wp + xr[1]
    tiles = _ceil(q_hi - q_lo, S)
    s.t_count = [tiles, 1, 1, 1]
    s.a_base[0] = q_lo - out_shift
    s.a_step[0][0] = S
    for (dy, dx, r) in row_taps:
        s.tap_off.append([dy * Wp + dx, 0, 0, 0])
        s.tap_c0.append(0)
        s.tap_bk.append(r * cin)
    s.e_base[0] = q_lo
    s.e_step[0][0] = S
    s.e_p1[0], s.e_p2[0] = Hp * Wp, Wp
    s.e_comp = [[(0, N, out.sn), (yr[0], yr[1], out.sy), (xr[0], xr[1], out.sx)], [ONE, ANY, ONE], [ONE, ANY, ONE], [ONE, ANY, ONE]]
    _set_weights_tmap(s, 32, kh * cin)
    s.out_dtype, s.out_cstride, s.out_elem_offset = out.dtype, out.sc, out.base
    s.act, s.act_slope = act, act_slope
    s.flops = 2 * tiles * 128 * 32 * kh * cin
    return s


def wmap_shift(w_shape, k: int, kdim: int, *, dgrad=False, rows=None, cols=None) -> torch.Tensor:
    """Packed operand of plan_shift_flat, [32][len(rows)*kdim]: row j*4 + c, column i*kdim + q, where filter row
    r = rows[i] and filter column s = cols[j] (defaults: all k rows; forward cols[j] = j, data gradient cols[j] = k-1-j).
    forward (OIHW, kdim = Cin):   W[c][q][r][s]
    data gradient (kdim = Cout):  W[q][c][r][s]   (c = input channel receiving the gradient)"""
    d0, d1 = w_shape[0], w_shape[1]
    rows = list(range(k)) if rows is None else list(rows)
    if cols is None:
        cols = list(range(k)) if not dgrad else [k - 1 - j for j in range(k)]
    nr = len(rows)
    idx = torch.full((8, 4, nr, kdim), -1, dtype=torch.int64)
    c = torch.arange(4).view(-1, 1, 1)
    r = torch.tensor(rows).view(1, -1, 1)
    q = torch.arange(kdim).view(1, 1, -1)
    for j, s_ in enumerate(cols):
        if not dgrad:
            valid = (c < d0) & (q < d1)
            flat = ((c * d1 + q) * k + r) * k + s_
        else:
            valid = (c < d1) & (q < d0)
            flat = ((q * d1 + c) * k + r) * k + s_
        idx[j] = torch.where(valid, flat, torch.full_like(flat, -1)).expand(4, nr, kdim)
    return idx.reshape(-1).to(torch.int32)


def plan_packed(xg: Geom, kh: int, kw: int, stride: int, off: int, cout: int, ho: int, wo: int, out: OutMap, *,
                act=L.ACT_NONE, act_slope=0.0, stats=False, per_sample_stats=False, note="", tf32=False) -> IgemmSpec:
    """Small-Cin plan ("packed rows"): one K chunk covers a whole filter row, because in NHWC the kw taps x C
    channels of a row are contiguous: window = Xpadded[n][stride*y + r + off][stride*x + off ...][0 : kw*C].
    The A view has overlapping strides (dim 1 advances by stride pixels, dim 0 spans 64*cchunks elements).
    Packed weights: W[co][r][s*C + c], zero beyond kw*C.  xg.c in {8, 16, 32}."""
    C, Hp, Wp, N = xg.c, xg.hp, xg.wp, xg.n
    s = IgemmSpec(kind=L.IGEMM_KMAJOR, note=note, tf32=tf32)
    kc, E = s.kc, s.esz
    assert C % 8 == 0 and C < 64
    s.block_n = _block_n(cout)
    s.n_tiles = _ceil(cout, s.block_n)
    s.n_valid = cout
    win = _ceil(kw * C, kc) * kc
    s.cchunks = win // kc
    bw, bh, bn = _choose_box(wo, ho, N, per_sample_stats or stride >= 2)
    tx, ty, tn = _ceil(wo, bw), _ceil(ho, bh), _ceil(N, bn)
    window = WINDOW and not tf32 and stride == 1 and C == 8 and kw <= 8 and wo >= 64
    flat_tiles = _ceil(N * Hp * Wp, 128)
    if window and not stats and flat_tiles < _ceil(wo, 128) * ho * N:
        # windowed form over the flattened padded grid: position q = (n, Y, X) computes output (Y, X) from the window
        # starting at q + (r + off)*Wp + off; rows that run over the end of an image row read the next row's pixels
        # and are not stored
        s.a_window = 8
        P = N * Hp * Wp
        s.t_count = [flat_tiles, 1, 1, 1]
        s.a_dims = [8, P, 1, 1, 1]
        s.a_strides = [0, 16, P * 16, P * 16, P * 16]
        s.a_box = [8, 128 + 7, 1, 1, 1]
        s.a_step[0][0] = 128
        for r in range(kh):
            s.tap_off.append([(r + off) * Wp + off, 0, 0, 0])
            s.tap_c0.append(0)
            s.tap_bk.append(r * win)
        s.e_step[0][0] = 128
        s.e_p1[0], s.e_p2[0] = Hp * Wp, Wp
        s.e_comp = [[(0, N, out.sn), (0, ho, out.sy), (0, wo, out.sx)], [ONE, ANY, ONE], [ONE, ANY, ONE], [ONE, ANY, ONE]]
        tx, ty, tn = flat_tiles, 1, 1
    elif window:
        # windowed form, one tile = up to 128 consecutive pixels of one output row
        s.a_window = 8
        bw = min(wo, 128)
        tx, ty, tn = _ceil(wo, bw), ho, N
        s.t_count = [tx, ty, tn, 1]
        s.a_dims = [8, Wp, Hp, N, 1]
        s.a_strides = [0, 16, Wp * 16, Hp * Wp * 16, N * Hp * Wp * 16]
        s.a_box = [8, bw + 7, 1, 1, 1]
        s.a_step[0][0], s.a_step[1][1], s.a_step[2][2] = bw, 1, 1
        for r in range(kh):
            s.tap_off.append([off, r + off, 0, 0])
            s.tap_c0.append(0)
            s.tap_bk.append(r * win)
        s.e_step[0][0], s.e_step[1][1], s.e_step[2][2] = bw, 1, 1
        s.e_comp = [[ONE, (0, wo, out.sx), ONE], [ONE, (0, ho, out.sy), ONE], [ONE, (0, N, out.sn), ONE], [ONE, ANY, ONE]]
        if stats:
            s.stats_mode = L.STATS_ON
            s.stats_dim, s.stats_comp = (2, 1) if per_sample_stats else (-1, 1)
    else:
        s.t_count = [tx, ty, tn, 1]
    if window:
        pass
    elif stride == 1:
        s.a_dims = [win, Wp, Hp, N, 1]
        s.a_strides = [0, C * E, Wp * C * E, Hp * Wp * C * E, N * Hp * Wp * C * E]
        s.a_box = [kc, bw, bh, bn, 1]
        s.a_step[0][0], s.a_step[1][1], s.a_step[2][2] = bw, bh, bn
        for r in range(kh):
            s.tap_off.append([off, r + off, 0, 0])
            s.tap_c0.append(0)
            s.tap_bk.append(r * win)
        s.e_step[0][0], s.e_step[1][1], s.e_step[2][2] = bw, bh, bn
        s.e_comp = [[ONE, (0, wo, out.sx), ONE], [ONE, (0, ho, out.sy), ONE], [ONE, (0, N, out.sn), ONE], [ONE, ANY, ONE]]
        if stats:
            s.stats_mode = L.STATS_ON
            s.stats_dim, s.stats_comp = (2, 1) if per_sample_stats else (-1, 1)
    else:
        st = stride
        assert st in (2, 4) and Hp % st == 0 and Wp % st == 0, "strided packed rows need padded extents that are multiples of the stride"
        # dims: (window, X [st pixels per step], row phase, Y (+ n*Hp/st)); the x offset `off` goes into the base pointer
        s.a_dims = [win, Wp // st, st, (Hp // st) * N, 1]
        s.a_strides = [0, st * C * E, Wp * C * E, st * Wp * C * E, (Hp // st) * N * st * Wp * C * E]
        s.a_box = [kc, bw, 1, bh, 1]
        s.a_elem_offset = off * C
        s.a_step[0][0] = bw
        s.a_step[1][2] = bh
        s.a_step[2][2] = Hp // st
        for r in range(kh):
            s.tap_off.append([0, (r + off) % st, (r + off) // st, 0])
            s.tap_c0.append(0)
            s.tap_bk.append(r * win)
        period = ty * bh
        s.e_step[0][0] = bw
        s.e_step[1][2] = bh
        s.e_step[2][2] = period
        s.e_p1[2] = period
        s.e_comp = [[ONE, (0, wo, out.sx), ONE], [ONE, ANY, ONE], [(0, N, out.sn), (0, ho, out.sy), ONE], [ONE, ANY, ONE]]
        if stats:
            s.stats_mode = L.STATS_ON
            s.stats_dim, s.stats_comp = (2, 0) if per_sample_stats else (-1, 1)
    _set_weights_tmap(s, cout, kh * win)
    s.out_dtype, s.out_cstride, s.out_elem_offset = out.dtype, out.sc, out.base
    s.act, s.act_slope = act, act_slope
    s.flops = 2 * tx * ty * tn * 128 * s.n_tiles * s.block_n * kh * win
    return s


# ------------------------------------------------------------------------------------------
# WGRAD plans
# ------------------------------------------------------------------------------------------
def _next_pow2(v):
    b = 1
    while b < v:
        b *= 2
    return b


def _pix_box64(w, h):
    """(bw, bh) with bw*bh == 64 covering a w x h pixel grid in ceil(w/bw) x ceil(h/bh) boxes of one sample.
    Boxes may over-cover: the M-side tensor is zero outside its interior (zero halo / TMA zero fill)."""
    bw = min(64, _next_pow2(w))
    bh = 64 // bw
    return bw, bh


def _ksplit_for(total_kb, out_tiles, paired=False, sms=148):
    """Split of the pixel loop: one work item per SM for paired plans (a pair shares one MMA stream), two per SM
    otherwise, so that both pipelines of an unpaired CTA have a tile."""
    slots = sms if paired else 2 * sms
    ks = max(1, min(total_kb, slots // max(out_tiles, 1)))
    # keep at least 8 K blocks per split so the pipeline fills
    while ks > 1 and total_kb // ks < 8:
        ks -= 1
    return ks


def plan_wgrad_box(mg: Geom, m_ch: int, ng: Geom, n_ch: int, taps: List[Tuple[int, int, int]], ph: int, pw: int,
                   n_stride: int, *, m_origin=(0, 0), n_packed_win: int = 0, m_packed_win: int = 0, pix_box=None,
                   note="", tf32=False) -> IgemmSpec:
    """Weight gradient over a ph x pw pixel grid per sample.
    M side: tensor mg (padded NHWC), pixel (n, y, x) read at padded (y + m_origin[0], x + m_origin[1]), m_ch channels.
    N side: tensor ng, pixel read at padded (n_stride*y + dy, n_stride*x + dx) for tap (dy, dx, kidx);
            n_ch channels (or, if n_packed_win > 0, the packed-row window of that many elements).
    Output fp32 [m_ch][ldo] with column = kidx*ncols + channel.
    m_packed_win > 0: the M side is read as a window of that many contiguous elements starting at the pixel (rows
    beyond m_ch are the following pixels' channels; they are computed and dropped).  pix_box = (bw, bh) overrides
    the 64-pixel box shape."""
    s = IgemmSpec(kind=L.IGEMM_WGRAD, note=note, tf32=tf32)
    kc, E = s.kc, s.esz          # TF32: boxes of 64 pixels x 32 channels (128 bytes), K step of the MMA = 8 pixels
    bw, bh = pix_box if pix_box else _pix_box64(pw, ph)
    assert bw * bh == 64
    bn = 1
    assert mg.c % 8 == 0 and ng.c % 8 == 0
    ncols = n_packed_win if n_packed_win else n_ch
    s.block_n = min(256, _ceil(ncols, 64) * 64)
    s.n_tiles = _ceil(ncols, s.block_n)
    s.m_tiles = _ceil(m_ch, 128)
    s.m_valid, s.wg_ncols = m_ch, ncols
    N = mg.n
    tx, ty, tn = _ceil(pw, bw), _ceil(ph, bh), N
    s.t_count = [tx, ty, tn, 1]
    # M-side view (c, x, y, n)
    s.a_dims = [m_packed_win if m_packed_win else mg.c, mg.wp, mg.hp, N, 1]
    s.a_strides = [0, mg.c * E, mg.wp * mg.c * E, mg.hp * mg.wp * mg.c * E, N * mg.hp * mg.wp * mg.c * E]
    s.a_box = [kc, bw, bh, bn, 1]
    s.a_base = [m_origin[1], m_origin[0], 0, 0]
    s.a_step[0][0], s.a_step[1][1], s.a_step[2][2] = bw, bh, bn
    C, Hp, Wp = ng.c, ng.hp, ng.wp
    d0 = n_packed_win if n_packed_win else C
    dys = sorted(dy for dy, _, _ in taps)
    kh = len(taps)
    if (TAPBOX and not tf32 and n_packed_win == 64 and n_stride == 1 and kh > 1 and len(set(dx for _, dx, _ in taps)) == 1
            and dys == list(range(dys[0], dys[0] + kh)) and ph + dys[0] + kh - 1 <= Hp and dys[0] >= 0):
        # filter rows in N: the B view gets a filter-row dimension (same stride as the image row) and 64-column box i of
        # the N tiles is the window of row dys[0] + i, so the 64-channel M operand is fetched once per N tile of four
        # filter rows instead of once per row.  Output column = i*64 + window element; s.box_taps[i] = kidx of box i.
        by_dy = {dy: kidx for dy, _, kidx in taps}
        s.box_taps = [by_dy[dys[0] + i] for i in range(kh)]
        s.wg_box_dim = 2
        s.block_n = 256
        s.n_tiles = _ceil(kh * 64, 256)
        s.wg_ncols = kh * 64
        s.b_dims = [64, Wp, kh, Hp - (kh - 1), N]
        s.b_strides = [0, C * 2, Wp * C * 2, Wp * C * 2, Hp * Wp * C * 2]
        s.b_box = [64, bw, 1, bh, 1]
        s.b_step[0][0], s.b_step[1][2], s.b_step[2][3] = bw, bh, bn
        s.tap_off.append([taps[0][1], 0, dys[0], 0])
        s.tap_c0.append(0)
        s.tap_bk.append(0)
        s.ldo = kh * 64
        s.b_rows, s.b_k = m_ch, s.ldo
        total_kb = tx * ty * tn
        s.pair = int(PAIRING and s.m_tiles % 2 == 0)
        s.ksplit = _ksplit_for(total_kb, s.m_tiles * s.n_tiles, bool(s.pair))
        s.flops = 2 * total_kb * 64 * 128 * s.m_tiles * s.n_tiles * s.block_n
        return s
    if n_stride == 1:
        s.b_dims = [d0, Wp, Hp, N, 1]
        s.b_strides = [0, C * E, Wp * C * E, Hp * Wp * C * E, N * Hp * Wp * C * E]
        s.b_box = [kc, bw, bh, bn, 1]
        s.b_step[0][0], s.b_step[1][1], s.b_step[2][2] = bw, bh, bn
        for (dy, dx, kidx) in taps:
            s.tap_off.append([dx, dy, 0, 0])
            s.tap_c0.append(0)
            s.tap_bk.append(kidx * ncols)
    else:
        assert n_stride == 2 and Hp % 2 == 0 and bn == 1
        if n_packed_win:
            s.b_dims = [d0, Wp // 2, 2, (Hp // 2) * N, 1]
            s.b_strides = [0, 2 * C * E, Wp * C * E, 2 * Wp * C * E, (Hp // 2) * N * 2 * Wp * C * E]
            s.b_box = [kc, bw, 1, bh, 1]
            s.b_step[0][0], s.b_step[1][2], s.b_step[2][2] = bw, bh, Hp // 2
            xoffs = set(dx for _, dx, _ in taps)
            assert len(xoffs) == 1
            s.b_elem_offset = xoffs.pop() * C
            for (dy, dx, kidx) in taps:
                s.tap_off.append([0, dy % 2, dy // 2, 0])
                s.tap_c0.append(0)
                s.tap_bk.append(kidx * ncols)
        else:
            assert Wp % 2 == 0
            s.b_dims = [d0, 2, Wp // 2, 2, (Hp // 2) * N]
            s.b_strides = [0, C * E, 2 * C * E, Wp * C * E, 2 * Wp * C * E]
            s.b_box = [kc, 1, bw, 1, bh]
            s.b_step[0][1], s.b_step[1][3], s.b_step[2][3] = bw, bh, Hp // 2
            for (dy, dx, kidx) in taps:
                s.tap_off.append([dx % 2, dx // 2, dy % 2, dy // 2])
                s.tap_c0.append(0)
                s.tap_bk.append(kidx * ncols)
    nk = max(k for _, _, k in taps) + 1
    s.ldo = nk * ncols
    s.b_rows, s.b_k = m_ch, s.ldo
    total_kb = tx * ty * tn
    s.pair = int(PAIRING and not tf32 and s.m_tiles % 2 == 0 and (s.block_n // 64) % 2 == 0)
    s.ksplit = _ksplit_for(total_kb, len(taps) * s.m_tiles * s.n_tiles, bool(s.pair))
    s.flops = 2 * total_kb * 64 * 128 * s.m_tiles * s.n_tiles * s.block_n * len(taps)
    return s


def plan_wgrad_small_cout(xg: Geom, cin: int, dyg: Geom, k: int, o: int, *, note="") -> IgemmSpec:
    """Weight gradient of a stride-1 k x k convolution with very few output channels (the generator head,
    networks.py:603-604: 64 -> 3): putting dY on the M side would pad 3 rows to 128, so the operands are swapped and the
    8-channel dY buffer is read as a packed window of 8 pixels x 8 channels on the N side:
        out[ci][r*64 + j*8 + co] = sum_{n, y', t} Xp[n][y'][t][ci] * dYp[n][y' + PD - r - o][t + PD - o - (k-1) + j][co]
                                 = dW[co][ci][r][k-1-j]                      (j < k; Xp, dYp: padded buffers, PD = dyg.pad)
    i.e. one MMA column block covers all k horizontal taps of a filter row as 8 pixel shifts of dY.  The M side is the
    channel vector of pixel t (window of 128 elements = 2 pixels; the second pixel's rows are dropped: m_valid = cin).
    o = xg.pad - conv padding.  Requires cin == xg.c == 64, dyg.c == 8, k <= 8, dyg.pad >= o + k - 1 and zero halos."""
    assert xg.c == cin == 64 and dyg.c == 8 and k <= 8
    PD = dyg.pad
    assert PD >= o + k - 1 and xg.pad * 2 - o <= PD, (PD, o, k, xg.pad)
    taps = [(PD - r - o, PD - o - (k - 1), r) for r in range(k)]
    s = plan_wgrad_box(xg, cin, dyg, 8, taps, xg.hp, xg.wp, 1, m_origin=(0, 0), n_packed_win=64, m_packed_win=128,
                       pix_box=(8, 8), note=note)
    s.swap_operands = True
    return s


def wmap_small_cout(w_shape, k: int, box_taps=None) -> torch.Tensor:
    """Scatter map of plan_wgrad_small_cout: packed[ci][i*64 + j*8 + co] -> W[co][ci][r][k-1-j] with r = i, or
    r = box_taps[i] when the filter rows ride in N (IgemmSpec.box_taps)."""
    cout, cin = w_shape[0], w_shape[1]
    idx = torch.full((cin, k, 8, 8), -1, dtype=torch.int64)
    ci = torch.arange(cin).view(-1, 1, 1, 1)
    r = torch.tensor(list(box_taps) if box_taps is not None else list(range(k))).view(1, -1, 1, 1)
    j = torch.arange(8).view(1, 1, -1, 1)
    co = torch.arange(8).view(1, 1, 1, -1)
    flat = ((co * cin + ci) * k + r) * k + (k - 1 - j)
    valid = (j < k) & (co < cout)
    idx = torch.where(valid, flat, torch.full_like(flat, -1)).expand(cin, k, 8, 8)
    return idx.reshape(-1).to(torch.int32)


# ------------------------------------------------------------------------------------------
# Weight index maps: packed[i] = W.flatten()[idx[i]] (or 0 when idx[i] < 0)
# ------------------------------------------------------------------------------------------
def wmap_taps(w_shape, rows, taps, cin, *, transposed_layout=False, swap=False) -> torch.Tensor:
    """Vectorised map for per-tap layouts: packed[row][kidx*cin + c] = W[...] with taps = [(r, s, kidx)].
    OIHW weights (nn.Conv2d): row = o, c = i  (swap=False)  -> W[o][i][r][s]
                              row = i, c = o  (swap=True)   -> W[o][i][r][s]   (data-gradient operand)
    IOHW weights (nn.ConvTranspose2d, transposed_layout=True): W[i][o][r][s]; row = o, c = i when swap=False
    (forward operand: contraction over the in-channels i), row = i, c = o when swap=True."""
    d0, d1, kh, kw = w_shape
    nk = max(k for _, _, k in taps) + 1
    idx = torch.full((rows, nk * cin), -1, dtype=torch.int64)
    rr = torch.arange(rows).view(-1, 1)
    cc = torch.arange(cin).view(1, -1)
    for (r, s_, kidx) in taps:
        if not transposed_layout:
            o, i = (cc, rr) if swap else (rr, cc)
            valid = (o < d0) & (i < d1)
            flat = ((o * d1 + i) * kh + r) * kw + s_
        else:
            # stored [i][o][kh][kw]
            o, i = (cc, rr) if swap else (rr, cc)
            valid = (i < d0) & (o < d1)
            flat = ((i * d1 + o) * kh + r) * kw + s_
        idx[:, kidx * cin:(kidx + 1) * cin] = torch.where(valid, flat, torch.full_like(flat, -1))
    return idx.reshape(-1).to(torch.int32)


def wmap_packed(w_shape, rows, kh, kw, c_buf, win, *, flip=False, swap=False) -> torch.Tensor:
    """Map for packed-row layouts: packed[row][r*win + s*c_buf + c] = W[row][c][r][s] (swap=False, OIHW)
    or W[c][row][kh-1-r][kw-1-s] (swap=True with flip: data-gradient through a stride-1 conv)."""
    d0, d1, _, _ = w_shape
    idx = torch.full((rows, kh * win), -1, dtype=torch.int64)
    rr = torch.arange(rows).view(-1, 1)
    cc = torch.arange(c_buf).view(1, -1)
    for r in range(kh):
        for s_ in range(kw):
            wr, ws = (kh - 1 - r, kw - 1 - s_) if flip else (r, s_)
            o, i = (cc, rr) if swap else (rr, cc)
            valid = (o < d0) & (i < d1)
            flat = ((o * d1 + i) * kh + wr) * kw + ws
            base = r * win + s_ * c_buf
            idx[:, base:base + c_buf] = torch.where(valid, flat, torch.full_like(flat, -1))
    return idx.reshape(-1).to(torch.int32)
