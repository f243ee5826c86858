"""Thin Python wrappers over the C ABI (include/pcgan_kernels.h): torch tensors in, raw
pointers out.  torch is used only for device memory and streams; every computation is a
launch of a kernel from libpcgan_kernels.so.  No fallback: a missing library or a
non-CUDA tensor raises.
"""
import ctypes as C

import torch

from . import _lib as L
from .plan import Geom, IgemmSpec, SLACK


def _ptr(t, elem_offset=0):
    if t is None:
        return None
    if not t.is_cuda:
        raise L.PcganError("pcgan ops need CUDA tensors (no CPU fallback); got %s" % t.device)
    if t.device.index != torch.cuda.current_device():
        # kernels go to the current device's stream (and the library sizes grids for the current device)
        raise L.PcganError("tensor on %s but the current CUDA device is %d: call torch.cuda.set_device (BaseModel.initialize does)"
                           % (t.device, torch.cuda.current_device()))
    return t.data_ptr() + elem_offset * t.element_size()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class Stats:
    """Launch bookkeeping: `launches` counts every kernel this package launches; when `igemm_events` is a list,
    every igemm launch is bracketed by CUDA events on its own stream (bench.py's roofline pass)."""
    launches = 0
    igemm_launches = 0
    igemm_events = None
    igemm_flops = 0
    op_events = None


def _count(n=1):
    Stats.launches += n


class _Timed:
    """When Stats.op_events is a list, every wrapped launch is bracketed by CUDA events (bench.py's per-kernel table);
    nbytes = the launch's algorithmic HBM traffic (tensors it must read and write once), for the HBM roofline."""

    def __init__(self, name, nbytes=0):
        self.name, self.e0, self.nbytes = name, None, int(nbytes)

    def __enter__(self):
        if Stats.op_events is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            Stats.op_events.append((self.name, self.e0, e1, self.nbytes))
        return False


class Igemm:
    """A planned implicit GEMM (one kernel launch per run)."""

    def __init__(self, spec: IgemmSpec):
        self.spec = spec
        self.lib = L.load()
        self._h = C.c_void_p()
        desc = spec.to_desc()
        L.check(self.lib.pcgan_igemm_plan_create(C.byref(desc), C.byref(self._h)), "igemm_plan_create(%s)" % spec.note)

    def __del__(self):
        try:
            if self._h:
                self.lib.pcgan_igemm_plan_destroy(self._h)
        except Exception:
            pass

    def run(self, a, b, out, bias=None, stats=None):
        s = self.spec
        Stats.launches += 1
        Stats.igemm_launches += 1
        ev = Stats.igemm_events
        if ev is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        L.check(self.lib.pcgan_igemm_run(self._h, _ptr(a, s.a_elem_offset), _ptr(b, s.b_elem_offset),
                                         _ptr(out, s.out_elem_offset), _ptr(bias), _ptr(stats), _stream()),
                "igemm_run(%s)" % s.note)
        if ev is not None:
            e1.record()
            ev.append((s.note, s.flops, e0, e1))


def alloc_act(g: Geom, device, zero=True):
    """Padded NHWC bf16 activation buffer (flat, with slack for packed-row windows)."""
    n = g.numel + SLACK
    return torch.zeros(n, dtype=torch.bfloat16, device=device) if zero else torch.empty(n, dtype=torch.bfloat16, device=device)


def gather_cast_bf16(src, idx, dst):
    _count()
    with _Timed("gather_cast_bf16"):
        L.check(L.load().pcgan_gather_cast_bf16(_ptr(src), _ptr(idx), _ptr(dst), idx.numel(), _stream()), "gather_cast_bf16")


def gather_tf32(src, idx, dst):
    """fp32 packed operand of a TF32 plan (values rounded to nearest TF32)"""
    _count()
    with _Timed("gather_tf32"):
        L.check(L.load().pcgan_gather_tf32(_ptr(src), _ptr(idx), _ptr(dst), idx.numel(), _stream()), "gather_tf32")


def scatter_f32(src, idx, dst, accumulate=False):
    _count()
    with _Timed("scatter_f32"):
        L.check(L.load().pcgan_scatter_f32(_ptr(src), _ptr(idx), _ptr(dst), idx.numel(), int(accumulate), _stream()), "scatter_f32")


def batch_table(items, device):
    """Device table of pcgan_batch_item for the batched gather / scatter: items = [(src, idx, dst)] tensors."""
    rows = [[t_src.data_ptr(), t_idx.data_ptr(), t_dst.data_ptr(), t_idx.numel()] for t_src, t_idx, t_dst in items]
    return torch.tensor(rows, dtype=torch.int64).to(device), max(r[3] for r in rows)


def gather_cast_bf16_batched(table, count, max_n):
    _count()
    with _Timed("gather_cast_bf16_batched"):
        L.check(L.load().pcgan_gather_cast_bf16_batched(_ptr(table), count, max_n, _stream()), "gather_cast_bf16_batched")


def scatter_f32_batched(table, count, max_n, accumulate=True):
    _count()
    with _Timed("scatter_f32_batched"):
        L.check(L.load().pcgan_scatter_f32_batched(_ptr(table), count, max_n, int(accumulate), _stream()), "scatter_f32_batched")


def running_table(items, device):
    """Device table of pcgan_running_item: items = [(stats, running_mean, running_var, num_batches_tracked, groups, c, count,
    momentum)] (tensors or None)."""
    import struct
    raw = bytearray()
    for it in items:
        st, rm, rv, nbt, groups, c, count, mom = it[:8]
        seq = int(it[8]) if len(it) > 8 else 0
        raw += struct.pack("<qqqqiiffii", st.data_ptr(), rm.data_ptr() if rm is not None else 0, rv.data_ptr() if rv is not None else 0,
                           nbt.data_ptr() if nbt is not None else 0, groups, c, float(count), float(mom), seq, 0)
    t = torch.frombuffer(raw, dtype=torch.uint8).clone().to(device)
    return t, max(it[5] for it in items)


def norm_running_batched(table, count, max_c):
    _count()
    with _Timed("norm_running_batched"):
        L.check(L.load().pcgan_norm_running_batched(_ptr(table), count, max_c, _stream()), "norm_running_batched")


def pack_nchw(src, dst, g: Geom, *, z=None, mul_out=None, mul_kind=L.ACT_TANH, halo=L.HALO_ZERO):
    """src: NCHW fp32 [n, cs, h, w] -> dst buffer of geometry g (resized to g.h x g.w when they differ)."""
    n, cs, h, w = src.shape
    assert src.dtype == torch.float32 and src.is_contiguous()
    a = L.PackArgs(src=_ptr(src), z=_ptr(z), mul_out=_ptr(mul_out), mul_kind=mul_kind, dst=_ptr(dst), n=n, cs=cs, h=h, w=w,
                   ho=g.h, wo=g.w, cd=g.c, pad=g.pad, halo=halo, dst_n_stride=0)
    _count()
    with _Timed("pack_nchw", src.numel() * 4 * (1 + (mul_out is not None)) + g.numel * 2):
        L.check(L.load().pcgan_pack_nchw(C.byref(a), _stream()), "pack_nchw")


def unpack_resize_bwd(gbuf, g: Geom, dst, *, accumulate=False, scale=1.0):
    """gbuf: NHWC bf16 gradient of geometry g -> dst NCHW fp32 [n, cd, h, w] (adjoint of the resize if sizes differ)."""
    n, cd, h, w = dst.shape
    assert dst.dtype == torch.float32 and dst.is_contiguous()
    a = L.UnpackArgs(g=_ptr(gbuf), dst=_ptr(dst), n=n, c=g.c, hs=g.h, ws=g.w, pad=g.pad, cd=cd, h=h, w=w,
                     accumulate=int(accumulate), scale=scale)
    _count()
    with _Timed("unpack_resize_bwd", g.numel * 2 + dst.numel() * 4):
        L.check(L.load().pcgan_unpack_resize_bwd(C.byref(a), _stream()), "unpack_resize_bwd")


def norm_finalize(stats, groups, c, count, *, eps=1e-5, momentum=0.1, gamma=None, beta=None, mean=None, rstd=None,
                  scale=None, shift=None, running_mean=None, running_var=None, drop_mask=None, in_groups=0):
    a = L.NormFinalizeArgs(stats=_ptr(stats), groups=groups, c=c, count=float(count), eps=eps, momentum=momentum,
                           gamma=_ptr(gamma), beta=_ptr(beta), mean=_ptr(mean), rstd=_ptr(rstd), scale=_ptr(scale),
                           shift=_ptr(shift), running_mean=_ptr(running_mean), running_var=_ptr(running_var),
                           drop_mask=_ptr(drop_mask), in_groups=in_groups)
    _count()
    with _Timed("norm_finalize"):
        L.check(L.load().pcgan_norm_finalize(C.byref(a), _stream()), "norm_finalize")


def norm_apply(x, xg: Geom, y, yg: Geom, *, y_halo=L.HALO_ZERO, scale=None, shift=None, groups=1, res=None, res_pad=0,
               res_scale=None, res_shift=None, res_groups=1, drop_mask=None, act=L.ACT_NONE, act_slope=0.0, post_mask=None,
               stats=None, count=0.0, eps=1e-5, gamma=None, beta=None, mean_out=None, rstd_out=None, scale_out=None, shift_out=None,
               y_c0=None):
    """stats given: the finalize (statistics -> scale / shift) is fused into this launch; scale / shift are ignored.
    y_c0 given: y is a buffer of yg.c >= xg.c channels and the xg.c output channels go to [y_c0, y_c0 + xg.c)."""
    if y_c0 is None:
        assert (xg.n, xg.h, xg.w, xg.c) == (yg.n, yg.h, yg.w, yg.c)
        y_c, y_c0 = 0, 0
    else:
        assert (xg.n, xg.h, xg.w) == (yg.n, yg.h, yg.w) and y_c0 + xg.c <= yg.c
        y_c = yg.c
    a = L.NormApplyArgs(x=_ptr(x), x_pad=xg.pad, res=_ptr(res), res_pad=res_pad, y=_ptr(y), y_pad=yg.pad, y_halo=y_halo,
                        n=xg.n, h=xg.h, w=xg.w, c=xg.c, scale=_ptr(scale), shift=_ptr(shift), groups=groups,
                        res_scale=_ptr(res_scale), res_shift=_ptr(res_shift), res_groups=res_groups,
                        drop_mask=_ptr(drop_mask), act=act, act_slope=act_slope, post_mask=_ptr(post_mask),
                        stats=_ptr(stats), count=float(count), eps=eps, gamma=_ptr(gamma), beta=_ptr(beta),
                        mean_out=_ptr(mean_out), rstd_out=_ptr(rstd_out), scale_out=_ptr(scale_out), shift_out=_ptr(shift_out),
                        y_c=y_c, y_c0=y_c0)
    _count()
    elems = xg.n * xg.h * xg.w * xg.c
    with _Timed("norm_apply", elems * 2 * (2 + (res is not None))):
        L.check(L.load().pcgan_norm_apply(C.byref(a), _stream()), "norm_apply")


def halo_fold(gpad, gg: Geom, out, out_pad, *, halo=L.HALO_REFLECT, add=None, add_pad=0):
    a = L.FoldArgs(gpad=_ptr(gpad), g_pad=gg.pad, halo=halo, add=_ptr(add), add_pad=add_pad, out=_ptr(out),
                   out_pad=out_pad, n=gg.n, h=gg.h, w=gg.w, c=gg.c)
    _count()
    with _Timed("halo_fold", 2 * gg.n * gg.c * (gg.hp * gg.wp + gg.h * gg.w * (1 + (add is not None)))):
        L.check(L.load().pcgan_halo_fold(C.byref(a), _stream()), "halo_fold")


def halo_accumulate(gpad, gg: Geom):
    """Reflect fold in place: afterwards the interior of the padded gradient `gpad` is the gradient of the un-padded tensor
    (consumers drop the halo: norm backward dy_fold=1, halo_fold(halo=HALO_ZERO))."""
    a = L.FoldArgs(gpad=_ptr(gpad), g_pad=gg.pad, halo=L.HALO_REFLECT, add=None, add_pad=0, out=None, out_pad=0,
                   n=gg.n, h=gg.h, w=gg.w, c=gg.c)
    _count()
    band = 2 * gg.pad * (gg.w + gg.h - 2 * gg.pad)
    with _Timed("halo_accumulate", 2 * gg.n * gg.c * band * 3):
        L.check(L.load().pcgan_halo_accumulate(C.byref(a), _stream()), "halo_accumulate")


def _bwd_args(dy, dy_pad, x, xg, *, res=None, res_pad=0, mean=None, rstd=None, scale=None, shift=None, groups=1,
              res_scale=None, res_shift=None, res_groups=1, drop_mask=None, act=L.ACT_NONE, act_slope=0.0, count=0.0,
              sums=None, dx=None, dx_pad=0, dres=None, dres_pad=0, dy_fold=0, affine=1, post_mask=None):
    return L.NormBwdArgs(dy=_ptr(dy), dy_pad=dy_pad, x=_ptr(x), x_pad=xg.pad, res=_ptr(res), res_pad=res_pad,
                         mean=_ptr(mean), rstd=_ptr(rstd), scale=_ptr(scale), shift=_ptr(shift), groups=groups,
                         res_scale=_ptr(res_scale), res_shift=_ptr(res_shift), res_groups=res_groups,
                         drop_mask=_ptr(drop_mask), act=act, act_slope=act_slope, n=xg.n, h=xg.h, w=xg.w, c=xg.c,
                         count=float(count), sums=_ptr(sums), dx=_ptr(dx), dx_pad=dx_pad, dres=_ptr(dres), dres_pad=dres_pad,
                         dy_fold=dy_fold, post_mask=_ptr(post_mask), affine=int(affine))


def norm_bwd_reduce(dy, dy_pad, x, xg, **kw):
    a = _bwd_args(dy, dy_pad, x, xg, **kw)
    _count()
    elems = xg.n * xg.h * xg.w * xg.c
    with _Timed("norm_bwd_reduce", elems * 2 * (2 + (kw.get("res") is not None))):
        L.check(L.load().pcgan_norm_bwd_reduce(C.byref(a), _stream()), "norm_bwd_reduce")


def norm_bwd_apply(dy, dy_pad, x, xg, **kw):
    a = _bwd_args(dy, dy_pad, x, xg, **kw)
    _count()
    elems = xg.n * xg.h * xg.w * xg.c
    with _Timed("norm_bwd_apply", elems * 2 * (3 + (kw.get("res") is not None) + (kw.get("dres") is not None))):
        L.check(L.load().pcgan_norm_bwd_apply(C.byref(a), _stream()), "norm_bwd_apply")


# PCGAN_NORM_FUSED=1 switches the one-launch cluster kernel on.  Measured on B200 at the ResnetBlock shape (64 x 32 x 32 x 256):
# 70 us per launch against 42 us for reduce + apply (15 resident clusters of 8 CTAs with 146 KB each run in lock step:
# load, reduce, store never overlap), so the two passes stay the default; the kernel is kept (and tested) as the base of a
# persistent, double-buffered variant.
NORM_FUSED = __import__("os").environ.get("PCGAN_NORM_FUSED", "0") == "1"


def norm_bwd(dy, dy_pad, x, xg, **kw):
    """Backward of one normalisation + activation unit: the one-launch cluster kernel when the arguments qualify (lean
    InstanceNorm path whose rows fit the cluster's shared memory), else reduce + apply."""
    a = _bwd_args(dy, dy_pad, x, xg, **kw)
    lib = L.load()
    if NORM_FUSED and lib.pcgan_norm_bwd_fused_supported(C.byref(a)):
        _count()
        elems = xg.n * xg.h * xg.w * xg.c
        with _Timed("norm_bwd_fused", elems * 2 * 3):
            L.check(lib.pcgan_norm_bwd_fused(C.byref(a), _stream()), "norm_bwd_fused")
        return
    norm_bwd_reduce(dy, dy_pad, x, xg, **kw)
    norm_bwd_apply(dy, dy_pad, x, xg, **kw)


def pool_out(h, pool_pad=1):
    """output side of the 3x3 stride-2 max pooling of an h x h map"""
    return (h + 2 * pool_pad - 3) // 2 + 1


def maxpool_fwd(x, xg: Geom, y, y_pad, idx, pool_pad=1):
    a = L.MaxpoolArgs(x=_ptr(x), x_pad=xg.pad, y=_ptr(y), y_pad=y_pad, idx=_ptr(idx), n=xg.n, h=xg.h, w=xg.w, c=xg.c, pool_pad=pool_pad)
    _count()
    with _Timed("maxpool_fwd", xg.n * xg.h * xg.w * xg.c * 2 + xg.n * xg.h * xg.w * xg.c * 3 // 4):
        L.check(L.load().pcgan_maxpool3x3s2_fwd(C.byref(a), _stream()), "maxpool_fwd")


def maxpool_bwd(dy, dy_pad, idx, dx, dx_pad, n, h, w, c, pool_pad=1):
    _count()
    with _Timed("maxpool_bwd", n * h * w * c * 2 + n * h * w * c * 3 // 4):
        L.check(L.load().pcgan_maxpool3x3s2_bwd(_ptr(dy), dy_pad, _ptr(idx), _ptr(dx), dx_pad, n, h, w, c, pool_pad, _stream()), "maxpool_bwd")


def act_bwd(dy, dy_pad, y, y_pad, dx, dx_pad, g: Geom, slope=0.0):
    """dx = y > 0 ? dy : slope * dy over the interior of geometry g (the three buffers have their own halo widths)"""
    _count()
    with _Timed("act_bwd", g.n * g.h * g.w * g.c * 2 * 3):
        L.check(L.load().pcgan_act_bwd(_ptr(dy), dy_pad, _ptr(y), y_pad, _ptr(dx), dx_pad, g.n, g.h, g.w, g.c, float(slope), _stream()), "act_bwd")


def nhwc_to_f32(buf, g: Geom, dst):
    """interior of the padded NHWC bf16 buffer -> contiguous fp32 [n, h, w, c]"""
    _count()
    with _Timed("nhwc_cast", g.n * g.h * g.w * g.c * 6):
        L.check(L.load().pcgan_nhwc_cast(_ptr(buf), _ptr(dst), g.pad, g.n, g.h, g.w, g.c, 1, _stream()), "nhwc_cast")


def f32_to_nhwc(src, buf, g: Geom):
    _count()
    with _Timed("nhwc_cast", g.n * g.h * g.w * g.c * 6):
        L.check(L.load().pcgan_nhwc_cast(_ptr(src), _ptr(buf), g.pad, g.n, g.h, g.w, g.c, 0, _stream()), "nhwc_cast")


def resize_nchw_fwd(src, dst):
    n, c, h, w = src.shape
    _count()
    with _Timed("resize_nchw_fwd", (src.numel() + dst.numel()) * 4):
        L.check(L.load().pcgan_resize_nchw_fwd(_ptr(src), _ptr(dst), n * c, h, w, dst.shape[2], dst.shape[3], _stream()), "resize_nchw_fwd")


def resize_nchw_bwd(gdst, gsrc):
    n, c, h, w = gsrc.shape
    _count()
    with _Timed("resize_nchw_bwd", (gdst.numel() + gsrc.numel()) * 4):
        L.check(L.load().pcgan_resize_nchw_bwd(_ptr(gdst), _ptr(gsrc), n * c, h, w, gdst.shape[2], gdst.shape[3], _stream()), "resize_nchw_bwd")


def loss(kind, p, target, *, per_sample=0, weight=1.0, weight_dev=None, loss_out=None, grad=None):
    a = L.LossArgs(kind=kind, p=_ptr(p), target=_ptr(target), n=p.numel(), per_sample=per_sample, weight=weight,
                   weight_dev=_ptr(weight_dev), loss=_ptr(loss_out), grad=_ptr(grad))
    _count()
    with _Timed("loss"):
        L.check(L.load().pcgan_loss(C.byref(a), _stream()), "loss")


def adam_table(items, device):
    """Device table of pcgan_adam_item: items = [(p, g, m, v)] fp32 tensors of equal numel."""
    rows = [[p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel()] for p, g, m, v in items]
    return torch.tensor(rows, dtype=torch.int64).to(device), max(r[4] for r in rows)


def adam_batched(table, count, max_n, lr, beta1, beta2, eps, step):
    """One Adam update of every tensor in the table (+ the one-thread launch that advances `step`)."""
    _count(2)
    with _Timed("adam_batched"):
        L.check(L.load().pcgan_adam_batched(_ptr(table), count, max_n, _ptr(lr), beta1, beta2, eps, _ptr(step), _stream()), "adam_batched")


def adam(p, g, m, v, lr, beta1, beta2, eps, step):
    _count()
    with _Timed("adam"):
        L.check(L.load().pcgan_adam(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), _ptr(lr), beta1, beta2, eps, _ptr(step), _stream()), "adam")
