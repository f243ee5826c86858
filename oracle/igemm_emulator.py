"""TEST INFRASTRUCTURE ONLY — CPU emulator of the pcgan_igemm kernel contract.

Executes a pcgan_b200.plan.IgemmSpec exactly as include/pcgan_kernels.h specifies it
(TMA box reads with zero fill outside the tensor, tap tables, mixed-radix tile
coordinates, the epilogue's output map, bias / statistics / activation), but with
torch on the CPU.  It validates the *planner* against torch.nn.functional
convolutions without a GPU; the product never imports it.
"""
import torch

ACT = {0: lambda x, s: x, 1: lambda x, s: torch.relu(x), 2: lambda x, s: torch.where(x > 0, x, x * s),
       3: lambda x, s: torch.tanh(x), 4: lambda x, s: torch.sigmoid(x)}


def tma_gather(flat, dims, strides_bytes, box, coords, esz=2):
    """flat: 1-D float tensor (bf16 values); coords: [T, 5] int64 start coordinates.
    Returns [T, rows, box[0]]; rows enumerate box dims 1..4 with dim 1 fastest."""
    T = coords.shape[0]
    inner = box[0]
    rows = box[1] * box[2] * box[3] * box[4]
    r = torch.arange(rows)
    loc = []
    rem = r
    for d in range(1, 5):
        loc.append(rem % box[d])
        rem = rem // box[d]
    k = torch.arange(inner)
    # element coordinates
    c0 = coords[:, 0].view(T, 1, 1) + k.view(1, 1, inner)
    valid = (c0 >= 0) & (c0 < dims[0])
    addr = c0.clone()
    for d in range(1, 5):
        cd = coords[:, d].view(T, 1, 1) + loc[d - 1].view(1, rows, 1)
        valid = valid & (cd >= 0) & (cd < dims[d])
        addr = addr + cd * (strides_bytes[d] // esz)
    addr = torch.where(valid, addr, torch.zeros_like(addr))
    if addr.numel() and (int(addr.max()) >= flat.numel() or int(addr.min()) < 0):
        raise IndexError("in-bounds TMA element outside the buffer: max %d numel %d" % (int(addr.max()), flat.numel()))
    vals = flat[addr.reshape(-1)].reshape(T, rows, inner)
    return torch.where(valid, vals, torch.zeros_like(vals))


def _digits(idx, count):
    out = []
    for j in range(4):
        out.append(idx % count[j])
        idx = idx // count[j]
    return out


def _coords(dig, base, step):
    return [base[d] + sum(dig[j] * step[j][d] for j in range(4)) for d in range(4)]


def run_kmajor(s, a_flat, b_flat, out_flat, bias=None, stats=None):
    """a_flat/b_flat: float32 1-D tensors holding bf16-representable values (already offset by *_elem_offset);
    out_flat: 1-D float tensor written in place (values rounded to bf16 when s.out_dtype == 0)."""
    m_tiles = s.t_count[0] * s.t_count[1] * s.t_count[2] * s.t_count[3]
    mt = torch.arange(m_tiles)
    dig = _digits(mt, s.t_count)
    ac = _coords(dig, s.a_base, s.a_step)
    a_window = getattr(s, "a_window", 0)
    tf32 = bool(getattr(s, "tf32", False))
    kc, esz = (32, 4) if tf32 else (64, 2)      # elements of a 128-byte K chunk, bytes per element
    a_rows = s.a_box[1] * s.a_box[2] * s.a_box[3] * s.a_box[4] - (7 if a_window else 0)
    assert 1 <= a_rows <= 128
    if a_window:
        assert a_window == 8 and s.a_box[0] == 8 and s.a_dims[0] == 8 and s.a_strides[1] == 16 and s.cchunks == 1
        assert s.a_box[2] == s.a_box[3] == s.a_box[4] == 1
    # epilogue row map
    r = torch.arange(128)
    valid = (r < a_rows).view(1, 128).expand(m_tiles, 128).clone()
    off = torch.zeros(m_tiles, 128, dtype=torch.int64)
    eb = _coords(dig, s.e_base, s.e_step)
    group = torch.zeros(m_tiles, dtype=torch.int64)
    rem = r
    for d in range(4):
        i = rem % s.a_box[d + 1]
        rem = rem // s.a_box[d + 1]
        g = eb[d].view(-1, 1) + i.view(1, 128)
        p1, p2 = s.e_p1[d], s.e_p2[d]
        if p1 > 0:
            c0, rm = torch.div(g, p1, rounding_mode="floor"), g % p1
        else:
            c0, rm = torch.zeros_like(g), g
        if p2 > 0:
            c1, c2 = torch.div(rm, p2, rounding_mode="floor"), rm % p2
        else:
            c1, c2 = rm, torch.zeros_like(rm)
        valid &= g >= 0
        for c, (lo, hi, st) in zip((c0, c1, c2), s.e_comp[d]):
            valid &= (c >= lo) & (c < hi)
            off += (c - lo) * st
        if d == s.stats_dim:
            g0 = eb[d]
            s0 = torch.div(g0, p1, rounding_mode="floor") if p1 > 0 else torch.zeros_like(g0)
            sr = g0 % p1 if p1 > 0 else g0
            s1 = torch.div(sr, p2, rounding_mode="floor") if p2 > 0 else sr
            group = s0 if s.stats_comp == 0 else s1
            if getattr(s, "stats_div", 0) > 1:
                group = torch.div(group, s.stats_div, rounding_mode="floor")
    for nt in range(s.n_tiles):
        acc = torch.zeros(m_tiles, 128, s.block_n, dtype=torch.float32)
        for t in range(s.num_taps):
            for cc in range(s.cchunks):
                coords = torch.stack([torch.full((m_tiles,), s.tap_c0[t] + cc * kc, dtype=torch.int64)] +
                                     [ac[d] + s.tap_off[t][d] for d in range(4)], dim=1)
                A = tma_gather(a_flat, s.a_dims, s.a_strides, s.a_box, coords, esz)  # [T, rows, kc]
                if a_window:
                    # windowed A: the box holds rows + 7 plain pixels of 8 channels; row m reads pixels m .. m+7
                    A = torch.cat([A[:, j:j + a_rows, :] for j in range(8)], dim=2)
                bc = torch.tensor([[s.tap_bk[t] + cc * kc, nt * s.block_n, 0, 0, 0]], dtype=torch.int64)
                B = tma_gather(b_flat, s.b_dims, s.b_strides, s.b_box, bc, esz)[0]  # [block_n, kc]
                acc[:, :a_rows] += A @ B.t()
        if getattr(s, "shift_taps", 0):
            # shift-sum epilogue: out[i][c] = sum_j acc[i + j][j*cpad + c]; the last shift_taps - 1 rows store nothing
            kw, cp = s.shift_taps, s.shift_cpad
            sh = torch.zeros(m_tiles, 128, s.block_n, dtype=torch.float32)
            for j in range(kw):
                sh[:, :128 - j, :cp] += acc[:, j:, j * cp:(j + 1) * cp]
            acc = sh
            valid = valid & (r < a_rows - (kw - 1)).view(1, 128)
        ncols = min(s.n_valid - nt * s.block_n, s.block_n)
        if ncols <= 0:
            continue
        v = acc[:, :, :ncols]
        if bias is not None:
            v = v + bias[nt * s.block_n: nt * s.block_n + ncols].view(1, 1, -1)
        if s.stats_mode:
            vm = torch.where(valid.unsqueeze(-1), v, torch.zeros_like(v))
            s1 = vm.sum(1)
            s2 = (vm * vm).sum(1)
            for ti in range(m_tiles):
                g = int(group[ti])
                stats[g, nt * s.block_n: nt * s.block_n + ncols, 0] += s1[ti]
                stats[g, nt * s.block_n: nt * s.block_n + ncols, 1] += s2[ti]
        v = ACT[s.act](v, s.act_slope)
        if s.out_dtype == 0:
            v = v.to(torch.bfloat16).to(torch.float32)
        ch = (nt * s.block_n + torch.arange(ncols)) * s.out_cstride
        addr = off.unsqueeze(-1) + ch.view(1, 1, -1)
        sel = valid.unsqueeze(-1).expand_as(addr)
        out_flat[addr[sel]] = v[sel]
    return out_flat


def run_wgrad(s, a_flat, b_flat, out_flat):
    """out_flat: float32 [m_valid * ldo] accumulated in place."""
    total_kb = s.t_count[0] * s.t_count[1] * s.t_count[2] * s.t_count[3]
    kb = torch.arange(total_kb)
    dig = _digits(kb, s.t_count)
    ac = _coords(dig, s.a_base, s.a_step)
    bc = _coords(dig, s.b_base, s.b_step)
    tf32 = bool(getattr(s, "tf32", False))
    kc, esz = (32, 4) if tf32 else (64, 2)
    nb = s.block_n // kc
    out = out_flat.view(-1)
    for t in range(s.num_taps):
        for mt in range(s.m_tiles):
            # A: [kb, 64 px, 128 ch]
            As = []
            for j in range(128 // kc):
                coords = torch.stack([torch.full((total_kb,), mt * 128 + j * kc, dtype=torch.int64)] + ac, dim=1)
                As.append(tma_gather(a_flat, s.a_dims, s.a_strides, s.a_box, coords, esz))
            A = torch.cat(As, dim=2)
            for nt in range(s.n_tiles):
                Bs = []
                bdim = getattr(s, "wg_box_dim", 0)
                for j in range(nb):
                    if bdim:   # filter rows in N: box nt*nb + j along tensor dim bdim, same channels
                        coords = torch.stack([torch.full((total_kb,), s.tap_c0[t], dtype=torch.int64)] +
                                             [bc[d] + s.tap_off[t][d] + (nt * nb + j if d + 1 == bdim else 0) for d in range(4)], dim=1)
                    else:
                        coords = torch.stack([torch.full((total_kb,), s.tap_c0[t] + nt * s.block_n + j * kc, dtype=torch.int64)] +
                                             [bc[d] + s.tap_off[t][d] for d in range(4)], dim=1)
                    Bs.append(tma_gather(b_flat, s.b_dims, s.b_strides, s.b_box, coords, esz))
                B = torch.cat(Bs, dim=2)  # [kb, 64, block_n]
                D = torch.einsum("kpm,kpn->mn", A, B)  # [128, block_n]
                rows = min(s.m_valid - mt * 128, 128)
                ncols = min(s.wg_ncols - nt * s.block_n, s.block_n)
                if rows <= 0 or ncols <= 0:
                    continue
                r = (mt * 128 + torch.arange(rows)).view(-1, 1) * s.ldo
                c = (s.tap_bk[t] + nt * s.block_n + torch.arange(ncols)).view(1, -1)
                out[(r + c).reshape(-1)] += D[:rows, :ncols].reshape(-1)
    return out_flat
