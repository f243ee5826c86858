"""Seeded sweep of ragged geometries through the planner + igemm emulator (CPU): every layer family of the path
(models/networks.py:578-605, 621-648, 747-775; models/resnet.py:20-28, 134) at odd / non-square / tiny / wider-than-a-tile
sizes and batch sizes, forward, data gradient and weight gradient, against torch.nn.functional on bf16-rounded operands.
The fixed cases of tests/test_plan_cpu.py pin the shapes of the benchmark; this pins the edges around them."""
import random

import pytest

from tests import test_plan_cpu as T

_rng = random.Random(20251018)


def _even(v):
    return v + (v & 1)


# family: (cin, cin_buf, cout, k, stride, cp, halo, xpad, even sizes?, (lo, hi) of H and W)
FWD_FAMILIES = {
    "res64": (64, 64, 64, 3, 1, 1, "reflect", 1, False, (2, 21)),
    "res128_n80": (128, 128, 80, 3, 1, 1, "reflect", 1, False, (2, 13)),
    "down_s2": (64, 64, 128, 3, 2, 1, "zero", 1, True, (4, 22)),
    "d4x4_s2": (64, 64, 128, 4, 2, 1, "zero", 1, True, (4, 22)),
    "d4x4_s1": (128, 128, 64, 4, 1, 1, "zero", 1, False, (4, 13)),
    "stem_overlap": (4, 8, 64, 7, 1, 3, "reflect", 3, False, (4, 40)),
    "stem_window": (4, 8, 64, 7, 1, 3, "reflect", 3, False, (64, 140)),
    "dstem_s2": (4, 8, 64, 4, 2, 1, "zero", 1, True, (4, 30)),
    "estem_s2": (3, 8, 64, 7, 2, 3, "zero", 3, True, (8, 40)),
    "head_n3": (64, 64, 3, 7, 1, 3, "reflect", 3, False, (4, 30)),
    "e1x1_s2": (64, 64, 128, 1, 2, 0, "zero", 1, True, (2, 16)),
    "tail_c32": (32, 32, 1, 3, 1, 1, "zero", 1, False, (3, 12)),
}


def _fwd_cases():
    out = []
    for name, (cin, cbuf, cout, k, s, cp, halo, xpad, even, (lo, hi)) in FWD_FAMILIES.items():
        for i in range(2):
            w = _rng.randint(lo, hi)
            h = _rng.randint(4, 7) if name == "stem_window" else _rng.randint(lo, min(hi, 24))   # the emulator is O(pixels * K)
            if even:
                h, w = _even(h), _even(w)
            n = _rng.randint(1, 3)
            out.append(("%s_%dx%dx%d" % (name, n, h, w), cin, cbuf, cout, k, s, cp, halo, xpad, h, w, n))
    return out


FWD_SWEEP = _fwd_cases()


@pytest.mark.parametrize("case", FWD_SWEEP, ids=[c[0] for c in FWD_SWEEP])
def test_conv_forward_plan_sweep(case):
    T.test_conv_forward_plan(case)


# family: (cin, cout, cout_buf, k, stride, cp, xpad, full_padded, dypad, even?, (lo, hi) of the square size)
DGRAD_FAMILIES = {
    "res_flat_full": (64, 64, 64, 3, 1, 1, 1, True, 1, False, (2, 19)),
    "res_flat_interior": (64, 128, 128, 3, 1, 1, 1, False, 1, False, (2, 13)),
    "head_packed_full": (64, 3, 8, 7, 1, 3, 3, True, 6, False, (4, 24)),
    "head_window_full": (64, 3, 8, 7, 1, 3, 3, True, 6, False, (58, 70)),
    "dhead_packed": (64, 1, 8, 4, 1, 1, 1, False, 2, False, (4, 14)),
    "d4x4_s1_box": (64, 64, 64, 4, 1, 1, 1, False, 0, False, (4, 12)),
    "down_s2": (64, 128, 128, 3, 2, 1, 1, False, 0, True, (4, 20)),
    "d4x4_s2": (64, 128, 128, 4, 2, 1, 1, False, 1, True, (4, 20)),
    "dstem_s2_to4": (4, 64, 64, 4, 2, 1, 1, False, 0, True, (4, 20)),
    "estem_s2_to3_shift": (3, 64, 64, 7, 2, 3, 3, False, 2, True, (8, 36)),
    "stem_to4_full": (4, 64, 64, 7, 1, 3, 3, True, 0, False, (4, 20)),
}


def _dgrad_cases():
    out = []
    for name, (cin, cout, cobuf, k, s, cp, xpad, full, dypad, even, (lo, hi)) in DGRAD_FAMILIES.items():
        h = _rng.randint(lo, hi)
        if even:
            h = _even(h)
        n = 1 if h > 40 else _rng.randint(1, 3)
        out.append(("%s_%dx%d" % (name, n, h), cin, cout, cobuf, k, s, cp, xpad, full, h, n, dypad))
    return out


DGRAD_SWEEP = _dgrad_cases()


@pytest.mark.parametrize("case", DGRAD_SWEEP, ids=[c[0] for c in DGRAD_SWEEP])
def test_conv_dgrad_plan_sweep(case):
    T.test_conv_dgrad_plan(case)


# family: (cin, cin_buf, cout, cout_buf, k, stride, cp, halo, xpad, dypad, even?, (lo, hi))
WGRAD_FAMILIES = {
    "res": (64, 64, 64, 64, 3, 1, 1, "reflect", 1, 1, False, (2, 18)),
    "res_c256": (256, 256, 128, 128, 3, 1, 1, "reflect", 1, 0, False, (2, 9)),
    "down_s2": (64, 64, 192, 192, 3, 2, 1, "zero", 1, 0, True, (4, 18)),
    "d4x4_s2": (64, 64, 128, 128, 4, 2, 1, "zero", 1, 1, True, (4, 18)),
    "d4x4_s1_odd": (64, 64, 64, 64, 4, 1, 1, "zero", 1, 1, False, (4, 11)),
    "stem_packed": (4, 8, 64, 64, 7, 1, 3, "reflect", 3, 0, False, (4, 26)),      # filter rows in N
    "dstem_s2_packed": (4, 8, 64, 64, 4, 2, 1, "zero", 1, 0, True, (4, 18)),
    "head_cout3": (64, 64, 3, 8, 7, 1, 3, "reflect", 3, 6, False, (4, 20)),       # swapped operands, filter rows in N
    "dhead_cout1": (64, 64, 1, 8, 4, 1, 1, "zero", 1, 2, False, (4, 12)),
}


def _wgrad_cases():
    out = []
    for name, (cin, cbuf, cout, cobuf, k, s, cp, halo, xpad, dypad, even, (lo, hi)) in WGRAD_FAMILIES.items():
        h = _rng.randint(lo, hi)
        if even:
            h = _even(h)
        n = _rng.randint(1, 3)
        out.append(("%s_%dx%d" % (name, n, h), cin, cbuf, cout, cobuf, k, s, cp, halo, xpad, h, n, dypad))
    return out


WGRAD_SWEEP = _wgrad_cases()


@pytest.mark.parametrize("case", WGRAD_SWEEP, ids=[c[0] for c in WGRAD_SWEEP])
def test_conv_wgrad_plan_sweep(case):
    T.test_conv_wgrad_plan(case)
