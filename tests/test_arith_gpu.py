"""The hand-written IEEE sequences of the packed kernels (stonefish_b200/csrc/sfh_packed.cuh) against the hardware's
correctly rounded instructions, run on the device by sfh_selftest (include/sfh.h)."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


def _run(ctx, which, n=0):
    tested, bad = C.c_uint64(), C.c_uint64()
    ctx._check(ctx.lib.sfh_selftest(ctx.ctx, which, n, C.byref(tested), C.byref(bad)), "sfh_selftest")
    if bad.value:
        print("selftest detail:", ctx.lib.sfh_last_error(ctx.ctx).decode())
    return tested.value, bad.value


def test_sqrt_sequence_exhaustive(ctx):
    tested, bad = _run(ctx, 0)
    assert tested == 0x7f7fffff - 0x0d000000 + 1  # every float in [2^-101, FLT_MAX]
    assert bad == 0


def test_div3_sequence_exhaustive(ctx):
    tested, bad = _run(ctx, 1)
    assert tested == 2 * (0x7f800000 - 0x0d800000) + 2  # +-[2^-100, FLT_MAX] and +-0
    assert bad == 0


def test_shared_reciprocal_division_random(ctx):
    tested, bad = _run(ctx, 2, 1 << 33)
    assert tested == 1 << 33
    assert bad == 0


def test_packed_operators_equal_scalar_unfused(ctx):
    tested, bad = _run(ctx, 3, 1 << 28)
    assert tested == 1 << 28
    assert bad == 0


def test_shared_reciprocal_division_wide_quotients(ctx):
    """ieee_div_r1 as the wave lookup uses it: world coordinate / grid size, magnitudes unrelated."""
    tested, bad = _run(ctx, 4, 1 << 31)
    assert tested == 1 << 31 and bad == 0
