"""The algebra behind the grouped taps-in-N kernel (nsc_b200/csrc/plane_conv.cu: tgroup_shift / tgroup_tap, plane_t_kernel<20, 5, 3>
and <20, 3, 5>), restated in numpy -- no GPU.

The kernel computes, per tap group g, P_g[r, slot] = X[r + a_g] . W[tap(g, slot)] for the rows r of the frame only (the descriptor row
shift a_g reads the image's zero rows beyond the frame), and the epilogue sums y[r] = sum_slot P[r + e_slot, slot] across TMEM lanes,
again over rows of the frame only.  That equals the SAME-padded conv (nn_core_operator.py:6-14) exactly iff a slot shift never
points outside the frame while the tap's input row is inside it -- which is what the one-sided outer groups guarantee."""
import numpy as np
import pytest


def tgroup_shift(groups, g):           # in units of the dilation (plane_conv.cu: tgroup_shift)
    return 2 * (g - 1) if groups == 3 else (-3, -1, 0, 1, 3)[g]


def tgroup_tap(groups, g, slot):       # -1: slot unused by this group (plane_conv.cu: tgroup_tap)
    if groups == 3:
        return slot + 2 * g if g <= slot <= g + 2 else -1
    if g < 2:
        return 2 * g + slot if slot <= 1 else -1
    if g == 2:
        return 4 if slot == 1 else -1
    return 2 * g - 2 + slot if slot >= 1 else -1


def conv_same(x, w, d):
    L, K = x.shape[0], w.shape[0]
    y = np.zeros((L, w.shape[2]))
    for t in range(K):
        s = (t - K // 2) * d
        for r in range(L):
            if 0 <= r + s < L:
                y[r] += x[r + s] @ w[t]
    return y


def grouped(x, w, d, groups, tap_of=tgroup_tap, shift_of=tgroup_shift):
    L = x.shape[0]
    slots = 5 if groups == 3 else 3
    xp = np.zeros((L + 16, x.shape[1]))
    xp[8:8 + L] = x                                   # the plane image: 8 zero rows above and below
    P = np.zeros((L, slots, w.shape[2]))
    for g in range(groups):
        a = shift_of(groups, g) * d
        assert abs(a) <= 8                            # the shift stays inside the image's halo rows
        for slot in range(slots):
            t = tap_of(groups, g, slot)
            if t >= 0:
                P[:, slot] += xp[8 + a:8 + a + L] @ w[t]
    y = np.zeros((L, w.shape[2]))
    for slot in range(slots):
        e = (slot - slots // 2) * d
        for r in range(L):
            if 0 <= r + e < L:                        # rows outside the frame are never computed
                y[r] += P[r + e, slot]
    return y


@pytest.mark.parametrize('groups', [3, 5])
@pytest.mark.parametrize('d', [1, 2])
def test_every_tap_lands_once_with_its_shift(groups, d):
    seen = {}
    slots = 5 if groups == 3 else 3
    for g in range(groups):
        for slot in range(slots):
            t = tgroup_tap(groups, g, slot)
            if t >= 0:
                assert t not in seen
                seen[t] = (tgroup_shift(groups, g) + slot - slots // 2) * d
    assert seen == {t: (t - 4) * d for t in range(9)}


@pytest.mark.parametrize('groups', [3, 5])
@pytest.mark.parametrize('d', [1, 2])
def test_grouped_form_is_the_same_padded_conv_at_frame_borders(groups, d):
    rng = np.random.RandomState(groups * 10 + d)
    x = rng.randn(128, 20)
    w = rng.randn(9, 20, 20)
    np.testing.assert_allclose(grouped(x, w, d, groups), conv_same(x, w, d), rtol=0, atol=1e-10)


def test_two_sided_outer_groups_would_lose_border_contributions():
    """Negative control: three groups shifted by -3d, 0, +3d with slots {-d, 0, +d} place every tap correctly in the interior but
    drop, e.g., tap 6 (shift +2d) on the first d rows: its slot shift -d points outside the frame while x[r + 2d] is inside."""
    d = 1
    rng = np.random.RandomState(0)
    x, w = rng.randn(128, 20), rng.randn(9, 20, 20)

    def shift_of(groups, g):
        return 3 * (g - 1)

    def tap_of(groups, g, slot):
        return 3 * g + slot - 1 if 1 <= slot <= 3 else -1     # 5-slot frame, only the middle three used

    got, want = grouped(x, w, d, 3, tap_of, shift_of), conv_same(x, w, d)
    np.testing.assert_allclose(got[8:-8], want[8:-8], rtol=0, atol=1e-10)     # interior rows agree
    assert np.abs(got[0] - want[0]).max() > 1e-3 and np.abs(got[-1] - want[-1]).max() > 1e-3
