"""Property tests (hypothesis) of the host-side integer logic of the hot path: tile starts, window index ranges, batch
padding, rank partitions, phase splits.  Bit-exact domains: every property is an equality or an ordering."""
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import tiling as otile
from terrain_diffusion_b200.inference.multiphase import build_timestep_ranges, phase_step_ranges
from terrain_diffusion_b200.inference.tiling import padded_batch_size, shard_rows, tile_starts, window_range
from terrain_diffusion_b200.scheduler import EDMDPMSolverMultistepScheduler

FAST = settings(max_examples=300, deadline=None)


@FAST
@given(st.integers(1, 20000), st.integers(1, 1024), st.integers(1, 1024))
def test_tile_starts_cover_the_canvas_and_match_the_oracle(length, tile, stride):
    s = tile_starts(length, tile, stride)
    assert s == otile.tile_starts(length, tile, stride)            # training/evaluation/__init__.py:16-22
    assert s[0] == 0 and s == sorted(set(s))
    if length > tile and stride <= tile:
        assert s[-1] == length - tile                              # the last tile is clamped to the edge
        covered = 0
        for a in s:
            assert a <= covered                                    # no gap between consecutive tiles
            covered = max(covered, a + tile)
        assert covered == length
    if length <= tile:
        assert s == [0]


@FAST
@given(st.integers(-5000, 5000), st.integers(1, 3000), st.integers(1, 600), st.integers(1, 600), st.integers(-300, 300))
def test_window_range_is_exactly_the_windows_that_intersect(a, span, size, stride, offset):
    b = a + span
    ks = window_range(a, b, size, stride, offset)

    def hits(k):
        lo = k * stride + offset
        return lo < b and lo + size > a
    assert all(hits(k) for k in ks)
    assert not hits(ks.start - 1) and not hits(ks.stop)
    assert list(ks) == list(otile.window_range(a, b, size, stride, offset))


@FAST
@given(st.integers(1, 64), st.sampled_from([1, 2, 4, 8, 16, 32]))
def test_padded_batch_size_is_the_next_power_of_two_capped(n, cap):
    p = padded_batch_size(n, cap)
    assert p <= cap and (p & (p - 1)) == 0
    assert p >= min(n, cap) and (p == cap or p < 2 * n)


@FAST
@given(st.integers(1, 200), st.integers(1, 16))
def test_shard_rows_is_a_contiguous_balanced_partition(n_rows, world):
    parts = [shard_rows(n_rows, world, r) for r in range(world)]
    flat = [i for p in parts for i in p]
    assert flat == list(range(n_rows))
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@FAST
@given(st.integers(2, 40), st.lists(st.floats(-2.0, 1.5, allow_nan=False), min_size=0, max_size=4))
def test_phase_split_is_an_ordered_partition_of_the_schedule(num_steps, thresholds):
    sched = EDMDPMSolverMultistepScheduler()
    sched.set_timesteps(num_steps)
    ts = sched.timesteps
    ranges = build_timestep_ranges(ts, thresholds)
    assert torch.equal(torch.cat(ranges), ts) and all(len(r) > 0 for r in ranges)
    th = sorted(thresholds, reverse=True)
    for r in ranges:                                                # no range straddles a threshold
        for t in th:
            assert bool((r >= t).all()) or bool((r < t).all())
    spans = phase_step_ranges(sched, num_steps, thresholds)
    assert spans[0][0] == 0 and spans[-1][1] == num_steps
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


@FAST
@given(st.integers(2, 60), st.floats(0.001, 0.05), st.floats(5.0, 200.0), st.floats(0.25, 1.0))
def test_scheduler_tables_and_step_coefficients_match_the_oracle_closed_form(n, smin, smax, sdata):
    """Sigma table (Karras, rho = 7; dpmsolver.py:329-342) and the per-step update coefficients (Appendix B) of the
    product scheduler against the oracle's fp64 closed form, for arbitrary step counts and sigma ranges."""
    from oracle import scheduler as osched
    s = EDMDPMSolverMultistepScheduler(sigma_min=smin, sigma_max=smax, sigma_data=sdata)
    s.set_timesteps(n)
    want = osched.karras_sigmas(n, smin, smax)
    assert torch.allclose(s.sigmas[:-1].double(), want.double()[:n], rtol=2e-6, atol=0) and float(s.sigmas[-1]) == 0.0
    ref = osched.step_coefficients(s.sigmas.double(), sdata)
    order = s.order_schedule()
    assert order[0] is False and order[-1] is False and all(order[1:-1])      # first / last step first order
    for i in (0, n // 2, n - 1):
        co = s.step_coefficients(i, order[i])
        for key in ("c_in", "t", "c_skip", "c_out", "r", "k"):
            assert abs(float(co[key]) - ref[i][key]) <= 2e-6 * max(1.0, abs(ref[i][key])), (i, key)
