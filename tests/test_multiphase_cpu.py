"""SURVEY 8(a) row a17 -- host side: the phase split of the multi-phase InfiniteDiffusion sampler is the reference's
`build_timestep_ranges` (golden vectors from the function itself, tests/golden/make_golden_phases.py), for the product
code and for the oracle restatement; sub-solve coefficient tables reset the multistep history at a phase start."""
from pathlib import Path

import numpy as np
import torch

from oracle import tiling as otile
from terrain_diffusion_b200.inference.multiphase import build_timestep_ranges, phase_step_ranges
from terrain_diffusion_b200.scheduler import EDMDPMSolverMultistepScheduler

G = np.load(Path(__file__).resolve().parent / "golden" / "phases_golden.npz")


def test_build_timestep_ranges_matches_the_reference_function():
    for name in ("demo", "unsorted", "none", "empty_phase", "edm12"):
        ts = torch.from_numpy(G["edm12_timesteps"] if name == "edm12" else G["demo.timesteps"])
        th = tuple(float(v) for v in G[f"{name}.thresholds"])
        for fn in (build_timestep_ranges, otile.build_timestep_ranges):
            r = fn(ts, th)
            assert [len(x) for x in r] == [int(v) for v in G[f"{name}.lens"]], (name, fn.__module__)
            assert np.array_equal(torch.cat([x.float() for x in r]).numpy(), G[f"{name}.concat"])


def test_phase_step_ranges_cover_the_schedule_and_match_the_golden_split():
    s = EDMDPMSolverMultistepScheduler(sigma_min=0.002, sigma_max=80, sigma_data=0.5)
    th = tuple(float(v) for v in G["edm12.thresholds"])
    ranges = phase_step_ranges(s, 12, th)
    assert ranges == [(0, 5), (5, 8), (8, 12)]
    assert np.array_equal(s.timesteps.numpy(), G["edm12_timesteps"])          # same table the golden was split on
    assert phase_step_ranges(s, 12, ()) == [(0, 12)]


def test_sub_solve_coefficients_drop_the_second_order_term_at_a_phase_start():
    s = EDMDPMSolverMultistepScheduler(sigma_min=0.002, sigma_max=80, sigma_data=0.5)
    s.set_timesteps(12)
    order = s.order_schedule()
    assert order[5] and order[6]                     # mid-schedule steps are second order in a full solve ...
    full = s.step_coefficients(5, order[5])
    reset = s.step_coefficients(5, False)            # ... but the first step after a blend has no x0 history
    assert full["k"] != 0.0 and reset["k"] == 0.0
    assert all(full[k] == reset[k] for k in ("c_in", "t", "c_skip", "c_out", "r"))
