"""ORACLE (test infrastructure, not product): CPU restatement of the reference's EDM DPM-Solver++(2M) scheduler.

Follows terrain_diffusion/scheduler/dpmsolver.py @ 82a0431 for the shipped option set (algorithm "dpmsolver++",
solver "midpoint", order 2, karras sigmas, prediction "epsilon", final_sigmas "zero"):

  _compute_karras_sigmas        dpmsolver.py:329-342      set_timesteps                 dpmsolver.py:285-326
  precondition_inputs           dpmsolver.py:226-229      trigflow_precondition_noise   dpmsolver.py:240-242
  precondition_outputs          dpmsolver.py:245-258      convert_model_output          dpmsolver.py:419-452
  first-order update            dpmsolver.py:454-490      second-order multistep update dpmsolver.py:492-561
  step (order schedule, state)  dpmsolver.py:650-726

All scalar arithmetic is done on 0-dim fp32 torch tensors in the reference's operation order so the tables are
fp32-identical.  Parity pinned by tests/test_oracle_golden.py against tables and step sequences recorded from the
unmodified reference (tests/golden/make_golden.py).
"""
from __future__ import annotations

import torch


def karras_sigmas(n: int, sigma_min=0.002, sigma_max=80.0, rho=7.0) -> torch.Tensor:
    ramp = torch.linspace(0, 1, n)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    return (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho


class OracleScheduler:
    def __init__(self, sigma_min=0.002, sigma_max=80.0, sigma_data=0.5, rho=7.0):
        self.sigma_min, self.sigma_max, self.sigma_data, self.rho = sigma_min, sigma_max, sigma_data, rho
        self.num_inference_steps = None

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        s = karras_sigmas(n, self.sigma_min, self.sigma_max, self.rho).to(torch.float32)
        self.timesteps = 0.25 * torch.log(s)
        self.sigmas = torch.cat([s, torch.tensor([0], dtype=torch.float32)])
        self.x0_prev = None
        self.lower_order_nums = 0
        self.step_index = None

    def precondition_inputs(self, sample, sigma):
        return sample * (1 / ((sigma ** 2 + self.sigma_data ** 2) ** 0.5))

    def trigflow_precondition_noise(self, sigma):
        return torch.atan(sigma / self.sigma_data)

    def precondition_outputs(self, sample, model_output, sigma):
        sd = self.sigma_data
        c_skip = sd ** 2 / (sigma ** 2 + sd ** 2)
        c_out = sigma * sd / (sigma ** 2 + sd ** 2) ** 0.5
        return c_skip * sample + c_out * model_output

    def _index_for_timestep(self, timestep):
        cand = (self.timesteps == timestep).nonzero()
        if len(cand) == 0:
            return len(self.timesteps) - 1
        return cand[1].item() if len(cand) > 1 else cand[0].item()

    def step(self, model_output, timestep, sample):
        if self.num_inference_steps is None:
            raise ValueError("set_timesteps was not called")
        if self.step_index is None:
            self.step_index = self._index_for_timestep(timestep)
        i = self.step_index
        n = len(self.timesteps)
        lower_order_final = i == n - 1  # final_sigmas_type == "zero"
        x0 = self.precondition_outputs(sample, model_output, self.sigmas[i])
        x0_prev, self.x0_prev = self.x0_prev, x0
        sigma_t, sigma_s0 = self.sigmas[i + 1], self.sigmas[i]
        lam_t = torch.log(torch.tensor(1)) - torch.log(sigma_t)
        lam_s0 = torch.log(torch.tensor(1)) - torch.log(sigma_s0)
        h = lam_t - lam_s0
        if self.lower_order_nums < 1 or lower_order_final:
            prev = (sigma_t / sigma_s0) * sample - (torch.tensor(1) * (torch.exp(-h) - 1.0)) * x0
        else:
            sigma_s1 = self.sigmas[i - 1]
            lam_s1 = torch.log(torch.tensor(1)) - torch.log(sigma_s1)
            h_0 = lam_s0 - lam_s1
            r0 = h_0 / h
            d0, d1 = x0, (1.0 / r0) * (x0 - x0_prev)
            a = torch.tensor(1) * (torch.exp(-h) - 1.0)
            prev = (sigma_t / sigma_s0) * sample - a * d0 - 0.5 * a * d1
        if self.lower_order_nums < 2:
            self.lower_order_nums += 1
        self.step_index += 1
        return prev


def step_coefficients(sigmas, sigma_data: float = 0.5):
    """Closed form of the update as linear coefficients in fp64:
        x0   = c_skip*x + c_out*F
        x'   = r*x + (1-r)*x0 + k*(x0 - x0_prev)         (k = 0 on the first and last step)
    Returns a list of dicts per step i with c_in, t (noise label), c_skip, c_out, r, k."""
    import math
    s = [float(v) for v in sigmas]
    n = len(s) - 1
    out = []
    for i in range(n):
        si, sn = s[i], s[i + 1]
        den = si * si + sigma_data * sigma_data
        r = sn / si
        k = 0.0
        if 0 < i < n - 1:
            h = math.log(si / sn)
            h0 = math.log(s[i - 1] / si)
            k = 0.5 * (1.0 - r) / (h0 / h)
        out.append(dict(c_in=1.0 / math.sqrt(den), t=math.atan(si / sigma_data), c_skip=sigma_data ** 2 / den,
                        c_out=si * sigma_data / math.sqrt(den), r=r, k=k))
    return out
