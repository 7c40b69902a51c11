"""Mirror of terrain_diffusion.scheduler (reference: terrain_diffusion/scheduler/)."""
from .dpmsolver import EDMDPMSolverMultistepScheduler, SchedulerOutput  # noqa: F401
