"""utils/schedule.py:4-8 of the reference (inactive in the shipped configuration: use_fast_schedule=False)."""
import torch


def get_fast_schedule(origial_timesteps, fast_after_steps, fast_rate):
    if fast_after_steps >= len(origial_timesteps) - 1:
        return origial_timesteps
    return torch.cat((origial_timesteps[:fast_after_steps], origial_timesteps[fast_after_steps + 1::fast_rate]), dim=0)
