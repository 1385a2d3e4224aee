"""Host-side DDIM schedule helpers (numpy, float64 -> float32), counterparts of
``make_ddim_timesteps`` / ``make_ddim_sampling_parameters`` in
ldm/modules/diffusionmodules/util.py:46-74.  Pure host code: runs once per ``sample()``.
"""
from __future__ import annotations

import numpy as np


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    if ddim_discr_method == 'uniform':
        stride = num_ddpm_timesteps // num_ddim_timesteps
        steps = np.arange(0, num_ddpm_timesteps, stride)
    elif ddim_discr_method == 'quad':
        steps = (np.linspace(0, np.sqrt(num_ddpm_timesteps * .8), num_ddim_timesteps) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    # +1: the final alpha values must be the ones from the first scale-to-data step
    steps_out = steps + 1
    if verbose:
        print(f'Selected timesteps for ddim sampler: {steps_out}')
    return steps_out


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    """Returns (sigmas, alphas, alphas_prev) with the reference's container types: ``alphacums`` is a CPU
    torch tensor, so ``alphas``/``sigmas`` come back as torch fp32 and ``alphas_prev`` as numpy."""
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    if verbose:
        print(f'Selected alphas for ddim sampler: a_t: {alphas}; a_(t-1): {alphas_prev}')
        print(f'For the chosen value of eta, which is {eta}, '
              f'this results in the following sigma_t schedule for ddim sampler {sigmas}')
    return sigmas, alphas, alphas_prev
