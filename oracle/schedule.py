"""Noise schedule + posterior coefficients (oracle, CPU).

Restates /root/reference/models/ligand_diffuser.py:
  * clip_noise_schedule        :620-633
  * polynomial_schedule        :636-650
  * PredefinedNoiseSchedule    :654-690   (gamma table, fp64 numpy -> fp32)
  * sigma / alpha              :232-238
  * sigma_and_alpha_t_given_s  :240-252
  * the per-step coefficient algebra of sample_p_zs_given_zt :505-527

Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np
import torch
import torch.nn.functional as fn


def clip_noise_schedule(alphas2, clip_value=0.001):
    # ligand_diffuser.py:620-633
    alphas2 = np.concatenate([np.ones(1), alphas2], axis=0)
    alphas_step = alphas2[1:] / alphas2[:-1]
    alphas_step = np.clip(alphas_step, a_min=clip_value, a_max=1.0)
    return np.cumprod(alphas_step, axis=0)


def polynomial_schedule(timesteps, s=1e-4, power=3.0):
    # ligand_diffuser.py:636-650
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    alphas2 = (1 - np.power(x / steps, power)) ** 2
    alphas2 = clip_noise_schedule(alphas2, clip_value=0.001)
    precision = 1 - 2 * s
    return precision * alphas2 + s


def gamma_table(timesteps=1000, precision=1e-5, power=2.0):
    """fp32 tensor [timesteps+1]; ligand_diffuser.py:659-686 with 'polynomial_2'."""
    alphas2 = polynomial_schedule(timesteps, s=precision, power=power)
    sigmas2 = 1 - alphas2
    log_a2_to_s2 = np.log(alphas2) - np.log(sigmas2)
    return torch.from_numpy(-log_a2_to_s2).float()


def gamma_lookup(gamma, t, timesteps):
    # ligand_diffuser.py:688-690
    t_int = torch.round(t * timesteps).long()
    return gamma[t_int]


def sigma(gamma):
    return torch.sqrt(torch.sigmoid(gamma))  # :232-234


def alpha(gamma):
    return torch.sqrt(torch.sigmoid(-gamma))  # :236-238


def sigma_and_alpha_t_given_s(gamma_t, gamma_s):
    # :240-252
    sigma2_t_given_s = -torch.expm1(fn.softplus(gamma_s) - fn.softplus(gamma_t))
    log_alpha2_t = fn.logsigmoid(-gamma_t)
    log_alpha2_s = fn.logsigmoid(-gamma_s)
    alpha_t_given_s = torch.exp(0.5 * (log_alpha2_t - log_alpha2_s))
    sigma_t_given_s = torch.sqrt(sigma2_t_given_s)
    return sigma2_t_given_s, sigma_t_given_s, alpha_t_given_s


def posterior_coefficients(gamma, s_int, timesteps):
    """(alpha_t|s, var_terms, sigma) for the reverse step s <- s+1, as fp32 scalars
    computed exactly the way sample_p_zs_given_zt does (:505-527), on 1-element
    fp32 tensors."""
    s = torch.full((1,), float(s_int)) / timesteps
    t = (torch.full((1,), float(s_int)) + 1) / timesteps
    gamma_s = gamma_lookup(gamma, s, timesteps)
    gamma_t = gamma_lookup(gamma, t, timesteps)
    sigma2_ts, sigma_ts, alpha_ts = sigma_and_alpha_t_given_s(gamma_t, gamma_s)
    sigma_s = sigma(gamma_s)
    sigma_t = sigma(gamma_t)
    var_terms = sigma2_ts / alpha_ts / sigma_t
    sig = sigma_ts * sigma_s / sigma_t
    return alpha_ts, var_terms, sig


def coefficient_table(gamma, timesteps):
    """fp32 [timesteps, 4]: row s = (alpha_t|s, var_terms, sigma, t=(s+1)/T).
    The t column is what the reference feeds the denoiser (:405-408, :513)."""
    rows = []
    for s_int in range(timesteps):
        a, v, sg = posterior_coefficients(gamma, s_int, timesteps)
        s_arr = torch.full((1,), s_int)
        t_arr = (s_arr + 1) / timesteps  # int64 tensor / int -> fp32, as in :405-408
        rows.append(torch.stack([a[0], v[0], sg[0], t_arr[0].float()]))
    return torch.stack(rows)
