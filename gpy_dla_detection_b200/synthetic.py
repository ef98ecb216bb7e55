"""
synthetic.py : seeded synthetic inputs of the published shapes.

The reference's learned model (`learned_qso_model_*.mat`), its QMC sample files
(`dla_samples_a03.mat`, `subdla_samples.mat`), its prior catalogue (`catalog.mat`) and the
SDSS FITS spectra are not available offline, so every test, the smoke run and `bench.py`
draw their inputs from here.  Shapes follow the reference:

* learned model : rest grid 911.75:0.25:1215.75 (1217 points), `mu`, `M` (1217 x 20),
  `log_omega`, `log_c_0`, `log_tau_0`, `log_beta`            (null_gp.py:390-422)
* DLA samples   : `offset_samples`, `log_nhi_samples`, `nhi_samples`  (dla_samples.py:67-77;
  generate_dla_samples.m:8-57 - Halton sequence + inverse-CDF of the log N_HI mixture)
* subDLA samples: log N_HI ~ U(19.5, 20) on the same offsets, `Z_lls`, `Z_dla`
  (subdla_samples.py:82-125; multi_dlas/set_lls_parameters.m:58-71)
* prior         : duck-typed `less_ind(z_qso)`                (model_priors.py:142-157)
* spectra       : BOSS-like log-lambda grid, (wavelengths, flux, noise_variance, pixel_mask)
  as `read_spec.read_spec` returns them                      (read_spec.py:22-71)

This module is data synthesis only (host NumPy); it is not on the measured path.
"""
from typing import Dict, Tuple

import numpy as np
from scipy.special import wofz

from .set_parameters import Parameters
from . import _tables as tables


# ----------------------------------------------------------------------------------------
# learned GP model
# ----------------------------------------------------------------------------------------
def make_learned_model(seed: int = 0, k: int = 20, rest_min: float = 911.75, rest_max: float = 1215.75) -> Dict[str, np.ndarray]:
    """
    Synthetic stand-in for the learned null-model file: grid rest_min:0.25:rest_max.  The defaults are
    the published DLA model's range; examples/gp_find_lls.py:102 trains on 850.75-1420.75 A.
    """
    rng = np.random.default_rng(seed)
    rest_wavelengths = rest_min + 0.25 * np.arange(int(round((rest_max - rest_min) / 0.25)) + 1)

    def bump(center, width, height):
        return height * np.exp(-0.5 * ((rest_wavelengths - center) / width) ** 2)

    mu = (
        1.0
        + 0.15 * (rest_wavelengths - 911.75) / 304.0
        + bump(1215.67, 9.0, 1.9)
        + bump(1025.72, 5.0, 0.35)
        + bump(1033.8, 4.0, 0.25)
        + bump(977.0, 3.0, 0.12)
    )

    # Gaussian-smoothed white noise columns, decaying scale
    N = rest_wavelengths.shape[0]
    kern_x = np.arange(-80, 81)
    kern = np.exp(-0.5 * (kern_x / 20.0) ** 2)
    kern /= np.sqrt(np.sum(kern**2))
    M = np.empty((N, k))
    for j in range(k):
        white = rng.standard_normal(N + kern_x.shape[0] - 1)
        M[:, j] = np.convolve(white, kern, "valid") * 0.2 * (0.85**j)

    t = (rest_wavelengths - 911.75) / 304.0
    log_omega = np.log(0.03 + 0.27 * (0.5 + 0.5 * np.cos(2.2 * np.pi * t + 0.3)) * (1 - 0.5 * t))

    return dict(
        rest_wavelengths=rest_wavelengths,
        mu=mu,
        M=np.ascontiguousarray(M),
        log_omega=log_omega,
        log_c_0=float(np.log(0.1)),
        log_tau_0=float(np.log(0.0023)),
        log_beta=float(np.log(3.65)),
    )


# ----------------------------------------------------------------------------------------
# QMC samples
# ----------------------------------------------------------------------------------------
def halton(n: int, base: int, skip: int = 1) -> np.ndarray:
    """First n points (after `skip`) of the van der Corput sequence in `base`."""
    out = np.zeros(n)
    idx = np.arange(skip, skip + n, dtype=np.int64)
    f = 1.0
    i = idx.copy()
    while np.any(i > 0):
        f = f / base
        out += f * (i % base)
        i //= base
    return out


def _mixture_inverse_cdf(u: np.ndarray, params: Parameters) -> np.ndarray:
    """Inverse CDF of the log N_HI mixture prior of dla_samples.py:106-125."""
    grid = np.linspace(params.fit_min_log_nhi, 25.0, 200001)
    unnorm = np.exp(-1.2695 * grid**2 + 50.863 * grid - 509.33)
    Z = np.trapezoid(unnorm, grid)
    uni = ((grid >= params.uniform_min_log_nhi) & (grid <= params.uniform_max_log_nhi)) / (
        params.uniform_max_log_nhi - params.uniform_min_log_nhi
    )
    pdf = params.alpha * unnorm / Z + (1 - params.alpha) * uni
    cdf = np.concatenate([[0.0], np.cumsum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(grid))])
    cdf /= cdf[-1]
    return np.interp(u, cdf, grid)


def make_dla_sample_arrays(params: Parameters, num_samples: int = None) -> Dict[str, np.ndarray]:
    S = params.num_dla_samples if num_samples is None else num_samples
    offset = halton(S, 2)
    log_nhi = _mixture_inverse_cdf(halton(S, 3), params)
    return dict(offset_samples=offset, log_nhi_samples=log_nhi, nhi_samples=10.0**log_nhi)


def make_subdla_sample_arrays(params: Parameters, num_samples: int = None) -> Dict[str, np.ndarray]:
    S = params.num_dla_samples if num_samples is None else num_samples
    offset = halton(S, 2)
    lo, hi = 19.5, 20.0
    log_nhi = lo + (hi - lo) * halton(S, 3)
    # partition functions (multi_dlas/set_lls_parameters.m:58-71): the DLA prior's peak value is
    # extrapolated uniformly down to `lo`
    grid = np.linspace(params.fit_min_log_nhi, 25.0, 200001)
    unnorm = np.exp(-1.2695 * grid**2 + 50.863 * grid - 509.33)
    Z_dla = float(np.trapezoid(unnorm, grid))
    Z_lls = float(unnorm[0] * (hi - lo))
    return dict(
        offset_samples=offset,
        log_nhi_samples=log_nhi,
        nhi_samples=10.0**log_nhi,
        Z_dla=Z_dla,
        Z_lls=Z_lls,
        extrapolate_min_log_nhi=lo,
    )


def make_lls_sample_arrays(num_samples: int, min_log_nhi: float = 17.0, max_log_nhi: float = 21.0) -> Dict[str, np.ndarray]:
    """Absorber samples for the Lyman-limit-system variant (examples/gp_find_lls.py:227-294): uniform in log N_HI."""
    offsets = halton(num_samples, 2)
    log_nhi = min_log_nhi + (max_log_nhi - min_log_nhi) * halton(num_samples, 3)
    return dict(offset_samples=offsets, log_nhi_samples=log_nhi, nhi_samples=10.0**log_nhi)


class SyntheticPrior:
    """
    Duck-typed stand-in for model_priors.PriorCatalog (only `less_ind` is on the path,
    model_priors.py:142-157): a synthetic (z_qsos, dla_ind) catalogue with p(DLA) ~ 0.1.
    """

    def __init__(self, params: Parameters, num_quasars: int = 50000, seed: int = 1):
        rng = np.random.default_rng(seed)
        self.params = params
        self.z_qsos = np.sort(2.15 + rng.gamma(2.0, 0.3, size=num_quasars))
        p = 0.04 + 0.05 * (self.z_qsos - 2.15)
        self.dla_ind = rng.random(num_quasars) < np.clip(p, 0.0, 0.4)

    def less_ind(self, z_qso: float) -> Tuple[float, float]:
        less_ind = self.z_qsos < (z_qso + self.params.prior_z_qso_increase)
        return np.sum(self.dla_ind[less_ind]), np.sum(less_ind)


# ----------------------------------------------------------------------------------------
# spectra
# ----------------------------------------------------------------------------------------
def _host_voigt_absorption(wavelengths, nhi, z_dla, num_lines=3):
    """scipy-based unbroadened absorption profile, for data synthesis only."""
    c = tables.SPEED_OF_LIGHT_CGS
    tot = np.zeros_like(wavelengths)
    for l in range(num_lines):
        vel = wavelengths * (c / (tables.TRANSITION_WAVELENGTHS[l] * (1 + z_dla)) / 1e8) - c
        zz = (vel + 1j * tables.GAMMAS[l]) / (np.sqrt(2) * tables.SIGMA)
        tot += -tables.LEADING_CONSTANTS[l] * np.real(wofz(zz)) / (np.sqrt(2 * np.pi) * tables.SIGMA)
    return np.exp(nhi * tot)


def sample_z_qsos(num: int, seed: int = 12345) -> np.ndarray:
    """DR12Q-like quasar redshift distribution: 2.15 + Gamma tail, clipped at 5."""
    rng = np.random.default_rng(seed)
    return np.minimum(2.15 + rng.gamma(2.0, 0.25, size=num), 5.0)


def make_spectrum(
    model: Dict[str, np.ndarray],
    z_qso: float,
    seed: int,
    num_pixels: int = 4650,
    params: Parameters = None,
) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """
    One BOSS-like spectrum: (wavelengths, flux, noise_variance, pixel_mask) with the
    meaning of read_spec.read_spec's return values (read_spec.py:22-71).
    """
    if params is None:
        params = Parameters()
    rng = np.random.default_rng(seed)

    loglam = 3.5523 + 1e-4 * np.arange(num_pixels)
    wavelengths = 10.0**loglam
    rest = wavelengths / (1 + z_qso)

    # continuum from the model inside its range; flat outside
    rw = model["rest_wavelengths"]
    inside = (rest >= rw[0]) & (rest <= rw[-1])
    cont = np.ones(num_pixels)
    cont[inside] = np.interp(rest[inside], rw, model["mu"])
    cont[rest < rw[0]] = 0.3
    # red side: the Ly-alpha emission wing decays to a unit continuum, so the 1310-1325 A
    # normalisation window sits at 1 like the (normalised) training spectra of the learned model
    mu_end = np.interp(rw[-1], rw, model["mu"])
    cont[rest > rw[-1]] = 1.0 + (mu_end - 1.0) * np.exp(-(rest[rest > rw[-1]] - rw[-1]) / 15.0)

    # GP draw on the modelled range
    xi = rng.standard_normal(model["M"].shape[1])
    gp_draw = np.zeros(num_pixels)
    for j in range(model["M"].shape[1]):
        gp_draw[inside] += np.interp(rest[inside], rw, model["M"][:, j]) * xi[j]

    # mean-flux suppression by the Lyman-series forest (Kim et al. parameters), the same functional
    # form the null model applies to its mean (effective_optical_depth.py:51-78)
    lya = params.lya_wavelength
    tw = tables.TRANSITION_WAVELENGTHS * 1e8
    tau_eff = np.zeros(num_pixels)
    for i in range(params.num_forest_lines):
        z_i = (wavelengths - tw[i]) / tw[i]
        tau_i = 0.0023 * tables.OSCILLATOR_STRENGTHS[i] / tables.OSCILLATOR_STRENGTHS[0] * tw[i] / tw[0]
        tau_eff += np.where(z_i <= z_qso, tau_i * np.maximum(1 + z_i, 0.0) ** 3.65, 0.0)
    suppress = np.exp(-tau_eff)

    # injected DLAs
    absorption = np.ones(num_pixels)
    num_dlas = rng.choice([0, 0, 0, 1, 1, 2, 3])
    z_lo = max(wavelengths[0] / lya - 1, (1 + z_qso) * params.lyman_limit / lya - 1) + 0.01
    z_hi = z_qso - 0.02
    dlas = []
    for _ in range(num_dlas):
        if z_hi <= z_lo:
            break
        zd = rng.uniform(z_lo, z_hi)
        ln = rng.uniform(20.0, 22.5)
        dlas.append((zd, ln))
        absorption *= _host_voigt_absorption(wavelengths, 10.0**ln, zd, 3)

    omega = np.zeros(num_pixels)
    omega[inside] = np.exp(np.interp(rest[inside], rw, model["log_omega"]))
    forest_noise = omega * rng.standard_normal(num_pixels) * (1 - np.exp(-tau_eff) + 0.1) * (rest < lya)

    clean = (cont + gp_draw + forest_noise) * suppress * absorption

    sigma_pix = np.exp(rng.normal(np.log(0.25), 0.35, size=num_pixels)) * rng.uniform(0.5, 1.6)
    noise_variance = sigma_pix**2
    normaliser = np.exp(rng.normal(np.log(4.0), 0.6))

    flux = (clean + sigma_pix * rng.standard_normal(num_pixels)) * normaliser
    noise_variance = noise_variance * normaliser**2

    # pixel mask: ivar == 0 (variance = inf, read_spec.py:59-63) or BRIGHTSKY (finite variance)
    ivar0 = rng.random(num_pixels) < 0.012
    bright = rng.random(num_pixels) < 0.004
    # a short run of consecutive bad pixels now and then (sky-line residuals)
    if rng.random() < 0.5:
        start = rng.integers(0, num_pixels - 8)
        ivar0[start : start + rng.integers(2, 8)] = True
    with np.errstate(divide="ignore"):
        noise_variance = np.where(ivar0, np.inf, noise_variance)
    flux = np.where(ivar0, 0.0, flux)
    pixel_mask = ivar0 | bright

    return wavelengths, flux, noise_variance, pixel_mask


# ----------------------------------------------------------------------------------------
# zQSO estimation (ZGP, BASELINE.json configs[4]): learned model over 910-3000 A and spectra
# that extend redwards of Ly-alpha with the usual broad emission lines
# ----------------------------------------------------------------------------------------
_EMISSION_LINES = (  # (rest wavelength, width, height): Ly-b/OVI, Ly-a, NV, SiIV, CIV, CIII], MgII
    (1033.0, 8.0, 0.35), (1215.67, 10.0, 2.2), (1240.0, 8.0, 0.5), (1398.0, 12.0, 0.35),
    (1549.0, 14.0, 1.1), (1909.0, 18.0, 0.6), (2799.0, 22.0, 0.5),
)


def make_zqso_model(seed: int = 0, k: int = 20) -> Dict[str, np.ndarray]:
    """Synthetic stand-in for learned_zqso_only_model_*.mat (zqso_gp.py:283-319): grid 910:0.25:3000."""
    rng = np.random.default_rng(seed + 1000)
    rest_wavelengths = 910.0 + 0.25 * np.arange(8361)
    mu = 0.9 * (rest_wavelengths / 1216.0) ** -1.2
    for center, width, height in _EMISSION_LINES:
        mu = mu + height * np.exp(-0.5 * ((rest_wavelengths - center) / width) ** 2)
    mu = np.where(rest_wavelengths < 1216.0, mu * (0.75 + 0.25 * (rest_wavelengths - 910.0) / 306.0), mu)
    N = rest_wavelengths.shape[0]
    kern_x = np.arange(-160, 161)
    kern = np.exp(-0.5 * (kern_x / 40.0) ** 2)
    kern /= np.sqrt(np.sum(kern**2))
    M = np.empty((N, k))
    for j in range(k):
        white = rng.standard_normal(N + kern_x.shape[0] - 1)
        M[:, j] = np.convolve(white, kern, "valid") * 0.15 * (0.88**j)
    return dict(
        rest_wavelengths=rest_wavelengths, mu=mu, M=np.ascontiguousarray(M),
        bluewards_mu=0.62, redwards_mu=0.21, bluewards_sigma=0.31, redwards_sigma=0.17,
    )


def make_zqso_spectrum(model: Dict[str, np.ndarray], z_qso: float, seed: int, num_pixels: int = 4650):
    """BOSS-like spectrum drawn from the zQSO model at redshift z_qso (no absorbers)."""
    rng = np.random.default_rng(seed + 5000)
    loglam = 3.5523 + 1e-4 * np.arange(num_pixels)
    wavelengths = 10.0**loglam
    rest = wavelengths / (1 + z_qso)
    rw = model["rest_wavelengths"]
    inside = (rest >= rw[0]) & (rest <= rw[-1])
    clean = np.where(rest < rw[0], model["bluewards_mu"], model["redwards_mu"]).astype(np.float64)
    clean[inside] = np.interp(rest[inside], rw, model["mu"])
    xi = rng.standard_normal(model["M"].shape[1])
    for j in range(model["M"].shape[1]):
        clean[inside] += np.interp(rest[inside], rw, model["M"][:, j]) * xi[j]
    sigma_pix = np.exp(rng.normal(np.log(0.2), 0.3, size=num_pixels)) * rng.uniform(0.5, 1.5)
    normaliser = np.exp(rng.normal(np.log(3.0), 0.5))
    flux = (clean + sigma_pix * rng.standard_normal(num_pixels)) * normaliser
    noise_variance = (sigma_pix * normaliser) ** 2
    ivar0 = rng.random(num_pixels) < 0.012
    bright = rng.random(num_pixels) < 0.004
    with np.errstate(divide="ignore"):
        noise_variance = np.where(ivar0, np.inf, noise_variance)
    flux = np.where(ivar0, 0.0, flux)
    return wavelengths, flux, noise_variance, ivar0 | bright


# ----------------------------------------------------------------------------------------
# the benchmark / parity-sweep workloads (bench.py, tools/parity_sweep.py, tests/golden/make_golden.py)
# ----------------------------------------------------------------------------------------
def make_workload(num_spectra: int, seed0: int = 0, num_dla_samples: int = 10000, num_lines: int = 3):
    """
    `num_spectra` synthetic BOSS-like spectra with the model, prior and sample arrays of a DLA catalogue run.
    Spectrum i is make_spectrum(model, z_qsos[i], seed = seed0 * 1000003 + i), z_qsos drawn with seed 12345 + seed0.
    Returns (params, model, prior, dla_arrays, subdla_arrays, z_qsos, spectra).
    """
    params = Parameters(num_dla_samples=num_dla_samples, num_lines=num_lines)
    model = make_learned_model(0)
    prior = SyntheticPrior(params)
    dla = make_dla_sample_arrays(params)
    sub = make_subdla_sample_arrays(params)
    z_qsos = sample_z_qsos(num_spectra, seed=12345 + seed0)
    spectra = [make_spectrum(model, z_qsos[i], seed=seed0 * 1000003 + i) for i in range(num_spectra)]
    return params, model, prior, dla, sub, z_qsos, spectra


def make_zqso_workload(num_spectra: int, seed0: int = 0):
    """`num_spectra` spectra for the zQSO sweep: (model, z_true, spectra); spectrum i has seed seed0 * 1000003 + i."""
    model = make_zqso_model(0)
    z_true = sample_z_qsos(num_spectra, seed=777 + seed0)
    spectra = [make_zqso_spectrum(model, float(z_true[i]), seed=seed0 * 1000003 + i) for i in range(num_spectra)]
    return model, z_true, spectra
