"""
read_spec.py : spectrum readers with the reference's return convention
(read_spec.py:22-71): (wavelengths, flux, noise_variance, pixel_mask).

FITS parsing is outside the hot path; `read_spec` is kept so that `process_qso` has the
reference's default argument.  It needs astropy (absent from the build image) and raises a clear
error without it.  For catalogue runs use `preload.preload` once and feed the engine from the
memory-mapped store (`preload.PreloadedSpectra`).
"""
from typing import Tuple

import numpy as np

BRIGHTSKY = 24


def arrays_from_boss_columns(loglam, flux, ivar, and_mask) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """The arithmetic of read_spec.py:49-69 on the four columns of a BOSS `COADD` table."""
    ivar = np.asarray(ivar, dtype=np.float64)
    wavelengths = 10.0 ** np.asarray(loglam, dtype=np.float64)
    zero = ivar == 0
    noise_variance = np.full(ivar.shape, np.nan)
    noise_variance[~zero] = 1.0 / ivar[~zero]
    pixel_mask = zero | (((np.asarray(and_mask).astype(np.int64) >> BRIGHTSKY) & 1).astype(bool))
    return wavelengths, np.asarray(flux, dtype=np.float64), noise_variance, pixel_mask


def read_spec(filename: str) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """SDSS DR12Q coadded "speclite" FITS file (read_spec.py:22-71)."""
    try:
        from astropy.io import fits
    except ImportError as e:  # pragma: no cover - astropy is not in the build image
        raise ImportError("read_spec needs astropy to parse FITS files; pass your own read_spec callable or "
                          "a preload.PreloadedSpectra store to process_qso") from e
    with fits.open(filename) as hdu:
        try:
            data = hdu["COADD"].data
        except KeyError:
            data = hdu[1].data
        return arrays_from_boss_columns(data["loglam"], data["flux"], data["ivar"], data["and_mask"])


def read_spec_dr14q(filename: str) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """SDSS DR14Q file: first binary table, no COADD extension (read_spec.py:74-117)."""
    try:
        from astropy.io import fits
    except ImportError as e:  # pragma: no cover
        raise ImportError("read_spec_dr14q needs astropy to parse FITS files") from e
    with fits.open(filename) as hdu:
        data = hdu[1].data
        return arrays_from_boss_columns(data["loglam"], data["flux"], data["ivar"], data["and_mask"])
