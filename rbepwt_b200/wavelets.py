"""Filter banks in PyWavelets' convention -- the `wavelet` argument of the reference's
pywt.dwt / pywt.idwt calls (/root/reference/rbepwt.py:2041, 2067).

Built-in (tools/gen_wavelets.py): haar, db1..db20, bior1.1 .. bior3.9 (spline pairs), bior4.4, and the
rbio mirror of every bior.  Any other name is looked up in PyWavelets when it is installed; a
pywt.Wavelet-like object (dec_lo/dec_hi/rec_lo/rec_hi) or a 4-tuple of sequences is used as given.
"""
import numpy as np

from ._wavelet_tables import TABLES


def _from_lowpass(dec_lo, rec_lo):
    dec_lo = np.asarray(dec_lo, dtype=np.float64)
    rec_lo = np.asarray(rec_lo, dtype=np.float64)
    sign = np.where(np.arange(dec_lo.size) % 2 == 0, 1.0, -1.0)
    rec_hi = sign * dec_lo          # rec_hi[i] = (-1)^i     dec_lo[i]
    dec_hi = -sign * rec_lo         # dec_hi[i] = (-1)^(i+1) rec_lo[i]
    return dec_lo, dec_hi, rec_lo, rec_hi


def wavelist():
    names = ["haar"] + sorted(TABLES)
    names += ["rbio" + n[4:] for n in sorted(TABLES) if n.startswith("bior")]
    return names


def filter_bank(wavelet):
    """(dec_lo, dec_hi, rec_lo, rec_hi) float64 arrays of equal, even length."""
    if isinstance(wavelet, str):
        name = "db1" if wavelet == "haar" else wavelet
        if name in TABLES:
            bank = _from_lowpass(*TABLES[name])
        elif name.startswith("rbio") and "bior" + name[4:] in TABLES:
            dl, dh, rl, rh = _from_lowpass(*TABLES["bior" + name[4:]])
            bank = (rl[::-1].copy(), rh[::-1].copy(), dl[::-1].copy(), dh[::-1].copy())
        else:
            try:
                import pywt  # optional
            except ImportError:
                raise ValueError("unknown wavelet %r (built in: %s; install PyWavelets for the rest)"
                                 % (wavelet, ", ".join(wavelist())))
            bank = tuple(np.asarray(f, dtype=np.float64) for f in pywt.Wavelet(wavelet).filter_bank)
    elif hasattr(wavelet, "dec_lo"):
        bank = tuple(np.asarray(f, dtype=np.float64)
                     for f in (wavelet.dec_lo, wavelet.dec_hi, wavelet.rec_lo, wavelet.rec_hi))
    else:
        bank = tuple(np.asarray(f, dtype=np.float64) for f in wavelet)
        if len(bank) != 4:
            raise ValueError("a filter bank is (dec_lo, dec_hi, rec_lo, rec_hi)")
    n = bank[0].size
    if n < 2 or n % 2 or any(f.ndim != 1 or f.size != n for f in bank):
        raise ValueError("filter bank must hold four 1-D filters of the same even length")
    return tuple(np.ascontiguousarray(f) for f in bank)
