"""Filter banks in PyWavelets' convention -- the `wavelet` argument of the reference's
pywt.dwt / pywt.idwt calls (/root/reference/rbepwt.py:2041, 2067).

Built-in (tools/gen_wavelets.py): haar, db1..db20, bior1.1 .. bior3.9 (spline pairs), bior4.4, and the
rbio mirror of every bior.  Any other name is looked up in PyWavelets when it is installed; a
pywt.Wavelet-like object (dec_lo/dec_hi/rec_lo/rec_hi) or a 4-tuple of sequences is used as given.
"""
import numpy as np

from ._wavelet_tables import TABLES as _EXACT_TABLES

# The built-in tables are the exact constructions rounded to float64 (tools/gen_wavelets.py).  PyWavelets tabulates the
# Daubechies filters as literals that differ from those in the 13th digit (db2[0] = -0.12940952255092145 there,
# -0.12940952255126037 exact), and the reference computes with PyWavelets' (rbepwt.py:2041): where the literals are
# known they are used, so that coefficients are bit-comparable with the reference's; when PyWavelets is importable
# its own filter bank is preferred for every name.  (dec_lo; rec_lo is its reverse for the orthogonal families.)
_PYWT_DEC_LO = {
    "db2": [-0.12940952255092145, 0.22414386804185735, 0.836516303737469, 0.48296291314469025],
    "db3": [0.035226291882100656, -0.08544127388224149, -0.13501102001039084,
            0.4598775021193313, 0.8068915093133388, 0.3326705529509569],
    "db4": [-0.010597401784997278, 0.032883011666982945, 0.030841381835986965,
            -0.18703481171888114, -0.02798376941698385, 0.6308807679295904,
            0.7148465705525415, 0.23037781330885523],
}
TABLES = dict(_EXACT_TABLES)
for _n, _lo in _PYWT_DEC_LO.items():
    TABLES[_n] = (tuple(_lo), tuple(_lo[::-1]))


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
        try:
            import pywt  # optional: the reference's own source of filter banks
            return tuple(np.ascontiguousarray(f, dtype=np.float64) for f in pywt.Wavelet(wavelet).filter_bank)
        except (ImportError, AttributeError):  # absent, or a stand-in module without Wavelet (oracle/ref_harness.py)
            pass
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
