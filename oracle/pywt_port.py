"""TEST INFRASTRUCTURE ONLY -- never imported by the product (rbepwt_b200/).

CPU restatement of the two PyWavelets calls on the reference's hot path:

    pywt.dwt (x, wavelet, 'periodization')      /root/reference/rbepwt.py:2041
    pywt.idwt(cA, cD, wavelet, 'periodization') /root/reference/rbepwt.py:2067

PyWavelets is a third-party dependency of the reference; it is NOT vendored under
/root/reference and is NOT installed in this image (no network), and the reference
pins no version (README.md:4-10 only lists names; the skimage API it uses dates the
environment to PyWavelets 0.5.x).  PARITY UNPINNED for this arithmetic: what pins
the restatement is (1) the PyWavelets test-suite vector for db2/periodization,
(2) the haar closed form, (3) perfect reconstruction, (4) the filter identities
checked in tests/test_oracle_dwt.py.

Algorithm (PyWavelets `downsampling_convolution_periodization` /
`upsampling_convolution_valid_sf_periodization`, restated):

  forward, x of even length N, filter length F:
      cA[o] = sum_{j=0..F-1} dec_lo[j] * x[(2o + F/2 - j) mod N]      o = 0..N/2-1
      cD[o] = same with dec_hi
  accumulated in ascending j, one multiply and one add per tap (no FMA), starting
  from 0.0 -- the order PyWavelets' C loop uses.  Odd N: x[N-1] is appended once.

  inverse:
      x[n] = S_lo[n] + S_hi[n],
      S_lo[n] = sum over taps m (ascending) and o with n == 2o - F/2 + 1 + m (mod N)
                of rec_lo[m] * cA[o];   S_hi likewise with rec_hi, cD.
  The wrap is modulo N even when N < F (several wraps).
"""
import numpy as np

_R = 0.7071067811865476  # 1/sqrt(2) as PyWavelets tabulates it

# rec_lo / dec_lo tables in PyWavelets' convention.
_DEC_LO = {
    "haar": [_R, _R],
    "db1": [_R, _R],
    "db2": [-0.12940952255092145, 0.22414386804185735, 0.836516303737469, 0.48296291314469025],
    "db3": [0.035226291882100656, -0.08544127388224149, -0.13501102001039084,
            0.4598775021193313, 0.8068915093133388, 0.3326705529509569],
    "db4": [-0.010597401784997278, 0.032883011666982945, 0.030841381835986965,
            -0.18703481171888114, -0.02798376941698385, 0.6308807679295904,
            0.7148465705525415, 0.23037781330885523],
    "bior4.4": [0.0, 0.03782845550726404, -0.023849465019556843, -0.11062440441843718,
                0.37740285561283066, 0.8526986790088938, 0.37740285561283066,
                -0.11062440441843718, -0.023849465019556843, 0.03782845550726404],
}
_REC_LO = {
    "bior4.4": [0.0, -0.06453888262869706, -0.04068941760916406, 0.41809227322161724,
                0.7884856164055829, 0.41809227322161724, -0.04068941760916406,
                -0.06453888262869706, 0.0, 0.0],
}


def filter_bank(wavelet):
    """(dec_lo, dec_hi, rec_lo, rec_hi) as float64 arrays.

    `wavelet` is a PyWavelets name from the small table above, a 4-tuple of
    sequences, or any object with dec_lo/dec_hi/rec_lo/rec_hi attributes.
    Orthogonal: rec_lo = reversed dec_lo.  Both families:
    rec_hi[i] = (-1)^i dec_lo[i], dec_hi[i] = (-1)^(i+1) rec_lo[i].
    """
    if isinstance(wavelet, str):
        if wavelet not in _DEC_LO:
            raise ValueError("oracle has no table for wavelet %r" % (wavelet,))
        dec_lo = np.array(_DEC_LO[wavelet], dtype=np.float64)
        if wavelet in _REC_LO:
            rec_lo = np.array(_REC_LO[wavelet], dtype=np.float64)
        else:
            rec_lo = dec_lo[::-1].copy()
        sign = np.array([(-1.0) ** i for i in range(len(dec_lo))])
        rec_hi = sign * dec_lo
        dec_hi = -sign * rec_lo
        return dec_lo, dec_hi, rec_lo, rec_hi
    if hasattr(wavelet, "dec_lo"):
        wavelet = (wavelet.dec_lo, wavelet.dec_hi, wavelet.rec_lo, wavelet.rec_hi)
    dec_lo, dec_hi, rec_lo, rec_hi = (np.asarray(f, dtype=np.float64) for f in wavelet)
    return dec_lo, dec_hi, rec_lo, rec_hi


def dwt(data, wavelet, mode="periodization"):
    if mode != "periodization":
        raise ValueError("only mode='periodization' is on the reference path")
    dec_lo, dec_hi, _, _ = filter_bank(wavelet)
    x = np.asarray(data, dtype=np.float64)
    if x.size % 2:
        x = np.append(x, x[-1])
    n, flen = x.size, len(dec_lo)
    o = np.arange(n // 2)
    ca = np.zeros(n // 2)
    cd = np.zeros(n // 2)
    for j in range(flen):
        xs = x[(2 * o + flen // 2 - j) % n]
        ca = ca + dec_lo[j] * xs
        cd = cd + dec_hi[j] * xs
    return ca, cd


def idwt(ca, cd, wavelet, mode="periodization"):
    if mode != "periodization":
        raise ValueError("only mode='periodization' is on the reference path")
    _, _, rec_lo, rec_hi = filter_bank(wavelet)
    ca = np.asarray(ca, dtype=np.float64)
    cd = np.asarray(cd, dtype=np.float64)
    half, flen = ca.size, len(rec_lo)
    n = 2 * half
    o = np.arange(half)
    s_lo = np.zeros(n)
    s_hi = np.zeros(n)
    for m in range(flen):
        idx = (2 * o - flen // 2 + 1 + m) % n  # injective in o
        s_lo[idx] = s_lo[idx] + rec_lo[m] * ca
        s_hi[idx] = s_hi[idx] + rec_hi[m] * cd
    return s_lo + s_hi


# ---- the tensor-product baseline (class Dwt, /root/reference/rbepwt.py:2249-2298) ----------------------------------
# pywt.wavedec2 / waverec2 with mode='periodization', restated from the two 1-D routines above: dwt2 transforms along
# axis 0, then along axis 1 (pywt.dwtn visits the axes in order), keys 'aa', 'da' (cH), 'ad' (cV), 'dd' (cD), and the
# next level transforms 'aa'.  PARITY UNPINNED like the 1-D routines (PyWavelets is absent here).

def _dwt_axis(a, wavelet, axis):
    lo = np.apply_along_axis(lambda v: dwt(v, wavelet)[0], axis, a)
    hi = np.apply_along_axis(lambda v: dwt(v, wavelet)[1], axis, a)
    return lo, hi


def _idwt_axis(lo, hi, wavelet, axis):
    lo, hi = np.moveaxis(lo, axis, -1), np.moveaxis(hi, axis, -1)
    out = np.stack([idwt(l, h, wavelet) for l, h in zip(lo.reshape(-1, lo.shape[-1]), hi.reshape(-1, hi.shape[-1]))])
    return np.moveaxis(out.reshape(lo.shape[:-1] + (2 * lo.shape[-1],)), -1, axis)


def wavedec2(data, wavelet, level, mode="periodization"):
    """[cA_L, (cH_L, cV_L, cD_L), ..., (cH_1, cV_1, cD_1)]"""
    a = np.asarray(data, dtype=np.float64)
    out = []
    for _ in range(level):
        lo0, hi0 = _dwt_axis(a, wavelet, 0)
        aa, ad = _dwt_axis(lo0, wavelet, 1)
        da, dd = _dwt_axis(hi0, wavelet, 1)
        out.append((da, ad, dd))
        a = aa
    return [a] + out[::-1]


def waverec2(coeffs, wavelet, mode="periodization"):
    a = coeffs[0]
    for (da, ad, dd) in coeffs[1:]:
        lo0 = _idwt_axis(a, ad, wavelet, 1)
        hi0 = _idwt_axis(da, dd, wavelet, 1)
        a = _idwt_axis(lo0, hi0, wavelet, 0)
    return a


def dwt2_baseline(img, levels, wavelet, ncoefs):
    """Dwt.encode -> threshold_coefs(ncoefs) -> decode + the clip of Image.decode_dwt (rbepwt.py:318-333, 2262-2298).
    Returns (decoded image, number of non-zero coefficients, sorted magnitudes of the kept coefficients)."""
    co = wavedec2(img, wavelet, levels)
    flat = np.concatenate([co[0].ravel()] + [np.concatenate([h.ravel(), v.ravel(), d.ravel()]) for h, v, d in co[1:]])
    th = np.zeros_like(flat)
    if ncoefs <= 0 or ncoefs >= flat.size:  # the reference's `count == ncoefs` test never fires: everything is kept
        th[:] = flat
    else:
        keep = np.argsort(np.abs(flat), kind="stable")[::-1][:ncoefs]
        th[keep] = flat[keep]
    size = co[0].size
    new = [th[:size].reshape(co[0].shape)]
    last = size
    for h, v, d in co[1:]:
        s = h.size
        new.append((th[last:last + s].reshape(h.shape), th[last + s:last + 2 * s].reshape(v.shape), th[last + 2 * s:last + 3 * s].reshape(d.shape)))
        last += 3 * s
    dec = waverec2(new, wavelet)
    return np.clip(dec, 0.0, 255.0), int(np.count_nonzero(th)), np.sort(np.abs(th[th != 0]))
