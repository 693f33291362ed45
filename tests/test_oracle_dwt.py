"""Known-answer tests pinning the PyWavelets restatement (oracle/pywt_port.py) -- SURVEY 8c."""
import numpy as np
import pytest

from oracle import pywt_port as pw


def test_pywt_testsuite_vector_db2_periodization():
    # PyWavelets' own test-suite vector (test_dwt_idwt / periodization, db2)
    x = [1, 2, 1, 5, -1, 8, 4, 6]
    ca, cd = pw.dwt(x, "db2", "periodization")
    np.testing.assert_allclose(ca, [4.053172, 3.05257099, 2.85381112, 8.42522221], atol=1e-7)
    np.testing.assert_allclose(cd, [0.18946869, 4.18258152, 4.33737503, 2.60428326], atol=1e-7)
    np.testing.assert_allclose(pw.idwt(ca, cd, "db2", "periodization"), x, atol=1e-13)


def test_haar_closed_form():
    ca, cd = pw.dwt([1, 2, 3, 4], "haar")
    r = np.sqrt(0.5)
    np.testing.assert_allclose(ca, [3 * r, 7 * r], rtol=1e-15)
    np.testing.assert_allclose(cd, [-r, -r], rtol=1e-15)


@pytest.mark.parametrize("wav", ["haar", "db1", "db2", "db3", "db4", "bior4.4"])
@pytest.mark.parametrize("n", [2, 4, 8, 16, 64, 1024])
def test_perfect_reconstruction_including_multiwrap(wav, n):
    x = np.random.default_rng(n).uniform(0, 255, n)
    ca, cd = pw.dwt(x, wav)
    assert ca.size == cd.size == n // 2
    np.testing.assert_allclose(pw.idwt(ca, cd, wav), x, atol=2e-9 * 255)


@pytest.mark.parametrize("wav", ["haar", "db2", "db3", "db4", "bior4.4"])
def test_filter_identities(wav):
    dec_lo, dec_hi, rec_lo, rec_hi = pw.filter_bank(wav)
    assert abs(dec_lo.sum() - np.sqrt(2)) < 1e-11 and abs(rec_lo.sum() - np.sqrt(2)) < 1e-11
    assert abs(dec_hi.sum()) < 1e-11 and abs(rec_hi.sum()) < 1e-11
    f = len(dec_lo)
    for s in range(0, f, 2):  # biorthogonality of the even shifts
        dot = sum(dec_lo[i] * rec_lo[f - 1 - (i - s)] for i in range(s, f))
        assert abs(dot - (1.0 if s == 0 else 0.0)) < 1e-10


def test_odd_length_is_padded_once():
    ca, cd = pw.dwt([1.0, 2.0, 3.0], "haar")
    ca2, cd2 = pw.dwt([1.0, 2.0, 3.0, 3.0], "haar")
    np.testing.assert_array_equal(ca, ca2)
    np.testing.assert_array_equal(cd, cd2)
