import numpy as np


def sdr_db(ref, est):
    ref = np.asarray(ref, dtype=np.float64)
    est = np.asarray(est, dtype=np.float64)
    num = np.sum(ref**2)
    den = np.sum((ref - est) ** 2)
    if den == 0:
        return np.inf
    return 10.0 * np.log10(num / den + 1e-300)


def rel_err(ref, est):
    ref = np.asarray(ref, dtype=np.float64)
    est = np.asarray(est, dtype=np.float64)
    return float(np.max(np.abs(ref - est)) / (np.max(np.abs(ref)) + 1e-30))
