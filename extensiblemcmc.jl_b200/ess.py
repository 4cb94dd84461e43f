"""Effective sample size (Geyer's initial positive sequence), host-side numpy.
The reference has no ESS code; BASELINE.json's second metric (ESS/sec) needs one."""
import numpy as np


def ess_geyer(x):
    """x: [n_samples, ...] -> ESS per trailing index."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    flat = x.reshape(n, -1)
    xc = flat - flat.mean(axis=0)
    nfft = 1 << int(np.ceil(np.log2(2 * n)))
    f = np.fft.rfft(xc, nfft, axis=0)
    acov = np.fft.irfft(f * np.conj(f), nfft, axis=0)[:n] / n
    var = acov[0]
    out = np.empty(flat.shape[1])
    for k in range(flat.shape[1]):
        if not var[k] > 0:
            out[k] = float(n) if n > 0 else 0.0
            continue
        rho = acov[:, k] / var[k]
        m = (n // 2) * 2
        pair = rho[0:m:2] + rho[1:m:2]            # Gamma_t = rho_2t + rho_2t+1
        neg = np.nonzero(pair <= 0)[0]
        T = neg[0] if neg.size else pair.shape[0]
        pair = np.minimum.accumulate(pair[:T]) if T > 0 else pair[:0]   # initial monotone sequence
        tau = -1.0 + 2.0 * pair.sum() if T > 0 else 1.0
        out[k] = n / max(tau, 1.0 / n)
    return out.reshape(x.shape[1:])
