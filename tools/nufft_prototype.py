"""Accuracy prototype of the type-1 NUFFT polarisation kernel (K7n): non-uniform frequencies
-> uniform time samples, ES ("exponential of semicircle") spreading kernel evaluated by
per-tap piecewise polynomials.  numpy only; used to choose w / polynomial degree."""
import numpy as np


def es_kernel(z, beta):
    out = np.zeros_like(z)
    m = np.abs(z) < 1
    out[m] = np.exp(beta * (np.sqrt(1 - z[m] ** 2) - 1))
    return out


def tap_polys(w, beta, deg):
    """coef[l, m]: phi_l(x) = sum_m coef[l,m] x^m, x in [-1/2, 1/2] = fractional offset;
    tap l sits at distance (l - (w-1)/2 - x) grid units from the point (w odd or even)."""
    # Chebyshev nodes on [-1/2, 1/2]
    nn = deg + 1
    xs = 0.5 * np.cos(np.pi * (np.arange(nn) + 0.5) / nn)
    coef = np.zeros((w, deg + 1))
    for l in range(w):
        dist = l - (w - 1) / 2.0 - xs
        ys = es_kernel(2 * dist / w, beta)
        coef[l] = np.polynomial.polynomial.polyfit(xs, ys, deg)
    return coef


def kernel_ft(w, beta, M, ks, nq=200):
    xq, wq = np.polynomial.legendre.leggauss(nq)
    phi = es_kernel(xq.copy(), beta)
    return (w / 2.0) * (wq * phi) @ np.cos(np.outer(xq, ks) * np.pi * w / M)


def nufft1(theta, c, N, M, w, deg=None):
    beta = 2.30 * w
    th = np.mod(theta, 2 * np.pi)
    cp = c * np.exp(1j * (N // 2) * th)
    g = th * M / (2 * np.pi)
    # nearest grid point for odd w / floor+1/2 for even: centre index
    if w % 2:
        ic = np.rint(g)
    else:
        ic = np.floor(g) + 0.5
    x = g - ic  # in [-1/2, 1/2]
    grid = np.zeros(M, complex)
    coef = tap_polys(w, beta, deg) if deg else None
    for l in range(w):
        dist = l - (w - 1) / 2.0 - x
        if coef is None:
            ph = es_kernel(2 * dist / w, beta)
        else:
            ph = np.polynomial.polynomial.polyval(x, coef[l], tensor=False)
        idx = (ic + (l - (w - 1) / 2.0)).astype(np.int64) % M
        np.add.at(grid, idx, cp * ph)
    ks = np.arange(N) - N // 2
    F = np.fft.ifft(grid) * M
    return F[ks % M] / kernel_ft(w, beta, M, ks)


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    N, M = 1000, 2048
    npt = 20000
    f = rng.normal(0, 300.0, npt)
    f[:2000] = np.round(f[:2000])  # clusters / exact grid hits
    c = rng.normal(size=npt) + 1j * rng.normal(size=npt)
    c /= np.sum(np.abs(c))
    dt = 0.01
    theta = -2 * np.pi * f * dt
    k = np.arange(N)
    ref = (c[None, :] * np.exp(1j * np.outer(k, theta))).sum(1)
    for w in (10, 11, 12, 13, 14):
        for deg in (None, w + 1, w + 3, w + 5):
            got = nufft1(theta, c, N, M, w, deg)
            print("w=%2d deg=%s  max err / sum|c| = %.2e" % (w, deg, np.max(np.abs(got - ref))))
    for w in (12, 13):
        for deg in range(7, 14):
            got = nufft1(theta, c, N, M, w, deg)
            print("w=%2d deg=%s  max err / sum|c| = %.2e" % (w, deg, np.max(np.abs(got - ref))))


def nufft1_real_folded(theta, c, N, M, w=12, deg=10):
    """Re of the type-1 transform with a HERMITIAN-FOLDED fine grid (groundwork for the next
    step of polar_nufft.cuh): only Re(sum_m g[m] e^{2 pi i k' m / M}) is needed, which equals the
    transform of h[m] = (g[m] + conj g[M-m]) / 2, and h is Hermitian, so cells 0 .. M/2 suffice
    (half the shared memory per warp).  A window cell m > M/2 is stored at M - m conjugated;
    cells 0 and M/2 keep only their real part at reconstruction."""
    beta = 2.30 * w
    th = np.mod(theta, 2 * np.pi)
    cp = c * np.exp(1j * (N // 2) * th)
    g = th * M / (2 * np.pi)
    fl = np.minimum(np.floor(g), M - 1)
    x = g - fl - 0.5
    coef = tap_polys(w, beta, deg)
    half = np.zeros(M // 2 + 1, complex)
    for l in range(w):
        ph = np.polynomial.polynomial.polyval(x, coef[l], tensor=False)
        m = (fl.astype(np.int64) - (w // 2 - 1) + l) % M
        fold = m > M // 2
        j = np.where(fold, M - m, m)
        v = cp * ph
        np.add.at(half, j, np.where(fold, np.conj(v), v))
    h = np.zeros(M, complex)
    h[1 : M // 2] = half[1 : M // 2] / 2
    h[M // 2 + 1 :] = np.conj(half[1 : M // 2][::-1]) / 2
    h[0] = half[0].real
    h[M // 2] = half[M // 2].real
    ks = np.arange(N) - N // 2
    F = np.fft.ifft(h) * M
    return (F[ks % M] / kernel_ft(w, beta, M, ks)).real


def check_folded():
    rng = np.random.default_rng(1)
    N, M, npt = 1000, 2048, 20000
    f = rng.normal(0, 300.0, npt)
    f[:2000] = np.round(f[:2000])
    f[2000:2100] = 50.0 + rng.normal(0, 1e-3, 100)  # u = -0.5: windows straddling M/2
    c = rng.normal(size=npt) + 1j * rng.normal(size=npt)
    c /= np.sum(np.abs(c))
    theta = -2 * np.pi * f * 0.01
    k = np.arange(N)
    ref = (c[None, :] * np.exp(1j * np.outer(k, theta))).sum(1).real
    got = nufft1_real_folded(theta, c, N, M)
    print("folded grid: max err / sum|c| = %.2e" % np.max(np.abs(got - ref)))


if __name__ == "__main__":
    check_folded()
