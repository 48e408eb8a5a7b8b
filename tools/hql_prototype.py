"""Numpy prototype of the Householder + implicit-QL eigensolver that eigh_hql.cuh implements.
Same loop structure and formulas as the CUDA kernels (zhetd2-lower / zung2r-in-place /
dsteqr-QL with recorded rotations / zlasr-style application), checked against np.linalg.eigh."""
import numpy as np

EPS = 2.0 ** -53
SAFMIN = 2.2250738585072014e-308


def tridiag_lower(A):
    """A (Hermitian, full) -> d, e (real), reflectors stored in A's lower triangle, tau."""
    A = A.copy()
    n = A.shape[0]
    d = np.zeros(n)
    e = np.zeros(max(n - 1, 0))
    tau = np.zeros(max(n - 1, 0), dtype=complex)
    for k in range(n - 1):
        m = n - k - 1
        alpha = A[k + 1, k]
        x = A[k + 2:, k]
        xnorm2 = np.sum(np.abs(x) ** 2)
        d[k] = A[k, k].real
        if xnorm2 == 0.0 and alpha.imag == 0.0:
            tau[k] = 0.0
            e[k] = alpha.real
            continue
        beta = -np.copysign(np.sqrt(alpha.real ** 2 + alpha.imag ** 2 + xnorm2), alpha.real)
        t = complex((beta - alpha.real) / beta, -alpha.imag / beta)
        scale = 1.0 / (alpha - beta)
        v = np.empty(m, dtype=complex)
        v[0] = 1.0
        v[1:] = x * scale
        A[k + 2:, k] = v[1:]
        e[k] = beta
        tau[k] = t
        A22 = A[k + 1:, k + 1:]
        p = t * (A22 @ v)
        al = -0.5 * t * np.vdot(p, v)
        w = p + al * v
        A22 -= np.outer(v, w.conj()) + np.outer(w, v.conj())
    d[n - 1] = A[n - 1, n - 1].real
    return d, e, A, tau


def form_q_inplace(A, tau):
    """Overwrite A (holding the reflectors below the first subdiagonal) with Q = H_0 ... H_{n-2}."""
    n = A.shape[0]
    Q = A  # in place
    for k in range(n - 2, -1, -1):
        m = n - k - 1
        t = tau[k]
        v1 = Q[k + 2:, k].copy()  # v[1:], v[0] = 1 implicit at row k+1
        if m > 1:
            sub = Q[k + 2:, k + 2:]
            u = v1.conj() @ sub  # row k+1 of those columns is zero before this step
            Q[k + 1, k + 2:] = -t * u
            sub -= np.outer(t * v1, u)
        Q[k + 1, k + 1] = 1.0 - t
        Q[k + 2:, k + 1] = -t * v1
        Q[k + 2:, k] = 0.0  # (column k is free now)
    Q[0, :] = 0.0
    Q[:, 0] = 0.0
    Q[0, 0] = 1.0
    return Q


def tql_record(d, e, maxit=60):
    """Implicit QL (dsteqr's QL branch, no scaling) on (d, e); returns eigenvalues (unsorted)
    and the recorded plane rotations as a list of sweeps (l, m, c[], s[]) where rotation j of a
    sweep acts on columns (i, i+1), i = m-1-j ... l  (zlasr side=R pivot=V direct=B)."""
    d = d.copy()
    n = len(d)
    e = np.concatenate([e.copy(), [0.0]])
    sweeps = []
    l = 0
    nit = 0
    eps2 = EPS * EPS * 4  # (2^-52)^2
    while l < n:
        # find small subdiagonal
        m = l
        while m < n - 1:
            tst = e[m] * e[m]
            if tst <= (eps2 * abs(d[m])) * abs(d[m + 1]) + SAFMIN:
                break
            m += 1
        if m < n - 1:
            e[m] = 0.0
        if m == l:
            l += 1
            continue
        if m == l + 1:
            # 2x2 block: dlaev2
            a, b, c_ = d[l], e[l], d[l + 1]
            rt1, rt2, cs, sn = laev2(a, b, c_)
            sweeps.append((l, l + 1, np.array([cs]), np.array([sn])))
            d[l], d[l + 1] = rt1, rt2
            e[l] = 0.0
            l += 2
            continue
        if nit >= maxit * n:
            raise RuntimeError("QL failed to converge")
        nit += 1
        p = d[l]
        g = (d[l + 1] - p) / (2.0 * e[l])
        r = np.hypot(g, 1.0)
        g = d[m] - p + e[l] / (g + np.copysign(r, g))
        s = 1.0
        c = 1.0
        p = 0.0
        cs = np.empty(m - l)
        ss = np.empty(m - l)
        for i in range(m - 1, l - 1, -1):
            f = s * e[i]
            b = c * e[i]
            # dlartg(g, f)
            r = np.hypot(g, f)
            if r == 0.0:
                c, s = 1.0, 0.0
            else:
                c, s = g / r, f / r
            if i != m - 1:
                e[i + 1] = r
            g = d[i + 1] - p
            r = (d[i] - g) * s + 2.0 * c * b
            p = s * r
            d[i + 1] = g + p
            g = c * r - b
            cs[m - 1 - i] = c
            ss[m - 1 - i] = -s
        d[l] -= p
        e[l] = g
        sweeps.append((l, m, cs, ss))
    return d, sweeps


def laev2(a, b, c):
    """LAPACK dlaev2: eigen-decomposition of [[a,b],[b,c]]; (cs1, sn1) is the unit right
    eigenvector for rt1: [cs1 sn1; -sn1 cs1] [[a,b],[b,c]] [cs1 -sn1; sn1 cs1] = diag(rt1, rt2)."""
    sm = a + c
    df = a - c
    adf = abs(df)
    tb = b + b
    ab = abs(tb)
    acmx, acmn = (a, c) if abs(a) > abs(c) else (c, a)
    if adf > ab:
        rt = adf * np.sqrt(1.0 + (ab / adf) ** 2)
    elif adf < ab:
        rt = ab * np.sqrt(1.0 + (adf / ab) ** 2)
    else:
        rt = ab * np.sqrt(2.0)
    if sm < 0.0:
        rt1 = 0.5 * (sm - rt)
        sgn1 = -1
        rt2 = (acmx / rt1) * acmn - (b / rt1) * b
    elif sm > 0.0:
        rt1 = 0.5 * (sm + rt)
        sgn1 = 1
        rt2 = (acmx / rt1) * acmn - (b / rt1) * b
    else:
        rt1 = 0.5 * rt
        rt2 = -0.5 * rt
        sgn1 = 1
    if df >= 0.0:
        cs = df + rt
        sgn2 = 1
    else:
        cs = df - rt
        sgn2 = -1
    acs = abs(cs)
    if acs > ab:
        ct = -tb / cs
        sn1 = 1.0 / np.sqrt(1.0 + ct * ct)
        cs1 = ct * sn1
    else:
        if ab == 0.0:
            cs1 = 1.0
            sn1 = 0.0
        else:
            tn = -cs / tb
            cs1 = 1.0 / np.sqrt(1.0 + tn * tn)
            sn1 = tn * cs1
    if sgn1 == sgn2:
        tn = cs1
        cs1 = -sn1
        sn1 = tn
    return rt1, rt2, cs1, sn1


def apply_sweeps(Z, sweeps):
    """zlasr(side=R, pivot=V, direct=B): for j = m-1 down to l: columns (j, j+1)."""
    Z = Z.copy()
    for (l, m, cs, ss) in sweeps:
        if m == l + 1 and len(cs) == 1:
            # 2x2 block from laev2: dsteqr stores work(l)=c, work(n-1+l)=s and calls
            # zlasr('R','V','B', n, 2, ...)
            pass
        for j in range(m - 1, l - 1, -1):
            c = cs[m - 1 - j]
            s = ss[m - 1 - j]
            temp = Z[:, j + 1].copy()
            Z[:, j + 1] = c * temp - s * Z[:, j]
            Z[:, j] = s * temp + c * Z[:, j]
    return Z


def eigh_hql(A):
    d, e, Ar, tau = tridiag_lower(A)
    Q = form_q_inplace(Ar, tau)
    lam, sweeps = tql_record(d, e)
    U = apply_sweeps(Q, sweeps)
    return lam, U, (d, e, Q, sweeps)


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    worst = 0
    for trial in range(60):
        n = int(rng.integers(1, 40))
        A = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
        A = A + A.conj().T
        kind = trial % 4
        if kind == 1:  # degenerate
            q, _ = np.linalg.qr(A)
            A = (q * (np.arange(n) // 3).astype(float)) @ q.conj().T
            A = 0.5 * (A + A.conj().T)
        elif kind == 2:  # sparse-ish, real
            A = np.diag(rng.normal(size=n)).astype(complex)
            if n > 2:
                A[0, n - 1] = A[n - 1, 0] = 0.3
        elif kind == 3:  # already tridiagonal / diagonal
            A = np.diag(rng.normal(size=n)).astype(complex)
        lam, U, (d, e, Q, sweeps) = eigh_hql(A)
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        e1 = np.abs(Q.conj().T @ A @ Q - T).max() if n > 1 else 0
        e2 = np.abs(A @ U - U * lam).max()
        e3 = np.abs(U.conj().T @ U - np.eye(n)).max()
        e4 = np.abs(np.sort(lam) - np.linalg.eigvalsh(A)).max()
        nrot = sum(len(s[2]) for s in sweeps)
        worst = max(worst, e1, e2, e3, e4)
        if max(e1, e2, e3, e4) > 1e-12:
            print("FAIL n=%d kind=%d tri %.1e res %.1e orth %.1e eval %.1e" % (n, kind, e1, e2, e3, e4))
    print("worst error", worst)
    for n in (32, 96):
        A = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
        A = A + A.conj().T
        lam, U, (d, e, Q, sweeps) = eigh_hql(A)
        nrot = sum(len(s[2]) for s in sweeps)
        print("n=%d: sweeps %d rotations %d (%.2f n^2) resid %.1e" % (n, len(sweeps), nrot, nrot / n / n, np.abs(A @ U - U * lam).max()))
