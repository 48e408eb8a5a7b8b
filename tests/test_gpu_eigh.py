"""Batched eigensolvers (musim_eigh through the C ABI) against numpy.linalg.eigh.
Eigenvectors are never compared directly (degeneracy): residual ||A U - U diag(l)||,
orthonormality and eigenvalues are."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _eigh(A, method):
    import torch

    from muspinsim_b200 import _lib

    A = np.ascontiguousarray(A, dtype=np.complex128)
    b, d, _ = A.shape
    At = torch.from_numpy(A).cuda()
    ev = torch.zeros(b, d, dtype=torch.float64, device="cuda")
    U = torch.zeros(b, d, d, dtype=torch.complex128, device="cuda")
    _lib.eigh_device(0, d, b, At.data_ptr(), ev.data_ptr(), U.data_ptr(), method)
    torch.cuda.synchronize()
    return ev.cpu().numpy(), U.cpu().numpy()


def _check(A, method, tol=5e-13):
    ev, U = _eigh(A, method)
    d = A.shape[-1]
    ref = np.linalg.eigvalsh(A)
    scale = max(1.0, np.abs(ref).max())
    assert np.all(np.diff(ev, axis=1) >= 0), "eigenvalues must be ascending"
    assert np.abs(ev - ref).max() / scale < tol
    assert np.abs(A @ U - U * ev[:, None, :]).max() / scale < tol
    assert np.abs(np.conj(np.transpose(U, (0, 2, 1))) @ U - np.eye(d)).max() < tol


def _rand_herm(rng, b, d):
    A = rng.normal(size=(b, d, d)) + 1j * rng.normal(size=(b, d, d))
    return np.ascontiguousarray(A + np.conj(np.transpose(A, (0, 2, 1))))


@pytest.mark.parametrize("method", [1, 2])
@pytest.mark.parametrize("d", [1, 2, 3, 4, 7, 8, 12, 16, 17, 24, 31, 32, 33, 48, 64, 83, 96])
def test_random_hermitian(method, d):
    if method == 1 and d > 96:
        pytest.skip("Jacobi limited by shared memory")
    rng = np.random.default_rng(d)
    _check(_rand_herm(rng, 24 if d < 64 else 6, d), method)


@pytest.mark.parametrize("method", [1, 2])
@pytest.mark.parametrize("d", [4, 12, 32, 96])
def test_degenerate_and_structured(method, d):
    rng = np.random.default_rng(100 + d)
    mats = []
    q, _ = np.linalg.qr(_rand_herm(rng, 1, d)[0])
    lam = (np.arange(d) // 4).astype(float)  # 4-fold degenerate levels
    mats.append((q * lam) @ q.conj().T)
    mats.append(np.zeros((d, d), dtype=complex))  # zero matrix (empty spin system)
    mats.append(np.eye(d, dtype=complex) * 3.0)  # multiple of identity
    mats.append(np.diag(rng.normal(size=d)).astype(complex))  # already diagonal
    T = np.diag(rng.normal(size=d)).astype(complex)  # real tridiagonal
    T += np.diag(np.ones(d - 1), 1) + np.diag(np.ones(d - 1), -1)
    mats.append(T)
    big = _rand_herm(rng, 1, d)[0] * 7e4  # ALC-scale norm (gamma_e * 2.6 T)
    mats.append(big)
    A = np.array([0.5 * (m + m.conj().T) for m in mats])
    _check(A, method)


@pytest.mark.parametrize("method", [1, 2])
def test_physical_hamiltonians(method):
    from muspinsim_b200 import workloads
    from muspinsim_b200.spinsys import system_from_spec

    for spec in (workloads.c5_large(2, 2), workloads.c3_alc(2, 2), workloads.c2_hfine_powder(2, 2)):
        s, _ = system_from_spec(spec)
        Z = s.zeeman_operators()
        rng = np.random.default_rng(5)
        A = np.array([s.hamiltonian + np.tensordot(rng.normal(size=3) * 2.0, Z, 1) for _ in range(4)])
        A[0] = s.hamiltonian  # zero field: massively degenerate
        _check(A, method)


def test_large_batch_all_matrices_processed():
    rng = np.random.default_rng(9)
    A = _rand_herm(rng, 3000, 12)
    _check(A, 2)


@pytest.mark.parametrize("d", [97, 112, 113, 128, 160, 200, 256, 384, 512, 1024])
def test_large_dimensions_global_memory_path(d):
    """d > 112 does not fit one SM's shared memory: working matrix in global memory
    (eigh_large.cuh); d <= 112 stays on the shared-memory kernels."""
    rng = np.random.default_rng(d)
    _check(_rand_herm(rng, 3 if d <= 256 else 2, d), 2, tol=5e-13 * max(1.0, d / 96.0))


def test_large_dimension_degenerate_and_physical():
    from muspinsim_b200.spinsys import system_from_spec

    d = 128
    rng = np.random.default_rng(7)
    q, _ = np.linalg.qr(_rand_herm(rng, 1, d)[0])
    lam = (np.arange(d) // 8).astype(float)
    mats = [(q * lam) @ q.conj().T, np.zeros((d, d), dtype=complex), np.eye(d, dtype=complex) * 2.0]
    s, _ = system_from_spec({"spins": ["mu", "e"] + ["H"] * 5,
                             "couplings": [{"type": "hyperfine", "i": 1, "value": np.diag([100.0, 100.0, 120.0])}]
                             + [{"type": "hyperfine", "i": 3 + k, "j": 2, "value": np.diag([3.0 + k, 4.0, 5.0])} for k in range(5)]})
    mats.append(s.hamiltonian)  # zero field: massively degenerate
    mats.append(s.hamiltonian + 0.3 * s.zeeman_operators()[2])
    A = np.array([0.5 * (m + m.conj().T) for m in mats])
    _check(A, 2, tol=2e-12)


@pytest.mark.parametrize("d", [33, 34, 40, 41, 47, 49, 56, 57, 63, 65, 72, 73, 80, 88, 90, 95])
def test_divide_and_conquer_range(d):
    """32 < d <= 96: tridiagonal divide and conquer (eigh_tdc.cuh) + compact-WY back-transformation
    (eigh_backwy.cuh); ragged dimensions (leaves cut at multiples of 8), degenerate and torn spectra."""
    rng = np.random.default_rng(1000 + d)
    _check(_rand_herm(rng, 8, d), 2)
    q, _ = np.linalg.qr(_rand_herm(rng, 1, d)[0])
    mats = [(q * (np.arange(d) // 3).astype(float)) @ q.conj().T, np.zeros((d, d), dtype=complex),
            np.eye(d, dtype=complex) * -2.0, np.diag(rng.normal(size=d)).astype(complex),
            np.diag(np.abs(np.arange(d) - (d - 1) / 2)).astype(complex) + np.diag(np.ones(d - 1), 1) + np.diag(np.ones(d - 1), -1),
            _rand_herm(rng, 1, d)[0] * 1e-7, _rand_herm(rng, 1, d)[0] * 7e4]
    _check(np.array([0.5 * (m + m.conj().T) for m in mats]), 2)


@pytest.mark.parametrize("d", [49, 56, 63, 64, 65, 71, 72, 73, 80, 88, 95, 96])
def test_structured_matrices_at_the_phase_boundaries(d):
    """The half-storage tridiagonalisation hands the trailing block from phase to phase (96 -> 72 -> 48 -> 32):
    sizes on both sides of every boundary, with the structures that exercise identity reflectors, dead tiles and
    deflation -- dense, block-diagonal, sparse, low rank + identity, already tridiagonal, graded."""
    rng = np.random.default_rng(1000 + d)
    mats = []
    for kind in range(6):
        A = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
        A = A + A.conj().T
        if kind == 1:
            m = d // 3
            A[:m, m:] = 0
            A[m:, :m] = 0
        elif kind == 2:
            A = A * (rng.uniform(size=(d, d)) < 0.08)
            A = A + A.conj().T
        elif kind == 3:
            v = rng.normal(size=(d, 3)) + 1j * rng.normal(size=(d, 3))
            A = v @ v.conj().T + 2 * np.eye(d)
        elif kind == 4:
            A = np.diag(rng.normal(size=d)).astype(complex) + np.diag(rng.normal(size=d - 1), 1)
            A = A + A.conj().T
        elif kind == 5:
            s = np.logspace(0, -8, d)
            A = A * s[:, None] * s[None, :]
        mats.append(A)
    _check(np.array(mats), 2, tol=1e-13)
