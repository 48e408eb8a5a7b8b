"""The per-call boundary (muspinsim_b200.hamiltonian.Hamiltonian) with the reference's own
known-answer tests and error behaviour (tests/test_hamiltonian.py:10-128, validation.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sx_sz():
    from muspinsim_b200.spinsys import spin_operators

    sx, sy, sz = spin_operators(0.5)
    return sx, sz


def test_creation_errors():
    from muspinsim_b200.hamiltonian import Hamiltonian

    H = Hamiltonian(np.array([[1, 0], [0, -1]]))
    assert H.dimension == (2,)
    with pytest.raises(ValueError):
        Hamiltonian(np.array([[1, 1], [0, 1]]))


def test_diag():
    from muspinsim_b200.hamiltonian import Hamiltonian

    sx, sz = _sx_sz()
    evals, evecs = Hamiltonian(sx).diag()
    assert np.allclose(evals, [-0.5, 0.5], atol=1e-15)
    evecsT = np.array([[1.0, 1.0], [-1.0, 1.0]]) / 2**0.5
    assert np.allclose(abs(np.dot(evecs, evecsT)), np.eye(2))


def test_evolve_fast_evolve_integrate():
    from muspinsim_b200.hamiltonian import Hamiltonian

    sx, sz = _sx_sz()
    rho0 = 0.5 * np.eye(2) + sz
    t = np.linspace(0, 1, 100)
    H = Hamiltonian(sx)
    evol = H.evolve(rho0, t, [sz])
    assert evol.shape == (100, 1)
    assert np.allclose(evol[:, 0], 0.5 * np.cos(2 * np.pi * t), atol=1e-12)
    avg = H.integrate_decaying(rho0, 1.0, [sz])
    assert np.isclose(avg[0], 0.5 / (1.0 + 4 * np.pi**2), atol=1e-13)
    H2 = Hamiltonian(np.kron(sx, np.eye(2)))
    fe = H2.fast_evolve(2 * sz, t, 2)
    assert np.allclose(fe, 0.5 * np.cos(2 * np.pi * t), atol=1e-12)


def test_evolve_without_operators_returns_the_density_matrices():
    """hamiltonian.py:108-116: with no operators the reference returns the density matrices rho(t);
    here as [nt, d, d].  Checked against the closed form, the numpy restatement of the reference's
    lines, and -- where oracle/_ref travelled -- the reference's own Hamiltonian.evolve."""
    from muspinsim_b200.hamiltonian import Hamiltonian

    sx, sz = _sx_sz()
    t = np.linspace(0, 1, 7)
    rho = Hamiltonian(sx).evolve(0.5 * np.eye(2) + sz, t)  # spin precessing about x
    assert rho.shape == (7, 2, 2)
    assert np.allclose(np.trace(rho, axis1=1, axis2=2), 1.0, atol=1e-13)
    assert np.allclose(np.einsum("tij,ji->t", rho, sz), 0.5 * np.cos(2 * np.pi * t), atol=1e-13)

    rng = np.random.default_rng(7)
    for d, nt in ((12, 33), (48, 9), (96, 5), (100, 3)):
        A = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
        Hm = A + A.conj().T
        B = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
        r0 = B @ B.conj().T
        r0 /= np.trace(r0).real
        times = np.sort(rng.uniform(0.0, 0.3, nt))
        got = Hamiltonian(Hm).evolve(r0, times)
        ev, V = np.linalg.eigh(Hm)
        r0e = V.conj().T @ r0 @ V
        ll = -2.0j * np.pi * (ev[:, None] - ev[None, :])
        want = np.array([V @ (np.exp(ll * tt) * r0e) @ V.conj().T for tt in times])
        assert np.max(np.abs(got - want)) < 1e-11
        assert np.allclose(got, np.conj(np.transpose(got, (0, 2, 1))), atol=1e-12)  # Hermitian at every time

    from oracle import ref_driver

    if ref_driver.available():
        ref_driver._import()
        from muspinsim.hamiltonian import Hamiltonian as RefH
        from muspinsim.spinop import DensityOperator

        d = 8
        A = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
        Hm = A + A.conj().T
        B = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
        r0 = B @ B.conj().T
        r0 /= np.trace(r0).real
        times = np.linspace(0.0, 0.5, 6)
        ref = RefH(Hm, dim=(2, 2, 2)).evolve(DensityOperator(r0, dim=(2, 2, 2)), times)
        ref = np.array([np.asarray(x.matrix.toarray() if hasattr(x.matrix, "toarray") else x.matrix) for x in ref])
        assert np.max(np.abs(Hamiltonian(Hm, dim=(2, 2, 2)).evolve(r0, times) - ref)) < 1e-11


def test_lindbladian_evolve_without_operators_returns_the_density_matrices():
    """lindbladian.py:103-109 (operators=[]): density matrices of a dissipative two-spin system; consistent
    with the expectation values of the same object, trace-preserving, Hermitian, and equal to the
    reference's own Lindbladian.evolve where oracle/_ref travelled."""
    from muspinsim_b200.hamiltonian import Hamiltonian
    from muspinsim_b200.lindbladian import Lindbladian

    rng = np.random.default_rng(11)
    d = 4
    A = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    Hm = 0.5 * (A + A.conj().T)
    J = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    B = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    r0 = B @ B.conj().T
    r0 /= np.trace(r0).real
    O = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    O = O + O.conj().T
    times = np.linspace(0.0, 0.4, 5)
    L = Lindbladian.from_hamiltonian(Hamiltonian(Hm, dim=(2, 2)), [(J, 0.3)])
    got = L.evolve(r0, times)
    assert got.shape == (5, d, d)
    assert np.allclose(got[0], r0, atol=1e-11)
    assert np.allclose(np.trace(got, axis1=1, axis2=2), 1.0, atol=1e-10)
    assert np.allclose(got, np.conj(np.transpose(got, (0, 2, 1))), atol=1e-12)
    assert np.allclose(np.einsum("tij,ji->t", got, O), L.evolve(r0, times, [O])[:, 0], atol=1e-10)
    from oracle import ref_driver

    if ref_driver.available():
        ref_driver._import()
        from muspinsim.hamiltonian import Hamiltonian as RefH
        from muspinsim.lindbladian import Lindbladian as RefL
        from muspinsim.spinop import DensityOperator, SpinOperator

        ref = RefL.from_hamiltonian(RefH(Hm, dim=(2, 2)), [(SpinOperator(J, dim=(2, 2)), 0.3)]).evolve(
            DensityOperator(r0, dim=(2, 2)), times)
        want = np.array([np.asarray(x.matrix.toarray() if hasattr(x.matrix, "toarray") else x.matrix) for x in ref])
        assert np.max(np.abs(got - want)) < 1e-9


def test_validation_errors():
    from muspinsim_b200.hamiltonian import Hamiltonian

    sx, sz = _sx_sz()
    H = Hamiltonian(sx)
    rho0 = 0.5 * np.eye(2) + sz
    with pytest.raises(ValueError):
        H.evolve(rho0, np.zeros((2, 2)), [sz])  # times not 1-D
    with pytest.raises(ValueError):
        H.integrate_decaying(rho0, -1.0, [sz])
    with pytest.raises(ValueError):
        H.integrate_decaying(rho0, 1.0, [])
    with pytest.raises(ValueError):
        H.evolve(np.eye(3), np.linspace(0, 1, 3), [sz])  # incompatible rho0


def test_lindbladian_closed_forms():
    """tests/test_lindbladian.py:40-80 and :130-159 of the reference, through the GPU path."""
    from muspinsim_b200.hamiltonian import Hamiltonian
    from muspinsim_b200.lindbladian import Lindbladian
    from muspinsim_b200.spinsys import spin_operators

    sx, sy, sz = spin_operators(0.5)
    sp, sm = sx + 1j * sy, sx - 1j * sy
    H = Hamiltonian(sz)
    rho0 = 0.5 * np.eye(2) + sx
    t = np.linspace(0, 1, 100)
    L = Lindbladian.from_hamiltonian(H)
    assert np.allclose(L.evolve(rho0, t, sx)[:, 0], 0.5 * np.cos(2 * np.pi * t), atol=1e-11)
    tau = 2.0
    assert abs(L.integrate_decaying(rho0, tau, sx)[0] - 0.5 * tau / (1 + 4 * np.pi**2 * tau**2)) < 1e-12
    for g in [1.0, 2.0, 5.0, 10.0]:
        L = Lindbladian.from_hamiltonian(H, [(sx, g)])
        ap = -0.5 * np.pi * g + ((0.5 * np.pi * g) ** 2 - 4 * np.pi**2 + 0j) ** 0.5
        am = -0.5 * np.pi * g - ((0.5 * np.pi * g) ** 2 - 4 * np.pi**2 + 0j) ** 0.5
        A = ap * am / (am - ap)
        solx = np.real(0.5 * A * (np.exp(ap * t) / ap - np.exp(am * t) / am))
        assert np.allclose(L.evolve(rho0, t, sx)[:, 0], solx, atol=1e-10)
        sol = np.real(0.5 * A * tau * (1 / ((1 - ap * tau) * ap) - 1 / ((1 - am * tau) * am)))
        assert abs(L.integrate_decaying(rho0, tau, sx)[0] - sol) < 1e-11
        L = Lindbladian.from_hamiltonian(H, [(sp, 1.5 * g), (sm, 0.5 * g)])
        ev = L.evolve(rho0, t, [sx, sz])
        assert np.allclose(ev[:, 0], 0.5 * np.cos(2 * np.pi * t) * np.exp(-2 * np.pi * g * t), atol=1e-10)
        assert np.allclose(ev[:, 1], 0.25 * (1 - np.exp(-4 * np.pi * g * t)), atol=1e-10)
    with pytest.raises(ValueError):
        Lindbladian.from_hamiltonian(sz)
    with pytest.raises(ValueError):
        L.evolve(np.eye(3), t, sx)


def test_resident_handle_across_runners_with_different_couplings():
    """Fitting loop (fitting.py:126-135 builds a new runner per function evaluation): runners
    sharing a handle cache reuse ONE device handle, re-uploading only H0 / Z; results equal
    those of fresh handles, also when the runners are used alternately."""
    import numpy as np

    from muspinsim_b200 import ExperimentRunner, workloads

    cache = {}
    specs = []
    for scale in (1.0, 1.3, 0.6):
        spec = workloads.c2_hfine_powder(n_orient=9, nt=100, n_h=1)
        for c in spec["couplings"]:
            c["value"] = np.asarray(c["value"]) * scale
        specs.append(spec)
    runners = [ExperimentRunner(s, device=0, handle_cache=cache) for s in specs]
    got = [r.run() for r in runners]
    assert len(cache) == 1 and all(r.handle is runners[0].handle for r in runners)
    for s, g in zip(specs, got):
        assert np.max(np.abs(ExperimentRunner(s, device=0).run() - g)) < 1e-13
    assert np.max(np.abs(runners[0].run() - got[0])) < 1e-13  # first runner again after the others
    assert np.max(np.abs(got[0] - got[1])) > 1e-3


def test_non_hermitian_operator_gives_the_complex_expectation_value():
    """The reference's SpinOperator API allows non-Hermitian observables such as S+ (hamiltonian.py:
    87-107 returns the complex trace); here they are evaluated as <Oh> + i <Oa>.  Checked against a
    numpy evaluation of Tr(rho(t) O), also through the Lindbladian boundary; a non-Hermitian rho0
    is rejected."""
    from muspinsim_b200.hamiltonian import Hamiltonian
    from muspinsim_b200.lindbladian import Lindbladian
    from muspinsim_b200.spinsys import spin_operators

    sx, sy, sz = spin_operators(0.5)
    sp = sx + 1j * sy
    rng = np.random.default_rng(5)
    A = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    Hm = 0.5 * (A + A.conj().T)
    R = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    rho0 = R @ R.conj().T
    rho0 /= np.trace(rho0).real
    O = np.kron(sp, np.eye(2)) + 0.3 * np.kron(np.eye(2), sz)
    t = np.linspace(0.0, 0.7, 23)
    lam, U = np.linalg.eigh(Hm)
    want = []
    for tk in t:
        E = U @ np.diag(np.exp(-2j * np.pi * lam * tk)) @ U.conj().T
        want.append(np.trace(E @ rho0 @ E.conj().T @ O))
    want = np.array(want)
    H = Hamiltonian(Hm)
    got = H.evolve(rho0, t, [O, O.conj().T])
    assert got.shape == (23, 2)
    assert np.max(np.abs(got[:, 0] - want)) < 1e-11 and np.max(np.abs(got[:, 1] - want.conj())) < 1e-11
    assert np.max(np.abs(want.imag)) > 1e-3  # the case is not trivially real
    gl = Lindbladian.from_hamiltonian(H).evolve(rho0, t, [O])
    assert np.max(np.abs(gl[:, 0] - want)) < 1e-10
    with pytest.raises(ValueError):
        H.evolve(R, t, [O])  # rho0 is not Hermitian
    assert len(H._handles) == 1  # one resident device handle served all operators


def test_resident_configuration_table_in_a_fitting_style_loop():
    """Runners that share a handle cache and a configuration table (a fit varies couplings only):
    from the second run on, the expanded table is found on the device (no upload, no expansion
    kernel), and the results equal the CPU oracle's for every coupling set."""
    from muspinsim_b200 import ExperimentRunner, workloads
    from oracle import muspin_oracle

    cache = {}
    hits = []
    for scale in (1.0, 1.2, 0.8, 1.0):
        spec = workloads.c2_hfine_powder(n_orient=11, nt=100, n_h=1)
        for c in spec["couplings"]:
            c["value"] = np.asarray(c["value"]) * scale
        r = ExperimentRunner(spec, device=0, handle_cache=cache)
        got = r.run()
        assert np.max(np.abs(got - muspin_oracle.run_spec(spec))) < 1e-9
        hits.append(r.handle.phase_ms("axes_resident_hits"))
    assert len(cache) == 1
    assert hits == [0.0, 1.0, 2.0, 3.0]


def test_workspaces_of_destroyed_handles_stay_in_the_pool_until_trimmed():
    """Every fresh runner of a process gets its workspaces from the device memory pool the previous
    handle returned them to (no cudaMalloc / cudaFree of the whole workspace per runner); results are
    unchanged across the re-use, and musim_trim_pool gives the memory back to the driver."""
    import torch

    from muspinsim_b200 import ExperimentRunner, _lib, workloads

    spec = workloads.c2_hfine_powder(n_orient=50, nt=64, n_h=2)
    ref = None
    for _ in range(3):
        r = ExperimentRunner(spec, device=0)
        out = r.run()
        r.handle.close()
        if ref is None:
            ref = out
        assert np.max(np.abs(out - ref)) < 1e-13  # (the NUFFT warps draw configurations dynamically: last-bit differences)
    torch.cuda.synchronize()
    free_before = torch.cuda.mem_get_info(0)[0]
    _lib.trim_pool(0)
    free_after = torch.cuda.mem_get_info(0)[0]
    assert free_after >= free_before
    r = ExperimentRunner(spec, device=0)  # and the library still works after a trim
    assert np.max(np.abs(r.run() - ref)) < 1e-13
    with pytest.raises(_lib.MusimError):
        _lib.trim_pool(999)
