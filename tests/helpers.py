"""Test helpers.  `OracleHandle` stands in for the CUDA handle so that the host-side logic
(configuration table, mode selection, sharding, result layout) can be tested without a GPU:
it evaluates each configuration with the CPU oracle from the ALREADY ROTATED (B, p, T) the host
code produced.  Test infrastructure only."""
import numpy as np

from oracle import muspin_oracle as mo


class OracleHandle:
    def __init__(self, spec):
        self.sys = mo.build_system(spec)
        self.calls = []
        self.launches = 0

    def set_option(self, k, v):
        pass

    def run_axes_host(self, mode, n_cfg, first, step, axes, times, tau, out):
        from muspinsim_b200.configs import expand_from_axes

        B, p, T, w, slot = expand_from_axes(axes, n_cfg, first, step)
        return self.run_host(mode, B, p, T, w, slot, times, tau, out)

    def run_host(self, mode, B, p, T, w, slot, times, tau, out):
        s = self.sys
        self.calls.append((mode, len(B)))
        for c in range(len(B)):
            H = s.H0 + mo.zeeman_matrix(s, B[c])
            O = s.muon_operator(p[c])
            if mode in (3, 4):
                L = mo.superop_lindbladian(H, mo.dissipation_operators(s, B[c], T[c]))
                rho0 = mo.rho0_matrix(s, B[c], p[c], T[c])
                if mode == 3:
                    val = mo.lindblad_evolve(L, rho0, times, O)
                else:
                    val = np.array([mo.lindblad_integrate(L, rho0, tau, O) / tau])
            elif mode == 1:
                d_other = s.d // s.dims[s.mu_i]
                # maximally mixed other spins: rho0 = rho_mu (x) 1/d_other in the muon's slot
                rho0 = mo.rho0_matrix(s, np.zeros(3), p[c], np.inf)
                val = mo.evolve_vectorised(H, rho0, times, O)
            elif mode == 0:
                val = mo.evolve_vectorised(H, mo.rho0_matrix(s, B[c], p[c], T[c]), times, O)
            elif mode == 2:
                val = np.array([mo.integrate_decaying(H, mo.rho0_matrix(s, B[c], p[c], T[c]), tau, O) / tau])
            elif mode == 5:
                rho0 = mo.rho0_matrix(s, np.zeros(3), p[c], np.inf)
                val = np.array([mo.integrate_decaying(H, rho0, tau, O) / tau])
            else:
                raise ValueError(mode)
            out[slot[c], :] += w[c] * np.real(val)
        return out
