"""Invariants of the CPU restatement of the reference's MD driver (oracle/md_oracle.py; src/dynamic.f90:140-702,
src/shake_module.f90, src/rmrottr.f90).  The reference holds no known-answer values for MD."""
import numpy as np

from oracle.md_oracle import MDOracle, KB


def _signs(n, seed):
    return np.where(np.random.default_rng(seed).random((n, 3)) > 0.5, -1.0, 1.0)


def test_shake_constraints_and_momenta(kat):
    s = kat["caffeine"]
    n = len(s["numbers"])
    md = MDOracle(s["numbers"], s["xyz"], signs=_signs(n, 1))   # xtb defaults: 4 fs, NVT, SHAKE on all bonds, H mass 4
    assert len(md.cons) == 25                                   # every bond of caffeine (24 atoms, 2 rings)
    ep, ek, te = md.run(5)
    d2 = np.array([np.sum((md.xyz[i] - md.xyz[j]) ** 2) for i, j in md.cons])
    assert np.max(np.abs(d2 - md.dist2) / md.dist2) < 1e-7       # tolshake
    p = (md.mass[:, None] * md.velo).sum(axis=0)
    com = (md.mass[:, None] * md.xyz).sum(axis=0) / md.mass.sum()
    ang = (md.mass[:, None] * np.cross(md.xyz - com, md.velo)).sum(axis=0)
    assert np.max(np.abs(p)) < 1e-9 and np.max(np.abs(ang)) < 1e-8
    assert np.all(np.isfinite(ep)) and np.all(te > 0)


def test_energy_conservation_without_thermostat(kat):
    s = kat["h2o_scc"]
    n = len(s["numbers"])
    md = MDOracle(s["numbers"], s["xyz"], tstep_fs=0.1, temperature=300.0, thermostat=False, shake=0, hmass=0,
                  signs=_signs(n, 3))
    ep, ek, te = md.run(60)
    etot = (ep + ek)[1:]       # the first step still carries the overall translation / rotation of the start velocities
    assert np.max(np.abs(etot - etot[0])) < 2e-5
    assert np.ptp(ep) > 1e-4   # the atoms do move
