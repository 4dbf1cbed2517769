"""GPU parity of the molecular dynamics propagation (src/dynamic.f90:140-702, src/shake_module.f90,
src/rmrottr.f90) through the C ABI (xtbb200_mdInit / xtbb200_mdRun) against the CPU restatement oracle/md_oracle.py on
identical start velocities: same SHAKE constraint lists, energies within 1e-7 Eh per step, positions and velocities after
the trajectory within 1e-6, identical SCC iteration counts are implied by the energies."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xb():
    import xtb_b200
    if xtb_b200.device_count() < 1:
        pytest.fail("no CUDA device: the engine has no CPU fallback")
    return xtb_b200


def _signs(n, seed):
    return np.where(np.random.default_rng(seed).random((n, 3)) > 0.5, -1.0, 1.0)


@pytest.mark.parametrize("solvent", [None, dict(solvent="water", model="alpb")])
def test_md_trajectory_matches_oracle(xb, kat, solvent):
    """xtb's default MD settings (4 fs, Berendsen 298.15 K, SHAKE on all bonds, H mass 4) on caffeine and on a
    rivaroxaban conformer as two replicas of one batch, in the gas phase and in ALPB water."""
    from oracle.md_oracle import MDOracle
    mols = [kat["caffeine"], kat["rivaroxaban"]]
    nsteps = 6
    orcs = [MDOracle(m["numbers"], m["xyz"], signs=_signs(len(m["numbers"]), 7 + k), solvent=solvent)
            for k, m in enumerate(mols)]
    b = xb.BatchCalculator([m["numbers"] for m in mols])
    if solvent:
        b.set_solvent(**solvent)
    ncons = b.md_init(np.concatenate([m["xyz"] for m in mols]), velocities=np.concatenate([o.velo for o in orcs]))
    assert ncons == sum(len(o.cons) for o in orcs)
    assert np.max(np.abs(b.array("md_mass") - np.concatenate([o.mass for o in orcs]))) < 1e-9
    r = b.md_run(nsteps)
    assert r["steps"] == nsteps
    xyz, vel = b.array("xyz").reshape(-1, 3), b.array("velocities").reshape(-1, 3)
    off = 0
    for k, o in enumerate(orcs):
        ep, ek, te = o.run(nsteps)
        n = len(o.numbers)
        assert np.max(np.abs(r["epot"][:, k] - ep)) < 1e-7, k
        assert np.max(np.abs(r["ekin"][:, k] - ek)) < 1e-8, k
        assert np.max(np.abs(r["temperature"][:, k] - te)) < 1e-3, k
        assert np.max(np.abs(xyz[off:off + n] - o.xyz)) < 1e-6, k
        assert np.max(np.abs(vel[off:off + n] - o.velo)) < 1e-8, k
        off += n
    # a second call continues the trajectory
    r2 = b.md_run(2)
    ep, ek, te = orcs[0].run(2)
    assert np.max(np.abs(r2["epot"][:, 0] - ep)) < 1e-7


def test_md_invariants_at_ensemble_size(xb, kat):
    """64 replicas of rivaroxaban with different start velocities: constraint lengths kept to SHAKE's tolerance, no overall
    linear or angular momentum, finite temperatures; X-H only SHAKE with the physical hydrogen mass at 1 fs."""
    s = kat["rivaroxaban"]
    n, nrep = len(s["numbers"]), 64
    b = xb.BatchCalculator([s["numbers"]] * nrep)
    x0 = np.concatenate([s["xyz"]] * nrep)
    ncons = b.md_init(x0, tstep_fs=1.0, temperature=400.0, shake=1, hmass=0, seed=11)
    nh = int(np.sum(np.asarray(s["numbers"]) == 1))
    assert ncons == nrep * nh
    v0 = b.array("velocities").reshape(nrep, n, 3)
    assert np.max(np.abs(v0[0] - v0[1])) > 0          # replicas differ
    r = b.md_run(5)
    assert r["steps"] == 5 and np.all(np.isfinite(r["epot"])) and np.all(r["temperature"] > 0)
    xyz = b.array("xyz").reshape(nrep, n, 3)
    vel = b.array("velocities").reshape(nrep, n, 3)
    mass = b.array("md_mass").reshape(nrep, n)
    x00 = s["xyz"]
    hs = [i for i in range(n) if s["numbers"][i] == 1]
    for rep in (0, 17, 63):
        for i in hs:
            j = int(np.argsort(np.sum((x00 - x00[i]) ** 2, axis=1))[1])
            d0 = np.sum((x00[i] - x00[j]) ** 2)
            d1 = np.sum((xyz[rep, i] - xyz[rep, j]) ** 2)
            assert abs(d1 - d0) / d0 < 1e-7
        p = (mass[rep][:, None] * vel[rep]).sum(axis=0)
        com = (mass[rep][:, None] * xyz[rep]).sum(axis=0) / mass[rep].sum()
        ang = (mass[rep][:, None] * np.cross(xyz[rep] - com, vel[rep])).sum(axis=0)
        assert np.max(np.abs(p)) < 1e-9 and np.max(np.abs(ang)) < 1e-7
