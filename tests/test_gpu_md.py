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


def test_md_start_state_from_the_reference_rules(xb, kat):
    """Start state of an MD run pinned on what the reference's source prescribes, without the Python MD restatement:
    masses (src/param/atomicmass.f90:37-39 times amutoau of src/mctc/convert.f90:51-55, hydrogens at hmass),
    the constraint list of init_shake (src/shake_module.f90:77-148: X-H bonds to the nearest neighbour, or every pair within
    1.2 x the covalent radii whose Wiberg bond order exceeds 0.5), the degrees of freedom 3N - ncons (src/dynamic.f90:292-294)
    and the start velocities of mdinitu (src/dynamic.f90:304-308, 850-877): every component has the magnitude
    sqrt(2 e / m) with e = T kB nfreedom / (2 * 3N), doubled for hydrogen."""
    s = kat["caffeine"]                                   # C8 H10 N4 O2: 25 covalent bonds, 10 of them C-H
    Z, xyz = np.asarray(s["numbers"]), np.asarray(s["xyz"], dtype=float).reshape(-1, 3)
    n = len(Z)
    amutoau = 1.660539040e-27 / 9.10938356e-31
    kB = 3.166808578545117e-06
    amu = {1: 1.00794075, 6: 12.01073590, 7: 14.00670321, 8: 15.99940492}
    # bonds from the geometry alone (1.2 x covalent radii of the reference's atomicRad table are ample here: use 1.3 A for
    # X-H and 1.7 A for pairs of C / N / O, which separates bonded from non-bonded pairs of caffeine by > 0.5 A)
    dist = np.linalg.norm(xyz[:, None] - xyz[None], axis=2) * 0.52917726
    isH = Z == 1
    bonded = (dist < np.where(isH[:, None] | isH[None, :], 1.3, 1.7)) & ~np.eye(n, dtype=bool)
    assert bonded.sum() // 2 == 25 and (bonded & (isH[:, None] ^ isH[None, :])).sum() // 2 == 10
    for mode, hmass, thermostat in ((2, 4, True), (1, 0, False)):
        b = xb.BatchCalculator([Z])
        ncons = b.md_init(xyz, temperature=298.15, thermostat=thermostat, shake=mode, hmass=hmass, seed=11)
        assert ncons == (25 if mode == 2 else 10)
        mass = b.array("md_mass")
        ref_mass = np.array([(hmass if (z == 1 and hmass > 0) else amu[int(z)]) * amutoau for z in Z])
        assert np.max(np.abs(mass - ref_mass) / ref_mass) < 1e-12
        nfree = 3.0 * n - ncons
        ekin = (1.0 if thermostat else 2.0) * 298.15 * 0.5 * kB * nfree
        eperat = ekin / (3.0 * n)
        v = b.array("velocities").reshape(n, 3)
        ref_v = np.sqrt(2.0 * eperat / ref_mass) * np.where(isH, 2.0, 1.0)
        assert np.max(np.abs(np.abs(v) - ref_v[:, None]) / ref_v[:, None]) < 1e-12
        assert set(np.sign(v).ravel().tolist()) == {-1.0, 1.0}
