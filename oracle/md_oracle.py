"""CPU restatement of the reference's molecular dynamics driver on top of the CPU oracle.

TEST INFRASTRUCTURE (like everything under oracle/): imported only by tests/ and bench.py's cpu_baseline leg.

Follows grimme-lab/xtb v6.7.1: src/dynamic.f90:140-702 (md: masses 237-252, degrees of freedom 277-280, the loop
395-690 with steps 1-3 at 553-626), src/shake_module.f90:50-148 (init_shake) and 150-245 (do_shake: sequential
corrections from the deviations evaluated at the start of each cycle), src/rmrottr.f90:19-125, src/dynamic.f90:767-776
(ekinet).  Parity: unpinned by reference values (the reference's tests hold none for MD and its start velocities come
from random_number); the integrator is checked through its invariants (constraint lengths, zero linear / angular momentum,
energy conservation without thermostat) and is the checker of the CUDA propagation on identical start velocities.
"""
import numpy as np

from .oracle import Oracle

AMUTOAU = 1.660539040e-27 * (1.0 / 9.10938356e-31)   # src/mctc/convert.f90:51-55
FSTOAU = 1.0e-15 / 2.4188843265857e-17                # src/mctc/convert.f90:58-60
KB = 3.166808578545117e-06                            # src/mctc/mctc_constants.f90:34
AUTOAA = 0.52917726

# src/param/atomicmass.f90:37-67 and src/param/atomicrad.f90:42-46 (the first 36 elements cover the test systems),
# src/shake_module.f90:36-41
MASS = [1.00794075, 4.00260193, 6.94003660, 9.01218307, 10.81102805, 12.01073590, 14.00670321, 15.99940492, 18.99840316,
        20.18004638, 22.98976928, 24.30505162, 26.98153853, 28.08549871, 30.97376200, 32.06478741, 35.45293758, 39.94779856,
        39.09830091, 40.07802251, 44.95590828, 47.86674496, 50.94146504, 51.99613176, 54.93804391, 55.84514443, 58.93319429,
        58.69334711, 63.54603995, 65.37778253, 69.72306607, 72.62755016, 74.92159457, 78.95938856, 79.90352778, 83.79800000]
ATOMIC_RAD_AA = [0.32, 0.37, 1.30, 0.99, 0.84, 0.75, 0.71, 0.64, 0.60, 0.62, 1.60, 1.40, 1.24, 1.14, 1.09, 1.04, 1.00, 1.01,
                 2.00, 1.74, 1.59, 1.48, 1.44, 1.30, 1.29, 1.24, 1.18, 1.17, 1.22, 1.20, 1.23, 1.20, 1.20, 1.18, 1.17, 1.16]
METAL = [0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0]


def init_shake(numbers, xyz, wbo, mode=2):
    """Constraint list [(i, j)] and squared reference distances (src/shake_module.f90:50-148)."""
    n = len(numbers)
    cons = []
    if mode == 1:
        for i in range(n):
            if numbers[i] != 1:
                continue
            r2 = np.sum((xyz - xyz[i]) ** 2, axis=1)
            r2[i] = np.inf
            jmin = int(np.argmin(r2))
            if numbers[jmin] == 1:
                if jmin > i:
                    cons.append((i, jmin))
            else:
                cons.append((i, jmin))
    elif mode == 2:
        for i in range(n):
            for j in range(i + 1, n):
                rij = float(np.sum((xyz[i] - xyz[j]) ** 2))
                aatoau = 1.0 / AUTOAA
                rco = (ATOMIC_RAD_AA[numbers[j] - 1] * aatoau + ATOMIC_RAD_AA[numbers[i] - 1] * aatoau) * AUTOAA
                if not 0.52917726 * np.sqrt(rij) < 1.2 * rco:
                    continue
                metalbond = METAL[numbers[i] - 1] == 1 or METAL[numbers[j] - 1] == 1
                wthr = 0.1 if metalbond else 0.5
                if not wbo[i, j] > wthr:
                    continue
                if numbers[i] in (3, 4) or numbers[j] in (3, 4):
                    metalbond = False
                if not metalbond:
                    cons.append((i, j))
    dist2 = np.array([np.sum((xyz[i] - xyz[j]) ** 2) for i, j in cons])
    return cons, dist2


def do_shake(cons, dist2, xyzo, xyz, vel, acc, mass, tstep, maxcyc=250, tol=1.0e-7):
    """src/shake_module.f90:150-245; returns the number of cycles (negative: not converged, nothing applied)."""
    if not cons:
        return 0
    xyzt = xyz.copy()
    dro = np.array([xyzo[i] - xyzo[j] for i, j in cons])
    icyc, conv = 0, False
    while True:
        dr = np.array([xyzt[i] - xyzt[j] for i, j in cons])
        d2 = np.sum(dr ** 2, axis=1)
        conv = np.max(np.abs(d2 - dist2) / dist2) < tol
        if not conv:
            for c, (i, j) in enumerate(cons):
                rmi, rmj = 1.0 / mass[i], 1.0 / mass[j]
                denom = 2.0 * (rmi + rmj) * np.dot(dr[c], dro[c])
                g = (dist2[c] - d2[c]) / denom
                xyzt[i] += rmi * g * dro[c]
                xyzt[j] -= rmj * g * dro[c]
        icyc += 1
        if conv or icyc > maxcyc:
            break
    if conv:
        vel += (xyzt - xyz) / tstep
        acc += (xyzt - xyz) / tstep ** 2
        xyz[:] = xyzt
        return icyc
    return -icyc


def rmrottr(mass, vel, xyz):
    """src/rmrottr.f90:19-125: remove the overall translation and rotation from the velocities."""
    tmass = mass.sum()
    com = (mass[:, None] * xyz).sum(axis=0) / tmass
    c = xyz - com
    angmom = (mass[:, None] * np.cross(c, vel)).sum(axis=0)
    inertia = np.zeros((3, 3))
    for m, r in zip(mass, c):
        inertia += m * (np.dot(r, r) * np.eye(3) - np.outer(r, r))
    omega = np.linalg.inv(inertia) @ angmom
    rlm = (mass[:, None] * vel).sum(axis=0)
    return vel - rlm / tmass - np.cross(omega, c)


def ekinet(vel, mass):
    return 0.5 * float(np.sum(mass * np.sum(vel ** 2, axis=1)))


def start_velocities(numbers, mass, nfree, temperature, thermostat, signs):
    """mdinitu, src/dynamic.f90:850-890, with the signs given by the caller (the reference draws them from random_number)."""
    n = len(numbers)
    f = 1.0 if thermostat else 2.0
    eperat = f * temperature * 0.5 * KB * nfree / (3.0 * n)
    v = np.sqrt(2.0 * eperat / mass) * np.where(np.asarray(numbers) == 1, 2.0, 1.0)
    return v[:, None] * signs


class MDOracle:
    def __init__(self, numbers, positions, charge=0.0, uhf=0, tstep_fs=4.0, temperature=298.15, thermostat=True, shake=2,
                 hmass=4, velocities=None, signs=None, solvent=None, threads=None):
        self.numbers = np.asarray(numbers, dtype=np.int32)
        self.xyz = np.array(positions, dtype=np.float64).reshape(-1, 3)
        self.o = Oracle(self.numbers, self.xyz, charge=charge, uhf=uhf, threads=threads)
        if solvent is not None:
            self.o.set_solvent(**solvent)
        n = len(self.numbers)
        assert self.o.singlepoint() == 0            # the single point of the main program (bond orders for init_shake)
        self._restart()
        wbo = self.o.get("wbo").reshape(n, n)
        self.mass = np.array([MASS[z - 1] for z in self.numbers]) * AMUTOAU
        if hmass > 0:
            self.mass[self.numbers == 1] = float(hmass) * AMUTOAU
        self.cons, self.dist2 = init_shake(self.numbers, self.xyz, wbo, shake) if shake else ([], np.zeros(0))
        self.nfree = 3.0 * n - (len(self.cons) if shake else 0)
        self.shake = bool(shake)
        self.thermostat = bool(thermostat)
        self.tstep0, self.tstep, self.taut, self.tinit = tstep_fs, tstep_fs * FSTOAU, 500.0, temperature
        if velocities is not None:
            self.velo = np.array(velocities, dtype=np.float64).reshape(-1, 3)
        else:
            self.velo = start_velocities(self.numbers, self.mass, self.nfree, temperature, thermostat, signs)
        self.first = True
        self.niter = []

    def _restart(self):
        self.o.set_restart(self.o.get("qsh"), self.o.get("dipm"), self.o.get("qp"))

    def _sp(self):
        self.o.set_positions(self.xyz)
        assert self.o.singlepoint() == 0
        self._restart()
        self.niter.append(self.o.info["niter"])
        return self.o.energies["e_total"], self.o.get("gradient").reshape(-1, 3)

    def run(self, nsteps):
        """Returns arrays epot, ekin, temperature of the steps (dynamic.f90:395-690)."""
        if self.first:
            self._sp()                      # dynamic.f90:383-385
            self.first = False
        ep, ek, te = [], [], []
        m = self.mass[:, None]
        for _ in range(nsteps):
            epot, grd = self._sp()
            acc = -grd / m
            xyzo = self.xyz.copy()
            veln = self.velo + 0.5 * self.tstep * acc
            T = 2.0 * ekinet(veln, self.mass) / self.nfree / KB
            xlam2 = np.sqrt(1.0 + (self.tstep0 / self.taut) * (self.tinit / T - 1.0))
            vel = self.velo + acc * self.tstep
            if self.thermostat:
                vel = xlam2 * vel
            self.xyz = xyzo + vel * self.tstep
            veln = 0.5 * (self.velo + vel)
            ekin = ekinet(veln, self.mass)
            T = 2.0 * ekin / self.nfree / KB
            if self.shake:
                do_shake(self.cons, self.dist2, xyzo, self.xyz, vel, acc, self.mass, self.tstep)
            self.velo = rmrottr(self.mass, vel, self.xyz)
            ep.append(epot); ek.append(ekin); te.append(T)
        return np.array(ep), np.array(ek), np.array(te)
