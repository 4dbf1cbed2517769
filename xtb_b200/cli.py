"""Thin command line front end: `python -m xtb_b200 mol.xyz --gfn 2 --grad` (the `xtb mol.xyz --gfn 2 --grad` call of
BASELINE configs[0]) on the CUDA engine, through the xtb.h C ABI.

Covers what the single point path of the reference's driver needs (src/prog/main.F90:224-399,721,1020-1048): geometry
readers for xyz and Turbomole coord files (the reference delegates them to mctc-lib, which is not part of its tree; the
formats are the published ones), the run-type flags of a single point with gradient, and the result files `energy`,
`gradient` (src/io/writer/turbomole.f90:32-123), `charges` (src/main/property.F90:189-192, one f14.8 per atom) and `wbo`
(src/main/property.F90:228-231,869-881, pairs above 0.1); implicit solvation (`--alpb / --gbsa SOLVENT`); and molecular
dynamics (`--md`, the $md block's time / step / temp / dump / shake / hmass / nvt as flags): trajectory `xtb.trj` (xyz frames
with the energy in the comment line, src/dynamic.f90:536-545, src/io/writer.f90:64-75), `mdrestart`
(src/dynamic.f90:814-827), the screen table (src/dynamic.f90:522-533) and the `xtbmdok` marker.  Everything else the
reference's CLI offers (optimisation, hessian, other Hamiltonians, metadynamics) is refused with an error instead of being
ignored.
"""
import argparse
import sys
import time

import numpy as np

SYMBOLS = ("H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr "
           "Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu "
           "Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn").split()
AATOAU = 1.0 / 0.529177210903   # mctc-lib's conversion (CODATA 2018), used by its geometry readers


def _number(sym):
    s = sym.strip().capitalize()
    if s.isdigit():
        return int(s)
    s = "".join(ch for ch in s if ch.isalpha())
    if s not in SYMBOLS:
        raise ValueError("unknown element symbol '%s'" % sym)
    return SYMBOLS.index(s) + 1


def read_xyz(text):
    """Xmol xyz: atom count, comment line, then `symbol x y z` in Angstrom."""
    lines = text.splitlines()
    n = int(lines[0].split()[0])
    numbers, xyz = [], []
    for ln in lines[2:2 + n]:
        t = ln.split()
        numbers.append(_number(t[0]))
        xyz.append([float(x.replace("d", "e").replace("D", "e")) for x in t[1:4]])
    if len(numbers) != n:
        raise ValueError("xyz file ends before %d atoms were read" % n)
    return np.array(numbers, dtype=np.int32), np.array(xyz) * AATOAU


def read_coord(text):
    """Turbomole coord: `$coord [angs]` block of `x y z symbol` lines (bohr unless `angs`), ended by the next `$` group."""
    numbers, xyz, scale, inside = [], [], 1.0, False
    for ln in text.splitlines():
        s = ln.strip()
        if not s or s.startswith("#"):
            continue
        if s.startswith("$"):
            if inside:
                break
            if s.lower().startswith("$coord"):
                inside = True
                if "angs" in s.lower():
                    scale = AATOAU
                elif "frac" in s.lower():
                    raise ValueError("fractional coordinates need a lattice: periodic systems are outside the engine's scope")
            continue
        if inside:
            t = s.split()
            xyz.append([float(x) for x in t[:3]])
            numbers.append(_number(t[3]))
    if not numbers:
        raise ValueError("no $coord block found")
    return np.array(numbers, dtype=np.int32), np.array(xyz) * scale


def read_geometry(path):
    with open(path) as f:
        text = f.read()
    if path.lower().endswith(".xyz"):
        return read_xyz(text)
    if "$coord" in text.lower():
        return read_coord(text)
    return read_xyz(text)


def _fortran_e(x, width, digits):
    """Fortran ESw.d: one non-zero digit before the point (Python's %E gives the same normalisation)."""
    return ("%" + "%d.%dE" % (width, digits)) % x


def write_turbomole(numbers, xyz, energy, gradient, iter_number=1, energy_file="energy", gradient_file="gradient"):
    """src/io/writer/turbomole.f90:96-123: `$energy` and `$grad` groups, each closed by `$end`."""
    with open(energy_file, "w") as f:
        f.write("$energy\n")
        f.write("%6d%18.11f%18.11f%18.11f\n" % (1, energy, energy, energy))
        f.write("$end\n")
    if gradient is None:
        return
    with open(gradient_file, "w") as f:
        f.write("$grad\n")
        f.write("  cycle = %6d    SCF energy =%18.11f   |dE/dxyz| =%10.6f\n" % (iter_number, energy, np.linalg.norm(gradient)))
        for z, r in zip(numbers, xyz):
            f.write("%20.14f  %20.14f  %20.14f      %-2s\n" % (r[0], r[1], r[2], SYMBOLS[z - 1].lower()))
        for g in gradient:
            f.write("".join(_fortran_e(v, 22, 13) for v in g) + "\n")
        f.write("$end\n")


def read_turbomole_gradient(path):
    """Inverse of write_turbomole for the last `cycle` of a gradient file: (energy, xyz, gradient)."""
    with open(path) as f:
        lines = [ln for ln in f.read().splitlines() if ln.strip() and not ln.startswith("$")]
    energy = float(lines[0].split("SCF energy =")[1].split()[0])
    body = lines[1:]
    n = len(body) // 2
    xyz = np.array([[float(x) for x in ln.split()[:3]] for ln in body[:n]])
    grad = np.array([[float(x.replace("D", "E")) for x in ln.split()[:3]] for ln in body[n:2 * n]])
    return energy, xyz, grad


def write_charges(q, path="charges"):
    with open(path, "w") as f:
        for v in q:
            f.write("%14.8f\n" % v)


def write_wbo(wbo, path="wbo", thr=0.1):
    n = wbo.shape[0]
    with open(path, "w") as f:
        for i in range(n):
            for j in range(i):
                if wbo[j, i] > thr:
                    f.write("%12d%12d   %.15f\n" % (j + 1, i + 1, wbo[j, i]))


def _fortran_d(x, width, digits):
    """Fortran Dw.d: mantissa in [0.1, 1), e.g. 0.37719066328018D+00."""
    if x == 0.0:
        return ("0." + "0" * digits + "D+00").rjust(width)
    e = int(np.floor(np.log10(abs(x)))) + 1
    ms = "%.*f" % (digits, abs(x) / 10.0 ** e)
    if ms.startswith("1."):      # rounded up to 1.0
        e += 1
        ms = "%.*f" % (digits, abs(x) / 10.0 ** e)
    return (("-" if x < 0 else "") + ms + "D%+03d" % e).rjust(width)


def write_trj_frame(f, numbers, xyz_bohr, energy, gnorm):
    """One xyz frame of xtb.trj (positions in Angstrom; comment line as writeMolecule builds it, src/io/writer.f90:64-75)."""
    f.write("%d\n energy: %s gnorm: %s xtb: b200\n" % (len(numbers), ("%20.12f" % energy).strip(), ("%20.12f" % gnorm).strip()))
    for z, r in zip(numbers, np.asarray(xyz_bohr) / AATOAU):
        f.write("%-4s %20.14f %20.14f %20.14f\n" % (SYMBOLS[z - 1], r[0], r[1], r[2]))
    f.flush()


def write_mdrestart(xyz, velo, path="mdrestart"):
    """src/dynamic.f90:814-827: '-1.0', then positions and velocities (atomic units) as 6D22.14 per atom."""
    with open(path, "w") as f:
        f.write(" -1.0\n")
        for r, v in zip(xyz, velo):
            f.write("".join(_fortran_d(x, 22, 14) for x in (*r, *v)) + "\n")


def read_mdrestart(path="mdrestart"):
    """src/dynamic.f90:829-844."""
    with open(path) as f:
        lines = f.read().splitlines()[1:]
    vals = np.array([[float(ln[22 * k:22 * (k + 1)].replace("D", "E").replace("d", "e")) for k in range(6)] for ln in lines if ln.strip()])
    return vals[:, :3].copy(), vals[:, 3:].copy()


def write_xtbrestart(qsh, dipm, qp, nat, nel, nopen, path="xtbrestart", gfn_method=2):
    """The binary restart file of the reference (writeRestart, src/restart.f90:182-200): Fortran sequential unformatted
    records (4-byte length markers, the default of the compilers xtb is built with): six int64 (method, method, atoms, shells,
    electrons, unpaired electrons), then the shell charges qsh(nshell), the atomic dipoles dipm(3, n) and the atomic
    quadrupoles qp(6, n) as float64."""
    qsh = np.ascontiguousarray(qsh, dtype="<f8").ravel()
    dipm = np.ascontiguousarray(dipm, dtype="<f8").ravel()
    qp = np.ascontiguousarray(qp, dtype="<f8").ravel()
    if dipm.size != 3 * nat or qp.size != 6 * nat:
        raise ValueError("dipm / qp do not match the number of atoms")
    head = np.array([gfn_method, gfn_method, nat, qsh.size, nel, nopen], dtype="<i8")
    with open(path, "wb") as f:
        for rec in (head, qsh, dipm, qp):
            raw = rec.tobytes()
            mark = np.array([len(raw)], dtype="<i4").tobytes()
            f.write(mark + raw + mark)


def read_xtbrestart(nat, nshell, nel=None, nopen=None, path="xtbrestart", gfn_method=2, warn=None):
    """readRestart (src/restart.f90:33-86): None when the file is missing, unreadable or written for another number of
    atoms / shells (`Dimension missmatch`); a different method, electron count or multiplicity only warns, as in the
    reference.  Returns (qsh, dipm[n][3], qp[n][6])."""
    warn = warn or (lambda msg: None)

    def record(f):
        m = f.read(4)
        if len(m) < 4:
            raise EOFError
        n = int(np.frombuffer(m, dtype="<i4")[0])
        raw = f.read(n)
        if len(raw) < n or f.read(4) != m:
            raise EOFError
        return raw

    try:
        with open(path, "rb") as f:
            head = np.frombuffer(record(f), dtype="<i8")
            if head.size != 6:
                raise EOFError
            iver, _, n8, nshell8, nel8, nopen8 = (int(x) for x in head)
            if iver != gfn_method:
                warn("Version number missmatch in restart file.")
            if nel is not None and nel8 != nel:
                warn("Number of electron missmatch in restart file.")
            if nopen is not None and nopen8 != nopen:
                warn("Multiplicity missmatch in restart file.")
            if n8 != nat or nshell8 != nshell:
                warn("Dimension missmatch in restart file.")
                return None
            qsh = np.frombuffer(record(f), dtype="<f8").copy()
            if qsh.size != nshell:
                raise EOFError
            if gfn_method > 1 and iver > 1:
                dipm = np.frombuffer(record(f), dtype="<f8").copy().reshape(nat, 3)
                qp = np.frombuffer(record(f), dtype="<f8").copy().reshape(nat, 6)
            else:
                dipm, qp = np.zeros((nat, 3)), np.zeros((nat, 6))
            return qsh, dipm, qp
    except FileNotFoundError:
        return None
    except (EOFError, ValueError):
        warn("Dimension missmatch in restart file.")
        return None


def read_xcontrol_md(text):
    """The `$md` block of an xcontrol input (keys of set_md, src/set_module.f90:1975-2072: temp, time, dump / dumpxyz /
    dumptrj, step, hmass, shake, nvt, restart; `key=value`, `#` comments, first occurrence wins as in the reference).
    Returns a dict with the argparse destination names; keys this driver does not implement are reported in 'ignored'."""
    keys = {"temp": ("temp", float), "time": ("time", float), "dump": ("dump", float), "dumpxyz": ("dump", float),
            "dumptrj": ("dump", float), "step": ("step", float), "hmass": ("hmass", int), "shake": ("shake", int),
            "nvt": ("nvt", None), "restart": ("restart", None)}
    out, ignored, inside = {}, [], False
    for raw in text.splitlines():
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        if line.startswith("$"):
            inside = line[1:].split()[0].lower() == "md" if len(line) > 1 else False
            continue
        if not inside:
            continue
        key, sep, val = line.replace(":", "=", 1).partition("=")
        key, val = key.strip().lower(), val.strip()
        if not sep or not val:
            continue
        if key not in keys:
            ignored.append(key)
            continue
        dest, conv = keys[key]
        if dest in out:
            continue
        if conv is None:
            v = val.lower()
            if v in ("true", "yes", "on", "1", ".true."):
                out[dest] = True
            elif v in ("false", "no", "off", "0", ".false."):
                out[dest] = False
            else:
                raise ValueError("$md: '%s' is not a boolean for %s" % (val, key))
        else:
            out[dest] = conv(float(val)) if conv is int else conv(val)
    if "shake" in out and out["shake"] not in (0, 1, 2):
        raise ValueError("$md: shake=%d is not available (0, 1, 2)" % out["shake"])
    out["ignored"] = ignored
    return out


def run_md(args, numbers, xyz, out=sys.stdout):
    """The MD run type (src/dynamic.f90:140-702 through xtbb200_mdInit / xtbb200_mdRun): one replica, frames every
    --dump fs starting with the first step, screen line every 200 steps, mdrestart and xtbmdok at the end."""
    import os

    import xtb_b200
    b = xtb_b200.BatchCalculator([numbers], charges=[args.chrg], uhf=[args.uhf], accuracy=args.acc, etemp=args.etemp,
                                 maxiter=args.iterations)
    if args.alpb:
        b.set_solvent(args.alpb, model="alpb")
    elif args.gbsa:
        b.set_solvent(args.gbsa, model="gbsa")
    velo = None
    if args.restart:
        xyz, velo = read_mdrestart()
    if os.path.exists("xtbmdok"):
        os.remove("xtbmdok")
    ncons = b.md_init(xyz, tstep_fs=args.step, temperature=args.temp, thermostat=not args.nve, shake=args.shake, hmass=args.hmass,
                      velocities=velo, seed=args.seed)
    nmax = int(args.time * 1000.0 / args.step)
    dumpstep = max(1, int(args.dump / args.step))
    if not args.silent:
        out.write(" MD time /ps        :%8.2f\n dt /fs             :%8.2f\n temperature /K     :%8.2f\n max steps          :%6d\n"
                  % (args.time, args.step, args.temp, nmax))
        out.write(" dumpstep(trj) /fs  :%8.2f%6d\n" % (args.dump, dumpstep))
        if args.hmass > 0:
            out.write(" H atoms mass (amu) :%6d\n" % args.hmass)
        out.write(" SHAKE %s\n" % ("on. # bonds  : %d all: %s" % (ncons, args.shake == 2) if args.shake else "off"))
        if not args.nve:
            out.write(" Berendsen THERMOSTAT on\n")
        out.write("   step   time/ps    <Epot>      Ekin   <T>   T     Etot\n")
    k, epav, tav, blown = 0, 0.0, 0.0, False
    with open("xtb.trj", "w") as trj:
        while k < nmax and not blown:
            frame_xyz = b.array("xyz").reshape(-1, 3)          # geometry of step k + 1: dumped with its energy
            chunk = min(dumpstep, nmax - k)
            r = b.md_run(1)
            blown = r["steps"] < 1
            if r["steps"]:
                write_trj_frame(trj, numbers, frame_xyz, r["epot"][0, 0], b.result(0)["gnorm"])
                rest = b.md_run(chunk - 1) if chunk > 1 else dict(steps=0, epot=np.zeros((0, 1)), ekin=np.zeros((0, 1)),
                                                                     temperature=np.zeros((0, 1)))
                blown = rest["steps"] < chunk - 1
                for part in (r, rest):
                    for i in range(part["steps"]):
                        k += 1
                        epav += part["epot"][i, 0]
                        tav += part["temperature"][i, 0]
                        if not args.silent and (k == 1 or k % 200 == 0):
                            out.write("%7d%8.2f%13.5f%9.4f%6.0f%6.0f%12.5f\n" % (
                                k, 0.001 * k * args.step, epav / k, part["ekin"][i, 0], tav / k, part["temperature"][i, 0],
                                part["epot"][i, 0] + part["ekin"][i, 0]))
    write_mdrestart(b.array("xyz").reshape(-1, 3), b.array("velocities").reshape(-1, 3))
    if blown:
        print("xtb_b200: MD is unstable, emergency exit", file=sys.stderr)
        return 1
    if not args.silent:
        out.write(" average properties\n Epot               :%16.8f\n T                  :%16.8f\n" % (epav / max(k, 1), tav / max(k, 1)))
        out.write("normal exit of md()\n")
    open("xtbmdok", "w").close()
    return 0


UNSUPPORTED = ["--opt", "--ohess", "--hess", "--omd", "--metadyn", "--cosmo", "--gff", "--gfnff",
               "--vipea", "--vip", "--vea", "--path", "--modef", "--scan", "--esp", "--stm", "--tblite", "--cpcmx", "--oniom",
               "--dipro", "--siman", "--screen", "--bhess", "--vfukui", "--vomega", "--pcem"]


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    for flag in argv:
        if flag.split("=")[0] in UNSUPPORTED:
            print("xtb_b200: %s is not available in the B200 GFN2-xTB single point engine" % flag, file=sys.stderr)
            return 2
    ap = argparse.ArgumentParser(prog="xtb_b200", description="GFN2-xTB single point (+ gradient) on the B200 engine")
    ap.add_argument("geometry")
    ap.add_argument("--gfn", type=int, default=2)
    ap.add_argument("--sp", "--scc", action="store_true", help="single point (the default run type)")
    ap.add_argument("--grad", action="store_true", help="write the Turbomole gradient file")
    ap.add_argument("-c", "--chrg", type=float, default=0.0)
    ap.add_argument("-u", "--uhf", type=int, default=0)
    ap.add_argument("-a", "--acc", type=float, default=1.0)
    ap.add_argument("--etemp", type=float, default=300.0)
    ap.add_argument("--iterations", type=int, default=250)
    ap.add_argument("--alpb", metavar="SOLVENT", default=None, help="ALPB implicit solvation (src/prog/argparser / set_gbsa)")
    ap.add_argument("-g", "--gbsa", metavar="SOLVENT", default=None, help="GBSA implicit solvation")
    ap.add_argument("--md", action="store_true", help="molecular dynamics (defaults of the reference's $md block)")
    ap.add_argument("--time", type=float, default=50.0, help="MD: simulation time / ps")
    ap.add_argument("--step", type=float, default=4.0, help="MD: time step / fs")
    ap.add_argument("--temp", type=float, default=298.15, help="MD: thermostat temperature / K")
    ap.add_argument("--dump", type=float, default=50.0, help="MD: trajectory dump interval / fs")
    ap.add_argument("--shake", type=int, default=2, choices=[0, 1, 2], help="MD: 0 off, 1 X-H bonds, 2 all bonds")
    ap.add_argument("--hmass", type=int, default=4, help="MD: hydrogen mass / amu (0: physical)")
    ap.add_argument("--nve", action="store_true", help="MD: no thermostat (nvt=false)")
    ap.add_argument("--seed", type=int, default=1, help="MD: seed of the start velocities' signs")
    ap.add_argument("--restart", action="store_true", help="MD: positions and velocities from ./mdrestart")
    ap.add_argument("-I", "--input", default=None, help="xcontrol file: its $md block sets the MD parameters not given as flags")
    ap.add_argument("--norestart", action="store_true", help="do not read ./xtbrestart (it is still written)")
    ap.add_argument("-v", "--verbose", action="store_true")
    ap.add_argument("--silent", action="store_true")
    args = ap.parse_args(argv)
    if args.gfn != 2:
        print("xtb_b200: only --gfn 2 is implemented", file=sys.stderr)
        return 2
    import xtb_b200
    numbers, xyz = read_geometry(args.geometry)
    if args.alpb and args.gbsa:
        print("xtb_b200: --alpb and --gbsa are mutually exclusive", file=sys.stderr)
        return 2
    if args.md:
        if args.input:
            with open(args.input) as f:
                blk = read_xcontrol_md(f.read())
            for key in blk.pop("ignored"):
                print("xtb_b200: $md key '%s' is not used by this driver" % key, file=sys.stderr)
            given = {a.split("=")[0] for a in argv if a.startswith("--")}
            for dest, val in blk.items():
                if dest == "nvt":
                    if "--nve" not in given:
                        args.nve = not val
                elif dest == "restart":
                    args.restart = args.restart or val
                elif "--" + dest not in given:
                    setattr(args, dest, val)
        return run_md(args, numbers, xyz)
    t0 = time.perf_counter()
    # one molecule through the batched ABI: it takes the restart wavefunction of an `xtbrestart` file (src/restart.f90)
    b = xtb_b200.BatchCalculator([numbers], charges=[args.chrg], uhf=[args.uhf], accuracy=args.acc, etemp=args.etemp,
                                 maxiter=args.iterations)
    if args.alpb:
        b.set_solvent(args.alpb, model="alpb")
    elif args.gbsa:
        b.set_solvent(args.gbsa, model="gbsa")
    sz, ne = b.sizes(0), b.electrons(0)
    restarted = False
    if not args.norestart:
        wfn = read_xtbrestart(len(numbers), sz["nshell"], ne["nel"], ne["nopen"],
                              warn=lambda msg: print("xtb_b200: warning: " + msg, file=sys.stderr))
        if wfn is not None:
            b.set_wavefunction(*wfn)
            restarted = True
            if not args.silent:
                print("q/qsh data taken from xtbrestart\nCAMM data taken from xtbrestart")
    r = b.singlepoint(xyz, restart=restarted)
    wall = time.perf_counter() - t0
    res = b.result(0)
    if not r["all_converged"]:
        print("xtb_b200: SCC did not converge in %d iterations" % int(res["niter"]), file=sys.stderr)
    energy, grad = float(r["energies"][0]), r["gradients"]
    q, dip = r["charges"], np.array([res["dipole_x"], res["dipole_y"], res["dipole_z"]])
    wbo = b.array("wbo").reshape(len(numbers), len(numbers))
    gap = res["hl_gap"]       # the engine's egap from ihomoa (src/scc_core.f90:501), not re-derived from occupations
    write_xtbrestart(b.array("qsh"), b.array("dipm"), b.array("qp"), len(numbers), ne["nel"], ne["nopen"])
    if not args.silent:
        print("convergence criteria %s after %d iterations" % ("satisfied" if r["all_converged"] else "cannot be satisfied", int(res["niter"])))
    write_turbomole(numbers, xyz, energy, grad if args.grad else None)
    write_charges(q)
    write_wbo(wbo)
    if not args.silent:
        print("         :::::::::::::::::::::::::::::::::::::::::::::::::::::")
        print("         ::                     SUMMARY                     ::")
        print("         :::::::::::::::::::::::::::::::::::::::::::::::::::::")
        print("         :: total energy        %18.12f Eh    ::" % energy)
        print("         :: gradient norm       %18.12f Eh/a0 ::" % np.linalg.norm(grad))
        print("         :: HOMO-LUMO gap       %18.12f eV    ::" % gap)
        print("         :::::::::::::::::::::::::::::::::::::::::::::::::::::")
        print("molecular dipole (au): %12.6f %12.6f %12.6f" % tuple(dip))
        print(" * wall-time: %.3f s (device: GFN2-xTB engine on B200)" % wall)
        print("normal termination of xtb_b200")
    return 0


if __name__ == "__main__":
    sys.exit(main())
