"""File formats of the command line front end (xtb_b200/cli.py): xyz and Turbomole coord readers, Turbomole energy /
gradient writers (src/io/writer/turbomole.f90:96-123), charges and wbo files; the GPU test runs the reference's
`xtb caffeine.coord --gfn 2 --grad` use case end to end."""
import os
import subprocess
import sys

import numpy as np
import pytest

from xtb_b200 import cli

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_readers_agree_between_xyz_and_coord(kat, tmp_path):
    s = kat["caffeine"]
    coord = "$coord\n" + "".join("%22.14f %22.14f %22.14f  %s\n" % (r[0], r[1], r[2], cli.SYMBOLS[z - 1].lower())
                                 for z, r in zip(s["numbers"], s["xyz"])) + "$periodic 0\n$end\n"
    z1, x1 = cli.read_coord(coord)
    assert np.array_equal(z1, s["numbers"]) and np.max(np.abs(x1 - s["xyz"])) < 1e-13
    xyz = "%d\ncaffeine\n" % len(z1) + "".join("%-2s %20.14f %20.14f %20.14f\n" % (cli.SYMBOLS[z - 1], *(r / cli.AATOAU))
                                               for z, r in zip(s["numbers"], s["xyz"]))
    z2, x2 = cli.read_xyz(xyz)
    assert np.array_equal(z2, s["numbers"]) and np.max(np.abs(x2 - s["xyz"])) < 1e-12
    p = tmp_path / "m.coord"
    p.write_text(coord.replace("$coord", "$coord angs"))
    z3, x3 = cli.read_geometry(str(p))
    assert np.max(np.abs(x3 - s["xyz"] * cli.AATOAU)) < 1e-12


def test_turbomole_writers_layout_and_round_trip(kat, tmp_path):
    s = kat["c3h4_api"]
    rng = np.random.default_rng(0)
    g = 1e-2 * rng.standard_normal(s["xyz"].shape)
    e = -8.3824793849585
    ef, gf = str(tmp_path / "energy"), str(tmp_path / "gradient")
    cli.write_turbomole(s["numbers"], s["xyz"], e, g, iter_number=1, energy_file=ef, gradient_file=gf)
    el = open(ef).read().splitlines()
    assert el[0] == "$energy" and el[-1] == "$end"
    assert el[1] == "%6d%18.11f%18.11f%18.11f" % (1, e, e, e) and len(el[1]) == 60      # (i6,3F18.11)
    gl = open(gf).read().splitlines()
    assert gl[0] == "$grad" and gl[-1] == "$end" and len(gl) == 3 + 2 * len(s["numbers"])
    assert gl[1].startswith("  cycle =      1    SCF energy =")
    assert len(gl[2 + len(s["numbers"])]) == 66                                          # (3ES22.13)
    e2, x2, g2 = cli.read_turbomole_gradient(gf)
    assert abs(e2 - e) < 1e-11 and np.max(np.abs(x2 - s["xyz"])) < 1e-13 and np.max(np.abs(g2 - g)) < 1e-15
    cli.write_charges([0.1, -0.25], str(tmp_path / "charges"))
    assert open(tmp_path / "charges").read() == "    0.10000000\n   -0.25000000\n"
    w = np.array([[0, 0.95, 0.05], [0.95, 0, 1.2], [0.05, 1.2, 0]])
    cli.write_wbo(w, str(tmp_path / "wbo"))
    rows = [ln.split() for ln in open(tmp_path / "wbo").read().splitlines()]
    assert [(int(a), int(b)) for a, b, _ in rows] == [(1, 2), (2, 3)]


def test_unsupported_run_types_are_refused():
    assert cli.main(["mol.xyz", "--opt"]) == 2
    assert cli.main(["mol.xyz", "--cosmo", "water"]) == 2
    assert cli.main(["mol.xyz", "--alpb", "water", "--gbsa", "water", "--gfn", "1"]) == 2
    assert cli.main(["mol.xyz", "--gfn", "1"]) == 2


@pytest.mark.gpu
def test_cli_single_point_with_gradient(kat, tmp_path):
    from oracle.oracle import Oracle
    s = kat["caffeine"]
    coord = "$coord\n" + "".join("%22.14f %22.14f %22.14f  %s\n" % (r[0], r[1], r[2], cli.SYMBOLS[z - 1].lower())
                                 for z, r in zip(s["numbers"], s["xyz"])) + "$end\n"
    (tmp_path / "caffeine.coord").write_text(coord)
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-m", "xtb_b200", "caffeine.coord", "--gfn", "2", "--grad"], cwd=tmp_path, env=env,
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "normal termination" in out.stdout
    o = Oracle(s["numbers"], s["xyz"])
    assert o.singlepoint() == 0
    e, x, g = cli.read_turbomole_gradient(str(tmp_path / "gradient"))
    assert abs(e - o.energies["e_total"]) < 1e-7
    assert np.max(np.abs(g - o.get("gradient").reshape(-1, 3))) < 1e-6
    q = np.loadtxt(tmp_path / "charges")
    assert np.max(np.abs(q - o.get("q"))) < 1e-6
    assert os.path.exists(tmp_path / "wbo") and os.path.exists(tmp_path / "energy")


@pytest.mark.gpu
def test_cli_alpb_single_point(kat, tmp_path):
    """`xtb caffeine.xyz --gfn 2 --grad --alpb water` on the engine against the oracle."""
    from oracle.oracle import Oracle
    s = kat["caffeine"]
    xyz_aa = s["xyz"] * 0.52917726
    (tmp_path / "caffeine.xyz").write_text("%d\n\n" % len(s["numbers"]) + "".join(
        "%-2s %20.14f %20.14f %20.14f\n" % (cli.SYMBOLS[z - 1], r[0], r[1], r[2]) for z, r in zip(s["numbers"], xyz_aa)))
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-m", "xtb_b200", "caffeine.xyz", "--gfn", "2", "--grad", "--alpb", "water"],
                         cwd=tmp_path, env=env, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    o = Oracle(s["numbers"], s["xyz"])
    o.set_solvent("water", "alpb")
    assert o.singlepoint() == 0
    e, x, g = cli.read_turbomole_gradient(str(tmp_path / "gradient"))
    assert abs(e - o.energies["e_total"]) < 1e-7
    assert np.max(np.abs(g - o.get("gradient").reshape(-1, 3))) < 1e-6


def test_xtbrestart_round_trip_and_mismatch_rules(tmp_path):
    """The binary restart file (src/restart.f90:33-86, 182-200): record layout byte for byte, round trip, and the
    reference's acceptance rules (dimension mismatch refuses, electron count / multiplicity / version only warn)."""
    rng = np.random.default_rng(1)
    nat, nsh = 5, 9
    qsh, dipm, qp = rng.normal(size=nsh), rng.normal(size=(nat, 3)), rng.normal(size=(nat, 6))
    p = str(tmp_path / "xtbrestart")
    cli.write_xtbrestart(qsh, dipm, qp, nat, nel=14, nopen=0, path=p)
    raw = open(p, "rb").read()
    # Fortran sequential unformatted records: marker, payload, marker
    assert len(raw) == (4 + 48 + 4) + (4 + 8 * nsh + 4) + (4 + 24 * nat + 4) + (4 + 48 * nat + 4)
    assert np.frombuffer(raw[:4], "<i4")[0] == 48 and np.frombuffer(raw[52:56], "<i4")[0] == 48
    assert np.frombuffer(raw[4:52], "<i8").tolist() == [2, 2, nat, nsh, 14, 0]
    assert np.frombuffer(raw[56:60], "<i4")[0] == 8 * nsh
    assert np.array_equal(np.frombuffer(raw[60:60 + 8 * nsh], "<f8"), qsh)
    # dipm(3, n), qp(6, n) in Fortran order = [atom][component] in C order
    off = 60 + 8 * nsh + 4 + 4
    assert np.array_equal(np.frombuffer(raw[off:off + 24 * nat], "<f8").reshape(nat, 3), dipm)
    got = cli.read_xtbrestart(nat, nsh, 14, 0, path=p)
    assert all(np.array_equal(a, b) for a, b in zip(got, (qsh, dipm, qp)))
    msgs = []
    assert cli.read_xtbrestart(nat, nsh, 16, 2, path=p, warn=msgs.append) is not None
    assert msgs == ["Number of electron missmatch in restart file.", "Multiplicity missmatch in restart file."]
    msgs.clear()
    assert cli.read_xtbrestart(nat + 1, nsh, 14, 0, path=p, warn=msgs.append) is None
    assert cli.read_xtbrestart(nat, nsh - 1, 14, 0, path=p, warn=msgs.append) is None
    assert msgs == ["Dimension missmatch in restart file."] * 2
    assert cli.read_xtbrestart(nat, nsh, path=str(tmp_path / "missing")) is None
    open(p, "wb").write(raw[:70])                        # truncated file
    assert cli.read_xtbrestart(nat, nsh, 14, 0, path=p) is None


@pytest.mark.gpu
def test_cli_restarts_from_xtbrestart(kat, tmp_path):
    """A second run in the same directory starts from the first run's xtbrestart (fewer SCC iterations, same energy); the
    file's shell charges and multipoles are the converged ones of the engine."""
    import re
    s = kat["caffeine"]
    (tmp_path / "caffeine.xyz").write_text("%d\n\n" % len(s["numbers"]) + "".join(
        "%-2s %20.14f %20.14f %20.14f\n" % (cli.SYMBOLS[z - 1], *(r / cli.AATOAU)) for z, r in zip(s["numbers"], s["xyz"])))
    env = dict(os.environ, PYTHONPATH=ROOT)
    runs = []
    for k in range(2):
        out = subprocess.run([sys.executable, "-m", "xtb_b200", "caffeine.xyz", "--gfn", "2"], cwd=tmp_path, env=env,
                             capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        niter = int(re.search(r"after (\d+) iterations", out.stdout).group(1))
        energy = float(re.search(r"total energy\s+(-?\d+\.\d+)", out.stdout).group(1))
        runs.append((niter, energy, "taken from xtbrestart" in out.stdout))
    assert runs[0][2] is False and runs[1][2] is True
    assert runs[1][0] < runs[0][0]
    assert abs(runs[1][1] - runs[0][1]) < 1e-7
    import xtb_b200
    b = xtb_b200.BatchCalculator([s["numbers"]])
    b.singlepoint(s["xyz"] if False else cli.read_geometry(str(tmp_path / "caffeine.xyz"))[1])
    qsh, dipm, qp = cli.read_xtbrestart(len(s["numbers"]), b.sizes(0)["nshell"], path=str(tmp_path / "xtbrestart"))
    assert np.max(np.abs(qsh - b.array("qsh"))) < 1e-5 and np.max(np.abs(dipm.ravel() - b.array("dipm"))) < 1e-5
    assert np.max(np.abs(qp.ravel() - b.array("qp"))) < 1e-5


def test_md_restart_and_trajectory_formats(tmp_path):
    """mdrestart (src/dynamic.f90:814-844: '-1.0' + 6D22.14 per atom) round trip and the xyz frames of xtb.trj."""
    rng = np.random.default_rng(0)
    x, v = rng.normal(size=(5, 3)) * 4.0, rng.normal(size=(5, 3)) * 1e-3
    x[0, 0], v[1, 2] = 0.0, -0.099999999999999999
    p = tmp_path / "mdrestart"
    cli.write_mdrestart(x, v, str(p))
    lines = p.read_text().splitlines()
    assert lines[0].strip() == "-1.0" and len(lines) == 6 and all(len(ln) == 132 for ln in lines[1:])
    assert lines[1][:22] == "  0.00000000000000D+00"
    assert cli._fortran_d(-0.0377190663280180, 22, 14) == " -0.37719066328018D-01"
    assert cli._fortran_d(0.99999999999999999, 22, 14) == "  0.10000000000000D+01"
    x2, v2 = cli.read_mdrestart(str(p))
    assert np.max(np.abs(x - x2)) < 1e-13 and np.max(np.abs(v - v2)) < 1e-16
    with open(tmp_path / "xtb.trj", "w") as f:
        cli.write_trj_frame(f, [8, 1, 1], x[:3], -5.0704513551, 0.0064574201)
        cli.write_trj_frame(f, [8, 1, 1], x[:3] + 0.1, -5.07, 0.007)
    text = (tmp_path / "xtb.trj").read_text().splitlines()
    assert text[0].strip() == "3" and text[1].startswith(" energy: -5.070451355100 gnorm: 0.006457420100 xtb:")
    assert len(text) == 10 and text[2].split()[0] == "O"
    z, xyz = cli.read_xyz("\n".join(text[:5]) + "\n")
    assert list(z) == [8, 1, 1] and np.max(np.abs(xyz - x[:3])) < 1e-12


@pytest.mark.gpu
def test_cli_md_run(kat, tmp_path):
    """`xtb caffeine.xyz --md` with a short $md block: trajectory frames, mdrestart, xtbmdok; restart continues."""
    s = kat["caffeine"]
    (tmp_path / "caffeine.xyz").write_text("%d\n\n" % len(s["numbers"]) + "".join(
        "%-2s %20.14f %20.14f %20.14f\n" % (cli.SYMBOLS[z - 1], r[0], r[1], r[2]) for z, r in zip(s["numbers"], s["xyz"] / cli.AATOAU)))
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "xtb_b200", "caffeine.xyz", "--md", "--time", "0.04", "--dump", "16", "--alpb", "water"]
    out = subprocess.run(cmd, cwd=tmp_path, env=env, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "normal exit of md()" in out.stdout and "SHAKE on. # bonds  : 25" in out.stdout
    assert (tmp_path / "xtbmdok").exists()
    frames = (tmp_path / "xtb.trj").read_text().splitlines()
    assert len(frames) == 3 * 26                                  # steps 1, 5, 9 of 10
    e = [float(frames[26 * k + 1].split()[1]) for k in range(3)]
    assert all(-42.3 < x < -42.0 for x in e)
    x1, v1 = cli.read_mdrestart(str(tmp_path / "mdrestart"))
    out = subprocess.run(cmd + ["--restart", "--time", "0.008"], cwd=tmp_path, env=env, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    x2, v2 = cli.read_mdrestart(str(tmp_path / "mdrestart"))
    assert 1e-4 < np.max(np.abs(x2 - x1)) < 0.5


def test_xcontrol_md_block():
    """`$md` block of an xcontrol file (keys of set_md, src/set_module.f90:1975-2072)."""
    text = """$chrg 0
$md
   temp=350.0 # in K
   time= 10.0
   dumptrj = 20.0
   step=  2.0
   velo=false
   nvt =false
   hmass=2
   shake=1
   sccacc=2.0
   time=99.0
$end
$scc
   temp=300
$end
"""
    blk = cli.read_xcontrol_md(text)
    assert blk == dict(temp=350.0, time=10.0, dump=20.0, step=2.0, nvt=False, hmass=2, shake=1, ignored=["velo", "sccacc"])
    assert cli.read_xcontrol_md("$opt\n maxcycle=5\n$end\n") == dict(ignored=[])
    with pytest.raises(ValueError):
        cli.read_xcontrol_md("$md\n shake=3\n$end\n")
