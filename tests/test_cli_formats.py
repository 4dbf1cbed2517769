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
