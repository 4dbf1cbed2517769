"""Stand-in for BASELINE configs[1] (crambin, whose coordinates are not available offline): a covalently
connected, clash-free random coil with crambin's 46-residue sequence, neutral termini and side chains, free thiols —
C202H321N55O64S6, 648 atoms, 1659 GFN2 AOs (crambin with its three disulfide bridges: C202H315N55O64S6, 642 atoms).
Built once by the seeded generator tools/make_polypeptide.py (seed 1) and kept as tests/golden/polypeptide_1crn.json."""
import json
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "polypeptide_1crn.json")


def make_polypeptide():
    """Atomic numbers, positions (bohr) and a description of the polypeptide."""
    with open(_PATH) as f:
        d = json.load(f)
    Z = np.array(d["numbers"], dtype=np.int32)
    xyz = np.array(d["xyz"], dtype=np.float64).reshape(-1, 3)
    return Z, xyz, {"formula": d["formula"], "seed": 1, "labels": d["labels"], "charge": d["charge"], "uhf": d["uhf"]}
