#!/usr/bin/env python3
"""Seeded generator of the stand-in for BASELINE configs[1] (crambin): a covalently connected, clash-free 46-residue
polypeptide with crambin's sequence (SURVEY 8(d) C2 — the coordinates of 1CRN are not available offline).  Neutral termini
(NH2 / COOH), neutral side chains (Asp / Glu as acids, Arg as neutral guanidine), total charge 0, closed shell.  The six
cysteines are free thiols: a random coil cannot close crambin's three disulfide bridges, so the composition is
C202H321N55O64S6 (648 atoms, 1659 GFN2 AOs) = crambin's C202H315N55O64S6 (642 atoms, 1653 AOs) plus the six thiol hydrogens.

The chain is a self-avoiding random coil: every residue draws its (phi, psi) from {helix, strand, PPII} and its side-chain
rotamers from {-60, 180, 60} with numpy.random.default_rng(seed); heavy atoms are placed from internal coordinates
(bond, angle, torsion), hydrogens from the hybridisation of their parent; a residue that clashes with the chain built so
far is redrawn, and the chain backtracks when a residue cannot be placed.

    python tools/make_polypeptide.py            # writes tests/golden/polypeptide_1crn.json (positions in bohr)
"""
import json
import os
import sys

import numpy as np

AA2BOHR = 1.0 / 0.52917721067
SEQ = "TTCCPSIVARSNFNVCRLPGTPEALCATYTGCIIIPGATCPGDYAN"      # crambin (1CRN), 46 residues
ZNUM = {"H": 1, "C": 6, "N": 7, "O": 8, "S": 16}
D2R = np.pi / 180.0


def place(a, b, c, r, theta, phi):
    """Position d with |cd| = r, angle(b, c, d) = theta, dihedral(a, b, c, d) = phi (degrees)."""
    bc = c - b
    bc /= np.linalg.norm(bc)
    n = np.cross(b - a, bc)
    n /= np.linalg.norm(n)
    m = np.cross(n, bc)
    t, p = theta * D2R, phi * D2R
    d2 = np.array([-r * np.cos(t), r * np.sin(t) * np.cos(p), r * np.sin(t) * np.sin(p)])
    return c + d2[0] * bc + d2[1] * m + d2[2] * n


# side chains: heavy atoms as (name, element, (ref a, ref b, ref c), bond / Angstrom, angle, torsion); a torsion given as
# ("chi", k, offset) is rotamer k of the residue plus the offset
SIDE = {
    "G": [],
    "A": [],
    "S": [("OG", "O", ("N", "CA", "CB"), 1.417, 111.0, ("chi", 0, 0.0))],
    "C": [("SG", "S", ("N", "CA", "CB"), 1.808, 114.0, ("chi", 0, 0.0))],
    "T": [("OG1", "O", ("N", "CA", "CB"), 1.433, 109.6, ("chi", 0, 0.0)),
          ("CG2", "C", ("N", "CA", "CB"), 1.521, 110.5, ("chi", 0, -120.0))],
    "V": [("CG1", "C", ("N", "CA", "CB"), 1.521, 110.5, ("chi", 0, 0.0)),
          ("CG2", "C", ("N", "CA", "CB"), 1.521, 110.5, ("chi", 0, 120.0))],
    "I": [("CG1", "C", ("N", "CA", "CB"), 1.530, 110.4, ("chi", 0, 0.0)),
          ("CG2", "C", ("N", "CA", "CB"), 1.521, 110.5, ("chi", 0, -120.0)),
          ("CD1", "C", ("CA", "CB", "CG1"), 1.513, 113.8, ("chi", 1, 0.0))],
    "L": [("CG", "C", ("N", "CA", "CB"), 1.530, 116.3, ("chi", 0, 0.0)),
          ("CD1", "C", ("CA", "CB", "CG"), 1.521, 110.5, ("chi", 1, 0.0)),
          ("CD2", "C", ("CA", "CB", "CG"), 1.521, 110.5, ("chi", 1, 120.0))],
    "P": [],      # ring atoms CD, CG are placed from the backbone (build_residue)
    "F": [("CG", "C", ("N", "CA", "CB"), 1.502, 113.8, ("chi", 0, 0.0)),
          ("CD1", "C", ("CA", "CB", "CG"), 1.390, 120.0, ("chi", 1, 0.0)),
          ("CD2", "C", ("CA", "CB", "CG"), 1.390, 120.0, ("chi", 1, 180.0)),
          ("CE1", "C", ("CB", "CG", "CD1"), 1.390, 120.0, 180.0),
          ("CE2", "C", ("CB", "CG", "CD2"), 1.390, 120.0, 180.0),
          ("CZ", "C", ("CG", "CD1", "CE1"), 1.390, 120.0, 0.0)],
    "Y": [("CG", "C", ("N", "CA", "CB"), 1.512, 113.8, ("chi", 0, 0.0)),
          ("CD1", "C", ("CA", "CB", "CG"), 1.390, 120.0, ("chi", 1, 0.0)),
          ("CD2", "C", ("CA", "CB", "CG"), 1.390, 120.0, ("chi", 1, 180.0)),
          ("CE1", "C", ("CB", "CG", "CD1"), 1.390, 120.0, 180.0),
          ("CE2", "C", ("CB", "CG", "CD2"), 1.390, 120.0, 180.0),
          ("CZ", "C", ("CG", "CD1", "CE1"), 1.390, 120.0, 0.0),
          ("OH", "O", ("CD1", "CE1", "CZ"), 1.376, 120.0, 180.0)],
    "N": [("CG", "C", ("N", "CA", "CB"), 1.516, 112.6, ("chi", 0, 0.0)),
          ("OD1", "O", ("CA", "CB", "CG"), 1.231, 120.8, ("chi", 1, 0.0)),
          ("ND2", "N", ("CA", "CB", "CG"), 1.328, 116.4, ("chi", 1, 180.0))],
    "D": [("CG", "C", ("N", "CA", "CB"), 1.516, 112.6, ("chi", 0, 0.0)),
          ("OD1", "O", ("CA", "CB", "CG"), 1.210, 123.0, ("chi", 1, 0.0)),
          ("OD2", "O", ("CA", "CB", "CG"), 1.340, 112.0, ("chi", 1, 180.0))],
    "E": [("CG", "C", ("N", "CA", "CB"), 1.520, 114.1, ("chi", 0, 0.0)),
          ("CD", "C", ("CA", "CB", "CG"), 1.516, 112.6, ("chi", 1, 0.0)),
          ("OE1", "O", ("CB", "CG", "CD"), 1.210, 123.0, ("chi", 2, 0.0)),
          ("OE2", "O", ("CB", "CG", "CD"), 1.340, 112.0, ("chi", 2, 180.0))],
    "R": [("CG", "C", ("N", "CA", "CB"), 1.520, 114.1, ("chi", 0, 0.0)),
          ("CD", "C", ("CA", "CB", "CG"), 1.520, 111.3, ("chi", 1, 0.0)),
          ("NE", "N", ("CB", "CG", "CD"), 1.461, 112.0, ("chi", 2, 0.0)),
          ("CZ", "C", ("CG", "CD", "NE"), 1.330, 124.2, ("chi", 3, 0.0)),
          ("NH1", "N", ("CD", "NE", "CZ"), 1.290, 120.0, 0.0),
          ("NH2", "N", ("CD", "NE", "CZ"), 1.360, 120.0, 180.0)],
}
# bonds between heavy atoms of a residue beyond N-CA, CA-C, C-O, CA-CB (used for the hydrogens and the clash exclusions)
BONDS = {
    "G": [], "A": [], "S": [("CB", "OG")], "C": [("CB", "SG")], "T": [("CB", "OG1"), ("CB", "CG2")],
    "V": [("CB", "CG1"), ("CB", "CG2")], "I": [("CB", "CG1"), ("CB", "CG2"), ("CG1", "CD1")],
    "L": [("CB", "CG"), ("CG", "CD1"), ("CG", "CD2")], "P": [("CB", "CG"), ("CG", "CD"), ("CD", "N")],
    "F": [("CB", "CG"), ("CG", "CD1"), ("CG", "CD2"), ("CD1", "CE1"), ("CD2", "CE2"), ("CE1", "CZ"), ("CE2", "CZ")],
    "Y": [("CB", "CG"), ("CG", "CD1"), ("CG", "CD2"), ("CD1", "CE1"), ("CD2", "CE2"), ("CE1", "CZ"), ("CE2", "CZ"), ("CZ", "OH")],
    "N": [("CB", "CG"), ("CG", "OD1"), ("CG", "ND2")], "D": [("CB", "CG"), ("CG", "OD1"), ("CG", "OD2")],
    "E": [("CB", "CG"), ("CG", "CD"), ("CD", "OE1"), ("CD", "OE2")],
    "R": [("CB", "CG"), ("CG", "CD"), ("CD", "NE"), ("NE", "CZ"), ("CZ", "NH1"), ("CZ", "NH2")],
}
# hydrogens per heavy atom: (count, kind); kinds: "sp3" (tetrahedral), "sp2" (planar)
HYD = {
    "CA": (1, "sp3"), "CB": (2, "sp3"),
    ("G", "CA"): (2, "sp3"), ("A", "CB"): (3, "sp3"), ("T", "CB"): (1, "sp3"), ("V", "CB"): (1, "sp3"), ("I", "CB"): (1, "sp3"),
    ("S", "OG"): (1, "sp3"), ("C", "SG"): (1, "sp3"), ("T", "OG1"): (1, "sp3"), ("T", "CG2"): (3, "sp3"),
    ("V", "CG1"): (3, "sp3"), ("V", "CG2"): (3, "sp3"), ("I", "CG1"): (2, "sp3"), ("I", "CG2"): (3, "sp3"), ("I", "CD1"): (3, "sp3"),
    ("L", "CG"): (1, "sp3"), ("L", "CD1"): (3, "sp3"), ("L", "CD2"): (3, "sp3"),
    ("P", "CG"): (2, "sp3"), ("P", "CD"): (2, "sp3"),
    ("F", "CD1"): (1, "sp2"), ("F", "CD2"): (1, "sp2"), ("F", "CE1"): (1, "sp2"), ("F", "CE2"): (1, "sp2"), ("F", "CZ"): (1, "sp2"),
    ("Y", "CD1"): (1, "sp2"), ("Y", "CD2"): (1, "sp2"), ("Y", "CE1"): (1, "sp2"), ("Y", "CE2"): (1, "sp2"), ("Y", "OH"): (1, "sp2"),
    ("N", "ND2"): (2, "sp2"), ("D", "OD2"): (1, "sp2"),
    ("E", "CG"): (2, "sp3"), ("E", "OE2"): (1, "sp2"),
    ("R", "CG"): (2, "sp3"), ("R", "CD"): (2, "sp3"), ("R", "NE"): (1, "sp2"), ("R", "NH1"): (1, "sp2"), ("R", "NH2"): (2, "sp2"),
    ("N", "CG"): (0, "sp2"), ("D", "CG"): (0, "sp2"), ("F", "CG"): (0, "sp2"), ("Y", "CG"): (0, "sp2"), ("Y", "CZ"): (0, "sp2"),
    ("E", "CD"): (0, "sp2"), ("R", "CZ"): (0, "sp2"),
}
RH = {"C": 1.09, "N": 1.01, "O": 0.97, "S": 1.34}
CONF = [(-57.0, -47.0), (-120.0, 130.0), (-75.0, 145.0)]     # helix, strand, polyproline II


def unit(v):
    return v / np.linalg.norm(v)


def add_hydrogens(center, el, nbrs, second, count, kind):
    """Hydrogens of a heavy atom at `center` bonded to the heavy atoms at nbrs; `second`: a neighbour of nbrs[0] other than
    the centre (torsion reference of terminal groups)."""
    r = RH[el]
    us = [unit(x - center) for x in nbrs]
    out = []
    if count == 0:
        return out
    if len(nbrs) == 3:                                  # CH
        out.append(center - r * unit(us[0] + us[1] + us[2]))
    elif len(nbrs) == 2 and kind == "sp3":              # CH2 (count 2) or a lone sp3 H
        b = -unit(us[0] + us[1])
        n = unit(np.cross(us[0], us[1]))
        al = 54.75 * D2R
        out.append(center + r * (np.cos(al) * b + np.sin(al) * n))
        if count == 2:
            out.append(center + r * (np.cos(al) * b - np.sin(al) * n))
    elif len(nbrs) == 2:                                # planar X-H between two neighbours
        out.append(center - r * unit(us[0] + us[1]))
    else:                                               # terminal group on one neighbour
        ang = 109.5 if kind == "sp3" else 120.0
        tors = {("sp3", 3): (60.0, 180.0, 300.0), ("sp3", 2): (60.0, 300.0), ("sp3", 1): (180.0,),
                ("sp2", 2): (0.0, 180.0), ("sp2", 1): (180.0,)}[(kind, count)]
        for t in tors:
            out.append(place(second, nbrs[0], center, r, ang, t))
    return out


class Residue:
    def __init__(self, aa):
        self.aa = aa
        self.heavy = {}         # name -> position
        self.el = {}
        self.hyd = []           # (position, parent name)


def build_residue(aa, prev, phi, psi_prev, chis, first):
    """Backbone + side chain heavy atoms of one residue.  prev: the previous Residue (None for the first)."""
    res = Residue(aa)
    if first:
        res.heavy["N"] = np.array([0.0, 0.0, 0.0])
        res.heavy["CA"] = np.array([1.458, 0.0, 0.0])
        res.heavy["C"] = place(np.array([0.0, 1.0, 0.0]), res.heavy["N"], res.heavy["CA"], 1.525, 111.0, phi)
    else:
        pN, pCA, pC = prev.heavy["N"], prev.heavy["CA"], prev.heavy["C"]
        res.heavy["N"] = place(pN, pCA, pC, 1.329, 116.2, psi_prev)
        res.heavy["CA"] = place(pCA, pC, res.heavy["N"], 1.458, 121.7, 180.0)
        res.heavy["C"] = place(pC, res.heavy["N"], res.heavy["CA"], 1.525, 111.0, phi)
    res.el.update({"N": "N", "CA": "C", "C": "C"})
    if aa != "G":
        # L configuration; proline: N-CA-CB closed to 103.5 degrees so that the ring can close (CB-CD 2.40 A)
        res.heavy["CB"] = place(res.heavy["N"], res.heavy["C"], res.heavy["CA"], 1.530, 110.1, 114.0 if aa == "P" else 122.6)
        res.el["CB"] = "C"
    if aa == "P":
        # CD on the amide nitrogen, cis to the preceding CA; CG closes the ring: 1.495 A from CB, 1.507 A from CD, on the side
        # away from N / CA, 0.35 A out of the plane (pucker)
        if first:
            cd = place(res.heavy["C"], res.heavy["CA"], res.heavy["N"], 1.473, 111.0, 15.0)
        else:
            cd = place(prev.heavy["CA"], prev.heavy["C"], res.heavy["N"], 1.473, 125.0, 0.0)
        cb = res.heavy["CB"]
        mid, ax = 0.5 * (cb + cd), unit(cd - cb)
        half = 0.5 * np.linalg.norm(cd - cb)
        # sphere intersection (the two bond lengths differ by 0.012 A: treated as equal, 1.501 A)
        h = np.sqrt(max(1.501 ** 2 - half ** 2, 0.0))
        away = mid - 0.5 * (res.heavy["N"] + res.heavy["CA"])
        away = unit(away - np.dot(away, ax) * ax)
        nrm = unit(np.cross(ax, away))
        ang = np.arcsin(min(0.35 / h, 1.0))
        res.heavy["CG"] = mid + h * (np.cos(ang) * away + np.sin(ang) * nrm)
        res.heavy["CD"] = cd
        res.el["CG"] = "C"
        res.el["CD"] = "C"
    for name, el, refs, r, th, tor in SIDE[aa]:
        if isinstance(tor, tuple):
            tor = chis[tor[1]] + tor[2]
        a, b, c = (res.heavy[x] for x in refs)
        res.heavy[name] = place(a, b, c, r, th, tor)
        res.el[name] = el
    return res


def finish_residue(res, prev, psi, last):
    """Carbonyl oxygen (needs psi), terminal groups and all hydrogens."""
    aa = res.aa
    N, CA, C = res.heavy["N"], res.heavy["CA"], res.heavy["C"]
    res.heavy["O"] = place(N, CA, C, 1.231, 120.5, psi + 180.0)
    res.el["O"] = "O"
    if last:
        res.heavy["OXT"] = place(N, CA, C, 1.340, 112.0, psi)
        res.el["OXT"] = "O"
    nb = {k: [] for k in res.heavy}
    bonds = [("N", "CA"), ("CA", "C"), ("C", "O")] + ([("CA", "CB")] if aa != "G" else []) + BONDS[aa]
    if last:
        bonds.append(("C", "OXT"))
    for a, b in bonds:
        nb[a].append(b)
        nb[b].append(a)
    res.hyd = []
    # backbone N: amide H (none for proline), NH2 at the N terminus
    if prev is None:
        n_h = 1 if aa == "P" else 2
        for t in ((60.0, 300.0) if n_h == 2 else (180.0,)):
            res.hyd.append((place(C, CA, N, RH["N"], 109.5, t), "N"))
    elif aa != "P":
        res.hyd.append((N - RH["N"] * unit(unit(prev.heavy["C"] - N) + unit(CA - N)), "N"))
    if last:
        res.hyd.append((place(CA, C, res.heavy["OXT"], RH["O"], 109.5, 180.0), "OXT"))
    for name in list(res.heavy):
        if name in ("N", "C", "O", "OXT"):
            continue
        cnt, kind = HYD.get((aa, name), HYD.get(name, (0, "sp3")))
        if cnt == 0:
            continue
        nbrs = [res.heavy[x] for x in nb[name]]
        first_nb = nb[name][0]
        second = None
        for x in nb[first_nb]:
            if x != name:
                second = res.heavy[x]
                break
        if second is None:
            second = N
        for h in add_hydrogens(res.heavy[name], res.el[name], nbrs, second, cnt, kind):
            res.hyd.append((h, name))
    return res


def atoms_of(res):
    el = [res.el[k] for k in res.heavy] + ["H"] * len(res.hyd)
    xyz = [res.heavy[k] for k in res.heavy] + [h for h, _ in res.hyd]
    return el, np.array(xyz)


def clashes(el_new, x_new, el_old, x_old, skip_bonded=None):
    """True when an atom of the new set comes too close to one of the old set."""
    if len(x_old) == 0:
        return False
    d = np.linalg.norm(x_new[:, None, :] - x_old[None, :, :], axis=2)
    hn = np.array([e == "H" for e in el_new])[:, None]
    ho = np.array([e == "H" for e in el_old])[None, :]
    lim = np.where(hn & ho, 1.9, np.where(hn | ho, 2.2, 2.9))
    if skip_bonded is not None:
        lim = np.where(skip_bonded, 0.0, lim)
    return bool(np.any(d < lim))


def internal_clash(res):
    """Atoms of one residue that are three bonds apart (1-4) or further must keep non-bonded distances."""
    el, x = atoms_of(res)
    names = list(res.heavy)
    n = len(x)
    adj = np.zeros((n, n), dtype=bool)
    bonds = [("N", "CA"), ("CA", "C"), ("C", "O")] + ([("CA", "CB")] if res.aa != "G" else []) + BONDS[res.aa]
    if "OXT" in res.heavy:
        bonds.append(("C", "OXT"))
    for a, b in bonds:
        adj[names.index(a), names.index(b)] = adj[names.index(b), names.index(a)] = True
    for k, (_, parent) in enumerate(res.hyd):
        adj[len(names) + k, names.index(parent)] = adj[names.index(parent), len(names) + k] = True
    a1 = adj.astype(int)
    a2 = (a1 @ a1) > 0
    a3 = (a1 @ a1 @ a1) > 0
    near = adj | a2 | np.eye(n, dtype=bool)          # 1-2, 1-3
    d = np.linalg.norm(x[:, None] - x[None, :], axis=2)
    h = np.array([e == "H" for e in el])
    hh, hx = h[:, None] & h[None, :], h[:, None] ^ h[None, :]
    lim14 = np.where(hh, 1.9, np.where(hx, 2.15, 2.45))
    lim15 = np.where(hh, 2.0, np.where(hx, 2.3, 2.8))
    lim = np.where(near, 0.0, np.where(a3, lim14, lim15))
    return bool(np.any(d < lim))


def generate(seed=1, max_tries=60):
    rng = np.random.default_rng(seed)
    chain = []
    psis = []
    tries = [0] * len(SEQ)
    i = 0
    while i < len(SEQ):
        aa = SEQ[i]
        ok = False
        while tries[i] < max_tries and not ok:
            tries[i] += 1
            phi, psi = CONF[rng.integers(0, 3)]
            if aa == "P":
                phi = -65.0
            phi += rng.normal(0.0, 8.0)
            psi += rng.normal(0.0, 8.0)
            chis = [(-60.0, 180.0, 60.0)[rng.integers(0, 3)] + rng.normal(0.0, 6.0) for _ in range(4)]
            if aa in ("F", "Y"):
                chis[1] = 90.0 + rng.normal(0.0, 10.0)
            prev = chain[i - 1] if i else None
            res = build_residue(aa, prev, phi, psis[i - 1] if i else 0.0, chis, i == 0)
            finish_residue(res, prev, psi, i == len(SEQ) - 1)
            if internal_clash(res):
                continue
            el_n, x_n = atoms_of(res)
            bad = False
            for k in range(i):
                el_o, x_o = atoms_of(chain[k])
                if k == i - 1:
                    # neighbouring residue: looser limits around the peptide bond
                    d = np.linalg.norm(x_n[:, None] - x_o[None, :], axis=2)
                    hn = np.array([e == "H" for e in el_n])[:, None]
                    ho = np.array([e == "H" for e in el_o])[None, :]
                    lim = np.where(hn & ho, 2.0, np.where(hn | ho, 2.2, 2.4))
                    near = np.zeros_like(d, dtype=bool)
                    names_n = list(res.heavy) + ["H:" + p for _, p in res.hyd]
                    names_o = list(chain[k].heavy) + ["H:" + p for _, p in chain[k].hyd]
                    for a, na in enumerate(names_n):
                        for b, nb_ in enumerate(names_o):
                            if na in ("N", "CA", "H:N", "CD", "H:CD", "H:CA", "C", "CB") and nb_ in ("C", "O", "CA", "H:CA", "N", "CB"):
                                near[a, b] = True
                    if np.any((d < lim) & ~near) or np.any(d < 0.9):
                        bad = True
                        break
                elif clashes(el_n, x_n, el_o, x_o):
                    bad = True
                    break
            if bad:
                continue
            ok = True
        if ok:
            if i < len(chain):
                chain[i] = res
                psis[i] = psi
            else:
                chain.append(res)
                psis.append(psi)
            i += 1
        else:
            tries[i] = 0
            i -= 1
            if i < 0:
                raise RuntimeError("generator failed")
            chain = chain[:i]
            psis = psis[:i]
    el, xyz, labels = [], [], []
    for i, res in enumerate(chain):
        e, x = atoms_of(res)
        el += e
        xyz.append(x)
        labels += ["%s%d:%s" % (res.aa, i + 1, k) for k in res.heavy] + ["%s%d:H(%s)" % (res.aa, i + 1, p) for _, p in res.hyd]
    xyz = np.concatenate(xyz)
    generate.labels = labels
    return el, xyz


def formula(el):
    from collections import Counter
    c = Counter(el)
    return "".join("%s%d" % (k, c[k]) for k in ("C", "H", "N", "O", "S"))


def main():
    el, xyz = generate(seed=1)
    f = formula(el)
    assert f == "C202H321N55O64S6", f
    d = np.linalg.norm(xyz[:, None] - xyz[None, :], axis=2) + 10 * np.eye(len(el))
    out = {"name": "polypeptide_1crn", "comment": "crambin-sequence random coil, seed 1 (tools/make_polypeptide.py)", "formula": f,
           "charge": 0, "uhf": 0, "numbers": [ZNUM[e] for e in el], "xyz": (xyz * AA2BOHR).round(10).tolist(),
           "labels": generate.labels, "min_distance_angstrom": float(d.min()), "extent_angstrom": (xyz.max(0) - xyz.min(0)).round(2).tolist()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "polypeptide_1crn.json")
    with open(path, "w") as fh:
        json.dump(out, fh)
    print(f, len(el), "atoms; min distance %.3f A; extent" % d.min(), out["extent_angstrom"], file=sys.stderr)


if __name__ == "__main__":
    main()
