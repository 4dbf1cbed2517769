#!/usr/bin/env python3
"""Extract the geometries and known-answer values that the reference's own tests hold for the
GFN2-xTB single point path into tests/golden/reference_kat.json.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

Sources: test/unit/molstock.f90 (geometries), test/unit/test_gfn2.f90:48-235,667-742,831-901,
987-1091 and (implicit solvation) 237-385,745-828,
862-900, test/api/c_api_example.c:30-163,298-419, test/unit/test_coulomb.f90:828-1079,
test/unit/test_repulsion.f90, test/unit/test_eeq.f90.
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
SYM = ("H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr "
       "Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu "
       "Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn").split()
NUM = re.compile(r'[-+]?\d+\.\d*(?:[eEdD][-+]?\d+)?(?=_wp)')


def molstock(name):
    src = open(os.path.join(REF, "test/unit/molstock.f90")).read()
    m = re.search(r'subroutine\s+%s\s*\(mol\)(.*?)end subroutine' % re.escape(name), src, re.S | re.I)
    body = m.group(1)
    ms = re.search(r'sym\(nat\)\s*=\s*\[character\(len=\d+\)\s*::(.*?)\]', body, re.S)
    if ms:
        syms = re.findall(r'"\s*([A-Za-z]+)\s*"', ms.group(1))
        numbers = [SYM.index(s.capitalize()) + 1 for s in syms]
    else:
        ma = re.search(r'at\(nat\)\s*=\s*&?\s*\[(.*?)\]', body, re.S)
        numbers = [int(t) for t in re.findall(r'\d+', ma.group(1))]
    xyz = re.search(r'xyz\(3,\s*nat\)\s*=\s*reshape\(\[(.*?)\]', body, re.S).group(1)
    xyz = [float(t.replace('d', 'e').replace('D', 'e')) for t in NUM.findall(xyz)]
    chrg = re.search(r'charge\s*=\s*([-+]?\d+\.?\d*)', body)
    uhf = re.search(r'uhf\s*=\s*(\d+)', body)
    assert len(xyz) == 3 * len(numbers), name
    return dict(numbers=numbers, xyz=xyz, charge=float(chrg.group(1)) if chrg else 0.0,
                uhf=int(uhf.group(1)) if uhf else 0)


# test_gfn2.f90:989 takes aatoau from mctc-lib (mctc_io_convert: CODATA 2018 Bohr radius), not from xtb
AATOAU = 1.0 / 0.529177210903
out = {}
for i in range(1, 11):
    out["mindless%02d" % i] = molstock("mindless%02d" % i)
for nm in ["caffeine", "rivaroxaban", "pdb_4qxx", "feco5", "co_cnx6", "fe_cnx6", "remdesivir", "taxol"]:
    out[nm] = molstock(nm)
for nm in ["co_cnx6", "fe_cnx6"]:  # given in Angstrom, test_gfn2.f90:1029,1064 multiply by aatoau
    out[nm]["xyz"] = [x * AATOAU for x in out[nm]["xyz"]]

e = [-30.348902328339, -24.069929673504, -23.718708414840, -22.751497890244, -27.735987617741,
     -18.559531652263, -33.423246445911, -29.982410709899, -20.549458882257, -25.648093031152]
g = [0.06081938165963, 0.05480379303387, 0.04179957306749, 0.06932568214054, 0.05133081215262,
     0.05897000784363, 0.04600837003153, 0.06088718433105, 0.05314451272641, 0.04588800056975]
h = [2.2611412120964, 1.2455369848896, 0.1490345448030, 1.2694246332861, 2.2362028272626,
     2.6053844166160, 2.6312636988491, 0.8502228061534, 3.2960732039022, 0.4185654147579]
for i in range(10):
    out["mindless%02d" % (i + 1)]["kat"] = dict(energy=e[i], gnorm=g[i], hl_gap=h[i], tol=1e-8,
                                                source="test/unit/test_gfn2.f90:698-712")
out["feco5"]["kat"] = dict(energy=-34.066853522474, gnorm=0.26985488354859, hl_gap=2.9239308006972, tol=1e-8,
                           source="test/unit/test_gfn2.f90:862-867")
out["co_cnx6"]["kat"] = dict(energy=-34.191439852632, gnorm=0.041770673817, hl_gap=3.769023080904, tol=1e-8,
                             wbo={"0,1": 0.889766438284303, "2,3": 1.728632460261216E-002, "1,9": 2.73380919723017},
                             source="test/unit/test_gfn2.f90:1047-1054")
out["fe_cnx6"]["kat"] = dict(energy=-33.314958144107, gnorm=0.018945181484, hl_gap=1.805948321662, tol=1e-8,
                             wbo={"0,1": 0.587095913328687, "2,3": 3.798539382576892E-002, "1,9": 2.81892857328157},
                             source="test/unit/test_gfn2.f90:1081-1089")
out["h2o_scc"] = dict(
    numbers=[8, 1, 1],
    xyz=[0.0, 0.0, -0.73578586109551, 1.44183152868459, 0.0, 0.36789293054775, -1.44183152868459, 0.0, 0.36789293054775],
    charge=0.0, uhf=0, guess="sad", maxiter=50,
    kat=dict(energy=-5.070451355118, gnorm=0.006457420125, hl_gap=14.450372368833, e_elec=-5.104362813671,
             e_es=0.031408028914, e_aes=0.000563300591, e_axc=-0.000813611055, e_disp=-0.000141250480,
             e_rep=0.033911458523, q1=0.28158903353422, qsh0=0.25886041477578, dipm_2_0=-0.55238805565739E-01,
             qp_0_5=0.41207363331933E-01, qp_2_0=0.95084219046109E-01, emo_homo=-12.166283951806,
             nshell=4, nao=6, ihomo=4, tol=1e-7, source="test/unit/test_gfn2.f90:121-164"))
out["c3h4_api"] = dict(
    numbers=[6, 6, 6, 1, 1, 1, 1],
    xyz=[0.0, 0.0, -1.79755622305860, 0.0, 0.0, 0.95338756106749, 0.0, 0.0, 3.22281255790261,
         -0.96412815539807, -1.66991895015711, -2.53624948351102, -0.96412815539807, 1.66991895015711, -2.53624948351102,
         1.92825631079613, 0.0, -2.53624948351102, 0.0, 0.0, 5.23010455462158],
    charge=0.0, uhf=0, maxiter=30,
    kat=dict(energy=-8.3824793849585, gnorm=0.11544410028854E-01, hl_gap=7.0005867526665,
             g_2_1=0.0, g_0_2=-0.74649908147372E-03, g_3_0=-0.28433755158510E-03, g_6_2=0.99750545315944E-02,
             q5=0.05184019996829, dipole2=-0.298279305689518, wbo9=2.89823984265213, tol=1e-9,
             source="test/unit/test_gfn2.f90:226-233; test/api/c_api_example.c:87-112"))
# water dimer in the field of a second dimer given as external point charges: test/api/c_api_example.c:298-419
out["h2o2_pcem_api"] = dict(
    numbers=[8, 1, 1, 8, 1, 1],
    xyz=[-2.75237178376284, 2.43247309226225, -0.01392519847964, -0.93157260886974, 2.79621404458590, -0.01863384029005,
         -3.43820531288547, 3.30583608421060, 1.42134539425148, -2.43247309226225, -2.75237178376284, 0.01392519847964,
         -2.79621404458590, -0.93157260886974, 0.01863384029005, -3.30583608421060, -3.43820531288547, -1.42134539425148],
    charge=0.0, uhf=0,
    pc_numbers=[8, 1, 1, 8, 1, 1],
    pc_charges=[-0.69645733, 0.36031084, 0.33614649, -0.69645733, 0.36031084, 0.33614649],
    pc_xyz=[2.75237178376284, -2.43247309226225, -0.01392519847964, 0.93157260886974, -2.79621404458590, -0.01863384029005,
            3.43820531288547, -3.30583608421060, 1.42134539425148, 2.43247309226225, 2.75237178376284, 0.01392519847964,
            2.79621404458590, 0.93157260886974, 0.01863384029005, 3.30583608421060, 3.43820531288547, -1.42134539425148],
    kat=dict(energy=-10.160927754235, g_4_0=-2.0326750053811E-03, g_1_1=3.3125460324302E-03, g_3_0=-1.1929405354823E-03,
             g_5_2=-1.6607683158587E-03, pcg_4_0=-0.000248319582, pcg_1_1=0.001420844475, pcg_3_0=0.003746685275,
             pcg_5_2=0.000651613444, tol_energy=1e-9, tol_grad=1e-8, source="test/api/c_api_example.c:298-419"))
# ---- implicit solvation (GBSA / ALPB), SURVEY 8f row f2
# C API sequence on the allene molecule above: gas phase, then GBSA(h2o), then ALPB(h2o), each restarted from the
# previous wavefunction (test/api/c_api_example.c:114-163)
out["c3h4_api"]["kat_solv"] = dict(
    gbsa=dict(solvent="h2o", energy=-8.38393864716134, q5=0.06090868805034, dipole2=-0.35455233974705, wbo9=2.89453979224265),
    alpb=dict(solvent="h2o", energy=-8.384076843892, q5=0.0644849340, dipole2=-0.3641866008, wbo9=2.8932146955),
    tol_energy=1e-9, tol_q=1e-8, tol_dipole=1e-6, tol_wbo=1e-8, source="test/api/c_api_example.c:114-163")
out["gbsa_h2o_api"] = dict(
    numbers=[7, 1, 1, 1, 15, 6, 6, 6, 7, 7, 7],
    xyz=[4.84112686339230, -1.15671648374837, 0.00000000000000, 5.87217277985151, -0.70323701391366, 1.54511189584035,
         5.87217277985151, -0.70323701391366, -1.54511189584035, 4.61871932563459, -3.05696083228965, 0.00000000000000,
         0.23751589819013, 1.28722286021158, 0.00000000000000, 0.00161949529588, -0.89075264835377, 2.57562301980395,
         -3.06580664031769, 2.16487781673832, 0.00000000000000, 0.00161949529588, -0.89075264835377, -2.57562301980395,
         -5.06585309511506, 3.00827014874972, 0.00000000000000, 0.00161949529588, -2.15043086045241, 4.34183098346579,
         0.00161949529588, -2.15043086045241, -4.34183098346579],
    charge=0.0, uhf=0, maxiter=30, solvent="h2o", model="gbsa", kernel="still",
    kat=dict(hl_gap=3.408607724814, energy=-22.002501380096, gnorm=0.19366866479022E-01, g_0_0=.9038674439127e-02,
             g_1_2=-.1394693523214e-02, g_7_0=0.17878350605942E-02, tol=5e-7, tol_gap=1e-5,
             source="test/unit/test_gfn2.f90:237-311"))
out["gbsa_salt_api"] = dict(
    numbers=[7, 7, 9, 1, 1, 1, 1, 1],
    xyz=[0.00000000000000, 1.30633555031959, 0.00000000000000, -1.26402080275201, -4.06871453465823, 0.00000000000000,
         1.42937939828404, 3.57212474255400, 0.00000000000000, -2.92500314615638, -3.12058224200551, 0.00000000000000,
         0.71133449039308, 0.35812010971703, 1.51289962417918, -1.25697212427635, -5.20406577474818, -1.53711646457211,
         0.71133449039308, 0.35812010971703, -1.51289962417918, -1.25697212427635, -5.20406577474818, 1.53711646457211],
    charge=0.0, uhf=0, maxiter=30, solvent="ch2cl2", model="gbsa", kernel="still", ion_strength=1.0e-3,
    ion_rad_angstrom=1.0,
    kat=dict(hl_gap=6.895830675032, energy=-13.027106170796, gnorm=0.014655633557, g_0_0=-0.6696141112967e-02,
             g_3_1=-0.7608884457863e-03, g_4_0=0.2359793849677e-02, g_6_2=0.6882248903260e-02, tol=5e-7, tol_gap=5e-5,
             source="test/unit/test_gfn2.f90:313-385"))
# ten "mindless" molecules in ten solvents, GBSA parameter sets with the Still kernel, ALPB switched on for every second
# molecule (test/unit/test_gfn2.f90:745-828)
ms_solv = ["h2o", "chcl3", "thf", "acetonitrile", "toluene", "ch2cl2", "ether", "methanol", "cs2", "dmso"]
ms_e = [-30.384994095620, -24.092790321708, -23.760000587724, -22.759033676135, -27.755756860587, -18.578906535415,
        -33.448544467530, -29.993777657049, -20.567481283491, -25.665458003353]
ms_g = [0.064382736361151, 0.055824257363022, 0.040696787567357, 0.073197307180171, 0.049852029310233,
        0.053051084845551, 0.046463198340453, 0.057936724779824, 0.055081716352177, 0.045339090185313]
ms_h = [2.5109978045214, 1.3559515037838, 0.1719065524538, 1.2428812469179, 2.1428443722148, 2.2548300583012,
        2.6935601499212, 0.83100059690964, 3.3270919380043, 0.3859217429034]
for i in range(10):
    out["mindless%02d" % (i + 1)]["kat_solv"] = dict(
        solvent=ms_solv[i], model="alpb" if (i + 1) % 2 == 0 else "gbsa", kernel="still", energy=ms_e[i], gnorm=ms_g[i],
        hl_gap=ms_h[i], tol=1e-8, source="test/unit/test_gfn2.f90:745-828")
# Fe(CO)5 in CH2Cl2: GBSA and ALPB, both with the Still kernel (test/unit/test_gfn2.f90:831-901)
out["feco5"]["kat_solv"] = dict(
    solvent="ch2cl2", kernel="still", tol=1e-8, source="test/unit/test_gfn2.f90:862-900",
    gbsa=dict(energy=-34.078580189077, gnorm=0.26591831790033, hl_gap=2.9163891345974),
    alpb=dict(energy=-34.078529449408, gnorm=0.26613440125774, hl_gap=2.9168589130125))
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kat.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(out), "systems")
