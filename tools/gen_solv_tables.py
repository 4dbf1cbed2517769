#!/usr/bin/env python3
"""Transcribe the numeric tables of the implicit solvation models (GBSA / ALPB) of the reference
(grimme-lab/xtb v6.7.1) into a generated C header.  Data only: no reference code is copied.

    python tools/gen_solv_tables.py /root/reference xtb_b200/csrc/tables

Sources (file:line of the reference):
  include/param_gbsa_<solvent>.fh    gfn2_<solvent>       GBSA parameter sets (Still kernel, src/solv/model.f90:289-307)
  include/param_alpb_<solvent>.fh    gfn2_alpb_<solvent>  ALPB parameter sets (P16 kernel, src/solv/model.f90:325-355)
  src/param/vdwradd3.f90:37-61       vanDerWaalsRadD3 (Angstrom)
  src/solv/lebedev.f90:542-577       ld0230: orbit parameters of the 230 point Lebedev-Laikov grid; the points are
                                     expanded here in the order of genOh1..genOh6 (lebedev.f90:126-317)
and, for the molecular dynamics driver (md_tables.h):
  src/param/atomicmass.f90:37-67     atomicMass (amu)
  src/shake_module.f90:36-41         metal(1:86) flags of init_shake
"""
import glob
import math
import os
import re
import sys

NUM = re.compile(r'[-+]?(?:\d+\.\d*|\.\d+)(?:[eEdD][-+]?\d+)?(?=_wp)')


def numbers(text):
    return [float(t.replace('d', 'e').replace('D', 'e')) for t in NUM.findall(text)]


def param_sets(ref, kind):
    """kind: 'gbsa' -> gfn2_<x>, 'alpb' -> gfn2_alpb_<x>.  Returns {normalised solvent name: 290 numbers}."""
    out = {}
    for path in sorted(glob.glob(os.path.join(ref, 'include', 'param_%s_*.fh' % kind))):
        text = open(path).read()
        blocks = re.split(r'(?=type\(gbsa_parameter\),parameter ::)', text)
        for b in blocks:
            m = re.match(r'type\(gbsa_parameter\),parameter :: (\w+) = gbsa_parameter', b)
            if not m:
                continue
            name = m.group(1)
            prefix = 'gfn2_alpb_' if kind == 'alpb' else 'gfn2_'
            if not name.startswith(prefix) or (kind == 'gbsa' and name.startswith('gfn2_alpb_')):
                continue
            vals = numbers(b)
            assert len(vals) == 8 + 3 * 94, (name, len(vals))
            key = name[len(prefix):]
            # names after normalizeSolventName (src/solv/model.f90:211-246) as matched in loadInternalParam
            key = {'h2o': 'water', 'nhexan': 'hexane'}.get(key, key)
            out[key] = vals
    return out


def vdw_radii(ref):
    text = open(os.path.join(ref, 'src', 'param', 'vdwradd3.f90')).read()
    m = re.search(r'vanDerWaalsRadD3\(1:94\) = aatoau \* \[(.*?)\]', text, re.S)
    v = numbers(m.group(1))
    assert len(v) == 94
    return v


def lebedev230(ref):
    text = open(os.path.join(ref, 'src', 'solv', 'lebedev.f90')).read()
    m = re.search(r'pure subroutine ld0230\(x, w, n\)(.*?)end subroutine ld0230', text, re.S)
    a = b = v = None
    pts = []

    def emit(rows, wt):
        for r in rows:
            pts.append((r[0], r[1], r[2], wt))

    def signs3(p):  # sign patterns in the order the reference writes them: first coordinate toggles fastest
        res = []
        for s3 in (1, -1):
            for s2 in (1, -1):
                for s1 in (1, -1):
                    res.append((s1 * p[0], s2 * p[1], s3 * p[2]))
        return res

    for line in m.group(1).splitlines():
        line = line.strip()
        mm = re.match(r'([abv]) = (\S+)_wp', line)
        if mm:
            val = float(mm.group(2))
            if mm.group(1) == 'a':
                a = val
            elif mm.group(1) == 'b':
                b = val
            else:
                v = val
            continue
        mm = re.match(r'call genOh(\d)', line)
        if not mm:
            continue
        code = int(mm.group(1))
        z = 0.0
        if code == 1:
            emit([(1, z, z), (-1, z, z), (z, 1, z), (z, -1, z), (z, z, 1), (z, z, -1)], v)
        elif code == 3:
            c = math.sqrt(1.0 / 3.0)
            emit(signs3((c, c, c)), v)
        elif code == 4:
            c = math.sqrt(1.0 - 2.0 * a * a)
            emit(signs3((a, a, c)) + signs3((a, c, a)) + signs3((c, a, a)), v)
        elif code == 5:
            c = math.sqrt(1.0 - a * a)

            def s2(p, q, pos):  # two non-zero entries p, q placed at positions pos, four sign patterns
                res = []
                for sq in (1, -1):
                    for sp in (1, -1):
                        r = [z, z, z]
                        r[pos[0]] = sp * p
                        r[pos[1]] = sq * q
                        res.append(tuple(r))
                return res
            emit(s2(a, c, (0, 1)) + s2(c, a, (0, 1)) + s2(a, c, (0, 2)) + s2(c, a, (0, 2))
                 + s2(a, c, (1, 2)) + s2(c, a, (1, 2)), v)
        elif code == 6:
            c = math.sqrt(1.0 - a * a - b * b)
            emit(signs3((a, b, c)) + signs3((a, c, b)) + signs3((b, a, c)) + signs3((b, c, a))
                 + signs3((c, a, b)) + signs3((c, b, a)), v)
        else:
            raise SystemExit('unexpected orbit code %d' % code)
    assert len(pts) == 230, len(pts)
    assert abs(sum(p[3] for p in pts) - 1.0) < 1e-13
    for p in pts:
        assert abs(p[0] ** 2 + p[1] ** 2 + p[2] ** 2 - 1.0) < 1e-14
    return pts


def md_tables(ref, outdir):
    text = open(os.path.join(ref, 'src', 'param', 'atomicmass.f90')).read()
    m = re.search(r'atomicMass\(1:118\) = amutoau \* \[(.*?)\]', text, re.S)
    mass = numbers(m.group(1))
    assert len(mass) == 118
    text = open(os.path.join(ref, 'src', 'shake_module.f90')).read()
    m = re.search(r'metal\(1:86\) = \[(.*?)\]', text, re.S)
    metal = [int(t) for t in re.findall(r'\b[01]\b', m.group(1))]
    assert len(metal) == 86
    with open(os.path.join(outdir, 'md_tables.h'), 'w') as f:
        f.write('// GENERATED by tools/gen_solv_tables.py from grimme-lab/xtb v6.7.1 parameter data. Do not edit.\n')
        f.write('#pragma once\n\n// atomic masses in amu (src/param/atomicmass.f90)\n')
        f.write('static const double md_atomic_mass_amu[118] = {%s};\n\n' % ', '.join(fmt(x) for x in mass))
        f.write('// metal flags of the SHAKE bond selection (src/shake_module.f90:36-41)\n')
        f.write('static const int md_metal[86] = {%s};\n' % ', '.join(str(x) for x in metal))
    print('wrote md_tables.h')


def fmt(v):
    return repr(float(v))


def main():
    ref, outdir = sys.argv[1], sys.argv[2]
    gbsa = param_sets(ref, 'gbsa')
    alpb = param_sets(ref, 'alpb')
    vdw = vdw_radii(ref)
    leb = lebedev230(ref)
    with open(os.path.join(outdir, 'solv_tables.h'), 'w') as f:
        f.write('// GENERATED by tools/gen_solv_tables.py from grimme-lab/xtb v6.7.1 parameter data. Do not edit.\n')
        f.write('#pragma once\n\n')
        f.write('// one internal GFN2-xTB parameter set: kernel 1 = Still (GBSA sets), 2 = P16 (ALPB sets)\n')
        f.write('struct SolvParamSet {\n    const char* name;\n    int kernel;\n'
                '    double epsv, smass, rhos, c1, rprobe, gshift, soset, alpha;\n'
                '    double gamscale[94], sx[94], tmp[94];\n};\n\n')
        f.write('static const int solv_nsets = %d;\n' % (len(gbsa) + len(alpb)))
        f.write('static const SolvParamSet solv_param_sets[%d] = {\n' % (len(gbsa) + len(alpb)))
        for kern, sets in ((1, gbsa), (2, alpb)):
            for name in sorted(sets):
                v = sets[name]
                f.write('  {"%s", %d, %s,\n' % (name, kern, ', '.join(fmt(x) for x in v[:8])))
                for k in range(3):
                    f.write('   {%s}%s\n' % (', '.join(fmt(x) for x in v[8 + 94 * k: 8 + 94 * (k + 1)]), ',' if k < 2 else ''))
                f.write('  },\n')
        f.write('};\n\n')
        f.write('// D3 van der Waals radii in Angstrom (multiplied by aatoau where they are used)\n')
        f.write('static const double solv_vdw_rad_d3_aa[94] = {%s};\n\n' % ', '.join(fmt(x) for x in vdw))
        f.write('// 230 point Lebedev-Laikov grid: x, y, z, weight (weights sum to 1)\n')
        f.write('static const int solv_nang = 230;\n')
        f.write('static const double solv_leb230[230][4] = {\n')
        for p in leb:
            f.write('  {%s},\n' % ', '.join(fmt(x) for x in p))
        f.write('};\n')
    print('wrote solv_tables.h: %d GBSA sets, %d ALPB sets' % (len(gbsa), len(alpb)))
    md_tables(ref, outdir)


if __name__ == '__main__':
    main()
