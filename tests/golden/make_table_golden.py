#!/usr/bin/env python3
"""Independent source for the element tables: the reference ships the GFN2-xTB parameters twice, as Fortran data statements
(src/xtb/gfn2.f90, from which tools/gen_tables.py generates xtb_b200/csrc/tables/gfn2_tables.h) and as the parameter file
param_gfn2-xtb.txt read by src/read_gfn_param.f90.  This script parses the FILE (its own format, its own parser) into
tests/golden/gfn2_param_file.json; tests/test_tables_independent.py compares the generated header against it for all 86
elements, so that a transcription error for an element outside the known-answer tests cannot pass unnoticed.

    python tests/golden/make_table_golden.py [/root/reference]
"""
import json
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out = {"source": "param_gfn2-xtb.txt (grimme-lab/xtb v6.7.1), scalings of src/read_gfn_param.f90:529-570", "globals": {}, "elements": {}}
section, z = None, None
with open(os.path.join(ref, "param_gfn2-xtb.txt")) as f:
    for raw in f:
        line = raw.strip()
        if not line:
            continue
        if line.startswith("$"):
            m = re.match(r"\$Z=\s*(\d+)", line)
            if m:
                section, z = "Z", int(m.group(1))
                out["elements"][str(z)] = {}
            else:
                section, z = line[1:].split()[0], None
            continue
        if section == "globpar":
            k, v = line.split()
            out["globals"][k] = float(v)
        elif section == "Z":
            k, v = line.split("=", 1)
            k = k.strip().lower()
            if k == "ao":
                shells = re.findall(r"(\d)([spdf])", v.strip())
                out["elements"][str(z)]["pqn"] = [int(n) for n, _ in shells]
                out["elements"][str(z)]["ang"] = ["spdf".index(l) for _, l in shells]
            else:
                out["elements"][str(z)][k] = [float(x) for x in v.split()]
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gfn2_param_file.json")
with open(path, "w") as f:
    json.dump(out, f, indent=0, sort_keys=True)
print("wrote", path, len(out["elements"]), "elements")
