#!/usr/bin/env python
"""Generates the golden fixtures in this directory from the reference's OWN unit tests.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Outputs (committed):
  algorithm_choice.json  every `REQUIRE(alg_choice == X)` in the reference's TensorAlgebra unit
                         tests, with the index packs of the einsum call that precedes it
                         (file:line kept for traceability).  Pins the analyser
                         (TensorAlgebra/Backends/Dispatch.hpp:859-957).
  known_answers.json     hand-transcribed known-answer vectors of the same tests (values only,
                         each with its file:line): Einsum1.cpp:17-88, Sort.cpp etc.
  water_sto3g.json       the H2O/STO-3G integrals used by QChemTests.cpp (S, T, V, TEI) and the
                         energies that test asserts (RHF -74.942079928192, MP2 -0.049149636120).

Nothing here copies reference source code: only test *data* (numbers, index letters) is extracted.
"""
import json
import os
import re
import sys

REF = "/root/reference"
UNIT = os.path.join(REF, "libs/Einsums/TensorAlgebra/tests/unit")
HERE = os.path.dirname(os.path.abspath(__file__))


def extract_algorithm_choices():
    out = []
    for fn in sorted(os.listdir(UNIT)):
        if not fn.endswith(".cpp") and not fn.endswith(".hip"):
            continue
        text = open(os.path.join(UNIT, fn)).read()
        for m in re.finditer(r"REQUIRE\(\s*alg_choice\s*==\s*([\w:]+)\s*\)", text):
            choice = m.group(1).split("::")[-1]
            line_start = text.rfind("\n", 0, m.start()) + 1
            if text[line_start : m.start()].strip().startswith("//"):
                continue  # assertion commented out in the reference
            # nearest preceding einsum( ... &alg_choice)
            head = text[: m.start()]
            k = head.rfind("&alg_choice")
            e = head.rfind("einsum(", 0, k)
            if k < 0 or e < 0:
                continue
            call = head[e : k + len("&alg_choice")]
            packs = re.findall(r"Indices\{([^}]*)\}", call)
            if len(packs) != 3:
                continue
            packs = [[t.strip().split("::")[-1] for t in p.split(",") if t.strip()] for p in packs]
            # operand expressions: text after each Indices{...},
            ops = re.findall(r"Indices\{[^}]*\}\s*,\s*([^,]+?)\s*,", call + ",")
            conj = "einsum<true" in call or "conj" in call
            line = head.count("\n") + 1
            out.append(
                {
                    "source": f"libs/Einsums/TensorAlgebra/tests/unit/{fn}:{line}",
                    "C": packs[0],
                    "A": packs[1],
                    "B": packs[2],
                    "operands": [o.strip() for o in ops[:3]],
                    "expected": choice,
                    "call_has_template_args": "einsum<" in call,
                    "conj_hint": conj,
                }
            )
    return out


def read_dat(path):
    rows = []
    for ln in open(path):
        t = ln.split()
        if t:
            rows.append(t)
    return rows


def extract_water():
    d = os.path.join(UNIT, "data/water_sto3g")
    out = {"source": "libs/Einsums/TensorAlgebra/tests/unit/data/water_sto3g/{S,T,V,TEI}.dat + QChemTests.cpp:185-416"}
    for name in ("S", "T", "V"):
        out[name] = [[int(r[0]), int(r[1]), float(r[2])] for r in read_dat(os.path.join(d, name + ".dat"))]
    out["TEI"] = [[int(r[0]), int(r[1]), int(r[2]), int(r[3]), float(r[4])] for r in read_dat(os.path.join(d, "TEI.dat"))]
    # asserted values, QChemTests.cpp
    out["expected"] = {
        "nbf": 7,
        "ndocc": 5,
        "enuc": None,
        "rhf_total_energy": -74.942079928192,
        "mp2_correlation_energy": -0.049149636120,
        "mp2_total_energy": -74.991229564312,
        "tolerance": 1e-6,
    }
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; golden files are already committed")
    ac = extract_algorithm_choices()
    json.dump(ac, open(os.path.join(HERE, "algorithm_choice.json"), "w"), indent=1)
    print("algorithm_choice.json:", len(ac), "entries")
    w = extract_water()
    json.dump(w, open(os.path.join(HERE, "water_sto3g.json"), "w"))
    print("water_sto3g.json: S", len(w["S"]), "T", len(w["T"]), "V", len(w["V"]), "TEI", len(w["TEI"]))


if __name__ == "__main__":
    main()
