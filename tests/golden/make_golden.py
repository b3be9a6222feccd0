"""Generates tests/golden/*.json from the reference's own test sources (run in the build container,
where /root/reference is mounted; the GPU box only sees the committed JSON).

  testkron01.json   <- tests/UnitTests_DMRGKron.cpp:39-252  (TestKron01: inputs via SetRow, expected
                       rows via CheckRow; SetRow stores value == column index, tests/UnitTests_Misc.cpp:15-18)
  opblocks.json     <- tests/UnitTests_Misc.cpp:82-136 + tests/UnitTests_DMRGBlock.cpp:76-131
                       (sector-respecting / sector-violating operators for CheckOperatorBlocks)
"""
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def ints(s):
    return [int(x) for x in re.findall(r"-?\d+", s)]


def testkron01():
    src = open(os.path.join(REF, "tests/UnitTests_DMRGKron.cpp")).read()
    body = src[src.index("PetscErrorCode TestKron01()\n{"):src.index("PetscErrorCode TestKron02()\n{")]
    out = {"source": "tests/UnitTests_DMRGKron.cpp:39-252", "blocks": {}, "expected": []}
    for name, nsites, qn, sizes in re.findall(r"(\w+)Block\.Initialize\(PETSC_COMM_WORLD,\s*(\d+),\s*\{([^}]*)\},\s*\{([^}]*)\}\)", body):
        out["blocks"][name] = {"nsites": int(nsites), "qn": [float(x) for x in qn.split(",")], "sizes": ints(sizes), "rows": []}
    for name, op, site, row, cols in re.findall(r"SetRow\(\s*(\w+)Block\.(S[zp])\((\d+)\),\s*(\d+),\s*\{([^}]*)\}\);", body):
        out["blocks"][name]["rows"].append({"op": op, "site": int(site), "row": int(row), "cols": ints(cols)})
    for op, site, row, cols, vals in re.findall(
            r"CheckRow\(BlockOut\.(S[zp])\((\d+)\),\s*\"[^\"]*\",\s*(\d+),\s*\{([^}]*)\},\s*\{([^}]*)\}\)", body):
        out["expected"].append({"op": op, "site": int(site), "row": int(row), "cols": ints(cols), "vals": [float(v) for v in ints(vals)]})
    assert len(out["blocks"]) == 2 and len(out["expected"]) == 120, (len(out["blocks"]), len(out["expected"]))
    return out


def opblocks():
    src = open(os.path.join(REF, "tests/UnitTests_Misc.cpp")).read()
    out = {"source": "tests/UnitTests_Misc.cpp:82-136, tests/UnitTests_DMRGBlock.cpp:76-131",
           "sectors": {"qn": [1.5, 0.5, -0.5, -1.5], "sizes": [2, 3, 2, 1]}, "ops": {}}
    for fn in ("SetSz0", "SetSp0", "SetSz1", "SetSp1"):
        body = src[src.index("PetscErrorCode %s(" % fn):]
        body = body[:body.index("return ierr;")]
        out["ops"][fn] = [{"row": int(r), "cols": ints(c)} for r, c in re.findall(r"SetRow\(\w+,\s*(\d+),\s*\{([^}]*)\}\)", body)]
    blk = open(os.path.join(REF, "tests/UnitTests_DMRGBlock.cpp")).read()
    # tests/UnitTests_DMRGBlock.cpp:76-131 — Sz(0), Sp(0) respect the sectors; Sz(1) row 7 does not and
    # MatOpCheckOperatorBlocks must return PETSC_ERR_ARG_OUTOFRANGE (= 63 in PETSc 3.8 petscerror.h)
    fn = blk[blk.index("PetscErrorCode Test_MatOpCheckOperatorBlocks()"):]
    fn = fn[:fn.index("return ierr;")]
    cases = {}
    for op, site, row, cols in re.findall(r"SetRow\(blk\.(S[zp])\((\d)\),\s*(\d+),\s*\{([^}]*)\}\)", fn):
        cases.setdefault("%s(%s)" % (op, site), []).append({"row": int(row), "cols": ints(cols)})
    out["check_cases"] = cases
    out["check_expect"] = {"Sz(0)": 0, "Sp(0)": 0, "Sz(1)": 63}
    assert "PETSC_ERR_ARG_OUTOFRANGE" in fn and len(cases) == 3
    return out


if __name__ == "__main__":
    json.dump(testkron01(), open(os.path.join(HERE, "testkron01.json"), "w"), indent=1)
    json.dump(opblocks(), open(os.path.join(HERE, "opblocks.json"), "w"), indent=1)
    print("written")
