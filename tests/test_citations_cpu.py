"""CPU (needs /root/reference, skipped elsewhere): every `file:line` citation of the reference in the C header, the oracle,
the design / integration documents and the source comments must name an existing reference file and a line range inside it."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PAT = re.compile(r"(?<![\w/.\-])((?:[\w.\-]+/)*[\w.\-]+\.(?:h|cpp|cu|cmake|msp)|CMakeLists\.txt):(\d+)(?:-(\d+))?")


def _sources():
    out = ["include/onika_b200.h", "DESIGN.md", "INTEGRATION.md", "README.md", "oracle/oracle.c", "oracle/oracle.h", "oracle/ref_driver.cpp",
           "oracle/ref_xs_driver.cpp", "oracle/__init__.py", "bench.py"]
    for sub in ("include/onika", "onika_b200", "tests"):
        for root, _, fs in os.walk(os.path.join(ROOT, sub)):
            if "_refplugins" in root or "__pycache__" in root:
                continue
            for f in fs:
                if f.endswith((".h", ".py", ".cu", ".cuh", ".md")):
                    out.append(os.path.relpath(os.path.join(root, f), ROOT))
    return out


def test_reference_citations_resolve():
    if not os.path.isdir(REF):
        pytest.skip("/root/reference is not present")
    index = {}
    for root, _, fs in os.walk(REF):
        if "/.git" in root:
            continue
        for f in fs:
            index.setdefault(f, []).append(os.path.relpath(os.path.join(root, f), REF))
    lines = {}
    bad, checked = [], 0
    for src in _sources():
        text = open(os.path.join(ROOT, src), errors="replace").read()
        for m in PAT.finditer(text):
            path, a, b = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            cands = [p for p in index.get(os.path.basename(path), []) if p.endswith(path)]
            if not cands:
                if os.path.exists(os.path.join(ROOT, path)) or any(os.path.basename(path) == os.path.basename(s) for s in _OWN):
                    continue                                    # a citation of this repository's own file
                bad.append((src, m.group(0), "no such reference file"))
                continue
            checked += 1
            ok = False
            for c in cands:
                if c not in lines:
                    with open(os.path.join(REF, c), errors="replace") as f:
                        lines[c] = sum(1 for _ in f)
                ok = ok or (a <= b <= lines[c])
            if not ok:
                bad.append((src, m.group(0), f"beyond the end of {cands}"))
    assert checked > 200 and not bad, bad[:40]


_OWN = [f for _, _, fs in os.walk(ROOT) for f in fs]
