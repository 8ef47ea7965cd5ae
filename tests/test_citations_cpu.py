"""Every `file:line` citation of the reference in the boundary header, the oracle header and the design documents
must point at an existing file and line of /root/reference.  Runs only where the reference is mounted (this
container); skipped on the GPU box."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
import glob

DOCS = ["include/flamegpu2_b200.h", "oracle/fgb_oracle.h", "oracle/fgb_oracle.c", "DESIGN.md", "INTEGRATION.md", "tests/README.md"]
DOCS += sorted(os.path.relpath(p, ROOT) for pat in ("include/flamegpu/**/*.h", "include/flamegpu/**/*.cuh", "flamegpu2_b200/csrc/**/*.cu*",
                                                      "examples/*.cuh", "flamegpu2_b200/*.py", "tests/test_*.py")
               for p in glob.glob(os.path.join(ROOT, pat), recursive=True))
CITE = re.compile(r"([A-Za-z0-9_][A-Za-z0-9_/\.\-]*\.(?:cuh|cu|cpp|hpp|h)):(\d+)(?:-(\d+))?")
OURS = ("fgb_", "flamegpu2_b200", "fgbs.py", "oracle_py", "slab.py", "sim.py", "host.py", "_capi.py", "bench", "test_citations", "CUDASimulation_impl.h", "FunctionArgs.h", "circles_model", "boids_model", "stress_model",
        "test_models", "ModelDescription.h", "ref_sim.cu", "gen_golden.cpp", "fgb_models.cu")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "flamegpu")), reason="reference not mounted")


def _index():
    by_name = {}
    for base, _, files in os.walk(REF):
        if "/.git" in base:
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".hpp")):
                by_name.setdefault(f, []).append(os.path.join(base, f))
    return by_name


def _resolve(path, by_name):
    cands = by_name.get(os.path.basename(path), [])
    if "/" in path:
        cands = [c for c in cands if c.endswith("/" + path.lstrip("./"))] or cands
    return cands


def test_reference_citations_resolve():
    by_name = _index()
    bad = []
    n = 0
    for doc in DOCS:
        text = open(os.path.join(ROOT, doc)).read()
        for m in CITE.finditer(text):
            path, lo, hi = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            if any(tag in path for tag in OURS):
                continue  # a citation of this repo's own files
            cands = _resolve(path, by_name)
            if not cands:
                bad.append(f"{doc}: {m.group(0)} -> no such file in the reference")
                continue
            n += 1
            if not any(sum(1 for _ in open(c, errors="replace")) >= max(lo, hi) for c in cands):
                bad.append(f"{doc}: {m.group(0)} -> beyond the end of {', '.join(os.path.relpath(c, REF) for c in cands)}")
    assert n > 100, "expected to find the citations"
    assert not bad, "\n".join(bad)
