"""Every `path:line[-line]` citation of the THUNDER tree in the headers, sources, oracle and documents points at an existing file
and at lines inside it (CPU; skipped where /root/reference is absent)."""
import os
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("THB_REFERENCE", "/root/reference"))
PAT = re.compile(r"((?:src|include|gpu|external|script|appsrc)/[\w/\.\-]+\.(?:cpp|h|cu|cuh|in|json)):(\d+)(?:-(\d+))?")


@pytest.mark.skipif(not (REF / "src").is_dir(), reason="reference tree not present")
def test_reference_citations_resolve():
    files = [ROOT / "DESIGN.md", ROOT / "INTEGRATION.md", ROOT / "README.md", ROOT / "bench.py", ROOT / "__graft_entry__.py"]
    for d, pats in (("include", ("*.h",)), ("thunder_b200", ("**/*.cuh", "**/*.cu", "**/*.cpp", "**/*.h", "**/*.py")),
                    ("oracle", ("*.py", "*.c", "*.cpp", "*.sh")), ("tests", ("*.py", "golden/*.py", "pf_host/*.cpp", "host_bind/*.cpp"))):
        for p in pats:
            files += [f for f in (ROOT / d).glob(p) if "_ref" not in f.parts]
    lines_of, n, bad = {}, 0, []
    for f in files:
        for m in PAT.finditer(f.read_text(errors="ignore")):
            path, a, b = REF / m.group(1), int(m.group(2)), m.group(3)
            n += 1
            if not path.exists():
                bad.append((f.name, m.group(0), "no such file"))
                continue
            if path not in lines_of:
                lines_of[path] = sum(1 for _ in open(path, errors="ignore"))
            hi = int(b) if b else a
            if a < 1 or hi < a or hi > lines_of[path]:
                bad.append((f.name, m.group(0), f"{lines_of[path]} lines"))
    assert n >= 150, n
    assert not bad, bad[:20]
