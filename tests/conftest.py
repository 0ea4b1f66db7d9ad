import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FIX = os.path.join(ROOT, "tests", "fixtures")
GOLDEN = os.path.join(ROOT, "tests", "golden", "sql_golden.json")
REF_ROOT = os.environ.get("CQ_REF", "/root/reference")
# the reference's shipped fixtures the "refdata" golden cases read (data/users.csv is BASELINE configs[0]):
# copied under tests/golden/ so that those cases also run where /root/reference does not exist (the GPU box)
REFDATA_ROOT = os.path.join(ROOT, "tests", "golden", "refdata")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden():
    with open(GOLDEN) as f:
        return json.load(f)


def parse_dump(text):
    """ref_dump / oracle_dump / cq_gpu_dump output -> (header, rows) with typed cells."""
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    if not lines or lines[0].startswith("#error"):
        return {"error": lines[0] if lines else "#empty"}
    head = lines[0].split()
    nrows, ncols = int(head[1]), int(head[3])
    cols = [ln[5:] for ln in lines[1:1 + ncols]]
    body = "\n".join(lines[1 + ncols:])
    rows = []
    # cells are separated by \t and rows by \n, but string cells carry their byte length
    pos = 0
    for _ in range(nrows):
        row = []
        while True:
            tag = body[pos]
            if tag == "N":
                row.append(("N",))
                pos += 1
            elif tag == "S":
                j = body.index(":", pos + 2)
                n = int(body[pos + 2:j])
                row.append(("S", body[j + 1:j + 1 + n]))
                pos = j + 1 + n
            else:
                j = pos
                while j < len(body) and body[j] not in "\t\n":
                    j += 1
                cell = body[pos:j]
                if tag == "I":
                    row.append(("I", int(cell[2:])))
                elif tag == "D":
                    row.append(("D", float.fromhex(cell[2:])))
                elif tag == "T":
                    row.append(("T", cell[2:]))
                else:
                    row.append(("?", cell))
                pos = j
            if pos < len(body) and body[pos] == "\t":
                pos += 1
                continue
            if pos < len(body) and body[pos] == "\n":
                pos += 1
            break
        rows.append(row)
    return {"cols": cols, "rows": rows}


def cells_equal(a, b, rel=0.0):
    if a[0] != b[0]:
        return False
    if a[0] == "D":
        x, y = a[1], b[1]
        if x == y or (x != x and y != y):
            return True
        return rel > 0 and abs(x - y) <= rel * max(abs(x), abs(y))
    return a == b


def dumps_equal(got, want, rel=0.0):
    if "error" in want or "error" in got:
        return got == want, "error mismatch"
    if got["cols"] != want["cols"]:
        return False, f"columns {got['cols']} != {want['cols']}"
    if len(got["rows"]) != len(want["rows"]):
        return False, f"row count {len(got['rows'])} != {len(want['rows'])}"
    for i, (r1, r2) in enumerate(zip(got["rows"], want["rows"])):
        if len(r1) != len(r2):
            return False, f"row {i} width"
        for c, (x, y) in enumerate(zip(r1, r2)):
            if not cells_equal(x, y, rel):
                return False, f"row {i} col {c}: {x} != {y}"
    return True, ""


def run_dump(binary, case, env=None):
    cwd = FIX if case["kind"] == "fix" else REFDATA_ROOT
    e = dict(os.environ)
    if env:
        e.update(env)
    p = subprocess.run([binary] + case["args"] + [case["sql"]], cwd=cwd, capture_output=True, timeout=300, env=e)
    return p.returncode, p.stdout.decode("latin1"), p.stderr.decode("latin1")


@pytest.fixture(scope="session")
def golden():
    return load_golden()
