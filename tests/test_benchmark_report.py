"""scripts/reference_benchmarks.py writes the table the reference's exporter writes (tests/utils/Performance.hpp:154-214)
under the file name of tests/main.cpp:39-43, parseable the way scripts/compare_tests_parse.py parses it."""
import os
import re
import subprocess
import sys

import pytest

from tests import harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse_like_reference(path):
    """Restatement of the row parser of the reference's scripts/compare_tests_parse.py:31-56."""
    rows = []
    for line in open(path, encoding="utf-8"):
        line = line.rstrip("\r\n")
        if re.match(r"^[ \t\-|]+$", line) or "|" not in line:
            continue
        parts = line.split("|")
        while parts and parts[-1].strip() == "":
            parts.pop()
        if len(parts) < 6:
            continue
        maxv, minv, avg, thr, status = (p.strip() for p in (parts[-1], parts[-2], parts[-3], parts[-4], parts[-5]))
        name = " | ".join(p.strip() for p in parts[:-5] if p.strip())
        if re.search(r"^[Bb]enchmark[ ]+[Nn]ame$", name) or not name:
            continue
        rows.append((name, status, float(thr), float(avg), float(minv), float(maxv)))
    return rows


@pytest.mark.skipif(not harness.available("oracle"), reason="oracle library not built")
def test_report_format(tmp_path):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "reference_benchmarks.py"), "--backend", "oracle",
                          "--samples", "2", "--filter", "1k", "--out", str(tmp_path)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    files = os.listdir(tmp_path)
    assert len(files) == 1 and re.fullmatch(r"benchmarks--\d{4}-\d{2}-\d{2}--\d{2}-\d{2}-\d{2}\.md", files[0])
    lines = open(tmp_path / files[0]).read().splitlines()
    assert lines[0].startswith("Benchmark Name") and lines[0].rstrip().endswith(
        "| Status | Threshold (ms) | Avg (ms) | Min (ms) | Max (ms) |")
    assert set(lines[1]) <= set("-| ")
    rows = parse_like_reference(tmp_path / files[0])
    assert {r[0] for r in rows} == {"Registry | Find all colliding pairs among 1k entities",
                                   "Registry | Find all colliding pairs among 1k bullets",
                                   "Registry | Find collisions 1-by-1 among 1k entities"}
    assert [r[3] for r in rows] == sorted(r[3] for r in rows)  # exportBenchmarks sorts by average time
    for name, status, thr, avg, lo, hi in rows:
        assert status == ("PASS" if avg <= thr else "FAIL") and lo <= avg <= hi
