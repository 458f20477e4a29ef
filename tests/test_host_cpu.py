"""Host-layer logic that needs no GPU: Go "%v" text and EqualFold as the predicate compiler sees
them (libquiverhost.so) against the oracle's restatement (oracle/gotypes.py)."""
import json
import random

import pytest


@pytest.fixture(scope="module")
def hostapi():
    from quiver_b200 import hostapi as h
    h.load()
    return h


def _typed(v):
    from oracle.gotypes import from_json
    if isinstance(v, bool) or v is None or isinstance(v, str):
        return v
    if isinstance(v, int):
        return {"int": v}
    if isinstance(v, float):
        return {"float": v}
    return from_json(v)


def test_sprint_v_matches_oracle(hostapi):
    from oracle.gotypes import sprint_v
    rnd = random.Random(5)
    vals = [0, 1, -7, 42, 1000000, 123456789012, 0.5, -2.25, 99.99, 1e6, 1234567.0, 1e21, 1e-5, 0.0001, 3e-7,
            1.7e9, 0.1 + 0.2, "abc", "", "Ünï", True, False, None]
    vals += [rnd.uniform(-1e7, 1e7) for _ in range(200)] + [rnd.random() * 10 ** rnd.randint(-12, 12) for _ in range(200)]
    for v in vals:
        text = json.dumps(v) if not isinstance(v, float) else repr(v)
        if isinstance(v, float) and "." not in text and "e" not in text:
            text += ".0"
        assert hostapi.sprint_v(text, typed=True) == sprint_v(_typed(v)), v
    # decoded JSON (every number a float64) and containers
    for doc in ['[1, 2.5, "x", true, null]', '{"b": 1, "a": {"z": [1e6, 10]}}', '[]', '{}', '[[1],[2,3]]']:
        from oracle.gotypes import from_json
        assert hostapi.sprint_v(doc, typed=False) == sprint_v(from_json(json.loads(doc))), doc


def test_equal_fold_matches_oracle(hostapi):
    from oracle.gotypes import equal_fold
    pairs = [("electronics", "ElectroNICS"), ("a", "b"), ("K", "k"), ("K", "k"), ("ſ", "S"), ("É", "é"),
             ("straße", "STRASSE"), ("Σ", "σ"), ("ς", "σ"), ("Я", "я"), ("abc", "abcd"), ("", ""), ("İ", "i")]
    for a, b in pairs:
        assert hostapi.equal_fold(a, b) == equal_fold(a, b), (a, b)


def test_host_library_reports_reference_errors_without_a_device(hostapi, capi):
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(hostapi.QuiverError) as e:
        hostapi.HybridIndex(8, "cosine")
    assert "no CPU fallback" in str(e.value)


def test_reference_arm_line_contract():
    """bench.py --impl reference: the JSON line the driver parses — same metric / unit as the native arm,
    --steps / --warmup honoured, every host core busy, e2e repeating the line's own value with zero copies."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--rows", "20000",
                        "--steps", "4", "--warmup", "3"], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout carries the JSON line and nothing else"
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"].startswith("exact k-NN QPS") and line["unit"] == "queries/s"
    assert line["steps"] == 4 and line["warmup"] == 3 and line["higher_is_better"] is True and line["n_gpus"] == 1
    cores = len(os.sched_getaffinity(0))
    assert line["cpu_baseline"]["cores"] == cores and line["cpu_baseline"]["kind"] == "port"
    assert line["cpu_baseline"]["value"] == line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0
