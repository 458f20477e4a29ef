"""Go dynamic values and fmt "%v" text, as far as the filter semantics need them. TEST INFRASTRUCTURE ONLY.

The reference's filter predicates are written over Go `interface{}` values and depend on the
DYNAMIC TYPE (int vs float64 vs string ...), on `fmt.Sprintf("%v", x)` text
(pkg/core/collection.go:601-634) and on `reflect.DeepEqual` (pkg/facets/facets.go:86,515-520).
A typed value is encoded exactly like tests/golden/*.json encodes the Go literals:

    "text"  True/False  None            string, bool, nil
    {"int": 42}                          int (any Go integer type)
    {"float": 99.99}                     float64   (what encoding/json produces for every number)
    {"float32": 1.5}                     float32
    {"strings": ["a", "b"]}              []string
    {"list": [v, ...]}                   []interface{}
    {"map": {"k": v}}                    map[string]interface{}
"""
from __future__ import annotations

import math
import struct


def from_json(obj):
    """What `json.Unmarshal(raw, &map[string]interface{})` yields (collection.go:725): every
    number is a float64, arrays are []interface{}, objects are map[string]interface{}."""
    if obj is None or isinstance(obj, (bool, str)):
        return obj
    if isinstance(obj, (int, float)):
        return {"float": float(obj)}
    if isinstance(obj, list):
        return {"list": [from_json(x) for x in obj]}
    if isinstance(obj, dict):
        return {"map": {k: from_json(v) for k, v in obj.items()}}
    raise TypeError(type(obj))


def kind(v) -> str:
    if v is None:
        return "nil"
    if isinstance(v, bool):
        return "bool"
    if isinstance(v, str):
        return "string"
    if isinstance(v, dict) and len(v) == 1:
        return next(iter(v))
    raise TypeError(f"not a typed Go value: {v!r}")


def is_numeric(v) -> bool:
    """facets.go:462-468 isNumeric / collection.go:577-599 asFloat64 (json.Number never occurs:
    the reference decodes with plain json.Unmarshal)."""
    return kind(v) in ("int", "float", "float32")


def to_float64(v) -> float:
    k = kind(v)
    if k == "int":
        return float(v["int"])
    if k == "float":
        return float(v["float"])
    if k == "float32":
        return struct.unpack("f", struct.pack("f", v["float32"]))[0]
    return 0.0


def _shortest_digits(x: float, bits: int):
    """(digits, decimal_point_position) of the shortest repr that round-trips."""
    if bits == 32:
        # shortest decimal that round-trips through float32
        for prec in range(1, 18):
            s = "%.*e" % (prec - 1, x)
            if struct.unpack("f", struct.pack("f", float(s)))[0] == x:
                break
    else:
        s = "%.17e" % x
        for prec in range(1, 18):
            s = "%.*e" % (prec - 1, x)
            if float(s) == x:
                break
    mant, exp = s.split("e")
    digits = mant.replace(".", "").lstrip("-")
    digits = digits.rstrip("0") or "0"
    return digits, int(exp) + 1


def format_float(x: float, bits: int = 64) -> str:
    """fmt "%v" of a float = strconv.FormatFloat(x, 'g', -1, bits): shortest round-trip digits;
    the exponent form is used when exp < -4 || exp >= 6 (strconv/ftoa.go: "if precision was the
    shortest possible, use precision 6 for this decision"), so 1000000.0 prints as 1e+06."""
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "+Inf" if x > 0 else "-Inf"
    neg = x < 0 or (x == 0 and math.copysign(1.0, x) < 0)
    ax = abs(x)
    if ax == 0:
        return "-0" if neg else "0"
    digits, dp = _shortest_digits(ax, bits)
    exp = dp - 1
    if exp < -4 or exp >= 6:
        body = digits[0] + ("." + digits[1:] if len(digits) > 1 else "")
        out = f"{body}e{'+' if exp >= 0 else '-'}{abs(exp):02d}"
    elif dp <= 0:
        out = "0." + "0" * (-dp) + digits
    elif dp >= len(digits):
        out = digits + "0" * (dp - len(digits))
    else:
        out = digits[:dp] + "." + digits[dp:]
    return ("-" if neg else "") + out


def sprint_v(v) -> str:
    """fmt.Sprintf("%v", v)."""
    k = kind(v)
    if k == "nil":
        return "<nil>"
    if k == "bool":
        return "true" if v else "false"
    if k == "string":
        return v
    if k == "int":
        return str(int(v["int"]))
    if k == "float":
        return format_float(float(v["float"]), 64)
    if k == "float32":
        return format_float(to_float64(v), 32)
    if k == "strings":
        return "[" + " ".join(v["strings"]) + "]"
    if k == "list":
        return "[" + " ".join(sprint_v(x) for x in v["list"]) + "]"
    if k == "map":
        items = sorted(v["map"].items())  # fmt prints maps in key order
        return "map[" + " ".join(f"{kk}:{sprint_v(vv)}" for kk, vv in items) + "]"
    raise TypeError(k)


def deep_equal(a, b) -> bool:
    """reflect.DeepEqual on the value shapes above: identical dynamic type and equal contents."""
    ka, kb = kind(a), kind(b)
    if ka != kb:
        return False
    if ka in ("nil",):
        return True
    if ka in ("bool", "string"):
        return a == b
    if ka == "int":
        return int(a["int"]) == int(b["int"])  # (same Go integer type assumed)
    if ka in ("float", "float32"):
        return to_float64(a) == to_float64(b)
    if ka == "strings":
        return list(a["strings"]) == list(b["strings"])
    if ka == "list":
        return len(a["list"]) == len(b["list"]) and all(deep_equal(x, y) for x, y in zip(a["list"], b["list"]))
    if ka == "map":
        return a["map"].keys() == b["map"].keys() and all(deep_equal(a["map"][k], b["map"][k]) for k in a["map"])
    raise TypeError(ka)


def equal_fold(a: str, b: str) -> bool:
    """strings.EqualFold: Unicode simple case folding. Python's casefold() is the FULL folding
    (e.g. 'ß' -> 'ss'), so compare rune by rune with the simple mappings instead."""
    if len(a) != len(b):
        return False
    for x, y in zip(a, b):
        if x == y:
            continue
        if _simple_fold(x) != _simple_fold(y):
            return False
    return True


def _simple_fold(ch: str) -> str:
    lo = ch.lower()
    if len(lo) != 1:
        lo = ch
    up = lo.upper()
    if len(up) == 1 and len(up.lower()) == 1:
        lo = up.lower()
    # the two ASCII orbits with a third member: K (KELVIN SIGN) and s (LONG S)
    if lo == "K":
        return "k"
    if lo == "ſ":
        return "s"
    return lo
