"""CPU restatement of the reference's two filter languages. TEST INFRASTRUCTURE ONLY.

  * facet filters      pkg/facets/facets.go:63-86 (Equality), :126-239 (Range), :289-329 (Set),
                       :365-380 (Exists), :397-429 (ExtractFacets), :432-459 (MatchesAllFilters)
  * metadata filters   pkg/core/collection.go:532-575 (matchesFilter), :577-634 (asFloat64,
                       valuesEqual, compareValues), and the scan-until-k loops of
                       Collection.Search (:701-752) and SearchWithFacets (:1177-1204)

Values use the typed encoding of oracle/gotypes.py (the same one tests/golden/*.json uses).
Pinned against tests/golden/filters.json (the reference's own truth tables).
"""
from __future__ import annotations

import json
import math

from .gotypes import (deep_equal, equal_fold, from_json, is_numeric, kind, sprint_v, to_float64)

# ---- facets.Filter ---------------------------------------------------------------------------


def facet_values_equal(a, b) -> bool:
    """facets.go:515-520 valuesEqual."""
    if is_numeric(a) and is_numeric(b):
        return to_float64(a) == to_float64(b)
    return deep_equal(a, b)


class EqualityFilter:
    """facets.go:39-86."""
    type = "equality"

    def __init__(self, field, value):
        self.field, self.value = field, value

    def match(self, value) -> bool:
        if self.value is None and value is None:
            return True
        if self.value is None or value is None:
            return False
        if kind(value) == "string" and kind(self.value) == "string":
            return equal_fold(value, self.value)
        if is_numeric(self.value) and is_numeric(value):
            return to_float64(self.value) == to_float64(value)
        return deep_equal(self.value, value)


class RangeFilter:
    """facets.go:93-239. Only numeric facet values can match; a bound of a non-numeric type makes
    its side false (the `default:` arms)."""
    type = "range"

    def __init__(self, field, min_, max_, include_min, include_max):
        self.field, self.min, self.max = field, min_, max_
        self.include_min, self.include_max = include_min, include_max

    def match(self, value) -> bool:
        if value is None or not is_numeric(value):
            return False
        # compareInt and compareFloat (facets.go:150-239) agree once both sides are widened; an int
        # value against int bounds is compared as integers, everything else as float64.
        both_int = kind(value) == "int"

        def side(bound, incl, lower):
            if bound is None:
                return True
            if not is_numeric(bound):
                return False
            if both_int and kind(bound) == "int":
                v, b = int(value["int"]), int(bound["int"])
            else:
                v, b = to_float64(value), to_float64(bound)
            if lower:
                return v >= b if incl else v > b
            return v <= b if incl else v < b

        return side(self.min, self.include_min, True) and side(self.max, self.include_max, False)


class SetFilter:
    """facets.go:267-329."""
    type = "set"

    def __init__(self, field, values):
        self.field = field
        self.values = list(values)

    def match(self, value) -> bool:
        if value is None:
            return False
        if kind(value) == "string":
            for v in self.values:
                if kind(v) == "string":
                    if equal_fold(value, v):
                        return True
                elif facet_values_equal(value, v):
                    return True
            return False
        if kind(value) in ("strings", "list"):
            items = value["strings"] if kind(value) == "strings" else value["list"]
            return any(facet_values_equal(item, v) for item in items for v in self.values)
        return any(facet_values_equal(value, v) for v in self.values)


class ExistsFilter:
    """facets.go:344-380."""
    type = "exists"

    def __init__(self, field, should_exist):
        self.field, self.should_exist = field, should_exist

    def match(self, value) -> bool:
        exists = value is not None
        if exists:
            k = kind(value)
            if k == "string":
                exists = len(value) > 0
            elif k in ("strings", "list"):
                exists = len(value[k]) > 0
            elif k == "map":
                exists = len(value["map"]) > 0
        return exists == self.should_exist


def extract_facets(metadata, facet_fields):
    """facets.go:397-429: dot-path lookup; nil values are dropped. `metadata` is a typed map."""
    if not facet_fields or metadata is None:
        return None
    out = []
    for field in facet_fields:
        value = metadata
        for part in field.split("."):
            if value is not None and kind(value) == "map" and part in value["map"]:
                value = value["map"][part]
            else:
                value = None
                break
        if value is not None:
            out.append((field, value))
    return out


def matches_all_filters(facets, filters) -> bool:
    """facets.go:432-459. `facets` is a list of (field, typed value)."""
    if not filters:
        return True
    if not facets:
        return False
    fmap = {}
    for f, v in facets:
        fmap[f] = v
    for flt in filters:
        exists = flt.field in fmap
        if not exists and flt.type != "exists":
            return False
        if not flt.match(fmap.get(flt.field)):
            return False
    return True


# ---- core.Filter (metadata) -------------------------------------------------------------------

OPERATORS = {"Equals": "=", "NotEquals": "!=", "GreaterThan": ">", "GreaterThanOrEqual": ">=",
             "LessThan": "<", "LessThanOrEqual": "<=", "In": "in", "NotIn": "not_in"}


def core_values_equal(a, b) -> bool:
    """collection.go:601-608: numeric pairs within 1e-9, everything else by "%v" text."""
    if is_numeric(a) and is_numeric(b):
        return abs(to_float64(a) - to_float64(b)) <= 1e-9
    return sprint_v(a) == sprint_v(b)


def core_compare_values(a, b) -> int:
    """collection.go:610-634."""
    if is_numeric(a) and is_numeric(b):
        af, bf = to_float64(a), to_float64(b)
        return -1 if af < bf else (1 if af > bf else 0)
    sa, sb = sprint_v(a).encode("utf-8"), sprint_v(b).encode("utf-8")  # Go compares bytes
    return -1 if sa < sb else (1 if sa > sb else 0)


def matches_filter(metadata: dict, field: str, op: str, value) -> bool:
    """collection.go:532-575. `metadata` maps field -> typed value; `op` is the operator text
    ("=", "!=", ">", ">=", "<", "<=", "in", "not_in"); anything else is false."""
    if field not in metadata:
        return False
    v = metadata[field]
    if op == "=":
        return core_values_equal(v, value)
    if op == "!=":
        return not core_values_equal(v, value)
    if op == ">":
        return core_compare_values(v, value) > 0
    if op == ">=":
        return core_compare_values(v, value) >= 0
    if op == "<":
        return core_compare_values(v, value) < 0
    if op == "<=":
        return core_compare_values(v, value) <= 0
    if op == "in":
        if value is not None and kind(value) == "list":
            return any(core_values_equal(v, x) for x in value["list"])
        return False
    if op == "not_in":
        if value is not None and kind(value) == "list":
            return not any(core_values_equal(v, x) for x in value["list"])
        return True
    return False


def metadata_mask(metadata_json, filters) -> list:
    """Per-row pass bit of Collection.Search's filter loop (collection.go:716-744): rows with no
    metadata, empty metadata or unparsable JSON never match; all filters are ANDed.
    `metadata_json[i]` is the row's raw JSON text (or None); filters are (field, op, typed value)."""
    out = []
    for raw in metadata_json:
        ok = False
        if raw:
            try:
                obj = json.loads(raw)
            except Exception:
                obj = None
            if isinstance(obj, dict):
                md = from_json(obj)["map"]
                ok = all(matches_filter(md, f, op, val) for f, op, val in filters)
            elif obj is None and raw.strip() == "null":
                # json.Unmarshal("null", &map) succeeds with a nil map: every lookup misses
                ok = len(filters) == 0
        out.append(ok)
    return out


def facet_mask(metadata_json, facet_fields, filters) -> list:
    """Per-row pass bit of SearchWithFacets (collection.go:1189-1204): the row must have a
    vectorFacets entry (set at Add time when the collection has facet fields and the row has
    metadata that parses, collection.go:160-176) and satisfy MatchesAllFilters."""
    out = []
    for raw in metadata_json:
        ok = False
        if raw and facet_fields:
            try:
                obj = json.loads(raw)
            except Exception:
                obj = None
            if isinstance(obj, dict):
                facets = extract_facets(from_json(obj), facet_fields)
                ok = matches_all_filters(facets, filters)
        out.append(ok)
    return out
