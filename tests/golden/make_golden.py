#!/usr/bin/env python3
"""Extract the reference's own known-answer vectors for the exact-search hot path.

Reads the Go test sources under /root/reference (read-only; only available in the build
container) and writes tests/golden/*.json. The Go literals are PARSED, not retyped by hand,
so the fixtures are exactly what the reference's tests assert. Re-run:

    python tests/golden/make_golden.py [/root/reference]

The JSON keeps Go's dynamic types:  {"int": 42}  {"float": 99.99}  {"float32": 1.414214}
{"strings": [...]} ([]string)  {"list": [...]} ([]interface{})  {"map": {...}}  {"ident": "Equals"}
{"call": "NewEqualityFilter", "args": [...]}; strings, bools and nil map to JSON directly.
"""
import json
import math
import os
import re
import struct
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

TOKEN = re.compile(r"""
    (?P<ws>\s+|//[^\n]*)
  | (?P<str>"(?:\\.|[^"\\])*")
  | (?P<raw>`[^`]*`)
  | (?P<num>-?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?)
  | (?P<id>[A-Za-z_][A-Za-z_0-9]*(?:\.[A-Za-z_][A-Za-z_0-9]*)*)
  | (?P<p>\[\]|[{}()\[\],:*])
""", re.X)


def tokenize(src):
    pos, out = 0, []
    while pos < len(src):
        m = TOKEN.match(src, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize at {src[pos:pos+40]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind != "ws":
            out.append((kind, m.group(kind)))
    return out


def f32(x):
    return struct.unpack("f", struct.pack("f", float(x)))[0]


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "")

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def expect(self, val):
        tok = self.next()
        if tok[1] != val:
            raise SyntaxError(f"expected {val!r}, got {tok!r} near token {self.i}")

    def parse_type_prefix(self):
        """Consumes a composite-literal type if one starts here; returns its name or None."""
        k, v = self.peek()
        if v == "[]":
            self.next()
            k2, v2 = self.next()
            if v2 == "interface":
                self.expect("{"); self.expect("}")
                return "[]interface{}"
            return "[]" + v2
        if k == "id" and v == "map":
            self.next(); self.expect("["); kt = self.next()[1]; self.expect("]")
            vt = self.next()[1]
            if vt == "interface":
                self.expect("{"); self.expect("}")
                vt = "interface{}"
            return f"map[{kt}]{vt}"
        return None

    def parse_value(self, elem_type=None):
        k, v = self.peek()
        if v == "{":  # untyped composite (element of a typed slice)
            return self.parse_body(elem_type)
        typ = self.parse_type_prefix()
        if typ is not None:
            return self.parse_body(typ)
        k, v = self.next()
        if k == "str":
            return json.loads(v)
        if k == "raw":
            return v[1:-1]
        if k == "num":
            if re.fullmatch(r"-?\d+", v):
                return {"int": int(v)}
            return {"float": float(v)}
        if k == "id":
            if v == "true":
                return True
            if v == "false":
                return False
            if v == "nil":
                return None
            if self.peek()[1] == "(":
                self.next()
                args = []
                while self.peek()[1] != ")":
                    args.append(self.parse_value())
                    if self.peek()[1] == ",":
                        self.next()
                self.expect(")")
                return self.call(v, args)
            if self.peek()[1] == "{":  # named composite literal: F32{...}, Filter{...}, FacetValue{...}
                return self.parse_body(v)
            if v == "math.MaxInt64":
                return {"int": 2 ** 63 - 1}
            return {"ident": v}
        raise SyntaxError(f"unexpected token {k} {v!r}")

    @staticmethod
    def num(x):
        if isinstance(x, dict):
            for key in ("int", "float", "float32"):
                if key in x:
                    return x[key]
        raise SyntaxError(f"not a number: {x!r}")

    def call(self, name, args):
        if name == "math.Sqrt":
            return {"float": math.sqrt(self.num(args[0]))}
        if name == "float32":
            return {"float32": f32(self.num(args[0]))}
        if name == "float64":
            return {"float": float(self.num(args[0]))}
        if name in ("int", "int32", "int64"):
            return {"int": int(self.num(args[0]))}
        return {"call": name, "args": args}

    def parse_body(self, typ):
        self.expect("{")
        items, keyed = [], []
        while self.peek()[1] != "}":
            # key: value ?
            if self.peek(1)[1] == ":" and self.peek()[0] in ("id", "str"):
                kk, kv = self.next()
                self.next()
                key = json.loads(kv) if kk == "str" else kv
                keyed.append((key, self.parse_value(self.elem_of(typ))))
            else:
                items.append(self.parse_value(self.elem_of(typ)))
            if self.peek()[1] == ",":
                self.next()
        self.expect("}")
        if typ in ("F32", "vectortypes.F32", "[]float32"):
            return {"f32vec": [f32(self.num(x)) for x in items]}
        if typ == "[]string":
            return {"strings": items}
        if typ == "[]interface{}":
            return {"list": items}
        if typ and typ.startswith("map["):
            return {"map": dict(keyed)}
        if keyed:
            return {"struct": typ, "fields": dict(keyed)}
        return {"struct": typ, "items": items}

    @staticmethod
    def elem_of(typ):
        if typ and typ.startswith("[]") and typ not in ("[]float32", "[]string", "[]interface{}"):
            return typ[2:]
        return None


def read(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


def balanced(src, start):
    """The text of the brace-balanced block starting at src[start] == '{' (string-aware)."""
    assert src[start] == "{"
    depth, i, in_str = 0, start, None
    while True:
        c = src[i]
        if in_str:
            if c == "\\" and in_str == '"':
                i += 1
            elif c == in_str:
                in_str = None
        elif c in '"`':
            in_str = c
        elif c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return src[start:i + 1]
        i += 1


def struct_table(src, func, var="tests"):
    """Parses `<var> := []struct { <fields> }{ <rows> }` inside function `func`."""
    start = src.index(f"func {func}(")
    m = re.compile(rf"{var}\s*:=\s*\[\]struct\s*\{{").search(src, start)
    body_start = m.end()
    depth, i = 1, body_start
    while depth:
        depth += {"{": 1, "}": -1}.get(src[i], 0)
        i += 1
    fields = []
    for line in src[body_start:i - 1].splitlines():
        line = line.split("//")[0].strip()
        if line:
            fields.append(line.split()[0])
    p = Parser(tokenize(balanced(src, src.index("{", i))))
    rows = p.parse_body("[]row")["items"]
    out = []
    for r in rows:
        if "fields" in r:
            out.append(r["fields"])
        else:
            out.append(dict(zip(fields, r["items"])))
    return out


def map_literal(src, func, var, typ_re):
    start = src.index(f"func {func}(")
    m = re.compile(rf"{var}\s*:=\s*{typ_re}").search(src, start)
    p = Parser(tokenize(balanced(src, m.end() - 1)))
    return p.parse_body("map[string]F32")["map"]


def main():
    cite = {}
    # ---- distances (tolerance 1e-6 in the reference) --------------------------------------------
    src = read("pkg/vectortypes/distances_test.go")
    dist = {}
    for func, metric in [("TestCosineDistance", "cosine"), ("TestEuclideanDistance", "euclidean"),
                         ("TestSquaredEuclideanDistance", "squared_euclidean"),
                         ("TestDotProductDistance", "dot_product"), ("TestManhattanDistance", "manhattan")]:
        dist[metric] = struct_table(src, func)
    cite["distances"] = "pkg/vectortypes/distances_test.go:9-204 (floatEquals tolerance 1e-6)"
    types_rows = struct_table(read("pkg/vectortypes/types_test.go"), "TestGetDistanceFuncByType")
    cite["distance_types"] = "pkg/vectortypes/types_test.go:9-80"
    hnsw_rows = struct_table(read("pkg/hnsw/hnsw_test.go"), "TestDistanceFunctions")
    cite["hnsw_distances"] = "pkg/hnsw/hnsw_test.go:376-455 (float32 variants, epsilon per row)"
    json.dump({"cite": cite, "vectortypes": dist, "types": types_rows, "hnsw_f32": hnsw_rows},
              open(os.path.join(OUT, "distances.json"), "w"), indent=1)

    # ---- exact search ---------------------------------------------------------------------------------
    src = read("pkg/hybrid/exact_test.go")
    vecs = map_literal(src, "TestExactIndex_Search", "testVectors", r"map\[string\]vectortypes\.F32\{")
    rows = struct_table(src, "TestExactIndex_Search")
    src2 = read("pkg/hybrid/hybrid_index_test.go")
    neg_vecs = map_literal(src2, "TestHybridIndex_SearchWithNegativeExample", "vectors",
                           r"map\[string\]vectortypes\.F32\{")
    json.dump({"cite": {"exact_search": "pkg/hybrid/exact_test.go:97-205 (cosine)",
                        "negative_example": "pkg/hybrid/hybrid_index_test.go:541-657 (cosine, query (.8,.2,0,0), "
                                            "negative = animal_fish, weight 0.5, k=3: 3 results, fish not first)",
                        "rerank_stability": "pkg/hybrid/hybrid_index_rerank_test.go:9-47 (euclidean; three copies "
                                            "of (1,0); q=(0,0); neg=(-1,0); w=0.5 => three equal distances)"},
               "exact_search": {"vectors": vecs, "cases": rows},
               "negative_example": {"vectors": neg_vecs, "query": [0.8, 0.2, 0.0, 0.0], "negative_id": "animal_fish",
                                    "weight": 0.5, "k": 3},
               "rerank_stability": {"vectors": {"1": [1.0, 0.0], "2": [1.0, 0.0], "3": [1.0, 0.0]},
                                    "query": [0.0, 0.0], "negative": [-1.0, 0.0], "weight": 0.5, "k": 3}},
              open(os.path.join(OUT, "exact_search.json"), "w"), indent=1)

    # ---- filters ------------------------------------------------------------------------------------
    src = read("pkg/facets/facets_test.go")
    facets = {name: struct_table(src, func) for name, func in
              [("equality", "TestEqualityFilter"), ("range", "TestRangeFilter"), ("set", "TestSetFilter"),
               ("exists", "TestExistsFilter"), ("matches_all", "TestMatchesAllFilters")]}
    core = struct_table(read("pkg/core/collection_test.go"), "TestMatchesFilter", var="testCases")
    json.dump({"cite": {"facets": "pkg/facets/facets_test.go:10-196,322-405; facets_numeric_test.go:7-23",
                        "matches_all_facets": "the facet list of TestMatchesAllFilters, facets_test.go:323-328",
                        "core": "pkg/core/collection_test.go:727-816 (TestMatchesFilter)"},
               "facets": facets,
               "matches_all_facets": {"category": "electronics", "price": {"float": 299.99}, "active": True,
                                      "tags": {"strings": ["smartphone", "android", "5G"]}},
               "core_matches_filter": core},
              open(os.path.join(OUT, "filters.json"), "w"), indent=1)
    print("wrote distances.json, exact_search.json, filters.json in", OUT)


if __name__ == "__main__":
    main()
