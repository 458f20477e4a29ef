// value.hpp — JSON values as Go's encoding/json produces them, and Go's fmt "%v" text.
//
// The reference's metadata predicates are written over `map[string]interface{}` decoded by plain
// json.Unmarshal (pkg/core/collection.go:725): every number is a float64, arrays are
// []interface{}, objects are map[string]interface{}. core.valuesEqual / compareValues
// (collection.go:601-634) fall back to fmt.Sprintf("%v", x) text, so that text has to be
// reproduced exactly (1000000.0 prints as 1e+06).
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace qh {

struct Value;
using ValuePtr = std::shared_ptr<Value>;

struct Value {
  enum Type { Null, Bool, Number, Int, String, Array, Object } type = Null;
  bool b = false;
  double num = 0.0;      // Number (float64) or Int (Go integer literal in a filter value)
  std::string str;
  std::vector<ValuePtr> arr;
  std::map<std::string, ValuePtr> obj;  // Go prints maps in key order

  bool is_numeric() const { return type == Number || type == Int; }
};

// Parses one JSON document. Numbers become Number (float64), like json.Unmarshal into
// interface{}. With `typed_literals` a number written without '.', 'e' or 'E' becomes Int: that is
// how a Go filter operand such as `Value: 42` (int) is distinguished from `42.0` (float64), which
// matters for its "%v" text (1000000 vs 1e+06). Returns nullptr and sets *err on a syntax error.
ValuePtr parse_json(const std::string& text, bool typed_literals, std::string* err);

// fmt.Sprintf("%v", v)
std::string sprint_v(const Value& v);
// strconv.FormatFloat(x, 'g', -1, 64) as used by %v
std::string format_float_v(double x);
// strings.EqualFold folding key: equal keys <=> EqualFold(a, b). UTF-8 in, UTF-8 out.
std::string fold_key(const std::string& s);
// reflect.DeepEqual for decoded JSON values / typed literals of the same shapes
bool deep_equal(const Value& a, const Value& b);

}  // namespace qh
