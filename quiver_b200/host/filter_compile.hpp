// filter_compile.hpp — lowers the reference's two predicate languages to the device normal form
// (include/quiver_gpu.h: qg_pred / qg_clause) and encodes metadata / facet values as columns.
//
//   core filters   pkg/core/collection.go:532-575 matchesFilter  (= != > >= < <= in not_in)
//   facet filters  pkg/facets/facets.go:63-86 Equality, :126-239 Range, :289-329 Set, :365-380 Exists,
//                  :432-459 MatchesAllFilters
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../../include/quiver_gpu.h"
#include "value.hpp"

namespace qh {

// One encoded column (a metadata key or a facet dot-path) over all rows.
struct Column {
  std::vector<uint8_t> kind;    // qg_value_kind (| 0x80 when a string / array / map is non-empty)
  std::vector<double> num;
  std::vector<int32_t> scode;   // rank of the value's "%v" text in `texts` (sorted, unique); -1 when absent
  std::vector<int32_t> fcode;   // strings: id of the case-folded string in `folded` (sorted, unique);
                                // arrays / maps: id of the whole value in `whole_keys`; -1 otherwise
  std::vector<std::string> texts;
  std::vector<std::string> folded;
  bool has_array_rows = false;  // a row holds an array / map value
  // Elements of the array-valued rows (CSR over all rows; a map counts as its own single element:
  // facets.go:322-328 compares it whole). elem_keys is the sorted, unique dictionary of element_key()
  // strings; arr_code holds indices into it.
  std::vector<int32_t> arr_off, arr_code;
  std::vector<std::string> elem_keys;
  // reflect.DeepEqual keys of the array / map values themselves (EqualityFilter on such a facet, facets.go:85)
  std::vector<std::string> whole_keys;
};

// Dictionary key of one value under facets.valuesEqual (facets.go:515-520): equal keys <=> valuesEqual.
// Numbers compare as float64 whatever their Go type, everything else by reflect.DeepEqual.
std::string element_key(const Value& v);

// value == nullptr: the field is absent (MISSING); no_row: the row has no metadata / facet entry.
struct CellRef {
  const Value* value;
  bool no_row;
};
void encode_column(const std::vector<CellRef>& cells, Column* out);

struct CoreFilter {      // types.Filter{Field, Operator, Value} (pkg/types/search.go:45-52)
  std::string field;
  std::string op;        // "=", "!=", ">", ">=", "<", "<=", "in", "not_in"; anything else never matches
  ValuePtr value;        // typed literal (Value::Int for Go integers)
};

struct FacetFilter {     // facets.Filter implementations
  enum Type { Equality, Range, Set, Exists } type = Equality;
  std::string field;
  ValuePtr value;        // Equality
  ValuePtr min, max;     // Range (nullptr / Null = open)
  bool include_min = true, include_max = true;
  std::vector<ValuePtr> values;  // Set
  bool should_exist = true;      // Exists
};

struct Program {
  std::vector<qg_pred> preds;
  std::vector<qg_clause> clauses;
  std::vector<int32_t> iset;
  std::vector<double> fset;
};

// `field_index(name)` returns the device column index of a field; `column(name)` its encoding.
struct ColumnSource {
  virtual ~ColumnSource() {}
  virtual int field_index(const std::string& name) = 0;
  virtual const Column& column(const std::string& name) = 0;
};

// Returns 0 (every predicate shape of the reference has a device form).
int compile_core_filters(const std::vector<CoreFilter>& filters, ColumnSource& cols, Program* out, std::string* err);
int compile_facet_filters(const std::vector<FacetFilter>& filters, ColumnSource& cols, Program* out, std::string* err);

}  // namespace qh
