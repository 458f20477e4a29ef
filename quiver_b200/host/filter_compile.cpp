// filter_compile.cpp — see filter_compile.hpp.
#include "filter_compile.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace qh {

namespace {
constexpr int K_MISSING = QG_KIND_MISSING, K_NULL = QG_KIND_NULL, K_STRING = QG_KIND_STRING,
              K_NUMBER = QG_KIND_NUMBER, K_BOOL = QG_KIND_BOOL, K_OTHER = QG_KIND_OTHER;
constexpr int M(int k) { return 1 << k; }
constexpr int PRESENT = M(K_NULL) | M(K_STRING) | M(K_NUMBER) | M(K_BOOL) | M(K_OTHER);
constexpr int NON_NUMBER = M(K_NULL) | M(K_STRING) | M(K_BOOL) | M(K_OTHER);

qg_clause clause(int op, int field) {
  qg_clause c{};
  c.op = op;
  c.field = field;
  return c;
}

int text_code(const Column& col, const std::string& text) {
  auto it = std::lower_bound(col.texts.begin(), col.texts.end(), text);
  if (it == col.texts.end() || *it != text) return -1;
  return (int)(it - col.texts.begin());
}
int folded_code(const Column& col, const std::string& folded) {
  auto it = std::lower_bound(col.folded.begin(), col.folded.end(), folded);
  if (it == col.folded.end() || *it != folded) return -1;
  return (int)(it - col.folded.begin());
}

struct Builder {
  Program* p;
  int begin_pred() { return (int)p->clauses.size(); }
  void end_pred(int first, bool negate) {
    qg_pred pd{};
    pd.first_clause = first;
    pd.n_clauses = (int)p->clauses.size() - first;
    pd.negate = negate ? 1 : 0;
    pd.require_row = 1;
    p->preds.push_back(pd);
  }
  void add(const qg_clause& c) { p->clauses.push_back(c); }
};

// clauses that are true iff core.valuesEqual(row value, v)   (collection.go:601-608)
void core_equal_clauses(Builder& b, int field, const Column& col, const Value& v) {
  const std::string text = sprint_v(v);
  const int code = text_code(col, text);
  if (v.is_numeric()) {
    qg_clause a = clause(QG_OP_NUM_EQ_TOL, field);
    a.fa = v.num;
    a.fb = 1e-9;
    b.add(a);
    if (code >= 0) {  // non-numeric rows compare by "%v" text
      qg_clause t = clause(QG_OP_SCODE_EQ, field);
      t.ia = code;
      t.ib = NON_NUMBER;
      b.add(t);
    }
  } else if (code >= 0) {
    qg_clause t = clause(QG_OP_SCODE_EQ, field);
    t.ia = code;
    t.ib = PRESENT;
    b.add(t);
  }
}

// clauses that are true iff core.compareValues(row value, v) <op> 0   (collection.go:610-634)
void core_compare_clauses(Builder& b, int field, const Column& col, const Value& v, const std::string& op) {
  const int num_op = op == "<" ? 0 : (op == "<=" ? 1 : (op == ">" ? 2 : 3));
  const std::string text = sprint_v(v);
  const int lb = (int)(std::lower_bound(col.texts.begin(), col.texts.end(), text) - col.texts.begin());
  const int ub = (int)(std::upper_bound(col.texts.begin(), col.texts.end(), text) - col.texts.begin());
  qg_clause t = clause(QG_OP_SCODE_CMP, field);
  // text(row) <  text(v) <=> scode <  lb ;  <= <=> scode < ub ;  > <=> scode >= ub ;  >= <=> scode >= lb
  if (op == "<") { t.ic = 0; t.ia = lb; }
  else if (op == "<=") { t.ic = 0; t.ia = ub; }
  else if (op == ">") { t.ic = 1; t.ia = ub; }
  else { t.ic = 1; t.ia = lb; }
  if (v.is_numeric()) {
    qg_clause a = clause(QG_OP_NUM_CMP, field);
    a.ia = num_op;
    a.fa = v.num;
    b.add(a);
    t.ib = NON_NUMBER;
  } else {
    t.ib = PRESENT;
  }
  b.add(t);
}

void exists_pred(Builder& b, int field) {
  const int first = b.begin_pred();
  qg_clause c = clause(QG_OP_KIND_IN, field);
  c.ia = PRESENT;
  b.add(c);
  b.end_pred(first, false);
}

std::string hex_bits(double x) {
  if (x == 0.0) x = 0.0;  // -0 == +0 under ==
  uint64_t u;
  static_assert(sizeof(u) == sizeof(x), "double is 64 bits");
  std::memcpy(&u, &x, 8);
  char buf[17];
  std::snprintf(buf, sizeof buf, "%016llx", (unsigned long long)u);
  return buf;
}

// reflect.DeepEqual below the top level is type-sensitive: a decoded float64 never equals a Go int.
std::string nested_key(const Value& v) {
  switch (v.type) {
    case Value::Null: return "z";
    case Value::Bool: return v.b ? "b1" : "b0";
    case Value::Number: return "f" + hex_bits(v.num);
    case Value::Int: return "i" + hex_bits(v.num);
    case Value::String: return "s" + std::to_string(v.str.size()) + ":" + v.str;
    case Value::Array: {
      std::string k = "a[";
      for (const ValuePtr& e : v.arr) k += (e ? nested_key(*e) : std::string("z")) + ",";
      return k + "]";
    }
    case Value::Object: {
      std::string k = "o{";
      for (const auto& kv : v.obj)
        k += std::to_string(kv.first.size()) + ":" + kv.first + "=" + (kv.second ? nested_key(*kv.second) : std::string("z")) + ",";
      return k + "}";
    }
  }
  return "?";
}

int whole_code(const Column& col, const std::string& key) {
  auto it = std::lower_bound(col.whole_keys.begin(), col.whole_keys.end(), key);
  if (it == col.whole_keys.end() || *it != key) return -1;
  return (int)(it - col.whole_keys.begin());
}

int elem_code(const Column& col, const std::string& key) {
  auto it = std::lower_bound(col.elem_keys.begin(), col.elem_keys.end(), key);
  if (it == col.elem_keys.end() || *it != key) return -1;
  return (int)(it - col.elem_keys.begin());
}

}  // namespace

std::string element_key(const Value& v) {
  if (v.is_numeric()) return "n" + hex_bits(v.num);  // toFloat64(a) == toFloat64(b)
  return nested_key(v);
}

void encode_column(const std::vector<CellRef>& cells, Column* out) {
  const size_t n = cells.size();
  out->kind.assign(n, (uint8_t)QG_KIND_MISSING);
  out->num.assign(n, 0.0);
  out->scode.assign(n, -1);
  out->fcode.assign(n, -1);
  out->texts.clear();
  out->folded.clear();
  out->has_array_rows = false;
  out->arr_off.assign(n + 1, 0);
  out->arr_code.clear();
  out->elem_keys.clear();
  out->whole_keys.clear();
  std::vector<std::string> whole(n);    // DeepEqual key of the array / map rows
  std::vector<std::string> elems;       // element keys in row order
  std::vector<int32_t> elem_count(n, 0);
  std::vector<std::string> text(n), fold(n);
  for (size_t i = 0; i < n; ++i) {
    if (cells[i].no_row) { out->kind[i] = QG_KIND_NOROW; continue; }
    const Value* v = cells[i].value;
    if (!v) continue;  // MISSING
    text[i] = sprint_v(*v);
    switch (v->type) {
      case Value::Null: out->kind[i] = K_NULL; break;
      case Value::Bool: out->kind[i] = K_BOOL; out->num[i] = v->b ? 1.0 : 0.0; break;
      case Value::Number:
      case Value::Int: out->kind[i] = K_NUMBER; out->num[i] = v->num; break;
      case Value::String:
        out->kind[i] = (uint8_t)(K_STRING | (v->str.empty() ? 0 : 0x80));
        fold[i] = fold_key(v->str);
        break;
      case Value::Array:
        out->kind[i] = (uint8_t)(K_OTHER | (v->arr.empty() ? 0 : 0x80));
        out->has_array_rows = true;
        whole[i] = nested_key(*v);
        for (const ValuePtr& e : v->arr) {
          static const Value kNil;
          elems.push_back(element_key(e ? *e : kNil));
        }
        elem_count[i] = (int32_t)v->arr.size();
        break;
      case Value::Object:
        out->kind[i] = (uint8_t)(K_OTHER | (v->obj.empty() ? 0 : 0x80));
        out->has_array_rows = true;
        whole[i] = nested_key(*v);
        elems.push_back(element_key(*v));  // a map is compared whole (facets.go:322-328)
        elem_count[i] = 1;
        break;
    }
  }
  for (size_t i = 0; i < n; ++i) {
    const int k = out->kind[i] & 0x7f;
    if (k == K_MISSING || k == QG_KIND_NOROW) continue;
    out->texts.push_back(text[i]);
    if (k == K_STRING) out->folded.push_back(fold[i]);
    if (k == K_OTHER) out->whole_keys.push_back(whole[i]);
  }
  std::sort(out->whole_keys.begin(), out->whole_keys.end());
  out->whole_keys.erase(std::unique(out->whole_keys.begin(), out->whole_keys.end()), out->whole_keys.end());
  std::sort(out->texts.begin(), out->texts.end());
  out->texts.erase(std::unique(out->texts.begin(), out->texts.end()), out->texts.end());
  std::sort(out->folded.begin(), out->folded.end());
  out->folded.erase(std::unique(out->folded.begin(), out->folded.end()), out->folded.end());
  out->elem_keys = elems;
  std::sort(out->elem_keys.begin(), out->elem_keys.end());
  out->elem_keys.erase(std::unique(out->elem_keys.begin(), out->elem_keys.end()), out->elem_keys.end());
  out->arr_code.reserve(elems.size());
  for (const std::string& k : elems) out->arr_code.push_back(elem_code(*out, k));
  for (size_t i = 0; i < n; ++i) out->arr_off[i + 1] = out->arr_off[i] + elem_count[i];
  for (size_t i = 0; i < n; ++i) {
    const int k = out->kind[i] & 0x7f;
    if (k == K_MISSING || k == QG_KIND_NOROW) continue;
    out->scode[i] = text_code(*out, text[i]);
    if (k == K_STRING) out->fcode[i] = folded_code(*out, fold[i]);
    if (k == K_OTHER) out->fcode[i] = whole_code(*out, whole[i]);
  }
}

int compile_core_filters(const std::vector<CoreFilter>& filters, ColumnSource& cols, Program* out, std::string* err) {
  Builder b{out};
  for (const CoreFilter& f : filters) {
    const int field = cols.field_index(f.field);
    const Column& col = cols.column(f.field);
    static const Value kNull;
    const Value& v = f.value ? *f.value : kNull;
    if (f.op == "=") {
      const int first = b.begin_pred();
      core_equal_clauses(b, field, col, v);
      b.end_pred(first, false);
    } else if (f.op == "!=") {
      exists_pred(b, field);  // a missing field is false even for != (collection.go:533-536)
      const int first = b.begin_pred();
      core_equal_clauses(b, field, col, v);
      b.end_pred(first, true);
    } else if (f.op == ">" || f.op == ">=" || f.op == "<" || f.op == "<=") {
      const int first = b.begin_pred();
      core_compare_clauses(b, field, col, v, f.op);
      b.end_pred(first, false);
    } else if (f.op == "in") {
      const int first = b.begin_pred();
      if (v.type == Value::Array)  // only []interface{} operands are searched (collection.go:551-558)
        for (const ValuePtr& e : v.arr) core_equal_clauses(b, field, col, *e);
      b.end_pred(first, false);
    } else if (f.op == "not_in") {
      exists_pred(b, field);
      if (v.type == Value::Array) {
        const int first = b.begin_pred();
        for (const ValuePtr& e : v.arr) core_equal_clauses(b, field, col, *e);
        b.end_pred(first, true);
      }  // a non-list operand matches every row that has the field (collection.go:571)
    } else {
      const int first = b.begin_pred();  // unknown operator: never matches (collection.go:572-574)
      b.end_pred(first, false);
    }
  }
  (void)err;
  return 0;
}

int compile_facet_filters(const std::vector<FacetFilter>& filters, ColumnSource& cols, Program* out, std::string* err) {
  Builder b{out};
  for (const FacetFilter& f : filters) {
    const int field = cols.field_index(f.field);
    const Column& col = cols.column(f.field);
    const int first = b.begin_pred();
    bool negate = false;
    switch (f.type) {
      case FacetFilter::Equality: {
        // a nil filter value only equals a nil facet, and nil facets are never stored (facets.go:423-425)
        if (!f.value || f.value->type == Value::Null) break;
        const Value& v = *f.value;
        if (v.type == Value::String) {  // strings.EqualFold (facets.go:73-77)
          const int code = folded_code(col, fold_key(v.str));
          if (code >= 0) {
            qg_clause c = clause(QG_OP_FCODE_EQ, field);
            c.ia = code;
            b.add(c);
          }
        } else if (v.is_numeric()) {    // toFloat64(a) == toFloat64(b) (facets.go:80-82)
          qg_clause c = clause(QG_OP_NUM_EQ, field);
          c.ia = M(K_NUMBER);
          c.fa = v.num;
          b.add(c);
        } else if (v.type == Value::Bool) {  // reflect.DeepEqual(bool, bool)
          qg_clause c = clause(QG_OP_NUM_EQ, field);
          c.ia = M(K_BOOL);
          c.fa = v.b ? 1.0 : 0.0;
          b.add(c);
        } else {
          // an array / map filter value: reflect.DeepEqual against the facet's own array / map (facets.go:85)
          const int code = whole_code(col, nested_key(v));
          if (code >= 0) {
            qg_clause c = clause(QG_OP_WHOLE_EQ, field);
            c.ia = code;
            b.add(c);
          }
        }
        break;
      }
      case FacetFilter::Range: {
        const bool has_lo = f.min && f.min->type != Value::Null, has_hi = f.max && f.max->type != Value::Null;
        // a bound of a non-numeric type makes its side false (the default: arms, facets.go:167,186)
        if ((has_lo && !f.min->is_numeric()) || (has_hi && !f.max->is_numeric())) break;
        qg_clause c = clause(QG_OP_NUM_RANGE, field);
        c.ia = (has_lo ? 1 : 0) | (f.include_min ? 2 : 0) | (has_hi ? 4 : 0) | (f.include_max ? 8 : 0);
        c.fa = has_lo ? f.min->num : 0.0;
        c.fb = has_hi ? f.max->num : 0.0;
        b.add(c);
        break;
      }
      case FacetFilter::Set: {
        // array facet: any element valuesEqual to any member; map facet: the whole value
        // (facets.go:308-328) — both through the column's element dictionary
        if (col.has_array_rows) {
          const int e0 = (int)out->iset.size();
          for (const ValuePtr& v : f.values) {
            static const Value kNil;
            const int code = elem_code(col, element_key(v ? *v : kNil));
            if (code >= 0) out->iset.push_back(code);
          }
          if ((int)out->iset.size() > e0) {
            qg_clause c = clause(QG_OP_ELEM_IN, field);
            c.ia = e0;
            c.ic = (int)out->iset.size() - e0;
            b.add(c);
          }
        }
        // string facet: EqualFold against the string members (facets.go:296-307)
        const int i0 = (int)out->iset.size();
        for (const ValuePtr& v : f.values)
          if (v && v->type == Value::String) {
            const int code = folded_code(col, fold_key(v->str));
            if (code >= 0) out->iset.push_back(code);
          }
        if ((int)out->iset.size() > i0) {
          qg_clause c = clause(QG_OP_FCODE_IN, field);
          c.ia = i0;
          c.ic = (int)out->iset.size() - i0;
          b.add(c);
        }
        // numeric facet: valuesEqual => float64 equality with the numeric members (facets.go:515-520)
        const int f0 = (int)out->fset.size();
        for (const ValuePtr& v : f.values)
          if (v && v->is_numeric()) out->fset.push_back(v->num);
        if ((int)out->fset.size() > f0) {
          qg_clause c = clause(QG_OP_NUM_IN, field);
          c.ia = f0;
          c.ic = (int)out->fset.size() - f0;
          c.ib = M(K_NUMBER);
          b.add(c);
        }
        // bool facet: reflect.DeepEqual with the bool members
        const int g0 = (int)out->fset.size();
        for (const ValuePtr& v : f.values)
          if (v && v->type == Value::Bool) out->fset.push_back(v->b ? 1.0 : 0.0);
        if ((int)out->fset.size() > g0) {
          qg_clause c = clause(QG_OP_NUM_IN, field);
          c.ia = g0;
          c.ic = (int)out->fset.size() - g0;
          c.ib = M(K_BOOL);
          b.add(c);
        }
        break;
      }
      case FacetFilter::Exists: {
        // exists = non-nil and, for strings / arrays / maps, non-empty (facets.go:365-380)
        qg_clause a = clause(QG_OP_KIND_IN, field);
        a.ia = M(K_NUMBER) | M(K_BOOL);
        b.add(a);
        qg_clause c = clause(QG_OP_KIND_IN, field);
        c.ia = M(K_STRING) | M(K_OTHER);
        c.ib = 1;  // non-empty
        b.add(c);
        negate = !f.should_exist;
        break;
      }
    }
    b.end_pred(first, negate);
  }
  return 0;
}

}  // namespace qh
