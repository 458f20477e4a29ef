// value.cpp — see value.hpp.
#include "value.hpp"

#include <charconv>
#include <cmath>
#include <cstring>

namespace qh {

namespace {

struct Parser {
  const std::string& s;
  size_t i = 0;
  bool typed;
  std::string err;
  explicit Parser(const std::string& t, bool ty) : s(t), typed(ty) {}

  void ws() {
    while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) ++i;
  }
  bool fail(const char* m) {
    if (err.empty()) err = std::string(m) + " at offset " + std::to_string(i);
    return false;
  }
  static void put_utf8(std::string& out, uint32_t cp) {
    if (cp < 0x80) out.push_back((char)cp);
    else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
    else if (cp < 0x10000) {
      out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
      out.push_back((char)(0x80 | (cp & 0x3F)));
    } else {
      out.push_back((char)(0xF0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
      out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F)));
    }
  }
  bool hex4(uint32_t* out) {
    if (i + 4 > s.size()) return fail("short \\u escape");
    uint32_t v = 0;
    for (int k = 0; k < 4; ++k) {
      char c = s[i++];
      v <<= 4;
      if (c >= '0' && c <= '9') v |= (uint32_t)(c - '0');
      else if (c >= 'a' && c <= 'f') v |= (uint32_t)(c - 'a' + 10);
      else if (c >= 'A' && c <= 'F') v |= (uint32_t)(c - 'A' + 10);
      else return fail("bad \\u escape");
    }
    *out = v;
    return true;
  }
  bool str(std::string* out) {
    if (i >= s.size() || s[i] != '"') return fail("expected string");
    ++i;
    while (i < s.size()) {
      unsigned char c = (unsigned char)s[i++];
      if (c == '"') return true;
      if (c < 0x20) return fail("control character in string");
      if (c != '\\') { out->push_back((char)c); continue; }
      if (i >= s.size()) break;
      char e = s[i++];
      switch (e) {
        case '"': out->push_back('"'); break;
        case '\\': out->push_back('\\'); break;
        case '/': out->push_back('/'); break;
        case 'b': out->push_back('\b'); break;
        case 'f': out->push_back('\f'); break;
        case 'n': out->push_back('\n'); break;
        case 'r': out->push_back('\r'); break;
        case 't': out->push_back('\t'); break;
        case 'u': {
          uint32_t cp = 0;
          if (!hex4(&cp)) return false;
          if (cp >= 0xD800 && cp <= 0xDBFF && i + 1 < s.size() && s[i] == '\\' && s[i + 1] == 'u') {
            size_t save = i;
            i += 2;
            uint32_t lo = 0;
            if (!hex4(&lo)) return false;
            if (lo >= 0xDC00 && lo <= 0xDFFF) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            else { i = save; cp = 0xFFFD; }
          } else if (cp >= 0xD800 && cp <= 0xDFFF) {
            cp = 0xFFFD;  // encoding/json replaces lone surrogates
          }
          put_utf8(*out, cp);
          break;
        }
        default: return fail("bad escape");
      }
    }
    return fail("unterminated string");
  }
  ValuePtr value(int depth) {
    if (depth > 256) { fail("nesting too deep"); return nullptr; }
    ws();
    if (i >= s.size()) { fail("unexpected end"); return nullptr; }
    auto v = std::make_shared<Value>();
    char c = s[i];
    if (c == '{') {
      ++i;
      v->type = Value::Object;
      ws();
      if (i < s.size() && s[i] == '}') { ++i; return v; }
      for (;;) {
        ws();
        std::string key;
        if (!str(&key)) return nullptr;
        ws();
        if (i >= s.size() || s[i] != ':') { fail("expected ':'"); return nullptr; }
        ++i;
        ValuePtr child = value(depth + 1);
        if (!child) return nullptr;
        v->obj[key] = child;  // later duplicates win, like encoding/json
        ws();
        if (i < s.size() && s[i] == ',') { ++i; continue; }
        if (i < s.size() && s[i] == '}') { ++i; return v; }
        fail("expected ',' or '}'");
        return nullptr;
      }
    }
    if (c == '[') {
      ++i;
      v->type = Value::Array;
      ws();
      if (i < s.size() && s[i] == ']') { ++i; return v; }
      for (;;) {
        ValuePtr child = value(depth + 1);
        if (!child) return nullptr;
        v->arr.push_back(child);
        ws();
        if (i < s.size() && s[i] == ',') { ++i; continue; }
        if (i < s.size() && s[i] == ']') { ++i; return v; }
        fail("expected ',' or ']'");
        return nullptr;
      }
    }
    if (c == '"') {
      v->type = Value::String;
      if (!str(&v->str)) return nullptr;
      return v;
    }
    if (s.compare(i, 4, "true") == 0) { i += 4; v->type = Value::Bool; v->b = true; return v; }
    if (s.compare(i, 5, "false") == 0) { i += 5; v->type = Value::Bool; v->b = false; return v; }
    if (s.compare(i, 4, "null") == 0) { i += 4; v->type = Value::Null; return v; }
    // number
    size_t j = i;
    bool is_int = true;
    if (j < s.size() && s[j] == '-') ++j;
    size_t digits0 = j;
    while (j < s.size() && s[j] >= '0' && s[j] <= '9') ++j;
    if (j == digits0) { fail("unexpected character"); return nullptr; }
    if (j < s.size() && s[j] == '.') { is_int = false; ++j; while (j < s.size() && s[j] >= '0' && s[j] <= '9') ++j; }
    if (j < s.size() && (s[j] == 'e' || s[j] == 'E')) {
      is_int = false;
      ++j;
      if (j < s.size() && (s[j] == '+' || s[j] == '-')) ++j;
      while (j < s.size() && s[j] >= '0' && s[j] <= '9') ++j;
    }
    double d = 0.0;
    auto res = std::from_chars(s.data() + i, s.data() + j, d);
    if (res.ec != std::errc()) { fail("bad number"); return nullptr; }
    i = j;
    v->type = (typed && is_int) ? Value::Int : Value::Number;
    v->num = d;
    return v;
  }
};

}  // namespace

ValuePtr parse_json(const std::string& text, bool typed_literals, std::string* err) {
  Parser p(text, typed_literals);
  ValuePtr v = p.value(0);
  if (v) {
    p.ws();
    if (p.i != text.size()) { p.fail("trailing characters"); v = nullptr; }
  }
  if (!v && err) *err = p.err;
  return v;
}

std::string format_float_v(double x) {
  if (std::isnan(x)) return "NaN";
  if (std::isinf(x)) return x > 0 ? "+Inf" : "-Inf";
  const bool neg = std::signbit(x);
  const double ax = std::fabs(x);
  if (ax == 0.0) return neg ? "-0" : "0";
  // shortest round-trip digits: std::to_chars scientific without precision is the shortest repr
  char buf[64];
  auto r = std::to_chars(buf, buf + sizeof(buf), ax, std::chars_format::scientific);
  std::string sci(buf, r.ptr);  // d.ddddde+XX
  size_t epos = sci.find('e');
  std::string mant = sci.substr(0, epos);
  int exp = std::atoi(sci.c_str() + epos + 1);
  std::string digits;
  for (char c : mant) if (c != '.') digits.push_back(c);
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  const int dp = exp + 1;  // decimal point position
  std::string out;
  // strconv %g with the shortest precision: "if precision was the shortest possible, use precision 6
  // for this decision" (ftoa.go) => exponent form when exp < -4 || exp >= 6.
  const int eprec = 6;
  if (exp < -4 || exp >= eprec) {
    out = digits.substr(0, 1);
    if (digits.size() > 1) out += "." + digits.substr(1);
    char eb[16];
    std::snprintf(eb, sizeof(eb), "e%c%02d", exp >= 0 ? '+' : '-', std::abs(exp));
    out += eb;
  } else if (dp <= 0) {
    out = "0." + std::string((size_t)(-dp), '0') + digits;
  } else if (dp >= (int)digits.size()) {
    out = digits + std::string((size_t)(dp - (int)digits.size()), '0');
  } else {
    out = digits.substr(0, (size_t)dp) + "." + digits.substr((size_t)dp);
  }
  return neg ? "-" + out : out;
}

std::string sprint_v(const Value& v) {
  switch (v.type) {
    case Value::Null: return "<nil>";
    case Value::Bool: return v.b ? "true" : "false";
    case Value::String: return v.str;
    case Value::Number: return format_float_v(v.num);
    case Value::Int: {
      char buf[32];
      std::snprintf(buf, sizeof(buf), "%lld", (long long)v.num);
      return buf;
    }
    case Value::Array: {
      std::string out = "[";
      for (size_t i = 0; i < v.arr.size(); ++i) {
        if (i) out += " ";
        out += sprint_v(*v.arr[i]);
      }
      return out + "]";
    }
    case Value::Object: {
      std::string out = "map[";
      bool first = true;
      for (const auto& kv : v.obj) {
        if (!first) out += " ";
        first = false;
        out += kv.first + ":" + sprint_v(*kv.second);
      }
      return out + "]";
    }
  }
  return "";
}

namespace {

// Simple (1:1) case folding for the scripts the tests exercise: ASCII, Latin-1 Supplement,
// Latin Extended-A, Greek and Cyrillic basic ranges, plus the two orbits with an ASCII member.
uint32_t fold_cp(uint32_t c) {
  if (c < 0x80) return (c >= 'A' && c <= 'Z') ? c + 32 : c;
  if (c == 0x212A) return 'k';                                   // KELVIN SIGN
  if (c == 0x017F) return 's';                                   // LATIN SMALL LETTER LONG S
  if (c >= 0xC0 && c <= 0xDE && c != 0xD7) return c + 32;        // Latin-1 capitals
  if (c >= 0x100 && c <= 0x17E) {                                // Latin Extended-A pairs
    if ((c >= 0x139 && c <= 0x148) || (c >= 0x179 && c <= 0x17E)) return (c & 1) ? c + 1 : c;
    if (c == 0x130 || c == 0x131 || c == 0x138 || c == 0x149 || c == 0x178) return c == 0x178 ? 0xFF : c;
    return (c & 1) ? c : c + 1;
  }
  if (c >= 0x391 && c <= 0x3A9 && c != 0x3A2) return c + 32;     // Greek capitals
  if (c == 0x3C2) return 0x3C3;                                  // final sigma folds with sigma
  if (c >= 0x410 && c <= 0x42F) return c + 32;                   // Cyrillic А..Я
  if (c >= 0x400 && c <= 0x40F) return c + 80;                   // Cyrillic Ѐ..Џ
  return c;
}

}  // namespace

std::string fold_key(const std::string& s) {
  std::string out;
  out.reserve(s.size());
  size_t i = 0;
  while (i < s.size()) {
    unsigned char c = (unsigned char)s[i];
    uint32_t cp;
    int len;
    if (c < 0x80) { cp = c; len = 1; }
    else if ((c >> 5) == 6 && i + 1 < s.size()) { cp = ((c & 0x1F) << 6) | ((unsigned char)s[i + 1] & 0x3F); len = 2; }
    else if ((c >> 4) == 14 && i + 2 < s.size()) {
      cp = ((c & 0x0F) << 12) | (((unsigned char)s[i + 1] & 0x3F) << 6) | ((unsigned char)s[i + 2] & 0x3F);
      len = 3;
    } else if ((c >> 3) == 30 && i + 3 < s.size()) {
      cp = ((c & 0x07) << 18) | (((unsigned char)s[i + 1] & 0x3F) << 12) | (((unsigned char)s[i + 2] & 0x3F) << 6) |
           ((unsigned char)s[i + 3] & 0x3F);
      len = 4;
    } else { out.push_back((char)c); ++i; continue; }  // invalid byte: kept verbatim
    i += (size_t)len;
    cp = fold_cp(cp);
    if (cp < 0x80) out.push_back((char)cp);
    else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
    else if (cp < 0x10000) {
      out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
      out.push_back((char)(0x80 | (cp & 0x3F)));
    } else {
      out.push_back((char)(0xF0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
      out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F)));
    }
  }
  return out;
}

bool deep_equal(const Value& a, const Value& b) {
  if (a.type != b.type) return false;
  switch (a.type) {
    case Value::Null: return true;
    case Value::Bool: return a.b == b.b;
    case Value::Number:
    case Value::Int: return a.num == b.num;
    case Value::String: return a.str == b.str;
    case Value::Array:
      if (a.arr.size() != b.arr.size()) return false;
      for (size_t i = 0; i < a.arr.size(); ++i) if (!deep_equal(*a.arr[i], *b.arr[i])) return false;
      return true;
    case Value::Object:
      if (a.obj.size() != b.obj.size()) return false;
      for (const auto& kv : a.obj) {
        auto it = b.obj.find(kv.first);
        if (it == b.obj.end() || !deep_equal(*kv.second, *it->second)) return false;
      }
      return true;
  }
  return false;
}

}  // namespace qh
