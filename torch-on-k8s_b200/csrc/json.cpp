// Recursive-descent JSON parser / serialiser (RFC 8259; \uXXXX incl. surrogate pairs -> UTF-8).
#include "json.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace tok {
namespace json {
namespace {

struct Parser {
  const char* p;
  const char* begin;
  std::string err;
  int depth = 0;

  bool fail(const char* msg) {
    char buf[160];
    snprintf(buf, sizeof(buf), "json: %s at byte %ld", msg, static_cast<long>(p - begin));
    err = buf;
    return false;
  }
  void ws() {
    while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') ++p;
  }
  static void utf8(std::string& out, unsigned cp) {
    if (cp < 0x80) {
      out += static_cast<char>(cp);
    } else if (cp < 0x800) {
      out += static_cast<char>(0xC0 | (cp >> 6));
      out += static_cast<char>(0x80 | (cp & 0x3F));
    } else if (cp < 0x10000) {
      out += static_cast<char>(0xE0 | (cp >> 12));
      out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F));
      out += static_cast<char>(0x80 | (cp & 0x3F));
    } else {
      out += static_cast<char>(0xF0 | (cp >> 18));
      out += static_cast<char>(0x80 | ((cp >> 12) & 0x3F));
      out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F));
      out += static_cast<char>(0x80 | (cp & 0x3F));
    }
  }
  bool hex4(unsigned* v) {
    unsigned x = 0;
    for (int k = 0; k < 4; ++k) {
      char c = *p++;
      x <<= 4;
      if (c >= '0' && c <= '9')
        x |= static_cast<unsigned>(c - '0');
      else if (c >= 'a' && c <= 'f')
        x |= static_cast<unsigned>(c - 'a' + 10);
      else if (c >= 'A' && c <= 'F')
        x |= static_cast<unsigned>(c - 'A' + 10);
      else
        return fail("bad \\u escape");
    }
    *v = x;
    return true;
  }
  bool string(std::string* out) {
    if (*p != '"') return fail("expected string");
    ++p;
    out->clear();
    for (;;) {
      unsigned char c = static_cast<unsigned char>(*p);
      if (c == 0) return fail("unterminated string");
      if (c == '"') {
        ++p;
        return true;
      }
      if (c < 0x20) return fail("control character in string");
      if (c != '\\') {
        *out += static_cast<char>(c);
        ++p;
        continue;
      }
      ++p;
      char e = *p++;
      switch (e) {
        case '"': *out += '"'; break;
        case '\\': *out += '\\'; break;
        case '/': *out += '/'; break;
        case 'b': *out += '\b'; break;
        case 'f': *out += '\f'; break;
        case 'n': *out += '\n'; break;
        case 'r': *out += '\r'; break;
        case 't': *out += '\t'; break;
        case 'u': {
          unsigned cp = 0;
          if (!hex4(&cp)) return false;
          if (cp >= 0xD800 && cp <= 0xDBFF && p[0] == '\\' && p[1] == 'u') {
            p += 2;
            unsigned lo = 0;
            if (!hex4(&lo)) return false;
            if (lo >= 0xDC00 && lo <= 0xDFFF)
              cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            else
              return fail("bad surrogate pair");
          }
          utf8(*out, cp);
          break;
        }
        default:
          return fail("bad escape");
      }
    }
  }
  bool number(Value* out) {
    const char* s = p;
    if (*p == '-') ++p;
    if (*p < '0' || *p > '9') return fail("bad number");
    if (*p == '0') {
      ++p;
    } else {
      while (*p >= '0' && *p <= '9') ++p;
    }
    bool integral = true;
    if (*p == '.') {
      integral = false;
      ++p;
      if (*p < '0' || *p > '9') return fail("bad fraction");
      while (*p >= '0' && *p <= '9') ++p;
    }
    if (*p == 'e' || *p == 'E') {
      integral = false;
      ++p;
      if (*p == '+' || *p == '-') ++p;
      if (*p < '0' || *p > '9') return fail("bad exponent");
      while (*p >= '0' && *p <= '9') ++p;
    }
    std::string tok(s, static_cast<size_t>(p - s));
    if (integral && tok.size() < 19) {
      *out = Value::integer(strtoll(tok.c_str(), nullptr, 10));
    } else {
      *out = Value::number(strtod(tok.c_str(), nullptr));
    }
    return true;
  }
  bool value(Value* out) {
    if (++depth > 256) return fail("nesting too deep");
    ws();
    bool ok = false;
    switch (*p) {
      case '{': {
        ++p;
        *out = Value::object();
        ws();
        if (*p == '}') {
          ++p;
          ok = true;
          break;
        }
        for (;;) {
          ws();
          std::string k;
          if (!string(&k)) return false;
          ws();
          if (*p != ':') return fail("expected ':'");
          ++p;
          Value v;
          if (!value(&v)) return false;
          if (Value* old = out->find(k))
            *old = std::move(v);  // last duplicate wins (encoding/json behaviour)
          else
            out->o.emplace_back(std::move(k), std::move(v));
          ws();
          if (*p == ',') {
            ++p;
            continue;
          }
          if (*p == '}') {
            ++p;
            ok = true;
            break;
          }
          return fail("expected ',' or '}'");
        }
        break;
      }
      case '[': {
        ++p;
        *out = Value::array();
        ws();
        if (*p == ']') {
          ++p;
          ok = true;
          break;
        }
        for (;;) {
          Value v;
          if (!value(&v)) return false;
          out->a.push_back(std::move(v));
          ws();
          if (*p == ',') {
            ++p;
            continue;
          }
          if (*p == ']') {
            ++p;
            ok = true;
            break;
          }
          return fail("expected ',' or ']'");
        }
        break;
      }
      case '"': {
        std::string s;
        if (!string(&s)) return false;
        *out = Value::str(s);
        ok = true;
        break;
      }
      case 't':
        if (strncmp(p, "true", 4) == 0) {
          p += 4;
          *out = Value::boolean(true);
          ok = true;
        }
        break;
      case 'f':
        if (strncmp(p, "false", 5) == 0) {
          p += 5;
          *out = Value::boolean(false);
          ok = true;
        }
        break;
      case 'n':
        if (strncmp(p, "null", 4) == 0) {
          p += 4;
          *out = Value();
          ok = true;
        }
        break;
      default:
        if (*p == '-' || (*p >= '0' && *p <= '9')) ok = number(out);
        if (!ok && err.empty()) return fail("unexpected character");
        break;
    }
    --depth;
    if (!ok && err.empty()) return fail("bad literal");
    return ok;
  }
};

void dump_string(const std::string& s, std::string& out) {
  out += '"';
  for (unsigned char c : s) {
    switch (c) {
      case '"': out += "\\\""; break;
      case '\\': out += "\\\\"; break;
      case '\n': out += "\\n"; break;
      case '\r': out += "\\r"; break;
      case '\t': out += "\\t"; break;
      case '\b': out += "\\b"; break;
      case '\f': out += "\\f"; break;
      default:
        if (c < 0x20) {
          char buf[8];
          snprintf(buf, sizeof(buf), "\\u%04x", c);
          out += buf;
        } else {
          out += static_cast<char>(c);
        }
    }
  }
  out += '"';
}

void dump_value(const Value& v, std::string& out) {
  switch (v.type) {
    case Value::Null: out += "null"; break;
    case Value::Bool: out += v.b ? "true" : "false"; break;
    case Value::Int: out += std::to_string(v.i); break;
    case Value::Double: {
      if (!isfinite(v.d)) {
        out += "null";
        break;
      }
      char buf[40];
      snprintf(buf, sizeof(buf), "%.17g", v.d);
      // prefer the shortest representation that round-trips
      for (int prec = 1; prec < 17; ++prec) {
        char t[40];
        snprintf(t, sizeof(t), "%.*g", prec, v.d);
        if (strtod(t, nullptr) == v.d) {
          memcpy(buf, t, sizeof(t));
          break;
        }
      }
      out += buf;
      if (!strpbrk(buf, ".eEn")) out += ".0";
      break;
    }
    case Value::String: dump_string(v.s, out); break;
    case Value::Array: {
      out += '[';
      for (size_t n = 0; n < v.a.size(); ++n) {
        if (n) out += ',';
        dump_value(v.a[n], out);
      }
      out += ']';
      break;
    }
    case Value::Object: {
      out += '{';
      for (size_t n = 0; n < v.o.size(); ++n) {
        if (n) out += ',';
        dump_string(v.o[n].first, out);
        out += ':';
        dump_value(v.o[n].second, out);
      }
      out += '}';
      break;
    }
  }
}

}  // namespace

bool parse(const char* text, Value* out, std::string* err) {
  if (!text) {
    if (err) *err = "json: null input";
    return false;
  }
  Parser ps;
  ps.p = ps.begin = text;
  Value v;
  if (!ps.value(&v)) {
    if (err) *err = ps.err;
    return false;
  }
  ps.ws();
  if (*ps.p != 0) {
    ps.fail("trailing characters");
    if (err) *err = ps.err;
    return false;
  }
  *out = std::move(v);
  return true;
}

std::string dump(const Value& v) {
  std::string out;
  dump_value(v, out);
  return out;
}

}  // namespace json
}  // namespace tok
