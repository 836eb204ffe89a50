// Minimal JSON document model for the TorchJob surface (manifests in, status/decisions out).
// Objects keep insertion order so that round-tripped manifests stay diff-able.
#pragma once
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

namespace tok {
namespace json {

struct Value {
  enum Type { Null, Bool, Int, Double, String, Array, Object };
  Type type = Null;
  bool b = false;
  int64_t i = 0;
  double d = 0;
  std::string s;
  std::vector<Value> a;
  std::vector<std::pair<std::string, Value>> o;

  static Value object() {
    Value v;
    v.type = Object;
    return v;
  }
  static Value array() {
    Value v;
    v.type = Array;
    return v;
  }
  static Value str(const std::string& x) {
    Value v;
    v.type = String;
    v.s = x;
    return v;
  }
  static Value integer(int64_t x) {
    Value v;
    v.type = Int;
    v.i = x;
    return v;
  }
  static Value number(double x) {
    Value v;
    v.type = Double;
    v.d = x;
    return v;
  }
  static Value boolean(bool x) {
    Value v;
    v.type = Bool;
    v.b = x;
    return v;
  }

  bool is_null() const { return type == Null; }
  bool is_object() const { return type == Object; }
  bool is_array() const { return type == Array; }
  bool is_string() const { return type == String; }
  bool is_number() const { return type == Int || type == Double; }

  const Value* find(const std::string& k) const {
    if (type != Object) return nullptr;
    for (const auto& kv : o)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
  Value* find(const std::string& k) {
    return const_cast<Value*>(static_cast<const Value*>(this)->find(k));
  }
  // object member access, inserting a null member when missing (turns Null into Object)
  Value& operator[](const std::string& k) {
    if (type == Null) type = Object;
    if (Value* v = find(k)) return *v;
    o.emplace_back(k, Value());
    return o.back().second;
  }
  void erase(const std::string& k) {
    for (size_t n = 0; n < o.size(); ++n)
      if (o[n].first == k) {
        o.erase(o.begin() + static_cast<long>(n));
        return;
      }
  }
  const Value* path(std::initializer_list<const char*> keys) const {
    const Value* cur = this;
    for (const char* k : keys) {
      if (!cur) return nullptr;
      cur = cur->find(k);
    }
    return cur;
  }
  int64_t as_int(int64_t dflt = 0) const {
    if (type == Int) return i;
    if (type == Double) return static_cast<int64_t>(d);
    return dflt;
  }
  double as_double(double dflt = 0) const {
    if (type == Int) return static_cast<double>(i);
    if (type == Double) return d;
    return dflt;
  }
  std::string as_string(const std::string& dflt = "") const { return type == String ? s : dflt; }
  bool as_bool(bool dflt = false) const { return type == Bool ? b : dflt; }
};

// Returns false and fills *err (with a byte offset) on malformed input.
bool parse(const char* text, Value* out, std::string* err);
std::string dump(const Value& v);

}  // namespace json
}  // namespace tok
