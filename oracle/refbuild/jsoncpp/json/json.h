/* Header-only stand-in for <jsoncpp/json/json.h>, used ONLY to compile the
 * unmodified reference under oracle/_ref (libjsoncpp is not installed in this
 * image).  Implements the small API surface the reference touches:
 *   Json::Reader::parse(std::istream&, Json::Value&)
 *   Json::Value::operator[](int / const char* / std::string), size(),
 *   asInt(), asDouble(), asBool(), asString(), isNull()
 * Semantics that matter to the reference: size() of a scalar or null is 0
 * (single-cell-with-interference.h:238-245 calls .size() on "video_bitrate": 0),
 * a missing key yields a null value whose asInt()/asDouble() are 0.
 * Test infrastructure; not part of the product.
 */
#ifndef RS_JSON_STANDIN_H_
#define RS_JSON_STANDIN_H_

#include <cstdlib>
#include <istream>
#include <iterator>
#include <map>
#include <string>
#include <vector>

namespace Json {

class Value {
 public:
  enum Kind { kNull, kNumber, kString, kBool, kArray, kObject };
  Value() : kind_(kNull), num_(0), flag_(false) {}

  unsigned size() const {
    if (kind_ == kArray) return (unsigned)items_.size();
    if (kind_ == kObject) return (unsigned)members_.size();
    return 0;
  }
  const Value& operator[](int i) const {
    if (kind_ != kArray || i < 0 || i >= (int)items_.size()) return Null();
    return items_[i];
  }
  const Value& operator[](unsigned i) const { return (*this)[(int)i]; }
  const Value& operator[](const char* key) const {
    if (kind_ != kObject) return Null();
    std::map<std::string, Value>::const_iterator it = members_.find(key);
    return it == members_.end() ? Null() : it->second;
  }
  const Value& operator[](const std::string& key) const { return (*this)[key.c_str()]; }
  int asInt() const { return kind_ == kNumber ? (int)num_ : (kind_ == kBool ? (int)flag_ : 0); }
  unsigned asUInt() const { return (unsigned)asInt(); }
  double asDouble() const { return kind_ == kNumber ? num_ : (kind_ == kBool ? (double)flag_ : 0.0); }
  bool asBool() const { return kind_ == kBool ? flag_ : (kind_ == kNumber ? num_ != 0 : false); }
  std::string asString() const { return kind_ == kString ? text_ : std::string(); }
  bool isNull() const { return kind_ == kNull; }

 private:
  friend class Reader;
  static const Value& Null() {
    static const Value v;
    return v;
  }
  Kind kind_;
  double num_;
  bool flag_;
  std::string text_;
  std::vector<Value> items_;
  std::map<std::string, Value> members_;
};

class Reader {
 public:
  bool parse(std::istream& is, Value& root) {
    std::string doc((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());
    return parse(doc, root);
  }
  bool parse(const std::string& doc, Value& root) {
    s_ = doc.c_str();
    p_ = 0;
    n_ = doc.size();
    root = Value();
    if (!ParseValue(root)) return false;
    Skip();
    return true;
  }

 private:
  void Skip() {
    while (p_ < n_) {
      char c = s_[p_];
      if (c == ' ' || c == '\t' || c == '\n' || c == '\r') {
        ++p_;
      } else if (c == '/' && p_ + 1 < n_ && s_[p_ + 1] == '/') {
        while (p_ < n_ && s_[p_] != '\n') ++p_;
      } else if (c == '/' && p_ + 1 < n_ && s_[p_ + 1] == '*') {
        p_ += 2;
        while (p_ + 1 < n_ && !(s_[p_] == '*' && s_[p_ + 1] == '/')) ++p_;
        p_ += 2;
      } else {
        break;
      }
    }
  }
  bool ParseString(std::string& out) {
    if (s_[p_] != '"') return false;
    ++p_;
    out.clear();
    while (p_ < n_ && s_[p_] != '"') {
      if (s_[p_] == '\\' && p_ + 1 < n_) {
        ++p_;
        char c = s_[p_];
        out.push_back(c == 'n' ? '\n' : c == 't' ? '\t' : c);
      } else {
        out.push_back(s_[p_]);
      }
      ++p_;
    }
    if (p_ >= n_) return false;
    ++p_;
    return true;
  }
  bool ParseValue(Value& v) {
    Skip();
    if (p_ >= n_) return false;
    char c = s_[p_];
    if (c == '{') {
      v.kind_ = Value::kObject;
      ++p_;
      Skip();
      if (p_ < n_ && s_[p_] == '}') { ++p_; return true; }
      while (true) {
        Skip();
        std::string key;
        if (!ParseString(key)) return false;
        Skip();
        if (p_ >= n_ || s_[p_] != ':') return false;
        ++p_;
        Value child;
        if (!ParseValue(child)) return false;
        v.members_[key] = child;
        Skip();
        if (p_ < n_ && s_[p_] == ',') { ++p_; continue; }
        if (p_ < n_ && s_[p_] == '}') { ++p_; return true; }
        return false;
      }
    }
    if (c == '[') {
      v.kind_ = Value::kArray;
      ++p_;
      Skip();
      if (p_ < n_ && s_[p_] == ']') { ++p_; return true; }
      while (true) {
        Value child;
        if (!ParseValue(child)) return false;
        v.items_.push_back(child);
        Skip();
        if (p_ < n_ && s_[p_] == ',') { ++p_; continue; }
        if (p_ < n_ && s_[p_] == ']') { ++p_; return true; }
        return false;
      }
    }
    if (c == '"') {
      v.kind_ = Value::kString;
      return ParseString(v.text_);
    }
    if (n_ - p_ >= 4 && std::string(s_ + p_, 4) == "true") { v.kind_ = Value::kBool; v.flag_ = true; p_ += 4; return true; }
    if (n_ - p_ >= 5 && std::string(s_ + p_, 5) == "false") { v.kind_ = Value::kBool; v.flag_ = false; p_ += 5; return true; }
    if (n_ - p_ >= 4 && std::string(s_ + p_, 4) == "null") { v.kind_ = Value::kNull; p_ += 4; return true; }
    char* end = 0;
    double d = std::strtod(s_ + p_, &end);
    if (end == s_ + p_) return false;
    v.kind_ = Value::kNumber;
    v.num_ = d;
    p_ = (size_t)(end - s_);
    return true;
  }
  const char* s_;
  size_t p_;
  size_t n_;
};

}  // namespace Json

#endif /* RS_JSON_STANDIN_H_ */
