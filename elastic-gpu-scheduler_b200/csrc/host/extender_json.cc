#include "extender_json.h"

#include <cstring>

namespace egs {

// ------------------------------------------------------------------------------------ interner
static uint64_t fnv1a(std::string_view s) {
  uint64_t h = 1469598103934665603ull;
  for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
  return h;
}
NodeInterner::NodeInterner() : table_(1024, -1) {}
void NodeInterner::Grow() {
  std::vector<int32_t> t(table_.size() * 2, -1);
  const size_t mask = t.size() - 1;
  for (int id = 0; id < (int)names_.size(); id++) {
    size_t i = fnv1a(names_[id]) & mask;
    while (t[i] >= 0) i = (i + 1) & mask;
    t[i] = id;
  }
  table_.swap(t);
}
int NodeInterner::Find(std::string_view name) const {
  const size_t mask = table_.size() - 1;
  for (size_t i = fnv1a(name) & mask;; i = (i + 1) & mask) {
    const int id = table_[i];
    if (id < 0) return -1;
    if (names_[id] == name) return id;
  }
}
int NodeInterner::Intern(std::string_view name) {
  int id = Find(name);
  if (id >= 0) return id;
  if ((names_.size() + 1) * 2 > table_.size()) Grow();
  id = (int)names_.size();
  names_.emplace_back(name);
  const size_t mask = table_.size() - 1;
  size_t i = fnv1a(name) & mask;
  while (table_[i] >= 0) i = (i + 1) & mask;
  table_[i] = id;
  return id;
}

// ------------------------------------------------------------------------------------ quantity
bool ParseQuantityValue(std::string_view q, int64_t *out) {
  size_t i = 0;
  bool neg = false;
  if (i < q.size() && (q[i] == '+' || q[i] == '-')) { neg = q[i] == '-'; i++; }
  // mantissa as integer `mant` scaled by 10^-frac
  unsigned __int128 mant = 0;
  int frac = 0, digits = 0;
  bool seen_dot = false;
  for (; i < q.size(); i++) {
    const char c = q[i];
    if (c >= '0' && c <= '9') { mant = mant * 10 + (unsigned)(c - '0'); digits++; if (seen_dot) frac++; if (digits > 30) return false; }
    else if (c == '.' && !seen_dot) seen_dot = true;
    else break;
  }
  if (digits == 0) return false;
  std::string_view suf = q.substr(i);
  int exp10 = 0; int exp2 = 0;
  if (suf.empty()) {}
  else if (suf == "m") exp10 = -3; else if (suf == "k") exp10 = 3; else if (suf == "M") exp10 = 6;
  else if (suf == "G") exp10 = 9; else if (suf == "T") exp10 = 12; else if (suf == "P") exp10 = 15; else if (suf == "E") exp10 = 18;
  else if (suf == "Ki") exp2 = 10; else if (suf == "Mi") exp2 = 20; else if (suf == "Gi") exp2 = 30;
  else if (suf == "Ti") exp2 = 40; else if (suf == "Pi") exp2 = 50; else if (suf == "Ei") exp2 = 60;
  else if ((suf[0] == 'e' || suf[0] == 'E') && suf.size() > 1) {
    size_t j = 1; bool eneg = false; int e = 0;
    if (suf[j] == '+' || suf[j] == '-') { eneg = suf[j] == '-'; j++; }
    if (j >= suf.size()) return false;
    for (; j < suf.size(); j++) { if (suf[j] < '0' || suf[j] > '9') return false; e = e * 10 + (suf[j] - '0'); if (e > 40) return false; }
    exp10 = eneg ? -e : e;
  } else return false;
  exp10 -= frac;
  unsigned __int128 v = mant;
  for (int k = 0; k < exp2; k++) { v <<= 1; if (v >> 100) return false; }
  for (; exp10 > 0; exp10--) { v *= 10; if (v >> 100) return false; }
  bool rem = false;
  for (; exp10 < 0; exp10++) { if (v % 10) rem = true; v /= 10; }
  if (rem && !neg) v += 1;                                   // Value() rounds up (toward +inf)
  if (v > (unsigned __int128)INT64_MAX) return false;
  *out = neg ? -(int64_t)v : (int64_t)v;
  return true;
}

// ------------------------------------------------------------------------------------ parser
namespace {
struct P {
  const char *p, *e;
  std::string err;
  void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
  bool fail(const char *m) { if (err.empty()) err = m; return false; }
  bool lit(const char *s) { size_t n = strlen(s); if ((size_t)(e - p) < n || memcmp(p, s, n)) return false; p += n; return true; }
  int depth = 0;                                              // nesting of the value being read (encoding/json caps it at 10000)
  bool enter() { if (++depth > 10000) return fail("exceeded max depth"); return true; }
  bool hex4(unsigned *cp) {
    if (e - p < 4) return false;
    unsigned v = 0;
    for (int k = 0; k < 4; k++) {
      const char h = *p++;
      if (h >= '0' && h <= '9') v = v * 16 + (h - '0'); else if ((h | 32) >= 'a' && (h | 32) <= 'f') v = v * 16 + ((h | 32) - 'a' + 10); else return false;
    }
    *cp = v;
    return true;
  }
  bool str(std::string *out) {                                // out may be null (skip)
    ws();
    if (p >= e || *p != '"') return fail("string expected");
    p++;
    while (p < e && *p != '"') {
      if (*p == '\\') {
        if (++p >= e) return fail("bad escape");
        char c = *p++;
        if (c == 'u') {
          unsigned cp = 0;
          if (!hex4(&cp)) return fail("invalid character in \\u hexadecimal character escape");
          if (cp >= 0xD800 && cp < 0xDC00) {                      // high surrogate: needs \uDC00..\uDFFF right after
            unsigned lo = 0; const char *save = p;
            if (e - p >= 6 && p[0] == '\\' && p[1] == 'u' && (p += 2, hex4(&lo)) && lo >= 0xDC00 && lo < 0xE000) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            else { p = save; cp = 0xFFFD; }                       // encoding/json: unpaired surrogate -> U+FFFD
          } else if (cp >= 0xDC00 && cp < 0xE000) cp = 0xFFFD;
          if (out) { if (cp < 0x80) out->push_back((char)cp); else if (cp < 0x800) { out->push_back((char)(0xC0 | cp >> 6)); out->push_back((char)(0x80 | (cp & 63))); }
                     else if (cp < 0x10000) { out->push_back((char)(0xE0 | cp >> 12)); out->push_back((char)(0x80 | ((cp >> 6) & 63))); out->push_back((char)(0x80 | (cp & 63))); }
                     else { out->push_back((char)(0xF0 | cp >> 18)); out->push_back((char)(0x80 | ((cp >> 12) & 63))); out->push_back((char)(0x80 | ((cp >> 6) & 63))); out->push_back((char)(0x80 | (cp & 63))); } }
        } else if (out) {
          out->push_back(c == 'n' ? '\n' : c == 't' ? '\t' : c == 'r' ? '\r' : c == 'b' ? '\b' : c == 'f' ? '\f' : c);
        }
      } else { if (out) out->push_back(*p); p++; }
    }
    if (p >= e) return fail("unterminated string");
    p++;
    return true;
  }
  bool skip() {                                               // any JSON value
    ws();
    if (p >= e) return fail("value expected");
    if (*p == '"') return str(nullptr);
    if (*p == '{' || *p == '[') {
      const char close = *p == '{' ? '}' : ']';
      const bool obj = *p == '{';
      if (!enter()) return false;
      p++; ws();
      if (p < e && *p == close) { p++; depth--; return true; }
      for (;;) {
        if (obj) { if (!str(nullptr)) return false; ws(); if (p >= e || *p++ != ':') return fail("':' expected"); }
        if (!skip()) return false;
        ws();
        if (p < e && *p == ',') { p++; continue; }
        if (p < e && *p == close) { p++; depth--; return true; }
        return fail("',' expected");
      }
    }
    while (p < e && *p != ',' && *p != '}' && *p != ']' && *p != ' ' && *p != '\n' && *p != '\t' && *p != '\r') p++;
    return true;
  }
  // object walker: calls f(key) for each member; f must consume the value
  template <class F> bool object(F f) {
    ws();
    if (lit("null")) return true;
    if (p >= e || *p != '{') return fail("object expected");
    if (!enter()) return false;
    p++; ws();
    if (p < e && *p == '}') { p++; depth--; return true; }
    for (;;) {
      std::string key;
      if (!str(&key)) return false;
      ws();
      if (p >= e || *p++ != ':') return fail("':' expected");
      if (!f(key)) return false;
      ws();
      if (p < e && *p == ',') { p++; continue; }
      if (p < e && *p == '}') { p++; depth--; return true; }
      return fail("',' or '}' expected");
    }
  }
  template <class F> bool array(F f) {
    ws();
    if (lit("null")) return true;
    if (p >= e || *p != '[') return fail("array expected");
    if (!enter()) return false;
    p++; ws();
    if (p < e && *p == ']') { p++; depth--; return true; }
    for (;;) {
      if (!f()) return false;
      ws();
      if (p < e && *p == ',') { p++; continue; }
      if (p < e && *p == ']') { p++; depth--; return true; }
      return fail("',' or ']' expected");
    }
  }
  bool string_map(std::map<std::string, std::string> *m) {
    return object([&](const std::string &k) { std::string v; if (!str(&v)) return false; (*m)[k] = v; return true; });
  }
};

bool parse_quantity(P &ps, int64_t *out) {                    // a quantity is a JSON string or a bare number
  ps.ws();
  std::string q;
  if (ps.p < ps.e && *ps.p == '"') { if (!ps.str(&q)) return false; }
  else { const char *b = ps.p; if (!ps.skip()) return false; q.assign(b, ps.p); }
  if (!ParseQuantityValue(q, out)) return ps.fail("bad quantity");
  return true;
}

bool parse_pod(P &ps, Pod *pod) {
  return ps.object([&](const std::string &k) {
    if (k == "metadata")
      return ps.object([&](const std::string &m) {
        if (m == "name") return ps.str(&pod->name);
        if (m == "namespace") return ps.str(&pod->ns);
        if (m == "uid") return ps.str(&pod->uid);
        if (m == "annotations") return ps.string_map(&pod->annotations);
        if (m == "labels") return ps.string_map(&pod->labels);
        return ps.skip();
      });
    if (k == "spec")
      return ps.object([&](const std::string &m) {
        if (m == "nodeName") return ps.str(&pod->node_name);
        if (m == "containers")
          return ps.array([&]() {
            Container c;
            bool ok = ps.object([&](const std::string &ck) {
              if (ck == "name") return ps.str(&c.name);
              if (ck == "resources")
                return ps.object([&](const std::string &rk) {
                  if (rk != "requests") return ps.skip();      // the path reads Requests only (pod.go:94-108)
                  return ps.object([&](const std::string &res) {
                    int64_t v;
                    if (res == kResourceGPUCore || res == kResourceGPUMemory) { if (!parse_quantity(ps, &v)) return false; c.requests[res] = v; return true; }
                    return ps.skip();
                  });
                });
              return ps.skip();
            });
            if (ok) pod->containers.push_back(std::move(c));
            return ok;
          });
        return ps.skip();
      });
    return ps.skip();
  });
}
}  // namespace

std::string ParseExtenderArgs(std::string_view json, NodeInterner *nodes, ExtenderArgs *out) {
  P ps{json.data(), json.data() + json.size(), ""};
  *out = ExtenderArgs();
  bool ok = ps.object([&](const std::string &k) {
    if (k == "pod" || k == "Pod") return parse_pod(ps, &out->pod);
    if (k == "nodenames" || k == "NodeNames") {
      ps.ws();
      if (ps.lit("null")) return true;
      out->has_nodenames = true;
      return ps.array([&]() {
        // fast path: a name without escapes is interned straight from the input buffer
        ps.ws();
        if (ps.p < ps.e && *ps.p == '"') {
          const char *b = ps.p + 1, *q = b;
          while (q < ps.e && *q != '"' && *q != '\\') q++;
          if (q < ps.e && *q == '"') { out->node_ids.push_back(nodes->Intern(std::string_view(b, q - b))); ps.p = q + 1; return true; }
        }
        std::string s;
        if (!ps.str(&s)) return false;
        out->node_ids.push_back(nodes->Intern(s));
        return true;
      });
    }
    return ps.skip();
  });
  if (ok) { ps.ws(); if (ps.p != ps.e) { ok = false; ps.err = "trailing data"; } }
  return ok ? "" : (ps.err.empty() ? "parse error" : ps.err);
}

std::string ParseBindingArgs(std::string_view json, BindingArgs *out) {
  P ps{json.data(), json.data() + json.size(), ""};
  *out = BindingArgs();
  bool ok = ps.object([&](const std::string &k) {
    if (k == "podName" || k == "PodName") return ps.str(&out->pod_name);
    if (k == "podNamespace" || k == "PodNamespace") return ps.str(&out->pod_namespace);
    if (k == "podUID" || k == "PodUID") return ps.str(&out->pod_uid);
    if (k == "node" || k == "Node") return ps.str(&out->node);
    return ps.skip();
  });
  return ok ? "" : (ps.err.empty() ? "parse error" : ps.err);
}

// ------------------------------------------------------------------------------------ encoders
void AppendJsonString(std::string *out, std::string_view s) {   // encoding/json: HTML-safe, \u00XX for controls
  static const char *hex = "0123456789abcdef";
  out->push_back('"');
  for (size_t i = 0; i < s.size(); i++) {
    const unsigned char c = (unsigned char)s[i];
    if (c == '"' || c == '\\') { out->push_back('\\'); out->push_back((char)c); }
    else if (c == '\n') *out += "\\n"; else if (c == '\r') *out += "\\r"; else if (c == '\t') *out += "\\t";
    else if (c < 0x20 || c == '<' || c == '>' || c == '&') { *out += "\\u00"; out->push_back(hex[c >> 4]); out->push_back(hex[c & 15]); }
    else if (c == 0xE2 && i + 2 < s.size() && (unsigned char)s[i + 1] == 0x80 && ((unsigned char)s[i + 2] == 0xA8 || (unsigned char)s[i + 2] == 0xA9)) {
      *out += (unsigned char)s[i + 2] == 0xA8 ? "\\u2028" : "\\u2029"; i += 2;
    } else out->push_back((char)c);
  }
  out->push_back('"');
}

std::string EncodeFilterResult(const std::vector<std::string> &node_names, const std::map<std::string, std::string> &failed,
                               const std::string &error, bool has_node_names) {
  // ExtenderFilterResult{Nodes omitempty (nil), NodeNames *[]string omitempty, FailedNodes omitempty, Error omitempty}.
  // predicate.go:33-37 sets NodeNames = &filterdNodes (a non-nil pointer: present even when empty); every error
  // path (predicate.go:21-31, routes.go:51-64) leaves NodeNames nil: the member is omitted.
  std::string o;
  o.reserve(32 + node_names.size() * 16);
  o += "{";
  bool any = false;
  if (has_node_names) {
    o += "\"nodenames\":[";
    for (size_t i = 0; i < node_names.size(); i++) { if (i) o.push_back(','); AppendJsonString(&o, node_names[i]); }
    o += "]";
    any = true;
  }
  if (!failed.empty()) {
    o += any ? ",\"failedNodes\":{" : "\"failedNodes\":{";
    any = true;
    bool first = true;
    for (const auto &kv : failed) { if (!first) o.push_back(','); first = false; AppendJsonString(&o, kv.first); o.push_back(':'); AppendJsonString(&o, kv.second); }
    o += "}";
  }
  if (!error.empty()) { o += any ? ",\"error\":" : "\"error\":"; AppendJsonString(&o, error); }
  o += "}";
  return o;
}

std::string EncodeHostPriorityList(const std::vector<std::pair<std::string, int64_t>> &scores) {
  std::string o = "[";
  for (size_t i = 0; i < scores.size(); i++) {
    if (i) o.push_back(',');
    o += "{\"host\":"; AppendJsonString(&o, scores[i].first); o += ",\"score\":"; o += std::to_string(scores[i].second); o += "}";
  }
  o += "]";
  return o;
}

std::string EncodeBindingResult(const std::string &error) {
  if (error.empty()) return "{}";
  std::string o = "{\"error\":"; AppendJsonString(&o, error); o += "}";
  return o;
}

}  // namespace egs
