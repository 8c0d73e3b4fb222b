// serenade_b200/csrc/avro_reader.cpp — loads the production on-disk VMIS index (Avro object container files)
// the way VMISIndex::new does (vmis_index.rs:85-313), without an Avro library:
//
//   container   magic "Obj\1", metadata map (avro.schema JSON, avro.codec), 16-byte sync marker, then blocks of
//               {record count, byte size, payload, sync}
//   codecs      null, deflate (raw RFC 1951 through zlib), snappy (decoder below + big-endian CRC32 trailer)
//   decoding    schema driven: the writer schema is parsed into a small node tree and the wanted fields are
//               picked BY NAME (serde derive at :184-192 / :249-255), whatever their position, through unions
//               (Spark writes nullable columns as ["type","null"]); every other field is skipped generically
//
// Files of one directory are decoded in parallel (the reference notes 161 s single threaded for the item
// index, :201) and merged in file-name order, a later record of the same key replacing an earlier one.
// Deviations, all towards failing loudly: a record that does not deserialize stops the LOAD with an error (the
// reference prints the error and silently skips the rest of that file, :229-232); directory iteration is in
// name order (the reference uses the unspecified fs::read_dir order, :207).
#include "avro_reader.h"

#include <dirent.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdint>
#include <atomic>
#include <cstring>
#include <map>
#include <thread>

namespace vmis {
namespace {

// ------------------------------------------------------------------------------------------------ JSON
struct Json {
  enum Type { Null, Bool, Num, Str, Arr, Obj } t = Null;
  bool b = false;
  double num = 0;
  std::string s;
  std::vector<Json> a;
  std::vector<std::pair<std::string, Json>> o;
  const Json* get(const char* key) const {
    for (auto& kv : o) if (kv.first == key) return &kv.second;
    return nullptr;
  }
};

struct JsonParser {
  const char* p; const char* e; std::string err;
  void ws() { while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
  bool fail(const char* m) { if (err.empty()) err = m; return false; }
  static void put_utf8(std::string& s, uint32_t c) {
    if (c < 0x80) s += (char)c;
    else if (c < 0x800) { s += (char)(0xC0 | (c >> 6)); s += (char)(0x80 | (c & 0x3F)); }
    else if (c < 0x10000) { s += (char)(0xE0 | (c >> 12)); s += (char)(0x80 | ((c >> 6) & 0x3F)); s += (char)(0x80 | (c & 0x3F)); }
    else { s += (char)(0xF0 | (c >> 18)); s += (char)(0x80 | ((c >> 12) & 0x3F)); s += (char)(0x80 | ((c >> 6) & 0x3F)); s += (char)(0x80 | (c & 0x3F)); }
  }
  bool str(std::string& out) {
    if (p >= e || *p != '"') return fail("JSON: string expected");
    ++p; out.clear();
    while (p < e && *p != '"') {
      if (*p == '\\') {
        if (++p >= e) return fail("JSON: bad escape");
        switch (*p) {
          case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
          case 'b': out += '\b'; break; case 'f': out += '\f'; break;
          case 'u': {
            if (e - p < 5) return fail("JSON: bad \\u escape");
            uint32_t c = 0;
            for (int i = 1; i <= 4; ++i) {
              const char h = p[i]; c <<= 4;
              if (h >= '0' && h <= '9') c |= (uint32_t)(h - '0'); else if (h >= 'a' && h <= 'f') c |= (uint32_t)(h - 'a' + 10);
              else if (h >= 'A' && h <= 'F') c |= (uint32_t)(h - 'A' + 10); else return fail("JSON: bad \\u escape");
            }
            put_utf8(out, c); p += 4; break;
          }
          default: out += *p;                      // \" \\ \/
        }
        ++p;
      } else out += *p++;
    }
    if (p >= e) return fail("JSON: unterminated string");
    ++p;
    return true;
  }
  bool value(Json& v, int depth = 0) {
    if (depth > 64) return fail("JSON: nesting too deep");
    ws();
    if (p >= e) return fail("JSON: unexpected end");
    if (*p == '{') {
      v.t = Json::Obj; ++p; ws();
      if (p < e && *p == '}') { ++p; return true; }
      for (;;) {
        ws(); std::string k;
        if (!str(k)) return false;
        ws(); if (p >= e || *p != ':') return fail("JSON: ':' expected");
        ++p; Json c;
        if (!value(c, depth + 1)) return false;
        v.o.emplace_back(std::move(k), std::move(c));
        ws(); if (p < e && *p == ',') { ++p; continue; }
        if (p < e && *p == '}') { ++p; return true; }
        return fail("JSON: ',' or '}' expected");
      }
    }
    if (*p == '[') {
      v.t = Json::Arr; ++p; ws();
      if (p < e && *p == ']') { ++p; return true; }
      for (;;) {
        Json c;
        if (!value(c, depth + 1)) return false;
        v.a.push_back(std::move(c));
        ws(); if (p < e && *p == ',') { ++p; continue; }
        if (p < e && *p == ']') { ++p; return true; }
        return fail("JSON: ',' or ']' expected");
      }
    }
    if (*p == '"') { v.t = Json::Str; return str(v.s); }
    if (e - p >= 4 && !std::memcmp(p, "true", 4)) { v.t = Json::Bool; v.b = true; p += 4; return true; }
    if (e - p >= 5 && !std::memcmp(p, "false", 5)) { v.t = Json::Bool; v.b = false; p += 5; return true; }
    if (e - p >= 4 && !std::memcmp(p, "null", 4)) { v.t = Json::Null; p += 4; return true; }
    char* end = nullptr;
    std::string num(p, std::min<size_t>((size_t)(e - p), 64));
    v.num = std::strtod(num.c_str(), &end);
    if (end == num.c_str()) return fail("JSON: value expected");
    v.t = Json::Num; p += end - num.c_str();
    return true;
  }
};

// ------------------------------------------------------------------------------------------------ schema
enum class Kind : uint8_t { Null, Boolean, Int, Long, Float, Double, Bytes, String, Record, Enum, Array, Map, Union, Fixed };
struct Node {
  Kind kind = Kind::Null;
  uint64_t fixed_size = 0;
  int items = -1;                                       // array / map element
  std::vector<std::pair<std::string, int>> fields;      // record
  std::vector<int> branches;                            // union
};
struct Schema {
  std::vector<Node> nodes;
  std::map<std::string, int> named;
  std::string err;
  int add(Kind k) { nodes.emplace_back(); nodes.back().kind = k; return (int)nodes.size() - 1; }
  int primitive(const std::string& n) {
    static const std::pair<const char*, Kind> prim[] = {{"null", Kind::Null}, {"boolean", Kind::Boolean}, {"int", Kind::Int},
        {"long", Kind::Long}, {"float", Kind::Float}, {"double", Kind::Double}, {"bytes", Kind::Bytes}, {"string", Kind::String}};
    for (auto& pk : prim) if (n == pk.first) return add(pk.second);
    return -1;
  }
  int parse(const Json& j, const std::string& ns, int depth = 0) {
    if (depth > 32) { err = "schema nesting too deep"; return -1; }
    if (j.t == Json::Str) {
      int n = primitive(j.s);
      if (n >= 0) return n;
      auto it = named.find(j.s.find('.') == std::string::npos && !ns.empty() ? ns + "." + j.s : j.s);
      if (it == named.end()) it = named.find(j.s);
      if (it == named.end()) { err = "unknown type '" + j.s + "'"; return -1; }
      return it->second;
    }
    if (j.t == Json::Arr) {
      const int u = add(Kind::Union);
      std::vector<int> br;
      for (auto& b : j.a) { const int n = parse(b, ns, depth + 1); if (n < 0) return -1; br.push_back(n); }
      nodes[u].branches = br;
      return u;
    }
    if (j.t != Json::Obj) { err = "bad schema node"; return -1; }
    const Json* ty = j.get("type");
    if (!ty) { err = "schema object without \"type\""; return -1; }
    if (ty->t != Json::Str) return parse(*ty, ns, depth + 1);            // {"type": {...}} / {"type": [...]}
    const std::string& t = ty->s;
    std::string my_ns = ns;
    if (const Json* n = j.get("namespace")) if (n->t == Json::Str) my_ns = n->s;
    auto full_name = [&]() {
      const Json* n = j.get("name");
      std::string nm = (n && n->t == Json::Str) ? n->s : std::string();
      if (nm.find('.') == std::string::npos && !my_ns.empty()) nm = my_ns + "." + nm;
      return nm;
    };
    auto reg = [&](int node) {
      const std::string fn = full_name();
      named[fn] = node;
      const size_t dot = fn.rfind('.');
      if (dot != std::string::npos) named.emplace(fn.substr(dot + 1), node);
    };
    if (t == "record" || t == "error") {
      const int r = add(Kind::Record);
      reg(r);
      const Json* fs = j.get("fields");
      if (!fs || fs->t != Json::Arr) { err = "record without fields"; return -1; }
      std::vector<std::pair<std::string, int>> fields;
      for (auto& f : fs->a) {
        const Json* fn = f.get("name"); const Json* ft = f.get("type");
        if (!fn || fn->t != Json::Str || !ft) { err = "bad record field"; return -1; }
        const int n = parse(*ft, my_ns, depth + 1);
        if (n < 0) return -1;
        fields.emplace_back(fn->s, n);
      }
      nodes[r].fields = fields;
      return r;
    }
    if (t == "enum") { const int n = add(Kind::Enum); reg(n); return n; }
    if (t == "fixed") {
      const int n = add(Kind::Fixed); reg(n);
      const Json* sz = j.get("size");
      if (!sz || sz->t != Json::Num || sz->num < 0) { err = "fixed without size"; return -1; }
      nodes[n].fixed_size = (uint64_t)sz->num;
      return n;
    }
    if (t == "array" || t == "map") {
      const int n = add(t == "array" ? Kind::Array : Kind::Map);
      const Json* it = j.get(t == "array" ? "items" : "values");
      if (!it) { err = t + " without element type"; return -1; }
      const int c = parse(*it, my_ns, depth + 1);
      if (c < 0) return -1;
      nodes[n].items = c;
      return n;
    }
    const int n = primitive(t);                                          // {"type":"long","logicalType":...}
    if (n >= 0) return n;
    auto it = named.find(t);
    if (it != named.end()) return it->second;
    err = "unknown type '" + t + "'";
    return -1;
  }
};

// ------------------------------------------------------------------------------------------------ binary decoding
struct Cursor {
  const uint8_t* p; const uint8_t* e; const char* err = nullptr;
  bool fail(const char* m) { if (!err) err = m; p = e; return false; }
  bool need(uint64_t n) { return (uint64_t)(e - p) >= n ? true : fail("truncated data"); }
  int64_t zigzag() {
    uint64_t v = 0; int shift = 0;
    for (;;) {
      if (p >= e) { fail("truncated varint"); return 0; }
      const uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7F) << shift;
      if (!(b & 0x80)) break;
      shift += 7;
      if (shift > 63) { fail("varint too long"); return 0; }
    }
    return (int64_t)(v >> 1) ^ -(int64_t)(v & 1);
  }
  void skip_bytes(uint64_t n) { if (need(n)) p += n; }
};

struct Decoder {
  const Schema& S;
  explicit Decoder(const Schema& s) : S(s) {}
  // follows union branch indices down to a concrete node; returns -1 on error
  int resolve(int n, Cursor& c) const {
    for (int guard = 0; guard < 16 && S.nodes[n].kind == Kind::Union; ++guard) {
      const int64_t b = c.zigzag();
      if (c.err) return -1;
      if (b < 0 || (size_t)b >= S.nodes[n].branches.size()) { c.fail("union branch out of range"); return -1; }
      n = S.nodes[n].branches[(size_t)b];
    }
    return n;
  }
  // a block-encoded array / map: calls item() count times
  template <class F>
  bool blocks(Cursor& c, F&& item) const {
    for (;;) {
      int64_t n = c.zigzag();
      if (c.err) return false;
      if (n == 0) return true;
      if (n == INT64_MIN) return c.fail("block count out of range");
      if (n < 0) { n = -n; (void)c.zigzag(); if (c.err) return false; }
      for (int64_t i = 0; i < n; ++i) if (!item()) return false;
    }
  }
  bool skip(int n, Cursor& c, int depth = 0) const {
    if (depth > 64) return c.fail("value nesting too deep");
    const Node& nd = S.nodes[n];
    switch (nd.kind) {
      case Kind::Null: return true;
      case Kind::Boolean: c.skip_bytes(1); return !c.err;
      case Kind::Int: case Kind::Long: case Kind::Enum: (void)c.zigzag(); return !c.err;
      case Kind::Float: c.skip_bytes(4); return !c.err;
      case Kind::Double: c.skip_bytes(8); return !c.err;
      case Kind::Bytes: case Kind::String: { const int64_t l = c.zigzag(); if (c.err) return false; if (l < 0) return c.fail("negative length"); c.skip_bytes((uint64_t)l); return !c.err; }
      case Kind::Fixed: c.skip_bytes(nd.fixed_size); return !c.err;
      case Kind::Record: for (auto& f : nd.fields) if (!skip(f.second, c, depth + 1)) return false; return true;
      case Kind::Array: return blocks(c, [&]() { return skip(nd.items, c, depth + 1); });
      case Kind::Map: return blocks(c, [&]() {
        const int64_t l = c.zigzag(); if (c.err) return false; if (l < 0) return c.fail("negative length");
        c.skip_bytes((uint64_t)l); return !c.err && skip(nd.items, c, depth + 1); });
      case Kind::Union: { const int r = resolve(n, c); return r >= 0 && skip(r, c, depth + 1); }
    }
    return false;
  }
  bool integer(int n, Cursor& c, int64_t* out) const {
    n = resolve(n, c); if (n < 0) return false;
    const Kind k = S.nodes[n].kind;
    if (k != Kind::Int && k != Kind::Long) return c.fail(k == Kind::Null ? "null where an integer is required" : "integer field has a non-integer type");
    *out = c.zigzag();
    return !c.err;
  }
  bool real(int n, Cursor& c, double* out) const {
    n = resolve(n, c); if (n < 0) return false;
    const Kind k = S.nodes[n].kind;
    if (k == Kind::Double) { if (!c.need(8)) return false; std::memcpy(out, c.p, 8); c.p += 8; return true; }
    if (k == Kind::Float) { if (!c.need(4)) return false; float f; std::memcpy(&f, c.p, 4); c.p += 4; *out = (double)f; return true; }
    if (k == Kind::Int || k == Kind::Long) { *out = (double)c.zigzag(); return !c.err; }
    return c.fail(k == Kind::Null ? "null where a double is required" : "double field has a non-numeric type");
  }
  bool boolean(int n, Cursor& c, bool* out) const {
    n = resolve(n, c); if (n < 0) return false;
    if (S.nodes[n].kind != Kind::Boolean) return c.fail(S.nodes[n].kind == Kind::Null ? "null where a boolean is required" : "boolean field has a non-boolean type");
    if (!c.need(1)) return false;
    *out = *c.p++ != 0;
    return true;
  }
  template <class Push>
  bool int_array(int n, Cursor& c, Push&& push) const {
    n = resolve(n, c); if (n < 0) return false;
    if (S.nodes[n].kind != Kind::Array) return c.fail(S.nodes[n].kind == Kind::Null ? "null where an array is required" : "array field has a non-array type");
    const int item = S.nodes[n].items;
    return blocks(c, [&]() { int64_t v; if (!integer(item, c, &v)) return false; push(v); return true; });
  }
};

// ------------------------------------------------------------------------------------------------ codecs
bool inflate_raw(const uint8_t* src, size_t n, std::vector<uint8_t>* out, std::string* err) {
  z_stream z; std::memset(&z, 0, sizeof z);
  if (n > 0x7FFFFFFFu) { *err = "deflate block larger than 2 GiB"; return false; }
  if (inflateInit2(&z, -15) != Z_OK) { *err = "zlib init failed"; return false; }
  out->resize(std::max<size_t>(n * 4, 1 << 16));
  z.next_in = const_cast<Bytef*>(src); z.avail_in = (uInt)n;
  size_t produced = 0;
  for (;;) {
    if (produced == out->size()) out->resize(out->size() * 2);
    z.next_out = out->data() + produced; z.avail_out = (uInt)std::min<size_t>(out->size() - produced, 1u << 30);
    const size_t before = z.avail_out;
    const int rc = inflate(&z, Z_NO_FLUSH);
    produced += before - z.avail_out;
    if (rc == Z_STREAM_END) break;
    if (rc != Z_OK || (z.avail_in == 0 && z.avail_out != 0)) { inflateEnd(&z); *err = "corrupt deflate block"; return false; }
  }
  inflateEnd(&z);
  out->resize(produced);
  return true;
}

// Snappy raw format: varint uncompressed length, then literal / copy elements.
bool snappy_uncompress(const uint8_t* s, size_t n, std::vector<uint8_t>* out, std::string* err) {
  const uint8_t* e = s + n;
  uint64_t len = 0; int shift = 0;
  for (;;) {
    if (s >= e || shift > 35) { *err = "corrupt snappy block (length)"; return false; }
    const uint8_t b = *s++; len |= (uint64_t)(b & 0x7F) << shift; shift += 7;
    if (!(b & 0x80)) break;
  }
  // a copy element is at least 2 bytes and yields at most 64: a valid stream cannot expand more than 32 x
  if (len > (uint64_t(1) << 32) || len > (uint64_t)n * 32 + 64) { *err = "corrupt snappy block (length)"; return false; }
  out->resize(len);
  uint8_t* d = out->data(); uint8_t* const d0 = d; uint8_t* const de = d + len;
  auto bad = [&]() { *err = "corrupt snappy block"; return false; };
  while (s < e) {
    const uint8_t tag = *s++;
    const uint32_t type = tag & 3u;
    if (type == 0) {
      uint64_t l = tag >> 2;
      if (l >= 60) {
        const uint32_t nb = (uint32_t)l - 59;
        if ((size_t)(e - s) < nb) return bad();
        l = 0; for (uint32_t i = 0; i < nb; ++i) l |= (uint64_t)s[i] << (8 * i);
        s += nb;
      }
      l += 1;
      if ((uint64_t)(e - s) < l || (uint64_t)(de - d) < l) return bad();
      std::memcpy(d, s, l); d += l; s += l;
      continue;
    }
    uint64_t l, off;
    if (type == 1) {
      if (s >= e) return bad();
      l = ((tag >> 2) & 7u) + 4; off = ((uint64_t)(tag >> 5) << 8) | *s++;
    } else if (type == 2) {
      if (e - s < 2) return bad();
      l = (tag >> 2) + 1; off = (uint64_t)s[0] | ((uint64_t)s[1] << 8); s += 2;
    } else {
      if (e - s < 4) return bad();
      l = (tag >> 2) + 1; off = (uint64_t)s[0] | ((uint64_t)s[1] << 8) | ((uint64_t)s[2] << 16) | ((uint64_t)s[3] << 24); s += 4;
    }
    if (off == 0 || off > (uint64_t)(d - d0) || (uint64_t)(de - d) < l) return bad();
    const uint8_t* from = d - off;
    for (uint64_t i = 0; i < l; ++i) d[i] = from[i];                    // overlapping copies repeat the pattern
    d += l;
  }
  if (d != de) return bad();
  return true;
}

// ------------------------------------------------------------------------------------------------ container
struct MappedFile {
  const uint8_t* data = nullptr; size_t size = 0; int fd = -1;
  bool open(const std::string& path, std::string* err) {
    fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) { *err = "cannot open " + path; return false; }
    struct stat st;
    if (fstat(fd, &st) != 0) { *err = "cannot stat " + path; return false; }
    size = (size_t)st.st_size;
    if (size == 0) { *err = path + " is empty"; return false; }
    void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) { *err = "cannot map " + path; return false; }
    data = static_cast<const uint8_t*>(m);
    return true;
  }
  ~MappedFile() { if (data) munmap(const_cast<uint8_t*>(data), size); if (fd >= 0) close(fd); }
};

// Decodes every record of one container file: on_record(decoder, root node, cursor) per record.
template <class F>
bool read_container(const std::string& path, F&& on_record, std::string* err) {
  MappedFile mf;
  if (!mf.open(path, err)) return false;
  Cursor c{mf.data, mf.data + mf.size};
  auto bad = [&](const std::string& m) { *err = path + ": " + m; return false; };
  if (mf.size < 4 || std::memcmp(c.p, "Obj\x01", 4)) return bad("not an Avro object container file");
  c.p += 4;
  std::string schema_json, codec = "null";
  for (;;) {                                                              // file metadata: map<string, bytes>
    int64_t n = c.zigzag();
    if (c.err) return bad(c.err);
    if (n == 0) break;
    if (n == INT64_MIN) return bad("corrupt header");
    if (n < 0) { n = -n; (void)c.zigzag(); }
    for (int64_t i = 0; i < n; ++i) {
      const int64_t kl = c.zigzag(); if (c.err || kl < 0 || !c.need((uint64_t)kl)) return bad("corrupt header");
      std::string key((const char*)c.p, (size_t)kl); c.p += kl;
      const int64_t vl = c.zigzag(); if (c.err || vl < 0 || !c.need((uint64_t)vl)) return bad("corrupt header");
      if (key == "avro.schema") schema_json.assign((const char*)c.p, (size_t)vl);
      else if (key == "avro.codec") codec.assign((const char*)c.p, (size_t)vl);
      c.p += vl;
    }
  }
  if (!c.need(16)) return bad("corrupt header");
  uint8_t sync[16]; std::memcpy(sync, c.p, 16); c.p += 16;
  if (schema_json.empty()) return bad("no avro.schema in the header");
  Json js; JsonParser jp{schema_json.data(), schema_json.data() + schema_json.size(), {}};
  if (!jp.value(js)) return bad("avro.schema: " + jp.err);
  Schema S;
  const int root = S.parse(js, "");
  if (root < 0) return bad("avro.schema: " + S.err);
  if (S.nodes[root].kind != Kind::Record) return bad("top-level schema is not a record");
  const bool is_null = codec == "null" || codec.empty(), is_deflate = codec == "deflate", is_snappy = codec == "snappy";
  if (!is_null && !is_deflate && !is_snappy) return bad("unsupported avro.codec '" + codec + "' (null, deflate and snappy are implemented)");
  Decoder D(S);
  std::vector<uint8_t> raw;
  while (c.p < c.e) {
    const int64_t count = c.zigzag(); const int64_t bytes = c.zigzag();
    if (c.err || count < 0 || bytes < 0 || !c.need((uint64_t)bytes + 16)) return bad("corrupt block header");
    Cursor b{c.p, c.p + bytes};
    if (is_deflate) {
      std::string e2;
      if (!inflate_raw(c.p, (size_t)bytes, &raw, &e2)) return bad(e2);
      b = Cursor{raw.data(), raw.data() + raw.size()};
    } else if (is_snappy) {
      if (bytes < 4) return bad("corrupt snappy block");
      std::string e2;
      if (!snappy_uncompress(c.p, (size_t)bytes - 4, &raw, &e2)) return bad(e2);
      const uint8_t* t = c.p + bytes - 4;
      const uint32_t want = ((uint32_t)t[0] << 24) | ((uint32_t)t[1] << 16) | ((uint32_t)t[2] << 8) | t[3];
      if ((uint32_t)crc32(0L, raw.data(), (uInt)raw.size()) != want) return bad("snappy block CRC mismatch");
      b = Cursor{raw.data(), raw.data() + raw.size()};
    }
    for (int64_t r = 0; r < count; ++r) {
      std::string e2;
      if (!on_record(D, S.nodes[root], b, &e2)) return bad(e2.empty() ? std::string(b.err ? b.err : "record does not deserialize") : e2);
    }
    if (b.p != b.e) return bad("trailing bytes in a block");
    c.p += bytes;
    if (std::memcmp(c.p, sync, 16)) return bad("sync marker mismatch");
    c.p += 16;
  }
  return true;
}

bool list_avro_files(const std::string& dir, std::vector<std::string>* out, std::string* err) {
  DIR* d = opendir(dir.c_str());
  if (!d) { *err = "cannot read directory " + dir; return false; }
  while (dirent* ent = readdir(d)) {
    const std::string n = ent->d_name;
    if (n.size() > 5 && n.compare(n.size() - 5, 5, ".avro") == 0) out->push_back(dir + "/" + n);   // :209, :271
  }
  closedir(d);
  std::sort(out->begin(), out->end());
  return true;
}

struct ItemChunk { std::vector<uint64_t> ids, off{0}; std::vector<uint32_t> sessions; std::vector<double> idf; std::vector<uint8_t> attr; };
struct SessionChunk { std::vector<uint32_t> index, ts; std::vector<uint64_t> off{0}, items; };

// struct ItemIdexAvroSchema { ItemId: i64, session_indices_time_ordered: Vec<i32>, idf: f64, ForSale: bool, IsAdult: bool } (:184-192)
bool read_item_file(const std::string& path, ItemChunk* ch, std::string* err) {
  return read_container(path, [&](const Decoder& D, const Node& rec, Cursor& c, std::string* e2) {
    int64_t id = 0; double idf = 0; bool for_sale = false, adult = false; unsigned seen = 0;
    for (auto& f : rec.fields) {
      bool ok;
      if (f.first == "ItemId") { ok = D.integer(f.second, c, &id); seen |= 1; }
      else if (f.first == "session_indices_time_ordered") {
        ok = D.int_array(f.second, c, [&](int64_t v) { ch->sessions.push_back((uint32_t)(int32_t)v); });   // `*x as u32` (:218)
        seen |= 2;
      }
      else if (f.first == "idf") { ok = D.real(f.second, c, &idf); seen |= 4; }
      else if (f.first == "ForSale") { ok = D.boolean(f.second, c, &for_sale); seen |= 8; }
      else if (f.first == "IsAdult") { ok = D.boolean(f.second, c, &adult); seen |= 16; }
      else ok = D.skip(f.second, c);
      if (!ok) { *e2 = "field '" + f.first + "': " + (c.err ? c.err : "decode error"); return false; }
    }
    if (seen != 31) { *e2 = "itemindex record lacks one of ItemId / session_indices_time_ordered / idf / ForSale / IsAdult"; return false; }
    ch->ids.push_back((uint64_t)id);                                       // `ItemId as u64` (:222)
    ch->off.push_back(ch->sessions.size());
    ch->idf.push_back(idf);
    ch->attr.push_back((uint8_t)(VMIS_ATTR_EXISTS | (for_sale ? VMIS_ATTR_FOR_SALE : 0) | (adult ? VMIS_ATTR_ADULT : 0)));
    return true;
  }, err);
}

// struct SessionIdexAvroSchema { SessionIndex: i32, item_ids_asc: Vec<i64>, Time: i32 } (:249-255)
bool read_session_file(const std::string& path, SessionChunk* ch, std::string* err) {
  return read_container(path, [&](const Decoder& D, const Node& rec, Cursor& c, std::string* e2) {
    int64_t idx = 0, time = 0; unsigned seen = 0;
    for (auto& f : rec.fields) {
      bool ok;
      if (f.first == "SessionIndex") { ok = D.integer(f.second, c, &idx); seen |= 1; }
      else if (f.first == "item_ids_asc") { ok = D.int_array(f.second, c, [&](int64_t v) { ch->items.push_back((uint64_t)v); }); seen |= 2; }
      else if (f.first == "Time") { ok = D.integer(f.second, c, &time); seen |= 4; }
      else ok = D.skip(f.second, c);
      if (!ok) { *e2 = "field '" + f.first + "': " + (c.err ? c.err : "decode error"); return false; }
    }
    if (seen != 7) { *e2 = "sessionindex record lacks one of SessionIndex / item_ids_asc / Time"; return false; }
    if (idx < 0 || idx >= 0x7FFFFFFFll) { *e2 = "SessionIndex out of range"; return false; }   // `as usize` of a negative i32 would not fit memory
    ch->index.push_back((uint32_t)idx);
    ch->ts.push_back((uint32_t)(int32_t)time);                             // `Time as u32` (:290)
    ch->off.push_back(ch->items.size());
    return true;
  }, err);
}

template <class Chunk, class F>
bool read_all(const std::vector<std::string>& files, std::vector<Chunk>* chunks, F&& read_one, std::string* err) {
  chunks->resize(files.size());
  std::vector<std::string> errs(files.size());
  std::atomic<size_t> next{0};
  auto work = [&]() {
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= files.size()) break;
      read_one(files[i], &(*chunks)[i], &errs[i]);
    }
  };
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const size_t nt = std::min<size_t>(files.size(), hw);
  std::vector<std::thread> th;
  for (size_t t = 1; t < nt; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  for (auto& e : errs) if (!e.empty()) { *err = e; return false; }
  return true;
}

}  // namespace

bool read_index_from_avro(const std::string& base_path, PrebuiltIndex* out, AvroLoadInfo* info, std::string* err) {
  *out = PrebuiltIndex();
  std::vector<std::string> item_files, session_files;
  if (!list_avro_files(base_path + "/itemindex", &item_files, err)) return false;          // :92
  if (!list_avro_files(base_path + "/sessionindex", &session_files, err)) return false;    // :99
  if (item_files.empty()) { *err = "no .avro files under " + base_path + "/itemindex"; return false; }
  if (session_files.empty()) { *err = "no .avro files under " + base_path + "/sessionindex"; return false; }
  std::vector<ItemChunk> ic; std::vector<SessionChunk> sc;
  if (!read_all(item_files, &ic, read_item_file, err)) return false;
  if (!read_all(session_files, &sc, read_session_file, err)) return false;

  // items: HashMap::insert semantics (:214-226) — the last record of an ItemId wins
  size_t n_rec = 0, n_post = 0;
  for (auto& c : ic) { n_rec += c.ids.size(); n_post += c.sessions.size(); }
  {
    std::vector<std::pair<uint64_t, uint64_t>> key; key.reserve(n_rec);      // (ItemId, global record number)
    uint64_t g = 0;
    for (auto& c : ic) for (uint64_t id : c.ids) key.emplace_back(id, g++);
    std::sort(key.begin(), key.end());
    std::vector<uint8_t> live(n_rec, 0);
    for (size_t i = 0; i < key.size(); ++i) if (i + 1 == key.size() || key[i + 1].first != key[i].first) live[key[i].second] = 1;
    out->item_ids.reserve(n_rec); out->post_off.assign(1, 0); out->post_sessions.reserve(n_post);
    g = 0;
    for (auto& c : ic) {
      for (size_t r = 0; r < c.ids.size(); ++r, ++g) {
        if (!live[g]) continue;
        out->item_ids.push_back(c.ids[r]); out->idf.push_back(c.idf[r]); out->attr.push_back(c.attr[r]);
        out->post_sessions.insert(out->post_sessions.end(), c.sessions.begin() + c.off[r], c.sessions.begin() + c.off[r + 1]);
        out->post_off.push_back(out->post_sessions.size());
      }
      std::vector<uint64_t>().swap(c.ids); std::vector<uint32_t>().swap(c.sessions);
    }
  }
  // sessions: dense vectors indexed by SessionIndex, truncated after the largest index used (:256-303)
  size_t n_srec = 0; uint32_t max_idx = 0;
  for (auto& c : sc) { n_srec += c.index.size(); for (uint32_t i : c.index) max_idx = std::max(max_idx, i); }
  if (n_srec == 0) { *err = "no session records under " + base_path + "/sessionindex"; return false; }
  const size_t S = (size_t)max_idx + 1;
  Sessions& ss = out->sessions;
  std::vector<uint64_t> len(S, 0);
  for (auto& c : sc) for (size_t r = 0; r < c.index.size(); ++r) len[c.index[r]] = c.off[r + 1] - c.off[r];   // last record wins
  ss.off.assign(S + 1, 0);
  for (size_t s = 0; s < S; ++s) ss.off[s + 1] = ss.off[s] + len[s];
  ss.items.assign(ss.off[S], 0); ss.ts.assign(S, 0);
  for (auto& c : sc) for (size_t r = 0; r < c.index.size(); ++r) {
    const uint32_t s = c.index[r];
    if (c.off[r + 1] - c.off[r] != len[s]) continue;                       // an earlier duplicate of another length; the last record always fits and is copied last
    std::copy(c.items.begin() + c.off[r], c.items.begin() + c.off[r + 1], ss.items.begin() + ss.off[s]);
    ss.ts[s] = c.ts[r];
  }
  if (info) { info->item_files = item_files.size(); info->session_files = session_files.size(); info->item_records = n_rec; info->session_records = n_srec; }
  return true;
}

// ------------------------------------------------------------------------------------------------ writer
namespace {

struct Out {
  std::vector<uint8_t> b;
  void zigzag(int64_t v) {
    uint64_t u = ((uint64_t)v << 1) ^ (uint64_t)(v >> 63);
    while (u >= 0x80) { b.push_back((uint8_t)(u | 0x80)); u >>= 7; }
    b.push_back((uint8_t)u);
  }
  void bytes(const void* p, size_t n) { const uint8_t* c = static_cast<const uint8_t*>(p); b.insert(b.end(), c, c + n); }
  void str(const std::string& s) { zigzag((int64_t)s.size()); bytes(s.data(), s.size()); }
};

const char* kItemSchema =
    "{\"type\":\"record\",\"name\":\"ItemIndex\",\"fields\":[{\"name\":\"ItemId\",\"type\":\"long\"},"
    "{\"name\":\"session_indices_time_ordered\",\"type\":{\"type\":\"array\",\"items\":\"int\"}},"
    "{\"name\":\"idf\",\"type\":\"double\"},{\"name\":\"ForSale\",\"type\":\"boolean\"},"
    "{\"name\":\"IsAdult\",\"type\":\"boolean\"}]}";
const char* kSessionSchema =
    "{\"type\":\"record\",\"name\":\"SessionIndex\",\"fields\":[{\"name\":\"SessionIndex\",\"type\":\"int\"},"
    "{\"name\":\"item_ids_asc\",\"type\":{\"type\":\"array\",\"items\":\"long\"}},"
    "{\"name\":\"Time\",\"type\":\"int\"}]}";

// One container file: records [lo, hi) produced by encode(i, out), blocks of ~1 MiB of encoded records.
template <class F>
bool write_container_file(const std::string& path, const char* schema, bool use_deflate, size_t lo, size_t hi, uint64_t seed,
                          F&& encode, std::string* err) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) { *err = "cannot create " + path; return false; }
  uint8_t sync[16];
  for (int i = 0; i < 16; ++i) { seed = seed * 6364136223846793005ull + 1442695040888963407ull; sync[i] = (uint8_t)(seed >> 56); }
  Out h;
  h.bytes("Obj\x01", 4);
  h.zigzag(2);
  h.str("avro.schema"); h.str(schema);
  h.str("avro.codec"); h.str(use_deflate ? "deflate" : "null");
  h.zigzag(0);
  h.bytes(sync, 16);
  bool ok = std::fwrite(h.b.data(), 1, h.b.size(), f) == h.b.size();
  Out blk; std::vector<uint8_t> comp;
  size_t in_block = 0;
  auto flush = [&]() {
    if (!in_block) return;
    const std::vector<uint8_t>* payload = &blk.b;
    if (use_deflate) {
      z_stream z; std::memset(&z, 0, sizeof z);
      if (deflateInit2(&z, Z_BEST_SPEED, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { ok = false; return; }
      comp.resize(deflateBound(&z, (uLong)blk.b.size()));
      z.next_in = blk.b.data(); z.avail_in = (uInt)blk.b.size();
      z.next_out = comp.data(); z.avail_out = (uInt)comp.size();
      if (deflate(&z, Z_FINISH) != Z_STREAM_END) ok = false;
      comp.resize(z.total_out);
      deflateEnd(&z);
      payload = &comp;
    }
    Out hd; hd.zigzag((int64_t)in_block); hd.zigzag((int64_t)payload->size());
    ok = ok && std::fwrite(hd.b.data(), 1, hd.b.size(), f) == hd.b.size() &&
         std::fwrite(payload->data(), 1, payload->size(), f) == payload->size() && std::fwrite(sync, 1, 16, f) == 16;
    blk.b.clear(); in_block = 0;
  };
  for (size_t i = lo; i < hi && ok; ++i) {
    if (encode(i, blk)) ++in_block;
    if (blk.b.size() >= (size_t(1) << 20)) flush();
  }
  flush();
  ok = (std::fclose(f) == 0) && ok;
  if (!ok) *err = "write to " + path + " failed";
  return ok;
}

}  // namespace

bool write_index_to_avro(const std::string& base_path, const PrebuiltIndex& p, const std::string& codec, size_t n_files,
                         std::string* err) {
  const bool deflate = codec == "deflate";
  if (!deflate && codec != "null") { *err = "codec must be \"null\" or \"deflate\""; return false; }
  if (n_files == 0) n_files = 1;
  for (const char* sub : {"", "/itemindex", "/sessionindex"}) {
    const std::string d = base_path + sub;
    if (mkdir(d.c_str(), 0777) != 0 && errno != EEXIST) { *err = "cannot create directory " + d; return false; }
  }
  const size_t I = p.item_ids.size(), S = p.sessions.size();
  std::vector<std::string> errs(2 * n_files);
  std::vector<std::thread> th;
  for (size_t fidx = 0; fidx < n_files; ++fidx) {
    th.emplace_back([&, fidx]() {
      char name[64]; std::snprintf(name, sizeof name, "/part-%05zu.avro", fidx);
      write_container_file(base_path + "/itemindex" + name, kItemSchema, deflate, I * fidx / n_files, I * (fidx + 1) / n_files,
                           0x1234 + fidx, [&](size_t i, Out& o) {
        o.zigzag((int64_t)p.item_ids[i]);                                   // ItemId: long (`as u64` on the way back)
        const uint64_t a = p.post_off[i], b = p.post_off[i + 1];
        if (b > a) { o.zigzag((int64_t)(b - a)); for (uint64_t e = a; e < b; ++e) o.zigzag((int64_t)(int32_t)p.post_sessions[e]); }
        o.zigzag(0);
        o.bytes(&p.idf[i], 8);
        o.b.push_back((p.attr[i] & VMIS_ATTR_FOR_SALE) ? 1 : 0);
        o.b.push_back((p.attr[i] & VMIS_ATTR_ADULT) ? 1 : 0);
        return true;
      }, &errs[2 * fidx]);
      write_container_file(base_path + "/sessionindex" + name, kSessionSchema, deflate, S * fidx / n_files,
                           S * (fidx + 1) / n_files, 0x9876 + fidx, [&](size_t s, Out& o) {
        const uint64_t a = p.sessions.off[s], b = p.sessions.off[s + 1];
        if (b == a) return false;                                           // unused session index
        o.zigzag((int64_t)s);
        o.zigzag((int64_t)(b - a));
        for (uint64_t e = a; e < b; ++e) o.zigzag((int64_t)p.sessions.items[e]);
        o.zigzag(0);
        o.zigzag((int64_t)(int32_t)p.sessions.ts[s]);
        return true;
      }, &errs[2 * fidx + 1]);
    });
  }
  for (auto& t : th) t.join();
  for (auto& e : errs) if (!e.empty()) { *err = e; return false; }
  return true;
}

}  // namespace vmis
