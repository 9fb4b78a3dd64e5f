// Host-side result formatting (no device code): the reference ends every CLI task with json.dump(results, indent=2)
// (pro_b_gan_infer.py:503-508) over Python lists made by .tolist() (:154, :162, :203, :208-209).  At 32768 triplets
// that is ~60 ms of interpreter time behind a 0.18 ms pass.  These writers produce the same bytes json.dumps would --
// floats as the shortest text that round-trips the double an fp32 value converts to (Python's float repr), lists one
// item per line at the given indentation -- straight from the result arrays.
#pragma once
#include <charconv>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace pbg_host {

struct JsonOut {
  char* p; char* end; size_t need = 0;   // `need` counts every byte, written or not (out may be too small / null)
  void put(char c) { ++need; if (p && p < end) *p++ = c; }
  void put(const char* s, size_t n) { need += n; if (p && p + n <= end) { memcpy(p, s, n); p += n; } else if (p) p = end; }
  void newline(int spaces) { put('\n'); for (int i = 0; i < spaces; ++i) put(' '); }
};

// repr(float(x)) for x an fp32 value: shortest round-trip digits of the double, Python's layout rules
// (float_repr_style 'short', format 'r': exponent form iff decpt <= -4 or decpt > 16; "-0.0"; NaN / Infinity as json.dumps).
inline void put_py_float(JsonOut& o, float f) {
  const double d = static_cast<double>(f);
  if (std::isnan(d)) { o.put("NaN", 3); return; }
  if (std::isinf(d)) { if (d < 0) o.put('-'); o.put("Infinity", 8); return; }
  char buf[40];
  const auto r = std::to_chars(buf, buf + sizeof buf, d, std::chars_format::scientific);   // [-]d[.ddd]e[+-]XX
  const char* s = buf;
  if (*s == '-') { o.put('-'); ++s; }
  const char* e = s;
  while (e < r.ptr && *e != 'e') ++e;
  char digits[24]; int nd = 0;
  for (const char* q = s; q < e; ++q) if (*q != '.') digits[nd++] = *q;
  int x = 0; { const char* q = e + 1; const bool neg = *q == '-'; if (*q == '-' || *q == '+') ++q; for (; q < r.ptr; ++q) x = x * 10 + (*q - '0'); if (neg) x = -x; }
  if (nd == 1 && digits[0] == '0') { o.put("0.0", 3); return; }
  const int decpt = x + 1;                         // value = 0.d1d2... x 10^decpt
  if (decpt <= -4 || decpt > 16) {                 // d[.ddd]e[+-]XX, at least two exponent digits, no ".0"
    o.put(digits[0]);
    if (nd > 1) { o.put('.'); o.put(digits + 1, static_cast<size_t>(nd - 1)); }
    o.put('e'); o.put(x < 0 ? '-' : '+');
    const int ax = x < 0 ? -x : x;
    char eb[8]; int ne = 0; int t = ax; do { eb[ne++] = static_cast<char>('0' + t % 10); t /= 10; } while (t);
    if (ne < 2) eb[ne++] = '0';
    while (ne) o.put(eb[--ne]);
  } else if (decpt <= 0) {                         // 0.000ddd
    o.put("0.", 2);
    for (int i = 0; i < -decpt; ++i) o.put('0');
    o.put(digits, static_cast<size_t>(nd));
  } else if (decpt >= nd) {                        // ddd000.0
    o.put(digits, static_cast<size_t>(nd));
    for (int i = nd; i < decpt; ++i) o.put('0');
    o.put(".0", 2);
  } else {                                         // dd.ddd
    o.put(digits, static_cast<size_t>(decpt));
    o.put('.');
    o.put(digits + decpt, static_cast<size_t>(nd - decpt));
  }
}

inline void put_i64(JsonOut& o, int64_t v) {
  char buf[24];
  const auto r = std::to_chars(buf, buf + sizeof buf, v);
  o.put(buf, static_cast<size_t>(r.ptr - buf));
}

// json.dumps(list, indent=indent) of a [rows, cols] array placed at nesting depth `depth` (the closing bracket is
// indented by indent * depth spaces).  cols == 0: a flat list of `rows` scalars.  indent < 0: the compact
// json.dumps(list) form (", " separators).  T = float or int64_t.
template <typename T, typename PutFn>
inline size_t format_rows(const T* v, size_t rows, int cols, int indent, int depth, char* out, size_t cap, PutFn put_one) {
  JsonOut o{out, out ? out + cap : nullptr};
  const bool pretty = indent >= 0;
  auto list = [&](const T* a, size_t n, int d) {
    if (n == 0) { o.put("[]", 2); return; }
    o.put('[');
    for (size_t i = 0; i < n; ++i) {
      if (i) { o.put(','); if (!pretty) o.put(' '); }
      if (pretty) o.newline(indent * (d + 1));
      put_one(o, a[i]);
    }
    if (pretty) o.newline(indent * d);
    o.put(']');
  };
  if (cols == 0) {
    list(v, rows, depth);
  } else if (rows == 0) {
    o.put("[]", 2);
  } else {
    o.put('[');
    for (size_t r = 0; r < rows; ++r) {
      if (r) { o.put(','); if (!pretty) o.put(' '); }
      if (pretty) o.newline(indent * (depth + 1));
      list(v + r * static_cast<size_t>(cols), static_cast<size_t>(cols), depth + 1);
    }
    if (pretty) o.newline(indent * depth);
    o.put(']');
  }
  return o.need;
}

}  // namespace pbg_host
