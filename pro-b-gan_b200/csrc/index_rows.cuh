// Host-side ingest of the CLI's index arrays (no device code).
//
// The reference decodes --input_triplets / --input_pairs / --input_entities with json.loads into nested Python lists
// (pro_b_gan_infer.py:485, :493, :501) and then walks them again with torch.tensor(...) (:135-136, :182, :226); at
// B = 32768 triplets that is ~70 ms of interpreter time in front of a 0.18 ms pass.  parse_index_rows reads the same
// JSON text once, straight into the int64 [rows, cols] buffer the kernels gather from.
#pragma once
#include <cstddef>
#include <cstdint>

namespace pbg_host {

struct RowParser {
  const char* p; const char* end; const char* begin;
  const char* what = nullptr;
  void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
  bool eat(char c) { ws(); if (p < end && *p == c) { ++p; return true; } return false; }
  bool fail(const char* w) { if (!what) what = w; return false; }
  // JSON integer: -? (0 | [1-9][0-9]*), nothing fractional / exponential, must fit int64
  bool integer(int64_t* out) {
    ws();
    bool neg = false;
    if (p < end && *p == '-') { neg = true; ++p; }
    if (p >= end || *p < '0' || *p > '9') return fail("expected an integer");
    if (*p == '0' && p + 1 < end && p[1] >= '0' && p[1] <= '9') return fail("leading zero in integer");
    uint64_t v = 0;
    const uint64_t lim = neg ? (uint64_t)1 << 63 : ((uint64_t)1 << 63) - 1;
    while (p < end && *p >= '0' && *p <= '9') {
      const uint64_t d = (uint64_t)(*p - '0');
      if (v > (lim - d) / 10) return fail("integer does not fit int64");
      v = v * 10 + d;
      ++p;
    }
    if (p < end && (*p == '.' || *p == 'e' || *p == 'E')) return fail("ids must be integers, found a float");
    *out = neg ? (int64_t)(0 - v) : (int64_t)v;
    return true;
  }
};

// cols == 1: "[a, b, ...]"      -> rows of one id        (--input_entities)
// cols >= 2: "[[a, b(, c)], ...]" -> rows of `cols` ids    (--input_pairs: 2, --input_triplets: 3)
// Returns the number of rows (out may be null to count only; rows beyond cap_rows are counted, not written),
// or -1 with *err_msg / *err_off set.
inline long long parse_index_rows(const char* text, size_t len, int cols, int64_t* out, size_t cap_rows,
                                  const char** err_msg, size_t* err_off) {
  RowParser ps{text, text + len, text};
  long long rows = 0;
  auto bad = [&](const char* w) { ps.fail(w); *err_msg = ps.what; *err_off = (size_t)(ps.p - ps.begin); return -1LL; };
  if (!ps.eat('[')) return bad("expected '['");
  if (!ps.eat(']')) {
    for (;;) {
      int64_t v[8];
      if (cols == 1) {
        if (!ps.integer(&v[0])) return bad("expected an integer");
      } else {
        if (!ps.eat('[')) return bad("expected '[' opening a row");
        for (int c = 0; c < cols; ++c) {
          if (c && !ps.eat(',')) return bad("row has too few ids");
          if (!ps.integer(&v[c])) return bad("expected an integer");
        }
        if (!ps.eat(']')) return bad("row has too many ids or is not closed");
      }
      if (out && (size_t)rows < cap_rows)
        for (int c = 0; c < cols; ++c) out[(size_t)rows * cols + c] = v[c];
      ++rows;
      if (ps.eat(',')) continue;
      if (ps.eat(']')) break;
      return bad("expected ',' or ']'");
    }
  }
  ps.ws();
  if (ps.p != ps.end) return bad("trailing characters after the array");
  return rows;
}

}  // namespace pbg_host
