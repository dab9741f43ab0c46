// hgef_io.cu -- MatrixMarket (.mtx) incidence reader, host side.
//
// The reference's standalone binaries load hypergraph incidence matrices from MatrixMarket coordinate files
// (include/dataloader/dataloader.hpp:22-104, mmio banner parsing): values are dropped, indices become 0-based,
// a `symmetric` matrix is mirrored and de-duplicated, and the coordinates are sorted row-major.  Same
// semantics here, as plain C-ABI calls (open -> sizes -> fill -> close), int64 coordinates ready for
// hg_csr_build_*; rows = vertices, columns = hyperedges.  Errors are return codes, never exit().
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "hgef_common.cuh"

struct hgMtx {
  int64_t nrow = 0, ncol = 0;
  std::vector<std::pair<int64_t, int64_t>> coords;   // sorted (row, col)
};

namespace hg {
namespace {

std::string lower(std::string s) {
  for (char &c : s) c = (char)std::tolower((unsigned char)c);
  return s;
}

// next whitespace-delimited token of [p, end); returns false at the end of the buffer
bool next_token(const char *&p, const char *end, const char *&tok, size_t &len) {
  while (p < end && std::isspace((unsigned char)*p)) ++p;
  if (p >= end) return false;
  tok = p;
  while (p < end && !std::isspace((unsigned char)*p)) ++p;
  len = (size_t)(p - tok);
  return true;
}

bool parse_i64(const char *tok, size_t len, int64_t &out) {
  if (len == 0 || len > 20) return false;
  char buf[24];
  std::memcpy(buf, tok, len);
  buf[len] = 0;
  char *e = nullptr;
  const long long v = std::strtoll(buf, &e, 10);
  if (e == buf || *e != 0) return false;
  out = v;
  return true;
}

}  // namespace
}  // namespace hg

using namespace hg;

extern "C" {

int hg_mtx_open(const char *path, hgMtx **out, int64_t *nrow, int64_t *ncol, int64_t *nnz) {
  HG_REQUIRE(path != nullptr && out != nullptr, "mtx_open: NULL argument");
  *out = nullptr;
  FILE *f = std::fopen(path, "rb");
  if (!f) return set_error(HG_EINVAL, "mtx_open: cannot open %s", path);
  std::fseek(f, 0, SEEK_END);
  const long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::string text((size_t)(sz > 0 ? sz : 0), '\0');
  const size_t got = sz > 0 ? std::fread(&text[0], 1, (size_t)sz, f) : 0;
  std::fclose(f);
  if (got != text.size()) return set_error(HG_EINVAL, "mtx_open: short read of %s", path);

  // banner: %%MatrixMarket matrix coordinate <field> <symmetry>
  const size_t eol = text.find('\n');
  const std::string banner = lower(text.substr(0, eol == std::string::npos ? text.size() : eol));
  char b0[64] = {0}, b1[64] = {0}, b2[64] = {0}, field[64] = {0}, symm[64] = {0};
  if (std::sscanf(banner.c_str(), "%63s %63s %63s %63s %63s", b0, b1, b2, field, symm) != 5 ||
      std::strcmp(b0, "%%matrixmarket") != 0 || std::strcmp(b1, "matrix") != 0)
    return set_error(HG_EINVAL, "mtx_open: %s has no MatrixMarket banner", path);
  if (std::strcmp(b2, "coordinate") != 0)
    return set_error(HG_EINVAL, "mtx_open: %s is not in coordinate format (got '%s')", path, b2);
  const bool has_value = !std::strcmp(field, "real") || !std::strcmp(field, "integer") || !std::strcmp(field, "double");
  const bool two_values = !std::strcmp(field, "complex");
  if (!has_value && !two_values && std::strcmp(field, "pattern") != 0)
    return set_error(HG_EINVAL, "mtx_open: unknown field '%s' in %s", field, path);
  const bool symmetric = !std::strcmp(symm, "symmetric") || !std::strcmp(symm, "skew-symmetric") || !std::strcmp(symm, "hermitian");
  if (!symmetric && std::strcmp(symm, "general") != 0)
    return set_error(HG_EINVAL, "mtx_open: unknown symmetry '%s' in %s", symm, path);

  // skip comment lines, then the size line
  const char *p = text.data() + (eol == std::string::npos ? text.size() : eol + 1), *end = text.data() + text.size();
  while (p < end) {
    const char *q = p;
    while (q < end && (*q == ' ' || *q == '\t' || *q == '\r')) ++q;
    if (q < end && *q == '%') { while (p < end && *p != '\n') ++p; if (p < end) ++p; continue; }
    if (q < end && *q == '\n') { p = q + 1; continue; }
    break;
  }
  const char *tok;
  size_t len;
  int64_t dims[3];
  for (int i = 0; i < 3; ++i)
    if (!next_token(p, end, tok, len) || !parse_i64(tok, len, dims[i]) || dims[i] < 0)
      return set_error(HG_EINVAL, "mtx_open: %s has no 'rows cols entries' line", path);
  hgMtx *m = new (std::nothrow) hgMtx();
  if (!m) return set_error(HG_ENOMEM, "mtx_open: out of host memory");
  m->nrow = dims[0];
  m->ncol = dims[1];
  m->coords.reserve((size_t)dims[2] * (symmetric ? 2 : 1));
  for (int64_t i = 0; i < dims[2]; ++i) {
    int64_t r, c;
    if (!next_token(p, end, tok, len) || !parse_i64(tok, len, r) || !next_token(p, end, tok, len) || !parse_i64(tok, len, c)) {
      delete m;
      return set_error(HG_EINVAL, "mtx_open: %s ends after %lld of %lld entries", path, (long long)i, (long long)dims[2]);
    }
    for (int k = 0; k < (two_values ? 2 : (has_value ? 1 : 0)); ++k)
      if (!next_token(p, end, tok, len)) {
        delete m;
        return set_error(HG_EINVAL, "mtx_open: entry %lld of %s has no value", (long long)i, path);
      }
    if (r < 1 || r > m->nrow || c < 1 || c > m->ncol) {   // (the reference stores such entries unchecked)
      delete m;
      return set_error(HG_EGRAPH, "mtx_open: entry %lld of %s is (%lld, %lld), outside %lld x %lld", (long long)i, path,
                       (long long)r, (long long)c, (long long)dims[0], (long long)dims[1]);
    }
    m->coords.emplace_back(r - 1, c - 1);                 // MatrixMarket is 1-based
    if (symmetric && r != c) m->coords.emplace_back(c - 1, r - 1);
  }
  std::sort(m->coords.begin(), m->coords.end());
  if (symmetric) m->coords.erase(std::unique(m->coords.begin(), m->coords.end()), m->coords.end());
  if (nrow) *nrow = m->nrow;
  if (ncol) *ncol = m->ncol;
  if (nnz) *nnz = (int64_t)m->coords.size();
  *out = m;
  return HG_OK;
}

int hg_mtx_fill(const hgMtx *m, int64_t *rows, int64_t *cols) {
  HG_REQUIRE(m != nullptr && (m->coords.empty() || (rows != nullptr && cols != nullptr)), "mtx_fill: NULL argument");
  for (size_t i = 0; i < m->coords.size(); ++i) {
    rows[i] = m->coords[i].first;
    cols[i] = m->coords[i].second;
  }
  return HG_OK;
}

int hg_mtx_close(hgMtx *m) {
  delete m;
  return HG_OK;
}

}  // extern "C"
