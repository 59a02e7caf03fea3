// Native TIFF / BigTIFF raster I/O (see tiff_io.h for the reference contract).
#include "tiff_io.h"

#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <atomic>
#include <cstring>
#include <mutex>
#include <thread>

namespace tdio {
namespace {

const int kTypeSize[19] = {0, 1, 1, 2, 4, 8, 1, 1, 2, 4, 8, 4, 8, 0, 0, 0, 8, 8, 8};

inline uint16_t bswap16(uint16_t v) { return (uint16_t)((v >> 8) | (v << 8)); }
inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }
inline uint64_t bswap64(uint64_t v) { return __builtin_bswap64(v); }

void swap_elems(uint8_t* p, size_t n, int sz) {
  if (sz == 2) { for (size_t i = 0; i < n; i++) { uint16_t v; memcpy(&v, p + 2 * i, 2); v = bswap16(v); memcpy(p + 2 * i, &v, 2); } }
  else if (sz == 4) { for (size_t i = 0; i < n; i++) { uint32_t v; memcpy(&v, p + 4 * i, 4); v = bswap32(v); memcpy(p + 4 * i, &v, 4); } }
  else if (sz == 8) { for (size_t i = 0; i < n; i++) { uint64_t v; memcpy(&v, p + 8 * i, 8); v = bswap64(v); memcpy(p + 8 * i, &v, 8); } }
}

// positional read on the descriptor: safe from several decoder threads at once
bool pread_all(FILE* fp, uint64_t off, void* buf, size_t n) {
  const int fd = fileno(fp);
  char* c = (char*)buf;
  while (n) {
    const ssize_t k = pread(fd, c, n, (off_t)off);
    if (k <= 0) return false;
    c += k; off += (uint64_t)k; n -= (size_t)k;
  }
  return true;
}

double tag_double(const RawTag& t, size_t i) {
  const uint8_t* p = t.data.data();
  switch (t.type) {
    case 1: case 7: return p[i];
    case 6: return (int8_t)p[i];
    case 3: { uint16_t v; memcpy(&v, p + 2 * i, 2); return v; }
    case 8: { int16_t v; memcpy(&v, p + 2 * i, 2); return v; }
    case 4: { uint32_t v; memcpy(&v, p + 4 * i, 4); return v; }
    case 9: { int32_t v; memcpy(&v, p + 4 * i, 4); return v; }
    case 11: { float v; memcpy(&v, p + 4 * i, 4); return v; }
    case 12: { double v; memcpy(&v, p + 8 * i, 8); return v; }
    case 16: case 18: { uint64_t v; memcpy(&v, p + 8 * i, 8); return (double)v; }
    case 17: { int64_t v; memcpy(&v, p + 8 * i, 8); return (double)v; }
    default: return 0;
  }
}
uint64_t tag_u64(const RawTag& t, size_t i) {
  const uint8_t* p = t.data.data();
  switch (t.type) {
    case 1: case 7: return p[i];
    case 3: { uint16_t v; memcpy(&v, p + 2 * i, 2); return v; }
    case 4: { uint32_t v; memcpy(&v, p + 4 * i, 4); return v; }
    case 16: case 18: { uint64_t v; memcpy(&v, p + 8 * i, 8); return v; }
    default: return (uint64_t)tag_double(t, i);
  }
}

// GDALCopyWords-style narrowing: round half away from zero, clamp, NaN -> 0.
template <typename T>
inline T narrow_from_double(double v, double lo, double hi) {
  if (std::isnan(v)) return 0;
  v = v >= 0 ? std::floor(v + 0.5) : std::ceil(v - 0.5);
  if (v < lo) v = lo;
  if (v > hi) v = hi;
  return (T)v;
}

inline double sample_as_double(const uint8_t* p, int bits, int fmt) {
  switch (bits) {
    case 8: return fmt == 2 ? (double)(int8_t)p[0] : (double)p[0];
    case 16: { if (fmt == 2) { int16_t v; memcpy(&v, p, 2); return v; } uint16_t v; memcpy(&v, p, 2); return v; }
    case 32:
      if (fmt == 3) { float v; memcpy(&v, p, 4); return v; }
      if (fmt == 2) { int32_t v; memcpy(&v, p, 4); return v; }
      { uint32_t v; memcpy(&v, p, 4); return v; }
    case 64:
      if (fmt == 3) { double v; memcpy(&v, p, 8); return v; }
      if (fmt == 2) { int64_t v; memcpy(&v, p, 8); return (double)v; }
      { uint64_t v; memcpy(&v, p, 8); return (double)v; }
  }
  return 0;
}

void convert_row(const uint8_t* src, int bits, int fmt, void* dst, DType type, long n) {
  const int sb = bits / 8;
  if (type == DT_F32) {
    float* d = (float*)dst;
    if (bits == 32 && fmt == 3) { memcpy(d, src, (size_t)n * 4); return; }
    for (long i = 0; i < n; i++) d[i] = (float)sample_as_double(src + (size_t)i * sb, bits, fmt);
  } else if (type == DT_I16) {
    int16_t* d = (int16_t*)dst;
    if (bits == 16 && fmt == 2) { memcpy(d, src, (size_t)n * 2); return; }
    for (long i = 0; i < n; i++)
      d[i] = narrow_from_double<int16_t>(sample_as_double(src + (size_t)i * sb, bits, fmt), -32768.0, 32767.0);
  } else {
    int32_t* d = (int32_t*)dst;
    if (bits == 32 && fmt == 2) { memcpy(d, src, (size_t)n * 4); return; }
    for (long i = 0; i < n; i++)
      d[i] = narrow_from_double<int32_t>(sample_as_double(src + (size_t)i * sb, bits, fmt), -2147483648.0, 2147483647.0);
  }
}

}  // namespace

// ------------------------------------------------------------------ LZW
bool lzw_decode(const uint8_t* in, size_t n, std::vector<uint8_t>* out, size_t expect) {
  // TIFF flavour: MSB-first codes, ClearCode 256, EOI 257, "early change" of the code width.
  out->resize(expect);
  uint8_t* o = out->data();
  size_t op = 0;
  struct Entry { int32_t prev; uint16_t len; uint8_t first, last; };
  Entry tab[4096 + 2];
  for (int i = 0; i < 256; i++) tab[i] = {-1, 1, (uint8_t)i, (uint8_t)i};
  int next = 258, bits = 9, prev = -1;
  uint64_t acc = 0; int nacc = 0; size_t pos = 0;
  while (op < expect) {
    while (nacc <= 56 && pos < n) { acc = (acc << 8) | in[pos++]; nacc += 8; }
    if (nacc < bits) break;
    const int code = (int)((acc >> (nacc - bits)) & ((1u << bits) - 1));
    nacc -= bits;
    if (code == 257) break;
    if (code == 256) { next = 258; bits = 9; prev = -1; continue; }
    if (prev < 0) {
      if (code >= 256) return false;
      o[op++] = (uint8_t)code;
      prev = code;
      continue;
    }
    int emit = code;
    bool kwk = false;
    if (code >= next) {            // KwKwK case
      if (code != next) return false;
      emit = prev; kwk = true;
    }
    const size_t len = tab[emit].len;
    const size_t total = len + (kwk ? 1 : 0);
    const uint8_t firstc = tab[emit].first;
    if (op + total <= expect) {
      int c = emit;
      for (size_t k = len; k-- > 0;) { o[op + k] = tab[c].last; c = tab[c].prev; }
      if (kwk) o[op + len] = firstc;
      op += total;
    } else {                       // last string runs past the expected size: keep the part that fits
      uint8_t tmp[4100];
      int c = emit;
      for (size_t k = len; k-- > 0;) { tmp[k] = tab[c].last; c = tab[c].prev; }
      if (kwk) tmp[len] = firstc;
      const size_t fit = expect - op;
      memcpy(o + op, tmp, fit);
      op = expect;
    }
    if (next < 4096) {
      tab[next] = {prev, (uint16_t)(tab[prev].len + 1), tab[prev].first, firstc};
      next++;
    }
    if (next == (1 << bits) - 1 && bits < 12) bits++;   // TIFF "early change"
    prev = code;
  }
  return op == expect;
}

void lzw_encode(const uint8_t* in, size_t n, std::vector<uint8_t>* out) {
  // TIFF flavour (see lzw_decode).  The string table is an open-addressing hash of (prefix code << 8 | byte) with a
  // generation stamp per slot, so that the frequent table resets of barely compressible data (float rasters) cost nothing.
  out->resize(n + n / 2 + 16);                   // 12-bit codes for 8-bit symbols: at most 1.5 bytes per byte (+ clear / EOI)
  uint8_t* o = out->data();
  size_t op = 0;
  uint64_t acc = 0; int nacc = 0;
  auto put = [&](int code, int bits) {
    acc = (acc << bits) | (uint32_t)code; nacc += bits;
    while (nacc >= 8) { o[op++] = (uint8_t)(acc >> (nacc - 8)); nacc -= 8; }
  };
  constexpr int HBITS = 14, HSIZE = 1 << HBITS;
  struct Slot { uint32_t key; uint16_t val, gen; };
  std::vector<Slot> tab(HSIZE, Slot{0, 0, 0});
  uint16_t gen = 1;
  auto reset = [&]() { if (++gen == 0) { std::fill(tab.begin(), tab.end(), Slot{0, 0, 0}); gen = 1; } };
  int next = 258, bits = 9;
  put(256, bits);
  if (n == 0) { put(257, bits); if (nacc) o[op++] = (uint8_t)(acc << (8 - nacc)); out->resize(op); return; }
  int cur = in[0];
  for (size_t i = 1; i < n; i++) {
    const int c = in[i];
    const uint32_t key = ((uint32_t)cur << 8) | (uint32_t)c;
    uint32_t h = (key * 2654435761u) >> (32 - HBITS);
    bool found = false;
    while (tab[h].gen == gen) {
      if (tab[h].key == key) { cur = tab[h].val; found = true; break; }
      h = (h + 1) & (HSIZE - 1);
    }
    if (found) continue;
    put(cur, bits);
    tab[h] = Slot{key, (uint16_t)next, gen};
    next++;
    if (next == (1 << bits) - 1 + 1 && bits < 12) bits++;   // writer lags reader by one entry
    if (next >= 4094) { put(256, bits); reset(); next = 258; bits = 9; }
    cur = c;
  }
  put(cur, bits);
  next++;
  if (next == (1 << bits) - 1 + 1 && bits < 12) bits++;
  put(257, bits);
  if (nacc) o[op++] = (uint8_t)((acc << (8 - nacc)) & 0xff);
  out->resize(op);
}

// ------------------------------------------------------------------ Raster
Raster::~Raster() { if (fp_) fclose(fp_); }

bool Raster::open(const std::string& path, std::string* err) {
  fp_ = fopen(path.c_str(), "rb");
  if (!fp_) { if (err) *err = "cannot open " + path; return false; }
  uint8_t hdr[16];
  if (!pread_all(fp_, 0, hdr, 8)) { *err = "short file"; return false; }
  if (hdr[0] == 'I' && hdr[1] == 'I') swap_ = false;
  else if (hdr[0] == 'M' && hdr[1] == 'M') swap_ = true;
  else { *err = "not a TIFF file"; return false; }
  uint16_t magic; memcpy(&magic, hdr + 2, 2); if (swap_) magic = bswap16(magic);
  uint64_t ifd_off;
  if (magic == 42) { uint32_t o; memcpy(&o, hdr + 4, 4); if (swap_) o = bswap32(o); ifd_off = o; big_ = false; }
  else if (magic == 43) {
    if (!pread_all(fp_, 0, hdr, 16)) { *err = "short file"; return false; }
    memcpy(&ifd_off, hdr + 8, 8); if (swap_) ifd_off = bswap64(ifd_off); big_ = true;
  } else { *err = "bad TIFF magic"; return false; }

  uint64_t nent;
  if (big_) { uint64_t v; if (!pread_all(fp_, ifd_off, &v, 8)) { *err = "bad IFD"; return false; } nent = swap_ ? bswap64(v) : v; ifd_off += 8; }
  else { uint16_t v; if (!pread_all(fp_, ifd_off, &v, 2)) { *err = "bad IFD"; return false; } nent = swap_ ? bswap16(v) : v; ifd_off += 2; }
  const int esz = big_ ? 20 : 12;
  // nothing in the file's own tables is trusted: every table must lie inside the file
  uint64_t fsize = 0;
  { struct stat sb; if (fstat(fileno(fp_), &sb) != 0) { *err = "cannot stat file"; return false; } fsize = (uint64_t)sb.st_size; }
  if (nent == 0 || nent > 65536 || ifd_off > fsize || nent * esz > fsize - ifd_off) { *err = "bad IFD (entry table outside the file)"; return false; }
  std::vector<uint8_t> ents(nent * esz);
  if (!pread_all(fp_, ifd_off, ents.data(), ents.size())) { *err = "bad IFD"; return false; }

  std::map<uint16_t, RawTag> tags;
  for (uint64_t e = 0; e < nent; e++) {
    const uint8_t* p = ents.data() + e * esz;
    uint16_t tag, type; memcpy(&tag, p, 2); memcpy(&type, p + 2, 2);
    if (swap_) { tag = bswap16(tag); type = bswap16(type); }
    uint64_t count;
    if (big_) { memcpy(&count, p + 4, 8); if (swap_) count = bswap64(count); }
    else { uint32_t c; memcpy(&c, p + 4, 4); if (swap_) c = bswap32(c); count = c; }
    if (type == 0 || type > 18 || kTypeSize[type] == 0) continue;
    if (count > fsize) { *err = "bad tag (count exceeds the file size)"; return false; }
    const uint64_t nbytes = count * kTypeSize[type];
    if (nbytes > fsize) { *err = "bad tag (data exceeds the file size)"; return false; }
    RawTag rt; rt.type = type; rt.count = count; rt.data.resize(nbytes);
    const int inl = big_ ? 8 : 4;
    const uint8_t* vp = p + (big_ ? 12 : 8);
    if (nbytes <= (uint64_t)inl) memcpy(rt.data.data(), vp, nbytes);
    else {
      uint64_t off;
      if (big_) { memcpy(&off, vp, 8); if (swap_) off = bswap64(off); }
      else { uint32_t o; memcpy(&o, vp, 4); if (swap_) o = bswap32(o); off = o; }
      if (off > fsize || nbytes > fsize - off || !pread_all(fp_, off, rt.data.data(), nbytes)) { *err = "bad tag data"; return false; }
    }
    if (swap_) swap_elems(rt.data.data(), (type == 5 || type == 10) ? count * 2 : count,
                          (type == 5 || type == 10) ? 4 : kTypeSize[type]);
    tags[tag] = std::move(rt);
  }
  auto geti = [&](uint16_t t, uint64_t def) -> uint64_t { auto it = tags.find(t); return it == tags.end() || it->second.count == 0 ? def : tag_u64(it->second, 0); };
  width_ = (uint32_t)geti(256, 0); height_ = (uint32_t)geti(257, 0);
  bits_ = (int)geti(258, 1); compression_ = (int)geti(259, 1);
  sample_format_ = (int)geti(339, 1); predictor_ = (int)geti(317, 1);
  if (sample_format_ == 4) sample_format_ = 1;
  if (geti(277, 1) != 1) { *err = "only single-band rasters are supported"; return false; }
  if (!(bits_ == 8 || bits_ == 16 || bits_ == 32 || bits_ == 64)) { *err = "unsupported bits per sample"; return false; }
  if (width_ == 0 || height_ == 0) { *err = "empty raster"; return false; }
  if (!(compression_ == 1 || compression_ == 5 || compression_ == 8 || compression_ == 32946)) { *err = "unsupported compression " + std::to_string(compression_); return false; }
  if (tags.count(322)) {
    tiled_ = true; block_w_ = (uint32_t)geti(322, 0); block_h_ = (uint32_t)geti(323, 0);
    if (!tags.count(324) || !tags.count(325)) { *err = "no tile offsets / byte counts"; return false; }
    const RawTag &o = tags[324], &c = tags[325];
    if (c.count < o.count) { *err = "fewer tile byte counts than tile offsets"; return false; }
    for (uint64_t i = 0; i < o.count; i++) { offsets_.push_back(tag_u64(o, i)); counts_.push_back(tag_u64(c, i)); }
  } else {
    tiled_ = false; block_w_ = width_;
    uint64_t rps = geti(278, height_); if (rps == 0 || rps > height_) rps = height_;
    block_h_ = (uint32_t)rps;
    if (!tags.count(273)) { *err = "no strip offsets"; return false; }
    const RawTag& o = tags[273];
    for (uint64_t i = 0; i < o.count; i++) offsets_.push_back(tag_u64(o, i));
    if (tags.count(279)) {
      const RawTag& c = tags[279];
      if (c.count < o.count) { *err = "fewer strip byte counts than strip offsets"; return false; }
      for (uint64_t i = 0; i < o.count; i++) counts_.push_back(tag_u64(c, i));
    }
    else for (uint64_t i = 0; i < o.count; i++) counts_.push_back((uint64_t)block_h_ * width_ * (bits_ / 8));
  }
  if (block_w_ == 0 || block_h_ == 0) { *err = "bad block size"; return false; }
  {
    // every block the raster needs must exist and lie inside the file (extra blocks are ignored)
    const uint64_t bx = ((uint64_t)width_ + block_w_ - 1) / block_w_, by = ((uint64_t)height_ + block_h_ - 1) / block_h_;
    const uint64_t need = tiled_ ? bx * by : by;
    if (offsets_.size() < need || counts_.size() < need) { *err = "block tables are shorter than the raster needs"; return false; }
    for (uint64_t i = 0; i < need; i++)
      if (offsets_[i] > fsize || counts_[i] > fsize - offsets_[i]) { *err = "block " + std::to_string(i) + " lies outside the file"; return false; }
    offsets_.resize(need); counts_.resize(need);
    if ((uint64_t)block_w_ * block_h_ * (bits_ / 8) > (1ull << 32)) { *err = "unreasonable block size"; return false; }
  }
  if (predictor_ == 2 && bits_ == 64) { *err = "horizontal predictor with 64-bit samples is not supported"; return false; }
  if (predictor_ == 3 && !(sample_format_ == 3 && (bits_ == 32 || bits_ == 64))) { *err = "floating point predictor on a non-float raster"; return false; }
  if (!(predictor_ == 1 || predictor_ == 2 || predictor_ == 3)) { *err = "unsupported predictor " + std::to_string(predictor_); return false; }

  // nodata (GDAL_NODATA, ASCII)
  if (tags.count(42113)) {
    const RawTag& t = tags[42113];
    std::string s((const char*)t.data.data(), t.data.size());
    char* endp = nullptr;
    double v = strtod(s.c_str(), &endp);
    if (endp != s.c_str()) { has_nodata_ = true; nodata_ = v; }
    else if (s.find("nan") != std::string::npos) { has_nodata_ = true; nodata_ = NAN; }
  }
  // georeferencing
  for (uint16_t t : {33550, 33922, 34264, 34735, 34736, 34737})
    if (tags.count(t)) geo_.geotags[t] = tags[t];
  if (tags.count(34264) && tags[34264].count >= 16) {
    const RawTag& m = tags[34264];
    geo_.gt[0] = tag_double(m, 3); geo_.gt[1] = tag_double(m, 0); geo_.gt[2] = tag_double(m, 1);
    geo_.gt[3] = tag_double(m, 7); geo_.gt[4] = tag_double(m, 4); geo_.gt[5] = tag_double(m, 5);
  } else if (tags.count(33550) && tags[33550].count >= 2) {
    const RawTag& s = tags[33550];
    double sx = tag_double(s, 0), sy = tag_double(s, 1);
    double ti = 0, tj = 0, tx = 0, ty = 0;
    if (tags.count(33922) && tags[33922].count >= 6) {
      const RawTag& tp = tags[33922];
      ti = tag_double(tp, 0); tj = tag_double(tp, 1); tx = tag_double(tp, 3); ty = tag_double(tp, 4);
    }
    geo_.gt[1] = sx; geo_.gt[5] = -sy; geo_.gt[2] = geo_.gt[4] = 0;
    geo_.gt[0] = tx - ti * sx; geo_.gt[3] = ty + tj * sy;
  }
  if (tags.count(34735)) {
    const RawTag& k = tags[34735];
    bool pixel_is_point = false;
    for (uint64_t i = 4; i + 3 < k.count; i += 4) {
      uint64_t key = tag_u64(k, i), loc = tag_u64(k, i + 1), val = tag_u64(k, i + 3);
      if (key == 1024 && loc == 0) geo_.is_geographic = (val == 2);
      if (key == 1025 && loc == 0) pixel_is_point = (val == 2);
    }
    if (pixel_is_point && !tags.count(34264)) {  // GDAL shifts PixelIsPoint rasters by half a cell
      geo_.gt[0] -= 0.5 * geo_.gt[1];
      geo_.gt[3] -= 0.5 * geo_.gt[5];
    }
  }
  return true;
}

void Raster::cell_sizes(std::vector<double>* dxc, std::vector<double>* dyc) const {
  // tiffIO.cpp:96-151.  PI is the reference's truncated literal (commonLib.h:76).
  const double PI = 3.14159265359;
  const double elipa = 6378137.000, elipb = 6356752.314, boa = elipb / elipa;
  const double dlon = std::fabs(geo_.gt[1]), dlat = std::fabs(geo_.gt[5]);
  dxc->assign(height_, dlon);
  dyc->assign(height_, dlat);
  if (!geo_.is_geographic) return;
  const double ytopedge = geo_.gt[3];
  const double yllcenter = ytopedge - (height_ * dlat) - dlat / 2.;
  for (uint32_t j = 0; j < height_; j++) {
    float rowlat = (float)(yllcenter + (height_ - j - 1) * dlat);   // float as in the reference
    double la = dlat * PI / 180., lo = dlon * PI / 180., lat = (double)rowlat * PI / 180.;
    double beta = atan(boa * tan(lat));
    double dbeta = la * boa * (cos(beta) / cos(lat)) * (cos(beta) / cos(lat));
    double ds2 = (pow(elipa * sin(beta), 2) + pow(elipb * cos(beta), 2)) * pow(dbeta, 2);
    (*dxc)[j] = elipa * cos(beta) * std::fabs(lo);
    (*dyc)[j] = sqrt(ds2);
  }
}

bool Raster::load_block(uint64_t idx, std::vector<uint8_t>* out, std::string* err) {
  const int sb = bits_ / 8;
  const size_t raw = (size_t)block_w_ * block_h_ * sb;
  if (idx >= offsets_.size()) { *err = "block index out of range"; return false; }
  // the last strip may be short
  size_t expect = raw;
  if (!tiled_) {
    uint64_t row0 = idx * block_h_;
    uint64_t rows = std::min<uint64_t>(block_h_, height_ - row0);
    expect = (size_t)rows * width_ * sb;
  }
  std::vector<uint8_t> comp;
  if (compression_ == 1) {                      // stored: straight into the block buffer
    if (counts_[idx] < expect) { *err = "raster block too short"; return false; }
    out->resize(counts_[idx]);
    if (!pread_all(fp_, offsets_[idx], out->data(), out->size())) { *err = "short read of raster block"; return false; }
  } else {
    comp.resize(counts_[idx]);
    if (!pread_all(fp_, offsets_[idx], comp.data(), comp.size())) { *err = "short read of raster block"; return false; }
  }
  if (compression_ == 1) {
  } else if (compression_ == 5) {
    if (!lzw_decode(comp.data(), comp.size(), out, expect)) { *err = "LZW decode failed"; return false; }
  } else {
    out->resize(expect);
    uLongf dl = (uLongf)expect;
    int rc = uncompress(out->data(), &dl, comp.data(), (uLong)comp.size());
    if (rc != Z_OK && rc != Z_BUF_ERROR) { *err = "deflate decode failed"; return false; }
    if (dl < expect) { *err = "deflate block too short"; return false; }
  }
  const size_t rows = expect / ((size_t)block_w_ * sb);
  if (predictor_ == 2) {
    for (size_t r = 0; r < rows; r++) {
      uint8_t* row = out->data() + r * block_w_ * sb;
      if (sb == 1) for (uint32_t i = 1; i < block_w_; i++) row[i] = (uint8_t)(row[i] + row[i - 1]);
      else if (sb == 2) { uint16_t* v = (uint16_t*)row; if (swap_) swap_elems(row, block_w_, 2); for (uint32_t i = 1; i < block_w_; i++) v[i] = (uint16_t)(v[i] + v[i - 1]); if (swap_) swap_elems(row, block_w_, 2); }
      else if (sb == 4) { uint32_t* v = (uint32_t*)row; if (swap_) swap_elems(row, block_w_, 4); for (uint32_t i = 1; i < block_w_; i++) v[i] += v[i - 1]; if (swap_) swap_elems(row, block_w_, 4); }
    }
  } else if (predictor_ == 3) {
    std::vector<uint8_t> tmp((size_t)block_w_ * sb);
    for (size_t r = 0; r < rows; r++) {
      uint8_t* row = out->data() + r * block_w_ * sb;
      const size_t nb = (size_t)block_w_ * sb;
      for (size_t i = 1; i < nb; i++) row[i] = (uint8_t)(row[i] + row[i - 1]);
      memcpy(tmp.data(), row, nb);
      for (uint32_t i = 0; i < block_w_; i++)
        for (int b = 0; b < sb; b++)   // planes are stored most-significant byte first
          row[(size_t)i * sb + (sb - 1 - b)] = tmp[(size_t)b * block_w_ + i];
    }
    return true;   // samples are now in host (little-endian) order whatever the file order
  }
  if (swap_ && sb > 1) swap_elems(out->data(), expect / sb, sb);
  return true;
}

bool Raster::read(long xstart, long ystart, long nrows, long ncols, void* dest, DType type,
                  std::string* err, long dest_stride) {
  if (xstart < 0 || ystart < 0 || xstart + ncols > (long)width_ || ystart + nrows > (long)height_) {
    *err = "read window outside raster"; return false;
  }
  if (dest_stride == 0) dest_stride = ncols;
  const int sb = bits_ / 8, db = dtype_bytes(type);
  const uint64_t blocks_across = tiled_ ? (width_ + block_w_ - 1) / block_w_ : 1;
  // the blocks (strips / tiles) that intersect the window, decoded by a small pool of threads: every block
  // writes a disjoint part of dest (LZW / Deflate decoding is the cost of reading real-world DEMs)
  struct Job { long by; uint64_t bx; };
  std::vector<Job> jobs;
  for (long by = ystart / block_h_; by <= (ystart + nrows - 1) / (long)block_h_; by++)
    for (uint64_t bx = (uint64_t)xstart / block_w_; bx <= (uint64_t)(xstart + ncols - 1) / block_w_; bx++) jobs.push_back({by, bx});
  std::atomic<size_t> next(0);
  std::atomic<bool> failed(false);
  std::mutex emu;
  // stored strips of the file's own sample type, whole rows: the file bytes ARE the destination bytes
  const bool direct = compression_ == 1 && !tiled_ && predictor_ == 1 && !(swap_ && sb > 1) && xstart == 0 && ncols == (long)width_ &&
                      dest_stride == ncols &&
                      ((type == DT_F32 && bits_ == 32 && sample_format_ == 3) || (type == DT_I16 && bits_ == 16 && sample_format_ == 2) ||
                       (type == DT_I32 && bits_ == 32 && sample_format_ == 2));
  auto work = [&]() {
    std::vector<uint8_t> blk;
    std::string e;
    for (;;) {
      const size_t j = next.fetch_add(1);
      if (j >= jobs.size() || failed.load()) return;
      const long by = jobs[j].by; const uint64_t bx = jobs[j].bx;
      if (direct) {
        const long r0 = std::max<long>(ystart, by * (long)block_h_), r1 = std::min<long>(ystart + nrows, (by + 1) * (long)block_h_);
        const size_t rowb = (size_t)width_ * sb;
        const uint64_t off = ((size_t)by < offsets_.size() ? offsets_[(size_t)by] : 0) + (uint64_t)(r0 - by * (long)block_h_) * rowb;
        const size_t want = (size_t)(r1 - r0) * rowb;
        if ((size_t)by >= offsets_.size() || (uint64_t)(r1 - by * (long)block_h_) * rowb > counts_[(size_t)by] ||
            !pread_all(fp_, off, (uint8_t*)dest + (size_t)(r0 - ystart) * rowb, want)) {
          std::lock_guard<std::mutex> g(emu);
          if (!failed.exchange(true)) *err = "short read of raster block";
          return;
        }
        continue;
      }
      if (!load_block((uint64_t)by * blocks_across + bx, &blk, &e)) {
        std::lock_guard<std::mutex> g(emu);
        if (!failed.exchange(true)) *err = e;
        return;
      }
      const long r0 = std::max<long>(ystart, by * (long)block_h_);
      const long r1 = std::min<long>(ystart + nrows, (by + 1) * (long)block_h_);
      const long c0 = std::max<long>(xstart, (long)(bx * block_w_));
      const long c1 = std::min<long>(xstart + ncols, (long)((bx + 1) * block_w_));
      for (long r = r0; r < r1; r++) {
        const uint8_t* src = blk.data() + ((size_t)(r - by * (long)block_h_) * block_w_ + (size_t)(c0 - (long)(bx * block_w_))) * sb;
        uint8_t* dst = (uint8_t*)dest + ((size_t)(r - ystart) * dest_stride + (size_t)(c0 - xstart)) * db;
        convert_row(src, bits_, sample_format_, dst, type, c1 - c0);
      }
    }
  };
  unsigned nthreads = std::min<unsigned>(compression_ == 1 ? 8u : 16u, std::max(1u, std::thread::hardware_concurrency()));
  nthreads = (unsigned)std::min<size_t>(nthreads, jobs.size());
  if (nthreads <= 1) work();
  else {
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthreads; t++) pool.emplace_back(work);
    for (auto& t : pool) t.join();
  }
  return !failed.load();
}

// ------------------------------------------------------------------ Writer
Writer::~Writer() { if (fp_) fclose(fp_); }

namespace {
struct OutTag { uint16_t tag, type; uint64_t count; std::vector<uint8_t> data; };
template <typename T> std::vector<uint8_t> bytes_of(const std::vector<T>& v) {
  std::vector<uint8_t> b(v.size() * sizeof(T)); if (!v.empty()) memcpy(b.data(), v.data(), b.size()); return b;
}
}  // namespace

bool Writer::create(const std::string& path, uint32_t width, uint32_t height, DType type,
                    double nodata, const GeoInfo& geo, int compression, std::string* err,
                    bool force_bigtiff) {
  width_ = width; height_ = height; type_ = type; compression_ = compression;
  const int cb = dtype_bytes(type);
  // reference rule (tiffIO.cpp:322-330); also needed whenever offsets overflow 32 bits
  const double fileGB = (double)cb * (double)width * (double)height / 1000000000.0;
  big_ = force_bigtiff || fileGB > 4.0;
  // strips of ~1 MiB keep the offset tables small (65536^2 float: 16384 strips)
  uint64_t rps = std::max<uint64_t>(1, (1u << 20) / ((uint64_t)width * cb));
  rows_per_strip_ = (uint32_t)std::min<uint64_t>(rps, height);
  const uint64_t nstrips = (height + rows_per_strip_ - 1) / rows_per_strip_;
  offsets_.assign(nstrips, 0); counts_.assign(nstrips, 0);
  fp_ = fopen(path.c_str(), "wb+");
  if (!fp_) { *err = "cannot create " + path; return false; }

  std::vector<OutTag> tags;
  auto add_short = [&](uint16_t t, uint16_t v) { tags.push_back({t, 3, 1, bytes_of(std::vector<uint16_t>{v})}); };
  auto add_long = [&](uint16_t t, uint32_t v) { tags.push_back({t, 4, 1, bytes_of(std::vector<uint32_t>{v})}); };
  add_long(256, width); add_long(257, height);
  add_short(258, (uint16_t)(cb * 8)); add_short(259, (uint16_t)compression);
  add_short(262, 1);
  const uint16_t otype = big_ ? 16 : 4;
  const int osz = big_ ? 8 : 4;
  tags.push_back({273, otype, nstrips, std::vector<uint8_t>(nstrips * osz, 0)});
  add_short(277, 1); add_long(278, rows_per_strip_);
  tags.push_back({279, otype, nstrips, std::vector<uint8_t>(nstrips * osz, 0)});
  add_short(284, 1);
  add_short(339, (uint16_t)(type == DT_F32 ? 3 : 2));
  for (auto& kv : geo.geotags) tags.push_back({kv.first, kv.second.type, kv.second.count, kv.second.data});
  if (!geo.geotags.count(33550) && !geo.geotags.count(34264)) {
    // no GeoTIFF tags on the source: still carry the geotransform the way GDAL would
    tags.push_back({33550, 12, 3, bytes_of(std::vector<double>{std::fabs(geo.gt[1]), std::fabs(geo.gt[5]), 0.0})});
    tags.push_back({33922, 12, 6, bytes_of(std::vector<double>{0, 0, 0, geo.gt[0], geo.gt[3], 0})});
  }
  {
    char buf[64];
    snprintf(buf, sizeof buf, "%.17g", nodata);     // GDAL writes the double repr of the nodata value
    std::string s(buf);
    std::vector<uint8_t> d(s.begin(), s.end()); d.push_back(0);
    tags.push_back({42113, 2, d.size(), d});
  }
  std::sort(tags.begin(), tags.end(), [](const OutTag& a, const OutTag& b) { return a.tag < b.tag; });

  // layout: header | IFD | out-of-line tag data | raster data
  const uint64_t hdr = big_ ? 16 : 8;
  const int esz = big_ ? 20 : 12, inl = big_ ? 8 : 4;
  const uint64_t ifd_size = (big_ ? 8 : 2) + tags.size() * esz + (big_ ? 8 : 4);
  uint64_t extra = hdr + ifd_size;
  std::vector<uint8_t> ifd(ifd_size, 0), tail;
  size_t p = 0;
  if (big_) { uint64_t n = tags.size(); memcpy(&ifd[p], &n, 8); p += 8; }
  else { uint16_t n = (uint16_t)tags.size(); memcpy(&ifd[p], &n, 2); p += 2; }
  for (auto& t : tags) {
    memcpy(&ifd[p], &t.tag, 2); memcpy(&ifd[p + 2], &t.type, 2);
    if (big_) memcpy(&ifd[p + 4], &t.count, 8); else { uint32_t c = (uint32_t)t.count; memcpy(&ifd[p + 4], &c, 4); }
    uint8_t* vp = &ifd[p + (big_ ? 12 : 8)];
    if (t.data.size() <= (size_t)inl) {
      memcpy(vp, t.data.data(), t.data.size());
      if (t.tag == 273) offsets_pos_ = hdr + (vp - ifd.data());
      if (t.tag == 279) counts_pos_ = hdr + (vp - ifd.data());
    } else {
      if (tail.size() & 1) tail.push_back(0);
      uint64_t off = extra + tail.size();
      if (big_) memcpy(vp, &off, 8); else { uint32_t o = (uint32_t)off; memcpy(vp, &o, 4); }
      if (t.tag == 273) offsets_pos_ = off;
      if (t.tag == 279) counts_pos_ = off;
      tail.insert(tail.end(), t.data.begin(), t.data.end());
    }
    p += esz;
  }
  uint8_t h[16] = {'I', 'I', 0, 0};
  if (big_) { h[2] = 43; h[4] = 8; uint64_t o = 16; memcpy(h + 8, &o, 8); }
  else { h[2] = 42; uint32_t o = 8; memcpy(h + 4, &o, 4); }
  if (fwrite(h, 1, hdr, fp_) != hdr || fwrite(ifd.data(), 1, ifd.size(), fp_) != ifd.size() ||
      (!tail.empty() && fwrite(tail.data(), 1, tail.size(), fp_) != tail.size())) { *err = "write failed"; return false; }
  data_start_ = extra + tail.size();
  data_start_ = (data_start_ + 15) & ~15ull;
  append_pos_ = data_start_;
  if (compression_ == 1) {
    // fixed layout: strip s starts at data_start_ + s * rows_per_strip * width * cb
    for (uint64_t s = 0; s < nstrips; s++) {
      uint64_t rows = std::min<uint64_t>(rows_per_strip_, height - s * rows_per_strip_);
      offsets_[s] = data_start_ + s * (uint64_t)rows_per_strip_ * width * cb;
      counts_[s] = rows * width * cb;
    }
    if (!big_ && offsets_.back() + counts_.back() > 0xffffffffull) { *err = "raster too large for classic TIFF"; return false; }
  }
  return true;
}

bool Writer::flush_batch(std::string* err) {
  // compress the queued strips on a small thread pool, then append them in strip order
  const size_t n = batch_.size();
  if (n == 0) return true;
  std::vector<std::vector<uint8_t>> comp(n);
  std::atomic<size_t> next{0};
  std::atomic<bool> failed{false};
  auto work = [&]() {
    for (size_t i; (i = next.fetch_add(1)) < n;) {
      const std::vector<uint8_t>& raw = batch_[i].second;
      if (compression_ == 5) lzw_encode(raw.data(), raw.size(), &comp[i]);
      else {
        uLongf dl = compressBound((uLong)raw.size()); comp[i].resize(dl);
        if (compress2(comp[i].data(), &dl, raw.data(), (uLong)raw.size(), 6) != Z_OK) { failed = true; return; }
        comp[i].resize(dl);
      }
    }
  };
  const unsigned nthreads = (unsigned)std::min<size_t>(n, std::min<unsigned>(32u, std::max(1u, std::thread::hardware_concurrency())));
  if (nthreads <= 1) work();
  else {
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthreads; t++) pool.emplace_back(work);
    for (auto& t : pool) t.join();
  }
  if (failed.load()) { *err = "deflate failed"; return false; }
  if (fseeko(fp_, (off_t)append_pos_, SEEK_SET) != 0) { *err = "write failed"; return false; }
  for (size_t i = 0; i < n; i++) {
    if (fwrite(comp[i].data(), 1, comp[i].size(), fp_) != comp[i].size()) { *err = "write failed"; return false; }
    offsets_[batch_[i].first] = append_pos_; counts_[batch_[i].first] = comp[i].size();
    append_pos_ += comp[i].size();
  }
  batch_.clear();
  if (!big_ && append_pos_ > 0xffffffffull) { *err = "raster too large for classic TIFF"; return false; }
  return true;
}

bool Writer::write_rows(long ystart, long nrows, const void* src, std::string* err, long src_stride) {
  if (src_stride == 0) src_stride = width_;
  const int cb = dtype_bytes(type_);
  const size_t rowb = (size_t)width_ * cb;
  if (ystart < 0 || ystart + nrows > (long)height_) { *err = "write window outside raster"; return false; }
  if (compression_ == 1) {
    if (src_stride == (long)width_) {
      if (fseeko(fp_, (off_t)(data_start_ + (uint64_t)ystart * rowb), SEEK_SET) != 0 ||
          fwrite(src, 1, rowb * nrows, fp_) != rowb * (size_t)nrows) { *err = "write failed"; return false; }
    } else {
      for (long r = 0; r < nrows; r++) {
        if (fseeko(fp_, (off_t)(data_start_ + (uint64_t)(ystart + r) * rowb), SEEK_SET) != 0 ||
            fwrite((const uint8_t*)src + (size_t)r * src_stride * cb, 1, rowb, fp_) != rowb) { *err = "write failed"; return false; }
      }
    }
    return true;
  }
  // compressed: accumulate whole strips (rows must arrive in order within a strip)
  for (long r = 0; r < nrows; r++) {
    const long y = ystart + r;
    const long s = y / rows_per_strip_;
    const long srow0 = s * (long)rows_per_strip_;
    const long srows = std::min<long>(rows_per_strip_, (long)height_ - srow0);
    if (pending_row0_ != srow0) {
      if (y != srow0) { *err = "compressed output needs strip-ordered rows"; return false; }
      pending_.assign((size_t)srows * rowb, 0); pending_row0_ = srow0;
    }
    memcpy(pending_.data() + (size_t)(y - srow0) * rowb, (const uint8_t*)src + (size_t)r * src_stride * cb, rowb);
    if (y == srow0 + srows - 1) {
      batch_.emplace_back((uint64_t)s, std::move(pending_));
      pending_.clear(); pending_row0_ = -1;
      if (batch_.size() >= 64 && !flush_batch(err)) return false;
    }
  }
  return true;
}

bool Writer::close(std::string* err) {
  if (!fp_) return true;
  bool ok = true;
  if (compression_ != 1) { std::string e; if (!flush_batch(&e)) ok = false; }
  const int osz = big_ ? 8 : 4;
  std::vector<uint8_t> ob(offsets_.size() * osz), cbuf(counts_.size() * osz);
  for (size_t i = 0; i < offsets_.size(); i++) {
    if (big_) { memcpy(&ob[i * 8], &offsets_[i], 8); memcpy(&cbuf[i * 8], &counts_[i], 8); }
    else { uint32_t o = (uint32_t)offsets_[i], c = (uint32_t)counts_[i]; memcpy(&ob[i * 4], &o, 4); memcpy(&cbuf[i * 4], &c, 4); }
  }
  if (fseeko(fp_, (off_t)offsets_pos_, SEEK_SET) != 0 || fwrite(ob.data(), 1, ob.size(), fp_) != ob.size()) ok = false;
  if (fseeko(fp_, (off_t)counts_pos_, SEEK_SET) != 0 || fwrite(cbuf.data(), 1, cbuf.size(), fp_) != cbuf.size()) ok = false;
  if (compression_ == 1) {   // make sure the file has its full length even if some rows were never written
    uint64_t end = offsets_.back() + counts_.back();
    if (fseeko(fp_, 0, SEEK_END) == 0 && (uint64_t)ftello(fp_) < end) {
      if (fseeko(fp_, (off_t)(end - 1), SEEK_SET) != 0 || fputc(0, fp_) == EOF) ok = false;
    }
  }
  if (fclose(fp_) != 0) ok = false;
  fp_ = nullptr;
  if (!ok && err) *err = "write failed while finalising TIFF";
  return ok;
}

std::string output_path_rule(const std::string& name) {
  static const char* known[6] = {".tif", ".img", ".sdat", ".bil", ".bin", ".tiff"};
  size_t dot = name.rfind('.');
  if (dot == std::string::npos) return name + ".tif";
  std::string ext = name.substr(dot);
  for (auto& c : ext) c = (char)tolower((unsigned char)c);
  for (auto k : known) if (ext == k) return name.substr(0, dot) + ext;   // reference lower-cases in place
  return name.substr(0, dot + 1) + "tif";
}

bool compare_rasters(const Raster& a, const std::string& aname, const Raster& b, const std::string& bname) {
  const double tol = 0.0001;
  if (a.width() != b.width()) { printf("Columns do not match: %d %d\n", (int)a.width(), (int)b.width()); return false; }
  if (a.height() != b.height()) { printf("Rows do not match: %d %d\n", (int)a.height(), (int)b.height()); return false; }
  std::vector<double> ax, ay, bx, by;
  a.cell_sizes(&ax, &ay); b.cell_sizes(&bx, &by);
  const double adx = std::fabs(ax[a.height() / 2]), ady = std::fabs(ay[a.height() / 2]);
  const double bdx = std::fabs(bx[b.height() / 2]), bdy = std::fabs(by[b.height() / 2]);
  if (std::fabs(adx - bdx) > tol) { printf("dx does not match: %lf %lf\n", adx, bdx); return false; }
  if (std::fabs(ady - bdy) > tol) { printf("dy does not match: %lf %lf\n", ady, bdy); return false; }
  if (std::fabs(a.geo().gt[0] - b.geo().gt[0]) > 0.0) {
    printf("Warning! Left edge does not match exactly:\n %lf in file %s\n %lf in file %s\n", a.geo().gt[0], aname.c_str(), b.geo().gt[0], bname.c_str());
  }
  if (std::fabs(a.geo().gt[3] - b.geo().gt[3]) > 0.0) {
    printf("Warning! Top edge does not match exactly:\n %lf in file %s\n %lf in file %s\n", a.geo().gt[3], aname.c_str(), b.geo().gt[3], bname.c_str());
  }
  return true;
}

}  // namespace tdio
