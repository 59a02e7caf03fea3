// Outlet points for aread8 / areadinf -o (reference: readoutlets, src/ReadOutlets.cpp:49-189, which goes through
// OGR).  Read natively here: ESRI shapefiles with Point / PointZ / PointM records and GeoJSON files with Point
// features — the two formats TauDEM workflows use.  A directory is a data source whose layers are its .shp files
// (layer name = file name without extension, layer number = alphabetical position).  Like the reference, only
// point layers are accepted and only the first coordinate pair of a feature is used.
#include <dirent.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/taudem_b200.h"

namespace td { void set_error(const std::string& msg); }

namespace {
bool ends_with(const std::string& s, const char* suf) {
  const size_t n = strlen(suf);
  if (s.size() < n) return false;
  for (size_t i = 0; i < n; ++i) if (tolower((unsigned char)s[s.size() - n + i]) != suf[i]) return false;
  return true;
}
std::string stem(const std::string& path) {
  const size_t sl = path.find_last_of("/\\");
  std::string f = sl == std::string::npos ? path : path.substr(sl + 1);
  const size_t dot = f.rfind('.');
  return dot == std::string::npos ? f : f.substr(0, dot);
}
uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
uint32_t le32(const unsigned char* p) { uint32_t v; memcpy(&v, p, 4); return v; }
double le64f(const unsigned char* p) { double v; memcpy(&v, p, 8); return v; }

bool read_shp(const std::string& path, std::vector<double>* x, std::vector<double>* y, std::string* err) {
  FILE* fp = fopen(path.c_str(), "rb");
  if (!fp) { *err = "cannot open " + path; return false; }
  std::vector<unsigned char> buf;
  unsigned char tmp[65536];
  size_t n;
  while ((n = fread(tmp, 1, sizeof tmp, fp)) > 0) buf.insert(buf.end(), tmp, tmp + n);
  fclose(fp);
  if (buf.size() < 100 || be32(buf.data()) != 9994u) { *err = path + " is not a shapefile"; return false; }
  const uint32_t type = le32(buf.data() + 32);
  if (type != 1 && type != 11 && type != 21) { *err = path + ": the outlet layer must hold points"; return false; }
  size_t pos = 100;
  while (pos + 8 <= buf.size()) {
    const size_t len = (size_t)be32(buf.data() + pos + 4) * 2;      // content length in 16-bit words
    const unsigned char* rec = buf.data() + pos + 8;
    if (pos + 8 + len > buf.size()) break;
    if (len >= 20) {
      const uint32_t st = le32(rec);
      if (st == 1 || st == 11 || st == 21) { x->push_back(le64f(rec + 4)); y->push_back(le64f(rec + 12)); }
    }                                                               // shape type 0 = null shape: no geometry, skipped
    pos += 8 + len;
  }
  return true;
}

// GeoJSON: every "coordinates" member whose value starts with a number is a Point position
bool read_geojson(const std::string& path, std::vector<double>* x, std::vector<double>* y, std::string* err) {
  FILE* fp = fopen(path.c_str(), "rb");
  if (!fp) { *err = "cannot open " + path; return false; }
  std::string s;
  char tmp[65536];
  size_t n;
  while ((n = fread(tmp, 1, sizeof tmp, fp)) > 0) s.append(tmp, n);
  fclose(fp);
  size_t pos = 0;
  bool other = false;
  while ((pos = s.find("\"coordinates\"", pos)) != std::string::npos) {
    pos += 13;
    size_t p = s.find('[', pos);
    if (p == std::string::npos) break;
    ++p;
    while (p < s.size() && isspace((unsigned char)s[p])) ++p;
    if (p < s.size() && s[p] == '[') { other = true; continue; }   // nested arrays: a line, polygon or multi-geometry
    char* e1 = nullptr;
    const double vx = strtod(s.c_str() + p, &e1);
    if (e1 == s.c_str() + p) continue;
    const char* q = e1;
    while (*q && (isspace((unsigned char)*q) || *q == ',')) ++q;
    char* e2 = nullptr;
    const double vy = strtod(q, &e2);
    if (e2 == q) continue;
    x->push_back(vx); y->push_back(vy);
  }
  if (x->empty() && other) { *err = path + ": the outlet layer must hold points"; return false; }
  return true;
}
}  // namespace

extern "C" int td_outlets_read(const char* datasrc, const char* lyrname, int uselyrname, int lyrno, double* x, double* y, int cap, int* n) {
  if (!datasrc || !n) { td::set_error("td_outlets_read: bad arguments"); return TD_ERR_ARG; }
  *n = 0;
  std::string path = datasrc, err;
  struct stat stt;
  if (stat(datasrc, &stt) != 0) { printf("Error Opening OGR Data Source .\n"); td::set_error(std::string("cannot open outlet data source ") + datasrc); return TD_ERR_IO; }
  if (S_ISDIR(stt.st_mode)) {
    std::vector<std::string> layers;
    if (DIR* d = opendir(datasrc)) {
      while (dirent* e = readdir(d)) if (ends_with(e->d_name, ".shp")) layers.push_back(e->d_name);
      closedir(d);
    }
    std::sort(layers.begin(), layers.end());
    std::string pick;
    if (uselyrname) { for (auto& l : layers) if (stem(l) == (lyrname ? lyrname : "")) pick = l; }
    else if (lyrno >= 0 && lyrno < (int)layers.size()) pick = layers[lyrno];
    if (pick.empty()) { td::set_error(std::string("outlet layer not found in ") + datasrc); return TD_ERR_IO; }
    path = std::string(datasrc) + "/" + pick;
  } else {
    // a file is a data source with one layer, named after the file
    if (uselyrname && lyrname && stem(path) != lyrname) { td::set_error(std::string("outlet layer ") + lyrname + " not found in " + datasrc); return TD_ERR_IO; }
    if (!uselyrname && lyrno != 0) { td::set_error("outlet layer number out of range"); return TD_ERR_IO; }
  }
  std::vector<double> vx, vy;
  const bool ok = (ends_with(path, ".json") || ends_with(path, ".geojson")) ? read_geojson(path, &vx, &vy, &err) : read_shp(path, &vx, &vy, &err);
  if (!ok) { td::set_error(err); return TD_ERR_IO; }
  *n = (int)vx.size();
  if (x && y) for (int i = 0; i < *n && i < cap; ++i) { x[i] = vx[i]; y[i] = vy[i]; }
  return TD_OK;
}
