// TEST INFRASTRUCTURE ONLY.  GDAL C-API subset over taudem_b200/csrc/tiff_io.*
// so that the reference's tiffIO.cpp runs unchanged (see oracle/shim/gdal.h).
#include "gdal.h"
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "../../taudem_b200/csrc/tiff_io.h"

namespace {
struct DS {
  std::string path;
  bool writing = false, update = false, created = false;
  std::unique_ptr<tdio::Raster> rd;
  std::unique_ptr<tdio::Writer> wr;
  int nx = 0, ny = 0;
  GDALDataType dt = GDT_Float32;
  double nodata = 0; bool has_nodata = false;
  tdio::GeoInfo geo;
  std::string proj;      // "GEOGCS[...]" for geographic rasters, "" otherwise
  // update mode (ranks > 0 of a multi-rank write): rows are patched in place
  FILE* upd = nullptr; uint64_t data_start = 0; int cellbytes = 4;
};
std::map<const char*, DS*> g_proj_owner;
std::string g_err;
tdio::DType to_dtype(GDALDataType t) { return t == GDT_Int16 ? tdio::DT_I16 : t == GDT_Int32 ? tdio::DT_I32 : tdio::DT_F32; }
}  // namespace

extern "C" {
void GDALAllRegister(void) {}
const char* CPLGetLastErrorMsg(void) { return g_err.c_str(); }
GDALDatasetH GDALOpen(const char* name, GDALAccess acc) {
  auto ds = new DS;
  ds->path = name;
  ds->rd.reset(new tdio::Raster);
  if (!ds->rd->open(name, &g_err)) { delete ds; return NULL; }
  ds->nx = ds->rd->width(); ds->ny = ds->rd->height();
  ds->has_nodata = ds->rd->has_nodata(); ds->nodata = ds->rd->has_nodata() ? ds->rd->nodata() : 0;
  ds->geo = ds->rd->geo();
  ds->proj = ds->geo.is_geographic ? "GEOGCS[\"WGS 84\"]" : "";
  int b = ds->rd->bits(), f = ds->rd->sample_format();
  ds->dt = (f == 3) ? (b == 64 ? GDT_Float64 : GDT_Float32) : (b == 16 ? GDT_Int16 : b == 32 ? GDT_Int32 : GDT_Byte);
  if (acc == GA_Update) {
    if (ds->rd->compression() != 1 || ds->rd->tiled()) { g_err = "update needs an uncompressed strip TIFF"; delete ds; return NULL; }
    ds->update = true; ds->data_start = ds->rd->block_offset(0); ds->cellbytes = b / 8;
    ds->rd.reset();
    ds->upd = fopen(name, "rb+");
    if (!ds->upd) { delete ds; return NULL; }
  }
  g_proj_owner[ds->proj.c_str()] = ds;
  return ds;
}
void GDALFlushCache(GDALDatasetH) {}
void GDALClose(GDALDatasetH h) {
  DS* ds = (DS*)h; if (!ds) return;
  if (ds->wr) { std::string e; if (!ds->wr->close(&e)) fprintf(stderr, "gdalshim: %s\n", e.c_str()); }
  if (ds->upd) fclose(ds->upd);
  // datasets opened read-only are kept alive by the reference (never closed); harmless
  for (auto it = g_proj_owner.begin(); it != g_proj_owner.end();) { if (it->second == ds) it = g_proj_owner.erase(it); else ++it; }
  delete ds;
}
GDALDriverH GDALGetDatasetDriver(GDALDatasetH) { return (GDALDriverH) "GTiff"; }
GDALDriverH GDALGetDriverByName(const char* n) { return strcmp(n, "GTiff") == 0 ? (GDALDriverH) "GTiff" : NULL; }
GDALDatasetH GDALCreate(GDALDriverH, const char* name, int nx, int ny, int, GDALDataType dt, char**) {
  auto ds = new DS; ds->path = name; ds->writing = true; ds->nx = nx; ds->ny = ny; ds->dt = dt;
  return ds;
}
const char* GDALGetProjectionRef(GDALDatasetH h) { return ((DS*)h)->proj.c_str(); }
CPLErr GDALSetProjection(GDALDatasetH h, const char* wkt) {
  DS* ds = (DS*)h; auto it = g_proj_owner.find(wkt);
  if (it != g_proj_owner.end()) { ds->geo = it->second->geo; ds->proj = it->second->proj; }
  return 0;
}
CPLErr GDALGetGeoTransform(GDALDatasetH h, double* gt) { memcpy(gt, ((DS*)h)->geo.gt, 6 * sizeof(double)); return 0; }
CPLErr GDALSetGeoTransform(GDALDatasetH h, double* gt) { memcpy(((DS*)h)->geo.gt, gt, 6 * sizeof(double)); return 0; }
GDALRasterBandH GDALGetRasterBand(GDALDatasetH h, int) { return h; }
int GDALGetRasterXSize(GDALDatasetH h) { return ((DS*)h)->nx; }
int GDALGetRasterYSize(GDALDatasetH h) { return ((DS*)h)->ny; }
const char* GDALGetRasterUnitType(GDALRasterBandH) { return ""; }
GDALDataType GDALGetRasterDataType(GDALRasterBandH h) { return ((DS*)h)->dt; }
double GDALGetRasterNoDataValue(GDALRasterBandH h, int* ok) { DS* ds = (DS*)h; if (ok) *ok = ds->has_nodata; return ds->nodata; }
CPLErr GDALSetRasterNoDataValue(GDALRasterBandH h, double v) { DS* ds = (DS*)h; ds->nodata = v; ds->has_nodata = true; return 0; }
CPLErr GDALRasterIO(GDALRasterBandH h, GDALRWFlag rw, int x0, int y0, int xs, int ys, void* buf, int, int, GDALDataType bt, int, int) {
  DS* ds = (DS*)h; std::string e;
  if (rw == GF_Read) {
    if (!ds->rd->read(x0, y0, ys, xs, buf, to_dtype(bt), &e)) { fprintf(stderr, "gdalshim read: %s\n", e.c_str()); return 3; }
    return 0;
  }
  if (bt != ds->dt && !(ds->update)) { fprintf(stderr, "gdalshim: write type mismatch\n"); return 3; }
  if (ds->update) {
    const size_t rowb = (size_t)ds->nx * ds->cellbytes;
    if (x0 != 0 || xs != ds->nx) { fprintf(stderr, "gdalshim: partial-row update unsupported\n"); return 3; }
    if (fseeko(ds->upd, (off_t)(ds->data_start + (uint64_t)y0 * rowb), SEEK_SET) != 0 || fwrite(buf, 1, rowb * ys, ds->upd) != rowb * (size_t)ys) return 3;
    return 0;
  }
  if (!ds->created) {
    ds->wr.reset(new tdio::Writer);
    if (!ds->wr->create(ds->path, ds->nx, ds->ny, to_dtype(ds->dt), ds->nodata, ds->geo, 1, &e)) { fprintf(stderr, "gdalshim create: %s\n", e.c_str()); return 3; }
    ds->created = true;
  }
  if (x0 != 0 || xs != ds->nx) { fprintf(stderr, "gdalshim: partial-row write unsupported\n"); return 3; }
  if (!ds->wr->write_rows(y0, ys, buf, &e)) { fprintf(stderr, "gdalshim write: %s\n", e.c_str()); return 3; }
  return 0;
}
char** CSLSetNameValue(char** l, const char*, const char*) { return l; }

OGRSpatialReferenceH OSRNewSpatialReference(const char* wkt) { return (OGRSpatialReferenceH)(wkt && strncmp(wkt, "GEOGCS", 6) == 0 ? "G" : "P"); }
int OSRIsGeographic(OGRSpatialReferenceH h) { return h && *(const char*)h == 'G'; }
int OSRIsProjected(OGRSpatialReferenceH h) { return !(h && *(const char*)h == 'G'); }
double OSRGetLinearUnits(OGRSpatialReferenceH, char** n) { if (n) *n = (char*)"unknown"; return 1.0; }
const char* OSRGetAttrValue(OGRSpatialReferenceH, const char*, int) { return NULL; }

// OGR (outlets, -o): a point shapefile is a data source with one layer named after the file; no attribute table
// (readoutlets then numbers the outlets itself, src/ReadOutlets.cpp:178-180).  Own little .shp parser: main header
// 100 bytes (shape type at byte 32, little endian), records = 8-byte big-endian header + content starting with
// the shape type and x, y as little-endian doubles.
struct ShimLayer { std::string name; int type = 0; std::vector<double> x, y; size_t next = 0; };
struct ShimFeature { double x, y; };
void OGRRegisterAll(void) {}
OGRDataSourceH OGROpen(const char* path, int, OGRSFDriverH*) {
  FILE* fp = fopen(path, "rb");
  if (!fp) return NULL;
  std::vector<unsigned char> b;
  unsigned char tmp[4096]; size_t n;
  while ((n = fread(tmp, 1, sizeof tmp, fp)) > 0) b.insert(b.end(), tmp, tmp + n);
  fclose(fp);
  if (b.size() < 100 || !(b[0] == 0 && b[1] == 0 && b[2] == 0x27 && b[3] == 0x0a)) return NULL;
  ShimLayer* L = new ShimLayer;
  std::string p = path;
  const size_t sl = p.find_last_of('/'), dot = p.rfind('.');
  L->name = p.substr(sl == std::string::npos ? 0 : sl + 1, dot - (sl == std::string::npos ? 0 : sl + 1));
  memcpy(&L->type, &b[32], 4);
  for (size_t pos = 100; pos + 8 <= b.size();) {
    const size_t len = 2 * (((size_t)b[pos + 4] << 24) | ((size_t)b[pos + 5] << 16) | ((size_t)b[pos + 6] << 8) | b[pos + 7]);
    if (pos + 8 + len > b.size()) break;
    int st; memcpy(&st, &b[pos + 8], 4);
    if (len >= 20 && (st == 1 || st == 11 || st == 21)) { double x, y; memcpy(&x, &b[pos + 12], 8); memcpy(&y, &b[pos + 20], 8); L->x.push_back(x); L->y.push_back(y); }
    pos += 8 + len;
  }
  return (OGRDataSourceH)L;
}
OGRLayerH OGR_DS_GetLayer(OGRDataSourceH ds, int i) { return i == 0 ? (OGRLayerH)ds : NULL; }
OGRLayerH OGR_DS_GetLayerByName(OGRDataSourceH ds, const char* nm) { return (ds && nm && ((ShimLayer*)ds)->name == nm) ? (OGRLayerH)ds : NULL; }
int OGR_DS_GetLayerCount(OGRDataSourceH ds) { return ds ? 1 : 0; }
void OGR_DS_Destroy(OGRDataSourceH ds) { delete (ShimLayer*)ds; }
const char* OGR_L_GetName(OGRLayerH l) { return l ? ((ShimLayer*)l)->name.c_str() : ""; }
OGRwkbGeometryType OGR_L_GetGeomType(OGRLayerH l) {
  const int t = l ? ((ShimLayer*)l)->type : 0;
  return (t == 1 || t == 11 || t == 21) ? wkbPoint : wkbUnknown;
}
OGRSpatialReferenceH OGR_L_GetSpatialRef(OGRLayerH) { return NULL; }
GIntBig OGR_L_GetFeatureCount(OGRLayerH l, int) { return l ? (GIntBig)((ShimLayer*)l)->x.size() : 0; }
OGRFeatureDefnH OGR_L_GetLayerDefn(OGRLayerH) { return NULL; }
void OGR_L_ResetReading(OGRLayerH l) { if (l) ((ShimLayer*)l)->next = 0; }
OGRFeatureH OGR_L_GetNextFeature(OGRLayerH l) {
  ShimLayer* L = (ShimLayer*)l;
  if (!L || L->next >= L->x.size()) return NULL;
  ShimFeature* f = new ShimFeature{L->x[L->next], L->y[L->next]};
  ++L->next;
  return (OGRFeatureH)f;
}
OGRFeatureH OGR_L_GetFeature(OGRLayerH, GIntBig) { return NULL; }
OGRGeometryH OGR_F_GetGeometryRef(OGRFeatureH f) { return (OGRGeometryH)f; }
int OGR_F_GetFieldIndex(OGRFeatureH, const char*) { return -1; }
int OGR_F_GetFieldAsInteger(OGRFeatureH, int) { return 0; }
GIntBig OGR_F_GetFieldAsInteger64(OGRFeatureH, int) { return 0; }
double OGR_F_GetFieldAsDouble(OGRFeatureH, int) { return 0; }
const char* OGR_F_GetFieldAsString(OGRFeatureH, int) { return ""; }
void OGR_F_Destroy(OGRFeatureH f) { delete (ShimFeature*)f; }
OGRFieldDefnH OGR_FD_GetFieldDefn(OGRFeatureDefnH, int) { return NULL; }
OGRFieldType OGR_Fld_GetType(OGRFieldDefnH) { return OFTInteger; }
double OGR_G_GetX(OGRGeometryH g, int) { return g ? ((ShimFeature*)g)->x : 0; }
double OGR_G_GetY(OGRGeometryH g, int) { return g ? ((ShimFeature*)g)->y : 0; }
}
