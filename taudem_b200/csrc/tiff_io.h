// Native TIFF / BigTIFF raster I/O for the TauDEM hot path (GDAL-free).
//
// Implements the file contract of the reference's raster layer
// (reference: src/tiffIO.cpp:54-185 open/header, :245-259 window read,
// :263-428 create/write, :434-445 geotoLength, :449-541 compareTiff,
// :580-598 geo<->grid coordinates) without GDAL: single-band rasters, classic
// TIFF or BigTIFF, strips or tiles, compression none / LZW / Deflate,
// predictors 1/2/3, integer and floating sample formats converted to the
// caller's working type the way GDALRasterIO does (round-to-nearest + clamp
// when narrowing to integers).  GeoTIFF tags and GDAL_NODATA pass through.
#pragma once
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace tdio {

enum DType { DT_I16 = 0, DT_I32 = 1, DT_F32 = 2 };  // SHORT_TYPE, LONG_TYPE, FLOAT_TYPE

inline int dtype_bytes(DType t) { return t == DT_I16 ? 2 : 4; }

struct RawTag {           // a TIFF tag kept verbatim for pass-through
  uint16_t type = 0;      // TIFF field type
  uint64_t count = 0;
  std::vector<uint8_t> data;  // little-endian payload
};

// Georeferencing + nodata carried from an input raster to the outputs derived
// from it (the reference copies projection + geotransform: tiffIO.cpp:343-352).
struct GeoInfo {
  double gt[6] = {0, 1, 0, 0, 0, -1};  // GDAL-style geotransform
  bool is_geographic = false;          // GTModelTypeGeoKey == 2
  std::map<uint16_t, RawTag> geotags;  // 33550 33922 34264 34735 34736 34737
};

class Raster {
 public:
  Raster() = default;
  ~Raster();
  Raster(const Raster&) = delete;
  Raster& operator=(const Raster&) = delete;

  // Opens an existing file; returns false (and sets err) on failure.
  bool open(const std::string& path, std::string* err);

  uint32_t width() const { return width_; }
  uint32_t height() const { return height_; }
  bool has_nodata() const { return has_nodata_; }
  // Nodata as tiffIO reports it: file value, else -9999 (tiffIO.cpp:162-167).
  double nodata() const { return has_nodata_ ? nodata_ : -9999.0; }
  const GeoInfo& geo() const { return geo_; }
  int bits() const { return bits_; }
  int sample_format() const { return sample_format_; }
  int compression() const { return compression_; }
  bool tiled() const { return tiled_; }
  uint64_t block_offset(size_t i) const { return offsets_[i]; }
  uint32_t rows_per_block() const { return block_h_; }

  // Per-row metric cell sizes (tiffIO.cpp:118-151): constant for projected
  // grids, WGS84 ellipsoid lengths per row for geographic grids.
  void cell_sizes(std::vector<double>* dxc, std::vector<double>* dyc) const;

  // Window read with type conversion (tiffIO.cpp:245-259).  dest is row-major
  // nrows x ncols of `type`, row stride `dest_stride` elements (0 = ncols).
  bool read(long xstart, long ystart, long nrows, long ncols, void* dest,
            DType type, std::string* err, long dest_stride = 0);

 private:
  bool load_block(uint64_t idx, std::vector<uint8_t>* out, std::string* err);
  FILE* fp_ = nullptr;
  bool big_ = false, swap_ = false;
  uint32_t width_ = 0, height_ = 0;
  int bits_ = 0, sample_format_ = 1, compression_ = 1, predictor_ = 1;
  bool tiled_ = false;
  uint32_t block_w_ = 0, block_h_ = 0;  // tile size, or (width, rows per strip)
  std::vector<uint64_t> offsets_, counts_;
  bool has_nodata_ = false;
  double nodata_ = 0;
  GeoInfo geo_;
};

// Streaming writer: header first, then rows in any order (strips of fixed
// height).  compression: 1 = none, 5 = LZW (the reference asks GDAL for LZW,
// tiffIO.cpp:316-318), 8 = Deflate.  BigTIFF is chosen by the reference's rule
// (tiffIO.cpp:322-330: cellbytes*X*Y/1e9 > 4.0) or when forced.
class Writer {
 public:
  Writer() = default;
  ~Writer();
  bool create(const std::string& path, uint32_t width, uint32_t height, DType type,
              double nodata, const GeoInfo& geo, int compression, std::string* err,
              bool force_bigtiff = false);
  // Writes rows [ystart, ystart+nrows) from src (row stride src_stride elements,
  // 0 = width).  Rows must be delivered in strip-aligned order for compressed
  // output (any order when uncompressed).
  bool write_rows(long ystart, long nrows, const void* src, std::string* err,
                  long src_stride = 0);
  bool close(std::string* err);

 private:
  FILE* fp_ = nullptr;
  bool big_ = false;
  uint32_t width_ = 0, height_ = 0, rows_per_strip_ = 0;
  DType type_ = DT_F32;
  int compression_ = 1;
  uint64_t data_start_ = 0, offsets_pos_ = 0, counts_pos_ = 0, append_pos_ = 0;
  std::vector<uint64_t> offsets_, counts_;
  std::vector<uint8_t> pending_;  // partial strip buffer for compressed output
  long pending_row0_ = -1;
  std::vector<std::pair<uint64_t, std::vector<uint8_t>>> batch_;  // finished raw strips awaiting compression
  bool flush_batch(std::string* err);
};

// The reference's output naming rule (tiffIO.cpp:268-306): known extensions
// keep their name (only .tif/.tiff are written natively here), a missing
// extension gets ".tif", an unknown one is replaced by "tif".
std::string output_path_rule(const std::string& name);

// Equivalent of tiffIO::compareTiff (tiffIO.cpp:449-541).
bool compare_rasters(const Raster& a, const std::string& aname, const Raster& b,
                     const std::string& bname);

// LZW / Deflate codecs (exposed for tests).
bool lzw_decode(const uint8_t* in, size_t n, std::vector<uint8_t>* out, size_t expect);
void lzw_encode(const uint8_t* in, size_t n, std::vector<uint8_t>* out);

}  // namespace tdio
