// points_io.cu -- the on-disk match format either side of the hot path.  Host code only.
//
// The reference stores the matches of every image pair with numpy.savetxt (linemod.py:168-171: pre_bbox, mkpts0,
// mkpts1, pre_K under data/<set>-points/<object>/{pre_bbox,mkpts0,mkpts1,pre_K}/<pair>.txt) and the pose regressor reads
// them back with numpy.loadtxt (pose/dataset.py).  numpy.savetxt(path, a) writes every value as '%.18e' % float(v),
// columns separated by one space, rows terminated by '\n' (a 1-D array is written one value per line).  glibc's printf
// performs the same correctly rounded binary -> decimal conversion as CPython, so the files produced here are
// byte-identical to numpy's (tests/test_points_io.py).  The batched entry point writes the files of all pairs of a batch
// straight from the pipeline's per-pair output slots on a pool of host threads.
#include <errno.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <charconv>
#include <cmath>
#include <string>
#include <thread>
#include <vector>

#include "pope_b200.h"

namespace {

// '%.18e' % float(v) as CPython prints it: correctly rounded 19 significant digits, exponent of at least two digits, 'inf' /
// '-inf' / 'nan' (CPython drops the sign of a NaN).  std::to_chars (Ryu printf) is the same correctly rounded conversion as
// glibc's printf at a fraction of the cost; sign + d.dddddddddddddddddd + e+XXX <= 27 characters.
inline int format_e18(char* out, double v) {
  if (std::isnan(v)) { memcpy(out, "nan", 3); return 3; }
  const std::to_chars_result r = std::to_chars(out, out + 40, v, std::chars_format::scientific, 18);
  return int(r.ptr - out);
}

template <typename T>
int write_txt(const char* path, const T* data, int64_t rows, int cols) {
  FILE* f = fopen(path, "wb");
  if (!f) return POPE_ERR_IO;
  std::vector<char> buf;
  buf.reserve(size_t(1) << 16);
  char tmp[40];
  int rc = POPE_OK;
  for (int64_t r = 0; r < rows && rc == POPE_OK; ++r) {
    for (int c = 0; c < cols; ++c) {
      const int n = format_e18(tmp, double(data[r * cols + c]));
      if (c) buf.push_back(' ');
      buf.insert(buf.end(), tmp, tmp + n);
    }
    buf.push_back('\n');
    if (buf.size() >= (size_t(1) << 16) - 64 * size_t(cols > 0 ? 1 : 0) - 64) {
      if (fwrite(buf.data(), 1, buf.size(), f) != buf.size()) rc = POPE_ERR_IO;
      buf.clear();
    }
  }
  if (rc == POPE_OK && !buf.empty() && fwrite(buf.data(), 1, buf.size(), f) != buf.size()) rc = POPE_ERR_IO;
  if (fclose(f) != 0 && rc == POPE_OK) rc = POPE_ERR_IO;
  return rc;
}

bool make_dir(const std::string& p) { return mkdir(p.c_str(), 0777) == 0 || errno == EEXIST; }

// numpy.loadtxt(path, delimiter=' ') for the files above: one row per line, values separated by blanks, '#' starts a
// comment, empty lines are skipped.  std::from_chars is the correctly rounded decimal -> binary conversion (CPython's float()
// is too), so values written with '%.18e' come back bit for bit.  Returns POPE_ERR_SHAPE for a ragged file or a token that is
// not a number.
int parse_txt(const char* path, std::vector<double>& vals, int64_t* rows, int* cols) {
  FILE* f = fopen(path, "rb");
  if (!f) return POPE_ERR_IO;
  std::vector<char> buf;
  char chunk[1 << 16];
  size_t got;
  while ((got = fread(chunk, 1, sizeof(chunk), f)) > 0) buf.insert(buf.end(), chunk, chunk + got);
  const bool bad = ferror(f) != 0;
  fclose(f);
  if (bad) return POPE_ERR_IO;
  *rows = 0;
  *cols = 0;
  const char* p = buf.data();
  const char* end = p + buf.size();
  while (p < end) {
    const char* eol = static_cast<const char*>(memchr(p, '\n', size_t(end - p)));
    const char* line_end = eol ? eol : end;
    const char* stop = static_cast<const char*>(memchr(p, '#', size_t(line_end - p)));
    if (!stop) stop = line_end;
    int n = 0;
    while (p < stop) {
      while (p < stop && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
      if (p >= stop) break;
      if (*p == '+') ++p;                                   // from_chars does not take a leading plus, float() does
      double v;
      const std::from_chars_result r = std::from_chars(p, stop, v);
      if (r.ec == std::errc::result_out_of_range) {         // float() overflows to inf and underflows to 0 silently
        const char* q = p;
        const bool neg = *q == '-';
        bool big = true;                                    // decide by the sign of the decimal exponent
        for (; q < r.ptr; ++q) if ((*q == 'e' || *q == 'E') && q + 1 < r.ptr && q[1] == '-') big = false;
        v = big ? (neg ? -HUGE_VAL : HUGE_VAL) : (neg ? -0.0 : 0.0);
      } else if (r.ec != std::errc()) {
        return POPE_ERR_SHAPE;
      }
      if (r.ptr < stop && !(*r.ptr == ' ' || *r.ptr == '\t' || *r.ptr == '\r')) return POPE_ERR_SHAPE;
      vals.push_back(v);
      ++n;
      p = r.ptr;
    }
    if (n > 0) {
      if (*rows == 0) *cols = n;
      else if (n != *cols) return POPE_ERR_SHAPE;
      ++*rows;
    }
    p = eol ? eol + 1 : end;
  }
  return POPE_OK;
}

}  // namespace

extern "C" int pope_savetxt_f32(const char* path, const float* data, int64_t rows, int cols) {
  if (!path || (!data && rows > 0) || rows < 0 || cols <= 0) return POPE_ERR_INVALID_ARG;
  return write_txt<float>(path, data, rows, cols);
}

extern "C" int pope_savetxt_f64(const char* path, const double* data, int64_t rows, int cols) {
  if (!path || (!data && rows > 0) || rows < 0 || cols <= 0) return POPE_ERR_INVALID_ARG;
  return write_txt<double>(path, data, rows, cols);
}

// ---- PNG crops (linemod.py:172-173 writes them with cv2.imwrite, pose/dataset.py:102-103 reads them with cv2.imread) --------
namespace {

void put_be32(std::vector<unsigned char>& v, uint32_t x) {
  v.push_back((unsigned char)(x >> 24)); v.push_back((unsigned char)(x >> 16)); v.push_back((unsigned char)(x >> 8)); v.push_back((unsigned char)x);
}
void put_chunk(std::vector<unsigned char>& out, const char* type, const unsigned char* data, size_t n) {
  put_be32(out, uint32_t(n));
  const size_t at = out.size();
  out.insert(out.end(), type, type + 4);
  if (n) out.insert(out.end(), data, data + n);
  put_be32(out, uint32_t(crc32(0L, out.data() + at, uInt(n + 4))));
}

// 8-bit image in OpenCV's memory order ([h, w, c] with c = 1 grey, 3 BGR, 4 BGRA; row pitch in bytes) -> PNG file: channels
// swapped to PNG's RGB(A) order, every scanline predicted from its left neighbour ('Sub' filter), one zlib stream.
// raw / z / out: scratch owned by the calling thread and reused from image to image (fresh megabyte-sized allocations per
// image serialise the threads of the batch writer in the kernel's page-fault path)
int write_png(const char* path, const unsigned char* img, int h, int w, int c, int64_t pitch, int level,
              std::vector<unsigned char>& raw, std::vector<unsigned char>& z, std::vector<unsigned char>& out) {
  if (h <= 0 || w <= 0 || (c != 1 && c != 3 && c != 4)) return POPE_ERR_SHAPE;
  const size_t row = size_t(w) * c;
  if (raw.size() < (row + 1) * size_t(h)) raw.resize((row + 1) * size_t(h));
  const size_t raw_len = (row + 1) * size_t(h);
  for (int y = 0; y < h; ++y) {
    const unsigned char* src = img + size_t(y) * pitch;
    unsigned char* dst = raw.data() + (row + 1) * size_t(y);
    *dst++ = 1;                                             // filter type Sub
    if (c == 3) {                                           // BGR -> RGB, the common case
      unsigned char pr = 0, pg = 0, pb = 0;
      for (int x = 0; x < w; ++x) {
        const unsigned char b = src[3 * x], g = src[3 * x + 1], r = src[3 * x + 2];
        dst[3 * x] = (unsigned char)(r - pr); dst[3 * x + 1] = (unsigned char)(g - pg); dst[3 * x + 2] = (unsigned char)(b - pb);
        pr = r; pg = g; pb = b;
      }
    } else {
      for (int x = 0; x < w; ++x)
        for (int k = 0; k < c; ++k) {
          const int ks = (c == 4 && k < 3) ? 2 - k : k;     // BGRA -> RGBA
          const unsigned char cur = src[size_t(x) * c + ks], left = x ? src[size_t(x - 1) * c + ks] : 0;
          dst[size_t(x) * c + k] = (unsigned char)(cur - left);
        }
    }
  }
  // run-length strategy: what OpenCV's encoder uses by default, fast on filtered scanlines
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  if (deflateInit2(&zs, level, Z_DEFLATED, 15, 8, Z_RLE) != Z_OK) return POPE_ERR_IO;
  uLongf zlen = deflateBound(&zs, uLong(raw_len));
  if (z.size() < zlen) z.resize(zlen);
  zs.next_in = raw.data(); zs.avail_in = uInt(raw_len);
  zs.next_out = z.data(); zs.avail_out = uInt(zlen);
  const int zrc = deflate(&zs, Z_FINISH);
  zlen = zs.total_out;
  deflateEnd(&zs);
  if (zrc != Z_STREAM_END) return POPE_ERR_IO;
  out.clear();
  out.reserve(zlen + 64);
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  out.insert(out.end(), sig, sig + 8);
  std::vector<unsigned char> hdr;
  put_be32(hdr, uint32_t(w));
  put_be32(hdr, uint32_t(h));
  const unsigned char tail[5] = {8, (unsigned char)(c == 1 ? 0 : c == 3 ? 2 : 6), 0, 0, 0};   // depth, colour type, deflate, filter, no interlace
  hdr.insert(hdr.end(), tail, tail + 5);
  put_chunk(out, "IHDR", hdr.data(), hdr.size());
  put_chunk(out, "IDAT", z.data(), zlen);
  put_chunk(out, "IEND", nullptr, 0);
  FILE* f = fopen(path, "wb");
  if (!f) return POPE_ERR_IO;
  const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
  return (fclose(f) == 0 && ok) ? POPE_OK : POPE_ERR_IO;
}

}  // namespace

extern "C" int pope_write_png(const char* path, const unsigned char* img, int height, int width, int channels, int64_t pitch,
                              int level) {
  if (!path || !img || pitch < int64_t(width) * channels || level < 0 || level > 9) return POPE_ERR_INVALID_ARG;
  std::vector<unsigned char> raw, z, out;
  return write_png(path, img, height, width, channels, pitch, level, raw, z, out);
}

extern "C" int pope_write_png_batch(const char* const* paths, const unsigned char* const* imgs, const int32_t* heights,
                                    const int32_t* widths, int channels, int n, int level, int n_threads) {
  if (!paths || !imgs || !heights || !widths || n < 0 || level < 0 || level > 9) return POPE_ERR_INVALID_ARG;
  if (n_threads <= 0) n_threads = int(sysconf(_SC_NPROCESSORS_ONLN));
  if (n_threads <= 0) n_threads = 1;
  if (n_threads > 64) n_threads = 64;
  if (n_threads > n) n_threads = n > 0 ? n : 1;
  std::atomic<int> next(0), status(POPE_OK);
  auto work = [&]() {
    std::vector<unsigned char> raw, z, out;
    for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) {
      if (!paths[i] || !imgs[i]) { status.store(POPE_ERR_INVALID_ARG); continue; }
      const int rc = write_png(paths[i], imgs[i], heights[i], widths[i], channels, int64_t(widths[i]) * channels, level, raw, z,
                               out);
      if (rc != POPE_OK) status.store(rc);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  return status.load();
}

extern "C" int pope_loadtxt_f64(const char* path, double* out, int64_t capacity, int64_t* rows, int* cols) {
  if (!path || !rows || !cols || capacity < 0 || (!out && capacity > 0)) return POPE_ERR_INVALID_ARG;
  std::vector<double> vals;
  const int rc = parse_txt(path, vals, rows, cols);
  if (rc != POPE_OK) return rc;
  if (int64_t(vals.size()) > capacity) return POPE_ERR_CAPACITY;
  if (!vals.empty()) memcpy(out, vals.data(), vals.size() * sizeof(double));
  return POPE_OK;
}

extern "C" int pope_read_match_files(const char* dir, const char* const* names, int n_pairs, float* mkpts0, float* mkpts1,
                                     int32_t* counts, int64_t capacity, int n_threads) {
  if (!dir || !names || !mkpts0 || !mkpts1 || !counts || n_pairs < 0 || capacity < 0) return POPE_ERR_INVALID_ARG;
  const std::string root(dir);
  if (n_threads <= 0) n_threads = int(sysconf(_SC_NPROCESSORS_ONLN));
  if (n_threads <= 0) n_threads = 1;
  if (n_threads > 64) n_threads = 64;
  if (n_threads > n_pairs) n_threads = n_pairs > 0 ? n_pairs : 1;
  std::atomic<int> next(0), status(POPE_OK);
  auto work = [&]() {
    std::vector<double> a, b;
    for (int p = next.fetch_add(1); p < n_pairs; p = next.fetch_add(1)) {
      counts[p] = -1;                                       // no such pair on disk (linemod.py skips pairs below 5 matches)
      if (!names[p]) continue;
      a.clear();
      b.clear();
      int64_t ra = 0, rb = 0;
      int ca = 0, cb = 0;
      const int rc0 = parse_txt((root + "/mkpts0/" + names[p] + ".txt").c_str(), a, &ra, &ca);
      if (rc0 == POPE_ERR_IO) continue;
      const int rc1 = parse_txt((root + "/mkpts1/" + names[p] + ".txt").c_str(), b, &rb, &cb);
      if (rc0 != POPE_OK || rc1 != POPE_OK || ra != rb || (ra > 0 && (ca != 2 || cb != 2))) {
        status.store(rc1 == POPE_ERR_IO ? POPE_ERR_IO : POPE_ERR_SHAPE);
        continue;
      }
      const int64_t m = ra > capacity ? capacity : ra;
      for (int64_t i = 0; i < m * 2; ++i) {
        mkpts0[size_t(p) * capacity * 2 + i] = float(a[i]);
        mkpts1[size_t(p) * capacity * 2 + i] = float(b[i]);
      }
      counts[p] = int32_t(m);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  return status.load();
}

extern "C" int pope_write_match_files(const char* dir, const char* const* names, int n_pairs, const float* mkpts0,
                                      const float* mkpts1, const int32_t* counts, int64_t capacity, int min_matches,
                                      int n_threads, int32_t* written) {
  if (!dir || !names || !mkpts0 || !mkpts1 || !counts || n_pairs < 0 || capacity < 0) return POPE_ERR_INVALID_ARG;
  const std::string root(dir);
  if (!make_dir(root) || !make_dir(root + "/mkpts0") || !make_dir(root + "/mkpts1")) return POPE_ERR_IO;
  if (n_threads <= 0) n_threads = int(sysconf(_SC_NPROCESSORS_ONLN));
  if (n_threads <= 0) n_threads = 1;
  if (n_threads > 64) n_threads = 64;
  if (n_threads > n_pairs) n_threads = n_pairs > 0 ? n_pairs : 1;
  std::atomic<int> next(0), status(POPE_OK), done(0);
  auto work = [&]() {
    for (int p = next.fetch_add(1); p < n_pairs; p = next.fetch_add(1)) {
      const int64_t m = counts[p] < 0 ? 0 : (counts[p] > capacity ? capacity : counts[p]);
      if (m < min_matches || !names[p]) continue;          // linemod.py:143-146: pairs with fewer than 5 matches are skipped
      const std::string a = root + "/mkpts0/" + names[p] + ".txt", b = root + "/mkpts1/" + names[p] + ".txt";
      int rc = write_txt<float>(a.c_str(), mkpts0 + size_t(p) * capacity * 2, m, 2);
      if (rc == POPE_OK) rc = write_txt<float>(b.c_str(), mkpts1 + size_t(p) * capacity * 2, m, 2);
      if (rc != POPE_OK) status.store(rc);
      else done.fetch_add(1);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  if (written) *written = done.load();
  return status.load();
}
