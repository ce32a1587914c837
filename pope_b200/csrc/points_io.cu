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
