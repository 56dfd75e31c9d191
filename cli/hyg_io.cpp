// cli/hyg_io.cpp -- see hyg_io.hpp.  Host only; depends on zlib.
#include "hyg_io.hpp"

#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <cerrno>
#include <charconv>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace hygio {

bool ends_with(const std::string& s, const std::string& suffix) {
  return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}

void mkdirs(const std::string& dir) {
  if (dir.empty() || dir == "." || dir == "/") return;
  struct stat st;
  if (stat(dir.c_str(), &st) == 0) {
    if (!S_ISDIR(st.st_mode)) throw Error("not a directory: " + dir);
    return;
  }
  const size_t slash = dir.find_last_of('/');
  if (slash != std::string::npos && slash > 0) mkdirs(dir.substr(0, slash));
  if (mkdir(dir.c_str(), 0777) != 0 && errno != EEXIST) throw Error("cannot create directory " + dir + ": " + std::strerror(errno));
}

void mkdirs_for_file(const std::string& path) {
  if (path.empty()) return;
  const size_t slash = path.find_last_of('/');
  if (slash == std::string::npos) return;
  mkdirs(path.substr(0, slash));
}

std::string read_text(const std::string& path) {
  gzFile f = gzopen(path.c_str(), "rb");
  if (!f) throw Error("cannot open " + path + ": " + std::strerror(errno));
  gzbuffer(f, 1 << 20);
  std::string out;
  std::vector<char> buf(1 << 22);
  for (;;) {
    const int n = gzread(f, buf.data(), static_cast<unsigned>(buf.size()));
    if (n < 0) {
      int err = 0;
      const char* msg = gzerror(f, &err);
      const std::string m = msg ? msg : "read error";
      gzclose(f);
      throw Error("error reading " + path + ": " + m);
    }
    if (n == 0) break;
    out.append(buf.data(), static_cast<size_t>(n));
  }
  gzclose(f);
  return out;
}

static double parse_cell(const char* b, const char* e, const std::string& path, size_t line) {
  while (b < e && (*b == ' ' || *b == '\t' || *b == '"')) b++;
  while (e > b && (e[-1] == ' ' || e[-1] == '\t' || e[-1] == '"' || e[-1] == '\r')) e--;
  if (b == e) return std::nan("");
  if ((e - b == 2 && b[0] == 'N' && b[1] == 'A') || (e - b == 3 && (b[0] == 'n' || b[0] == 'N') && (b[1] == 'a' || b[1] == 'A') && (b[2] == 'n' || b[2] == 'N')))
    return std::nan("");
  // format(scientific = FALSE) writes hundreds of decimals for tiny probabilities: no fixed-size buffer
  char small[64];
  const size_t n = static_cast<size_t>(e - b);
  std::string big;
  char* tmp = small;
  if (n >= sizeof(small)) { big.assign(b, n); tmp = &big[0]; }
  else { std::memcpy(small, b, n); small[n] = 0; }
  char* end = nullptr;
  const double v = std::strtod(tmp, &end);
  if (end != tmp + n) throw Error(path + ": line " + std::to_string(line) + ": not a number: '" + std::string(tmp, n) + "'");
  return v;
}

Table read_csv_numeric(const std::string& path, bool first_line_is_header) { return read_delimited_numeric(path, first_line_is_header, ','); }

Table read_delimited_numeric(const std::string& path, bool first_line_is_header, char sep) {
  const std::string txt = read_text(path);
  Table t;
  size_t pos = 0, line = 0;
  bool header_done = !first_line_is_header;
  while (pos < txt.size()) {
    size_t eol = txt.find('\n', pos);
    if (eol == std::string::npos) eol = txt.size();
    const char* b = txt.data() + pos;
    const char* e = txt.data() + eol;
    pos = eol + 1;
    line++;
    if (e > b && e[-1] == '\r') e--;
    if (b == e) continue;  // blank line
    if (!header_done) {
      const char* c = b;
      while (c <= e) {
        const char* q = static_cast<const char*>(std::memchr(c, sep, static_cast<size_t>(e - c)));
        if (!q) q = e;
        const char* hb = c;
        const char* he = q;
        while (hb < he && (*hb == '"' || *hb == ' ')) hb++;
        while (he > hb && (he[-1] == '"' || he[-1] == ' ')) he--;
        t.header.emplace_back(hb, he);
        c = q + 1;
      }
      t.cols = t.header.size();
      header_done = true;
      continue;
    }
    size_t ncol = 0;
    const char* c = b;
    while (c <= e) {
      const char* q = static_cast<const char*>(std::memchr(c, sep, static_cast<size_t>(e - c)));
      if (!q) q = e;
      t.v.push_back(parse_cell(c, q, path, line));
      ncol++;
      c = q + 1;
    }
    if (t.cols == 0) t.cols = ncol;
    if (ncol != t.cols) throw Error(path + ": line " + std::to_string(line) + " has " + std::to_string(ncol) + " fields, expected " + std::to_string(t.cols));
    t.rows++;
  }
  return t;
}

Writer::Writer(const std::string& path) : path_(path) {
  if (ends_with(path, ".gz")) {
    gz_ = gzopen(path.c_str(), "wb6");
    if (!gz_) throw Error("cannot open " + path + " for writing: " + std::strerror(errno));
    gzbuffer(static_cast<gzFile>(gz_), 1 << 20);
  } else {
    fp_ = std::fopen(path.c_str(), "wb");
    if (!fp_) throw Error("cannot open " + path + " for writing: " + std::strerror(errno));
  }
}
Writer::~Writer() {
  try { close(); } catch (...) {}
}
void Writer::write(const std::string& s) {
  if (s.empty()) return;
  if (gz_) {
    size_t off = 0;
    while (off < s.size()) {
      const unsigned chunk = static_cast<unsigned>(std::min<size_t>(s.size() - off, 1u << 30));
      if (gzwrite(static_cast<gzFile>(gz_), s.data() + off, chunk) <= 0) throw Error("write error on " + path_);
      off += chunk;
    }
  } else if (std::fwrite(s.data(), 1, s.size(), static_cast<FILE*>(fp_)) != s.size()) {
    throw Error("write error on " + path_);
  }
}
void Writer::close() {
  if (gz_) {
    const int rc = gzclose(static_cast<gzFile>(gz_));
    gz_ = nullptr;
    if (rc != Z_OK) throw Error("close error on " + path_);
  }
  if (fp_) {
    const int rc = std::fclose(static_cast<FILE*>(fp_));
    fp_ = nullptr;
    if (rc != 0) throw Error("close error on " + path_);
  }
}

// ---- R's format(x, scientific = FALSE) --------------------------------------------------------------------------------
namespace {
constexpr int KP_MAX = 27;
const long double kTbl[KP_MAX + 1] = {1e0L,  1e1L,  1e2L,  1e3L,  1e4L,  1e5L,  1e6L,  1e7L,  1e8L,  1e9L,  1e10L, 1e11L, 1e12L, 1e13L,
                                      1e14L, 1e15L, 1e16L, 1e17L, 1e18L, 1e19L, 1e20L, 1e21L, 1e22L, 1e23L, 1e24L, 1e25L, 1e26L, 1e27L};

// for |x| = alpha * 10^kpower, 1 <= alpha < 10: kpower, the significant digits needed (<= digits) and whether rounding to
// `digits` significant digits widens the fixed representation (R src/main/format.c: scientific())
void r_scientific(double x, int digits, int& neg, int& kpower, int& nsig, bool& widens) {
  if (x == 0.0) { kpower = 0; nsig = 1; neg = 0; widens = false; return; }
  neg = x < 0.0;
  const double r = std::fabs(x);
  int kp = static_cast<int>(std::floor(std::log10(r))) - digits + 1;
  long double r_prec = r;
  if (std::abs(kp) < 10) {
    if (kp > 0) r_prec /= kTbl[kp];
    else if (kp < 0) r_prec *= kTbl[-kp];
  } else if (kp <= -308) {
    r_prec = (r * 1e+303L) / powl(10.0L, kp + 303);
  } else {
    r_prec /= powl(10.0L, kp);
  }
  if (r_prec < kTbl[digits - 1]) { r_prec *= 10.0L; kp--; }
  double alpha = static_cast<double>(nearbyintl(r_prec));
  nsig = digits;
  for (int j = 1; j <= digits; j++) {
    alpha /= 10.0;
    if (alpha == std::floor(alpha)) nsig--;
    else break;
  }
  if (nsig == 0 && digits > 0) { nsig = 1; kp += 1; }
  kpower = kp + digits - 1;
  widens = kpower > 0 && kpower <= KP_MAX && r < static_cast<double>(kTbl[kpower]);
}
}  // namespace

std::vector<std::string> r_format_fixed(const std::vector<double>& x, int digits) {
  const int scipen = 100;  // format.default(scientific = FALSE)
  int rgt = INT_MIN, mxl = INT_MIN, mxsl = INT_MIN, mxns = INT_MIN, mxe = INT_MIN, mne = INT_MAX, neg = 0;
  bool naflag = false, nanflag = false, posinf = false, neginf = false;
  for (double v : x) {
    if (!std::isfinite(v)) {
      if (std::isnan(v)) nanflag = true;
      else if (v > 0) posinf = true;
      else neginf = true;
      continue;
    }
    int neg_i, kpower, nsig;
    bool widens;
    r_scientific(v, digits, neg_i, kpower, nsig, widens);
    int left = kpower + 1;
    if (widens) left--;
    const int sleft = neg_i + ((left <= 0) ? 1 : left);
    const int right = nsig - left;
    if (neg_i) neg = 1;
    rgt = std::max(rgt, right);
    mxl = std::max(mxl, left);
    mxsl = std::max(mxsl, sleft);
    mxns = std::max(mxns, nsig);
    mxe = std::max(mxe, kpower);
    mne = std::min(mne, kpower);
  }
  (void)naflag;
  int w = 0, d = 0, e = 0;
  if (mxl != INT_MIN) {
    if (mxl < 0) mxsl = 1 + neg;
    if (rgt < 0) rgt = 0;
    const int wF = mxsl + rgt + (rgt != 0);
    e = (mxe >= 100 || mne <= -99) ? 2 : 1;
    d = mxns - 1;
    w = neg + (d > 0) + d + 4 + e;
    if (wF <= w + scipen) { e = 0; d = rgt; w = wF; }
  }
  if (nanflag && w < 3) w = 3;
  if (posinf && w < 3) w = 3;
  if (neginf && w < 4) w = 4;
  std::vector<std::string> out;
  out.reserve(x.size());
  std::vector<char> buf(1024 + static_cast<size_t>(std::max(0, w)));
  for (double v : x) {
    if (std::isnan(v)) std::snprintf(buf.data(), buf.size(), "%*s", w, "NaN");
    else if (!std::isfinite(v)) std::snprintf(buf.data(), buf.size(), "%*s", w, v > 0 ? "Inf" : "-Inf");
    else if (e) std::snprintf(buf.data(), buf.size(), d ? "%#*.*e" : "%*.*e", w, d, v);
    else std::snprintf(buf.data(), buf.size(), "%*.*f", w, d, v);
    out.emplace_back(buf.data());
  }
  return out;
}

// shortest round-trip digits of |x| > 0 and the decimal exponent of the first digit
static void shortest_digits(double x, std::string& digits, int& exp10) {
  char buf[64];
  const auto res = std::to_chars(buf, buf + sizeof(buf), std::fabs(x), std::chars_format::scientific);
  const std::string s(buf, res.ptr);
  const size_t epos = s.find('e');
  digits.clear();
  for (size_t i = 0; i < epos; i++)
    if (s[i] != '.') digits.push_back(s[i]);
  exp10 = std::atoi(s.c_str() + epos + 1);
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
}

std::string readr_double(double x) {
  // readr/vroom write the shortest digits that round-trip (grisu3).  Layout: plain decimal notation unless the exponent
  // form is strictly shorter (or the number is an integer with more than four trailing zeros), exponent without padding.
  // Values are exact; for corner cases the layout of readr itself could not be checked here (no R in this image).
  if (std::isnan(x)) return "NA";
  if (!std::isfinite(x)) return x > 0 ? "Inf" : "-Inf";
  if (x == 0.0) return "0";
  std::string dg;
  int e10;
  shortest_digits(x, dg, e10);
  const int len = static_cast<int>(dg.size());
  const int d_exp = e10 - (len - 1);  // value = dg * 10^d_exp
  const std::string sign = x < 0 ? "-" : "";
  std::string sci = dg.substr(0, 1);
  if (len > 1) sci += "." + dg.substr(1);
  sci += "e" + std::to_string(e10);
  if (d_exp >= 0) return (d_exp <= 4) ? sign + dg + std::string(static_cast<size_t>(d_exp), '0') : sign + sci;
  const std::string fixed = fixed_double(std::fabs(x));
  return sign + (fixed.size() <= sci.size() ? fixed : sci);
}

std::string fixed_double(double x) {
  if (std::isnan(x)) return "NA";
  if (!std::isfinite(x)) return x > 0 ? "Inf" : "-Inf";
  if (x == 0.0) return "0";
  std::string dg;
  int e10;
  shortest_digits(x, dg, e10);
  const int len = static_cast<int>(dg.size());
  const int decpt = e10 + 1;
  std::string s = x < 0 ? "-" : "";
  if (decpt <= 0) s += "0." + std::string(static_cast<size_t>(-decpt), '0') + dg;
  else if (decpt >= len) s += dg + std::string(static_cast<size_t>(decpt - len), '0');
  else s += dg.substr(0, static_cast<size_t>(decpt)) + "." + dg.substr(static_cast<size_t>(decpt));
  return s;
}

std::string py_repr_double(double x) {
  if (std::isnan(x)) return "nan";
  if (!std::isfinite(x)) return x > 0 ? "inf" : "-inf";
  if (x == 0.0) return std::signbit(x) ? "-0.0" : "0.0";
  std::string dg;
  int e10;
  shortest_digits(x, dg, e10);
  const int len = static_cast<int>(dg.size());
  const int decpt = e10 + 1;
  std::string s = x < 0 ? "-" : "";
  if (decpt > 16 || decpt < -3) {
    s += dg.substr(0, 1);
    if (len > 1) s += "." + dg.substr(1);
    char eb[16];
    std::snprintf(eb, sizeof(eb), "e%c%02d", e10 < 0 ? '-' : '+', std::abs(e10));
    s += eb;
  } else if (decpt <= 0) {
    s += "0." + std::string(static_cast<size_t>(-decpt), '0') + dg;
  } else if (decpt >= len) {
    s += dg + std::string(static_cast<size_t>(decpt - len), '0') + ".0";
  } else {
    s += dg.substr(0, static_cast<size_t>(decpt)) + "." + dg.substr(static_cast<size_t>(decpt));
  }
  return s;
}

void savetxt_e18(const std::string& path, const double* a, size_t rows, size_t cols) {
  Writer w(path);
  std::string chunk;
  chunk.reserve(1 << 20);
  char buf[64];
  for (size_t r = 0; r < rows; r++) {
    for (size_t c = 0; c < cols; c++) {
      const int n = std::snprintf(buf, sizeof(buf), "%.18e", a[r * cols + c]);
      if (c) chunk.push_back(',');
      chunk.append(buf, static_cast<size_t>(n));
    }
    chunk.push_back('\n');
    if (chunk.size() > (1 << 20) - 4096) { w.write(chunk); chunk.clear(); }
  }
  w.write(chunk);
  w.close();
}

// ---- .npy / .npz ------------------------------------------------------------------------------------------------------
static std::string npy_header(const std::string& descr, const std::vector<size_t>& shape) {
  std::string dict = "{'descr': '" + descr + "', 'fortran_order': False, 'shape': (";
  for (size_t i = 0; i < shape.size(); i++) {
    dict += std::to_string(shape[i]);
    if (shape.size() == 1 || i + 1 < shape.size()) dict += ",";
    if (i + 1 < shape.size()) dict += " ";
  }
  dict += "), }";
  // magic (6) + version (2) + header length (2) + dict + padding + '\n' is a multiple of 64
  size_t total = 10 + dict.size() + 1;
  const size_t pad = (64 - total % 64) % 64;
  dict += std::string(pad, ' ');
  dict += "\n";
  std::string h = "\x93NUMPY";
  h.push_back('\x01');
  h.push_back('\x00');
  const uint16_t hl = static_cast<uint16_t>(dict.size());
  h.push_back(static_cast<char>(hl & 0xFF));
  h.push_back(static_cast<char>(hl >> 8));
  return h + dict;
}

static void put16(std::string& s, uint16_t v) { s.push_back(static_cast<char>(v & 0xFF)); s.push_back(static_cast<char>(v >> 8)); }
static void put32(std::string& s, uint32_t v) { for (int i = 0; i < 4; i++) s.push_back(static_cast<char>((v >> (8 * i)) & 0xFF)); }

void save_npz(const std::string& path, const std::string& descr, const std::vector<size_t>& shape, const void* data, size_t nbytes) {
  const std::string head = npy_header(descr, shape);
  const size_t usize = head.size() + nbytes;
  if (usize >= 0xFFFFFFFFull) throw Error("save_npz: member larger than 4 GiB is not supported: " + path);
  uLong crc = crc32(0L, Z_NULL, 0);
  crc = crc32(crc, reinterpret_cast<const Bytef*>(head.data()), static_cast<uInt>(head.size()));
  {
    const Bytef* p = static_cast<const Bytef*>(data);
    size_t left = nbytes;
    while (left) {
      const uInt n = static_cast<uInt>(std::min<size_t>(left, 1u << 30));
      crc = crc32(crc, p, n);
      p += n;
      left -= n;
    }
  }
  // raw deflate of header + payload
  z_stream zs;
  std::memset(&zs, 0, sizeof(zs));
  if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw Error("deflateInit2 failed");
  std::vector<unsigned char> comp(deflateBound(&zs, static_cast<uLong>(usize)) + 64);
  zs.next_out = comp.data();
  zs.avail_out = static_cast<uInt>(comp.size());
  zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(head.data()));
  zs.avail_in = static_cast<uInt>(head.size());
  if (deflate(&zs, Z_NO_FLUSH) != Z_OK) { deflateEnd(&zs); throw Error("deflate failed"); }
  zs.next_in = static_cast<Bytef*>(const_cast<void*>(data));
  zs.avail_in = static_cast<uInt>(nbytes);
  if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { deflateEnd(&zs); throw Error("deflate failed"); }
  const size_t csize = comp.size() - zs.avail_out;
  deflateEnd(&zs);

  const std::string name = "arr_0.npy";
  std::string lh;
  put32(lh, 0x04034b50u); put16(lh, 20); put16(lh, 0); put16(lh, 8); put16(lh, 0); put16(lh, 0x21);  // 1980-01-01
  put32(lh, static_cast<uint32_t>(crc)); put32(lh, static_cast<uint32_t>(csize)); put32(lh, static_cast<uint32_t>(usize));
  put16(lh, static_cast<uint16_t>(name.size())); put16(lh, 0);
  lh += name;
  std::string cd;
  put32(cd, 0x02014b50u); put16(cd, 20); put16(cd, 20); put16(cd, 0); put16(cd, 8); put16(cd, 0); put16(cd, 0x21);
  put32(cd, static_cast<uint32_t>(crc)); put32(cd, static_cast<uint32_t>(csize)); put32(cd, static_cast<uint32_t>(usize));
  put16(cd, static_cast<uint16_t>(name.size())); put16(cd, 0); put16(cd, 0); put16(cd, 0); put16(cd, 0); put32(cd, 0); put32(cd, 0);
  cd += name;
  std::string eocd;
  put32(eocd, 0x06054b50u); put16(eocd, 0); put16(eocd, 0); put16(eocd, 1); put16(eocd, 1);
  put32(eocd, static_cast<uint32_t>(cd.size())); put32(eocd, static_cast<uint32_t>(lh.size() + csize)); put16(eocd, 0);

  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw Error("cannot open " + path + " for writing: " + std::strerror(errno));
  bool ok = std::fwrite(lh.data(), 1, lh.size(), f) == lh.size();
  ok = ok && std::fwrite(comp.data(), 1, csize, f) == csize;
  ok = ok && std::fwrite(cd.data(), 1, cd.size(), f) == cd.size();
  ok = ok && std::fwrite(eocd.data(), 1, eocd.size(), f) == eocd.size();
  ok = (std::fclose(f) == 0) && ok;
  if (!ok) throw Error("write error on " + path);
}

// ---- np.load of an .npz: first member of the archive, stored or deflated (zip64 sizes in the central directory included) --
static uint16_t get16(const unsigned char* p) { return static_cast<uint16_t>(p[0] | (p[1] << 8)); }
static uint32_t get32(const unsigned char* p) { return static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 8) | (static_cast<uint32_t>(p[2]) << 16) | (static_cast<uint32_t>(p[3]) << 24); }
static uint64_t get64(const unsigned char* p) { return static_cast<uint64_t>(get32(p)) | (static_cast<uint64_t>(get32(p + 4)) << 32); }

NpyArray load_npz(const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw Error("cannot open " + path + ": " + std::strerror(errno));
  std::vector<unsigned char> buf;
  {
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    buf.resize(n > 0 ? static_cast<size_t>(n) : 0);
    const bool ok = buf.empty() || std::fread(buf.data(), 1, buf.size(), f) == buf.size();
    std::fclose(f);
    if (!ok) throw Error("read error on " + path);
  }
  if (buf.size() < 22) throw Error(path + ": not a zip archive");
  size_t eocd = std::string::npos;
  for (size_t i = buf.size() - 22 + 1; i-- > 0;) {
    if (get32(&buf[i]) == 0x06054b50u) { eocd = i; break; }
    if (buf.size() - i > 22 + 65536) break;
  }
  if (eocd == std::string::npos) throw Error(path + ": end of central directory not found");
  uint64_t cd_off = get32(&buf[eocd + 16]);
  if (cd_off == 0xFFFFFFFFull) {   // zip64 end of central directory, through its locator
    if (eocd < 20 || get32(&buf[eocd - 20]) != 0x07064b50u) throw Error(path + ": zip64 locator not found");
    const uint64_t e64 = get64(&buf[eocd - 20 + 8]);
    if (e64 + 56 > buf.size() || get32(&buf[e64]) != 0x06064b50u) throw Error(path + ": bad zip64 end of central directory");
    cd_off = get64(&buf[e64 + 48]);
  }
  if (cd_off + 46 > buf.size() || get32(&buf[cd_off]) != 0x02014b50u) throw Error(path + ": bad central directory");
  const unsigned char* cd = &buf[cd_off];
  const uint16_t method = get16(cd + 10);
  uint64_t csize = get32(cd + 20), usize = get32(cd + 24), lho = get32(cd + 42);
  const uint16_t nlen = get16(cd + 28), xlen = get16(cd + 30);
  {   // zip64 extra field: the 8-byte values appear in the order usize, csize, offset, only for fields that are 0xFFFFFFFF
    const unsigned char* x = cd + 46 + nlen;
    const unsigned char* xe = x + xlen;
    while (x + 4 <= xe) {
      const uint16_t id = get16(x), sz = get16(x + 2);
      if (id == 0x0001) {
        const unsigned char* q = x + 4;
        if (usize == 0xFFFFFFFFull) { usize = get64(q); q += 8; }
        if (csize == 0xFFFFFFFFull) { csize = get64(q); q += 8; }
        if (lho == 0xFFFFFFFFull) { lho = get64(q); q += 8; }
      }
      x += 4 + sz;
    }
  }
  if (lho + 30 > buf.size() || get32(&buf[lho]) != 0x04034b50u) throw Error(path + ": bad local header");
  const size_t data = lho + 30 + get16(&buf[lho + 26]) + get16(&buf[lho + 28]);
  if (data + csize > buf.size()) throw Error(path + ": truncated member");
  std::vector<unsigned char> raw(usize);
  if (method == 0) {
    if (csize != usize) throw Error(path + ": stored member with differing sizes");
    std::memcpy(raw.data(), &buf[data], usize);
  } else if (method == 8) {
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) throw Error("inflateInit2 failed");
    zs.next_in = &buf[data];
    zs.avail_in = static_cast<uInt>(csize);
    zs.next_out = raw.data();
    zs.avail_out = static_cast<uInt>(usize);
    const int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    if (rc != Z_STREAM_END) throw Error(path + ": inflate failed");
  } else {
    throw Error(path + ": unsupported zip compression method " + std::to_string(method));
  }
  // .npy: magic, version, header length, a Python dict literal
  if (raw.size() < 10 || std::memcmp(raw.data(), "\x93NUMPY", 6) != 0) throw Error(path + ": member is not an .npy array");
  size_t hlen, hoff;
  if (raw[6] == 1) { hlen = get16(&raw[8]); hoff = 10; } else { hlen = get32(&raw[8]); hoff = 12; }
  if (hoff + hlen > raw.size()) throw Error(path + ": truncated .npy header");
  const std::string head(reinterpret_cast<const char*>(&raw[hoff]), hlen);
  NpyArray a;
  {
    const size_t d = head.find("'descr'");
    const size_t q0 = head.find('\'', head.find(':', d) + 1), q1 = head.find('\'', q0 + 1);
    if (d == std::string::npos || q0 == std::string::npos || q1 == std::string::npos) throw Error(path + ": no descr in the .npy header");
    a.descr = head.substr(q0 + 1, q1 - q0 - 1);
    if (head.find("'fortran_order': False") == std::string::npos) throw Error(path + ": Fortran-ordered arrays are not supported");
    const size_t s0 = head.find('(', head.find("'shape'")), s1 = head.find(')', s0);
    if (s0 == std::string::npos || s1 == std::string::npos) throw Error(path + ": no shape in the .npy header");
    const std::string sh = head.substr(s0 + 1, s1 - s0 - 1);
    size_t pos = 0;
    while (pos < sh.size()) {
      while (pos < sh.size() && (sh[pos] == ' ' || sh[pos] == ',')) pos++;
      if (pos >= sh.size()) break;
      char* end = nullptr;
      a.shape.push_back(static_cast<size_t>(std::strtoull(sh.c_str() + pos, &end, 10)));
      pos = static_cast<size_t>(end - sh.c_str());
    }
  }
  a.data.assign(raw.begin() + static_cast<long>(hoff + hlen), raw.end());
  size_t n = 1;
  for (size_t d : a.shape) n *= d;
  const size_t item = a.descr.size() >= 3 ? static_cast<size_t>(std::strtoul(a.descr.c_str() + 2, nullptr, 10)) : 0;
  if (item == 0 || a.data.size() != n * item) throw Error(path + ": payload size does not match shape and dtype " + a.descr);
  return a;
}

// ---- pandas.read_csv(sep='\t') of the matrices `hygeia aggregate` writes: a header line, first column = the index ----------
IndexedIntMatrix read_indexed_int_matrix(const std::string& path, char sep) {
  const std::string txt = read_text(path);
  IndexedIntMatrix m;
  size_t pos = 0, line = 0;
  bool header_done = false;
  while (pos < txt.size()) {
    size_t eol = txt.find('\n', pos);
    if (eol == std::string::npos) eol = txt.size();
    const char* b = txt.data() + pos;
    const char* e = txt.data() + eol;
    pos = eol + 1;
    line++;
    if (e > b && e[-1] == '\r') e--;
    if (b == e) continue;
    if (!header_done) {
      size_t n = 0;
      for (const char* c = b; c < e; c++) n += (*c == sep);
      m.cols = n;
      m.index_name.assign(b, static_cast<const char*>(std::memchr(b, sep, static_cast<size_t>(e - b))) ? static_cast<const char*>(std::memchr(b, sep, static_cast<size_t>(e - b))) : e);
      header_done = true;
      continue;
    }
    const char* c = b;
    size_t k = 0;
    while (c <= e) {
      const char* q = static_cast<const char*>(std::memchr(c, sep, static_cast<size_t>(e - c)));
      if (!q) q = e;
      // integers, possibly written as floats ("3.0"); an empty cell is not a number here
      bool neg = false;
      const char* d = c;
      if (d < q && (*d == '-' || *d == '+')) { neg = (*d == '-'); d++; }
      long long v = 0;
      const char* d0 = d;
      while (d < q && *d >= '0' && *d <= '9') { v = v * 10 + (*d - '0'); d++; }
      if (d == d0 || (d < q && *d != '.')) {
        v = static_cast<long long>(parse_cell(c, q, path, line));
        neg = false;
      }
      if (neg) v = -v;
      if (k == 0) m.index.push_back(v);
      else {
        if (v < -32768 || v > 32767) throw Error(path + ": line " + std::to_string(line) + ": value out of the int16 range");
        m.v.push_back(static_cast<int16_t>(v));
      }
      k++;
      c = q + 1;
    }
    if (k != m.cols + 1) throw Error(path + ": line " + std::to_string(line) + " has " + std::to_string(k) + " fields, expected " + std::to_string(m.cols + 1));
    m.rows++;
  }
  return m;
}

}  // namespace hygio
