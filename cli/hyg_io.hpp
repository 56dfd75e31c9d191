// cli/hyg_io.hpp -- file formats of the reference's command-line contract (host only, no CUDA).
//
// What is reproduced (paths relative to /root/reference):
//   input tables      readr::read_csv as used by src/single_group/src/r/input_output_functions.R:10-20,80-103 (first line is ALWAYS a
//                     header, gz by content) and pandas.read_table(sep=',', header=None) of src/two_group/run_inference_two_groups.py:177-191
//   regimes CSV       format(., scientific = FALSE) + readr::write_csv, src/single_group/bin/estimate_parameters_and_regimes:326-338
//   theta / p / omega / kappa CSVs   readr::write_csv of doubles, bin/estimate_parameters_and_regimes:343-379, input_output_functions.R:4-7,54-73
//   np.savetxt default format, np.savez_compressed                   run_inference_two_groups.py:246-255,304-322
#ifndef HYG_IO_HPP
#define HYG_IO_HPP

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace hygio {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

bool ends_with(const std::string& s, const std::string& suffix);
// dir.create(dirname(path), recursive = TRUE) (input_output_functions.R:23-29)
void mkdirs_for_file(const std::string& path);
void mkdirs(const std::string& dir);

// whole file as text; gzip is detected from the content (zlib reads plain files transparently)
std::string read_text(const std::string& path);

struct Table {
  std::vector<std::string> header;  // empty when the file was read without a header
  size_t rows = 0, cols = 0;
  std::vector<double> v;            // row-major rows x cols
  double at(size_t r, size_t c) const { return v[r * cols + c]; }
};
// Comma-separated numeric table.  first_line_is_header = true is readr::read_csv's behaviour (the first line is consumed as
// column names whatever it holds: SURVEY C-1, the first CpG site of a header-less file is dropped); false is header=None.
Table read_csv_numeric(const std::string& path, bool first_line_is_header);
Table read_delimited_numeric(const std::string& path, bool first_line_is_header, char sep);

// pandas.read_csv(sep='\t').set_index(first column) of an integer matrix (aggregate_results.py:165-206 writes them,
// get_dmps.py:58-61 reads them): header line, then index value + `cols` small integers per row
struct IndexedIntMatrix {
  std::string index_name;
  std::vector<long long> index;
  size_t rows = 0, cols = 0;
  std::vector<int16_t> v;   // row-major rows x cols
};
IndexedIntMatrix read_indexed_int_matrix(const std::string& path, char sep = '\t');

// np.load(path)['arr_0'] of an .npz with one member (np.savez_compressed / np.savez; zip64 aware): raw little-endian C-order data
struct NpyArray {
  std::string descr;           // e.g. "<i2"
  std::vector<size_t> shape;
  std::vector<unsigned char> data;
};
NpyArray load_npz(const std::string& path);

// Text sink; gzip-compressed when the path ends in ".gz" (what readr::write_csv and np.savetxt do)
class Writer {
 public:
  explicit Writer(const std::string& path);
  ~Writer();
  void write(const std::string& s);
  void close();

 private:
  void* gz_ = nullptr;
  void* fp_ = nullptr;
  std::string path_;
};

// R: format(x, scientific = FALSE) with getOption("digits") = 7 -- common number of decimals, common width, right-justified
// (src/main/format.c formatReal/scientific, restated).
std::vector<std::string> r_format_fixed(const std::vector<double>& x, int digits = 7);
// shortest digits that round-trip, laid out the way readr/vroom's grisu3 writer does it
std::string readr_double(double x);
// shortest digits that round-trip in plain decimal notation (data.table::fwrite(scipen = 999))
std::string fixed_double(double x);
// Python repr(float)
std::string py_repr_double(double x);

// np.savetxt(path, a, delimiter=',') with the default fmt '%.18e'
void savetxt_e18(const std::string& path, const double* a, size_t rows, size_t cols);

// np.savez_compressed(path, arr): one deflated member "arr_0.npy".  descr e.g. "<i2", "<f4"; data is little-endian C order.
void save_npz(const std::string& path, const std::string& descr, const std::vector<size_t>& shape, const void* data, size_t nbytes);

}  // namespace hygio
#endif
