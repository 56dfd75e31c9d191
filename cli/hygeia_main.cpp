// cli/hygeia_main.cpp -- `hygeia`: the reference's command-line contract on top of libhygeia_b200.so (C ABI).
//
// Sub-commands and the reference entry points they stand in for (paths relative to /root/reference):
//   hygeia estimate_parameters_and_regimes ...   src/single_group/bin/estimate_parameters_and_regimes (flags :12-204, flow :206-379)
//   hygeia infer ...                             src/two_group/run_inference_two_groups.py (flags :19-73, flow :92-322)
//   hygeia preprocess ...                        src/two_group/preprocess_bed.py (host-side ETL; parity unpinned, see there)
//   hygeia get_chrom_segments ...                src/two_group/get_chrom_segments.py
//   hygeia aggregate ...                         src/two_group/aggregate_results.py
//   hygeia get_dmps ...                          src/two_group/get_dmps.py (+ multiple_testing.py)
//   hygeia make_bed_file ...                     src/single_group/bin/make_bed_file
//   hygeia --version | -v | version              src/single_group/hygeia.docker:44-46, src/two_group/hygeia.docker:54-56
// The Nextflow modules (modules/single_group/2_estimate_parameters.nf:39-51, 3_estimate_regimes.nf:35-46,
// modules/two_group/2_estimate_parameters_and_regimes.nf:39-52, 4_infer.nf:44-49) call these unchanged.
// All computation goes through the C ABI (include/hygeia_b200.h); there is no CPU fallback.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <set>
#include <string>
#include <vector>

#include <unistd.h>

#include "../include/hygeia_b200.h"
#include "hyg_io.hpp"

using hygio::Error;

namespace {

#ifndef HYGEIA_CLI_VERSION
#define HYGEIA_CLI_VERSION "0.1.0+b200"
#endif

// ---- argument parsing: `--name value`, `--name=value`, boolean flags with or without a value ----------------------------
struct Args {
  std::map<std::string, std::vector<std::string>> kv;
  bool has(const std::string& k) const { return kv.count(k) != 0; }
  std::string str(const std::string& k, const std::string& def) const {
    auto it = kv.find(k);
    return (it == kv.end() || it->second.empty()) ? def : it->second.back();
  }
};

bool parse_bool(const std::string& s, bool& out) {
  std::string t;
  for (char c : s) t.push_back(static_cast<char>(std::tolower(static_cast<unsigned char>(c))));
  if (t == "true" || t == "t" || t == "1" || t == "yes") { out = true; return true; }
  if (t == "false" || t == "f" || t == "0" || t == "no") { out = false; return true; }
  return false;
}

// `flags`: names that may appear without a value (argparser flag = TRUE; absl booleans, which also accept --noNAME)
Args parse_args(int argc, char** argv, int first, const std::set<std::string>& known, const std::set<std::string>& flags) {
  Args a;
  for (int i = first; i < argc; i++) {
    std::string tok = argv[i];
    if (tok.rfind("--", 0) != 0) throw Error("unexpected argument '" + tok + "'");
    tok = tok.substr(2);
    std::string name = tok, value;
    bool has_value = false;
    const size_t eq = tok.find('=');
    if (eq != std::string::npos) { name = tok.substr(0, eq); value = tok.substr(eq + 1); has_value = true; }
    if (!known.count(name)) {
      if (name.rfind("no", 0) == 0 && flags.count(name.substr(2)) && !has_value) { a.kv[name.substr(2)].push_back("FALSE"); continue; }
      throw Error("unknown flag --" + name);
    }
    if (!has_value) {
      if (flags.count(name)) {
        bool b;
        if (i + 1 < argc && std::strncmp(argv[i + 1], "--", 2) != 0 && parse_bool(argv[i + 1], b)) { value = argv[++i]; }
        else value = "TRUE";
      } else {
        if (i + 1 >= argc) throw Error("flag --" + name + " needs a value");
        value = argv[++i];
      }
    }
    a.kv[name].push_back(value);
  }
  return a;
}

std::vector<double> parse_list(const std::string& s, const std::string& what) {
  std::vector<double> v;
  size_t pos = 0;
  while (pos <= s.size()) {
    size_t c = s.find(',', pos);
    if (c == std::string::npos) c = s.size();
    const std::string t = s.substr(pos, c - pos);
    char* end = nullptr;
    const double x = std::strtod(t.c_str(), &end);
    if (t.empty() || end != t.c_str() + t.size()) throw Error("--" + what + ": not a number: '" + t + "'");
    v.push_back(x);
    pos = c + 1;
  }
  return v;
}

bool get_bool(const Args& a, const std::string& k, bool def) {
  if (!a.has(k)) return def;
  bool b;
  if (!parse_bool(a.str(k, ""), b)) throw Error("--" + k + ": expected TRUE or FALSE");
  return b;
}
long get_int(const Args& a, const std::string& k, long def) {
  if (!a.has(k)) return def;
  const std::string s = a.str(k, "");
  char* end = nullptr;
  const long v = std::strtol(s.c_str(), &end, 10);
  if (s.empty() || *end) throw Error("--" + k + ": expected an integer, got '" + s + "'");
  return v;
}
double get_double(const Args& a, const std::string& k, double def) {
  if (!a.has(k)) return def;
  const std::string s = a.str(k, "");
  char* end = nullptr;
  const double v = std::strtod(s.c_str(), &end);
  if (s.empty() || *end) throw Error("--" + k + ": expected a number, got '" + s + "'");
  return v;
}

struct Ctx {
  hyg_ctx* c = nullptr;
  Ctx() {
    c = hyg_create(0);
    if (!c) throw Error(std::string("cannot initialise the GPU: ") + hyg_create_error());
  }
  ~Ctx() { hyg_destroy(c); }
  void check(int rc, const char* what) {
    if (rc < 0) throw Error(std::string(what) + ": " + hyg_last_error(c));
  }
};

// counts: table rows = sites, cols = samples -> uint16 [S][T] (site fastest), the device layout
std::vector<uint16_t> to_counts(const hygio::Table& t, const std::string& what, size_t row0 = 0, size_t row1 = SIZE_MAX) {
  row1 = std::min(row1, t.rows);
  const size_t T = row1 - row0, S = t.cols;
  std::vector<uint16_t> out(S * T);
  for (size_t r = 0; r < T; r++)
    for (size_t s = 0; s < S; s++) {
      const double v = t.at(row0 + r, s);
      if (!(v >= 0.0 && v <= 65535.0) || v != std::floor(v)) throw Error(what + ": read counts must be integers in 0..65535 (row " + std::to_string(row0 + r + 1) + ")");
      out[s * T + r] = static_cast<uint16_t>(v);
    }
  return out;
}

// ======================================================================================================================
// hygeia estimate_parameters_and_regimes
// ======================================================================================================================
int cmd_single_group(int argc, char** argv) {
  const std::set<std::string> known = {"mu", "sigma", "u", "kappa", "omega", "p_input_csv_file", "kappa_input_csv_file", "omega_input_csv_file",
                                       "n_methylated_reads_csv_file", "genomic_positions_csv_file", "n_total_reads_csv_file",
                                       "regime_probabilities_csv_file", "theta_trace_csv_file", "omega_csv_file", "kappa_csv_file", "p_csv_file",
                                       "theta_file", "is_kappa_fixed", "n_particles", "estimate_regime_probabilities", "estimate_parameters",
                                       "epsilon", "normalise_gradients", "use_adam", "n_steps_without_parameter_update", "learning_rate_exponent",
                                       "learning_rate_factor", "root_dir", "randomise_rng_seed", "rng_seed",
                                       // extensions (absent from the reference CLI; defaults reproduce it)
                                       "segment_sites", "segment_halo"};
  const std::set<std::string> flags = {"estimate_regime_probabilities", "estimate_parameters"};
  const Args a = parse_args(argc, argv, 2, known, flags);

  // defaults of bin/estimate_parameters_and_regimes:12-204
  const std::vector<double> mu = parse_list(a.str("mu", "0.99,0.01,0.80,0.20,0.50,0.50"), "mu");
  const std::vector<double> sigma = parse_list(a.str("sigma", "0.05,0.05,0.20,0.20,0.20,0.2886751"), "sigma");
  const long u = get_int(a, "u", 2);
  const size_t R = mu.size();
  if (sigma.size() != R) throw Error("--mu and --sigma differ in length");
  std::vector<double> kappa, omega;
  if (!a.str("kappa_input_csv_file", "").empty()) {
    const hygio::Table t = hygio::read_csv_numeric(a.str("kappa_input_csv_file", ""), true);
    for (size_t r = 0; r < t.rows; r++) kappa.push_back(t.at(r, 0));
  } else kappa = parse_list(a.str("kappa", "2,2,2,2,2,2"), "kappa");
  if (!a.str("omega_input_csv_file", "").empty()) {
    const hygio::Table t = hygio::read_csv_numeric(a.str("omega_input_csv_file", ""), true);
    for (size_t r = 0; r < t.rows; r++) omega.push_back(t.at(r, 0));
  } else omega = parse_list(a.str("omega", "0.995,0.975,0.950,0.925,0.900,0.900"), "omega");
  if (kappa.size() != R || omega.size() != R) throw Error("kappa / omega must have one entry per regime");
  // p: R x R, row-major here
  std::vector<double> p(R * R, 0.0);
  if (!a.str("p_input_csv_file", "").empty()) {
    const hygio::Table t = hygio::read_csv_numeric(a.str("p_input_csv_file", ""), true);
    if (t.rows != R || t.cols != R) throw Error("--p_input_csv_file must hold an R x R matrix");
    p = t.v;
  } else {
    // hard-wired 1/5 whatever R is (bin/estimate_parameters_and_regimes:243-247)
    for (size_t i = 0; i < R; i++)
      for (size_t j = 0; j < R; j++) p[i * R + j] = (i == j) ? 0.0 : 1.0 / 5.0;
  }
  const bool is_kappa_fixed = get_bool(a, "is_kappa_fixed", true);
  const long n_particles = get_int(a, "n_particles", 250);
  const bool est_regimes = get_bool(a, "estimate_regime_probabilities", false);
  const bool est_params = get_bool(a, "estimate_parameters", false);
  const double epsilon = get_double(a, "epsilon", 0.01);
  const bool normalise = get_bool(a, "normalise_gradients", false);
  const bool use_adam = get_bool(a, "use_adam", true);
  const long n_steps = get_int(a, "n_steps_without_parameter_update", 200);
  const double lr_exp = get_double(a, "learning_rate_exponent", 0.1);
  const double lr_fac = get_double(a, "learning_rate_factor", 0.01);
  const bool randomise = get_bool(a, "randomise_rng_seed", true);
  const long rng_seed = get_int(a, "rng_seed", -73);
  // throughput mode of the regimes pass: cut the chromosome into concurrent segments (0 = one sequential run, the reference's
  // behaviour; "auto" = size chosen for the device).  Ignored with --estimate_parameters (theta evolves along the chain).
  const std::string seg_s = a.str("segment_sites", "0");
  const uint64_t segment_sites = (seg_s == "auto") ? HYG_SEGMENT_AUTO : static_cast<uint64_t>(std::stoull(seg_s));
  const long segment_halo = get_int(a, "segment_halo", 5000);
  if (segment_halo < 0) throw Error("--segment_halo must be >= 0");
  if (!is_kappa_fixed) throw Error("--is_kappa_fixed FALSE is not supported (the reference's kappa-gradient path is broken: SURVEY.md C-7)");

  const std::string f_meth = a.str("n_methylated_reads_csv_file", ""), f_pos = a.str("genomic_positions_csv_file", ""),
                    f_tot = a.str("n_total_reads_csv_file", ""), f_reg = a.str("regime_probabilities_csv_file", ""),
                    f_trace = a.str("theta_trace_csv_file", ""), f_p = a.str("p_csv_file", "p.csv"), f_omega = a.str("omega_csv_file", "omega.csv"),
                    f_kappa = a.str("kappa_csv_file", "kappa.csv"), f_theta = a.str("theta_file", "p.csv");
  if (f_meth.empty() || f_pos.empty() || f_tot.empty()) throw Error("--n_methylated_reads_csv_file, --n_total_reads_csv_file and --genomic_positions_csv_file are required");
  if (est_regimes && f_reg.empty()) throw Error("--regime_probabilities_csv_file is required with --estimate_regime_probabilities");
  if (est_params && f_trace.empty()) throw Error("--theta_trace_csv_file is required with --estimate_parameters");
  // create_dirs_for_file on every file argument (:250-262)
  for (const std::string& f : {f_meth, f_pos, f_tot, f_reg, f_trace, f_p, f_omega, f_kappa, f_theta}) hygio::mkdirs_for_file(f);

  // get_known_parameters (model_functions.R:36-63)
  std::vector<double> vartheta = {static_cast<double>(u), static_cast<double>(R)};
  std::vector<double> alpha(R), beta(R);
  for (size_t r = 0; r < R; r++) {
    const double nu = mu[r] * (1.0 - mu[r]) / (sigma[r] * sigma[r]) - 1.0;
    alpha[r] = mu[r] * nu;
    beta[r] = (1.0 - mu[r]) * nu;
  }
  vartheta.insert(vartheta.end(), alpha.begin(), alpha.end());
  vartheta.insert(vartheta.end(), beta.begin(), beta.end());
  vartheta.push_back(1.0);
  vartheta.insert(vartheta.end(), kappa.begin(), kappa.end());
  const size_t D = R * R;

  uint64_t seed = static_cast<uint64_t>(static_cast<int64_t>(rng_seed));
  if (randomise) {
    std::random_device rd;
    seed = (static_cast<uint64_t>(rd()) << 32) ^ rd();
  }
  std::vector<double> theta(D);
  if (est_params) {
    // sampleFromParameterPriorCpp (singleGroup.cpp:18-35)
    if (hyg_sg_sample_theta_prior(static_cast<uint32_t>(D), seed, theta.data()) < 0) throw Error("hyg_sg_sample_theta_prior failed");
  } else {
    // convert_model_parameters_to_theta (model_functions.R:65-78): diag(p) <- -1; c(log(p[p != -1]), logit(omega)).
    // `p[...]` walks the matrix in COLUMN-major order: the transition matrix goes in transposed.
    size_t k = 0;
    for (size_t col = 0; col < R; col++)
      for (size_t row = 0; row < R; row++) {
        if (row == col) continue;
        if (p[row * R + col] == -1.0) continue;
        if (k >= R * (R - 1)) throw Error("p has more than R(R-1) off-diagonal entries different from -1");
        theta[k++] = std::log(p[row * R + col]);
      }
    if (k != R * (R - 1)) throw Error("p must not contain -1 off the diagonal");
    for (size_t r = 0; r < R; r++) theta[k++] = std::log(omega[r] / (1.0 - omega[r]));
  }

  // data: site per row, first line consumed as header (readr::read_csv)
  std::fprintf(stderr, "Reading data from file: %s\n", f_pos.c_str());
  const hygio::Table tpos = hygio::read_csv_numeric(f_pos, true);
  std::fprintf(stderr, "Reading data from file: %s\n", f_tot.c_str());
  const hygio::Table ttot = hygio::read_csv_numeric(f_tot, true);
  std::fprintf(stderr, "Reading data from file: %s\n", f_meth.c_str());
  const hygio::Table tmeth = hygio::read_csv_numeric(f_meth, true);
  const size_t T = tpos.rows, S = ttot.cols;
  if (T == 0) throw Error("no CpG sites in " + f_pos);
  if (ttot.rows != T || tmeth.rows != T || tmeth.cols != S) throw Error("positions, n_total_reads and n_methylated_reads differ in shape");
  std::vector<uint32_t> pos(T);
  for (size_t t = 0; t < T; t++) {
    const double v = tpos.at(t, 0);
    if (!(v >= 0.0 && v < 4294967296.0)) throw Error("genomic position out of range at row " + std::to_string(t + 1));
    pos[t] = static_cast<uint32_t>(v);
  }
  const std::vector<uint16_t> nt = to_counts(ttot, f_tot), nm = to_counts(tmeth, f_meth);

  Ctx ctx;
  hyg_sg_run_args ra;
  hyg_sg_default_run_args(&ra);
  ra.n_particles_max = static_cast<uint32_t>(n_particles);
  ra.smc_proposal_type = 1;
  ra.smc_resample_type = 2;
  ra.use_online_marginal_smoothing = est_regimes;
  ra.epsilon = epsilon;
  ra.use_online_parameter_estimation = est_params;
  ra.normalise_gradients = normalise;
  ra.use_adam = use_adam;
  ra.n_steps_without_parameter_update = static_cast<uint32_t>(n_steps);
  ra.learning_rate_exponent = lr_exp;
  ra.learning_rate_factor = lr_fac;
  ctx.check(hyg_sg_set_segmentation(ctx.c, segment_sites, static_cast<uint64_t>(segment_halo), static_cast<uint64_t>(segment_halo)),
            "hyg_sg_set_segmentation");
  std::vector<double> probs(est_regimes ? T * (1 + R) : 0), trace(est_params ? T * D : 0);
  double seconds = 0.0;
  int32_t status[HYG_SG_STATUS_WORDS] = {0};
  ctx.check(hyg_sg_run_online_combined_inference(ctx.c, vartheta.data(), static_cast<uint32_t>(vartheta.size()), theta.data(), static_cast<uint32_t>(D), T,
                                                 static_cast<uint32_t>(S), pos.data(), nt.data(), nm.data(), &ra, seed, nullptr,
                                                 est_regimes ? probs.data() : nullptr, est_params ? trace.data() : nullptr, nullptr, &seconds, status),
            "runOnlineCombinedInference");
  std::fprintf(stderr, "inference: %zu sites x %zu samples in %.3f s\n", T, S, seconds);
  // The kernel's status words (include/hygeia_b200.h): anything that makes the CSV differ from the reference's estimator is
  // either fatal or said out loud -- a lag-set overflow already failed the call above (HYG_ERR_CAPACITY).
  if (status[5] > 0)
    std::fprintf(stderr, "note: at %d site(s) two particles had exactly equal weights and only one survived the resampling; the reference "
                         "breaks such ties by the (unspecified) order std::sort leaves, this program by (regime, sojourn).\n", status[5]);
  if (status[2] > 0 || status[7] > 0) {
    std::fprintf(stderr, "error: --segment_sites %llu with halos of %lld sites is not safe for this data: %d owned site(s) were still pending at the "
                         "end of a right halo, %d overlap row(s) differ by more than 1e-6 from the next segment's (left halo too short). "
                         "Use larger halos or whole chains (--segment_sites 0).\n",
                 static_cast<unsigned long long>(segment_sites), static_cast<long long>(segment_halo), status[2], status[7]);
    return 3;
  }

  if (est_regimes) {
    // columns genomic_position, regime_1..R; every column through format(., scientific = FALSE) (:326-338)
    std::vector<std::vector<std::string>> cols(1 + R);
    std::vector<double> col(T);
    for (size_t c = 0; c <= R; c++) {
      for (size_t t = 0; t < T; t++) col[t] = probs[t * (1 + R) + c];
      cols[c] = hygio::r_format_fixed(col);
    }
    hygio::Writer w(f_reg);
    std::string line = "genomic_position";
    for (size_t r = 1; r <= R; r++) line += ",regime_" + std::to_string(r);
    line += "\n";
    w.write(line);
    std::string chunk;
    for (size_t t = 0; t < T; t++) {
      for (size_t c = 0; c <= R; c++) {
        if (c) chunk.push_back(',');
        chunk += cols[c][t];
      }
      chunk.push_back('\n');
      if (chunk.size() > (1u << 20)) { w.write(chunk); chunk.clear(); }
    }
    w.write(chunk);
    w.close();
  }
  if (est_params) {
    {  // theta trace: theta_1..theta_D, one row per site (:343-348)
      hygio::Writer w(f_trace);
      std::string line;
      for (size_t k = 1; k <= D; k++) line += (k > 1 ? ",theta_" : "theta_") + std::to_string(k);
      line += "\n";
      w.write(line);
      std::string chunk;
      for (size_t t = 0; t < T; t++) {
        for (size_t k = 0; k < D; k++) {
          if (k) chunk.push_back(',');
          chunk += hygio::readr_double(trace[t * D + k]);
        }
        chunk.push_back('\n');
        if (chunk.size() > (1u << 20)) { w.write(chunk); chunk.clear(); }
      }
      w.write(chunk);
      w.close();
    }
    // convert_theta_to_model_parameters on the last row (model_functions.R:81-111)
    const double* th = trace.data() + (T - 1) * D;
    std::vector<double> pf(R * R, 0.0), om(R);
    for (size_t rr = 0; rr < R; rr++) {
      const double* blk = th + rr * (R - 1);
      double mx = blk[0];
      for (size_t k = 1; k + 1 < R; k++) mx = std::max(mx, blk[k]);
      double sm = 0.0;
      for (size_t k = 0; k + 1 < R; k++) sm += std::exp(blk[k] - mx);
      const double lz = mx + std::log(sm);
      size_t k = 0;
      for (size_t c = 0; c < R; c++)
        if (c != rr) pf[rr * R + c] = std::exp(blk[k++] - lz);
    }
    for (size_t r = 0; r < R; r++) om[r] = 1.0 / (1.0 + std::exp(-th[R * (R - 1) + r]));
    {
      hygio::Writer w(f_p);
      std::string s;
      for (size_t r = 1; r <= R; r++) s += (r > 1 ? ",regime_" : "regime_") + std::to_string(r);
      s += "\n";
      for (size_t i = 0; i < R; i++) {
        for (size_t j = 0; j < R; j++) s += (j ? "," : "") + hygio::readr_double(pf[i * R + j]);
        s += "\n";
      }
      w.write(s);
      w.close();
    }
    auto write_column = [](const std::string& file, const std::string& name, const double* v, size_t n) {
      hygio::Writer w(file);
      std::string s = name + "\n";
      for (size_t i = 0; i < n; i++) s += hygio::readr_double(v[i]) + "\n";
      w.write(s);
      w.close();
    };
    write_column(f_omega, "omega", om.data(), R);
    write_column(f_kappa, "kappa", kappa.data(), R);
    write_column(f_theta, "data", th, D);
  }
  return 0;
}

// ======================================================================================================================
// hygeia infer
// ======================================================================================================================
int cmd_infer(int argc, char** argv) {
  const std::set<std::string> known = {"mu", "sigma", "minimum_duration", "omega_case", "merge_log_prob", "split_prob", "num_resampled_particles",
                                       "num_samples_backward", "multinomial", "chrom", "results_dir", "data_dir", "single_group_dir", "seed", "batch",
                                       "segment_size", "buffer_size", "hazard"};
  const std::set<std::string> flags = {"multinomial"};
  const Args a = parse_args(argc, argv, 2, known, flags);
  // defaults of run_inference_two_groups.py:19-73
  const std::string s_mu = a.str("mu", "0.95,0.05,0.80,0.20,0.50,0.50"), s_sigma = a.str("sigma", "0.05,0.05,0.1,0.1,0.1,0.2886751");
  const std::vector<double> mu = parse_list(s_mu, "mu"), sigma = parse_list(s_sigma, "sigma");
  const size_t R = mu.size();
  if (sigma.size() != R) throw Error("--mu and --sigma differ in length");
  const long u = get_int(a, "minimum_duration", 3);
  const double omega_case = get_double(a, "omega_case", 0.8);
  const double merge_log_prob = get_double(a, "merge_log_prob", std::log(0.1));
  const double split_prob = get_double(a, "split_prob", 0.01);
  std::vector<long> Ms;
  if (a.has("num_resampled_particles"))
    for (const std::string& s : a.kv.at("num_resampled_particles")) Ms.push_back(std::strtol(s.c_str(), nullptr, 10));
  else Ms.push_back(50);
  const long B = get_int(a, "num_samples_backward", 25);
  if (get_bool(a, "multinomial", false)) throw Error("--multinomial is not supported (the default, optimal finite-state + systematic resampling, is what is built)");
  const std::string chrom = a.str("chrom", "22"), results_dir = a.str("results_dir", "../test"), data_dir = a.str("data_dir", "data"),
                    sg_dir = a.str("single_group_dir", "test_data/single_group_results");
  const long seed = get_int(a, "seed", 0), batch = get_int(a, "batch", 0), segment = get_int(a, "segment_size", 100000), buffer = get_int(a, "buffer_size", 5000);
  if (batch < 0 || segment <= 0 || buffer < 0) throw Error("--batch, --segment_size and --buffer_size must be non-negative");
  // not a flag of the reference: "reference" (default) evaluates the sojourn hazard as its fp32 TensorFlow code does, fixed
  // value 0.1 included (case_control_regime_model.py:111-168); "exact" is the negative-binomial hazard in fp64
  const std::string hazard = a.str("hazard", "reference");
  if (hazard != "reference" && hazard != "exact") throw Error("--hazard must be reference or exact");

  // results_dir/chrom_{chrom}_{batch}/ and the flag dump (:101-108)
  const std::string path = results_dir + "/chrom_" + chrom + "_" + std::to_string(batch);
  hygio::mkdirs(path);
  {
    std::string s;
    s += "--mu=" + s_mu + "\n--sigma=" + s_sigma + "\n--minimum_duration=" + std::to_string(u) + "\n--omega_case=" + hygio::py_repr_double(omega_case);
    s += "\n--merge_log_prob=" + hygio::py_repr_double(merge_log_prob) + "\n--split_prob=" + hygio::py_repr_double(split_prob);
    for (long M : Ms) s += "\n--num_resampled_particles=" + std::to_string(M);
    s += "\n--num_samples_backward=" + std::to_string(B) + "\n--nomultinomial\n--chrom=" + chrom + "\n--results_dir=" + results_dir;
    s += "\n--data_dir=" + data_dir + "\n--single_group_dir=" + sg_dir + "\n--seed=" + std::to_string(seed) + "\n--batch=" + std::to_string(batch);
    s += "\n--segment_size=" + std::to_string(segment) + "\n--buffer_size=" + std::to_string(buffer);
    std::printf("specified flags:\n%s\n", s.c_str());
    hygio::Writer w(path + "/flags" + std::to_string(seed) + ".txt");
    w.write(s);
    w.close();
  }

  // get_estimated_control_group_param (:76-89): column `data` of theta_<chrom>.csv.gz
  const hygio::Table th = hygio::read_csv_numeric(sg_dir + "/theta_" + chrom + ".csv.gz", true);
  if (th.rows < R * R) throw Error("theta file has fewer than R*R rows");
  size_t data_col = 0;
  for (size_t c = 0; c < th.header.size(); c++)
    if (th.header[c] == "data") data_col = c;
  std::vector<double> logp(R * R, -HUGE_VAL), omega_control(R), omega_k(R, omega_case);
  {
    size_t i = 0;
    for (size_t r = 0; r < R; r++) {
      double sum = 0.0;
      std::vector<double> row(R, 0.0);
      for (size_t r1 = 0; r1 < R; r1++)
        if (r != r1) { row[r1] = std::exp(th.at(i++, data_col)); sum += row[r1]; }
      for (size_t r1 = 0; r1 < R; r1++)
        if (r != r1) logp[r * R + r1] = std::log(row[r1] / sum);
    }
    // the reference applies sigmoid twice and the model takes logit once (:145-150, case_control_regime_model.py:117-119)
    for (size_t r = 0; r < R; r++) omega_control[r] = 1.0 / (1.0 + std::exp(-th.at(th.rows - R + r, data_col)));
  }

  // whole-chromosome inputs, no header (:176-191)
  const hygio::Table tpos = hygio::read_csv_numeric(data_dir + "/positions_" + chrom + ".txt.gz", false);
  const hygio::Table ttc = hygio::read_csv_numeric(data_dir + "/n_total_reads_control_" + chrom + ".txt.gz", false);
  const hygio::Table tmc = hygio::read_csv_numeric(data_dir + "/n_methylated_reads_control_" + chrom + ".txt.gz", false);
  const hygio::Table ttk = hygio::read_csv_numeric(data_dir + "/n_total_reads_case_" + chrom + ".txt.gz", false);
  const hygio::Table tmk = hygio::read_csv_numeric(data_dir + "/n_methylated_reads_case_" + chrom + ".txt.gz", false);
  const size_t n_sites = tpos.rows;
  if (ttc.rows != n_sites || tmc.rows != n_sites || ttk.rows != n_sites || tmk.rows != n_sites) throw Error("input matrices differ in the number of sites");
  if (tmc.cols != ttc.cols || tmk.cols != ttk.cols) throw Error("total and methylated read matrices differ in the number of samples");
  // segment selection (:194-219)
  if (static_cast<size_t>(batch) * static_cast<size_t>(segment) > n_sites) {
    std::printf("Batch index is too large for the chromosome\n");
    return 0;
  }
  const size_t lo = static_cast<size_t>(std::max<long>(0, batch * segment - buffer));
  const size_t hi = std::min<size_t>(static_cast<size_t>((batch + 1) * segment + buffer), n_sites);
  const size_t T = hi - lo;
  size_t ret0, ret1;
  if (batch == 0) { ret0 = 0; ret1 = std::min<size_t>(T, static_cast<size_t>(segment)); }
  else { ret0 = static_cast<size_t>(buffer); ret1 = std::min<size_t>(T, static_cast<size_t>(buffer + segment)); }
  if (ret0 > ret1) ret0 = ret1;
  if (T == 0) throw Error("empty segment");
  const std::vector<uint16_t> ntc = to_counts(ttc, "n_total_reads_control", lo, hi), nmc = to_counts(tmc, "n_methylated_reads_control", lo, hi);
  const std::vector<uint16_t> ntk = to_counts(ttk, "n_total_reads_case", lo, hi), nmk = to_counts(tmk, "n_methylated_reads_case", lo, hi);
  const size_t Sc = ttc.cols, Sk = ttk.cols;
  for (size_t i = 0; i < ntk.size(); i++)
    if (nmk[i] > ntk[i]) throw Error("assertion failed: case methylated reads exceed total reads");   // :210

  // echo the inputs of the kept window (:246-255): int16 values, np.savetxt default format
  auto echo = [&](const std::string& name, const hygio::Table& t) {
    std::vector<double> v((ret1 - ret0) * t.cols);
    for (size_t r = ret0; r < ret1; r++)
      for (size_t c = 0; c < t.cols; c++) v[(r - ret0) * t.cols + c] = static_cast<double>(static_cast<int16_t>(static_cast<int64_t>(t.at(lo + r, c))));
    hygio::savetxt_e18(path + "/" + name, v.data(), ret1 - ret0, t.cols);
  };
  echo("observations_control.csv.gz", tmc);
  echo("observations_case.csv.gz", tmk);
  echo("n_total_reads_control.csv.gz", ttc);
  echo("n_total_reads_case.csv.gz", ttk);
  {
    std::vector<double> v((ret1 - ret0) * tpos.cols);
    for (size_t r = ret0; r < ret1; r++)
      for (size_t c = 0; c < tpos.cols; c++) v[(r - ret0) * tpos.cols + c] = tpos.at(lo + r, c);
    hygio::savetxt_e18(path + "/positions.csv.gz", v.data(), ret1 - ret0, tpos.cols);
  }

  Ctx ctx;
  std::vector<double> alpha(R), beta(R), two(R, 2.0);
  for (size_t r = 0; r < R; r++) {
    const double nu = mu[r] * (1.0 - mu[r]) / (sigma[r] * sigma[r]) - 1.0;   // case_control_regime_model.py:19-23
    alpha[r] = mu[r] * nu;
    beta[r] = (1.0 - mu[r]) * nu;
  }
  ctx.check(hyg_sg_set_model(ctx.c, static_cast<uint32_t>(R), static_cast<uint32_t>(u), alpha.data(), beta.data(), 1, two.data()), "hyg_sg_set_model");
  ctx.check(hyg_sg_add_dataset(ctx.c, T, static_cast<uint32_t>(Sc), ntc.data(), nmc.data(), 0, T), "hyg_sg_add_dataset");
  ctx.check(hyg_sg_add_dataset(ctx.c, T, static_cast<uint32_t>(Sk), ntk.data(), nmk.data(), 0, T), "hyg_sg_add_dataset");
  ctx.check(hyg_sg_emission(ctx.c), "hyg_sg_emission");

  std::string s_logz = "{", s_time = "{";
  for (size_t mi = 0; mi < Ms.size(); mi++) {
    const long M = Ms[mi];
    std::printf("%ld\n", M);
    const long N = M * static_cast<long>(2 * R + R * R);
    const auto t0 = std::chrono::steady_clock::now();
    hyg_tg_model m;
    std::memset(&m, 0, sizeof(m));
    m.R = static_cast<uint32_t>(R); m.minimum_duration = static_cast<uint32_t>(u); m.num_resampled = static_cast<uint32_t>(M); m.num_backward = static_cast<uint32_t>(B);
    m.log_p_control = logp.data(); m.omega_control = omega_control.data(); m.omega_case = omega_k.data();
    m.kappa_control = two.data(); m.kappa_case = two.data();
    m.merge_prob = std::exp(merge_log_prob); m.split_prob = split_prob;
    m.hazard_mode = (hazard == "exact") ? HYG_TG_HAZARD_EXACT : HYG_TG_HAZARD_REFERENCE;
    ctx.check(hyg_tg_set_model(ctx.c, &m, T), "hyg_tg_set_model");
    std::vector<int32_t> traj(T * static_cast<size_t>(B) * 5);
    double log_norm = 0.0;
    hyg_tg_chain ch;
    std::memset(&ch, 0, sizeof(ch));
    ch.control_dataset = 0; ch.case_dataset = 1; ch.seed = static_cast<uint64_t>(seed); ch.chain_id = static_cast<uint32_t>(batch);
    ch.trajectories = traj.data(); ch.log_normalizing_constant = &log_norm;
    float ms = 0.0f;
    ctx.check(hyg_tg_run(ctx.c, &ch, 1, &ms), "hyg_tg_run");
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    // test functions averaged over the trajectories (:233-240,294-297) -- over the whole window, halo included, as the reference does
    std::vector<float> split(T), reg(T * 2 * R, 0.0f);
    std::vector<int16_t> merged((ret1 - ret0) * B), control((ret1 - ret0) * B * 2), cse((ret1 - ret0) * B * 2);
    for (size_t t = 0; t < T; t++) {
      int n_split = 0;
      std::vector<int> cc(R, 0), ck(R, 0);
      for (long b = 0; b < B; b++) {
        const int32_t* x = traj.data() + (t * B + b) * 5;
        n_split += (x[0] == 0);
        cc[x[2]]++; ck[x[4]]++;
        if (t >= ret0 && t < ret1) {
          const size_t o = (t - ret0) * B + b;
          merged[o] = static_cast<int16_t>(x[0]);
          control[2 * o] = static_cast<int16_t>(x[1]); control[2 * o + 1] = static_cast<int16_t>(x[2]);
          cse[2 * o] = static_cast<int16_t>(x[3]); cse[2 * o + 1] = static_cast<int16_t>(x[4]);
        }
      }
      split[t] = static_cast<float>(n_split) / static_cast<float>(B);
      for (size_t r = 0; r < R; r++) { reg[t * 2 * R + r] = static_cast<float>(cc[r]) / B; reg[t * 2 * R + R + r] = static_cast<float>(ck[r]) / B; }
    }
    const std::string tag = "_" + std::to_string(N) + "_" + std::to_string(seed) + ".npz";
    const size_t Tr = ret1 - ret0;
    hygio::save_npz(path + "/optimal_backward_particles_merged_state" + tag, "<i2", {Tr, static_cast<size_t>(B)}, merged.data(), merged.size() * 2);
    hygio::save_npz(path + "/optimal_backward_particles_control_state" + tag, "<i2", {Tr, static_cast<size_t>(B), 2}, control.data(), control.size() * 2);
    hygio::save_npz(path + "/optimal_backward_particles_case_state" + tag, "<i2", {Tr, static_cast<size_t>(B), 2}, cse.data(), cse.size() * 2);
    hygio::save_npz(path + "/optimal_split_probs" + tag, "<f4", {T}, split.data(), split.size() * 4);
    hygio::save_npz(path + "/optimal_regime_probs" + tag, "<f4", {T, 2 * R}, reg.data(), reg.size() * 4);
    s_logz += (mi ? ", " : "") + std::to_string(N) + ": " + hygio::py_repr_double(log_norm);
    s_time += (mi ? ", " : "") + std::to_string(N) + ": " + hygio::py_repr_double(wall);
    std::fprintf(stderr, "M=%ld: %zu sites, device %.1f ms, wall %.3f s, log normalising constant %.6f\n", M, T, ms, wall, log_norm);
  }
  s_logz += "}\n";
  s_time += "}\n";
  { hygio::Writer w(path + "/log_normalizing_constants_optimal_" + std::to_string(seed) + ".txt"); w.write(s_logz); w.close(); }
  { hygio::Writer w(path + "/optimal_time_" + std::to_string(seed) + ".txt"); w.write(s_time); w.close(); }
  { hygio::Writer w(path + "/optimal_time_backward_" + std::to_string(seed) + ".txt"); w.write("{}\n"); w.close(); }
  return 0;
}

// ======================================================================================================================
// hygeia aggregate (src/two_group/aggregate_results.py) and hygeia get_dmps (src/two_group/get_dmps.py)
// ======================================================================================================================
bool path_exists(const std::string& p) { return ::access(p.c_str(), F_OK) == 0; }

void append_int(std::string& s, long long v) {
  char b[24];
  const int n = std::snprintf(b, sizeof(b), "%lld", v);
  s.append(b, static_cast<size_t>(n));
}

// DataFrame.set_index(pos).to_csv(sep = '\t', compression = 'gzip') of an integer matrix: header "pos\t0\t1...", then rows
template <class Tv>
void write_indexed_matrix(const std::string& path, const std::vector<long long>& pos, const Tv* v, size_t rows, size_t cols) {
  hygio::Writer w(path);
  std::string buf = "pos";
  for (size_t c = 0; c < cols; c++) { buf += '\t'; append_int(buf, static_cast<long long>(c)); }
  buf += '\n';
  for (size_t r = 0; r < rows; r++) {
    append_int(buf, pos[r]);
    for (size_t c = 0; c < cols; c++) { buf += '\t'; append_int(buf, static_cast<long long>(v[r * cols + c])); }
    buf += '\n';
    if (buf.size() > (1u << 20)) { w.write(buf); buf.clear(); }
  }
  w.write(buf);
  w.close();
}

// one of the echoes `infer` writes with np.savetxt(delimiter = ','): the reference reads it with sep = ' ', which yields the one
// numeric column of a one-sample group (aggregate_results.py:100-108) and cannot be cast for more samples; here every sample
// column is kept (a superset of what the reference can do)
hygio::Table read_echo(const std::string& path) { return hygio::read_delimited_numeric(path, false, ','); }

int cmd_aggregate(int argc, char** argv) {
  const std::set<std::string> known = {"results_dir", "output_dir", "seeds", "chrom", "num_batches", "num_particles", "compute_freqs"};
  const Args a = parse_args(argc, argv, 2, known, {"compute_freqs"});
  const std::string results_dir = a.str("results_dir", "../test"), output_dir = a.str("output_dir", "../test/results"), chrom = a.str("chrom", "22");
  const long seeds = get_int(a, "seeds", 10), num_batches = get_int(a, "num_batches", 30), N = get_int(a, "num_particles", 2400);
  const bool compute_freqs = get_bool(a, "compute_freqs", false);
  if (seeds < 1) throw Error("--seeds must be positive");
  std::printf("Results directory: %s\nOutput directory: %s\nProcessing chromosome: %s\n", results_dir.c_str(), output_dir.c_str(), chrom.c_str());
  hygio::mkdirs(output_dir);

  std::vector<long long> pos;
  std::vector<int8_t> merged, creg, kreg;
  std::vector<int16_t> cdur, kdur, ntc, ntk, nmc, nmk;
  size_t P = 0, Sc = 0, Sk = 0;
  long processed = 0;
  for (long batch = 0; batch < num_batches; batch++) {
    const std::string dir = results_dir + "/chrom_" + chrom + "_" + std::to_string(batch);
    std::printf("\nProcessing batch %ld\nLooking for data in: %s\n", batch, dir.c_str());
    if (!path_exists(dir)) { std::printf("Directory does not exist: %s\n", dir.c_str()); break; }
    if (!path_exists(dir + "/positions.csv.gz")) { std::printf("positions.csv.gz not found in %s\n", dir.c_str()); break; }
    const hygio::Table tp = read_echo(dir + "/positions.csv.gz");
    const hygio::Table e_ntc = read_echo(dir + "/n_total_reads_control.csv.gz"), e_ntk = read_echo(dir + "/n_total_reads_case.csv.gz");
    const hygio::Table e_nmc = read_echo(dir + "/observations_control.csv.gz"), e_nmk = read_echo(dir + "/observations_case.csv.gz");
    const size_t Tb = tp.rows;
    if (e_ntc.rows != Tb || e_ntk.rows != Tb || e_nmc.rows != Tb || e_nmk.rows != Tb) throw Error(dir + ": the echoed count files differ in length");
    if (processed == 0) { Sc = e_ntc.cols; Sk = e_ntk.cols; }
    if (e_ntc.cols != Sc || e_nmc.cols != Sc || e_ntk.cols != Sk || e_nmk.cols != Sk) throw Error(dir + ": the number of samples changed between batches");
    // per seed: merged [Tb][B], control / case [Tb][B][2] = (duration, regime), int16 (aggregate_results.py:110-124)
    std::vector<hygio::NpyArray> am, ac, ak;
    size_t Pb = 0;
    for (long sd = 0; sd < seeds; sd++) {
      const std::string tag = "_" + std::to_string(N) + "_" + std::to_string(sd) + ".npz";
      am.push_back(hygio::load_npz(dir + "/optimal_backward_particles_merged_state" + tag));
      ac.push_back(hygio::load_npz(dir + "/optimal_backward_particles_control_state" + tag));
      ak.push_back(hygio::load_npz(dir + "/optimal_backward_particles_case_state" + tag));
      const hygio::NpyArray &m = am.back(), &c = ac.back(), &k = ak.back();
      if (m.descr != "<i2" || c.descr != "<i2" || k.descr != "<i2") throw Error(dir + ": trajectories are not int16");
      if (m.shape.size() != 2 || m.shape[0] != Tb || c.shape != std::vector<size_t>({Tb, m.shape[1], 2}) || k.shape != c.shape)
        throw Error(dir + ": trajectory shapes of seed " + std::to_string(sd) + " do not match the window");
      Pb += m.shape[1];
    }
    std::printf("Successfully processed %ld seeds out of %ld\n", seeds, seeds);
    if (processed == 0) P = Pb;
    if (Pb != P) throw Error(dir + ": the number of trajectories changed between batches");
    const size_t T0 = pos.size();
    pos.resize(T0 + Tb);
    merged.resize((T0 + Tb) * P); creg.resize((T0 + Tb) * P); kreg.resize((T0 + Tb) * P); cdur.resize((T0 + Tb) * P); kdur.resize((T0 + Tb) * P);
    for (size_t t = 0; t < Tb; t++) {
      pos[T0 + t] = static_cast<int32_t>(static_cast<long long>(tp.at(t, 0)));   // .astype(np.int32), :161
      size_t p = 0;
      for (long sd = 0; sd < seeds; sd++) {
        const size_t B = am[sd].shape[1];
        const int16_t* m = reinterpret_cast<const int16_t*>(am[sd].data.data()) + t * B;
        const int16_t* c = reinterpret_cast<const int16_t*>(ac[sd].data.data()) + t * B * 2;
        const int16_t* k = reinterpret_cast<const int16_t*>(ak[sd].data.data()) + t * B * 2;
        for (size_t b = 0; b < B; b++, p++) {
          const size_t o = (T0 + t) * P + p;
          merged[o] = static_cast<int8_t>(m[b]);
          cdur[o] = c[2 * b]; creg[o] = static_cast<int8_t>(c[2 * b + 1]);
          kdur[o] = k[2 * b]; kreg[o] = static_cast<int8_t>(k[2 * b + 1]);
        }
      }
    }
    auto app = [&](std::vector<int16_t>& dst, const hygio::Table& e) {
      for (size_t i = 0; i < e.v.size(); i++) dst.push_back(static_cast<int16_t>(e.v[i]));
    };
    app(ntc, e_ntc); app(ntk, e_ntk); app(nmc, e_nmc); app(nmk, e_nmk);
    processed++;
    std::printf("Successfully processed batch %ld\n", batch);
  }
  std::printf("\nProcessing complete. Successfully processed %ld batches\n", processed);
  if (pos.empty()) { std::printf("No data was processed. Check the input directories and file paths.\n"); return 1; }
  const size_t T = pos.size();
  for (size_t i = 0; i < T * P; i++)
    if (creg[i] < 0 || creg[i] > 7 || kreg[i] < 0 || kreg[i] > 7) throw Error("regime labels outside 0..7");

  // split probabilities = mean(merged == 0) over all seeds' trajectories (:125,173) and, on request, the regime frequencies
  // (:208-214): one pass of the site-statistics kernel
  Ctx ctx;
  const uint32_t R = 8;
  std::vector<double> split(T), nul(T), cf, kf;
  if (compute_freqs) { cf.resize(T * R); kf.resize(T * R); }
  float ms = 0.0f;
  ctx.check(hyg_tg_site_statistics(ctx.c, T, static_cast<uint32_t>(P), R, merged.data(), creg.data(), kreg.data(), 0, split.data(), nul.data(),
                                   compute_freqs ? cf.data() : nullptr, compute_freqs ? kf.data() : nullptr, nullptr, &ms), "hyg_tg_site_statistics");
  std::fprintf(stderr, "aggregate: %zu sites x %zu trajectories, site statistics %.3f ms on the device\n", T, P, ms);

  std::printf("Concatenating results...\n");
  const std::string o = output_dir + "/";
  write_indexed_matrix(o + "control_regimes_chrom_" + chrom + ".csv.gz", pos, creg.data(), T, P);
  write_indexed_matrix(o + "case_regimes_chrom_" + chrom + ".csv.gz", pos, kreg.data(), T, P);
  write_indexed_matrix(o + "merge_states_chrom_" + chrom + ".csv.gz", pos, merged.data(), T, P);
  {
    hygio::Writer w(o + "split_probs_" + chrom + ".csv.gz");
    std::string buf = "pos\t0\n";
    for (size_t t = 0; t < T; t++) { append_int(buf, pos[t]); buf += '\t'; buf += hygio::py_repr_double(split[t]); buf += '\n'; }
    w.write(buf);
    w.close();
  }
  write_indexed_matrix(o + "n_total_reads_control_chrom_" + chrom + ".csv.gz", pos, ntc.data(), T, Sc);
  write_indexed_matrix(o + "n_total_reads_case_chrom_" + chrom + ".csv.gz", pos, ntk.data(), T, Sk);
  write_indexed_matrix(o + "n_meth_reads_control_chrom_" + chrom + ".csv.gz", pos, nmc.data(), T, Sc);
  write_indexed_matrix(o + "n_meth_reads_case_chrom_" + chrom + ".csv.gz", pos, nmk.data(), T, Sk);
  write_indexed_matrix(o + "control_durations_chrom_" + chrom + ".csv.gz", pos, cdur.data(), T, P);
  write_indexed_matrix(o + "case_durations_chrom_" + chrom + ".csv.gz", pos, kdur.data(), T, P);
  if (compute_freqs) {
    // row-wise value_counts(normalize = True): one column per label that occurs anywhere, labels ascending, empty where absent
    auto write_freq = [&](const std::string& path, const std::vector<double>& f) {
      bool present[8] = {false, false, false, false, false, false, false, false};
      for (size_t t = 0; t < T; t++)
        for (uint32_t r = 0; r < R; r++) present[r] = present[r] || f[t * R + r] > 0.0;
      hygio::Writer w(path);
      std::string buf = "pos";
      for (uint32_t r = 0; r < R; r++) if (present[r]) { buf += '\t'; append_int(buf, r); }
      buf += '\n';
      for (size_t t = 0; t < T; t++) {
        append_int(buf, pos[t]);
        for (uint32_t r = 0; r < R; r++) if (present[r]) { buf += '\t'; if (f[t * R + r] > 0.0) buf += hygio::py_repr_double(f[t * R + r]); }
        buf += '\n';
      }
      w.write(buf);
      w.close();
    };
    write_freq(o + "case_regimes_freq_" + chrom + ".csv", kf);
    write_freq(o + "control_regimes_freq_" + chrom + ".csv", cf);
  }
  return 0;
}

std::string fmt4(double x) {
  char b[64];
  std::snprintf(b, sizeof(b), "%.4f", x);
  return b;
}

int cmd_get_dmps(int argc, char** argv) {
  const std::set<std::string> known = {"fdr_thresholds", "results_dir", "output_dir", "n_regimes", "chrom", "test_regime_combinations"};
  const Args a = parse_args(argc, argv, 2, known, {"test_regime_combinations"});
  std::vector<double> thresholds;
  if (a.has("fdr_thresholds"))
    for (const std::string& s : a.kv.at("fdr_thresholds")) for (double v : parse_list(s, "fdr_thresholds")) thresholds.push_back(v);
  else thresholds = {0.01, 0.05};
  const std::string path = a.str("results_dir", "../test"), output_dir = a.str("output_dir", "../test/dmp"), chrom = a.str("chrom", "21");
  const long Rl = get_int(a, "n_regimes", 6);
  const bool combos = get_bool(a, "test_regime_combinations", false);
  if (Rl < 1 || Rl > 8) throw Error("--n_regimes must be in 1..8");
  const uint32_t R = static_cast<uint32_t>(Rl);
  hygio::mkdirs(output_dir);

  const hygio::IndexedIntMatrix mc = hygio::read_indexed_int_matrix(path + "/control_regimes_chrom_" + chrom + ".csv.gz");
  const hygio::IndexedIntMatrix mk = hygio::read_indexed_int_matrix(path + "/case_regimes_chrom_" + chrom + ".csv.gz");
  if (mc.rows != mk.rows || mc.cols != mk.cols) throw Error("control and case regime matrices differ in shape");
  const size_t T = mc.rows, P = mc.cols;
  if (T == 0 || P == 0) throw Error("empty regime matrices");
  std::vector<int8_t> creg(T * P), kreg(T * P);
  for (size_t i = 0; i < T * P; i++) {
    if (mc.v[i] < 0 || mc.v[i] >= static_cast<int>(R) || mk.v[i] < 0 || mk.v[i] >= static_cast<int>(R)) throw Error("regime label outside 0..n_regimes-1");
    creg[i] = static_cast<int8_t>(mc.v[i]); kreg[i] = static_cast<int8_t>(mk.v[i]);
  }
  // positions come from the index of split_probs_<chrom>.csv.gz (:75-83)
  std::vector<long long> pos;
  {
    const std::string txt = hygio::read_text(path + "/split_probs_" + chrom + ".csv.gz");
    size_t p = txt.find('\n');
    while (p != std::string::npos && p + 1 < txt.size()) {
      pos.push_back(std::strtoll(txt.c_str() + p + 1, nullptr, 10));
      p = txt.find('\n', p + 1);
    }
  }
  if (pos.size() != T) throw Error("split_probs and the regime matrices differ in length");

  Ctx ctx;
  std::vector<double> split(T), nul(T), cf(T * R), kf(T * R), pair;
  if (combos) pair.resize(T * R * R);
  float ms = 0.0f;
  // test statistic 1 - #(control != case) / P and the regime frequencies bincount / P (:63-64,117-126): the site-statistics kernel
  ctx.check(hyg_tg_site_statistics(ctx.c, T, static_cast<uint32_t>(P), R, creg.data(), creg.data(), kreg.data(), 0, split.data(), nul.data(), cf.data(),
                                   kf.data(), combos ? pair.data() : nullptr, &ms), "hyg_tg_site_statistics");
  std::fprintf(stderr, "get_dmps: %zu sites x %zu trajectories, site statistics %.3f ms on the device\n", T, P, ms);

  // weights (:78-79,103-109): 1/3 (diff_1 + diff_2 + diff_3) of the positions, NaN -> 1e5, reciprocal
  std::vector<double> w_fp(T, 1.0), w_fn(T);
  for (size_t t = 0; t < T; t++) {
    double d = 1e+5;
    if (t >= 3) {
      const double d1 = static_cast<double>(pos[t] - pos[t - 1]), d2 = static_cast<double>(pos[t] - pos[t - 2]), d3 = static_cast<double>(pos[t] - pos[t - 3]);
      d = (1.0 / 3.0) * ((d1 + d2) + d3);
    }
    w_fn[t] = 1.0 / d;
  }

  auto write_rows = [&](const std::string& file, const std::vector<size_t>& rows, const std::vector<double>& stat, const double* weights, bool with_regimes) {
    hygio::Writer w(file);
    std::string buf = "chrom,position,null_stats,false_negative_weight";
    if (with_regimes) {
      for (uint32_t r = 0; r < R; r++) buf += ",Control_METEOR_" + std::to_string(r + 1);
      for (uint32_t r = 0; r < R; r++) buf += ",Case_METEOR_" + std::to_string(r + 1);
    }
    buf += '\n';
    for (size_t i : rows) {
      const double fnw = weights ? weights[i] : 1.0;
      buf += chrom; buf += ','; append_int(buf, pos[i]); buf += ',';
      if (with_regimes) {   // float_format = "%.4f"
        buf += fmt4(stat[i]) + "," + fmt4(fnw);
        for (uint32_t r = 0; r < R; r++) buf += "," + fmt4(cf[i * R + r]);
        for (uint32_t r = 0; r < R; r++) buf += "," + fmt4(kf[i * R + r]);
      } else {
        buf += hygio::py_repr_double(stat[i]) + "," + hygio::py_repr_double(fnw);
      }
      buf += '\n';
    }
    w.write(buf);
    w.close();
  };
  auto plain = [&](const std::vector<double>& stat, double thr, const std::string& file, bool with_regimes) {
    uint64_t k = 0;
    double Qk = 0.0, threshold = 0.0;
    ctx.check(hyg_fdr_procedure(ctx.c, T, stat.data(), thr, &k, &Qk, &threshold), "hyg_fdr_procedure");
    std::vector<size_t> rows;
    for (size_t t = 0; t < T; t++) if (stat[t] < threshold) rows.push_back(t);
    write_rows(file, rows, stat, nullptr, with_regimes);
  };
  auto weighted = [&](const std::vector<double>& stat, double thr, const std::string& file, bool with_regimes) {
    std::vector<uint64_t> idx(T);
    uint64_t n_sel = 0;
    double Nk = 0.0;
    ctx.check(hyg_weighted_fdr_procedure(ctx.c, T, stat.data(), thr, w_fp.data(), w_fn.data(), &n_sel, idx.data(), &Nk), "hyg_weighted_fdr_procedure");
    std::vector<size_t> rows(idx.begin(), idx.begin() + static_cast<long>(n_sel));
    std::sort(rows.begin(), rows.end());
    write_rows(file, rows, stat, w_fn.data(), with_regimes);
  };
  std::vector<double> stat_ij(T);
  for (double thr : thresholds) {
    const std::string ts = hygio::py_repr_double(thr);
    plain(nul, thr, output_dir + "/dmp_" + ts + ".csv", true);
    if (combos)
      for (uint32_t i = 0; i < R; i++)
        for (uint32_t j = 0; j < R; j++)
          if (i != j) {
            for (size_t t = 0; t < T; t++) stat_ij[t] = pair[(t * R + i) * R + j];
            plain(stat_ij, thr, output_dir + "/dmp_" + std::to_string(i) + "_" + std::to_string(j) + "_" + ts + ".csv", false);
          }
    weighted(nul, thr, output_dir + "/weighted_dmp_" + ts + ".csv", true);
    if (combos)
      for (uint32_t i = 0; i < R; i++)
        for (uint32_t j = 0; j < R; j++)
          if (i != j) {
            for (size_t t = 0; t < T; t++) stat_ij[t] = pair[(t * R + i) * R + j];
            weighted(stat_ij, thr, output_dir + "/weighted_dmp_" + std::to_string(i) + "_" + std::to_string(j) + "_" + ts + ".csv", false);
          }
  }
  return 0;
}

// ======================================================================================================================
// hygeia preprocess (src/two_group/preprocess_bed.py; polars 1.8.2, src/two_group/requirements.txt:49): BED-format methylation
// calls -> the per-chromosome count matrices `infer` and `estimate_parameters_and_regimes` read.  Host-side ETL, no GPU.
// PARITY UNPINNED: polars is not installable in the build image, so this is a restatement of the script's data flow as read
// (full joins without key coalescing, nulls -> 0, f64::round = half away from zero), checked against an independent pandas
// restatement (oracle/preprocess_oracle.py) and hand-made cases, not against the script itself.
// ======================================================================================================================
std::vector<std::string> split_tab(const char* b, const char* e) {
  std::vector<std::string> out;
  const char* c = b;
  while (c <= e) {
    const char* q = static_cast<const char*>(std::memchr(c, '\t', static_cast<size_t>(e - c)));
    if (!q) q = e;
    out.emplace_back(c, q);
    c = q + 1;
  }
  return out;
}

struct CollapsedSite { long long start; double total, avg; };

// read_bed_file + collapse_strands (:124-263): rows of this chromosome with ref_genotype CG, the two strands of a CpG joined on
// (+).end == (-).start, coverage summed, methylation percentage coverage-weighted, position = (+).start or (-).start - 1
std::vector<CollapsedSite> read_bed_collapsed(const std::string& path, const std::string& chrom) {
  const std::string txt = hygio::read_text(path);
  struct Row { long long start, end; double cov, pm; };
  std::vector<Row> pos, neg;
  size_t p = txt.find('\n');                       // skip_rows = 1: the header line
  p = (p == std::string::npos) ? txt.size() : p + 1;
  size_t line = 1;
  while (p < txt.size()) {
    size_t eol = txt.find('\n', p);
    if (eol == std::string::npos) eol = txt.size();
    const char* b = txt.data() + p;
    const char* e = txt.data() + eol;
    p = eol + 1;
    line++;
    if (e > b && e[-1] == '\r') e--;
    if (b == e) continue;
    const std::vector<std::string> f = split_tab(b, e);
    if (f.size() < 12) throw Error(path + ": line " + std::to_string(line) + " has fewer than 12 tab-separated fields");
    if (f[0] != chrom || f[11] != "CG") continue;
    Row r;
    r.start = std::strtoll(f[1].c_str(), nullptr, 10); r.end = std::strtoll(f[2].c_str(), nullptr, 10);
    r.cov = std::strtod(f[9].c_str(), nullptr); r.pm = std::strtod(f[10].c_str(), nullptr);
    if (f[5] == "+") pos.push_back(r);
    else if (f[5] == "-") neg.push_back(r);
  }
  std::map<long long, size_t> neg_by_start;
  for (size_t i = 0; i < neg.size(); i++) neg_by_start.emplace(neg[i].start, i);   // first occurrence wins; CpG calls are unique per strand
  std::vector<char> used(neg.size(), 0);
  std::vector<CollapsedSite> out;
  auto emit = [&](bool has_pos, const Row* a, const Row* b2) {
    const double cp = has_pos ? a->cov : 0.0, pp = has_pos ? a->pm : 0.0;
    const double cn = b2 ? b2->cov : 0.0, pn = b2 ? b2->pm : 0.0;
    const double tot = cp + cn;
    if (!(tot > 0.0)) return;                                             // sites without coverage are dropped
    CollapsedSite c;
    c.start = has_pos ? a->start : b2->start - 1;
    c.total = tot;
    c.avg = ((cp * pp) + (cn * pn)) / tot;
    out.push_back(c);
  };
  for (const Row& r : pos) {
    auto it = neg_by_start.find(r.end);
    if (it != neg_by_start.end()) { used[it->second] = 1; emit(true, &r, &neg[it->second]); }
    else emit(true, &r, nullptr);
  }
  for (size_t i = 0; i < neg.size(); i++) if (!used[i]) emit(false, nullptr, &neg[i]);
  std::stable_sort(out.begin(), out.end(), [](const CollapsedSite& x, const CollapsedSite& y) { return x.start < y.start; });
  return out;
}

int cmd_preprocess(int argc, char** argv) {
  const std::set<std::string> known = {"cpg_file_path", "output_path", "case_data_path", "case_id_names", "control_data_path", "control_id_names", "chromosome", "verbose"};
  const Args a = parse_args(argc, argv, 2, known, {"verbose"});
  if (!a.has("cpg_file_path")) throw Error("Required flag --cpg_file_path not provided");
  const std::string cpg_path = a.str("cpg_file_path", ""), out_dir = a.str("output_path", "../test"), chrom = a.str("chromosome", "22");
  auto multi = [&](const char* k) { return a.has(k) ? a.kv.at(k) : std::vector<std::string>(); };
  const std::vector<std::string> case_paths = multi("case_data_path"), control_paths = multi("control_data_path");
  std::vector<std::string> case_ids = multi("case_id_names"), control_ids = multi("control_id_names");
  if (case_paths.empty() && control_paths.empty()) throw Error("Must provide either case samples, control samples, or both");
  if (!case_ids.empty() && case_ids.size() != case_paths.size()) throw Error("Number of case data paths must match number of case ID names");
  if (!control_ids.empty() && control_ids.size() != control_paths.size()) throw Error("Number of control data paths must match number of control ID names");
  if (!path_exists(cpg_path)) throw Error("CpG file not found: " + cpg_path);
  hygio::mkdirs(out_dir);

  // load_cpg_sites (:97-122): tab-separated with a header; rows whose seqID equals the chromosome; Pos0 = start - 1
  std::vector<long long> pos0;
  {
    const std::string txt = hygio::read_text(cpg_path);
    size_t eol = txt.find('\n');
    if (eol == std::string::npos) throw Error(cpg_path + ": empty file");
    const char* hb = txt.data();
    const char* he = txt.data() + eol;
    if (he > hb && he[-1] == '\r') he--;
    const std::vector<std::string> hdr = split_tab(hb, he);
    size_t c_seq = SIZE_MAX, c_start = SIZE_MAX;
    for (size_t i = 0; i < hdr.size(); i++) { if (hdr[i] == "seqID") c_seq = i; if (hdr[i] == "start") c_start = i; }
    if (c_seq == SIZE_MAX || c_start == SIZE_MAX) throw Error(cpg_path + ": columns seqID and start are required");
    size_t p = eol + 1;
    while (p < txt.size()) {
      size_t e2 = txt.find('\n', p);
      if (e2 == std::string::npos) e2 = txt.size();
      const char* b = txt.data() + p;
      const char* e = txt.data() + e2;
      p = e2 + 1;
      if (e > b && e[-1] == '\r') e--;
      if (b == e) continue;
      const std::vector<std::string> f = split_tab(b, e);
      if (f.size() <= std::max(c_seq, c_start)) continue;
      if (f[c_seq] == chrom) pos0.push_back(std::strtoll(f[c_start].c_str(), nullptr, 10) - 1);
    }
  }
  if (pos0.empty()) throw Error("No CpG sites found for chromosome " + chrom);
  std::stable_sort(pos0.begin(), pos0.end());
  const size_t T = pos0.size();
  std::map<long long, size_t> row_of;
  for (size_t i = 0; i < T; i++) row_of.emplace(pos0[i], i);

  // process_sample_data (:265-358): per sample, methylated = round(coverage x pct / 100), unmethylated = round(coverage x (100 -
  // pct) / 100), joined onto the CpG list; a call at a position that is not in the list ends up with a null position and is dropped
  // (:379-383); a listed site without a call is null -> 0 (:402).  NaN marks "null" until then.
  const double NaN = std::nan("");
  bool any_null = false;
  auto group = [&](const std::vector<std::string>& paths, std::vector<double>& meth, std::vector<double>& unmeth) {
    const size_t S = paths.size();
    meth.assign(T * S, NaN); unmeth.assign(T * S, NaN);
    for (size_t s = 0; s < S; s++) {
      if (!path_exists(paths[s])) { std::fprintf(stderr, "File not found: %s\n", paths[s].c_str()); continue; }
      for (const CollapsedSite& c : read_bed_collapsed(paths[s], chrom)) {
        auto it = row_of.find(c.start);
        if (it == row_of.end()) continue;
        meth[it->second * S + s] = std::round(c.total * c.avg / 100.0);
        unmeth[it->second * S + s] = std::round(c.total * (100.0 - c.avg) / 100.0);
      }
    }
    for (double v : meth) any_null = any_null || std::isnan(v);
  };
  std::vector<double> mc, uc, mk, uk;
  group(control_paths, mc, uc);
  group(case_paths, mk, uk);

  // save_results (:430-470): np.savetxt(fmt = '%s', delimiter = ','); a frame with nulls went through float64 ("12.0"), one without
  // stays integer ("12")
  auto cell = [&](double v) {
    if (std::isnan(v)) v = 0.0;
    std::string t = std::to_string(static_cast<long long>(v));
    return any_null ? t + ".0" : t;
  };
  auto save = [&](const std::string& name, const std::vector<double>& x, const std::vector<double>* y, size_t S) {
    hygio::Writer w(out_dir + "/" + name + "_" + chrom + ".txt.gz");
    std::string buf;
    for (size_t t = 0; t < T; t++) {
      for (size_t s = 0; s < S; s++) {
        double v = std::isnan(x[t * S + s]) ? 0.0 : x[t * S + s];
        if (y) v += std::isnan((*y)[t * S + s]) ? 0.0 : (*y)[t * S + s];
        if (s) buf += ',';
        buf += cell(v);
      }
      buf += '\n';
      if (buf.size() > (1u << 20)) { w.write(buf); buf.clear(); }
    }
    w.write(buf);
    w.close();
  };
  {
    hygio::Writer w(out_dir + "/positions_" + chrom + ".txt.gz");
    std::string buf;
    for (size_t t = 0; t < T; t++) { append_int(buf, pos0[t]); buf += '\n'; }
    w.write(buf);
    w.close();
  }
  { hygio::Writer w(out_dir + "/cpg_sites_merged_" + chrom + ".txt.gz"); w.write(std::to_string(T) + "\n"); w.close(); }
  if (!control_paths.empty()) { save("n_methylated_reads_control", mc, nullptr, control_paths.size()); save("n_total_reads_control", mc, &uc, control_paths.size()); }
  if (!case_paths.empty()) { save("n_methylated_reads_case", mk, nullptr, case_paths.size()); save("n_total_reads_case", mk, &uk, case_paths.size()); }
  std::printf("Successfully processed %zu CpG sites for chromosome %s\n", T, chrom.c_str());
  return 0;
}

// ======================================================================================================================
// hygeia get_chrom_segments (src/two_group/get_chrom_segments.py): the list of `infer` batches of a chromosome
// ======================================================================================================================
int cmd_get_chrom_segments(int argc, char** argv) {
  const Args a = parse_args(argc, argv, 2, {"input_file", "chromosome", "segment_size", "output_csv"}, {});
  const std::string in = a.str("input_file", "positions.txt"), chrom = a.str("chromosome", "22"), out = a.str("output_csv", "chrom_segments.csv");
  const long segment = get_int(a, "segment_size", 100000);
  if (segment <= 0) throw Error("--segment_size must be positive");
  const hygio::Table t = hygio::read_csv_numeric(in, false);             // pd.read_csv(header=None): every line is a position
  const size_t n_segments = 1 + t.rows / static_cast<size_t>(segment);    // :33
  hygio::mkdirs_for_file(out);
  hygio::Writer w(out);
  std::string buf = "chrom,segment_index\n";
  for (size_t i = 0; i < n_segments; i++) buf += chrom + "," + std::to_string(i) + "\n";
  w.write(buf);
  w.close();
  std::printf("Segment information saved to %s\n", out.c_str());
  return 0;
}

// ======================================================================================================================
// hygeia make_bed_file (src/single_group/bin/make_bed_file:19-75)
// ======================================================================================================================
int cmd_make_bed(int argc, char** argv) {
  const Args a = parse_args(argc, argv, 2, {"chr", "regimes_file", "output_file"}, {});
  const std::string chr = a.str("chr", ""), in = a.str("regimes_file", ""), out = a.str("output_file", "");
  if (in.empty() || out.empty()) throw Error("--regimes_file and --output_file are required");
  const hygio::Table t = hygio::read_csv_numeric(in, true);
  size_t pos_col = SIZE_MAX;
  std::vector<size_t> reg_cols;
  for (size_t c = 0; c < t.header.size(); c++) {
    if (t.header[c] == "genomic_position") pos_col = c;
    else reg_cols.push_back(c);
  }
  if (pos_col == SIZE_MAX) throw Error("no genomic_position column in " + in);
  static const char* colours[] = {"248,118,109", "183,159,0", "0,186,56", "0,191,196", "97,156,255", "245,100,227", "128,128,128"};
  struct Row { long start; std::string line; };
  std::vector<Row> rows(t.rows);
  for (size_t r = 0; r < t.rows; r++) {
    double best = -HUGE_VAL;
    size_t arg = 0, ties = 0;
    for (size_t k = 0; k < reg_cols.size(); k++) {
      const double v = t.at(r, reg_cols[k]);
      if (v > best) { best = v; arg = k; }
    }
    for (size_t k = 0; k < reg_cols.size(); k++) ties += (t.at(r, reg_cols[k]) == best);
    const bool tie = ties > 1;
    const std::string name = tie ? "equiprobable" : t.header[reg_cols[arg]];
    // colours are keyed by (regime columns..., equiprobable); the reference's table has 6 + 1 entries
    const size_t ci = tie ? 6 : arg;
    const std::string rgb = ci < 7 ? colours[ci] : "";
    const long gp = static_cast<long>(t.at(r, pos_col));
    rows[r].start = gp - 1;
    rows[r].line = chr + "\t" + std::to_string(gp - 1) + "\t" + std::to_string(gp + 1) + "\t" + name + "\t" + hygio::fixed_double(best) + "\t.\t" +
                   std::to_string(gp - 1) + "\t" + std::to_string(gp + 1) + "\t" + rgb + "\n";
  }
  std::stable_sort(rows.begin(), rows.end(), [](const Row& x, const Row& y) { return x.start < y.start; });   // setkey(bed, chr, start)
  hygio::mkdirs_for_file(out);
  hygio::Writer w(out);
  std::string chunk;
  for (const Row& r : rows) {
    chunk += r.line;
    if (chunk.size() > (1u << 20)) { w.write(chunk); chunk.clear(); }
  }
  w.write(chunk);
  w.close();
  std::fprintf(stderr, "Completed processing for chromosome %s\n", chr.c_str());
  return 0;
}

// hidden helpers for the CPU-side tests of the file formats (no GPU needed)
int cmd_selftest(int argc, char** argv) {
  const std::string what = argc > 2 ? argv[2] : "";
  if (what == "format") {          // hygeia _selftest format 1 10 100  -> one formatted value per line, between '|'
    std::vector<double> x;
    for (int i = 3; i < argc; i++) x.push_back(std::strtod(argv[i], nullptr));
    for (const std::string& s : hygio::r_format_fixed(x)) std::printf("|%s|\n", s.c_str());
    return 0;
  }
  if (what == "readr" || what == "pyrepr") {
    for (int i = 3; i < argc; i++) {
      const double v = std::strtod(argv[i], nullptr);
      std::printf("%s\n", (what == "readr" ? hygio::readr_double(v) : hygio::py_repr_double(v)).c_str());
    }
    return 0;
  }
  if (what == "read" && argc > 4) {  // hygeia _selftest read <file> <header 0|1>  -> shape, header, checksum
    const hygio::Table t = hygio::read_csv_numeric(argv[3], std::atoi(argv[4]) != 0);
    double sum = 0.0;
    for (double v : t.v) sum += v;
    std::printf("%zu %zu %zu %.17g\n", t.rows, t.cols, t.header.size(), sum);
    return 0;
  }
  if (what == "npz" && argc > 3) {   // writes a small int16 and float32 archive next to argv[3]
    const int16_t a[6] = {1, -2, 3, 4, 5, 32767};
    const float b[3] = {0.25f, 0.5f, 1.0f};
    hygio::save_npz(std::string(argv[3]) + "_i2.npz", "<i2", {3, 2}, a, sizeof(a));
    hygio::save_npz(std::string(argv[3]) + "_f4.npz", "<f4", {3}, b, sizeof(b));
    const double c[4] = {12.0, 3.0, 10000.0, 0.0};
    hygio::savetxt_e18(std::string(argv[3]) + "_txt.csv.gz", c, 2, 2);
    return 0;
  }
  if (what == "loadnpz" && argc > 3) {   // hygeia _selftest loadnpz <file>  -> descr, shape, sum of the (int16 / int32 / float32 / float64) payload
    const hygio::NpyArray a = hygio::load_npz(argv[3]);
    std::string shape;
    size_t n = 1;
    for (size_t d : a.shape) { shape += (shape.empty() ? "" : "x") + std::to_string(d); n *= d; }
    double sum = 0.0;
    for (size_t i = 0; i < n; i++) {
      if (a.descr == "<i2") sum += reinterpret_cast<const int16_t*>(a.data.data())[i];
      else if (a.descr == "<i4") sum += reinterpret_cast<const int32_t*>(a.data.data())[i];
      else if (a.descr == "<f4") sum += reinterpret_cast<const float*>(a.data.data())[i];
      else if (a.descr == "<f8") sum += reinterpret_cast<const double*>(a.data.data())[i];
    }
    std::printf("%s %s %.17g\n", a.descr.c_str(), shape.c_str(), sum);
    return 0;
  }
  if (what == "readmatrix" && argc > 3) {   // hygeia _selftest readmatrix <file.tsv[.gz]>  -> rows, cols, index name, sums
    const hygio::IndexedIntMatrix m = hygio::read_indexed_int_matrix(argv[3]);
    long long si = 0, sv = 0;
    for (long long v : m.index) si += v;
    for (int16_t v : m.v) sv += v;
    std::printf("%zu %zu %s %lld %lld\n", m.rows, m.cols, m.index_name.c_str(), si, sv);
    return 0;
  }
  throw Error("unknown self test");
}

void show_help() {
  std::printf("Usage: hygeia [command] [arguments...]\n\nAvailable commands:\n"
              "  estimate_parameters_and_regimes   - Estimate parameters and regimes (single group)\n"
              "  preprocess                        - BED-format methylation calls -> per-chromosome count matrices\n"
              "  get_chrom_segments                - List the segments (batches) of a chromosome for infer\n"
              "  infer                             - Two-group (case/control) inference for one chromosome segment\n"
              "  aggregate                         - Aggregate the two-group results of all segments and seeds of a chromosome\n"
              "  get_dmps                          - Differentially methylated positions at given FDR thresholds\n"
              "  make_bed_file                     - Create a BED file\n\nOther options:\n"
              "  version, -v, --version            - Display version information\n"
              "  help, -h, --help                  - Display this help message\n");
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) { show_help(); return 1; }
  const std::string cmd = argv[1];
  try {
    if (cmd == "version" || cmd == "-v" || cmd == "--version") {
      const char* v = std::getenv("HYGEIA_VERSION");
      std::printf("hygeia version %s\n", (v && *v) ? v : HYGEIA_CLI_VERSION);
      return 0;
    }
    if (cmd == "help" || cmd == "-h" || cmd == "--help") { show_help(); return 0; }
    if (cmd == "estimate_parameters_and_regimes") return cmd_single_group(argc, argv);
    if (cmd == "infer") return cmd_infer(argc, argv);
    if (cmd == "preprocess") return cmd_preprocess(argc, argv);
    if (cmd == "get_chrom_segments") return cmd_get_chrom_segments(argc, argv);
    if (cmd == "aggregate") return cmd_aggregate(argc, argv);
    if (cmd == "get_dmps") return cmd_get_dmps(argc, argv);
    if (cmd == "make_bed_file") return cmd_make_bed(argc, argv);
    if (cmd == "_selftest") return cmd_selftest(argc, argv);
    std::fprintf(stderr, "Error: Invalid command '%s'\nValid commands are: estimate_parameters_and_regimes preprocess get_chrom_segments infer aggregate get_dmps make_bed_file\nUse 'hygeia help' for more information\n", cmd.c_str());
    return 2;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "hygeia %s: error: %s\n", cmd.c_str(), e.what());
    return 1;
  }
}
