// hygeia_b200/csrc/hyg_dev_structs.h -- POD descriptors passed from the host side of the C ABI to the kernels.
#ifndef HYG_DEV_STRUCTS_H
#define HYG_DEV_STRUCTS_H

#include <stdint.h>

#ifndef HYG_RMAX
#define HYG_RMAX 8
#endif
#define HYG_OVL_ROWS 32   // rows of the next segment every segment recomputes for the left-halo check

namespace hyg {

// Model constants for one theta (ModelParameters of singleGroup.h:69-455 after setUnknownParameters).
struct SgModelDev {
  int R;            // number of regimes
  int u;            // minimum sojourn
  int n_particles;  // N_max (<= 256)
  uint32_t dcap;    // table entries per regime; entry dcap-1 is the terminal (steady) value used for every d >= dcap
  double P[HYG_RMAX][HYG_RMAX];     // P[r'][r], zero diagonal
  double logP[HYG_RMAX][HYG_RMAX];  // log P (only used on the exact log-domain path)
  // tab[r * dcap + (d-1)] = { c_new(d,r), lc(d,r) }:
  //   c_new = 0 if d < u; 1 if exit flag; else rho(d,r)          (new-segment factor, singleGroup.h:584-596)
  //   lc    = log(1 - rho(d,r)) if !exit && rho <= 1, else -inf   (continuation,      singleGroup.h:597-605)
  const double2* tab;
  // tabg[r * dcap + (d-1)] = d log rho(d,r) / d theta_omega_r (parameter-estimation mode only; may be null)
  const double* tabg;
};

// One chain = one (dataset, seed) pair.
struct SgChainDev {
  unsigned long long T;
  const double* logobs;   // T x R, device (output of K1)
  const double* unif;     // T injected uniforms (device) or null -> Philox(seed, chain_id, t)
  unsigned long long seed;
  uint32_t chain_id;
  double* probs;          // T x (1+R) rows (position, p_1..p_R), written whole -- device memory, or pinned host memory mapped
                          // into the device address space (the rows then stream to the host as sites are finalised) -- or null
  const uint32_t* pos;    // T genomic positions (device) or null -> global site index
  double* logz;           // T running log Z_t (device) or null
  // optional step-level taps for parity tests (device, may be null)
  int* k_kept;            // T : K of optimal resampling; -1 growth; -2 keep-largest
  unsigned char* drew;    // T : uniform consumed at site t
  int* n_pending;         // T : lag-set size after site t
  int* n_curr;            // T : particle count after site t
  int* finalised_at;      // T : step at which site t was emitted
  unsigned long long* support_hash;   // T : order-independent hash of the finite-weight support {(d, r)} after site t
  unsigned char* tie_flags;           // T : bit 0 exact tie among the sorted weights, bit 1 a tie decided a particle's fate
  int* status;            // 8 ints accumulated over the chain's segments: [0] forced emissions (lag set full), [1] max lag-set size,
                          // [2] owned sites emitted by force at the end of a segment's right halo, [3] sites stepped through,
                          // [4] sites whose sort was redone on the full words (weights equal in their top 56 bits), [5] sites where an exact tie of weights
                          // decided a particle's fate, [6] double systematic draws repaired, [7] reserved
  // Segmented execution (hyg_sg_set_segmentation): this descriptor covers the sites [t_off, t_off + T) of its chain -- every
  // pointer above is already offset to local site 0 -- and OWNS the local sites [own_lo, own_hi): rows outside that range are
  // warm-up (left halo: the filter forgets its initial condition) or run-out (right halo: until every owned site is finalised)
  // and are not written.  Whole-chain execution: t_off = 0, own_lo = 0, own_hi = T, last_segment = 1.
  unsigned long long t_off;
  unsigned long long own_lo, own_hi;
  int last_segment;       // T-1 is the end of the chain (the reference's forced finalisation there is genuine)
  double* seg_inc;        // out: log Z_{own_hi-1} - log Z_{own_lo-1} (device) or null
  double* ovl;            // out: [HYG_OVL_ROWS][R] posterior rows of the first sites AFTER own_hi that this segment finalised in its
                          // right halo (NaN where none) -- compared with the next segment's rows after the launch: an
                          // insufficient LEFT halo of the next segment shows up as a difference.  null: no check
  // parameter-estimation mode
  const double* theta0;   // D initial theta (device) or null
  double* theta_trace;    // T x D (device) or null
};

// Segmented execution: one unit per segment j >= 1 of a chain with a logz output.  The segment's rows hold log Z relative to
// the site before the segment; the fix-up adds seg_inc[0] + ... + seg_inc[j-1] (summed in that order).
struct SgLogzFix {
  double* logz;            // first owned row of the segment
  const double* seg_inc;   // the chain's per-segment increments
  unsigned long long len;  // owned rows
  unsigned int j;          // segment index
  unsigned int pad_;
};

// Segmented execution, left-halo check: compare the rows a segment recomputed in its right halo with the rows the next
// segment wrote (one unit per segment boundary).
struct SgOvlCheck {
  const double* ovl;              // [HYG_OVL_ROWS][R]
  const double* probs;            // the chain's output row of the first site after the boundary: [..][1 + R]
  unsigned long long* max_bits;   // per chain: bits of the largest |difference| (non-negative doubles order like integers)
  int* status;                    // per chain status words ([7] counts rows that differ by more than 1e-6)
  unsigned int rows;              // rows that exist after the boundary (<= HYG_OVL_ROWS)
  unsigned int R;
};

struct SgRunDev {
  int use_smoothing;
  double epsilon;
  int lcap;               // capacity of the lag set per CTA
  double* psi_ws;         // workspace base
  unsigned long long psi_stride;  // doubles per CTA
  unsigned int* queue;    // atomic chain counter
  int n_chains;
  int force_full_sort;    // test hook: every resampling site takes the exact sort on the full (key, regime, sojourn) words
  // parameter-estimation mode (K3)
  int use_param_est;
  int normalise_gradients;
  int use_adam;
  unsigned int n_steps_without_update;
  double lr_exponent;
  double lr_factor;
  double kappa[HYG_RMAX];
  double* pe_ws;          // per-CTA table workspace: tab (double2) | tabg | wh | wg, each R x pe_dcap
  unsigned long long pe_stride;   // doubles per CTA
  uint32_t pe_dcap;
};

}  // namespace hyg
#endif
