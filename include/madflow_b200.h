/* C ABI of the process-independent library  (libmadflow_b200.so).
 *
 * HELAS external wavefunctions, ALOHA vertices (test hooks), RAMBO phase space with cuts and
 * boost, Philox/VEGAS sampling, accumulation and grid refinement, PDF / alpha_s interpolation, and an FP64
 * peak probe.
 * It replaces the TensorFlow graphs of python_package/madflow/wavefunctions_flow.py and
 * phasespace.py and the vegasflow calls of scripts/madflow_exec.py:487-525 (reference has no
 * native ABI for these; its only native boundary is the per-process TF custom op, see
 * madflow_b200_process.h).
 *
 * Conventions as in madflow_b200_process.h: `d_` = device pointer, caller owns all memory,
 * asynchronous on `stream` (cudaStream_t as void*), 0 = success, negative = error,
 * mf_last_error() gives the text.
 */
#ifndef MADFLOW_B200_H
#define MADFLOW_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MF_VEGAS_BINS 50
#define MF_VEGAS_HEADER 4   /* accumulators: sum t, sum t^2, #events with non-zero t, reserved */
#define MF_MAX_DIM 32

const char* mf_last_error(void);
int mf_version(void);

/* ---- HELAS external wavefunctions (wavefunctions_flow.py:33-154) -------------------------------
 * kind: 0 ixxxxx, 1 oxxxxx, 2 vxxxxx, 3 sxxxxx.  d_p: (nevt,4) momenta (E,px,py,pz).
 * d_out: (6,nevt) complex128 interleaved ((3,nevt) for sxxxxx), the reference's output layout. */
int mf_wavefunction(int kind, const double* d_p, int64_t nevt, double mass, int nhel, int nsf, double sqh,
                    double* d_out, void* stream);

/* ---- ALOHA vertex routines (test hook; PyOut_create_aloha.py output) ---------------------------
 * id: 0 FFV1_0, 1 FFV1_1, 2 FFV1_2, 3 VVV1P0_1, 4 VVV1_0, 5 FFV1P0_3, 6/7/8 VVVV{1,3,4}_0,
 * 9/10/11 VVVV{1,3,4}P0_1.  Inputs: up to four (6,nevt) complex wavefunctions in the routine's
 * argument order (unused = NULL).  d_out: (nevt,) complex for amplitudes, (6,nevt) otherwise.  */
int mf_aloha(int id, const double* d_a, const double* d_b, const double* d_c, const double* d_d, int64_t nevt,
             double coup_re, double coup_im, double mass, double width, double* d_out, void* stream);

/* ---- phase space (phasespace.py) --------------------------------------------------------------- */
typedef struct mf_cut {
  int32_t var;       /* 0 pt, 1 mt, 2 mt2; pairs (extension): 3 invariant mass, 4 Delta R */
  int32_t particle;  /* pair variables: i + 256 * j */
  int32_t has_min, has_max;
  double vmin, vmax;
} mf_cut;

typedef struct mf_ps_const {
  double pi, acc, gev2pb;   /* reference (float32-rounded) or exact, see DESIGN.md */
} mf_ps_const;

/* rambo (phasespace.py:145-212): d_x (nevt, 4*nout) uniforms; sqrts scalar, or per event when
 * d_sqrts != NULL; masses NULL or nout doubles (host).  d_p (nevt,nout,4), d_w (nevt,).          */
int mf_rambo(int nout, const double* d_x, int64_t nevt, double sqrts, const double* d_sqrts, const double* masses,
             const mf_ps_const* k, double* d_p, double* d_w, void* stream);

/* ramboflow + cuts + boost (phasespace.py:257-319, 480-520): d_x (nevt, 4*(next-2)+2).
 * Outputs for EVERY event (no compaction): d_p (nevt,next,4), d_w, d_x1, d_x2 (nevt,),
 * d_pass (nevt,) uint8 = all cuts passed (evaluated on the COM momenta); lab_frame boosts d_p.   */
int mf_phasespace(int next, const double* d_x, int64_t nevt, double com_sqrts, const double* masses,
                  const mf_ps_const* k, const mf_cut* cuts, int ncuts, int lab_frame, double* d_p, double* d_w,
                  double* d_x1, double* d_x2, uint8_t* d_pass, void* stream);

/* _boost_to_lab (phasespace.py:322-356) in place on (nevt,next,4)                                 */
int mf_boost_to_lab(int next, double* d_p, const double* d_x1, const double* d_x2, int64_t nevt, void* stream);

/* ---- VEGAS ([EXT] vegasflow; DESIGN.md "VEGAS") ------------------------------------------------ */
/* Philox uniforms in [0,1): (nevt, ndim), global event index = first_event + row                  */
int mf_philox_uniform(uint64_t seed, uint32_t iteration, uint64_t first_event, int64_t nevt, int ndim,
                      double* d_out, void* stream);
/* sample: d_x (nevt,ndim) mapped points, d_xjac (nevt,) = vegas weight * inv_total_events,
 * d_bins (ndim,nevt) uint8 bin indices.  d_grid (ndim,51).                                        */
int mf_vegas_sample(const double* d_grid, int ndim, uint64_t seed, uint32_t iteration, uint64_t first_event,
                    int64_t nevt, double inv_total_events, double* d_x, double* d_xjac, uint8_t* d_bins,
                    void* stream);
int mf_vegas_blocks(void);
/* accumulate: t = xjac*f; partial (nblocks, 4+ndim*50): header (see MF_VEGAS_HEADER), histogram of t^2 */
int mf_vegas_accumulate(const double* d_f, const double* d_xjac, const uint8_t* d_bins, int64_t nevt, int ndim,
                        int with_hist, double* d_partial, int nblocks, void* stream);
/* the same with t = sum_i d_f[i][e] * d_w[i][e] over nterms <= 16 subprocesses evaluated on the same events
 * (scripts/madflow_exec.py:444-455); d_f / d_w: host arrays of device pointers                    */
int mf_vegas_accumulate_sum(int nterms, const double* const* d_f, const double* const* d_w, const uint8_t* d_bins,
                            int64_t nevt, int ndim, int with_hist, double* d_partial, int nblocks, void* stream);
/* sums (4+ndim*50) = [add ? sums : 0] + sum over blocks, in fixed order (deterministic)           */
int mf_vegas_reduce(const double* d_partial, int nblocks, int ndim, int add, double* d_sums, void* stream);
/* Lepage refinement (alpha = 1.5) of d_grid in place from the histogram in d_sums[4:]             */
int mf_vegas_refine(double* d_grid, const double* d_sums, int ndim, void* stream);

/* ---- event output (python_package/madflow/lhe_writer.py:151-239, example/compare_mg5_hists.py:16-57) ----
 * The events are the slots of an event buffer in device memory: momenta d_mom (nevt, nexternal, 4) and a
 * weight per slot = d_w1[i] * d_w2[i] (d_w2 may be NULL); slots of weight 0 (cut events, padding) are ignored.
 * mfp_integrand_events() of a process library exposes the buffer of the last integrand call.
 *
 * histogram: d_hist[nbins + 2] += weights, [0] underflow, [nbins + 1] overflow (and NaN); `observable` of
 * particle `particle`: 0 pt, 1 pseudorapidity, 2 rapidity, 3 energy, 4 invariant mass.                     */
int mf_event_histogram(const double* d_mom, const double* d_w1, const double* d_w2, int64_t nevt, int nexternal,
                       int particle, int observable, double lo, double hi, int nbins, double* d_hist, void* stream);
/* weight statistics: d_partial (nblocks, 3) = per block {max |w|, sum |w|, sum w^2}; the caller reduces the rows    */
int mf_weight_stats_blocks(void);
int mf_weight_stats(const double* d_w1, const double* d_w2, int64_t nevt, double* d_partial, int nblocks, void* stream);
/* unweighting + compaction: slot i is kept with probability |w_i| / wmax (Philox: key seed, counter
 * first_index + i) and appended -- momenta, weight sign(w) * max(|w|, wmax), global index first_index + i --
 * at position (*d_count)++ of the output arrays (capacity slots; *d_count keeps counting past it).        */
int mf_select_events(const double* d_mom, const double* d_w1, const double* d_w2, int64_t nevt, int nexternal, double wmax,
                     uint64_t seed, uint64_t first_index, double* d_out_mom, double* d_out_w, int64_t* d_out_index,
                     int32_t* d_count, int64_t capacity, void* stream);

/* ---- PDFs and alpha_s from an LHAPDF lhagrid1 set (what the reference asks pdfflow for:
 * scripts/madflow_exec.py:412-413 pdf.xfxQ2(pids, x, q2), :431 pdf.alphasQ2(q2)) ------------------------
 * d_table: one member packed into doubles (header, subgrid descriptors, knots and their logarithms, values;
 * layout in madflow_b200/csrc/pdf.cuh, packer madflow_b200/pdf.py).  columns: for each requested flavour
 * its column in the table (host ints).  d_out: (nevt, ncolumns) x*f(x, Q2); log-bicubic interpolation
 * (LHAPDF LogBicubicInterpolator), frozen at the grid edges.                                             */
#define MF_PDF_MAX_FLAVOURS 16
int mf_pdf_xfxq2(const double* d_table, const int32_t* columns, int ncolumns, const double* d_x, const double* d_q2,
                 int64_t nevt, double* d_out, void* stream);
/* alpha_s(Q2) from the set's AlphaS_Qs / AlphaS_Vals table (LHAPDF AlphaS_Ipol)                          */
int mf_pdf_alphasq2(const double* d_table, const double* d_q2, int64_t nevt, double* d_out, void* stream);

/* ---- measurement -------------------------------------------------------------------------------
 * FP64 FMA throughput of the current device, measured: `iters` dependent DFMA per chain, 8 chains
 * per thread, full grid.  Synchronous.  Returns TFLOP/s in *tflops (2 flop per DFMA).            */
int mf_fp64_peak(int iters, double* tflops, double* ms);
/* the same for the FP64 tensor instruction (mma.sync.m8n8k4.f64): `iters` x 4 independent accumulators per warp,
 * 4 warps per block, 4 blocks per SM; 512 flop per instruction.  DFMA and DMMA share one pipe on sm_100a.          */
int mf_dmma_peak(int iters, double* tflops, double* ms);

#ifdef __cplusplus
}
#endif
#endif
