/* C ABI of one compiled process library  (libmfp_<process>.so).
 *
 * One library per MG5 subprocess, produced by madflow_b200.codegen from the process IR
 * (built-in generator or the `pyout` MG5 plugin in CUDA mode).  It replaces what the reference
 * builds per subprocess with `--custom_op`:  the TensorFlow custom op `Matrix<proc>` registered by
 * python_package/madflow/custom_op/constants.py:104-161 and driven once PER HELICITY from
 * `cusmatrix` (custom_op/generation.py:234-294).  Here one call evaluates all helicities.
 *
 * Conventions: plain pointers and sizes; `d_` pointers are DEVICE addresses, others host;
 * the caller owns every buffer; every call is asynchronous on `stream` (a cudaStream_t passed as
 * void*, NULL = default stream) unless stated; return 0 on success, negative on error with the
 * text available from mfp_last_error().  No global state besides the last-error string.
 */
#ifndef MADFLOW_B200_PROCESS_H
#define MADFLOW_B200_PROCESS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MFP_MAX_PARAMS 8     /* masses and widths (real)            */
#define MFP_MAX_COUPLINGS 8  /* couplings (complex)                 */
#define MFP_MAX_OUT 8        /* outgoing particles                  */
#define MFP_MAX_CUTS 16
#define MFP_MAX_CHANNELS 32  /* initial-state flavour pairs of one subprocess (incl. mirrored) */
#define MFP_LAYOUT_AOS 0     /* (nevt, nexternal, 4)  -- the reference's layout, phasespace.py:4-5 */
#define MFP_LAYOUT_SOA 1     /* (nexternal, 4, nevt)  -- coalesced                                 */

typedef struct mfp_info {
  char name[64];            /* MG5 shell string, e.g. "1_gg_ttx"  (matrix_method_python.inc:77)  */
  int32_t nexternal, ninitial, ncomb, ncolor, ndiags, namps, nwavefuncs;
  int32_t nparams, ncouplings;
  int32_t ndim;             /* 4*(nexternal-2)+2 random numbers per event (madflow_exec.py:360)  */
  int32_t block_threads;    /* CUDA block size the kernels were compiled for                     */
  double denominator;       /* helicity/colour/identical-particle average (inc:104)              */
  double flops_per_event;   /* algorithmic FP64 flop of smatrix, reference call list */
} mfp_info;

int mfp_get_info(mfp_info* out);
/* names of the real parameters / complex couplings, in the order smatrix expects them
 * (sorted masses+widths, then sorted couplings: PyOut_exporter.py:244,407)                      */
const char* mfp_param_name(int i);
const char* mfp_coupling_name(int i);
/* coupling i = (re + i*im) * G^power, G = 2 sqrt(pi alpha_s)   (parameters.py:13-15)            */
int mfp_coupling_def(int i, double* re, double* im, int* power);
/* helicity table entry (generated order)                                                         */
int mfp_helicity(int icomb, int leg);

/* Matrix_<proc>.smatrix (matrix_method_python.inc:80-104): |M|^2 summed over helicities and
 * colours, averaged.  d_p: momenta (E,px,py,pz) in `layout`; par: nparams host doubles;
 * d_coup: complex couplings, interleaved re/im, shape (ncouplings, nevt) if coup_stride==1 or
 * (ncouplings, 1) if coup_stride==0 (frozen model, parameters.py:44-53);
 * sqh: the sqrt(1/2) to use (reference: float32-rounded, wavefunctions_flow.py:10);
 * d_out: nevt doubles.                                                                           */
int mfp_smatrix(const double* d_p, int layout, int64_t nevt, const double* par, const double* d_coup,
                int64_t coup_stride, double sqh, double* d_out, void* stream);

/* Same call with HOST buffers (pageable or pinned): copies in, runs, copies out, synchronises. */
int mfp_smatrix_host(const double* h_p, int layout, int64_t nevt, const double* par, const double* h_coup,
                     int64_t coup_stride, double sqh, double* h_out);

/* Per-helicity values Matrix_<proc>.matrix (inc:106-138) for one helicity row; test hook.      */
int mfp_matrix_hel(const double* d_p, int layout, int64_t nevt, int icomb, const double* par,
                   const double* d_coup, int64_t coup_stride, double sqh, double* d_out, void* stream);

typedef struct mfp_cut {
  int32_t var;        /* 0 = pt, 1 = mt, 2 = mt2   (phasespace.py:405-422); extension for PAIRS of particles:
                       * 3 = invariant mass, 4 = Delta R, with particle = i + 256 * j */
  int32_t particle;   /* index into the nexternal momenta                     */
  int32_t has_min, has_max;
  double vmin, vmax;  /* strict: vmin < var < vmax (phasespace.py:444-461)   */
} mfp_cut;

typedef struct mfp_integrand_args {
  /* VEGAS sampling ([EXT] vegasflow; DESIGN.md "VEGAS") */
  const double* d_grid;     /* (ndim, 51) bin edges                                              */
  uint64_t seed;            /* Philox key                                                        */
  uint32_t iteration;       /* Philox counter word 2                                             */
  uint64_t first_event;     /* global index of this call's first event (Philox counter 0-1)      */
  int64_t nevents;          /* events this call generates                                        */
  double inv_total_events;  /* 1/N of the whole iteration: xjac = vegas weight / N               */
  /* phase space (phasespace.py:359-520) */
  double com_sqrts;
  double masses[MFP_MAX_OUT];
  int32_t lab_frame;        /* boost to the lab before the matrix element (com_output=False)     */
  int32_t ncuts;
  mfp_cut cuts[MFP_MAX_CUTS];
  double pi, acc, gev2pb;   /* constants: reference (float32-rounded) or exact                   */
  /* model (parameters.py) */
  double par[MFP_MAX_PARAMS];
  int32_t alpha_mode;       /* 0: frozen alpha_s;  1: one-loop running at q2 = (sum mT / 2)^2;
                             * 2: the PDF set's alpha_s table at q2 (pdf.alphasQ2, madflow_exec.py:431) */
  double alpha_s;           /* frozen value, or alpha_s(mz2) for the running                      */
  double mz2, b0;           /* running: alpha_s / (1 + alpha_s*b0*log(q2/mz2))                    */
  double sqh;
  /* output */
  double* d_partial;        /* (nblocks, 4 + ndim*50) block partials, fully overwritten:
                             * [sum t, sum t^2, #events that reached the matrix element, 0] + hist */
  int32_t nblocks;          /* grid size, from mfp_integrand_blocks()                             */
  int32_t accumulate_hist;  /* 0 when the grid is frozen                                          */
  /* scratch in HBM for the accepted events of this call (three-stage pipeline of the
   * helicity-parallel flavour); size from mfp_integrand_workspace(); may be NULL/0 otherwise   */
  void* d_workspace;
  int64_t workspace_bytes;
  /* parton luminosity (madflow_exec.py:410-417, 450-454): sum over the subprocess's initial-state flavour
   * pairs of xf_a(x1,q2) xf_b(x2,q2) / x1 / x2, multiplied into the event weight.  d_pdf: one member of an
   * lhagrid1 set packed by mf_pdf_* (madflow_b200.h; layout in csrc/pdf.cuh); NULL = --no_pdf, luminosity 1
   * (:437-438).  chan_fl1/2: column of the flavour in the table, for hadron 1 / 2.                          */
  const double* d_pdf;
  int32_t nchannels;
  int8_t chan_fl1[MFP_MAX_CHANNELS], chan_fl2[MFP_MAX_CHANNELS];
  double fixed_q2;          /* > 0: muF^2 = muR^2 fixed (madflow -q, :376-386); else q2 = (sum mT/2)^2    */
  /* several subprocesses on the same events (`p p > ...`: ret += luminosity_i * smatrix_i, :444-455): every
   * library runs generation + matrix element into its own workspace with skip_accumulate = 1, then
   * mf_vegas_accumulate_sum (madflow_b200.h) sums the terms per event.  Helicity-parallel flavour only.    */
  int32_t skip_accumulate;
} mfp_integrand_args;

/* recommended persistent grid size for the current device (multiple of the SM count)          */
int mfp_integrand_blocks(void);
/* Fix the grid size mfp_integrand_blocks() reports for the helicity-parallel flavour (0 = back to automatic).
 * The event buffer is cut into 4 segments per block, so libraries that evaluate different subprocesses on the
 * SAME events (skip_accumulate, below) must agree on it: slot i then holds the same event in all of them.   */
int mfp_set_integrand_blocks(int nblocks);
/* bytes of device scratch mfp_integrand needs for a call generating `nevents` events (0 for the
 * one-event-per-thread flavour, whose single kernel keeps everything on chip)                   */
int64_t mfp_integrand_workspace(int64_t nevents);
/* One pass of the integrand of scripts/madflow_exec.py:422-470 over `nevents` events:
 * Philox -> VEGAS map -> x1,x2 -> RAMBO -> cuts -> boost -> scale, alpha_s, luminosity -> smatrix -> weight ->
 * block partial sums of xjac*f, (xjac*f)^2 and the per-dimension histogram of (xjac*f)^2.      */
int mfp_integrand(const mfp_integrand_args* args, void* stream);

/* The events of the LAST mfp_integrand call on this workspace (helicity-parallel flavour): the slots
 * that reached the matrix element, grouped in segments; slots of weight 0 are padding.  The integrand value
 * of slot i is d_me[i] * d_weight[i] (the weight includes the parton luminosity).  What the reference passes to its LHE writer from inside the integrand
 * (scripts/madflow_exec.py:462-464: all_ps and weight * ret).  Returns -2 for the one-event-per-thread
 * flavour, whose single kernel keeps the events on chip.                                                   */
typedef struct mfp_event_view {
  const double* d_mom;      /* (capacity, nexternal, 4) momenta the matrix element was evaluated on       */
  const double* d_weight;   /* (capacity) xjac * phase-space weight (0 = empty slot)                       */
  const double* d_me;       /* (capacity) |M|^2                                                             */
  const double* d_alpha_s;  /* (capacity) alpha_s of the event                                              */
  int64_t capacity;
  const uint8_t* d_bins;    /* (ndim, capacity) VEGAS bin of every dimension                                */
} mfp_event_view;
int mfp_integrand_events(void* d_workspace, int64_t nevents, mfp_event_view* out);

/* Kernel flavour (DESIGN.md "Kernel mapping"): 0 = the process's default, 1 = one event per thread,
 * 2 = helicity-parallel thread blocks.  Both flavours compute the same numbers; the switch exists so
 * that the choice per process is made on measurements.  mfp_get_variant returns 1 or 2.            */
int mfp_set_variant(int variant);
int mfp_get_variant(void);

const char* mfp_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
