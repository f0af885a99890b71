/*
 * fsweep.h — C ABI of libfsweep.so: the B200-native frequency-bin sweep engine.
 *
 * The reference (gdalsanto/flamo, pure Python/PyTorch) has no FFI; its boundary for this
 * path is the Python class API.  This header is the single new seam underneath that API:
 * the host-side mirror (flamo_b200/processor/{dsp,system}.py) lowers a module tree into a
 * flat "sweep program" and calls the two entry points below through ctypes.  Each op kind
 * cites the reference code it replaces (paths relative to /root/reference/flamo/).
 *
 *   fsweep_forward   replaces  Series.forward (processor/system.py:279-301),
 *                              Recursion.forward (system.py:397-425) and every
 *                              freq_response/freq_convolve lambda it reaches
 *                              (processor/dsp.py:466-468, 552-554, 922-924, 1520-1526,
 *                               2226-2232, 2587-2593, 3356-3374, 3406-3408, 3504-3530)
 *   fsweep_backward  replaces  the autograd graph torch builds through those same calls
 *                              (einsum / rfft / prod / div / where / exp / linalg.solve
 *                              backward), as driven by loss.backward() in
 *                              optimize/trainer.py:190
 *   fsweep_forward_loss / fsweep_backward_loss
 *                    replace   the same plus the Shell's |.| output layer and an MSE criterion
 *                              (optimize/loss.py:90-103 or torch.nn.MSELoss): one launch yields
 *                              the loss and its gradients (optimize/trainer.py:177-190)
 *   fsweep_expm_* / fsweep_sparsity_* / fsweep_weighted_total
 *                    replace   the parameter-sized pieces of a training step (dsp.py:649,
 *                              loss.py:36-63, trainer.py:184-188) so that a captured step is a
 *                              handful of launches
 *   fsweep_allreduce_p2p       the one exchange of the multi-GPU step (no reference counterpart)
 *
 * Which kernel family serves a plan is an implementation detail (fsweep_plan_kernel_family):
 * row-distributed interpreter, FDN-loop kernels, thread-per-bin kernels for small loops,
 * CTA-per-bin kernels for wide loops, streaming kernels for TABLE-heavy programs, and the
 * table / deferred-gradient kernels for large section cascades (flamo_b200/csrc/).
 *
 * Conventions
 *   - plain C, no C++ exceptions cross the boundary; every pointer marked "device" is a CUDA
 *     device pointer owned by the caller; the library allocates nothing on the device and
 *     creates no streams; all work is enqueued on the caller's stream and is CUDA-graph
 *     capture safe (no sync, no allocation, no host callback).
 *   - dtype: FSWEEP_C64 computes in float32 / complex64, FSWEEP_C128 in float64 / complex128.
 *     "real" below means float or double accordingly; "cplx" is interleaved (re, im) of real.
 *   - a bin k is the rFFT bin index, omega_k = 2*pi*k/nfft, z^-1 = exp(-j*omega_k); the
 *     anti-aliasing radius gamma = 10^(-|alias_decay_db| / (20*nfft)) (dsp.py:294-307).
 *   - signals: x is (batch, n_bins, n_in, cols) cplx, y is (batch, n_bins, n_out, cols) cplx
 *     (or real when the epilogue is FSWEEP_EPI_ABS); the last three dims are contiguous, the
 *     batch stride is given in elements so bin-range slices of a larger tensor can be passed
 *     without a copy.  x and y address the FIRST PROCESSED bin (absolute index bin_begin).
 */
#ifndef FSWEEP_H
#define FSWEEP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define FSWEEP_API __attribute__((visibility("default")))
#else
#define FSWEEP_API
#endif

#define FSWEEP_VERSION 2

/* return codes */
#define FSWEEP_OK 0
#define FSWEEP_E_BADARG (-1)
#define FSWEEP_E_UNSUPPORTED (-2)
#define FSWEEP_E_WORKSPACE (-3)
#define FSWEEP_E_CUDA (-4)

/* dtypes */
#define FSWEEP_C64 0
#define FSWEEP_C128 1
/* may be OR-ed into FSWEEP_C128 at plan creation: the caller's PARAMETERS are float32 (a float32 model swept in float64
 * arithmetic because float32 arithmetic misses the 1e-4 bar), so coefficient gradients of deferred section cascades may
 * be accumulated in float32 arithmetic (buffers stay float64) */
#define FSWEEP_DT_GRAD32 256

/* epilogues (output layer fused into the sweep) */
#define FSWEEP_EPI_NONE 0 /* y = Y                      (cplx out)                           */
#define FSWEEP_EPI_ABS 1  /* y = |Y|                    (real out)  Transform(torch.abs)      */

/* op kinds.  Coefficient layout per kind ("coef"), and the gradient buffer layout ("grad",
 * same shape and dtype as coef unless noted).  Delays are ALWAYS float64. */
enum fsweep_op_kind {
  /* dense real matrix, constant over bins: y = W x.  dsp.Gain / dsp.Matrix after `map`
   * (dsp.py:466-468, 642-665).  coef: real[n_out][n_in]. */
  FSWEEP_OP_GAIN = 1,
  /* diagonal real gain: y_n = w_n x_n.  dsp.parallelGain (dsp.py:552-554).  coef: real[n]. */
  FSWEEP_OP_PGAIN = 2,
  /* dense cascade of K second-order sections per (out,in) pair:
   *   H[m][n] = prod_s B_s / prod_s A_s,  B_s = b0 + b1*g*z^-1 + b2*g^2*z^-2  (g = gamma),
   *   H = eps if |prod A| == 0   (dsp.py:1520-1526 = 2226-2232 = 2587-2593: Biquad, SVF, GEQ
   *   after their maps; equals the reference's zero-padded rfft of the 3 taps).
   * coef: real[K][n_in][n_out][2][8] "packed sections": two Taylor blocks per section,
   *   block 0 (used for bins with cos(omega_k) >= 0, expansion point w0 = +1):
   *     {b0+b1+b2, b1+2*b2, b2, 0,  a0+a1+a2, a1+2*a2, a2, 0}
   *   block 1 (cos(omega_k) < 0, w0 = -1):
   *     {b0-b1+b2, b1-2*b2, b2, 0,  a0-a1+a2, a1-2*a2, a2, 0}
   * so that B(w) = B(w0) + B'(w0) v + b2 v^2 with v = w - w0 formed without cancellation
   * (w = gamma*z^-1).  The caller forms the sums in float64.  grad has the same layout: each bin
   * contributes d/d{B(w0), B'(w0), b2, A(w0), A'(w0), a2} to its block; the chain rule through the
   * (linear) packing is the caller's. */
  FSWEEP_OP_SOS = 3,
  /* diagonal cascade (parallelBiquad / parallelSVF / parallelGEQ).  coef: real[K][n][2][8]. */
  FSWEEP_OP_PSOS = 4,
  /* dense delays: H[m][n] = gamma^d * exp(-j*omega_k*d), d = coef[m][n] samples
   * (dsp.py:3356-3374).  With FSWEEP_F_ISINT d is rounded and the phase index (k*d mod nfft)
   * is formed in integers; otherwise k*d/nfft is range-reduced in float64.
   * coef: double[n_out][n_in]; grad: double[n_out][n_in] (dL/dd; zero for ISINT). */
  FSWEEP_OP_DELAY = 5,
  /* diagonal delays (dsp.parallelDelay, dsp.py:3504-3530).  coef: double[n]. */
  FSWEEP_OP_PDELAY = 6,
  /* dense precomputed response streamed from HBM (generic dsp.Filter, dsp.py:901-924):
   * coef: cplx[M][n_out][n_in] indexed by ABSOLUTE bin; grad: same (dL/dH, PyTorch
   * convention dL/dRe + j dL/dIm). */
  FSWEEP_OP_TABLE = 7,
  /* diagonal table (dsp.parallelFilter, dsp.py:1021-1023).  coef: cplx[M][n]. */
  FSWEEP_OP_PTABLE = 8,
  /* closed loop y = (I - F*Fb)^-1 F x (system.py:417-425).  The next n_ff ops form the
   * feedforward chain F (applied left to right), the n_fb ops after them the feedback chain
   * Fb.  Owns no coefficient slot.  n_out = loop width, n_in = input channels of F. */
  FSWEEP_OP_RECURSION = 9
};

/* op flags */
#define FSWEEP_F_ISINT 1u /* DELAY/PDELAY: integer delays, exact integer phase          */
#define FSWEEP_F_GRAD 2u  /* a coefficient gradient is wanted for this op in backward   */

typedef struct fsweep_op {
  int32_t kind;       /* enum fsweep_op_kind                                               */
  int32_t n_out;      /* output channels (rows of H)                                       */
  int32_t n_in;       /* input channels  (cols of H); == n_out for diagonal kinds          */
  int32_t n_sections; /* K for SOS/PSOS, else 0                                            */
  uint32_t flags;
  int32_t n_ff;       /* RECURSION only                                                    */
  int32_t n_fb;       /* RECURSION only                                                    */
  int32_t per_item;   /* != 0: the slot holds ONE COEFFICIENT SET PER BATCH ITEM, [batch][numel] (hyper-conditioning:
                         flamo's NN-in-the-loop examples call the module once per item with ext_param,
                         examples/e7_biquad_nn.py:149-156, e4_recursion_nn.py:243-250; here item b of x meets set b
                         inside ONE launch).  A plan with such an op runs the generic kernels with one grid slice per
                         item; its gradient buffer is [batch][numel] as well, the gradients of the other slots are
                         summed over the items as usual.  0: one set for the whole batch.                  */
} fsweep_op_t;

typedef struct fsweep_plan fsweep_plan_t; /* opaque, immutable after create, host memory only */

FSWEEP_API int fsweep_version(void);
FSWEEP_API const char* fsweep_last_error(void); /* thread-local message of the last failing call */

/* ops: flat pre-order program (top-level series; at most one RECURSION, whose chains follow
 * it).  Every non-RECURSION op owns coefficient slot i = its rank among non-RECURSION ops. */
FSWEEP_API int fsweep_plan_create(const fsweep_op_t* ops, int n_ops, int64_t nfft, double alias_decay_db,
                       int dtype, fsweep_plan_t** out);
FSWEEP_API int fsweep_plan_destroy(fsweep_plan_t* plan);
FSWEEP_API int fsweep_plan_num_coeffs(const fsweep_plan_t* plan);
/* number of reals (for TABLE kinds: of cplx) in slot `slot`'s coefficient / gradient buffer
 * given the total bin count M (only TABLE kinds depend on M). */
FSWEEP_API int64_t fsweep_plan_coeff_numel(const fsweep_plan_t* plan, int slot, int64_t M);

/* name of the kernel family a forward (backward != 0: backward) call over n_bins bins will launch for this plan
 * (bench / profile bookkeeping; static string) */
FSWEEP_API const char* fsweep_plan_kernel_family(const fsweep_plan_t* plan, int64_t n_bins, int backward);

/* scratch the caller must provide to fsweep_backward / fsweep_*_loss (fsweep_forward needs none) */
FSWEEP_API size_t fsweep_workspace_bytes(const fsweep_plan_t* plan, int64_t batch, int64_t cols, int64_t n_bins);

FSWEEP_API int fsweep_forward(const fsweep_plan_t* plan,
                   const void* const* coeffs,      /* host array [num_coeffs] of device pointers */
                   const void* x, int64_t x_batch_stride, /* device; stride in cplx elements */
                   void* y, int64_t y_batch_stride,       /* device; stride in output elements */
                   int64_t batch, int64_t cols, int64_t bin_begin, int64_t n_bins,
                   int epilogue, void* stream /* cudaStream_t */);

FSWEEP_API int fsweep_backward(const fsweep_plan_t* plan, const void* const* coeffs,
                    const void* x, int64_t x_batch_stride,
                    const void* grad_y, int64_t gy_batch_stride, /* device; dL/dy, layout of y */
                    void* const* grad_coeffs, /* host array [num_coeffs]; NULL entries skipped;
                                                 buffers are OVERWRITTEN (TABLE kinds: only the
                                                 processed bin range is written) */
                    void* grad_x, int64_t gx_batch_stride, /* device or NULL */
                    int64_t batch, int64_t cols, int64_t bin_begin, int64_t n_bins,
                    int epilogue, void* workspace, size_t workspace_bytes, void* stream);

/* ---- criteria fused into the sweep -----------------------------------------------------------------
 * The |.| output layer (Transform(torch.abs), every BASELINE config) and an MSE criterion are evaluated
 * inside the sweep kernel: neither |Y| nor dL/d|Y| touches HBM, and because the backward kernel
 * recomputes the forward states anyway, ONE launch of fsweep_backward_loss yields the loss AND its
 * gradients (upstream gradient 1; the caller scales them by dL/dloss).  Replaces, in one training step
 * (optimize/trainer.py:177-190): the output layer, the criterion forward (optimize/loss.py:90-103 or
 * torch.nn.MSELoss), its autograd backward and the forward sweep.  Only for programs with cols == 1. */
#define FSWEEP_CRIT_MSE 1       /* nn.MSELoss()(|Y|, t): e = |Y[b,k,r]| - t[b,k,r]   (examples/e7_biquad.py:67) */
#define FSWEEP_CRIT_MSE_CHSUM 2 /* mse_loss (optimize/loss.py:90-103): e = sum_r |Y[b,k,r]| - t[b,k]           */

typedef struct fsweep_criterion {
  int32_t kind;
  int32_t reserved;
  const void* target;          /* device real: (batch, n_bins, n_out) [MSE] or (batch, n_bins) [MSE_CHSUM];
                                  addresses the FIRST PROCESSED bin like x */
  int64_t target_batch_stride; /* in elements */
  double scale;                /* loss = scale * sum e^2 over the processed range; for the reference's mean the
                                  caller passes 1 / (number of elements of the FULL target) */
  void* loss;                  /* device real[1], overwritten */
} fsweep_criterion_t;

/* loss only (validation) */
FSWEEP_API int fsweep_forward_loss(const fsweep_plan_t* plan, const void* const* coeffs, const void* x,
                        int64_t x_batch_stride, const fsweep_criterion_t* crit, int64_t batch,
                        int64_t bin_begin, int64_t n_bins, void* workspace, size_t workspace_bytes, void* stream);
/* loss and d loss / d coeffs (and d loss / d x when grad_x != NULL); buffers as in fsweep_backward */
FSWEEP_API int fsweep_backward_loss(const fsweep_plan_t* plan, const void* const* coeffs, const void* x,
                         int64_t x_batch_stride, const fsweep_criterion_t* crit, void* const* grad_coeffs,
                         void* grad_x, int64_t gx_batch_stride, int64_t batch, int64_t bin_begin, int64_t n_bins,
                         void* workspace, size_t workspace_bytes, void* stream);

/* E = exp(S) for the orthogonal map of dsp.Matrix (reference dsp.py:649, functional.py:42-56):
 * skew != 0: S = triu(P,1) - triu(P,1)^T, else S = P.  P, E, G, gP: device real[n][n] row-major, real = float
 * for dtype FSWEEP_C64 and double for FSWEEP_C128 (the parameter's own dtype: no cast kernels around the call);
 * the arithmetic is float64 either way.
 * backward: gP = dL/dP given G = dL/dE.  One CTA, entirely on the device (torch.matrix_exp reads its norm
 * back to the host, which cannot be captured in a CUDA graph).
 * n <= 2*fsweep_expm_max_n() for forward, n <= fsweep_expm_max_n() for backward. */
FSWEEP_API int fsweep_expm_max_n(void);
FSWEEP_API int fsweep_expm_forward(const void* P, void* E, int n, int skew, int dtype, void* stream);
FSWEEP_API int fsweep_expm_backward(const void* P, const void* G, void* gP, int n, int skew, int dtype, void* stream);
/* The same with the sparsity_loss of the result (optimize/loss.py:36-63, below) riding along: forward also writes
 * sparsity = (sum |E| - n sqrt n) / (n (1 - sqrt n)) (device real[1], or NULL); backward takes dL/dE in G (or NULL) and
 * dL/dsparsity in gsparsity (device real[1], or NULL; E = the forward result) and returns the gradient of both through
 * the map — the parameter-sized part of a colorless-FDN training step is then ONE launch each way. */
FSWEEP_API int fsweep_expm_forward_sp(const void* P, void* E, int n, int skew, int dtype, void* sparsity, void* stream);
FSWEEP_API int fsweep_expm_backward_sp(const void* P, const void* G, void* gP, int n, int skew, int dtype, const void* E,
                                       const void* gsparsity, void* stream);

/* dsp.Biquad / parallelBiquad, low-pass or high-pass prototype (reference dsp.py:1494-1563, functional.py:376-470): raw
 * parameter (K, 2, n_out, n_in) [parallel: (K, 2, n), n_out = n_in = n] -> bounded map -> RBJ taps -> packed Taylor blocks
 * of FSWEEP_OP_SOS (layout above), float64, in ONE launch.  Forward: packed != NULL (double[K][n_in][n_out][2][8]).
 * Adjoint: packed == NULL, gpacked = dL/dpacked (double, same layout), gparam = dL/dparam (the parameter's dtype: float
 * for FSWEEP_C64, double for FSWEEP_C128). */
FSWEEP_API int fsweep_biquad_design(const void* param, int K, int n_out, int n_in, int parallel, int highpass, int dtype,
                                    void* packed, const void* gpacked, void* gparam, void* stream);

/* dsp.SVF / parallelSVF with the general mixing (filter_type = None; reference dsp.py:2214-2232, 2234-2347): raw parameter
 * (5, K, n_out, n_in) [parallel: (5, K, n)] -> activations -> taps -> packed Taylor blocks, float64; forward when
 * packed != NULL, adjoint when packed == NULL (arguments as fsweep_biquad_design). */
FSWEEP_API int fsweep_svf_design(const void* param, int K, int n_out, int n_in, int parallel, int dtype, void* packed,
                                 const void* gpacked, void* gparam, void* stream);

/* sparsity_loss of the mapped feedback matrix (reference optimize/loss.py:36-63), A: device real[n_mats][n][n]:
 *   loss = mean_i ((sum |A_i| - n sqrt n) / (n (1 - sqrt n)));  backward: gA = gloss * dloss/dA (gloss: device real[1]).
 * One launch each way, capture safe. */
FSWEEP_API int fsweep_sparsity_forward(const void* A, int n_mats, int n, int dtype, void* loss, void* stream);
FSWEEP_API int fsweep_sparsity_backward(const void* A, const void* gloss, int n_mats, int n, int dtype, void* gA,
                                        void* stream);

/* Weighted total of the step's criteria (reference optimize/trainer.py:184-188): parts[i] are device real[1];
 * vals (device real[n+1]) receives vals[i] = scales[i] * part_i (i < n) followed by sum_i alphas[i] * vals[i].
 * One launch. */
#define FSWEEP_MAX_CRITERIA 8
FSWEEP_API int fsweep_weighted_total(const void* const* parts, const double* alphas, const double* scales, int n,
                                     int dtype, void* vals, void* stream);
/* The same, and the n + 1 values are ALSO stored into host_vals — mapped pinned host memory (cudaHostAlloc /
 * cudaHostRegister; with unified addressing the host pointer is the device pointer), real[n+1] — followed, behind a
 * system-wide fence, by the number of this launch (1, 2, ...; seq_counter: device int32[1], zero at the start) into
 * host_seq (mapped pinned int32[1]).  A host thread that polls host_seq has the step's losses (reference
 * optimize/trainer.py:190 `loss.item()`) as soon as they exist: no copy node in the captured step, no wait for the
 * kernels behind the criteria.  host_vals == NULL: plain fsweep_weighted_total.
 * Layout of host_vals: FSWEEP_C128 - double[n+1], published by a system-wide fence before host_seq is written;
 * FSWEEP_C64 - (n+1) pairs {float value; int32 launch number}, each written with ONE aligned 8-byte store and no
 * fence: the host waits until host_seq AND every pair carry the number it expects. */
FSWEEP_API int fsweep_weighted_total_notify(const void* const* parts, const double* alphas, const double* scales, int n,
                                            int dtype, void* vals, void* host_vals, void* host_seq, void* seq_counter,
                                            void* stream);

/* torch.optim.Adam's update (the Trainer's optimizer, reference optimize/trainer.py:42; no weight decay, no amsgrad)
 * for up to FSWEEP_ADAM_MAX_TENSORS parameter tensors in ONE launch.  All pointers are device pointers; param / grad /
 * exp_avg / exp_avg_sq are real[numel] (float for FSWEEP_C64, double for FSWEEP_C128), step is a float32[1] counter
 * private to the tensor (incremented by the call), lr a float32[1].  Capture safe. */
#define FSWEEP_ADAM_MAX_TENSORS 32
typedef struct fsweep_adam_tensor {
  void* param;
  const void* grad;
  void* exp_avg;
  void* exp_avg_sq;
  void* step;
  int64_t numel;
} fsweep_adam_tensor_t;
FSWEEP_API int fsweep_adam_step(const fsweep_adam_tensor_t* tensors, int n, int dtype, const void* lr, double beta1,
                                double beta2, double eps, void* stream);
/* The same launch with a rider: one extra block evaluates the weighted total of the step's criteria (and notifies the
 * host, see fsweep_weighted_total_notify) - the values are only read by the host, so inside a captured step they need
 * no launch of their own.  total == NULL: plain fsweep_adam_step. */
typedef struct fsweep_total_job {
  const void* parts[FSWEEP_MAX_CRITERIA];
  double alphas[FSWEEP_MAX_CRITERIA];
  double scales[FSWEEP_MAX_CRITERIA];
  int32_t n;
  int32_t reserved;
  void* vals;        /* device real[n+1] */
  void* host_vals;   /* optional, with host_seq and seq_counter */
  void* host_seq;
  void* seq_counter;
} fsweep_total_job_t;
/* fsweep_expm_backward_sp with the same rider (it runs BEFORE the optimizer: the host has the losses while the adjoint of
 * the map and the optimizer are still running). */
FSWEEP_API int fsweep_expm_backward_sp_total(const void* P, const void* G, void* gP, int n, int skew, int dtype,
                                             const void* E, const void* gsparsity, const fsweep_total_job_t* total,
                                             void* stream);
FSWEEP_API int fsweep_adam_step_total(const fsweep_adam_tensor_t* tensors, int n, int dtype, const void* lr, double beta1,
                                      double beta2, double eps, const fsweep_total_job_t* total, void* stream);

/* One-shot all-reduce (sum, then * scale) of a small float32 buffer that lives in symmetric / peer-mapped memory on
 * every rank of one node — the single exchange of a multi-GPU training step (flamo has no multi-device path; this
 * replaces the NCCL all-reduce of flamo_b200/parallel.py).  peer_buffers / peer_signal_pads: DEVICE arrays of `world`
 * device pointers (rank r's buffer / signal pad, as torch.distributed._symmetric_memory hands them out; the pads must
 * be zero before the first call and at least 2 KiB).  epoch_counter: device uint32[2], zero before the first call, private
 * to this communicator: [0] the epoch, [1] a sticky error flag (1 + rank of a peer that did not arrive within the
 * bounded spin; the results of that call are then undefined and the caller must not use them).  One kernel, in place, capture safe, bit-identical results on every rank.
 * n <= fsweep_allreduce_p2p_max_n(). */
FSWEEP_API int fsweep_allreduce_p2p_max_n(void);
FSWEEP_API int fsweep_allreduce_p2p(void* const* peer_buffers, void* const* peer_signal_pads, int rank, int world, int n,
                                    double scale, void* epoch_counter, void* stream);

/* The same exchange as ONE push kernel that also packs and unpacks: `segs` are the step's float32 gradient tensors (and the
 * loss values), reduced IN PLACE.  Every rank's receive area (peer_buffers[r], symmetric memory, 8-byte aligned, zeroed
 * once) holds 2 * world * cap 8-BYTE slots (double buffered on the epoch parity); cap >= the total number of values.
 * Gathers the segments, stores each value together with the step's epoch ({value, epoch}, one 8-byte store) into slot
 * [rank] of every peer's area over NVLink, polls its own area until every slot carries the epoch, sums the slots in
 * rank order (bit-identical on every rank), scales and scatters back: no fence, no flag round.  epoch_counter as
 * above; the signal pads are not used any more (kept in the signature).  Capture safe. */
#define FSWEEP_AR_MAX_SEGS 32
typedef struct fsweep_seg {
  void* ptr;     /* device float32 */
  int64_t numel;
} fsweep_seg_t;
FSWEEP_API int fsweep_allreduce_push(const fsweep_seg_t* segs, int n_segs, void* const* peer_buffers,
                                     void* const* peer_signal_pads, int rank, int world, int cap, double scale,
                                     void* epoch_counter, void* stream);
/* The same, and the LAST segment (the step's loss values) is also stored into mapped pinned host memory in
 * fsweep_weighted_total_notify's float32 layout ({value, launch number} pairs, host_seq, seq_counter): the host has the
 * exchanged losses while the optimizer is still running.  host_vals == NULL: plain fsweep_allreduce_push. */
FSWEEP_API int fsweep_allreduce_push_notify(const fsweep_seg_t* segs, int n_segs, void* const* peer_buffers,
                                            void* const* peer_signal_pads, int rank, int world, int cap, double scale,
                                            void* epoch_counter, void* host_vals, void* host_seq, void* seq_counter,
                                            void* stream);

/* Real-to-complex FFT of the excitation along the time axis (reference flamo/processor/dsp.py:69-93 dsp.FFT and
 * :122-163 dsp.FFTAntiAlias: torch.fft.rfft(x [* envelope], n=nfft, dim=1)): x float32 [batch][n_time][channels]
 * (time stride = channels, batch stride given in elements), zero padded / cropped to nfft, X complex64
 * [batch][nfft/2 + 1][channels] = scale * DFT.  Two launches of a four-step FFT (fsweep_fft.cu) instead of cuFFT's five.
 * `table`: fsweep_rfft_table_entries(nfft) complex64 twiddles, filled once per (nfft, device) by fsweep_rfft_table
 * (float64 inside, laid out in the order the kernels' threads read them); `workspace`: fsweep_rfft_workspace_bytes(nfft, batch * channels); `envelope`: NULL or nfft float32 factors.
 * fsweep_rfft_supported(nfft) == 0 (and FSWEEP_E_UNSUPPORTED from fsweep_rfft): nfft is odd, < 512, or nfft / 2 does
 * not split into two factors <= 1024 made of radices <= 64 - the caller keeps cuFFT for those.  Capture safe. */
FSWEEP_API int fsweep_rfft_supported(int64_t nfft);
FSWEEP_API size_t fsweep_rfft_workspace_bytes(int64_t nfft, int64_t signals);
FSWEEP_API int64_t fsweep_rfft_table_entries(int64_t nfft);
FSWEEP_API int fsweep_rfft_table(void* table, int64_t nfft, void* stream);
FSWEEP_API int fsweep_rfft(const void* x, int64_t batch, int64_t n_time, int64_t channels, int64_t x_batch_stride,
                           int64_t nfft, double scale, const void* envelope, const void* table, void* workspace,
                           size_t workspace_bytes, void* X, void* stream);

/* Upload of a step's batch (reference optimize/trainer.py:176 move_to_device: inputs and targets) from PINNED host memory
 * (cudaHostAlloc / cudaHostRegister: with unified addressing the host pointer is a device pointer) into device buffers as
 * ONE kernel that reads the host tensors over PCIe, instead of one DMA copy per tensor.  n <= FSWEEP_UPLOAD_MAX segments
 * of bytes[i] bytes; the host must not rewrite a source before the stream has passed the call (as with
 * cudaMemcpyAsync). */
#define FSWEEP_UPLOAD_MAX 4
FSWEEP_API int fsweep_upload(const void* const* host_src, void* const* dev_dst, const int64_t* bytes, int n, void* stream);

/* FP32 FMA peak probe (bench.py's roofline denominator for the compute-bound sweeps; SURVEY.md section 8d "derive +
 * measure"): `blocks` blocks of 256 threads, 64 independent FFMAs per thread and round; fsweep_fma_probe_flops gives the
 * flop count of one launch, the caller times it with CUDA events.  out: device float[1] (never written in practice). */
FSWEEP_API int fsweep_fma_probe(void* out, int blocks, int iters, void* stream);
FSWEEP_API double fsweep_fma_probe_flops(int blocks, int iters);

/* number of kernels the last forward / backward call of this thread enqueued (bench bookkeeping) */
FSWEEP_API int fsweep_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FSWEEP_H */
