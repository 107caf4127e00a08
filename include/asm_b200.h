/*
 * asm_b200.h -- C ABI of the B200-native band-limited angular-spectrum propagation path.
 *
 * The reference (WeijieXie/learned_hologram_gan) has no FFI of its own: the boundary it
 * exposes is the Python surface of learnedMethodForHologram/angular_spectrum_method.py,
 * whose arithmetic is torch.fft + ATen.  This header is the boundary a replacement
 * library exports; every entry point names the reference lines it replaces
 * ("asm.py:L" = learnedMethodForHologram/angular_spectrum_method.py, "util.py:L" =
 * learnedMethodForHologram/utilities.py).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - Every tensor and the scratch buffer are owned by the caller (PyTorch's caching
 *     allocator); the library owns only the opaque plan (twiddle / permutation tables).
 *   - All tensor pointers are DEVICE pointers on the plan's device, contiguous, fp32 or
 *     interleaved complex64 (re,im).  Layout is [plane][row][col]; a "plane" is one
 *     (sample, colour) image, colour fastest:  plane = sample * n_colour + colour.
 *   - Calls are asynchronous on the stream passed in; the library never synchronises.
 *   - Return 0 on success, a negative asm_status otherwise; asm_last_error() returns a
 *     thread-local message.  Nothing throws, nothing exits.
 *   - Plans are immutable after creation; entry points are re-entrant and set the CUDA
 *     device themselves (autograd calls backward from another host thread).
 */
#ifndef ASM_B200_H
#define ASM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASM_B200_VERSION 103

typedef struct asm_plan asm_plan;
typedef void* asm_stream; /* cudaStream_t */

enum asm_status {
  ASM_OK = 0,
  ASM_EINVAL = -1,           /* bad argument / inconsistent descriptor */
  ASM_EUNSUPPORTED_SIZE = -2,/* transform length cannot be planned */
  ASM_ECUDA = -3,            /* CUDA runtime error (message has the detail) */
  ASM_EWORKSPACE = -4        /* scratch buffer too small */
};

/* ---- input stage (what the row-forward pass reads) ------------------------------ */
enum asm_in_kind {
  ASM_IN_PHASE = 0,     /* in1 = phase f32 [P,R,C];              x = exp(i*s*phase)      asm.py:136,390 */
  ASM_IN_AMP_PHASE = 1, /* in0 = amp, in1 = phase f32 [P,R,C];   x = a*exp(i*s*phase)    asm.py:88,382,550 */
  ASM_IN_COMPLEX = 2,   /* in0 = complex64 [P,R,C] (also: cotangent of a complex output) */
  ASM_IN_SPECTRUM = 3,  /* in0 = complex64 [P,Rp,Cp], natural order, DC at [0,0]         asm.py:527,539 */
  ASM_IN_COTANGENT = 4  /* adjoint prologue: in0 = saved complex field y [P,R,C];
                           ybar = cot_abs*y/|y| + cot_angle*(i*y)/|y|^2 + 2*cot_abs2*y
                                + cot_scale*(|y|-cot_target)*y/|y|      (each term optional) */
};

/* ---- spectral filter applied between the column transforms ------------------------ */
enum asm_filter_kind {
  ASM_FILTER_NONE = 0, /* identity (mask flag may still apply)                           asm.py:551 */
  ASM_FILTER_H = 1     /* H = exp(-2*pi*i*z*w), generated on the fly, fp32-faithful      asm.py:206-211 */
};
enum asm_filter_flags {
  ASM_FILTER_CONJ = 1,      /* use conj(H): the adjoint, and the reference's "/ H"       asm.py:366,383 */
  ASM_FILTER_CIRC_MASK = 2  /* multiply by the circular low-pass of the plan             asm.py:91,333 */
};

/* ---- output stage (what the row-inverse pass, or the column pass, writes) ---------- */
enum asm_out_kind {
  ASM_OUT_ABS = 0,        /* out0 = |y| f32 [P,R,C]                                       asm.py:92,522 */
  ASM_OUT_ANGLE = 1,      /* out0 = angle(y)                                                             */
  ASM_OUT_ABS_ANGLE = 2,  /* out0 = |y|, out1 = angle(y)                                  asm.py:424,531 */
  ASM_OUT_COMPLEX = 3,    /* out0 = y complex64 [P,R,C]                                   asm.py:383 */
  ASM_OUT_ABS2 = 4,       /* out0 = |y|^2                                                 asm.py:139 */
  ASM_OUT_SPECTRUM = 5,   /* out0 = complex64 [P,Rp,Cp] natural order (no inverse passes) asm.py:391,551 */
  ASM_OUT_GRAD_PHASE = 6  /* adjoint epilogue: with x = a*exp(i*s*phase) (aux_amp may be NULL = 1)
                             out0 = d/dphase = s*a*Im(conj(e)*xbar), out1 (optional) = d/da = Re(conj(e)*xbar) */
};

/*
 * One call = one pass of the hot path over n_samples * n_colour input planes.
 *
 *   reduce_depth == 0:  each input plane (s,c) is propagated to n_depth output planes,
 *                       out plane = (s*n_depth + d)*n_colour + c        (asm.py:516-518)
 *   reduce_depth == 1:  adjoint shape: n_depth input planes (s,d,c), filtered and SUMMED
 *                       over d inside the column pass into one output plane (s,c).
 *   The depth of (s,d) is z_dev[depth_index ? depth_index[s*n_depth+d] : d]  (asm.py:536-537).
 */
typedef struct asm_io {
  int32_t struct_bytes;   /* = sizeof(asm_io), checked */
  int32_t n_samples;
  int32_t n_depth;
  int32_t reduce_depth;

  int32_t in_kind;
  int32_t filter_kind;
  int32_t filter_flags;
  int32_t out_kind;

  const void* in0;
  const void* in1;
  const float* cot_abs;    /* ASM_IN_COTANGENT terms, f32 [P,R,C] or NULL */
  const float* cot_angle;
  const float* cot_abs2;
  const float* cot_target;
  float cot_scale;
  float phase_scale;       /* s above; 1.0f normally, fl32(2*pi) for asm.py:549 */

  const float* wm_grid;    /* f32 [n_colour,Rp,Cp]: |value| = w_grid (asm.py:155-171), sign bit set = bin is
                              outside the circular mask (util.py:234-241).  The caller builds it with the
                              reference's own host ops, because torch's CPU sqrt (MKL VML) is not correctly
                              rounded and a 1-ulp change of w moves H by up to 1.6e-3 rad.  NULL = the
                              kernels generate w and the mask themselves with IEEE-rounded fp32 arithmetic. */
  const float* z_dev;      /* f32 [n_z] propagation distances (device) */
  const int32_t* depth_index; /* device, [n_samples*n_depth] or NULL */
  int32_t n_z;
  int32_t reserved0;

  void* out0;
  void* out1;
  void* save_field;        /* optional complex64 [P_out,R,C]: y before abs/angle (for backward) */
  const float* aux_phase;  /* ASM_OUT_GRAD_PHASE: the forward's phase / amplitude */
  const float* aux_amp;
  float out_scale;         /* applied once at the last stage, e.g. 1/(Rp*Cp)  (ifft2 norm, asm.py:92) */
  int32_t loss_partial_len;
  const float* loss_target;/* optional fused amplitude-L2: sum((|y|-target)^2) block partials */
  float* loss_partial;     /* f32 [loss_partial_len], fully overwritten; fixed summation order */

  void* workspace;         /* >= asm_workspace_bytes(...) */
  size_t workspace_bytes;

  const void* wm_tiled;    /* optional: wm_grid re-ordered for the tiled column kernel by asm_build_wm_tiled
                              (asm_wm_tiled_bytes(plan) bytes).  NULL = the run-time-planned kernels are used
                              whenever the call needs w or the mask. */

  /* Fused training step (optional, asm_fused_step_supported(plan) == 1): when adj_grad_phase is non-NULL the
   * call is the forward pass described above with out_kind = ASM_OUT_ABS, loss_target and loss_partial
   * (out0 may then be NULL: |y| is not written), FOLLOWED BY its adjoint: the cotangent
   * adj_cot_scale * (|y| - loss_target) * y/|y| of every output plane goes back through the same pipeline with
   * conj(filter), summed over depth, into adj_grad_phase = d/dphase (f32, shaped like in1; d/damp is not
   * produced).  It equals the two calls  {ABS + loss + save_field}  and  {ASM_IN_COTANGENT(cot_target,
   * cot_scale) -> ASM_OUT_GRAD_PHASE, reduce_depth, CONJ toggled}  bit for bit, but the row-inverse pass of the
   * first and the row-forward pass of the second are one kernel: |y|, the saved field and their 24 bytes of
   * traffic per output sample never exist. */
  float* adj_grad_phase;
  float adj_cot_scale;
  int32_t loss_target_u8;  /* fused step only: 1 = loss_target points at uint8 samples v [P,R,C] and the target amplitude
                              is fl(v / 255) (the reference's 8-bit image convention, util.py:44: `.div(255)`), converted
                              as the row kernel reads it: a quarter of the bytes over the host link and out of HBM */
} asm_io;

/* ---- grids the reference keeps as attributes ---------------------------------------- */
enum asm_grid_kind {
  ASM_GRID_W = 0,          /* f32 [n_colour,Rp,Cp]   w_grid                       asm.py:155-171 */
  ASM_GRID_CIRC_MASK = 1,  /* f32 [Rp,Cp] {0,1}      diffraction_limited_mask     util.py:206-243 */
  ASM_GRID_RADIAL = 2,     /* f32 [Rp,Cp]            sqrt(u^2+v^2)*min(Rp,Cp)     util.py:276-296 */
  ASM_GRID_H = 3,          /* c64 [D,n_colour,Rp,Cp] transfer function            asm.py:195-213 */
  ASM_GRID_BAND_LIMIT = 4  /* u8  [D,n_colour,Rp,Cp] band-limit mask              asm.py:173-193 */
};

int asm_version(void);
/* sizeof(asm_io) as compiled into the library, for binding generators to check their struct layout */
int asm_sizeof_io(void);
const char* asm_last_error(void);

/*
 * Geometry of one propagator (replaces asm.py:30-63).  rows/cols: un-padded hologram;
 * pad_rows/pad_cols: zero-pad on each side; pitch in metres (double, as the reference's
 * Python float); wavelengths: host fp32 array; mask_radius = min(Rp,Cp)*coefficient
 * evaluated by the caller in double (asm.py:152), must be <= min(Rp,Cp)/2 (util.py:225-229).
 */
int asm_plan_create(asm_plan** out, int device, int rows, int cols, int pad_rows, int pad_cols,
                    double pitch, const float* wavelengths, int n_colour, double mask_radius);
int asm_plan_destroy(asm_plan* plan);
/* sizes: out[0]=Rp, out[1]=Cp, out[2]=1 if both lengths use the mixed-radix path, 0 if Bluestein */
int asm_plan_info(const asm_plan* plan, int32_t* out, int n);

/* Scratch bytes needed by asm_propagate for this descriptor (only sizes/kinds are read). */
size_t asm_workspace_bytes(const asm_plan* plan, const asm_io* io);

/* Builders for the attributes.  ASM_GRID_H evaluates exp(-2 pi i z w) from wm_grid (see asm_io; NULL =
 * device-generated w); the other kinds are the device-generated (IEEE-rounded) grids. z_dev: f32 [n_depth]. */
int asm_build_grid(const asm_plan* plan, int grid_kind, const float* wm_grid, const float* z_dev,
                   int n_depth, int filter_flags, void* out_dev, asm_stream stream);

/* The compile-time planned column kernel reads w and the mask in ITS tile order (column tile, scrambled
 * row position, column in tile) so that one tile is one contiguous run.  asm_wm_tiled_bytes returns the size
 * of that copy (0: this geometry has no such kernel); asm_build_wm_tiled fills it from wm_grid (NULL =
 * device-generated IEEE-rounded values), once per geometry; pass the result as asm_io.wm_tiled. */
size_t asm_wm_tiled_bytes(const asm_plan* plan);
int asm_build_wm_tiled(const asm_plan* plan, const float* wm_grid, void* out_dev, asm_stream stream);

/* The fused pipeline:  [prologue + row FFT] -> [column FFT * filter * column IFFT, depth loop]
 * -> [row IFFT + crop + epilogue].  Replaces asm.py:87-92 and every variant of it. */
int asm_propagate(const asm_plan* plan, const asm_io* io, asm_stream stream);

/* 1 when this geometry runs on the compile-time planned kernels, i.e. asm_io.adj_grad_phase is honoured;
 * 0 otherwise (asm_propagate then rejects a descriptor with adj_grad_phase set: issue the two calls). */
int asm_fused_step_supported(const asm_plan* plan);

/* Finishes the fused amplitude-L2 loss on the device: out_dev[0] = scale * sum(loss_partial[0..len)), summed by ONE
 * block in a fixed order with a double accumulator (no float atomics: bit-identical from run to run).  Replaces the
 * `sum()` of F.mse_loss over the partials; scale = 1/numel gives the mean (asm.py callers: F.mse_loss). */
int asm_loss_finish(const float* loss_partial, int len, double scale, float* out_dev, asm_stream stream);

/* ---- measurement hooks (bench.py) ----------------------------------------------------------
 * Kernels launched by this library since load (all plans, all threads). */
long long asm_launch_count(void);
/* When enabled, every kernel launch of asm_propagate is bracketed by CUDA events on the launching
 * stream.  asm_profile_collect waits for the recorded events, ADDS per-kernel device time (ms) and
 * launch counts into out_ms[k] / out_launches[k] (k = 0 row-forward, 1 column, 2 row-inverse,
 * 3 fused row-inverse + row-forward of the fused step; slots >= n are dropped),
 * and returns the events to a pool.  asm_profile_enable(1) fills that pool (2048 events) BEFORE the
 * region it measures, so that no event is created between two launches of a timed step.
 * Not thread-safe against concurrent asm_propagate calls. */
int asm_profile_enable(int on);
int asm_profile_collect(double* out_ms, long long* out_launches, int n);

#ifdef __cplusplus
}
#endif
#endif /* ASM_B200_H */
