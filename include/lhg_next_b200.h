/*
 * lhg_next_b200.h -- C ABI of the stages either side of the propagation path.
 *
 * Same library (libasm_b200.so) and same conventions as asm_b200.h: extern "C", plain pointers and sizes,
 * device pointers owned by the caller (fp32, contiguous [plane][row][col]), asynchronous on the stream passed
 * in, 0 on success / negative lhg_status otherwise, lhg_next_last_error() for the thread-local message.
 * Reductions are two-stage with a fixed order (per-block partials, one finishing block): no float atomics,
 * bit-identical from run to run.  Every entry point names the reference lines it replaces:
 *   "loss.py:L" = learnedMethodForHologram/watermelon_hologram/loss_func.py
 *   "util.py:L" = learnedMethodForHologram/utilities.py
 *   "ap2poh.py:L" = learnedMethodForHologram/watermelon_hologram/AP2POH.py
 *   "nn.py:L"   = learnedMethodForHologram/neural_network_components.py
 *   "dl.py:L"   = learnedMethodForHologram/watermelon_hologram/data_loader.py
 */
#ifndef LHG_NEXT_B200_H
#define LHG_NEXT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LHG_NEXT_VERSION 105

typedef void* lhg_stream; /* cudaStream_t */

enum lhg_status { LHG_OK = 0, LHG_EINVAL = -1, LHG_ECUDA = -3, LHG_EWORKSPACE = -4 };

int lhg_next_version(void);
const char* lhg_next_last_error(void);
/* kernels launched by the entry points of this header since the library was loaded */
long long lhg_next_launch_count(void);

/* Scratch (floats) the reductions below need for `planes` planes of rows x cols: pass at least this much. */
size_t lhg_next_partial_floats(long long planes, int rows, int cols);

/* ---- N1: loss epilogues (loss.py:66-104, watermelon.py:418-445) ----------------------------------------
 * lhg_amp_loss_terms: ONE pass over hat (and target) produces
 *   terms[0] = F.mse_loss(hat, target)                                   loss.py:101   (NaN slot if target == NULL)
 *   terms[1] = total_variation(hat)    = mean|dx hat| + mean|dy hat|     loss.py:66-77
 *   terms[2] = total_variation(target)                                   loss.py:96
 *   terms[3] = |terms[1] - terms[2]|   = total_variation_loss            loss.py:92-96
 *   terms[4] = terms[0] + alpha*terms[3] = amp_loss                      loss.py:99-103
 * hat/target: f32 [planes,rows,cols]; terms: device f32 [5]; partial: device scratch of partial_floats floats. */
int lhg_amp_loss_terms(const float* hat, const float* target, long long planes, int rows, int cols, float alpha,
                       float* partial, size_t partial_floats, float* terms, lhg_stream stream);

/* Gradient with respect to hat of  sum_i g[i]*terms[i]  (g: device f32 [5], the upstream cotangents of the five
 * terms; terms: the device f32 [5] lhg_amp_loss_terms wrote for the same inputs, read for the sign of
 * TV(hat)-TV(target); abs' = sgn as in torch).  grad_hat f32 [planes,rows,cols].  With target == NULL only
 * g[1] (total_variation) is used and terms may be NULL. */
int lhg_amp_loss_backward(const float* hat, const float* target, const float* g, const float* terms, float alpha,
                          long long planes, int rows, int cols, float* grad_hat, lhg_stream stream);

/* focal_sincos_phase_gradient_loss (loss.py:135-163): with S = cat(sin, cos) of each phase and
 * d1 = |dx S_fake - dx S_real|, d2 = |dy ...|,  loss = mean(d1*(d1/max d1)) + mean(d2*(d2/max d2)).
 * The weights are constants of the graph, so one pass gives sum d^2 and max d together:
 *   terms[0] = max d1, terms[1] = max d2, terms[2] = loss,
 *   terms[3] = mean d1 + mean d2 = phase_sincos_gradient_loss (loss.py:165-183, the un-weighted variant of
 *   watermelon.py:921) from the same pass.   terms: device f32 [4].  Phases f32 [planes,rows,cols], any range. */
int lhg_focal_phase_loss_terms(const float* fake_phase, const float* real_phase, long long planes, int rows,
                               int cols, float* partial, size_t partial_floats, float* terms, lhg_stream stream);

/* d loss / d fake_phase * g[0] (g: device f32 [1]); terms from lhg_focal_phase_loss_terms of the same inputs. */
int lhg_focal_phase_loss_backward(const float* fake_phase, const float* real_phase, const float* terms,
                                  const float* g, long long planes, int rows, int cols, float* grad_fake,
                                  lhg_stream stream);

/* d phase_sincos_gradient_loss / d fake_phase * g[0] (g: device f32 [1]): the same stencil with sgn differences. */
int lhg_phase_gradient_loss_backward(const float* fake_phase, const float* real_phase, const float* g,
                                     long long planes, int rows, int cols, float* grad_fake, lhg_stream stream);

/* Point-wise phase losses (loss_func.py:186-208), ONE pass for both:
 *   terms[0] = max d over both channels of d = |sin f - sin r|, |cos f - cos r|
 *   terms[1] = focal_sincos_phase_loss = mean(d * d / max d)        (loss_func.py:186-203)
 *   terms[2] = plain_phase_loss        = mean |f - r|               (loss_func.py:206-208)
 * partial as for the other losses (lhg_next_partial_floats). */
int lhg_phase_point_loss_terms(const float* fake_phase, const float* real_phase, long long planes, int rows, int cols,
                               float* partial, size_t partial_floats, float* terms, lhg_stream stream);
/* d loss / d fake_phase * g[0]; focal != 0: focal_sincos_phase_loss (needs terms), 0: plain_phase_loss. */
int lhg_phase_point_loss_backward(const float* fake_phase, const float* real_phase, const float* terms, const float* g,
                                  int focal, long long planes, int rows, int cols, float* grad_fake, lhg_stream stream);

/* ---- N4: focal-stack export (util.py:69-84 tensor_normalizor_2D, util.py:179-203 -> plt.imsave) ---------
 * minmax: device f32 [planes,2] = (min, max) over rows x cols of every plane (NaN propagates as in torch). */
int lhg_plane_minmax(const float* x, long long planes, long long plane_elems, float* partial,
                     size_t partial_floats, float* minmax, lhg_stream stream);
/* out = (x - min) / (max - min) per plane, IEEE fp32 ops in the reference's order (util.py:83). */
int lhg_normalize_planes(const float* x, const float* minmax, long long planes, long long plane_elems,
                         float* out, lhg_stream stream);
/* out = x / (max * 1.01) per plane: amplitude_normalizor (util.py:53-66), same fp32 operations. */
int lhg_amplitude_normalize(const float* x, const float* minmax, long long planes, long long plane_elems,
                            float* out, lhg_stream stream);
/* 8-bit image writer: x f32 [images,3,rows,cols] -> out u8 [images,rows,cols,out_channels], out_channels 3 (RGB)
 * or 4 (RGBA, alpha 255: what plt.imsave stores); value = (uint8)(fl32(normalised * 255)), truncation, the
 * float-RGB branch of matplotlib's to_rgba(bytes=True).  minmax == NULL packs x itself (already in [0,1]). */
int lhg_pack_rgb_u8(const float* x, const float* minmax, long long images, int rows, int cols, int out_channels,
                    uint8_t* out, lhg_stream stream);

/* ---- N2: AP2POH tail (ap2poh.py:86-116) --------------------------------------------------------------
 * field: complex64 [planes,rows,cols] (plane = sample*3 + colour), the output of propagate_AP2C_backward.
 * m = ChannelWiseSymmetricConv(re) + i*ChannelWiseSymmetricConv(im)  (nn.py:35-95: one k x k kernel and one
 * bias per colour, zero padding (k-1)/2; weights f32 [3,k,k], bias f32 [3], device);
 * a = |m| / (1.01 * max over the plane of |m|)  (util.py:53-66);  p = angle(m);
 * POH = p + acos(a) where (row+col) is even, p - acos(a) where it is odd (checkerboard cell 1, ap2poh.py:35-47,86-95).
 * plane_max: device f32 [planes] (written, then read by the second kernel). */
int lhg_ap2poh_tail(const void* field, const float* weights, const float* bias, int ksize, long long planes,
                    int rows, int cols, float* partial, size_t partial_floats, float* plane_max, float* poh,
                    lhg_stream stream);

/* Backward of lhg_ap2poh_tail (the training step differentiates AP2POH.forward, ap2poh.py:104-116): g_poh = dL/dPOH
 * f32 [planes,rows,cols] -> grad_field complex64 [planes,rows,cols] (dL/dre + i dL/dim, torch's convention for the
 * real/imag split of ap2poh.py:108-110), grad_weights f32 [3,k,k], grad_bias f32 [3].  The gradient of the per-plane
 * maximum goes to the arg-max pixel, as torch.max does.  scratch: device f32, at least
 * lhg_ap2poh_tail_backward_floats(...) floats, 8-byte aligned. */
size_t lhg_ap2poh_tail_backward_floats(int ksize, long long planes, int rows, int cols);
int lhg_ap2poh_tail_backward(const void* field, const float* weights, const float* bias, int ksize,
                             const float* g_poh, long long planes, int rows, int cols, float* scratch,
                             size_t scratch_floats, void* grad_field, float* grad_weights, float* grad_bias,
                             lhg_stream stream);

/* ---- N3: raw .bin dataset reader (dl.py:8-123) --------------------------------------------------------
 * Host side: copy the first copy_bytes of items idx[0..n) of a memory-mapped [N, item_bytes] file into a
 * (pinned) staging buffer, dst[i] = base[idx[i]], with `threads` host threads (0 = pick).  Returns LHG_EINVAL
 * for an index outside [0, n_items). */
int lhg_bin_gather(const void* base, long long n_items, size_t item_bytes, size_t copy_bytes,
                   const int64_t* idx, int n, void* dst, int threads);
/* Device side: out[i] = cat(img[i] (3 planes), depth[i] plane 0) -> f32 [n,4,plane_elems]   dl.py:43-51;
 * depth holds depth_planes planes per item (1 when only the first plane was uploaded). */
int lhg_assemble_rgbd(const float* img, const float* depth, int depth_planes, long long n, long long plane_elems,
                      float* out, lhg_stream stream);
/* out = fl32(2*pi) * x  (dl.py:85: "2 * torch.pi * phs") */
int lhg_scale_two_pi(const float* x, long long n, float* out, lhg_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* LHG_NEXT_B200_H */
