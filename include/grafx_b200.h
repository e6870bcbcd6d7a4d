/* grafx_b200 -- C ABI of the B200 (sm_100a) batched audio-DSP hot path.
 *
 * The reference (sh-lee97/grafx @ 474e5dc) has no FFI of its own: its operator interface is the
 * Python processor protocol `nn.Module.forward(input_signals[N,C,L], **parameter_tensors)`
 * (sphinx-doc/source/introduction/processors.rst:170-192, call site render/graph.py:143-145).
 * Each entry point below replaces the O(samples) inner loop of one reference function; the
 * comment on each cites the reference file:line it stands in for.  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous memory unless the name ends in `_host`;
 *   - all functions are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy
 *     default stream), never allocate, never synchronise, never throw;
 *   - scratch memory is supplied by the caller: ask `*_workspace_bytes` first;
 *   - return value: GFX_OK (0) or a negative error code; after GFX_ERR_CUDA the failing
 *     cudaError_t is available from gfx_last_cuda_error();
 *   - input and output buffers must not overlap.
 */
#ifndef GRAFX_B200_H
#define GRAFX_B200_H

#include <stddef.h>

#if defined(__GNUC__)
#define GFX_API __attribute__((visibility("default")))
#else
#define GFX_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define GFX_OK 0
#define GFX_ERR_INVALID (-1)
#define GFX_ERR_WORKSPACE (-2)
#define GFX_ERR_CUDA (-3)
#define GFX_ERR_UNSUPPORTED (-4)

/* ---- library ---------------------------------------------------------------------------- */
GFX_API int gfx_version(void);              /* major*10000 + minor*100 + patch */
GFX_API int gfx_last_cuda_error(void);      /* cudaError_t of the last failed runtime call, 0 if none */
GFX_API const char* gfx_error_string(int code);
GFX_API unsigned long long gfx_kernel_launch_count(void); /* kernels this library has launched since it was loaded */
GFX_API int gfx_device_sm_count(void);      /* SMs of the current device (148 on B200), <0 on error */
/* Measurement aid (bench.py roofline: the fp32-FMA fraction beside the HBM one): enqueues a kernel of independent
 * FFMA chains on every SM and returns the number of FMAs it executes (negative: error code); time it with events. */
GFX_API long long gfx_fma_probe_f32(float* out, int iters, void* stream);

/* ---- exact biquad cascade ------------------------------------------------------------------
 * Replaces IIRFilter._process_lfilter (processors/core/iir.py:154-184: K sequential
 * torchaudio.functional.lfilter calls) and IIRFilter._process_ssm (iir.py:186-261).
 *   x  [batch, c_sig, L]          Bs, As [batch, c_filt, K, 3] (b0,b1,b2 / a0,a1,a2, not normalised)
 *   y  [batch, max(c_sig,c_filt), L]
 * Channel broadcasting follows iir.py:158-167: c_sig==c_filt, or either one is 1.
 * Each section is normalised by its own a0; zero initial state; no clamping. */
GFX_API size_t gfx_biquad_cascade_workspace_bytes(int batch, int c_sig, int c_filt, int K, int elem_size);
GFX_API int gfx_biquad_cascade_f32(const float* x, float* y, const float* Bs, const float* As, int batch,
                           int c_sig, int c_filt, int K, long long L, void* workspace,
                           size_t workspace_bytes, void* stream);
/* The cascade with the two options render_grafx's batched (4-D) sources call for (render/graph.py:63-75, 132-147;
 * render/core.py:6-33); with xcopy == NULL, src_outer == src_inner == 0 and coef_repeat == 1 it is gfx_biquad_cascade_f32.
 *  - coef_repeat > 1: runs of coef_repeat consecutive batch items share ONE coefficient row -- Bs, As are
 *    [batch / coef_repeat, c_filt, K, 3] (upstream expands every node's parameters over the batch of renders and
 *    designs / filters each copy; here the expansion never exists in memory).
 *  - xcopy != NULL (the FIRST render order: upstream copies the sources into the signal buffer, then the first processor
 *    reads that copy back): the signal is read from the caller's sources  x [src_outer, src_inner, c_sig, L]  (renders x
 *    source nodes, src_outer * src_inner == batch); item b of the launch (node-major like the signal buffer:
 *    b = node * src_outer + render) takes source item (b % src_outer) * src_inner + b / src_outer, and every input row is
 *    ALSO written to  xcopy [batch, c_sig, L]  -- the source slice of the signal buffer -- from the tile the kernel staged
 *    anyway, so the separate copy pass (gfx_node_copy_f32: one more read of every source) is not needed.  Requires
 *    c_sig == max(c_sig, c_filt).
 * y [batch, max(c_sig, c_filt), L]; workspace as gfx_biquad_cascade_workspace_bytes (elem_size 4). */
GFX_API int gfx_biquad_cascade_ex_f32(const float* x, float* xcopy, float* y, const float* Bs, const float* As, int batch,
                              int c_sig, int c_filt, int K, long long L, int src_outer, int src_inner, int coef_repeat,
                              void* workspace, size_t workspace_bytes, void* stream);
/* float64 variant: the reference's only known-answer test runs in double
 * (tests/processors/test_filter.py:215-233). */
GFX_API int gfx_biquad_cascade_f64(const double* x, double* y, const double* Bs, const double* As, int batch,
                           int c_sig, int c_filt, int K, long long L, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ---- parameter activations -> biquad coefficients ----------------------------------------------
 * Replaces the elementwise coefficient designers of processors/filter.py (:144-154 BiquadFilter,
 * :303-338 StateVariableFilter, :373-383/:592-604 activations, :416-556 LP/HP/BP/BR/AP,
 * :645-754 Peaking/LowShelf/HighShelf) and the band layout of ParametricEqualizer (eq.py:300-314)
 * with one launch.  n_rows = product of the leading dims of the parameter tensors, each [n_rows, K].
 * family: 0 parametric EQ (flags&1: low shelf / peaks / high shelf layout), 1 peaking, 2 low shelf,
 * 3 high shelf (p0=w0, p1=q_inv, p2=log_gain); 4 low-pass, 5 high-pass, 6 band-pass, 7 band-reject,
 * 8 all-pass (p0=w0, p1=q_inv); 9 stability-constrained direct form (p0=Bs[n_rows,K,3], p1=A1_pre,
 * p2=A2_pre, p3=A0 used when flags&2); 10 state-variable (p0=twoR, p1=G, p2=c_hp, p3=c_bp, p4=c_lp).
 * Output Bs, As [n_rows, K, 3]. */
GFX_API int gfx_biquad_design_f32(int family, const float* p0, const float* p1, const float* p2, const float* p3,
                                  const float* p4, float* Bs, float* As, int n_rows, int K, int flags,
                                  void* stream);

/* ---- stereo <-> mid/side -------------------------------------------------------------------
 * Replaces lr_to_ms / ms_to_lr (processors/core/midside.py:4-17).  x, y [batch, 2, L]:
 *   y[:,0] = (x[:,0] + x[:,1]) * mult,  y[:,1] = (x[:,0] - x[:,1]) * mult
 * mult = 0.5 is lr_to_ms, mult = 1 is ms_to_lr. */
GFX_API int gfx_midside_f32(const float* x, float* y, int batch, long long L, float mult, void* stream);

/* ---- FIR convolution (overlap-save, in-kernel FFT) --------------------------------------------
 * Replaces convolve() / FIRConvolution._native_forward (processors/core/convolution.py:82-83,
 * 119-134: pad, rfft, multiply, irfft, slice through torch.fft / cuFFT) with its intended
 * semantics (true linear convolution; the shipped code is only correct for even Lx+Lh-1):
 *   y[b,c,n] = sum_k h[b,c',k] x[b,c'',n + shift - k],  n < L,  shift = 0 (causal) | filter_len/2 (zerophase)
 *   x [batch, cx, L], h [batch, ch, filter_len], y [batch, max(cx,ch), L]; cx == ch or one of them is 1.
 * `plan` is the twiddle table for n = gfx_fir_fft_size(filter_len): allocate gfx_fft_plan_bytes(n)
 * bytes once per device and fill them with gfx_fft_plan_init.  The workspace holds filter (and,
 * for filter_len > 16384, input-block and output-block) spectra; any size >= the one for batch = 1
 * works, the size returned for the full batch is the fastest.
 * filter_repeat (>= 1, divides batch): that many consecutive batch items share one filter, i.e. h is
 * [batch / filter_repeat, ch, filter_len] -- the 4-D path of render_grafx repeats every node's parameters over
 * the batch of renders (render/graph.py:63-75); filter spectra are then formed once per filter.
 * gfx_fir_set_tuning(long_n, mid_n): FFT size of the partitioned path (4096 | 8192 | 16384) and of the
 * single-partition path for 2048 < filter_len <= mid_n/2 (8192 | 16384); 0 keeps a value.  Changes what
 * gfx_fir_fft_size returns, so fetch the plan again afterwards. */
GFX_API int gfx_fir_fft_size(int filter_len);
GFX_API size_t gfx_fft_plan_bytes(int n);
GFX_API int gfx_fft_plan_init(void* plan, int n, void* stream);
GFX_API size_t gfx_fir_conv_workspace_bytes(int batch, int cx, int ch, long long L, int filter_len, int zerophase);
GFX_API int gfx_fir_conv_f32(const float* x, const float* h, float* y, int batch, int cx, int ch, long long L,
                             int filter_len, int zerophase, int filter_repeat, const void* plan, void* workspace,
                             size_t workspace_bytes, void* stream);
GFX_API int gfx_fir_set_tuning(int long_n, int mid_n);
/* long filters (> 16384 taps): 0 (default, fastest as measured) = four kernels per sweep with the spectra in HBM;
 * 1 = one persistent pipelined launch whose spectra stay in L2 (2.4x less DRAM traffic, ~10 % slower; used when the
 * filter has <= 12 partitions of 8192 taps).  Affects the
 * workspace size: query it again afterwards. */
GFX_API int gfx_fir_set_long_mode(int mode, int lookahead); /* lookahead: pipeline depth in batch items (0 keeps it) */
/* per-bin multiply-accumulate of the long-filter path: 2 (default) = fir_mac3_kernel where it applies (all partitions in
 * one pass, <= 24, rows of <= 32 blocks: packed fp32x2 products, the row's input spectra resident in shared memory),
 * otherwise as 1; 1 = fir_mac2_kernel (one complex bin per thread, up to 24 partitions per pass, inputs through a
 * cp.async ring); 0 = fir_mac_kernel (<= 12 partitions per pass, the round-1 kernel, used when the filter has <= 12
 * partitions; kept for A/B measurements).  They differ in summation order only (~1e-7 relative). */
GFX_API int gfx_fir_set_mac_form(int form);
/* spectra workspace per sweep of the long-filter path in MiB (default 6144; affects gfx_fir_conv_workspace_bytes).  Small
 * sweeps keep the spectra in L2 between the kernels of a sweep at the price of more, smaller launches. */
GFX_API int gfx_fir_set_sweep_mb(int mb);
/* FIRFilter.forward (processors/filter.py:65-77): taps = normalize_impulse(tanh(fir_raw)) -- the activation and the
 * unit-energy scale are applied while the filter spectra are formed (one small reduction kernel + the convolution).
 * fir_raw [batch, ch, filter_len]; workspace: gfx_fir_conv_workspace_bytes(...) + 256 * ceil(4 batch / 256) bytes. */
GFX_API int gfx_fir_filter_f32(const float* x, const float* fir_raw, float* y, int batch, int cx, int ch, long long L,
                               int filter_len, const void* plan, void* workspace, size_t workspace_bytes, void* stream);
/* Same convolution (causal) with the filter given as the UN-NORMALISED impulse response of
 * gfx_reverb_ir_f32 (mode 0: mid/side rows, ms_to_lr = 0; mode 3: left/right rows, ms_to_lr = 1) plus the
 * energies of its raw mid/side rows: normalize_impulse (processors/reverb.py:215-228, core/utils.py:14-18)
 * is applied while the filter spectra are formed, so the normalised IR never exists in memory.
 * ir_raw [batch / filter_repeat, 2, ir_len], energy [batch / filter_repeat, 2], x [batch, cx, L] (cx = 1 or 2),
 * y [batch, 2, L].
 * Plan / workspace as for gfx_fir_conv_f32 with ch = 2. */
GFX_API int gfx_fir_conv_midside_ir_f32(const float* x, const float* ir_raw, const float* energy, float* y, int batch,
                                        int cx, long long L, int ir_len, int ms_to_lr, int filter_repeat,
                                        const void* plan, void* workspace, size_t workspace_bytes, void* stream);

/* ---- reverb impulse-response synthesis ---------------------------------------------------------
 * Replaces STFTMaskedNoiseReverb.compute_stft_mask / compute_ir (processors/reverb.py:161-200:
 * mask, torch.istft) and the ms_to_lr + normalize_impulse of _process_* (reverb.py:215-228).
 *   noise_stft [*, 2, bins, frames] complex64 (re,im interleaved); noise_batch_stride = complex
 *     elements between batch items, 0 when one fixed noise is shared (the default upstream);
 *   init/delta_log_magnitude [batch, 2, bins]; gain_env_log_magnitude [batch, 2, frames] or NULL;
 *   window [n_fft] (the registered hann buffer); ir [batch, 2, ir_len] (output);
 *   energy [batch, 2] (output): sum_t of the squared RAW mid/side rows;
 *   workspace: gfx_reverb_ir_workspace_bytes(batch, n_fft, hop, ir_len) bytes of scratch;
 *   mode 0: ir = raw mid/side response; mode 3: raw left/right (= mid +- side) response -- feed either,
 *     with `energy`, to gfx_fir_conv_midside_ir_f32, which folds the normalisation into the filter spectra;
 *   mode 1: mid/side normalised to unit energy in place; mode 2: left/right, normalised (pseudo_midside).
 * bins = n_fft/2+1, frames = 1 + ir_len/hop.  Geometry: n_fft = 384, hop = 192 (the reference default) runs the tuned
 * kernel; any power-of-two n_fft in 32..4096 with 1 <= hop <= n_fft runs the general two-kernel path; anything else
 * returns GFX_ERR_UNSUPPORTED (and a workspace size of 0). */
GFX_API size_t gfx_reverb_ir_workspace_bytes(int batch, int n_fft, int hop, int ir_len);
GFX_API int gfx_reverb_ir_f32(const float* noise_stft, long long noise_batch_stride, const float* init_log_magnitude,
                              const float* delta_log_magnitude, const float* gain_env_log_magnitude,
                              const float* window, float* ir, float* energy, void* workspace,
                              size_t workspace_bytes, int batch, int n_fft, int hop, int ir_len, int mode,
                              void* stream);

/* ---- filtered-noise-shaping reverb: impulse-response synthesis -----------------------------------
 * Replaces the envelope broadcast of FilteredNoiseShapingReverb.forward (processors/reverb.py:364-380):
 *   ir[b,c,t] = sum_k noise[c,k,offset+t] * gain[b,c,k] * (exp(t*decay[b,c,k]) - fade_gain[b,c,k] * exp(t*fade[b,c,k]))
 * noise [channels, bands, noise_len] (the band-filtered noise buffer); decay/gain/fade/fade_gain [batch, channels,
 * bands] ALREADY activated (decay, fade: log-slopes per sample; fade, fade_gain NULL without `use_fade_in`);
 * ir [batch, channels, ir_len] un-normalised; energy [batch, channels] = sum_t ir^2 (for normalize_impulse, e.g.
 * through gfx_fir_conv_midside_ir_f32 when channels = 2).  bands <= 32. */
GFX_API size_t gfx_noise_shaping_ir_workspace_bytes(int batch, int channels, int ir_len);
GFX_API int gfx_noise_shaping_ir_f32(const float* noise, long long noise_len, long long noise_offset, const float* decay,
                                     const float* gain, const float* fade, const float* fade_gain, float* ir,
                                     float* energy, void* workspace, size_t workspace_bytes, int batch, int channels,
                                     int bands, int ir_len, void* stream);

/* ---- dry/wet mix -----------------------------------------------------------------------------
 * Replaces the mix in DryWet.forward (processors/container.py:62-67):
 *   y[b] = weight[b] * wet[b] + (1 - weight[b]) * dry[b],   dry/wet/y [batch, inner], weight [batch]. */
GFX_API int gfx_drywet_f32(const float* dry, const float* wet, const float* weight, float* y, int batch,
                           long long inner, void* stream);

/* ---- memoryless processors ---------------------------------------------------------------------
 * One streaming pass for the sample-wise processors of processors/stereo.py and processors/nonlinear.py and
 * the weighted accumulation of ParallelMix (processors/container.py:203-216).  x, y [batch, channels, L];
 * every parameter pointer has leading dim batch; dc [batch*channels] (row means to subtract first, from
 * gfx_row_mean_f32: `remove_dc`) or NULL.
 *   op 0 gain          y = x * exp(p0[b,c])                                   (StereoGain, stereo.py:31-38)
 *   op 1 side gain     channels = 2: side of the mid/side pair * exp(p0[b])   (SideGainImager, stereo.py:71-84)
 *   op 2 tanh          p0 log_pre_gain|NULL, p1 log_post_gain|NULL, p2 bias|NULL  (TanhDistortion, nonlinear.py:64-89)
 *   op 3 piecewise tanh p0 log_hardness[b,2], p1 z_threshold[b,2], p2 log_pre_gain|NULL, p3 log_post_gain|NULL
 *                                                                              (PiecewiseTanhDistortion, :159-205)
 *   op 4 power series  p0 basis_weights[b,order], p1 log_pre_gain|NULL        (PowerDistortion, :268-285)
 *   op 5 chebyshev     same parameters                                         (ChebyshevDistortion, :349-384)
 *   op 6 scale-add     y = (flags&4 ? y : 0) + p0[b] * x
 * flags: 1 = post gain is 1 / pre gain (inverse_post_gain), 2 = tanh on every basis function (use_tanh), 4 = accumulate. */
GFX_API int gfx_row_mean_f32(const float* x, float* mean, int rows, long long L, void* stream);
/* mean over the row of x^2: the energy term of rms_difference (processors/core/utils.py:7-11) used by
 * GainStagingRegularization (processors/container.py:284-292); rows = batch (row length channels * L). */
GFX_API int gfx_row_mean_square_f32(const float* x, float* mean, int rows, long long L, void* stream);
GFX_API int gfx_pointwise_f32(int op, const float* x, float* y, int batch, int channels, long long L, const float* p0,
                              const float* p1, const float* p2, const float* p3, const float* dc, int order, int flags,
                              void* stream);

/* ---- node-axis aggregation of the render loop --------------------------------------------------
 * Replaces aggregate_tensor "sum" / "scatter" (render/core.py:101-112, torch.sum / torch_geometric
 * scatter).  src is a strided view [batch, n_src, inner], dst a strided view [batch, n_dst, inner]
 * (strides in elements, so slices of the signal buffer are read and written in place):
 *   dst[b, j] = sum_{i : index[i] == j} src[b, i]      (index == NULL: all sources go to j = 0). */
GFX_API int gfx_node_sum_f32(const float* src, float* dst, const int* index, int batch, int n_src, int n_dst,
                             long long inner, long long src_batch_stride, long long src_node_stride,
                             long long dst_batch_stride, long long dst_node_stride, void* stream);

/* Source write of create_signal_buffer (render/core.py:6-33): strided block copy of a [batch, nodes, inner]
 * view into another (strides in elements), i.e. dst[b, j] = src[b, j]; transposes a batched input
 * [B, V0, C, L] into the node-major signal buffer in one pass. */
GFX_API int gfx_node_copy_f32(const float* src, float* dst, int batch, int nodes, long long inner,
                              long long src_batch_stride, long long src_node_stride, long long dst_batch_stride,
                              long long dst_node_stride, void* stream);

/* ---- backward pass of the biquad cascade ---------------------------
 * Upstream differentiates IIRFilter._process_lfilter (processors/core/iir.py:154-196) through torchaudio's lfilter
 * autograd.  Here the adjoint of a section is the same recursion on the time-reversed gradient (gfx_biquad_cascade_f32
 * with the section's coefficients), and the coefficient gradients are lagged inner products:
 *   out0[row][j] = sum_n u[n] s0[n-j], j = 0..2, and the same for (s1, out1) unless both are NULL (the section's
 *   input and output signals share one read of u); u_reversed != 0 reads u back to front (u is the reversed-time
 *   signal the adjoint recursion produced).  u, s0, s1 [rows, L]; out0, out1 [rows, 3]. */
GFX_API int gfx_lag_dots_f32(const float* u, const float* s0, const float* s1, float* out0, float* out1, int rows,
                             long long L, int u_reversed, void* stream);

/* ---- dynamics: Compressor / NoiseGate, and fused serial chains of them -------------------------
 * Replaces Compressor.forward / NoiseGate.forward (processors/dynamics.py:361-419,598-651), the
 * knees (:443-489, :675-721), TruncatedOnePoleIIRFilter and Ballistics
 * (processors/core/envelope.py:34-60, 84-101) and, with n_stages > 1, a SerialChain of such
 * processors (processors/container.py:116-140) without materialising the intermediate signal.
 *   x, y [batch, channels, L];  every parameter pointer is a device array with leading dim batch.
 * `stages` is a HOST array of n_stages (<= 4) descriptors, consumed before the call returns. */
typedef struct gfx_dynamics_stage {
    int kind;               /* 0 compressor, 1 noise gate */
    int knee;               /* 0 hard, 1 quadratic, 2 exponential, 3 quadratic as shipped in ApproxNoiseGate
                             * (dynamics.py:186-204: R = exp(log_ratio), knee / (2 (W + 1e-3)); kind 1 only) */
    int energy_smoother;    /* 0 none, 1 truncated one-pole ("iir"), 2 ballistics */
    int gain_smoother;      /* same coding */
    int gain_smooth_in_log; /* gain smoother runs on the log-gain */
    int reserved;
    const float* log_threshold; /* [batch] */
    const float* log_ratio;     /* [batch] */
    const float* log_knee;      /* [batch], may be NULL for the hard knee */
    const float* z_alpha_pre;   /* [batch,1] iir / [batch,2] ballistics / NULL */
    const float* z_alpha_post;  /* same for the gain smoother */
    float* hist_pre;            /* [batch, L] scratch; required for stages > 0 with an iir energy smoother when iir_len < L */
    float* hist_post;           /* [batch, L] scratch; required for an iir gain smoother when iir_len < L */
} gfx_dynamics_stage;
GFX_API size_t gfx_dynamics_workspace_bytes(int batch, int n_stages);
/* threads per CTA of the scan (one-pole smoother) variant: 0 = chosen from the row count (default), or 128 / 256 */
GFX_API int gfx_dynamics_set_tuning(int scan_threads);
/* attack / release ballistics: 1 (default) = independent chunks with a warm-up (dynamics_spec_kernel) for every row whose
 * followers forget their state within 16 chunks, the row walk for the others; 0 = always the row walk */
GFX_API int gfx_dynamics_set_ballistics_mode(int mode);
GFX_API int gfx_dynamics_f32(const float* x, float* y, int batch, int channels, long long L,
                             const gfx_dynamics_stage* stages, int n_stages, int iir_len,
                             void* workspace, size_t workspace_bytes, void* stream);
/* The same with parameter rows shared by runs of param_repeat consecutive batch items (render_grafx's 4-D sources,
 * render/graph.py:132-147: upstream expands every node's parameters over the batch of renders): the parameter tensors of
 * every stage are [batch / param_repeat, ...]; hist_pre / hist_post stay [batch, L].  param_repeat == 1 is gfx_dynamics_f32. */
GFX_API int gfx_dynamics_rep_f32(const float* x, float* y, int batch, int channels, long long L,
                                 const gfx_dynamics_stage* stages, int n_stages, int iir_len, int param_repeat,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* Stand-alone envelope smoothers and followers on the same kernel.
 * Replaces: TruncatedOnePoleIIRFilter.forward (processors/core/envelope.py:34-60), Ballistics.forward
 * (core/envelope.py:84-101), BaseEnvelopeFollower.forward (processors/dynamics.py:745-767).
 *   x [batch, channels, L] -> y [batch, L]
 *   detect:   0 mean over channels of x^2 ("energy"), 1 mean of |x| ("amplitude"), 2 x itself (channels == 1)
 *   smoother: 1 truncated one-pole (z [batch,1], relu as upstream), 2 ballistics (z [batch,2], starts from 1)
 *   log_out:  y = log(envelope + 1e-5) (the envelope followers)
 * workspace: gfx_dynamics_workspace_bytes(batch, 1). */
GFX_API int gfx_envelope_f32(const float* x, float* y, int batch, int channels, long long L, int smoother,
                             const float* z, int detect, int log_out, int iir_len, void* workspace,
                             size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GRAFX_B200_H */
