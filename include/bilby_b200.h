/*
 * bilby_b200 - C ABI of the B200-native compact-binary likelihood hot path.
 *
 * Drop-in boundary for bilby's  parameters -> log_likelihood_ratio  path.  The reference (bilby,
 * pure Python) has no FFI of its own; each entry point below names the reference interface it
 * replaces (file:line relative to the bilby source tree).  Plain pointers and sizes only - no
 * torch / numpy types.  "dev" pointers are CUDA device pointers on the handle's device, "host"
 * pointers are ordinary host memory.  All functions return 0 on success, non-zero on failure;
 * bb_last_error() describes the most recent failure of the calling thread.
 *
 * There is NO CPU execution path in this library: every compute entry point launches sm_100a
 * kernels and fails if no CUDA device is usable.
 */
#ifndef BILBY_B200_H
#define BILBY_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define BB_ABI_VERSION 1
#define BB_MAX_DET 4
#define BB_NPARAM 16

/* Column order of the per-sample parameter matrix, row-major double[n][BB_NPARAM].
 * These are the *converted* source-model parameters, i.e. what
 * bilby/gw/waveform_generator.py:260-269 (_format_parameters -> parameter_conversion,
 * bilby/gw/conversion.py:182-283) hands to bilby/gw/source.py:269-348 (lal_binary_black_hole) plus
 * the extrinsic parameters bilby/gw/detector/interferometer.py:303-368 reads.
 * chi_i is the aligned spin component a_i*cos(tilt_i) (conversion.py:146-153). */
enum bb_param {
    BB_MASS_1 = 0,            /* solar masses */
    BB_MASS_2 = 1,
    BB_CHI_1 = 2,
    BB_CHI_2 = 3,
    BB_LUMINOSITY_DISTANCE = 4, /* Mpc */
    BB_THETA_JN = 5,
    BB_PSI = 6,
    BB_PHASE = 7,
    BB_RA = 8,
    BB_DEC = 9,
    BB_GEOCENT_TIME = 10,     /* GPS s; with time marginalisation: start_time (+ jitter added on device) */
    BB_TIME_JITTER = 11,
    BB_LAMBDA_1 = 12,
    BB_LAMBDA_2 = 13
};

enum bb_approximant { BB_IMRPHENOMD = 0, BB_TAYLORF2 = 1 };

enum bb_marginalization {
    BB_MARG_PHASE = 1,    /* bilby/gw/likelihood/base.py:786-792 */
    BB_MARG_DISTANCE = 2, /* base.py:775-784, 879-885 */
    BB_MARG_TIME = 4      /* base.py:794-820 */
};

typedef struct bb_handle bb_handle;

const char* bb_last_error(void);
int bb_abi_version(void);

/* One handle per (process, device): owns the device-resident tiles and scratch. */
int bb_create(int device, bb_handle** out);
void bb_destroy(bb_handle* h);

/* Upload the detector network and data: replaces what GravitationalWaveTransient reads from
 * InterferometerList on every call (base.py:432-439; interferometer.py:551-564 PSD array,
 * strain_data.py:142-159 frequency_mask, :212-233 frequency_domain_strain,
 * detector/geometry.py detector_tensor & vertex).
 *   detector_tensors host double[n_det][9], vertices host double[n_det][3] (metres),
 *   strain host double[n_det][n_freq][2] (re, im), psd host double[n_det][n_freq] (+inf allowed),
 *   mask host uint8[n_det][n_freq].  n_freq = round(duration*sampling_frequency/2)+1. */
int bb_set_network(bb_handle* h, int n_det, int n_freq, double duration, double sampling_frequency,
                   double start_time, const double* detector_tensors, const double* vertices,
                   const double* strain, const double* psd, const unsigned char* mask);

/* Source model selection: replaces WaveformGenerator(frequency_domain_source_model=lal_binary_black_hole /
 * lal_binary_neutron_star, waveform_arguments={waveform_approximant, reference_frequency,
 * minimum_frequency, maximum_frequency}) - source.py:338-342, 422-426.  maximum_frequency <= 0 means
 * "last bin of the grid" (source.py default frequency_array[-1]). */
int bb_set_waveform(bb_handle* h, int approximant, double reference_frequency, double minimum_frequency,
                    double maximum_frequency);

/* Marginalisation state: replaces GravitationalWaveTransient.__init__ set-up (base.py:183-223).
 *   flags: OR of bb_marginalization.
 *   distance table = the FITPACK representation (tx, ty, c) of base.py:929-934's
 *   BoundedRectBivariateSpline over (_d_inner_h_ref_array, _optimal_snr_squared_ref_array), with its
 *   bounding box; ref_dist = priors['luminosity_distance'].rescale(0.5) (base.py:216).
 *   time prior: Uniform [time_min, time_max] (base.py:799-806); jitter = base.py:189-197. */
int bb_set_marginalization(bb_handle* h, int flags, double ref_dist,
                           const double* tx, int nx, const double* ty, int ny, const double* c,
                           double xmin, double xmax, double ymin, double ymax,
                           double time_min, double time_max, int jitter_time);

/* Batched log-likelihood ratio: replaces GravitationalWaveTransient.log_likelihood_ratio
 * (base.py:419-446) evaluated for n parameter rows.  Invalid waveform domains give the reference's
 * sentinel np.nan_to_num(-inf) = -DBL_MAX (base.py:424-425).
 * _device: params/out are device pointers, asynchronous on `stream` (a cudaStream_t, may be NULL).
 * _host:   params/out are host pointers; copies host->device, computes, copies back, synchronises. */
int bb_log_likelihood_ratio_device(bb_handle* h, const double* params_dev, long n, double* out_dev,
                                   void* stream);
int bb_log_likelihood_ratio_host(bb_handle* h, const double* params_host, long n, double* out_host);

/* Cubic-spline calibration (bilby/gw/detector/calibration.py:257-384, applied in
 * interferometer.py:364).  bb_set_calibration describes the model once: n_points spline nodes per
 * detector at linspace(log10_fmin[d], log10_fmax[d], n_points) and the n_points x n_points matrix
 * CubicSpline.nodes_to_spline_coefficients (calibration.py:302-325, row-major, host).  n_points = 0
 * removes the model.  The *_cal_* entry points take, next to the parameter rows, the calibration
 * parameters double[n][n_det][2][n_points] = recalib_{IFO}_amplitude_i, recalib_{IFO}_phase_i. */
int bb_set_calibration(bb_handle* h, int n_points, const double* log10_fmin, const double* log10_fmax,
                       const double* nodes_to_spline_coefficients);
int bb_log_likelihood_ratio_cal_device(bb_handle* h, const double* params_dev, const double* cal_params_dev,
                                       long n, double* out_dev, void* stream);
int bb_log_likelihood_ratio_cal_host(bb_handle* h, const double* params_host, const double* cal_params_host,
                                     long n, double* out_host);
int bb_inner_products_cal_device(bb_handle* h, const double* params_dev, const double* cal_params_dev, long n,
                                 double* out_dev, void* stream);

/* Per-detector inner products: replaces GravitationalWaveTransient.calculate_snrs (base.py:260-354)
 * / Interferometer.inner_product + optimal_snr_squared (interferometer.py:607-640).
 * out double[n][n_det][3] = (Re <h|d>, Im <h|d>, <h|h>) with the reference's 4/T normalisation
 * (gw/utils.py:118-138).  With a frequency-sharded network (bb_set_frequency_shard) these are the
 * rank-local partial sums to be all-reduced. */
int bb_inner_products_device(bb_handle* h, const double* params_dev, long n, double* out_dev, void* stream);

/* Likelihood from (all-reduced) inner products: replaces compute_log_likelihood_from_snrs
 * (base.py:448-477) for the phase / distance marginalised and plain cases.
 * snrs double[n][n_det][3] as produced by bb_inner_products_device. */
int bb_likelihood_from_inner_products_device(bb_handle* h, const double* params_dev, const double* snrs_dev,
                                             long n, double* out_dev, void* stream);

/* Detector-based sky frame and detector time reference: replaces GravitationalWaveTransient.get_sky_frame_parameters
 * (base.py:1091-1137) -> zenith_azimuth_to_ra_dec (gw/utils.py:232-256, gw/geometry.py:215-258, 346-377).
 *   rotation: host double[9] = rotation_matrix_from_delta(vertex_1 - vertex_2) of the two reference detectors
 *             (reference_frame="H1L1"), or NULL for reference_frame="sky".  When set, columns BB_RA / BB_DEC of the
 *             parameter rows hold (azimuth, zenith).
 *   time_reference_vertex: host double[3], vertex [m] of the time-reference detector (time_reference="H1"), or NULL
 *             for "geocent".  When set, column BB_GEOCENT_TIME holds the arrival time at that detector.
 * The conversion runs on the device in front of the prologue kernel for every evaluation entry point. */
int bb_set_reference_frame(bb_handle* h, const double* rotation, const double* time_reference_vertex);
/* (ra, dec, geocent_time) for each parameter row under the current reference frame; out double[n][3]. */
int bb_sky_frame_parameters_device(bb_handle* h, const double* params_dev, long n, double* out_dev, void* stream);

/* Restrict this handle to the contiguous bin range [k_begin, k_end) (frequency sharding of long
 * signals across GPUs).  Pass (0, n_freq) to undo. */
int bb_set_frequency_shard(bb_handle* h, int k_begin, int k_end);

/* Polarisations on the full frequency grid: replaces WaveformGenerator.frequency_domain_strain
 * (waveform_generator.py:113-141) for injections and tests.
 * out double[n][2][n_freq][2]: [plus|cross][bin][re|im]. */
int bb_frequency_domain_strain_device(bb_handle* h, const double* params_dev, long n, double* out_dev,
                                      void* stream);

/* Polarisations on a frequency SEQUENCE: replaces _base_waveform_frequency_sequence (bilby/gw/source.py:
 * 1068-1140, lalsim SimInspiralChooseFDWaveformSequence) as used by the ROQ and relative-binning source models.
 * Every frequency is evaluated (no f_min / f_max masking); first_frequency is the sequence's first element (the
 * f_min of the upstream domain check).  frequencies_dev device double[n_nodes]; out double[n][2][n_nodes][2]. */
int bb_frequency_sequence_strain_device(bb_handle* h, const double* params_dev, long n, const double* frequencies_dev,
                                        int n_nodes, double first_frequency, double* out_dev, void* stream);

/* Detector-frame strain: replaces Interferometer.get_detector_response (interferometer.py:303-368).
 * out double[n][n_det][n_freq][2]. */
int bb_detector_response_device(bb_handle* h, const double* params_dev, long n, double* out_dev, void* stream);

/* Distance-marginalisation lookup table: replaces _create_lookup_table (base.py:994-1018).
 * x_ref host double[nx] (_d_inner_h_ref_array), y_ref host double[ny] (_optimal_snr_squared_ref_array),
 * distance / prior host double[nd]; table_out host double[ny][nx]. */
int bb_build_distance_table(bb_handle* h, const double* x_ref, int nx, const double* y_ref, int ny,
                            const double* distance, const double* prior, int nd, double ref_dist,
                            int phase_marginalization, double* table_out);

/* Antenna patterns and delays (tests / diagnostics): replaces Interferometer.antenna_response
 * (interferometer.py:267-301) and time_delay_from_geocenter (:571-590).
 * out double[n][n_det][3] = (F+, Fx, delay). */
int bb_antenna_response_device(bb_handle* h, const double* params_dev, long n, double* out_dev, void* stream);

/* ln I0 (gw/utils.py:1006-1022), elementwise on device arrays. */
int bb_ln_i0_device(bb_handle* h, const double* x_dev, long n, double* out_dev, void* stream);

/* Interferometer.get_detector_response for caller-supplied polarisation arrays
 * (interferometer.py:303-368; injection path).  plus/cross: device complex double[n_freq][2];
 * params_dev: ONE parameter row; det: detector index; out: device double[n_freq][2]. */
int bb_project_polarizations_device(bb_handle* h, int det, const double* plus_dev, const double* cross_dev,
                                    const double* params_dev, double* out_dev, void* stream);

/* noise_weighted_inner_product over a detector's mask (gw/utils.py:118-138 via
 * interferometer.py:607-640): sum conj(a) b / S * 4/T.  a, b: device complex double[n_freq][2];
 * b == NULL means the detector's own data (Interferometer.inner_product).  out: device double[2]. */
int bb_noise_weighted_inner_product_device(bb_handle* h, int det, const double* a_dev, const double* b_dev,
                                           double* out_dev, void* stream);

/* ---- Frequency-sharded evaluation with the exchange fused into the kernels ---------
 * The reference has no multi-device path (bilby/core/sampler/base_sampler.py:772-800 fans single evaluations out
 * over a process pool); this is the long-signal partition the north star names: every rank (one process per GPU)
 * owns the bin range set with bb_set_frequency_shard and evaluates all samples on it.  The partial
 * (Re<d|h>, Im<d|h>, <h|h>) of every (sample, detector) are stored by K1 itself into every rank's exchange buffer
 * (peer memory over NVLink), a one-warp kernel exchanges arrival flags, and the epilogue sums the partials: no
 * NCCL call and no host synchronisation on the data path.
 *   bb_exchange_create : allocates this rank's buffer for up to max_rows samples and returns its 64-byte CUDA IPC
 *                        handle (ipc_handle_out); world <= 8.
 *   bb_exchange_connect: ipc_handles = the `world` handles in rank order (gathered by the host code, e.g. with
 *                        torch.distributed.all_gather); opens the peers' buffers.
 *   bb_log_likelihood_ratio_sharded_device: collective - every rank calls it with the same rows; out_dev [n] holds
 *                        the full likelihood on every rank (same marginalisations as bb_log_likelihood_ratio_device
 *                        except time / calibration marginalisation and the reduced-order likelihoods).
 *   bb_exchange_status : 0, or 1 + r when rank r's arrival flag was not seen within ~10 s (the wait gives up
 *                        instead of hanging the device; the results of that call are invalid). */
int bb_exchange_create(bb_handle* h, int world, int rank, long max_rows, void* ipc_handle_out);
int bb_exchange_connect(bb_handle* h, const void* ipc_handles);
int bb_log_likelihood_ratio_sharded_device(bb_handle* h, const double* params_dev, long n, double* out_dev, void* stream);
int bb_exchange_status(bb_handle* h, int* status_out);
int bb_exchange_destroy(bb_handle* h);

/* ---- Reduced-order likelihoods ------------------------------------
 * Once one of the two set-up calls below has succeeded, bb_inner_products[_cal]_device and
 * bb_log_likelihood_ratio[_cal]_{device,host} evaluate that likelihood instead of the full-grid one
 * (same parameter rows, same outputs); n_edges = 0 / n_linear = 0 switches back.
 *
 * Relative binning: replaces RelativeBinningGravitationalWaveTransient.calculate_snrs and
 * compute_waveform_ratio_per_interferometer (bilby/gw/likelihood/relative.py:365-430) with the source model
 * lal_binary_*_relative_binning evaluated at the bin edges (bilby/gw/source.py:724-799, fiducial = 0).
 *   bin_freqs host double[n_edges]                  (relative.py:179-240 setup_bins)
 *   fiducial  host double[n_det][n_edges][2]        per_detector_fiducial_waveform_points (relative.py:236-240)
 *   summary   host double[n_det][4][n_edges-1][2]   a0, a1, b0, b1 (relative.py:319-363 compute_summary_data)
 * With BB_MARG_TIME the full-grid reconstruction of relative.py:380-421 is used: pass also
 *   fiducial_grid host double[n_det][n_freq][2] (per_detector_fiducial_waveforms) and bin_inds host int[n_edges];
 *   both may be NULL when time marginalisation is off. */
int bb_set_relative_binning(bb_handle* h, int n_edges, const double* bin_freqs, const double* fiducial,
                            const double* summary, const double* fiducial_grid, const int* bin_inds);

/* Multi-banding (S. Morisaki, arXiv:2104.07813): replaces MBGravitationalWaveTransient.calculate_snrs
 * (bilby/gw/likelihood/multiband.py:728-765, linear-interpolation form of (h, h)) with the source model
 * binary_*_frequency_sequence (bilby/gw/source.py:901-1140) evaluated at the banded frequency points:
 *   <d|h> = conj( sum_k h_det(f_k) linear_coeffs[k] ),  <h|h> = sum_k |h_det(f_k)|^2 quadratic_coeffs[k].
 *   frequencies      host double[n_points]             banded_frequency_points (multiband.py:449-478; duplicates
 *                                                      between neighbouring bands are evaluated twice)
 *   linear_coeffs    host double[n_det][n_points][2]   linear_coeffs[ifo.name]    (multiband.py:529-549)
 *   quadratic_coeffs host double[n_det][n_points]      quadratic_coeffs[ifo.name] (multiband.py:551-611)
 * Runs on the relative-binning kernel K5 in its edge form (no neighbour term); n_points = 0 switches back to the
 * full grid.  Time marginalisation and the IFFT-FFT form of (h, h) are not provided. */
int bb_set_multiband(bb_handle* h, int n_points, const double* frequencies, const double* linear_coeffs,
                     const double* quadratic_coeffs);

/* Time marginalisation of the multi-banded likelihood (MBGravitationalWaveTransient._setup_time_marginalization_multiband
 * and calculate_snrs, bilby/gw/likelihood/multiband.py:714-726, 789-797): the reference's FFT of the scattered
 * strain * linear_coeffs array of n_full = Nbs[-1] / 2 points, evaluated for the times inside the geocent_time prior.
 *   full_index [n_points]: position of every banded point in that array (int(f * durations[0]));  delta_tc = durations[0] / n_full;
 *   the antenna response is taken at beam_pattern_reference_time (Interferometer.reference_time).  Call after
 *   bb_set_multiband; needs bb_set_marginalization with BB_MARG_TIME.  n_full = 0 switches it off. */
int bb_set_multiband_time_marginalization(bb_handle* h, long n_full, const int* full_index, double delta_tc,
                                          double beam_pattern_reference_time);

/* The IFFT-FFT form of (h, h) of the multi-banded likelihood (linear_interpolation=False; bilby/gw/likelihood/
 * multiband.py:613-646 _setup_quadratic_coefficients_ifft_fft, :766-787 calculate_snrs).  Band 0 and the even bins of
 * every band's 2 M-point spectrum are per-point weights and belong in bb_set_multiband's quadratic_coeffs; this call
 * hands over what the odd bins need, for the bands b >= 1 (n_bands of them):
 *   band_m[b] = M^(b) (power of two), band_ks / band_ke = Ks_Ke[b], band_start[b] = start_end_idxs[b][0],
 *   band_norm[b] = 4 / That^(b), sqrt_window [n_points] = square_root_windows,
 *   i_odd = for every band in turn [n_det][M^(b) / 2] = Ibcs[ifo][b].real[1::2].
 * n_bands = 0 switches it off (linear-interpolation form). */
int bb_set_multiband_ifft_fft(bb_handle* h, int n_bands, const int* band_m, const int* band_ks, const int* band_ke,
                              const int* band_start, const double* band_norm, const double* sqrt_window,
                              const double* i_odd);

/* Batched complex FFT, `batch` contiguous transforms of 2^log2n points (2^8 .. 2^18), forward sign e^{-2 pi i},
 * out-of-place, device memory (csrc/bb_fft.cuh: what numpy.fft.fft does in the reference's multiband.py:766-797). */
int bb_fft_device(bb_handle* h, const double* in, double* out, long batch, int log2n, void* stream);

/* ROQ: replaces ROQGravitationalWaveTransient.calculate_snrs, _closest_time_indices, _interp_five_samples and
 * _calculate_d_inner_h_array (bilby/gw/likelihood/roq.py:467-651) with the source model binary_*_roq
 * (bilby/gw/source.py:693-721, 802-898) evaluated at the ROQ frequency nodes.
 *   nodes_linear host double[n_linear], nodes_quadratic host double[n_quadratic]   (weights['frequency_nodes_*'])
 *   weights_linear    host double[n_det][n_time][n_linear][2]   weights['{IFO}_linear'][0]   (roq.py:849-916)
 *   weights_quadratic host double[n_det][n_quadratic]           weights['{IFO}_quadratic'][0] (roq.py:976-1004)
 *   time_samples = (time_start_index + i) * time_step, i < n_time   weights['time_samples']   (roq.py:747-766)
 * Time marginalisation (BB_MARG_TIME): the likelihood's own time grid (roq.py:320-331)
 *   marg_times = marg_time_start + j * marg_delta_tc, j < n_marg_times, antenna response and delays at
 *   beam_pattern_reference_time; the all-times contraction W conj(h_linear) of roq.py:638 runs as one
 *   double-complex GEMM per detector. */
int bb_set_roq(bb_handle* h, int n_linear, const double* nodes_linear, int n_quadratic, const double* nodes_quadratic,
               int n_time, long time_start_index, double time_step, const double* weights_linear,
               const double* weights_quadratic, int n_marg_times, double marg_time_start, double marg_delta_tc,
               double beam_pattern_reference_time);

/* Marginalised-parameter reconstruction, batched: replaces
 * GravitationalWaveTransient.generate_posterior_sample_from_marginalized_likelihood and
 * generate_{time,distance,phase}_sample_from_marginalized_likelihood (bilby/gw/likelihood/base.py:502-773), which
 * bilby.gw.conversion.generate_posterior_samples_from_marginalized_likelihood (conversion.py:2366-2449) maps over the
 * posterior rows with a process pool.  The marginalisation flags of bb_set_marginalization select the steps (time ->
 * distance -> phase, each seeing the values drawn before it).
 *   bb_set_reconstruction_grid: likelihood._distance_array and likelihood.distance_prior_array (base.py:916-919),
 *     needed when distance marginalisation is on.
 *   uniforms_dev [n][3]: the unit-interval draws Interped.sample() (core/prior/base.py:143-164,
 *     core/prior/interpolated.py:88-94) would make for (time, distance, phase); columns of steps that are off are
 *     not read.
 *   out_dev [n][3]: new geocent_time, luminosity_distance, phase (for steps that are off: the time column is not
 *     written, distance / phase are copied from the row).  NaN for waveform-domain errors.
 * With calibration marginalisation loaded (bb_set_calibration_marginalization; no time marginalisation, cal_params_dev
 * NULL) column 0 changes meaning: uniforms_dev[.][0] is the unit-interval draw of Generator.choice over the response
 * curves (base.py:544-578) and out_dev[.][0] the drawn recalib_index; the distance and phase steps then use the inner
 * products of that curve (base.py:289-290).
 * Restrictions: full-grid likelihood (not ROQ / relative binning), uniform geocent_time prior not wider than 0.249 s,
 * 32768 / sampling_frequency a power of two. */
int bb_set_reconstruction_grid(bb_handle* h, const double* distance_array, const double* distance_prior_array,
                               int n_distance);
int bb_reconstruct_marginalized_device(bb_handle* h, const double* params_dev, const double* cal_params_dev, long n,
                                       const double* uniforms_dev, double* out_dev, void* stream);

/* Calibration marginalisation: GravitationalWaveTransient(calibration_marginalization=True)
 * (bilby/gw/likelihood/base.py:333-346 the per-curve arrays of calculate_snrs, :860-877
 * calibration_marginalized_likelihood, :1037-1051 set-up).  curves: likelihood.calibration_draws, as
 * [n_det][n_curves][n_freq] complex (re, im) on the FULL frequency grid (entries outside the frequency mask are
 * ignored).  After this call bb_log_likelihood_ratio_{device,host} return the calibration-marginalised likelihood,
 * combined with phase / distance marginalisation as bb_set_marginalization says; time marginalisation and
 * per-sample calibration parameters are refused.  n_curves = 0 switches it off. */
int bb_set_calibration_marginalization(bb_handle* h, int n_curves, const double* curves);

/* The dense FP64 contraction of the path on the tensor cores (mma.sync.m8n8k4.f64, csrc/bb_gemm.cuh): what the
 * reference writes as `linear_matrix @ conj(h_linear)` (bilby/gw/likelihood/roq.py:604-651), the calibration-curve
 * products of bilby/gw/likelihood/base.py:305-346 and the per-basis-element inverse FFTs of roq.py:849-918.
 *   C[batch][m][n] (+)= alpha * sum_seg sum_k A_seg[batch][m][k] * B_seg[batch][n][k]
 * A [m x k] and B [n x k] are K-contiguous row-major, C [m x n] row-major; is_complex: elements are (re, im) pairs.
 * Leading dimensions and strides are in elements; device memory.  (Inside the library the operands are produced
 * directly in the kernel's packed tile layout; this entry packs row-major inputs first, then synchronises.) */
int bb_contract_device(bb_handle* h, int is_complex, int m, int n, int k, int n_seg, long seg_stride_a, long seg_stride_b,
                       int n_batch, long batch_stride_a, long batch_stride_b, long batch_stride_c, double alpha,
                       const double* a, long lda, const double* b, long ldb, int accumulate, double* c, long ldc,
                       void* stream);

/* Set-up artefact builder (no handle needed): the linear ROQ weights of n_det detectors sharing one basis,
 * ROQGravitationalWaveTransient._set_weights_linear (bilby/gw/likelihood/roq.py:849-918) - one zero-padded inverse FFT
 * of data * conj(basis_b) / PSD per basis element, of which only the time samples [lo, lo + n_win) are kept.
 *   d_over_s   [n_det][n_freq_sel] complex (re, im): strain / PSD at the basis frequencies that overlap the data
 *   basis      [n_basis][n_freq_sel] complex: linear basis at the same frequencies (NOT conjugated)
 *   bin_index  [n_freq_sel]: index of each frequency in the n_time-point transform (ifo_idxs + f_min * T, roq.py:901)
 *   out        [n_det][n_win][n_basis] complex = (4 n_time / T) * ifft(...)[lo : lo + n_win].T */
int bb_build_roq_linear_weights(int device, int n_det, int n_freq_sel, const double* d_over_s, int n_basis, const double* basis,
                                const int* bin_index, long n_time, long lo, int n_win, double duration, double* out);

/* Set-up artefact builder: the quadratic ROQ weights of n_det detectors sharing one basis,
 * ROQGravitationalWaveTransient._set_weights_quadratic (bilby/gw/likelihood/roq.py:976-1004):
 *   out[d][b] = (4 / T) sum_j basis_real[b][j] * inv_psd[d][j]       (host arrays in, host array out) */
int bb_build_roq_quadratic_weights(int device, int n_det, int n_freq_sel, const double* inv_psd, int n_basis,
                                   const double* basis_real, double duration, double* out);

/* Set-up artefact builder: the summary data of RelativeBinningGravitationalWaveTransient.compute_summary_data
 * (bilby/gw/likelihood/relative.py:319-363) from the handle's data tiles.  Bin b covers the grid bins
 * [bin_start[b], bin_start[b + 1]) and has centre frequency centre[b]; fiducial = per-detector fiducial waveform on the
 * full grid, [n_det][n_freq] complex (re, im), host.  out [n_det][4][n_bins] complex (re, im), host: a0, a1, b0, b1. */
int bb_build_relbin_summary_data(bb_handle* h, int n_bins, const int* bin_start, const double* centre,
                                 const double* fiducial, double* out);

/* Device-resident sampling front end (a batched sampler whose points never cross PCIe).
 * Replaces, per sample, PriorDict.rescale (bilby/core/prior/dict.py:647-666) over analytic priors
 * (bilby/core/prior/analytical.py: DeltaFunction :45, PowerLaw / LogUniform :107, Uniform :214, Cosine :415, Sine :475,
 * Gaussian :535), the parameter conversion the waveform generator applies (bilby/gw/conversion.py:182-283
 * convert_to_lal_binary_black_hole_parameters, :286-348 convert_to_lal_binary_neutron_star_parameters, :1826-1985
 * generate_component_masses, tidal maps :1187-1264) and the packing into parameter rows (enum bb_param).
 *   kind      prior class; (a, b, c) = (minimum, maximum, alpha) for PowerLaw (LogUniform: alpha = -1),
 *             (minimum, maximum) for Uniform / Sine / Cosine, (mu, sigma) for Gaussian, (peak) for DeltaFunction
 *   key       which sampled parameter the dimension is (enum bb_source_key, bilby's parameter names; with a
 *             detector-based sky frame / time reference azimuth, zenith and {IFO}_time take the ra, dec and
 *             geocent_time keys, like the rows)
 * bb_set_sampling_priors: the sampled dimensions in the sampler's order plus the fixed parameters (DeltaFunction /
 * float priors, the side effects of the marginalisations); neutron_star selects the tidal conversion.
 * bb_rows_from_unit_cube_device: unit_dev [n][n_dim] -> rows_dev [n][16] (and theta_dev [n][n_dim] if not NULL);
 * bb_rows_from_theta_device: theta_dev [n][n_dim] sampled parameters -> rows.  Precessing spins (tilts other than 0
 * or pi) give NaN spin columns, which the evaluation kernels answer with the waveform-error sentinel. */
enum bb_prior_kind {
    BB_PRIOR_DELTA = 0, BB_PRIOR_UNIFORM = 1, BB_PRIOR_POWERLAW = 2, BB_PRIOR_SINE = 3, BB_PRIOR_COSINE = 4,
    BB_PRIOR_GAUSSIAN = 5
};
enum bb_source_key {
    BB_KEY_MASS_1 = 0, BB_KEY_MASS_2 = 1, BB_KEY_CHIRP_MASS = 2, BB_KEY_MASS_RATIO = 3, BB_KEY_TOTAL_MASS = 4,
    BB_KEY_SYMMETRIC_MASS_RATIO = 5, BB_KEY_CHI_1 = 6, BB_KEY_CHI_2 = 7, BB_KEY_A_1 = 8, BB_KEY_A_2 = 9,
    BB_KEY_TILT_1 = 10, BB_KEY_TILT_2 = 11, BB_KEY_COS_TILT_1 = 12, BB_KEY_COS_TILT_2 = 13,
    BB_KEY_LUMINOSITY_DISTANCE = 14, BB_KEY_THETA_JN = 15, BB_KEY_COS_THETA_JN = 16, BB_KEY_PSI = 17, BB_KEY_PHASE = 18,
    BB_KEY_DELTA_PHASE = 19, BB_KEY_RA = 20, BB_KEY_DEC = 21, BB_KEY_GEOCENT_TIME = 22, BB_KEY_TIME_JITTER = 23,
    BB_KEY_LAMBDA_1 = 24, BB_KEY_LAMBDA_2 = 25, BB_KEY_LAMBDA_TILDE = 26, BB_KEY_DELTA_LAMBDA_TILDE = 27,
    BB_KEY_COUNT = 28
};
typedef struct bb_prior_spec {
    int kind;
    int key;
    double a, b, c;
} bb_prior_spec;
int bb_set_sampling_priors(bb_handle* h, int n_dim, const bb_prior_spec* specs, int n_fixed, const int* fixed_keys,
                           const double* fixed_values, int neutron_star);
int bb_rows_from_unit_cube_device(bb_handle* h, const double* unit_dev, long n, double* theta_dev, double* rows_dev,
                                  void* stream);
int bb_rows_from_theta_device(bb_handle* h, const double* theta_dev, long n, double* rows_dev, void* stream);

/* The device math layer on arrays (csrc/bb_math.cuh: what the per-bin loops use instead of numpy's sin / cos / arctan
 * and of IEEE division), exposed so that it can be checked on its own.  function 0: out[2 i], out[2 i + 1] =
 * sin(pi x_i), cos(pi x_i) (|x| < 2^50, NaN beyond); 1: out[i] = atan(x_i); 2: out[i] = 1 / x_i for normal x_i > 0. */
int bb_math_device(bb_handle* h, int function, const double* x_dev, long n, double* out_dev, void* stream);

/* Measurement hooks (bench.py).  With profiling enabled the handle brackets every launch of the
 * dominant kernel (K1, the fused inner-product kernel) with CUDA events on the launching stream;
 * bb_profile_read synchronises and returns the summed duration and the number of launches since the
 * last read.  bb_fp64_peak runs a register-resident DFMA stream kernel and returns the achieved
 * TFLOP/s - the FP64 roofline denominator MEASURED_PEAKS.json does not provide. */
int bb_profile_enable(bb_handle* h, int on);
int bb_profile_read(bb_handle* h, double* k1_ms, long* k1_launches);
int bb_fp64_peak(bb_handle* h, double* tflops);
/* The same for the FP64 tensor path (mma.sync.m8n8k4.f64 stream): the denominator of the dense contractions. */
int bb_fp64_tensor_peak(bb_handle* h, double* tflops);

/* Number of kernel launches issued through this handle so far (bench.py's gpu_launches claim). */
long bb_launch_count(bb_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* BILBY_B200_H */
