// Device-resident sampling front end (included by bb_kernels.cu): unit cube -> sampled parameters -> source-model
// parameters -> parameter rows, one thread per sample, so that a batched sampler's points never cross PCIe.
//
//   bb_prior_rescale      bilby/core/prior/dict.py:647-666 PriorDict.rescale with the analytic priors of
//                         bilby/core/prior/analytical.py: DeltaFunction :45, PowerLaw / LogUniform :107, Uniform :214,
//                         Cosine :415, Sine :475, Gaussian :535 (operation order of the reference)
//   bb_convert_and_pack   bilby/gw/conversion.py:1826-1985 generate_component_masses (require_add=False),
//                         :182-283 convert_to_lal_binary_black_hole_parameters (aligned spins, cos_* angles,
//                         delta_phase), :286-348 convert_to_lal_binary_neutron_star_parameters and the tidal maps
//                         :1187-1264, then the row layout of include/bilby_b200.h (enum bb_param)
//
// Which keys are sampled / fixed is the same for every sample of a run, so the presence tests of the reference's
// dictionaries become a bit mask that is uniform over the grid.  Precessing spins (tilts other than 0 or pi) are
// outside the aligned-spin source models: the host path raises, this kernel writes NaN into the spin column, which
// the evaluation kernels answer with the waveform-error sentinel.
#pragma once

struct BBSampling {
    int n_dim = 0;
    unsigned present = 0;                    // keys that are sampled or fixed
    int bns = 0;                             // neutron-star conversion (tidal keys)
    bb_prior_spec* d_specs = nullptr;        // [n_dim]
    double fixed[BB_KEY_COUNT];
};

__device__ __forceinline__ double bb_prior_rescale(const bb_prior_spec& p, double u) {
    switch (p.kind) {
        case BB_PRIOR_DELTA:
            return p.a;                                                      // peak * val ** 0
        case BB_PRIOR_UNIFORM:
            return p.a + u * (p.b - p.a);
        case BB_PRIOR_POWERLAW: {
            if (p.c == -1.0) return p.a * exp(u * log(p.b / p.a));
            const double e = 1.0 + p.c;
            const double lo = pow(p.a, e), hi = pow(p.b, e);
            return pow(lo + u * (hi - lo), 1.0 / e);
        }
        case BB_PRIOR_SINE: {
            const double norm = 1.0 / (cos(p.a) - cos(p.b));
            return acos(cos(p.a) - u / norm);
        }
        case BB_PRIOR_COSINE: {
            const double norm = 1.0 / (sin(p.b) - sin(p.a));
            return asin(u / norm + sin(p.a));
        }
        case BB_PRIOR_GAUSSIAN:
            return p.a + erfinv(2.0 * u - 1.0) * 1.4142135623730951 * p.b;  // mu + erfinv(2 u - 1) 2**0.5 sigma
    }
    return nan("");
}

__device__ __forceinline__ double bb_eta_to_q(double eta) {      // conversion.py:969-989
    const double temp = 1.0 / eta / 2.0 - 1.0;
    return temp - sqrt(temp * temp - 1.0);
}

__device__ __forceinline__ double bb_sign(double x) { return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : x); }

__device__ __forceinline__ void bb_convert_and_pack(double* v, unsigned m, int bns, double* row) {
#define BB_HAS(k) ((m >> (k)) & 1u)
#define BB_PUT(k, val) do { v[k] = (val); m |= 1u << (k); } while (0)
    // ---- generate_component_masses
    if (BB_HAS(BB_KEY_MASS_1)) {
        if (!BB_HAS(BB_KEY_MASS_2)) {
            if (BB_HAS(BB_KEY_TOTAL_MASS)) {
                BB_PUT(BB_KEY_MASS_2, v[BB_KEY_TOTAL_MASS] - v[BB_KEY_MASS_1]);
            } else {
                if (!BB_HAS(BB_KEY_MASS_RATIO) && BB_HAS(BB_KEY_SYMMETRIC_MASS_RATIO))
                    BB_PUT(BB_KEY_MASS_RATIO, bb_eta_to_q(v[BB_KEY_SYMMETRIC_MASS_RATIO]));
                if (BB_HAS(BB_KEY_MASS_RATIO)) BB_PUT(BB_KEY_MASS_2, v[BB_KEY_MASS_RATIO] * v[BB_KEY_MASS_1]);
            }
        }
    } else if (BB_HAS(BB_KEY_MASS_2)) {
        if (BB_HAS(BB_KEY_TOTAL_MASS)) {
            BB_PUT(BB_KEY_MASS_1, v[BB_KEY_TOTAL_MASS] - v[BB_KEY_MASS_2]);
        } else {
            if (!BB_HAS(BB_KEY_MASS_RATIO) && BB_HAS(BB_KEY_SYMMETRIC_MASS_RATIO))
                BB_PUT(BB_KEY_MASS_RATIO, bb_eta_to_q(v[BB_KEY_SYMMETRIC_MASS_RATIO]));
            if (BB_HAS(BB_KEY_MASS_RATIO)) BB_PUT(BB_KEY_MASS_1, 1.0 / v[BB_KEY_MASS_RATIO] * v[BB_KEY_MASS_2]);
        }
    } else {
        if (BB_HAS(BB_KEY_TOTAL_MASS)) {
            if (BB_HAS(BB_KEY_MASS_RATIO)) {
            } else if (BB_HAS(BB_KEY_SYMMETRIC_MASS_RATIO)) {
                BB_PUT(BB_KEY_MASS_RATIO, bb_eta_to_q(v[BB_KEY_SYMMETRIC_MASS_RATIO]));
            } else if (BB_HAS(BB_KEY_CHIRP_MASS)) {
                BB_PUT(BB_KEY_SYMMETRIC_MASS_RATIO, pow(v[BB_KEY_CHIRP_MASS] / v[BB_KEY_TOTAL_MASS], 5.0 / 3.0));
                BB_PUT(BB_KEY_MASS_RATIO, bb_eta_to_q(v[BB_KEY_SYMMETRIC_MASS_RATIO]));
            }
        } else if (BB_HAS(BB_KEY_CHIRP_MASS)) {
            if (!BB_HAS(BB_KEY_MASS_RATIO) && BB_HAS(BB_KEY_SYMMETRIC_MASS_RATIO))
                BB_PUT(BB_KEY_MASS_RATIO, bb_eta_to_q(v[BB_KEY_SYMMETRIC_MASS_RATIO]));
            if (BB_HAS(BB_KEY_MASS_RATIO)) {
                const double q = v[BB_KEY_MASS_RATIO];
                BB_PUT(BB_KEY_TOTAL_MASS, v[BB_KEY_CHIRP_MASS] * pow(1.0 + q, 1.2) / pow(q, 0.6));
            }
        }
        if (BB_HAS(BB_KEY_TOTAL_MASS) && BB_HAS(BB_KEY_MASS_RATIO)) {
            const double m1 = v[BB_KEY_TOTAL_MASS] / (1.0 + v[BB_KEY_MASS_RATIO]);
            BB_PUT(BB_KEY_MASS_1, m1);
            BB_PUT(BB_KEY_MASS_2, m1 * v[BB_KEY_MASS_RATIO]);
        }
    }
    const double m1 = BB_HAS(BB_KEY_MASS_1) ? v[BB_KEY_MASS_1] : nan("");
    const double m2 = BB_HAS(BB_KEY_MASS_2) ? v[BB_KEY_MASS_2] : nan("");
    // ---- spins (conversion.py:229-253) and the aligned-spin shortcut a cos(tilt)
    double chi_row[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int k_chi = BB_KEY_CHI_1 + i, k_a = BB_KEY_A_1 + i, k_t = BB_KEY_TILT_1 + i, k_ct = BB_KEY_COS_TILT_1 + i;
        if (BB_HAS(k_chi)) {
            if (!BB_HAS(k_a)) {
                BB_PUT(k_a, fabs(v[k_chi]));
                BB_PUT(k_ct, bb_sign(v[k_chi]));
            } else {
                BB_PUT(k_ct, v[k_a] == 0.0 ? 1.0 : v[k_chi] / v[k_a]);
            }
        }
        if (BB_HAS(k_ct)) BB_PUT(k_t, acos(v[k_ct]));
        const double a = BB_HAS(k_a) ? v[k_a] : 0.0, t = BB_HAS(k_t) ? v[k_t] : 0.0;
        const bool aligned = (a == 0.0) || (t == 0.0) || (t == 3.141592653589793);
        chi_row[i] = aligned ? a * cos(t) : nan("");
    }
    if (BB_HAS(BB_KEY_COS_THETA_JN)) BB_PUT(BB_KEY_THETA_JN, acos(v[BB_KEY_COS_THETA_JN]));
    const double psi = BB_HAS(BB_KEY_PSI) ? v[BB_KEY_PSI] : 0.0;
    if (BB_HAS(BB_KEY_DELTA_PHASE)) {         // conversion.py:266-272
        const double two_pi = 6.283185307179586;
        double r = fmod(v[BB_KEY_DELTA_PHASE] - bb_sign(cos(v[BB_KEY_THETA_JN])) * psi, two_pi);
        if (r != 0.0 && r < 0.0) r += two_pi;                  // numpy's mod takes the divisor's sign
        BB_PUT(BB_KEY_PHASE, r);
    }
    // ---- tides
    double l1 = 0.0, l2 = 0.0;
    if (bns && (BB_HAS(BB_KEY_LAMBDA_1) || BB_HAS(BB_KEY_LAMBDA_2) || BB_HAS(BB_KEY_LAMBDA_TILDE) ||
                BB_HAS(BB_KEY_DELTA_LAMBDA_TILDE))) {
        double eta = (m1 * m2) / ((m1 + m2) * (m1 + m2));
        eta = fmin(eta, 0.25);
        const double eta2 = eta * eta;
        if (BB_HAS(BB_KEY_DELTA_LAMBDA_TILDE)) {          // conversion.py:1187-1231
            const double lt = v[BB_KEY_LAMBDA_TILDE], dlt = v[BB_KEY_DELTA_LAMBDA_TILDE];
            const double sq = sqrt(1.0 - 4.0 * eta);
            const double c1 = 1.0 + 7.0 * eta - 31.0 * eta2;
            const double c2 = sq * (1.0 + 9.0 * eta - 11.0 * eta2);
            const double c3 = sq * (1.0 - 13272.0 / 1319.0 * eta + 8944.0 / 1319.0 * eta2);
            const double c4 = 1.0 - 15910.0 / 1319.0 * eta + 32850.0 / 1319.0 * eta2 + 3380.0 / 1319.0 * (eta2 * eta);
            l1 = (13.0 * lt / 8.0 * (c3 - c4) - 2.0 * dlt * (c1 - c2)) / ((c1 + c2) * (c3 - c4) - (c1 - c2) * (c3 + c4));
            l2 = (13.0 * lt / 8.0 * (c3 + c4) - 2.0 * dlt * (c1 + c2)) / ((c1 - c2) * (c3 + c4) - (c1 + c2) * (c3 - c4));
        } else if (BB_HAS(BB_KEY_LAMBDA_TILDE)) {         // conversion.py:1234-1264
            const double q = m2 / m1, qm5 = pow(q, -5.0);
            l1 = 13.0 / 8.0 * v[BB_KEY_LAMBDA_TILDE] /
                 ((1.0 + 7.0 * eta - 31.0 * eta2) * (1.0 + qm5)
                  + sqrt(1.0 - 4.0 * eta) * (1.0 + 9.0 * eta - 11.0 * eta2) * (1.0 - qm5));
            l2 = l1 / pow(q, 5.0);
        } else {
            l1 = BB_HAS(BB_KEY_LAMBDA_1) ? v[BB_KEY_LAMBDA_1] : 0.0;
            // lambda_2 follows lambda_1 m1^5 / m2^5 when only lambda_1 is given (conversion.py:339-346)
            l2 = BB_HAS(BB_KEY_LAMBDA_2) ? v[BB_KEY_LAMBDA_2]
                                         : (BB_HAS(BB_KEY_LAMBDA_1) ? l1 * pow(m1, 5.0) / pow(m2, 5.0) : 0.0);
        }
    }
    row[BB_MASS_1] = m1;
    row[BB_MASS_2] = m2;
    row[BB_CHI_1] = chi_row[0];
    row[BB_CHI_2] = chi_row[1];
    row[BB_LUMINOSITY_DISTANCE] = BB_HAS(BB_KEY_LUMINOSITY_DISTANCE) ? v[BB_KEY_LUMINOSITY_DISTANCE] : nan("");
    row[BB_THETA_JN] = BB_HAS(BB_KEY_THETA_JN) ? v[BB_KEY_THETA_JN] : nan("");
    row[BB_PSI] = psi;
    row[BB_PHASE] = BB_HAS(BB_KEY_PHASE) ? v[BB_KEY_PHASE] : nan("");
    row[BB_RA] = BB_HAS(BB_KEY_RA) ? v[BB_KEY_RA] : 0.0;
    row[BB_DEC] = BB_HAS(BB_KEY_DEC) ? v[BB_KEY_DEC] : 0.0;
    row[BB_GEOCENT_TIME] = BB_HAS(BB_KEY_GEOCENT_TIME) ? v[BB_KEY_GEOCENT_TIME] : 0.0;
    row[BB_TIME_JITTER] = BB_HAS(BB_KEY_TIME_JITTER) ? v[BB_KEY_TIME_JITTER] : 0.0;
    row[BB_LAMBDA_1] = l1;
    row[BB_LAMBDA_2] = l2;
    row[14] = 0.0;
    row[15] = 0.0;
#undef BB_HAS
#undef BB_PUT
}

// from_unit != 0: `in` holds unit-cube points, rescaled through the priors (and written to theta_out if given);
// from_unit == 0: `in` already holds the sampled parameters in the order of the prior table.
__global__ void __launch_bounds__(128)
bb_sample_rows_kernel(const double* __restrict__ in, long n, int n_dim, const bb_prior_spec* __restrict__ specs,
                      BBSampling cfg, int from_unit, double* __restrict__ theta_out, double* __restrict__ rows) {
    __shared__ bb_prior_spec sp[BB_KEY_COUNT];
    for (int i = threadIdx.x; i < n_dim; i += blockDim.x) sp[i] = specs[i];
    __syncthreads();
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    double v[BB_KEY_COUNT];
#pragma unroll
    for (int k = 0; k < BB_KEY_COUNT; ++k) v[k] = cfg.fixed[k];
    for (int j = 0; j < n_dim; ++j) {
        const double x = in[s * n_dim + j];
        const double th = from_unit ? bb_prior_rescale(sp[j], x) : x;
        if (theta_out) theta_out[s * n_dim + j] = th;
        // a dynamic key: the value array lives in local memory, which this set-up-rate kernel can afford
        v[sp[j].key] = th;
    }
    double row[BB_NPARAM];
    bb_convert_and_pack(v, cfg.present, cfg.bns, row);
#pragma unroll
    for (int k = 0; k < BB_NPARAM; ++k) rows[s * BB_NPARAM + k] = row[k];
}

static void bb_sampling_release(bb_handle* h) {
    if (!h->sampling) return;
    cudaFree(h->sampling->d_specs);
    delete h->sampling;
    h->sampling = nullptr;
}

extern "C" int bb_set_sampling_priors(bb_handle* h, int n_dim, const bb_prior_spec* specs, int n_fixed,
                                      const int* fixed_keys, const double* fixed_values, int neutron_star) {
    if (!h) return bb_fail("bb_set_sampling_priors: null handle");
    if (n_dim < 0 || n_dim > BB_KEY_COUNT || n_fixed < 0 || n_fixed > BB_KEY_COUNT)
        return bb_fail("bb_set_sampling_priors: at most BB_KEY_COUNT sampled and fixed keys");
    if ((n_dim && !specs) || (n_fixed && (!fixed_keys || !fixed_values))) return bb_fail("bb_set_sampling_priors: null table");
    BB_CUDA(cudaSetDevice(h->device));
    if (!h->sampling) h->sampling = new BBSampling();
    BBSampling& sm = *h->sampling;
    unsigned present = 0;
    for (int k = 0; k < BB_KEY_COUNT; ++k) sm.fixed[k] = 0.0;
    for (int j = 0; j < n_dim; ++j) {
        const bb_prior_spec& p = specs[j];
        if (p.key < 0 || p.key >= BB_KEY_COUNT) return bb_fail("bb_set_sampling_priors: unknown key");
        if (p.kind < BB_PRIOR_DELTA || p.kind > BB_PRIOR_GAUSSIAN) return bb_fail("bb_set_sampling_priors: unknown prior kind");
        if (present >> p.key & 1u) return bb_fail("bb_set_sampling_priors: a key appears twice");
        present |= 1u << p.key;
    }
    for (int j = 0; j < n_fixed; ++j) {
        const int k = fixed_keys[j];
        if (k < 0 || k >= BB_KEY_COUNT) return bb_fail("bb_set_sampling_priors: unknown fixed key");
        if (present >> k & 1u) return bb_fail("bb_set_sampling_priors: a key appears twice");
        present |= 1u << k;
        sm.fixed[k] = fixed_values[j];
    }
    cudaFree(sm.d_specs);
    sm.d_specs = nullptr;
    if (n_dim) {
        BB_CUDA(cudaMalloc(&sm.d_specs, n_dim * sizeof(bb_prior_spec)));
        BB_CUDA(cudaMemcpy(sm.d_specs, specs, n_dim * sizeof(bb_prior_spec), cudaMemcpyHostToDevice));
    }
    sm.n_dim = n_dim;
    sm.present = present;
    sm.bns = neutron_star ? 1 : 0;
    return 0;
}

static int bb_sample_rows(bb_handle* h, const double* in_dev, long n, int from_unit, double* theta_dev, double* rows_dev,
                          void* stream) {
    if (!h || !h->sampling) return bb_fail("sampling front end: bb_set_sampling_priors was not called");
    if (n <= 0) return 0;
    if ((h->sampling->n_dim && !in_dev) || !rows_dev) return bb_fail("sampling front end: null buffer");
    BB_CUDA(cudaSetDevice(h->device));
    bb_sample_rows_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        in_dev, n, h->sampling->n_dim, h->sampling->d_specs, *h->sampling, from_unit, theta_dev, rows_dev);
    h->launches++;
    BB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int bb_rows_from_unit_cube_device(bb_handle* h, const double* unit_dev, long n, double* theta_dev,
                                             double* rows_dev, void* stream) {
    return bb_sample_rows(h, unit_dev, n, 1, theta_dev, rows_dev, stream);
}

extern "C" int bb_rows_from_theta_device(bb_handle* h, const double* theta_dev, long n, double* rows_dev, void* stream) {
    return bb_sample_rows(h, theta_dev, n, 0, nullptr, rows_dev, stream);
}
