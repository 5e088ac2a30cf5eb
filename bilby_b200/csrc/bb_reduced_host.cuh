// Host side of the reduced-order likelihoods: set-up entry points of the C ABI and kernel launchers
// (included by bb_kernels.cu after bb_reduced.cuh).
#pragma once

static void bb_reduced_clear(bb_handle* h) {
    for (void* p : h->red_bufs) cudaFree(p);
    h->red_bufs.clear();
    delete h->rb;
    delete h->rq;
    h->rb = nullptr;
    h->rq = nullptr;
    cudaFree(h->d_roq_V); cudaFree(h->d_roq_Y); cudaFree(h->d_roq_hh);
    h->d_roq_V = h->d_roq_Y = nullptr;
    h->d_roq_hh = nullptr;
    h->roq_chunk = 0;
    h->roq_y_elems = 0;
    cudaFree(h->d_mb_idx); cudaFree(h->d_mb_E); cudaFree(h->d_mb_V); cudaFree(h->d_mb_Y); cudaFree(h->d_mb_hh);
    h->d_mb_idx = nullptr;
    h->d_mb_E = h->d_mb_V = h->d_mb_Y = nullptr;
    h->d_mb_hh = nullptr;
    h->mb_nfull = 0;
    h->mb_chunk = 0;
    h->mb_y_cap = 0;
    h->mb_nrow = 0;
    h->mb_bands.clear();
    cudaFree(h->d_mb_sqrtw); cudaFree(h->d_mb_Z); cudaFree(h->d_mb_Z2); cudaFree(h->d_mb_Z3);
    h->d_mb_sqrtw = nullptr;
    h->d_mb_Z = h->d_mb_Z2 = h->d_mb_Z3 = nullptr;
    h->mb_z_cap = 0;
    h->kind = 0;
}

template <typename T>
static int bb_red_upload(bb_handle* h, const T* host, size_t count, const T** dev) {
    T* d = nullptr;
    BB_CUDA(cudaMalloc(&d, count * sizeof(T)));
    h->red_bufs.push_back(d);
    BB_CUDA(cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *dev = d;
    return 0;
}

static int bb_upload_nodes(bb_handle* h, const double* f, int n, BBNodes* out) {
    std::vector<double> u(n), lf(n), q34(n), t3(n), x3(n), u7(n);
    for (int i = 0; i < n; ++i) {
        if (!(f[i] > 0.0)) return bb_fail("frequency nodes must be positive");
        u[i] = pow(f[i], -1.0 / 6.0);
        lf[i] = log(f[i]);
        q34[i] = pow(f[i], 0.75);
        // the powers bb_wave forms per node, in its operation order (bb_wave_cols reads them instead)
        t3[i] = u[i] * u[i];
        x3[i] = f[i] * t3[i] * t3[i];
        u7[i] = u[i] * (t3[i] * t3[i] * t3[i]);
    }
    out->n = n;
    if (bb_red_upload(h, f, n, &out->f)) return 1;
    if (bb_red_upload(h, u.data(), n, &out->u)) return 1;
    if (bb_red_upload(h, lf.data(), n, &out->lf)) return 1;
    if (bb_red_upload(h, q34.data(), n, &out->q34)) return 1;
    if (bb_red_upload(h, t3.data(), n, &out->t3)) return 1;
    if (bb_red_upload(h, x3.data(), n, &out->x3)) return 1;
    if (bb_red_upload(h, u7.data(), n, &out->u7)) return 1;
    const int rows = (n + 31) / 32;
    std::vector<double> blk((size_t)rows * BB_NB_ROW);
    for (int i = 0; i < rows * 32; ++i) {
        const int j = i < n ? i : n - 1;
        double* b = blk.data() + (size_t)(i / 32) * BB_NB_ROW + (i % 32);
        b[BB_NB_F] = f[j]; b[BB_NB_T3] = t3[j]; b[BB_NB_X3] = x3[j]; b[BB_NB_U7] = u7[j];
        b[BB_NB_LF] = lf[j]; b[BB_NB_Q34] = q34[j];
    }
    if (bb_red_upload(h, blk.data(), blk.size(), &out->blk)) return 1;
    return 0;
}

// K5's edge tables in rows of 32 edges: etab[row][det]{lin_c[32] (double2), cross_g[32] (double2, if given), quad_e[32]}
static int bb_upload_edge_rows(bb_handle* h, int nd, int np, const std::vector<double2>& lc, const std::vector<double>& qe,
                               const std::vector<double2>* cg, const double** out) {
    const int et = cg ? 160 : 96, rows = np / 32;
    std::vector<double> tab((size_t)rows * nd * et, 0.0);
    for (int d = 0; d < nd; ++d)
        for (int j = 0; j < np; ++j) {
            double* t = tab.data() + ((size_t)(j / 32) * nd + d) * et;
            const int l = j % 32;
            t[2 * l] = lc[(size_t)d * np + j].x;
            t[2 * l + 1] = lc[(size_t)d * np + j].y;
            if (cg) {
                t[64 + 2 * l] = (*cg)[(size_t)d * np + j].x;
                t[64 + 2 * l + 1] = (*cg)[(size_t)d * np + j].y;
            }
            t[(cg ? 128 : 64) + l] = qe[(size_t)d * np + j];
        }
    return bb_red_upload(h, tab.data(), tab.size(), out);
}

static double2 bb_cinv(double re, double im) {
    // 1 / (re + i im); zero fiducial strain gives inf/nan like the reference's division (relative.py:373)
    const double den = re * re + im * im;
    return make_double2(re / den, -im / den);
}

extern "C" int bb_set_relative_binning(bb_handle* h, int n_edges, const double* bin_freqs, const double* fiducial,
                                       const double* summary, const double* fiducial_grid, const int* bin_inds) {
    if (!h || !h->have_network) return bb_fail("bb_set_relative_binning: network not set");
    BB_CUDA(cudaSetDevice(h->device));
    bb_reduced_clear(h);
    if (n_edges == 0) return 0;
    if (n_edges < 2 || !bin_freqs || !fiducial || !summary) return bb_fail("bb_set_relative_binning: bad arguments");
    const int nd = h->net.n_det, nb = n_edges - 1, nf = h->net.n_freq;
    BBRelbinDev* rb = new BBRelbinDev();
    memset(rb, 0, sizeof(*rb));
    h->rb = rb;
    if (bb_upload_nodes(h, bin_freqs, n_edges, &rb->edges)) return 1;
    std::vector<double2> ginv((size_t)nd * n_edges), a0((size_t)nd * nb), a1((size_t)nd * nb);
    std::vector<double> b0((size_t)nd * nb), b1((size_t)nd * nb), iw(nb), centre(nb);
    for (int d = 0; d < nd; ++d) {
        for (int j = 0; j < n_edges; ++j) {
            const double* g = fiducial + ((size_t)d * n_edges + j) * 2;
            ginv[(size_t)d * n_edges + j] = bb_cinv(g[0], g[1]);
        }
        const double* sd = summary + (size_t)d * 4 * nb * 2;
        for (int b = 0; b < nb; ++b) {
            a0[(size_t)d * nb + b] = make_double2(sd[(0 * nb + b) * 2], sd[(0 * nb + b) * 2 + 1]);
            a1[(size_t)d * nb + b] = make_double2(sd[(1 * nb + b) * 2], sd[(1 * nb + b) * 2 + 1]);
            b0[(size_t)d * nb + b] = sd[(2 * nb + b) * 2];
            b1[(size_t)d * nb + b] = sd[(3 * nb + b) * 2];
        }
    }
    for (int b = 0; b < nb; ++b) {
        iw[b] = 1.0 / (bin_freqs[b + 1] - bin_freqs[b]);
        centre[b] = (bin_freqs[b + 1] + bin_freqs[b]) / 2;
    }
    if (bb_red_upload(h, ginv.data(), ginv.size(), &rb->ginv)) return 1;
    if (bb_red_upload(h, a0.data(), a0.size(), &rb->a0)) return 1;
    if (bb_red_upload(h, a1.data(), a1.size(), &rb->a1)) return 1;
    if (bb_red_upload(h, b0.data(), b0.size(), &rb->b0)) return 1;
    if (bb_red_upload(h, b1.data(), b1.size(), &rb->b1)) return 1;
    if (bb_red_upload(h, iw.data(), iw.size(), &rb->inv_width)) return 1;
    if (bb_red_upload(h, centre.data(), centre.size(), &rb->centre)) return 1;
    {
        // edge form of the sums (bb_relbin_edge_sample in bb_reduced.cuh), zero-padded to whole rows of 32 edges
        const int np = (n_edges + 31) / 32 * 32;
        std::vector<double2> lc((size_t)nd * np, make_double2(0.0, 0.0)), cg((size_t)nd * np, make_double2(0.0, 0.0));
        std::vector<double> qe((size_t)nd * np, 0.0);
        for (int d = 0; d < nd; ++d)
            for (int j = 0; j < n_edges; ++j) {
                const double2 gi = ginv[(size_t)d * n_edges + j];
                double cx = 0.0, cy = 0.0, e = 0.0;
                if (j >= 1) {                 // bin j-1, whose right edge is j
                    const size_t b = (size_t)d * nb + j - 1;
                    cx += 0.5 * a0[b].x + a1[b].x * iw[j - 1];
                    cy += 0.5 * a0[b].y + a1[b].y * iw[j - 1];
                    e += 0.25 * b0[b] + b1[b] * iw[j - 1];
                    const double2 gl = ginv[(size_t)d * n_edges + j - 1];
                    const double hb = 0.5 * b0[b];
                    // (1/h0_j) conj(1/h0_{j-1})
                    cg[(size_t)d * np + j] = make_double2(hb * (gi.x * gl.x + gi.y * gl.y), hb * (gi.y * gl.x - gi.x * gl.y));
                }
                if (j <= nb - 1) {            // bin j, whose left edge is j
                    const size_t b = (size_t)d * nb + j;
                    cx += 0.5 * a0[b].x - a1[b].x * iw[j];
                    cy += 0.5 * a0[b].y - a1[b].y * iw[j];
                    e += 0.25 * b0[b] - b1[b] * iw[j];
                }
                // C = conj(1/h0_j) (cx + i cy)
                lc[(size_t)d * np + j] = make_double2(cx * gi.x + cy * gi.y, cy * gi.x - cx * gi.y);
                qe[(size_t)d * np + j] = e * (gi.x * gi.x + gi.y * gi.y);
            }
        if (bb_red_upload(h, lc.data(), lc.size(), &rb->lin_c)) return 1;
        if (bb_red_upload(h, qe.data(), qe.size(), &rb->quad_e)) return 1;
        if (bb_red_upload(h, cg.data(), cg.size(), &rb->cross_g)) return 1;
        if (bb_upload_edge_rows(h, nd, np, lc, qe, &cg, &rb->etab)) return 1;
        rb->ne_pad = np;
    }
    if (fiducial_grid && bin_inds) {
        // series factor of relative.py:417-420: (4/T) h0 conj(d) / S on the full grid.  d/S comes from the tiles'
        // source arrays, which the handle no longer has on the host: read d_ds back (it already carries 4/T).
        std::vector<double2> ds((size_t)nd * h->n_pad);
        BB_CUDA(cudaMemcpy(ds.data(), h->d_ds, ds.size() * sizeof(double2), cudaMemcpyDeviceToHost));
        std::vector<double2> pg((size_t)nd * nf);
        for (int d = 0; d < nd; ++d)
            for (int k = 0; k < nf; ++k) {
                const double* g = fiducial_grid + ((size_t)d * nf + k) * 2;
                const double2 q = ds[(size_t)d * h->n_pad + k];          // (4/T) d / S
                // h0 conj(d/S): (g0 + i g1)(qx - i qy)
                pg[(size_t)d * nf + k] = make_double2(g[0] * q.x + g[1] * q.y, g[1] * q.x - g[0] * q.y);
            }
        std::vector<int> bok(nf, -1);
        for (int b = 0; b < nb; ++b) {
            const int hi = (b == nb - 1) ? bin_inds[b + 1] + 1 : bin_inds[b + 1];     // bin_sizes[-1] += 1
            for (int k = bin_inds[b]; k < hi && k < nf; ++k) bok[k] = b;
        }
        if (bb_red_upload(h, pg.data(), pg.size(), &rb->pgrid)) return 1;
        if (bb_red_upload(h, bok.data(), bok.size(), &rb->bin_of_k)) return 1;
    }
    h->rb_fmin = bin_freqs[0];
    h->kind = 1;
    return 0;
}

extern "C" int bb_set_multiband(bb_handle* h, int n_points, const double* frequencies, const double* linear_coeffs,
                                const double* quadratic_coeffs) {
    if (!h || !h->have_network) return bb_fail("bb_set_multiband: network not set");
    BB_CUDA(cudaSetDevice(h->device));
    bb_reduced_clear(h);
    if (n_points == 0) return 0;
    if (n_points < 1 || !frequencies || !linear_coeffs || !quadratic_coeffs) return bb_fail("bb_set_multiband: bad arguments");
    const int nd = h->net.n_det;
    BBRelbinDev* rb = new BBRelbinDev();
    memset(rb, 0, sizeof(*rb));
    h->rb = rb;
    if (bb_upload_nodes(h, frequencies, n_points, &rb->edges)) return 1;
    // edge form of K5: C_k = conj(linear_coeffs[k]), E_k = quadratic_coeffs[k], no neighbour term (cross_g stays NULL)
    const int np = (n_points + 31) / 32 * 32;
    std::vector<double2> lc((size_t)nd * np, make_double2(0.0, 0.0));
    std::vector<double> qe((size_t)nd * np, 0.0);
    for (int d = 0; d < nd; ++d)
        for (int k = 0; k < n_points; ++k) {
            const double* c = linear_coeffs + ((size_t)d * n_points + k) * 2;
            lc[(size_t)d * np + k] = make_double2(c[0], -c[1]);
            qe[(size_t)d * np + k] = quadratic_coeffs[(size_t)d * n_points + k];
        }
    if (bb_red_upload(h, lc.data(), lc.size(), &rb->lin_c)) return 1;
    if (bb_red_upload(h, qe.data(), qe.size(), &rb->quad_e)) return 1;
    if (bb_upload_edge_rows(h, nd, np, lc, qe, nullptr, &rb->etab)) return 1;
    rb->ne_pad = np;
    h->rb_fmin = frequencies[0];
    for (int k = 1; k < n_points; ++k) if (frequencies[k] < h->rb_fmin) h->rb_fmin = frequencies[k];
    h->kind = 1;
    return 0;
}

extern "C" int bb_set_multiband_time_marginalization(bb_handle* h, long n_full, const int* full_index, double delta_tc,
                                                     double beam_pattern_reference_time) {
    if (!h || !h->rb || h->kind != 1 || h->rb->cross_g) return bb_fail("bb_set_multiband_time_marginalization: call bb_set_multiband first");
    BB_CUDA(cudaSetDevice(h->device));
    cudaFree(h->d_mb_idx); cudaFree(h->d_mb_E);
    h->d_mb_idx = nullptr;
    h->d_mb_E = nullptr;
    h->mb_nfull = 0;
    h->mb_nrow = 0;
    if (n_full <= 0) return 0;
    if (!full_index || !(delta_tc > 0.0)) return bb_fail("bb_set_multiband_time_marginalization: bad arguments");
    const int np = h->rb->edges.n;
    for (int p = 0; p < np; ++p)
        if (full_index[p] < 0 || full_index[p] >= n_full) return bb_fail("bb_set_multiband_time_marginalization: index outside the full array");
    BB_CUDA(cudaMalloc(&h->d_mb_idx, (size_t)np * sizeof(int)));
    BB_CUDA(cudaMemcpy(h->d_mb_idx, full_index, (size_t)np * sizeof(int), cudaMemcpyHostToDevice));
    h->mb_nfull = n_full;
    h->mb_dtc = delta_tc;
    h->mb_ref_time = beam_pattern_reference_time;
    return 0;
}

static long bb_red_grid(const bb_handle* h, long n);

extern "C" int bb_set_multiband_ifft_fft(bb_handle* h, int n_bands, const int* band_m, const int* band_ks,
                                        const int* band_ke, const int* band_start, const double* band_norm,
                                        const double* sqrt_window, const double* i_odd) {
    if (!h || !h->rb || h->kind != 1 || h->rb->cross_g) return bb_fail("bb_set_multiband_ifft_fft: call bb_set_multiband first");
    BB_CUDA(cudaSetDevice(h->device));
    h->mb_bands.clear();
    cudaFree(h->d_mb_sqrtw);
    h->d_mb_sqrtw = nullptr;
    if (n_bands <= 0) return 0;
    if (!band_m || !band_ks || !band_ke || !band_start || !band_norm || !sqrt_window || !i_odd)
        return bb_fail("bb_set_multiband_ifft_fft: bad arguments");
    const int nd = h->net.n_det, np = h->rb->edges.n;
    size_t total = 0;
    for (int b = 0; b < n_bands; ++b) {
        const int M = band_m[b];
        int l2 = 0;
        while ((1 << l2) < M) ++l2;
        if ((1 << l2) != M || l2 < 8 || l2 > 18) return bb_fail("bb_set_multiband_ifft_fft: M^(b) must be 2^8 .. 2^18");
        if (band_ks[b] < 1 || band_ke[b] >= M / 2 || band_ke[b] < band_ks[b] || band_start[b] < 0
            || band_start[b] + (band_ke[b] - band_ks[b]) >= np)
            return bb_fail("bb_set_multiband_ifft_fft: band bins outside (0, M / 2) or outside the banded points");
        total += (size_t)nd * (M / 2);
    }
    double* d_io = nullptr;
    if (bb_red_upload(h, i_odd, total, (const double**)&d_io)) return 1;
    BB_CUDA(cudaMalloc(&h->d_mb_sqrtw, (size_t)np * sizeof(double)));
    BB_CUDA(cudaMemcpy(h->d_mb_sqrtw, sqrt_window, (size_t)np * sizeof(double), cudaMemcpyHostToDevice));
    size_t off = 0;
    for (int b = 0; b < n_bands; ++b) {
        BBMbBand bd;
        bd.M = band_m[b];
        bd.log2M = 0;
        while ((1 << bd.log2M) < bd.M) ++bd.log2M;
        bd.Ks = band_ks[b]; bd.Ke = band_ke[b]; bd.start = band_start[b];
        bd.norm = band_norm[b];
        bd.i_odd = d_io + off;
        off += (size_t)nd * (bd.M / 2);
        h->mb_bands.push_back(bd);
    }
    return 0;
}

// adds the odd-bin terms of every band b >= 1 to target[(s * NDET + d) * stride + offset], s in [s0, s0 + m)
template <int NDET, int APPROX, bool CAL>
static int bb_mb_add_ifft_fft_terms(bb_handle* h, long s0, long m, double* target, int stride, int offset, cudaStream_t st) {
    const BBRelbinDev& rb = *h->rb;
    const size_t smem = (size_t)BB_RED_WARPS * (BC_NCOEF + (CAL ? NDET * 4 * h->cal.n_points : 0)) * sizeof(double);
    BB_CUDA(cudaFuncSetAttribute(bb_mb_band_fill_kernel<NDET, APPROX, CAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int maxM = 0;
    for (const BBMbBand& bd : h->mb_bands) maxM = bd.M > maxM ? bd.M : maxM;
    // transforms per pass: three buffers of chunk * NDET * M complex numbers, ~1.5 GB in total
    long chunk = (long)(0.5e9 / ((double)NDET * maxM * sizeof(double2)));
    if (chunk > m) chunk = m;
    if (chunk < 1) chunk = 1;
    const size_t need = (size_t)chunk * NDET * maxM;
    if (need > h->mb_z_cap) {
        cudaFree(h->d_mb_Z); cudaFree(h->d_mb_Z2); cudaFree(h->d_mb_Z3);
        h->d_mb_Z = h->d_mb_Z2 = h->d_mb_Z3 = nullptr;
        h->mb_z_cap = 0;
        BB_CUDA(cudaMalloc(&h->d_mb_Z, need * sizeof(double2)));
        BB_CUDA(cudaMalloc(&h->d_mb_Z2, need * sizeof(double2)));
        BB_CUDA(cudaMalloc(&h->d_mb_Z3, need * sizeof(double2)));
        h->mb_z_cap = need;
    }
    for (long c0 = 0; c0 < m; c0 += chunk) {
        const long mc = (m - c0) < chunk ? (m - c0) : chunk;
        for (const BBMbBand& bd : h->mb_bands) {
            const long nt = mc * NDET;
            const size_t elems = (size_t)nt * bd.M;
            BB_CUDA(cudaMemsetAsync(h->d_mb_Z, 0, elems * sizeof(double2), st));
            bb_mb_band_fill_kernel<NDET, APPROX, CAL><<<(unsigned)bb_red_grid(h, mc), BB_RED_THREADS, smem, st>>>(
                h->d_coef, s0 + c0, mc, rb, bd, h->d_mb_sqrtw, h->d_calrec, h->cal, h->d_mb_Z);
            BB_CUDA(cudaGetLastError());
            if (bb_fft_forward(h->d_mb_Z, h->d_mb_Z2, h->d_mb_Z3, nt, bd.log2M, h->sm_count, st)) return 1;
            bb_mb_band_modulate_kernel<<<(unsigned)(h->sm_count * 8), 256, 0, st>>>(h->d_mb_Z3, (long)elems, bd.M);
            if (bb_fft_forward(h->d_mb_Z3, h->d_mb_Z2, h->d_mb_Z, nt, bd.log2M, h->sm_count, st)) return 1;
            bb_mb_band_reduce_kernel<<<(unsigned)((nt * 32 + 255) / 256), 256, 0, st>>>(
                h->d_mb_Z, mc, NDET, bd, target + (size_t)c0 * NDET * stride, stride, offset);
            BB_CUDA(cudaGetLastError());
            h->launches += 7;
        }
    }
    return 0;
}

extern "C" int bb_set_roq(bb_handle* h, int n_linear, const double* nodes_linear, int n_quadratic,
                          const double* nodes_quadratic, int n_time, long time_start_index, double time_step,
                          const double* weights_linear, const double* weights_quadratic, int n_marg_times,
                          double marg_time_start, double marg_delta_tc, double beam_pattern_reference_time) {
    if (!h || !h->have_network) return bb_fail("bb_set_roq: network not set");
    BB_CUDA(cudaSetDevice(h->device));
    bb_reduced_clear(h);
    if (n_linear == 0) return 0;
    if (n_linear < 1 || n_quadratic < 1 || n_time < 5 || !nodes_linear || !nodes_quadratic || !weights_linear
        || !weights_quadratic || !(time_step > 0.0))
        return bb_fail("bb_set_roq: bad arguments");
    const int nd = h->net.n_det;
    BBRoqDev* rq = new BBRoqDev();
    memset(rq, 0, sizeof(*rq));
    h->rq = rq;
    if (bb_upload_nodes(h, nodes_linear, n_linear, &rq->lin)) return 1;
    if (bb_upload_nodes(h, nodes_quadratic, n_quadratic, &rq->quad)) return 1;
    {
        // K7's GEMM operand, packed (bb_gemm.cuh bb_pk: tiles of 128 ROQ times x slabs of 16 nodes, zero padded)
        const long S = (n_linear + 15) / 16;
        const size_t per_det = bb_pk_elems(n_time, n_linear, BB_GEMM_TR_B);
        std::vector<double2> wp((size_t)nd * per_det, make_double2(0.0, 0.0));
        const double2* w = (const double2*)weights_linear;
        for (int d = 0; d < nd; ++d)
            for (int t = 0; t < n_time; ++t)
                for (int j = 0; j < n_linear; ++j)
                    wp[(size_t)d * per_det + bb_pk(t, j, BB_GEMM_TR_B, S)] = w[((size_t)d * n_time + t) * n_linear + j];
        if (bb_red_upload(h, wp.data(), wp.size(), &rq->W)) return 1;
    }
    {
        // node-blocked copy for K6: W2[d][p][t][l] = W[d][t][32 p + l]
        const int nblk = (n_linear + 31) / 32;
        std::vector<double2> w2((size_t)nd * nblk * n_time * 32, make_double2(0.0, 0.0));
        const double2* w = (const double2*)weights_linear;
        for (int d = 0; d < nd; ++d)
            for (int t = 0; t < n_time; ++t)
                for (int j = 0; j < n_linear; ++j)
                    w2[(((size_t)d * nblk + (j >> 5)) * n_time + t) * 32 + (j & 31)] = w[((size_t)d * n_time + t) * n_linear + j];
        if (bb_red_upload(h, w2.data(), w2.size(), &rq->W2)) return 1;
    }
    if (bb_red_upload(h, weights_quadratic, (size_t)nd * n_quadratic, &rq->wq)) return 1;
    rq->n_time = n_time;
    rq->time_start_index = time_start_index;
    rq->time_step = time_step;
    rq->n_marg = n_marg_times;
    rq->marg_start = marg_time_start;
    rq->marg_dtc = marg_delta_tc;
    h->marg.roq_dtc = marg_delta_tc;
    h->roq_ref_time = beam_pattern_reference_time;
    double fmin = nodes_linear[0];
    for (int i = 0; i < n_linear; ++i) fmin = nodes_linear[i] < fmin ? nodes_linear[i] : fmin;
    for (int i = 0; i < n_quadratic; ++i) fmin = nodes_quadratic[i] < fmin ? nodes_quadratic[i] : fmin;
    h->roq_fmin = fmin;           // first of the unique (sorted) nodes: f_min of the sequence call
    h->kind = 2;
    return 0;
}

// ---- launchers ------------------------------------------------------------------------------------
static long bb_red_grid(const bb_handle* h, long n) {
    long blocks = (n + BB_RED_WARPS - 1) / BB_RED_WARPS;
    // 32 CTAs of 256 threads per SM (BB_RED_GRID_PER_SM overrides it): 3 are resident, the others start as the first
    // ones drain.  Measured on relative binning, 1e6 samples: 4 per SM 6.07e8 eval/s, 8: 6.42e8, 16: 6.53e8, 32: 6.59e8,
    // 64: 6.59e8, one CTA per 8 samples: 6.27e8; exactly the resident CTAs (no partial last wave): 6.15e8
    static const long per_sm = [] { const char* e = getenv("BB_RED_GRID_PER_SM"); const long v = e ? atol(e) : 0; return v > 0 ? v : 32L; }();
    const long cap = (long)h->sm_count * per_sm;
    return blocks < cap ? blocks : cap;
}

template <int NDET, int APPROX, bool CAL>
static int bb_launch_reduced_t(bb_handle* h, long n, double* out, cudaStream_t st) {
    const size_t smem = (size_t)2 * BB_RED_WARPS * (BC_NCOEF + (CAL ? NDET * 4 * h->cal.n_points : 0)) * sizeof(double);
    const long grid = bb_red_grid(h, n);
    BBProfScope prof(h, st);
    if (h->kind == 1) {
        if (h->rb->cross_g) {
            BB_CUDA(cudaFuncSetAttribute(bb_relbin_kernel<NDET, APPROX, CAL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            bb_relbin_kernel<NDET, APPROX, CAL, true><<<(unsigned)grid, BB_RED_THREADS, smem, st>>>(
                h->d_coef, n, *h->rb, h->d_calrec, h->cal, out);
        } else {            // multi-banding (bb_set_multiband)
            BB_CUDA(cudaFuncSetAttribute(bb_relbin_kernel<NDET, APPROX, CAL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            bb_relbin_kernel<NDET, APPROX, CAL, false><<<(unsigned)grid, BB_RED_THREADS, smem, st>>>(
                h->d_coef, n, *h->rb, h->d_calrec, h->cal, out);
            // IFFT-FFT form of (h, h): the odd-bin terms of the bands b >= 1 join <h|h> of every detector
            if (!h->mb_bands.empty() && bb_mb_add_ifft_fft_terms<NDET, APPROX, CAL>(h, 0, n, out, 3, 2, st)) return 1;
        }
    } else {
        const size_t smem_k6 = (size_t)2 * BB_ROQ_WARPS * (BC_NCOEF + (CAL ? NDET * 4 * h->cal.n_points : 0)) * sizeof(double)
                               + (size_t)BB_ROQ_WARPS * NDET * 5 * 32 * sizeof(double2) + BB_ROQ_WARPS * sizeof(unsigned long long);
        long grid_k6 = (n + BB_ROQ_WARPS - 1) / BB_ROQ_WARPS;
        static const long k6_per_sm = [] { const char* e = getenv("BB_ROQ_GRID_PER_SM"); const long v = e ? atol(e) : 0; return v > 0 ? v : (long)BB_ROQ_CTAS; }();
        if (grid_k6 > k6_per_sm * h->sm_count) grid_k6 = k6_per_sm * h->sm_count;
        BB_CUDA(cudaFuncSetAttribute(bb_roq_kernel<NDET, APPROX, CAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_k6));
        bb_roq_kernel<NDET, APPROX, CAL><<<(unsigned)grid_k6, BB_ROQ_THREADS, smem_k6, st>>>(
            h->d_coef, n, *h->rq, h->d_calrec, h->cal, out);
    }
    h->launches++;
    BB_CUDA(cudaGetLastError());
    return 0;
}

template <int NDET, int APPROX, bool CAL>
static int bb_launch_roq_time_marg_t(bb_handle* h, long n, double* out, cudaStream_t st) {
    const BBRoqDev& rq = *h->rq;
    if (rq.n_marg < 1) return bb_fail("ROQ time marginalisation: bb_set_roq was called without a time grid");
    const int nl = rq.lin.n, nt = rq.n_time;
    // ROQ rows any sample can touch.  Only times with t + jitter inside the geocent_time prior contribute (base.py:799-806)
    // and the detector time is (t + jitter) - start + delay with |delay| <= |vertex| / c, so the five-point stencils stay
    // inside [tmin - start - dmax, tmax - start + dmax]: the reference contracts all n_time rows (roq.py:604-651), the
    // weights span twice the light-crossing time (roq.py:747-765), a quarter of them is never read.
    double dmax = 0.0;
    for (int d = 0; d < NDET; ++d) {
        const double* v = h->net.vertex[d];
        const double r = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) / BB_C_SI;
        dmax = r > dmax ? r : dmax;
    }
    dmax *= 1.0 + 1e-9;
    const double ts0 = (double)rq.time_start_index * rq.time_step, space = (double)(rq.time_start_index + 1) * rq.time_step - ts0;
    long r_lo = (long)floor((h->marg.time_min - h->net.start_time - dmax - ts0) / space) - 3;
    long r_hi = (long)floor((h->marg.time_max - h->net.start_time + dmax - ts0) / space) + 3;
    if (r_lo < 0) r_lo = 0;
    if (r_hi > nt - 1) r_hi = nt - 1;
    if (getenv("BB_ROQ_NOTRIM") || r_hi < r_lo) { r_lo = 0; r_hi = nt - 1; }
    const int row0 = (int)(r_lo / BB_GEMM_TR_B) * BB_GEMM_TR_B;          // whole row tiles of the packed weights
    const int nrow = (int)(r_hi - row0 + 1);
    // chunk the batch so that one Y buffer [NDET][chunk][nrow] stays below ~1 GB; two buffers: the interpolation +
    // logsumexp kernel of chunk c runs on the auxiliary stream under the GEMM of chunk c + 1
    size_t chunk = (size_t)(1.0e9 / ((double)NDET * nrow * sizeof(double2)));
    chunk = chunk / 64 * 64;
    if (chunk > (size_t)n) chunk = (size_t)n;
    if (chunk < 1) chunk = 1;
    const size_t v_elems = (size_t)NDET * bb_pk_elems((long)chunk, nl, BB_GEMM_TR_A(true));
    const size_t y_elems = (size_t)NDET * chunk * nrow;
    if (chunk > h->roq_chunk || y_elems > h->roq_y_elems) {
        cudaFree(h->d_roq_V); cudaFree(h->d_roq_Y); cudaFree(h->d_roq_hh);
        h->d_roq_V = h->d_roq_Y = nullptr;
        h->d_roq_hh = nullptr;
        h->roq_chunk = 0;
        BB_CUDA(cudaMalloc(&h->d_roq_V, v_elems * sizeof(double2)));
        BB_CUDA(cudaMemsetAsync(h->d_roq_V, 0, v_elems * sizeof(double2), st));
        BB_CUDA(cudaMalloc(&h->d_roq_Y, 2 * y_elems * sizeof(double2)));
        BB_CUDA(cudaMalloc(&h->d_roq_hh, 2 * chunk * sizeof(double)));
        h->roq_chunk = chunk;
        h->roq_y_elems = y_elems;
    }
    const int n_chunks = (int)((n + (long)chunk - 1) / (long)chunk);
    if (!h->aux) BB_CUDA(cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking));
    while ((int)h->tm_events.size() < 2 * n_chunks) {
        cudaEvent_t e;
        BB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->tm_events.push_back(e);
    }
    const size_t smem = (size_t)BB_RED_WARPS * (BC_NCOEF + (CAL ? NDET * 4 * h->cal.n_points : 0)) * sizeof(double);
    BB_CUDA(cudaFuncSetAttribute(bb_roq_hlinear_kernel<NDET, APPROX, CAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BBProfScope prof(h, st);
    for (int c = 0; c < n_chunks; ++c) {
        const long s0 = (long)c * (long)chunk;
        const long m = (n - s0) < (long)chunk ? (n - s0) : (long)chunk;
        const long grid = bb_red_grid(h, m);
        double2* Y = h->d_roq_Y + (size_t)(c & 1) * h->roq_y_elems;
        double* hh = h->d_roq_hh + (size_t)(c & 1) * h->roq_chunk;
        if (c >= 2) BB_CUDA(cudaStreamWaitEvent(st, h->tm_events[2 * (c - 2) + 1], 0));      // Y / hh buffer free again
        bb_roq_hlinear_kernel<NDET, APPROX, CAL><<<(unsigned)grid, BB_RED_THREADS, smem, st>>>(
            h->d_coef, s0, m, rq, h->d_calrec, h->cal, h->d_roq_V, hh);
        BB_CUDA(cudaGetLastError());
        {
            // Y_d[s][t - row0] = sum_i V_d[s][i] W_d[t][i] for every detector: one batched DMMA GEMM (bb_gemm.cuh)
            BBGemmArgs ga{};
            ga.A[0] = h->d_roq_V;
            ga.B[0] = rq.W + (size_t)(row0 / BB_GEMM_TR_B) * ((nl + 15) / 16) * BB_GEMM_TR_B * BB_PK;
            ga.C = reinterpret_cast<double*>(Y);
            ga.slabs_a = ga.slabs_b = (nl + 15) / 16;
            ga.slab0 = 0; ga.n_slabs = (nl + 15) / 16;
            ga.batch_a = (long)bb_pk_elems(m, nl, BB_GEMM_TR_A(true));
            ga.batch_b = (long)bb_pk_elems(nt, nl, BB_GEMM_TR_B);
            ga.batch_c = (long)m * nrow;
            ga.ldc = nrow;
            ga.M = (int)m; ga.N = nrow; ga.n_seg = 1; ga.n_batch = NDET; ga.accumulate = 0; ga.alpha = 1.0;
            if (bb_gemm_nt(true, ga, h->sm_count, st)) return 1;
        }
        BB_CUDA(cudaEventRecord(h->tm_events[2 * c], st));
        BB_CUDA(cudaStreamWaitEvent(h->aux, h->tm_events[2 * c], 0));
        // 128-thread CTAs (14K registers): one becomes resident per SM in what the next chunk's GEMM (2 x 128 threads x
        // 192 registers) leaves free
        const long grid_e = (m + 3) / 4 < 4L * h->sm_count ? (m + 3) / 4 : 4L * h->sm_count;
        bb_roq_time_marg_kernel<NDET><<<(unsigned)grid_e, 128, 0, h->aux>>>(
            h->d_coef, s0, m, rq, Y, row0, nrow, hh, h->marg, h->net.start_time, out);
        BB_CUDA(cudaGetLastError());
        BB_CUDA(cudaEventRecord(h->tm_events[2 * c + 1], h->aux));
        h->launches += 3;
    }
    BB_CUDA(cudaStreamWaitEvent(st, h->tm_events[2 * (n_chunks - 1) + 1], 0));      // join
    if (n_chunks > 1) BB_CUDA(cudaStreamWaitEvent(st, h->tm_events[2 * (n_chunks - 2) + 1], 0));
    return 0;
}

template <int NDET, int APPROX, bool CAL>
static int bb_launch_mb_time_marg_t(bb_handle* h, long n, double* out, cudaStream_t st) {
    const BBRelbinDev& rb = *h->rb;
    if (h->mb_nfull <= 0) return bb_fail("multi-banded time marginalisation: bb_set_multiband_time_marginalization was not called");
    const int np = rb.edges.n;
    const long S = (np + 15) / 16;
    const double dtc = h->mb_dtc, start = h->net.start_time;
    // rows of the transform any sample can use: t_j + jitter inside the prior with |jitter| <= 1 / fs (the reference's
    // jitter prior is +- 1 / fs, base.py:189-197); a sample that needs a row outside the window gets NaN, never a
    // silently truncated sum
    const double jmax = 1.0 / h->net.sampling_frequency;
    long r_lo = (long)floor((h->marg.time_min - jmax - start) / dtc) - 2, r_hi = (long)ceil((h->marg.time_max + jmax - start) / dtc) + 2;
    if (r_lo < 0) r_lo = 0;
    if (r_hi > h->mb_nfull - 1) r_hi = h->mb_nfull - 1;
    if (r_hi < r_lo) r_hi = r_lo;
    const int nrow = (int)(r_hi - r_lo + 1);
    if (!h->d_mb_E || h->mb_row0 != r_lo || h->mb_nrow != nrow) {
        cudaFree(h->d_mb_E);
        h->d_mb_E = nullptr;
        const size_t e_elems = bb_pk_elems(nrow, np, BB_GEMM_TR_B);
        BB_CUDA(cudaMalloc(&h->d_mb_E, e_elems * sizeof(double2)));
        BB_CUDA(cudaMemsetAsync(h->d_mb_E, 0, e_elems * sizeof(double2), st));
        bb_mb_phase_kernel<<<1024, 256, 0, st>>>(h->d_mb_idx, np, r_lo, nrow, h->mb_nfull, h->d_mb_E);
        BB_CUDA(cudaGetLastError());
        h->mb_row0 = r_lo;
        h->mb_nrow = nrow;
        h->launches++;
    }
    size_t chunk = (size_t)(1.0e9 / ((double)S * 16 * 1.25 * sizeof(double2)));
    chunk = chunk / 64 * 64;
    if (chunk > (size_t)n) chunk = (size_t)n;
    if (chunk < 1) chunk = 1;
    if (chunk > h->mb_chunk) {
        cudaFree(h->d_mb_V); cudaFree(h->d_mb_Y); cudaFree(h->d_mb_hh);
        h->d_mb_V = h->d_mb_Y = nullptr;
        h->d_mb_hh = nullptr;
        h->mb_chunk = 0;
        const size_t v_elems = bb_pk_elems((long)chunk, np, BB_GEMM_TR_A(true));
        BB_CUDA(cudaMalloc(&h->d_mb_V, v_elems * sizeof(double2)));
        BB_CUDA(cudaMemsetAsync(h->d_mb_V, 0, v_elems * sizeof(double2), st));
        BB_CUDA(cudaMalloc(&h->d_mb_hh, chunk * sizeof(double)));
        h->mb_chunk = chunk;
        h->mb_y_cap = 0;
    }
    if (h->mb_chunk * (size_t)nrow > h->mb_y_cap) {
        cudaFree(h->d_mb_Y);
        h->d_mb_Y = nullptr;
        h->mb_y_cap = 0;
        BB_CUDA(cudaMalloc(&h->d_mb_Y, h->mb_chunk * (size_t)nrow * sizeof(double2)));
        h->mb_y_cap = h->mb_chunk * (size_t)nrow;
    }
    const size_t smem = (size_t)BB_RED_WARPS * (BC_NCOEF + (CAL ? NDET * 4 * h->cal.n_points : 0)) * sizeof(double);
    BB_CUDA(cudaFuncSetAttribute(bb_mb_series_kernel<NDET, APPROX, CAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BBProfScope prof(h, st);
    for (long s0 = 0; s0 < n; s0 += (long)chunk) {
        const long m = (n - s0) < (long)chunk ? (n - s0) : (long)chunk;
        const long grid = bb_red_grid(h, m);
        bb_mb_series_kernel<NDET, APPROX, CAL><<<(unsigned)grid, BB_RED_THREADS, smem, st>>>(
            h->d_coef, s0, m, rb, h->d_calrec, h->cal, h->d_mb_V, h->d_mb_hh);
        BB_CUDA(cudaGetLastError());
        if (!h->mb_bands.empty()) {       // IFFT-FFT form of (h, h): odd-bin terms per detector, folded into <h|h>
            double* tmp = nullptr;
            BB_CUDA(cudaMallocAsync(&tmp, (size_t)m * NDET * sizeof(double), st));
            BB_CUDA(cudaMemsetAsync(tmp, 0, (size_t)m * NDET * sizeof(double), st));
            if (bb_mb_add_ifft_fft_terms<NDET, APPROX, CAL>(h, s0, m, tmp, 1, 0, st)) return 1;
            bb_mb_hh_fold_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(tmp, m, NDET, h->d_mb_hh);
            BB_CUDA(cudaFreeAsync(tmp, st));
        }
        BBGemmArgs ga{};
        ga.A[0] = h->d_mb_V;
        ga.B[0] = h->d_mb_E;
        ga.C = reinterpret_cast<double*>(h->d_mb_Y);
        ga.slabs_a = ga.slabs_b = S;
        ga.slab0 = 0; ga.n_slabs = (int)S;
        ga.ldc = nrow;
        ga.M = (int)m; ga.N = nrow; ga.n_seg = 1; ga.n_batch = 1; ga.accumulate = 0; ga.alpha = 1.0;
        if (bb_gemm_nt(true, ga, h->sm_count, st)) return 1;
        const long grid_e = (m + 3) / 4 < 8L * h->sm_count ? (m + 3) / 4 : 8L * h->sm_count;
        bb_mb_time_marg_kernel<<<(unsigned)grid_e, 128, 0, st>>>(h->d_coef, s0, m, h->d_mb_Y, r_lo, nrow, h->mb_nfull, dtc,
                                                                  h->d_mb_hh, h->marg, start, out);
        BB_CUDA(cudaGetLastError());
        h->launches += 3;
    }
    return 0;
}

template <int NDET, int APPROX, bool CAL>
static int bb_launch_relbin_time_marg_t(bb_handle* h, long n, double* out, cudaStream_t st) {
    const BBRelbinDev& rb = *h->rb;
    if (!rb.pgrid) return bb_fail("relative binning time marginalisation: bb_set_relative_binning was called without "
                                  "the full-grid fiducial waveforms");
    if (h->nfft == 0) return bb_fail("time marginalisation needs n_freq - 1 to be a power of two");
    const int nfft = h->nfft, nb = rb.edges.n - 1;
    int log2n = 0;
    while ((1 << log2n) < nfft) ++log2n;
    int plan_a, plan_b, ps;
    bb_tm_plan(log2n, &plan_a, &plan_b, &ps);
    const size_t smem = (bb_tm_series_elems(nfft, ps) + (size_t)nb * NDET * 2) * sizeof(double2)
                        + (BC_NCOEF + 33 + (CAL ? NDET * 4 * h->cal.n_points : 0)) * sizeof(double);
    if (smem > 227 * 1024) return bb_fail("relative binning time marginalisation: series + bins do not fit shared memory");
    BB_CUDA(cudaFuncSetAttribute(bb_relbin_time_marg_kernel<NDET, APPROX, CAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
    long grid = (long)h->sm_count * per_sm;
    if (grid > n) grid = n;
    {
        BBProfScope prof(h, st);
        bb_relbin_time_marg_kernel<NDET, APPROX, CAL><<<(unsigned)grid, BB_TM_THREADS, smem, st>>>(
            h->d_coef, n, rb, h->net.n_freq, h->net.df, nfft, log2n, h->d_twiddle, h->marg, h->net.start_time,
            h->net.duration, h->d_calrec, h->cal, out);
    }
    h->launches++;
    BB_CUDA(cudaGetLastError());
    return 0;
}

// what: 0 inner products (K5 / K6), 1 time-marginalised likelihood
template <int NDET>
static int bb_launch_reduced_n(bb_handle* h, long n, double* out, cudaStream_t st, int what) {
    const bool pd = h->wf.approximant == BB_IMRPHENOMD;
    const bool cal = h->cal_params != nullptr;
#define BB_RED_DISPATCH(FN)                                                                              \
    do {                                                                                                 \
        if (cal) return pd ? FN<NDET, BB_IMRPHENOMD, true>(h, n, out, st) : FN<NDET, BB_TAYLORF2, true>(h, n, out, st); \
        return pd ? FN<NDET, BB_IMRPHENOMD, false>(h, n, out, st) : FN<NDET, BB_TAYLORF2, false>(h, n, out, st);        \
    } while (0)
    if (what == 0) BB_RED_DISPATCH(bb_launch_reduced_t);
    if (h->kind == 1 && !h->rb->cross_g) BB_RED_DISPATCH(bb_launch_mb_time_marg_t);
    if (h->kind == 1) BB_RED_DISPATCH(bb_launch_relbin_time_marg_t);
    BB_RED_DISPATCH(bb_launch_roq_time_marg_t);
#undef BB_RED_DISPATCH
}

static int bb_launch_reduced(bb_handle* h, long n, double* out, cudaStream_t st, int what) {
    if (h->shard_lo != 0 || h->shard_hi != h->net.n_freq)
        return bb_fail("reduced-order likelihoods are sample-sharded only (bb_set_frequency_shard is active)");
    switch (h->net.n_det) {
        case 1: return bb_launch_reduced_n<1>(h, n, out, st, what);
        case 2: return bb_launch_reduced_n<2>(h, n, out, st, what);
        case 3: return bb_launch_reduced_n<3>(h, n, out, st, what);
        case 4: return bb_launch_reduced_n<4>(h, n, out, st, what);
    }
    return bb_fail("bad n_det");
}
