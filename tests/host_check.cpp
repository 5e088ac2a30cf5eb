// TEST INFRASTRUCTURE ONLY.  Compiles the library's host+device inline math (bilby_b200/csrc/*.cuh)
// with g++ so the per-sample prologue, the per-bin amplitude/phase evaluation and the epilogue special
// functions can be checked against the oracle in the GPU-less build container
// (tests/test_host_math.py).  The shipped library contains no CPU execution path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define BB_CONST_QUAL static const
#include "../bilby_b200/csrc/bb_common.cuh"
#include "../bilby_b200/csrc/bb_geometry.cuh"
#include "../bilby_b200/csrc/bb_phenomd.cuh"
#ifdef BB_HAVE_TAYLORF2
#include "../bilby_b200/csrc/bb_taylorf2.cuh"
#endif
#include "../bilby_b200/csrc/bb_special.cuh"
#include "../bilby_b200/csrc/qnm_table.inc"
#include "../bilby_b200/csrc/phenomd_fit.inc"

static std::vector<double> read_all(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<double> v(sz / 8);
    if (fread(v.data(), 8, v.size(), f) != v.size()) exit(3);
    fclose(f);
    return v;
}
static void write_all(const char* path, const std::vector<double>& v) {
    FILE* f = fopen(path, "wb");
    fwrite(v.data(), 8, v.size(), f);
    fclose(f);
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: host_check MODE in.bin out.bin\n"); return 1; }
    std::vector<double> in = read_all(argv[2]);
    std::vector<double> out;
    if (!strcmp(argv[1], "wave")) {
        // in: n, n_det, n_freq, duration, fs, start_time, approximant, f_ref, f_min, f_max, k_lo, k_hi,
        //     tensors[n_det*9], vertices[n_det*3], params[n*16]
        size_t o = 0;
        const long n = (long)in[o++];
        BBNetwork net;
        memset(&net, 0, sizeof(net));
        net.n_det = (int)in[o++];
        net.n_freq = (int)in[o++];
        net.duration = in[o++];
        net.sampling_frequency = in[o++];
        net.start_time = in[o++];
        BBWaveformConfig wf;
        memset(&wf, 0, sizeof(wf));
        wf.approximant = (int)in[o++];
        wf.add_jitter = 0;
        wf.f_ref = in[o++];
        wf.f_min = in[o++];
        wf.f_max = in[o++];
        net.k_lo = (int)in[o++];
        net.k_hi = (int)in[o++];
        net.df = (net.sampling_frequency / 2) / (double)(net.n_freq - 1);
        for (int d = 0; d < net.n_det; ++d) for (int i = 0; i < 9; ++i) net.detector_tensor[d][i] = in[o++];
        for (int d = 0; d < net.n_det; ++d) for (int i = 0; i < 3; ++i) net.vertex[d][i] = in[o++];
        BBQnmTable qnm = {bb_qnm_x, bb_qnm_fring, bb_qnm_fring_d2, bb_qnm_fdamp, bb_qnm_fdamp_d2, BB_QNM_N};
        // out: layout header [BC_NCOEF, BC_DET, BC_DSTRIDE, BC_KMIN], then per sample coef[BC_NCOEF] and (A, Phi) per bin
        out.assign(4 + (size_t)n * (BC_NCOEF + 2 * (size_t)net.n_freq), 0.0);
        out[0] = BC_NCOEF; out[1] = BC_DET; out[2] = BC_DSTRIDE; out[3] = BC_KMIN;
        for (long s = 0; s < n; ++s) {
            double* c = &out[4 + (size_t)s * (BC_NCOEF + 2 * (size_t)net.n_freq)];
            const double* p = &in[o + s * BB_NPARAM];
            if (wf.approximant == 0) bb_phenomd_prologue(p, net, wf, qnm, bb_phenomd_fit, c);
#ifdef BB_HAVE_TAYLORF2
            else bb_taylorf2_prologue(p, net, wf, c);
#endif
            double* ap = c + BC_NCOEF;
            const int k0 = (int)c[BC_KMIN], k1 = (int)c[BC_KMAX];
            for (int k = k0; k < k1; ++k) {
                const double f = (double)k * net.df;
                const double u = pow(f, -1.0 / 6.0), t = u * u, x = f * t * t;
                if (wf.approximant == 0) {
                    ap[2 * k] = bb_phenomd_amp(c, f, u, t, x);
                    ap[2 * k + 1] = bb_phenomd_phase(c, f, t, x, log(f), pow(f, 0.75));
                }
#ifdef BB_HAVE_TAYLORF2
                else {
                    ap[2 * k] = bb_taylorf2_amp(c, u, t);
                    ap[2 * k + 1] = bb_taylorf2_phase(c, f, t, x, log(f));
                }
#endif
            }
        }
    } else if (!strcmp(argv[1], "lni0")) {
        out.resize(in.size());
        for (size_t i = 0; i < in.size(); ++i) out[i] = bb_ln_i0(in[i], bb_i0e_a, bb_i0e_b);
    } else if (!strcmp(argv[1], "bispev")) {
        // in: nx, ny, nq, xmin, xmax, ymin, ymax, tx[nx], ty[ny], c[(nx-4)(ny-4)], xq[nq], yq[nq]
        size_t o = 0;
        BBSpline2D s;
        s.nx = (int)in[o++];
        s.ny = (int)in[o++];
        const long nq = (long)in[o++];
        s.xmin = in[o++]; s.xmax = in[o++]; s.ymin = in[o++]; s.ymax = in[o++];
        s.tx = &in[o]; o += s.nx;
        s.ty = &in[o]; o += s.ny;
        s.c = &in[o]; o += (size_t)(s.nx - 4) * (s.ny - 4);
        const double* xq = &in[o];
        const double* yq = xq + nq;
        out.resize(nq);
        for (long i = 0; i < nq; ++i) out[i] = bb_bispev(s, xq[i], yq[i]);
    } else if (!strcmp(argv[1], "gmst")) {
        out.resize(in.size());
        for (size_t i = 0; i < in.size(); ++i) out[i] = bb_gmst(in[i]);
    } else {
        return 1;
    }
    write_all(argv[3], out);
    return 0;
}
