// FP64 tensor-core GEMM for the dense contractions of the path (included by bb_kernels.cu) - replaces cuBLAS
// ZGEMM / DGEMM in
//   K7   ROQ all-times contraction  W conj(h_lin)            bilby/gw/likelihood/roq.py:604-651
//   calibration marginalisation     [samples x bins] x [bins x curves]   bilby/gw/likelihood/base.py:305-346, 860-877
//   ROQ linear-weight build         phase matrix x (d/S conj(basis))     bilby/gw/likelihood/roq.py:849-918
//
// One form covers all three:   C[m][n] (+)= alpha * sum_seg sum_k A_seg[m][k] * B_seg[n][k]
// with A [M x K] and B [N x K] (both contracted along their rows, operands in the packed layout below), C [M x N]
// row-major, complex (interleaved re, im) or real doubles, an optional batch dimension (detectors) and up to
// BB_GEMM_MAX_SEG K-segments (detectors concatenated along the contraction axis).
//
// Blackwell has no tcgen05 kind for FP64: the FP64 tensor path is mma.sync.aligned.m8n8k4.f64 (SASS DMMA.884), measured
// here at 37.1 TFLOP/s (tools/micro/dmma_peak.cu) = the nominal 148 SM x 64 FMA/clk.  One DMMA occupies a
// sub-partition's FP64 pipe for 16 cycles, so the issue slots are nearly free and the kernel is built to keep two
// independent DMMAs in flight per sub-partition:
//   * persistent CTAs of 4 warps, TWO per SM, no producer warp: the warp that is last to finish a stage refills it.
//     Two independent CTAs per SM because the tile epilogue (stores from the fragments) and the slab hand-overs leave
//     the tensor pipe idle for ~12k cycles per tile when all warps of an SM belong to one CTA (K = 256: 32.9 TFLOP/s
//     with one 8-warp CTA, tools/gemm_shapes.py); with two CTAs one computes while the other stores;
//   * CTA tile 64 x 64 complex (warp tile 32 x 32 = 4 x 4 fragments, 4 DMMAs per fragment pair and k-step:
//     re += ar br, re += (-ai) bi, im += ar bi, im += ai br: 128 accumulator registers) or 128 x 64 real (warp tile
//     64 x 32 = 8 x 4 fragments);
//   * K slabs of 16 staged through a 2-deep shared-memory ring (40 KB per stage) by two bulk copies per stage
//     (cp.async.bulk -> SASS UBLKCP, completion on the stage's mbarrier, SYNCS) of operands that are stored PACKED
//     (bb_pk below); rows are padded to a pitch = 64 (complex) / 32 (real) mod 128 bytes so that the fragment loads
//     (LDS.128 / LDS.64) are bank-conflict free;
//   * epilogue straight from the accumulator fragments: each quad of lanes writes 128 contiguous bytes of a C row.
// Rows beyond M / N of the last tiles only reach accumulators that are never stored; the K tail of the last slab is
// zero in both operands.
#pragma once

#define BB_GEMM_MAX_SEG 4
#define BB_GEMM_BK 16
#define BB_GEMM_CONSUMERS 4
#define BB_GEMM_THREADS (BB_GEMM_CONSUMERS * 32)
#define BB_PK 20          // packed row pitch in ELEMENTS: 16 of a K slab + 4 of padding (320 B complex, 160 B real)

// ---- packed operand layout.  A bulk copy costs the TMA unit ~55 cycles however small it is (tools/micro/
// bulk_copy_rate.cu: 256-byte rows stream at 4.6 B/clk/SM, >= 20 KB pieces at 50 B/clk/SM), so a tile cannot be
// fetched row by row from a row-major matrix.  Every operand is therefore WRITTEN by its producer (the kernels of
// this library, or the set-up upload) in the shape the GEMM reads it: row tiles of TR rows x K slabs of 16, each
// (tile, slab) one contiguous piece of TR rows of BB_PK elements - the shared-memory image including the bank
// padding - so a stage is filled by TWO bulk copies (20 - 40 KB each).  Element (r, k) of a matrix with S slabs:
__host__ __device__ __forceinline__ size_t bb_pk(long r, long k, int TR, long S) {
    return ((size_t)((r / TR) * S + (k >> 4)) * TR + (size_t)(r % TR)) * BB_PK + (size_t)(k & 15);
}
static inline size_t bb_pk_elems(long rows, long K, int TR) {
    return (size_t)((rows + TR - 1) / TR) * (size_t)((K + 15) / 16) * TR * BB_PK;
}

struct BBGemmArgs {
    const void* A[BB_GEMM_MAX_SEG];     // packed operands (batch 0) of every K segment: A tiles of BM rows,
    const void* B[BB_GEMM_MAX_SEG];     // B tiles of BN rows
    long slabs_a, slabs_b;              // S of the packed A / B matrices
    int slab0, n_slabs;                 // contraction window: slabs [slab0, slab0 + n_slabs) of every segment
    long batch_a, batch_b, batch_c;     // element strides between batches
    double* C;                          // row-major [M][ldc] (complex: interleaved re, im)
    long ldc;
    int M, N, n_seg, n_batch;
    int accumulate;                     // 0: C = alpha A B^T, 1: C += alpha A B^T
    double alpha;
    int skew_from;                      // CTAs with blockIdx >= skew_from start half a tile late (set by bb_gemm_nt)
};

template <bool CPLX>
struct BBGemmCfg {
    static constexpr int ELT = CPLX ? 16 : 8;                 // bytes per element
    static constexpr int FM = CPLX ? 4 : 8, FN = 4;             // fragments (8 x 8) per warp tile (real 8 x 8 spills)
    static constexpr int WM = 2, WN = 2;                      // warps per CTA tile
    static constexpr int BM = WM * FM * 8, BN = WN * FN * 8;  // 64 x 64 complex, 128 x 64 real
    static constexpr int STRIDE = BB_PK * ELT;                // 320 / 160 bytes: = 64 / 32 mod 128 (conflict-free fragments)
    static constexpr int STAGE_BYTES = (BM + BN) * STRIDE;    // 40960 complex, 30720 real
    // ring depth: a real slab holds half the DMMAs of a complex one (4096 cycles of pipe time), too short to cover the
    // refill latency with two stages
    static constexpr int STAGES = CPLX ? 2 : 3;
    static constexpr int SMEM = STAGES * STAGE_BYTES + STAGES * 16 + 128;
};
#define BB_GEMM_TR_A(cplx) ((cplx) ? 64 : 128)
#define BB_GEMM_TR_B_OF(cplx) 64
#define BB_GEMM_TR_B 64          // complex operands (the real one is only used by the calibration marginalisation)

__device__ __forceinline__ void bb_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void bb_mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bb_smem_u32(bar)) : "memory");
}

template <bool CPLX>
__global__ void __launch_bounds__(BB_GEMM_THREADS, 2) bb_gemm_nt_kernel(BBGemmArgs g) {
    using Cfg = BBGemmCfg<CPLX>;
    extern __shared__ __align__(128) unsigned char gemm_smem[];
    unsigned char* stages = gemm_smem;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(gemm_smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    int* done = reinterpret_cast<int*>(full + Cfg::STAGES);          // warps finished with each stage
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) {
            bb_mbar_init(&full[s], 1);
            done[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // The two CTAs of an SM start together and run at the same rate, so their epilogues would coincide and leave the
    // tensor pipe idle: the second wave of CTAs starts half a tile late (and stays half a tile out of step).
    if ((int)blockIdx.x >= g.skew_from) {
        // half a tile, but no more than a few epilogues' worth (long-K tiles dwarf their epilogue anyway)
        const long long half = (long long)g.n_slabs * g.n_seg * 4096;
        const long long t0 = clock64(), wait = half < 40000 ? half : 40000;
        while (clock64() - t0 < wait) __nanosleep(1000);
    }
    const int tiles_m = (g.M + Cfg::BM - 1) / Cfg::BM, tiles_n = (g.N + Cfg::BN - 1) / Cfg::BN;
    const long tiles_mn = (long)tiles_m * tiles_n;
    const long n_tiles = tiles_mn * g.n_batch;
    const int n_slabs = g.n_slabs * g.n_seg;

    // ---- producer side.  There is no producer warp: the slabs of this CTA's tiles form one stream; every warp tracks
    // the cursor of the next slab to fetch, and the warp that is LAST to finish reading a stage refills it (its 32 lanes
    // issue one bulk copy per tile row), so nobody ever waits for a free stage.
    long p_tile = blockIdx.x, p_it = 0;
    int p_sl = 0;
    auto issue = [&]() {
        if (p_tile >= n_tiles) return;
        const int b = (int)(p_tile / tiles_mn);
        const long r = p_tile - (long)b * tiles_mn;
        const int tm = (int)(r / tiles_n), tn = (int)(r - (long)tm * tiles_n);       // n fastest: the A tile is reused from L2
        const int seg = p_sl / g.n_slabs, ks = g.slab0 + (p_sl - seg * g.n_slabs);
        const int stage = (int)(p_it % Cfg::STAGES);
        unsigned char* sa = stages + (size_t)stage * Cfg::STAGE_BYTES;
        unsigned char* sb = sa + Cfg::BM * Cfg::STRIDE;
        if (lane == 0) {
            bb_mbar_expect_tx(&full[stage], (unsigned)Cfg::STAGE_BYTES);
            const char* ga = reinterpret_cast<const char*>(g.A[seg])
                             + ((size_t)b * g.batch_a + (size_t)(tm * g.slabs_a + ks) * Cfg::BM * BB_PK) * Cfg::ELT;
            const char* gb = reinterpret_cast<const char*>(g.B[seg])
                             + ((size_t)b * g.batch_b + (size_t)(tn * g.slabs_b + ks) * Cfg::BN * BB_PK) * Cfg::ELT;
            bb_bulk_g2s(sa, ga, (unsigned)(Cfg::BM * Cfg::STRIDE), &full[stage]);
            bb_bulk_g2s(sb, gb, (unsigned)(Cfg::BN * Cfg::STRIDE), &full[stage]);
        }
        __syncwarp();
    };
    auto advance = [&]() {
        ++p_it;
        if (++p_sl == n_slabs) { p_sl = 0; p_tile += gridDim.x; }
    };
    for (int s = 0; s < Cfg::STAGES; ++s) {
        if (warp == 0) issue();
        advance();
    }

    // ---------------- consumer warps
    const int wm = warp / Cfg::WN, wn = warp - wm * Cfg::WN;
    const int gq = lane >> 2, tq = lane & 3;           // fragment row / k index of this lane
    long it = 0;
    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = (int)(tile / ((long)tiles_m * tiles_n));
        const long r = tile - (long)b * tiles_m * tiles_n;
        const int tm = (int)(r / tiles_n), tn = (int)(r - (long)tm * tiles_n);
        const int m0 = tm * Cfg::BM, n0 = tn * Cfg::BN;
        double acc[Cfg::FM][Cfg::FN][CPLX ? 4 : 2];
#pragma unroll
        for (int i = 0; i < Cfg::FM; ++i)
#pragma unroll
            for (int j = 0; j < Cfg::FN; ++j)
#pragma unroll
                for (int c = 0; c < (CPLX ? 4 : 2); ++c) acc[i][j][c] = 0.0;

        for (int sl = 0; sl < n_slabs; ++sl, ++it) {
            const int stage = (int)(it % Cfg::STAGES);
            const unsigned par = (unsigned)((it / Cfg::STAGES) & 1);
            bb_mbar_wait(&full[stage], par);
            const unsigned char* sa = stages + (size_t)stage * Cfg::STAGE_BYTES
                                      + (size_t)(wm * Cfg::FM * 8 + gq) * Cfg::STRIDE + tq * Cfg::ELT;
            const unsigned char* sb = stages + (size_t)stage * Cfg::STAGE_BYTES + Cfg::BM * Cfg::STRIDE
                                      + (size_t)(wn * Cfg::FN * 8 + gq) * Cfg::STRIDE + tq * Cfg::ELT;
            // (fetching the fragments of k-step ks + 1 under the DMMAs of k-step ks - a second register set - measured the
            // same: the other CTA's warp on the sub-partition already covers the LDS latency; 186 instead of 230 registers
            // leave room for a co-resident epilogue kernel)
#pragma unroll 1
            for (int ks = 0; ks < BB_GEMM_BK / 4; ++ks) {
                if (CPLX) {
                    double2 a[Cfg::FM], bb[Cfg::FN];
#pragma unroll
                    for (int i = 0; i < Cfg::FM; ++i)
                        a[i] = *reinterpret_cast<const double2*>(sa + i * 8 * Cfg::STRIDE + ks * 4 * Cfg::ELT);
#pragma unroll
                    for (int j = 0; j < Cfg::FN; ++j)
                        bb[j] = *reinterpret_cast<const double2*>(sb + j * 8 * Cfg::STRIDE + ks * 4 * Cfg::ELT);
                    // 32 independent DMMAs, then the 32 that depend on them
#pragma unroll
                    for (int i = 0; i < Cfg::FM; ++i)
#pragma unroll
                        for (int j = 0; j < Cfg::FN; ++j) {
                            bb_dmma(acc[i][j][0], acc[i][j][1], a[i].x, bb[j].x);
                            bb_dmma(acc[i][j][2], acc[i][j][3], a[i].x, bb[j].y);
                        }
#pragma unroll
                    for (int i = 0; i < Cfg::FM; ++i) {
                        const double nai = -a[i].y;
#pragma unroll
                        for (int j = 0; j < Cfg::FN; ++j) {
                            bb_dmma(acc[i][j][0], acc[i][j][1], nai, bb[j].y);
                            bb_dmma(acc[i][j][2], acc[i][j][3], a[i].y, bb[j].x);
                        }
                    }
                } else {
                    double a[Cfg::FM], bb[Cfg::FN];
#pragma unroll
                    for (int i = 0; i < Cfg::FM; ++i)
                        a[i] = *reinterpret_cast<const double*>(sa + i * 8 * Cfg::STRIDE + ks * 4 * Cfg::ELT);
#pragma unroll
                    for (int j = 0; j < Cfg::FN; ++j)
                        bb[j] = *reinterpret_cast<const double*>(sb + j * 8 * Cfg::STRIDE + ks * 4 * Cfg::ELT);
#pragma unroll
                    for (int i = 0; i < Cfg::FM; ++i)
#pragma unroll
                        for (int j = 0; j < Cfg::FN; ++j) bb_dmma(acc[i][j][0], acc[i][j][1], a[i], bb[j]);
                }
            }
            // this warp is done reading the stage; the last of the 4 warps refills it with the slab STAGES ahead
            __syncwarp();
            int prev = 0;
            if (lane == 0) prev = atomicAdd(&done[stage], 1);
            prev = __shfl_sync(0xffffffffu, prev, 0);
            if (prev == BB_GEMM_CONSUMERS - 1) {
                if (lane == 0) done[stage] = 0;
                issue();
            }
            advance();
        }

        // ---- epilogue: fragment (i, j): rows m0 + wm*FM*8 + i*8 + gq, columns n0 + wn*FN*8 + j*8 + 2 tq + {0, 1}
        double* cbase = g.C + (size_t)b * g.batch_c * (CPLX ? 2 : 1);
        if (!g.accumulate && g.alpha == 1.0 && m0 + Cfg::BM <= g.M && n0 + Cfg::BN <= g.N) {
            // interior tile: no bounds checks, constant offsets from one row pointer
            constexpr int W = CPLX ? 2 : 1;
            double* p0 = cbase + ((size_t)(m0 + wm * Cfg::FM * 8 + gq) * g.ldc + (n0 + wn * Cfg::FN * 8 + 2 * tq)) * W;
            const size_t row8 = (size_t)8 * g.ldc * W;
#pragma unroll
            for (int i = 0; i < Cfg::FM; ++i) {
                double* p = p0 + i * row8;
#pragma unroll
                for (int j = 0; j < Cfg::FN; ++j) {
                    if (CPLX) {
                        *reinterpret_cast<double2*>(p + j * 16) = make_double2(acc[i][j][0], acc[i][j][2]);
                        *reinterpret_cast<double2*>(p + j * 16 + 2) = make_double2(acc[i][j][1], acc[i][j][3]);
                    } else {
                        p[j * 8] = acc[i][j][0];           // (16-byte alignment of a real row is not guaranteed)
                        p[j * 8 + 1] = acc[i][j][1];
                    }
                }
            }
            continue;
        }
#pragma unroll
        for (int i = 0; i < Cfg::FM; ++i) {
            const int m = m0 + (wm * Cfg::FM + i) * 8 + gq;
            if (m >= g.M) continue;
#pragma unroll
            for (int j = 0; j < Cfg::FN; ++j) {
                const int n = n0 + (wn * Cfg::FN + j) * 8 + 2 * tq;
                if (n >= g.N) continue;
                if (CPLX) {
                    double* p = cbase + ((size_t)m * g.ldc + n) * 2;
                    double v0 = g.alpha * acc[i][j][0], v1 = g.alpha * acc[i][j][2];     // element n: re, im
                    double v2 = g.alpha * acc[i][j][1], v3 = g.alpha * acc[i][j][3];     // element n + 1
                    const bool two = n + 1 < g.N;
                    if (g.accumulate) {
                        v0 += p[0]; v1 += p[1];
                        if (two) { v2 += p[2]; v3 += p[3]; }
                    }
                    *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
                    if (two) *reinterpret_cast<double2*>(p + 2) = make_double2(v2, v3);
                } else {
                    double* p = cbase + (size_t)m * g.ldc + n;
                    double v0 = g.alpha * acc[i][j][0], v1 = g.alpha * acc[i][j][1];
                    const bool two = n + 1 < g.N;
                    if (g.accumulate) {
                        v0 += p[0];
                        if (two) v1 += p[1];
                    }
                    if (two && ((reinterpret_cast<size_t>(p) & 15) == 0)) *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
                    else {
                        p[0] = v0;
                        if (two) p[1] = v1;
                    }
                }
            }
        }
    }
}

// Host launcher (operands packed with bb_pk: A in tiles of BB_GEMM_TR_A rows, B in tiles of BB_GEMM_TR_B rows).
static int bb_gemm_nt(bool cplx, const BBGemmArgs& g, int sm_count, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0 || g.n_slabs <= 0) return 0;
    if (g.n_seg < 1 || g.n_seg > BB_GEMM_MAX_SEG || g.n_batch < 1) return bb_fail("bb_gemm_nt: bad segment / batch count");
    if (g.slab0 < 0 || g.slab0 + g.n_slabs > g.slabs_a || g.slab0 + g.n_slabs > g.slabs_b)
        return bb_fail("bb_gemm_nt: contraction window outside the packed operands");
    for (int s = 0; s < g.n_seg; ++s)
        if ((reinterpret_cast<size_t>(g.A[s]) | reinterpret_cast<size_t>(g.B[s])) & 15)
            return bb_fail("bb_gemm_nt: operands must be 16-byte aligned");
    const int smem = cplx ? BBGemmCfg<true>::SMEM : BBGemmCfg<false>::SMEM;
    if (cplx) BB_CUDA(cudaFuncSetAttribute(bb_gemm_nt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    else BB_CUDA(cudaFuncSetAttribute(bb_gemm_nt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int bm = cplx ? BBGemmCfg<true>::BM : BBGemmCfg<false>::BM, bn = cplx ? BBGemmCfg<true>::BN : BBGemmCfg<false>::BN;
    long tiles = (long)((g.M + bm - 1) / bm) * ((g.N + bn - 1) / bn) * g.n_batch;
    const unsigned grid = (unsigned)(tiles < 2L * sm_count ? tiles : 2L * sm_count);       // two CTAs per SM
    BBGemmArgs ga = g;
    ga.skew_from = (tiles >= 4L * sm_count) ? sm_count : (int)grid;      // only when every CTA has several tiles
    if (cplx) bb_gemm_nt_kernel<true><<<grid, BB_GEMM_THREADS, smem, st>>>(ga);
    else bb_gemm_nt_kernel<false><<<grid, BB_GEMM_THREADS, smem, st>>>(ga);
    BB_CUDA(cudaGetLastError());
    return 0;
}

// row-major [rows][ld] (device) -> packed (bb_pk) with zero padding: for callers whose operands are not produced by
// this library's kernels (bb_contract_device)
template <typename T>
__global__ void bb_gemm_pack_kernel(const T* __restrict__ src, long rows, long K, long ld, int TR, T* __restrict__ dst) {
    const long S = (K + 15) / 16, rows_p = (rows + TR - 1) / TR * TR;
    const long total = rows_p * S * 16;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / (S * 16), k = i - r * (S * 16);
        T v{};
        if (r < rows && k < K) v = src[r * ld + k];
        dst[bb_pk(r, k, TR, S)] = v;
    }
}
