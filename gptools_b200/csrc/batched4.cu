// Batched ll + gradient: FOUR independent theta streams per SM.
//
// One persistent CTA per theta: phase 1 left-looking Cholesky with generated K tiles and inverted diagonal tiles,
// back substitution, phase 2 XT = L^{-T} in place, phase 3 K^{-1} tiles contracted with regenerated dK tiles
// (tools/tile_model.py is the numpy model of the tile recurrences).  Cut for latency tolerance: measurements on
// the first-generation kernel (256 threads, 2 CTAs/SM; git history, profiles/r01a_*): the FP64 pipe -- which DMMA and scalar DFMA share on B200 (profiles/microbench/fp64_overlap.cu)
// -- was only ~63% busy because each CTA spends half of its time in latency-bound non-GEMM phases and only two
// CTAs (= two thetas) fit on an SM.  Here a CTA is 4 warps / 128 threads with a ~39 KB footprint, so four CTAs
// (four thetas) share an SM and the chance that nobody feeds the tensor pipe drops from ~29% to ~8%:
//   * one 64x64 output tile per job, 32x32 warp tiles (16 DMMA.8x8x4 per k-step, 64 accumulator registers);
//   * operands stream through a double-buffered cp.async ring of 64x16 chunks (A, B) with an XOR swizzle;
//   * no dedicated diagonal-tile buffer: the Gauss-Jordan sweep runs in registers (32 entries per thread) and
//     its result goes straight to the workspace; panel products take their B fragments directly from L2.
// Algorithmic work: M^3 flop per theta; roofline = FP64 tensor pipe.
#include <stdio.h>
#include <stdlib.h>

#include "se_fast.cuh"

namespace {

using namespace sefast;

constexpr int TB = 64;
constexpr int BK = 16;
#ifndef GPT_B4_STAGES
#define GPT_B4_STAGES 2  // 2 stages = 39 KB per CTA: the smaller carve-out leaves ~90 KB of L1 per SM (measured +3% over 3)
#endif
#ifndef GPT_B4_MINB
#define GPT_B4_MINB 4
#endif
constexpr int STAGES = GPT_B4_STAGES;
constexpr int THREADS = 128;
constexpr int CHUNK = TB * BK;         // doubles per operand chunk
constexpr int STAGE_D = 2 * CHUNK;     // A, B
constexpr int LDT = 68;
constexpr int ST_D = TB * LDT;         // 4352
// ring (2 stages: 4096 doubles = 32 KB); the same storage is the staging tile (64 x 68) + staged row points
constexpr int R_D = (STAGES * STAGE_D > ST_D + 256) ? STAGES * STAGE_D : ST_D + 256;
constexpr int MAXT = 32;
constexpr int TILE = TB * TB;
constexpr int PTS_OFF = ST_D;          // staged rows behind the staging tile: x[64][FD] (<= 128), alpha[64], orders
constexpr int PTS_ALPHA = 128;
constexpr int PTS_ORD = 192;

struct Smem {
    double R[R_D];
    double piv[8];  // product of the 8 pivots of each 8x8 pivot block of the tile being factored
    double exptab[64];  // 2^(j/64), for exp_nonpos_tab
    double rk[TB], zk[TB];
    double red[4][GPT_MAX_PARAMS + 2];
    const double* a[MAXT];
    const double* b[MAXT];
    unsigned char flag[MAXT];  // bit0: A operand upper triangular, bit2: B operand upper triangular
    int skip_upper;            // output is a diagonal tile: its upper-right 32x32 block is not needed
    CovParams cp;
    double noise2;
    int theta;
    int info;
};

__device__ __forceinline__ double* slot(double* ws, int I, int J) { return ws + (size_t)(I * (I + 1) / 2 + J) * TILE; }
__device__ __forceinline__ double* slotDT(double* ws, int nT, int I) {
    return ws + (size_t)(nT * (nT + 1) / 2 + I) * TILE;
}

struct Lane {
    int tid, warp, lane, g, t, wr, wc;
};

__device__ __forceinline__ void issue_chunk(Smem& sm, const Lane& L, int q) {
    const int s = q >> 2, kc = q & 3;
    double* base = sm.R + (q % STAGES) * STAGE_D;
    const double* srcs[2] = {sm.a[s], sm.b[s]};
#pragma unroll
    for (int op = 0; op < 2; op++) {
        const double* src = srcs[op];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int idx = L.tid + u * THREADS;
            const int r = idx >> 3, c2 = (idx & 7) * 2;
            cp_async16(base + op * CHUNK + r * BK + (c2 ^ ((r & 3) << 2)), src + r * TB + kc * BK + c2);
        }
    }
}

// acc += A[s] * B[s]^T over the steps of the job table
__device__ __forceinline__ void run_job(Smem& sm, const Lane& L, int nsteps, double (&acc)[4][4][2]) {
    const int total = nsteps * 4;
    __syncthreads();  // table visible, ring free
#pragma unroll
    for (int q = 0; q < STAGES - 1; q++) {
        if (q < total) issue_chunk(sm, L, q);
        cp_async_commit();
    }
    const bool dead = sm.skip_upper && L.wr == 0 && L.wc == 1;
    for (int q = 0; q < total; q++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (q + STAGES - 1 < total) issue_chunk(sm, L, q + STAGES - 1);
        cp_async_commit();
        const int fl = sm.flag[q >> 2];
        const bool half_zero = ((fl & 1) && L.wr == 1) || ((fl & 4) && L.wc == 1);
        if (!dead && !(half_zero && (q & 3) < 2)) {
            const double* aS = sm.R + (q % STAGES) * STAGE_D + (L.wr * 32 + L.g) * BK;
            const double* bS = sm.R + (q % STAGES) * STAGE_D + CHUNK + (L.wc * 32 + L.g) * BK;
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                const int col = ((kk ^ (L.g & 3)) << 2) + L.t;
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; i++) a[i] = aS[i * 8 * BK + col];
#pragma unroll
                for (int j = 0; j < 4; j++) b[j] = bS[j * 8 * BK + col];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();  // ring may now be reused as staging
}

__device__ __forceinline__ void zero_acc(double (&acc)[4][4][2]) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
}

__device__ __forceinline__ void acc_to_tile(double* tile, const Lane& L, const double (&acc)[4][4][2]) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
            *reinterpret_cast<double2*>(tile + (L.wr * 32 + i * 8 + L.g) * LDT + L.wc * 32 + j * 8 + 2 * L.t) = v;
        }
}

__device__ __forceinline__ void acc_to_global(double* tile, const Lane& L, const double (&acc)[4][4][2], double scale) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double2 v = make_double2(scale * acc[i][j][0], scale * acc[i][j][1]);
            *reinterpret_cast<double2*>(tile + (L.wr * 32 + i * 8 + L.g) * TB + L.wc * 32 + j * 8 + 2 * L.t) = v;
        }
}

// out = St * Binv^T, St in shared (stride LDT), Binv a lower-triangular 64x64 tile in GLOBAL memory (row-major):
// the B fragments are fetched straight from L2, four k-steps per batch, so no shared buffer is needed for them.
__device__ __forceinline__ void mult_lower_global(const double* St, const double* Binv, const Lane& L,
                                                  double (&out)[4][4][2]) {
    zero_acc(out);
    const int kmax = (L.wc + 1) * 32;  // Binv[n][c] = 0 for c > n
    const double* aS = St + (L.wr * 32 + L.g) * LDT + L.t;
    const double* bG = Binv + (size_t)(L.wc * 32 + L.g) * TB + L.t;
#pragma unroll 1
    for (int kb = 0; kb < kmax; kb += 16) {
        double b[4][4];
#pragma unroll
        for (int kk = 0; kk < 4; kk++)
#pragma unroll
            for (int j = 0; j < 4; j++) b[kk][j] = bG[(size_t)j * 8 * TB + kb + kk * 4];
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            double a[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = aS[i * 8 * LDT + kb + kk * 4];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(out[i][j][0], out[i][j][1], a[i], b[kk][j]);
        }
    }
}

// ---- diagonal tile: X = chol(tile)^{-1} by a BLOCKED in-place Gauss-Jordan sweep (8x8 blocks) -------------------
// tools/tile_model.py: blocked_gj_inverse_factor.  Scalar FP64 chains are extremely slow next to CTAs that stream
// DMMAs through the shared FP64 pipe (the register sweep of the first kernel: up to 230k cycles per tile), so
// everything except the eight 8x8 pivot factorisations is done as 8x8x8 products = two DMMA.8x8x4 each.
// In-place layout of the 64x64 tile V (stride LDT) while pivot block J is processed:
//   (I,K), K > J : Schur complement;  (I,J): panel L_IJ, then Y_IJ;  K < J: Y_IK (rows > J) / X_JK (rows <= J).
__device__ __forceinline__ double* blk8(double* V, int I, int K) { return V + (8 * I) * LDT + 8 * K; }

// C (8x8 block, in place) = sc * C + sa * A * B^T   (nt) or  sc * C + sa * A * B  (nn); one warp, k = 8.
template <bool NN>
__device__ __forceinline__ void blk_mma(double* C, const double* A, const double* B, double sa, double sc, int g, int t) {
    const double a0 = sa * A[g * LDT + t], a1 = sa * A[g * LDT + 4 + t];
    double b0, b1;
    if (NN) {
        b0 = B[t * LDT + g];
        b1 = B[(4 + t) * LDT + g];
    } else {
        b0 = B[g * LDT + t];
        b1 = B[g * LDT + 4 + t];
    }
    double2 c = make_double2(0.0, 0.0);
    if (sc != 0.0) {
        c = *reinterpret_cast<const double2*>(C + g * LDT + 2 * t);
        c.x *= sc;
        c.y *= sc;
    }
    dmma884(c.x, c.y, a0, b0);
    dmma884(c.x, c.y, a1, b1);
    __syncwarp();  // C may alias A or B (in-place products): every lane has read its fragments before any lane stores
    *reinterpret_cast<double2*>(C + g * LDT + 2 * t) = c;
}

// Pivot block: P (8x8 SPD, lower valid) -> chol(P)^{-1} (lower, zeros above), every lane of ONE warp redundantly
// (no shuffles, fully unrolled scalar Gauss-Jordan sweep in registers).  Returns prod of the pivots.
__device__ __forceinline__ double diag8(Smem& sm, double* P, int lane, int row0) {
    double p[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) p[i][j] = P[i * LDT + j];
    __syncwarp();  // the result overwrites P: all lanes have read it
    double dprod = 1.0;
    double dsave[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        double d = p[j][j];
        if (!(d > 0.0)) {
            if (lane == 0 && sm.info == 0) sm.info = row0 + j + 1;
            d = 1.0;
        }
        dsave[j] = d;
        dprod *= d;
        const double rinv = fast_rcp_pos(d);
        double w[8];
#pragma unroll
        for (int c = 0; c < 8; c++) w[c] = (c > j) ? p[c][j] : ((c < j) ? p[j][c] : 0.0);
#pragma unroll
        for (int r = j + 1; r < 8; r++) {
            const double mult = w[r] * rinv;
#pragma unroll
            for (int c = 0; c <= r; c++) {
                if (c == j) p[r][c] = -mult;
                else p[r][c] -= mult * w[c];
            }
        }
    }
    // every lane holds every value: uniform (same address, same data) shared-memory stores; 64 lane-predicated
    // branches here cost several times the elimination itself
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const double rs = fast_rsqrt_pos(dsave[i]);
#pragma unroll
        for (int j = 0; j < 8; j++) P[i * LDT + j] = (j < i) ? p[i][j] * rs : ((j == i) ? rs : 0.0);
    }
    __syncwarp();
    return dprod;
}

// St: 64x64 SPD tile (stride LDT, lower triangle valid) -> X = chol(tile)^{-1} in place (zeros above the diagonal);
// sum log L_ii in sm.red[0][0]; sm.info on a non-positive pivot.
__device__ void potrf_inv_tile(Smem& sm, const Lane& L, double* V, int row0) {
    const int g = L.g, t = L.t;
    if (L.warp == 3) {
        const double dp = diag8(sm, blk8(V, 0, 0), L.lane, row0);
        if (L.lane == 0) sm.piv[0] = dp;
    }
    __syncthreads();
#pragma unroll 1
    for (int J = 0; J < 8; J++) {
        const double* Xp = blk8(V, J, J);
        // phase A: multiply by the pivot inverses.  Items:
        //   [0, J)      : X_JK = Xp_J Y_JK (K < J); for K = J-1 first Y_{J,J-1} = -L_{J,J-1} Xp_{J-1}
        //   [J, J+n)    : L_IJ = V_IJ Xp_J^T (I > J)
        //   [J+n, J+2n) : Y_{I,J-1} = -L_{I,J-1} Xp_{J-1} (I > J), the deferred last stage of step J-1
        const int n = 7 - J;
        const int nitems = J + n + (J > 0 ? n : 0);
        for (int item = L.warp; item < nitems; item += 4) {
            if (item < J) {
                if (item == J - 1) {
                    blk_mma<true>(blk8(V, J, item), blk8(V, J, item), blk8(V, J - 1, J - 1), -1.0, 0.0, g, t);
                    __syncwarp();
                }
                blk_mma<true>(blk8(V, J, item), Xp, blk8(V, J, item), 1.0, 0.0, g, t);
            } else if (item < J + n) {
                const int I = J + 1 + (item - J);
                blk_mma<false>(blk8(V, I, J), blk8(V, I, J), Xp, 1.0, 0.0, g, t);
            } else {
                const int I = J + 1 + (item - J - n);
                blk_mma<true>(blk8(V, I, J - 1), blk8(V, I, J - 1), blk8(V, J - 1, J - 1), -1.0, 0.0, g, t);
            }
        }
        __syncthreads();
        if (J == 7) break;
        // phase B: warp 3 finalises and factors the NEXT pivot block (the serial scalar chain) while warps 0..2
        // apply the rest of the rank-8 update: V_IK -= L_IJ L_KJ^T (J < K <= I), Y_IK -= L_IJ X_JK (K < J < I)
        if (L.warp == 3) {
            blk_mma<false>(blk8(V, J + 1, J + 1), blk8(V, J + 1, J), blk8(V, J + 1, J), -1.0, 1.0, g, t);
            __syncwarp();
            const double dp = diag8(sm, blk8(V, J + 1, J + 1), L.lane, row0 + 8 * (J + 1));
            if (L.lane == 0) sm.piv[J + 1] = dp;
        } else {
            // row I > J carries I items: K in [0, I] without K == J; warp w takes items w, w + 3, ... of the
            // row-major list ((J+1, J+1) is warp 3's)
            int I = J + 1, q = L.warp;
            for (;;) {
                while (I < 8 && q >= I) {
                    q -= I;
                    I++;
                }
                if (I >= 8) break;
                const int K = (q < J) ? q : q + 1;
                if (K < J) blk_mma<true>(blk8(V, I, K), blk8(V, I, J), blk8(V, J, K), -1.0, 1.0, g, t);
                else if (!(I == J + 1 && K == J + 1))
                    blk_mma<false>(blk8(V, I, K), blk8(V, I, J), blk8(V, K, J), -1.0, 1.0, g, t);
                q += 3;
            }
        }
        __syncthreads();
    }
    // zero the blocks above the block diagonal (the diagonal blocks already carry their zeros)
    for (int idx = L.tid; idx < TB * TB; idx += THREADS) {
        const int r = idx >> 6, c = idx & (TB - 1);
        if ((c >> 3) > (r >> 3)) V[r * LDT + c] = 0.0;
    }
    __syncthreads();
    if (L.warp == 0) {  // sum log L_ii = 1/2 sum log(pivot products): one log per lane, off the pivot chain
        const double lg = warp_sum(L.lane < 8 ? log(sm.piv[L.lane]) : 0.0);
        if (L.lane == 0) sm.red[0][0] = 0.5 * lg;
    }
    __syncthreads();
}

__device__ __forceinline__ double ktot_entry(const Smem& sm, const BatchedParams& p, int gi, int gj) {
    if (gi < p.M && gj < p.M) {
        double v = cov_eval(sm.cp, p.X + (size_t)gi * p.D, p.n + (size_t)gi * p.D, p.X + (size_t)gj * p.D,
                            p.n + (size_t)gj * p.D, -1);
        if (gi == gj) v += sm.noise2 + p.diag[gi];
        return v;
    }
    return (gi == gj) ? 1.0 : 0.0;
}

template <int FD>
__device__ __forceinline__ void stage_rows(Smem& sm, const BatchedParams& p, const Lane& L, int I, const double* avec) {
    if constexpr (FD == 1 || FD == 2) {
        if (L.tid < TB) {
            const int r = L.tid;
            const int gi = I * TB + r;
            const bool ok = gi < p.M;
            double* pts = sm.R + PTS_OFF;
            int pk = 0;
#pragma unroll
            for (int d = 0; d < FD; d++) {
                pts[r * FD + d] = ok ? p.X[(size_t)gi * FD + d] : 0.0;
                pk |= (ok ? (p.n[(size_t)gi * FD + d] & 255) : 0) << (8 * d);
            }
            reinterpret_cast<int*>(pts + PTS_ORD)[r] = pk;
            if (avec != nullptr) pts[PTS_ALPHA + r] = ok ? avec[gi] : 0.0;
        }
    }
}

template <int FD>
__device__ __forceinline__ PointReg<FD> staged_point(const double* pts, int r) {
    PointReg<FD> q;
    const int pk = reinterpret_cast<const int*>(pts + PTS_ORD)[r];
#pragma unroll
    for (int d = 0; d < FD; d++) {
        q.x[d] = pts[r * FD + d];
        q.n[d] = (pk >> (8 * d)) & 255;
    }
    return q;
}

// C = K_tot - S in place on the staging tile.  Thread tid owns column c = tid & 63 and rows (tid >> 6) + 2u.
template <int FD>
__device__ __forceinline__ void gen_ktot_tile(const Smem& sm, const BatchedParams& p, double* St, int tid, int I, int Jc) {
    const bool diag_tile = (I == Jc);
    const int c = tid & (TB - 1);
    const int gj = Jc * TB + c;
    if constexpr (FD == 0) {
        unsigned want = 0;  // requested kernel parameters (bit q), gall has 1 + GPT_MAX_DIM slots
        for (int q = 0; q < p.nidx; q++)
            if (p.idx[q] < p.nparams && p.idx[q] < 1 + GPT_MAX_DIM) want |= 1u << p.idx[q];
#pragma unroll 1
        for (int u = 0; u < 32; u++) {
            const int r = (tid >> 6) + 2 * u;
            if (diag_tile && c > r) continue;
            St[r * LDT + c] = ktot_entry(sm, p, I * TB + r, gj) - St[r * LDT + c];
        }
    } else {
        SEHoist<FD> h = se_hoist<FD>(sm.cp);
        h.etab = sm.exptab;
        const bool col_ok = gj < p.M;
        const PointReg<FD> pj = load_point<FD>(p.X, p.n, col_ok ? gj : 0);
        const double dj = col_ok ? sm.noise2 + __ldg(p.diag + gj) : 0.0;
        const double* pts = sm.R + PTS_OFF;
        if (FD <= 2 && p.low_order) {
#pragma unroll 8
            for (int u = 0; u < 32; u++) {
                const int r = (tid >> 6) + 2 * u;
                const int gi = I * TB + r;
                const PointReg<FD> pi = staged_point<FD>(pts, r);
                double v = se_value_low<FD, true>(h, pi, pj);
                v = (gi == gj) ? v + dj : v;
                const double pad = (gi == gj) ? 1.0 : 0.0;
                v = (col_ok && gi < p.M) ? v : pad;
                if (!(diag_tile && c > r)) St[r * LDT + c] = v - St[r * LDT + c];
            }
        } else {
#pragma unroll 2
            for (int u = 0; u < 32; u++) {
                const int r = (tid >> 6) + 2 * u;
                if (diag_tile && c > r) continue;
                const int gi = I * TB + r;
                double v;
                if (col_ok && gi < p.M) {
                    PointReg<FD> pi;
                    if constexpr (FD <= 2) pi = staged_point<FD>(pts, r);
                    else pi = load_point<FD>(p.X, p.n, gi);
                    v = se_value<FD>(h, pi, pj);
                    if (gi == gj) v += dj;
                } else {
                    v = (gi == gj) ? 1.0 : 0.0;
                }
                St[r * LDT + c] = v - St[r * LDT + c];
            }
        }
    }
}

template <int FD>
__device__ __forceinline__ void grad_tile(const Smem& sm, const BatchedParams& p, const double* St, int tid, int I, int J,
                                          const double* __restrict__ avec, double (&gall)[1 + GPT_MAX_DIM],
                                          double& tr_kinv) {
    const bool diag_tile = (I == J);
    const int c = tid & (TB - 1);
    const int gj = J * TB + c;
    if (gj >= p.M) return;
    const double aj = avec[gj];
    if constexpr (FD == 0) {
        unsigned want = 0;  // requested kernel parameters (bit q), gall has 1 + GPT_MAX_DIM slots
        for (int q = 0; q < p.nidx; q++)
            if (p.idx[q] < p.nparams && p.idx[q] < 1 + GPT_MAX_DIM) want |= 1u << p.idx[q];
#pragma unroll 1
        for (int u = 0; u < 32; u++) {
            const int r = (tid >> 6) + 2 * u;
            const int gi = I * TB + r;
            if (gi >= p.M || (diag_tile && c > r)) continue;
            const double kinv = St[r * LDT + c];
            double w = avec[gi] * aj - kinv;
            if (diag_tile && c == r) {
                tr_kinv += kinv;
                w *= 0.5;
            }
            if (sm.cp.kid == GPT_KERNEL_SE) {
                double dk[2 + GPT_MAX_DIM];
                se_cov_all(sm.cp, p.X + (size_t)gi * p.D, p.n + (size_t)gi * p.D, p.X + (size_t)gj * p.D,
                           p.n + (size_t)gj * p.D, dk);
#pragma unroll
                for (int q = 0; q < 1 + GPT_MAX_DIM; q++)
                    if (q <= p.D) gall[q] += w * dk[1 + q];
            } else {
                // Matern / Gibbs: dual-number closed forms (covfn_hyper.cuh), only the requested parameters
#pragma unroll
                for (int q = 0; q < 1 + GPT_MAX_DIM; q++)
                    if ((want >> q) & 1u)
                        gall[q] += w * cov_hyper_eval(sm.cp, p.X + (size_t)gi * p.D, p.n + (size_t)gi * p.D,
                                                      p.X + (size_t)gj * p.D, p.n + (size_t)gj * p.D, q);
            }
        }
    } else {
        SEHoist<FD> h = se_hoist<FD>(sm.cp);
        h.etab = sm.exptab;
        const PointReg<FD> pj = load_point<FD>(p.X, p.n, gj);
        const double* pts = sm.R + PTS_OFF;
        double wk = 0.0;
        if (FD <= 2 && p.low_order) {
            double trl = 0.0;
#pragma unroll 8
            for (int u = 0; u < 32; u++) {
                const int r = (tid >> 6) + 2 * u;
                const int gi = I * TB + r;
                const bool use = (gi < p.M) && !(diag_tile && c > r);
                const bool on_diag = diag_tile && (c == r);
                const double kinv = St[r * LDT + c];
                const PointReg<FD> pi = staged_point<FD>(pts, r);
                const double ai = pts[PTS_ALPHA + r];
                double w = ai * aj - kinv;
                w = on_diag ? 0.5 * w : w;
                w = use ? w : 0.0;
                trl += (use && on_diag) ? kinv : 0.0;
                double K, dl[FD];
                se_value_grad_low<FD, true>(h, pi, pj, K, dl);
                wk = fma(w, K, wk);
#pragma unroll
                for (int d = 0; d < FD; d++) gall[1 + d] = fma(w, dl[d], gall[1 + d]);
            }
            tr_kinv += trl;
        } else {
#pragma unroll 2
            for (int u = 0; u < 32; u++) {
                const int r = (tid >> 6) + 2 * u;
                const int gi = I * TB + r;
                if (gi >= p.M || (diag_tile && c > r)) continue;
                const double kinv = St[r * LDT + c];
                PointReg<FD> pi;
                double ai;
                if constexpr (FD <= 2) {
                    pi = staged_point<FD>(pts, r);
                    ai = pts[PTS_ALPHA + r];
                } else {
                    pi = load_point<FD>(p.X, p.n, gi);
                    ai = avec[gi];
                }
                double w = ai * aj - kinv;
                if (diag_tile && c == r) {
                    tr_kinv += kinv;
                    w *= 0.5;
                }
                double K, dl[FD];
                se_value_grad<FD>(h, pi, pj, K, dl);
                wk += w * K;
#pragma unroll
                for (int d = 0; d < FD; d++) gall[1 + d] += w * dl[d];
            }
        }
        gall[0] += (h.sig != 0.0) ? 2.0 * wk / h.sig : 0.0;
    }
}

template <int FD>
__global__ void __launch_bounds__(THREADS, GPT_B4_MINB) ll_batched4_kernel(BatchedParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    Lane L;
    L.tid = threadIdx.x;
    L.warp = L.tid >> 5;
    L.lane = L.tid & 31;
    L.g = L.lane >> 2;
    L.t = L.lane & 3;
    L.wr = L.warp >> 1;
    L.wc = L.warp & 1;
    const int nT = p.nT;
    const int np1 = p.nparams + 1;
    double* ws = p.workspace + (size_t)blockIdx.x * p.ws_per_cta;
    double* zvec = ws + (size_t)(nT * (nT + 1) / 2 + nT) * TILE;
    double* rvec = zvec + (size_t)nT * TB;
    double* avec = rvec + (size_t)nT * TB;
    double* St = sm.R;
    double acc[4][4][2];
    if (L.tid < 64) sm.exptab[L.tid] = GPT_EXP2_64[L.tid];  // published by the first barrier of the theta loop

    for (;;) {
        __syncthreads();
        if (L.tid == 0) {
            sm.theta = atomicAdd(p.counter, 1);
            sm.info = 0;
        }
        __syncthreads();
        const int b = sm.theta;
        if (b >= p.B) break;
        if (L.tid == 0) {
            const double* th = p.thetas + (size_t)b * np1;
            cov_params_init(sm.cp, p.kid, p.D, p.nparams, th);
            sm.noise2 = th[p.nparams] * th[p.nparams];
        }
        __syncthreads();
        const double* yb = p.y + (size_t)b * p.y_stride;
        double logdet = 0.0, zz = 0.0;  // thread 0
        PT_DECL;

        // =========================== phase 1: Cholesky ===========================
        for (int k = 0; k < nT; k++) {
            for (int I = k; I < nT; I++) {
                if (L.tid < k) {
                    sm.a[L.tid] = slot(ws, I, L.tid);
                    sm.b[L.tid] = slot(ws, k, L.tid);
                    sm.flag[L.tid] = 0;
                }
                if (L.tid == 0) sm.skip_upper = (I == k) ? 1 : 0;
                zero_acc(acc);
                PT_MARK(7);
                run_job(sm, L, k, acc);
                PT_MARK(0);
                // S -> staging, then C = K_tot - S in place (K_tot generated from the closed forms, never stored)
                acc_to_tile(St, L, acc);
                stage_rows<FD>(sm, p, L, I, nullptr);
                __syncthreads();
                gen_ktot_tile<FD>(sm, p, St, L.tid, I, k);
                PT_MARK(1);
                if (I == k) {
                    // residual r_k = y_k - sum_j L(k,j) z_j : 2 threads per row, 32 columns each
                    {
                        const int r = L.tid >> 1, h2 = L.tid & 1;
                        double s = 0.0;
                        for (int j = 0; j < k; j++) {
                            const double* row = slot(ws, k, j) + r * TB + h2 * 32;
                            const double* zj = zvec + j * TB + h2 * 32;
#pragma unroll
                            for (int c = 0; c < 32; c++) s += row[c] * zj[c];
                        }
                        s += __shfl_xor_sync(0xffffffffu, s, 1);
                        if (h2 == 0) {
                            const int gi = k * TB + r;
                            sm.rk[r] = ((gi < p.M) ? yb[gi] : 0.0) - s;
                        }
                    }
                    __syncthreads();
                    PT_MARK(4);
                    potrf_inv_tile(sm, L, St, k * TB);
                    PT_MARK(2);
                    if (L.tid == 0) logdet += sm.red[0][0];
                    {
                        // Inv_k, Inv_k^T to the workspace; z_k = Inv_k r_k
                        double* Dk = slot(ws, k, k);
                        double* DTk = slotDT(ws, nT, k);
                        for (int idx = L.tid; idx < TILE; idx += THREADS) {
                            const int r = idx >> 6, c = idx & (TB - 1);
                            Dk[idx] = St[r * LDT + c];
                            DTk[idx] = St[c * LDT + r];
                        }
                        if (L.tid < TB) {
                            double s = 0.0;
                            for (int c = 0; c <= L.tid; c++) s += St[L.tid * LDT + c] * sm.rk[c];
                            zvec[k * TB + L.tid] = s;
                            sm.zk[L.tid] = s;
                        }
                    }
                    __syncthreads();
                    if (L.tid == 0) {
                        double s = 0.0;
                        for (int c = 0; c < TB; c++) s += sm.zk[c] * sm.zk[c];
                        zz += s;
                    }
                } else {
                    __syncthreads();
                    double out[4][4][2];
                    mult_lower_global(St, slot(ws, k, k), L, out);
                    acc_to_global(slot(ws, I, k), L, out, 1.0);
                }
                __syncthreads();
                PT_MARK(3);
            }
        }
        __threadfence_block();

        const bool need_alpha = (p.nidx > 0) || (p.alpha_out != nullptr);
        if (need_alpha) {
            PT_MARK(7);
            // ======================= alpha = L^{-T} z (block back substitution) =======================
            for (int i = L.tid; i < nT * TB; i += THREADS) rvec[i] = zvec[i];
            __syncthreads();
            for (int J = nT - 1; J >= 0; J--) {
                {
                    const int a = L.tid >> 1, h2 = L.tid & 1;
                    const double* row = slotDT(ws, nT, J) + a * TB + h2 * 32;
                    const double* rj = rvec + J * TB + h2 * 32;
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < 32; c++) s += row[c] * rj[c];
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    if (h2 == 0) {
                        sm.zk[a] = s;
                        avec[J * TB + a] = s;
                    }
                }
                __syncthreads();
                for (int I = 0; I < J; I++) {
                    const int c = L.tid >> 1, h2 = L.tid & 1;
                    const double* tile = slot(ws, J, I) + (h2 * 32) * TB + c;
                    double s = 0.0;
#pragma unroll
                    for (int r = 0; r < 32; r++) s += tile[r * TB] * sm.zk[h2 * 32 + r];
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    if (h2 == 0) rvec[I * TB + c] -= s;
                }
                __syncthreads();
            }
            if (p.alpha_out != nullptr)
                for (int i = L.tid; i < p.M; i += THREADS) p.alpha_out[(size_t)b * p.M + i] = avec[i];
            PT_MARK(5);
        }

        double gall[1 + GPT_MAX_DIM];
#pragma unroll
        for (int q = 0; q < 1 + GPT_MAX_DIM; q++) gall[q] = 0.0;
        double tr_kinv = 0.0;

        if (p.nidx > 0 && sm.info == 0) {
            // =========================== phase 2: XT = L^{-T} in place ===========================
            for (int I = 1; I < nT; I++) {
                for (int J = 0; J < I; J++) {
                    const int nsteps = I - J;
                    if (L.tid < nsteps) {
                        const int m = J + L.tid;
                        sm.a[L.tid] = (m == J) ? slotDT(ws, nT, J) : slot(ws, m, J);
                        sm.b[L.tid] = slot(ws, I, m);
                        sm.flag[L.tid] = (m == J) ? 1 : 0;
                    }
                    if (L.tid == 0) sm.skip_upper = 0;
                    zero_acc(acc);
                    PT_MARK(7);
                    run_job(sm, L, nsteps, acc);
                    PT_MARK(0);
                    acc_to_tile(St, L, acc);
                    __syncthreads();
                    double out[4][4][2];
                    mult_lower_global(St, slot(ws, I, I), L, out);
                    acc_to_global(slot(ws, I, J), L, out, -1.0);
                    __syncthreads();
                    PT_MARK(3);
                }
            }
            __threadfence_block();
            // ================== phase 3: K^{-1} tiles + gradient contraction ==================
            for (int J = 0; J < nT; J++) {
                for (int I = J; I < nT; I++) {
                    const int nsteps = nT - I;
                    if (L.tid < nsteps) {
                        const int m = I + L.tid;
                        int fl = 0;
                        const double *aa, *bb;
                        if (m == I) {
                            aa = slotDT(ws, nT, I);
                            bb = (I == J) ? slotDT(ws, nT, J) : slot(ws, I, J);
                            fl = 1 | ((I == J) ? 4 : 0);
                        } else {
                            aa = slot(ws, m, I);
                            bb = slot(ws, m, J);
                        }
                        sm.a[L.tid] = aa;
                        sm.b[L.tid] = bb;
                        sm.flag[L.tid] = (unsigned char)fl;
                    }
                    if (L.tid == 0) sm.skip_upper = (I == J) ? 1 : 0;
                    zero_acc(acc);
                    PT_MARK(7);
                    run_job(sm, L, nsteps, acc);
                    PT_MARK(0);
                    acc_to_tile(St, L, acc);
                    stage_rows<FD>(sm, p, L, I, avec);
                    __syncthreads();
                    grad_tile<FD>(sm, p, St, L.tid, I, J, avec, gall, tr_kinv);
                    PT_MARK(6);
                }
            }
        }

        // =========================== outputs ===========================
        __syncthreads();
        if (p.nidx > 0) {
#pragma unroll
            for (int q = 0; q < 1 + GPT_MAX_DIM; q++) {
                const double s = warp_sum(gall[q]);
                if (L.lane == 0) sm.red[L.warp][q] = s;
            }
            const double s = warp_sum(tr_kinv);
            if (L.lane == 0) sm.red[L.warp][GPT_MAX_PARAMS] = s;
            __syncthreads();
            bool want_noise = false;
            for (int q = 0; q < p.nidx; q++) want_noise |= (p.idx[q] == p.nparams);
            double aa = 0.0;
            if (want_noise) {
                double part = 0.0;
                for (int i = L.tid; i < p.M; i += THREADS) part += avec[i] * avec[i];
                part = warp_sum(part);
                if (L.lane == 0) sm.red[L.warp][GPT_MAX_PARAMS + 1] = part;
                __syncthreads();
                for (int w = 0; w < 4; w++) aa += sm.red[w][GPT_MAX_PARAMS + 1];
            }
            if (L.tid == 0) {
                double tr = 0.0;
                for (int w = 0; w < 4; w++) tr += sm.red[w][GPT_MAX_PARAMS];
                const double sn = p.thetas[(size_t)b * np1 + p.nparams];
                for (int q = 0; q < p.nidx; q++) {
                    double gsum = 0.0;
                    if (p.idx[q] == p.nparams) {
                        gsum = sn * (aa - tr);  // gaussian_process.py:1484-1488: noise kernel derivative 2 sigma_n I
                    } else {
                        for (int w = 0; w < 4; w++) gsum += sm.red[w][p.idx[q]];
                    }
                    p.grad[(size_t)b * p.nidx + q] = (sm.info == 0) ? gsum : 0.0;
                }
            }
        }
        if (L.tid == 0) {
            p.ll[b] = -0.5 * zz - logdet - 0.5 * p.M * 1.8378770664093453;  // log(2 pi)
            p.status[b] = sm.info;
#ifdef GPT_PHASE_TIMING
            PT_MARK(7);
            if (p.phase_cycles)
                for (int q = 0; q < 8; q++)
                    atomicAdd((unsigned long long*)p.phase_cycles + q, (unsigned long long)pt_acc[q]);
#endif
        }
    }
}

// experiments: GPT_B4_CTAS_PER_SM=2|3 pads the dynamic shared memory request so that fewer CTAs fit on an SM
static size_t smem_request() {
    size_t bytes = sizeof(Smem);
    if (const char* e = getenv("GPT_B4_CTAS_PER_SM")) {
        const int v = atoi(e);
        if (v == 3) bytes = 74 * 1024;
        if (v == 2) bytes = 112 * 1024;
    }
    return bytes;
}

template <int FD>
void launch_t(const BatchedParams& p, int num_ctas, cudaStream_t s) {
    const size_t smem_bytes = smem_request();
    cudaFuncSetAttribute(ll_batched4_kernel<FD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
#ifdef GPT_PHASE_TIMING
    {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ll_batched4_kernel<FD>, THREADS, sizeof(Smem));
        fprintf(stderr, "[batched4] smem %zu B per CTA, occupancy %d CTAs/SM, grid %d\n", sizeof(Smem), nb, num_ctas);
    }
#endif
    ll_batched4_kernel<FD><<<num_ctas, THREADS, smem_bytes, s>>>(p);
}

}  // namespace

size_t batched_ws_doubles_per_cta(int nT) {
    return (size_t)(nT * (nT + 1) / 2 + nT) * TILE + (size_t)3 * nT * TB;
}

int batched4_ctas_per_sm() {
    if (const char* e = getenv("GPT_B4_CTAS_PER_SM")) {
        const int v = atoi(e);
        if (v >= 2 && v <= GPT_B4_MINB) return v;
    }
    return GPT_B4_MINB;
}

void launch_ll_batched4(const BatchedParams& p, int num_ctas, cudaStream_t s) {
    if (p.kid == GPT_KERNEL_SE && p.D == 1) launch_t<1>(p, num_ctas, s);
    else if (p.kid == GPT_KERNEL_SE && p.D == 2) launch_t<2>(p, num_ctas, s);
    else if (p.kid == GPT_KERNEL_SE && p.D == 3) launch_t<3>(p, num_ctas, s);
    else launch_t<0>(p, num_ctas, s);
}
