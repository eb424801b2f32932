// Batched log-marginal-likelihood + gradient for many hyper-parameter vectors theta at once --
// the unit of work of GaussianProcess.update_hyperparameters (gaussian_process.py:1332-1416 ->
// compute_K_L_alpha_ll :1418-1522) evaluated for B thetas in ONE launch (emcee walkers, optimizer
// restarts, ll grids).  The reference farms these out to worker processes, one theta per call.
//
// One persistent CTA (256 threads, 2 CTAs per SM) owns one theta at a time (dynamic scheduler) and keeps
// that theta's factor in a private, L2/HBM-resident workspace of 64x64 tiles:
//   phase 1  left-looking blocked Cholesky.  K_tot tiles are GENERATED from the closed forms
//            (covfn.cuh) directly into the DMMA accumulator registers (K is never stored); the
//            diagonal tile is factored and explicitly inverted in shared memory so that the panel
//            solve is a DMMA product as well.  z = L^{-1} y and sum(log L_ii) ride along.
//   backsolve alpha = L^{-T} z
//   phase 2  XT = L^{-T} in place (block substitution, all DMMA)
//   phase 3  K^{-1} tiles = sum_m XT XT^T accumulate in registers and are immediately contracted
//            with the regenerated dK/dtheta tiles: g_p = 1/2 tr((alpha alpha^T - K^{-1}) dK_p).
//            Neither K^{-1} nor any dK is ever written to memory.
// Every matrix product is mma.sync.m8n8k4.f64 (DMMA) with operands streamed global->shared by a
// 3-stage cp.async pipeline in 64x16 chunks (XOR-swizzled, conflict-free fragment reads).
// Algorithmic work: M^3 flop per theta (potrf M^3/3 + inverse 2M^3/3); roofline = FP64 tensor pipe.
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

// Per-phase cycle accounting (development aid): build with -DGPT_PHASE_TIMING and pass BatchedParams::phase_cycles.
#ifdef GPT_PHASE_TIMING
#define PT_DECL long long pt_t0 = clock64(), pt_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PT_MARK(slot)                          \
    do {                                       \
        const long long pt_t1 = clock64();     \
        pt_acc[slot] += pt_t1 - pt_t0;         \
        pt_t0 = pt_t1;                         \
    } while (0)
#else
#define PT_DECL
#define PT_MARK(slot)
#endif
// slots: 0 GEMM jobs, 1 K-tile generation, 2 potrf+inverse of the diagonal tile, 3 panel products + stores,
//        4 residual / z, 5 back substitution, 6 gradient contraction, 7 everything else

namespace {

constexpr int TB = 64;
constexpr int BK = 16;
constexpr int STAGES = 3;
constexpr int THREADS = 256;
constexpr int CHUNK = TB * BK;         // doubles per operand chunk
constexpr int STAGE_D = 3 * CHUNK;     // A0, A1, B
constexpr int R_D = STAGES * STAGE_D;  // pipeline ring (also: staging for two 64x68 tiles)
constexpr int LDT = 68;                // padded tile stride (== 4 mod 16: conflict-free DMMA fragment reads)
constexpr int DG_D = TB * LDT;
constexpr int MAXT = 32;
constexpr int TILE = TB * TB;

struct Smem {
    double R[R_D];
    double Dg[DG_D];
    double dvec[TB], xdiag[TB], rk[TB], zk[TB];
    double red[8][GPT_MAX_PARAMS + 2];
    const double* a0[MAXT];
    const double* a1[MAXT];
    const double* b[MAXT];
    unsigned char flag[MAXT];  // per step: bit0 A0 upper-triangular, bit1 A1 upper-triangular, bit2 B upper-triangular
    int skip_th;               // job: tile (0/1) whose OUTPUT is a diagonal tile (upper 32x32 block not needed), or -1
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
    int tid, warp, lane, g, t, th, wr, wc;
};

// ---- pipelined product: acc(th) += A_th[s] * B[s]^T over steps s (each a 64x64x64 tile product) ----
__device__ __forceinline__ void issue_chunk(Smem& sm, const Lane& L, int q) {
    const int s = q >> 2, kc = q & 3;
    double* base = sm.R + (q % STAGES) * STAGE_D;
    const double* srcs[3] = {sm.a0[s], sm.a1[s], sm.b[s]};
#pragma unroll
    for (int op = 0; op < 3; op++) {
        const double* src = srcs[op];
        if (src == nullptr) continue;
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int idx = L.tid + u * THREADS;
            const int r = idx >> 3, c2 = (idx & 7) * 2;
            cp_async16(base + op * CHUNK + r * BK + (c2 ^ ((r & 3) << 2)), src + r * TB + kc * BK + c2);
        }
    }
}

__device__ __forceinline__ void run_job(Smem& sm, const Lane& L, int nsteps, double (&acc)[4][4][2]) {
    const int total = nsteps * 4;
    __syncthreads();  // step table visible, ring buffer free
#pragma unroll
    for (int q = 0; q < STAGES - 1; q++) {
        if (q < total) issue_chunk(sm, L, q);
        cp_async_commit();
    }
    for (int q = 0; q < total; q++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (q + STAGES - 1 < total) issue_chunk(sm, L, q + STAGES - 1);
        cp_async_commit();
        const int s = q >> 2;
        const double* ap = L.th ? sm.a1[s] : sm.a0[s];
        // structural zeros, skipped at warp (32x32) granularity: an upper-triangular 64x64 operand ([row][c],
        // c >= row) has nothing in columns < 32 of its rows >= 32; the upper-right block of a symmetric
        // diagonal output tile is never used.
        const int fl = sm.flag[s];
        const bool half_zero = (((fl >> L.th) & 1) && L.wr == 1) || ((fl & 4) && L.wc == 1);
        const bool dead = (sm.skip_th == L.th) && L.wr == 0 && L.wc == 1;
        if (ap != nullptr && !dead && !(half_zero && (q & 3) < 2)) {
            const double* aS = sm.R + (q % STAGES) * STAGE_D + L.th * CHUNK + (L.wr * 32 + L.g) * BK;
            const double* bS = sm.R + (q % STAGES) * STAGE_D + 2 * CHUNK + (L.wc * 32 + L.g) * BK;
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
    __syncthreads();  // ring buffer may now be reused as staging
}

__device__ __forceinline__ void zero_acc(double (&acc)[4][4][2]) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
}

// accumulator fragments -> 64x68 staging tile
__device__ __forceinline__ void acc_to_tile(double* tile, const Lane& L, const double (&acc)[4][4][2]) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
            *reinterpret_cast<double2*>(tile + (L.wr * 32 + i * 8 + L.g) * LDT + L.wc * 32 + j * 8 + 2 * L.t) = v;
        }
}

// out = St * Binv^T with Binv lower triangular (rows n, contraction c <= n): both operands in shared (stride LDT)
__device__ __forceinline__ void mult_lower(const double* St, const double* Binv, const Lane& L, double (&out)[4][4][2]) {
    zero_acc(out);
    const int kmax = (L.wc + 1) * 32;
    const double* aS = St + (L.wr * 32 + L.g) * LDT + L.t;
    const double* bS = Binv + (L.wc * 32 + L.g) * LDT + L.t;
#pragma unroll 2
    for (int k0 = 0; k0 < kmax; k0 += 4) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = aS[i * 8 * LDT + k0];
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = bS[j * 8 * LDT + k0];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(out[i][j][0], out[i][j][1], a[i], b[j]);
    }
}

// accumulator fragments -> global 64x64 tile (row-major), scaled
__device__ __forceinline__ void acc_to_global(double* tile, const Lane& L, const double (&acc)[4][4][2], double scale) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double2 v = make_double2(scale * acc[i][j][0], scale * acc[i][j][1]);
            *reinterpret_cast<double2*>(tile + (L.wr * 32 + i * 8 + L.g) * TB + L.wc * 32 + j * 8 + 2 * L.t) = v;
        }
}

// Replace the 64x64 SPD tile in sm.Dg (stride LDT; lower triangle valid) by X = chol(tile)^{-1} (lower
// triangular, zeros above).  Returns sum(log L_ii) via sm.red[0][0]; flags sm.info (LAPACK-style).
//
// In-place Gauss-Jordan sweep held in REGISTERS (tools/tile_model.py: gj_inverse_factor).  Thread (r, q) =
// (tid >> 2, tid & 3) owns the 16 entries V[r][q + 4s].  After step j, position (r, c) holds, for c > j, the
// Schur complement a_rc and, for c <= j, e_rc = (L_unit^{-1})_rc where tile = L_unit D L_unit^T; finally
// X = D^{-1/2} L_unit^{-1}.  Per column one barrier: the pivot column / row vector w_j (64 doubles) is
// published through a double buffer in shared memory (sm.dvec / sm.xdiag).  The factor L itself is never
// needed by the batched path (only its inverse, the pivots and z_k).  ~64 x (18 LDS + 16 DFMA + 1 rcp) per
// thread instead of the serialized shared-memory rank-1 updates of the first version (35% of the phase-1
// samples in profiles/r01b).
__device__ void potrf_inv_tile(Smem& sm, const Lane& L, int row0) {
    double* Dg = sm.Dg;
    const int tid = L.tid;
    const int r = tid >> 2, q = tid & 3;
    double v[16];
#pragma unroll
    for (int s = 0; s < 16; s++) {
        const int c = q + 4 * s;
        v[s] = (c <= r) ? Dg[r * LDT + c] : 0.0;
    }
    double* const wA = sm.dvec;
    double* const wB = sm.xdiag;
    if (q == 0) wA[r] = v[0];  // w_0 = column 0 (pivot at r = 0)
    __syncthreads();
    double dr = 1.0;  // pivot of my row
    // 16 unrolled groups of 4 columns (the 4 columns of a group run in a rolled loop): register t always holds
    // column q + 4t, so every shared-memory offset below is a compile-time constant relative to W + q.
    // Entries above the diagonal carry harmless garbage (never published, never written back).
#pragma unroll
    for (int s = 0; s < 16; s++) {
#pragma unroll 1
        for (int qq = 0; qq < 4; qq++) {
            const int j = 4 * s + qq;
            const double* W = (qq & 1) ? wB : wA;  // j & 1 == qq & 1
            double* Wn = (qq & 1) ? wA : wB;
            double d = W[j];
            if (!(d > 0.0)) {
                if (tid == 0 && sm.info == 0) sm.info = row0 + j + 1;
                d = 1.0;
            }
            if (r == j) dr = d;
            if (r > j) {
                const double mult = W[r] * __drcp_rn(d);
                const double* Wq = W + q;
#pragma unroll
                for (int t = 0; t < 16; t++) {
                    if (t == s) v[t] = (q == qq) ? -mult : v[t] - mult * Wq[4 * t];
                    else v[t] -= mult * Wq[4 * t];
                }
            }
            const int jn = j + 1;
            if (jn < TB) {
                // column jn of the Schur complement (rows >= jn) ...
                if (q == (jn & 3) && r >= jn) Wn[r] = (qq == 3) ? v[(s + 1) & 15] : v[s];
                // ... and row jn of L_unit^{-1} (columns < jn)
                if (r == jn) {
#pragma unroll
                    for (int t = 0; t < 16; t++)
                        if (q + 4 * t < jn) Wn[q + 4 * t] = v[t];
                }
            }
            __syncthreads();
        }
    }
    const double rs = 1.0 / sqrt(dr);
#pragma unroll
    for (int t = 0; t < 16; t++) {
        const int c = q + 4 * t;
        Dg[r * LDT + c] = (c < r) ? v[t] * rs : ((c == r) ? rs : 0.0);
    }
    {
        double s = (q == 0) ? 0.5 * log(dr) : 0.0;
        s = warp_sum(s);
        if (L.lane == 0) sm.zk[L.warp] = s;  // zk is free here (z_k is formed after this call)
    }
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; w++) s += sm.zk[w];
        sm.red[0][0] = s;
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

// ---- register-resident squared-exponential evaluation, input dimension known at compile time ----------
// (the generic cov_eval keeps per-dimension arrays in local memory and re-reads the parameters from shared
// memory for every entry; ncu showed ~60% of all warp samples there, profiles/r01a_*)
template <int D>
struct SEHoist {
    double sig2, sig;
    double il[D];
};

template <int D>
__device__ __forceinline__ SEHoist<D> se_hoist(const CovParams& cp) {
    SEHoist<D> h;
    h.sig2 = cp.sig2;
    h.sig = cp.p[0];
#pragma unroll
    for (int d = 0; d < D; d++) h.il[d] = cp.inv_l[d];
    return h;
}

template <int D>
struct PointReg {
    double x[D];
    int n[D];
};

template <int D>
__device__ __forceinline__ PointReg<D> load_point(const double* __restrict__ X, const int32_t* __restrict__ n, int gi) {
    PointReg<D> q;
#pragma unroll
    for (int d = 0; d < D; d++) {
        q.x[d] = __ldg(X + (size_t)gi * D + d);
        q.n[d] = __ldg(n + (size_t)gi * D + d);
    }
    return q;
}

// value only
template <int D>
__device__ __forceinline__ double se_value(const SEHoist<D>& h, const PointReg<D>& a, const PointReg<D>& b) {
    double r2 = 0.0, prod = 1.0;
    int sj = 0;
#pragma unroll
    for (int d = 0; d < D; d++) {
        const double tau = a.x[d] - b.x[d];
        double tl = tau * h.il[d];
        if (tau == 0.0) tl = 0.0;
        r2 += tl * tl;
        sj += b.n[d];
        double f, g;
        se_dim_factor(tau, h.il[d], a.n[d] + b.n[d], false, f, g);
        prod *= f;
    }
    double k = h.sig2 * exp_nonpos(-0.5 * r2) * prod;
    return (sj & 1) ? -k : k;
}

// value K and dK/dl_d for every dimension (dK/dsigma = 2K/sigma is formed by the caller)
template <int D>
__device__ __forceinline__ void se_value_grad(const SEHoist<D>& h, const PointReg<D>& a, const PointReg<D>& b,
                                              double& K, double (&dl)[D]) {
    double r2 = 0.0;
    int sj = 0;
    double f[D], g[D];
#pragma unroll
    for (int d = 0; d < D; d++) {
        const double tau = a.x[d] - b.x[d];
        double tl = tau * h.il[d];
        if (tau == 0.0) tl = 0.0;
        r2 += tl * tl;
        sj += b.n[d];
        se_dim_factor(tau, h.il[d], a.n[d] + b.n[d], true, f[d], g[d]);
    }
    double base = h.sig2 * exp_nonpos(-0.5 * r2);
    if (sj & 1) base = -base;
    double prod = 1.0;
#pragma unroll
    for (int d = 0; d < D; d++) prod *= f[d];
    K = base * prod;
#pragma unroll
    for (int d = 0; d < D; d++) {
        double pr = g[d];
#pragma unroll
        for (int e = 0; e < D; e++)
            if (e != d) pr *= f[e];
        dl[d] = base * pr;
    }
}

// Branch-free versions for derivative orders <= 1 on both sides (m_d <= 2): selects only, so that the compiler
// can interleave the unrolled evaluations (a conditional branch per entry serialises them -- measured).
template <int D>
__device__ __forceinline__ double se_value_low(const SEHoist<D>& h, const PointReg<D>& a, const PointReg<D>& b) {
    double r2 = 0.0, prod = 1.0;
    int sj = 0;
#pragma unroll
    for (int d = 0; d < D; d++) {
        const double tau = a.x[d] - b.x[d];
        const double tl = (tau == 0.0) ? 0.0 : tau * h.il[d];
        r2 = fma(tl, tl, r2);
        sj += b.n[d];
        double f, g;
        se_dim_factor_low(tau, h.il[d], a.n[d] + b.n[d], f, g);
        prod *= f;
    }
    const double k = h.sig2 * exp_nonpos_nobranch(-0.5 * r2) * prod;
    return (sj & 1) ? -k : k;
}

template <int D>
__device__ __forceinline__ void se_value_grad_low(const SEHoist<D>& h, const PointReg<D>& a, const PointReg<D>& b,
                                                  double& K, double (&dl)[D]) {
    double r2 = 0.0;
    int sj = 0;
    double f[D], g[D];
#pragma unroll
    for (int d = 0; d < D; d++) {
        const double tau = a.x[d] - b.x[d];
        const double tl = (tau == 0.0) ? 0.0 : tau * h.il[d];
        r2 = fma(tl, tl, r2);
        sj += b.n[d];
        se_dim_factor_low(tau, h.il[d], a.n[d] + b.n[d], f[d], g[d]);
    }
    double base = h.sig2 * exp_nonpos_nobranch(-0.5 * r2);
    base = (sj & 1) ? -base : base;
    double prod = 1.0;
#pragma unroll
    for (int d = 0; d < D; d++) prod *= f[d];
    K = base * prod;
#pragma unroll
    for (int d = 0; d < D; d++) {
        double pr = g[d];
#pragma unroll
        for (int e = 0; e < D; e++)
            if (e != d) pr *= f[e];
        dl[d] = base * pr;
    }
}

// ---- row points of the (up to) two tiles of a job, staged in the 512 spare doubles behind the staging tiles ----
// With 2 x 113 KB of shared memory per SM the L1 data cache is gone: every per-entry global load of a point
// is an L2 round trip (measured: 1200 cycles per entry, 35% of the kernel).  Layout (doubles from PTS_OFF):
// x[set][r][d] at (set*64+r)*FD+d (< 256), alpha[set][r] at 256.., packed orders (int) at double offset 384.
constexpr int PTS_OFF = 2 * DG_D;
constexpr int PTS_ALPHA = 256;
constexpr int PTS_ORD = 384;

template <int FD>
__device__ __forceinline__ void stage_rows(Smem& sm, const BatchedParams& p, const Lane& L, int I0, int I1, bool has1,
                                           const double* avec) {
    if constexpr (FD == 1 || FD == 2) {
        if (L.tid < 128) {
            const int set = L.tid >> 6, r = L.tid & 63;
            const int gi = (set ? I1 : I0) * TB + r;
            const bool ok = (set == 0 || has1) && gi < p.M;
            double* pts = sm.R + PTS_OFF;
            int pk = 0;
#pragma unroll
            for (int d = 0; d < FD; d++) {
                pts[(set * 64 + r) * FD + d] = ok ? p.X[(size_t)gi * FD + d] : 0.0;
                pk |= (ok ? (p.n[(size_t)gi * FD + d] & 255) : 0) << (8 * d);
            }
            reinterpret_cast<int*>(pts + PTS_ORD)[set * 64 + r] = pk;
            if (avec != nullptr) pts[PTS_ALPHA + set * 64 + r] = ok ? avec[gi] : 0.0;
        }
    }
}

template <int FD>
__device__ __forceinline__ PointReg<FD> staged_point(const double* pts, int set, int r) {
    PointReg<FD> q;
    const int pk = reinterpret_cast<const int*>(pts + PTS_ORD)[set * 64 + r];
#pragma unroll
    for (int d = 0; d < FD; d++) {
        q.x[d] = pts[(set * 64 + r) * FD + d];
        q.n[d] = (pk >> (8 * d)) & 255;
    }
    return q;
}

// K_tot tile (Imine, Jc) into the staging tile St. Thread tl owns column c = tl & 63 and rows (tl >> 6) + 2u.
// C = K_tot - S in place: St holds S (the accumulated L L^T part); the result goes to Cout (St itself, or the
// diagonal-tile buffer).  Running after the accumulators were dumped keeps their 64 registers free for ILP.
template <int FD>
__device__ __forceinline__ void gen_ktot_tile(const Smem& sm, const BatchedParams& p, const double* St, double* Cout,
                                              int tl, int set, int Imine, int Jc) {
    const bool diag_tile = (Imine == Jc);
    const int c = tl & (TB - 1);
    const int gj = Jc * TB + c;
    if constexpr (FD == 0) {
#pragma unroll 1
        for (int u = 0; u < 32; u++) {
            const int r = (tl >> 6) + 2 * u;
            if (diag_tile && c > r) continue;  // the factorisation only reads the lower triangle
            Cout[r * LDT + c] = ktot_entry(sm, p, Imine * TB + r, gj) - St[r * LDT + c];
        }
    } else {
        const SEHoist<FD> h = se_hoist<FD>(sm.cp);
        const bool col_ok = gj < p.M;
        const PointReg<FD> pj = load_point<FD>(p.X, p.n, col_ok ? gj : 0);
        const double dj = col_ok ? sm.noise2 + __ldg(p.diag + gj) : 0.0;
        const double* pts = sm.R + PTS_OFF;
        if (FD <= 2 && p.low_order) {
            // branch-free body: always evaluate (staged rows are zero-filled when out of range), select, predicated store
#pragma unroll 8
            for (int u = 0; u < 32; u++) {
                const int r = (tl >> 6) + 2 * u;
                const int gi = Imine * TB + r;
                const PointReg<FD> pi = staged_point<FD>(pts, set, r);
                double v = se_value_low<FD>(h, pi, pj);
                v = (gi == gj) ? v + dj : v;
                const double pad = (gi == gj) ? 1.0 : 0.0;
                v = (col_ok && gi < p.M) ? v : pad;
                if (!(diag_tile && c > r)) Cout[r * LDT + c] = v - St[r * LDT + c];
            }
        } else {
#pragma unroll 2
            for (int u = 0; u < 32; u++) {
                const int r = (tl >> 6) + 2 * u;
                if (diag_tile && c > r) continue;
                const int gi = Imine * TB + r;
                double v;
                if (col_ok && gi < p.M) {
                    PointReg<FD> pi;
                    if constexpr (FD <= 2) pi = staged_point<FD>(pts, set, r);
                    else pi = load_point<FD>(p.X, p.n, gi);
                    v = se_value<FD>(h, pi, pj);
                    if (gi == gj) v += dj;
                } else {
                    v = (gi == gj) ? 1.0 : 0.0;
                }
                Cout[r * LDT + c] = v - St[r * LDT + c];
            }
        }
    }
}

// gradient contraction of one K^{-1} tile held in St:  gall[q] += sum w_ij dK_ij/dparam_q  (SE kernel)
template <int FD>
__device__ __forceinline__ void grad_tile(const Smem& sm, const BatchedParams& p, const double* St, int tl, int set,
                                          int Imine, int J, const double* __restrict__ avec,
                                          double (&gall)[1 + GPT_MAX_DIM], double& tr_kinv) {
    const bool diag_tile = (Imine == J);
    const int c = tl & (TB - 1);
    const int gj = J * TB + c;
    if (gj >= p.M) return;
    const double aj = avec[gj];
    if constexpr (FD == 0) {
#pragma unroll 1
        for (int u = 0; u < 32; u++) {
            const int r = (tl >> 6) + 2 * u;
            const int gi = Imine * TB + r;
            if (gi >= p.M || (diag_tile && c > r)) continue;
            const double kinv = St[r * LDT + c];
            double w = avec[gi] * aj - kinv;
            if (diag_tile && c == r) {
                tr_kinv += kinv;
                w *= 0.5;
            }
            double dk[2 + GPT_MAX_DIM];
            se_cov_all(sm.cp, p.X + (size_t)gi * p.D, p.n + (size_t)gi * p.D, p.X + (size_t)gj * p.D,
                       p.n + (size_t)gj * p.D, dk);
#pragma unroll
            for (int q = 0; q < 1 + GPT_MAX_DIM; q++)
                if (q <= p.D) gall[q] += w * dk[1 + q];
        }
    } else {
        const SEHoist<FD> h = se_hoist<FD>(sm.cp);
        const PointReg<FD> pj = load_point<FD>(p.X, p.n, gj);
        const double* pts = sm.R + PTS_OFF;
        double wk = 0.0;  // sum w * K  (-> dK/dsigma = 2K/sigma)
        if (FD <= 2 && p.low_order) {
            double trl = 0.0;
#pragma unroll 8
            for (int u = 0; u < 32; u++) {
                const int r = (tl >> 6) + 2 * u;
                const int gi = Imine * TB + r;
                const bool use = (gi < p.M) && !(diag_tile && c > r);
                const bool on_diag = diag_tile && (c == r);
                const double kinv = St[r * LDT + c];
                const PointReg<FD> pi = staged_point<FD>(pts, set, r);
                const double ai = pts[PTS_ALPHA + set * 64 + r];
                double w = ai * aj - kinv;
                w = on_diag ? 0.5 * w : w;
                w = use ? w : 0.0;
                trl += (use && on_diag) ? kinv : 0.0;
                double K, dl[FD];
                se_value_grad_low<FD>(h, pi, pj, K, dl);
                wk = fma(w, K, wk);
#pragma unroll
                for (int d = 0; d < FD; d++) gall[1 + d] = fma(w, dl[d], gall[1 + d]);
            }
            tr_kinv += trl;
        } else {
#pragma unroll 2
            for (int u = 0; u < 32; u++) {
                const int r = (tl >> 6) + 2 * u;
                const int gi = Imine * TB + r;
                if (gi >= p.M || (diag_tile && c > r)) continue;
                const double kinv = St[r * LDT + c];
                PointReg<FD> pi;
                double ai;
                if constexpr (FD <= 2) {
                    pi = staged_point<FD>(pts, set, r);
                    ai = pts[PTS_ALPHA + set * 64 + r];
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
__global__ void __launch_bounds__(THREADS, 2) ll_batched_kernel_t(BatchedParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    Lane L;
    L.tid = threadIdx.x;
    L.warp = L.tid >> 5;
    L.lane = L.tid & 31;
    L.g = L.lane >> 2;
    L.t = L.lane & 3;
    L.th = L.warp >> 2;
    L.wr = (L.warp >> 1) & 1;
    L.wc = L.warp & 1;
    const int nT = p.nT;
    const int np1 = p.nparams + 1;
    double* ws = p.workspace + (size_t)blockIdx.x * p.ws_per_cta;
    double* zvec = ws + (size_t)(nT * (nT + 1) / 2 + nT) * TILE;  // z = L^{-1} y      (nT*64)
    double* rvec = zvec + (size_t)nT * TB;                        // back-substitution residual
    double* avec = rvec + (size_t)nT * TB;                        // alpha
    double acc[4][4][2];

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
        PT_DECL;
        double logdet = 0.0, zz = 0.0;  // meaningful in thread 0

        // =========================== phase 1: Cholesky ===========================
        for (int k = 0; k < nT; k++) {
            for (int I0 = k; I0 < nT; I0 += 2) {
                const int I1 = I0 + 1;
                const bool has1 = I1 < nT;
                if (L.tid < k) {
                    sm.a0[L.tid] = slot(ws, I0, L.tid);
                    sm.a1[L.tid] = has1 ? slot(ws, I1, L.tid) : nullptr;
                    sm.b[L.tid] = slot(ws, k, L.tid);
                    sm.flag[L.tid] = 0;
                }
                if (L.tid == 0) sm.skip_th = (I0 == k) ? 0 : -1;
                zero_acc(acc);
                PT_MARK(7);
                run_job(sm, L, k, acc);
                PT_MARK(0);
                // K_tot tile generated into the (now idle) staging area, then C = K_tot - acc in registers
                const int Imine = L.th ? I1 : I0;
                const bool active = L.th ? has1 : true;
                // dump S = sum_j L(I,j) L(k,j)^T to the staging tiles, then C = K_tot - S in place (K_tot generated
                // from the closed forms; never stored anywhere else).  The diagonal tile is written to Dg.
                acc_to_tile(sm.R + L.th * DG_D, L, acc);
                stage_rows<FD>(sm, p, L, I0, I1, has1, nullptr);
                __syncthreads();
                if (active) {
                    double* St = sm.R + L.th * DG_D;
                    gen_ktot_tile<FD>(sm, p, St, (Imine == k) ? sm.Dg : St, L.tid & 127, L.th, Imine, k);
                }
                PT_MARK(1);
                if (I0 == k) {
                    // residual r_k = y_k - sum_j L(k,j) z_j  (4 threads per row, 16 columns each)
                    {
                        const int r = L.tid >> 2, qd = L.tid & 3;
                        double s = 0.0;
                        for (int j = 0; j < k; j++) {
                            const double* row = slot(ws, k, j) + r * TB + qd * 16;
                            const double* zj = zvec + j * TB + qd * 16;
#pragma unroll
                            for (int c = 0; c < 16; c++) s += row[c] * zj[c];
                        }
                        s += __shfl_xor_sync(0xffffffffu, s, 1);
                        s += __shfl_xor_sync(0xffffffffu, s, 2);
                        if (qd == 0) {
                            const int gi = k * TB + r;
                            sm.rk[r] = ((gi < p.M) ? yb[gi] : 0.0) - s;
                        }
                    }
                    __syncthreads();
                    PT_MARK(4);
                    potrf_inv_tile(sm, L, k * TB);
                    PT_MARK(2);
                    if (L.tid == 0) logdet += sm.red[0][0];
                    // Inv_k and Inv_k^T to the workspace; z_k = Inv_k r_k
                    {
                        double* Dk = slot(ws, k, k);
                        double* DTk = slotDT(ws, nT, k);
                        for (int idx = L.tid; idx < TILE; idx += THREADS) {
                            const int r = idx >> 6, c = idx & (TB - 1);
                            Dk[idx] = sm.Dg[r * LDT + c];
                            DTk[idx] = sm.Dg[c * LDT + r];
                        }
                        if (L.tid < TB) {
                            double s = 0.0;
                            for (int c = 0; c <= L.tid; c++) s += sm.Dg[L.tid * LDT + c] * sm.rk[c];
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
                    if (has1) {
                        double out[4][4][2];
                        if (L.th == 1) {
                            mult_lower(sm.R + DG_D, sm.Dg, L, out);
                            acc_to_global(slot(ws, I1, k), L, out, 1.0);
                        }
                    }
                } else {
                    __syncthreads();
                    if (active) {
                        double out[4][4][2];
                        mult_lower(sm.R + L.th * DG_D, sm.Dg, L, out);
                        acc_to_global(slot(ws, Imine, k), L, out, 1.0);
                    }
                }
                __syncthreads();  // L(I,k) tiles / z visible to the whole CTA before the next job reads them
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
                    // alpha_J = Inv_J^T r_J : row a of DT(J) (upper triangular), 4 threads per row x 16 columns
                    const int a = L.tid >> 2, qd = L.tid & 3;
                    const double* row = slotDT(ws, nT, J) + a * TB + qd * 16;
                    const double* rj = rvec + J * TB + qd * 16;
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < 16; c++) s += row[c] * rj[c];
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);
                    if (qd == 0) {
                        sm.zk[a] = s;
                        avec[J * TB + a] = s;
                    }
                }
                __syncthreads();
                // r_I -= L(J,I)^T alpha_J for I < J : 4 threads per column (16 rows each), 64 columns per pass
                for (int I = 0; I < J; I++) {
                    const int c = L.tid >> 2, qd = L.tid & 3;
                    const double* tile = slot(ws, J, I) + (qd * 16) * TB + c;
                    double s = 0.0;
#pragma unroll
                    for (int r = 0; r < 16; r++) s += tile[r * TB] * sm.zk[qd * 16 + r];
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);
                    if (qd == 0) rvec[I * TB + c] -= s;
                }
                __syncthreads();
            }
            if (p.alpha_out != nullptr)
                for (int i = L.tid; i < p.M; i += THREADS) p.alpha_out[(size_t)b * p.M + i] = avec[i];
            PT_MARK(5);
        }

        double gall[1 + GPT_MAX_DIM];  // sum w * dK/dparam for every SE parameter (sigma_f, l_1..l_D)
#pragma unroll
        for (int q = 0; q < 1 + GPT_MAX_DIM; q++) gall[q] = 0.0;
        double tr_kinv = 0.0;

        if (p.nidx > 0 && sm.info == 0) {
            // =========================== phase 2: XT = L^{-T} in place ===========================
            for (int I = 1; I < nT; I++) {
                __syncthreads();
                {
                    const double* Di = slot(ws, I, I);
                    for (int idx = L.tid; idx < TILE; idx += THREADS) sm.Dg[(idx >> 6) * LDT + (idx & (TB - 1))] = Di[idx];
                }
                for (int J0 = 0; J0 < I; J0 += 2) {
                    const int J1 = J0 + 1;
                    const bool has1 = J1 < I;
                    const int nsteps = I - J0;
                    if (L.tid < nsteps) {
                        const int m = J0 + L.tid;
                        const double *a0, *a1 = nullptr;
                        a0 = (m == J0) ? slotDT(ws, nT, J0) : slot(ws, m, J0);
                        if (has1 && m >= J1) a1 = (m == J1) ? slotDT(ws, nT, J1) : slot(ws, m, J1);
                        sm.a0[L.tid] = a0;
                        sm.a1[L.tid] = a1;
                        sm.b[L.tid] = slot(ws, I, m);
                        sm.flag[L.tid] = (unsigned char)((m == J0 ? 1 : 0) | ((has1 && m == J1) ? 2 : 0));
                    }
                    if (L.tid == 0) sm.skip_th = -1;
                    zero_acc(acc);
                    PT_MARK(7);
                    run_job(sm, L, nsteps, acc);
                    PT_MARK(0);
                    const bool active = L.th ? has1 : true;
                    acc_to_tile(sm.R + L.th * DG_D, L, acc);
                    __syncthreads();
                    if (active) {
                        double out[4][4][2];
                        mult_lower(sm.R + L.th * DG_D, sm.Dg, L, out);
                        acc_to_global(slot(ws, I, L.th ? J1 : J0), L, out, -1.0);
                    }
                    __syncthreads();
                    PT_MARK(3);
                }
            }
            __threadfence_block();
            // ================== phase 3: K^{-1} tiles + gradient contraction ==================
            for (int J = 0; J < nT; J++) {
                for (int I0 = J; I0 < nT; I0 += 2) {
                    const int I1 = I0 + 1;
                    const bool has1 = I1 < nT;
                    const int nsteps = nT - I0;
                    if (L.tid < nsteps) {
                        const int m = I0 + L.tid;
                        const double *a0, *a1 = nullptr, *bb;
                        int fl = 0;
                        if (m == I0) {
                            a0 = slotDT(ws, nT, I0);
                            bb = (I0 == J) ? slotDT(ws, nT, J) : slot(ws, I0, J);
                            fl = 1 | ((I0 == J) ? 4 : 0);
                        } else {
                            a0 = slot(ws, m, I0);
                            bb = slot(ws, m, J);
                            if (has1) {
                                a1 = (m == I1) ? slotDT(ws, nT, I1) : slot(ws, m, I1);
                                if (m == I1) fl = 2;
                            }
                        }
                        sm.a0[L.tid] = a0;
                        sm.a1[L.tid] = a1;
                        sm.b[L.tid] = bb;
                        sm.flag[L.tid] = (unsigned char)fl;
                    }
                    if (L.tid == 0) sm.skip_th = (I0 == J) ? 0 : -1;
                    zero_acc(acc);
                    PT_MARK(7);
                    run_job(sm, L, nsteps, acc);
                    PT_MARK(0);
                    const int Imine = L.th ? I1 : I0;
                    const bool active = L.th ? has1 : true;
                    acc_to_tile(sm.R + L.th * DG_D, L, acc);
                    stage_rows<FD>(sm, p, L, I0, I1, has1, avec);
                    __syncthreads();
                    if (active)
                        grad_tile<FD>(sm, p, sm.R + L.th * DG_D, L.tid & 127, L.th, Imine, J, avec, gall, tr_kinv);
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
                for (int w = 0; w < 8; w++) aa += sm.red[w][GPT_MAX_PARAMS + 1];
            }
            if (L.tid == 0) {
                double tr = 0.0;
                for (int w = 0; w < 8; w++) tr += sm.red[w][GPT_MAX_PARAMS];
                const double sn = p.thetas[(size_t)b * np1 + p.nparams];
                for (int q = 0; q < p.nidx; q++) {
                    double gsum = 0.0;
                    if (p.idx[q] == p.nparams) {
                        // gaussian_process.py:1484-1488: noise kernel derivative 2 sigma_n I
                        gsum = sn * (aa - tr);
                    } else {
                        for (int w = 0; w < 8; w++) gsum += sm.red[w][p.idx[q]];
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
            if (p.phase_cycles) for (int q = 0; q < 8; q++) atomicAdd((unsigned long long*)p.phase_cycles + q, (unsigned long long)pt_acc[q]);
#endif
        }
    }
}

}  // namespace

size_t batched_ws_doubles_per_cta(int nT) {
    return (size_t)(nT * (nT + 1) / 2 + nT) * TILE + (size_t)3 * nT * TB;
}

int batched_max_ctas(int device) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int per_sm = 2;
    if (const char* e = getenv("GPT_BATCHED_CTAS_PER_SM")) {  // experiments only
        const int v = atoi(e);
        if (v == 1 || v == 2) per_sm = v;
    }
    return sms * per_sm;
}

template <int FD>
static void launch_t(const BatchedParams& p, int num_ctas, cudaStream_t s) {
    cudaFuncSetAttribute(ll_batched_kernel_t<FD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    ll_batched_kernel_t<FD><<<num_ctas, THREADS, sizeof(Smem), s>>>(p);
}

void launch_ll_batched(const BatchedParams& p, int num_ctas, cudaStream_t s) {
    // SE kernel with D <= 3: register-resident fast evaluators; everything else: generic closed forms
    if (p.kid == GPT_KERNEL_SE && p.D == 1) launch_t<1>(p, num_ctas, s);
    else if (p.kid == GPT_KERNEL_SE && p.D == 2) launch_t<2>(p, num_ctas, s);
    else if (p.kid == GPT_KERNEL_SE && p.D == 3) launch_t<3>(p, num_ctas, s);
    else launch_t<0>(p, num_ctas, s);
}
