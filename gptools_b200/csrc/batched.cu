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
#include "common.cuh"
#include "internal.h"

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
        if (ap != nullptr) {
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

// Factor the 64x64 SPD tile in sm.Dg (stride LDT) and replace it by the inverse of its Cholesky factor
// (lower triangular, zeros above).  Returns sum(log L_ii) to thread 0 via sm.red[0][0]; flags sm.info.
__device__ void potrf_inv_tile(Smem& sm, const Lane& L, int row0) {
    double* Dg = sm.Dg;
    const int tid = L.tid;
    const int ta = tid >> 4, tb = tid & 15;
    for (int j = 0; j < TB - 1; j++) {
        double d = Dg[j * LDT + j];
        if (!(d > 0.0)) {
            if (tid == 0 && sm.info == 0) sm.info = row0 + j + 1;
            d = 1.0;
        }
        const double invd = 1.0 / d;
        for (int i = j + 1 + ta; i < TB; i += 16) {
            const double lij = Dg[i * LDT + j] * invd;
            for (int c = j + 1 + tb; c <= i; c += 16) Dg[i * LDT + c] -= lij * Dg[c * LDT + j];
        }
        __syncthreads();
    }
    if (tid < TB) {
        double d = Dg[tid * LDT + tid];
        if (!(d > 0.0)) {
            if (tid == TB - 1 && sm.info == 0) sm.info = row0 + TB;
            d = 1.0;
        }
        const double sd = sqrt(d);
        sm.dvec[tid] = sd;
        sm.xdiag[tid] = 1.0 / sd;
    }
    __syncthreads();
    for (int idx = tid; idx < TB * TB; idx += THREADS) {
        const int r = idx >> 6, c = idx & (TB - 1);
        if (c < r) Dg[r * LDT + c] *= sm.xdiag[c];
    }
    if (L.warp == 0) {
        double s = log(sm.dvec[L.lane]) + log(sm.dvec[L.lane + 32]);
        s = warp_sum(s);
        if (L.lane == 0) sm.red[0][0] = s;
    }
    __syncthreads();
    // X = L^{-1}: column j by the lane quad (4j..4j+3); X^T lives in the strict upper triangle
    {
        const int j = tid >> 2, h = tid & 3;
        const double xjj = sm.xdiag[j];
        for (int i = 1; i < TB; i++) {
            double s = 0.0;
            if (i > j) {
                for (int m = j + h; m < i; m += 4) {
                    const double x = (m == j) ? xjj : Dg[j * LDT + m];
                    s += Dg[i * LDT + m] * x;
                }
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (i > j && h == 0) Dg[j * LDT + i] = -s * sm.xdiag[i];
            __syncwarp();
        }
    }
    __syncthreads();
    // in place: lower <- X, diagonal <- 1/L_ii, upper <- 0
    for (int idx = tid; idx < TB * TB; idx += THREADS) {
        const int r = idx >> 6, c = idx & (TB - 1);
        if (c < r) {
            const double x = Dg[c * LDT + r];
            Dg[r * LDT + c] = x;
            Dg[c * LDT + r] = 0.0;
        } else if (c == r) {
            Dg[r * LDT + r] = sm.xdiag[r];
        }
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

__global__ void __launch_bounds__(THREADS, 2) ll_batched_kernel(BatchedParams p) {
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
                }
                zero_acc(acc);
                run_job(sm, L, k, acc);
                // K_tot tile generated into the (now idle) staging area, then C = K_tot - acc in registers
                const int Imine = L.th ? I1 : I0;
                const bool active = L.th ? has1 : true;
                if (active) {
                    double* St = sm.R + L.th * DG_D;
                    const int tl = L.tid & 127;
                    const bool diag_tile = (Imine == k);
#pragma unroll 1
                    for (int u = 0; u < 32; u++) {
                        const int e = tl + 128 * u;
                        const int r = e >> 6, c = e & (TB - 1);
                        if (diag_tile && c > r) continue;  // the factorisation only reads the lower triangle
                        St[r * LDT + c] = ktot_entry(sm, p, Imine * TB + r, k * TB + c);
                    }
                }
                __syncthreads();
                if (active) {
                    const double* St = sm.R + L.th * DG_D;
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const double2 kv = *reinterpret_cast<const double2*>(
                                St + (L.wr * 32 + i * 8 + L.g) * LDT + L.wc * 32 + j * 8 + 2 * L.t);
                            acc[i][j][0] = kv.x - acc[i][j][0];
                            acc[i][j][1] = kv.y - acc[i][j][1];
                        }
                }
                if (I0 == k) {
                    // diagonal tile -> Dg ; optional second tile -> staging 1
                    if (L.th == 0) acc_to_tile(sm.Dg, L, acc);
                    else if (has1) acc_to_tile(sm.R + DG_D, L, acc);
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
                    potrf_inv_tile(sm, L, k * TB);
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
                    acc_to_tile(sm.R + L.th * DG_D, L, acc);
                    __syncthreads();
                    if (active) {
                        double out[4][4][2];
                        mult_lower(sm.R + L.th * DG_D, sm.Dg, L, out);
                        acc_to_global(slot(ws, Imine, k), L, out, 1.0);
                    }
                }
                __syncthreads();  // L(I,k) tiles / z visible to the whole CTA before the next job reads them
            }
        }
        __threadfence_block();

        const bool need_alpha = (p.nidx > 0) || (p.alpha_out != nullptr);
        if (need_alpha) {
            // ======================= alpha = L^{-T} z (block back substitution) =======================
            for (int i = L.tid; i < nT * TB; i += THREADS) rvec[i] = zvec[i];
            __syncthreads();
            for (int J = nT - 1; J >= 0; J--) {
                if (L.tid < TB) {
                    // alpha_J = Inv_J^T r_J : row a of DT(J)
                    const double* row = slotDT(ws, nT, J) + L.tid * TB;
                    double s = 0.0;
                    for (int c = L.tid; c < TB; c++) s += row[c] * rvec[J * TB + c];
                    sm.zk[L.tid] = s;
                    avec[J * TB + L.tid] = s;
                }
                __syncthreads();
                // r_I -= L(J,I)^T alpha_J for I < J : one thread per column
                for (int col = L.tid; col < J * TB; col += THREADS) {
                    const int I = col >> 6, c = col & (TB - 1);
                    const double* tile = slot(ws, J, I) + c;
                    double s = 0.0;
#pragma unroll 8
                    for (int r = 0; r < TB; r++) s += tile[r * TB] * sm.zk[r];
                    rvec[col] -= s;
                }
                __syncthreads();
            }
            if (p.alpha_out != nullptr)
                for (int i = L.tid; i < p.M; i += THREADS) p.alpha_out[(size_t)b * p.M + i] = avec[i];
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
                    }
                    zero_acc(acc);
                    run_job(sm, L, nsteps, acc);
                    const bool active = L.th ? has1 : true;
                    acc_to_tile(sm.R + L.th * DG_D, L, acc);
                    __syncthreads();
                    if (active) {
                        double out[4][4][2];
                        mult_lower(sm.R + L.th * DG_D, sm.Dg, L, out);
                        acc_to_global(slot(ws, I, L.th ? J1 : J0), L, out, -1.0);
                    }
                    __syncthreads();
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
                        if (m == I0) {
                            a0 = slotDT(ws, nT, I0);
                            bb = (I0 == J) ? slotDT(ws, nT, J) : slot(ws, I0, J);
                        } else {
                            a0 = slot(ws, m, I0);
                            bb = slot(ws, m, J);
                            if (has1) a1 = (m == I1) ? slotDT(ws, nT, I1) : slot(ws, m, I1);
                        }
                        sm.a0[L.tid] = a0;
                        sm.a1[L.tid] = a1;
                        sm.b[L.tid] = bb;
                    }
                    zero_acc(acc);
                    run_job(sm, L, nsteps, acc);
                    const int Imine = L.th ? I1 : I0;
                    const bool active = L.th ? has1 : true;
                    acc_to_tile(sm.R + L.th * DG_D, L, acc);
                    __syncthreads();
                    if (active) {
                        const double* St = sm.R + L.th * DG_D;
                        const int tl = L.tid & 127;
                        const bool diag_tile = (Imine == J);
#pragma unroll 1
                        for (int u = 0; u < 32; u++) {
                            const int e = tl + 128 * u;
                            const int r = e >> 6, c = e & (TB - 1);
                            const int gi = Imine * TB + r, gj = J * TB + c;
                            if (gi >= p.M || gj >= p.M || (diag_tile && c > r)) continue;
                            const double kinv = St[r * LDT + c];
                            double w = avec[gi] * avec[gj] - kinv;
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
                    }
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
            if (L.tid == 0) {
                double tr = 0.0, aa = 0.0;
                for (int w = 0; w < 8; w++) tr += sm.red[w][GPT_MAX_PARAMS];
                for (int i = 0; i < p.M; i++) aa += avec[i] * avec[i];
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
        }
    }
}

}  // namespace

size_t batched_ws_doubles_per_cta(int nT) {
    return (size_t)(nT * (nT + 1) / 2 + nT) * TILE + (size_t)3 * nT * TB;
}

int batched_max_ctas(int device) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 148 * 2;
    return prop.multiProcessorCount * 2;
}

void launch_ll_batched(const BatchedParams& p, int num_ctas, cudaStream_t s) {
    cudaFuncSetAttribute(ll_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    ll_batched_kernel<<<num_ctas, THREADS, sizeof(Smem), s>>>(p);
}
