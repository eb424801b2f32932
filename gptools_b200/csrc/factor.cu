// Single-matrix blocked Cholesky building blocks (the GEMM-shaped parts live in gemm.cu):
//   potrf_diag   : 128x128 diagonal block factorisation + explicit inverse of the block (so that the
//                  panel solve, the triangular solves of predict and trtri all become DMMA GEMMs),
//                  the block's share of z = L^{-1} y, its log-determinant share and LAPACK-style info
//   panel_trsm   : one-pass panel solve P = A21 L11^{-T} (+ panel copy, right-hand-side update)
//   backsolve    : alpha = L^{-T} z, one launch (chain of CTAs)
//   grad_reduce  : 1/2 tr((a a^T - S) dK_p) with the dK tiles regenerated on the fly (never stored)
// Replaces scipy.linalg.cholesky / cho_solve in compute_K_L_alpha_ll (gaussian_process.py:1452-1504).
#include "common.cuh"
#include "internal.h"

namespace {

constexpr int NB = GPT_NB;
constexpr int LDB = NB + 4;  // stride == 4 (mod 16): conflict-free DMMA fragment loads from the block
constexpr int NBLK = NB / 8;
constexpr int PCH = NB * 20;  // one 128 x 16 chunk of the previous panel block, rows padded to 20
constexpr size_t POTRF_SMEM = ((size_t)NB * LDB + 2 * NB + 2 * PCH) * sizeof(double);

// ---- 128x128 diagonal block: blocked in-place Gauss-Jordan sweep on 8x8 sub-blocks ------------------------------
// The first version of this kernel eliminated one column per __syncthreads with scalar FP64 (207 us per block,
// and it sits on the critical path of the blocked Cholesky: 13% of an M = 24576 factorisation, all of an M = 4000
// one).  Here everything except the sixteen 8x8 pivot factorisations is an 8x8x8 product = two DMMA.8x8x4
// (same scheme as batched4.cu::potrf_inv_tile, tools/tile_model.py: blocked_gj_inverse_factor).
// In-place layout of V while pivot block J is processed:
//   (I,K), K > J : Schur complement;  (I,J): panel L_IJ (also written to A), later Y_IJ = -L_IJ Xp_J;
//   K < J: Y_IK (rows > J) / X_JK (rows <= J, final inverse).
__device__ __forceinline__ double* blk8(double* V, int I, int K) { return V + (8 * I) * LDB + 8 * K; }

// One 8x8x8 block product: C (in place) = sc * C + sa * A * B^T (nt) or sc * C + sa * A * B (nn), G != null: also
// stored to global.  A warp runs up to FOUR independent products at a time (loads of all four, then the DMMAs
// interleaved, then the stores) -- the products are latency chains (LDS -> DMMA -> DMMA -> STS), and the kernel is
// a single CTA on the critical path of the blocked Cholesky, so instruction-level parallelism is all there is.
struct BlkItem {
    double* C;
    const double* A;
    const double* B;
    double* G;
    int nn;
    bool valid;
};

__device__ __forceinline__ void blk_batch(const BlkItem (&it)[4], double sa, double sc, long ldg, int g, int t) {
    double a0[4], a1[4], b0[4], b1[4];
    double2 c[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        if (it[u].valid) {
            const int ob0 = it[u].nn ? t * LDB + g : g * LDB + t;
            const int ob1 = it[u].nn ? (4 + t) * LDB + g : g * LDB + 4 + t;
            a0[u] = it[u].A[g * LDB + t];
            a1[u] = it[u].A[g * LDB + 4 + t];
            b0[u] = it[u].B[ob0];
            b1[u] = it[u].B[ob1];
            c[u] = (sc != 0.0) ? *reinterpret_cast<const double2*>(it[u].C + g * LDB + 2 * t) : make_double2(0.0, 0.0);
        }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
        if (it[u].valid) {
            a0[u] *= sa;
            a1[u] *= sa;
            c[u].x *= sc;
            c[u].y *= sc;
            dmma884(c[u].x, c[u].y, a0[u], b0[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < 4; u++)
        if (it[u].valid) dmma884(c[u].x, c[u].y, a1[u], b1[u]);
    __syncwarp();  // C may alias A or B (in-place products): every lane has read its fragments before any lane stores
#pragma unroll
    for (int u = 0; u < 4; u++) {
        if (it[u].valid) {
            *reinterpret_cast<double2*>(it[u].C + g * LDB + 2 * t) = c[u];
            if (it[u].G) *reinterpret_cast<double2*>(it[u].G + g * ldg + 2 * t) = c[u];
        }
    }
}

// Pivot block P (8x8 SPD, lower valid): in place -> chol(P)^{-1} (zeros above the diagonal); chol(P) itself goes
// to the global block G (zeros above); the eight pivots go to dsm (their logs are taken in parallel at the end).
// Every lane of one warp redundantly, fully unrolled in registers.
__device__ __forceinline__ void pivot8(double* P, double* G, long ldg, double* dsm, double* lsm, int lane, int row0,
                                       int* s_info) {
    double p[8][8], lc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) p[i][j] = P[i * LDB + j];
    __syncwarp();  // the result overwrites P: all lanes have read it
    double dsave[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        double d = p[j][j];
        if (!(d > 0.0)) {
            if (lane == 0 && *s_info == 0) *s_info = row0 + j + 1;
            d = 1.0;
        }
        dsave[j] = d;
        const double rinv = fast_rcp_pos(d);
        double w[8];
#pragma unroll
        for (int c = 0; c < 8; c++) w[c] = (c > j) ? p[c][j] : ((c < j) ? p[j][c] : 0.0);
#pragma unroll
        for (int r = j + 1; r < 8; r++) lc[r][j] = w[r];  // unscaled column j of the Cholesky factor
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
    double rs[8];
#pragma unroll
    for (int i = 0; i < 8; i++) rs[i] = fast_rsqrt_pos(dsave[i]);
    // Every lane holds every value: store them with uniform (same address, same data) shared-memory writes --
    // 64 lane-predicated branches cost ~7000 cycles here, five times the elimination itself.  The factor goes
    // through the scratch block lsm and is copied to global by all lanes.
#pragma unroll
    for (int i = 0; i < 8; i++) {
        dsm[i] = dsave[i];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            P[i * LDB + j] = (j < i) ? p[i][j] * rs[i] : ((j == i) ? rs[i] : 0.0);
            lsm[i * 8 + j] = (j < i) ? lc[i][j] * rs[j] : ((j == i) ? dsave[i] * rs[i] : 0.0);
        }
    }
    __syncwarp();
    G[(lane >> 3) * ldg + (lane & 7)] = lsm[lane];
    G[((lane >> 3) + 4) * ldg + (lane & 7)] = lsm[lane + 32];
    __syncwarp();
}

// Pprev != null: the block still lacks the rank-128 / rank-256 update of the previous step(s), A -= Pprev Pprev^T
// (Pprev = the 128 x (16 pchunks) block of the previous panel(s) that sits beside this diagonal block, leading
// dimension ldp); applying it here takes that update off the critical path of the blocked Cholesky.
__global__ void __launch_bounds__(256, 1) potrf_diag_kernel(double* __restrict__ A, long lda, double* __restrict__ inv,
                                                            double* __restrict__ yk, double* __restrict__ logdet_part,
                                                            int* __restrict__ info, int row0,
                                                            const double* __restrict__ Pprev, long ldp, int pchunks) {
    extern __shared__ __align__(16) double sm[];
    double* V = sm;                  // NB x LDB
    double* yv = sm + NB * LDB;      // right-hand side block
    double* dsm = yv + NB;           // the NB pivots
    __shared__ int s_info;
    __shared__ double lsm[64];       // pivot8 scratch: one 8x8 block of the factor
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;

    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx >> 7, c = idx & (NB - 1);
        V[r * LDB + c] = (c <= r) ? A[(long)r * lda + c] : 0.0;
    }
    if (tid < NB && yk != nullptr) yv[tid] = yk[tid];
    if (tid == 0) s_info = 0;
    __syncthreads();
    if (Pprev != nullptr) {
        // 136 lower 8x8 blocks, 17 per warp, accumulated over eight 16-wide chunks of Pprev (double-buffered)
        double* Pc = dsm + NB;
        double2 acc[17];
        int offI[17], offK[17];
#pragma unroll
        for (int j = 0; j < 17; j++) {
            const int q = warp + 8 * j;
            int I = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
            while ((I + 1) * (I + 2) / 2 <= q) I++;
            while (I * (I + 1) / 2 > q) I--;
            const int K = q - I * (I + 1) / 2;
            offI[j] = 8 * I;
            offK[j] = 8 * K;
            acc[j] = make_double2(0.0, 0.0);
        }
        auto load_chunk = [&](int st, int kc) {
            for (int idx = tid; idx < NB * 8; idx += 256) {
                const int r = idx >> 3, c2 = (idx & 7) * 2;
                cp_async16(Pc + st * PCH + r * 20 + c2, Pprev + (long)r * ldp + kc * 16 + c2);
            }
        };
        load_chunk(0, 0);
        cp_async_commit();
        for (int kc = 0; kc < pchunks; kc++) {
            if (kc + 1 < pchunks) load_chunk((kc + 1) & 1, kc + 1);
            cp_async_commit();
            cp_async_wait<1>();
            __syncthreads();
            const double* Ps = Pc + (kc & 1) * PCH;
#pragma unroll
            for (int j = 0; j < 17; j++) {
#pragma unroll
                for (int kk = 0; kk < 4; kk++)
                    dmma884(acc[j].x, acc[j].y, Ps[(offI[j] + g) * 20 + kk * 4 + t], Ps[(offK[j] + g) * 20 + kk * 4 + t]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < 17; j++) {
            double2* v = reinterpret_cast<double2*>(V + (offI[j] + g) * LDB + offK[j] + 2 * t);
            double2 cur = *v;
            cur.x -= acc[j].x;
            cur.y -= acc[j].y;
            *v = cur;
        }
        __syncthreads();
    }
    // the factor is a GEMM operand later: the blocks above the block diagonal must be clean zeros
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx >> 7, c = idx & (NB - 1);
        if ((c >> 3) > (r >> 3)) A[(long)r * lda + c] = 0.0;
    }
    if (warp == 7) pivot8(blk8(V, 0, 0), A, lda, dsm, lsm, lane, row0, &s_info);
    __syncthreads();

#pragma unroll 1
    for (int J = 0; J < NBLK; J++) {
        double* Xp = blk8(V, J, J);
        // ---- phase A: multiply by the pivot inverses.
        //   chain (J > 0, warp 0) : Y_{J,J-1} = -L_{J,J-1} Xp_{J-1}, then X_{J,J-1} = Xp_J Y_{J,J-1}
        //   a [0, na)             : X_JK = Xp_J Y_JK, K < J-1                               (nn, sa = +1)
        //   b [na, na+n)          : L_IJ = V_IJ Xp_J^T, I > J, the final factor -> also to A (nt, sa = +1)
        //   c [na+n, na+2n)       : Y_{I,J-1} = -L_{I,J-1} Xp_{J-1}, I > J (deferred from step J-1) (nn, sa = -1)
        const int n = NBLK - 1 - J;
        const int na = (J > 0) ? J - 1 : 0;
        if (J > 0 && warp == 0) {
            BlkItem it[4];
            it[0].C = blk8(V, J, J - 1); it[0].A = blk8(V, J, J - 1); it[0].B = blk8(V, J - 1, J - 1);
            it[0].G = nullptr; it[0].nn = 1; it[0].valid = true;
            it[1].valid = it[2].valid = it[3].valid = false;
            blk_batch(it, -1.0, 0.0, lda, g, t);
            __syncwarp();
            it[0].A = Xp; it[0].B = blk8(V, J, J - 1);
            blk_batch(it, 1.0, 0.0, lda, g, t);
        }
        for (int base = warp; base < na + n; base += 32) {  // sa = +1 items
            BlkItem it[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int item = base + 8 * u;
                it[u].valid = item < na + n;
                if (item < na) {
                    it[u].C = blk8(V, J, item); it[u].A = Xp; it[u].B = blk8(V, J, item); it[u].G = nullptr; it[u].nn = 1;
                } else {
                    const int I = J + 1 + (item - na);
                    it[u].C = blk8(V, I, J); it[u].A = blk8(V, I, J); it[u].B = Xp;
                    it[u].G = A + (long)(8 * I) * lda + 8 * J; it[u].nn = 0;
                }
            }
            blk_batch(it, 1.0, 0.0, lda, g, t);
        }
        if (J > 0) {
            for (int base = warp; base < n; base += 32) {  // sa = -1 items
                BlkItem it[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int I = J + 1 + base + 8 * u;
                    it[u].valid = (base + 8 * u) < n;
                    it[u].C = blk8(V, I, J - 1); it[u].A = blk8(V, I, J - 1); it[u].B = blk8(V, J - 1, J - 1);
                    it[u].G = nullptr; it[u].nn = 1;
                }
                blk_batch(it, -1.0, 0.0, lda, g, t);
            }
        }
        __syncthreads();
        if (J + 1 == NBLK) break;
        // ---- phase B: warp 7 finalises and factors the next pivot block while the others apply the rank-8 update
        if (warp == 7) {
            BlkItem it[4];
            it[0].C = blk8(V, J + 1, J + 1); it[0].A = blk8(V, J + 1, J); it[0].B = blk8(V, J + 1, J);
            it[0].G = nullptr; it[0].nn = 0; it[0].valid = true;
            it[1].valid = it[2].valid = it[3].valid = false;
            blk_batch(it, -1.0, 1.0, lda, g, t);
            __syncwarp();
            pivot8(blk8(V, J + 1, J + 1), A + (long)(8 * (J + 1)) * lda + 8 * (J + 1), lda, dsm + 8 * (J + 1), lsm, lane,
                   row0 + 8 * (J + 1), &s_info);
        } else {
            // Row I > J carries I items: K in [0, I] without K == J (K < J: inverse part Y_IK -= L_IJ X_JK,
            // K > J: trailing part V_IK -= L_IJ L_KJ^T).  Warp w takes the items w, w + 7, ... of the row-major
            // list; (J+1, J+1) is warp 7's.
            int I = J + 1, q = warp;
            bool more = true;
            while (more) {
                BlkItem it[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    while (I < NBLK && q >= I) {
                        q -= I;
                        I++;
                    }
                    const bool in = I < NBLK;
                    const int K = (q < J) ? q : q + 1;
                    it[u].valid = in && !(I == J + 1 && K == J + 1);
                    const int Ic = in ? I : J + 1, Kc = in ? K : 0;
                    it[u].C = blk8(V, Ic, Kc);
                    it[u].A = blk8(V, Ic, J);
                    it[u].nn = (Kc < J) ? 1 : 0;
                    it[u].B = (Kc < J) ? blk8(V, J, Kc) : blk8(V, Kc, J);
                    it[u].G = nullptr;
                    q += 7;
                    more = in;
                }
                blk_batch(it, -1.0, 1.0, lda, g, t);
            }
        }
        __syncthreads();
    }

    // V now holds X = L^{-1} in its lower block triangle (diagonal blocks carry their own zeros)
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx >> 7, c = idx & (NB - 1);
        inv[r * NB + c] = ((c >> 3) <= (r >> 3)) ? V[r * LDB + c] : 0.0;
    }
    if (yk != nullptr) {
        // z_k = X y_k : two lanes per row
        const int r = tid >> 1, h = tid & 1;
        double sacc = 0.0;
        for (int c = h; c <= r; c += 2) sacc += V[r * LDB + c] * yv[c];
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
        if (h == 0) yk[r] = sacc;
    }
    __shared__ double red[4];
    if (warp < 4) {  // sum(log L_ii) = 1/2 sum(log d_i), one pivot per thread
        const double lg = warp_sum(log(dsm[tid]));
        if (lane == 0) red[warp] = lg;
    }
    __syncthreads();
    if (tid == 0) {
        *logdet_part = 0.5 * ((red[0] + red[1]) + (red[2] + red[3]));
        if (s_info != 0) atomicCAS(info, 0, s_info);
    }
}

// ---- panel solve: P = A21 L11^{-T}, one pass, in place -----------------------------------------------------------
// Blocked forward substitution on 8-column blocks, each warp owning 8 rows of the panel with the whole 8 x 128 row
// slab in accumulator registers (16 DMMA C-fragments): for J = 0..15: P_J = A_J Xp_J^T (Xp_J = the 8x8 diagonal
// block inverse, i.e. the diagonal block of Inv), then A_K -= P_J L_KJ^T for K > J.  Warps never synchronise with
// each other after the operand load.  Replaces "multiply by the explicit 128x128 inverse + one refinement step"
// (3 GEMM launches + 2 copies + the rhs update, ~75 us on the critical path of the blocked Cholesky) by one ~10 us
// launch that is backward stable at the 8x8 block level.  Also writes the panel copy the rank-128 update reads and
// applies the right-hand-side update y -= P z_k.
constexpr int TR_ROWS = 64;
constexpr size_t TRSM_SMEM = ((size_t)NB * LDB + (size_t)TR_ROWS * LDB + NBLK * 64) * sizeof(double);

__global__ void __launch_bounds__(256, 1) panel_trsm_kernel(double* __restrict__ A21, long lda,
                                                            const double* __restrict__ L11, long ldl,
                                                            const double* __restrict__ inv_k,
                                                            double* __restrict__ out0, long ld0, int split,
                                                            double* __restrict__ out1, long ld1,
                                                            int rows, const double* __restrict__ zk,
                                                            double* __restrict__ y) {
    extern __shared__ __align__(16) double sm[];
    double* Ls = sm;                       // NB x LDB: L11 (lower; the zeros above are loaded too)
    double* As = Ls + NB * LDB;            // TR_ROWS x LDB: the row slab, A21 on entry, P on exit
    double* Xs = As + TR_ROWS * LDB;       // NBLK blocks of 8 x 8: diagonal blocks of Inv
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = blockIdx.x * TR_ROWS;

    // TMA-staged panel: every 1 KB row of L11 (128 rows) and of the A21 slab (64 rows) is ONE bulk async copy
    // (cp.async.bulk -> UBLKCP) issued by one thread and completing on an mbarrier with a transaction count; no
    // per-thread 16-byte copies, no wait_group.  The 8x8 diagonal blocks of Inv (64 B pieces) stay on cp.async.
    __shared__ unsigned long long bar;
    const int nrows_here = (rows - r0 < TR_ROWS) ? (rows - r0) : TR_ROWS;
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int idx = tid; idx < (TR_ROWS - nrows_here) * NB; idx += 256)   // rows beyond the matrix: zeros
        As[(nrows_here + idx / NB) * LDB + (idx % NB)] = 0.0;
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) mbar_expect_tx(&bar, (uint32_t)((NB + nrows_here) * NB * sizeof(double)));
    __syncthreads();  // the expected byte count is posted before any copy can complete
    if (tid < NB) bulk_g2s(Ls + tid * LDB, L11 + (long)tid * ldl, NB * sizeof(double), &bar);
    else if (tid - NB < nrows_here) bulk_g2s(As + (tid - NB) * LDB, A21 + (long)(r0 + tid - NB) * lda, NB * sizeof(double), &bar);
    for (int idx = tid; idx < NBLK * 32; idx += 256) {
        const int J = idx >> 5, rr = (idx >> 2) & 7, cc = (idx & 3) * 2;
        cp_async16(Xs + J * 64 + rr * 8 + cc, inv_k + (long)(8 * J + rr) * NB + 8 * J + cc);
    }
    cp_async_commit();
    cp_async_wait<0>();
    mbar_wait(&bar, 0);
    __syncthreads();

    double* Aw = As + (warp * 8 + g) * LDB;  // this lane's row of the warp's 8-row slab
    double2 acc[NBLK];
#pragma unroll
    for (int K = 0; K < NBLK; K++) acc[K] = *reinterpret_cast<const double2*>(Aw + 8 * K + 2 * t);
#pragma unroll
    for (int J = 0; J < NBLK; J++) {
        // A_J (C-fragment layout) -> A-fragment layout through the slab
        *reinterpret_cast<double2*>(Aw + 8 * J + 2 * t) = acc[J];
        __syncwarp();
        const double a0 = Aw[8 * J + t], a1 = Aw[8 * J + 4 + t];
        const double x0 = Xs[J * 64 + g * 8 + t], x1 = Xs[J * 64 + g * 8 + 4 + t];
        double2 pj = make_double2(0.0, 0.0);
        dmma884(pj.x, pj.y, a0, x0);
        dmma884(pj.x, pj.y, a1, x1);
        __syncwarp();
        *reinterpret_cast<double2*>(Aw + 8 * J + 2 * t) = pj;  // final P_J
        __syncwarp();
        const double n0 = -Aw[8 * J + t], n1 = -Aw[8 * J + 4 + t];
#pragma unroll
        for (int K = 0; K < NBLK; K++) {
            if (K > J) {
                const double b0 = Ls[(8 * K + g) * LDB + 8 * J + t], b1 = Ls[(8 * K + g) * LDB + 8 * J + 4 + t];
                dmma884(acc[K].x, acc[K].y, n0, b0);
                dmma884(acc[K].x, acc[K].y, n1, b1);
            }
        }
    }
    __syncwarp();
    // write P to the panel buffer and back into A21; y -= P z_k
    double2 za = make_double2(0.0, 0.0), zb = za;
    if (y != nullptr) {
        za = *reinterpret_cast<const double2*>(zk + lane * 4);
        zb = *reinterpret_cast<const double2*>(zk + lane * 4 + 2);
    }
#pragma unroll
    for (int rr = 0; rr < 8; rr++) {
        const int row = r0 + warp * 8 + rr;
        const double2 va = *reinterpret_cast<const double2*>(As + (warp * 8 + rr) * LDB + lane * 4);
        const double2 vb = *reinterpret_cast<const double2*>(As + (warp * 8 + rr) * LDB + lane * 4 + 2);
        double sdot = (va.x * za.x + va.y * za.y) + (vb.x * zb.x + vb.y * zb.y);
        sdot = warp_sum(sdot);
        if (row < rows) {
            // panel copy for the rank updates: rows < split go to out0, the others (re-based) to out1
            double* pout = (row < split) ? out0 + (long)row * ld0 : out1 + (long)(row - split) * ld1;
            *reinterpret_cast<double2*>(pout + lane * 4) = va;
            *reinterpret_cast<double2*>(pout + lane * 4 + 2) = vb;
            *reinterpret_cast<double2*>(A21 + (long)row * lda + lane * 4) = va;
            *reinterpret_cast<double2*>(A21 + (long)row * lda + lane * 4 + 2) = vb;
            if (lane == 0 && y != nullptr) y[row] -= sdot;
        }
    }
}

__global__ void transpose_kernel(double* __restrict__ out, long ldo, const double* __restrict__ in, long ldi,
                                 int rows, int cols, long stride_out, long stride_in) {
    __shared__ double tile[32][33];
    out += (long)blockIdx.z * stride_out;
    in += (long)blockIdx.z * stride_in;
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int ir = by + r, ic = bx + threadIdx.x;
        tile[r][threadIdx.x] = (ir < rows && ic < cols) ? in[(long)ir * ldi + ic] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int orow = bx + r, ocol = by + threadIdx.x;
        if (orow < cols && ocol < rows) out[(long)orow * ldo + ocol] = tile[threadIdx.x][r];
    }
}

__global__ void copy2d_kernel(double* __restrict__ out, long ldo, const double* __restrict__ in, long ldi,
                              int rows, int cols) {
    // 8 rows per CTA (a warp per row, lanes stride the columns); rows live in grid.x, which has no 65535 limit
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const double* src = in + (long)r * ldi;
    double* dst = out + (long)r * ldo;
    for (int c = threadIdx.x & 31; c < cols; c += 32) dst[c] = src[c];
}

__global__ void add_diag_kernel(double* __restrict__ A, long lda, const double* __restrict__ d, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[(long)i * lda + i] += d[i];
}

__global__ void fill_kernel(double* __restrict__ p, long n, double v) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void identity_pad_kernel(double* __restrict__ A, long lda, int n_valid, int n_pad) {
    // rows/cols >= n_valid: zero, with ones on the diagonal
    // only the padding is touched: full rows r >= n_valid, and the column strip c >= n_valid of the other rows
    const int r = blockIdx.x;  // rows in grid.x (no 65535 limit)
    if (r >= n_pad) return;
    const int c0 = (r >= n_valid) ? 0 : n_valid;
    for (int c = c0 + threadIdx.x; c < n_pad; c += blockDim.x) A[(long)r * lda + c] = (r == c) ? 1.0 : 0.0;
}

constexpr int GT = 64;

__global__ void __launch_bounds__(256) grad_reduce_kernel(GradReduceParams p) {
    __shared__ double sx_r[GT * GPT_MAX_DIM], sx_c[GT * GPT_MAX_DIM];
    __shared__ int32_t sn_r[GT * GPT_MAX_DIM], sn_c[GT * GPT_MAX_DIM];
    __shared__ double red[8][GPT_MAX_PARAMS];
    const int bx = blockIdx.x, by = blockIdx.y;
    const int cta = by * gridDim.x + bx;
    const int tid = threadIdx.x;
    double acc[GPT_MAX_PARAMS];
#pragma unroll
    for (int q = 0; q < GPT_MAX_PARAMS; q++) acc[q] = 0.0;
    if (bx <= by) {
        const int D = p.cp.D;
        const int r0 = by * GT, c0 = bx * GT;
        for (int i = tid; i < GT * D; i += 256) {
            const int r = i / D, d = i - r * D;
            const bool okr = (r0 + r) < p.N, okc = (c0 + r) < p.N;
            sx_r[r * GPT_MAX_DIM + d] = okr ? p.X[(long)(r0 + r) * D + d] : 0.0;
            sn_r[r * GPT_MAX_DIM + d] = okr ? p.n[(long)(r0 + r) * D + d] : 0;
            sx_c[r * GPT_MAX_DIM + d] = okc ? p.X[(long)(c0 + r) * D + d] : 0.0;
            sn_c[r * GPT_MAX_DIM + d] = okc ? p.n[(long)(c0 + r) * D + d] : 0;
        }
        __syncthreads();
        const int ty = tid >> 4, tx = tid & 15;
#pragma unroll 1
        for (int a = 0; a < 4; a++) {
            const int lr = ty + 16 * a, r = r0 + lr;
#pragma unroll 1
            for (int b = 0; b < 4; b++) {
                const int lc = tx + 16 * b, c = c0 + lc;
                if (r >= p.N || c > r) continue;
                double w = -p.S[(long)r * p.lds + c];
                if (p.a) w += p.a[r] * p.a[c];
                if (r == c) w *= 0.5;
                if (p.cp.kid == GPT_KERNEL_SE) {
                    double dk[2 + GPT_MAX_DIM];
                    se_cov_all(p.cp, sx_r + lr * GPT_MAX_DIM, sn_r + lr * GPT_MAX_DIM, sx_c + lc * GPT_MAX_DIM,
                               sn_c + lc * GPT_MAX_DIM, dk);
                    for (int q = 0; q < p.nidx; q++) acc[q] += w * dk[1 + p.idx[q]];
                } else {
                    for (int q = 0; q < p.nidx; q++)
                        acc[q] += w * cov_eval(p.cp, sx_r + lr * GPT_MAX_DIM, sn_r + lr * GPT_MAX_DIM,
                                               sx_c + lc * GPT_MAX_DIM, sn_c + lc * GPT_MAX_DIM, p.idx[q]);
                }
            }
        }
    }
    const int warp = tid >> 5, lane = tid & 31;
    for (int q = 0; q < p.nidx; q++) {
        const double s = warp_sum(acc[q]);
        if (lane == 0) red[warp][q] = s;
    }
    __syncthreads();
    if (tid < p.nidx) {
        double s = 0.0;
        for (int w = 0; w < 8; w++) s += red[w][tid];
        p.partials[(long)cta * p.nidx + tid] = s;
    }
}

__global__ void grad_final_kernel(const double* __restrict__ partials, int nctas, int nidx, double* __restrict__ out) {
    // fixed-order (deterministic) sum of the per-CTA partials
    const int q = blockIdx.x;
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < nctas; i += 256) s += partials[(long)i * nidx + q];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[q] = sh[0];
}

__global__ void trace_sumsq_kernel(const double* __restrict__ A, long lda, const double* __restrict__ v, int n,
                                   double* __restrict__ out) {
    __shared__ double sh[2][256];
    double tr = 0.0, ss = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        tr += A[(long)i * lda + i];
        ss += v[i] * v[i];
    }
    sh[0][threadIdx.x] = tr;
    sh[1][threadIdx.x] = ss;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = sh[0][0];
        out[1] = sh[1][0];
    }
}

// out[0] = -1/2 z^T z - sum_k logdet_k - M/2 log(2 pi)   (gaussian_process.py:1463-1468), one CTA, fixed summation order
__global__ void ll_reduce_kernel(const double* __restrict__ z, int M, const double* __restrict__ logdet, int nblk,
                                 double* __restrict__ out) {
    __shared__ double sh[2][256];
    double zz = 0.0, ld = 0.0;
    for (int i = threadIdx.x; i < M; i += 256) zz += z[i] * z[i];
    for (int i = threadIdx.x; i < nblk; i += 256) ld += logdet[i];
    sh[0][threadIdx.x] = zz;
    sh[1][threadIdx.x] = ld;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = -0.5 * sh[0][0] - sh[1][0] - 0.5 * M * 1.8378770664093453;  // log(2 pi)
}

}  // namespace

void launch_ll_reduce(const double* z, int M, const double* logdet, int nblk, double* out, cudaStream_t s) {
    ll_reduce_kernel<<<1, 256, 0, s>>>(z, M, logdet, nblk, out);
}

void launch_potrf_diag(double* Ablk, long lda, double* inv, double* yk, double* logdet_part, int* info, int row0,
                       const double* Pprev, long ldp, int pcols, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM);
        attr_set = true;
    }
    potrf_diag_kernel<<<1, 256, POTRF_SMEM, s>>>(Ablk, lda, inv, yk, logdet_part, info, row0, Pprev, ldp, pcols / 16);
}

void launch_panel_trsm(double* A21, long lda, const double* L11, long ldl, const double* inv_k, double* out0, long ld0,
                       int split, double* out1, long ld1, int rows, const double* zk, double* y, cudaStream_t s) {
    if (rows <= 0) return;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(panel_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSM_SMEM);
        attr_set = true;
    }
    panel_trsm_kernel<<<(rows + TR_ROWS - 1) / TR_ROWS, 256, TRSM_SMEM, s>>>(A21, lda, L11, ldl, inv_k, out0, ld0, split, out1,
                                                                             ld1, rows, zk, y);
}

// ---- alpha = L^{-T} z in ONE launch: a chain of CTAs, one per 128-row block ----------------------------------------
// CTA i owns block b = bhi-i of the solution; launches cover at most one CTA per SM (see the launcher).  It folds
// z_b -= L[k, b]^T alpha_k for k = nblk-1 .. b+1 as the alpha_k are published, then alpha_b = Inv_b^T z_b.  The two
// operands on the critical path -- the tile L[b+1, b] needed last and Inv_b -- are fetched before the wait (into
// registers and shared memory), so a link of the chain costs a flag poll + ~130 FMAs per thread, not a kernel
// launch plus two cold 128 KB reads (the per-block launches took 29 us each: 15% of an M = 4000 factorisation).
constexpr size_t BSC_SMEM = ((size_t)NB * NB + 6 * NB) * sizeof(double);

__global__ void __launch_bounds__(256, 1) backsolve_chain_kernel(const double* __restrict__ L, long ld, int nblk, int bhi,
                                                                 const double* __restrict__ inv,
                                                                 const double* __restrict__ z, double* alpha,
                                                                 int* flags) {
    extern __shared__ __align__(16) double sm[];
    double* invs = sm;                 // Inv_b, NB x NB
    double* zs = sm + NB * NB;         // z_b
    double* ak = zs + NB;              // alpha_k
    double* part = ak + NB;            // 4 x NB partial sums
    const int tid = threadIdx.x;
    const int b = bhi - blockIdx.x;  // this launch covers blocks bhi, bhi-1, ... (one per CTA)
    const int c = tid & (NB - 1), hh = tid >> 7;  // column, half of the rows

    const double* invb = inv + (size_t)b * NB * NB;
    for (int idx = tid; idx < NB * NB / 2; idx += 256) cp_async16(invs + 2 * idx, invb + 2 * idx);
    cp_async_commit();
    if (tid < NB) zs[tid] = z[b * NB + tid];
    // the tile needed last, L[b+1, b]: rows hh*64 .. +63 of column c, in registers
    double last[64];
    if (b + 1 < nblk) {
        const double* tile = L + ((long)(b + 1) * NB + hh * 64) * ld + (long)b * NB + c;
#pragma unroll
        for (int r = 0; r < 64; r++) last[r] = tile[(long)r * ld];
    }
    __syncthreads();

    for (int k = nblk - 1; k > b; k--) {
        if (tid == 0) {
            while (atomicAdd(flags + k, 0) == 0) __nanosleep(64);
            __threadfence();
        }
        __syncthreads();
        if (tid < NB) ak[tid] = __ldcg(alpha + (long)k * NB + tid);
        __syncthreads();
        double s0 = 0.0, s1 = 0.0;
        if (k == b + 1) {
#pragma unroll
            for (int r = 0; r < 64; r += 2) {
                s0 = fma(last[r], ak[hh * 64 + r], s0);
                s1 = fma(last[r + 1], ak[hh * 64 + r + 1], s1);
            }
        } else {
            const double* tile = L + ((long)k * NB + hh * 64) * ld + (long)b * NB + c;
#pragma unroll 16
            for (int r = 0; r < 64; r += 2) {
                s0 = fma(tile[(long)r * ld], ak[hh * 64 + r], s0);
                s1 = fma(tile[(long)(r + 1) * ld], ak[hh * 64 + r + 1], s1);
            }
        }
        part[hh * NB + c] = s0 + s1;
        __syncthreads();
        if (tid < NB) zs[tid] -= part[tid] + part[NB + tid];
    }
    cp_async_wait<0>();
    __syncthreads();
    {   // alpha_b = Inv_b^T z_b: output c, rows i >= c only (Inv is lower triangular), split two ways
        double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
        for (int i = hh * 64; i < hh * 64 + 64; i += 2) {
            s0 = fma(invs[i * NB + c], zs[i], s0);
            s1 = fma(invs[(i + 1) * NB + c], zs[i + 1], s1);
        }
        part[hh * NB + c] = s0 + s1;
    }
    __syncthreads();
    if (tid < NB) alpha[(long)b * NB + tid] = part[tid] + part[NB + tid];
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicExch(flags + b, 1);
}

void launch_backsolve_chain(const double* L, long ld, int nblk, const double* inv, const double* z, double* alpha,
                            int* flags, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(backsolve_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BSC_SMEM);
        attr_set = true;
    }
    cudaMemsetAsync(flags, 0, sizeof(int) * nblk, s);
    // A CTA spins on flags published by CTAs of higher blocks.  One CTA fits per SM, so a launch never holds more CTAs
    // than there are SMs: everything a CTA waits for is either co-resident or finished in an earlier launch, and the
    // chain cannot deadlock whatever order the hardware dispatches blocks in.
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms < 1) num_sms = 1;
    }
    for (int bhi = nblk - 1; bhi >= 0; bhi -= num_sms) {
        const int cnt = (bhi + 1 < num_sms) ? bhi + 1 : num_sms;
        backsolve_chain_kernel<<<cnt, 256, BSC_SMEM, s>>>(L, ld, nblk, bhi, inv, z, alpha, flags);
    }
}

void launch_transpose(double* out, long ldo, const double* in, long ldi, int rows, int cols, cudaStream_t s) {
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    transpose_kernel<<<grid, block, 0, s>>>(out, ldo, in, ldi, rows, cols, 0, 0);
}

void launch_transpose_batched(double* out, long ldo, long stride_out, const double* in, long ldi, long stride_in,
                              int rows, int cols, int batch, cudaStream_t s) {
    if (batch <= 0 || rows <= 0 || cols <= 0) return;
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch), block(32, 8);
    transpose_kernel<<<grid, block, 0, s>>>(out, ldo, in, ldi, rows, cols, stride_out, stride_in);
}

void launch_copy2d(double* out, long ldo, const double* in, long ldi, int rows, int cols, cudaStream_t s) {
    if (rows <= 0 || cols <= 0) return;
    copy2d_kernel<<<(rows + 7) / 8, 256, 0, s>>>(out, ldo, in, ldi, rows, cols);
}

void launch_add_diag(double* A, long lda, const double* d, int n, cudaStream_t s) {
    add_diag_kernel<<<(n + 255) / 256, 256, 0, s>>>(A, lda, d, n);
}

void launch_fill(double* p, long n, double v, cudaStream_t s) {
    if (n <= 0) return;
    long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    fill_kernel<<<(int)blocks, 256, 0, s>>>(p, n, v);
}

void launch_set_identity_pad(double* A, long lda, int n_valid, int n_pad, cudaStream_t s) {
    if (n_pad <= n_valid) return;
    identity_pad_kernel<<<n_pad, 128, 0, s>>>(A, lda, n_valid, n_pad);
}

void launch_grad_reduce(const GradReduceParams& p, cudaStream_t s) {
    const int nt = (p.N + GT - 1) / GT;
    dim3 grid(nt, nt);
    grad_reduce_kernel<<<grid, 256, 0, s>>>(p);
    grad_final_kernel<<<p.nidx, 256, 0, s>>>(p.partials, nt * nt, p.nidx, p.out);
}

// d ll / d sigma_f for a kernel k = sigma_f^2 g without a closed-form pass: with K_tot = K + D (D diagonal),
//   1/2 (a' dK a - tr(K_tot^-1 dK)) = (sum_i [a_i (y_i - a_i D_i) + (K_tot^-1)_ii D_i] - n) / sigma_f,  D_i = d[i] + noise2
__global__ void sigma_identity_kernel(const double* __restrict__ Kinv, long ld, const double* __restrict__ a,
                                      const double* __restrict__ y, const double* __restrict__ d, double noise2, int n,
                                      double sigma_f, double* __restrict__ out) {
    __shared__ double sh[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const double Di = d[i] + noise2;
        acc += a[i] * (y[i] - a[i] * Di) + Kinv[(long)i * ld + i] * Di;
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (sigma_f != 0.0) ? (sh[0] - (double)n) / sigma_f : 0.0;
}

void launch_sigma_identity(const double* Kinv, long ld, const double* a, const double* y, const double* d, double noise2,
                           int n, double sigma_f, double* out, cudaStream_t s) {
    sigma_identity_kernel<<<1, 256, 0, s>>>(Kinv, ld, a, y, d, noise2, n, sigma_f, out);
}

void launch_trace_and_sumsq(const double* A, long lda, const double* v, int n, double* out, cudaStream_t s) {
    trace_sumsq_kernel<<<1, 256, 0, s>>>(A, lda, v, n, out);
}
