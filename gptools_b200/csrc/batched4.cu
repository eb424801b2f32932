// Batched ll + gradient (+ prediction): FOUR independent theta streams per SM.
//
// One persistent CTA per theta, one job loop over three tile sweeps (tools/tile_model.py is the numpy model of the tile
// recurrences): sweep 1 left-looking Cholesky with generated K tiles and inverted diagonal tiles, sweep 2 XT = L^{-T} in
// place, sweep 3 K^{-1} tiles contracted with regenerated dK tiles.  Cut for latency tolerance: the FP64 pipe -- which DMMA
// and scalar DFMA share on B200 (profiles/microbench/fp64_overlap.cu) -- was only ~63% busy in the first-generation kernel
// (256 threads, 2 CTAs/SM; git history, profiles/r01a_*) because each CTA spends half of its time in latency-bound
// non-GEMM phases.  Here a CTA is 4 warps / 128 threads with a ~39 KB footprint, so four CTAs (four thetas) share an SM:
//   * one 64x64 output tile per job, 32x32 warp tiles (16 DMMA.8x8x4 per k-step, 64 accumulator registers);
//   * operands stream through a ring of 64xBK chunks (A, B) filled by BULK ASYNC COPIES (cp.async.bulk, the TMA engine)
//     that complete on mbarriers: the workspace tiles are stored chunk-major and pre-swizzled (tix()), so a chunk is
//     one contiguous copy issued by a single lane; warps wait on the chunk's mbarrier, never on each other (the
//     warp that releases a slot last refills it), so the four warps of a CTA -- which sit on four different SM
//     sub-partitions -- drift by up to a ring's depth instead of meeting at a CTA barrier after every chunk;
//   * structural zeros are removed at compile-time granularity, never by predication inside a DMMA stream: diagonal
//     output tiles run a split job (two warps issue the 10 lower 8x8 products of their 16, the other two share the
//     off-diagonal block), triangular operands skip whole chunks and use a half-height chunk_mma instance in the first
//     chunk that reaches a warp's rows, the panel products follow the exact 8x8-block triangle of the inverse tile;
//   * no dedicated diagonal-tile buffer: the Gauss-Jordan sweep works in the staging tile, its 8x8 pivot blocks
//     distributed over the lanes of one warp; panel products take their B fragments from L2, one group ahead;
//   * nothing is rebuilt from re-read tiles: the residual y - sum L z, alpha = L^{-T} z and (prediction batches) the
//     predictive mean and |L^{-1} k*|^2 are accumulated from the panel tiles while they are still in registers;
//   * prediction batches (gpt_predict_batched): the test points are extra tile rows of sweep 1.
// Algorithmic work: M^3 flop per theta; roofline = FP64 tensor pipe (DESIGN.md sections 3, 4).
#include <stdio.h>
#include <stdlib.h>

#include "se_fast.cuh"

namespace {

using namespace sefast;

constexpr int TB = 64;
#ifndef GPT_B4_BK
#define GPT_B4_BK 16
#endif
#ifndef GPT_B4_STAGES
#define GPT_B4_STAGES 2  // BK x STAGES = 32 columns in flight = 32 KB per CTA (the staging tile aliases the ring)
#endif
constexpr int BK = GPT_B4_BK;
constexpr int CPT = TB / BK;           // chunks per 64-deep tile step
static_assert(BK == 8 || BK == 16, "chunk depth");
#ifndef GPT_B4_MINB
#define GPT_B4_MINB 4
#endif
#ifndef GPT_B4_UNROLL
#define GPT_B4_UNROLL 4  // entries in flight per thread in the short closed-form loops (8: +40% code, measured slower)
#endif
constexpr int STAGES = GPT_B4_STAGES;
constexpr int UNROLL = GPT_B4_UNROLL;
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
    unsigned long long full[STAGES];  // mbarriers: chunk landed (transaction count)
    unsigned int cnt[STAGES];         // warps done with the slot (mod 4); the last one refills it
    double piv[8];  // product of the 8 pivots of each 8x8 pivot block of the tile being factored
    double exptab[64];  // 2^(j/64), for exp_nonpos_tab
    __align__(16) double rk[TB];
    __align__(16) double zk[TB];  // read as double2 by acc_times_vec
    double red[4][GPT_MAX_PARAMS + 3];
    const double* a[MAXT];
    const double* b[MAXT];
    unsigned char flag[MAXT];  // bit0: A operand upper triangular, bit2: B operand upper triangular
    CovParams cp;
    double noise2;
    int theta;
    int info;
    int use_tab;  // this theta takes the short closed forms (SE, D <= 2, orders <= 1, every 1/l finite)
};

__device__ __forceinline__ double* slot(double* ws, int I, int J) { return ws + (size_t)(I * (I + 1) / 2 + J) * TILE; }
__device__ __forceinline__ double* slotDT(double* ws, int nT, int I) {
    return ws + (size_t)(nT * (nT + 1) / 2 + I) * TILE;
}

struct Lane {
    int tid, warp, lane, g, t, wr, wc;
};

// Row gi of the (extended) point arrays carries a point: a training row, or a test row of a prediction batch.
__device__ __forceinline__ bool row_ok(const BatchedParams& p, int gi) {
    return gi < p.M || (gi >= p.nT * TB && gi < p.nT * TB + p.Ms);
}
// Tile (t, J) of the test rows of a prediction batch
__device__ __forceinline__ double* tslot(double* ws, const BatchedParams& p, int t, int J) {
    return ws + p.ts_off + ((size_t)t * p.nT + J) * TILE;
}

// ---- workspace tile layout: chunk-major, pre-swizzled -------------------------------------------------------------
// A 64x64 tile is stored as CPT chunks of 64 rows x BK columns; inside a chunk row r holds its BK columns with the
// XOR swizzle that makes the DMMA fragment reads conflict-free.  A chunk is therefore ONE contiguous block of
// CHUNK doubles in global memory and lands in the ring exactly as the fragment loader wants it.
__device__ __forceinline__ int swz(int r) { return (BK == 16) ? ((r & 3) << 2) : (((r >> 1) & 1) << 2); }
__device__ __forceinline__ int tix(int r, int c) { return (c / BK) * CHUNK + r * BK + ((c % BK) ^ swz(r)); }

// one lane: chunk q of the current job -> ring slot q % STAGES
__device__ __forceinline__ void issue_chunk(Smem& sm, int q) {
    const int s = q / CPT, kc = q % CPT, slt = q % STAGES;
    double* base = sm.R + slt * STAGE_D;
    mbar_expect_tx(&sm.full[slt], 2 * CHUNK * sizeof(double));
#ifdef GPT_B4_NO_L2_HINTS
    const unsigned long long pa = l2_policy_evict_first(), pb = pa;
#else
    const unsigned long long pa = l2_policy_evict_first(), pb = l2_policy_evict_last();
#endif
    bulk_g2s(base, sm.a[s] + kc * CHUNK, CHUNK * sizeof(double), &sm.full[slt], pa);
    bulk_g2s(base + CHUNK, sm.b[s] + kc * CHUNK, CHUNK * sizeof(double), &sm.full[slt], pb);
}

// acc += A[arow.., :] * B[brow.., :]^T over one chunk; LOWER: only the 8x8 products on and below the block diagonal
// TRI: the A operand is upper triangular and this chunk's 16 columns start at the first row of block IMAX - 2, so the
// last block (IMAX - 1, eight rows further down) only meets the chunk's upper eight columns (kk >= 2).
template <bool LOWER, int IMAX = 4, bool TRI = false>
__device__ __forceinline__ void chunk_mma(const double* stage, int arow, int brow, const Lane& L, double (&acc)[4][4][2]) {
    const double* aS = stage + (arow + L.g) * BK;
    const double* bS = stage + CHUNK + (brow + L.g) * BK;
    const int sw = swz(L.g);
#pragma unroll
    for (int kk = 0; kk < BK / 4; kk++) {
        const int col = ((kk << 2) ^ sw) + L.t;
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < IMAX; i++) a[i] = aS[i * 8 * BK + col];
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = bS[j * 8 * BK + col];
#pragma unroll
        for (int i = 0; i < IMAX; i++) {
            if (TRI && i == IMAX - 1 && kk < 2) continue;
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (!LOWER || j <= i) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
}

// acc += A[s] * B[s]^T over the steps of the job table.  DIAG (the output is a diagonal tile, only its lower
// triangle is consumed): warps 0 and 3 own the diagonal 32x32 blocks (lower 8x8 products only), warps 1 and 2 share
// block (1,0), taking alternate chunks -- warp 1's partial sums end up in the (unused) upper-right block of the
// staging tile and are folded in by st_lower().
__device__ __forceinline__ void run_job(Smem& sm, const Lane& L, int nsteps, const bool DIAG, double (&acc)[4][4][2],
                                        uint32_t& phase) {
    const int total = nsteps * CPT;
    fence_proxy_async();
    __syncthreads();  // table visible, ring free (its last use as staging tile is over)
    if (L.tid == 0) {
        const int pre = total < STAGES ? total : STAGES;
        for (int q = 0; q < pre; q++) issue_chunk(sm, q);
    }
    const bool offd = DIAG && (L.warp == 1 || L.warp == 2);
    const int arow = (DIAG ? (L.warp != 0) : L.wr) * 32;  // block row of the A operand / of the output
    const int brow = (DIAG ? (L.warp == 3) : L.wc) * 32;
    for (int q = 0; q < total; q++) {
        const int slt = q % STAGES;
        mbar_wait(&sm.full[slt], (phase >> slt) & 1u);
        phase ^= 1u << slt;
        const int fl = sm.flag[q / CPT];
        // structural zeros: operand upper triangular (bit0: A, bit2: B) => rows >= 32 vanish for k < 32
        const bool zero = ((q % CPT) < CPT / 2) && (((fl & 1) && arow) || ((fl & 4) && brow));
        const bool mine = !offd || ((q & 1) == (L.warp & 1));
        if (!zero && mine) {
            const double* stage = sm.R + slt * STAGE_D;
            if (DIAG && !offd) chunk_mma<true>(stage, arow, brow, L, acc);
            // A upper triangular (A[r][k] = 0 for k < r): in the first chunk that reaches a warp's rows (chunk 0 for
            // rows 0..31, chunk 2 for rows 32..63) only its first 16 rows meet non-zero columns, and in that chunk and
            // the next the last 8-row block only the upper eight columns: the exact 8x8-block triangle, at compile time
            else if ((fl & 1) && (q % CPT) == (arow ? CPT / 2 : 0)) chunk_mma<false, 2, true>(stage, arow, brow, L, acc);
            else if ((fl & 1) && (q % CPT) == (arow ? CPT / 2 : 0) + 1) chunk_mma<false, 4, true>(stage, arow, brow, L, acc);
            else chunk_mma<false>(stage, arow, brow, L, acc);
        }
        __syncwarp();
        if (L.lane == 0) {
            // release / acquire on the slot counter: the fragment reads of every warp happen-before the refill that the
            // last warp out issues (generic-proxy reads -> async-proxy write of the same bytes: proxy fence in between).
            // (An explicit "empty" mbarrier per slot on top of this -- the textbook producer/consumer pair -- costs 1 %
            // and leaves compute-sanitizer's racecheck report unchanged: profiles/r02_sanitizer.txt.)
            const unsigned old = atom_add_acq_rel_cta(&sm.cnt[slt], 1u);
            if ((old & 3u) == 3u && q + STAGES < total) {
                fence_proxy_async();
                issue_chunk(sm, q + STAGES);
            }
        }
    }
    __syncthreads();  // ring may now be reused as staging
}

// Entry (r, c), c <= r, of the lower triangle that a DIAG job left in the staging tile.
__device__ __forceinline__ double st_lower(const double* St, int r, int c) {
    double s = St[r * LDT + c];
    if (r >= 32 && c < 32) s += St[(r - 32) * LDT + c + 32];
    return s;
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

// tile = +-acc: the sign flip is an integer operation on the high word (a multiplication by -1 would be 32 more FP64
// instructions per thread waiting for the shared FP64 pipe)
__device__ __forceinline__ double flip_sign(double x, int flip) {
    return __hiloint2double(__double2hiint(x) ^ flip, __double2loint(x));
}
__device__ __forceinline__ void acc_to_global(double* tile, const Lane& L, const double (&acc)[4][4][2], bool negate) {
    const int flip = negate ? (int)0x80000000 : 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double2 v = make_double2(flip_sign(acc[i][j][0], flip), flip_sign(acc[i][j][1], flip));
            *reinterpret_cast<double2*>(tile + tix(L.wr * 32 + i * 8 + L.g, L.wc * 32 + j * 8 + 2 * L.t)) = v;
        }
}

// Row sums of (accumulator tile) x v over this warp's 32 columns -> part[wc * 64 + row]; v: 64 doubles in shared memory.
// The two column halves are added by the caller after a CTA barrier (fixed order: deterministic).
__device__ __forceinline__ void acc_times_vec(const double (&out)[4][4][2], const double* v, const Lane& L, double* part) {
    double2 vv[4];
#pragma unroll
    for (int j = 0; j < 4; j++) vv[j] = *reinterpret_cast<const double2*>(v + L.wc * 32 + j * 8 + 2 * L.t);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) s = fma(out[i][j][1], vv[j].y, fma(out[i][j][0], vv[j].x, s));
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (L.t == 0) part[L.wc * 64 + L.wr * 32 + i * 8 + L.g] = s;
    }
}

// Row sums of squares of the accumulator tile over this warp's 32 columns -> part[wc * 64 + row]
__device__ __forceinline__ void acc_row_sumsq(const double (&out)[4][4][2], const Lane& L, double* part) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) s = fma(out[i][j][1], out[i][j][1], fma(out[i][j][0], out[i][j][0], s));
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (L.t == 0) part[L.wc * 64 + L.wr * 32 + i * 8 + L.g] = s;
    }
}

// out = St * Binv^T, St in shared (stride LDT), Binv a lower-triangular 64x64 tile in GLOBAL memory (tix layout).
// The B fragments are fetched straight from L2 (no shared buffer), eight columns per group, ONE GROUP AHEAD of the
// products that consume them (the loads of a group used to sit exposed in front of its DMMAs: the panel products ran
// at a third of the efficiency of the ring-fed jobs).  Binv[n][c] = 0 for c > n: the four groups that cross the warp's
// 32 rows of B use compile-time row ranges (8-row block j of B meets column group s only for j >= s).
template <int JMIN>
__device__ __forceinline__ void panel_load_b(double (&b)[2][4], const double* bG, int kcol, int sw) {
#pragma unroll
    for (int kk = 0; kk < 2; kk++)
#pragma unroll
        for (int j = JMIN; j < 4; j++)
            b[kk][j] = bG[((kcol + kk * 4) / BK) * CHUNK + j * 8 * BK + (((kcol + kk * 4) % BK) ^ sw)];
}
template <int JMIN>
__device__ __forceinline__ void panel_mma(double (&out)[4][4][2], const double* aS, int kcol, const double (&b)[2][4]) {
#pragma unroll
    for (int kk = 0; kk < 2; kk++) {
        double a[4];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = aS[i * 8 * LDT + kcol + kk * 4];
#pragma unroll
        for (int j = JMIN; j < 4; j++) {
#pragma unroll
            for (int i = 0; i < 4; i++) dmma884(out[i][j][0], out[i][j][1], a[i], b[kk][j]);
        }
    }
}
__device__ __forceinline__ void mult_lower_global(const double* St, const double* Binv, const Lane& L,
                                                  double (&out)[4][4][2]) {
    zero_acc(out);
    const double* aS = St + (L.wr * 32 + L.g) * LDT + L.t;
    const double* bG = Binv + (L.wc * 32 + L.g) * BK + L.t;
    const int sw = swz(L.g);
    const int base = L.wc * 32;  // first column of the triangular part of this warp's rows of B
    double bc[2][4], bn[2][4];
    panel_load_b<0>(bc, bG, 0, sw);
    if (L.wc) {  // columns 0..31: dense for rows 32..63
#pragma unroll 1
        for (int kcol = 0; kcol < 32; kcol += 8) {
            panel_load_b<0>(bn, bG, kcol + 8, sw);  // the last one is group 0 of the triangular part
            panel_mma<0>(out, aS, kcol, bc);
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
#pragma unroll
                for (int j2 = 0; j2 < 4; j2++) bc[kk][j2] = bn[kk][j2];
        }
    }
    panel_load_b<1>(bn, bG, base + 8, sw);
    panel_mma<0>(out, aS, base, bc);
    panel_load_b<2>(bc, bG, base + 16, sw);
    panel_mma<1>(out, aS, base + 8, bn);
    panel_load_b<3>(bn, bG, base + 24, sw);
    panel_mma<2>(out, aS, base + 16, bc);
    panel_mma<3>(out, aS, base + 24, bn);
}

// sum_{c < 32} tile(r, c0 + c) * v[c] for a workspace tile (tix layout), c0 a multiple of 32; 16-byte loads
__device__ __forceinline__ double row_dot32(const double* tile, int r, int c0, const double* v) {
    const int sw = swz(r);
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
        const double2 t2 = *reinterpret_cast<const double2*>(tile + ((c0 + c) / BK) * CHUNK + r * BK + (((c0 + c) % BK) ^ sw));
        s += t2.x * v[c];
        s += t2.y * v[c + 1];
    }
    return s;
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

// Pivot block: P (8x8 SPD, lower valid) -> chol(P)^{-1} (lower, zeros above) by ONE warp, in place: an LDL^T
// Gauss-Jordan sweep with the block distributed over the lanes like a DMMA C fragment (lane = 4 row + q holds columns
// 2q, 2q+1 of its row), pivot row / column broadcast by shuffles.  Every FP64 instruction of this serial chain waits
// for a slot between the DMMAs of the co-resident CTAs, so what counts is their number: ~70 here, against ~290 when
// every lane swept the whole block redundantly.  Returns the product of the pivots.
__device__ __forceinline__ double diag8(Smem& sm, double* P, int lane, int row0) {
    const int r = lane >> 2, q = lane & 3;
    const int c0 = 2 * q, c1 = c0 + 1;
    const double2 pv = *reinterpret_cast<const double2*>(P + r * LDT + c0);
    double p0 = pv.x, p1 = pv.y;  // entries above the diagonal (c > r) are never read by the sweep
    double dprod = 1.0, myd = 1.0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const double pub = (j & 1) ? p1 : p0;                               // this lane's entry of column pair j / 2
        const double rj0 = __shfl_sync(0xffffffffu, p0, j * 4 + q);         // p[j][c0]
        const double rj1 = __shfl_sync(0xffffffffu, p1, j * 4 + q);         // p[j][c1]
        double d = __shfl_sync(0xffffffffu, pub, j * 4 + (j >> 1));         // p[j][j]
        const double colr = __shfl_sync(0xffffffffu, pub, r * 4 + (j >> 1));   // p[r][j]
        const double cc0 = __shfl_sync(0xffffffffu, pub, c0 * 4 + (j >> 1));   // p[c0][j]
        const double cc1 = __shfl_sync(0xffffffffu, pub, c1 * 4 + (j >> 1));   // p[c1][j]
        if (!(d > 0.0)) {
            if (lane == 0 && sm.info == 0) sm.info = row0 + j + 1;
            d = 1.0;
        }
        myd = (r == j) ? d : myd;
        dprod *= d;
        const double rinv = fast_rcp_pos(d);
        const double w0 = (c0 > j) ? cc0 : ((c0 < j) ? rj0 : 0.0);
        const double w1 = (c1 > j) ? cc1 : ((c1 < j) ? rj1 : 0.0);
        if (r > j) {
            const double mult = colr * rinv;
            p0 = (c0 == j) ? -mult : fma(-mult, w0, p0);
            p1 = (c1 == j) ? -mult : fma(-mult, w1, p1);
        }
    }
    const double rs = fast_rsqrt_pos(myd);
    double2 o;
    o.x = (c0 < r) ? p0 * rs : ((c0 == r) ? rs : 0.0);
    o.y = (c1 < r) ? p1 * rs : ((c1 == r) ? rs : 0.0);
    *reinterpret_cast<double2*>(P + r * LDT + c0) = o;
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
    if (row_ok(p, gi) && gj < p.M) {
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
            const bool ok = row_ok(p, gi);
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

// ---- short closed forms (SE, D <= 2, derivative orders <= 1, every l finite and non-zero) ---------------------------
// Scalar FP64 instructions are arbitrated one-for-one against the DMMAs of the co-resident CTAs, so each of them costs
// the issuing warp ~20 cycles: the closed forms are cut to the minimum.  The (-1)^{sum nj} sign and sigma^2 are hoisted
// per column, the common factor 1/l of the length-scale derivative is applied once per thread, and the gradient pass
// re-reads sigma^2 exp(-r^2/2) (eb, written by phase 1, prefetched into L2 while the tile's products run) instead of
// recomputing the exponential.  (A theta-independent table of the coordinate differences was measured and lost: its
// L2 reads stall the loops for longer than the two subtractions they replace, profiles/r02d_*.)
//   m = 0: f = 1,               g = il u            (u = tau^2 / l^2, il = 1 / l, f: value factor, g: d f / d l
//   m = 1: f = -tau il^2,       g = il f (u - 2)     up to the common exponential whose own l-derivative il u f is
//   m = 2: f = il^2 (u - 1),    g = il il^2 (2 + u (u - 5))          included in g)
// staging-tile entry (r, c) of a job's result: DIAG jobs leave part of block (1,0) in the upper-right block
__device__ __forceinline__ double st_get(const double* St, int r, int c, bool diag) {
    const double s = St[r * LDT + c];
    const bool fold = diag && (r >= 32) && (c < 32);
    const double e = St[fold ? (r - 32) * LDT + c + 32 : r * LDT + c];  // always a valid address: select, no branch
    return fold ? s + e : s;
}

template <int FD>
__device__ __forceinline__ void gen_ktot_tab(const Smem& sm, const BatchedParams& p, double* St, int tid, int I, int Jc,
                                             double* __restrict__ eb) {
    const bool DIAG = (I == Jc), EB = (eb != nullptr);
    const int c = tid & (TB - 1);
    const int gj = Jc * TB + c;
    const bool col_ok = gj < p.M;
    int nj[FD], sj = 0;
    double il2[FD], nil2[FD], nh[FD], xj[FD];
#pragma unroll
    for (int d = 0; d < FD; d++) {
        nj[d] = col_ok ? __ldg(p.n + (size_t)gj * FD + d) : 0;
        sj += nj[d];
        il2[d] = sm.cp.inv_l[d] * sm.cp.inv_l[d];
        nil2[d] = -il2[d];
        nh[d] = -0.5 * il2[d];
        xj[d] = col_ok ? __ldg(p.X + (size_t)gj * FD + d) : 0.0;
    }
    const double ssig2 = (sj & 1) ? -sm.cp.sig2 : sm.cp.sig2;
    const double dj = col_ok ? sm.noise2 + __ldg(p.diag + gj) : 0.0;
    const double* pts = sm.R + PTS_OFF;
    const int* ord = reinterpret_cast<const int*>(pts + PTS_ORD);
#pragma unroll UNROLL
    for (int u = 0; u < 32; u++) {
        const int r = (tid >> 6) + 2 * u;
        const int gi = I * TB + r;
        double tau[FD], q[FD];
#pragma unroll
        for (int d = 0; d < FD; d++) {
            tau[d] = pts[r * FD + d] - xj[d];
            q[d] = tau[d] * tau[d];
        }
        double arg = q[0] * nh[0];
#pragma unroll
        for (int d = 1; d < FD; d++) arg = fma(q[d], nh[d], arg);
        const double base = ssig2 * exp_nonpos_tab(arg, sm.exptab);
        if (EB) eb[tid + u * THREADS] = base;
        const int pk = ord[r];
        double v = base;
#pragma unroll
        for (int d = 0; d < FD; d++) {
            const int m = ((pk >> (8 * d)) & 255) + nj[d];
            const double f1 = tau[d] * nil2[d];
            const double f2 = il2[d] * fma(q[d], il2[d], -1.0);
            double f = (m == 1) ? f1 : 1.0;
            f = (m == 2) ? f2 : f;
            v *= f;
        }
        const double vd = v + dj;
        v = (gi == gj) ? vd : v;
        const double pad = (gi == gj) ? 1.0 : 0.0;
        v = (col_ok && row_ok(p, gi)) ? v : pad;
        const double out = v - st_get(St, r, c, DIAG);
        if (!DIAG || c <= r) St[r * LDT + c] = out;
    }
}


// grad_tab with fewer FP64 instructions per entry (each one waits for a slot between the DMMAs of the co-resident
// CTAs, so the contraction's time is its FP64 instruction count).  With u = tau^2 the per-dimension factors are
//   f = s (c0 + c2 u),  g = s (g0 + g2 u + g4 u^2),  s = tau for m = 1 and 1 otherwise,
//   m = 0: (1, 0 | 0, il2, 0)   m = 1: (-il2, 0 | 2 il2, -il2^2, 0)   m = 2: (-il2, il2^2 | 2 il2, -5 il2^2, il2^3)
// (g without its common factor 1 / l).  A thread's column fixes nj, the row order is 0 or 1, so the coefficients are two
// register-resident sets per dimension picked by selects (ALU pipe), the polynomials are 1 + 2 fused multiply-adds, and
// s_0 s_1 multiplies the weight once.  Diagonal tiles: the loop takes the lower triangle with the diagonal at full
// weight; the half weight of the diagonal entries and tr K^-1 are a per-thread correction after the loop (each thread
// owns at most one diagonal entry).
template <int FD>
__device__ __forceinline__ void grad_tab(const Smem& sm, const BatchedParams& p, const double* St, int tid, int I, int J,
                                         const double* __restrict__ avec, const double* __restrict__ eb,
                                         double (&gall)[1 + GPT_MAX_DIM], double& tr_kinv) {
    const bool DIAG = (I == J);
    const int c = tid & (TB - 1);
    const int gj = J * TB + c;
    if (gj >= p.M) return;
    const double aj = avec[gj];
    double xj[FD], c0[FD][2], c2B[FD], g0[FD][2], g2[FD][2], g4B[FD];
    int nj[FD];
#pragma unroll
    for (int d = 0; d < FD; d++) {
        nj[d] = __ldg(p.n + (size_t)gj * FD + d);
        xj[d] = __ldg(p.X + (size_t)gj * FD + d);
        const double il2 = sm.cp.inv_l[d] * sm.cp.inv_l[d];
        const double il4 = il2 * il2;
        const bool j1 = (nj[d] != 0);
        // row order 0: m = nj (0 or 1); row order 1: m = nj + 1 (1 or 2)
        c0[d][0] = j1 ? -il2 : 1.0;
        g0[d][0] = j1 ? 2.0 * il2 : 0.0;
        g2[d][0] = j1 ? -il4 : il2;
        c0[d][1] = -il2;
        c2B[d] = j1 ? il4 : 0.0;
        g0[d][1] = 2.0 * il2;
        g2[d][1] = j1 ? -5.0 * il4 : -il4;
        g4B[d] = j1 ? il4 * il2 : 0.0;
    }
    const double* pts = sm.R + PTS_OFF;
    const int* ord = reinterpret_cast<const int*>(pts + PTS_ORD);
    double wk = 0.0, gl[FD];
#pragma unroll
    for (int d = 0; d < FD; d++) gl[d] = 0.0;
    // contribution of entry (r, c) with weight w (already multiplied by the cached exponential)
    auto entry = [&](int r, double wb, double& awk, double (&agl)[FD]) {
        const int pk = ord[r];
        double F[FD], G[FD], tau[FD];
        bool odd[FD];
#pragma unroll
        for (int d = 0; d < FD; d++) {
            tau[d] = pts[r * FD + d] - xj[d];
            const double q = tau[d] * tau[d];
            const bool b = ((pk >> (8 * d)) & 255) != 0;
            odd[d] = (b != (nj[d] != 0));
            const double cc0 = b ? c0[d][1] : c0[d][0];
            const double cc2 = b ? c2B[d] : 0.0;
            const double gg0 = b ? g0[d][1] : g0[d][0];
            const double gg2 = b ? g2[d][1] : g2[d][0];
            const double gg4 = b ? g4B[d] : 0.0;
            F[d] = fma(cc2, q, cc0);
            G[d] = fma(fma(gg4, q, gg2), q, gg0);
        }
        if constexpr (FD == 1) {
            const double ws = odd[0] ? wb * tau[0] : wb;
            awk = fma(ws, F[0], awk);
            agl[0] = fma(ws, G[0], agl[0]);
        } else {
            const double pp = tau[0] * tau[1];
            double sel = odd[0] ? tau[0] : 1.0;
            sel = odd[1] ? (odd[0] ? pp : tau[1]) : sel;
            const double ws = wb * sel;
            const double t0 = ws * F[0];
            awk = fma(t0, F[1], awk);
            agl[1] = fma(t0, G[1], agl[1]);
            agl[0] = fma(ws * G[0], F[1], agl[0]);
        }
    };
    const int u0 = (DIAG && c >= 32) ? 16 : 0;  // warp-uniform: columns >= 32 of a diagonal tile start at row 32
#pragma unroll UNROLL
    for (int u = u0; u < 32; u++) {
        const int r = (tid >> 6) + 2 * u;
        const int gi = I * TB + r;
        const bool use = (gi < p.M) && (!DIAG || c <= r);
        const double kinv = st_get(St, r, (DIAG && c > r) ? r : c, DIAG);
        double w = fma(pts[PTS_ALPHA + r], aj, -kinv);
        w = use ? w : 0.0;
        entry(r, w * eb[tid + u * THREADS], wk, gl);
    }
    if (DIAG && ((c & 1) == (tid >> 6))) {  // this thread's diagonal entry (r = c): half weight, trace
        const int r = c;
        const double kinv = st_get(St, r, c, true);
        tr_kinv += kinv;
        const double w = -0.5 * fma(pts[PTS_ALPHA + r], aj, -kinv);
        entry(r, w * eb[tid + ((c - (tid >> 6)) >> 1) * THREADS], wk, gl);
    }
#pragma unroll
    for (int d = 0; d < FD; d++) gall[1 + d] += gl[d] * sm.cp.inv_l[d];
    gall[0] += (sm.cp.p[0] != 0.0) ? 2.0 * wk / sm.cp.p[0] : 0.0;
}

// ---- Matern 5/2 short forms (KIND == 1: D <= 2, at most one first derivative per point) ---------------------------------
// With il2_d = 1 / l_d^2, tau = x_i - x_j, r2 = sum tau_d^2 il2_d, s = sqrt(5 r2), e = exp(-s), Bf = (1 + s) e
// (reference kernel/src/matern.c:61-186; the same closed forms as matern52_cov / matern52_cov_dual in covfn.cuh):
//   value            k00 = sig2 (1 + s + 5/3 r2) e
//   row deriv a      k10 = -5/3 sig2 Bf tau_a il2_a              col deriv b   k01 = +5/3 sig2 Bf tau_b il2_b
//   both             k11 = 5/3 sig2 (Bf [a == b] il2_a - 5 e tau_a tau_b il2_a il2_b)
// and their derivatives with respect to l_d through dr2/dl_d = -2 tau_d^2 il2_d / l_d, d il2_d / dl_d = -2 il2_d / l_d,
// d(1 + s + 5/3 r2)e / dr2 = -5/6 Bf, dBf/dr2 = -5/2 e, de/dr2 = -5 e / (2 s).  sigma_f: trace identity at the outputs.
template <int FD>
struct M52Entry {
    double e, Bf, s, r2, tau[FD];
};
template <int FD>
__device__ __forceinline__ M52Entry<FD> m52_entry(const double* pts, int r, const double (&xj)[FD], const double (&il2)[FD],
                                                   const double* etab) {
    M52Entry<FD> t;
    t.r2 = 0.0;
#pragma unroll
    for (int d = 0; d < FD; d++) {
        t.tau[d] = pts[r * FD + d] - xj[d];
        t.r2 = fma(t.tau[d] * t.tau[d], il2[d], t.r2);
    }
    const double q = 5.0 * t.r2;
    t.s = (q > 0.0) ? q * fast_rsqrt_pos(q) : 0.0;
    t.e = exp_nonpos_tab(-t.s, etab);
    t.Bf = (1.0 + t.s) * t.e;
    return t;
}

template <int FD>
__device__ __forceinline__ void m52_gen(const Smem& sm, const BatchedParams& p, double* St, int tid, int I, int Jc) {
    const bool DIAG = (I == Jc);
    const int c = tid & (TB - 1);
    const int gj = Jc * TB + c;
    const bool col_ok = gj < p.M;
    double il2[FD], xj[FD];
    int bcol = -1;  // dimension of the column point's derivative
#pragma unroll
    for (int d = 0; d < FD; d++) {
        il2[d] = sm.cp.inv_l[d] * sm.cp.inv_l[d];
        xj[d] = col_ok ? __ldg(p.X + (size_t)gj * FD + d) : 0.0;
        if (col_ok && __ldg(p.n + (size_t)gj * FD + d) != 0) bcol = d;
    }
    const double dj = col_ok ? sm.noise2 + __ldg(p.diag + gj) : 0.0;
    const double c53 = 1.6666666666666667 * sm.cp.sig2;
    const double* pts = sm.R + PTS_OFF;
    const int* ord = reinterpret_cast<const int*>(pts + PTS_ORD);
#pragma unroll UNROLL
    for (int u = 0; u < 32; u++) {
        const int r = (tid >> 6) + 2 * u;
        const int gi = I * TB + r;
        const M52Entry<FD> t = m52_entry<FD>(pts, r, xj, il2, sm.exptab);
        const int pk = ord[r];
        int arow = -1;
#pragma unroll
        for (int d = 0; d < FD; d++)
            if ((pk >> (8 * d)) & 255) arow = d;
        double ta = 0.0, ia = 0.0, tb = 0.0, ib = 0.0;
#pragma unroll
        for (int d = 0; d < FD; d++) {
            ta = (d == arow) ? t.tau[d] : ta;
            ia = (d == arow) ? il2[d] : ia;
            tb = (d == bcol) ? t.tau[d] : tb;
            ib = (d == bcol) ? il2[d] : ib;
        }
        const double v00 = sm.cp.sig2 * fma(1.6666666666666667, t.r2, 1.0 + t.s) * t.e;
        const double v10 = -c53 * t.Bf * ta * ia;
        const double v01 = c53 * t.Bf * tb * ib;
        const double v11 = c53 * (((arow == bcol) ? t.Bf * ia : 0.0) - 5.0 * t.e * (ta * ia) * (tb * ib));
        double v = (arow < 0) ? ((bcol < 0) ? v00 : v01) : ((bcol < 0) ? v10 : v11);
        const double vd = v + dj;
        v = (gi == gj) ? vd : v;
        const double pad = (gi == gj) ? 1.0 : 0.0;
        v = (col_ok && row_ok(p, gi)) ? v : pad;
        const double out = v - st_get(St, r, c, DIAG);
        if (!DIAG || c <= r) St[r * LDT + c] = out;
    }
}

template <int FD>
__device__ __forceinline__ void m52_grad(const Smem& sm, const BatchedParams& p, const double* St, int tid, int I, int J,
                                         const double* __restrict__ avec, double (&gall)[1 + GPT_MAX_DIM], double& tr_kinv) {
    const bool DIAG = (I == J);
    const int c = tid & (TB - 1);
    const int gj = J * TB + c;
    if (gj >= p.M) return;
    const double aj = avec[gj];
    double il2[FD], xj[FD];
    int bcol = -1;
#pragma unroll
    for (int d = 0; d < FD; d++) {
        il2[d] = sm.cp.inv_l[d] * sm.cp.inv_l[d];
        xj[d] = __ldg(p.X + (size_t)gj * FD + d);
        if (__ldg(p.n + (size_t)gj * FD + d) != 0) bcol = d;
    }
    const double c53 = 1.6666666666666667 * sm.cp.sig2;
    const double* pts = sm.R + PTS_OFF;
    const int* ord = reinterpret_cast<const int*>(pts + PTS_ORD);
    double gl[FD];
#pragma unroll
    for (int d = 0; d < FD; d++) gl[d] = 0.0;
    const int u0 = (DIAG && c >= 32) ? 16 : 0;
#pragma unroll 2
    for (int u = u0; u < 32; u++) {
        const int r = (tid >> 6) + 2 * u;
        const int gi = I * TB + r;
        const bool use = (gi < p.M) && (!DIAG || c <= r);
        const bool on_diag = DIAG && (c == r);
        const double kinv = st_get(St, r, (DIAG && c > r) ? r : c, DIAG);
        double w = fma(pts[PTS_ALPHA + r], aj, -kinv);
        w = on_diag ? 0.5 * w : w;
        w = use ? w : 0.0;
        if (use && on_diag) {
            tr_kinv += kinv;
            gall[0] += kinv * (sm.noise2 + p.diag[gi]);  // sum_i K^-1_ii D_i for the sigma_f identity
        }
        const M52Entry<FD> t = m52_entry<FD>(pts, r, xj, il2, sm.exptab);
        const int pk = ord[r];
        int arow = -1;
#pragma unroll
        for (int d = 0; d < FD; d++)
            if ((pk >> (8 * d)) & 255) arow = d;
        double ta = 0.0, ia = 0.0, tb = 0.0, ib = 0.0;
#pragma unroll
        for (int d = 0; d < FD; d++) {
            ta = (d == arow) ? t.tau[d] : ta;
            ia = (d == arow) ? il2[d] : ia;
            tb = (d == bcol) ? t.tau[d] : tb;
            ib = (d == bcol) ? il2[d] : ib;
        }
        const double es = (t.s > 0.0) ? t.e / t.s : 0.0;   // e / s: multiplies tau_a tau_b tau_d^2 (-> 0 with s)
        const double wc = w * c53;
#pragma unroll
        for (int d = 0; d < FD; d++) {
            // g = dk/dl_d * l_d / (5/3 sig2)   (the common 1 / l_d is applied once at the end)
            const double q = t.tau[d] * t.tau[d] * il2[d];         // tau_d^2 / l_d^2 = -l_d/2 dr2/dl_d
            const double g00 = t.Bf * q;
            const double g10 = -(ta * ia) * fma(5.0 * t.e, q, (d == arow) ? -2.0 * t.Bf : 0.0);
            const double g01 = (tb * ib) * fma(5.0 * t.e, q, (d == bcol) ? -2.0 * t.Bf : 0.0);
            const double dab = (arow == bcol) ? ia * fma(5.0 * t.e, q, (d == arow) ? -2.0 * t.Bf : 0.0) : 0.0;
            const double cnt = ((d == arow) ? 1.0 : 0.0) + ((d == bcol) ? 1.0 : 0.0);
            const double g11 = dab - 5.0 * (ta * ia) * (tb * ib) * fma(5.0 * es, q, -2.0 * t.e * cnt);
            const double g = (arow < 0) ? ((bcol < 0) ? g00 : g01) : ((bcol < 0) ? g10 : g11);
            gl[d] = fma(wc, g, gl[d]);
        }
    }
#pragma unroll
    for (int d = 0; d < FD; d++) gall[1 + d] += gl[d] * sm.cp.inv_l[d];
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
            St[r * LDT + c] = ktot_entry(sm, p, I * TB + r, gj) - (diag_tile ? st_lower(St, r, c) : St[r * LDT + c]);
        }
    } else {
        SEHoist<FD> h = se_hoist<FD>(sm.cp);
        h.etab = sm.exptab;
        const bool col_ok = gj < p.M;
        const PointReg<FD> pj = load_point<FD>(p.X, p.n, col_ok ? gj : 0);
        const double dj = col_ok ? sm.noise2 + __ldg(p.diag + gj) : 0.0;
        const double* pts = sm.R + PTS_OFF;
        if (FD <= 2 && p.low_order) {
#pragma unroll 2
            for (int u = 0; u < 32; u++) {
                const int r = (tid >> 6) + 2 * u;
                const int gi = I * TB + r;
                const PointReg<FD> pi = staged_point<FD>(pts, r);
                double v = se_value_low<FD, true>(h, pi, pj);
                v = (gi == gj) ? v + dj : v;
                const double pad = (gi == gj) ? 1.0 : 0.0;
                v = (col_ok && row_ok(p, gi)) ? v : pad;
                if (!(diag_tile && c > r)) St[r * LDT + c] = v - (diag_tile ? st_lower(St, r, c) : St[r * LDT + c]);
            }
        } else {
#pragma unroll 2
            for (int u = 0; u < 32; u++) {
                const int r = (tid >> 6) + 2 * u;
                if (diag_tile && c > r) continue;
                const int gi = I * TB + r;
                double v;
                if (col_ok && row_ok(p, gi)) {
                    PointReg<FD> pi;
                    if constexpr (FD <= 2) pi = staged_point<FD>(pts, r);
                    else pi = load_point<FD>(p.X, p.n, gi);
                    v = se_value<FD>(h, pi, pj);
                    if (gi == gj) v += dj;
                } else {
                    v = (gi == gj) ? 1.0 : 0.0;
                }
                St[r * LDT + c] = v - (diag_tile ? st_lower(St, r, c) : St[r * LDT + c]);
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
        // k = sigma_f^2 g for every kernel: dk/dsigma_f = 2 k / sigma_f, and with K_tot = K + D (D = noise^2 + err^2 + jitter)
        //   1/2 (a' dK a - tr(K_tot^-1 dK)) = (a'(y - D a) - M + sum_i K_tot^-1_ii D_i) / sigma_f
        // -- no closed-form evaluation at all (the sums over i are taken at the outputs); a composite has no common factor
        const bool sig_id = (sm.cp.kid != GPT_KERNEL_SE) && (sm.cp.kid != GPT_KERNEL_COMPOSITE) && (want & 1u);
        if (sig_id) want &= ~1u;
#pragma unroll 1
        for (int u = 0; u < 32; u++) {
            const int r = (tid >> 6) + 2 * u;
            const int gi = I * TB + r;
            if (gi >= p.M || (diag_tile && c > r)) continue;
            const double kinv = diag_tile ? st_lower(St, r, c) : St[r * LDT + c];
            double w = avec[gi] * aj - kinv;
            if (diag_tile && c == r) {
                tr_kinv += kinv;
                if (sig_id) gall[0] += kinv * (sm.noise2 + p.diag[gi]);  // sum_i K^-1_ii D_i for the sigma_f identity
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
                        gall[q] += w * cov_eval(sm.cp, p.X + (size_t)gi * p.D, p.n + (size_t)gi * p.D,
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
#pragma unroll 2
            for (int u = 0; u < 32; u++) {
                const int r = (tid >> 6) + 2 * u;
                const int gi = I * TB + r;
                const bool use = (gi < p.M) && !(diag_tile && c > r);
                const bool on_diag = diag_tile && (c == r);
                const double kinv = (diag_tile && c <= r) ? st_lower(St, r, c) : St[r * LDT + c];
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
                const double kinv = diag_tile ? st_lower(St, r, c) : St[r * LDT + c];
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

// Operands of contraction step `st` of the job that produces tile (I, J) in sweep ph; returns the structure flags
// (bit0: A upper triangular, bit2: B upper triangular).
__device__ __forceinline__ int job_step(double* ws, const BatchedParams& p, int nT, int ph, int I, int J, int st,
                                        const double*& aa, const double*& bb) {
    int fl = 0;
    if (ph == 1) {  // S(I,k) = sum_{j<k} L(I,j) L(k,j)^T   (I >= nT: a test row block of a prediction batch)
        aa = (I < nT) ? slot(ws, I, st) : tslot(ws, p, I - nT, st);
        bb = slot(ws, J, st);
    } else if (ph == 2) {  // sum_{m=J}^{I-1} XT(J,m)-as-stored * L(I,m)^T
        const int m = J + st;
        aa = (m == J) ? slotDT(ws, nT, J) : slot(ws, m, J);
        bb = slot(ws, I, m);
        fl = (m == J) ? 1 : 0;
    } else {  // K^{-1}(I,J) = sum_{m>=I} XT(I,m) XT(J,m)^T
        const int m = I + st;
        if (m == I) {
            aa = slotDT(ws, nT, I);
            bb = (I == J) ? slotDT(ws, nT, J) : slot(ws, I, J);
            fl = 1 | ((I == J) ? 4 : 0);
        } else {
            aa = slot(ws, m, I);
            bb = slot(ws, m, J);
        }
    }
    return fl;
}

template <int FD, int KIND = 0>
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
    if (L.tid == 0) {
        for (int q = 0; q < STAGES; q++) {
            mbar_init(&sm.full[q], 1);
            sm.cnt[q] = 0u;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t ring_phase = 0u;  // parity of the next completion of every ring slot (same in all threads)

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
            if (p.kid == GPT_KERNEL_COMPOSITE) {  // the leaves of this theta live in the CTA's workspace
                CovComposite* cc = reinterpret_cast<CovComposite*>(ws + p.comp_off);
                comp_init(*cc, p.D, p.comp_nleaf, p.comp_kids, p.comp_nps, p.comp_nterms, p.comp_masks, th);
                sm.cp.comp = cc;
            }
            sm.noise2 = th[p.nparams] * th[p.nparams];
            int tab = (p.short_forms != 0) ? 1 : 0;
            for (int d = 0; d < p.D; d++)
                if (!(sm.cp.inv_l[d] <= 1.79769313486231570e+308)) tab = 0;  // l = 0 (or NaN): the guarded closed forms
            sm.use_tab = tab;
        }
        __syncthreads();
        const double* yb = p.y + (size_t)b * p.y_stride;
        // running residual r = y - sum_j L(., j) z_j: every panel tile subtracts its share as it is produced (below), so
        // the diagonal tile finds r_k complete instead of re-reading its block row from memory
        if (L.tid < TB) {
            for (int k = 0; k < nT; k++) {
                const int gi = k * TB + L.tid;
                rvec[gi] = (gi < p.M) ? yb[gi] : 0.0;
            }
            for (int k = 0; k < 2 * p.nTs; k++) ws[p.pv_off + (size_t)k * TB + L.tid] = 0.0;  // running mean, sum of squares
        }
        double logdet = 0.0, zz = 0.0;  // thread 0
        PT_DECL;

        double gall[1 + GPT_MAX_DIM];
#pragma unroll
        for (int q = 0; q < 1 + GPT_MAX_DIM; q++) gall[q] = 0.0;
        double tr_kinv = 0.0;
        const bool need_alpha = (p.nidx > 0) || (p.alpha_out != nullptr);
        // with a gradient request sweep 2 produces L^{-T} tile by tile: alpha = L^{-T} z accumulates in its epilogues
        const bool fused_alpha = (p.nidx > 0);

        // One job loop for the three tile sweeps, so that the operand ring, the panel product and the tile epilogues are
        // each instantiated ONCE: the four CTAs of an SM sit in different phases, and the instruction cache has to
        // hold all of them at the same time (a third of the kernel's code size was worth 8% of its run time).
        //   sweep 1: left-looking Cholesky, tile (I, J = k): outer k, inner I >= k
        //   sweep 2: XT = L^{-T} in place,  tile (I, J):      outer I, inner J < I
        //   sweep 3: K^{-1} tiles + gradient contraction:     outer J, inner I >= J
#pragma unroll 1
        for (int ph = 1; ph <= 3; ph++) {
            if (ph == 2) {
                __threadfence_block();
                if (need_alpha && !fused_alpha) {
                    PT_MARK(7);
                    // ======================= alpha = L^{-T} z (block back substitution) =======================
                    for (int i = L.tid; i < nT * TB; i += THREADS) rvec[i] = zvec[i];
                    __syncthreads();
                    for (int J = nT - 1; J >= 0; J--) {
                        {
                            const int a = L.tid >> 1, h2 = L.tid & 1;
                            double s = row_dot32(slotDT(ws, nT, J), a, h2 * 32, rvec + J * TB + h2 * 32);
                            s += __shfl_xor_sync(0xffffffffu, s, 1);
                            if (h2 == 0) {
                                sm.zk[a] = s;
                                avec[J * TB + a] = s;
                            }
                        }
                        __syncthreads();
                        for (int I = 0; I < J; I++) {
                            const int c = L.tid >> 1, h2 = L.tid & 1;
                            const double* tile = slot(ws, J, I);
                            double s = 0.0;
#pragma unroll
                            for (int r = 0; r < 32; r++) s += tile[tix(h2 * 32 + r, c)] * sm.zk[h2 * 32 + r];
                            s += __shfl_xor_sync(0xffffffffu, s, 1);
                            if (h2 == 0) rvec[I * TB + c] -= s;
                        }
                        __syncthreads();
                    }
                    if (p.alpha_out != nullptr)
                        for (int i = L.tid; i < p.M; i += THREADS) p.alpha_out[(size_t)b * p.M + i] = avec[i];
                    PT_MARK(5);
                }
                if (!(p.nidx > 0 && sm.info == 0)) break;
            }
            if (ph == 3) {
                // alpha is complete; its sweep-2 updates were reductions performed at L2: read them back past L1 and
                // store them again so that the plain loads of sweep 3 see them
                __threadfence_block();
                __syncthreads();
                if (L.tid < TB)
                    for (int k = 0; k < nT; k++) avec[k * TB + L.tid] = __ldcg(avec + k * TB + L.tid);
                __syncthreads();
                if (p.alpha_out != nullptr)
                    for (int i = L.tid; i < p.M; i += THREADS) p.alpha_out[(size_t)b * p.M + i] = avec[i];
            }
#pragma unroll 1
            for (int o = (ph == 2) ? 1 : 0; o < nT; o++) {
                // sweep 1 of a prediction batch: the test row blocks follow the training rows of every column
                const int i0 = (ph == 2) ? 0 : o, i1 = (ph == 2) ? o : ((ph == 1) ? nT + p.nTs : nT);
#pragma unroll 1
                for (int i = i0; i < i1; i++) {
                    const int I = (ph == 2) ? o : i, J = (ph == 2) ? i : o;
                    const bool diag = (I == J);  // never in sweep 2
                    const int nsteps = (ph == 1) ? J : ((ph == 2) ? I - J : nT - I);
                    if (L.tid < nsteps) {
                        const double *aa, *bb;
                        const int fl = job_step(ws, p, nT, ph, I, J, L.tid, aa, bb);
                        sm.a[L.tid] = aa;
                        sm.b[L.tid] = bb;
                        sm.flag[L.tid] = (unsigned char)fl;
                    }
                    if (ph == 2 && i == i0 && L.tid < TB) sm.zk[L.tid] = __ldcg(zvec + I * TB + L.tid);  // published by run_job's barrier
                    zero_acc(acc);
                    const double* ebt = ws + p.eb_off + (size_t)((I < nT ? I * (I + 1) / 2 : 0) + J) * TILE;
                    if (ph == 3 && sm.use_tab) {  // the cached exponentials of this tile: on their way to L2 while the products run
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(ebt + L.tid * 16));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(ebt + 2048 + L.tid * 16));
                    }
                    PT_MARK(7);
                    run_job(sm, L, nsteps, diag, acc, ring_phase);
                    PT_MARK(0);
                    acc_to_tile(St, L, acc);
                    if (ph != 2) stage_rows<FD>(sm, p, L, I, (ph == 3) ? avec : nullptr);
                    __syncthreads();
                    if (ph == 1) {
                        // C = K_tot - S in place (K_tot generated from the closed forms, never stored)
                        if constexpr (KIND == 1) {
                            m52_gen<FD>(sm, p, St, L.tid, I, J);
                        } else if constexpr (FD == 1 || FD == 2) {
                            if (sm.use_tab) gen_ktot_tab<FD>(sm, p, St, L.tid, I, J, (p.eb_off && I < nT) ? const_cast<double*>(ebt) : nullptr);
                            else gen_ktot_tile<FD>(sm, p, St, L.tid, I, J);
                        } else {
                            gen_ktot_tile<FD>(sm, p, St, L.tid, I, J);
                        }
                        PT_MARK(1);
                        if (diag) {
                            const int k = J;
                            // residual r_k = y_k - sum_j L(k,j) z_j: accumulated by the panel epilogues of the columns before
                            if (L.tid < TB) sm.rk[L.tid] = __ldcg(rvec + k * TB + L.tid);
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
                                    Dk[tix(r, c)] = St[r * LDT + c];
                                    DTk[tix(r, c)] = St[c * LDT + r];
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
                            if (fused_alpha && L.tid >= TB) {  // diagonal term of alpha_k = sum_I (L^{-1}(I,k))^T z_I
                                const int a = L.tid - TB;
                                double s = 0.0;
                                for (int c = a; c < TB; c++) s = fma(St[c * LDT + a], sm.zk[c], s);
                                avec[k * TB + a] = s;
                            }
                        } else {
                            __syncthreads();
                        }
                    }
                    if (ph == 2 || (ph == 1 && !diag)) {
                        // panel: L(I,k) = C(I,k) Inv_k^T   /   XT tile: -(sum) Inv_I^T
                        double out[4][4][2];
                        mult_lower_global(St, (ph == 1) ? slot(ws, J, J) : slot(ws, I, I), L, out);
                        // sweep 1: r_I -= L(I,k) z_k;  sweep 2: alpha_J += (L^{-1}(I,J))^T z_I  (the stored tile is -out)
                        // test row block: mean_t += L*(t,k) z_k,  sumsq_t += |L*(t,k)|^2 row by row
                        acc_times_vec(out, sm.zk, L, sm.R + PTS_OFF);
                        if (I >= nT) acc_row_sumsq(out, L, sm.R + PTS_OFF + 128);
                        acc_to_global((I < nT) ? slot(ws, I, J) : tslot(ws, p, I - nT, J), L, out, ph != 1);
                    }
                    if (ph != 3) {
                        __syncthreads();
                        if (!diag && L.tid < TB) {
                            const double* part = sm.R + PTS_OFF;
                            if (I < nT) {
                                double* target = (ph == 1) ? rvec + I * TB : avec + J * TB;
                                atomicAdd(target + L.tid, -(part[L.tid] + part[64 + L.tid]));  // one update per address and job
                            } else {
                                double* pm = ws + p.pv_off + (size_t)(I - nT) * TB;
                                atomicAdd(pm + L.tid, part[L.tid] + part[64 + L.tid]);
                                atomicAdd(pm + (size_t)p.nTs * TB + L.tid, part[128 + L.tid] + part[192 + L.tid]);
                            }
                        }
                        PT_MARK(3);
                    } else {
                        if constexpr (KIND == 1) {
                            m52_grad<FD>(sm, p, St, L.tid, I, J, avec, gall, tr_kinv);
                        } else if constexpr (FD == 1 || FD == 2) {
                            if (sm.use_tab) grad_tab<FD>(sm, p, St, L.tid, I, J, avec, ebt, gall, tr_kinv);
                            else grad_tile<FD>(sm, p, St, L.tid, I, J, avec, gall, tr_kinv);
                        } else {
                            grad_tile<FD>(sm, p, St, L.tid, I, J, avec, gall, tr_kinv);
                        }
                        PT_MARK(6);
                    }
                }
            }
        }

        // =========================== outputs ===========================
        __syncthreads();
        if (p.Ms > 0) {  // prediction batch: mean = L* z; |L*|^2 row by row (predict_var_finish_kernel adds the prior variance)
            const double* pm = ws + p.pv_off;
            for (int s2 = L.tid; s2 < p.Ms; s2 += THREADS) {
                p.pmean[(size_t)b * p.Ms + s2] = __ldcg(pm + s2);
                p.pvar[(size_t)b * p.Ms + s2] = __ldcg(pm + (size_t)p.nTs * TB + s2);
            }
        }
        if (p.nidx > 0) {
#pragma unroll
            for (int q = 0; q < 1 + GPT_MAX_DIM; q++) {
                const double s = warp_sum(gall[q]);
                if (L.lane == 0) sm.red[L.warp][q] = s;
            }
            const double s = warp_sum(tr_kinv);
            if (L.lane == 0) sm.red[L.warp][GPT_MAX_PARAMS] = s;
            __syncthreads();
            bool want_noise = false, want_sig = false;
            for (int q = 0; q < p.nidx; q++) {
                want_noise |= (p.idx[q] == p.nparams);
                want_sig |= (p.idx[q] == 0);
            }
            // the sigma_f identity of grad_tile (generic closed forms, single kernels)
            const bool sig_id = (FD == 0 || KIND == 1) && want_sig && sm.cp.kid != GPT_KERNEL_SE && sm.cp.kid != GPT_KERNEL_COMPOSITE;
            double aa = 0.0, ay = 0.0;
            if (want_noise || sig_id) {
                double part = 0.0, part2 = 0.0;
                for (int i = L.tid; i < p.M; i += THREADS) {
                    const double a = avec[i];
                    part += a * a;
                    part2 += a * (yb[i] - a * (sm.noise2 + p.diag[i]));
                }
                part = warp_sum(part);
                part2 = warp_sum(part2);
                if (L.lane == 0) {
                    sm.red[L.warp][GPT_MAX_PARAMS + 1] = part;
                    sm.red[L.warp][GPT_MAX_PARAMS + 2] = part2;
                }
                __syncthreads();
                for (int w = 0; w < 4; w++) {
                    aa += sm.red[w][GPT_MAX_PARAMS + 1];
                    ay += sm.red[w][GPT_MAX_PARAMS + 2];
                }
            }
            if (L.tid == 0) {
                double tr = 0.0;
                for (int w = 0; w < 4; w++) tr += sm.red[w][GPT_MAX_PARAMS];
                const double sn = p.thetas[(size_t)b * np1 + p.nparams];
                for (int q = 0; q < p.nidx; q++) {
                    double gsum = 0.0;
                    if (p.idx[q] == p.nparams) {
                        gsum = sn * (aa - tr);  // gaussian_process.py:1484-1488: noise kernel derivative 2 sigma_n I
                    } else if (sig_id && p.idx[q] == 0) {
                        double kd = 0.0;
                        for (int w = 0; w < 4; w++) kd += sm.red[w][0];
                        gsum = (sm.cp.p[0] != 0.0) ? (ay - (double)p.M + kd) / sm.cp.p[0] : 0.0;
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

// var[b][s] = k_theta_b(x*_s, x*_s) - |L^-1 k*_s|^2: the prior variance of every test point from the closed forms (kept out of
// the persistent kernel, whose SE instantiations would otherwise carry the generic closed forms and their stack frame).
// One CTA per theta.
__global__ void predict_var_finish_kernel(BatchedParams p) {
    __shared__ CovParams cp;
    __shared__ CovComposite comp;
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        const double* th = p.thetas + (size_t)b * (p.nparams + 1);
        cov_params_init(cp, p.kid, p.D, p.nparams, th);
        if (p.kid == GPT_KERNEL_COMPOSITE) {
            comp_init(comp, p.D, p.comp_nleaf, p.comp_kids, p.comp_nps, p.comp_nterms, p.comp_masks, th);
            cp.comp = &comp;
        }
    }
    __syncthreads();
    for (int s2 = threadIdx.x; s2 < p.Ms; s2 += blockDim.x) {
        const size_t gi = (size_t)p.nT * TB + s2;
        const double kss = cov_eval(cp, p.X + gi * p.D, p.n + gi * p.D, p.X + gi * p.D, p.n + gi * p.D, -1);
        p.pvar[(size_t)b * p.Ms + s2] = kss - p.pvar[(size_t)b * p.Ms + s2];
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

template <int FD, int KIND = 0>
void launch_t(const BatchedParams& p, int num_ctas, cudaStream_t s) {
    const size_t smem_bytes = smem_request();
    cudaFuncSetAttribute(ll_batched4_kernel<FD, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
#ifdef GPT_PHASE_TIMING
    {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ll_batched4_kernel<FD, KIND>, THREADS, sizeof(Smem));
        fprintf(stderr, "[batched4] smem %zu B per CTA, occupancy %d CTAs/SM, grid %d\n", sizeof(Smem), nb, num_ctas);
    }
#endif
    ll_batched4_kernel<FD, KIND><<<num_ctas, THREADS, smem_bytes, s>>>(p);
}

}  // namespace

size_t batched_lower_tiles(int nT) { return (size_t)nT * (nT + 1) / 2; }

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
    // Matern 5/2 in one or two dimensions with at most first derivatives: its own short closed forms
    else if (p.kid == GPT_KERNEL_MATERN52 && p.D == 1 && p.low_order && !getenv("GPT_B4_LONG_FORMS")) launch_t<1, 1>(p, num_ctas, s);
    else if (p.kid == GPT_KERNEL_MATERN52 && p.D == 2 && p.low_order && !getenv("GPT_B4_LONG_FORMS")) launch_t<2, 1>(p, num_ctas, s);
    else launch_t<0>(p, num_ctas, s);
    if (p.Ms > 0) predict_var_finish_kernel<<<p.B, 128, 0, s>>>(p);
}
