// ntm_quad.cuh -- short horizons (N <= 24) of the fused loop, literal Gamma, EC-power box QP: FOUR LANES PER SCENARIO.
//
// Round 1 gave every scenario a whole warp with lane j = horizon index j: at N = 20 twelve of 32 lanes idle (22 of 32 at
// N = 10), the triangular Toeplitz loop kept a third of the remaining lane-slots busy, and every reduction / scan paid
// five shuffle stages -- 1,510 warp instructions per re-linearisation, 67 % of the issue slots, 78 % of the
// shared-memory pipe (profiles/r01_closed_loop_v6_*).  Here a scenario is a QUAD: lane l of the quad owns the E
// consecutive horizon indices l*E .. l*E+E-1 (E = ceil(N / 4), a template parameter, so all per-index state is a
// register array and every loop over it is unrolled), a warp advances eight scenarios in lockstep, reductions are two
// shuffle stages, the stage-map scans are E serial steps + a 4-lane exchange, and the lags of the Toeplitz table are
// dealt to the lanes in a boustrophedon order (l, 7-l, 8+l, 15-l, ...) so every lane sums the same number of entries.
// The Hessian is kept as the packed lower triangle (210 doubles at N = 20): 3.4 KB per scenario, 64 scenarios per SM.
//
// Same arithmetic as the one-warp kernel where the order of operations is visible to the reference's bit-sensitive
// stop rule (NTM_MPC_Sim.m:123): G and F are the same running sums in the same order, the QP mat-vecs run k = 0..N-1,
// bound components are exactly umin / umax.  Quads desynchronise under the eps_break policy (1..i_sim inner iterations
// per step), so every quad carries its own (scenario, k, it) and pulls the next scenario from the queue by itself; the
// warp only shares the instruction stream ("pass" = condense, stop rule / plant, QP, rollout + re-scheduling).
#pragma once
#include "ntm_device.cuh"
#include "ntm_kernels.h"

namespace ntm {

#define NTM_QFULL 0xffffffffu
#define NTM_QHCAP 6          // free sets up to this size are factorised in shared memory, larger ones in the global slab

__device__ __forceinline__ double quad_sum(double v) {
    v += __shfl_xor_sync(NTM_QFULL, v, 1);
    v += __shfl_xor_sync(NTM_QFULL, v, 2);
    return v;
}
__device__ __forceinline__ int quad_sum(int v) {
    v += __shfl_xor_sync(NTM_QFULL, v, 1);
    v += __shfl_xor_sync(NTM_QFULL, v, 2);
    return v;
}
// lane-local best (v, idx) -> quad best; the lowest index wins ties; idx >= (1 << 20) means "none"
__device__ __forceinline__ void quad_argmin(double &v, int &idx) {
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
        const double ov = __shfl_xor_sync(NTM_QFULL, v, o);
        const int oi = __shfl_xor_sync(NTM_QFULL, idx, o);
        if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
}
__device__ __forceinline__ bool quad_all(bool p, int lane) {
    const unsigned b = __ballot_sync(NTM_QFULL, p);
    return ((b >> (lane & 28)) & 0xFu) == 0xFu;
}
__device__ __forceinline__ bool quad_any(bool p, int lane) {
    const unsigned b = __ballot_sync(NTM_QFULL, p);
    return ((b >> (lane & 28)) & 0xFu) != 0u;
}

__host__ __device__ __forceinline__ int qtri(int j) { return (j * (j + 1)) >> 1; }

// doubles of one quad's shared-memory region (even): G | params | Fx | scratch
__host__ __device__ inline int quad_region_doubles(int N) {
    const int g = (qtri(N) + 1) & ~1;
    const int scratch = 6 * N + 64;
    int r = g + 16 + ((N + 1) & ~1) + scratch;
    r = (r + 1) & ~1;
    if ((r & 15) != 4) r += ((4 - (r & 15)) + 16) & 15;      // region stride = 4 mod 16 doubles: spreads the quads over the banks
    return r;
}

struct QuadWork {
    double *G;            // packed lower triangle, row j at qtri(j)
    Params *prm;
    double *Fx;           // [N] gradient handed from the lag owners to the variable owners
    // scratch, build phase
    double2 *P12;         // [N]      p_d
    double2 *QP12;        // [N + 8]  2 Q p_d, zero beyond N
    double2 *QE12;        // [N + 8]  2 Q (v_i - r), zero beyond N
    // scratch, QP phase (aliases the build arrays)
    double2 *cand;        // [N] {solution before last, last solution}
    double2 *cand0;       // [N] {lb, ub}
    int2 *hst;            // [N] partition states of the two previous solutions (y-space)
    double *H;            // [NTM_QHCAP x NTM_QHCAP]
    int *idx;             // [N]
    double *uv, *sol;     // [N] each; ALIAS cand (the candidates are only read by the two start passes of the warp solver)
};

__device__ inline QuadWork quad_carve(unsigned char *base, int N) {
    QuadWork w;
    double *d = reinterpret_cast<double *>(base);
    w.G = d; d += (qtri(N) + 1) & ~1;
    w.prm = reinterpret_cast<Params *>(d); d += 16;
    w.Fx = d; d += (N + 1) & ~1;
    double2 *v = reinterpret_cast<double2 *>(d);
    w.P12 = v; w.QP12 = v + N; w.QE12 = v + 2 * N + 8;                  // 3N + 16 double2 = 6N + 32 doubles
    w.cand = v; w.cand0 = v + N;                                       // 4N doubles
    w.uv = d; w.sol = d + N;                                           // alias cand
    double *q = d + 4 * N;
    w.hst = reinterpret_cast<int2 *>(q); q += N;                       // 5N
    w.H = q; q += NTM_QHCAP * NTM_QHCAP;                               // 5N + 36
    w.idx = reinterpret_cast<int *>(q);                                // N ints: 5.5N + 36 <= 6N + 64
    return w;
}

// affine stage map z -> [a 0; c s] z + (k1, k2)
struct QAff { double a, c, s, k1, k2; };
__device__ __forceinline__ QAff qaff_compose(const QAff &L, const QAff &E) {      // L after E
    QAff r;
    r.a = L.a * E.a;
    r.c = fma(L.c, E.a, L.s * E.c);
    r.s = L.s * E.s;
    r.k1 = fma(L.a, E.k1, L.k1);
    r.k2 = fma(L.c, E.k1, fma(L.s, E.k2, L.k2));
    return r;
}
__device__ __forceinline__ QAff qaff_shfl_up(const QAff &m, int off) {
    QAff r;
    r.a = __shfl_up_sync(NTM_QFULL, m.a, off, 4);
    r.c = __shfl_up_sync(NTM_QFULL, m.c, off, 4);
    r.s = __shfl_up_sync(NTM_QFULL, m.s, off, 4);
    r.k1 = __shfl_up_sync(NTM_QFULL, m.k1, off, 4);
    r.k2 = __shfl_up_sync(NTM_QFULL, m.k2, off, 4);
    return r;
}
// exclusive prefix over the four lanes of a quad of the lanes' total maps (lane 0: identity)
__device__ __forceinline__ QAff quad_exclusive(QAff t, int l) {
    QAff e = qaff_shfl_up(t, 1);
    if (l >= 1) t = qaff_compose(t, e);
    e = qaff_shfl_up(t, 2);
    if (l >= 2) t = qaff_compose(t, e);
    e = qaff_shfl_up(t, 1);
    if (l == 0) { e.a = 1.0; e.c = 0.0; e.s = 1.0; e.k1 = 0.0; e.k2 = 0.0; }
    return e;
}

// G(j, k) of the packed symmetric Hessian
__device__ __forceinline__ double qG(const double *__restrict__ G, int j, int k) {
    return (k <= j) ? G[qtri(j) + k] : G[qtri(k) + j];
}

// lane 0 of the quad: m x m SPD solve H x = rhs (LDL', in place), H row-major with pitch ldh.  Returns false on a
// non-positive pivot.
__device__ inline bool quad_ldl_solve(int m, double *__restrict__ H, int ldh, double *__restrict__ x) {
    bool ok = true;
    for (int k = 0; k < m; ++k) {
        const double d = H[k * ldh + k];
        if (!(d > 0.0)) { ok = false; break; }
        const double inv = 1.0 / d;
        for (int a = k + 1; a < m; ++a) {
            const double lak = H[a * ldh + k] * inv;
            for (int b = k + 1; b <= a; ++b) H[a * ldh + b] = fma(-lak, H[b * ldh + k], H[a * ldh + b]);
            H[k * ldh + a] = lak;
            x[a] = fma(-lak, x[k], x[a]);
        }
    }
    if (!ok) return false;
    for (int k = 0; k < m; ++k) x[k] = x[k] / H[k * ldh + k];
    for (int k = m - 1; k >= 1; --k) {
        const double xk = x[k];
        for (int a = 0; a < k; ++a) x[a] = fma(-H[a * ldh + k], xk, x[a]);
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// Slow path of the QP: the previous solutions are not optimal, so the active set has to be found.  Running that in
// lockstep over the eight quads of a warp makes all eight pay for the slowest one (measured: 87 % of the passes took the
// slow path although only ~16 % of the QPs need it, 1,124 instructions per re-linearisation); instead the whole WARP
// solves the quads that need it one after the other with the one-warp algorithm of qp_solve<1> (lane j = variable j:
// start = best of four candidates by objective, exact primal active set, LDL' of the free block by ldl_solve<1>), reading
// the quad's packed Hessian and candidate arrays in shared memory.  Returns the status; Uout / state_out per lane.
// ------------------------------------------------------------------------------------------------
__device__ inline int qp_solve_warp_packed(int N, int j, const QuadWork &w, double *hslab, double Fj, int hn, int max_iter,
                                           int &iters_out, double &Uout, int &state_out) {
    using Gp = Group<1>;
    const bool act = j < N;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const double2 bnd = act ? w.cand0[j] : make_double2(0.0, 0.0);
    const double2 prev = act ? w.cand[j] : make_double2(0.0, 0.0);
    const int2 hs = act ? w.hst[j] : make_int2(-1, -1);
    const double lbj = bnd.x, ubj = bnd.y;
    const bool pinned = !(ubj > lbj);
    const double cu2 = (hn >= 2) ? prev.x : lbj, cu1 = (hn >= 1) ? prev.y : lbj;
    const int s2h = hs.x, s1h = hs.y;
    double g2 = 0.0, g3 = 0.0, s2 = 0.0, s3 = 0.0;
    int state = -1;
    double u = lbj, g = Fj, sc = fabs(Fj);
    bool solved = false;
    if (hn >= 1) {
        if (act) {
            for (int k = 0; k < N; ++k) {
                const double gk = qG(w.G, j, k), ga = fabs(gk);
                const double2 cb = w.cand[k];
                const double c2 = (hn >= 2) ? cb.x : w.cand0[k].x;
                g2 = fma(gk, c2, g2); g3 = fma(gk, cb.y, g3);
                s2 = fma(ga, fabs(c2), s2); s3 = fma(ga, fabs(cb.y), s3);
            }
        }
        {   // last solution first
            const double t = g3 + Fj, sa = s3 + fabs(Fj);
            const bool ok = !act || pinned || (s1h < 0 && t >= -NTM_QP_EPS_G * sa) || (s1h > 0 && -t >= -NTM_QP_EPS_G * sa);
            if (Gp::all(ok, nullptr)) { solved = true; state = pinned ? -1 : s1h; u = cu1; g = t; sc = sa; }
        }
        if (!solved && hn >= 2) {
            const double t = g2 + Fj, sa = s2 + fabs(Fj);
            const bool ok = !act || pinned || (s2h < 0 && t >= -NTM_QP_EPS_G * sa) || (s2h > 0 && -t >= -NTM_QP_EPS_G * sa);
            if (Gp::all(ok, nullptr)) { solved = true; state = pinned ? -1 : s2h; u = cu2; g = t; sc = sa; }
        }
    }
    bool exact = true;
    if (!solved) {
        double g0 = 0.0, g1 = 0.0, s0 = 0.0, s1 = 0.0;
        if (act) {
            for (int k = 0; k < N; ++k) {
                const double gk = qG(w.G, j, k), ga = fabs(gk);
                const double2 ca = w.cand0[k];
                g0 = fma(gk, ca.x, g0); g1 = fma(gk, ca.y, g1);
                s0 = fma(ga, fabs(ca.x), s0); s1 = fma(ga, fabs(ca.y), s1);
            }
        }
        const double q0 = Gp::sum(act ? lbj * fma(0.5, g0, Fj) : 0.0, nullptr);
        const double q1 = Gp::sum(act ? ubj * fma(0.5, g1, Fj) : 0.0, nullptr);
        double qb = q0;
        state = -1; u = lbj; g = g0 + Fj; sc = s0;
        if (q1 < qb) { qb = q1; state = 1; u = ubj; g = g1 + Fj; sc = s1; }
        if (hn >= 2) {
            const double q2 = Gp::sum(act ? cu2 * fma(0.5, g2, Fj) : 0.0, nullptr);
            if (q2 < qb) { qb = q2; state = s2h; u = cu2; g = g2 + Fj; sc = s2; }
        }
        if (hn >= 1) {
            const double q3 = Gp::sum(act ? cu1 * fma(0.5, g3, Fj) : 0.0, nullptr);
            if (q3 < qb) { qb = q3; state = s1h; u = cu1; g = g3 + Fj; sc = s3; }
        }
        sc += fabs(Fj);
        if (pinned) { state = -1; u = lbj; }
    }
    __syncwarp();                                        // the candidate arrays are dead from here on: uv / sol alias them
    int status = solved ? NTM_SCN_OK : NTM_SCN_QP_ITER_CAP, it = 1;
    bool broke = false;
    for (it = 1; !solved && it <= max_iter; ++it) {
        const bool isfree = act && state == 0;
        int m;
        const int pos = Gp::prefix(isfree, nullptr, m);
        if (m > 0) {
            if (isfree) { w.idx[pos] = j; w.sol[pos] = -g; }
            __syncwarp();
            double *H = (m <= NTM_QHCAP) ? w.H : hslab;                   // big free sets go to the global slab
            const int ldh = (m <= NTM_QHCAP) ? NTM_QHCAP : N;
            if (j < m) {
                const int cb = w.idx[j];
                for (int aa = 0; aa < m; ++aa) H[aa * ldh + j] = qG(w.G, w.idx[aa], cb);
            }
            __syncwarp();
            broke |= ldl_solve<1>(m, j, H, ldh, w.sol);                   // sol[0..m) = Newton step on the face
            const double pj = isfree ? w.sol[pos] : 0.0;
            double aj = INF;
            if (isfree) {
                if (pj < 0.0) aj = (lbj - u) / pj;
                else if (pj > 0.0) aj = (ubj - u) / pj;
            }
            int jblk;
            const double amin = Gp::argmin(aj, j, nullptr, nullptr, jblk);
            const bool blocked = amin < 1.0;
            const double alpha = blocked ? fmax(amin, 0.0) : 1.0;
            if (isfree) u = fma(alpha, pj, u);
            exact = false;
            if (blocked) {                                                // a bound blocks: fix it, stay on the arc
                if (act) {
                    double dg = 0.0;
                    for (int aa = 0; aa < m; ++aa) dg = fma(qG(w.G, j, w.idx[aa]), w.sol[aa], dg);
                    g = fma(alpha, dg, g);
                }
                if (j == jblk) { state = (pj < 0.0) ? -1 : 1; u = (pj < 0.0) ? lbj : ubj; }
                __syncwarp();
                continue;
            }
        }
        if (!exact) {                                                     // exact gradient and its scale at the face minimiser
            if (act) w.uv[j] = u;
            __syncwarp();
            double t = Fj, sa = fabs(Fj);
            if (act) {
                for (int k = 0; k < N; ++k) {
                    const double gk = qG(w.G, j, k), uk_ = w.uv[k];
                    t = fma(gk, uk_, t);
                    sa = fma(fabs(gk), fabs(uk_), sa);
                }
            }
            g = t; sc = sa; exact = true;
        }
        double lam = INF;
        if (act && !pinned) {
            if (state < 0) lam = g / sc;
            else if (state > 0) lam = -g / sc;
        }
        int jw;
        const double lmin = Gp::argmin(lam, j, nullptr, nullptr, jw);
        if (!(lmin < -NTM_QP_EPS_G)) { status = NTM_SCN_OK; break; }
        if (j == jw) state = 0;
        __syncwarp();
    }
    if (it > max_iter) it = max_iter;
    const bool nonfinite = Gp::any(act && !(isfinite(u) && isfinite(g)), nullptr);
    double Uj = (state < 0) ? lbj : ((state > 0) ? ubj : fmin(fmax(u, lbj), ubj));
    if (nonfinite) Uj = nan("");
    if (nonfinite || broke) status = NTM_SCN_NONFINITE;
    __syncwarp();
    Uout = Uj;
    state_out = state;
    iters_out = it;
    return status;
}

// ------------------------------------------------------------------------------------------------
// The kernel.  E = horizon indices per lane (N <= 4E), EXT = 1 adds the RK4 plant option.
// One pass of the main loop = for every quad of the warp: [fetch a scenario] -> condense G, F (:66,:72-73 / :119-121)
// -> stop rule of the iteration that just finished (:123-127) and, at the end of a time step, the plant (:130) -> QP
// (:97) -> rollout with the old rho and re-scheduling (:110-117).  All shuffles / ballots sit at warp-uniform points;
// everything per-quad is predicated.
// ------------------------------------------------------------------------------------------------
template <int E, int EXT>
__global__ void __launch_bounds__(128, 2) closed_loop_quad_kernel(LoopArgs a, unsigned int region_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, l = lane & 3;
    const int quad_in_cta = threadIdx.x >> 2;
    const int quads_per_cta = blockDim.x >> 2;
    const int qlead = lane & 28;                                   // first lane of this quad
    const int N = a.N, S = a.S, flags = a.flags, layout = a.layout;
    const QuadWork w = quad_carve(smem_raw + (size_t)quad_in_cta * region_bytes, N);
    double *hslab = a.hscratch + ((size_t)blockIdx.x * quads_per_cta + quad_in_cta) * (size_t)N * N;
    const bool fxk = (flags & NTM_PROFILE_F_XK) != 0;
    const bool fixed = (flags & NTM_PROFILE_INNER_FIXED) != 0;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const int EX = 2 * (a.k_sim + 1);
    const int qp_cap = 10 * N + 20;
    const int j0 = l * E;                                          // first index of this lane
    const Params &P = *w.prm;
    const bool rec = a.rec_ld > 0;

    // per-quad scenario state
    int s = -1, k = 0, it = 0, inner = 0, qpit = 0, status = 0, hn = 0;
    bool live = false, exhausted = false, fresh = false;
    double x1 = 0.0, x2 = 0.0, x01 = 0.0, x02 = 0.0, cost = 0.0;
    double a11[E], a21[E], bb[E], U[E], Uold[E], hU1[E], hU2[E];
    int hs1[E], hs2[E];
#pragma unroll
    for (int e = 0; e < E; ++e) { a11[e] = 1.0; a21[e] = 0.0; bb[e] = 1.0; U[e] = 0.0; Uold[e] = 1.0; hU1[e] = 0.0; hU2[e] = 0.0; hs1[e] = -1; hs2[e] = -1; }

    for (;;) {
        // ------------------------------------------------------------------ (0) quads without a scenario pull one
        {
            int t = 0;
            if (!live && !exhausted && l == 0) t = (int)atomicAdd(a.counter, 1u);
            t = __shfl_sync(NTM_QFULL, t, qlead);
            if (!live && !exhausted) {
                if (t >= S) exhausted = true;
                else {
                    s = t; live = true; fresh = true;
                    const int ss = (a.params_count == 1) ? 0 : s, SS = (a.params_count == 1) ? 1 : a.params_count;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int pi = 4 * l + q;
                        const double v = __ldg(a.params + elem(layout, SS, NTM_NPARAM, ss, pi == 15 ? 7 : pi));
                        reinterpret_cast<double *>(w.prm)[pi] = (pi == 15) ? 1.0 / v : v;      // slot 15: 1 / w_dep
                    }
                    x01 = __ldg(a.x0 + elem(layout, S, 2, s, 0)); x02 = __ldg(a.x0 + elem(layout, S, 2, s, 1));
                }
            }
            if (__all_sync(NTM_QFULL, !live)) break;
            __syncwarp();
            if (fresh) {
                // offline build, NTM_MPC_Sim.m:63-65: rho(x0) repeated over the horizon
                fresh = false;
                k = 0; it = 0; inner = 0; qpit = 0; status = 0; hn = 0; cost = 0.0;
                x1 = x01; x2 = x02;
                double sa, sc_, sb;
                schedule(P, flags, x1, x2, sa, sc_, sb);
#pragma unroll
                for (int e = 0; e < E; ++e) { a11[e] = sa; a21[e] = sc_; bb[e] = sb; U[e] = 0.0; Uold[e] = 1.0; hU1[e] = 0.0; hU2[e] = 0.0; hs1[e] = -1; hs2[e] = -1; }
                if (l == 0) {
                    const size_t rb = (size_t)s * (size_t)a.rec_ld;
                    a.xk[rec ? rb : elem(layout, S, EX, s, 0)] = x1;
                    a.xk[rec ? rb + 1 : elem(layout, S, EX, s, 1)] = x2;
                }
            }
        }
        const size_t rbase = (size_t)(s < 0 ? 0 : s) * (size_t)a.rec_ld;

        // ------------------------------------------------------------------ (1) condense: G (packed), Fy
        double Fy[E];
        {
            const double xF1 = fxk ? x1 : x01, xF2 = fxk ? x2 : x02;
            const double a22 = P.a22, C1 = P.C1, C2 = P.C2;
            QAff tot = {1.0, 0.0, 1.0, 0.0, 0.0};
#pragma unroll
            for (int e = 0; e < E; ++e) {
                if (j0 + e < N) { const QAff m = {a11[e], a21[e], a22, C1, C2}; tot = qaff_compose(m, tot); }
            }
            const QAff ex = quad_exclusive(tot, l);
            double p1 = ex.a, p2 = ex.c;                                   // first column of A_{j0-1} ... A_0
            double v1 = fma(ex.a, xF1, ex.k1), v2 = fma(ex.c, xF1, fma(ex.s, xF2, ex.k2));
            const double q11 = P.q11, q12 = P.q12, q22 = P.q22, r1 = P.r1, r2 = P.r2;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int j = j0 + e;
                if (j < N) {
                    w.P12[j] = make_double2(p1, p2);
                    w.QP12[j] = make_double2(2.0 * (q11 * p1 + q12 * p2), 2.0 * (q12 * p1 + q22 * p2));
                    const double nv1 = fma(a11[e], v1, C1);
                    const double nv2 = fma(a22, v2, fma(a21[e], v1, C2));
                    v1 = nv1; v2 = nv2;
                    const double e1 = v1 - r1, e2 = v2 - r2;
                    w.QE12[j] = make_double2(2.0 * (q11 * e1 + q12 * e2), 2.0 * (q12 * e1 + q22 * e2));
                    const double np1 = a11[e] * p1;
                    const double np2 = fma(a21[e], p1, a22 * p2);
                    p1 = np1; p2 = np2;
                }
            }
            w.QP12[N + 2 * l] = make_double2(0.0, 0.0); w.QP12[N + 2 * l + 1] = make_double2(0.0, 0.0);
            w.QE12[N + 2 * l] = make_double2(0.0, 0.0); w.QE12[N + 2 * l + 1] = make_double2(0.0, 0.0);
            __syncwarp();
            // correlation sums, two lags at a time: lag slots t, t+1 of this lane are 4t + l and 4(t+1) + 3 - l
#pragma unroll
            for (int t = 0; t < E; t += 2) {
                const int eA = 4 * t + l;
                const int eB = 4 * (t + 1) + 3 - l;
                constexpr bool dummy = false; (void)dummy;
                const bool hasB = (t + 1 < E);
                double gA = 0.0, fA = 0.0, gB = 0.0, fB = 0.0;
                const int mmax = N - 4 * t;
                int addrA = qtri(N - 1) + (N - 1 - eA), addrB = qtri(N - 1) + (N - 1 - eB);
                const double2 *qpa = w.QP12 + eA, *qea = w.QE12 + eA, *qpb = w.QP12 + eB, *qeb = w.QE12 + eB;
                for (int m = 0; m < mmax; ++m) {
                    const double2 p = w.P12[m];
                    {
                        const double2 qa = qpa[m], ea = qea[m];
                        fA = fma(p.x, ea.x, fA); fA = fma(p.y, ea.y, fA);
                        gA = fma(p.x, qa.x, gA); gA = fma(p.y, qa.y, gA);
                        if (m <= N - 1 - eA) w.G[addrA] = gA;
                    }
                    if (hasB) {
                        const double2 qb = qpb[m], eb = qeb[m];
                        fB = fma(p.x, eb.x, fB); fB = fma(p.y, eb.y, fB);
                        gB = fma(p.x, qb.x, gB); gB = fma(p.y, qb.y, gB);
                        if (m <= N - 1 - eB) w.G[addrB] = gB;
                    }
                    addrA -= N - m; addrB -= N - m;
                }
                if (eA < N) w.Fx[eA] = fA;
                if (hasB && eB < N) w.Fx[eB] = fB;
            }
            __syncwarp();
#pragma unroll
            for (int e = 0; e < E; ++e) Fy[e] = (j0 + e < N) ? w.Fx[j0 + e] : 0.0;
        }

        // ------------------------------------------------------------------ (2) stop rule, plant, outputs
        {
            double dsum = 0.0;
            if (!fixed) {                                                                  // the fixed policy never looks at |Uold - U|
                double dl = 0.0;
#pragma unroll
                for (int e = 0; e < E; ++e) if (j0 + e < N) dl += fabs(Uold[e] - U[e]);
                dsum = quad_sum(dl);                                                       // :123
            }
            const double u0 = __shfl_sync(NTM_QFULL, U[0], qlead);                         // :107  uk(:,k) = U(1)
            if (live && it > 0) {
                inner = it;
                const bool brk = !fixed && dsum < a.eps;                                   // :124-125
                const bool stop = brk || it == a.i_sim;                                    // :94
                if (!brk) {
#pragma unroll
                    for (int e = 0; e < E; ++e) Uold[e] = U[e];                            // :127 (skipped by the break)
                }
                if (stop) {
                    double nw, nom;
                    if constexpr (EXT != 0) plant_of(P, flags, P.a22, 0.0, x1, x2, u0, nw, nom);   // :130, or its RK4 refinement
                    else plant_euler(P, flags, x1, x2, u0, nw, nom);                       // :130
                    x1 = nw; x2 = nom;
                    const double e1 = x1 - P.r1, e2 = x2 - P.r2;
                    cost += e1 * (P.q11 * e1 + P.q12 * e2) + e2 * (P.q12 * e1 + P.q22 * e2);
                    if (l == 0) {
                        a.xk[rec ? rbase + 2 * (k + 1) : elem(layout, S, EX, s, 2 * (k + 1))] = x1;
                        a.xk[rec ? rbase + 2 * (k + 1) + 1 : elem(layout, S, EX, s, 2 * (k + 1) + 1)] = x2;
                        a.uk[rec ? rbase + k : elem(layout, S, a.k_sim, s, k)] = u0;
                        if (a.inner) a.inner[elem(layout, S, a.k_sim, s, k)] = inner;
                        if (a.qpit) a.qpit[elem(layout, S, a.k_sim, s, k)] = qpit;
                    }
                    ++k; it = 0; qpit = 0;
                }
            }
            if (live && k >= a.k_sim) {                                                    // the scenario is complete
                if (l == 0) {
                    if (!(isfinite(x1) && isfinite(x2) && isfinite(cost))) status = max(status, (int)NTM_SCN_NONFINITE);
                    if (a.cost) a.cost[rec ? rbase : (size_t)s] = cost;
                    if (a.status) a.status[s] = status;
                    if (rec && a.rec_status) a.rec_status[rbase] = (double)status;
                }
                live = false;
            }
        }
        if (live) ++it;

        // ------------------------------------------------------------------ (4) QP in the variables y = b .* U  (:97)
        double lo[E], hi[E];
        int st_y[E];                               // partition state of the answer in y-space
        double yv[E];                              // the answer
        int qp_status = NTM_SCN_OK, nit = 1;
        {
            bool act[E], pin[E];
            double c1[E], c2[E];
            int s1y[E], s2y[E];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                act[e] = j0 + e < N;
                const double yl = bb[e] * P.umin, yh = bb[e] * P.umax;
                lo[e] = fmin(yl, yh); hi[e] = fmax(yl, yh);
                pin[e] = !(hi[e] > lo[e]);
                const bool neg = bb[e] < 0.0;
                c1[e] = (hn >= 1) ? bb[e] * hU1[e] : lo[e];
                c2[e] = (hn >= 2) ? bb[e] * hU2[e] : lo[e];
                s1y[e] = neg ? -hs1[e] : hs1[e];
                s2y[e] = neg ? -hs2[e] : hs2[e];
                if (act[e]) {
                    w.cand[j0 + e] = make_double2(c2[e], c1[e]); w.cand0[j0 + e] = make_double2(lo[e], hi[e]);
                    w.hst[j0 + e] = make_int2(s2y[e], s1y[e]);
                }
            }
            __syncwarp();
            bool solved = !live;
            double u[E], g[E], sc[E];
            int state[E];
#pragma unroll
            for (int e = 0; e < E; ++e) { u[e] = lo[e]; g[e] = Fy[e]; sc[e] = fabs(Fy[e]); state[e] = -1; }
            // ---- fast path: is one of the two previous solutions a vertex with multipliers of the right sign?  (strict
            //      test t >= 0: no gradient scale needed; the borderline cases go to the slow path with its tolerance)
            double g2[E], g3[E];
#pragma unroll
            for (int e = 0; e < E; ++e) { g2[e] = 0.0; g3[e] = 0.0; }
            for (int kk = 0; kk < N; ++kk) {
                const double2 cb = w.cand[kk];
                const int tk = qtri(kk);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int j = j0 + e;
                    const double gk = (kk <= j) ? w.G[qtri(j) + kk] : w.G[tk + j];
                    g2[e] = fma(gk, cb.x, g2[e]); g3[e] = fma(gk, cb.y, g3[e]);
                }
            }
            {
                bool ok1 = true, ok2 = true;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const double t1 = g3[e] + Fy[e], t2 = g2[e] + Fy[e];
                    ok1 = ok1 && (!act[e] || pin[e] || (s1y[e] < 0 && t1 >= 0.0) || (s1y[e] > 0 && -t1 >= 0.0));
                    ok2 = ok2 && (!act[e] || pin[e] || (s2y[e] < 0 && t2 >= 0.0) || (s2y[e] > 0 && -t2 >= 0.0));
                }
                const bool all1 = quad_all(ok1, lane), all2 = quad_all(ok2, lane);
                if (!solved && hn >= 1 && all1) {
                    solved = true;
#pragma unroll
                    for (int e = 0; e < E; ++e) { state[e] = pin[e] ? -1 : s1y[e]; u[e] = c1[e]; g[e] = g3[e] + Fy[e]; }
                } else if (!solved && hn >= 2 && all2) {
                    solved = true;
#pragma unroll
                    for (int e = 0; e < E; ++e) { state[e] = pin[e] ? -1 : s2y[e]; u[e] = c2[e]; g[e] = g2[e] + Fy[e]; }
                }
            }
            // ---- slow path: the warp solves the quads that need it one after the other (qp_solve_warp_packed)
            {
                unsigned slow_mask = __ballot_sync(NTM_QFULL, !solved && l == 0);          // bit 4q: quad q of this warp
                const int warp_quad0 = (threadIdx.x >> 5) << 3;
                while (slow_mask) {
                    const int ql = __ffs(slow_mask) - 1;
                    slow_mask &= slow_mask - 1;
                    const int qq = ql >> 2;
                    const QuadWork wq = quad_carve(smem_raw + (size_t)(warp_quad0 + qq) * region_bytes, N);
                    double *slab = a.hscratch + ((size_t)blockIdx.x * quads_per_cta + warp_quad0 + qq) * (size_t)N * N;
                    const int hnq = __shfl_sync(NTM_QFULL, hn, ql);
                    const double Fj = (lane < N) ? wq.Fx[lane] : 0.0;
                    int its = 1, stj = -1;
                    double Uq = 0.0;
                    const int stq = qp_solve_warp_packed(N, lane, wq, slab, Fj, hnq, qp_cap, its, Uq, stj);
                    if (lane < N) { wq.uv[lane] = Uq; wq.idx[lane] = stj; }
                    __syncwarp();
                    if ((lane >> 2) == qq) {
                        solved = true; qp_status = stq; nit = its;
#pragma unroll
                        for (int e = 0; e < E; ++e) {
                            if (act[e]) { u[e] = w.uv[j0 + e]; state[e] = w.idx[j0 + e]; g[e] = 0.0; }
                        }
                    }
                    __syncwarp();
                }
            }
            // answer (bound components exactly lb / ub), IEEE-faithful on non-finite data
            bool nf = false;
#pragma unroll
            for (int e = 0; e < E; ++e) nf = nf || (act[e] && !(isfinite(u[e]) && isfinite(g[e])));
            const bool nonfinite = quad_any(nf, lane);
#pragma unroll
            for (int e = 0; e < E; ++e) {
                yv[e] = (state[e] < 0) ? lo[e] : ((state[e] > 0) ? hi[e] : fmin(fmax(u[e], lo[e]), hi[e]));
                if (nonfinite) yv[e] = nan("");
                st_y[e] = state[e];
            }
            if (nonfinite) qp_status = NTM_SCN_NONFINITE;
        }
        // back to U-space; history of the last two solutions (warm starts)
        if (live) {
            status = max(status, qp_status);
            qpit += nit;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const bool neg = bb[e] < 0.0;
                const int su = neg ? -st_y[e] : st_y[e];
                double Uj = (su < 0 || bb[e] == 0.0) ? P.umin : ((su > 0) ? P.umax : fmin(fmax(yv[e] / bb[e], P.umin), P.umax));
                if (!(yv[e] == yv[e])) Uj = yv[e];                                          // NaN stays NaN
                const int sn = (bb[e] == 0.0) ? -1 : su;
                hU2[e] = hU1[e]; hs2[e] = hs1[e];
                hU1[e] = Uj; hs1[e] = sn;
                U[e] = Uj;
                if (a.Uk != nullptr && j0 + e < N) a.Uk[elem(layout, S, N * a.k_sim, s, k * N + j0 + e)] = Uj;   // :106
            }
            hn = min(hn + 1, 2);
        }

        // ------------------------------------------------------------------ (5) rollout with the OLD rho, re-scheduling
        {
            const double a22 = P.a22, C1 = P.C1, C2 = P.C2;
            QAff tot = {1.0, 0.0, 1.0, 0.0, 0.0};
#pragma unroll
            for (int e = 0; e < E; ++e) {
                if (j0 + e < N) { const QAff m = {a11[e], a21[e], a22, fma(bb[e], U[e], C1), C2}; tot = qaff_compose(m, tot); }
            }
            const QAff ex = quad_exclusive(tot, l);
            double z1 = fma(ex.a, x1, ex.k1), z2 = fma(ex.c, x1, fma(ex.s, x2, ex.k2));    // xN(:, j0)
#pragma unroll
            for (int e = 0; e < E; ++e) {
                if (j0 + e < N) {
                    double na, nc, nb;
                    schedule<true>(P, flags, z1, z2, na, nc, nb);                          // :114-116 on xN(:,i)
                    const double n1 = fma(a11[e], z1, fma(bb[e], U[e], C1));               // :113 with the old rho
                    const double n2 = fma(a22, z2, fma(a21[e], z1, C2));
                    z1 = n1; z2 = n2;
                    if (live) { a11[e] = na; a21[e] = nc; bb[e] = nb; }
                }
            }
            __syncwarp();
        }
    }
    // the last quad to leave re-arms the work queue for the next launch
    if (l == 0) {
        const unsigned int groups = gridDim.x * quads_per_cta;
        __threadfence();
        if (atomicAdd(a.counter + 1, 1u) == groups - 1) { a.counter[0] = 0u; a.counter[1] = 0u; __threadfence(); }
    }
}

}  // namespace ntm
