/* CPU oracle, C restatement  (TEST INFRASTRUCTURE ONLY -- never linked into the product).
 *
 * Scalar fp64 restatement of the literal, minimally repaired reference (see ntm_oracle.py for
 * the defect ledger).  It exists for two jobs only: (1) bulk parity checks in tests/ at sizes the
 * NumPy oracle cannot reach in seconds, (2) the `cpu_baseline` / `--impl reference` legs of
 * bench.py (OpenMP over scenarios, core count reported).  tests/test_oracle.py pins it against the
 * NumPy oracle and the golden vectors.  Compile with -ffp-contract=off: MATLAB does not fuse.
 *
 * Reference lines followed (relative to the upstream tree):
 *   NTM_MPC_Sim.m:24,25,37      kappa, zeta, C
 *   rho1.m:2 | rhos.m:18, rho2.m:2, rho3.m:2-3
 *   A.m:2, B.m:2
 *   Rho_to_PhiGammaLambda.m:17-52
 *   NTM_MPC_Sim.m:67-73,120-121  G, F
 *   NTM_MPC_Sim.m:93-131         closed loop
 *
 * PARITY UNPINNED at the QP boundary: the reference calls closed-source quadprog
 * (NTM_MPC_Sim.m:97); the box QP here is an exact primal active-set method certified by KKT
 * residuals in the tests.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NTM_ORACLE_MAXN 128

/* physics block, order shared with ntm_oracle.py:PHYS_ORDER */
typedef struct {
    double j_BS, w_dep, w_marg, w_sat, tau_r, rs, a, eta_CD, tau_E0, mu0, Lq, B_pol, m, Cw,
        tau_A0, tau_w, omega0, Ts, umin, umax, r1, r2, q11, q12, q22,
        c_tauE; /* tau_E(w) hook (PF_TAUE_W): tau_E = tau_E0 * (1 - c_tauE * w); NTM_MPC_Sim.m:14 "NOT EXACT FORMULA" */
} ntm_phys;
#define NTM_NPHYS 26

enum { PF_RHO1_SQ = 1, PF_GAMMA_I = 2, PF_F_XK = 4, PF_PLANT_C = 8, PF_INNER_FIXED = 16, PF_PLANT_RK4 = 64, PF_TAUE_W = 128 };

static const double PI_ = 3.141592653589793;

static double kappa_of(const ntm_phys *p) { /* NTM_MPC_Sim.m:24 */
    return 16 * p->mu0 * p->Lq * (p->rs * p->rs) / (0.82 * p->tau_r * p->B_pol * PI_);
}
static double zeta_of(const ntm_phys *p) { /* NTM_MPC_Sim.m:25 */
    return p->m * p->Cw * (p->tau_A0 * p->tau_A0) * p->tau_w * (p->a * p->a * p->a);
}
static void C_of(const ntm_phys *p, double C[2]) { /* NTM_MPC_Sim.m:37 */
    double kappa = kappa_of(p);
    C[0] = -4.0 / 3.0 * (kappa * p->Ts * p->j_BS * p->w_sat) / (p->w_sat * p->w_sat + p->w_marg * p->w_marg);
    C[1] = p->Ts * p->omega0 / p->tau_E0;
}

static double rho1f(const double x[2], double wmarg, int sq) { /* rho1.m:2 | rhos.m:18 */
    return sq ? 1.0 / (x[0] * x[0] + wmarg * wmarg) : 1.0 / (x[0] + wmarg * wmarg);
}
static double rho2f(const double x[2]) { return (x[0] * x[0]) / x[1]; } /* rho2.m:2 */
static double rho3f(const double x[2], double w_dep) {                    /* rho3.m:2-3 */
    double ws = x[0] / w_dep;
    return (0.25 + 0.24 * ws) / (1 + 1.5 * ws + 0.43 * (ws * ws) + 0.64 * (ws * ws * ws));
}

/* A.m:2 -- row-major 2x2 {a11,a12,a21,a22} */
static void Af(double r1, double r2, double kappa, double taur, double Ts, double zeta, double rs,
               double a, double TE, double A[4]) {
    A[0] = ((4.0 / 3.0) * (kappa * rs / (0.82 * taur)) * Ts * r1 + 1);
    A[1] = 0.0;
    A[2] = ((r2 * Ts) / (zeta * (a * a * a)));
    A[3] = (1 - Ts / TE);
}
/* B.m:2 as the column [b;0] (defect D7) */
static void Bf(double r3, double wdep, double kappa, double Ts, double etaCD, double B[2]) {
    B[0] = (kappa * Ts * etaCD / wdep) * r3;
    B[1] = 0.0;
}

typedef struct { double kappa, zeta, taur, Ts, rs, a, TE, wdep, etaCD; } model_consts;

static void Am(const model_consts *c, double r1, double r2, double A[4]) {
    Af(r1, r2, c->kappa, c->taur, c->Ts, c->zeta, c->rs, c->a, c->TE, A);
}
static void Bm(const model_consts *c, double r3, double B[2]) { Bf(r3, c->wdep, c->kappa, c->Ts, c->etaCD, B); }

static void mv2(const double A[4], const double v[2], double out[2]) {
    double o0 = A[0] * v[0] + A[1] * v[1];
    double o1 = A[2] * v[0] + A[3] * v[1];
    out[0] = o0; out[1] = o1;
}

/* Rho_to_PhiGammaLambda.m.  Column-major outputs like MATLAB: Phi 2N x 2, Gamma 2N x N, Lambda 2N. */
static void condense(const model_consts *c, int N, const double *R1, const double *R2, const double *R3,
                     const double C[2], int gamma_i, double *Phi, double *Gam, double *Lam) {
    const int ld = 2 * N;
    double A[4], t[2], v[2];
    memset(Phi, 0, sizeof(double) * 4 * N);
    memset(Gam, 0, sizeof(double) * 2 * N * N);
    /* Phi :17-22 (left-multiplying, repair D5) */
    Am(c, R1[0], R2[0], A);
    Phi[0] = A[0]; Phi[1] = A[2]; Phi[ld] = A[1]; Phi[ld + 1] = A[3];
    for (int j = 2; j <= N; ++j) {
        Am(c, R1[j - 1], R2[j - 1], A);
        for (int col = 0; col < 2; ++col) {
            v[0] = Phi[col * ld + 2 * (j - 2)]; v[1] = Phi[col * ld + 2 * (j - 2) + 1];
            mv2(A, v, t);
            Phi[col * ld + 2 * (j - 1)] = t[0]; Phi[col * ld + 2 * (j - 1) + 1] = t[1];
        }
    }
    /* Gamma :26-40 */
    Bm(c, R3[0], t);
    Gam[0] = t[0]; Gam[1] = t[1];
    for (int i = 2; i <= N; ++i)
        for (int j = 1; j <= i; ++j) {
            double *dst = Gam + (size_t)(j - 1) * ld + 2 * (i - 1);
            if (i != j) {
                int k = gamma_i ? i : (i - j); /* :32 literal index i-j */
                Am(c, R1[k - 1], R2[k - 1], A);
                const double *src = Gam + (size_t)(j - 1) * ld + 2 * (i - 2);
                v[0] = src[0]; v[1] = src[1];
                mv2(A, v, t);
                dst[0] = t[0]; dst[1] = t[1];
            } else {
                Bm(c, R3[j - 1], t);
                dst[0] = t[0]; dst[1] = t[1];
            }
        }
    /* Lambda :47-52 */
    Lam[0] = C[0]; Lam[1] = C[1];
    for (int i = 2; i <= N; ++i) {
        Am(c, R1[i - 1], R2[i - 1], A);
        v[0] = Lam[2 * (i - 2)]; v[1] = Lam[2 * (i - 2) + 1];
        mv2(A, v, t);
        Lam[2 * (i - 1)] = t[0] + C[0]; Lam[2 * (i - 1) + 1] = t[1] + C[1];
    }
}

/* NTM_MPC_Sim.m:72-73: G = 2 Gamma' Omega Gamma, F = 2 Gamma' Omega (Phi x + Lambda - R), Omega = I (x) Q.
 * G row-major N x N (symmetric). */
static void hessian_grad(int N, const double *Phi, const double *Gam, const double *Lam, const double x[2],
                         const double r[2], const double Q[3], double *G, double *F, double *work) {
    const int ld = 2 * N;
    double *OG = work;            /* Omega*Gamma, 2N x N column-major */
    double *e = work + 2 * N * N; /* Omega*(Phi x + Lambda - R) */
    for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i) {
            double g0 = Gam[(size_t)j * ld + 2 * i], g1 = Gam[(size_t)j * ld + 2 * i + 1];
            OG[(size_t)j * ld + 2 * i] = Q[0] * g0 + Q[1] * g1;
            OG[(size_t)j * ld + 2 * i + 1] = Q[1] * g0 + Q[2] * g1;
        }
    for (int i = 0; i < N; ++i) {
        double v0 = Phi[2 * i] * x[0] + Phi[ld + 2 * i] * x[1] + Lam[2 * i] - r[0];
        double v1 = Phi[2 * i + 1] * x[0] + Phi[ld + 2 * i + 1] * x[1] + Lam[2 * i + 1] - r[1];
        e[2 * i] = Q[0] * v0 + Q[1] * v1;
        e[2 * i + 1] = Q[1] * v0 + Q[2] * v1;
    }
    for (int a = 0; a < N; ++a) {
        for (int b = 0; b <= a; ++b) {
            double s = 0.0;
            for (int k = 2 * a; k < ld; ++k) s += Gam[(size_t)a * ld + k] * OG[(size_t)b * ld + k];
            G[a * N + b] = G[b * N + a] = 2 * s;
        }
        double s = 0.0;
        for (int k = 2 * a; k < ld; ++k) s += Gam[(size_t)a * ld + k] * e[k];
        F[a] = 2 * s;
    }
}

/* Cholesky solve of the m x m SPD system (in place, row-major lower).  Returns 0 on success. */
static int chol_solve(int m, double *Hm, double *b) {
    for (int j = 0; j < m; ++j) {
        double d = Hm[j * m + j];
        for (int k = 0; k < j; ++k) d -= Hm[j * m + k] * Hm[j * m + k];
        if (!(d > 0.0)) return 1;
        d = sqrt(d);
        Hm[j * m + j] = d;
        for (int i = j + 1; i < m; ++i) {
            double s = Hm[i * m + j];
            for (int k = 0; k < j; ++k) s -= Hm[i * m + k] * Hm[j * m + k];
            Hm[i * m + j] = s / d;
        }
    }
    for (int i = 0; i < m; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= Hm[i * m + k] * b[k];
        b[i] = s / Hm[i * m + i];
    }
    for (int i = m - 1; i >= 0; --i) {
        double s = b[i];
        for (int k = i + 1; k < m; ++k) s -= Hm[k * m + i] * b[k];
        b[i] = s / Hm[i * m + i];
    }
    return 0;
}

/* Exact primal active-set box QP in scaled variables (same algorithm as ntm_oracle.py:qp_box).
 * work: >= 2*N*N + 6*N doubles.  Returns status (0 ok, 1 iteration cap, 2 non-finite). */
int ntm_oracle_qp_box(int N, const double *G, const double *F, const double *lb, const double *ub,
                      double *U, int *iters, double *work) {
    double *Hs = work, *Hm = work + N * N, *fs = Hm + N * N, *t = fs + N, *g = t + N, *p = g + N,
           *rng = p + N, *rhs = rng + N;
    int state[NTM_ORACLE_MAXN], idx[NTM_ORACLE_MAXN];
    const double tol_g = 1e-13;
    int status = 1, it = 0, max_iter = 10 * N + 20;
    for (int i = 0; i < N; ++i) {
        if (!isfinite(F[i])) { for (int k = 0; k < N; ++k) U[k] = NAN; *iters = 0; return 2; }
        for (int j = 0; j < N; ++j)
            if (!isfinite(G[i * N + j])) { for (int k = 0; k < N; ++k) U[k] = NAN; *iters = 0; return 2; }
    }
    for (int i = 0; i < N; ++i) rng[i] = ub[i] - lb[i];
    for (int i = 0; i < N; ++i) {
        double s = 0.0;
        for (int j = 0; j < N; ++j) { Hs[i * N + j] = G[i * N + j] * rng[i] * rng[j]; s += G[i * N + j] * lb[j]; }
        fs[i] = (F[i] + s) * rng[i];
        t[i] = 0.0; state[i] = -1;
    }
    for (it = 1; it <= max_iter; ++it) {
        int m = 0, block = -1, bstate = 0;
        double alpha = 1.0;
        for (int i = 0; i < N; ++i) if (state[i] == 0) idx[m++] = i;
        if (m > 0) {
            for (int i = 0; i < N; ++i) { double s = 0.0; for (int j = 0; j < N; ++j) s += Hs[i * N + j] * t[j]; g[i] = s + fs[i]; }
            for (int a = 0; a < m; ++a) {
                for (int b = 0; b < m; ++b) Hm[a * m + b] = Hs[idx[a] * N + idx[b]];
                rhs[a] = -g[idx[a]];
            }
            if (chol_solve(m, Hm, rhs)) { status = 2; break; }
            for (int i = 0; i < N; ++i) p[i] = 0.0;
            for (int a = 0; a < m; ++a) p[idx[a]] = rhs[a];
            for (int a = 0; a < m; ++a) {
                int i = idx[a];
                if (p[i] < 0) { double ai = (0.0 - t[i]) / p[i]; if (ai < alpha) { alpha = ai; block = i; bstate = -1; } }
                else if (p[i] > 0) { double ai = (1.0 - t[i]) / p[i]; if (ai < alpha) { alpha = ai; block = i; bstate = 1; } }
            }
            for (int i = 0; i < N; ++i) t[i] = t[i] + alpha * p[i];
        }
        if (block >= 0) { state[block] = bstate; t[block] = (bstate == -1) ? 0.0 : 1.0; continue; }
        double worst_v = INFINITY; int worst = 0;
        for (int i = 0; i < N; ++i) {
            double s = 0.0, sc = 0.0;
            for (int j = 0; j < N; ++j) { s += Hs[i * N + j] * t[j]; sc += fabs(Hs[i * N + j]) * fabs(t[j]); }
            s += fs[i]; sc += fabs(fs[i]) + 1e-300;
            double lam = (state[i] == -1) ? s / sc : (state[i] == 1 ? -s / sc : 0.0);
            if (lam < worst_v) { worst_v = lam; worst = i; }
        }
        if (worst_v >= -tol_g) { status = 0; break; }
        state[worst] = 0;
    }
    if (it > max_iter) it = max_iter;
    /* polish on the final partition, raw data */
    {
        int m = 0;
        for (int i = 0; i < N; ++i) { U[i] = (state[i] == 1) ? ub[i] : lb[i]; if (state[i] == 0) idx[m++] = i; }
        if (m > 0 && status != 2) {
            for (int a = 0; a < m; ++a) {
                int i = idx[a];
                double s = 0.0;
                for (int j = 0; j < N; ++j) if (state[j] != 0) s += G[i * N + j] * U[j];
                rhs[a] = -(F[i] + s) * rng[i];
                for (int b = 0; b < m; ++b) Hm[a * m + b] = G[i * N + idx[b]] * rng[i] * rng[idx[b]];
            }
            if (chol_solve(m, Hm, rhs)) status = 2;
            for (int a = 0; a < m; ++a) {
                int i = idx[a];
                double v = rhs[a] * rng[i];
                v = v < lb[i] ? lb[i] : v;
                v = v > ub[i] ? ub[i] : v;
                U[i] = v;
            }
        }
    }
    *iters = it;
    return status;
}

/* ---- materialising entry points used by the tests ---- */
void ntm_oracle_condense(const double *phys, int N, const double *R1, const double *R2, const double *R3, int flags,
                         double *Phi, double *Gam, double *Lam) {
    const ntm_phys *p = (const ntm_phys *)phys;
    model_consts c = { kappa_of(p), zeta_of(p), p->tau_r, p->Ts, p->rs, p->a, p->tau_E0, p->w_dep, p->eta_CD };
    double C[2];
    C_of(p, C);
    condense(&c, N, R1, R2, R3, C, (flags & PF_GAMMA_I) != 0, Phi, Gam, Lam);
}

void ntm_oracle_hessian_grad(int N, const double *Phi, const double *Gam, const double *Lam, const double *x,
                             const double *r, const double *Q3, double *G, double *F) {
    double *work = (double *)malloc(sizeof(double) * (2 * (size_t)N * N + 2 * N));
    hessian_grad(N, Phi, Gam, Lam, x, r, Q3, G, F, work);
    free(work);
}

/* NTM_MPC_Sim.m:130: x+ = A(rho(x)) x + B(rho(x)) u (+ C with PF_PLANT_C) */
static void plant_euler(const model_consts *c, const ntm_phys *p, const double C[2], int flags, const double x[2],
                        double u, double out[2]) {
    double A[4], B[2], t[2];
    model_consts cx = *c;
    if (flags & PF_TAUE_W) cx.TE = p->tau_E0 * (1 - p->c_tauE * x[0]);   /* the vector field carries tau_E of its own state */
    Am(&cx, rho1f(x, p->w_marg, (flags & PF_RHO1_SQ) != 0), rho2f(x), A); Bm(c, rho3f(x, p->w_dep), B);
    mv2(A, x, t);
    out[0] = t[0] + B[0] * u; out[1] = t[1] + B[1] * u;
    if (flags & PF_PLANT_C) { out[0] += C[0]; out[1] += C[1]; }
}

/* One scenario of NTM_MPC_Sim.m:63-73 + :80-131.  Outputs: xk[2*(k_sim+1)] column-major (2 x k_sim+1),
 * uk[k_sim], Uk[N*k_sim] column-major (may be NULL), inner[k_sim], qpit[k_sim], cost, returns status. */
static int closed_loop_one(const ntm_phys *p, const double x0[2], int N, int k_sim, int i_sim, double eps, int flags,
                           double *xk, double *uk, double *Uk, int *inner, int *qpit, double *cost, double *ws) {
    model_consts c = { kappa_of(p), zeta_of(p), p->tau_r, p->Ts, p->rs, p->a, p->tau_E0, p->w_dep, p->eta_CD };
    double C[2], A[4], B[2];
    C_of(p, C);
    if (flags & PF_TAUE_W) c.TE = p->tau_E0 * (1 - p->c_tauE * x0[0]);    /* offline build :63-73 */
    const int sq = (flags & PF_RHO1_SQ) != 0, gi = (flags & PF_GAMMA_I) != 0;
    const double r[2] = { p->r1, p->r2 }, Q[3] = { p->q11, p->q12, p->q22 };
    double *R1 = ws, *R2 = R1 + N, *R3 = R2 + N, *Phi = R3 + N, *Gam = Phi + 4 * N, *Lam = Gam + 2 * N * N,
           *G = Lam + 2 * N, *F = G + N * N, *U = F + N, *Uold = U + N, *lb = Uold + N, *ub = lb + N,
           *xN = ub + N, *work = xN + 2 * (N + 1);
    int status = 0;
    for (int i = 0; i < N; ++i) {
        R1[i] = rho1f(x0, p->w_marg, sq); R2[i] = rho2f(x0); R3[i] = rho3f(x0, p->w_dep);   /* :63-65 */
        Uold[i] = 1.0; lb[i] = p->umin; ub[i] = p->umax;                                      /* :86 */
    }
    condense(&c, N, R1, R2, R3, C, gi, Phi, Gam, Lam);                                        /* :66 */
    hessian_grad(N, Phi, Gam, Lam, x0, r, Q, G, F, work);                                     /* :72-73 */
    xk[0] = x0[0]; xk[1] = x0[1];
    for (int k = 0; k < k_sim; ++k) {
        const double *xc = xk + 2 * k;
        inner[k] = 0; qpit[k] = 0;
        /* tau_E(w) hook: the workspace tau_E is re-evaluated from the measured width at the top of each time step and
         * held over the horizon (ntm_oracle.py:closed_loop); G, F on hand keep the previous step's value */
        if (flags & PF_TAUE_W) c.TE = p->tau_E0 * (1 - p->c_tauE * xc[0]);
        for (int it = 1; it <= i_sim; ++it) {
            int nit = 0;
            int st = ntm_oracle_qp_box(N, G, F, lb, ub, U, &nit, work);                       /* :97 */
            if (st > status) status = st;
            qpit[k] += nit;
            if (Uk) for (int i = 0; i < N; ++i) Uk[(size_t)k * N + i] = U[i];                 /* :106 */
            uk[k] = U[0];                                                                     /* :107 */
            xN[0] = xc[0]; xN[1] = xc[1];                                                     /* :110 */
            for (int i = 0; i < N; ++i) {                                                     /* :112-117 */
                Am(&c, R1[i], R2[i], A); Bm(&c, R3[i], B);
                double t[2];
                mv2(A, xN + 2 * i, t);
                xN[2 * (i + 1)] = t[0] + B[0] * U[i] + C[0];
                xN[2 * (i + 1) + 1] = t[1] + B[1] * U[i] + C[1];
                R1[i] = rho1f(xN + 2 * i, p->w_marg, sq); R2[i] = rho2f(xN + 2 * i); R3[i] = rho3f(xN + 2 * i, p->w_dep);
            }
            condense(&c, N, R1, R2, R3, C, gi, Phi, Gam, Lam);                                /* :119 */
            hessian_grad(N, Phi, Gam, Lam, (flags & PF_F_XK) ? xc : x0, r, Q, G, F, work);    /* :120-121 */
            inner[k] = it;
            double d = 0.0;
            for (int i = 0; i < N; ++i) d += fabs(Uold[i] - U[i]);
            if (!(flags & PF_INNER_FIXED) && d < eps) break;                                  /* :123-126 */
            for (int i = 0; i < N; ++i) Uold[i] = U[i];                                       /* :127 */
        }
        {                                                                                     /* :130 */
            double n[2];
            if (flags & PF_PLANT_RK4) {                       /* fidelity option: RK4 of the same vector field */
                double kk[4][2], y[2] = { xc[0], xc[1] }, e[2];
                for (int st = 0; st < 4; ++st) {
                    plant_euler(&c, p, C, flags, y, uk[k], e);
                    kk[st][0] = e[0] - y[0]; kk[st][1] = e[1] - y[1];
                    const double h = (st < 2) ? 0.5 : 1.0;
                    y[0] = xc[0] + h * kk[st][0]; y[1] = xc[1] + h * kk[st][1];
                }
                for (int d = 0; d < 2; ++d)
                    n[d] = xc[d] + ((kk[0][d] + 2.0 * kk[1][d]) + (2.0 * kk[2][d] + kk[3][d])) / 6.0;
            } else {
                plant_euler(&c, p, C, flags, xc, uk[k], n);
            }
            xk[2 * (k + 1)] = n[0]; xk[2 * (k + 1) + 1] = n[1];
        }
    }
    double cs = 0.0;
    for (int k = 1; k <= k_sim; ++k) {
        double e0 = xk[2 * k] - r[0], e1 = xk[2 * k + 1] - r[1];
        if (!isfinite(e0) || !isfinite(e1)) status = status > 2 ? status : 2;
        cs += e0 * (Q[0] * e0 + Q[1] * e1) + e1 * (Q[1] * e0 + Q[2] * e1);
    }
    *cost = cs;
    return status;
}

static size_t ws_doubles(int N) { return (size_t)3 * N + 4 * N + 2 * N * N + 2 * N + N * N + 5 * N + 2 * (N + 1) + 2 * N * N + 8 * N + 64; }

/* Batch driver, scenario-slowest ("MATLAB") layouts: phys[S][26], x0[S][2], xk[S][2*(k_sim+1)], uk[S][k_sim],
 * Uk[S][N*k_sim] or NULL, inner/qpit[S][k_sim], cost[S], status[S].  threads<=0 -> all OpenMP threads.
 * Returns the number of threads used. */
int ntm_oracle_closed_loop_batch(int S, int N, int k_sim, int i_sim, double eps, int flags, const double *phys,
                                 const double *x0, double *xk, double *uk, double *Uk, int *inner, int *qpit,
                                 double *cost, int *status, int threads) {
    int used = 1;
    if (N > NTM_ORACLE_MAXN) return -1;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();    /* NB: 1 under torchrun (it exports OMP_NUM_THREADS=1): pass a count */
    omp_set_dynamic(0);
#pragma omp parallel num_threads(threads)
#endif
    {
        double *ws = (double *)malloc(sizeof(double) * ws_doubles(N));
#ifdef _OPENMP
#pragma omp single
        used = omp_get_num_threads();                      /* the team that actually runs, not the request */
#endif
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (int s = 0; s < S; ++s) {
            status[s] = closed_loop_one((const ntm_phys *)(phys + (size_t)s * NTM_NPHYS), x0 + 2 * (size_t)s, N, k_sim, i_sim,
                                        eps, flags, xk + (size_t)s * 2 * (k_sim + 1), uk + (size_t)s * k_sim,
                                        Uk ? Uk + (size_t)s * N * k_sim : NULL, inner + (size_t)s * k_sim,
                                        qpit + (size_t)s * k_sim, cost + s, ws);
        }
        free(ws);
    }
    return used;
}

int ntm_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
