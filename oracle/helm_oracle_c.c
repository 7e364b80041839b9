/*
 * helm_oracle_c.c -- C/OpenMP port of the CPU oracle's solve path, used ONLY as the timed CPU arm
 * (bench.py `cpu_baseline` and `--impl reference`) and cross-checked against oracle/helm_oracle.py in
 * tests/.  TEST INFRASTRUCTURE: the product never links or calls this file.
 *
 * It restates how the reference runs this path on a CPU (JuliaInv/Helmholtz.jl + its un-vendored
 * engines): an assembled sparse shifted operator (src/GetHelmholtz.jl:33-50,81-83; Kronecker Laplacian
 * src/PlainNodalLaplacian.jl:18-46), a Galerkin hierarchy A_{l+1} = R A_l P with full weighting /
 * linear interpolation (Multigrid.MGsetup, called at src/ShiftedLaplacianMultigridSolver.jl:65), damped
 * Jacobi, V or W cycle, an inexact Jacobi-GMRES coarsest solve, and right-preconditioned FGMRES(m)
 * (solveGMRES_MG, :89), every product being a multi-threaded sparse mat-vec over the N x nrhs block
 * (ParSpMatVec: y = A^H' x on a CSC of the transpose == row-wise CSR products, OpenMP over rows,
 * src/GetHelmholtz.jl:85-95).  "parity unpinned" applies as stated in helm_oracle.py.
 *
 * Build:  make -C oracle     (gcc -O3 -fopenmp -shared)
 */
#include <complex.h>
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef double _Complex zc;

typedef struct {
    int64_t nrows, ncols;
    int64_t* rowptr;
    int32_t* col;
    zc* val;
} Csr;

typedef struct {
    int dim;
    int64_t n[3];
    int64_t N;
    Csr A;      /* operator of this level (shifted) */
    Csr P, R;   /* to / from the next coarser level (absent on the coarsest) */
    zc* dinv;   /* relax_param / diag(A) */
    zc *x, *b, *r; /* work vectors N x kcap */
} Level;

typedef struct Oracle {
    int dim, levels, npre, npost, cycle, coarse_iters, kcap;
    double relax_param, shift, wre;
    Level* L;
    zc* shiftdiag; /* i*shift*wre^2*m: H x = SH x - shiftdiag .* x  (src/GetHelmholtz.jl:85-95) */
    double setup_seconds;
    int64_t n_prec;
} Oracle;

static double now(void) { return omp_get_wtime(); }

static void csr_free(Csr* a) {
    free(a->rowptr);
    free(a->col);
    free(a->val);
    memset(a, 0, sizeof(*a));
}

/* y(:,r) = alpha*A*x(:,r) + beta*y(:,r) for nrhs columns (leading dimensions = rows/cols of A) */
static void spmv(const Csr* A, const zc* x, zc* y, int nrhs, zc alpha, zc beta) {
    const int64_t nr = A->nrows, nc = A->ncols;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nr; ++i) {
        const int64_t p0 = A->rowptr[i], p1 = A->rowptr[i + 1];
        for (int r = 0; r < nrhs; ++r) {
            const zc* xr = x + (int64_t)r * nc;
            zc s = 0;
            for (int64_t p = p0; p < p1; ++p) s += A->val[p] * xr[A->col[p]];
            zc* yr = y + (int64_t)r * nr;
            yr[i] = (beta == 0 ? 0 : beta * yr[i]) + alpha * s;
        }
    }
}

/* ---- operator assembly: row-wise statement of kron(I,D1)+kron(D2,I)(+...) + diag(mass) ---------- */
static void assemble_fine(Oracle* o, Level* L, const double* h, const double* m, const double* gamma, double wre,
                          double wim, int neumann_top, int sommerfeld, int order_bc, double shift) {
    const int dim = o->dim;
    const int64_t n0 = L->n[0], n1 = L->n[1], n2 = L->n[2], N = L->N;
    const double BC = order_bc == 2 ? 2.0 : 1.0;
    const zc w = wre + wim * I, w2 = w * w;
    Csr* A = &L->A;
    A->nrows = A->ncols = N;
    A->rowptr = (int64_t*)malloc((N + 1) * sizeof(int64_t));
    const int per = 2 * dim + 1;
    A->col = (int32_t*)malloc((size_t)N * per * sizeof(int32_t));
    A->val = (zc*)malloc((size_t)N * per * sizeof(zc));
    /* row lengths */
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < N; ++p) {
        const int64_t i = p % n0, j = (p / n0) % n1, k = p / (n0 * n1);
        int c = 1 + (i > 0) + (i < n0 - 1) + (j > 0) + (j < n1 - 1);
        if (dim == 3) c += (k > 0) + (k < n2 - 1);
        A->rowptr[p + 1] = c;
    }
    A->rowptr[0] = 0;
    for (int64_t p = 0; p < N; ++p) A->rowptr[p + 1] += A->rowptr[p];
    const int64_t nn[3] = {n0, n1, n2};
    const int64_t st[3] = {1, n0, n0 * n1};
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < N; ++p) {
        const int64_t id[3] = {p % n0, (p / n0) % n1, p / (n0 * n1)};
        int64_t q = A->rowptr[p];
        /* mass = -w^2 m (1 - i gamma / Re w)  (GetHelmholtz.jl:41) */
        zc diag = -w2 * m[p] * (1.0 - I * gamma[p] / wre);
        /* Sommerfeld: mass -= (-i Re(w) (2/h_d) sqrt(m)) on every boundary face (GetHelmholtz.jl:43-47,222-247;
         * getSommerfeldBC is always called with its default second-order BC) */
        if (sommerfeld) {
            for (int d = 0; d < dim; ++d) {
                const int top_face = (d == dim - 1);
                if (id[d] == 0 && !(top_face && neumann_top)) diag += I * wre * (2.0 / h[d]) * sqrt(m[p]);
                if (id[d] == nn[d] - 1) diag += I * wre * (2.0 / h[d]) * sqrt(m[p]);
            }
        }
        diag += I * shift * wre * wre * m[p]; /* GetHelmholtzShiftOP (GetHelmholtz.jl:81-83) */
        for (int d = 0; d < dim; ++d) diag += ((id[d] == 0 || id[d] == nn[d] - 1) ? BC : 2.0) / (h[d] * h[d]);
        /* columns in ascending order: k-, j-, i-, centre, i+, j+, k+ */
        for (int d = dim - 1; d >= 0; --d)
            if (id[d] > 0) {
                A->col[q] = (int32_t)(p - st[d]);
                A->val[q++] = -((id[d] == nn[d] - 1) ? BC : 1.0) / (h[d] * h[d]);
            }
        A->col[q] = (int32_t)p;
        A->val[q++] = diag;
        for (int d = 0; d < dim; ++d)
            if (id[d] < nn[d] - 1) {
                A->col[q] = (int32_t)(p + st[d]);
                A->val[q++] = -((id[d] == 0) ? BC : 1.0) / (h[d] * h[d]);
            }
    }
}

/* linear interpolation P (fine x coarse) and full weighting R = 2^-dim P^T as CSR */
static void build_transfers(Level* F, Level* C, int dim) {
    const int64_t Nf = F->N, Nc = C->N;
    const int64_t cst[3] = {1, C->n[0], C->n[0] * C->n[1]};
    const int64_t fst[3] = {1, F->n[0], F->n[0] * F->n[1]};
    Csr* P = &F->P;
    P->nrows = Nf;
    P->ncols = Nc;
    P->rowptr = (int64_t*)malloc((Nf + 1) * sizeof(int64_t));
    P->rowptr[0] = 0;
    for (int64_t p = 0; p < Nf; ++p) {
        const int64_t id[3] = {p % F->n[0], (p / F->n[0]) % F->n[1], p / (F->n[0] * F->n[1])};
        int c = 1;
        for (int d = 0; d < dim; ++d) c *= (id[d] & 1) ? 2 : 1;
        P->rowptr[p + 1] = P->rowptr[p] + c;
    }
    P->col = (int32_t*)malloc((size_t)P->rowptr[Nf] * sizeof(int32_t));
    P->val = (zc*)malloc((size_t)P->rowptr[Nf] * sizeof(zc));
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < Nf; ++p) {
        const int64_t id[3] = {p % F->n[0], (p / F->n[0]) % F->n[1], p / (F->n[0] * F->n[1])};
        int64_t q = P->rowptr[p];
        const int o2 = dim == 3 ? (int)(id[2] & 1) : 0, o1 = (int)(id[1] & 1), o0 = (int)(id[0] & 1);
        const double w = 1.0 / ((1 << o0) * (1 << o1) * (1 << o2));
        for (int a2 = 0; a2 <= o2; ++a2)
            for (int a1 = 0; a1 <= o1; ++a1)
                for (int a0 = 0; a0 <= o0; ++a0) {
                    const int64_t c = (id[0] / 2 + a0) * cst[0] + (id[1] / 2 + a1) * cst[1] +
                                      (dim == 3 ? (id[2] / 2 + a2) * cst[2] : 0);
                    P->col[q] = (int32_t)c;
                    P->val[q++] = w;
                }
    }
    Csr* R = &F->R;
    R->nrows = Nc;
    R->ncols = Nf;
    R->rowptr = (int64_t*)malloc((Nc + 1) * sizeof(int64_t));
    R->rowptr[0] = 0;
    for (int64_t c = 0; c < Nc; ++c) {
        const int64_t id[3] = {c % C->n[0], (c / C->n[0]) % C->n[1], c / (C->n[0] * C->n[1])};
        int cnt = 1;
        for (int d = 0; d < dim; ++d) cnt *= 1 + (id[d] > 0) + (id[d] < C->n[d] - 1);
        R->rowptr[c + 1] = R->rowptr[c] + cnt;
    }
    R->col = (int32_t*)malloc((size_t)R->rowptr[Nc] * sizeof(int32_t));
    R->val = (zc*)malloc((size_t)R->rowptr[Nc] * sizeof(zc));
    const double sc = dim == 3 ? 0.125 : 0.25;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < Nc; ++c) {
        const int64_t id[3] = {c % C->n[0], (c / C->n[0]) % C->n[1], c / (C->n[0] * C->n[1])};
        int64_t q = R->rowptr[c];
        for (int e2 = (dim == 3 ? -1 : 0); e2 <= (dim == 3 ? 1 : 0); ++e2) {
            const int64_t f2 = dim == 3 ? 2 * id[2] + e2 : 0;
            if (f2 < 0 || f2 >= F->n[2]) continue;
            for (int e1 = -1; e1 <= 1; ++e1) {
                const int64_t f1 = 2 * id[1] + e1;
                if (f1 < 0 || f1 >= F->n[1]) continue;
                for (int e0 = -1; e0 <= 1; ++e0) {
                    const int64_t f0 = 2 * id[0] + e0;
                    if (f0 < 0 || f0 >= F->n[0]) continue;
                    R->col[q] = (int32_t)(f0 * fst[0] + f1 * fst[1] + f2 * fst[2]);
                    R->val[q++] = sc * (e0 ? 0.5 : 1.0) * (e1 ? 0.5 : 1.0) * (e2 ? 0.5 : 1.0);
                }
            }
        }
    }
}

static int cmp_i32(const void* a, const void* b) { return (*(const int32_t*)a > *(const int32_t*)b) - (*(const int32_t*)a < *(const int32_t*)b); }

/* Ac = R A P (Gustavson, one coarse row at a time, OpenMP over rows, two passes) */
static void galerkin(const Csr* R, const Csr* A, const Csr* P, Csr* Ac) {
    const int64_t Nc = R->nrows;
    Ac->nrows = Ac->ncols = Nc;
    Ac->rowptr = (int64_t*)calloc(Nc + 1, sizeof(int64_t));
    for (int pass = 0; pass < 2; ++pass) {
#pragma omp parallel
        {
            int64_t* mark = (int64_t*)malloc(Nc * sizeof(int64_t));
            zc* acc = (zc*)malloc(Nc * sizeof(zc));
            int32_t list[512];
            for (int64_t c = 0; c < Nc; ++c) mark[c] = -1;
#pragma omp for schedule(static)
            for (int64_t Irow = 0; Irow < Nc; ++Irow) {
                int cnt = 0;
                for (int64_t pr = R->rowptr[Irow]; pr < R->rowptr[Irow + 1]; ++pr) {
                    const int32_t i = R->col[pr];
                    const zc rv = R->val[pr];
                    for (int64_t pa = A->rowptr[i]; pa < A->rowptr[i + 1]; ++pa) {
                        const int32_t k = A->col[pa];
                        const zc ra = rv * A->val[pa];
                        for (int64_t pp = P->rowptr[k]; pp < P->rowptr[k + 1]; ++pp) {
                            const int32_t J = P->col[pp];
                            if (mark[J] != Irow) {
                                mark[J] = Irow;
                                acc[J] = 0;
                                list[cnt++] = J;
                            }
                            acc[J] += ra * P->val[pp];
                        }
                    }
                }
                if (pass == 0) {
                    Ac->rowptr[Irow + 1] = cnt;
                } else {
                    qsort(list, cnt, sizeof(int32_t), cmp_i32);
                    int64_t q = Ac->rowptr[Irow];
                    for (int t = 0; t < cnt; ++t) {
                        Ac->col[q] = list[t];
                        Ac->val[q++] = acc[list[t]];
                    }
                }
            }
            free(mark);
            free(acc);
        }
        if (pass == 0) {
            for (int64_t c = 0; c < Nc; ++c) Ac->rowptr[c + 1] += Ac->rowptr[c];
            Ac->col = (int32_t*)malloc((size_t)Ac->rowptr[Nc] * sizeof(int32_t));
            Ac->val = (zc*)malloc((size_t)Ac->rowptr[Nc] * sizeof(zc));
        }
    }
}

static void level_dinv(Level* L, double relax_param) {
    L->dinv = (zc*)malloc(L->N * sizeof(zc));
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < L->N; ++i) {
        zc d = 0;
        for (int64_t p = L->A.rowptr[i]; p < L->A.rowptr[i + 1]; ++p)
            if (L->A.col[p] == i) d = L->A.val[p];
        L->dinv[i] = relax_param / d;
    }
}

Oracle* horc_create(int dim, const int64_t* n_nodes, const double* h, const double* m, const double* gamma, double wre,
                    double wim, int neumann_top, int sommerfeld, int order_bc, double shift, int levels,
                    double relax_param, int npre, int npost, int cycle, int coarse_iters) {
    const double t0 = now();
    Oracle* o = (Oracle*)calloc(1, sizeof(Oracle));
    o->dim = dim;
    o->levels = levels;
    o->npre = npre;
    o->npost = npost;
    o->cycle = cycle;
    o->coarse_iters = coarse_iters;
    o->relax_param = relax_param;
    o->shift = shift;
    o->wre = wre;
    o->L = (Level*)calloc(levels, sizeof(Level));
    Level* L0 = &o->L[0];
    L0->dim = dim;
    L0->N = 1;
    for (int d = 0; d < 3; ++d) {
        L0->n[d] = d < dim ? n_nodes[d] : 1;
        L0->N *= L0->n[d];
    }
    assemble_fine(o, L0, h, m, gamma, wre, wim, neumann_top, sommerfeld, order_bc, shift);
    o->shiftdiag = (zc*)malloc(L0->N * sizeof(zc));
    for (int64_t p = 0; p < L0->N; ++p) o->shiftdiag[p] = I * shift * wre * wre * m[p];
    for (int l = 0; l < levels; ++l) {
        Level* F = &o->L[l];
        level_dinv(F, relax_param);
        if (l == levels - 1) break;
        Level* C = &o->L[l + 1];
        C->dim = dim;
        C->N = 1;
        for (int d = 0; d < 3; ++d) {
            if (d < dim) {
                if (F->n[d] < 3 || (F->n[d] % 2) != 1) {
                    fprintf(stderr, "horc_create: cannot coarsen level %d\n", l);
                    return NULL;
                }
                C->n[d] = (F->n[d] + 1) / 2;
            } else {
                C->n[d] = 1;
            }
            C->N *= C->n[d];
        }
        build_transfers(F, C, dim);
        galerkin(&F->R, &F->A, &F->P, &C->A);
    }
    o->setup_seconds = now() - t0;
    return o;
}

void horc_destroy(Oracle* o) {
    if (!o) return;
    for (int l = 0; l < o->levels; ++l) {
        Level* L = &o->L[l];
        csr_free(&L->A);
        csr_free(&L->P);
        csr_free(&L->R);
        free(L->dinv);
        free(L->x);
        free(L->b);
        free(L->r);
    }
    free(o->L);
    free(o->shiftdiag);
    free(o);
}

double horc_setup_seconds(const Oracle* o) { return o->setup_seconds; }
int64_t horc_level_nnz(const Oracle* o, int level) { return o->L[level].A.rowptr[o->L[level].N]; }
int64_t horc_level_size(const Oracle* o, int level) { return o->L[level].N; }

static void ensure_work(Oracle* o, int nrhs) {
    if (nrhs <= o->kcap) return;
    for (int l = 0; l < o->levels; ++l) {
        Level* L = &o->L[l];
        free(L->x);
        free(L->b);
        free(L->r);
        L->x = (zc*)malloc((size_t)L->N * nrhs * sizeof(zc));
        L->b = (zc*)malloc((size_t)L->N * nrhs * sizeof(zc));
        L->r = (zc*)malloc((size_t)L->N * nrhs * sizeof(zc));
    }
    o->kcap = nrhs;
}

/* x += dinv .* (b - A x), nsweeps times; first sweep from zero: x = dinv .* b */
static void jacobi(const Level* L, zc* x, const zc* b, zc* r, int nrhs, int nsweeps, int x_is_zero) {
    const int64_t N = L->N;
    for (int s = 0; s < nsweeps; ++s) {
        if (s == 0 && x_is_zero) {
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < N; ++i)
                for (int c = 0; c < nrhs; ++c) x[(int64_t)c * N + i] = L->dinv[i] * b[(int64_t)c * N + i];
        } else {
            memcpy(r, b, (size_t)N * nrhs * sizeof(zc));
            spmv(&L->A, x, r, nrhs, -1.0, 1.0);
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < N; ++i)
                for (int c = 0; c < nrhs; ++c) x[(int64_t)c * N + i] += L->dinv[i] * r[(int64_t)c * N + i];
        }
    }
}

static zc zdot(const zc* a, const zc* b, int64_t n) {
    double re = 0, im = 0;
#pragma omp parallel for schedule(static) reduction(+ : re, im)
    for (int64_t i = 0; i < n; ++i) {
        const zc v = conj(a[i]) * b[i];
        re += creal(v);
        im += cimag(v);
    }
    return re + im * I;
}
static void zaxpy(zc* y, zc a, const zc* x, int64_t n) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) y[i] += a * x[i];
}

/* nsteps of Jacobi-preconditioned GMRES (one cycle, MGS) per column, from a zero guess */
static void coarse_gmres(const Level* L, const zc* b, zc* x, int nrhs, int nsteps) {
    const int64_t N = L->N;
    zc* V = (zc*)malloc((size_t)(nsteps + 1) * N * sizeof(zc));
    zc* z = (zc*)malloc((size_t)N * sizeof(zc));
    zc* H = (zc*)calloc((size_t)(nsteps + 1) * nsteps, sizeof(zc));
    zc* cs = (zc*)calloc(nsteps, sizeof(zc));
    zc* sn = (zc*)calloc(nsteps, sizeof(zc));
    zc* s = (zc*)calloc(nsteps + 1, sizeof(zc));
    zc* y = (zc*)calloc(nsteps, sizeof(zc));
    const int ldh = nsteps + 1;
    for (int c = 0; c < nrhs; ++c) {
        const zc* bc = b + (int64_t)c * N;
        zc* xc = x + (int64_t)c * N;
        memset(xc, 0, N * sizeof(zc));
        const double beta = sqrt(creal(zdot(bc, bc, N)));
        if (beta == 0.0) continue;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i) V[i] = bc[i] / beta;
        memset(s, 0, (nsteps + 1) * sizeof(zc));
        s[0] = beta;
        int jd = 0;
        for (int j = 0; j < nsteps; ++j) {
            zc* vj = V + (int64_t)j * N;
            zc* w = V + (int64_t)(j + 1) * N;
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < N; ++i) z[i] = L->dinv[i] * vj[i];
            spmv(&L->A, z, w, 1, 1.0, 0.0);
            for (int i = 0; i <= j; ++i) {
                const zc hij = zdot(V + (int64_t)i * N, w, N);
                H[i + j * ldh] = hij;
                zaxpy(w, -hij, V + (int64_t)i * N, N);
            }
            const double hn = sqrt(creal(zdot(w, w, N)));
            H[j + 1 + j * ldh] = hn;
            if (hn > 0) {
#pragma omp parallel for schedule(static)
                for (int64_t i = 0; i < N; ++i) w[i] /= hn;
            }
            for (int k = 0; k < j; ++k) {
                const zc t = cs[k] * H[k + j * ldh] + sn[k] * H[k + 1 + j * ldh];
                H[k + 1 + j * ldh] = -conj(sn[k]) * H[k + j * ldh] + cs[k] * H[k + 1 + j * ldh];
                H[k + j * ldh] = t;
            }
            const zc a = H[j + j * ldh];
            const double aa = cabs(a), den = sqrt(aa * aa + hn * hn);
            if (aa == 0) {
                cs[j] = 0;
                sn[j] = 1;
            } else {
                cs[j] = aa / den;
                sn[j] = (a / aa) * hn / den;
            }
            H[j + j * ldh] = cs[j] * a + sn[j] * hn;
            H[j + 1 + j * ldh] = 0;
            s[j + 1] = -conj(sn[j]) * s[j];
            s[j] = cs[j] * s[j];
            jd = j + 1;
            if (hn == 0) break;
        }
        for (int i = jd - 1; i >= 0; --i) {
            zc acc = s[i];
            for (int k = i + 1; k < jd; ++k) acc -= H[i + k * ldh] * y[k];
            y[i] = acc / H[i + i * ldh];
        }
        /* x = dinv .* (V y) */
        memset(z, 0, N * sizeof(zc));
        for (int i = 0; i < jd; ++i) zaxpy(z, y[i], V + (int64_t)i * N, N);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i) xc[i] = L->dinv[i] * z[i];
    }
    free(V);
    free(z);
    free(H);
    free(cs);
    free(sn);
    free(s);
    free(y);
}

static void cycle(Oracle* o, int l, const zc* b, zc* x, int x_is_zero, int nrhs) {
    Level* F = &o->L[l];
    if (l == o->levels - 1) {
        coarse_gmres(F, b, x, nrhs, o->coarse_iters);
        return;
    }
    Level* C = &o->L[l + 1];
    jacobi(F, x, b, F->r, nrhs, o->npre, x_is_zero);
    memcpy(F->r, b, (size_t)F->N * nrhs * sizeof(zc));
    spmv(&F->A, x, F->r, nrhs, -1.0, 1.0);
    spmv(&F->R, F->r, C->b, nrhs, 1.0, 0.0);
    if (l + 1 == o->levels - 1) {
        coarse_gmres(C, C->b, C->x, nrhs, o->coarse_iters);
    } else {
        cycle(o, l + 1, C->b, C->x, 1, nrhs);
        if (o->cycle == 1) cycle(o, l + 1, C->b, C->x, 0, nrhs);
    }
    spmv(&F->P, C->x, x, nrhs, 1.0, 1.0);
    jacobi(F, x, b, F->r, nrhs, o->npost, 0);
}

void horc_cycle(Oracle* o, const zc* B, zc* Z, int nrhs) {
    ensure_work(o, nrhs);
    if (o->levels == 1) coarse_gmres(&o->L[0], B, Z, nrhs, o->coarse_iters);
    else cycle(o, 0, B, Z, 1, nrhs);
    o->n_prec += nrhs;
}

/* Y = H X (shifted = 0: SH X - shiftdiag .* X, the reference's Afun) or SH X */
void horc_apply(Oracle* o, const zc* X, zc* Y, int nrhs, int shifted) {
    const int64_t N = o->L[0].N;
    spmv(&o->L[0].A, X, Y, nrhs, 1.0, 0.0);
    if (!shifted) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i)
            for (int c = 0; c < nrhs; ++c) Y[(int64_t)c * N + i] -= o->shiftdiag[i] * X[(int64_t)c * N + i];
    }
}

/* Right-preconditioned restarted FGMRES(inner) on every column (MGS, Givens residual estimate), all columns
 * advancing together so that every sparse product runs over the whole N x nrhs block like the reference's
 * multi-RHS SpMatMul.  Stops a column at ||r||/||b|| <= tol; `max_prec` (> 0) bounds the number of
 * preconditioner applications per column (the bounded timing sample of bench.py).  Returns wall seconds. */
double horc_solve_fgmres(Oracle* o, const zc* B, zc* X, int nrhs, int inner, int max_cycles, double tol, int max_prec,
                         int32_t* iters, double* relres) {
    const double t0 = now();
    const int64_t N = o->L[0].N;
    ensure_work(o, nrhs);
    const int m = inner, ldh = m + 1;
    zc* V = (zc*)malloc((size_t)(m + 1) * N * nrhs * sizeof(zc));
    zc* Z = (zc*)malloc((size_t)m * N * nrhs * sizeof(zc));
    zc* H = (zc*)calloc((size_t)ldh * m * nrhs, sizeof(zc));
    zc* cs = (zc*)calloc((size_t)m * nrhs, sizeof(zc));
    zc* sn = (zc*)calloc((size_t)m * nrhs, sizeof(zc));
    zc* s = (zc*)calloc((size_t)ldh * nrhs, sizeof(zc));
    zc* y = (zc*)calloc((size_t)m, sizeof(zc));
    double* bn = (double*)calloc(nrhs, sizeof(double));
    int* done = (int*)calloc(nrhs, sizeof(int));
    int* jd = (int*)calloc(nrhs, sizeof(int));
    const int64_t vs = N * nrhs;
    memset(X, 0, (size_t)vs * sizeof(zc));
    for (int c = 0; c < nrhs; ++c) {
        bn[c] = sqrt(creal(zdot(B + (int64_t)c * N, B + (int64_t)c * N, N)));
        iters[c] = 0;
        relres[c] = bn[c] == 0 ? 0 : 1;
        done[c] = bn[c] == 0;
    }
    memcpy(V, B, (size_t)vs * sizeof(zc)); /* r0 = b */
    int nprec = 0, stop = 0;
    for (int cyc = 0; cyc < max_cycles && !stop; ++cyc) {
        for (int c = 0; c < nrhs; ++c) {
            zc* v0 = V + (int64_t)c * N;
            const double beta = done[c] ? 0.0 : sqrt(creal(zdot(v0, v0, N)));
            memset(s + (int64_t)c * ldh, 0, ldh * sizeof(zc));
            s[(int64_t)c * ldh] = beta;
            jd[c] = 0;
            if (!done[c] && cyc > 0) {
                relres[c] = beta / bn[c];
                if (relres[c] <= tol) done[c] = 1;
            }
            const double sc = (done[c] || beta == 0) ? 0.0 : 1.0 / beta;
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < N; ++i) v0[i] *= sc;
        }
        int all = 1;
        for (int c = 0; c < nrhs; ++c) all &= done[c];
        if (all) break;
        for (int j = 0; j < m; ++j) {
            zc* Vj = V + (int64_t)j * vs;
            zc* Zj = Z + (int64_t)j * vs;
            zc* W = V + (int64_t)(j + 1) * vs;
            horc_cycle(o, Vj, Zj, nrhs);
            horc_apply(o, Zj, W, nrhs, 0);
            ++nprec;
            for (int c = 0; c < nrhs; ++c) {
                if (done[c]) continue;
                zc* w = W + (int64_t)c * N;
                zc* Hc = H + (int64_t)c * ldh * m + (int64_t)j * ldh;
                zc* csc = cs + (int64_t)c * m;
                zc* snc = sn + (int64_t)c * m;
                zc* sc_ = s + (int64_t)c * ldh;
                for (int i = 0; i <= j; ++i) {
                    const zc* vi = V + (int64_t)i * vs + (int64_t)c * N;
                    Hc[i] = zdot(vi, w, N);
                    zaxpy(w, -Hc[i], vi, N);
                }
                const double hn = sqrt(creal(zdot(w, w, N)));
                Hc[j + 1] = hn;
                const double isc = hn > 0 ? 1.0 / hn : 0.0;
#pragma omp parallel for schedule(static)
                for (int64_t i = 0; i < N; ++i) w[i] *= isc;
                for (int k = 0; k < j; ++k) {
                    const zc t = csc[k] * Hc[k] + snc[k] * Hc[k + 1];
                    Hc[k + 1] = -conj(snc[k]) * Hc[k] + csc[k] * Hc[k + 1];
                    Hc[k] = t;
                }
                const zc a = Hc[j];
                const double aa = cabs(a), den = sqrt(aa * aa + hn * hn);
                if (aa == 0) {
                    csc[j] = 0;
                    snc[j] = 1;
                } else {
                    csc[j] = aa / den;
                    snc[j] = (a / aa) * hn / den;
                }
                Hc[j] = csc[j] * a + snc[j] * hn;
                Hc[j + 1] = 0;
                sc_[j + 1] = -conj(snc[j]) * sc_[j];
                sc_[j] = csc[j] * sc_[j];
                relres[c] = cabs(sc_[j + 1]) / bn[c];
                jd[c] = j + 1;
                iters[c] += 1;
                if (relres[c] <= tol) done[c] = 1;
            }
            all = 1;
            for (int c = 0; c < nrhs; ++c) all &= done[c];
            if (max_prec > 0 && nprec >= max_prec) stop = 1;
            if (all || stop) break;
        }
        for (int c = 0; c < nrhs; ++c) {
            const int J = jd[c];
            const zc* Hc = H + (int64_t)c * ldh * m;
            const zc* sc_ = s + (int64_t)c * ldh;
            for (int i = J - 1; i >= 0; --i) {
                zc acc = sc_[i];
                for (int k = i + 1; k < J; ++k) acc -= Hc[i + (int64_t)k * ldh] * y[k];
                y[i] = acc / Hc[i + (int64_t)i * ldh];
            }
            for (int i = 0; i < J; ++i) zaxpy(X + (int64_t)c * N, y[i], Z + (int64_t)i * vs + (int64_t)c * N, N);
        }
        all = 1;
        for (int c = 0; c < nrhs; ++c) all &= done[c];
        if (all || stop) break;
        /* restart: r = b - H x */
        horc_apply(o, X, V, nrhs, 0);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < vs; ++i) V[i] = B[i] - V[i];
    }
    free(V);
    free(Z);
    free(H);
    free(cs);
    free(sn);
    free(s);
    free(y);
    free(bn);
    free(done);
    free(jd);
    return now() - t0;
}

int horc_num_threads(void) { return omp_get_max_threads(); }
/* torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm (rank 0 only) asks for all host cores explicitly */
void horc_set_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
}
