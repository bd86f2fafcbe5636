/* ip_oracle.c -- C restatement of the reference's interior-penalty assembly and of
 * the COO -> sparse -> A*x path.  TEST / BASELINE INFRASTRUCTURE ONLY: nothing in
 * stefanproblem_b200/ links or calls this.  It is (a) cross-checked against the
 * NumPy oracle (tests/test_oracle_c.py) and (b) the CPU baseline that bench.py
 * times (`cpu_baseline.kind = "port"`, and `--impl reference`).
 *
 * Follows /root/reference/StefanDG/InteriorPenalty/:
 *   gradient_operator              IP_gradient_operator.jl:1-9
 *   flux_operator                  IP_flux_operator.jl:1-33
 *   assemble_face_flux_operator!   IP_flux_operator.jl:35-118
 *   boundary face flux             IP_flux_operator.jl:330-357
 *   mass_operator                  IP_penalty_operator.jl:1-18
 *   assemble_face_penalty_operator! IP_penalty_operator.jl:21-65
 *   assemble_face_robin_operator!  IP_robin_interface_operator.jl:1-45
 *   driver loops                   IP_assembly.jl:1-80 (per-cell loops, one face from
 *                                  the lower cell id), plus the mesh-aligned interface
 *                                  extension of oracle/ip.py
 *   CutCellDG.assemble_couple_cell_matrix! / sparse_operator  
 * As in the reference, every element matrix is recomputed per cell / per face by
 * looping over quadrature points, and appended to a growable COO list; the list is
 * then compressed (duplicates summed, zeros dropped) and applied as a CSR SpMV.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>
#include <unistd.h>

#define MAXN1 4
#define MAXNF 16
#define MAXQ 8

typedef struct {
  int64_t n, cap;
  int64_t *rows, *cols;
  double* vals;
} coo_t;

typedef struct {
  int64_t nrows, nnz;
  int64_t* indptr;
  int32_t* indices;
  double* data;
  double assemble_seconds, compress_seconds;
  int64_t coo_entries;
} csr_t;

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* minimal parallel-for over row ranges (this image's gcc has no libgomp) */
typedef void (*range_fn)(int64_t lo, int64_t hi, void* ctx);
typedef struct {
  range_fn fn;
  int64_t lo, hi;
  void* ctx;
} task_t;
static void* run_task(void* p) {
  task_t* t = (task_t*)p;
  t->fn(t->lo, t->hi, t->ctx);
  return NULL;
}
static int hw_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}
static void parallel_for(int64_t n, int nthreads, range_fn fn, void* ctx) {
  if (nthreads <= 0) nthreads = hw_threads();
  if (nthreads > 256) nthreads = 256;
  if (nthreads == 1 || n < 1024) {
    fn(0, n, ctx);
    return;
  }
  pthread_t th[256];
  task_t tk[256];
  for (int t = 0; t < nthreads; ++t) {
    tk[t].fn = fn;
    tk[t].lo = n * t / nthreads;
    tk[t].hi = n * (t + 1) / nthreads;
    tk[t].ctx = ctx;
    pthread_create(&th[t], NULL, run_task, &tk[t]);
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
}

static void coo_reserve(coo_t* c, int64_t extra) {
  if (c->n + extra <= c->cap) return;
  int64_t cap = c->cap ? c->cap : 1 << 20;
  while (cap < c->n + extra) cap *= 2;
  c->rows = (int64_t*)realloc(c->rows, cap * sizeof(int64_t));
  c->cols = (int64_t*)realloc(c->cols, cap * sizeof(int64_t));
  c->vals = (double*)realloc(c->vals, cap * sizeof(double));
  c->cap = cap;
}

/* ---------------------------------------------------------------- basis */
static void gauss_legendre(int n, double* x, double* w) {
  for (int i = 0; i < n; ++i) {
    double z = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 0.0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; ++j) {
        double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0);
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      double dz = p1 / pp;
      z -= dz;
      if (fabs(dz) < 1e-16) break;
    }
    x[n - 1 - i] = z;
    w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
}

static void lagrange_1d(int order, double x, double* val, double* der) {
  int n = order + 1;
  double nodes[MAXN1];
  for (int a = 0; a < n; ++a) nodes[a] = -1.0 + 2.0 * a / order;
  for (int a = 0; a < n; ++a) {
    double v = 1.0, d = 0.0;
    for (int b = 0; b < n; ++b)
      if (b != a) v *= (x - nodes[b]) / (nodes[a] - nodes[b]);
    for (int c = 0; c < n; ++c) {
      if (c == a) continue;
      double t = 1.0 / (nodes[a] - nodes[c]);
      for (int b = 0; b < n; ++b)
        if (b != a && b != c) t *= (x - nodes[b]) / (nodes[a] - nodes[b]);
      d += t;
    }
    val[a] = v;
    der[a] = d;
  }
}

/* basis(p) and transform_gradient(gradient(basis, p), jac): kron(N(xi), N(eta)) */
static void basis2d(int order, const double* p, double* V) {
  int n = order + 1;
  double nx[MAXN1], dx[MAXN1], ny[MAXN1], dy[MAXN1];
  lagrange_1d(order, p[0], nx, dx);
  lagrange_1d(order, p[1], ny, dy);
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) V[a * n + b] = nx[a] * ny[b];
}

static void gradient2d(int order, const double* p, const double* jac, double* G /* NF x 2 */) {
  int n = order + 1;
  double nx[MAXN1], dx[MAXN1], ny[MAXN1], dy[MAXN1];
  lagrange_1d(order, p[0], nx, dx);
  lagrange_1d(order, p[1], ny, dy);
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) {
      G[(a * n + b) * 2 + 0] = dx[a] * ny[b] / jac[0];
      G[(a * n + b) * 2 + 1] = nx[a] * dy[b] / jac[1];
    }
}

typedef struct {
  int n;
  double p[MAXQ * MAXQ][2];
  double w[MAXQ * MAXQ];
} quadrule_t;

static void cell_quadrature(int nq, quadrule_t* q) {
  double x[MAXQ], w[MAXQ];
  gauss_legendre(nq, x, w);
  q->n = nq * nq;
  for (int i = 0; i < nq; ++i)
    for (int j = 0; j < nq; ++j) {
      q->p[i * nq + j][0] = x[i];
      q->p[i * nq + j][1] = x[j];
      q->w[i * nq + j] = w[i] * w[j];
    }
}

static void face_quadrature(int faceid, int nq, quadrule_t* q) {
  double x[MAXQ], w[MAXQ];
  gauss_legendre(nq, x, w);
  q->n = nq;
  for (int i = 0; i < nq; ++i) {
    q->w[i] = w[i];
    switch (faceid) {
      case 0: q->p[i][0] = x[i]; q->p[i][1] = -1.0; break;
      case 1: q->p[i][0] = 1.0; q->p[i][1] = x[i]; break;
      case 2: q->p[i][0] = x[i]; q->p[i][1] = 1.0; break;
      default: q->p[i][0] = -1.0; q->p[i][1] = x[i]; break;
    }
  }
}

static const double kNormals[4][2] = {{0, -1}, {1, 0}, {0, 1}, {-1, 0}};
static const int kOpposite[4] = {2, 3, 0, 1};

/* ------------------------------------------------------ element matrices */
/* IP_gradient_operator.jl:1-9 */
static void gradient_operator(int order, const quadrule_t* quad, double k, const double* jac,
                              double detjac, double* M) {
  int nf = (order + 1) * (order + 1);
  double G[MAXNF * 2];
  memset(M, 0, sizeof(double) * nf * nf);
  for (int q = 0; q < quad->n; ++q) {
    gradient2d(order, quad->p[q], jac, G);
    for (int a = 0; a < nf; ++a)
      for (int b = 0; b < nf; ++b)
        M[a * nf + b] += k * (G[a * 2] * G[b * 2] + G[a * 2 + 1] * G[b * 2 + 1]) * detjac * quad->w[q];
  }
}

/* IP_flux_operator.jl:1-33 (constant normal and scale area on mesh faces) */
static void flux_operator(int order, const quadrule_t* q1, const quadrule_t* q2, const double* normal,
                          double k, const double* jac, double sa, double* M) {
  int nf = (order + 1) * (order + 1);
  double G[MAXNF * 2], V[MAXNF];
  memset(M, 0, sizeof(double) * nf * nf);
  for (int i = 0; i < q1->n; ++i) {
    gradient2d(order, q1->p[i], jac, G);
    basis2d(order, q2->p[i], V);
    for (int a = 0; a < nf; ++a) {
      double gn = G[a * 2] * normal[0] + G[a * 2 + 1] * normal[1];
      for (int b = 0; b < nf; ++b) M[a * nf + b] += k * gn * V[b] * sa * q1->w[i];
    }
  }
}

/* IP_penalty_operator.jl:1-18 */
static void mass_operator(int order, const quadrule_t* q1, const quadrule_t* q2, double sa, double* M) {
  int nf = (order + 1) * (order + 1);
  double V1[MAXNF], V2[MAXNF];
  memset(M, 0, sizeof(double) * nf * nf);
  for (int i = 0; i < q1->n; ++i) {
    basis2d(order, q1->p[i], V1);
    basis2d(order, q2->p[i], V2);
    for (int a = 0; a < nf; ++a)
      for (int b = 0; b < nf; ++b) M[a * nf + b] += V1[a] * V2[b] * sa * q1->w[i];
  }
}

/* CutCellDG.assemble_couple_cell_matrix!(sys, nodeids1, nodeids2, 1, vec(M)):
 * rows from cell1, cols from cell2; M given row-major here, appended column-major */
static void assemble_couple(coo_t* c, int nf, int64_t cell1, int64_t cell2, const double* M,
                            double scale) {
  coo_reserve(c, nf * nf);
  for (int b = 0; b < nf; ++b)
    for (int a = 0; a < nf; ++a) {
      c->rows[c->n] = cell1 * nf + a;
      c->cols[c->n] = cell2 * nf + b;
      c->vals[c->n] = scale * M[a * nf + b];
      c->n++;
    }
}

static void add_t(int nf, double sa, const double* A, double sb, const double* B /* transposed */,
                  double* out) {
  for (int a = 0; a < nf; ++a)
    for (int b = 0; b < nf; ++b) out[a * nf + b] = sa * A[a * nf + b] + sb * B[b * nf + a];
}

/* IP_flux_operator.jl:35-118 */
static void assemble_face_flux_operator(coo_t* c, int order, const quadrule_t* q1, const quadrule_t* q2,
                                        const double* normal, double k1, double k2,
                                        const double* jac, double sa, int64_t cell1, int64_t cell2) {
  int nf = (order + 1) * (order + 1);
  double M11[MAXNF * MAXNF], M12[MAXNF * MAXNF], M21[MAXNF * MAXNF], M22[MAXNF * MAXNF];
  double T[MAXNF * MAXNF];
  memset(T, 0, sizeof(T));
  flux_operator(order, q1, q1, normal, k1, jac, sa, M11);
  flux_operator(order, q1, q2, normal, k1, jac, sa, M12);
  flux_operator(order, q2, q1, normal, k2, jac, sa, M21);
  flux_operator(order, q2, q2, normal, k2, jac, sa, M22);
  for (int i = 0; i < nf * nf; ++i) {
    M11[i] *= -0.5;
    M12[i] *= -0.5;
    M21[i] *= -0.5;
    M22[i] *= -0.5;
  }
  add_t(nf, 1.0, M11, 1.0, M11, T);
  assemble_couple(c, nf, cell1, cell1, T, 1.0);
  add_t(nf, 1.0, M21, -1.0, M12, T);
  assemble_couple(c, nf, cell2, cell1, T, 1.0);
  add_t(nf, -1.0, M12, 1.0, M21, T);
  assemble_couple(c, nf, cell1, cell2, T, 1.0);
  add_t(nf, -1.0, M22, -1.0, M22, T);
  assemble_couple(c, nf, cell2, cell2, T, 1.0);
}

/* IP_penalty_operator.jl:21-65 and IP_robin_interface_operator.jl:1-45 */
static void assemble_face_mass_blocks(coo_t* c, int order, const quadrule_t* q1, const quadrule_t* q2,
                                      double coef, double offsign, double sa, int64_t cell1,
                                      int64_t cell2) {
  int nf = (order + 1) * (order + 1);
  double M[MAXNF * MAXNF];
  mass_operator(order, q1, q1, sa, M);
  assemble_couple(c, nf, cell1, cell1, M, coef);
  mass_operator(order, q1, q2, sa, M);
  assemble_couple(c, nf, cell1, cell2, M, offsign * coef);
  mass_operator(order, q2, q1, sa, M);
  assemble_couple(c, nf, cell2, cell1, M, offsign * coef);
  mass_operator(order, q2, q2, sa, M);
  assemble_couple(c, nf, cell2, cell2, M, coef);
}

enum { T_VOLUME = 1, T_IE_FLUX = 2, T_IE_PEN = 4, T_IF_FLUX = 8, T_IF_PEN = 16, T_B_FLUX = 32,
       T_B_PEN = 64, T_ROBIN = 128 };

static int64_t neighbour(int nx, int ny, int f, int64_t cell) {
  int64_t i = cell / ny, j = cell % ny;
  switch (f) {
    case 0: return j > 0 ? cell - 1 : -1;
    case 1: return i < nx - 1 ? cell + ny : -1;
    case 2: return j < ny - 1 ? cell + 1 : -1;
    default: return i > 0 ? cell - ny : -1;
  }
}



typedef struct {
  int64_t col;
  double val;
} entry_t;

static int cmp_entry(const void* a, const void* b) {
  int64_t ca = ((const entry_t*)a)->col, cb = ((const entry_t*)b)->col;
  return ca < cb ? -1 : (ca > cb ? 1 : 0);
}

/* dropzeros!(sparse(rows, cols, vals, N, N)) */
typedef struct {
  entry_t* ent;
  int64_t *cnt, *rownnz;
  csr_t* A;
} compress_ctx;

static void compress_rows(int64_t lo, int64_t hi, void* p) {
  compress_ctx* c = (compress_ctx*)p;
  for (int64_t r = lo; r < hi; ++r) {
    int64_t b = c->cnt[r], e = c->cnt[r + 1], k = b;
    qsort(c->ent + b, e - b, sizeof(entry_t), cmp_entry);
    for (int64_t i = b; i < e;) {
      int64_t col = c->ent[i].col;
      double v = 0.0;
      while (i < e && c->ent[i].col == col) v += c->ent[i++].val;
      if (v != 0.0) {
        c->ent[k].col = col;
        c->ent[k].val = v;
        k++;
      }
    }
    c->rownnz[r] = k - b;
  }
}

static void compress_copy(int64_t lo, int64_t hi, void* p) {
  compress_ctx* c = (compress_ctx*)p;
  for (int64_t r = lo; r < hi; ++r)
    for (int64_t i = 0; i < c->rownnz[r]; ++i) {
      c->A->indices[c->A->indptr[r] + i] = (int32_t)c->ent[c->cnt[r] + i].col;
      c->A->data[c->A->indptr[r] + i] = c->ent[c->cnt[r] + i].val;
    }
}

static void compress(const coo_t* c, int64_t nrows, csr_t* A) {
  int64_t* cnt = (int64_t*)calloc(nrows + 1, sizeof(int64_t));
  for (int64_t e = 0; e < c->n; ++e) cnt[c->rows[e] + 1]++;
  for (int64_t r = 0; r < nrows; ++r) cnt[r + 1] += cnt[r];
  entry_t* ent = (entry_t*)malloc(sizeof(entry_t) * (c->n ? c->n : 1));
  int64_t* pos = (int64_t*)malloc(sizeof(int64_t) * nrows);
  memcpy(pos, cnt, sizeof(int64_t) * nrows);
  for (int64_t e = 0; e < c->n; ++e) {
    int64_t r = c->rows[e];
    ent[pos[r]].col = c->cols[e];
    ent[pos[r]].val = c->vals[e];
    pos[r]++;
  }
  A->nrows = nrows;
  A->indptr = (int64_t*)malloc(sizeof(int64_t) * (nrows + 1));
  int64_t* rownnz = (int64_t*)calloc(nrows, sizeof(int64_t));
  compress_ctx ctx = {ent, cnt, rownnz, A};
  parallel_for(nrows, 0, compress_rows, &ctx);
  A->indptr[0] = 0;
  for (int64_t r = 0; r < nrows; ++r) A->indptr[r + 1] = A->indptr[r] + rownnz[r];
  A->nnz = A->indptr[nrows];
  A->indices = (int32_t*)malloc(sizeof(int32_t) * (A->nnz ? A->nnz : 1));
  A->data = (double*)malloc(sizeof(double) * (A->nnz ? A->nnz : 1));
  parallel_for(nrows, 0, compress_copy, &ctx);
  free(cnt);
  free(ent);
  free(pos);
  free(rownnz);
}

/* The whole of assemble_interior_penalty_linear_system! (+ aligned interface and
 * Robin passes), then sparse_operator. */
csr_t* ip_oracle_assemble(int order, int numqp, int nx, int ny, double hx, double hy, double k1,
                          double k2, double penalty, double lambda, int variant, int terms,
                          const int8_t* phase) {
  const int nf = (order + 1) * (order + 1);
  const int64_t ncells = (int64_t)nx * ny;
  if ((int64_t)ncells * nf > 2000000000LL) return NULL;
  const double jac[2] = {0.5 * hx, 0.5 * hy};
  const double detjac = jac[0] * jac[1];
  const double facedetjac[4] = {jac[0], jac[1], jac[0], jac[1]};
  quadrule_t cq, fq[4];
  cell_quadrature(numqp, &cq);
  for (int f = 0; f < 4; ++f) face_quadrature(f, numqp, &fq[f]);
  coo_t c = {0, 0, NULL, NULL, NULL};
  double M[MAXNF * MAXNF], T[MAXNF * MAXNF];
  memset(T, 0, sizeof(T));
  double t0 = now_s();
#define SIGN(cell) (phase ? (int)phase[cell] : 1)
#define COND(s) ((s) > 0 ? k1 : k2)
  /* assemble_gradient_operator!  IP_gradient_operator.jl:29-57 */
  if (terms & T_VOLUME)
    for (int64_t cell = 0; cell < ncells; ++cell) {
      gradient_operator(order, &cq, COND(SIGN(cell)), jac, detjac, M);
      assemble_couple(&c, nf, cell, cell, M, 1.0);
    }
  /* assemble_interelement_flux_operator!  IP_flux_operator.jl:202-256 */
  if (terms & T_IE_FLUX)
    for (int64_t cell = 0; cell < ncells; ++cell)
      for (int f = 0; f < 4; ++f) {
        int64_t nbr = neighbour(nx, ny, f, cell);
        if (nbr >= 0 && cell < nbr && SIGN(nbr) == SIGN(cell)) {
          double k = COND(SIGN(cell));
          assemble_face_flux_operator(&c, order, &fq[f], &fq[kOpposite[f]], kNormals[f], k, k, jac,
                                      facedetjac[f], cell, nbr);
        }
      }
  /* assemble_interelement_penalty_operator!  IP_penalty_operator.jl:134-181 */
  if (terms & T_IE_PEN)
    for (int64_t cell = 0; cell < ncells; ++cell)
      for (int f = 0; f < 4; ++f) {
        int64_t nbr = neighbour(nx, ny, f, cell);
        if (nbr >= 0 && cell < nbr && SIGN(nbr) == SIGN(cell))
          assemble_face_mass_blocks(&c, order, &fq[f], &fq[kOpposite[f]], penalty, -1.0,
                                    facedetjac[f], cell, nbr);
      }
  /* mesh-aligned interface: side 1 = the +1 cell (oracle/ip.py) */
  for (int pass = 0; pass < 3; ++pass) {
    int bit = pass == 0 ? T_IF_FLUX : (pass == 1 ? T_IF_PEN : T_ROBIN);
    if (!(terms & bit)) continue;
    for (int64_t cell = 0; cell < ncells; ++cell) {
      if (SIGN(cell) != 1) continue;
      for (int f = 0; f < 4; ++f) {
        int64_t nbr = neighbour(nx, ny, f, cell);
        if (nbr < 0 || SIGN(nbr) != -1) continue;
        if (pass == 0)
          assemble_face_flux_operator(&c, order, &fq[f], &fq[kOpposite[f]], kNormals[f], k1, k2,
                                      jac, facedetjac[f], cell, nbr);
        else if (pass == 1)
          assemble_face_mass_blocks(&c, order, &fq[f], &fq[kOpposite[f]], penalty, -1.0,
                                    facedetjac[f], cell, nbr);
        else
          assemble_face_mass_blocks(&c, order, &fq[f], &fq[kOpposite[f]], 0.25 * lambda, 1.0,
                                    facedetjac[f], cell, nbr);
      }
    }
  }
  /* assemble_boundary_flux_operator!  IP_flux_operator.jl:393-444 */
  if (terms & T_B_FLUX)
    for (int64_t cell = 0; cell < ncells; ++cell)
      for (int f = 0; f < 4; ++f)
        if (neighbour(nx, ny, f, cell) < 0) {
          flux_operator(order, &fq[f], &fq[f], kNormals[f], COND(SIGN(cell)), jac, facedetjac[f], M);
          if (variant == 0)
            add_t(nf, -1.0, M, -1.0, M, T); /* -(M + M') */
          else
            add_t(nf, 0.0, M, -1.0, M, T); /* legacy: -M' */
          assemble_couple(&c, nf, cell, cell, T, 1.0);
        }
  /* assemble_boundary_penalty_operator!  IP_penalty_operator.jl:286-329 */
  if (terms & T_B_PEN)
    for (int64_t cell = 0; cell < ncells; ++cell)
      for (int f = 0; f < 4; ++f)
        if (neighbour(nx, ny, f, cell) < 0) {
          mass_operator(order, &fq[f], &fq[f], facedetjac[f], M);
          assemble_couple(&c, nf, cell, cell, M, penalty);
        }
  double t1 = now_s();
  csr_t* A = (csr_t*)calloc(1, sizeof(csr_t));
  compress(&c, ncells * nf, A);
  double t2 = now_s();
  A->assemble_seconds = t1 - t0;
  A->compress_seconds = t2 - t1;
  A->coo_entries = c.n;
  free(c.rows);
  free(c.cols);
  free(c.vals);
  return A;
}

void ip_oracle_info(const csr_t* A, int64_t* nrows, int64_t* nnz, int64_t* coo_entries,
                    double* assemble_seconds, double* compress_seconds) {
  *nrows = A->nrows;
  *nnz = A->nnz;
  *coo_entries = A->coo_entries;
  *assemble_seconds = A->assemble_seconds;
  *compress_seconds = A->compress_seconds;
}

void ip_oracle_csr(const csr_t* A, int64_t* indptr, int32_t* indices, double* data) {
  memcpy(indptr, A->indptr, sizeof(int64_t) * (A->nrows + 1));
  memcpy(indices, A->indices, sizeof(int32_t) * A->nnz);
  memcpy(data, A->data, sizeof(double) * A->nnz);
}

/* y = A x; nthreads <= 0: all hardware threads (Julia's SparseArrays `*` is serial: 1) */
typedef struct {
  const csr_t* A;
  const double* x;
  double* y;
} spmv_ctx;

static void spmv_rows(int64_t lo, int64_t hi, void* p) {
  spmv_ctx* c = (spmv_ctx*)p;
  const csr_t* A = c->A;
  for (int64_t r = lo; r < hi; ++r) {
    double v = 0.0;
    for (int64_t i = A->indptr[r]; i < A->indptr[r + 1]; ++i) v += A->data[i] * c->x[A->indices[i]];
    c->y[r] = v;
  }
}

void ip_oracle_spmv(const csr_t* A, const double* x, double* y, int nthreads) {
  spmv_ctx ctx = {A, x, y};
  parallel_for(A->nrows, nthreads, spmv_rows, &ctx);
}

int ip_oracle_max_threads(void) { return hw_threads(); }

void ip_oracle_free(csr_t* A) {
  if (!A) return;
  free(A->indptr);
  free(A->indices);
  free(A->data);
  free(A);
}
