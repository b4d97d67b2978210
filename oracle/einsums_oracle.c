/*
 * einsums_oracle.c — CPU ORACLE (test infrastructure; never part of the product path).
 * Plain-C restatement of the reference's host algorithms for einsum / permute / BLAS seam.
 * See einsums_oracle.h for the file:line map. Parity: PINNED by tests/test_oracle_golden.py.
 */
#define _GNU_SOURCE
#include "einsums_oracle.h"

#include <complex.h>
#include <ctype.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ index-pack algebra
 * TA/Detail/Utilities.hpp:609-751 (Intersect / Difference / Unique keep first-pack order) */
typedef struct { int n; int v[ORC_MAX_UNIQUE]; } pack;
typedef struct { int n; int label[4 * ORC_MAX_UNIQUE]; int pos[4 * ORC_MAX_UNIQUE]; } pospack;

static int has(const pack *p, int x) { for (int i = 0; i < p->n; i++) if (p->v[i] == x) return 1; return 0; }
static pack mk(int n, const int *v) { pack p; p.n = n; for (int i = 0; i < n; i++) p.v[i] = v[i]; return p; }
static pack intersect(const pack *a, const pack *b) { pack r; r.n = 0; for (int i = 0; i < a->n; i++) if (has(b, a->v[i])) r.v[r.n++] = a->v[i]; return r; }
static pack difference(const pack *a, const pack *b) { pack r; r.n = 0; for (int i = 0; i < a->n; i++) if (!has(b, a->v[i])) r.v[r.n++] = a->v[i]; return r; }
static pack unique(const pack *a) { pack r; r.n = 0; for (int i = 0; i < a->n; i++) if (!has(&r, a->v[i])) r.v[r.n++] = a->v[i]; return r; }
static int same_pack(const pack *a, const pack *b) { if (a->n != b->n) return 0; for (int i = 0; i < a->n; i++) if (a->v[i] != b->v[i]) return 0; return 1; }

/* find_type_with_position(Ts, Us), Utilities.hpp:63-152: for each T in order, every position of T in Us */
static pospack find_pos(const pack *ts, const pack *us) {
    pospack r; r.n = 0;
    for (int i = 0; i < ts->n; i++)
        for (int p = 0; p < us->n; p++)
            if (us->v[p] == ts->v[i]) { r.label[r.n] = ts->v[i]; r.pos[r.n] = p; r.n++; }
    return r;
}
/* contiguous_positions, Utilities.hpp:289-324 */
static int contiguous(const pospack *x) { for (int i = 0; i + 1 < x->n; i++) if (x->pos[i] != x->pos[i + 1] - 1) return 0; return 1; }
/* is_same_ordering, Utilities.hpp:326-343 */
static int same_ordering(const pospack *x, const pospack *y) {
    if (x->n == 0 || y->n == 0) return 0;
    if (x->n != y->n) return 0;
    for (int i = 0; i < x->n; i++) if (x->label[i] != y->label[i]) return 0;
    return 1;
}

typedef struct {
    pack C, A, B, linksAB, links, C_unique, link_unique, CminusA, CminusB;
    pospack link_in_A, link_in_B, target_in_A, target_in_B, A_target_in_C, B_target_in_C;
} analysis;

static void analyse(analysis *z, const pack *C, const pack *A, const pack *B) {
    z->C = *C; z->A = *A; z->B = *B;
    z->linksAB = intersect(A, B);
    z->links = difference(&z->linksAB, C);
    z->C_unique = unique(C);
    z->link_unique = unique(&z->links);
    z->CminusA = difference(C, A);
    z->CminusB = difference(C, B);
    z->link_in_A = find_pos(&z->link_unique, A);
    z->link_in_B = find_pos(&z->link_unique, B);
    z->target_in_A = find_pos(&z->C_unique, A);
    z->target_in_B = find_pos(&z->C_unique, B);
    z->A_target_in_C = find_pos(A, C);
    z->B_target_in_C = find_pos(B, C);
}

static int is_hadamard(const pack *p) { pack u = unique(p); return u.n != p->n; }

/* Dispatch.hpp:236-262 */
static int is_outer(const analysis *z, int conj_a, int conj_b) {
    if (!(z->linksAB.n == 0 && contiguous(&z->target_in_A) && contiguous(&z->target_in_B))) return 0;
    if (z->A_target_in_C.n == 0) return 0; /* std::get<1> of an empty tuple does not compile in the reference */
    int swap_ab = z->A_target_in_C.pos[0] != 0;
    return (!conj_a && !swap_ab) || (!conj_b && swap_ab);
}
/* Dispatch.hpp:407-439; (A = matrix pack, B = vector pack) */
static int is_matvec(const analysis *z, int conj_a, int conj_b) {
    int cond = contiguous(&z->link_in_A) && contiguous(&z->link_in_B) && contiguous(&z->target_in_A) &&
               same_ordering(&z->link_in_A, &z->link_in_B) && same_ordering(&z->target_in_A, &z->A_target_in_C) &&
               !same_ordering(&z->target_in_B, &z->B_target_in_C) && z->B_target_in_C.n == 0 && !conj_b;
    if (!cond) return 0;
    int transpose_a = z->link_in_A.pos[0] == 0;
    return transpose_a || !conj_a;
}
/* Dispatch.hpp:565-609 */
static int is_matmat(const analysis *z, int conj_a, int conj_b) {
    int cond = z->CminusA.n > 0 && z->CminusB.n > 0 && contiguous(&z->link_in_A) && contiguous(&z->link_in_B) &&
               contiguous(&z->target_in_A) && contiguous(&z->target_in_B) && contiguous(&z->A_target_in_C) &&
               contiguous(&z->B_target_in_C) && same_ordering(&z->link_in_A, &z->link_in_B) &&
               same_ordering(&z->target_in_A, &z->A_target_in_C) && same_ordering(&z->target_in_B, &z->B_target_in_C);
    if (!cond) return 0;
    int tA = z->link_in_A.pos[0] == 0, tB = z->link_in_B.pos[0] != 0, tC = z->A_target_in_C.pos[0] != 0;
    return (tC && (tA || !conj_a) && (tB || !conj_b)) || (!tC && (!tA || !conj_a) && (!tB || !conj_b));
}

/* 0 generic, 1 dot, 2 direct, 3 ger, 4 gemv (A matrix), 14 gemv swapped (B matrix), 5 gemm */
static int classify_static_ex(const pack *C, const pack *A, const pack *B, int conj_a, int conj_b, int same_types) {
    if (is_hadamard(C) || is_hadamard(A) || is_hadamard(B) || !same_types) return ORC_GENERIC; /* :896-899 */
    if (C->n == 0 && same_pack(A, B)) return ORC_DOT;                                        /* :900, :211-215 */
    if (same_pack(C, A) && same_pack(C, B) && !conj_a && !conj_b) return ORC_DIRECT;         /* :921, :220-231 */
    analysis z; analyse(&z, C, A, B);
    if (is_outer(&z, conj_a, conj_b)) return ORC_GER;                                        /* :931 */
    if (is_matvec(&z, conj_a, conj_b)) return ORC_GEMV;                                      /* :935 */
    analysis zs; analyse(&zs, C, B, A);
    if (is_matvec(&zs, conj_a, conj_b)) return 10 + ORC_GEMV;                                /* :939 (ConjA,ConjB passed unswapped) */
    if (C->n >= 2 && A->n >= 2 && B->n >= 2 && is_matmat(&z, conj_a, conj_b)) return ORC_GEMM; /* :945-951 */
    return ORC_GENERIC;
}

int orc_classify_static(int c_rank, const int *c_idx, int a_rank, const int *a_idx, int b_rank, const int *b_idx, int conj_a, int conj_b,
                        int same_types) {
    pack C = mk(c_rank, c_idx), A = mk(a_rank, a_idx), B = mk(b_rank, b_idx);
    int r = classify_static_ex(&C, &A, &B, conj_a, conj_b, same_types);
    return r >= 10 ? r - 10 : r;
}

/* ------------------------------------------------------------------ run-time view checks */
static int64_t t_size(const orc_tensor *t) { int64_t s = 1; for (int i = 0; i < t->rank; i++) s *= t->dim[i]; return s; }
static int t_row_major(const orc_tensor *t) { /* Tensor/DeviceTensor.hpp:1106-1116 */
    if (t->rank < 2) return 1;
    int64_t s0 = t->stride[0], sl = t->stride[t->rank - 1];
    if (s0 > sl) return 1;
    if (s0 == sl) return t->dim[0] > t->dim[t->rank - 1];
    return 0;
}
static int t_totally_vectorable(const orc_tensor *t) { /* Tensor/DeviceTensor.hpp:1130-1155 */
    if (t->rank == 0) return 0;
    if (t->rank == 1) return 1;
    int64_t s0 = t->stride[0], sl = t->stride[t->rank - 1];
    if (s0 < sl) return t->dim[t->rank - 1] * sl == t_size(t) * s0;
    return t->dim[0] * s0 == t_size(t) * sl;
}
static int64_t product_dims(const pospack *p, const orc_tensor *t) { int64_t r = 1; for (int i = 0; i < p->n; i++) r *= t->dim[p->pos[i]]; return r; }
static int64_t last_stride(const pospack *p, const orc_tensor *t) { return p->n ? t->stride[p->pos[p->n - 1]] : 0; }

/* the "gemmable view" walk, Dispatch.hpp:654-693 */
static int view_ok(const orc_tensor *t, int64_t s0, int64_t s1) {
    if (t->rank == 2) return !(t->stride[0] != 1 && t->stride[1] != 1);
    int64_t cutoff = s0 > s1 ? s0 : s1, check = 1;
    if (t_row_major(t)) {
        for (int i = t->rank - 1; i >= 0; i--) {
            if (t->stride[i] == cutoff) check = cutoff;
            if (t->stride[i] != check) return 0;
            check *= t->dim[i];
        }
    } else {
        for (int i = 0; i < t->rank; i++) {
            if (t->stride[i] == cutoff) check = cutoff;
            if (t->stride[i] != check) return 0;
            check *= t->dim[i];
        }
    }
    return 1;
}

typedef struct { int64_t dA[2], sA[2], dB[2], sB[2], dC[2], sC[2]; int tA, tB, tC; } mm_views;

static int do_matrix_product_check(const analysis *z, const orc_tensor *C, const orc_tensor *A, const orc_tensor *B, mm_views *v) {
    int tA = z->link_in_A.pos[0] == 0, tB = z->link_in_B.pos[0] != 0, tC = z->A_target_in_C.pos[0] != 0;
    int64_t t;
    v->dA[0] = product_dims(&z->A_target_in_C, C); v->dA[1] = product_dims(&z->link_in_A, A);
    v->sA[0] = last_stride(&z->target_in_A, A);    v->sA[1] = last_stride(&z->link_in_A, A);
    if (!view_ok(A, v->sA[0], v->sA[1])) return 0;
    if (tA) { t = v->dA[0]; v->dA[0] = v->dA[1]; v->dA[1] = t; t = v->sA[0]; v->sA[0] = v->sA[1]; v->sA[1] = t; }
    v->dB[0] = product_dims(&z->link_in_B, B);     v->dB[1] = product_dims(&z->B_target_in_C, C);
    v->sB[0] = last_stride(&z->link_in_B, B);      v->sB[1] = last_stride(&z->target_in_B, B);
    if (!view_ok(B, v->sB[0], v->sB[1])) return 0;
    if (tB) { t = v->dB[0]; v->dB[0] = v->dB[1]; v->dB[1] = t; t = v->sB[0]; v->sB[0] = v->sB[1]; v->sB[1] = t; }
    v->dC[0] = product_dims(&z->A_target_in_C, C); v->dC[1] = product_dims(&z->B_target_in_C, C);
    v->sC[0] = last_stride(&z->A_target_in_C, C);  v->sC[1] = last_stride(&z->B_target_in_C, C);
    if (!view_ok(C, v->sC[0], v->sC[1])) return 0;
    if (tC) { t = v->dC[0]; v->dC[0] = v->dC[1]; v->dC[1] = t; t = v->sC[0]; v->sC[0] = v->sC[1]; v->sC[1] = t; }
    if (v->sA[1] != 1 || v->sB[1] != 1 || v->sC[1] != 1) return 0; /* :818-821 (device views) */
    v->tA = tA; v->tB = tB; v->tC = tC;
    return 1;
}

typedef struct { int64_t dA[2], sA[2], dB, sB, dC, sC; int tA; } mv_views;
/* A = matrix, B = vector */
static int do_matrix_vector_check(const analysis *z, const orc_tensor *C, const orc_tensor *A, const orc_tensor *B, mv_views *v) {
    int tA = z->link_in_A.pos[0] == 0;
    int64_t t;
    if (!t_totally_vectorable(B) || !t_totally_vectorable(C)) return 0; /* :471-473 */
    v->dA[0] = product_dims(&z->A_target_in_C, C); v->dA[1] = product_dims(&z->link_in_A, A);
    v->sA[0] = last_stride(&z->target_in_A, A);    v->sA[1] = last_stride(&z->link_in_A, A);
    if (!view_ok(A, v->sA[0], v->sA[1])) return 0;
    if (tA) { t = v->dA[0]; v->dA[0] = v->dA[1]; v->dA[1] = t; t = v->sA[0]; v->sA[0] = v->sA[1]; v->sA[1] = t; }
    v->dB = product_dims(&z->link_in_B, B); v->sB = last_stride(&z->link_in_B, B);
    v->dC = product_dims(&z->A_target_in_C, C); v->sC = last_stride(&z->A_target_in_C, C);
    v->tA = tA;
    return 1;
}

static int do_outer_check(const analysis *z, const orc_tensor *C, const orc_tensor *A, const orc_tensor *B) {
    if (!t_totally_vectorable(A) || !t_totally_vectorable(B)) return 0; /* :287-290 */
    int64_t s0 = last_stride(&z->A_target_in_C, C), s1 = last_stride(&z->B_target_in_C, C);
    return view_ok(C, s0, s1);
}

static int same_dtype3(const orc_tensor *C, const orc_tensor *A, const orc_tensor *B) { return C->dtype == A->dtype && C->dtype == B->dtype; }

int orc_classify(const orc_tensor *C, const orc_tensor *A, const orc_tensor *B, int conj_a, int conj_b) {
    pack pc = mk(C->rank, C->index), pa = mk(A->rank, A->index), pb = mk(B->rank, B->index);
    int s = classify_static_ex(&pc, &pa, &pb, conj_a, conj_b, same_dtype3(C, A, B));
    analysis z;
    switch (s) {
    case ORC_GER: analyse(&z, &pc, &pa, &pb); return do_outer_check(&z, C, A, B) ? ORC_GER : ORC_GENERIC;
    case ORC_GEMV: { mv_views v; analyse(&z, &pc, &pa, &pb); return do_matrix_vector_check(&z, C, A, B, &v) ? ORC_GEMV : ORC_GENERIC; }
    case 10 + ORC_GEMV: { mv_views v; analyse(&z, &pc, &pb, &pa); return do_matrix_vector_check(&z, C, B, A, &v) ? ORC_GEMV : ORC_GENERIC; }
    case ORC_GEMM: { mm_views v; analyse(&z, &pc, &pa, &pb); return do_matrix_product_check(&z, C, A, B, &v) ? ORC_GEMM : ORC_GENERIC; }
    default: return s;
    }
}

/* ------------------------------------------------------------------ generic algorithm */
typedef struct { int n; int label[ORC_MAX_UNIQUE]; int64_t dim[ORC_MAX_UNIQUE], sC[ORC_MAX_UNIQUE], sA[ORC_MAX_UNIQUE], sB[ORC_MAX_UNIQUE]; } loopset;

/* get_stride_for, Utilities.hpp:560-592: sum of the strides of every axis carrying the label */
static int64_t stride_for(const orc_tensor *t, int label) { int64_t s = 0; for (int i = 0; i < t->rank; i++) if (t->index[i] == label) s += t->stride[i]; return s; }
static int64_t dim_for(const orc_tensor *t, int label) { for (int i = 0; i < t->rank; i++) if (t->index[i] == label) return t->dim[i]; return -1; }

static int build_loops(const orc_tensor *C, const orc_tensor *A, const orc_tensor *B, int full_space, loopset *tg, loopset *lk) {
    pack pc = mk(C->rank, C->index), pa = mk(A->rank, A->index), pb = mk(B->rank, B->index);
    pack cu = unique(&pc);
    pack lab = intersect(&pa, &pb), links = difference(&lab, &pc);
    pack lu = unique(&links);
    if (full_space) { /* every index not in C, in order of first appearance in A then B */
        pack ab; ab.n = 0;
        for (int i = 0; i < pa.n; i++) ab.v[ab.n++] = pa.v[i];
        for (int i = 0; i < pb.n; i++) ab.v[ab.n++] = pb.v[i];
        pack nc = difference(&ab, &pc);
        lu = unique(&nc);
    }
    tg->n = cu.n;
    for (int i = 0; i < cu.n; i++) {
        int l = cu.v[i];
        tg->label[i] = l;
        tg->dim[i] = dim_for(C, l); /* Dispatch.hpp:173: target dims come from C */
        tg->sC[i] = stride_for(C, l); tg->sA[i] = stride_for(A, l); tg->sB[i] = stride_for(B, l);
    }
    lk->n = lu.n;
    for (int i = 0; i < lu.n; i++) {
        int l = lu.v[i];
        lk->label[i] = l;
        int64_t d = dim_for(A, l); /* Dispatch.hpp:174: link dims come from A */
        if (d < 0) d = dim_for(B, l);
        lk->dim[i] = d;
        lk->sC[i] = 0; lk->sA[i] = stride_for(A, l); lk->sB[i] = stride_for(B, l);
    }
    return 0;
}

#define GEN_IMPL(NAME, T, CONJ, ZERO, FROMPAIR)                                                                               \
    static T NAME##_link(const loopset *lk, int level, const T *A, const T *B, int64_t ia, int64_t ib, int ca, int cb) {       \
        if (level == lk->n) {                                                                                                   \
            T a = A[ia], b = B[ib];                                                                                             \
            if (ca) a = CONJ(a);                                                                                                \
            if (cb) b = CONJ(b);                                                                                                \
            return a * b;                                                                                                       \
        }                                                                                                                       \
        T sum = ZERO;                                                                                                           \
        for (int64_t i = 0; i < lk->dim[level]; i++)                                                                            \
            sum += NAME##_link(lk, level + 1, A, B, ia + i * lk->sA[level], ib + i * lk->sB[level], ca, cb);                    \
        return sum;                                                                                                             \
    }                                                                                                                           \
    static void NAME##_target(const loopset *tg, const loopset *lk, int level, T *C, const T *A, const T *B, int64_t ic,        \
                              int64_t ia, int64_t ib, T alpha, int ca, int cb) {                                                \
        if (level == tg->n) {                                                                                                   \
            C[ic] += alpha * NAME##_link(lk, 0, A, B, ia, ib, ca, cb);                                                          \
            return;                                                                                                             \
        }                                                                                                                       \
        for (int64_t i = 0; i < tg->dim[level]; i++)                                                                            \
            NAME##_target(tg, lk, level + 1, C, A, B, ic + i * tg->sC[level], ia + i * tg->sA[level], ib + i * tg->sB[level],   \
                          alpha, ca, cb);                                                                                       \
    }                                                                                                                           \
    static void NAME##_scale(const orc_tensor *Ct, int axis, T *C, int64_t off, T beta, int zero) {                             \
        if (axis == Ct->rank) { C[off] = zero ? ZERO : C[off] * beta; return; }                                                 \
        for (int64_t i = 0; i < Ct->dim[axis]; i++) NAME##_scale(Ct, axis + 1, C, off + i * Ct->stride[axis], beta, zero);      \
    }                                                                                                                           \
    static void NAME##_run(const double beta[2], orc_tensor *Ct, const double alpha[2], const orc_tensor *At,                   \
                           const orc_tensor *Bt, int ca, int cb, int full) {                                                    \
        loopset tg, lk;                                                                                                         \
        build_loops(Ct, At, Bt, full, &tg, &lk);                                                                                \
        T b = FROMPAIR(beta), a = FROMPAIR(alpha);                                                                              \
        int zero = (beta[0] == 0.0 && beta[1] == 0.0);                                                                          \
        /* BaseAlgebra.hpp:117-137: C->zero() when beta == 0 (never reads C) else *C *= beta, on the whole tensor */            \
        NAME##_scale(Ct, 0, (T *)Ct->data, 0, b, zero);                                                                         \
        NAME##_target(&tg, &lk, 0, (T *)Ct->data, (const T *)At->data, (const T *)Bt->data, 0, 0, 0, a, ca, cb);                \
    }

#define ID(x) (x)
#define PAIR_REAL_D(p) ((double)(p)[0])
#define PAIR_REAL_F(p) ((float)(p)[0])
#define PAIR_CD(p) ((p)[0] + (p)[1] * I)
#define PAIR_CF(p) ((float complex)((float)(p)[0] + (float)(p)[1] * I))
GEN_IMPL(gen_d, double, ID, 0.0, PAIR_REAL_D)
GEN_IMPL(gen_s, float, ID, 0.0f, PAIR_REAL_F)
GEN_IMPL(gen_z, double complex, conj, 0.0, PAIR_CD)
GEN_IMPL(gen_c, float complex, conjf, 0.0f, PAIR_CF)

int orc_einsum_generic(const double beta[2], orc_tensor *C, const double alpha[2], const orc_tensor *A, const orc_tensor *B, int conj_a,
                       int conj_b, int full_space) {
    if (C->dtype != A->dtype || C->dtype != B->dtype) return -1;
    switch (C->dtype) {
    case ORC_F64: gen_d_run(beta, C, alpha, A, B, conj_a, conj_b, full_space); return 0;
    case ORC_F32: gen_s_run(beta, C, alpha, A, B, conj_a, conj_b, full_space); return 0;
    case ORC_C128: gen_z_run(beta, C, alpha, A, B, conj_a, conj_b, full_space); return 0;
    case ORC_C64: gen_c_run(beta, C, alpha, A, B, conj_a, conj_b, full_space); return 0;
    }
    return -1;
}

/* ------------------------------------------------------------------ permute (Permute.hpp:375-406) */
#define PERM_IMPL(NAME, T, CONJ, FROMPAIR)                                                                                      \
    static void NAME(const double beta[2], orc_tensor *C, const double alpha[2], const orc_tensor *A, const int *perm, int cj) { \
        T       b = FROMPAIR(beta), a = FROMPAIR(alpha);                                                                         \
        int     zero = (beta[0] == 0.0 && beta[1] == 0.0);                                                                       \
        int     r = A->rank;                                                                                                     \
        int64_t idx[ORC_MAX_RANK] = {0};                                                                                         \
        int64_t total = 1;                                                                                                       \
        for (int i = 0; i < r; i++) total *= A->dim[i];                                                                          \
        int inv[ORC_MAX_RANK];                                                                                                   \
        for (int i = 0; i < r; i++) inv[perm[i]] = i; /* A axis -> C axis */                                                     \
        for (int64_t x = 0; x < total; x++) {                                                                                    \
            int64_t oa = 0, oc = 0;                                                                                              \
            for (int i = 0; i < r; i++) { oa += idx[i] * A->stride[i]; oc += idx[i] * C->stride[inv[i]]; }                       \
            T v = ((const T *)A->data)[oa];                                                                                      \
            if (cj) v = CONJ(v);                                                                                                 \
            T *dst = (T *)C->data + oc;                                                                                          \
            *dst = zero ? a * v : b * (*dst) + a * v;                                                                            \
            for (int i = r - 1; i >= 0; i--) { if (++idx[i] < A->dim[i]) break; idx[i] = 0; }                                    \
        }                                                                                                                        \
    }
PERM_IMPL(perm_d, double, ID, PAIR_REAL_D)
PERM_IMPL(perm_s, float, ID, PAIR_REAL_F)
PERM_IMPL(perm_z, double complex, conj, PAIR_CD)
PERM_IMPL(perm_c, float complex, conjf, PAIR_CF)

int orc_permute(const double beta[2], orc_tensor *C, const double alpha[2], const orc_tensor *A, const int *perm, int conj_a) {
    if (C->dtype != A->dtype || C->rank != A->rank) return -1;
    for (int i = 0; i < A->rank; i++) if (A->dim[i] == 0) return 0;
    switch (A->dtype) {
    case ORC_F64: perm_d(beta, C, alpha, A, perm, conj_a); return 0;
    case ORC_F32: perm_s(beta, C, alpha, A, perm, conj_a); return 0;
    case ORC_C128: perm_z(beta, C, alpha, A, perm, conj_a); return 0;
    case ORC_C64: perm_c(beta, C, alpha, A, perm, conj_a); return 0;
    }
    return -1;
}

/* ------------------------------------------------------------------ reference-BLAS loops (column-major) */
#define GEMM_IMPL(NAME, T, CONJ)                                                                                                 \
    static void NAME(char ta, char tb, int64_t m, int64_t n, int64_t k, T alpha, const T *A, int64_t lda, const T *B, int64_t ldb, \
                     T beta, T *C, int64_t ldc) {                                                                                 \
        ta = (char)tolower(ta); tb = (char)tolower(tb);                                                                           \
        if (m == 0 || n == 0 || k == 0) return; /* the reference's vendor layer returns early, HBV/src/gemm.hip:101 */            \
        for (int64_t j = 0; j < n; j++)                                                                                           \
            for (int64_t i = 0; i < m; i++) {                                                                                     \
                T s = 0;                                                                                                          \
                for (int64_t l = 0; l < k; l++) {                                                                                 \
                    T a = (ta == 'n') ? A[i + l * lda] : A[l + i * lda];                                                          \
                    T b = (tb == 'n') ? B[l + j * ldb] : B[j + l * ldb];                                                          \
                    if (ta == 'c') a = CONJ(a);                                                                                   \
                    if (tb == 'c') b = CONJ(b);                                                                                   \
                    s += a * b;                                                                                                   \
                }                                                                                                                 \
                C[i + j * ldc] = (beta == (T)0) ? alpha * s : alpha * s + beta * C[i + j * ldc];                                  \
            }                                                                                                                     \
    }
GEMM_IMPL(gemm_d, double, ID)
GEMM_IMPL(gemm_s, float, ID)
GEMM_IMPL(gemm_z, double complex, conj)

void orc_dgemm(char ta, char tb, int64_t m, int64_t n, int64_t k, double alpha, const double *A, int64_t lda, const double *B, int64_t ldb,
               double beta, double *C, int64_t ldc) { gemm_d(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc); }
void orc_sgemm(char ta, char tb, int64_t m, int64_t n, int64_t k, float alpha, const float *A, int64_t lda, const float *B, int64_t ldb,
               float beta, float *C, int64_t ldc) { gemm_s(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc); }
void orc_zgemm(char ta, char tb, int64_t m, int64_t n, int64_t k, const double alpha[2], const double *A, int64_t lda, const double *B,
               int64_t ldb, const double beta[2], double *C, int64_t ldc) {
    gemm_z(ta, tb, m, n, k, alpha[0] + alpha[1] * I, (const double complex *)A, lda, (const double complex *)B, ldb, beta[0] + beta[1] * I,
           (double complex *)C, ldc);
}
void orc_dgemv(char t, int64_t m, int64_t n, double alpha, const double *A, int64_t lda, const double *x, int64_t incx, double beta,
               double *y, int64_t incy) {
    t = (char)tolower(t);
    if (m == 0 || n == 0) return;
    int64_t leny = t == 'n' ? m : n, lenx = t == 'n' ? n : m;
    const double *xb = incx < 0 ? x - (lenx - 1) * incx : x;
    double       *yb = incy < 0 ? y - (leny - 1) * incy : y;
    for (int64_t i = 0; i < leny; i++) {
        double s = 0;
        for (int64_t l = 0; l < lenx; l++) s += (t == 'n' ? A[i + l * lda] : A[l + i * lda]) * xb[l * incx];
        yb[i * incy] = beta == 0.0 ? alpha * s : alpha * s + beta * yb[i * incy];
    }
}
double orc_ddot(int64_t n, const double *x, int64_t incx, const double *y, int64_t incy) {
    double s = 0;
    const double *xb = incx < 0 ? x - (n - 1) * incx : x, *yb = incy < 0 ? y - (n - 1) * incy : y;
    for (int64_t i = 0; i < n; i++) s += xb[i * incx] * yb[i * incy];
    return s;
}
void orc_dger(int64_t m, int64_t n, double alpha, const double *x, int64_t incx, const double *y, int64_t incy, double *A, int64_t lda) {
    const double *xb = incx < 0 ? x - (m - 1) * incx : x, *yb = incy < 0 ? y - (n - 1) * incy : y;
    for (int64_t j = 0; j < n; j++)
        for (int64_t i = 0; i < m; i++) A[i + j * lda] += alpha * xb[i * incx] * yb[j * incy];
}
/* HBV/src/direct_prod.hip:20-200: z += alpha*x*y */
void orc_ddirprod(int64_t n, double alpha, const double *x, int64_t incx, const double *y, int64_t incy, double *z, int64_t incz) {
    for (int64_t i = 0; i < n; i++) z[i * incz] += alpha * x[i * incx] * y[i * incy];
}

/* ------------------------------------------------------------------ optional OpenBLAS (timing baseline) */
typedef void (*dgemm32_t)(const char *, const char *, const int *, const int *, const int *, const double *, const double *, const int *,
                          const double *, const int *, const double *, double *, const int *);
typedef void (*dgemm64_t)(const char *, const char *, const int64_t *, const int64_t *, const int64_t *, const double *, const double *,
                          const int64_t *, const double *, const int64_t *, const double *, double *, const int64_t *);
typedef void (*dgemv32_t)(const char *, const int *, const int *, const double *, const double *, const int *, const double *, const int *,
                          const double *, double *, const int *);
typedef void (*dgemv64_t)(const char *, const int64_t *, const int64_t *, const double *, const double *, const int64_t *, const double *,
                          const int64_t *, const double *, double *, const int64_t *);
static void *g_blas = NULL;
static dgemm32_t g_dgemm32; static dgemm64_t g_dgemm64; static dgemv32_t g_dgemv32; static dgemv64_t g_dgemv64;
static void (*g_set_threads)(int);
static char g_blas_info[512] = "none";

int orc_blas_load(const char *path) {
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) { snprintf(g_blas_info, sizeof g_blas_info, "dlopen failed: %s", dlerror()); return -1; }
    g_dgemm32 = NULL; g_dgemm64 = NULL; g_dgemv32 = NULL; g_dgemv64 = NULL;
    const char *sym = NULL;
    if ((g_dgemm64 = (dgemm64_t)dlsym(h, "scipy_dgemm_64_"))) { g_dgemv64 = (dgemv64_t)dlsym(h, "scipy_dgemv_64_"); sym = "scipy_dgemm_64_ (ILP64)"; }
    else if ((g_dgemm32 = (dgemm32_t)dlsym(h, "scipy_dgemm_"))) { g_dgemv32 = (dgemv32_t)dlsym(h, "scipy_dgemv_"); sym = "scipy_dgemm_ (LP64)"; }
    else if ((g_dgemm32 = (dgemm32_t)dlsym(h, "dgemm_"))) { g_dgemv32 = (dgemv32_t)dlsym(h, "dgemv_"); sym = "dgemm_ (LP64)"; }
    if (!sym) { snprintf(g_blas_info, sizeof g_blas_info, "no dgemm symbol in %s", path); dlclose(h); return -2; }
    g_set_threads = (void (*)(int))dlsym(h, "scipy_openblas_set_num_threads64_");
    if (!g_set_threads) g_set_threads = (void (*)(int))dlsym(h, "scipy_openblas_set_num_threads");
    if (!g_set_threads) g_set_threads = (void (*)(int))dlsym(h, "openblas_set_num_threads");
    char *(*cfg)(void) = (char *(*)(void))dlsym(h, "scipy_openblas_get_config64_");
    if (!cfg) cfg = (char *(*)(void))dlsym(h, "scipy_openblas_get_config");
    if (!cfg) cfg = (char *(*)(void))dlsym(h, "openblas_get_config");
    snprintf(g_blas_info, sizeof g_blas_info, "%s via %s [%s]", path, sym, cfg ? cfg() : "?");
    g_blas = h;
    return 0;
}
int orc_blas_loaded(void) { return g_blas != NULL; }
const char *orc_blas_info(void) { return g_blas_info; }
void orc_blas_set_threads(int n) { if (g_set_threads) g_set_threads(n); }

int orc_blas_dgemm(char ta, char tb, int64_t m, int64_t n, int64_t k, double alpha, const double *A, int64_t lda, const double *B,
                   int64_t ldb, double beta, double *C, int64_t ldc) {
    if (g_dgemm64) { g_dgemm64(&ta, &tb, &m, &n, &k, &alpha, A, &lda, B, &ldb, &beta, C, &ldc); return 0; }
    if (g_dgemm32) {
        int m_ = (int)m, n_ = (int)n, k_ = (int)k, lda_ = (int)lda, ldb_ = (int)ldb, ldc_ = (int)ldc;
        g_dgemm32(&ta, &tb, &m_, &n_, &k_, &alpha, A, &lda_, B, &ldb_, &beta, C, &ldc_);
        return 0;
    }
    orc_dgemm(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
    return 1;
}
int orc_blas_dgemv(char t, int64_t m, int64_t n, double alpha, const double *A, int64_t lda, const double *x, int64_t incx, double beta,
                   double *y, int64_t incy) {
    if (g_dgemv64) { g_dgemv64(&t, &m, &n, &alpha, A, &lda, x, &incx, &beta, y, &incy); return 0; }
    if (g_dgemv32) {
        int m_ = (int)m, n_ = (int)n, lda_ = (int)lda, ix = (int)incx, iy = (int)incy;
        g_dgemv32(&t, &m_, &n_, &alpha, A, &lda_, x, &ix, &beta, y, &iy);
        return 0;
    }
    orc_dgemv(t, m, n, alpha, A, lda, x, incx, beta, y, incy);
    return 1;
}

/* ------------------------------------------------------------------ CPU baseline of the whole path (FP64) */
/* OpenMP version of the generic loops: parallel over the outermost target loop, as
 * BaseAlgebra.hpp:75 (EINSUMS_OMP_PARALLEL_FOR on every level; inner levels are nested => serial). */
static void gen_d_target_omp(const loopset *tg, const loopset *lk, double *C, const double *A, const double *B, double alpha) {
    if (tg->n == 0) { C[0] += alpha * gen_d_link(lk, 0, A, B, 0, 0, 0, 0); return; }
    /* collapse the two outermost target loops when present to expose enough parallelism */
    int64_t d0 = tg->dim[0], d1 = tg->n > 1 ? tg->dim[1] : 1;
#pragma omp parallel for schedule(static)
    for (int64_t x = 0; x < d0 * d1; x++) {
        int64_t i = x / d1, j = x - i * d1;
        if (tg->n > 1)
            gen_d_target(tg, lk, 2, C, A, B, i * tg->sC[0] + j * tg->sC[1], i * tg->sA[0] + j * tg->sA[1], i * tg->sB[0] + j * tg->sB[1],
                         alpha, 0, 0);
        else
            gen_d_target(tg, lk, 1, C, A, B, i * tg->sC[0], i * tg->sA[0], i * tg->sB[0], alpha, 0, 0);
    }
}

int orc_einsum_f64(double beta, orc_tensor *C, double alpha, const orc_tensor *A, const orc_tensor *B) {
    if (C->dtype != ORC_F64 || A->dtype != ORC_F64 || B->dtype != ORC_F64) return -1;
    pack pc = mk(C->rank, C->index), pa = mk(A->rank, A->index), pb = mk(B->rank, B->index);
    int s = classify_static_ex(&pc, &pa, &pb, 0, 0, 1);
    analysis z;
    if (s == ORC_GEMM) {
        mm_views v;
        analyse(&z, &pc, &pa, &pb);
        if (do_matrix_product_check(&z, C, A, B, &v)) {
            /* row-major views: C(r x c) = op(A) op(B); column-major BLAS sees the transposes, so the
             * operands are exchanged (LinearAlgebra/Bases/gemm.hpp:109-242, GPULinearAlgebra.hpp:126-139). */
            const double *pA = (const double *)A->data, *pB = (const double *)B->data;
            double *pC = (double *)C->data;
            if (!v.tC) {
                int64_t m = v.dC[0], n = v.dC[1], k = v.tA ? v.dA[0] : v.dA[1];
                orc_blas_dgemm(v.tB ? 't' : 'n', v.tA ? 't' : 'n', n, m, k, alpha, pB, v.sB[0], pA, v.sA[0], beta, pC, v.sC[0]);
            } else {
                /* C^T = op(B)^T op(A)^T : gemm(transB', transA', tB, tA) on the swapped operands (Dispatch.hpp:831-833) */
                int64_t m = v.dC[0], n = v.dC[1], k = v.tA ? v.dA[0] : v.dA[1];
                orc_blas_dgemm(!v.tA ? 't' : 'n', !v.tB ? 't' : 'n', n, m, k, alpha, pA, v.sA[0], pB, v.sB[0], beta, pC, v.sC[0]);
            }
            return ORC_GEMM;
        }
        s = ORC_GENERIC;
    }
    if (s == ORC_GEMV || s == 10 + ORC_GEMV) {
        mv_views v;
        const orc_tensor *M = (s == ORC_GEMV) ? A : B, *V = (s == ORC_GEMV) ? B : A;
        if (s == ORC_GEMV) analyse(&z, &pc, &pa, &pb); else analyse(&z, &pc, &pb, &pa);
        if (do_matrix_vector_check(&z, C, M, V, &v)) {
            /* row-major matrix (rows x cols, ld = sA[0]) seen by column-major BLAS as its transpose */
            int64_t rows = v.dA[0], cols = v.dA[1];
            orc_blas_dgemv(v.tA ? 'n' : 't', cols, rows, alpha, (const double *)M->data, v.sA[0], (const double *)V->data, v.sB, beta,
                           (double *)C->data, v.sC);
            return ORC_GEMV;
        }
        s = ORC_GENERIC;
    }
    /* everything else: generic loops (DOT/DIRECT/GER are not on any benchmark config) */
    loopset tg, lk;
    build_loops(C, A, B, 1, &tg, &lk);
    gen_d_scale(C, 0, (double *)C->data, 0, beta, beta == 0.0);
    gen_d_target_omp(&tg, &lk, (double *)C->data, (const double *)A->data, (const double *)B->data, alpha);
    return s >= 10 ? ORC_GENERIC : (s == ORC_GEMM || s == ORC_GEMV ? ORC_GENERIC : s);
}

/* cache-blocked transpose over the (fastest-A, fastest-C) plane; other axes looped outside */
int orc_permute_fast_f64(double beta, orc_tensor *C, double alpha, const orc_tensor *A, const int *perm) {
    int r = A->rank;
    if (r == 0) { double *c = (double *)C->data; *c = (beta == 0 ? 0 : beta * *c) + alpha * *(const double *)A->data; return 0; }
    int inv[ORC_MAX_RANK];
    for (int i = 0; i < r; i++) inv[perm[i]] = i;
    int fa = 0, fc = 0; /* A axes with the smallest A stride / smallest C stride */
    for (int i = 1; i < r; i++) {
        if (A->stride[i] < A->stride[fa]) fa = i;
        if (C->stride[inv[i]] < C->stride[inv[fc]]) fc = i;
    }
    int64_t outer = 1; int oax[ORC_MAX_RANK], no = 0;
    for (int i = 0; i < r; i++) if (i != fa && i != fc) { oax[no++] = i; outer *= A->dim[i]; }
    const double *pa = (const double *)A->data; double *pcd = (double *)C->data;
    const int64_t BLK = 32;
#pragma omp parallel for schedule(static)
    for (int64_t o = 0; o < outer; o++) {
        int64_t x = o, offa = 0, offc = 0;
        for (int q = no - 1; q >= 0; q--) { int ax = oax[q]; int64_t i = x % A->dim[ax]; x /= A->dim[ax]; offa += i * A->stride[ax]; offc += i * C->stride[inv[ax]]; }
        if (fa == fc) {
            for (int64_t i = 0; i < A->dim[fa]; i++) {
                double *d = pcd + offc + i * C->stride[inv[fa]];
                *d = (beta == 0 ? 0 : beta * *d) + alpha * pa[offa + i * A->stride[fa]];
            }
        } else {
            for (int64_t c0 = 0; c0 < A->dim[fc]; c0 += BLK)
                for (int64_t a0 = 0; a0 < A->dim[fa]; a0 += BLK)
                    for (int64_t ic = c0; ic < c0 + BLK && ic < A->dim[fc]; ic++)
                        for (int64_t ia = a0; ia < a0 + BLK && ia < A->dim[fa]; ia++) {
                            double *d = pcd + offc + ia * C->stride[inv[fa]] + ic * C->stride[inv[fc]];
                            *d = (beta == 0 ? 0 : beta * *d) + alpha * pa[offa + ia * A->stride[fa] + ic * A->stride[fc]];
                        }
        }
    }
    return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
