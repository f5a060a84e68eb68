/* orc_internal.h -- shared macros/helpers of the oracle sources (test infrastructure only) */
#ifndef ORC_INTERNAL_H
#define ORC_INTERNAL_H
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "adflow_oracle.h"

#define IRHO 0
#define IVX 1
#define IVY 2
#define IVZ 3
#define IRHOE 4
#define ITU1 5
#define IMX IVX
#define IMY IVY
#define IMZ IVZ

/* src/modules/constants.F90:22-24,71-101 */
static const double __attribute__((unused)) zero = 0.0, one = 1.0, two = 2.0, three = 3.0, four = 4.0, five = 5.0;
static const double __attribute__((unused)) half = 0.5, fourth = 0.25, eighth = 0.125;
static const double __attribute__((unused)) third = 1.0 / 3.0, sixth = 1.0 / 6.0;
static const double __attribute__((unused)) eps_ = 1.e-25;
static const double __attribute__((unused)) thresholdReal = 1.e-10;

typedef struct Dims {
    int nx, ny, nz, il, jl, kl, ie, je, ke, ib, jb, kb;
    long NI, NJ, NK, N; /* box extents and size */
    long sI, sJ, sK;    /* strides */
} Dims;

static Dims dims_of(const OrcBlock* b) {
    Dims d;
    d.nx = b->nx; d.ny = b->ny; d.nz = b->nz;
    d.il = d.nx + 1; d.jl = d.ny + 1; d.kl = d.nz + 1;
    d.ie = d.nx + 2; d.je = d.ny + 2; d.ke = d.nz + 2;
    d.ib = d.nx + 3; d.jb = d.ny + 3; d.kb = d.nz + 3;
    d.NI = d.ib + 1; d.NJ = d.jb + 1; d.NK = d.kb + 1;
    d.N = d.NI * d.NJ * d.NK;
    d.sI = 1; d.sJ = d.NI; d.sK = d.NI * d.NJ;
    return d;
}
#define IDX(i, j, k) ((long)(i) + d.NI * ((long)(j) + d.NJ * (long)(k)))
#define W(c, l) b->w[(long)(l) * d.N + (c)]
#define DW(c, l) b->dw[(long)(l) * d.N + (c)]
#define FW(c, l) b->fw[(long)(l) * d.N + (c)]
#define X(c, m) b->x[(long)(m) * d.N + (c)]
#define GR(c, m) b->grad[(long)(m) * d.N + (c)]

static inline double dmax(double a, double b) { return a > b ? a : b; }
static inline double dmin(double a, double b) { return a < b ? a : b; }
/* Fortran DIM(x,y) = max(x-y,0) */
static inline double fdim_(double a, double b) { return a - b > 0.0 ? a - b : 0.0; }


#endif
