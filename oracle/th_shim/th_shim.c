/* TEST INFRASTRUCTURE ONLY -- bodies for oracle/th_shim/TH/TH.h. */
#include <stdio.h>
#include "TH/TH.h"

static long shim_numel(const THShimTensor *t) {
    long n = 1;
    for (int i = 0; i < 4; ++i) n *= (t->size[i] > 0 ? t->size[i] : 1);
    return n;
}
float *THFloatTensor_data(THFloatTensor *t) { return (float *)t->data; }
int *THIntTensor_data(THIntTensor *t) { return (int *)t->data; }
long *THLongTensor_data(THLongTensor *t) { return (long *)t->data; }
unsigned char *THByteTensor_data(THByteTensor *t) { return (unsigned char *)t->data; }
long THFloatTensor_size(const THFloatTensor *t, int dim) { return t->size[dim]; }
void THFloatTensor_resize4d(THFloatTensor *t, long s0, long s1, long s2, long s3) {
    if (s0 * s1 * s2 * s3 > t->capacity) {
        fprintf(stderr, "th_shim: resize4d beyond caller-provided capacity\n");
        abort();
    }
    t->size[0] = s0; t->size[1] = s1; t->size[2] = s2; t->size[3] = s3;
}
void THFloatTensor_zero(THFloatTensor *t) { memset(t->data, 0, (size_t)shim_numel(t) * sizeof(float)); }
int THLongTensor_isContiguous(const void *t) { (void)t; return 1; }
THByteTensor *THByteTensor_newWithSize1d(long n) {
    THByteTensor *t = (THByteTensor *)calloc(1, sizeof(THByteTensor));
    t->size[0] = n; t->itemsize = 1; t->capacity = n;
    t->data = malloc((size_t)(n > 0 ? n : 1));
    return t;
}
void THByteTensor_fill(THByteTensor *t, unsigned char v) { memset(t->data, v, (size_t)t->size[0]); }
void THByteTensor_free(THByteTensor *t) { free(t->data); free(t); }
void THShim_argcheck(int cond, int argn, const char *msg) {
    if (!cond) { fprintf(stderr, "th_shim: THArgCheck failed (arg %d): %s\n", argn, msg); abort(); }
}
