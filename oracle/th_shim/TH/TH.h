/* TEST INFRASTRUCTURE ONLY -- minimal stand-in for the long-gone <TH/TH.h>.
 *
 * The reference's CPU sources (lib/roi_align/src/crop_and_resize.c, lib/nms/src/nms.c)
 * include <TH/TH.h>, which does not exist in torch >= 1.0.  This header supplies exactly
 * the handful of names those two files reference so they can be
 * compiled *unmodified*, where they lie under /root/reference, into oracle/_ref/.
 * One struct serves every tensor type; data is caller-owned (ctypes/numpy buffers).
 */
#ifndef FI_ORACLE_TH_SHIM_H
#define FI_ORACLE_TH_SHIM_H
#include <stdlib.h>
#include <string.h>

typedef struct THShimTensor {
    long size[4];     /* crop_and_resize.c reads ->size[i] directly            */
    void *data;       /* caller-owned storage                                  */
    long itemsize;    /* bytes per element                                     */
    long capacity;    /* elements available behind data (resize may not grow)  */
} THShimTensor;

typedef THShimTensor THFloatTensor;
typedef THShimTensor THIntTensor;
typedef THShimTensor THLongTensor;
typedef THShimTensor THByteTensor;

float *THFloatTensor_data(THFloatTensor *t);
int *THIntTensor_data(THIntTensor *t);
long *THLongTensor_data(THLongTensor *t);
unsigned char *THByteTensor_data(THByteTensor *t);
long THFloatTensor_size(const THFloatTensor *t, int dim);
void THFloatTensor_resize4d(THFloatTensor *t, long s0, long s1, long s2, long s3);
void THFloatTensor_zero(THFloatTensor *t);
int THLongTensor_isContiguous(const void *t);
THByteTensor *THByteTensor_newWithSize1d(long n);
void THByteTensor_fill(THByteTensor *t, unsigned char v);
void THByteTensor_free(THByteTensor *t);
void THShim_argcheck(int cond, int argn, const char *msg);
#define THArgCheck(cond, argn, msg) THShim_argcheck((cond) != 0, (argn), (msg))

#endif
