/* TEST INFRASTRUCTURE ONLY (oracle/): minimal <cblas.h> stand-in.
 *
 * The reference dispatches its GEMM/GEMV/DOTU to any CBLAS
 * (/root/reference/include/jet/TensorHelpers.hpp:11-15,54-60,84-88,108-110).
 * This image has no system CBLAS, but the scipy wheel bundles an LP64 OpenBLAS
 * whose exports carry a `scipy_` prefix.  This header declares the six entry
 * points the reference calls and forwards them to those exports.
 */
#ifndef JETB200_ORACLE_CBLAS_SHIM_H
#define JETB200_ORACLE_CBLAS_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum CBLAS_TRANSPOSE {
    CblasNoTrans = 111,
    CblasTrans = 112,
    CblasConjTrans = 113
} CBLAS_TRANSPOSE;

void scipy_cblas_cgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int, int, int, const void *,
                       const void *, int, const void *, int, const void *, void *, int);
void scipy_cblas_zgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int, int, int, const void *,
                       const void *, int, const void *, int, const void *, void *, int);
void scipy_cblas_cgemv(CBLAS_ORDER, CBLAS_TRANSPOSE, int, int, const void *, const void *, int,
                       const void *, int, const void *, void *, int);
void scipy_cblas_zgemv(CBLAS_ORDER, CBLAS_TRANSPOSE, int, int, const void *, const void *, int,
                       const void *, int, const void *, void *, int);
void scipy_cblas_cdotu_sub(int, const void *, int, const void *, int, void *);
void scipy_cblas_zdotu_sub(int, const void *, int, const void *, int, void *);
void scipy_openblas_set_num_threads(int);
int scipy_openblas_get_num_threads(void);
char *scipy_openblas_get_config(void);

#define cblas_cgemm scipy_cblas_cgemm
#define cblas_zgemm scipy_cblas_zgemm
#define cblas_cgemv scipy_cblas_cgemv
#define cblas_zgemv scipy_cblas_zgemv
#define cblas_cdotu_sub scipy_cblas_cdotu_sub
#define cblas_zdotu_sub scipy_cblas_zdotu_sub

#ifdef __cplusplus
}
#endif
#endif
