/*
 * TEST INFRASTRUCTURE -- not product code.
 *
 * Batched driver around the reference's scalar `matern52()` (declared in
 * /root/reference/gptools/kernel/include/matern.h:24-26, defined in
 * /root/reference/gptools/kernel/src/matern.c:165-186).  It plays the role of
 * the reference's Cython loop (/root/reference/gptools/kernel/_matern.pyx:28-31)
 * so that the unmodified C source can be exercised through ctypes.  It is
 * compiled TOGETHER WITH the reference's own matern.c (which stays where it
 * lies under /root/reference; see oracle/Makefile) into oracle/_ref/.
 */
#include <stdint.h>

double matern52(const double *xi, const double *xj, const int32_t *ni,
                const int32_t *nj, int32_t d, const double *var);

void matern52_pairs(const double *Xi, const double *Xj, const int32_t *ni,
                    const int32_t *nj, int64_t npairs, int32_t d,
                    const double *var, double *out)
{
    int64_t p;
    for (p = 0; p < npairs; p++)
        out[p] = matern52(Xi + p * d, Xj + p * d, ni + p * d, nj + p * d, d, var);
}
