/*
 * mpp_layout.h -- the two FMS layout rules this path depends on, ONE C source shared by the product library
 * (mom5_b200/csrc/capi.cu), the CPU oracle (oracle/mom5adv_oracle.c) and any C/C++ host that wants to predict a layout.
 * They are bookkeeping, not arithmetic under test: a rank's extents must simply be the same numbers FMS would hand it.
 *   mpp_define_layout2D   src/shared/mpp/include/mpp_domains_define.inc:28-55
 *   mpp_compute_extent    src/shared/mpp/include/mpp_domains_define.inc:187-273 (no user extent): mirror-symmetric uneven split
 * (mom5_b200/domain.py carries a Python twin for hosts without the library; tests/test_domain_plan.py checks the two agree.)
 */
#ifndef MPP_LAYOUT_H
#define MPP_LAYOUT_H

#include <math.h>

/* layout2[0] = idiv (x), layout2[1] = jdiv (y):  idiv = nint(sqrt(float(ndivs*isz)/jsz)), decreased until it divides ndivs */
static inline void mpp_define_layout2d_c(int ni_g, int nj_g, int ndivs, int *layout2)
{
    /* float() is default real; -r8 builds make it binary64 */
    const double q = (double)((long)ndivs * (long)ni_g) / (double)nj_g;
    int idiv = (int)lround(sqrt(q));
    if (idiv < 1) idiv = 1;
    while (ndivs % idiv != 0) idiv--;
    layout2[0] = idiv;
    layout2[1] = ndivs / idiv;
}

/* ibegin[d], iend[d] (inclusive, d = 0..ndivs-1) of the compute domains along one axis; returns 0 on success,
 * 1 = an extent is not positive definite, 2 = the extents do not span the axis (the reference's two FATAL checks) */
static inline int mpp_compute_extent_c(int isg, int ieg, int ndivs, int *ibegin, int *iend)
{
    const int npts = ieg - isg + 1;
    const int even_n = (ndivs % 2 == 0), even_p = (npts % 2 == 0);
    const int symmetrize = (even_n && even_p) || (!even_n && !even_p) || (!even_n && even_p && ndivs < npts / 2);
    int is = isg, ie = 0, imaxv = ieg, ndmax = ndivs;
    for (int ndiv = 0; ndiv < ndivs; ndiv++) {
        if (ndiv < (ndivs - 1) / 2 + 1) {
            ie = is + (int)ceil((double)(imaxv - is + 1) / (double)(ndmax - ndiv)) - 1;
            const int ndmirror = (ndivs - 1) - ndiv;
            if (ndmirror > ndiv && symmetrize) {
                const int mb = isg + ieg - ie, me = isg + ieg - is;
                ibegin[ndmirror] = mb > ie + 1 ? mb : ie + 1;
                iend[ndmirror] = me > ie + 1 ? me : ie + 1;
                imaxv = ibegin[ndmirror] - 1;
                ndmax = ndmax - 1;
            }
        } else if (symmetrize) {
            is = ibegin[ndiv];
            ie = iend[ndiv];
        } else {
            ie = is + (int)ceil((double)(imaxv - is + 1) / (double)(ndmax - ndiv)) - 1;
        }
        ibegin[ndiv] = is;
        iend[ndiv] = ie;
        if (ie < is) return 1;
        if (ndiv == ndivs - 1 && iend[ndiv] != ieg) return 2;
        is = ie + 1;
    }
    return 0;
}

#endif
