/* rrtmg_lw_kdata_layout.h -- interchange format for the UNREDUCED (16 g-point)
 * k-distribution data that rrtmg_lw_ini() consumes.
 *
 * The reference fills these arrays in lw_kgb01..16 (src/rrtmg_lw_k_g.f90 or
 * src/rrtmg_lw_read_nc.f90 + data/rrtmg_lw.nc; both are absent from the
 * reference mount).  Array names and shapes are those declared in
 * modules/rrlw_kg01.f90 .. rrlw_kg16.f90 (the "...o" originals, no1..no16 = 16).
 *
 * The blob is a flat little-endian float64 array: the arrays below, in table
 * order, each in Fortran element order (first subscript fastest).
 *   gfirst == 0 : shape (lead, 16)   element (i,g) at i + lead*g    (kao, selfrefo, ...)
 *   gfirst == 1 : shape (16, lead)   element (g,j) at g + 16*j      (fracrefao/fracrefbo)
 * `lead` is the product of all non-g extents, e.g. kao(9,5,13,16) -> 585,
 * kbo(5,5,13:59,16) -> 1175, kao_mn2o(9,19,16) -> 171.
 * weighted == 1 arrays are reduced 16 -> ngc(band) with the rwgt weights,
 * weighted == 0 (Planck fractions) by plain summation (src/rrtmg_lw_init.f90:385-2034).
 *
 * This header is a data-format specification only; it holds no algorithm.
 */
#ifndef RRTMG_LW_KDATA_LAYOUT_H
#define RRTMG_LW_KDATA_LAYOUT_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int band;          /* 1..16 */
    const char* name;  /* reference array name in modules/rrlw_kgNN.f90 */
    int lead;          /* product of non-g extents */
    int gfirst;        /* 1: g-point is the FIRST subscript (fracref arrays) */
    int weighted;      /* 1: rwgt-weighted reduction, 0: plain sum */
} rrtmg_lw_karray_desc;

#define RRTMG_LW_NKARRAYS 115

static const rrtmg_lw_karray_desc RRTMG_LW_KARRAYS[RRTMG_LW_NKARRAYS] = {
    /* band 1  (modules/rrlw_kg01.f90) */
    {1, "kao", 65, 0, 1}, {1, "kbo", 235, 0, 1}, {1, "selfrefo", 10, 0, 1}, {1, "forrefo", 4, 0, 1},
    {1, "fracrefao", 1, 1, 0}, {1, "fracrefbo", 1, 1, 0}, {1, "kao_mn2", 19, 0, 1}, {1, "kbo_mn2", 19, 0, 1},
    /* band 2 */
    {2, "kao", 65, 0, 1}, {2, "kbo", 235, 0, 1}, {2, "selfrefo", 10, 0, 1}, {2, "forrefo", 4, 0, 1},
    {2, "fracrefao", 1, 1, 0}, {2, "fracrefbo", 1, 1, 0},
    /* band 3 */
    {3, "kao", 585, 0, 1}, {3, "kbo", 1175, 0, 1}, {3, "selfrefo", 10, 0, 1}, {3, "forrefo", 4, 0, 1},
    {3, "fracrefao", 9, 1, 0}, {3, "fracrefbo", 5, 1, 0}, {3, "kao_mn2o", 171, 0, 1}, {3, "kbo_mn2o", 95, 0, 1},
    /* band 4 */
    {4, "kao", 585, 0, 1}, {4, "kbo", 1175, 0, 1}, {4, "selfrefo", 10, 0, 1}, {4, "forrefo", 4, 0, 1},
    {4, "fracrefao", 9, 1, 0}, {4, "fracrefbo", 5, 1, 0},
    /* band 5 */
    {5, "kao", 585, 0, 1}, {5, "kbo", 1175, 0, 1}, {5, "selfrefo", 10, 0, 1}, {5, "forrefo", 4, 0, 1},
    {5, "fracrefao", 9, 1, 0}, {5, "fracrefbo", 5, 1, 0}, {5, "kao_mo3", 171, 0, 1}, {5, "ccl4o", 1, 0, 1},
    /* band 6 */
    {6, "kao", 65, 0, 1}, {6, "selfrefo", 10, 0, 1}, {6, "forrefo", 4, 0, 1}, {6, "fracrefao", 1, 1, 0},
    {6, "kao_mco2", 19, 0, 1}, {6, "cfc11adjo", 1, 0, 1}, {6, "cfc12o", 1, 0, 1},
    /* band 7 */
    {7, "kao", 585, 0, 1}, {7, "kbo", 235, 0, 1}, {7, "selfrefo", 10, 0, 1}, {7, "forrefo", 4, 0, 1},
    {7, "fracrefao", 9, 1, 0}, {7, "fracrefbo", 1, 1, 0}, {7, "kao_mco2", 171, 0, 1}, {7, "kbo_mco2", 19, 0, 1},
    /* band 8 */
    {8, "kao", 65, 0, 1}, {8, "kbo", 235, 0, 1}, {8, "selfrefo", 10, 0, 1}, {8, "forrefo", 4, 0, 1},
    {8, "fracrefao", 1, 1, 0}, {8, "fracrefbo", 1, 1, 0}, {8, "kao_mco2", 19, 0, 1}, {8, "kao_mn2o", 19, 0, 1},
    {8, "kao_mo3", 19, 0, 1}, {8, "kbo_mco2", 19, 0, 1}, {8, "kbo_mn2o", 19, 0, 1}, {8, "cfc12o", 1, 0, 1},
    {8, "cfc22adjo", 1, 0, 1},
    /* band 9 */
    {9, "kao", 585, 0, 1}, {9, "kbo", 235, 0, 1}, {9, "selfrefo", 10, 0, 1}, {9, "forrefo", 4, 0, 1},
    {9, "fracrefao", 9, 1, 0}, {9, "fracrefbo", 1, 1, 0}, {9, "kao_mn2o", 171, 0, 1}, {9, "kbo_mn2o", 19, 0, 1},
    /* band 10 */
    {10, "kao", 65, 0, 1}, {10, "kbo", 235, 0, 1}, {10, "selfrefo", 10, 0, 1}, {10, "forrefo", 4, 0, 1},
    {10, "fracrefao", 1, 1, 0}, {10, "fracrefbo", 1, 1, 0},
    /* band 11 */
    {11, "kao", 65, 0, 1}, {11, "kbo", 235, 0, 1}, {11, "selfrefo", 10, 0, 1}, {11, "forrefo", 4, 0, 1},
    {11, "fracrefao", 1, 1, 0}, {11, "fracrefbo", 1, 1, 0}, {11, "kao_mo2", 19, 0, 1}, {11, "kbo_mo2", 19, 0, 1},
    /* band 12 */
    {12, "kao", 585, 0, 1}, {12, "selfrefo", 10, 0, 1}, {12, "forrefo", 4, 0, 1}, {12, "fracrefao", 9, 1, 0},
    /* band 13 */
    {13, "kao", 585, 0, 1}, {13, "selfrefo", 10, 0, 1}, {13, "forrefo", 4, 0, 1}, {13, "fracrefao", 9, 1, 0},
    {13, "fracrefbo", 1, 1, 0}, {13, "kao_mco2", 171, 0, 1}, {13, "kao_mco", 171, 0, 1}, {13, "kbo_mo3", 19, 0, 1},
    /* band 14 */
    {14, "kao", 65, 0, 1}, {14, "kbo", 235, 0, 1}, {14, "selfrefo", 10, 0, 1}, {14, "forrefo", 4, 0, 1},
    {14, "fracrefao", 1, 1, 0}, {14, "fracrefbo", 1, 1, 0},
    /* band 15 */
    {15, "kao", 585, 0, 1}, {15, "selfrefo", 10, 0, 1}, {15, "forrefo", 4, 0, 1}, {15, "fracrefao", 9, 1, 0},
    {15, "kao_mn2", 171, 0, 1},
    /* band 16 */
    {16, "kao", 585, 0, 1}, {16, "kbo", 235, 0, 1}, {16, "selfrefo", 10, 0, 1}, {16, "forrefo", 4, 0, 1},
    {16, "fracrefao", 9, 1, 0}, {16, "fracrefbo", 1, 1, 0},
};

/* total number of doubles in the blob = 16 * sum(lead) = 16 * 13241 = 211856 */
#define RRTMG_LW_KDATA_DOUBLES 211856

#ifdef __cplusplus
}
#endif
#endif
