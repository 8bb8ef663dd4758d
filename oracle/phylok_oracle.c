/* C restatement of the reference node-pair DP  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Follows /root/reference/phyloK2.py:158-209 (PhyloKernel.kernel) on the flat post-order arrays that
 * oracle/phylok_oracle.py::annotate builds (phyloK2.py:132-146):
 *     prod[i]        production of internal node i            (1..3, or any class id in labelled mode)
 *     cprod[2i+s]    production of child slot s               (0 = tip)
 *     cidx[2i+s]     post-order index of an internal child    (-1 = tip)
 *     bl[2i+s]       branch length of child slot s
 * Same loop nest (post-order x post-order), same expanded Gaussian argument sq1 + sq2 - 2*dot
 * (phyloK2.py:184-185), same ordered-slot child rule (phyloK2.py:187-199), fp64 throughout.
 * Pinned by tests/test_oracle.py against tests/golden/kernel_kat.json (values from the reference itself).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 *
 * Build: make -C oracle   (gcc -O2 -fPIC -shared; -ffp-contract=off so no FMA reassociates the argument)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* returns K, or NaN on allocation failure; counts (may be NULL) receives {node pairs, matched cells M, child reads R} */
double pk_oracle_kernel(int32_t n1, const uint8_t *prod1, const uint8_t *cprod1, const int32_t *cidx1, const double *bl1,
                        int32_t n2, const uint8_t *prod2, const uint8_t *cprod2, const int32_t *cidx2, const double *bl2,
                        double lambda, double tau, double sigma, int64_t *counts)
{
    double *dp = (double *)malloc(sizeof(double) * (size_t)n1 * (size_t)n2);
    double *sq1 = (double *)malloc(sizeof(double) * (size_t)n1);
    double *sq2 = (double *)malloc(sizeof(double) * (size_t)n2);
    double k = 0.0;
    int64_t m = 0, r = 0;
    if (!dp || !sq1 || !sq2) { free(dp); free(sq1); free(sq2); return NAN; }
    for (int32_t i = 0; i < n1; ++i) sq1[i] = bl1[2 * i] * bl1[2 * i] + bl1[2 * i + 1] * bl1[2 * i + 1];   /* :146 */
    for (int32_t j = 0; j < n2; ++j) sq2[j] = bl2[2 * j] * bl2[2 * j] + bl2[2 * j + 1] * bl2[2 * j + 1];

    for (int32_t i = 0; i < n1; ++i) {                       /* :177  children before parents */
        for (int32_t j = 0; j < n2; ++j) {                   /* :181 */
            if (prod1[i] != prod2[j]) continue;              /* :183 */
            double dot = bl1[2 * i] * bl2[2 * j] + bl1[2 * i + 1] * bl2[2 * j + 1];
            double res = lambda * exp(-1. / tau * (sq1[i] + sq2[j] - 2 * dot));   /* :184-185 */
            ++m;
            for (int s = 0; s < 2; ++s) {                    /* :187 ordered slots */
                uint8_t c1 = cprod1[2 * i + s], c2 = cprod2[2 * j + s];
                if (c1 != c2) continue;                      /* :191-192 */
                if (c1 == 0) {
                    res *= sigma + lambda;                   /* :194-196 tip / tip */
                } else {
                    res *= sigma + dp[(size_t)cidx1[2 * i + s] * n2 + cidx2[2 * j + s]];   /* :198-199 */
                    ++r;
                }
            }
            dp[(size_t)i * n2 + j] = res;                    /* :203 */
            k += res;                                        /* :204 */
        }
    }
    if (counts) { counts[0] = (int64_t)n1 * n2; counts[1] = m; counts[2] = r; }
    free(dp); free(sq1); free(sq2);
    return k;
}

int pk_oracle_abi(void) { return 1; }
