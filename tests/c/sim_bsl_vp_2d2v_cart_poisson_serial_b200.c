/*
 * The GPU counterpart of the reference executable
 *     sim_bsl_vp_2d2v_cart_poisson_serial <namelist file without .nml>
 * (simulations/parallel/bsl_vp_2d2v_cart_poisson_serial/sim_bsl_vp_2d2v_cart_poisson_serial.F90): a plain C program
 * on top of the C ABI, the way a host written in any language drives it.  Reads the same namelist, writes the same
 * thdiag.dat ('(13g20.12)' rows) into the working directory.
 *   usage: sim_bsl_vp_2d2v_cart_poisson_serial_b200 <namelist> [thdiag file]
 */
#include <stdio.h>

#include "sll_b200.h"

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s <namelist file, with or without .nml> [thdiag file]\n", argv[0]);
        return 2;
    }
    const char *out = argc > 2 ? argv[2] : "thdiag.dat";
    int rc = sllb_init(0);
    if (rc == SLLB_OK) rc = sllb_sim4d_run_namelist(argv[1], NULL, out);
    if (rc != SLLB_OK) {
        /* the reference prints and stops (SLL_ERROR); so do we */
        fprintf(stderr, "sim_bsl_vp_2d2v_cart_poisson_serial_b200: error %d: %s\n", rc, sllb_last_error());
        return 1;
    }
    printf("#run finished: %s\n", out);
    return 0;
}
