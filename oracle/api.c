/* oracle/api.c -- lifecycle of the CPU oracle (test infrastructure only).
 * Mirrors RAD:Initialize (GEOS_RadiationGridComp.F90:564-580: set_inhomogeneity(ih=1),
 * initialize_cloud_subcol_gen defaults) followed by rrtmg_lw_ini / rrtmg_sw_ini. */
#include "internal.h"

static int g_ready = 0;

int oracle_init(const char *blob_path) {
    if (g_ready) return 0;
    int rc = blob_load(blob_path);
    if (rc) return rc;
    rc = lw_init();
    if (rc) return -10 + rc;
    rc = sw_init();
    if (rc) return -20 + rc;
    rc = oracle_set_mcica(1, 0);
    if (rc) return -30 + rc;
    g_ready = 1;
    return 0;
}

void oracle_finalize(void) {
    if (!g_ready) return;
    lw_free();
    sw_free();
    blob_free();
    g_ready = 0;
}
