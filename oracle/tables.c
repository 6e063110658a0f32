/* oracle/tables.c -- reader for the extracted reference data blob (test infrastructure only).
 * Format documented in tools/extract_tables.py. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "internal.h"

static unsigned char *g_blob = NULL;
static int g_nent = 0;
static BlobEntry *g_ent = NULL;

int blob_load(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    blob_free();
    g_blob = (unsigned char *)malloc((size_t)sz);
    if (fread(g_blob, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); return -2; }
    fclose(f);
    if (memcmp(g_blob, "RRTMGTB1", 8) != 0) return -3;
    memcpy(&g_nent, g_blob + 8, 4);
    g_ent = (BlobEntry *)calloc((size_t)g_nent, sizeof(BlobEntry));
    size_t pos = 12;
    for (int i = 0; i < g_nent; ++i) {
        memcpy(g_ent[i].name, g_blob + pos, 48);
        memcpy(&g_ent[i].dtype, g_blob + pos + 48, 4);
        memcpy(&g_ent[i].ndim, g_blob + pos + 52, 4);
        memcpy(g_ent[i].dims, g_blob + pos + 56, 24);
        memcpy(&g_ent[i].offset, g_blob + pos + 80, 8);
        memcpy(&g_ent[i].nbytes, g_blob + pos + 88, 8);
        pos += 96;
    }
    return 0;
}

void blob_free(void) {
    free(g_blob); g_blob = NULL;
    free(g_ent); g_ent = NULL;
    g_nent = 0;
}

static const BlobEntry *find(const char *name) {
    for (int i = 0; i < g_nent; ++i)
        if (strcmp(g_ent[i].name, name) == 0) return &g_ent[i];
    return NULL;
}

const double *blob_f64(const char *name, int *n) {
    const BlobEntry *e = find(name);
    if (!e || e->dtype != 0) { if (n) *n = 0; return NULL; }
    if (n) *n = (int)(e->nbytes / 8);
    return (const double *)(g_blob + e->offset);
}

const int *blob_i32(const char *name, int *n) {
    const BlobEntry *e = find(name);
    if (!e || e->dtype != 1) { if (n) *n = 0; return NULL; }
    if (n) *n = (int)(e->nbytes / 4);
    return (const int *)(g_blob + e->offset);
}
