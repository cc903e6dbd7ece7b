/* TEST INFRASTRUCTURE ONLY -- never part of the product path.
 *
 * Library-level harness around the UNMODIFIED reference sources (compiled in place from
 * /root/reference/src by oracle/Makefile into oracle/_ref/libigdref.so).  It supplies what the
 * reference CLI's translation unit supplies -- the process globals of src/igd.c:14-19 and the
 * open/teardown sequence of src/igd_search.c:917-925,974 -- so that tests can call getOverlaps,
 * getOverlaps_v, seqOverlaps, getMap ... directly and read hits[] / sm[] / hitmap as binary.
 */
#include "igd_search.h"

void *hc;
iGD_t *IGD;
gdata_t *gData = NULL;
gdata0_t *gData0 = NULL;
int32_t preIdx, preChr, tile_size;
FILE *fP;

/* returns nFiles, or -1 */
int ref_open(const char *igdName)
{
    char tmp[1024];
    FILE *fi = fopen(igdName, "rb");
    if (!fi) return -1;
    fclose(fi);
    IGD = get_igdinfo((char *)igdName);
    snprintf(tmp, sizeof(tmp) - 16, "%s", igdName);
    char *dot = strrchr(tmp, '.');
    if (!dot) return -1;
    *dot = '\0';
    strcat(tmp, "_index.tsv");
    IGD->finfo = get_fileinfo(tmp, &IGD->nFiles);
    if (!IGD->finfo) return -1;
    fP = fopen(igdName, "rb");
    preChr = -6, preIdx = -8;
    return IGD->nFiles;
}

void ref_close(void)
{
    if (fP) fclose(fP);
    fP = NULL;
    if (gData) free(gData);
    gData = NULL;
    if (gData0) free(gData0);
    gData0 = NULL;
    /* the reference leaks IGD's members on purpose-built paths; a test process does not care */
}

int32_t ref_nfiles(void) { return IGD->nFiles; }
int32_t ref_file_nr(int32_t i) { return IGD->finfo[i].nr; }
const char *ref_file_name(int32_t i) { return IGD->finfo[i].fileName; }
int32_t ref_nbp(void) { return IGD->nbp; }
int32_t ref_nctg(void) { return IGD->nCtg; }
int32_t ref_ntile(int32_t c) { return IGD->nTile[c]; }
int32_t ref_gtype(void) { return IGD->gType; }

/* hitmap as one dense row-major uint32[nFiles*nFiles] */
int64_t ref_getmap_dense(uint32_t *dense, int32_t v)
{
    int32_t n = IGD->nFiles, i;
    uint32_t **rows = malloc(n * sizeof(uint32_t *));
    for (i = 0; i < n; i++) rows[i] = dense + (size_t)i * n;
    int64_t r = v > 0 ? getMap_v(rows, v) : getMap(rows);
    free(rows);
    return r;
}
