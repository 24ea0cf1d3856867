/* LD_PRELOAD shim: time() returns $MIMC3_FAKE_TIME (the reference seeds rand() with time(NULL),
 * MIMC_module.c:516).  Test infrastructure only. */
#include <stdlib.h>
#include <time.h>
time_t time(time_t *t) {
    const char *s = getenv("MIMC3_FAKE_TIME");
    time_t v = s ? (time_t)atoll(s) : (time_t)1700000000;
    if (t) *t = v;
    return v;
}
