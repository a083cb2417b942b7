/* TEST INFRASTRUCTURE (oracle). Link shim used with -Wl,--wrap=time so that the
 * clock-seeded reference programs (e.g. CASCL_1024_L8.c:164, BP_1024.c:136) can be run
 * with a chosen seed without touching their source: time() returns $POLAR_REF_TIME. */
#include <stdlib.h>
long __wrap_time(long *t)
{
    const char *s = getenv("POLAR_REF_TIME");
    long v = s ? atol(s) : 0;
    if (t) *t = v;
    return v;
}
