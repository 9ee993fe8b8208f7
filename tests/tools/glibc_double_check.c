#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdint.h>
#include "oracle_common.h"
static uint64_t s = 88172645463325252ULL;
static uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static double urand(void) { return (rnd() >> 11) * (1.0 / 9007199254740992.0); }
int main(void) {
    long bad_log = 0, bad_exp = 0, n = 40000000;
    for (long i = 0; i < n; i++) {
        double x;
        int m = i % 4;
        if (m == 0) x = exp(-64 * urand() + 4);                 /* energies 1e-28 .. 50 */
        else if (m == 1) x = 0.9 + 0.2 * urand();               /* near 1 */
        else if (m == 2) { uint64_t u = rnd() & 0x7fefffffffffffffULL; memcpy(&x, &u, 8); }
        else x = (float)exp(-28 * urand());                      /* float-valued, like the encoder */
        double a = log(x), b = og_log(x);
        if (memcmp(&a, &b, 8) && !(a != a && b != b)) { if (bad_log < 5) printf("log %a: %a vs %a\n", x, a, b); bad_log++; }
        double y = (m == 2) ? (urand() - 0.5) * 1020 : -40 * urand() + 3;
        a = exp(y); b = og_exp(y);
        if (memcmp(&a, &b, 8) && !(a != a && b != b)) { if (bad_exp < 5) printf("exp %a: %a vs %a\n", y, a, b); bad_exp++; }
    }
    printf("n=%ld mismatches log=%ld exp=%ld\n", n, bad_log, bad_exp);
    return 0;
}
