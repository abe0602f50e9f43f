/*
 * oracle/ref/faketime.c -- TEST INFRASTRUCTURE.
 * LD_PRELOAD shim: time() returns $HW1F_FAKE_TIME so that the reference
 * binaries' time(NULL)-derived seeds (src/1_bond_pricing.cu:53,
 * src/2_option_pricing.cu:128,223, src/3_sensitivity_analysis.cu:539,713,
 * src/benchmark_reductions.cu:95) are reproducible without touching the
 * reference sources.
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <time.h>

time_t time(time_t* out)
{
    const char* s = getenv("HW1F_FAKE_TIME");
    time_t t = s ? (time_t)strtoll(s, NULL, 10) : (time_t)1700000000;
    if (out) *out = t;
    return t;
}
