/*
 * oracle/ref_harness_at3.cpp — TEST INFRASTRUCTURE (see ref_harness.cpp header).
 * ATRAC3 stage taps; compiled with -fno-access-control so it can read encoder internals.
 * (filled in as the ATRAC3 path is built)
 */
#include <cstdint>
extern "C" int ref_at3_taps_version(void) { return 0; }
