/* oracle/gsl_shim: intentionally empty.  The reference's sampletau/c_sample_tau.c:11-19
 * includes this GSL header but uses no symbol from it; GSL is not installed in this image.
 * Test infrastructure only (see oracle/README.md). */
