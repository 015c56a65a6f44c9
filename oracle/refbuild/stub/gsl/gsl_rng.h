/* GSL stand-in: only the type and the one call system.c:set_random_numbers makes. */
#ifndef GL_STUB_GSL_RNG_H
#define GL_STUB_GSL_RNG_H
typedef struct gl_stub_gsl_rng gsl_rng;
double gsl_rng_uniform(gsl_rng *r);
#endif
