/* stand-in for R_ext/Random.h (see ../Rinternals.h) */
#ifndef MCLB_R_STUB_RANDOM_H
#define MCLB_R_STUB_RANDOM_H
void GetRNGstate(void);
void PutRNGstate(void);
double unif_rand(void);
#endif
