/* stand-in for R.h (see Rinternals.h in this directory) */
#ifndef MCLB_R_STUB_R_H
#define MCLB_R_STUB_R_H
#include <stdint.h>
#endif
