// oracle/marg_oracle.cpp -- the marginalization restatement lives in ba_oracle.cpp (it shares the factor
// evaluation code of that translation unit); this file only keeps the Makefile's source list stable.
#include "oracle.h"
