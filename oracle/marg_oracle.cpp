// oracle/marg_oracle.cpp -- placeholder translation unit; marginalization restatement lands here.
#include "oracle.h"
extern "C" int oracle_marginalize(const bvio_window*, const bvio_opts*, int, bvio_prior_out*) { return BVIO_ERR_UNSUPPORTED; }
