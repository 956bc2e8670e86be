// rf_lib.cu -- library-level entry points of include/rf_b200.h.
#include "rf_common.h"

extern "C" const char *rf_last_error(void) { return rf::last_error_ref().c_str(); }
