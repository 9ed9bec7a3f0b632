#include "host_util.h"
#include "sdm_b200.h"
extern "C" const char* b2_last_error(void) { return b2::last_error(); }
extern "C" int b2_version(void) { return 100; }
