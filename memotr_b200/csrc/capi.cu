// capi.cu -- ABI bookkeeping entry points of libmemotr_b200.so (include/memotr_b200.h).
#include "common.cuh"

extern "C" int memotr_abi_version(void) { return MEMOTR_ABI_VERSION; }
extern "C" const char *memotr_last_error(void) { return memotr::err_buf(); }
