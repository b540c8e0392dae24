// capi.cu -- ABI bookkeeping entry points of libmemotr_b200.so (include/memotr_b200.h).
#include "common.cuh"

using memotr::fail;

extern "C" int memotr_abi_version(void) { return MEMOTR_ABI_VERSION; }
extern "C" const char *memotr_last_error(void) { return memotr::err_buf(); }

// ---- SM budget of the persistent kernels (include/memotr_b200.h: memotr_set_sm_budget) ---------------------------------
static int g_sm_budget = 0;
extern "C" int memotr_set_sm_budget(int n_sm) {
  MEMOTR_REQUIRE(n_sm >= 0, "set_sm_budget: negative budget");
  g_sm_budget = n_sm;
  return MEMOTR_OK;
}
namespace memotr {
int sm_limit(int n_sm) { return (g_sm_budget > 0 && g_sm_budget < n_sm) ? g_sm_budget : n_sm; }
}  // namespace memotr

// ---- device-side interval timer usable inside CUDA graphs ------------------------------------------------------------
// bench.py measures the dominant kernel's duration live, inside the timed region, on the launching stream.  Events
// recorded with cudaEventRecordExternal become event-record nodes when the stream is being captured, and -- unlike
// plain captured events -- may be read with cudaEventElapsedTime after each graph launch.
struct MemotrTimer {
  int n;
  cudaEvent_t *ev;
};

extern "C" void *memotr_timer_create(int n_events) {
  if (n_events <= 0) return nullptr;
  MemotrTimer *t = new MemotrTimer{n_events, new cudaEvent_t[n_events]};
  for (int i = 0; i < n_events; ++i)
    if (cudaEventCreate(&t->ev[i]) != cudaSuccess) return nullptr;
  return t;
}
extern "C" void memotr_timer_destroy(void *timer) {
  MemotrTimer *t = (MemotrTimer *)timer;
  if (!t) return;
  for (int i = 0; i < t->n; ++i) cudaEventDestroy(t->ev[i]);
  delete[] t->ev;
  delete t;
}
extern "C" int memotr_timer_record(void *timer, int idx, void *stream) {
  MemotrTimer *t = (MemotrTimer *)timer;
  MEMOTR_REQUIRE(t && idx >= 0 && idx < t->n, "timer_record: bad arguments");
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing((cudaStream_t)stream, &cs);
  cudaError_t e = cudaEventRecordWithFlags(t->ev[idx], (cudaStream_t)stream,
                                           cs == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault);
  if (e != cudaSuccess) return memotr::fail(MEMOTR_ECUDA, "timer_record: %s", cudaGetErrorString(e));
  return MEMOTR_OK;
}
extern "C" int memotr_timer_elapsed_ms(void *timer, int idx_start, int idx_stop, float *ms) {
  MemotrTimer *t = (MemotrTimer *)timer;
  MEMOTR_REQUIRE(t && ms && idx_start >= 0 && idx_stop >= 0 && idx_start < t->n && idx_stop < t->n,
                 "timer_elapsed: bad arguments");
  cudaError_t e = cudaEventElapsedTime(ms, t->ev[idx_start], t->ev[idx_stop]);
  if (e != cudaSuccess) return memotr::fail(MEMOTR_ECUDA, "timer_elapsed: %s", cudaGetErrorString(e));
  return MEMOTR_OK;
}
