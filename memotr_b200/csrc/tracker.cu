// tracker.cu -- the per-frame tracker glue on the device ("next" row 1).
//
// The reference runs this part on the host, one `.item()`-style synchronisation per track and per field
// (models/runtime_tracker.py:29-101, models/query_updater.py:243-254, submit_engine.py:89-102): score thresholds, the
// disappear counter, identity assignment for newborn tracks, concatenation + `ids >= 0` filtering of the TrackInstances
// and the result filter.  Here the track table is a FIXED-CAPACITY structure-of-arrays in HBM (rows [0, *n_active) are
// live, the rest is padding that the attention kernels mask as padded keys), so the whole frame loop -- transformer,
// tracker glue, query updater, feedback of the track queries -- is one replayable CUDA graph with no host round trip.
//
// Ordering is the reference's: surviving previous tracks in their old order, then newborns in detect-query order
// (TrackInstances.cat_tracked_instances + boolean mask); identities are max_obj_id + rank.  All integer outputs are
// bit-exact against oracle/tracker.py (tests/test_tracker_gpu.py).
#include "common.cuh"

namespace memotr {
namespace trk {

constexpr int NT = 1024;

__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.f / (1.f + expf(-x)); }  // Tensor.sigmoid (fp32)

// exclusive prefix sum of `flag` over the block (NT threads); returns the position, `total` = block sum
__device__ __forceinline__ int block_scan(int flag, int &total, int *warp_sums) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = flag;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += n;
    }
    warp_sums[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  const int base = warp ? warp_sums[warp - 1] : 0;
  total = warp_sums[31];
  __syncthreads();  // warp_sums is reused by the next call
  return base + incl - flag;
}

// RuntimeTracker.update decisions (runtime_tracker.py:42-57,60-62,76,85-89) + select_active_tracks (query_updater.py:253-254)
__global__ void __launch_bounds__(NT)
tracker_decide_kernel(const float *__restrict__ logits, int nd, int ncls, int cap, const long long *__restrict__ ids,
                      const long long *__restrict__ labels, const long long *__restrict__ dis,
                      const int *__restrict__ n_active_in, float det_thr, float trk_thr, int miss_tol,
                      long long *__restrict__ max_obj_id, long long *__restrict__ ids_o, long long *__restrict__ labels_o,
                      long long *__restrict__ dis_o, int *__restrict__ n_active_o, int *__restrict__ src,
                      unsigned char *__restrict__ pad, int *__restrict__ overflow) {
  pdl_grid_sync();
  __shared__ int warp_sums[32];
  const int tid = threadIdx.x;
  const int na = min(max(*n_active_in, 0), cap);
  const long long first_id = *max_obj_id;
  int running = 0;
  // previous tracks: score of the track's own label against the track threshold, disappear counter, death
  for (int start = 0; start < cap; start += NT) {
    const int i = start + tid;
    int flag = 0;
    long long id = -1, lab = 0, d = 0;
    if (i < na) {
      lab = labels[i];
      const float sc = sigmoidf_ref(logits[(long)(nd + i) * ncls + lab]);
      d = sc < trk_thr ? dis[i] + 1 : 0;
      id = d >= miss_tol ? -1 : ids[i];
      flag = id >= 0;
    }
    int total;
    const int pos = running + block_scan(flag, total, warp_sums);
    if (flag) ids_o[pos] = id, labels_o[pos] = lab, dis_o[pos] = d, src[pos] = i;
    running += total;
  }
  const int kept = running;
  // newborn tracks: detect queries whose best class score reaches the detection threshold
  for (int start = 0; start < nd; start += NT) {
    const int j = start + tid;
    int flag = 0, lab = 0;
    if (j < nd) {
      float m = -1.f;
      for (int c = 0; c < ncls; ++c) {
        const float s = sigmoidf_ref(logits[(long)j * ncls + c]);
        if (s > m) m = s, lab = c;  // first maximum, as torch.max(dim)
      }
      flag = m >= det_thr;
    }
    int total;
    const int pos = running + block_scan(flag, total, warp_sums);
    if (flag && pos < cap) ids_o[pos] = first_id + (pos - kept), labels_o[pos] = lab, dis_o[pos] = 0, src[pos] = cap + j;
    running += total;
  }
  const int n_out = min(running, cap);
  for (int r = tid; r < cap; r += NT) {
    pad[r] = r >= n_out;
    if (r >= n_out) ids_o[r] = -1, labels_o[r] = 0, dis_o[r] = 0, src[r] = -1;
  }
  if (tid == 0) {
    *n_active_o = n_out;
    *max_obj_id = first_id + (n_out - kept);
    if (running > cap) atomicAdd(overflow, running - cap);  // newborns dropped for lack of rows: the host must look
  }
}

struct Table {  // device view of memotr_track_table
  long long *ids, *labels, *dis;
  float *qe, *oe, *lo, *lm, *ref, *box, *logit;
  int *n_active;
};
struct Frame {
  const float *logits, *boxes, *outputs, *last_ref, *auxq;
};

// field hand-off (runtime_tracker.py:43-45,60-71; query_updater.py:246-251): row r of the new table <- its source
__global__ void __launch_bounds__(256)
tracker_gather_kernel(Frame f, Table a, Table b, const int *__restrict__ src, int nd, int ncls, int cap, int C) {
  pdl_grid_sync();
  const int r = blockIdx.x, si = src[r], tid = threadIdx.x;
  const bool prev = si >= 0 && si < cap;
  const int frow = prev ? nd + si : si - cap;  // row of the frame outputs
  for (int c = tid; c < C; c += blockDim.x) {
    float qe = 0.f, oe = 0.f, lo = 0.f, lm = 0.f;
    if (si >= 0) {
      oe = f.outputs[(long)frow * C + c];
      if (prev) qe = a.qe[(long)si * C + c], lo = a.lo[(long)si * C + c], lm = a.lm[(long)si * C + c];
      else qe = f.auxq[(long)frow * C + c], lo = oe, lm = qe;
    }
    b.qe[(long)r * C + c] = qe, b.oe[(long)r * C + c] = oe, b.lo[(long)r * C + c] = lo, b.lm[(long)r * C + c] = lm;
  }
  if (tid < 4) {
    float rf = 0.f, bx = 0.f;
    if (si >= 0) {
      bx = f.boxes[(long)frow * 4 + tid];
      rf = prev ? a.ref[(long)si * 4 + tid] : f.last_ref[(long)frow * 4 + tid];
    }
    b.ref[(long)r * 4 + tid] = rf, b.box[(long)r * 4 + tid] = bx;
  }
  for (int c = tid; c < ncls; c += blockDim.x) b.logit[(long)r * ncls + c] = si >= 0 ? f.logits[(long)frow * ncls + c] : 0.f;
}

__global__ void __launch_bounds__(256) tracker_commit_kernel(Table a, Table b, int ncls, int cap, int C) {
  pdl_grid_sync();
  const int r = blockIdx.x, tid = threadIdx.x;
  for (int c = tid; c < C; c += blockDim.x) {
    const long o = (long)r * C + c;
    a.qe[o] = b.qe[o], a.oe[o] = b.oe[o], a.lo[o] = b.lo[o], a.lm[o] = b.lm[o];
  }
  if (tid < 4) a.ref[(long)r * 4 + tid] = b.ref[(long)r * 4 + tid], a.box[(long)r * 4 + tid] = b.box[(long)r * 4 + tid];
  for (int c = tid; c < ncls; c += blockDim.x) a.logit[(long)r * ncls + c] = b.logit[(long)r * ncls + c];
  if (tid == 0) {
    a.ids[r] = b.ids[r], a.labels[r] = b.labels[r], a.dis[r] = b.dis[r];
    if (r == 0) *a.n_active = *b.n_active;
  }
}

// submit_engine.py:89-102: score and area filters, cxcywh -> xyxy in pixels of the original image
__global__ void tracker_results_kernel(Table a, int ncls, int cap, float score_thr, float area_thr, float ow, float oh,
                                       long long *__restrict__ ids_o, float *__restrict__ xyxy,
                                       float *__restrict__ scores, unsigned char *__restrict__ keep) {
  pdl_grid_sync();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= cap) return;
  const bool live = r < *a.n_active;
  float m = -1.f;
  for (int c = 0; c < ncls; ++c) m = fmaxf(m, sigmoidf_ref(a.logit[(long)r * ncls + c]));
  const float cx = a.box[r * 4], cy = a.box[r * 4 + 1], w = a.box[r * 4 + 2], h = a.box[r * 4 + 3];
  const float area = __fmul_rn(__fmul_rn(__fmul_rn(w, ow), h), oh);  // boxes[:,2] * ori_w * boxes[:,3] * ori_h
  const float hw = __fmul_rn(0.5f, w), hh = __fmul_rn(0.5f, h);
  ids_o[r] = live ? a.ids[r] : -1;
  scores[r] = live ? m : 0.f;
  xyxy[r * 4 + 0] = live ? __fmul_rn(__fsub_rn(cx, hw), ow) : 0.f;
  xyxy[r * 4 + 1] = live ? __fmul_rn(__fsub_rn(cy, hh), oh) : 0.f;
  xyxy[r * 4 + 2] = live ? __fmul_rn(__fadd_rn(cx, hw), ow) : 0.f;
  xyxy[r * 4 + 3] = live ? __fmul_rn(__fadd_rn(cy, hh), oh) : 0.f;
  keep[r] = live && m > score_thr && area > area_thr;
}

static bool table_ok(const memotr_track_table *t) {
  return t && t->ids && t->labels && t->disappear_time && t->query_embed && t->output_embed && t->last_output &&
         t->long_memory && t->ref_pts && t->boxes && t->logits && t->n_active;
}
static Table view(const memotr_track_table *t) {
  return Table{t->ids, t->labels, t->disappear_time, t->query_embed, t->output_embed, t->last_output, t->long_memory,
               t->ref_pts, t->boxes, t->logits, t->n_active};
}

}  // namespace trk
}  // namespace memotr

using namespace memotr;

extern "C" int memotr_tracker_update(const memotr_frame_outputs *frame, int n_det, int ncls, int C,
                                     const memotr_track_table *tracks, const memotr_track_table *scratch, int capacity,
                                     float det_thresh, float track_thresh, int miss_tolerance, long long *max_obj_id,
                                     int *src_index, unsigned char *track_pad, int *overflow, void *stream) {
  MEMOTR_REQUIRE(frame && frame->pred_logits && frame->pred_boxes && frame->outputs && frame->last_ref_pts &&
                     frame->aux_queries,
                 "tracker_update: null frame output");
  MEMOTR_REQUIRE(trk::table_ok(tracks) && trk::table_ok(scratch), "tracker_update: null track table field");
  MEMOTR_REQUIRE(max_obj_id && src_index && track_pad && overflow, "tracker_update: null pointer");
  MEMOTR_REQUIRE(n_det >= 0 && ncls >= 1 && C >= 1 && capacity >= 1 && miss_tolerance >= 0, "tracker_update: bad size");
  cudaStream_t st = (cudaStream_t)stream;
  const trk::Table a = trk::view(tracks), b = trk::view(scratch);
  const trk::Frame f{frame->pred_logits, frame->pred_boxes, frame->outputs, frame->last_ref_pts, frame->aux_queries};
  MEMOTR_LAUNCH(trk::tracker_decide_kernel, 1, trk::NT, 0, st, f.logits, n_det, ncls, capacity, (const long long *)a.ids,
                (const long long *)a.labels, (const long long *)a.dis, (const int *)a.n_active, det_thresh, track_thresh,
                miss_tolerance, max_obj_id, b.ids, b.labels, b.dis, b.n_active, src_index, track_pad, overflow);
  MEMOTR_LAUNCH(trk::tracker_gather_kernel, capacity, 256, 0, st, f, a, b, (const int *)src_index, n_det, ncls, capacity, C);
  MEMOTR_LAUNCH(trk::tracker_commit_kernel, capacity, 256, 0, st, a, b, ncls, capacity, C);
  return check_launch("tracker_update");
}

extern "C" int memotr_tracker_results(const memotr_track_table *tracks, int capacity, int ncls, float score_thresh,
                                      float area_thresh, float ori_w, float ori_h, long long *ids, float *boxes_xyxy,
                                      float *scores, unsigned char *keep, void *stream) {
  MEMOTR_REQUIRE(trk::table_ok(tracks) && ids && boxes_xyxy && scores && keep, "tracker_results: null pointer");
  MEMOTR_REQUIRE(capacity >= 1 && ncls >= 1, "tracker_results: bad size");
  MEMOTR_LAUNCH(trk::tracker_results_kernel, ceil_div(capacity, 128), 128, 0, (cudaStream_t)stream, trk::view(tracks),
                ncls, capacity, score_thresh, area_thresh, ori_w, ori_h, ids, boxes_xyxy, scores, keep);
  return check_launch("tracker_results");
}
