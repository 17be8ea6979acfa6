// frame_assembler.cpp — what a "frame" is in live operation (host side, no CUDA).
//
// Restates the reference's runtime-sized approximate-time synchroniser
//   skeleton_3d/include/my_message_filters/sync_policies/approximate_time_vec.h:170-217 (add),
//   :262-388 (deque/past bookkeeping), :390-480 (candidate boundaries, virtual times), :488-626 (process)
//   skeleton_3d/include/my_message_filters/synchronizer_vec.h:147-188 (signal / cb)
// configured as in skeleton_3d's main() (queue max(3, 1 + C/4), inter-message lower bound 20 ms, age
// penalty 2.0; S3D:1218-1223), followed by the worker loop's per-frame gating (S3D:1029-1057): the pivot is
// the newest stamp, a frame whose pivot does not advance is skipped, and cameras lagging the pivot by more
// than 67 ms are blanked (treated as "no persons").
//
// Messages are handled as (stamp, caller id) pairs - the payload stays with the caller, who packs the frames
// the assembler emits into the [n_frames][C][p_max] arrays of the batch ABI. Time arithmetic follows
// ros::Time / ros::Duration: integer nanoseconds, and Duration * double goes through seconds as a double and
// back (DurationBase::operator*, fromSec) exactly as roscpp does. roscpp is not available in the build
// container: parity of this component is pinned only by the independent Python restatement under oracle/.
#include <cmath>
#include <cstdint>
#include <deque>
#include <limits>
#include <vector>

#include "ses3d.h"

namespace {

const int64_t kNs = 1000000000LL;

struct Msg {
  int64_t stamp;  // ns
  int64_t id;
  bool valid = false;
};

// ros::Duration::toSec / fromSec round trip used by Duration * double
double dur_to_sec(int64_t ns) {
  int64_t sec = ns / kNs, nsec = ns % kNs;
  if (nsec < 0) { nsec += kNs; sec -= 1; }   // roscpp keeps nsec in [0, 1e9)
  return static_cast<double>(sec) + 1e-9 * static_cast<double>(nsec);
}
int64_t dur_from_sec(double d) {
  const int64_t sec = static_cast<int64_t>(std::floor(d));
  int64_t nsec = static_cast<int64_t>(std::llround((d - static_cast<double>(sec)) * 1e9));  // boost::math::round
  return sec * kNs + nsec;   // the rollover normalisation of roscpp is the identity on a single ns count
}
int64_t dur_scale(int64_t ns, double scale) { return dur_from_sec(dur_to_sec(ns) * scale); }
double time_to_sec(int64_t ns) { return static_cast<double>(ns / kNs) + 1e-9 * static_cast<double>(ns % kNs); }

struct Frame {
  std::vector<Msg> msgs;
  std::vector<uint8_t> blank;
  int pivot = -1;
};

}  // namespace

struct ses3d_assembler_s {
  ses3d_assembler_config cfg;
  uint32_t n = 0, queue_size = 0, NO_PIVOT = 0;
  std::vector<std::deque<Msg>> deques;
  std::vector<std::vector<Msg>> past;
  std::vector<Msg> candidate;
  uint32_t num_non_empty = 0;
  int64_t candidate_start = 0, candidate_end = 0, pivot_time = 0;
  uint32_t pivot = 0;
  int64_t max_interval = std::numeric_limits<int64_t>::max();
  std::vector<char> has_dropped;
  std::vector<int64_t> lower_bound;
  std::deque<Frame> ready;
  double last_stamp = 0.0;   // S3D:1010
  int64_t n_emitted = 0, n_skipped = 0, n_blanked = 0, n_dropped = 0, n_signalled = 0;

  // ---- ATV:262-388
  void deque_delete_front(uint32_t i) {
    deques[i].pop_front();
    if (deques[i].empty()) --num_non_empty;
  }
  void deque_move_front_to_past(uint32_t i) {
    past[i].push_back(deques[i].front());
    deques[i].pop_front();
    if (deques[i].empty()) --num_non_empty;
  }
  void make_candidate() {
    for (uint32_t i = 0; i < n; ++i) candidate[i] = deques[i].front();
    for (uint32_t i = 0; i < n; ++i) past[i].clear();
  }
  void recover_n(size_t count, uint32_t i) {
    while (count > 0) { deques[i].push_front(past[i].back()); past[i].pop_back(); --count; }
    if (!deques[i].empty()) ++num_non_empty;
  }
  void recover(uint32_t i) {
    while (!past[i].empty()) { deques[i].push_front(past[i].back()); past[i].pop_back(); }
    if (!deques[i].empty()) ++num_non_empty;
  }
  void recover_and_delete(uint32_t i) {
    while (!past[i].empty()) { deques[i].push_front(past[i].back()); past[i].pop_back(); }
    deques[i].pop_front();
    if (!deques[i].empty()) ++num_non_empty;
  }
  void publish_candidate() {  // ATV:371-386 -> SYV:147-161 -> worker gating
    signal(candidate);
    for (Msg& m : candidate) m.valid = false;
    pivot = NO_PIVOT;
    num_non_empty = 0;
    for (uint32_t i = 0; i < n; ++i) recover_and_delete(i);
  }

  // ---- ATV:390-480
  void boundary(uint32_t& index, int64_t& time, bool end) const {
    time = deques[0].front().stamp;
    index = 0;
    for (uint32_t i = 1; i < n; ++i) {
      const int64_t t = deques[i].front().stamp;
      if ((t < time) ^ end) { time = t; index = i; }
    }
  }
  int64_t virtual_time(uint32_t i) const {
    if (deques[i].empty()) {
      const int64_t lb = past[i].back().stamp + lower_bound[i];
      return lb > pivot_time ? lb : pivot_time;
    }
    return deques[i].front().stamp;
  }
  void virtual_boundary(uint32_t& index, int64_t& time, bool end) const {
    std::vector<int64_t> vt(n);
    for (uint32_t i = 0; i < n; ++i) vt[i] = virtual_time(i);
    time = vt[0];
    index = 0;
    for (uint32_t i = 0; i < n; ++i)
      if ((vt[i] < time) ^ end) { time = vt[i]; index = i; }
  }

  // ---- ATV:488-626
  void process() {
    const double scale = 1.0 + cfg.age_penalty;
    while (num_non_empty == n) {
      int64_t end_time, start_time;
      uint32_t end_index, start_index;
      boundary(end_index, end_time, true);
      boundary(start_index, start_time, false);
      for (uint32_t i = 0; i < n; ++i)
        if (i != end_index) has_dropped[i] = 0;
      if (pivot == NO_PIVOT) {
        if (end_time - start_time > max_interval) { deque_delete_front(start_index); continue; }
        if (has_dropped[end_index]) { deque_delete_front(start_index); continue; }
        make_candidate();
        candidate_start = start_time;
        candidate_end = end_time;
        pivot = end_index;
        pivot_time = end_time;
        deque_move_front_to_past(start_index);
      } else {
        if (dur_scale(end_time - candidate_end, scale) >= (start_time - candidate_start)) {
          deque_move_front_to_past(start_index);
        } else {
          make_candidate();
          candidate_start = start_time;
          candidate_end = end_time;
          deque_move_front_to_past(start_index);
        }
      }
      if (start_index == pivot) {
        publish_candidate();
      } else if (dur_scale(end_time - candidate_end, scale) >= (pivot_time - candidate_start)) {
        publish_candidate();
      } else if (num_non_empty < n) {
        std::vector<int> virtual_moves(n, 0);
        while (true) {
          int64_t v_end, v_start;
          uint32_t v_end_i, v_start_i;
          virtual_boundary(v_end_i, v_end, true);
          virtual_boundary(v_start_i, v_start, false);
          if (dur_scale(v_end - candidate_end, scale) >= (pivot_time - candidate_start)) {
            publish_candidate();
            break;
          }
          if (dur_scale(v_end - candidate_end, scale) < (v_start - candidate_start)) {
            num_non_empty = 0;
            for (uint32_t i = 0; i < n; ++i) recover_n((size_t)virtual_moves[i], i);
            break;
          }
          deque_move_front_to_past(v_start_i);
          virtual_moves[v_start_i]++;
        }
      }
    }
  }

  // ---- ATV:170-217
  void add(uint32_t i, int64_t stamp, int64_t id) {
    Msg m;
    m.stamp = stamp; m.id = id; m.valid = true;
    deques[i].push_back(m);
    if (deques[i].size() == 1) {
      ++num_non_empty;
      if (num_non_empty == n) process();
    }
    if (deques[i].size() + past[i].size() > queue_size) {
      num_non_empty = 0;
      for (uint32_t j = 0; j < n; ++j) recover(j);
      deques[i].pop_front();
      has_dropped[i] = 1;
      ++n_dropped;
      if (pivot != NO_PIVOT) {
        for (Msg& c : candidate) c.valid = false;
        pivot = NO_PIVOT;
        process();
      }
    }
  }

  // ---- worker gating, S3D:1029-1057
  void signal(const std::vector<Msg>& tuple) {
    ++n_signalled;
    if (cfg.max_sync_diff_s < 0.0) {   // synchroniser only: every tuple is handed on as it is (the node gates itself)
      Frame f;
      f.msgs = tuple;
      f.blank.assign(n, 0);
      f.pivot = 0;
      for (uint32_t i = 1; i < n; ++i)
        if (tuple[i].stamp > tuple[f.pivot].stamp) f.pivot = (int)i;
      ready.push_back(f);
      ++n_emitted;
      return;
    }
    double t_max = 0.0;
    int t_max_idx = -1;
    for (uint32_t i = 0; i < n; ++i) {
      const double t = time_to_sec(tuple[i].stamp);
      if (t > t_max) { t_max = t; t_max_idx = (int)i; }
    }
    if (t_max_idx < 0) { ++n_skipped; return; }
    const double delta_t = t_max - last_stamp;
    if (delta_t <= 0.0) { ++n_skipped; return; }   // re-used message or time jumped backwards (S3D:1043-1046)
    last_stamp = t_max;
    Frame f;
    f.msgs = tuple;
    f.blank.assign(n, 0);
    f.pivot = t_max_idx;
    for (uint32_t i = 0; i < n; ++i) {
      const double dt = t_max - time_to_sec(tuple[i].stamp);
      if (dt > cfg.max_sync_diff_s) { f.blank[i] = 1; ++n_blanked; }   // S3D:1049-1057
    }
    ready.push_back(f);
    ++n_emitted;
  }
};

extern "C" {

int ses3d_assembler_default_config(int32_t n_cams, ses3d_assembler_config* cfg) {
  if (!cfg || n_cams < 1) return SES3D_E_INVALID;
  cfg->n_cams = n_cams;
  const uint32_t q = 1u + (uint32_t)n_cams / 4u;
  cfg->queue_size = q > 3u ? q : 3u;                 // std::max(3u, 1 + NUM_CAMERAS / 4), S3D:1219
  cfg->inter_message_lower_bound_ns = 20000000LL;    // ros::Duration(0.020), S3D:1220
  cfg->age_penalty = 2.0;                            // S3D:1221
  cfg->max_interval_ns = -1;                         // ros::DURATION_MAX (ATV:84)
  cfg->max_sync_diff_s = 0.067;                      // g_max_sync_diff, S3D:64
  return SES3D_OK;
}

int ses3d_assembler_create(const ses3d_assembler_config* cfg, ses3d_assembler* out) {
  if (!cfg || !out || cfg->n_cams < 1 || cfg->queue_size < 1 || cfg->age_penalty < 0.0 ||
      cfg->inter_message_lower_bound_ns < 0)
    return SES3D_E_INVALID;
  ses3d_assembler_s* a = new (std::nothrow) ses3d_assembler_s;
  if (!a) return SES3D_E_NOMEM;
  a->cfg = *cfg;
  a->n = (uint32_t)cfg->n_cams;
  a->queue_size = cfg->queue_size;
  a->NO_PIVOT = a->n;
  a->pivot = a->NO_PIVOT;
  a->deques.resize(a->n);
  a->past.resize(a->n);
  a->candidate.resize(a->n);
  a->has_dropped.assign(a->n, 0);
  a->lower_bound.assign(a->n, cfg->inter_message_lower_bound_ns);
  if (cfg->max_interval_ns >= 0) a->max_interval = cfg->max_interval_ns;
  *out = a;
  return SES3D_OK;
}

int ses3d_assembler_destroy(ses3d_assembler a) {
  delete a;
  return SES3D_OK;
}

int ses3d_assembler_add(ses3d_assembler a, int32_t cam, int64_t stamp_ns, int64_t msg_id) {
  if (!a || cam < 0 || (uint32_t)cam >= a->n || stamp_ns < 0) return SES3D_E_INVALID;
  const size_t before = a->ready.size();
  a->add((uint32_t)cam, stamp_ns, msg_id);
  return (int)(a->ready.size() - before);
}

int ses3d_assembler_pop(ses3d_assembler a, int64_t* ids, int64_t* stamps_ns, uint8_t* blank, int32_t* pivot) {
  if (!a || !ids) return SES3D_E_INVALID;
  if (a->ready.empty()) return 0;
  const Frame& f = a->ready.front();
  for (uint32_t i = 0; i < a->n; ++i) {
    ids[i] = f.msgs[i].id;
    if (stamps_ns) stamps_ns[i] = f.msgs[i].stamp;
    if (blank) blank[i] = f.blank[i];
  }
  if (pivot) *pivot = f.pivot;
  a->ready.pop_front();
  return 1;
}

// The 1-slot latest-wins mailbox between the synchroniser callback (ROS spinner thread) and the worker thread
// (skeletonCallback / skeletonThreadCallback, S3D:999-1025) as a deterministic replay. The callback stores frame i in
// the slot at t_ready[i] and overwrites whatever is still there; the worker, when idle, takes the slot's content (or
// sleeps until the next store) and is busy for busy[i]. taken[i] = 1: processed, 0: overwritten before the worker saw it.
int ses3d_mailbox_replay(int32_t n, const int64_t* t_ready_ns, const int64_t* busy_ns, uint8_t* taken,
                         int64_t* t_start_ns) {
  if (n < 0 || (n > 0 && (!t_ready_ns || !busy_ns || !taken))) return SES3D_E_INVALID;
  for (int32_t i = 1; i < n; ++i)
    if (t_ready_ns[i] < t_ready_ns[i - 1]) return SES3D_E_INVALID;
  int64_t free_at = INT64_MIN;   // the worker starts out waiting on the condition variable
  int processed = 0;
  int32_t i = 0;
  while (i < n) {
    int32_t j = i;   // newest frame stored by the time the worker looks at the slot
    if (free_at > t_ready_ns[i])
      while (j + 1 < n && t_ready_ns[j + 1] <= free_at) ++j;
    for (int32_t k = i; k < j; ++k) {
      taken[k] = 0;
      if (t_start_ns) t_start_ns[k] = -1;
    }
    const int64_t start = free_at > t_ready_ns[j] ? free_at : t_ready_ns[j];
    taken[j] = 1;
    if (t_start_ns) t_start_ns[j] = start;
    free_at = start + (busy_ns[j] > 0 ? busy_ns[j] : 0);
    ++processed;
    i = j + 1;
  }
  return processed;
}

int ses3d_assembler_stats(ses3d_assembler a, int64_t stats[5]) {
  if (!a || !stats) return SES3D_E_INVALID;
  stats[0] = a->n_emitted; stats[1] = a->n_skipped; stats[2] = a->n_blanked; stats[3] = a->n_dropped;
  stats[4] = a->n_signalled;
  return SES3D_OK;
}

}  // extern "C"
