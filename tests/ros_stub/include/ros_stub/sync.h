// ros_stub/sync.h — message_filters::Subscriber plus the reference's runtime-sized approximate-time synchroniser
// (my_message_filters::SynchronizerVec / sync_policies::ApproximateTimeVec, skeleton_3d/include/my_message_filters),
// API as the shim node uses it (S3D:1172-1178, 1218-1223). The policy itself is NOT restated here: the stub drives the
// product's frame assembler (ses3d_assembler_*, csrc/frame_assembler.cpp) in its synchroniser-only mode, so the replay
// test also exercises that code. TEST INFRASTRUCTURE ONLY (see core.h).
#pragma once
#include <ses3d.h>

#include "ros_stub/msgs.h"

namespace message_filters {

template <class M>
class Subscriber {
 public:
  typedef std::shared_ptr<M const> MConstPtr;
  Subscriber() {}
  void subscribe(ros::NodeHandle& nh, const std::string& topic, uint32_t queue,
                 const ros::TransportHints& hints = ros::TransportHints()) {
    sub_ = nh.subscribe<M>(topic, queue, [this](const MConstPtr& m) { if (cb_) cb_(m); }, ros::VoidConstPtr(), hints);
  }
  void registerCallback(const std::function<void(const MConstPtr&)>& cb) { cb_ = cb; }
 private:
  ros::Subscriber sub_;
  std::function<void(const MConstPtr&)> cb_;
};

namespace sync_policies {
template <class M>
struct ApproximateTimeVec {
  uint32_t queue_size, n;
  ros::Duration lower_bound{0, 0};
  double age_penalty = 0.1;
  ApproximateTimeVec(uint32_t queue_size_, uint32_t n_) : queue_size(queue_size_), n(n_) {}
  void setInterMessageLowerBound(ros::Duration d) { lower_bound = d; }
  void setAgePenalty(double p) { age_penalty = p; }
  typedef M Message;
};
}  // namespace sync_policies

template <class Policy>
class SynchronizerVec {
 public:
  typedef typename Policy::Message M;
  typedef std::shared_ptr<M const> MConstPtr;
  SynchronizerVec(const Policy& policy, std::vector<Subscriber<M>>& subs) : policy_(policy) {
    ses3d_assembler_config cfg;
    ses3d_assembler_default_config((int32_t)policy.n, &cfg);
    cfg.queue_size = policy.queue_size;
    cfg.inter_message_lower_bound_ns = policy.lower_bound.toNSec();
    cfg.age_penalty = policy.age_penalty;
    cfg.max_sync_diff_s = -1.0;   // synchroniser only; the node's worker does the gating itself
    if (ses3d_assembler_create(&cfg, &asm_) != SES3D_OK) throw std::runtime_error("ses3d_assembler_create failed");
    for (uint32_t i = 0; i < policy.n; ++i)
      subs[i].registerCallback([this, i](const MConstPtr& m) { add(i, m); });
  }
  ~SynchronizerVec() { ses3d_assembler_destroy(asm_); }
  SynchronizerVec(const SynchronizerVec&) = delete;
  void registerCallback(const std::function<void(const std::vector<MConstPtr>&)>& cb) { cb_ = cb; }

 private:
  void add(uint32_t cam, const MConstPtr& m) {
    const int64_t id = (int64_t)store_.size();
    store_.push_back(m);
    const int ready = ses3d_assembler_add(asm_, (int32_t)cam, (int64_t)m->header.stamp.toNSec(), id);
    std::vector<int64_t> ids(policy_.n);
    for (int k = 0; k < ready; ++k) {
      if (ses3d_assembler_pop(asm_, ids.data(), nullptr, nullptr, nullptr) != 1) break;
      std::vector<MConstPtr> tuple(policy_.n);
      for (uint32_t i = 0; i < policy_.n; ++i) tuple[i] = store_[(size_t)ids[i]];
      if (cb_) cb_(tuple);
    }
  }
  Policy policy_;
  ses3d_assembler asm_ = nullptr;
  std::vector<MConstPtr> store_;   // replay-sized; a live system would drop what the policy has released
  std::function<void(const std::vector<MConstPtr>&)> cb_;
};

}  // namespace message_filters
