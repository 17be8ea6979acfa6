// ros_stub/core.h — a minimal, single-process stand-in for the parts of roscpp the shim nodes under ros_shim/ use.
//
// TEST INFRASTRUCTURE ONLY. ROS is not installed in this image, so the shim nodes (the files a maintainer drops into
// the catkin workspace) could otherwise not even be syntax-checked. This stub lets them be compiled unchanged and RUN:
// a scenario file (parameters, tf transforms, a time-ordered list of serialised messages) stands in for the ROS master,
// the bag player and tf; everything the node publishes is serialised (ROS 1 wire format) into an output file. The wire
// format here is written independently of smartedgesensor3dhumanpose_b200/csrc/wire.cpp, so the replay test also
// cross-checks the two codecs. Nothing of the product links against or includes this directory.
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace ros {

// ---------------------------------------------------------------------------------------------- time
struct Duration {
  int32_t sec = 0, nsec = 0;
  Duration() {}
  Duration(int32_t s, int32_t ns) : sec(s), nsec(ns) {}
  explicit Duration(double t) { fromSec(t); }
  Duration& fromSec(double t) {
    const int64_t ns = (int64_t)std::floor(t * 1e9 + 0.5);
    sec = (int32_t)std::floor((double)ns / 1e9);
    nsec = (int32_t)(ns - (int64_t)sec * 1000000000LL);
    return *this;
  }
  double toSec() const { return (double)sec + 1e-9 * (double)nsec; }
  int64_t toNSec() const { return (int64_t)sec * 1000000000LL + nsec; }
  bool sleep() const { return true; }   // the stub never waits
};

struct Time {
  uint32_t sec = 0, nsec = 0;
  Time() {}
  Time(uint32_t s, uint32_t ns) : sec(s), nsec(ns) {}
  explicit Time(double t) { fromNSec((uint64_t)std::floor(t * 1e9 + 0.5)); }
  Time& fromNSec(uint64_t t) { sec = (uint32_t)(t / 1000000000ULL); nsec = (uint32_t)(t % 1000000000ULL); return *this; }
  double toSec() const { return (double)sec + 1e-9 * (double)nsec; }
  uint64_t toNSec() const { return (uint64_t)sec * 1000000000ULL + nsec; }
  bool operator<(const Time& o) const { return toNSec() < o.toNSec(); }
  bool operator>(const Time& o) const { return toNSec() > o.toNSec(); }
  bool operator==(const Time& o) const { return sec == o.sec && nsec == o.nsec; }
  bool operator!=(const Time& o) const { return !(*this == o); }
  bool isZero() const { return sec == 0 && nsec == 0; }
  static Time now();
};

struct Rate {
  explicit Rate(double) {}
  bool sleep() { return true; }
};

// ---------------------------------------------------------------------------------------- serialisation
namespace stub {

struct Writer {
  std::vector<uint8_t> buf;
  void raw(const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; buf.insert(buf.end(), b, b + n); }
  template <class T> void pod(T v) { raw(&v, sizeof(T)); }
};
struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  Reader(const uint8_t* b, size_t n) : p(b), end(b + n) {}
  void raw(void* dst, size_t n) {
    if (!ok || (size_t)(end - p) < n) { ok = false; std::memset(dst, 0, n); return; }
    std::memcpy(dst, p, n);
    p += n;
  }
  template <class T> T pod() { T v; raw(&v, sizeof(T)); return v; }
  size_t left() const { return (size_t)(end - p); }
};

// primitives (little endian hosts only, like the rest of the repo)
#define ROS_STUB_POD(T)                                             \
  inline void ser(Writer& w, const T& v) { w.pod<T>(v); }           \
  inline void de(Reader& r, T& v) { v = r.pod<T>(); }
ROS_STUB_POD(uint8_t) ROS_STUB_POD(int8_t) ROS_STUB_POD(uint16_t) ROS_STUB_POD(int16_t) ROS_STUB_POD(uint32_t)
ROS_STUB_POD(int32_t) ROS_STUB_POD(uint64_t) ROS_STUB_POD(int64_t) ROS_STUB_POD(float) ROS_STUB_POD(double)
#undef ROS_STUB_POD
inline void ser(Writer& w, const bool& v) { w.pod<uint8_t>(v ? 1 : 0); }
inline void de(Reader& r, bool& v) { v = r.pod<uint8_t>() != 0; }
inline void ser(Writer& w, const std::string& s) { w.pod<uint32_t>((uint32_t)s.size()); w.raw(s.data(), s.size()); }
inline void de(Reader& r, std::string& s) {
  const uint32_t n = r.pod<uint32_t>();
  if (!r.ok || r.left() < n) { r.ok = false; s.clear(); return; }
  s.assign((const char*)r.p, n);
  r.p += n;
}
inline void ser(Writer& w, const Time& t) { w.pod<uint32_t>(t.sec); w.pod<uint32_t>(t.nsec); }
inline void de(Reader& r, Time& t) { t.sec = r.pod<uint32_t>(); t.nsec = r.pod<uint32_t>(); }
inline void ser(Writer& w, const Duration& t) { w.pod<int32_t>(t.sec); w.pod<int32_t>(t.nsec); }
inline void de(Reader& r, Duration& t) { t.sec = r.pod<int32_t>(); t.nsec = r.pod<int32_t>(); }

// T[] (u32 count + elements) and T[N] (elements only)
template <class T> void ser(Writer& w, const std::vector<T>& v) {
  w.pod<uint32_t>((uint32_t)v.size());
  for (const T& e : v) ser(w, e);
}
template <class T> void de(Reader& r, std::vector<T>& v) {
  const uint32_t n = r.pod<uint32_t>();
  v.clear();
  if (!r.ok || n > r.left()) { r.ok = false; return; }   // every element takes at least one byte
  v.resize(n);
  for (T& e : v) de(r, e);
}
template <class T, size_t N> struct FixedArray {   // boost::array stand-in
  T elems[N];
  FixedArray() { for (size_t i = 0; i < N; ++i) elems[i] = T(); }
  FixedArray(std::initializer_list<T> il) {
    size_t i = 0;
    for (const T& v : il) if (i < N) elems[i++] = v;
    for (; i < N; ++i) elems[i] = T();
  }
  T& operator[](size_t i) { return elems[i]; }
  const T& operator[](size_t i) const { return elems[i]; }
  T& at(size_t i) { if (i >= N) throw std::out_of_range("FixedArray"); return elems[i]; }
  T* begin() { return elems; }
  T* end() { return elems + N; }
  const T* begin() const { return elems; }
  const T* end() const { return elems + N; }
  T* data() { return elems; }
  const T* data() const { return elems; }
  static constexpr size_t size() { return N; }
};
template <class T, size_t N> void ser(Writer& w, const FixedArray<T, N>& a) { for (size_t i = 0; i < N; ++i) ser(w, a[i]); }
template <class T, size_t N> void de(Reader& r, FixedArray<T, N>& a) { for (size_t i = 0; i < N; ++i) de(r, a[i]); }

// ------------------------------------------------------------------------------------------ the "master"
struct ParamValue {
  enum Kind { STRING, BOOL, DOUBLE, INT, STRINGS } kind = STRING;
  std::string s;
  bool b = false;
  double d = 0;
  int64_t i = 0;
  std::vector<std::string> ss;
};
struct QueuedMsg {
  int64_t deliver_ns;
  int32_t phase;   // 0 = available during set-up (spinOnce), 1 = delivered by spin()
  std::string topic;
  std::vector<uint8_t> bytes;
};
struct TfEntry { std::string target, source; double t[3]; double q[4]; /* x y z w */ };

struct Master {
  std::mutex mu;
  std::map<std::string, ParamValue> params;
  std::vector<TfEntry> tf;
  std::vector<QueuedMsg> queue;
  std::vector<bool> delivered;
  std::multimap<std::string, std::function<void(const std::vector<uint8_t>&)>> subs;
  std::vector<std::pair<std::string, std::vector<uint8_t>>> out;   // published (topic, bytes) in publication order
  std::vector<std::string> log;
  bool ok = true;
  int64_t now_ns = 0;

  static Master& get() { static Master m; return m; }

  static std::string rd_str(std::ifstream& f) {
    uint32_t n = 0;
    f.read((char*)&n, 4);
    std::string s(n, '\0');
    if (n) f.read(&s[0], n);
    return s;
  }
  // scenario file: "RSTB" u32 version | u32 n_params {str name, u8 kind, value} | u32 n_tf {str target, str source,
  // f64 t[3], f64 q[4]} | u32 n_msgs {i64 deliver_ns, i32 phase, str topic, u32 len, bytes}
  bool load(const char* path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    char magic[4];
    uint32_t ver = 0, n = 0;
    f.read(magic, 4);
    f.read((char*)&ver, 4);
    if (std::memcmp(magic, "RSTB", 4) != 0 || ver != 1) return false;
    f.read((char*)&n, 4);
    for (uint32_t i = 0; i < n; ++i) {
      const std::string name = rd_str(f);
      uint8_t kind = 0;
      f.read((char*)&kind, 1);
      ParamValue v;
      v.kind = (ParamValue::Kind)kind;
      if (kind == ParamValue::STRING) v.s = rd_str(f);
      else if (kind == ParamValue::BOOL) { uint8_t b = 0; f.read((char*)&b, 1); v.b = b != 0; }
      else if (kind == ParamValue::DOUBLE) f.read((char*)&v.d, 8);
      else if (kind == ParamValue::INT) f.read((char*)&v.i, 8);
      else { uint32_t m = 0; f.read((char*)&m, 4); for (uint32_t j = 0; j < m; ++j) v.ss.push_back(rd_str(f)); }
      params[name] = v;
    }
    f.read((char*)&n, 4);
    for (uint32_t i = 0; i < n; ++i) {
      TfEntry e;
      e.target = rd_str(f);
      e.source = rd_str(f);
      f.read((char*)e.t, 24);
      f.read((char*)e.q, 32);
      tf.push_back(e);
    }
    f.read((char*)&n, 4);
    for (uint32_t i = 0; i < n; ++i) {
      QueuedMsg q;
      f.read((char*)&q.deliver_ns, 8);
      f.read((char*)&q.phase, 4);
      q.topic = rd_str(f);
      uint32_t len = 0;
      f.read((char*)&len, 4);
      q.bytes.resize(len);
      if (len) f.read((char*)q.bytes.data(), len);
      queue.push_back(std::move(q));
    }
    delivered.assign(queue.size(), false);
    return (bool)f;
  }
  // output file: u32 n {str topic, u32 len, bytes} | u32 n_log {str}
  bool dump(const char* path) {
    std::ofstream f(path, std::ios::binary);
    auto wr_str = [&](const std::string& s) { uint32_t n = (uint32_t)s.size(); f.write((const char*)&n, 4); f.write(s.data(), n); };
    uint32_t n = (uint32_t)out.size();
    f.write((const char*)&n, 4);
    for (auto& o : out) {
      wr_str(o.first);
      uint32_t len = (uint32_t)o.second.size();
      f.write((const char*)&len, 4);
      f.write((const char*)o.second.data(), len);
    }
    n = (uint32_t)log.size();
    f.write((const char*)&n, 4);
    for (auto& l : log) wr_str(l);
    return (bool)f;
  }
  void deliver(int phase) {
    for (size_t i = 0; i < queue.size(); ++i) {
      if (delivered[i] || queue[i].phase != phase) continue;
      auto range = subs.equal_range(queue[i].topic);
      if (range.first == range.second) continue;   // nobody listens (yet)
      delivered[i] = true;
      now_ns = queue[i].deliver_ns;
      std::vector<std::function<void(const std::vector<uint8_t>&)>> cbs;
      for (auto it = range.first; it != range.second; ++it) cbs.push_back(it->second);
      for (auto& cb : cbs) cb(queue[i].bytes);
    }
  }
  void logf(const char* level, const char* fmt, ...) __attribute__((format(printf, 3, 4)));
};

inline void Master::logf(const char* level, const char* fmt, ...) {
  char line[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(line, sizeof line, fmt, ap);
  va_end(ap);
  std::lock_guard<std::mutex> lk(mu);
  log.push_back(std::string(level) + ": " + line);
}

}  // namespace stub

inline Time Time::now() { Time t; t.fromNSec((uint64_t)stub::Master::get().now_ns); return t; }

#define ROS_INFO(...) ::ros::stub::Master::get().logf("INFO", __VA_ARGS__)
#define ROS_WARN(...) ::ros::stub::Master::get().logf("WARN", __VA_ARGS__)
#define ROS_ERROR(...) ::ros::stub::Master::get().logf("ERROR", __VA_ARGS__)
#define ROS_INFO_STREAM(x) do { } while (0)

// ---------------------------------------------------------------------------------------------- node API
inline void init(int&, char**, const std::string&) {}
inline bool ok() { return stub::Master::get().ok; }
inline void shutdown() { stub::Master::get().ok = false; }
inline void spinOnce() { stub::Master::get().deliver(0); }
inline void spin() {   // replay: every queued run-phase message in order, then the node shuts down
  stub::Master::get().deliver(0);
  stub::Master::get().deliver(1);
  stub::Master::get().ok = false;
}

struct TransportHints {
  TransportHints& tcpNoDelay(bool = true) { return *this; }
};
typedef std::shared_ptr<void const> VoidConstPtr;

class Publisher {
 public:
  Publisher() {}
  explicit Publisher(const std::string& t) : topic_(t) {}
  template <class M> void publish(const M& m) const {
    stub::Writer w;
    ser(w, m);
    stub::Master& ms = stub::Master::get();
    std::lock_guard<std::mutex> lk(ms.mu);
    ms.out.emplace_back(topic_, std::move(w.buf));
  }
  const std::string& getTopic() const { return topic_; }
 private:
  std::string topic_;
};

class Subscriber {
 public:
  Subscriber() {}
  explicit Subscriber(const std::string& t) : topic_(t) {}
  void shutdown() {}
  const std::string& getTopic() const { return topic_; }
 private:
  std::string topic_;
};

class NodeHandle {
 public:
  explicit NodeHandle(const std::string& ns = "") : ns_(ns) {}
  std::string resolve(const std::string& name) const { return ns_ == "~" ? "~" + name : name; }

  bool lookup(const std::string& name, stub::ParamValue& v) const {
    stub::Master& m = stub::Master::get();
    auto it = m.params.find(resolve(name));
    if (it == m.params.end()) return false;
    v = it->second;
    return true;
  }
  bool param(const std::string& name, std::string& out, const std::string& def) const {
    stub::ParamValue v;
    if (lookup(name, v) && v.kind == stub::ParamValue::STRING) { out = v.s; return true; }
    out = def; return false;
  }
  bool param(const std::string& name, bool& out, const bool& def) const {
    stub::ParamValue v;
    if (lookup(name, v) && v.kind == stub::ParamValue::BOOL) { out = v.b; return true; }
    out = def; return false;
  }
  bool param(const std::string& name, double& out, const double& def) const {
    stub::ParamValue v;
    if (lookup(name, v) && (v.kind == stub::ParamValue::DOUBLE || v.kind == stub::ParamValue::INT)) {
      out = v.kind == stub::ParamValue::DOUBLE ? v.d : (double)v.i; return true;
    }
    out = def; return false;
  }
  bool param(const std::string& name, int& out, const int& def) const {
    stub::ParamValue v;
    if (lookup(name, v) && v.kind == stub::ParamValue::INT) { out = (int)v.i; return true; }
    out = def; return false;
  }
  bool param(const std::string& name, std::vector<std::string>& out, const std::vector<std::string>& def) const {
    stub::ParamValue v;
    if (lookup(name, v) && v.kind == stub::ParamValue::STRINGS) { out = v.ss; return true; }
    out = def; return false;
  }
  template <class T> bool param(const std::string& name, T& out, const T& def) const {
    return param(name, out, def);
  }

  template <class M> Publisher advertise(const std::string& topic, uint32_t /*queue*/, bool /*latch*/ = false) {
    return Publisher(topic);
  }
  template <class M>
  Subscriber subscribe(const std::string& topic, uint32_t /*queue*/,
                       const std::function<void(const std::shared_ptr<M const>&)>& cb,
                       const VoidConstPtr& = VoidConstPtr(), const TransportHints& = TransportHints()) {
    stub::Master::get().subs.emplace(topic, [cb](const std::vector<uint8_t>& bytes) {
      auto m = std::make_shared<M>();
      stub::Reader r(bytes.data(), bytes.size());
      de(r, *m);
      if (!r.ok) { ROS_ERROR("malformed message dropped"); return; }
      cb(m);
    });
    return Subscriber(topic);
  }
 private:
  std::string ns_;
};

}  // namespace ros
