"""Pure-Python restatement of the reference's frame assembly, TEST INFRASTRUCTURE ONLY (checker for
csrc/frame_assembler.cpp): ApproximateTimeVec (my_message_filters/sync_policies/approximate_time_vec.h:170-217,
262-480, 488-626) + SynchronizerVec::signal (synchronizer_vec.h:147-161) + the worker gating S3D:1029-1057.
roscpp is absent here, so ros::Time/Duration arithmetic (integer ns; Duration*double via toSec/fromSec) is
restated from the roscpp sources' documented behaviour: parity unpinned."""
import math
from collections import deque

NS = 1_000_000_000


def _dur_to_sec(ns):
    sec, nsec = divmod(ns, NS)            # python floors: nsec in [0, 1e9) like roscpp
    return float(sec) + 1e-9 * float(nsec)


def _round_half_away(x):
    return int(math.floor(x + 0.5)) if x >= 0 else int(math.ceil(x - 0.5))


def _dur_from_sec(d):
    sec = math.floor(d)
    return int(sec) * NS + _round_half_away((d - sec) * 1e9)


def _scale(ns, s):
    return _dur_from_sec(_dur_to_sec(ns) * s)


def _time_to_sec(ns):
    return float(ns // NS) + 1e-9 * float(ns % NS)


class RefAssembler:
    def __init__(self, n, queue_size=None, lower_bound_ns=20_000_000, age_penalty=2.0, max_interval_ns=None,
                 max_sync_diff_s=0.067):
        self.n = n
        self.queue_size = queue_size if queue_size is not None else max(3, 1 + n // 4)
        self.lb = [lower_bound_ns] * n
        self.scale = 1.0 + age_penalty
        self.max_interval = max_interval_ns
        self.max_sync = max_sync_diff_s
        self.deques = [deque() for _ in range(n)]
        self.past = [[] for _ in range(n)]
        self.candidate = [None] * n
        self.nne = 0
        self.pivot = None
        self.pivot_time = self.cstart = self.cend = 0
        self.dropped = [False] * n
        self.ready = []
        self.last_stamp = 0.0
        self.stats = dict(emitted=0, skipped_backwards=0, blanked_cameras=0, dropped_messages=0, signalled=0)

    # bookkeeping ATV:262-388
    def _del_front(self, i):
        self.deques[i].popleft()
        if not self.deques[i]:
            self.nne -= 1

    def _to_past(self, i):
        self.past[i].append(self.deques[i].popleft())
        if not self.deques[i]:
            self.nne -= 1

    def _make_candidate(self):
        self.candidate = [d[0] for d in self.deques]
        for p in self.past:
            p.clear()

    def _recover(self, i, count=None):
        k = len(self.past[i]) if count is None else count
        for _ in range(k):
            self.deques[i].appendleft(self.past[i].pop())
        if self.deques[i]:
            self.nne += 1

    def _publish(self):
        self._signal(list(self.candidate))
        self.candidate = [None] * self.n
        self.pivot = None
        self.nne = 0
        for i in range(self.n):
            while self.past[i]:
                self.deques[i].appendleft(self.past[i].pop())
            self.deques[i].popleft()
            if self.deques[i]:
                self.nne += 1

    def _boundary(self, times, end):
        t, idx = times[0], 0
        for i in range(1, self.n):
            if (times[i] < t) ^ end:
                t, idx = times[i], i
        return idx, t

    def _virtual_times(self):
        out = []
        for i in range(self.n):
            if not self.deques[i]:
                out.append(max(self.past[i][-1][0] + self.lb[i], self.pivot_time))
            else:
                out.append(self.deques[i][0][0])
        return out

    def _vboundary(self, end):
        vt = self._virtual_times()
        t, idx = vt[0], 0
        for i in range(self.n):
            if (vt[i] < t) ^ end:
                t, idx = vt[i], i
        return idx, t

    def _process(self):
        while self.nne == self.n:
            heads = [d[0][0] for d in self.deques]
            end_i, end_t = self._boundary(heads, True)
            start_i, start_t = self._boundary(heads, False)
            for i in range(self.n):
                if i != end_i:
                    self.dropped[i] = False
            if self.pivot is None:
                if self.max_interval is not None and end_t - start_t > self.max_interval:
                    self._del_front(start_i)
                    continue
                if self.dropped[end_i]:
                    self._del_front(start_i)
                    continue
                self._make_candidate()
                self.cstart, self.cend, self.pivot, self.pivot_time = start_t, end_t, end_i, end_t
                self._to_past(start_i)
            else:
                if _scale(end_t - self.cend, self.scale) >= start_t - self.cstart:
                    self._to_past(start_i)
                else:
                    self._make_candidate()
                    self.cstart, self.cend = start_t, end_t
                    self._to_past(start_i)
            if start_i == self.pivot:
                self._publish()
            elif _scale(end_t - self.cend, self.scale) >= self.pivot_time - self.cstart:
                self._publish()
            elif self.nne < self.n:
                moves = [0] * self.n
                while True:
                    ve_i, ve_t = self._vboundary(True)
                    vs_i, vs_t = self._vboundary(False)
                    if _scale(ve_t - self.cend, self.scale) >= self.pivot_time - self.cstart:
                        self._publish()
                        break
                    if _scale(ve_t - self.cend, self.scale) < vs_t - self.cstart:
                        self.nne = 0
                        for i in range(self.n):
                            self._recover(i, moves[i])
                        break
                    self._to_past(vs_i)
                    moves[vs_i] += 1

    def add(self, cam, stamp_ns, mid):
        n0 = len(self.ready)
        self.deques[cam].append((stamp_ns, mid))
        if len(self.deques[cam]) == 1:
            self.nne += 1
            if self.nne == self.n:
                self._process()
        if len(self.deques[cam]) + len(self.past[cam]) > self.queue_size:
            self.nne = 0
            for j in range(self.n):
                self._recover(j)
            self.deques[cam].popleft()
            self.dropped[cam] = True
            self.stats["dropped_messages"] += 1
            if self.pivot is not None:
                self.candidate = [None] * self.n
                self.pivot = None
                self._process()
        return len(self.ready) - n0

    # worker gating S3D:1029-1057
    def _signal(self, tup):
        self.stats["signalled"] += 1
        t_max, idx = 0.0, -1
        for i, (st, _) in enumerate(tup):
            t = _time_to_sec(st)
            if t > t_max:
                t_max, idx = t, i
        if idx < 0 or t_max - self.last_stamp <= 0.0:
            self.stats["skipped_backwards"] += 1
            return
        self.last_stamp = t_max
        blank = [(t_max - _time_to_sec(st)) > self.max_sync for st, _ in tup]
        self.stats["blanked_cameras"] += sum(blank)
        self.stats["emitted"] += 1
        self.ready.append(dict(ids=[m for _, m in tup], stamps_ns=[s for s, _ in tup], blank=blank, pivot=idx))


def mailbox_replay(t_ready, busy):
    """The 1-slot latest-wins mailbox between skeletonCallback (S3D:999-1006: store under the mutex, notify) and the
    worker loop (S3D:1017-1025: wait for `updated`, copy, clear) as an event simulation over the store events, written
    independently of csrc/frame_assembler.cpp. Returns (taken flags, start times)."""
    n = len(t_ready)
    taken, start = [0] * n, [-1] * n
    state = {"slot": None, "busy_until": float("-inf")}

    def take(at):
        k = state["slot"]
        state["slot"] = None
        taken[k], start[k] = 1, at
        state["busy_until"] = at + max(busy[k], 0)

    for i in range(n):
        t = t_ready[i]
        # the worker came free strictly before this store and found a frame waiting: it took it then
        if state["slot"] is not None and state["busy_until"] < t:
            take(state["busy_until"])
        state["slot"] = i                      # the store; an unread frame is overwritten
        if state["busy_until"] <= t:           # an idle worker is woken by the notify
            take(t)
    if state["slot"] is not None:              # what is left in the slot when the stream ends is processed last
        take(state["busy_until"])
    return taken, start
