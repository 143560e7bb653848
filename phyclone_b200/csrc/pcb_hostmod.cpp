// _pcbhost: CPython extension with the host-side bookkeeping of the tile store (see tiles.py).
//
// TileBook = memo tables + pending-op lists + dependency levels + flush packing of `tiles.TileStore`,
// moved out of the interpreter because a PG iteration records ~15 k tile ops and resolves ~5 k scores
// (config #2) and each of them was 1-3 us of Python.  Pure host code: no CUDA here, the packed plan it
// produces is executed by `engine.DeviceArena.run_plan` (or by the NumPy executor of the CPU tests).
#define PY_SSIZE_T_CLEAN
#include <Python.h>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <map>
#include <vector>

namespace {

constexpr int JOB_SCAN = 1, JOB_ADD_PRIOR = 2;

struct PendAdd {
  int level, a, b, out;
  bool neg;
};
struct PendJob {
  int level, base, out, slot;
  uint32_t kid_off, n_kids;
};

struct Book {
  PyObject_HEAD
  int64_t top, permanent, capacity;
  int prior_id;
  std::unordered_map<uint64_t, int>* add_memo;
  std::unordered_map<std::string, int>* node_memo;
  std::unordered_map<std::string, int>* score_memo;
  std::vector<int>* level;            // per tile id; 0 = materialised
  std::vector<int>* touched;          // ids whose level is non-zero in the pending batch
  std::vector<PendAdd>* adds;
  std::vector<PendJob>* jobs;
  std::vector<int>* kids;             // flat child ids of the pending jobs
  std::vector<double>* sc_lse;        // per score id (generation-wide)
  std::vector<double>* sc_last;
  std::vector<char>* sc_ready;
  std::vector<int>* pending_scores;   // score ids in slot order of the pending batch
  int64_t n_adds, n_jobs, n_scores, n_hits, n_pairs;
  int64_t n_flushes, n_launches, d2h_bytes, staged_ints;  // of the in-extension flush path (run)
  double run_s;
};

inline int level_of(Book* b, int id) { return (size_t)id < b->level->size() ? (*b->level)[id] : 0; }

inline void set_level(Book* b, int id, int lv) {
  if ((size_t)id >= b->level->size()) b->level->resize((size_t)id + 1024, 0);
  (*b->level)[id] = lv;
  b->touched->push_back(id);
}

// parse a Python sequence of ints (tuple fast path) into `out`; returns false on error
inline bool ints_of(PyObject* seq, std::vector<int>& out) {
  out.clear();
  if (PyTuple_CheckExact(seq)) {
    const Py_ssize_t n = PyTuple_GET_SIZE(seq);
    out.reserve((size_t)n);
    for (Py_ssize_t i = 0; i < n; ++i) {
      const long v = PyLong_AsLong(PyTuple_GET_ITEM(seq, i));
      if (v == -1 && PyErr_Occurred()) return false;
      out.push_back((int)v);
    }
    return true;
  }
  PyObject* fast = PySequence_Fast(seq, "expected a sequence of tile ids");
  if (!fast) return false;
  const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
  for (Py_ssize_t i = 0; i < n; ++i) {
    const long v = PyLong_AsLong(PySequence_Fast_GET_ITEM(fast, i));
    if (v == -1 && PyErr_Occurred()) {
      Py_DECREF(fast);
      return false;
    }
    out.push_back((int)v);
  }
  Py_DECREF(fast);
  return true;
}

thread_local std::vector<int> g_tmp;

int Book_init(Book* self, PyObject* args, PyObject*) {
  long long permanent = 0, capacity = 0;
  int prior_id = -1;
  if (!PyArg_ParseTuple(args, "LL|i", &permanent, &capacity, &prior_id)) return -1;
  self->permanent = self->top = permanent;
  self->capacity = capacity;
  self->prior_id = prior_id;
  self->add_memo = new std::unordered_map<uint64_t, int>();
  self->node_memo = new std::unordered_map<std::string, int>();
  self->score_memo = new std::unordered_map<std::string, int>();
  self->level = new std::vector<int>();
  self->touched = new std::vector<int>();
  self->adds = new std::vector<PendAdd>();
  self->jobs = new std::vector<PendJob>();
  self->kids = new std::vector<int>();
  self->sc_lse = new std::vector<double>();
  self->sc_last = new std::vector<double>();
  self->sc_ready = new std::vector<char>();
  self->pending_scores = new std::vector<int>();
  self->n_adds = self->n_jobs = self->n_scores = self->n_hits = self->n_pairs = 0;
  self->n_flushes = self->n_launches = self->d2h_bytes = self->staged_ints = 0;
  self->run_s = 0.0;
  return 0;
}

void Book_dealloc(Book* self) {
  delete self->add_memo;
  delete self->node_memo;
  delete self->score_memo;
  delete self->level;
  delete self->touched;
  delete self->adds;
  delete self->jobs;
  delete self->kids;
  delete self->sc_lse;
  delete self->sc_last;
  delete self->sc_ready;
  delete self->pending_scores;
  Py_TYPE(self)->tp_free((PyObject*)self);
}

void clear_pending(Book* b) {
  for (int id : *b->touched)
    if ((size_t)id < b->level->size()) (*b->level)[id] = 0;
  b->touched->clear();
  b->adds->clear();
  b->jobs->clear();
  b->kids->clear();
  b->pending_scores->clear();
}

PyObject* Book_reset(Book* self, PyObject*) {
  self->top = self->permanent;
  self->add_memo->clear();
  self->node_memo->clear();
  self->score_memo->clear();
  self->sc_lse->clear();
  self->sc_last->clear();
  self->sc_ready->clear();
  clear_pending(self);
  Py_RETURN_NONE;
}

// every op returns -1 (no exception) when the arena is full: the caller grows it and retries
inline bool full(Book* b) { return b->top >= b->capacity; }

// the three recording ops, callable from C++ (ForestState below) as well as from Python
int book_add(Book* self, int a, int b, bool neg) {
  const uint64_t key = ((uint64_t)(uint32_t)a << 33) | ((uint64_t)(uint32_t)b << 1) | (neg ? 1u : 0u);
  auto it = self->add_memo->find(key);
  if (it != self->add_memo->end()) {
    ++self->n_hits;
    return it->second;
  }
  if (full(self)) return -1;
  const int t = (int)self->top++;
  const int lv = 1 + std::max(level_of(self, a), level_of(self, b));
  set_level(self, t, lv);
  self->adds->push_back(PendAdd{lv, a, b, t, neg});
  (*self->add_memo)[key] = t;
  return t;
}

PyObject* add_impl(Book* self, PyObject* const* args, Py_ssize_t nargs, bool neg) {
  if (nargs != 2) {
    PyErr_SetString(PyExc_TypeError, "add(a, b)");
    return nullptr;
  }
  const long a = PyLong_AsLong(args[0]), b = PyLong_AsLong(args[1]);
  if ((a == -1 || b == -1) && PyErr_Occurred()) return nullptr;
  return PyLong_FromLong(book_add(self, (int)a, (int)b, neg));
}

PyObject* Book_add(Book* self, PyObject* const* args, Py_ssize_t nargs) { return add_impl(self, args, nargs, false); }
PyObject* Book_sub(Book* self, PyObject* const* args, Py_ssize_t nargs) { return add_impl(self, args, nargs, true); }

inline std::string key_of(int base, const std::vector<int>& kids) {
  std::string k;
  k.resize((kids.size() + 1) * sizeof(int));
  std::memcpy(&k[0], &base, sizeof(int));
  if (!kids.empty()) std::memcpy(&k[sizeof(int)], kids.data(), kids.size() * sizeof(int));
  return k;
}

int book_node(Book* self, int base, const std::vector<int>& kids) {
  std::string key = key_of(base, kids);
  auto it = self->node_memo->find(key);
  if (it != self->node_memo->end()) {
    ++self->n_hits;
    return it->second;
  }
  if (full(self)) return -1;
  const int t = (int)self->top++;
  int lv = level_of(self, base);
  for (int k : kids) lv = std::max(lv, level_of(self, k));
  ++lv;
  set_level(self, t, lv);
  self->jobs->push_back(PendJob{lv, base, t, -1, (uint32_t)self->kids->size(), (uint32_t)kids.size()});
  self->kids->insert(self->kids->end(), kids.begin(), kids.end());
  self->node_memo->emplace(std::move(key), t);
  return t;
}

PyObject* Book_node(Book* self, PyObject* const* args, Py_ssize_t nargs) {
  if (nargs != 2) {
    PyErr_SetString(PyExc_TypeError, "node(base, kids)");
    return nullptr;
  }
  const long base = PyLong_AsLong(args[0]);
  if (base == -1 && PyErr_Occurred()) return nullptr;
  std::vector<int>& kids = g_tmp;
  if (!ints_of(args[1], kids)) return nullptr;
  return PyLong_FromLong(book_node(self, (int)base, kids));
}

int book_score(Book* self, const std::vector<int>& kids);

// score(kids) -> score id (resolve with value())
PyObject* Book_score(Book* self, PyObject* arg) {
  std::vector<int>& kids = g_tmp;
  if (!ints_of(arg, kids)) return nullptr;
  return PyLong_FromLong(book_score(self, kids));
}

int book_score(Book* self, const std::vector<int>& kids) {
  std::string key = key_of(-7, kids);
  auto it = self->score_memo->find(key);
  if (it != self->score_memo->end()) {
    ++self->n_hits;
    return it->second;
  }
  const int sid = (int)self->sc_ready->size();
  self->sc_ready->push_back(0);
  self->sc_lse->push_back(0.0);
  self->sc_last->push_back(0.0);
  int lv = 0;
  for (int k : kids) lv = std::max(lv, level_of(self, k));
  ++lv;
  const int slot = (int)self->pending_scores->size();
  self->pending_scores->push_back(sid);
  self->jobs->push_back(PendJob{lv, -1, -1, slot, (uint32_t)self->kids->size(), (uint32_t)kids.size()});
  self->kids->insert(self->kids->end(), kids.begin(), kids.end());
  self->score_memo->emplace(std::move(key), sid);
  return sid;
}

// value(score id) -> (lse_sum, last_sum) or None while the score is still pending
PyObject* Book_value(Book* self, PyObject* arg) {
  const long sid = PyLong_AsLong(arg);
  if (sid == -1 && PyErr_Occurred()) return nullptr;
  if (sid < 0 || (size_t)sid >= self->sc_ready->size()) {
    PyErr_SetString(PyExc_IndexError, "bad score id");
    return nullptr;
  }
  if (!(*self->sc_ready)[sid]) Py_RETURN_NONE;
  return Py_BuildValue("(dd)", (*self->sc_lse)[sid], (*self->sc_last)[sid]);
}

PyObject* Book_has_pending(Book* self, PyObject*) {
  if (self->adds->empty() && self->jobs->empty()) Py_RETURN_FALSE;
  Py_RETURN_TRUE;
}

// pack() -> (packed int32 bytes, plan list, n_slots)
//   plan entries: ("add", offset, n, sign) | ("jobs", offset, n, n_kids, conv_pairs); same layout as
//   engine.DeviceArena.run_plan expects: adds [a | b | out], jobs [n x 8 table | kids]
// packs the pending batch: `packed` = every id list / job table, `raw` = the launch plan as 6 int32 per record
// (what pcb_run_plan_dev reads), `plan` (may be NULL) = the same plan as a Python list for engine.run_plan
bool pack_core(Book* self, std::vector<int32_t>& packed, std::vector<int32_t>& raw, PyObject* plan) {
  int max_level = 0;
  for (const PendAdd& a : *self->adds) max_level = std::max(max_level, a.level);
  for (const PendJob& j : *self->jobs) max_level = std::max(max_level, j.level);
  packed.clear();
  raw.clear();
  packed.reserve(self->adds->size() * 3 + self->jobs->size() * 8 + self->kids->size());
  std::vector<const PendAdd*> sel;
  std::vector<const PendJob*> selj;
  for (int lv = 1; lv <= max_level; ++lv) {
    for (int neg = 0; neg < 2; ++neg) {
      sel.clear();
      for (const PendAdd& a : *self->adds)
        if (a.level == lv && (int)a.neg == neg) sel.push_back(&a);
      if (sel.empty()) continue;
      const size_t off = packed.size(), n = sel.size();
      for (const PendAdd* a : sel) packed.push_back(a->a);
      for (const PendAdd* a : sel) packed.push_back(a->b);
      for (const PendAdd* a : sel) packed.push_back(a->out);
      const int32_t rr[6] = {0, (int32_t)off, (int32_t)n, 0, neg, 0};
      raw.insert(raw.end(), rr, rr + 6);
      if (plan) {
        PyObject* rec = Py_BuildValue("(snnd)", "add", (Py_ssize_t)off, (Py_ssize_t)n, neg ? -1.0 : 1.0);
        if (!rec) return false;
        PyList_Append(plan, rec);
        Py_DECREF(rec);
      }
      self->n_adds += (int64_t)n;
    }
    selj.clear();
    for (const PendJob& j : *self->jobs)
      if (j.level == lv) selj.push_back(&j);
    if (selj.empty()) continue;
    const size_t off = packed.size(), n = selj.size();
    int32_t koff = 0;
    for (const PendJob* j : selj) {
      const int32_t rec[8] = {koff, (int32_t)j->n_kids, j->base, j->out,
                              j->slot >= 0 ? (JOB_SCAN | JOB_ADD_PRIOR) : JOB_SCAN, j->slot, 0, 0};
      packed.insert(packed.end(), rec, rec + 8);
      koff += (int32_t)j->n_kids;
    }
    for (const PendJob* j : selj)
      packed.insert(packed.end(), self->kids->begin() + j->kid_off, self->kids->begin() + j->kid_off + j->n_kids);
    const int32_t rr[6] = {1, (int32_t)off, (int32_t)n, koff, 0, 0};
    raw.insert(raw.end(), rr, rr + 6);
    if (plan) {
      PyObject* rec = Py_BuildValue("(snnnn)", "jobs", (Py_ssize_t)off, (Py_ssize_t)n, (Py_ssize_t)koff,
                                    (Py_ssize_t)(koff - (int32_t)n));
      if (!rec) return false;
      PyList_Append(plan, rec);
      Py_DECREF(rec);
    }
    self->n_jobs += (int64_t)n;
    self->n_pairs += (int64_t)(koff - (int32_t)n);
  }
  return true;
}

PyObject* Book_pack(Book* self, PyObject*) {
  std::vector<int32_t> packed, raw;
  PyObject* plan = PyList_New(0);
  if (!plan) return nullptr;
  if (!pack_core(self, packed, raw, plan)) {
    Py_DECREF(plan);
    return nullptr;
  }
  PyObject* bytes = PyBytes_FromStringAndSize((const char*)packed.data(), (Py_ssize_t)(packed.size() * 4));
  PyObject* raw_bytes = PyBytes_FromStringAndSize((const char*)raw.data(), (Py_ssize_t)(raw.size() * 4));
  PyObject* out = Py_BuildValue("(NNnN)", bytes, plan, (Py_ssize_t)self->pending_scores->size(), raw_bytes);
  return out;
}

void take_scores(Book* self, const double* p) {
  const size_t n = self->pending_scores->size();
  for (size_t s = 0; s < n; ++s) {
    const int sid = (*self->pending_scores)[s];
    (*self->sc_lse)[sid] = p[s];
    (*self->sc_last)[sid] = p[n + s];
    (*self->sc_ready)[sid] = 1;
  }
  self->n_scores += (int64_t)n;
}

// run(run_plan_fn, last_error_fn, arena, host, host_cap_ints, dev, prior, rows_lse, rows_last, slot_cap, sums_dev,
//     sums_host, stream) -> None | (ints needed, slots needed)
//   The whole flush without leaving the extension: pack, copy into the pinned staging buffer, ONE call of
//   pcb_run_plan_dev (through its address in libphyclone_b200.so), read the scores back.  Returns the sizes
//   instead when a buffer is too small or the batch has no score to wait for; the caller then takes the generic
//   path (TileStore.flush), which also (re)allocates.  All addresses are plain integers.
typedef int (*run_plan_fn)(const void*, const int32_t*, int64_t, int32_t*, const int32_t*, int32_t, double, double*, double*,
                           int64_t, double*, double*, void* const*, void*);
typedef const char* (*last_error_fn)(void);

PyObject* Book_run(Book* self, PyObject* args) {
  unsigned long long fn_a, err_a, arena_a, host_a, dev_a, rows_l_a, rows_t_a, sums_dev_a, sums_host_a, stream_a;
  long long host_cap, slot_cap;
  double prior;
  if (!PyArg_ParseTuple(args, "KKKKLKdKKLKKK", &fn_a, &err_a, &arena_a, &host_a, &host_cap, &dev_a, &prior, &rows_l_a,
                        &rows_t_a, &slot_cap, &sums_dev_a, &sums_host_a, &stream_a))
    return nullptr;
  const long long n_ints = (long long)(self->adds->size() * 3 + self->jobs->size() * 8 + self->kids->size());
  const long long n_slots = (long long)self->pending_scores->size();
  if (n_slots == 0 || n_ints > host_cap || n_slots > slot_cap || n_ints == 0)
    return Py_BuildValue("(LL)", n_ints, n_slots);
  std::vector<int32_t> packed, raw;
  if (!pack_core(self, packed, raw, nullptr)) return nullptr;
  std::memcpy((void*)(uintptr_t)host_a, packed.data(), packed.size() * 4);
  const auto t0 = std::chrono::steady_clock::now();
  int rc;
  Py_BEGIN_ALLOW_THREADS
  rc = ((run_plan_fn)(uintptr_t)fn_a)((const void*)(uintptr_t)arena_a, (const int32_t*)(uintptr_t)host_a,
                                      (int64_t)packed.size(), (int32_t*)(uintptr_t)dev_a, raw.data(),
                                      (int32_t)(raw.size() / 6), prior, (double*)(uintptr_t)rows_l_a,
                                      (double*)(uintptr_t)rows_t_a, (int64_t)n_slots, (double*)(uintptr_t)sums_dev_a,
                                      (double*)(uintptr_t)sums_host_a, nullptr, (void*)(uintptr_t)stream_a);
  Py_END_ALLOW_THREADS
  self->run_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (rc != 0) {
    const char* msg = ((last_error_fn)(uintptr_t)err_a)();
    PyErr_Format(PyExc_RuntimeError, "pcb_run_plan_dev failed (%d): %s", rc, msg ? msg : "?");
    return nullptr;
  }
  take_scores(self, (const double*)(uintptr_t)sums_host_a);
  ++self->n_flushes;
  self->n_launches += (int64_t)(raw.size() / 6) + 1;
  self->d2h_bytes += 16 * n_slots;
  self->staged_ints += (int64_t)packed.size();
  clear_pending(self);
  Py_RETURN_NONE;
}

// done(scores or None): scores = contiguous float64 buffer [2, n_slots]; marks the batch as executed
PyObject* Book_done(Book* self, PyObject* arg) {
  const size_t n = self->pending_scores->size();
  if (n) {
    Py_buffer view;
    if (PyObject_GetBuffer(arg, &view, PyBUF_C_CONTIGUOUS) != 0) return nullptr;
    if ((size_t)view.len < 2 * n * sizeof(double)) {
      PyBuffer_Release(&view);
      PyErr_SetString(PyExc_ValueError, "score buffer too small");
      return nullptr;
    }
    take_scores(self, (const double*)view.buf);
    PyBuffer_Release(&view);
  }
  clear_pending(self);
  Py_RETURN_NONE;
}

// reserve(n) -> first id of a block of n consecutive tiles handed to somebody else (the device-resident SMC sweep
// allocates its tiles itself), or -1 if the arena is too small: the caller grows it and asks again
PyObject* Book_reserve(Book* self, PyObject* arg) {
  const long long n = PyLong_AsLongLong(arg);
  if (n == -1 && PyErr_Occurred()) return nullptr;
  if (n < 0 || self->top + n > self->capacity) return PyLong_FromLong(-1);
  const long long first = self->top;
  self->top += n;
  return PyLong_FromLongLong(first);
}

PyObject* Book_set_capacity(Book* self, PyObject* arg) {
  const long long c = PyLong_AsLongLong(arg);
  if (c == -1 && PyErr_Occurred()) return nullptr;
  self->capacity = c;
  Py_RETURN_NONE;
}

PyObject* Book_counters(Book* self, PyObject*) {
  return Py_BuildValue("{s:L,s:L,s:L,s:L,s:L,s:L,s:L,s:L,s:L,s:L,s:d}", "adds", (long long)self->n_adds, "node_jobs",
                       (long long)self->n_jobs, "scores", (long long)self->n_scores, "memo_hits",
                       (long long)self->n_hits, "conv_pairs", (long long)self->n_pairs, "top", (long long)self->top,
                       "fast_flushes", (long long)self->n_flushes, "fast_launches", (long long)self->n_launches,
                       "fast_d2h_bytes", (long long)self->d2h_bytes, "fast_staged_ints", (long long)self->staged_ints,
                       "fast_run_s", self->run_s);
}

PyMethodDef Book_methods[] = {
    {"reset", (PyCFunction)Book_reset, METH_NOARGS, "start a new arena generation"},
    {"add", (PyCFunction)(void (*)(void))Book_add, METH_FASTCALL, "tile(a) + tile(b) -> id (or -1 if the arena is full)"},
    {"sub", (PyCFunction)(void (*)(void))Book_sub, METH_FASTCALL, "tile(a) - tile(b) -> id (or -1 if the arena is full)"},
    {"node", (PyCFunction)(void (*)(void))Book_node, METH_FASTCALL, "tile(base) + log_S(kids) -> id (or -1)"},
    {"score", (PyCFunction)Book_score, METH_O, "score handle of a dummy root folding kids"},
    {"value", (PyCFunction)Book_value, METH_O, "(lse_sum, last_sum) or None while pending"},
    {"has_pending", (PyCFunction)Book_has_pending, METH_NOARGS, ""},
    {"run", (PyCFunction)Book_run, METH_VARARGS, "flush the pending batch through pcb_run_plan_dev without leaving the extension"},
    {"pack", (PyCFunction)Book_pack, METH_NOARGS, "(packed bytes, plan, n_slots, plan as int32 x 6 records) of the pending batch"},
    {"done", (PyCFunction)Book_done, METH_O, "hand back the [2, n_slots] scores of the packed batch"},
    {"set_capacity", (PyCFunction)Book_set_capacity, METH_O, ""},
    {"reserve", (PyCFunction)Book_reserve, METH_O, "first id of n consecutive tiles set aside (or -1 if the arena is full)"},
    {"counters", (PyCFunction)Book_counters, METH_NOARGS, ""},
    {nullptr, nullptr, 0, nullptr}};

PyTypeObject BookType = {PyVarObject_HEAD_INIT(nullptr, 0)};

// =====================================================================================================
// ForestState: root-level snapshot of an SMC forest (see smc/forest_state.py for the semantics).
// Immutable; extensions create new objects.  All tile ops go straight into a TileBook.
// =====================================================================================================
#include <cmath>

std::vector<double> g_log;    // g_log[n] = log(n) exactly as NumPy computes it (set from Python)
std::vector<double> g_rterm;  // FS-CRP root-count term per number of roots (tree/distributions.py:103-133)

inline double lf(long n) { return std::lgamma((double)n + 1.0); }
inline double tlog(long n) { return (size_t)n < g_log.size() ? g_log[(size_t)n] : std::log((double)n); }

inline uint64_t mix64(uint64_t x) {
  x ^= x >> 30;
  x *= 0xbf58476d1ce4e5b9ULL;
  x ^= x >> 27;
  x *= 0x94d049bb133111ebULL;
  x ^= x >> 31;
  return x;
}
inline uint64_t edge_hash(long e, long s, long d) {
  return mix64(mix64(mix64((uint64_t)e + 0x9e3779b97f4a7c15ULL) ^ (uint64_t)s) ^ (uint64_t)d);
}
inline uint64_t data_hash(long node, long pos, long idx) {
  return mix64(mix64(mix64((uint64_t)node + 0x632be59bd9b4e019ULL) ^ (uint64_t)pos) ^ (uint64_t)idx) ^ 0x5bd1e995ULL;
}

enum { K_ATTACH = 0, K_NEW = 1, K_OUTLIER = 2 };

struct FState {
  PyObject_HEAD
  std::vector<long>* roots;
  std::vector<int>*lr, *lp, *ndata, *size, *slot, *gidx;
  std::vector<std::vector<int>>* kids;
  long n_nodes, n_slots, n_edges, n_out, last_added;
  int has_last;
  double crp, mult, oprior, omarg;
  uint64_t skey;
  PyObject* roots_tuple;
};

extern PyTypeObject FStateType;

FState* fstate_alloc() {
  FState* s = PyObject_New(FState, &FStateType);
  if (!s) return nullptr;
  s->roots = new std::vector<long>();
  s->lr = new std::vector<int>();
  s->lp = new std::vector<int>();
  s->ndata = new std::vector<int>();
  s->size = new std::vector<int>();
  s->slot = new std::vector<int>();
  s->gidx = new std::vector<int>();
  s->kids = new std::vector<std::vector<int>>();
  s->n_nodes = 0;
  s->n_slots = 1;
  s->n_edges = s->n_out = 0;
  s->last_added = 0;
  s->has_last = 0;
  s->crp = s->mult = s->oprior = s->omarg = 0.0;
  s->skey = 0;
  s->roots_tuple = nullptr;
  return s;
}

void FState_dealloc(FState* s) {
  delete s->roots;
  delete s->lr;
  delete s->lp;
  delete s->ndata;
  delete s->size;
  delete s->slot;
  delete s->gidx;
  delete s->kids;
  Py_XDECREF(s->roots_tuple);
  PyObject_Del(s);
}

PyObject* roots_tuple_of(FState* s) {
  if (!s->roots_tuple) {
    PyObject* t = PyTuple_New((Py_ssize_t)s->roots->size());
    if (!t) return nullptr;
    for (size_t i = 0; i < s->roots->size(); ++i) PyTuple_SET_ITEM(t, (Py_ssize_t)i, PyLong_FromLong((*s->roots)[i]));
    s->roots_tuple = t;
  }
  Py_INCREF(s->roots_tuple);
  return s->roots_tuple;
}

inline int index_of(FState* s, long node) {
  for (size_t j = 0; j < s->roots->size(); ++j)
    if ((*s->roots)[j] == node) return (int)j;
  return -1;
}

struct DataPt {
  int idx, value_id, single_id;
  double op, opn, marg;
};
inline bool parse_dp(PyObject* o, DataPt& d) {
  return PyArg_ParseTuple(o, "iiiddd", &d.idx, &d.value_id, &d.single_id, &d.op, &d.opn, &d.marg) != 0;
}

inline bool longs_of(PyObject* seq, std::vector<long>& out) {
  out.clear();
  PyObject* it = PyObject_GetIter(seq);
  if (!it) return false;
  while (PyObject* item = PyIter_Next(it)) {
    const long v = PyLong_AsLong(item);
    Py_DECREF(item);
    if (v == -1 && PyErr_Occurred()) {
      Py_DECREF(it);
      return false;
    }
    out.push_back(v);
  }
  Py_DECREF(it);
  return !PyErr_Occurred();
}

// (log_p prior, log_p_one prior) from running statistics: tree/distributions.py:39-133, same operation
// order as smc/forest_state.py::_priors
inline void priors(double log_alpha, long n, double crp, double mult, const int* sizes, size_t n_sizes, double oprior,
                   double omarg, double& lp, double& lp1) {
  const double base = (double)n * log_alpha + crp;
  lp = base - (double)(n - 1) * tlog(n + 1) - mult;
  double ways = 0.0;
  for (size_t i = 0; i < n_sizes; ++i) ways += (double)(sizes[i] - 1) * tlog(sizes[i]);
  lp1 = base + (-ways + g_rterm[n_sizes]) - mult;
  const double extra = oprior + omarg;
  lp += extra;
  lp1 += extra;
}

// forest_state(roots, lr, lp, kids, ndata, size, slot, gidx, n_nodes, n_slots, n_edges, crp, mult, oprior, omarg,
//              n_out, skey, last_added | None)
PyObject* mod_forest_state(PyObject*, PyObject* args) {
  PyObject *roots, *lr, *lp, *kids, *ndata, *size, *slot, *gidx, *last;
  long n_nodes, n_slots, n_edges, n_out;
  double crp, mult, oprior, omarg;
  unsigned long long skey;
  if (!PyArg_ParseTuple(args, "OOOOOOOOlllddddlKO", &roots, &lr, &lp, &kids, &ndata, &size, &slot, &gidx, &n_nodes, &n_slots,
                        &n_edges, &crp, &mult, &oprior, &omarg, &n_out, &skey, &last))
    return nullptr;
  FState* s = fstate_alloc();
  if (!s) return nullptr;
  bool ok = longs_of(roots, *s->roots) && ints_of(lr, *s->lr) && ints_of(lp, *s->lp) && ints_of(ndata, *s->ndata) &&
            ints_of(size, *s->size) && ints_of(slot, *s->slot) && ints_of(gidx, *s->gidx);
  if (ok) {
    PyObject* it = PyObject_GetIter(kids);
    ok = it != nullptr;
    if (ok) {
      while (PyObject* item = PyIter_Next(it)) {
        std::vector<int> k;
        const bool good = ints_of(item, k);
        Py_DECREF(item);
        if (!good) {
          ok = false;
          break;
        }
        s->kids->push_back(std::move(k));
      }
      Py_DECREF(it);
    }
  }
  if (!ok) {
    Py_DECREF(s);
    return nullptr;
  }
  s->n_nodes = n_nodes;
  s->n_slots = n_slots;
  s->n_edges = n_edges;
  s->n_out = n_out;
  s->crp = crp;
  s->mult = mult;
  s->oprior = oprior;
  s->omarg = omarg;
  s->skey = skey;
  if (last != Py_None) {
    s->last_added = PyLong_AsLong(last);
    s->has_last = 1;
  }
  return (PyObject*)s;
}

// extend(book, kind, dp, arg) -> thawed successor state, or None if the arena is full
// status: 0 = ok, 1 = arena full (nothing allocated; grow and call again), -1 = Python error set
FState* fstate_extend(FState* self, Book* book, int kind, const DataPt& dp, PyObject* arg, int* status);

PyObject* FState_extend(FState* self, PyObject* args) {
  Book* book;
  int kind;
  PyObject *dpo, *arg;
  if (!PyArg_ParseTuple(args, "O!iOO", &BookType, &book, &kind, &dpo, &arg)) return nullptr;
  DataPt dp;
  if (!parse_dp(dpo, dp)) return nullptr;
  int status = 0;
  FState* s = fstate_extend(self, book, kind, dp, arg, &status);
  if (status == 1) Py_RETURN_NONE;
  return (PyObject*)s;
}

FState* fstate_extend(FState* self, Book* book, int kind, const DataPt& dp, PyObject* arg, int* status) {
  *status = -1;
  const size_t R = self->roots->size();
  long node = 0;
  std::vector<long> taken_names;
  int lp_new = -1, lr_new = -1, j_at = -1;
  std::vector<int> take, keep, nk;
  // record the tile ops first: nothing is allocated if the arena turns out to be full
  if (kind == K_ATTACH) {
    node = PyLong_AsLong(arg);
    if (node == -1 && PyErr_Occurred()) return nullptr;
    j_at = index_of(self, node);
    if (j_at < 0) {
      PyErr_SetString(PyExc_KeyError, "node is not a root of this state");
      return nullptr;
    }
    lp_new = (*self->lp)[j_at] == book->prior_id ? dp.single_id : book_add(book, (*self->lp)[j_at], dp.value_id, false);
    if (lp_new < 0) {
      *status = 1;
      return nullptr;
    }
    lr_new = (*self->kids)[j_at].empty() ? lp_new : book_node(book, lp_new, (*self->kids)[j_at]);
    if (lr_new < 0) {
      *status = 1;
      return nullptr;
    }
  } else if (kind == K_NEW) {
    if (!longs_of(arg, taken_names)) return nullptr;
    for (size_t j = 0; j < R; ++j) {
      bool in = false;
      for (long t : taken_names) in |= (t == (*self->roots)[j]);
      (in ? take : keep).push_back((int)j);
    }
    for (int j : take) nk.push_back((*self->lr)[j]);  // descending slot order = thawed child order
    lp_new = dp.single_id;
    lr_new = nk.empty() ? lp_new : book_node(book, lp_new, nk);
    if (lr_new < 0) {
      *status = 1;
      return nullptr;
    }
  }
  FState* s = fstate_alloc();
  if (!s) return nullptr;
  s->n_out = self->n_out;
  s->omarg = self->omarg;
  s->oprior = self->oprior;
  s->has_last = 1;
  if (kind == K_OUTLIER) {
    *s->roots = *self->roots;
    *s->lr = *self->lr;
    *s->lp = *self->lp;
    *s->kids = *self->kids;
    *s->ndata = *self->ndata;
    *s->size = *self->size;
    *s->slot = *self->slot;
    *s->gidx = *self->gidx;
    s->n_nodes = self->n_nodes;
    s->n_slots = self->n_slots;
    s->n_edges = self->n_edges;
    s->crp = self->crp;
    s->mult = self->mult;
    s->skey = self->skey ^ data_hash(-1, self->n_out, dp.idx);
    s->n_out = self->n_out + 1;
    s->omarg = self->omarg + dp.marg;
    if (dp.op != 0.0) s->oprior = self->oprior + dp.op;
    s->last_added = -1;
    {
    *status = 0;
    return s;
  }
  }
  if (dp.op != 0.0) s->oprior = self->oprior + dp.opn;
  if (kind == K_ATTACH) {
    const int n = (*self->ndata)[j_at];
    *s->roots = *self->roots;
    *s->kids = *self->kids;
    *s->size = *self->size;
    *s->slot = *self->slot;
    *s->gidx = *self->gidx;
    *s->lp = *self->lp;
    *s->lr = *self->lr;
    *s->ndata = *self->ndata;
    (*s->lp)[j_at] = lp_new;
    (*s->lr)[j_at] = lr_new;
    (*s->ndata)[j_at] = n + 1;
    s->n_nodes = self->n_nodes;
    s->n_slots = self->n_slots;
    s->n_edges = self->n_edges;
    s->mult = self->mult;
    s->crp = self->crp + (n >= 1 ? lf(n) - lf(n - 1) : 0.0);
    s->skey = self->skey ^ data_hash(node, n, dp.idx);
    s->last_added = node;
    {
    *status = 0;
    return s;
  }
  }
  const long name = self->n_nodes, gi = self->n_slots, root_gi = 0;
  s->roots->push_back(name);
  s->lr->push_back(lr_new);
  s->lp->push_back(lp_new);
  s->kids->push_back(nk);
  s->ndata->push_back(1);
  int sz = 1;
  for (int j : take) sz += (*self->size)[j];
  s->size->push_back(sz);
  s->slot->push_back((int)self->n_edges);
  s->gidx->push_back((int)gi);
  for (int j : keep) {
    s->roots->push_back((*self->roots)[j]);
    s->lr->push_back((*self->lr)[j]);
    s->lp->push_back((*self->lp)[j]);
    s->kids->push_back((*self->kids)[j]);
    s->ndata->push_back((*self->ndata)[j]);
    s->size->push_back((*self->size)[j]);
    s->slot->push_back((*self->slot)[j]);
    s->gidx->push_back((*self->gidx)[j]);
  }
  s->n_nodes = self->n_nodes + 1;
  s->n_slots = self->n_slots + 1;
  s->n_edges = self->n_edges + 1;
  const long k = (long)take.size();
  s->mult = self->mult - lf((long)R) + lf((long)R - k + 1) + lf(k);
  s->crp = self->crp;
  uint64_t key = self->skey ^ edge_hash(self->n_edges, root_gi, gi) ^ data_hash(name, 0, dp.idx);
  for (int j : take) {
    const long e = (*self->slot)[j], c = (*self->gidx)[j];
    key ^= edge_hash(e, root_gi, c) ^ edge_hash(e, gi, c);
  }
  s->skey = key;
  s->last_added = name;
  {
    *status = 0;
    return s;
  }
}

// extend_built(book, kind, dp, arg) -> AS-BUILT successor state, or None if the arena is full.
// The state of the tree the reference builds by editing in place (no from_dict round trip), as the retained path of a
// conditional sweep does (smc/samplers/conditional.py:19-74): `arg` = node name (ATTACH) or the children in ADOPTION
// order (NEW).  Differences from `extend` (the re-read tree): a new node's children fold in reversed adoption order
// (every adopted edge goes to the head of the adjacency list), the node is created empty and then takes the data point,
// and a node that takes a data point gets `log_r += value` instead of a re-fold (tree/tree_node.py:42-52).
FState* fstate_extend_built(FState* self, Book* book, int kind, const DataPt& dp, PyObject* arg, int* status) {
  *status = -1;
  if (kind == K_OUTLIER) return fstate_extend(self, book, kind, dp, arg, status);
  const size_t R = self->roots->size();
  if (kind == K_ATTACH) {
    const long node = PyLong_AsLong(arg);
    if (node == -1 && PyErr_Occurred()) return nullptr;
    const int j = index_of(self, node);
    if (j < 0) {
      PyErr_SetString(PyExc_KeyError, "node is not a root of this state");
      return nullptr;
    }
    const int lp_old = (*self->lp)[j], lr_old = (*self->lr)[j];
    const int lp_new = lp_old == book->prior_id ? dp.single_id : book_add(book, lp_old, dp.value_id, false);
    if (lp_new < 0) {
      *status = 1;
      return nullptr;
    }
    const int lr_new = lr_old == lp_old ? lp_new : book_add(book, lr_old, dp.value_id, false);
    if (lr_new < 0) {
      *status = 1;
      return nullptr;
    }
    // everything but the two tiles is what `extend` computes
    FState* s = fstate_alloc();
    if (!s) return nullptr;
    const int n = (*self->ndata)[j];
    *s->roots = *self->roots;
    *s->kids = *self->kids;
    *s->size = *self->size;
    *s->slot = *self->slot;
    *s->gidx = *self->gidx;
    *s->lp = *self->lp;
    *s->lr = *self->lr;
    *s->ndata = *self->ndata;
    (*s->lp)[j] = lp_new;
    (*s->lr)[j] = lr_new;
    (*s->ndata)[j] = n + 1;
    s->n_nodes = self->n_nodes;
    s->n_slots = self->n_slots;
    s->n_edges = self->n_edges;
    s->n_out = self->n_out;
    s->omarg = self->omarg;
    s->oprior = (dp.op != 0.0) ? self->oprior + dp.opn : self->oprior;
    s->mult = self->mult;
    s->crp = self->crp + (n >= 1 ? lf(n) - lf(n - 1) : 0.0);
    s->skey = self->skey ^ data_hash(node, n, dp.idx);
    s->last_added = node;
    s->has_last = 1;
    *status = 0;
    return s;
  }
  // NEW
  std::vector<long> adopted;
  if (!longs_of(arg, adopted)) return nullptr;
  std::vector<int> order;
  for (long c : adopted) {
    const int j = index_of(self, c);
    if (j < 0) {
      PyErr_SetString(PyExc_KeyError, "child is not a root of this state");
      return nullptr;
    }
    order.push_back(j);
  }
  std::vector<int> nk;
  for (size_t i = order.size(); i-- > 0;) nk.push_back((*self->lr)[order[i]]);
  // the node is created without data (log_p = the prior tile, log_r = log_p + log_S(children) or log_p itself) ...
  const int lr_empty = nk.empty() ? book->prior_id : book_node(book, book->prior_id, nk);
  if (lr_empty < 0) {
    *status = 1;
    return nullptr;
  }
  // ... and then takes the point: log_p becomes the singleton tile, log_r += value (its own tile even for a leaf)
  const int lr_new = book_add(book, lr_empty, dp.value_id, false);
  if (lr_new < 0) {
    *status = 1;
    return nullptr;
  }
  std::vector<char> taken(R, 0);
  for (int j : order) taken[(size_t)j] = 1;
  FState* s = fstate_alloc();
  if (!s) return nullptr;
  const long name = self->n_nodes, gi = self->n_slots, root_gi = 0;
  int sz = 1;
  for (int j : order) sz += (*self->size)[j];
  s->roots->push_back(name);
  s->lr->push_back(lr_new);
  s->lp->push_back(dp.single_id);
  s->kids->push_back(nk);
  s->ndata->push_back(1);
  s->size->push_back(sz);
  s->slot->push_back((int)self->n_edges);
  s->gidx->push_back((int)gi);
  for (size_t j = 0; j < R; ++j)
    if (!taken[j]) {
      s->roots->push_back((*self->roots)[j]);
      s->lr->push_back((*self->lr)[j]);
      s->lp->push_back((*self->lp)[j]);
      s->kids->push_back((*self->kids)[j]);
      s->ndata->push_back((*self->ndata)[j]);
      s->size->push_back((*self->size)[j]);
      s->slot->push_back((*self->slot)[j]);
      s->gidx->push_back((*self->gidx)[j]);
    }
  const long k = (long)order.size();
  s->n_nodes = self->n_nodes + 1;
  s->n_slots = self->n_slots + 1;
  s->n_edges = self->n_edges + 1;
  s->n_out = self->n_out;
  s->omarg = self->omarg;
  s->oprior = (dp.op != 0.0) ? self->oprior + dp.opn : self->oprior;
  s->mult = self->mult - lf((long)R) + lf((long)R - k + 1) + lf(k);
  s->crp = self->crp;
  uint64_t key = self->skey ^ edge_hash(self->n_edges, root_gi, gi) ^ data_hash(name, 0, dp.idx);
  for (int j : order) {
    const long e = (*self->slot)[j], c = (*self->gidx)[j];
    key ^= edge_hash(e, root_gi, c) ^ edge_hash(e, gi, c);
  }
  s->skey = key;
  s->last_added = name;
  s->has_last = 1;
  *status = 0;
  return s;
}

PyObject* FState_extend_built(FState* self, PyObject* args) {
  Book* book;
  int kind;
  PyObject *dpo, *arg;
  if (!PyArg_ParseTuple(args, "O!iOO", &BookType, &book, &kind, &dpo, &arg)) return nullptr;
  DataPt dp;
  if (!parse_dp(dpo, dp)) return nullptr;
  int status = 0;
  FState* s = fstate_extend_built(self, book, kind, dp, arg, &status);
  if (status == 1) Py_RETURN_NONE;
  return (PyObject*)s;
}

// score_self(book, log_alpha) -> (prior_p, prior_one, sid or -1): the state's own tree score
// (tree/distributions.py:186-206), the data terms as a lazy score handle
PyObject* FState_score_self(FState* self, PyObject* args) {
  Book* book;
  double log_alpha;
  if (!PyArg_ParseTuple(args, "O!d", &BookType, &book, &log_alpha)) return nullptr;
  double lp, lp1;
  priors(log_alpha, self->n_nodes, self->crp, self->mult, self->size->data(), self->size->size(), self->oprior,
         self->omarg, lp, lp1);
  const int sid = self->lr->empty() ? -1 : book_score(book, *self->lr);
  return Py_BuildValue("(ddi)", lp, lp1, sid);
}

// score_attach_all(book, dp, nodes, log_alpha) -> [(prior_p, prior_one, score id, n children of node)] | None
PyObject* FState_score_attach_all(FState* self, PyObject* args) {
  Book* book;
  PyObject *dpo, *nodes_o;
  double log_alpha;
  if (!PyArg_ParseTuple(args, "O!OOd", &BookType, &book, &dpo, &nodes_o, &log_alpha)) return nullptr;
  DataPt dp;
  if (!parse_dp(dpo, dp)) return nullptr;
  std::vector<long> nodes;
  if (!longs_of(nodes_o, nodes)) return nullptr;
  double oprior = self->oprior;
  if (dp.op != 0.0) oprior = oprior + dp.opn;
  const long n = self->n_nodes;
  const double head = (double)n * log_alpha;
  double ways = 0.0;
  for (int z : *self->size) ways += (double)(z - 1) * tlog(z);
  const double tail_p = -(double)(n - 1) * tlog(n + 1) - self->mult + (oprior + self->omarg);
  const double tail_one = (-ways + g_rterm[self->size->size()]) - self->mult + (oprior + self->omarg);
  std::vector<int> kids_t;
  PyObject* out = PyList_New(0);
  if (!out) return nullptr;
  for (long node : nodes) {
    const int j = index_of(self, node);
    if (j < 0) {
      Py_DECREF(out);
      PyErr_SetString(PyExc_KeyError, "node is not a root of the base state");
      return nullptr;
    }
    const int m = (*self->ndata)[j];
    const double crp = head + (self->crp + (m >= 1 ? lf(m) - lf(m - 1) : 0.0));
    const int lr = book_add(book, (*self->lr)[j], dp.value_id, false);  // log_r += value (tree/tree_node.py:47-52)
    if (lr < 0) {
      Py_DECREF(out);
      Py_RETURN_NONE;
    }
    kids_t = *self->lr;
    kids_t[j] = lr;
    const int sid = book_score(book, kids_t);
    PyObject* rec = Py_BuildValue("(ddin)", crp + tail_p, crp + tail_one, sid, (Py_ssize_t)(*self->kids)[j].size());
    PyList_Append(out, rec);
    Py_DECREF(rec);
  }
  return out;
}

// score_new(book, dp, children in adoption order, log_alpha) -> (prior_p, prior_one, sid, roots tuple, k) | None
PyObject* FState_score_new(FState* self, PyObject* args) {
  Book* book;
  PyObject *dpo, *children_o;
  double log_alpha;
  if (!PyArg_ParseTuple(args, "O!OOd", &BookType, &book, &dpo, &children_o, &log_alpha)) return nullptr;
  DataPt dp;
  if (!parse_dp(dpo, dp)) return nullptr;
  std::vector<long> children;
  if (!longs_of(children_o, children)) return nullptr;
  double oprior = self->oprior;
  if (dp.op != 0.0) oprior = oprior + dp.opn;
  const size_t R = self->roots->size();
  std::vector<int> order;
  for (long c : children) {
    const int j = index_of(self, c);
    if (j < 0) {
      PyErr_SetString(PyExc_KeyError, "child is not a root of the base state");
      return nullptr;
    }
    order.push_back(j);
  }
  // children adopted in the given order, each new edge at the HEAD of the adjacency list: fold order is reversed
  std::vector<int> kids_t;
  for (size_t i = order.size(); i-- > 0;) kids_t.push_back((*self->lr)[order[i]]);
  const int lp_t = dp.single_id;
  const int lr = kids_t.empty() ? lp_t : book_node(book, lp_t, kids_t);
  if (lr < 0) Py_RETURN_NONE;
  std::vector<char> taken(R, 0);
  for (int j : order) taken[(size_t)j] = 1;
  const long k = (long)order.size();
  const double mult = self->mult - lf((long)R) + lf((long)R - k + 1) + lf(k);
  std::vector<int> sizes, root_kids;
  int sz = 1;
  for (int j : order) sz += (*self->size)[j];
  sizes.push_back(sz);
  root_kids.push_back(lr);
  PyObject* roots = PyTuple_New((Py_ssize_t)(R - (size_t)k + 1));
  if (!roots) return nullptr;
  PyTuple_SET_ITEM(roots, 0, PyLong_FromLong(self->n_nodes));
  Py_ssize_t pos = 1;
  for (size_t j = 0; j < R; ++j)
    if (!taken[j]) {
      sizes.push_back((*self->size)[j]);
      root_kids.push_back((*self->lr)[j]);
      PyTuple_SET_ITEM(roots, pos++, PyLong_FromLong((*self->roots)[j]));
    }
  double lp, lp1;
  priors(log_alpha, self->n_nodes + 1, self->crp, mult, sizes.data(), sizes.size(), oprior, self->omarg, lp, lp1);
  const int sid = book_score(book, root_kids);
  return Py_BuildValue("(ddiNn)", lp, lp1, sid, roots, (Py_ssize_t)k);
}

// score_outlier(book, dp, log_alpha) -> (prior_p, prior_one, sid or -1, roots tuple)
PyObject* FState_score_outlier(FState* self, PyObject* args) {
  Book* book;
  PyObject* dpo;
  double log_alpha;
  if (!PyArg_ParseTuple(args, "O!Od", &BookType, &book, &dpo, &log_alpha)) return nullptr;
  DataPt dp;
  if (!parse_dp(dpo, dp)) return nullptr;
  double oprior = self->oprior;
  if (dp.op != 0.0) oprior = oprior + dp.op;
  const double omarg = self->omarg + dp.marg;
  double lp, lp1;
  priors(log_alpha, self->n_nodes, self->crp, self->mult, self->size->data(), self->size->size(), oprior, omarg, lp, lp1);
  const int sid = self->lr->empty() ? -1 : book_score(book, *self->lr);
  PyObject* roots = roots_tuple_of(self);
  if (!roots) return nullptr;
  return Py_BuildValue("(ddiN)", lp, lp1, sid, roots);
}

PyObject* FState_get_roots(FState* self, void*) { return roots_tuple_of(self); }
PyObject* FState_get_n_nodes(FState* self, void*) { return PyLong_FromLong(self->n_nodes); }
PyObject* FState_get_skey(FState* self, void*) { return PyLong_FromUnsignedLongLong(self->skey); }
PyObject* FState_get_last(FState* self, void*) {
  if (!self->has_last) Py_RETURN_NONE;
  return PyLong_FromLong(self->last_added);
}
PyObject* FState_get_lr(FState* self, void*) {
  PyObject* t = PyTuple_New((Py_ssize_t)self->lr->size());
  if (!t) return nullptr;
  for (size_t i = 0; i < self->lr->size(); ++i) PyTuple_SET_ITEM(t, (Py_ssize_t)i, PyLong_FromLong((*self->lr)[i]));
  return t;
}

PyObject* ints_tuple(const std::vector<int>& v) {
  PyObject* t = PyTuple_New((Py_ssize_t)v.size());
  if (!t) return nullptr;
  for (size_t i = 0; i < v.size(); ++i) PyTuple_SET_ITEM(t, (Py_ssize_t)i, PyLong_FromLong(v[i]));
  return t;
}

// export() -> (roots, lr, lp, kids, ndata, size, slot, gidx, n_nodes, n_slots, n_edges, n_out, crp, mult, oprior, omarg,
//              skey): everything the device-resident sweep needs to take the state over (smc/device_sweep.py)
PyObject* FState_export(FState* self, PyObject*) {
  PyObject* roots = roots_tuple_of(self);
  PyObject* kids = PyTuple_New((Py_ssize_t)self->kids->size());
  if (!roots || !kids) {
    Py_XDECREF(roots);
    Py_XDECREF(kids);
    return nullptr;
  }
  for (size_t i = 0; i < self->kids->size(); ++i) PyTuple_SET_ITEM(kids, (Py_ssize_t)i, ints_tuple((*self->kids)[i]));
  return Py_BuildValue("(NNNNNNNNlllldddd K)", roots, ints_tuple(*self->lr), ints_tuple(*self->lp), kids,
                       ints_tuple(*self->ndata), ints_tuple(*self->size), ints_tuple(*self->slot), ints_tuple(*self->gidx),
                       self->n_nodes, self->n_slots, self->n_edges, self->n_out, self->crp, self->mult, self->oprior,
                       self->omarg, (unsigned long long)self->skey);
}

// pack_states(thawed states, as-built states, rmax) -> (th_int, ba_int, th_dbl, ba_dbl, kids) as bytes:
//   the record layout of pcb_sweep_desc (include/phyclone_b200.h): int record [5 + 9*rmax padded to 4] = R, n_nodes,
//   n_slots, n_edges, n_out, then rmax-strided name, lr, lp, ndata, size, slot, gidx, kid offset, kid count; double
//   record [8] = crp, mult, oprior, omarg, 0...; kids = the child tile lists the kid offsets index.
PyObject* mod_pack_states(PyObject*, PyObject* args) {
  PyObject *th_o, *ba_o;
  int rmax;
  if (!PyArg_ParseTuple(args, "OOi", &th_o, &ba_o, &rmax)) return nullptr;
  PyObject* lists[2] = {PySequence_Fast(th_o, "expected a sequence of states"), nullptr};
  if (!lists[0]) return nullptr;
  lists[1] = PySequence_Fast(ba_o, "expected a sequence of states");
  if (!lists[1]) {
    Py_DECREF(lists[0]);
    return nullptr;
  }
  const Py_ssize_t T = PySequence_Fast_GET_SIZE(lists[0]);
  const int irec = (5 + 9 * rmax + 3) & ~3;
  std::vector<int32_t> ints[2];
  std::vector<double> dbls[2];
  std::vector<int32_t> kids;
  bool ok = PySequence_Fast_GET_SIZE(lists[1]) == T;
  if (!ok) PyErr_SetString(PyExc_ValueError, "pack_states: the two lists differ in length");
  for (int w = 0; ok && w < 2; ++w) {
    ints[w].assign((size_t)T * irec, 0);
    dbls[w].assign((size_t)T * 8, 0.0);
  }
  for (Py_ssize_t t = 0; ok && t < T; ++t) {
    for (int w = 0; w < 2 && ok; ++w) {
      PyObject* o = PySequence_Fast_GET_ITEM(lists[w], t);
      if (!PyObject_TypeCheck(o, &FStateType)) {
        PyErr_SetString(PyExc_TypeError, "pack_states: not a ForestState");
        ok = false;
        break;
      }
      FState* s = (FState*)o;
      const int R = (int)s->roots->size();
      if (R > rmax) {
        PyErr_SetString(PyExc_OverflowError, "pack_states: forest with more roots than rmax");
        ok = false;
        break;
      }
      int32_t* rec = ints[w].data() + (size_t)t * irec;
      rec[0] = R;
      rec[1] = (int32_t)s->n_nodes;
      rec[2] = (int32_t)s->n_slots;
      rec[3] = (int32_t)s->n_edges;
      rec[4] = (int32_t)s->n_out;
      int32_t* body = rec + 5;
      for (int r = 0; r < R; ++r) {
        body[0 * rmax + r] = (int32_t)(*s->roots)[r];
        body[1 * rmax + r] = (*s->lr)[r];
        body[2 * rmax + r] = (*s->lp)[r];
        body[3 * rmax + r] = (*s->ndata)[r];
        body[4 * rmax + r] = (*s->size)[r];
        body[5 * rmax + r] = (*s->slot)[r];
        body[6 * rmax + r] = (*s->gidx)[r];
        body[7 * rmax + r] = (int32_t)kids.size();
        body[8 * rmax + r] = (int32_t)(*s->kids)[r].size();
        kids.insert(kids.end(), (*s->kids)[r].begin(), (*s->kids)[r].end());
      }
      double* d = dbls[w].data() + (size_t)t * 8;
      d[0] = s->crp;
      d[1] = s->mult;
      d[2] = s->oprior;
      d[3] = s->omarg;
    }
  }
  Py_DECREF(lists[0]);
  Py_DECREF(lists[1]);
  if (!ok) return nullptr;
  return Py_BuildValue("(y#y#y#y#y#)", (const char*)ints[0].data(), (Py_ssize_t)(ints[0].size() * 4),
                       (const char*)ints[1].data(), (Py_ssize_t)(ints[1].size() * 4), (const char*)dbls[0].data(),
                       (Py_ssize_t)(dbls[0].size() * 8), (const char*)dbls[1].data(), (Py_ssize_t)(dbls[1].size() * 8),
                       kids.empty() ? "" : (const char*)kids.data(), (Py_ssize_t)(kids.size() * 4));
}

PyMethodDef FState_methods[] = {
    {"export", (PyCFunction)FState_export, METH_NOARGS, "all fields as a tuple"},
    {"extend", (PyCFunction)FState_extend, METH_VARARGS, "thawed successor state (None if the arena is full)"},
    {"extend_built", (PyCFunction)FState_extend_built, METH_VARARGS, "as-built successor state (None if the arena is full)"},
    {"score_self", (PyCFunction)FState_score_self, METH_VARARGS, "(prior_p, prior_one, score id | -1) of the state's own tree"},
    {"score_attach_all", (PyCFunction)FState_score_attach_all, METH_VARARGS, ""},
    {"score_new", (PyCFunction)FState_score_new, METH_VARARGS, ""},
    {"score_outlier", (PyCFunction)FState_score_outlier, METH_VARARGS, ""},
    {nullptr, nullptr, 0, nullptr}};

PyGetSetDef FState_getset[] = {{"roots", (getter)FState_get_roots, nullptr, "root names, in order", nullptr},
                               {"n_nodes", (getter)FState_get_n_nodes, nullptr, "", nullptr},
                               {"skey", (getter)FState_get_skey, nullptr, "structure hash", nullptr},
                               {"last_added", (getter)FState_get_last, nullptr, "", nullptr},
                               {"lr", (getter)FState_get_lr, nullptr, "log_r tile ids of the roots", nullptr},
                               {nullptr, nullptr, nullptr, nullptr, nullptr}};

PyTypeObject FStateType = {PyVarObject_HEAD_INIT(nullptr, 0)};

PyObject* mod_set_tables(PyObject*, PyObject* args) {
  PyObject *logs, *rterms;
  if (!PyArg_ParseTuple(args, "OO", &logs, &rterms)) return nullptr;
  for (int which = 0; which < 2; ++which) {
    PyObject* fast = PySequence_Fast(which ? rterms : logs, "expected a sequence of floats");
    if (!fast) return nullptr;
    std::vector<double>& dst = which ? g_rterm : g_log;
    dst.clear();
    for (Py_ssize_t i = 0; i < PySequence_Fast_GET_SIZE(fast); ++i) dst.push_back(PyFloat_AsDouble(PySequence_Fast_GET_ITEM(fast, i)));
    Py_DECREF(fast);
    if (PyErr_Occurred()) return nullptr;
  }
  Py_RETURN_NONE;
}

PyObject* mod_edge_hash(PyObject*, PyObject* args) {
  long e, s, d;
  if (!PyArg_ParseTuple(args, "lll", &e, &s, &d)) return nullptr;
  return PyLong_FromUnsignedLongLong(edge_hash(e, s, d));
}
PyObject* mod_data_hash(PyObject*, PyObject* args) {
  long node, pos, idx;
  if (!PyArg_ParseTuple(args, "lll", &node, &pos, &idx)) return nullptr;
  return PyLong_FromUnsignedLongLong(data_hash(node, pos, idx));
}

// =====================================================================================================
// Holder: one SMC candidate = thawed(parent) + one edit, scored from root-level states (smc/swarm.py adds the
// tree replay on top as a Python subclass).  Mirrors the read interface of the reference's TreeHolder
// (smc/swarm/tree_holder.py:6-110): log_p, log_p_one, tree_roots, node last added to, ...
// =====================================================================================================
#include <structmember.h>

struct Holder {
  PyObject_HEAD
  long uid;
  int kind;
  PyObject* dp;
  PyObject* arg;
  PyObject* parent_holder;
  PyObject* store;         // TileStore: flush() / _grow() / .book
  Book* book;              // strong reference
  FState* parent_state;    // thawed state of the parent particle
  FState* tstate;          // thawed state of this candidate, built on demand
  PyObject* roots_tuple;   // as-built root order
  PyObject* roots_arr;     // the same as an array (made by the Python side on first use, or shared)
  PyObject* vkey;
  PyObject* thawed;        // cache slot of the Python subclass (replayed tree)
  PyObject* grid_size;
  long node_last_added_to, nkids;
  double prior_p, prior_one, log_p, log_p_one, log_pdf;
  int sid, resolved;
  DataPt dpt;
};

extern PyTypeObject HolderType;
long g_next_uid = -1;  // negative: never collides with the Python-side holder counter

Holder* holder_new(PyTypeObject* cls) {
  Holder* h = (Holder*)cls->tp_alloc(cls, 0);
  if (!h) return nullptr;
  h->uid = g_next_uid--;
  h->kind = 0;
  h->dp = h->arg = h->parent_holder = h->store = nullptr;
  h->book = nullptr;
  h->parent_state = h->tstate = nullptr;
  h->roots_tuple = h->roots_arr = h->vkey = h->thawed = h->grid_size = nullptr;
  h->node_last_added_to = h->nkids = 0;
  h->prior_p = h->prior_one = h->log_p = h->log_p_one = h->log_pdf = 0.0;
  h->sid = -1;
  h->resolved = 0;
  return h;
}

void Holder_dealloc(Holder* h) {
  // a genealogy is as long as the data set (2000 at config #5, more for unclustered WGS): hand the parent chain down
  // iteratively -- each ancestor is released with its own parent link already taken, so no call nests deeper than one
  PyObject* parent = h->parent_holder;
  h->parent_holder = nullptr;
  while (parent && parent != Py_None && PyObject_TypeCheck(parent, &HolderType) && Py_REFCNT(parent) == 1) {
    Holder* q = (Holder*)parent;
    PyObject* next = q->parent_holder;
    q->parent_holder = nullptr;
    Py_DECREF(parent);
    parent = next;
  }
  Py_XDECREF(parent);
  Py_XDECREF(h->dp);
  Py_XDECREF(h->arg);
  Py_XDECREF(h->store);
  Py_XDECREF((PyObject*)h->book);
  Py_XDECREF((PyObject*)h->parent_state);
  Py_XDECREF((PyObject*)h->tstate);
  Py_XDECREF(h->roots_tuple);
  Py_XDECREF(h->roots_arr);
  Py_XDECREF(h->vkey);
  Py_XDECREF(h->thawed);
  Py_XDECREF(h->grid_size);
  Py_TYPE(h)->tp_free((PyObject*)h);
}

inline void set_ref(PyObject** slot, PyObject* v) {
  Py_XINCREF(v);
  Py_XDECREF(*slot);
  *slot = v;
}

// common fields of a freshly scored candidate
void holder_fill(Holder* h, int kind, PyObject* dp, const DataPt& dpt, PyObject* arg, PyObject* parent_holder,
                 PyObject* store, Book* book, FState* parent_state, PyObject* grid_size) {
  h->kind = kind;
  h->dpt = dpt;
  set_ref(&h->dp, dp);
  set_ref(&h->arg, arg);
  set_ref(&h->parent_holder, parent_holder);
  set_ref(&h->store, store);
  Py_INCREF((PyObject*)book);
  h->book = book;
  Py_INCREF((PyObject*)parent_state);
  h->parent_state = parent_state;
  set_ref(&h->grid_size, grid_size);
}

bool holder_resolve(Holder* h) {
  if (h->resolved) return true;
  double lp = h->prior_p, lp1 = h->prior_one;
  if (h->sid >= 0) {
    Book* b = h->book;
    if ((size_t)h->sid >= b->sc_ready->size()) {
      PyErr_SetString(PyExc_IndexError, "bad score id");
      return false;
    }
    if (!(*b->sc_ready)[h->sid]) {  // first score read of a batch: run everything that is pending
      PyObject* r = PyObject_CallMethod(h->store, "flush", nullptr);
      if (!r) return false;
      Py_DECREF(r);
      if (!(*b->sc_ready)[h->sid]) {
        PyErr_SetString(PyExc_RuntimeError, "score still pending after flush");
        return false;
      }
    }
    lp += (*b->sc_lse)[h->sid];
    lp1 += (*b->sc_last)[h->sid];
  }
  h->log_p = lp;
  h->log_p_one = lp1;
  h->resolved = 1;
  return true;
}

PyObject* Holder_get_log_p(Holder* h, void*) {
  if (!holder_resolve(h)) return nullptr;
  return PyFloat_FromDouble(h->log_p);
}
PyObject* Holder_get_log_p_one(Holder* h, void*) {
  if (!holder_resolve(h)) return nullptr;
  return PyFloat_FromDouble(h->log_p_one);
}

// thawed_state(): state of from_dict(to_dict(parent + edit)), built on first use
PyObject* Holder_thawed_state(Holder* h, PyObject*) {
  if (!h->tstate) {
    for (int attempt = 0; attempt < 2 && !h->tstate; ++attempt) {
      int status = 0;
      PyObject* arg = h->arg;
      PyObject* tmp = nullptr;
      if (h->kind == K_NEW) {  // the thawed child order is the slot order: any iteration order of the set will do
        tmp = PySequence_Tuple(arg);
        if (!tmp) return nullptr;
        arg = tmp;
      }
      FState* s = fstate_extend(h->parent_state, h->book, h->kind, h->dpt, arg, &status);
      Py_XDECREF(tmp);
      if (status == -1) return nullptr;
      if (status == 1) {
        PyObject* r = PyObject_CallMethod(h->store, "_grow", nullptr);
        if (!r) return nullptr;
        Py_DECREF(r);
        continue;
      }
      h->tstate = s;
    }
    if (!h->tstate) {
      PyErr_SetString(PyExc_MemoryError, "tile arena exhausted");
      return nullptr;
    }
  }
  Py_INCREF((PyObject*)h->tstate);
  return (PyObject*)h->tstate;
}

PyObject* Holder_get_vkey(Holder* h, void*) {
  if (!h->vkey) {
    PyObject* st = Holder_thawed_state(h, nullptr);
    if (!st) return nullptr;
    FState* s = (FState*)st;
    PyObject* last = s->has_last ? PyLong_FromLong(s->last_added) : (Py_INCREF(Py_None), Py_None);
    h->vkey = Py_BuildValue("(KNl)", (unsigned long long)s->skey, last, s->n_nodes + 1);
    Py_DECREF(st);
    if (!h->vkey) return nullptr;
  }
  Py_INCREF(h->vkey);
  return h->vkey;
}

PyObject* Holder_get_outlier_name(Holder*, void*) { return PyLong_FromLong(-1); }

PyMemberDef Holder_members[] = {
    {"uid", T_LONG, offsetof(Holder, uid), READONLY, ""},
    {"kind", T_INT, offsetof(Holder, kind), READONLY, ""},
    {"dp", T_OBJECT, offsetof(Holder, dp), READONLY, ""},
    {"arg", T_OBJECT, offsetof(Holder, arg), READONLY, ""},
    {"_parent_holder", T_OBJECT, offsetof(Holder, parent_holder), READONLY, ""},
    {"_store", T_OBJECT, offsetof(Holder, store), READONLY, ""},
    {"_grid_size", T_OBJECT, offsetof(Holder, grid_size), READONLY, ""},
    {"_roots_tuple", T_OBJECT, offsetof(Holder, roots_tuple), READONLY, ""},
    {"_roots_arr", T_OBJECT, offsetof(Holder, roots_arr), 0, ""},
    {"_thawed", T_OBJECT, offsetof(Holder, thawed), 0, ""},
    {"node_last_added_to", T_LONG, offsetof(Holder, node_last_added_to), READONLY, ""},
    {"num_children_on_node_that_matters", T_LONG, offsetof(Holder, nkids), READONLY, ""},
    {"log_pdf", T_DOUBLE, offsetof(Holder, log_pdf), 0, ""},
    {nullptr, 0, 0, 0, nullptr}};

PyGetSetDef Holder_getset[] = {{"log_p", (getter)Holder_get_log_p, nullptr, "", nullptr},
                               {"log_p_one", (getter)Holder_get_log_p_one, nullptr, "", nullptr},
                               {"vkey", (getter)Holder_get_vkey, nullptr, "value key of the thawed candidate", nullptr},
                               {"outlier_node_name", (getter)Holder_get_outlier_name, nullptr, "", nullptr},
                               {nullptr, nullptr, nullptr, nullptr, nullptr}};

PyMethodDef Holder_methods[] = {
    {"thawed_state", (PyCFunction)Holder_thawed_state, METH_NOARGS, "root-level state of this candidate's own round trip"},
    {nullptr, nullptr, 0, nullptr}};

PyTypeObject HolderType = {PyVarObject_HEAD_INIT(nullptr, 0)};

// shared prior terms of all attach candidates of one base state (same operation order as score_attach_all)
struct AttachTerms {
  double head, tail_p, tail_one;
};
inline AttachTerms attach_terms(FState* base, const DataPt& dp, double log_alpha) {
  double oprior = base->oprior;
  if (dp.op != 0.0) oprior = oprior + dp.opn;
  const long n = base->n_nodes;
  AttachTerms t;
  t.head = (double)n * log_alpha;
  double ways = 0.0;
  for (int z : *base->size) ways += (double)(z - 1) * tlog(z);
  t.tail_p = -(double)(n - 1) * tlog(n + 1) - base->mult + (oprior + base->omarg);
  t.tail_one = (-ways + g_rterm[base->size->size()]) - base->mult + (oprior + base->omarg);
  return t;
}

// attach_holders(cls, store, base_state, parent_state, parent_holder, dp, dp_tuple, nodes, log_alpha, roots_arr)
//   -> [holder per node] | None (arena full)
PyObject* mod_attach_holders(PyObject*, PyObject* args) {
  PyTypeObject* cls;
  PyObject *store, *parent_holder, *dp, *dpo, *nodes_o, *roots_arr;
  FState *base, *pstate;
  double log_alpha;
  if (!PyArg_ParseTuple(args, "O!OO!O!OOOOdO", &PyType_Type, &cls, &store, &FStateType, &base, &FStateType, &pstate,
                        &parent_holder, &dp, &dpo, &nodes_o, &log_alpha, &roots_arr))
    return nullptr;
  DataPt dpt;
  if (!parse_dp(dpo, dpt)) return nullptr;
  std::vector<long> nodes;
  if (!longs_of(nodes_o, nodes)) return nullptr;
  PyObject* book_o = PyObject_GetAttrString(store, "book");
  if (!book_o) return nullptr;
  if (!PyObject_TypeCheck(book_o, &BookType)) {
    Py_DECREF(book_o);
    PyErr_SetString(PyExc_TypeError, "store.book is not a TileBook");
    return nullptr;
  }
  Book* book = (Book*)book_o;
  PyObject* grid = PyObject_GetAttrString(dp, "grid_size");
  if (!grid) {
    Py_DECREF(book_o);
    return nullptr;
  }
  const AttachTerms t = attach_terms(base, dpt, log_alpha);
  PyObject* roots_t = roots_tuple_of(base);
  PyObject* out = PyList_New(0);
  std::vector<int> kids_t;
  bool ok = roots_t && out;
  bool full = false;
  for (size_t q = 0; ok && q < nodes.size(); ++q) {
    const long node = nodes[q];
    const int j = index_of(base, node);
    if (j < 0) {
      PyErr_SetString(PyExc_KeyError, "node is not a root of the base state");
      ok = false;
      break;
    }
    const int m = (*base->ndata)[j];
    const double crp = t.head + (base->crp + (m >= 1 ? lf(m) - lf(m - 1) : 0.0));
    const int lr = book_add(book, (*base->lr)[j], dpt.value_id, false);  // log_r += value (tree/tree_node.py:47-52)
    if (lr < 0) {
      full = true;
      break;
    }
    kids_t = *base->lr;
    kids_t[j] = lr;
    Holder* h = holder_new(cls);
    if (!h) {
      ok = false;
      break;
    }
    PyObject* node_o = PyLong_FromLong(node);
    holder_fill(h, K_ATTACH, dp, dpt, node_o, parent_holder, store, book, pstate, grid);
    Py_DECREF(node_o);
    h->prior_p = crp + t.tail_p;
    h->prior_one = crp + t.tail_one;
    h->sid = book_score(book, kids_t);
    h->node_last_added_to = node;
    h->nkids = (long)(*base->kids)[j].size();
    set_ref(&h->roots_tuple, roots_t);
    if (roots_arr != Py_None) set_ref(&h->roots_arr, roots_arr);
    PyList_Append(out, (PyObject*)h);
    Py_DECREF((PyObject*)h);
  }
  Py_XDECREF(roots_t);
  Py_DECREF(grid);
  Py_DECREF(book_o);
  if (!ok || full) {
    Py_XDECREF(out);
    if (!ok) return nullptr;
    Py_RETURN_NONE;
  }
  return out;
}

// edit_holder(cls, store, base_state, parent_state, parent_holder, kind, dp, dp_tuple, arg, log_alpha)
//   -> holder of base + (NEW root adopting `arg` | OUTLIER) | None (arena full)
PyObject* mod_edit_holder(PyObject*, PyObject* args) {
  PyTypeObject* cls;
  PyObject *store, *parent_holder, *dp, *dpo, *arg;
  FState *base, *pstate;
  int kind;
  double log_alpha;
  if (!PyArg_ParseTuple(args, "O!OO!O!OiOOOd", &PyType_Type, &cls, &store, &FStateType, &base, &FStateType, &pstate,
                        &parent_holder, &kind, &dp, &dpo, &arg, &log_alpha))
    return nullptr;
  if (kind != K_NEW && kind != K_OUTLIER) {
    PyErr_SetString(PyExc_ValueError, "edit_holder: kind must be NEW or OUTLIER");
    return nullptr;
  }
  DataPt dpt;
  if (!parse_dp(dpo, dpt)) return nullptr;
  PyObject* book_o = PyObject_GetAttrString(store, "book");
  if (!book_o) return nullptr;
  if (!PyObject_TypeCheck(book_o, &BookType)) {
    Py_DECREF(book_o);
    PyErr_SetString(PyExc_TypeError, "store.book is not a TileBook");
    return nullptr;
  }
  Book* book = (Book*)book_o;
  // the scoring itself is the ForestState method: (prior_p, prior_one, sid, roots[, k])
  PyObject* rec;
  if (kind == K_NEW) {
    PyObject* children = PySequence_Tuple(arg);  // adoption order = the set's iteration order
    if (!children) {
      Py_DECREF(book_o);
      return nullptr;
    }
    PyObject* a = Py_BuildValue("(OOOd)", book_o, dpo, children, log_alpha);
    Py_DECREF(children);
    rec = a ? FState_score_new(base, a) : nullptr;
    Py_XDECREF(a);
  } else {
    PyObject* a = Py_BuildValue("(OOd)", book_o, dpo, log_alpha);
    rec = a ? FState_score_outlier(base, a) : nullptr;
    Py_XDECREF(a);
  }
  if (!rec || rec == Py_None) {
    Py_DECREF(book_o);
    return rec;  // error, or None: arena full
  }
  PyObject* grid = PyObject_GetAttrString(dp, "grid_size");
  Holder* h = grid ? holder_new(cls) : nullptr;
  if (!h) {
    Py_XDECREF(grid);
    Py_DECREF(rec);
    Py_DECREF(book_o);
    return nullptr;
  }
  holder_fill(h, kind, dp, dpt, arg, parent_holder, store, book, pstate, grid);
  Py_DECREF(grid);
  Py_DECREF(book_o);
  h->prior_p = PyFloat_AsDouble(PyTuple_GET_ITEM(rec, 0));
  h->prior_one = PyFloat_AsDouble(PyTuple_GET_ITEM(rec, 1));
  h->sid = (int)PyLong_AsLong(PyTuple_GET_ITEM(rec, 2));
  set_ref(&h->roots_tuple, PyTuple_GET_ITEM(rec, 3));
  if (kind == K_NEW) {
    h->nkids = PyLong_AsLong(PyTuple_GET_ITEM(rec, 4));
    h->node_last_added_to = base->n_nodes;
  } else {
    h->nkids = 0;
    h->node_last_added_to = -1;
  }
  Py_DECREF(rec);
  return (PyObject*)h;
}

// =====================================================================================================
// Particle (smc/swarm/particle.py): log_q, parent particle, holder; the incremental weight
// log_w = log_p - parent.log_p - log_q (smc/kernels/base.py:38-57) is evaluated on first read, so that every new
// candidate of an SMC step is scored by one batched flush.
// =====================================================================================================
struct Particle {
  PyObject_HEAD
  PyObject* log_q;
  PyObject* parent;
  PyObject* holder;
  double log_w;
  int has_w;
};

extern PyTypeObject ParticleType;

bool holder_scores(PyObject* holder, double* lp, double* lp1) {
  if (PyObject_TypeCheck(holder, &HolderType)) {
    Holder* h = (Holder*)holder;
    if (!holder_resolve(h)) return false;
    *lp = h->log_p;
    *lp1 = h->log_p_one;
    return true;
  }
  PyObject* a = PyObject_GetAttrString(holder, "log_p");
  if (!a) return false;
  *lp = PyFloat_AsDouble(a);
  Py_DECREF(a);
  if (lp1) {
    PyObject* b = PyObject_GetAttrString(holder, "log_p_one");
    if (!b) return false;
    *lp1 = PyFloat_AsDouble(b);
    Py_DECREF(b);
  }
  return !PyErr_Occurred();
}

Particle* particle_make(PyObject* log_q, PyObject* parent, PyObject* holder) {
  Particle* p = PyObject_New(Particle, &ParticleType);
  if (!p) return nullptr;
  Py_INCREF(log_q);
  Py_INCREF(parent);
  Py_INCREF(holder);
  p->log_q = log_q;
  p->parent = parent;
  p->holder = holder;
  p->log_w = 0.0;
  p->has_w = 0;
  return p;
}

PyObject* Particle_new(PyTypeObject*, PyObject* args, PyObject*) {
  PyObject *log_q, *parent, *holder;
  if (!PyArg_ParseTuple(args, "OOO", &log_q, &parent, &holder)) return nullptr;
  return (PyObject*)particle_make(log_q, parent, holder);
}

void Particle_dealloc(Particle* p) {
  // genealogies are as long as the data set: release the parent chain iteratively, not recursively
  PyObject* parent = p->parent;
  Py_XDECREF(p->log_q);
  Py_XDECREF(p->holder);
  PyObject_Del(p);
  while (parent && Py_TYPE(parent) == &ParticleType && Py_REFCNT(parent) == 1) {
    Particle* q = (Particle*)parent;
    PyObject* next = q->parent;
    Py_XDECREF(q->log_q);
    Py_XDECREF(q->holder);
    PyObject_Del(q);
    parent = next;
  }
  Py_XDECREF(parent);
}

bool particle_log_w(Particle* p, double* out) {
  if (!p->has_w) {
    double lp, unused;
    if (!holder_scores(p->holder, &lp, &unused)) return false;
    const double lq = PyFloat_AsDouble(p->log_q);
    if (lq == -1.0 && PyErr_Occurred()) return false;
    double w = lp;
    if (p->parent != Py_None) {
      double plp;
      if (Py_TYPE(p->parent) == &ParticleType) {
        if (!holder_scores(((Particle*)p->parent)->holder, &plp, &unused)) return false;
      } else {
        PyObject* a = PyObject_GetAttrString(p->parent, "log_p");
        if (!a) return false;
        plp = PyFloat_AsDouble(a);
        Py_DECREF(a);
      }
      w = w - plp;
    }
    p->log_w = w - lq;
    p->has_w = 1;
  }
  *out = p->log_w;
  return true;
}

PyObject* Particle_get_log_w(Particle* p, void*) {
  double w;
  if (!particle_log_w(p, &w)) return nullptr;
  return PyFloat_FromDouble(w);
}
PyObject* Particle_get_log_p(Particle* p, void*) {
  double lp, lp1;
  if (!holder_scores(p->holder, &lp, &lp1)) return nullptr;
  return PyFloat_FromDouble(lp);
}
PyObject* Particle_get_log_p_one(Particle* p, void*) {
  double lp, lp1;
  if (!holder_scores(p->holder, &lp, &lp1)) return nullptr;
  return PyFloat_FromDouble(lp1);
}
PyObject* Particle_get_log_pdf(Particle*, void*) { return PyFloat_FromDouble(0.0); }
PyObject* Particle_get_holder(Particle* p, void*) {
  Py_INCREF(p->holder);
  return p->holder;
}
PyObject* Particle_get_tree_roots(Particle* p, void*) { return PyObject_GetAttrString(p->holder, "tree_roots"); }
PyObject* Particle_get_tree_nodes(Particle* p, void*) { return PyObject_GetAttrString(p->holder, "tree_nodes"); }
PyObject* Particle_get_tree(Particle* p, void*) { return PyObject_GetAttrString(p->holder, "tree"); }

PyMemberDef Particle_members[] = {{"_log_q", T_OBJECT, offsetof(Particle, log_q), 0, ""},
                                  {"parent_particle", T_OBJECT, offsetof(Particle, parent), READONLY, ""},
                                  {"_holder", T_OBJECT, offsetof(Particle, holder), READONLY, ""},
                                  {nullptr, 0, 0, 0, nullptr}};

PyGetSetDef Particle_getset[] = {{"log_w", (getter)Particle_get_log_w, nullptr, "", nullptr},
                                 {"log_p", (getter)Particle_get_log_p, nullptr, "", nullptr},
                                 {"log_p_one", (getter)Particle_get_log_p_one, nullptr, "", nullptr},
                                 {"log_pdf", (getter)Particle_get_log_pdf, nullptr, "", nullptr},
                                 {"holder", (getter)Particle_get_holder, nullptr, "", nullptr},
                                 {"tree_roots", (getter)Particle_get_tree_roots, nullptr, "", nullptr},
                                 {"tree_nodes", (getter)Particle_get_tree_nodes, nullptr, "", nullptr},
                                 {"tree", (getter)Particle_get_tree, nullptr, "", nullptr},
                                 {nullptr, nullptr, nullptr, nullptr, nullptr}};

PyTypeObject ParticleType = {PyVarObject_HEAD_INIT(nullptr, 0)};

// particle_log_ws(particles, last_step) -> [log_w] (+ log_p_one - log_p on the last SMC step, samplers/base.py:52-58)
PyObject* mod_particle_log_ws(PyObject*, PyObject* args) {
  PyObject* seq;
  int last;
  if (!PyArg_ParseTuple(args, "Op", &seq, &last)) return nullptr;
  PyObject* fast = PySequence_Fast(seq, "expected a sequence of particles");
  if (!fast) return nullptr;
  const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
  PyObject* out = PyList_New(n);
  for (Py_ssize_t i = 0; out && i < n; ++i) {
    PyObject* o = PySequence_Fast_GET_ITEM(fast, i);
    double w;
    if (Py_TYPE(o) == &ParticleType) {
      Particle* p = (Particle*)o;
      if (!particle_log_w(p, &w)) {
        Py_CLEAR(out);
        break;
      }
      if (last) {
        double lp, lp1;
        if (!holder_scores(p->holder, &lp, &lp1)) {
          Py_CLEAR(out);
          break;
        }
        w = w - lp + lp1;
      }
    } else {
      PyErr_SetString(PyExc_TypeError, "particle_log_ws: not a Particle");
      Py_CLEAR(out);
      break;
    }
    PyList_SET_ITEM(out, i, PyFloat_FromDouble(w));
  }
  Py_DECREF(fast);
  return out;
}

// finish_particles(parents, dist_of, kind, pick, kids, koff, cands, log_qs, meta, new_holder) -> [Particle]
//   Turns the draws of one SMC step (csrc/pcb_host.cpp) into particles: particle i extends parents[i] with the
//   candidate it drew from proposal dist_of[i] -- an enumerated one (kind 0: cands[j][pick]) or a new root adopting
//   kids[koff[i]:koff[i+1]] (made by the Python callback new_holder(j, frozenset), which owns the value-keyed LRU) --
//   and carries the proposal log-probability of semi_adapted.py:36-61.
//   meta[j] = (parent is empty, log 1/2, log(R + 1), R); log_qs[j]: float64 buffer of the normalised candidate scores.
PyObject* finish_particles_core(PyObject* parents_o, const std::vector<long>& dist, const std::vector<long>& kind,
                                const std::vector<long>& pick, const std::vector<long>& kids,
                                const std::vector<long>& koff, PyObject* cands_o, PyObject* logq_o, PyObject* meta_o,
                                PyObject* cb);

PyObject* mod_finish_particles(PyObject*, PyObject* args) {
  PyObject *parents_o, *dist_o, *kind_o, *pick_o, *kids_o, *koff_o, *cands_o, *logq_o, *meta_o, *cb;
  if (!PyArg_ParseTuple(args, "OOOOOOOOOO", &parents_o, &dist_o, &kind_o, &pick_o, &kids_o, &koff_o, &cands_o, &logq_o,
                        &meta_o, &cb))
    return nullptr;
  std::vector<long> dist, kind, pick, kids, koff;
  if (!longs_of(dist_o, dist) || !longs_of(kind_o, kind) || !longs_of(pick_o, pick) || !longs_of(kids_o, kids) ||
      !longs_of(koff_o, koff))
    return nullptr;
  return finish_particles_core(parents_o, dist, kind, pick, kids, koff, cands_o, logq_o, meta_o, cb);
}

PyObject* finish_particles_core(PyObject* parents_o, const std::vector<long>& dist, const std::vector<long>& kind,
                                const std::vector<long>& pick, const std::vector<long>& kids,
                                const std::vector<long>& koff, PyObject* cands_o, PyObject* logq_o, PyObject* meta_o,
                                PyObject* cb) {
  PyObject* parents = PySequence_Fast(parents_o, "parents must be a sequence");
  if (!parents) return nullptr;
  const Py_ssize_t n = PySequence_Fast_GET_SIZE(parents);
  if ((size_t)n != dist.size() || (size_t)n != kind.size() || (size_t)n != pick.size() || koff.size() != (size_t)n + 1) {
    Py_DECREF(parents);
    PyErr_SetString(PyExc_ValueError, "finish_particles: array lengths differ");
    return nullptr;
  }
  struct Made {
    PyObject* holder;
    double log_q;
  };
  std::map<std::pair<long, std::vector<long>>, Made> memo;  // same proposal, same children: same holder and log_q
  PyObject* out = PyList_New(n);
  bool ok = out != nullptr;
  for (Py_ssize_t i = 0; ok && i < n; ++i) {
    const long j = dist[(size_t)i];
    PyObject* meta = PySequence_GetItem(meta_o, j);
    if (!meta) {
      ok = false;
      break;
    }
    int empty;
    double log_half, log_r1;
    long R;
    ok = PyArg_ParseTuple(meta, "pddl", &empty, &log_half, &log_r1, &R) != 0;
    Py_DECREF(meta);
    if (!ok) break;
    PyObject* holder = nullptr;
    double lq = 0.0;
    if (kind[(size_t)i] == 0) {
      PyObject* cl = PySequence_GetItem(cands_o, j);
      holder = cl ? PySequence_GetItem(cl, pick[(size_t)i]) : nullptr;
      Py_XDECREF(cl);
      PyObject* arr = holder ? PySequence_GetItem(logq_o, j) : nullptr;
      Py_buffer view;
      if (!arr || PyObject_GetBuffer(arr, &view, PyBUF_C_CONTIGUOUS) != 0) {
        Py_XDECREF(arr);
        Py_XDECREF(holder);
        ok = false;
        break;
      }
      if ((size_t)view.len < (size_t)(pick[(size_t)i] + 1) * sizeof(double)) {
        PyErr_SetString(PyExc_IndexError, "finish_particles: pick outside the candidate scores");
        ok = false;
      } else {
        const double q = ((const double*)view.buf)[pick[(size_t)i]];
        lq = empty ? q : log_half + q;
      }
      PyBuffer_Release(&view);
      Py_DECREF(arr);
      if (!ok) {
        Py_XDECREF(holder);
        break;
      }
    } else {
      std::pair<long, std::vector<long>> key(j, std::vector<long>(kids.begin() + koff[(size_t)i], kids.begin() + koff[(size_t)i + 1]));
      auto it = memo.find(key);
      if (it == memo.end()) {
        PyObject* lst = PyList_New(0);
        for (long c : key.second) {
          PyObject* v = PyLong_FromLong(c);
          PyList_Append(lst, v);
          Py_DECREF(v);
        }
        PyObject* fs = PyFrozenSet_New(lst);
        Py_DECREF(lst);
        PyObject* h = nullptr;
        if (fs && PyTuple_Check(cb)) {
          // (holder class, data point, its host tuple, log alpha, [(store, parent state, parent holder) per proposal]):
          // the new-root candidate is made right here; within a step (proposal, child set) is unique in `memo`, which
          // is all the reference's value-keyed cache (semi_adapted.py:186-194) can ever hit on
          PyObject *cls, *dp, *dpo, *per;
          double la;
          if (PyArg_ParseTuple(cb, "OOOdO", &cls, &dp, &dpo, &la, &per)) {
            PyObject* ctx = PySequence_GetItem(per, j);
            PyObject *store, *pstate, *pholder;
            if (ctx && PyArg_ParseTuple(ctx, "OOO", &store, &pstate, &pholder)) {
              for (int attempt = 0; attempt < 2; ++attempt) {
                PyObject* a = Py_BuildValue("(OOOOOiOOOd)", cls, store, pstate, pstate, pholder, (int)K_NEW, dp, dpo, fs, la);
                h = a ? mod_edit_holder(nullptr, a) : nullptr;
                Py_XDECREF(a);
                if (h != Py_None) break;
                Py_DECREF(h);  // arena full: grow it and record again
                h = nullptr;
                PyObject* r = PyObject_CallMethod(store, "_grow", nullptr);
                if (!r) break;
                Py_DECREF(r);
              }
              if (!h && !PyErr_Occurred()) PyErr_SetString(PyExc_MemoryError, "tile arena exhausted");
            }
            Py_XDECREF(ctx);
          }
        } else if (fs) {
          h = PyObject_CallFunction(cb, "lO", j, fs);
        }
        Py_XDECREF(fs);
        if (!h) {
          ok = false;
          break;
        }
        long nk;
        if (PyObject_TypeCheck(h, &HolderType)) {
          nk = ((Holder*)h)->nkids;
        } else {
          PyObject* a = PyObject_GetAttrString(h, "num_children_on_node_that_matters");
          nk = a ? PyLong_AsLong(a) : -1;
          Py_XDECREF(a);
          if (PyErr_Occurred()) {
            Py_DECREF(h);
            ok = false;
            break;
          }
        }
        const double lbc = lf(R) - lf(nk) - lf(R - nk);  // utils/math.py:180-182
        Made m{h, log_half - (log_r1 + lbc)};
        it = memo.emplace(std::move(key), m).first;  // the map keeps the reference
      }
      holder = it->second.holder;
      Py_INCREF(holder);
      lq = it->second.log_q;
    }
    PyObject* lqo = PyFloat_FromDouble(lq);
    Particle* p = lqo ? particle_make(lqo, PySequence_Fast_GET_ITEM(parents, i), holder) : nullptr;
    Py_XDECREF(lqo);
    Py_DECREF(holder);
    if (!p) {
      ok = false;
      break;
    }
    PyList_SET_ITEM(out, i, (PyObject*)p);
  }
  for (auto& kv : memo) Py_DECREF(kv.second.holder);
  Py_DECREF(parents);
  if (!ok) {
    Py_XDECREF(out);
    return nullptr;
  }
  return out;
}

// propose_step(draws_fn, bitgen, parents, dist_of, q_list, roots_list, cands, log_qs, meta, new_holder) -> [Particle]
//   One SMC step after the proposals exist: flattens the proposals' q vectors and root lists, makes every particle's
//   draws with ONE call of pcb_smc_draws (csrc/pcb_host.cpp: numpy's own Generator code on the same bit generator, in
//   swarm order; called through its address) and turns them into particles (finish_particles).
typedef int (*smc_draws_fn)(void*, int64_t, const int32_t*, const double*, const int64_t*, const int64_t*, const int64_t*,
                            const uint8_t*, int32_t*, int64_t*, int64_t*, int64_t*);

PyObject* mod_propose_step(PyObject*, PyObject* args) {
  unsigned long long fn_a, bitgen_a;
  PyObject *parents_o, *dist_o, *q_list, *roots_list, *cands_o, *logq_o, *meta_o, *cb;
  if (!PyArg_ParseTuple(args, "KKOOOOOOOO", &fn_a, &bitgen_a, &parents_o, &dist_o, &q_list, &roots_list, &cands_o, &logq_o,
                        &meta_o, &cb))
    return nullptr;
  std::vector<long> dist;
  if (!longs_of(dist_o, dist)) return nullptr;
  const Py_ssize_t nd = PySequence_Size(q_list);
  if (nd < 0 || PySequence_Size(roots_list) != nd || PySequence_Size(meta_o) != nd) {
    PyErr_SetString(PyExc_ValueError, "propose_step: per-proposal lists differ in length");
    return nullptr;
  }
  std::vector<double> q_flat;
  std::vector<int64_t> q_off(1, 0), r_flat, r_off(1, 0);
  std::vector<uint8_t> empty;
  for (Py_ssize_t j = 0; j < nd; ++j) {
    for (int which = 0; which < 2; ++which) {
      PyObject* arr = PySequence_GetItem(which ? roots_list : q_list, j);
      if (!arr) return nullptr;
      Py_buffer view;
      if (PyObject_GetBuffer(arr, &view, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) != 0) {
        Py_DECREF(arr);
        return nullptr;
      }
      const bool ok = view.itemsize == 8 && view.format && (which ? (view.format[0] == 'l' || view.format[0] == 'q')
                                                                  : view.format[0] == 'd');
      if (ok) {
        const Py_ssize_t m = view.len / 8;
        if (which) {
          r_flat.insert(r_flat.end(), (const int64_t*)view.buf, (const int64_t*)view.buf + m);
          r_off.push_back((int64_t)r_flat.size());
        } else {
          q_flat.insert(q_flat.end(), (const double*)view.buf, (const double*)view.buf + m);
          q_off.push_back((int64_t)q_flat.size());
        }
      }
      PyBuffer_Release(&view);
      Py_DECREF(arr);
      if (!ok) {
        PyErr_SetString(PyExc_TypeError, "propose_step: q must be float64 and roots int64 (contiguous)");
        return nullptr;
      }
    }
    PyObject* meta = PySequence_GetItem(meta_o, j);
    if (!meta) return nullptr;
    const int is_empty = PyObject_IsTrue(PyTuple_Check(meta) && PyTuple_GET_SIZE(meta) > 0 ? PyTuple_GET_ITEM(meta, 0) : Py_False);
    Py_DECREF(meta);
    empty.push_back(is_empty ? 1 : 0);
  }
  const size_t n = dist.size();
  std::vector<int32_t> dist32(n), kind32(n);
  size_t kid_cap = 1;
  for (size_t i = 0; i < n; ++i) {
    if (dist[i] < 0 || dist[i] >= nd) {
      PyErr_SetString(PyExc_IndexError, "propose_step: proposal index out of range");
      return nullptr;
    }
    dist32[i] = (int32_t)dist[i];
    kid_cap += (size_t)(r_off[(size_t)dist[i] + 1] - r_off[(size_t)dist[i]]);
  }
  std::vector<int64_t> pick64(n), kids64(kid_cap), koff64(n + 1);
  if (q_flat.empty()) q_flat.push_back(0.0);
  if (r_flat.empty()) r_flat.push_back(0);
  const int rc = ((smc_draws_fn)(uintptr_t)fn_a)((void*)(uintptr_t)bitgen_a, (int64_t)n, dist32.data(), q_flat.data(),
                                                 q_off.data(), r_flat.data(), r_off.data(), empty.data(), kind32.data(),
                                                 pick64.data(), kids64.data(), koff64.data());
  if (rc != 0) {
    PyErr_Format(PyExc_RuntimeError, "pcb_smc_draws failed (rc=%d)", rc);
    return nullptr;
  }
  std::vector<long> kind(kind32.begin(), kind32.end()), pick(pick64.begin(), pick64.end()),
      koff(koff64.begin(), koff64.end()), kids(kids64.begin(), kids64.begin() + (n ? koff64[n] : 0));
  return finish_particles_core(parents_o, dist, kind, pick, kids, koff, cands_o, logq_o, meta_o, cb);
}

// =====================================================================================================
// Data-point Gibbs move: scores of re-attaching one data point to every node (and the outlier set), recorded from
// the ONE current tree (mcmc/gibbs_mh.py:35-70 in the reference copies the tree once per candidate).
// gibbs_scores(book, parent, kids, lp, lr, root, cand, cand_ndata, oi, n_old, n_nodes, crp, mult, sizes, oprior, omarg,
//              dp_tuple, log_alpha, with_outlier) -> [(prior_p, prior_one, sid)] | None (arena full)
//   parent / kids / lp / lr are indexed by graph slot; cand lists the candidate nodes in `tree.nodes` order;
//   oi = slot of the node the point sits in now (-1: it is an outlier), n_old = points in that node.
// Floating-point operations in the order of DataPointSampler._score_candidates / forest_state._priors.
// =====================================================================================================
PyObject* mod_gibbs_scores(PyObject*, PyObject* args) {
  Book* book;
  PyObject *parent_o, *kids_o, *lp_o, *lr_o, *cand_o, *cand_n_o, *sizes_o, *dpo;
  long root, oi, n_old, n_nodes;
  double crp, mult, oprior, omarg, log_alpha;
  int with_outlier;
  if (!PyArg_ParseTuple(args, "O!OOOOlOOlllddOddOdp", &BookType, &book, &parent_o, &kids_o, &lp_o,
                        &lr_o, &root, &cand_o, &cand_n_o, &oi, &n_old, &n_nodes, &crp, &mult, &sizes_o, &oprior, &omarg,
                        &dpo, &log_alpha, &with_outlier))
    return nullptr;
  DataPt dp;
  if (!parse_dp(dpo, dp)) return nullptr;
  std::vector<int> parent, lp, lr, cand, cand_n, sizes;
  if (!ints_of(parent_o, parent) || !ints_of(lp_o, lp) || !ints_of(lr_o, lr) || !ints_of(cand_o, cand) ||
      !ints_of(cand_n_o, cand_n) || !ints_of(sizes_o, sizes))
    return nullptr;
  std::vector<std::vector<int>> kids;
  {
    PyObject* it = PyObject_GetIter(kids_o);
    if (!it) return nullptr;
    while (PyObject* item = PyIter_Next(it)) {
      std::vector<int> k;
      const bool good = ints_of(item, k);
      Py_DECREF(item);
      if (!good) {
        Py_DECREF(it);
        return nullptr;
      }
      kids.push_back(std::move(k));
    }
    Py_DECREF(it);
    if (PyErr_Occurred()) return nullptr;
  }
  const size_t n_slots = parent.size();
  if (kids.size() != n_slots || lp.size() != n_slots || lr.size() != n_slots || cand.size() != cand_n.size() ||
      root < 0 || (size_t)root >= n_slots || oi >= (long)n_slots) {
    PyErr_SetString(PyExc_ValueError, "gibbs_scores: inconsistent tree arrays");
    return nullptr;
  }
  const int NONE = -0x7fffffff;
  std::vector<int> over(n_slots, NONE), over2(n_slots, NONE);
  std::vector<int> touched2, kt;
  int lp_old = -1;
  double crp_removed = crp;
  if (oi < 0) {  // currently an outlier: removing it touches no tile (tree/tree.py:446-454)
    if (dp.op != 0.0) oprior = oprior - dp.op + dp.opn;
    omarg = omarg - dp.marg;
  } else {
    // removal: new log_p of the old node, refreshed log_r along old node .. root
    lp_old = book_add(book, lp[(size_t)oi], dp.value_id, true);
    if (lp_old < 0) Py_RETURN_NONE;
    long i = oi;
    bool first = true;
    while (i != root) {
      kt.clear();
      for (int c : kids[(size_t)i]) kt.push_back(over[(size_t)c] != NONE ? over[(size_t)c] : lr[(size_t)c]);
      const int base = first ? lp_old : lp[(size_t)i];
      const int t = kt.empty() ? base : book_node(book, base, kt);
      if (t < 0) Py_RETURN_NONE;
      over[(size_t)i] = t;
      first = false;
      i = parent[(size_t)i];
      if (i < 0) {
        PyErr_SetString(PyExc_ValueError, "gibbs_scores: node without a path to the root");
        return nullptr;
      }
    }
    crp_removed = crp - (lf(n_old - 1) - lf(n_old - 2));
  }
  auto cur = [&](int c) { return over[(size_t)c] != NONE ? over[(size_t)c] : lr[(size_t)c]; };
  PyObject* out = PyList_New(0);
  if (!out) return nullptr;
  const std::vector<int>& roots = kids[(size_t)root];
  for (size_t q = 0; q < cand.size(); ++q) {
    const int ni = cand[q];
    const int m = cand_n[q] - (ni == oi ? 1 : 0);  // points in the node once the point is removed
    for (int t : touched2) over2[(size_t)t] = NONE;
    touched2.clear();
    int t0 = book_add(book, cur(ni), dp.value_id, false);
    if (t0 < 0) {
      Py_DECREF(out);
      Py_RETURN_NONE;
    }
    over2[(size_t)ni] = t0;
    touched2.push_back(ni);
    long i = parent[(size_t)ni];
    while (i != root) {
      if (i < 0) {
        Py_DECREF(out);
        PyErr_SetString(PyExc_ValueError, "gibbs_scores: node without a path to the root");
        return nullptr;
      }
      kt.clear();
      for (int c : kids[(size_t)i]) kt.push_back(over2[(size_t)c] != NONE ? over2[(size_t)c] : cur(c));
      const int t = book_node(book, i == oi ? lp_old : lp[(size_t)i], kt);
      if (t < 0) {
        Py_DECREF(out);
        Py_RETURN_NONE;
      }
      over2[(size_t)i] = t;
      touched2.push_back((int)i);
      i = parent[(size_t)i];
    }
    kt.clear();
    for (int c : roots) kt.push_back(over2[(size_t)c] != NONE ? over2[(size_t)c] : cur(c));
    const double c2 = crp_removed + (m >= 1 ? lf(m) - lf(m - 1) : 0.0);
    double p, p1;
    priors(log_alpha, n_nodes, c2, mult, sizes.data(), sizes.size(), oprior, omarg, p, p1);
    PyObject* rec = Py_BuildValue("(ddi)", p, p1, book_score(book, kt));
    PyList_Append(out, rec);
    Py_DECREF(rec);
  }
  if (with_outlier) {
    kt.clear();
    for (int c : roots) kt.push_back(cur(c));
    double op = oprior;
    if (dp.op != 0.0) op = oprior - dp.opn + dp.op;
    double p, p1;
    priors(log_alpha, n_nodes, crp_removed, mult, sizes.data(), sizes.size(), op, omarg + dp.marg, p, p1);
    PyObject* rec = Py_BuildValue("(ddi)", p, p1, kt.empty() ? -1 : book_score(book, kt));
    PyList_Append(out, rec);
    Py_DECREF(rec);
  }
  return out;
}

// log_sum_exp(float64 buffer) (utils/math.py:102-118): max shift, terms summed left to right with libm exp / log
PyObject* mod_log_sum_exp(PyObject*, PyObject* arg) {
  Py_buffer view;
  if (PyObject_GetBuffer(arg, &view, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) != 0) return nullptr;
  if (view.itemsize != 8 || !view.format || view.format[0] != 'd') {
    PyBuffer_Release(&view);
    PyErr_SetString(PyExc_TypeError, "log_sum_exp: expected a contiguous float64 buffer");
    return nullptr;
  }
  const double* x = (const double*)view.buf;
  const Py_ssize_t n = view.len / 8;
  if (n == 0) {
    PyBuffer_Release(&view);
    PyErr_SetString(PyExc_ValueError, "log_sum_exp of an empty array");
    return nullptr;
  }
  double m = x[0];
  bool nan = std::isnan(m);
  for (Py_ssize_t i = 1; i < n; ++i) {
    nan |= std::isnan(x[i]);
    if (x[i] > m) m = x[i];
  }
  double out;
  if (nan) {
    out = std::nan("");
  } else if (std::isinf(m)) {
    out = m;
  } else {
    double total = 0.0;
    for (Py_ssize_t i = 0; i < n; ++i) total += std::exp(x[i] - m);
    out = std::log(total) + m;
  }
  PyBuffer_Release(&view);
  return PyFloat_FromDouble(out);
}

PyMethodDef mod_methods[] = {
    {"forest_state", mod_forest_state, METH_VARARGS, "build a ForestState from Python sequences"},
    {"pack_states", mod_pack_states, METH_VARARGS, "retained-path records of the device-resident sweep"},
    {"log_sum_exp", mod_log_sum_exp, METH_O, "log sum exp of a float64 buffer, reference summation order"},
    {"particle_log_ws", mod_particle_log_ws, METH_VARARGS, "incremental weights of a list of particles"},
    {"propose_step", mod_propose_step, METH_VARARGS, "draws of one SMC step (pcb_smc_draws) and the particles they make"},
    {"finish_particles", mod_finish_particles, METH_VARARGS, "particles of one SMC step from its draws"},
    {"gibbs_scores", mod_gibbs_scores, METH_VARARGS, "scores of moving one data point to every node / the outliers"},
    {"attach_holders", mod_attach_holders, METH_VARARGS, "scored candidates: the data point joins each given root"},
    {"edit_holder", mod_edit_holder, METH_VARARGS, "scored candidate: new root adopting some roots, or outlier"},
    {"set_tables", mod_set_tables, METH_VARARGS, "log(n) and r_term(n) tables"},
    {"edge_hash", mod_edge_hash, METH_VARARGS, "structure-hash contribution of an edge (slot, src, dst)"},
    {"data_hash", mod_data_hash, METH_VARARGS, "structure-hash contribution of a data point (node, position, idx)"},
    {nullptr, nullptr, 0, nullptr}};


PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_pcbhost", "host-side bookkeeping of the phyclone_b200 tile store", -1,
                      mod_methods};

}  // namespace

PyMODINIT_FUNC PyInit__pcbhost(void) {
  BookType.tp_name = "_pcbhost.TileBook";
  BookType.tp_basicsize = sizeof(Book);
  BookType.tp_flags = Py_TPFLAGS_DEFAULT;
  BookType.tp_new = PyType_GenericNew;
  BookType.tp_init = (initproc)Book_init;
  BookType.tp_dealloc = (destructor)Book_dealloc;
  BookType.tp_methods = Book_methods;
  if (PyType_Ready(&BookType) < 0) return nullptr;
  PyObject* m = PyModule_Create(&moddef);
  if (!m) return nullptr;
  Py_INCREF(&BookType);
  PyModule_AddObject(m, "TileBook", (PyObject*)&BookType);
  FStateType.tp_name = "_pcbhost.ForestState";
  FStateType.tp_basicsize = sizeof(FState);
  FStateType.tp_flags = Py_TPFLAGS_DEFAULT;
  FStateType.tp_dealloc = (destructor)FState_dealloc;
  FStateType.tp_methods = FState_methods;
  FStateType.tp_getset = FState_getset;
  if (PyType_Ready(&FStateType) < 0) return nullptr;
  Py_INCREF(&FStateType);
  PyModule_AddObject(m, "ForestState", (PyObject*)&FStateType);
  HolderType.tp_name = "_pcbhost.Holder";
  HolderType.tp_basicsize = sizeof(Holder);
  HolderType.tp_flags = Py_TPFLAGS_DEFAULT | Py_TPFLAGS_BASETYPE;
  HolderType.tp_dealloc = (destructor)Holder_dealloc;
  HolderType.tp_members = Holder_members;
  HolderType.tp_getset = Holder_getset;
  HolderType.tp_methods = Holder_methods;
  if (PyType_Ready(&HolderType) < 0) return nullptr;
  Py_INCREF(&HolderType);
  PyModule_AddObject(m, "Holder", (PyObject*)&HolderType);
  ParticleType.tp_name = "_pcbhost.Particle";
  ParticleType.tp_basicsize = sizeof(Particle);
  ParticleType.tp_flags = Py_TPFLAGS_DEFAULT;
  ParticleType.tp_new = Particle_new;
  ParticleType.tp_dealloc = (destructor)Particle_dealloc;
  ParticleType.tp_members = Particle_members;
  ParticleType.tp_getset = Particle_getset;
  if (PyType_Ready(&ParticleType) < 0) return nullptr;
  Py_INCREF(&ParticleType);
  PyModule_AddObject(m, "Particle", (PyObject*)&ParticleType);
  return m;
}
