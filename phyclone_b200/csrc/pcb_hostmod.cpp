// _pcbhost: CPython extension with the host-side bookkeeping of the tile store (see tiles.py).
//
// TileBook = memo tables + pending-op lists + dependency levels + flush packing of `tiles.TileStore`,
// moved out of the interpreter because a PG iteration records ~15 k tile ops and resolves ~5 k scores
// (config #2) and each of them was 1-3 us of Python.  Pure host code: no CUDA here, the packed plan it
// produces is executed by `engine.DeviceArena.run_plan` (or by the NumPy executor of the CPU tests).
#define PY_SSIZE_T_CLEAN
#include <Python.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
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
  if (!PyArg_ParseTuple(args, "LL", &permanent, &capacity)) return -1;
  self->permanent = self->top = permanent;
  self->capacity = capacity;
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

PyObject* add_impl(Book* self, PyObject* const* args, Py_ssize_t nargs, bool neg) {
  if (nargs != 2) {
    PyErr_SetString(PyExc_TypeError, "add(a, b)");
    return nullptr;
  }
  const long a = PyLong_AsLong(args[0]), b = PyLong_AsLong(args[1]);
  if ((a == -1 || b == -1) && PyErr_Occurred()) return nullptr;
  const uint64_t key = ((uint64_t)(uint32_t)a << 33) | ((uint64_t)(uint32_t)b << 1) | (neg ? 1u : 0u);
  auto it = self->add_memo->find(key);
  if (it != self->add_memo->end()) {
    ++self->n_hits;
    return PyLong_FromLong(it->second);
  }
  if (full(self)) return PyLong_FromLong(-1);
  const int t = (int)self->top++;
  const int lv = 1 + std::max(level_of(self, (int)a), level_of(self, (int)b));
  set_level(self, t, lv);
  self->adds->push_back(PendAdd{lv, (int)a, (int)b, t, neg});
  (*self->add_memo)[key] = t;
  return PyLong_FromLong(t);
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

PyObject* Book_node(Book* self, PyObject* const* args, Py_ssize_t nargs) {
  if (nargs != 2) {
    PyErr_SetString(PyExc_TypeError, "node(base, kids)");
    return nullptr;
  }
  const long base = PyLong_AsLong(args[0]);
  if (base == -1 && PyErr_Occurred()) return nullptr;
  std::vector<int>& kids = g_tmp;
  if (!ints_of(args[1], kids)) return nullptr;
  std::string key = key_of((int)base, kids);
  auto it = self->node_memo->find(key);
  if (it != self->node_memo->end()) {
    ++self->n_hits;
    return PyLong_FromLong(it->second);
  }
  if (full(self)) return PyLong_FromLong(-1);
  const int t = (int)self->top++;
  int lv = level_of(self, (int)base);
  for (int k : kids) lv = std::max(lv, level_of(self, k));
  ++lv;
  set_level(self, t, lv);
  self->jobs->push_back(PendJob{lv, (int)base, t, -1, (uint32_t)self->kids->size(), (uint32_t)kids.size()});
  self->kids->insert(self->kids->end(), kids.begin(), kids.end());
  self->node_memo->emplace(std::move(key), t);
  return PyLong_FromLong(t);
}

// score(kids) -> score id (resolve with value())
PyObject* Book_score(Book* self, PyObject* arg) {
  std::vector<int>& kids = g_tmp;
  if (!ints_of(arg, kids)) return nullptr;
  std::string key = key_of(-7, kids);
  auto it = self->score_memo->find(key);
  if (it != self->score_memo->end()) {
    ++self->n_hits;
    return PyLong_FromLong(it->second);
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
  return PyLong_FromLong(sid);
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
PyObject* Book_pack(Book* self, PyObject*) {
  int max_level = 0;
  for (const PendAdd& a : *self->adds) max_level = std::max(max_level, a.level);
  for (const PendJob& j : *self->jobs) max_level = std::max(max_level, j.level);
  std::vector<int32_t> packed;
  packed.reserve(self->adds->size() * 3 + self->jobs->size() * 8 + self->kids->size());
  PyObject* plan = PyList_New(0);
  if (!plan) return nullptr;
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
      PyObject* rec = Py_BuildValue("(snnd)", "add", (Py_ssize_t)off, (Py_ssize_t)n, neg ? -1.0 : 1.0);
      PyList_Append(plan, rec);
      Py_DECREF(rec);
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
    PyObject* rec = Py_BuildValue("(snnnn)", "jobs", (Py_ssize_t)off, (Py_ssize_t)n, (Py_ssize_t)koff,
                                  (Py_ssize_t)(koff - (int32_t)n));
    PyList_Append(plan, rec);
    Py_DECREF(rec);
    self->n_jobs += (int64_t)n;
    self->n_pairs += (int64_t)(koff - (int32_t)n);
  }
  PyObject* bytes = PyBytes_FromStringAndSize((const char*)packed.data(), (Py_ssize_t)(packed.size() * 4));
  PyObject* out = Py_BuildValue("(NNn)", bytes, plan, (Py_ssize_t)self->pending_scores->size());
  return out;
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
    const double* p = (const double*)view.buf;
    for (size_t s = 0; s < n; ++s) {
      const int sid = (*self->pending_scores)[s];
      (*self->sc_lse)[sid] = p[s];
      (*self->sc_last)[sid] = p[n + s];
      (*self->sc_ready)[sid] = 1;
    }
    PyBuffer_Release(&view);
    self->n_scores += (int64_t)n;
  }
  clear_pending(self);
  Py_RETURN_NONE;
}

PyObject* Book_set_capacity(Book* self, PyObject* arg) {
  const long long c = PyLong_AsLongLong(arg);
  if (c == -1 && PyErr_Occurred()) return nullptr;
  self->capacity = c;
  Py_RETURN_NONE;
}

PyObject* Book_counters(Book* self, PyObject*) {
  return Py_BuildValue("{s:L,s:L,s:L,s:L,s:L,s:L}", "adds", (long long)self->n_adds, "node_jobs",
                       (long long)self->n_jobs, "scores", (long long)self->n_scores, "memo_hits",
                       (long long)self->n_hits, "conv_pairs", (long long)self->n_pairs, "top", (long long)self->top);
}

PyMethodDef Book_methods[] = {
    {"reset", (PyCFunction)Book_reset, METH_NOARGS, "start a new arena generation"},
    {"add", (PyCFunction)(void (*)(void))Book_add, METH_FASTCALL, "tile(a) + tile(b) -> id (or -1 if the arena is full)"},
    {"sub", (PyCFunction)(void (*)(void))Book_sub, METH_FASTCALL, "tile(a) - tile(b) -> id (or -1 if the arena is full)"},
    {"node", (PyCFunction)(void (*)(void))Book_node, METH_FASTCALL, "tile(base) + log_S(kids) -> id (or -1)"},
    {"score", (PyCFunction)Book_score, METH_O, "score handle of a dummy root folding kids"},
    {"value", (PyCFunction)Book_value, METH_O, "(lse_sum, last_sum) or None while pending"},
    {"has_pending", (PyCFunction)Book_has_pending, METH_NOARGS, ""},
    {"pack", (PyCFunction)Book_pack, METH_NOARGS, "(packed bytes, plan, n_slots) of the pending batch"},
    {"done", (PyCFunction)Book_done, METH_O, "hand back the [2, n_slots] scores of the packed batch"},
    {"set_capacity", (PyCFunction)Book_set_capacity, METH_O, ""},
    {"counters", (PyCFunction)Book_counters, METH_NOARGS, ""},
    {nullptr, nullptr, 0, nullptr}};

PyTypeObject BookType = {PyVarObject_HEAD_INIT(nullptr, 0)};

PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_pcbhost", "host-side bookkeeping of the phyclone_b200 tile store", -1,
                      nullptr};

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
  return m;
}
