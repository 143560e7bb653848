"""Device-resident tile arena + batched node-refresh jobs (host side of the hot path).

PyTorch is used only as plumbing here: it owns the HBM allocations and the CUDA stream; all
arithmetic is done by the kernels behind the C ABI (`pcb_*_dev` entry points).

Tiles live in HBM in the padded layout described in include/phyclone_b200.h.  Tile ids are
plain integers managed by the caller (see `TileAllocator`).
"""

import ctypes as C

import numpy as np
import torch

from . import _lib

JOB_SCAN = 1
JOB_ADD_PRIOR = 2
JOB_INTS = 8


class TileAllocator:
    """Bump allocator over arena slots with a permanent prefix and a resettable region."""

    def __init__(self, capacity):
        self.capacity = capacity
        self.permanent = 0
        self.top = 0

    def alloc(self, n=1):
        first = self.top
        self.top += n
        if self.top > self.capacity:
            raise _lib.PcbError("tile arena exhausted (%d slots)" % self.capacity)
        return first

    def freeze(self):
        """Everything allocated so far survives `reset()`."""
        self.permanent = self.top

    def reset(self):
        self.top = self.permanent


class DeviceArena:
    def __init__(self, S, G, capacity, device=None, variant=-1):
        self.lib = _lib.require_device()
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.layout = _lib.Layout()
        _lib.check(self.lib.pcb_make_layout(S, G, variant, C.byref(self.layout)))
        self.S, self.G = S, G
        self.capacity = int(capacity)
        pitch = self.layout.pitch
        with torch.cuda.device(self.device):
            # pads must be (and stay) zero: kernels only ever write the G payload columns
            self.lg = torch.zeros((self.capacity, S, pitch), dtype=torch.float64, device=self.device)
            self.lin = torch.zeros((self.capacity, S, pitch), dtype=torch.float64, device=self.device)
            self.rmax = torch.zeros((self.capacity, S), dtype=torch.float64, device=self.device)
        self.c = _lib.Arena(self.layout, self.lg.data_ptr(), self.lin.data_ptr(), self.rmax.data_ptr(), self.capacity)
        self.alloc = TileAllocator(self.capacity)
        self.prior = -float(np.log(G))
        self._staging = {}
        self._rows = self._sums = self._sums_host = None
        self._pinned_vals = self._staged_vals = None
        self.staged_bytes = 0  # host->device bytes of job tables / id lists so far
        self._plan_synced = True
        self._plan_np = self._sums_np = None

    LATENCY = 4  # PCB_KERNEL_LATENCY (include/phyclone_b200.h)

    def set_latency(self, on=True):
        """Ask for the warp-per-row mapping of small launches (layout.kernel bit 4: csrc/pcb_api.cu).  Results equal
        the default mapping's to rounding (a different association of the same non-negative sums)."""
        if on:
            self.layout.kernel |= self.LATENCY
        else:
            self.layout.kernel &= ~self.LATENCY
        self.c = _lib.Arena(self.layout, self.lg.data_ptr(), self.lin.data_ptr(), self.rmax.data_ptr(), self.capacity)

    def grow(self, new_capacity):
        """Enlarge the arena in place (tile ids stay valid): new zero-filled tensors, old content copied over."""
        new_capacity = int(new_capacity)
        if new_capacity <= self.capacity:
            return
        with torch.cuda.device(self.device):
            for name in ("lg", "lin", "rmax"):
                old = getattr(self, name)
                new = torch.zeros((new_capacity,) + tuple(old.shape[1:]), dtype=old.dtype, device=self.device)
                new[: self.capacity].copy_(old)
                setattr(self, name, new)
        self.capacity = new_capacity
        self.alloc.capacity = new_capacity
        self.c = _lib.Arena(self.layout, self.lg.data_ptr(), self.lin.data_ptr(), self.rmax.data_ptr(), self.capacity)

    # ---------------------------------------------------------------- plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _to_dev_i32(self, arr, key):
        """Host int32 array -> device tensor through a reusable pinned staging buffer."""
        arr = np.ascontiguousarray(arr, dtype=np.int32).reshape(-1)
        n = arr.shape[0]
        st = self._staging.get(key)
        if st is None or st[0].shape[0] < n:
            cap = max(1024, int(n * 1.5))
            st = (
                torch.empty(cap, dtype=torch.int32, pin_memory=True),
                torch.empty(cap, dtype=torch.int32, device=self.device),
                torch.cuda.Event(),
            )
            self._staging[key] = st
        else:
            st[2].synchronize()  # the previous async copy out of this pinned buffer must have drained
        host, dev, ev = st
        self.staged_bytes += 4 * n
        host.numpy()[:n] = arr
        dev[:n].copy_(host[:n], non_blocking=True)
        ev.record(torch.cuda.current_stream(self.device))
        return dev

    # ------------------------------------------------------------------- tiles
    def import_tiles(self, dense, ids):
        """dense: [n,S,G] numpy (host) or CUDA tensor of log tiles -> arena slots `ids`."""
        if isinstance(dense, np.ndarray):
            dense = torch.from_numpy(np.ascontiguousarray(dense, dtype=np.float64)).to(self.device)
        dense = dense.contiguous()
        n = dense.shape[0]
        assert dense.shape[1:] == (self.S, self.G) and len(ids) == n
        ids_dev = self._to_dev_i32(ids, "ids")
        _lib.check(
            self.lib.pcb_arena_import_dev(C.byref(self.c), dense.data_ptr(), ids_dev.data_ptr(), n, self._stream())
        )
        self._keep = dense  # keep alive until the stream has consumed it

    def upload_dense(self, values, ids):
        """Host [n,S,G] float64 -> pinned staging -> HBM -> arena slots `ids`; returns the bytes sent."""
        values = np.ascontiguousarray(values, dtype=np.float64)
        if self._pinned_vals is None or self._pinned_vals.shape != values.shape:
            self._pinned_vals = torch.empty(values.shape, dtype=torch.float64, pin_memory=True)
            self._staged_vals = torch.empty(values.shape, dtype=torch.float64, device=self.device)
        self._pinned_vals.numpy()[...] = values
        self._staged_vals.copy_(self._pinned_vals, non_blocking=True)
        self.import_tiles(self._staged_vals, ids)
        return values.nbytes

    def export_tiles(self, ids):
        n = len(ids)
        out = torch.empty((n, self.S, self.G), dtype=torch.float64, device=self.device)
        ids_dev = self._to_dev_i32(ids, "ids")
        _lib.check(self.lib.pcb_arena_export_dev(C.byref(self.c), ids_dev.data_ptr(), n, out.data_ptr(), self._stream()))
        return out

    def map_trees(self, lp_ids, tree_off, post_order, child_off, child_idx):
        """MAP prevalence indices (process_trace/map.py:48-166) of a batch of trees whose node log_p tiles live in
        this arena: tiles are gathered on the device, one launch walks every (tree, sample) pair, only the
        [N,S] index table comes back.  Arguments as `ops.map_trees`, with tile ids instead of dense log_p."""
        n, n_trees, n_edges = len(lp_ids), len(tree_off) - 1, len(child_idx)
        if n == 0:
            return np.zeros((0, self.S), dtype=np.int32)
        with torch.cuda.device(self.device):
            dense = self.export_tiles(lp_ids)
            ints = np.concatenate([np.asarray(x, dtype=np.int32).reshape(-1)
                                   for x in (tree_off, post_order, child_off, child_idx)])
            d = self._to_dev_i32(ints, "map")
            p_tree = d.data_ptr()
            p_post = p_tree + 4 * (n_trees + 1)
            p_coff = p_post + 4 * n
            p_cidx = p_coff + 4 * (n + 1)
            r_max = torch.empty_like(dense)
            d_choice = torch.empty((max(n_edges, 1), self.S, self.G), dtype=torch.int32, device=self.device)
            s_choice = torch.empty((n, self.S, self.G), dtype=torch.int32, device=self.device)
            max_idx = torch.zeros((n, self.S), dtype=torch.int32, device=self.device)
            _lib.check(self.lib.pcb_map_trees_dev(dense.data_ptr(), p_tree, p_post, p_coff, p_cidx, n_trees, n, n_edges,
                                                  self.S, self.G, r_max.data_ptr(), d_choice.data_ptr(),
                                                  s_choice.data_ptr(), max_idx.data_ptr(), self._stream()))
            return max_idx.cpu().numpy()

    def tile_add(self, a_ids, b_ids, out_ids, addend=0.0, b_sign=1.0):
        """out = a (+/- b) + addend in the log domain; refreshes lin / rmax of `out`."""
        n = len(out_ids)
        if n == 0:
            return
        # one packed staging copy: [a | b | out]
        packed = np.empty(3 * n, dtype=np.int32)
        packed[:n] = a_ids
        packed[n : 2 * n] = b_ids if b_ids is not None else -1
        packed[2 * n :] = out_ids
        d = self._to_dev_i32(packed, "add")
        base = d.data_ptr()
        _lib.check(
            self.lib.pcb_tile_add_dev(C.byref(self.c), base, base + 4 * n, base + 8 * n, n, float(b_sign), float(addend),
                                      self._stream())
        )

    def tile_sub(self, a_ids, b_ids, out_ids):
        self.tile_add(a_ids, b_ids, out_ids, 0.0, -1.0)

    def launch_jobs(self, tab, kids, rows):
        """tab: int32 [n,8] job records; kids: flat child ids; rows: CUDA [2,n_slots,S] tensor or None."""
        n = tab.shape[0]
        nk = kids.shape[0]
        packed = np.empty(n * JOB_INTS + max(nk, 1), dtype=np.int32)
        packed[: n * JOB_INTS] = tab.reshape(-1)
        packed[n * JOB_INTS : n * JOB_INTS + nk] = kids
        d = self._to_dev_i32(packed, "jobs")
        base = d.data_ptr()
        pl = pt = None
        if rows is not None:
            pl, pt = rows[0].data_ptr(), rows[1].data_ptr()
        _lib.check(
            self.lib.pcb_node_jobs_dev(C.byref(self.c), base, base + 4 * n * JOB_INTS, n, self.prior, pl, pt, self._stream())
        )

    def flush_context(self):
        """Addresses `TileBook.run` needs to execute a flush without coming back to Python: the library entry points,
        the arena descriptor, the pinned staging buffer and its device twin, the score buffers, the stream.  Buffers
        are (re)allocated by `run_plan`; call again after it has run."""
        st = self._staging.get("plan")
        if st is None or self._rows is None:
            return None
        if not self._plan_synced:  # the in-extension path overwrites the pinned buffer without asking
            torch.cuda.current_stream(self.device).synchronize()
            self._plan_synced = True
        lib = self.lib
        fn = C.cast(lib.pcb_run_plan_dev, C.c_void_p).value
        err = C.cast(lib.pcb_last_error, C.c_void_p).value
        stream = torch.cuda.current_stream(self.device).cuda_stream
        return (fn, err, C.addressof(self.c), st[0].data_ptr(), st[0].shape[0], st[1].data_ptr(), float(self.prior),
                self._rows[0].data_ptr(), self._rows[1].data_ptr(), self._rows.shape[1], self._sums.data_ptr(),
                self._sums_host.data_ptr(), stream)

    def run_plan(self, packed, plan, n_slots, plan_raw=None):
        """Execute one flush: `packed` (int32) holds every id list / job table of the batch and is sent to the
        device in ONE pinned copy; `plan` is a list of launches over it:
            ("add", offset, n, sign)          ids at offset: [a | b | out], n each
            ("jobs", offset, n, n_kids, cp)   table at offset (n*8 ints) followed by n_kids child ids
        (`plan_raw`: the same as bytes of 6 int32 per record).  The whole flush is one `pcb_run_plan_dev` call.
        Returns a host float64 array [2, n_slots] with the summed score rows (synchronises), or None."""
        n_rec = len(plan)
        if n_rec == 0:
            return None
        if plan_raw is None:
            raw = np.zeros((n_rec, 6), dtype=np.int32)
            for r, rec in enumerate(plan):
                raw[r, :5] = (0, rec[1], rec[2], 0, rec[3] < 0) if rec[0] == "add" else (1, rec[1], rec[2], rec[3], 0)
            plan_raw = raw.tobytes()
        stream_obj = torch.cuda.current_stream(self.device)
        n_ints = packed.shape[0]
        st = self._staging.get("plan")
        if st is None or st[0].shape[0] < n_ints:
            stream_obj.synchronize()  # a copy out of the old pinned buffer may still be in flight
            cap = max(4096, int(n_ints * 1.5))
            st = self._staging["plan"] = (torch.empty(cap, dtype=torch.int32, pin_memory=True),
                                          torch.empty(cap, dtype=torch.int32, device=self.device), None)
            self._plan_np = st[0].numpy()
        elif not self._plan_synced:
            stream_obj.synchronize()  # the previous copy out of the pinned buffer must have drained
        self._plan_np[:n_ints] = packed
        self.staged_bytes += 4 * n_ints
        pl = pt = sd = sh = None
        if n_slots:
            if self._rows is None or self._rows.shape[1] < n_slots:
                cap = max(4096, int(n_slots * 1.5))
                self._rows = torch.empty((2, cap, self.S), dtype=torch.float64, device=self.device)
                self._sums = torch.empty(2 * cap, dtype=torch.float64, device=self.device)
                self._sums_host = torch.empty(2 * cap, dtype=torch.float64, pin_memory=True)
                self._sums_np = self._sums_host.numpy()
            pl, pt = self._rows[0].data_ptr(), self._rows[1].data_ptr()
            sd, sh = self._sums.data_ptr(), self._sums_host.data_ptr()
        _lib.check(self.lib.pcb_run_plan_dev(C.byref(self.c), st[0].data_ptr(), n_ints, st[1].data_ptr(), plan_raw, n_rec,
                                             self.prior, pl, pt, n_slots, sd, sh, None,
                                             C.c_void_p(stream_obj.cuda_stream)))
        self._plan_synced = bool(n_slots)
        if not n_slots:
            return None
        return self._sums_np[: 2 * n_slots].reshape(2, n_slots)

    def reduce_rows(self, rows):
        """rows: contiguous [2, n_slots, S] -> [2, n_slots] (sum over samples in order), one launch."""
        n_slots = rows.shape[1]
        out = torch.empty((2, n_slots), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.pcb_reduce_rows_dev(rows.data_ptr(), 2 * n_slots, self.S, out.data_ptr(), self._stream()))
        return out

    # -------------------------------------------------------------------- jobs
    def run_jobs(self, jobs, children, n_slots=0):
        """Enqueue node-refresh jobs.  jobs: int32 [n,8]; children: int32 flat tile ids.

        Returns (lse_sum, last_sum) CUDA tensors of length n_slots (or None)."""
        jobs = np.ascontiguousarray(jobs, dtype=np.int32).reshape(-1, JOB_INTS)
        n = jobs.shape[0]
        if n == 0:
            return None
        jd = self._to_dev_i32(jobs, "jobs")
        cd = self._to_dev_i32(children if len(children) else np.zeros(1, np.int32), "kids")
        rows_l = rows_t = None
        pl = pt = None
        if n_slots:
            rows_l = torch.empty((n_slots, self.S), dtype=torch.float64, device=self.device)
            rows_t = torch.empty((n_slots, self.S), dtype=torch.float64, device=self.device)
            pl, pt = rows_l.data_ptr(), rows_t.data_ptr()
        _lib.check(
            self.lib.pcb_node_jobs_dev(C.byref(self.c), jd.data_ptr(), cd.data_ptr(), n, self.prior, pl, pt, self._stream())
        )
        if not n_slots:
            return None
        out = torch.empty((2, n_slots), dtype=torch.float64, device=self.device)
        st = self._stream()
        _lib.check(self.lib.pcb_reduce_rows_dev(rows_l.data_ptr(), n_slots, self.S, out[0].data_ptr(), st))
        _lib.check(self.lib.pcb_reduce_rows_dev(rows_t.data_ptr(), n_slots, self.S, out[1].data_ptr(), st))
        return out


def make_job(child_off, n_children, base=-1, out=-1, flags=JOB_SCAN, slot=-1):
    return (child_off, n_children, base, out, flags, slot, 0, 0)
