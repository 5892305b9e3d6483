"""Streaming file plane of library generation.

The reference writes one ``library__pNNNNNN.npz`` per sample on the critical path of each MPI rank
(``holodeck/librarian/gen_lib.py:236-337``: ``np.savez`` of ~1 MB, zip + CRC, right after ``run_model``) and, when
all ranks are done, re-reads every file into the combined arrays (``combine.py:86-272, 366-441``).  On a B200 a
sample takes ~10 ms, so that file plane *is* the critical path (round 1: 615 -> 350 samples/s).  Here

* the combined layout exists from the start: one ``.npy`` per dataset in ``<output>/library_store/`` with the
  reference's combined shapes -- ``gwb (S,F,R)``, ``hc_ss (S,F,R,L)``, ``hc_bg (S,F,R)``, ``sspar (S,4,F,R,L)``,
  ``bgpar (S,7,F,R)`` (``combine.py:140-175``) -- plus ``status (S,) uint8`` (0 = to do, 1 = done, 2 = failed:
  what the existence / ``fail`` key of a per-sample file encodes in the reference) and ``failures.log``.  Every rank
  opens the same files and writes only the rows of its own samples (one positioned write per row): no merge pass,
  no re-read; readers (resume check, combine) memory-map them;
* ``run_model`` leaves its products on the device; :class:`AsyncSampleWriter` copies them into a ring of PINNED
  staging slots on a side stream and a background thread writes each finished slot into the store, so neither
  the device->host copy nor the file write sits between two samples;
* ``combine.sam_lib_combine`` turns the store into the reference's single-file library when asked.
"""
import json
import queue
import threading
from pathlib import Path

import numpy as np

DIRNAME_LIBRARY_STORE = "library_store"
STATUS_TODO, STATUS_DONE, STATUS_FAIL = 0, 1, 2
_LAYOUT_FNAME = "layout.json"


def dataset_shapes(nsamples, nfreqs, nreals, nloudest, gwb_flag, ss_flag, params_flag):
    """Combined-file shapes of the datasets a run produces (``combine.py:140-175``)."""
    S, F, R = int(nsamples), int(nfreqs), int(nreals)
    L = int(nloudest) if ss_flag else 1                      # lib_tools.py:790 (`nloudest if singles_flag else 1`)
    shapes = {}
    if gwb_flag:
        shapes["gwb"] = (S, F, R)
    if ss_flag:
        shapes["hc_ss"] = (S, F, R, L)
        shapes["hc_bg"] = (S, F, R)
    if params_flag:
        shapes["sspar"] = (S, 4, F, R, L)
        shapes["bgpar"] = (S, 7, F, R)
    return shapes


class LibraryStore:
    """The memory-mapped combined layout of one library (``<output>/library_store``)."""

    def __init__(self, path, layout, maps):
        self.path = Path(path)
        self.layout = layout
        self.maps = maps                  # name -> np.memmap (S, ...)
        self.status = maps["status"]
        self.fobs = None
        self._fds = {}                    # name -> (file descriptor, byte offset of row 0, bytes per row): see `put`

    @classmethod
    def _dir(cls, output):
        return Path(output).joinpath(DIRNAME_LIBRARY_STORE)

    @classmethod
    def exists(cls, output):
        return cls._dir(output).joinpath(_LAYOUT_FNAME).exists()

    @classmethod
    def create(cls, output, nsamples, nfreqs, nreals, nloudest, gwb_flag, ss_flag, params_flag, fobs_cents, fobs_edges):
        """Allocate the files (rank 0, before the barrier).  An existing store with the same layout is kept (resume);
        a different layout is an error -- a resumed run must not silently mix shapes."""
        path = cls._dir(output)
        path.mkdir(parents=True, exist_ok=True)
        shapes = dataset_shapes(nsamples, nfreqs, nreals, nloudest, gwb_flag, ss_flag, params_flag)
        layout = dict(nsamples=int(nsamples), nfreqs=int(nfreqs), nreals=int(nreals), nloudest=int(nloudest),
                      gwb_flag=bool(gwb_flag), ss_flag=bool(ss_flag), params_flag=bool(params_flag),
                      datasets={kk: list(vv) for kk, vv in shapes.items()})
        lfile = path.joinpath(_LAYOUT_FNAME)
        if lfile.exists():
            old = json.loads(lfile.read_text())
            if old != layout:
                raise RuntimeError(f"{path} holds a library with a different layout ({old}) than requested ({layout}); "
                                   "use a new output directory or `--recreate`")
            return cls.open(output)
        for name, shape in shapes.items():
            np.lib.format.open_memmap(path.joinpath(f"{name}.npy"), mode="w+", dtype=np.float64, shape=shape).flush()
        np.lib.format.open_memmap(path.joinpath("status.npy"), mode="w+", dtype=np.uint8, shape=(int(nsamples),)).flush()
        np.save(path.joinpath("fobs_cents.npy"), np.asarray(fobs_cents, dtype=np.float64))
        np.save(path.joinpath("fobs_edges.npy"), np.asarray(fobs_edges, dtype=np.float64))
        lfile.write_text(json.dumps(layout))
        return cls.open(output)

    @classmethod
    def open(cls, output, mode="r+"):
        path = cls._dir(output)
        layout = json.loads(path.joinpath(_LAYOUT_FNAME).read_text())
        maps = {name: np.load(path.joinpath(f"{name}.npy"), mmap_mode=mode) for name in layout["datasets"]}
        maps["status"] = np.load(path.joinpath("status.npy"), mmap_mode=mode)
        store = cls(path, layout, maps)
        store.fobs = (np.load(path.joinpath("fobs_cents.npy")), np.load(path.joinpath("fobs_edges.npy")))
        return store

    def reset(self):
        """`--recreate`: every sample is to do again."""
        self.status[:] = STATUS_TODO
        self.status.flush()

    def is_done(self, pnum):
        return int(self.status[pnum]) == STATUS_DONE

    def _row_file(self, name):
        """(fd, offset of row 0, row bytes) of a dataset's ``.npy`` file, opened once per process."""
        ent = self._fds.get(name)
        if ent is None:
            import os
            mm = self.maps[name]
            row = int(mm.dtype.itemsize * (np.prod(mm.shape[1:]) if mm.ndim > 1 else 1))
            ent = (os.open(self.path.joinpath(f"{name}.npy"), os.O_RDWR), int(mm.offset), row)
            self._fds[name] = ent
        return ent

    def _write_row(self, name, pnum, arr):
        """One positioned write per row (``pwrite`` straight into the page cache).  Assigning into the shared memory
        map instead takes a page fault per 4 KB page -- 270 per sample -- and with eight ranks mapping the same files
        those faults serialise in the kernel (measured: 8-GPU library generation at 73 % of 8 x one GPU)."""
        import os
        fd, off, row = self._row_file(name)
        buf = np.ascontiguousarray(arr, dtype=self.maps[name].dtype)
        assert buf.nbytes == row, (name, buf.shape, self.maps[name].shape)
        view = memoryview(buf).cast("B")
        done = 0
        while done < row:
            done += os.pwrite(fd, view[done:], off + int(pnum) * row + done)

    def put(self, pnum, data):
        """Write the products of sample ``pnum`` (host arrays keyed like ``run_model``'s dict)."""
        for name in self.layout["datasets"]:
            self._write_row(name, pnum, data[name])
        self._write_row("status", pnum, np.array([STATUS_DONE], dtype=np.uint8))

    def put_failure(self, pnum, message):
        """A failed sample: NaN rows, as ``combine.py:404-417`` produces from a ``fail`` file."""
        for name in self.layout["datasets"]:
            self._write_row(name, pnum, np.full(self.maps[name].shape[1:], np.nan))
        self._write_row("status", pnum, np.array([STATUS_FAIL], dtype=np.uint8))
        with open(self.path.joinpath("failures.log"), "a") as ff:
            ff.write(f"{int(pnum)}\t{message}\n")

    def flush(self):
        """Nothing to do: rows were written with `pwrite` (page cache; as durable as the reference's `np.savez`)."""


class AsyncSampleWriter:
    """Device -> pinned slot -> memory map, off the sample loop's critical path.

    ``submit(pnum, tensors)`` enqueues asynchronous device->host copies of the sample's CUDA tensors into one of
    ``nslots`` pinned staging slots on a dedicated copy stream (ordered after the producing stream by an event) and
    returns immediately; a daemon thread waits for the slot's copy event and writes the slot into the store.  With no
    CUDA device (CPU tests, gloo) the same path runs synchronously on host arrays.
    """

    def __init__(self, store, nslots=4, also_npz=None):
        self.store = store
        self.nslots = int(nslots)
        self.also_npz = also_npz          # optional callable(pnum, host_dict): the reference's per-sample file
        self._free = queue.Queue()
        self._work = queue.Queue()
        self._slots = []
        self._error = None
        self._cuda = False
        try:
            import torch
            self._cuda = torch.cuda.is_available()
        except Exception:   # noqa: BLE001
            pass
        if self._cuda:
            import torch
            self._copy_stream = torch.cuda.Stream()
            shapes = {kk: tuple(vv[1:]) for kk, vv in store.layout["datasets"].items()}
            for ii in range(self.nslots):
                self._slots.append({kk: torch.empty(ss, dtype=torch.float64).pin_memory() for kk, ss in shapes.items()})
                self._free.put(ii)
        self._thread = threading.Thread(target=self._drain, daemon=True)
        self._thread.start()

    def submit(self, pnum, data):
        """``data``: dict name -> CUDA tensor (or host array) for every dataset of the store's layout."""
        if self._error is not None:
            raise self._error
        names = list(self.store.layout["datasets"])
        if not self._cuda or not all(hasattr(data[nn], "is_cuda") and data[nn].is_cuda for nn in names):
            host = {nn: (data[nn].cpu().numpy() if hasattr(data[nn], "cpu") else np.asarray(data[nn])) for nn in names}
            self._work.put((int(pnum), host, None, None))
            return
        import torch
        slot = self._free.get()                 # blocks only when the disk is nslots samples behind
        ready = torch.cuda.Event()
        ready.record()                          # products are complete on the producing stream
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            for nn in names:
                self._slots[slot][nn].copy_(data[nn], non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        # keep the device tensors alive until the copy has run
        self._work.put((int(pnum), self._slots[slot], done, (slot, [data[nn] for nn in names])))

    def submit_failure(self, pnum, message):
        self._work.put((int(pnum), None, None, str(message)))

    def _drain(self):
        while True:
            item = self._work.get()
            if item is None:
                return
            pnum, host, done, extra = item
            try:
                if host is None:
                    self.store.put_failure(pnum, extra)
                    continue
                if done is not None:
                    done.synchronize()
                    host = {nn: tt.numpy() for nn, tt in host.items()}
                self.store.put(pnum, host)
                if self.also_npz is not None:
                    self.also_npz(pnum, host)
            except Exception as err:   # noqa: BLE001
                self._error = err
            finally:
                if done is not None:
                    self._free.put(extra[0])

    def close(self):
        """Wait until every submitted sample is in the store, then flush the maps."""
        self._work.put(None)
        self._thread.join()
        self.store.flush()
        if self._error is not None:
            raise self._error
