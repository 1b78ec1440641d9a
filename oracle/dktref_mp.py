"""The reference on SEVERAL ranks inside one container: oracle/_ref/libdktref_mp_morton.so is the reference's
own sources built over oracle/shim_mp (a multi-process MPI stand-in: fork()ed ranks, shared-memory messages).
TEST INFRASTRUCTURE: only tests/ and bench.py's reference legs may import this module.

    results = run(nranks, job, *args)      # job(rank, nranks, R, *args) runs in every rank; R = dktref.Reference-like session

Every rank is a fork of the calling process taken AFTER the shared world exists; results come back pickled through pipes."""
import ctypes as C
import os
import pickle
import signal
import sys
import time

import numpy as np

import dktref

SFC = "mp_morton"


def available():
    return dktref.available(SFC)


def _world_lib():
    L = dktref._lib(SFC)
    L.dktmp_world_create.argtypes = [C.c_int, C.c_size_t]
    L.dktmp_set_rank.argtypes = [C.c_int]
    L.dktmp_bytes_sent.restype = C.c_long
    L.dktref_da_local_info.argtypes = [C.c_void_p, C.c_void_p]
    return L


def local_info(da):
    """(local nodes, local begin, total = ghosted nodes, ranks, rank, global nodes) of a RefDA."""
    out = np.zeros(6, dtype=np.int64)
    da.ref.L.dktref_da_local_info(da.h, out.ctypes.data_as(C.c_void_p))
    return tuple(int(v) for v in out)


def run(nranks, job, *args, arena_bytes=1 << 31, timeout=600):
    L = _world_lib()
    if L.dktmp_world_create(nranks, arena_bytes):
        raise RuntimeError("dktmp_world_create failed")
    pipes, pids = [], []
    for r in range(nranks):
        rd, wr = os.pipe()
        pid = os.fork()
        if pid == 0:
            code = 0
            try:
                os.close(rd)
                L.dktmp_set_rank(r)
                res = job(r, nranks, *args)
                with os.fdopen(wr, "wb") as f:
                    pickle.dump(res, f)
            except BaseException as e:  # noqa: BLE001 - the child must never return into the parent's stack
                sys.stderr.write("[dktref_mp rank %d] %r\n" % (r, e))
                code = 1
            os._exit(code)
        os.close(wr)
        pipes.append(rd)
        pids.append(pid)
    out, t0 = [], time.time()
    try:
        for r, rd in enumerate(pipes):
            with os.fdopen(rd, "rb") as f:
                data = f.read()
            _, status = os.waitpid(pids[r], 0)
            if status != 0 or not data:
                raise RuntimeError("reference rank %d failed (status %d)" % (r, status))
            out.append(pickle.loads(data))
            if time.time() - t0 > timeout:
                raise RuntimeError("timeout")
    finally:
        for p in pids:
            try:
                os.kill(p, signal.SIGKILL)
            except ProcessLookupError:
                pass
    return out


def session(dim, max_depth):
    """A dktref.Reference bound to the multi-process library (call inside a job)."""
    return dktref.Reference(dim, max_depth, SFC)
