"""Concurrent use of the C ABI: several host threads, each on its own stream, through the entry points that
share process-wide state (completion tickets of ga_nn_distance_fwd_bwd, launch counters, tuning knobs).
Results must equal the serial ones bit for bit."""
import ctypes
import threading

import numpy as np
import pytest
import torch

from util import cloud

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
p = ctypes.c_void_p


def _bufs(b, n, seed):
    a = torch.from_numpy(cloud(seed, (b, n, 3))).to(DEV)
    c = torch.from_numpy(cloud(seed + 50, (b, n, 3))).to(DEV)
    g = torch.full((b, n), 1.0 / n, device=DEV)
    out = dict(d1=torch.empty(b, n, device=DEV), i1=torch.empty(b, n, dtype=torch.int32, device=DEV),
               d2=torch.empty(b, n, device=DEV), i2=torch.empty(b, n, dtype=torch.int32, device=DEV),
               o1=torch.empty(b, n, 3, device=DEV), o2=torch.empty(b, n, 3, device=DEV))
    return a, c, g, out


def _one_call(lib, a, c, g, o, stream):
    from geometric_adv_b200 import _lib
    b, n, _ = a.shape
    _lib.check(lib.ga_nn_distance_fwd_bwd(b, n, n, p(a.data_ptr()), p(c.data_ptr()), p(g.data_ptr()), p(g.data_ptr()),
                                          p(o["d1"].data_ptr()), p(o["i1"].data_ptr()), p(o["d2"].data_ptr()),
                                          p(o["i2"].data_ptr()), p(o["o1"].data_ptr()), p(o["o2"].data_ptr()), 0,
                                          p(stream.cuda_stream)))


def _two_calls(lib, a, c, g, o, stream):
    from geometric_adv_b200 import _lib
    b, n, _ = a.shape
    _lib.check(lib.ga_nn_distance_fwd(b, n, n, p(a.data_ptr()), p(c.data_ptr()), p(o["d1"].data_ptr()),
                                      p(o["i1"].data_ptr()), p(o["d2"].data_ptr()), p(o["i2"].data_ptr()), 0,
                                      p(stream.cuda_stream)))
    _lib.check(lib.ga_nn_distance_bwd(b, n, n, p(a.data_ptr()), p(c.data_ptr()), p(g.data_ptr()), p(o["i1"].data_ptr()),
                                      p(g.data_ptr()), p(o["i2"].data_ptr()), p(o["o1"].data_ptr()), p(o["o2"].data_ptr()),
                                      p(stream.cuda_stream)))


@pytest.mark.parametrize("shape", [(50, 2048), (12, 1024)])
def test_threads_and_streams_equal_serial(ga, shape):
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    b, n = shape
    nthreads, reps = 4, 12
    work = [_bufs(b, n, 10 * t) for t in range(nthreads)]
    serial = []
    s0 = torch.cuda.Stream()
    for a, c, g, o in work:
        _one_call(lib, a, c, g, o, s0)
        s0.synchronize()
        serial.append({k: v.clone() for k, v in o.items()})
    errors = []

    def worker(t):
        try:
            torch.cuda.set_device(0)
            a, c, g, o = work[t]
            st = torch.cuda.Stream()
            for r in range(reps):
                for v in o.values():
                    v.zero_()
                torch.cuda.current_stream().synchronize()
                (_one_call if (r + t) % 2 == 0 else _two_calls)(lib, a, c, g, o, st)
                if t == 0:
                    lib.ga_set_tuning(18, 1)  # a knob rewritten (to its default) while others are inside the library
                st.synchronize()
                for k, v in o.items():
                    if not torch.equal(v, serial[t][k]):
                        errors.append("thread %d rep %d: %s differs" % (t, r, k))
                        return
        except Exception as e:  # noqa: BLE001
            errors.append("thread %d: %r" % (t, e))

    th = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errors, errors


def test_stale_ticket_is_not_reused(ga):
    """One-call entry arms a ticket record; a later plain backward on the same buffers after a DIFFERENT forward
    must not start early on the stale record: results equal the reference order anyway."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    a, c, g, o = _bufs(20, 2048, 3)
    st = torch.cuda.Stream()
    _one_call(lib, a, c, g, o, st)
    st.synchronize()
    first = {k: v.clone() for k, v in o.items()}
    a2 = torch.from_numpy(cloud(99, (20, 2048, 3))).to(DEV)
    # plain forward on other inputs into the SAME idx buffers, then a plain backward
    _two_calls(lib, a2, c, g, o, st)
    st.synchronize()
    want = ga.nn_distance_grad(a2, c, g, o["i1"], g, o["i2"])
    assert torch.equal(o["o1"], want[0]) and torch.equal(o["o2"], want[1])
    assert not torch.equal(o["i1"], first["i1"])
