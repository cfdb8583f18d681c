"""Model check of the peer-memory halo protocol (csrc/comm.cu: k_peer_push, csrc/solver.cu:
exchange_halo) on CPU threads: every "rank" owns a slab of two twin arrays with halo cells, a
"pass" reads one twin (own cells + halos) and writes the own cells of the other, an "exchange"
stores the new boundary cells straight into the neighbours' halo cells, increments the neighbours'
arrival counters and waits until its own counters show one more arrival per neighbour -- the same
counters-only-increase scheme the kernels use.  Random delays shake the interleavings; the result
must equal the sequential computation, which is what the argument in exchange_halo's comment
claims (consecutive exchanges target different arrays, a pass never reads the halo of the array it
produces, a rank cannot run more than one exchange ahead of a neighbour).

This exercises the protocol, not the CUDA kernels: those are checked bit for bit against the
single-GPU result on 2, 4 and 8 GPUs in tests/test_gpu_sharded.py."""
import random
import threading
import time

import numpy as np
import pytest

HALO = 2


def sequential(u0, steps):
    u = u0.copy()
    for _ in range(steps):
        p = np.pad(u, 1)                              # zero boundary
        u = 0.25 * p[:-2] + 0.5 * p[1:-1] + 0.25 * p[2:] + 1.0
    return u


class Rank:
    def __init__(self, r, P, own):
        self.r, self.P = r, P
        n = len(own)
        self.n = n
        # twin arrays: [halo_up (HALO) | own (n) | halo_down (HALO)]
        self.a = [np.zeros(n + 2 * HALO), np.zeros(n + 2 * HALO)]
        self.a[0][HALO:HALO + n] = own
        self.arrivals = [0, 0]                        # from rank-1, from rank+1 (written by the neighbours)
        self.consumed = [0, 0]
        self.lock = threading.Lock()                  # stands for the atomicity of atomicAdd_system


def run_rank(me, ranks, steps, rng, errors):
    try:
        cur = 0
        up = ranks[me.r - 1] if me.r > 0 else None
        dn = ranks[me.r + 1] if me.r < me.P - 1 else None

        def exchange(which):
            n = me.n
            arr = me.a[which]
            time.sleep(rng.random() * 2e-4)
            if up is not None:                        # my first rows -> its lower halo, then raise its counter
                up.a[which][HALO + up.n:HALO + up.n + HALO] = arr[HALO:2 * HALO]
                with up.lock:
                    up.arrivals[1] += 1
            if dn is not None:
                dn.a[which][0:HALO] = arr[n:n + HALO]
                with dn.lock:
                    dn.arrivals[0] += 1
            for d, nb in ((0, up), (1, dn)):
                if nb is None:
                    continue
                want = me.consumed[d] + 1
                t0 = time.time()
                while me.arrivals[d] - want < 0:
                    if time.time() - t0 > 20:
                        raise RuntimeError(f"rank {me.r}: deadlock waiting for neighbour {d}")
                    time.sleep(0)
                me.consumed[d] = want

        exchange(cur)                                 # like form_rhs / set_fields: halos of the input valid
        for _ in range(steps):
            src, dst = me.a[cur], me.a[1 - cur]
            time.sleep(rng.random() * 2e-4)
            lo = HALO - 1
            p = src[lo:HALO + me.n + 1].copy()
            if up is None:
                p[0] = 0.0
            if dn is None:
                p[-1] = 0.0
            dst[HALO:HALO + me.n] = 0.25 * p[:-2] + 0.5 * p[1:-1] + 0.25 * p[2:] + 1.0
            cur = 1 - cur
            exchange(cur)
        me.result = me.a[cur][HALO:HALO + me.n].copy()
    except Exception as e:                            # noqa: BLE001 - reported by the main thread
        errors.append(e)


@pytest.mark.parametrize("P,seed", [(2, 0), (4, 1), (8, 2)])
def test_counter_protocol_matches_sequential(P, seed):
    rng = random.Random(seed)
    per, steps = 6, 60
    u0 = np.random.default_rng(seed).standard_normal(P * per)
    ranks = [Rank(r, P, u0[r * per:(r + 1) * per]) for r in range(P)]
    errors = []
    threads = [threading.Thread(target=run_rank, args=(rk, ranks, steps, random.Random(rng.random()), errors)) for rk in ranks]
    for t in threads:
        t.start()
    for t in threads:
        t.join(60)
    assert not errors, errors
    got = np.concatenate([rk.result for rk in ranks])
    assert np.array_equal(got, sequential(u0, steps))
