"""Host-side logic of the multi-GPU path on CPU: world-size-2 gloo process group driving the C-ABI
all-gather callback, the interleaved shard map, and the host-side addition of XYZZ partials
(`pm_host_sum_partials`) against the oracle.  No GPU needed."""
import os
import random
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from polymath_b200 import sharded
        from polymath_b200.lib import load
        lib = load()
        sharded.bind(lib)
        cb = sharded.make_allgather()
        rc = lib.pm_allgather_selftest(cb, None, rank, world)
        # a realistic payload: each rank contributes an XYZZ partial; everyone adds them on the host
        from oracle import curve
        from oracle.fields import Q_MOD
        from polymath_b200 import codec
        import ctypes as C
        k = 1000 + 7 * rank
        pt = curve.g1_mul(curve.G1_GEN, k)
        zz = pow(3 + rank, 2, Q_MOD)
        zzz = pow(3 + rank, 3, Q_MOD)            # (x*zz, y*zzz, zz, zzz) represents pt
        rec = b"".join(codec.fq_to_wire(v) for v in (pt[0] * zz % Q_MOD, pt[1] * zzz % Q_MOD, zz, zzz))
        send = C.create_string_buffer(rec, 192)
        recv = C.create_string_buffer(192 * world)
        rc2 = cb(None, C.addressof(send), 192, C.addressof(recv))
        total = sharded.host_sum_partials(recv.raw, world)
        want = curve.g1_mul(curve.G1_GEN, sum(1000 + 7 * r for r in range(world)))
        q.put((rank, rc, rc2, total == want))
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        q.put((rank, -1, -1, repr(e)))


def test_allgather_callback_over_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, rc, rc2, ok in sorted(results):
        assert rc == 0 and rc2 == 0 and ok is True, results


def test_shard_map_partitions_every_vector():
    from polymath_b200.sharded import shard_indices
    for total in (0, 1, 5, 103, 2 ** 12 + 7):
        for world in (1, 2, 3, 8):
            seen = sorted(i for r in range(world) for i in shard_indices(total, r, world))
            assert seen == list(range(total))
            sizes = [len(shard_indices(total, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_host_sum_partials_matches_oracle():
    from oracle import curve
    from oracle.fields import Q_MOD
    from polymath_b200 import codec, sharded
    rnd = random.Random(5)
    parts, ks = b"", []
    for i in range(6):
        k = rnd.randrange(1, 2 ** 64)
        ks.append(k)
        pt = curve.g1_mul(curve.G1_GEN, k)
        z = rnd.randrange(1, Q_MOD)
        zz, zzz = z * z % Q_MOD, z * z * z % Q_MOD
        parts += b"".join(codec.fq_to_wire(v) for v in (pt[0] * zz % Q_MOD, pt[1] * zzz % Q_MOD, zz, zzz))
    parts += bytes(192)                       # an infinity partial (ZZ = 0)
    assert sharded.host_sum_partials(parts, 7) == curve.g1_mul(curve.G1_GEN, sum(ks))
    # P + (-P) = infinity, doubling branch
    pt = curve.g1_mul(curve.G1_GEN, 9)
    one = lambda p: b"".join(codec.fq_to_wire(v) for v in (p[0], p[1], 1, 1))
    assert sharded.host_sum_partials(one(pt) + one(curve.g1_neg(pt)), 2) is None
    assert sharded.host_sum_partials(one(pt) + one(pt), 2) == curve.g1_mul(curve.G1_GEN, 18)


def _ntt_worker(rank, world, port, q, log_n, inverse):
    try:
        sys.path.insert(0, ROOT)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from oracle import poly
        from oracle.fields import R_MOD
        from polymath_b200 import codec, sharded
        n = 1 << log_n
        rnd = random.Random(99)
        x = [rnd.randrange(R_MOD) for _ in range(n)]
        dom, sub = poly.Domain(n), poly.Domain(n // world)
        w = pow(dom.group_gen, R_MOD - 2, R_MOD) if inverse else dom.group_gen
        # local transform of the interleaved subsequence (the GPU does this with pm_ntt_dist_local) ...
        mine = x[rank::world]
        y = sub.ifft(mine) if inverse else sub.fft(mine)
        # ... twiddled and packed in the product's send order
        order = sharded.ntt_send_order(log_n, world)
        send = torch.frombuffer(bytearray(codec.frs_to_wire([y[k2] * pow(w, rank * k2, R_MOD) % R_MOD for k2 in order])),
                                dtype=torch.uint8)
        recv = torch.empty_like(send)
        sharded.ntt_exchange(send, recv, world)
        got = codec.frs_from_wire(recv.numpy().tobytes())
        per = n // world // world
        wg = pow(w, n // world, R_MOD)
        scale = pow(world, R_MOD - 2, R_MOD) if inverse else 1
        out = [sum(got[g * per + b] * pow(wg, g * k1, R_MOD) for g in range(world)) * scale % R_MOD
               for k1 in range(world) for b in range(per)]
        full = dom.ifft(x) if inverse else dom.fft(x)
        q.put((rank, out == full[rank::world]))
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))


@pytest.mark.parametrize("log_n,inverse", [(6, False), (7, True)])
def test_sharded_ntt_exchange_over_gloo_world2(log_n, inverse):
    """The decomposition behind pm_ntt_dist_local / all-to-all / pm_ntt_dist_combine, on CPU: oracle transforms for
    the local step, the product's send order and exchange, checked against the oracle's full transform."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ntt_worker, args=(r, world, port, q, log_n, inverse)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok in sorted(results):
        assert ok is True, results
