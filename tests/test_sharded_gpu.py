"""Sharded MSM path on a single GPU: `world` sharded contexts in one process stand in for `world`
ranks; their partial sums are concatenated in rank order (what the all-gather delivers) and every
"rank" must finish to the oracle's proof.  Bit-exact."""
import ctypes as C

import pytest

from oracle import polymath as opm, r1cs as orc
from oracle.fields import R_MOD
from oracle.merlin import MerlinFieldTranscript
from oracle.rng import StdRng as ORng, fr_rand

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3, 8])
def test_virtual_ranks_reproduce_the_oracle_proof(pmlib, world):
    from polymath_b200 import codec, sharded
    from polymath_b200.api import R1CS, _lib
    from polymath_b200.lib import check
    lib = _lib()
    sharded.bind(lib)
    consts = [11, 22, 33, 44, 55, 66, 77]
    circ = orc.MiMCDemo(None, None, consts)
    cs = orc.synthesize(circ, setup_mode=True)
    a, b, c = cs.to_matrices()
    orng = ORng.seed_from_u64(77)
    pk_or = opm.generate_proving_key(circ, orng)
    x, z = pk_or.trapdoor["x"], pk_or.trapdoor["z"]
    r1cs = R1CS(cs.num_instance_variables, cs.num_witness_variables, a, b, c)
    ctxs = []
    for rank in range(world):
        h = C.c_void_p()
        xg2, zg2 = C.create_string_buffer(192), C.create_string_buffer(192)
        check(lib.pm_setup_sharded(C.byref(r1cs.view), codec.fr_to_wire(x), codec.fr_to_wire(z), rank, world, C.byref(h), xg2, zg2))
        ctxs.append(h)
    pcs = orc.synthesize(orc.MiMCDemo(5, 6, consts), setup_mode=False)
    inst, wit = pcs.instance_assignment, pcs.witness_assignment
    trace = {}
    want = opm.create_proof_with_assignment(pk_or, inst, wit, orng, trace=trace)
    ra = codec.frs_to_wire(trace["ra"])
    # phase 1
    parts = b""
    for h in ctxs:
        check(lib.pm_ctx_set_assignment(h, codec.frs_to_wire(inst), codec.frs_to_wire(wit)))
        out = C.create_string_buffer(384)
        check(lib.pm_prove_phase1_partial(h, ra, out))
        parts += out.raw
    for h in ctxs:
        ao, co = C.create_string_buffer(96), C.create_string_buffer(96)
        check(lib.pm_prove_phase1_finish(h, parts, world, ao, co))
        assert codec.g1_from_wire(ao.raw) == want.a_g1 and codec.g1_from_wire(co.raw) == want.c_g1
    # phase 2 (replicated)
    for h in ctxs:
        ev = C.create_string_buffer(32)
        check(lib.pm_prove_phase2(h, codec.fr_to_wire(trace["x1"]), codec.fr_to_wire(trace["y1_alpha"]), ev))
        assert codec.fr_from_wire(ev.raw) == want.a_at_x1
    # phase 3
    parts = b""
    for h in ctxs:
        out = C.create_string_buffer(192)
        check(lib.pm_prove_phase3_partial(h, codec.fr_to_wire(trace["x2"]), codec.fr_to_wire(trace["c_at_x1"]), out))
        parts += out.raw
    for h in ctxs:
        do = C.create_string_buffer(96)
        check(lib.pm_prove_phase3_finish(h, parts, world, do))
        assert codec.g1_from_wire(do.raw) == want.d_g1
    assert opm.verify_proof(pk_or.vk, want, inst[1:])
    for h in ctxs:
        lib.pm_ctx_destroy(h)


@pytest.mark.parametrize("log_n,world", [(2, 2), (4, 4), (6, 8), (9, 2), (12, 4), (13, 8), (20, 8)])
@pytest.mark.parametrize("inverse", [False, True])
def test_sharded_ntt_virtual_ranks(pmlib, log_n, world, inverse):
    """pm_ntt_dist_local / pm_ntt_dist_combine with `world` virtual ranks on one GPU (the all-to-all is emulated by
    slicing): every rank must hold the interleaved slice of the full transform.  Bit-exact."""
    import random
    import torch
    from oracle import poly
    from polymath_b200 import codec, kernels, sharded
    n = 1 << log_n
    rnd = random.Random(log_n * 10 + world)
    seeds = [rnd.randrange(R_MOD) for _ in range(61)]
    x = [seeds[i % 61] * (i + 1) % R_MOD for i in range(n)]
    if log_n <= 13:
        dom = poly.Domain(n)
        full = dom.ifft(x) if inverse else dom.fft(x)
    else:
        full = kernels.ntt_fr(x, inverse=inverse)        # checked against the oracle in test_kernels_gpu.py
    ranks = [sharded.ShardedNtt(log_n, r, world) for r in range(world)]
    for r, sn in enumerate(ranks):
        sn.data.copy_(torch.frombuffer(bytearray(codec.frs_to_wire(x[r::world])), dtype=torch.uint8))
        sn.local_step(inverse)
    blk = (n // world // world) * 32
    for h, sn in enumerate(ranks):
        sn.recv.copy_(torch.cat([src.send[h * blk:(h + 1) * blk] for src in ranks]))
        sn.combine_step(inverse)
    torch.cuda.synchronize()
    for h, sn in enumerate(ranks):
        assert codec.frs_from_wire(sn.out.cpu().numpy().tobytes()) == full[h::world]


def test_collective_phases_with_a_one_rank_communicator(pmlib):
    """pm_ctx_attach_nccl + pm_prove_phase{1,3}_collective (NCCL all-gather on the device inside the phases) on a
    communicator of ONE rank: the plumbing (dlopen'd NCCL, gather buffers, per-rank decoding of the partial sums)
    must reproduce the oracle's proof bytes.  The multi-rank exchange itself is covered by bench.py --gpus N
    (proof_verified) and, for the host logic, by the gloo tests."""
    from polymath_b200 import codec, sharded
    from polymath_b200.api import R1CS, StdRng, _lib
    from polymath_b200.lib import check
    lib = _lib()
    sharded.bind(lib)
    consts = [3, 1, 4, 1, 5, 9, 2, 6]
    circ = orc.MiMCDemo(None, None, consts)
    cs = orc.synthesize(circ, setup_mode=True)
    a, b, c = cs.to_matrices()
    seed = 4242
    orng, drng = ORng.seed_from_u64(seed), StdRng.seed_from_u64(seed)
    pk_or = opm.generate_proving_key(circ, orng)
    r1cs = R1CS(cs.num_instance_variables, cs.num_witness_variables, a, b, c)
    h = C.c_void_p()
    vk = C.create_string_buffer(392)
    check(lib.pm_polymath_setup_sharded(C.byref(r1cs.view), drng._h, 0, 1, C.byref(h), vk))
    assert vk.raw == pk_or.vk.serialize_compressed()
    path = sharded.loaded_nccl_path()
    pb = path.encode() if path else None
    ident = C.create_string_buffer(128)
    check(lib.pm_nccl_unique_id(pb, ident))
    check(lib.pm_ctx_attach_nccl(h, pb, ident.raw))
    assert lib.pm_ctx_has_collective(h) == 1
    pcs = orc.synthesize(orc.MiMCDemo(7, 8, consts), setup_mode=False)
    inst, wit = pcs.instance_assignment, pcs.witness_assignment
    want = opm.create_proof_with_assignment(pk_or, inst, wit, orng)
    got = C.create_string_buffer(176)
    null_cb = C.cast(None, sharded.ALLGATHER_FN)
    # `drng` has consumed exactly the setup draws, like `orng` had before its proof
    check(lib.pm_polymath_prove_sharded(h, codec.frs_to_wire(inst), codec.frs_to_wire(wit), 1, drng._h, null_cb, None, got))
    assert got.raw == want.serialize_compressed()
    lib.pm_ctx_destroy(h)


@pytest.mark.parametrize("log_n", [10, 14, 20])
def test_resident_kernels_with_virtual_ranks(pmlib, log_n):
    """The multi-GPU data path on ONE GPU (`pm_ctx_selftest_resident`): after a proof whose bytes the oracle / the golden
    file pins, the polynomial work is replayed the way 2, 4 and 8 ranks of the sharded-resident flow run it — strided SAP
    rows (row-range SpMV), sharded transforms with the all-to-all emulated by copies between the virtual ranks, the
    ring exchange, local quotient checks and scalar assembly, and the (X - x1) division by chunk ranges with its carry
    exchange — and every slice / range must equal what the unsharded path computed.  2^20 exercises the two-level carry
    propagation inside a rank's range; the real NCCL exchange is covered by bench.py --gpus N (golden proof_check)."""
    from polymath_b200 import circuits
    from polymath_b200.api import Polymath, StdRng, _lib
    from polymath_b200.lib import check
    lib = _lib()
    lib.pm_ctx_selftest_resident.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
    n = 1 << log_n
    r1cs, inst, wit, rng = circuits.synthetic_mimc(n, seed=3)
    pk, vk_bytes = Polymath.setup(r1cs, rng)
    proof = Polymath.prove(pk, inst, wit, rng)
    assert Polymath.verify(vk_bytes, inst[1:], proof)
    if log_n <= 10:        # byte-identical to the oracle at the size it can reach
        from oracle import fast
        from oracle.poly import Domain
        import numpy as np
        # same key (exported), witness and blinding through the CPU prove
        from polymath_b200 import keydump
        import bench
        key = bench._key_from_device(pk)
        cpu = bench.CpuProver(r1cs, inst, wit, key, "device key")
        rng2 = StdRng.seed_from_u64(77)
        ra_rng = StdRng.seed_from_u64(77)
        ra = [ra_rng.fr_rand(), ra_rng.fr_rand()]
        assert cpu.prove(ra).serialize_compressed() == Polymath.prove(pk, inst, wit, rng2)
    for world in (2, 4, 8):
        if n < 8 * world * world:
            continue
        bad = C.c_uint64(123)
        check(lib.pm_ctx_selftest_resident(pk._h, world, C.byref(bad)))
        assert bad.value == 0, (log_n, world, bad.value)
    # an unsharded context only, after a complete proof
    pk.close()
