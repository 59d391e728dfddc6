"""Multi-GPU proving: one process per GPU, MSMs split by point range.

Every rank holds the points g = k*world + rank of the proving key, runs the (small) polynomial work
redundantly and its share of each MSM; the per-rank partial sums are exchanged with one all-gather per
MSM phase and added on the host by every rank (group addition is not an NCCL reduction).  Two transports:

* "nccl" (default on an NCCL process group): the library owns an NCCL communicator (`attach_nccl` ->
  `pm_ctx_attach_nccl`, NCCL bound with dlopen to the libnccl the process already uses) and the phases
  all-gather the raw per-window sums on the DEVICE, stream-ordered right behind the MSM kernels
  (`pm_prove_phase{1,3}_collective`): no host round trip between compute and collective;
* "callback": the host-side 192-byte partial sums travel through a `torch.distributed` all-gather
  supplied as a C callback (gloo in the CPU tests).

The protocol flow itself stays in the C++ host mirror (`pm_polymath_prove_sharded`).
"""
import ctypes as C
import os
import time

from . import codec
from .api import _lib, R1CS, StdRng, ProvingKey
from .lib import check, load

ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


def shard_indices(total: int, rank: int, world: int):
    """Global indices owned by `rank` under the interleaved split (mirrors ProverCtx::local_count)."""
    return range(rank, total, world)


def make_allgather(group=None, device=None):
    """Build the `pm_allgather_fn` callback on top of torch.distributed."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")

    trace = os.environ.get("PM_TRACE_ALLGATHER")
    stats = {"calls": 0, "s": 0.0}
    bufs = {}     # nbytes -> (pinned send, device send, device recv, pinned recv): no allocation on the hot path

    def _buffers(nbytes):
        b = bufs.get(nbytes)
        if b is None:
            pin = device.type == "cuda"
            hs = torch.empty(nbytes, dtype=torch.uint8, pin_memory=pin)
            hr = torch.empty(world * nbytes, dtype=torch.uint8, pin_memory=pin)
            b = bufs[nbytes] = (hs, torch.empty(nbytes, dtype=torch.uint8, device=device),
                                torch.empty(world * nbytes, dtype=torch.uint8, device=device), hr)
        return b

    def _cb(_user, send, nbytes, recv):
        t0 = time.perf_counter()
        try:
            hs, ds, dr, hr = _buffers(nbytes)
            C.memmove(hs.data_ptr(), send, nbytes)
            ds.copy_(hs, non_blocking=True)
            dist.all_gather_into_tensor(dr, ds, group=group)
            hr.copy_(dr, non_blocking=True)
            if device.type == "cuda":
                torch.cuda.current_stream(device).synchronize()
            C.memmove(recv, hr.data_ptr(), world * nbytes)
            if trace:
                stats["calls"] += 1
                stats["s"] += time.perf_counter() - t0
                if stats["calls"] % 16 == 0 and dist.get_rank(group) == 0:
                    print("[allgather] %d calls, %.3f ms avg" % (stats["calls"], stats["s"] / stats["calls"] * 1e3), flush=True)
            return 0
        except Exception:  # the C side turns a non-zero return into PM_ERR_STATE
            import traceback
            traceback.print_exc()
            return 1

    return ALLGATHER_FN(_cb)


def bind(lib):
    if getattr(lib, "_sharded_bound", False):
        return
    vp, u8p = C.c_void_p, C.c_char_p
    from .api import R1CSView, PKView
    lib.pm_polymath_setup_sharded.argtypes = [C.POINTER(R1CSView), vp, C.c_int, C.c_int, C.POINTER(vp), u8p]
    lib.pm_polymath_prove_sharded.argtypes = [vp, u8p, u8p, C.c_int, vp, ALLGATHER_FN, vp, u8p]
    lib.pm_allgather_selftest.argtypes = [ALLGATHER_FN, vp, C.c_int, C.c_int]
    lib.pm_setup_sharded.argtypes = [C.POINTER(R1CSView), u8p, u8p, C.c_int, C.c_int, C.POINTER(vp), u8p, u8p]
    lib.pm_ctx_create_sharded.argtypes = [C.POINTER(PKView), C.c_int, C.c_int, C.POINTER(vp)]
    lib.pm_prove_phase1_partial.argtypes = [vp, u8p, u8p]
    lib.pm_prove_phase1_finish.argtypes = [vp, u8p, C.c_int, u8p, u8p]
    lib.pm_prove_phase3_partial.argtypes = [vp, u8p, u8p, u8p]
    lib.pm_prove_phase3_finish.argtypes = [vp, u8p, C.c_int, u8p]
    lib.pm_host_sum_partials.argtypes = [u8p, C.c_int, C.c_size_t, u8p]
    lib.pm_nccl_unique_id.argtypes = [u8p, u8p]
    lib.pm_ctx_attach_nccl.argtypes = [vp, u8p, u8p]
    lib.pm_ctx_has_collective.argtypes = [vp]
    lib.pm_prove_phase1_collective.argtypes = [vp, u8p, u8p, u8p]
    lib.pm_prove_phase3_collective.argtypes = [vp, u8p, u8p, u8p]
    lib.pm_ntt_dist_local.argtypes = [vp, vp, C.c_uint, C.c_uint, C.c_uint, C.c_int, vp]
    lib.pm_ntt_dist_combine.argtypes = [vp, vp, C.c_uint, C.c_uint, C.c_int, vp]
    lib._sharded_bound = True


def loaded_nccl_path():
    """Path of the libnccl this process already has mapped (torch's bundled one), so that one NCCL serves the process."""
    try:
        with open("/proc/self/maps") as fh:
            for line in fh:
                if "libnccl.so" in line:
                    return line.split()[-1]
    except OSError:
        pass
    try:
        import nvidia.nccl as n
        cand = os.path.join(list(n.__path__)[0], "lib", "libnccl.so.2")
        if os.path.exists(cand):
            return cand
    except Exception:
        pass
    return None


def attach_nccl(lib, ctx_handle, rank: int, world: int, group=None):
    """Give a sharded context its own NCCL communicator (`pm_ctx_attach_nccl`): rank 0 creates the id, the bytes travel
    through the torch process group, every rank joins.  Afterwards the phases all-gather their partial sums on the
    device (`pm_prove_phase{1,3}_collective`) and `pm_polymath_prove_sharded` needs no callback."""
    import torch
    import torch.distributed as dist
    path = loaded_nccl_path()
    pb = path.encode() if path else None
    buf = C.create_string_buffer(128)
    if rank == 0:
        check(lib.pm_nccl_unique_id(pb, buf))
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(dev)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ident = bytes(t.cpu().numpy().tobytes())
    check(lib.pm_ctx_attach_nccl(ctx_handle, pb, ident))


def ntt_send_order(log_n: int, world: int):
    """k2 index (into the rank's local (N/world)-point transform) of every position of the send buffer of the
    sharded NTT: destination-major, block h = the k2 = h (mod world) in increasing order (mirrors
    k_dist_twiddle_pack in csrc/ntt.cu)."""
    n2 = (1 << log_n) // world
    per = n2 // world
    return [h + world * b for h in range(world) for b in range(per)]


def ntt_exchange(send, recv, world: int, group=None):
    """The one all-to-all of the sharded NTT: equal blocks, block h of `send` goes to rank h."""
    import torch.distributed as dist
    if world == 1:
        recv.copy_(send)
    else:
        dist.all_to_all_single(recv, send, group=group)


class ShardedNtt:
    """Fr NTT of size 2^log_n sharded over the ranks of a process group (four-step, ONE all-to-all).

    Rank g holds the interleaved subsequence x[j*world + g] as a CUDA uint8 tensor of 32-byte Montgomery elements
    (`self.data`).  `run()` = local (N/G)-point transform + twiddle/pack (`pm_ntt_dist_local`), one
    `all_to_all_single` of equal blocks over NCCL/NVLink, G-point combine (`pm_ntt_dist_combine`).  The result
    `self.out` holds X[k] for k = rank (mod world) at local index (k - rank)/world.  Everything is enqueued on
    torch's current stream.  `exchange` can be replaced (the single-GPU tests emulate the ranks in one process).
    """

    def __init__(self, log_n: int, rank: int, world: int, group=None, device=None):
        import torch
        self.lib = _lib()
        bind(self.lib)
        assert world in (1, 2, 4, 8) and log_n >= 2 * (world.bit_length() - 1)
        self.log_n, self.rank, self.world, self.group = log_n, rank, world, group
        self.log_g = world.bit_length() - 1
        self.local_elems = (1 << log_n) // world
        device = device or torch.device("cuda", torch.cuda.current_device())
        nbytes = self.local_elems * 32
        self.data = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.send = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.recv = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.out = torch.empty(nbytes, dtype=torch.uint8, device=device)

    def local_step(self, inverse=False):
        import torch
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(self.lib.pm_ntt_dist_local(C.c_void_p(self.data.data_ptr()), C.c_void_p(self.send.data_ptr()), self.log_n,
                                         self.log_g, self.rank, 1 if inverse else 0, stream))

    def exchange(self):
        ntt_exchange(self.send, self.recv, self.world, self.group)

    def combine_step(self, inverse=False):
        import torch
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(self.lib.pm_ntt_dist_combine(C.c_void_p(self.recv.data_ptr()), C.c_void_p(self.out.data_ptr()), self.log_n,
                                           self.log_g, 1 if inverse else 0, stream))

    def run(self, inverse=False):
        self.local_step(inverse)
        self.exchange()
        self.combine_step(inverse)
        return self.out


def host_sum_partials(parts: bytes, count: int, stride: int = 192):
    """Host-only: canonical affine sum of XYZZ partial records (no GPU needed)."""
    lib = load()
    bind(lib)
    out = C.create_string_buffer(96)
    check(lib.pm_host_sum_partials(parts, count, stride, out))
    return codec.g1_from_wire(out.raw)


class ShardedProver:
    """`Polymath::setup` + `prove` across `world` processes (one GPU each)."""

    def __init__(self, r1cs: R1CS, rng: StdRng, rank: int, world: int, group=None, collective="auto"):
        """collective: "nccl" = the library's own NCCL all-gather inside the phases (device buffers, stream-ordered);
        "callback" = host partial sums exchanged by a torch.distributed callback (works on gloo: the CPU tests);
        "auto" = nccl when the process group is NCCL, else callback."""
        import torch.distributed as dist
        self.lib = _lib()
        bind(self.lib)
        self.rank, self.world = rank, world
        if collective == "auto":
            collective = os.environ.get("PM_SHARDED_COLLECTIVE") or \
                ("nccl" if world > 1 and dist.is_initialized() and dist.get_backend(group) == "nccl" else "callback")
        self.collective = collective
        self._cb = make_allgather(group) if (world > 1 and collective == "callback") else ALLGATHER_FN(lambda *_: 1)
        h = C.c_void_p()
        vk = C.create_string_buffer(392)
        check(self.lib.pm_polymath_setup_sharded(C.byref(r1cs.view), rng._h, rank, world, C.byref(h), vk))
        self.pk = ProvingKey(h, vk.raw)
        self.pk._r1cs = r1cs
        self.vk_bytes = vk.raw
        self._proof = C.create_string_buffer(176)
        if collective == "nccl":
            attach_nccl(self.lib, self.pk._h, rank, world, group)
            self._cb = C.cast(None, ALLGATHER_FN)      # NULL: the phases run their own collective

    def set_assignment(self, instance_ptr, witness_ptr):
        check(self.lib.pm_ctx_set_assignment(self.pk._h, instance_ptr, witness_ptr))

    def prove(self, instance_wire, witness_ptr, rng: StdRng) -> bytes:
        check(self.lib.pm_polymath_prove_sharded(self.pk._h, instance_wire, witness_ptr, 1, rng._h, self._cb, None, self._proof))
        return self._proof.raw

    def prove_resident(self, instance_wire, rng: StdRng) -> bytes:
        check(self.lib.pm_polymath_prove_sharded(self.pk._h, instance_wire, None, 0, rng._h, self._cb, None, self._proof))
        return self._proof.raw

    def phase_ms(self):
        return self.pk.phase_ms()

    def close(self):
        self.pk.close()
