"""CPU oracle for the Polymath proving hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and
there only as the checker (or as the timed CPU baseline), never on the CUDA
product path.  The product path (``polymath_b200``) fails loudly when its CUDA
library is missing; it never falls back to this package.

PARITY UNPINNED: the reference (sigma0-dev/polymath, pure Rust over un-vendored
arkworks / merlin / rand crates) cannot be compiled or executed in this
environment (no cargo/rustc, no network) and its own tests hold no golden
vectors.  This oracle restates the reference's algorithm
from its source (cited file:line everywhere) and the published algorithms of
its dependencies (arkworks 0.4.x, merlin 3.0, rand 0.8 StdRng).  It is pinned
only by: public curve/field constants, the zcash-format generator encoding,
the merlin "test protocol" vector, RFC 7539's ChaCha block vector, pairing
bilinearity, naive-vs-fast cross checks and ``verify(prove(..)) == true`` on the
reference's own test circuits (tests/dummy.rs, tests/mimc.rs).
"""
