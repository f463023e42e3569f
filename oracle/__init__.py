"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  Nothing under zkvm_b200/ does.  Parity status: unpinned against the reference (no
reference source or fixtures exist); pinned against RFC 9496 vectors and
libsodium 1.0.20 (tests/golden/)."""
