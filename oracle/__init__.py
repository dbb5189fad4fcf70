"""ORACLE -- test infrastructure only.

CPU restatement of the reference's per-slot inference-and-decode hot path (functional torch fp32 on
the host).  It is the checker the CUDA engine is compared with; it is never imported by
genesis_b200/.  Allowed importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline leg and
--impl reference).

Parity pinning: the reference holds NO golden vectors or tests for this path,
so the oracle is pinned against outputs of the reference itself, imported in the build container
from /root/reference by oracle/make_golden.py, and committed as tests/golden/*.npz.
"""
