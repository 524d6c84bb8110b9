"""CPU oracle for the RandLA-Net hot path (TEST INFRASTRUCTURE ONLY).

Nothing under ``myria3d_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs use it, and there only as the checker / the timed CPU reference.

PARITY UNPINNED: the reference (IGNF/myria3d) ships no numerical golden
vectors for this path and its third-party arithmetic (torch_geometric 2.4,
torch_cluster, torch_scatter) cannot be installed here; the oracle restates
those libraries' published semantics (see ``randla_oracle.py``) and is pinned
only structurally (strict load of the shipped checkpoint, reference shape
tests) plus cross-checks against independent formulations (scipy cKDTree,
``torch.nn.functional``).
"""
