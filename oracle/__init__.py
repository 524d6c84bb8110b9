"""CPU oracle for the RandLA-Net hot path (TEST INFRASTRUCTURE ONLY).

Nothing under ``myria3d_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs use it, and there only as the checker / the timed CPU reference.

PARITY: PINNED TO THE REFERENCE'S OWN CODE, NOT TO ITS THIRD-PARTY KERNELS.  The reference (IGNF/myria3d) ships no
numerical golden vectors for this path and torch_geometric 2.4 / torch_cluster / torch_scatter cannot be installed here.
What could be done, and is (round 2):

* ``gen_golden_ref_model.py`` executes the reference's model file (``myria3d/models/modules/pyg_randla_net.py``,
  unmodified) on stand-ins for the PyG primitives it imports (``pyg_standin.py``: MLP, MessagePassing.propagate,
  knn_graph, knn_interpolate, softmax, scatter -- their published semantics).  ``randla_oracle.py`` reproduces those
  runs BIT FOR BIT (eval logits, train logits, loss, every gradient, BatchNorm buffers, the random stream), also under
  the shipped trained checkpoint (``gen_golden_ckpt.py``).  So the wiring, operator order, parameter names and RNG
  consumption of the oracle are the reference file's.
* ``gen_golden_ref.py`` executes the reference's ``split_cloud_into_samples`` / ``get_mosaic_of_centers``, the node-budget
  transforms, ``NormalizePos`` and ``Interpolator.reduce_predicted_logits`` the same way; ``sample_prep_oracle.py`` /
  ``stitch_oracle.py`` reproduce them bit for bit.

What stays UNPINNED: the arithmetic inside PyG's / torch_cluster's own kernels (kd-tree tie order, scatter summation order
on their CUDA path), restated from their documentation and cross-checked against independent formulations (scipy
cKDTree vs brute force, ``torch.nn.functional``).
"""
